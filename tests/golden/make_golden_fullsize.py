#!/usr/bin/env python
"""Generates tests/golden/reference_golden_fullsize.npz by RUNNING THE UNMODIFIED REFERENCE at BASELINE's own sizes
(config B: 1080p / 1000 features / 3 levels; config C: 4K / 10 000 features / 4 levels; config D: a 1080p sequence in
sequentialMode with per-frame replacement).  Same mechanics as make_golden.py: the reference's .py modules are imported from
/root/reference, its two Cython modules from oracle/_ref.  Only the small result arrays are stored; the frames are
regenerated from their seeds by pyfeaturetrack_b200.synth (NumPy + SciPy only).

    python oracle/build_ref.py && python tests/golden/make_golden_fullsize.py      # ~2 minutes of reference CPU time
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KLT_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, REF)
warnings.simplefilter("ignore")

from PIL import Image  # noqa: E402
import klt  # noqa: E402
import selectGoodFeatures as sgf  # noqa: E402
import trackFeatures as tf  # noqa: E402
import goodFeaturesUtils as gfu  # noqa: E402
from pyfeaturetrack_b200 import synth  # noqa: E402

assert klt.__file__.startswith(REF), klt.__file__
sgf.KLT_verbose = 0
tf.KLT_verbose = 0
G = {}


def fl_arrays(fl):
    return np.stack([np.array([float(f.x) for f in fl]), np.array([float(f.y) for f in fl]), np.array([float(f.val) for f in fl])])


def context(**kw):
    tc = klt.KLT_TrackingContext()
    for k, v in kw.items():
        setattr(tc, k, v)
    tc.KLTUpdateTCBorder()
    return tc


# config B: synth.frame_pair(1080, 1920, seed=0)
imgs = [Image.fromarray(a) for a in synth.frame_pair(1080, 1920, seed=0)]
tc = context(nPyramidLevels=3, subsampling=2, max_residue=10.0)
fl = sgf.KLTSelectGoodFeatures(tc, imgs[0], 1000)
G["B_sel1000"] = fl_arrays(fl)
tf.KLTTrackFeatures(tc, imgs[0], imgs[1], fl)
G["B_trk1000"] = fl_arrays(fl)

# config C: synth.frame_pair(2160, 3840, seed=0), selection only (the reference needs ~7 s per 4K tracking call and ~20 s to select)
imgs = [Image.fromarray(a) for a in synth.frame_pair(2160, 3840, seed=0)]
tc = context(nPyramidLevels=4, subsampling=2, max_residue=10.0)
fl = sgf.KLTSelectGoodFeatures(tc, imgs[0], 10000)
G["C_sel10000"] = fl_arrays(fl)
tf.KLTTrackFeatures(tc, imgs[0], imgs[1], fl)
G["C_trk10000"] = fl_arrays(fl)

# config D: synth.frames(1080, 1920, 6 x sequence_shifts(5), seed=101), sequentialMode, replacement through the reference's own
# ScanImageForGoodFeatures + sort + _enforceMinimumDistance(overwriteAllFeatures=False) on tc.pyramid_last's gradients
nfr = 5
shifts = [(s[0] * 6, s[1] * 6) for s in synth.sequence_shifts(nfr)]
frames = [Image.fromarray(a) for a in synth.frames(1080, 1920, shifts, seed=101)]
tc = context(nPyramidLevels=3, subsampling=2, max_residue=10.0, sequentialMode=True)
fl = sgf.KLTSelectGoodFeatures(tc, frames[0], 1000)
seq = [fl_arrays(fl)]
for k in range(1, nfr):
    tf.KLTTrackFeatures(tc, frames[k - 1], frames[k], fl)
    seq.append(fl_arrays(fl))
    gx, gy = tc.pyramid_last_gradx.img[0], tc.pyramid_last_grady.img[0]
    bx, hw = int(max(tc.borderx, tc.window_width / 2)), int(tc.window_width / 2)
    px, py, pv = gfu.ScanImageForGoodFeatures(gx, gy, bx, bx, hw, hw, tc.nSkippedPixels)
    pts = list(zip(pv, px, py))
    pts.sort()
    pts.reverse()
    sgf._enforceMinimumDistance(pts, fl, 1920, 1080, tc.mindist, tc.min_eigenvalue, False)
    seq.append(fl_arrays(fl))
G["D_seq1000"] = np.stack(seq)          # [select, track1, replace1, track2, replace2, ...][x|y|val][1000]

np.savez_compressed(os.path.join(HERE, "reference_golden_fullsize.npz"), **G)
print("wrote", {k: v.shape for k, v in G.items()})
