#!/usr/bin/env python
"""Generates tests/golden/reference_golden.npz by RUNNING THE UNMODIFIED REFERENCE (TimSC/PyFeatureTrack).

Run in the build container, where /root/reference is mounted:
    python oracle/build_ref.py && python tests/golden/make_golden.py
The reference's Python modules are imported straight from /root/reference (interpreted, unmodified); its two
Cython extension modules come from oracle/_ref (built from the reference's .pyx by oracle/build_ref.py).
The reference ships no tests or golden vectors of its own, so these vectors - outputs of the
reference itself on its two PGM fixtures and on seeded synthetic frames - are what pins the oracle and the
CUDA path on the GPU box, where /root/reference does not exist.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KLT_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))   # compiled goodFeaturesUtils / trackFeaturesUtils
sys.path.insert(0, REF)                                      # klt.py, convolve.py, ... interpreted from the mount
warnings.simplefilter("ignore")

from PIL import Image  # noqa: E402
import klt  # noqa: E402
import convolve  # noqa: E402
import pyramid  # noqa: E402
import selectGoodFeatures as sgf  # noqa: E402
import trackFeatures as tf  # noqa: E402
import trackFeaturesUtils as tfu  # noqa: E402
import goodFeaturesUtils as gfu  # noqa: E402
from pyfeaturetrack_b200 import synth  # noqa: E402

assert klt.__file__.startswith(REF), klt.__file__
sgf.KLT_verbose = 0
tf.KLT_verbose = 0

G = {}


def fl_arrays(fl):
    return (np.array([float(f.x) for f in fl]), np.array([float(f.y) for f in fl]), np.array([int(f.val) for f in fl], np.int32))


def context(**kw):
    tc = klt.KLT_TrackingContext()
    for k, v in kw.items():
        setattr(tc, k, v)
    tc.KLTUpdateTCBorder()
    return tc


# ---- kernels and borders -------------------------------------------------------------------------------
for s in (0.7, 1.0, 1.5, 1.8, 3.6, 7.2):
    g, d = convolve._computeKernels(s)
    G["taps_g_%s" % s] = np.array(g)
    G["taps_d_%s" % s] = np.array(d)
borders = []
for (w, L, ss) in ((7, 2, 4), (7, 3, 2), (7, 4, 2), (15, 3, 2), (7, 2, 2), (7, 2, 8), (7, 3, 8), (5, 1, 2), (15, 2, 4), (9, 3, 4)):
    tc = context(window_width=w, window_height=w, nPyramidLevels=L, subsampling=ss)
    borders.append((w, L, ss, tc.borderx))
G["borders"] = np.array(borders, np.float64)
tc = klt.KLT_TrackingContext()
G["default_ctx"] = np.array([tc.nPyramidLevels, tc.subsampling, tc.borderx, tc.bordery], np.float64)
pyr_choices = []
for sr in (1, 3, 5, 10, 15, 20, 31, 32, 50, 100, 400):
    tc = klt.KLT_TrackingContext()
    tc.KLTChangeTCPyramid(sr)
    pyr_choices.append((sr, tc.nPyramidLevels, tc.subsampling))
G["pyramid_choices"] = np.array(pyr_choices, np.float64)

# ---- config A: the reference's own fixtures ---------------------------------------------------------------
img0 = Image.open(os.path.join(REF, "img0.pgm"))
img1 = Image.open(os.path.join(REF, "img1.pgm"))
f0 = np.array(img0.convert("F"))
tc = context(max_residue=10.0)
sm = convolve.KLTComputeSmoothedImage(f0, 0.7)
gx, gy = convolve.KLTComputeGradients(sm, 1.0)
G["A_smooth"] = sm
G["A_gradx"] = gx
G["A_grady"] = gy
p = pyramid.KLTPyramid(320, 240, 4, 2)
p.Compute(sm, 0.9)
G["A_pyr1"] = p.img[1]
g1x, g1y = convolve.KLTComputeGradients(p.img[1], 1.0)
G["A_pyr1_gradx"] = g1x
G["A_pyr1_grady"] = g1y
px, py, pv = gfu.ScanImageForGoodFeatures(gx, gy, 30, 30, 3, 3, 0)
G["A_scan_val"] = np.array(pv, np.float32).reshape(240 - 60, 320 - 60)
G["A_scan_x0"] = np.array(px[:5], np.int32)
G["A_scan_y0"] = np.array(py[:5], np.int32)
G["A_patch"] = tfu.extractImagePatchSlow(sm, 100.3, 57.8, 7, 7)
for n in (50, 100):
    tc = context(max_residue=10.0)
    fl = sgf.KLTSelectGoodFeatures(tc, img0, n)
    G["A_sel%d" % n] = np.stack(fl_arrays(fl))
    tf.KLTTrackFeatures(tc, img0, img1, fl)
    G["A_trk%d" % n] = np.stack(fl_arrays(fl))
    tf.KLTTrackFeatures(tc, img1, img0, fl)          # example1.py:53-55 tracks back and forth
    G["A_trk%d_back" % n] = np.stack(fl_arrays(fl))
# no residue check, retainTrackers, skipped pixels
tc = context()
fl = sgf.KLTSelectGoodFeatures(tc, img0, 60)
tf.KLTTrackFeatures(tc, img0, img1, fl)
G["A_trk60_nores"] = np.stack(fl_arrays(fl))
tc = context(retainTrackers=True)
fl = sgf.KLTSelectGoodFeatures(tc, img0, 60)
tf.KLTTrackFeatures(tc, img0, img1, fl)
G["A_trk60_retain"] = np.stack(fl_arrays(fl))
tc = context(nSkippedPixels=2, mindist=15, min_eigenvalue=500)
fl = sgf.KLTSelectGoodFeatures(tc, img0, 40)
G["A_sel40_skip2"] = np.stack(fl_arrays(fl))
tc = context(smoothBeforeSelecting=False)
fl = sgf.KLTSelectGoodFeatures(tc, img0, 40)
G["A_sel40_nosmooth"] = np.stack(fl_arrays(fl))
# sequential mode over img0 -> img1 -> img0 with replacement through _enforceMinimumDistance(..., False)
tc = context(max_residue=10.0, sequentialMode=True)
fl = sgf.KLTSelectGoodFeatures(tc, img0, 80)
seq = []
for a, b in ((img0, img1), (img1, img0), (img0, img1)):
    tf.KLTTrackFeatures(tc, a, b, fl)
    seq.append(np.stack(fl_arrays(fl)))
    gxl, gyl = tc.pyramid_last_gradx.img[0], tc.pyramid_last_grady.img[0]
    px, py, pv = gfu.ScanImageForGoodFeatures(gxl, gyl, int(tc.borderx), int(tc.bordery), 3, 3, tc.nSkippedPixels)
    pl = list(zip(pv, px, py))
    pl.sort()
    pl.reverse()
    sgf._enforceMinimumDistance(pl, fl, 320, 240, tc.mindist, tc.min_eigenvalue, False)
    seq.append(np.stack(fl_arrays(fl)))
G["A_seq80"] = np.stack(seq)

# ---- synthetic: 640x480, L=3, ss=2 (a small config B) and 15x15 windows (a small config E, translational) ----
imgs = synth.frame_pair(480, 640, seed=0)
tc = context(max_residue=10.0, nPyramidLevels=3, subsampling=2)
I = [Image.fromarray(a) for a in imgs]
fl = sgf.KLTSelectGoodFeatures(tc, I[0], 300)
G["S640_sel300"] = np.stack(fl_arrays(fl))
tf.KLTTrackFeatures(tc, I[0], I[1], fl)
G["S640_trk300"] = np.stack(fl_arrays(fl))
f = np.array(I[1].convert("F"))
sm = convolve.KLTComputeSmoothedImage(f, 0.7)
p = pyramid.KLTPyramid(640, 480, 2, 3)
p.Compute(sm, 0.9)
G["S640_img1_pyr2"] = p.img[2]
g2x, g2y = convolve.KLTComputeGradients(p.img[2], 1.0)
G["S640_img1_pyr2_gradx"] = g2x

imgs = synth.frame_pair(480, 640, seed=1)
tc = context(window_width=15, window_height=15, nPyramidLevels=2, subsampling=4, max_residue=8.0)
I = [Image.fromarray(a) for a in imgs]
fl = sgf.KLTSelectGoodFeatures(tc, I[0], 200)
G["S640w15_sel200"] = np.stack(fl_arrays(fl))
tf.KLTTrackFeatures(tc, I[0], I[1], fl)
G["S640w15_trk200"] = np.stack(fl_arrays(fl))

# a harder pair: larger shift so that some features hit MAX_ITERATIONS / SMALL_DET / OOB paths
imgs = synth.frame_pair(240, 320, seed=2, shift=(6.2, -9.4))
tc = context(nPyramidLevels=2, subsampling=2, max_residue=5.0, max_iterations=4, min_determinant=2000.0)
I = [Image.fromarray(a) for a in imgs]
fl = sgf.KLTSelectGoodFeatures(tc, I[0], 150)
G["H320_sel150"] = np.stack(fl_arrays(fl))
tf.KLTTrackFeatures(tc, I[0], I[1], fl)
G["H320_trk150"] = np.stack(fl_arrays(fl))

# SMALL_DET path: a determinant threshold in the middle of the observed range
imgs = synth.frame_pair(240, 320, seed=3)
tc = context(nPyramidLevels=2, subsampling=2, min_determinant=3.0e7)
I = [Image.fromarray(a) for a in imgs]
fl = sgf.KLTSelectGoodFeatures(tc, I[0], 120)
G["D320_sel120"] = np.stack(fl_arrays(fl))
tf.KLTTrackFeatures(tc, I[0], I[1], fl)
G["D320_trk120"] = np.stack(fl_arrays(fl))

out = os.path.join(HERE, "reference_golden.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out), "bytes;", len(G), "arrays")
from collections import Counter
for k in ("A_trk100", "S640_trk300", "S640w15_trk200", "H320_trk150", "D320_trk120", "A_trk60_retain"):
    print(k, Counter(G[k][2].astype(int).tolist()))
