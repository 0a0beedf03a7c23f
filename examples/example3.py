#!/usr/bin/env python
"""Sequence tracking with replacement, after example3.c of the C library the reference was ported from (the reference
itself ships only example1.py): select features in the first frame, then for every new frame track them, replace the
lost ones and store the list in a feature table; finally print how long each slot's tracks lived.

Runs on synthetic frames (`pyfeaturetrack_b200.synth`), needs a B200 and the built library.
usage: python example3.py [nframes]"""
from __future__ import print_function
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pyfeaturetrack_b200
pyfeaturetrack_b200.install_dropin()
from klt import KLT_TrackingContext, KLTCountRemainingFeatures                                     # noqa: E402
from selectGoodFeatures import KLTSelectGoodFeatures, KLTReplaceLostFeatures                        # noqa: E402
from trackFeatures import KLTTrackFeatures                                                          # noqa: E402
from storeFeatures import KLTCreateFeatureTable, KLTStoreFeatureList, table_arrays                 # noqa: E402
from writeFeatures import KLTWriteFeatureListToPPM                                                  # noqa: E402
import selectGoodFeatures, trackFeatures                                                           # noqa: E402
from pyfeaturetrack_b200 import synth                                                               # noqa: E402

selectGoodFeatures.KLT_verbose = trackFeatures.KLT_verbose = 0
nFeatures = 150
nFrames = int(sys.argv[1]) if len(sys.argv) > 1 else 10
frames = synth.fast_frames(480, 640, nFrames, seed=11)

tc = KLT_TrackingContext()
tc.sequentialMode = True              # keep the previous frame's pyramid on the device
tc.writeInternalImages = False
tc.affineConsistencyCheck = -1
tc.max_residue = 10.0
ft = KLTCreateFeatureTable(nFrames, nFeatures)

fl = KLTSelectGoodFeatures(tc, frames[0], nFeatures)
KLTStoreFeatureList(fl, ft, 0)
for i in range(1, nFrames):
    KLTTrackFeatures(tc, frames[i - 1], frames[i], fl)
    tracked = KLTCountRemainingFeatures(fl)
    KLTReplaceLostFeatures(tc, frames[i], fl)
    KLTStoreFeatureList(fl, ft, i)
    print("frame %2d: %3d of %d features tracked, %3d replaced" % (i, tracked, nFeatures, nFeatures - tracked))

x, y, v = table_arrays(ft)
restarts = (v[:, 1:] > 0).sum(axis=1)          # val > 0 marks a freshly selected feature (its eigenvalue)
print("slots never lost: %d of %d; most restarts in one slot: %d" % ((restarts == 0).sum(), nFeatures, restarts.max()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example3_last.ppm")
KLTWriteFeatureListToPPM(fl, frames[-1], out)
print("wrote", out)
