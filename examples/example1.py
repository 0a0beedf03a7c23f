#!/usr/bin/env python
"""The reference's example1.py (example1.py:17-65) on the GPU: select the best features of img0, track them to img1 and
back 100 times, print the time per KLTTrackFeatures call.  Only the import preamble differs from the reference script
(and time.clock(), which no longer exists, became time.perf_counter())."""
from __future__ import print_function
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyfeaturetrack_b200
pyfeaturetrack_b200.install_dropin()

from klt import *                      # noqa: E402,F401,F403
from PIL import Image                  # noqa: E402
from selectGoodFeatures import *       # noqa: E402,F401,F403
from writeFeatures import *            # noqa: E402,F401,F403
from trackFeatures import *            # noqa: E402,F401,F403
import selectGoodFeatures, trackFeatures  # noqa: E402


def main():
    tc = KLT_TrackingContext()
    nFeatures = 50
    tc.nSkippedPixels = 0
    tc.max_residue = 10.0
    KLTPrintTrackingContext(tc)
    g = os.path.join(ROOT, "tests", "golden")
    img1 = Image.open(os.path.join(g, "img0.pgm"))
    img2 = Image.open(os.path.join(g, "img1.pgm"))
    fl = KLTSelectGoodFeatures(tc, img1, nFeatures)
    print("\nIn first image:")
    for i, feat in enumerate(fl):
        print("Feature #{0}:  ({1},{2}) with value of {3}".format(i, feat.x, feat.y, feat.val))
    KLTWriteFeatureListToPPM(fl, img1, "feat1.ppm")
    selectGoodFeatures.KLT_verbose = trackFeatures.KLT_verbose = 0
    count = 0
    ti = time.perf_counter()
    for i in range(100):
        KLTTrackFeatures(tc, img1, img2, fl)
        KLTTrackFeatures(tc, img2, img1, fl)
        count += 2
    print((time.perf_counter() - ti) / count)
    print("\nIn second image:")
    for i, feat in enumerate(fl):
        print("Feature #{0}:  ({1},{2}) with value of {3}".format(i, feat.x, feat.y, feat.val))
    KLTWriteFeatureListToPPM(fl, img2, "feat2.ppm")


if __name__ == "__main__":
    main()
