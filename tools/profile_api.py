#!/usr/bin/env python
"""cProfile of the drop-in calls (one 1080p pair, 1000 features): where the host time of KLTTrackFeatures /
KLTSelectGoodFeatures / KLTReplaceLostFeatures goes.  usage: python tools/profile_api.py [n_calls]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from PIL import Image
from pyfeaturetrack_b200 import klt, selectGoodFeatures as sgf, trackFeatures as tf, synth, _capi

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sgf.KLT_verbose = 0
tf.KLT_verbose = 0
a, b = synth.frame_pair(1080, 1920, seed=3)
ia, ib = Image.fromarray(a), Image.fromarray(b)
tc = klt.KLT_TrackingContext()
tc.nPyramidLevels, tc.subsampling, tc.max_residue = 3, 2, 10.0
tc.KLTUpdateTCBorder()
fl = sgf.KLTSelectGoodFeatures(tc, ia, 1000)
import copy
for name, fn in (("KLTTrackFeatures", lambda: tf.KLTTrackFeatures(tc, ia, ib, copy.copy(fl0))),
                 ("KLTSelectGoodFeatures", lambda: sgf.KLTSelectGoodFeatures(tc, ia, 1000))):
    fl0 = [copy.copy(f) for f in fl]
    for _ in range(3):
        fn()
    _capi.default_ctx().sync()
    t0 = time.perf_counter()
    for _ in range(n_calls):
        fl0 = [copy.copy(f) for f in fl]
        fn()
    _capi.default_ctx().sync()
    print("%s: %.3f ms per call (incl. list copy)" % (name, (time.perf_counter() - t0) / n_calls * 1e3))
    pr = cProfile.Profile()
    for _ in range(n_calls):
        fl0 = [copy.copy(f) for f in fl]          # (the list copy stays outside the profile)
        pr.enable()
        fn()
        pr.disable()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("cumtime").print_stats(28)
