"""Experiment: device-resident pair steps (workload B, 64 pairs) issued to ONE context vs alternating between TWO contexts
(own streams, pyramids and lists; shared read-only frames), so that the tracking kernel of step k can overlap the pyramid
builds of step k + 1.  python tools/two_ctx_probe.py [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    from pyfeaturetrack_b200 import _capi, klt, synth, selectGoodFeatures as sgf, trackFeatures as tf
    sgf.KLT_verbose = 0
    H, W, n, L, ss, B = 1080, 1920, 1000, 3, 2, 64
    tc = klt.KLT_TrackingContext()
    tc.nPyramidLevels, tc.subsampling, tc.max_residue = L, ss, 10.0
    tc.KLTUpdateTCBorder()
    params, taps = sgf.make_params(tc), tf._taps_for_one_image(tc)
    lib = _capi.lib()
    ctxs = [_capi.default_ctx(), _capi.Context(_capi.default_ctx().device)]
    pairs = [synth.frame_pair(H, W, seed=s) for s in range(4)]
    f1 = np.stack([pairs[i % 4][0] for i in range(B)]); f2 = np.stack([pairs[i % 4][1] for i in range(B)])
    sel = []
    for a, _ in pairs:
        fl = sgf.KLTSelectGoodFeatures(tc, a, n)
        sel.append((np.array([float(f.x) for f in fl]), np.array([float(f.y) for f in fl]), np.array([f.val for f in fl], np.int32)))
    x0 = np.stack([sel[i % 4][0] for i in range(B)]); y0 = np.stack([sel[i % 4][1] for i in range(B)]); v0 = np.stack([sel[i % 4][2] for i in range(B)])
    c0 = ctxs[0]
    d_f1, d_f2 = c0.device_alloc(f1.nbytes), c0.device_alloc(f2.nbytes)
    c0.memcpy(d_f1, f1, f1.nbytes); c0.memcpy(d_f2, f2, f2.nbytes)
    d0 = [c0.device_alloc(a.nbytes) for a in (x0, y0, v0)]
    for d, a in zip(d0, (x0, y0, v0)):
        c0.memcpy(d, a, a.nbytes)
    c0.sync()
    lanes = []
    for c in ctxs:
        lanes.append(dict(c=c, p1=_capi.Pyramid(c, W, H, L, ss, B), p2=_capi.Pyramid(c, W, H, L, ss, B),
                          d=[c.device_alloc(a.nbytes) for a in (x0, y0, v0)]))

    def step(lane):
        c = lane["c"]
        for dst, src, a in zip(lane["d"], d0, (x0, y0, v0)):
            c.memcpy(dst, src, a.nbytes)
        c.check(lib.klt_track_pairs_u8(c.handle, C.byref(params), C.byref(taps), _capi.PRECISION_FAST_WINDOWED, lane["p1"].handle,
                                       lane["p2"].handle, d_f1, d_f2, W, W * H, n, lane["d"][0], lane["d"][1], lane["d"][2]))

    def run(nl):
        for k in range(6):
            step(lanes[k % nl])
        for c in ctxs:
            c.sync()
        t0 = time.perf_counter()
        for k in range(steps):
            step(lanes[k % nl])
        for c in ctxs:
            c.sync()
        return (time.perf_counter() - t0) * 1e3 / steps

    for rep in range(2):
        a, b = run(1), run(2)
        print("one context %.4f ms/step (%.1f k pairs/s)   two contexts alternating %.4f ms/step (%.1f k pairs/s)" % (a, B / a, b, B / b))
    hv = np.empty_like(v0)
    for lane in lanes:
        lane["c"].memcpy(hv, lane["d"][2], hv.nbytes); lane["c"].sync()
        print("tracked", int((hv == 0).sum()))


if __name__ == "__main__":
    main()
