"""A small pass over every kernel family of the FINAL build for compute-sanitizer (memcheck / racecheck / initcheck):
strict + fast pyramid builds, batched selection (all paths of the walk incl. small chunks), windowed / fast / strict
tracking, the affine check, the async pair pipeline and klt_sequence with graph replay.  Small images keep it short.

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from pyfeaturetrack_b200 import _capi, klt, synth, selectGoodFeatures as sgf, trackFeatures as tf, config
    sgf.KLT_verbose = 0
    tf.KLT_verbose = 0
    H, W, n = 240, 320, 100
    shifts = synth.sequence_shifts(6)
    seqs = [synth.frames(H, W, [(a * 4, b * 4) for a, b in shifts], seed=100 + s) for s in range(3)]
    # drop-in API in the three tracking modes + both selection modes
    for mode in ("strict", "fast", "windowed"):
        for sel in ("strict", "fast"):
            config.set_precision(track=mode, select=sel)
            tc = klt.KLT_TrackingContext()
            tc.max_residue = 10.0
            tc.sequentialMode = True
            fl = sgf.KLTSelectGoodFeatures(tc, seqs[0][0], n)
            for k in range(1, 4):
                tf.KLTTrackFeatures(tc, seqs[0][k - 1], seqs[0][k], fl)
                sgf.KLTReplaceLostFeatures(tc, seqs[0][k], fl)
    # affine consistency check, 15x15
    config.set_precision(track="fast", select="strict")
    tc = klt.KLT_TrackingContext()
    tc.window_width = tc.window_height = 15
    tc.affineConsistencyCheck = 2
    tc.KLTUpdateTCBorder()
    fl = sgf.KLTSelectGoodFeatures(tc, seqs[1][0], 40)
    for k in range(1, 4):
        tf.KLTTrackFeatures(tc, seqs[1][0], seqs[1][k], fl)
    # sequences: strict/strict, windowed/fast, fast/strict; enough steps for the graph replay
    ctx = _capi.default_ctx()
    tc = klt.KLT_TrackingContext()
    tc.nPyramidLevels, tc.subsampling, tc.max_residue = 2, 2, 10.0
    tc.KLTUpdateTCBorder()
    for prec, sm in ((_capi.PRECISION_STRICT, _capi.SELECT_STRICT), (_capi.PRECISION_FAST_WINDOWED, _capi.SELECT_FAST),
                     (_capi.PRECISION_FAST, _capi.SELECT_STRICT)):
        q = _capi.Sequence(ctx, sgf.make_params(tc), tf._taps_for_one_image(tc), W, H, 3, n, prec, sm)
        q.start(np.ascontiguousarray(np.stack([s[0] for s in seqs])))
        for k in range(1, 6):
            q.step(np.ascontiguousarray(np.stack([s[k] for s in seqs])))
        q.sync()
        assert q.uses_graph()
        q.close()
    # image-only builds of every streaming-kernel generation: fused levels 0 + 1 (width % 8 == 0), two-strip packed level 0 +
    # packed decimation (width % 8 == 4), whole-row CTAs with bulk stores (width >= 960), one level only
    tc3 = klt.KLT_TrackingContext()
    tc3.nPyramidLevels, tc3.subsampling = 3, 2
    tc3.KLTUpdateTCBorder()
    rng = np.random.default_rng(5)
    for (h3, w3, l3) in ((96, 328, 3), (96, 324, 3), (80, 964, 2), (72, 1920, 1), (64, 248, 2)):
        fr = (rng.random((2, h3, w3)) * 255).astype(np.uint8)
        p3 = _capi.Pyramid(ctx, w3, h3, l3, 2, 2)
        tc3.nPyramidLevels = l3
        p3.build_u8(fr, tf._taps_for_one_image(tc3), _capi.PRECISION_FAST_WINDOWED)
        ctx.sync()
        p3.close()
    # the walk's rare paths
    os.environ["KLT_B200_SELECT_CHUNK"] = "64"
    c2 = _capi.Context(ctx.device)
    del os.environ["KLT_B200_SELECT_CHUNK"]
    yy, xx = np.mgrid[0:H, 0:W]
    checker = (((xx // 8) + (yy // 8)) % 2 * 200 + 20).astype(np.uint8)
    pyr = _capi.Pyramid(c2, W, H, 2, 2, 2)
    pyr.build_u8(np.ascontiguousarray(np.stack([seqs[2][0], checker])), tf._taps_for_one_image(tc), _capi.PRECISION_STRICT)
    x, y, v = np.full((2, 300), -1.0), np.full((2, 300), -1.0), np.full((2, 300), -1, np.int32)
    params = sgf.make_params(tc)
    for rep in (0, 1):
        c2.check(_capi.lib().klt_select_good_features_batch(c2.handle, C.byref(params), pyr.handle, 300, rep, _capi.SELECT_STRICT,
                                                           x.ctypes.data, y.ctypes.data, v.ctypes.data))
        v[:, ::3] = -4
    pyr.close()
    c2.close()
    # batched pairs, host frames, async
    B = 4
    f1 = ctx.pinned_array((B, H, W), np.uint8); f2 = ctx.pinned_array((B, H, W), np.uint8)
    for i in range(B):
        f1[i], f2[i] = seqs[i % 3][0], seqs[i % 3][1]
    hx, hy, hv = ctx.pinned_array((B, n), np.float64), ctx.pinned_array((B, n), np.float64), ctx.pinned_array((B, n), np.int32)
    fl = sgf.KLTSelectGoodFeatures(tc, seqs[0][0], n)
    hx[:] = [float(f.x) for f in fl]; hy[:] = [float(f.y) for f in fl]; hv[:] = [int(f.val) for f in fl]
    p1, p2 = _capi.Pyramid(ctx, W, H, 2, 2, B), _capi.Pyramid(ctx, W, H, 2, 2, B)
    taps = tf._taps_for_one_image(tc)
    lib = _capi.lib()
    for prec in (_capi.PRECISION_FAST_WINDOWED, _capi.PRECISION_FAST):
        ctx.check(lib.klt_track_pairs_u8_async(ctx.handle, C.byref(params), C.byref(taps), prec, p1.handle, p2.handle, f1.ctypes.data,
                                               f2.ctypes.data, W, W * H, n, hx.ctypes.data, hy.ctypes.data, hv.ctypes.data))
        ctx.check(lib.klt_async_result(ctx.handle))
    print("sanitize_smoke done: %d launches" % ctx.launch_count())


if __name__ == "__main__":
    main()
