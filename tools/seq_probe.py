"""Sequence-step probe (config D shape): frames/s of klt_sequence for several batch sizes and arithmetic modes, device-timed,
plus the per-kernel profile of one configuration.  python tools/seq_probe.py [--frames 60]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--batches", default="1,8,16")
    ap.add_argument("--H", type=int, default=1080)
    ap.add_argument("--W", type=int, default=1920)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--modes", default="fast,mixed,strict")
    args = ap.parse_args()
    from pyfeaturetrack_b200 import _capi, klt, synth, selectGoodFeatures as sgf, trackFeatures as tf
    ctx = _capi.default_ctx()
    tc = klt.KLT_TrackingContext()
    tc.nPyramidLevels, tc.subsampling, tc.max_residue, tc.sequentialMode = 3, 2, 10.0, True
    tc.KLTUpdateTCBorder()
    params, taps = sgf.make_params(tc), tf._taps_for_one_image(tc)
    H, W, n = args.H, args.W, args.n
    distinct = 16
    base = [synth.fast_frames(H, W, distinct, seed=7 + s) for s in range(4)]
    out = []
    for B in [int(b) for b in args.batches.split(",")]:
        frames = ctx.pinned_array((distinct, B, H, W), np.uint8)
        for k in range(distinct):
            for s in range(B):
                frames[k, s] = base[s % 4][(k + s // 4) % distinct]
        dfr = ctx.device_alloc(frames.nbytes)
        ctx.memcpy(dfr, frames, frames.nbytes)
        ctx.sync()
        all_modes = {"fast": ("windowed", _capi.PRECISION_FAST_WINDOWED, "fast", _capi.SELECT_FAST),
                     "mixed": ("windowed", _capi.PRECISION_FAST_WINDOWED, "strict", _capi.SELECT_STRICT),
                     "strict": ("strict", _capi.PRECISION_STRICT, "strict", _capi.SELECT_STRICT)}
        for (pname, prec, sname, smode) in [all_modes[m] for m in args.modes.split(",")]:
            q = _capi.Sequence(ctx, params, taps, W, H, B, n, prec, smode)
            q.start(dfr)
            for k in range(1, 6):
                q.step(dfr + (k % distinct) * B * H * W)
            q.sync()
            l0 = ctx.launch_count()
            ctx.timer_start()
            t0 = time.perf_counter()
            for k in range(args.frames):
                q.step(dfr + ((k + 6) % distinct) * B * H * W)
            ctx.timer_stop()
            ms = ctx.timer_elapsed_ms()
            q.sync()
            wall = (time.perf_counter() - t0) * 1e3
            x, y, v, vt = q.features()
            st = q.select_stats()
            rec = dict(walk_consumed_mean=float(st[:, 0].mean()), walk_fallbacks=int(st[:, 3].sum()), B=B, precision=pname, select=sname, ms_per_step=round(ms / args.frames, 4), wall_ms_per_step=round(wall / args.frames, 4),
                       frames_per_sec=round(B * args.frames / (ms * 1e-3), 1), graph=q.uses_graph(),
                       launches_per_step=(ctx.launch_count() - l0) / args.frames,
                       tracked_frac=float((vt == 0).mean()), filled_frac=float((v >= 0).mean()))
            # host-fed: pinned frames uploaded every step, lists downloaded every step
            hx, hy = ctx.pinned_array((B, n), np.float64), ctx.pinned_array((B, n), np.float64)
            hv, hvt = ctx.pinned_array((B, n), np.int32), ctx.pinned_array((B, n), np.int32)
            for k in range(3):
                q.step(frames[k % distinct], out=(hx, hy, hv, hvt))
            q.sync()
            t0 = time.perf_counter()
            for k in range(args.frames):
                q.step(frames[(k + 3) % distinct], out=(hx, hy, hv, hvt))
            q.sync()
            e2e = (time.perf_counter() - t0) * 1e3
            rec["e2e_ms_per_step"] = round(e2e / args.frames, 4)
            rec["e2e_frames_per_sec"] = round(B * args.frames / (e2e * 1e-3), 1)
            if B == 8:
                ctx.profile_reset(); ctx.profile(True)
                for k in range(8):
                    q.step(dfr + (k % distinct) * B * H * W)
                ctx.profile(False)
                prof = ctx.profile_read()
                rec["kernel_us_per_step"] = {k2: round(1e3 * v2["ms"] / 8, 2) for k2, v2 in prof.items()}
                rec["kernel_gbps"] = {k2: round(v2["bytes"] / (v2["ms"] * 1e-3) / 1e9, 1) for k2, v2 in prof.items() if v2["bytes"] and v2["ms"]}
            q.close()
            out.append(rec)
            print(json.dumps(rec), flush=True)
        ctx.device_free(dfr)


if __name__ == "__main__":
    main()
