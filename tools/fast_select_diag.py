"""How the fused fast selection differs from the reference's (oracle) selection: exact set overlap, and for every feature
that differs whether it is a near-tie (a neighbour within mindist, or a swap at the cut-off value)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from pyfeaturetrack_b200 import klt, synth, selectGoodFeatures as sgf, config
    from oracle import klt_oracle as O
    sgf.KLT_verbose = 0
    for (H, W, n, L, seeds) in ((1080, 1920, 1000, 3, (0, 1, 2, 3)), (2160, 3840, 10000, 4, (0,))):
        for seed in seeds:
            tc = klt.KLT_TrackingContext()
            tc.nPyramidLevels, tc.subsampling = L, 2
            tc.KLTUpdateTCBorder()
            p = O.Params(nPyramidLevels=L, subsampling=2)
            f = synth.frames(H, W, [(0.0, 0.0)], seed=seed)[0]
            res = {}
            for mode in ("fast", "strict"):
                config.set_precision(select=mode)
                fl = sgf.KLTSelectGoodFeatures(tc, f, n)
                res[mode] = {(int(a.x), int(a.y)): int(a.val) for a in fl if a.val > 0}
            wx, wy, wv = O.select_good_features(p, f, n)
            want = {(int(a), int(b)): int(c) for a, b, c in zip(wx, wy, wv) if c > 0}
            assert res["strict"] == want
            got = res["fast"]
            miss = [k for k in want if k not in got]
            extra = [k for k in got if k not in want]
            cutoff = min(want.values())
            near = 0
            for (x, y) in extra:
                d = [max(abs(x - a), abs(y - b)) for (a, b) in miss]
                if d and min(d) <= 9:
                    near += 1
            print("H=%d seed=%d n=%d overlap=%.4f differing=%d of which within mindist of the feature they replace=%d; cutoff val=%d; extra vals=%s" %
                  (H, seed, n, len(set(got) & set(want)) / float(len(want)), len(extra), near, cutoff, sorted(got[k] for k in extra)[:12]), flush=True)


if __name__ == "__main__":
    main()
