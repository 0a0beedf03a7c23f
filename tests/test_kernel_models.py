"""CPU models of the row bookkeeping of the streaming kernels (pyfeaturetrack_b200/csrc/klt_stream.cu).

The GPU tests compare the kernels with the oracle at a handful of shapes; these models replay the kernels' loop structure
(segments, warm-up rows, the look-ahead, reflect indices, which rows a segment stores) in Python for MANY heights and segment
sizes -- odd heights, one-segment images, segments that touch both borders -- against a direct evaluation of the same filters.
They follow the CUDA code statement by statement (same variable names) and have to be kept in step with it."""
import numpy as np
import pytest


def reflect1(i, n):
    return -i - 1 if i < 0 else (2 * n - i - 1 if i >= n else i)


def model_level01_column(col, s5, p11, rows_per_seg):
    """stream_level01_kernel / l01_rows for ONE column (vertical passes only): u8 column -> smoothed level-0 column and
    decimated level-1 column (output Y centred on level-0 row 2Y + 1, 11 taps)."""
    H = len(col)
    OH = H // 2
    L0 = np.full(H, np.nan)
    L1 = np.full(OH, np.nan)
    stores0 = np.zeros(H, int)
    for ys in range(0, OH, rows_per_seg):
        ye = min(OH, ys + rows_per_seg)
        r_end = H if ye == OH else 2 * ye
        sa = [0.0] * 4                               # pending smoothed rows (vertical accumulate-and-shift, radius 2)
        P = [0.0] * 5                                # pending level-1 rows
        j0, j1 = ys - 2, ye + 3
        t0, t_last = 2 * j0 - 2, 2 * (j1 - 1) + 3
        tn = [t0]

        def load_row():
            t = min(tn[0], t_last)
            tn[0] += 1
            return float(col[reflect1(t, H)])

        def smooth_row(x):                           # returns the completed smoothed row (input row t - 2)
            out = s5[4] * x + sa[0]
            for m in range(3):
                sa[m] = s5[3 - m] * x + sa[m + 1]
            sa[3] = s5[0] * x
            return out

        wa, wb = load_row(), load_row()
        for _ in range(2):
            smooth_row(wa); wa = load_row()
            smooth_row(wb); wb = load_row()
        Y_out = ys
        for j in range(j0, j1):
            r = 2 * j
            he = smooth_row(wa); wa = load_row()
            if 2 * ys <= r < r_end:
                L0[r] = he; stores0[r] += 1
            ho = smooth_row(wb); wb = load_row()
            if 2 * ys <= r + 1 < r_end:
                L0[r + 1] = ho; stores0[r + 1] += 1
            o = p11[10] * he + P[0]
            a = [p11[8] * he + P[1], p11[6] * he + P[2], p11[4] * he + P[3], p11[2] * he + P[4], p11[0] * he]
            P = [p11[9] * ho + a[0], p11[7] * ho + a[1], p11[5] * ho + a[2], p11[3] * ho + a[3], p11[1] * ho + a[4]]
            if j - 3 >= ys:
                assert Y_out == j - 3 < ye
                L1[Y_out] = o
                Y_out += 1
        assert Y_out == ye
    assert (stores0 == 1).all(), "every level-0 row is stored exactly once"
    return L0, L1


def direct_level01_column(col, s5, p11):
    H = len(col)
    ext = lambda a, i: a[reflect1(i, len(a))]
    # smoothed image on the reflect-extended domain = smoothing of the reflect-extended column (what the kernel computes)
    sm = lambda r: sum(s5[j] * float(ext(col, r + j - 2)) for j in range(5))
    L0 = np.array([sm(r) for r in range(H)])
    L1 = np.array([sum(p11[j] * sm(2 * Y + 1 + j - 5) for j in range(11)) for Y in range(H // 2)])
    return L0, L1


@pytest.mark.parametrize("H", [32, 33, 47, 64, 97, 135, 270])
@pytest.mark.parametrize("rows_per_seg", [16, 17, 34, 1000])
def test_level01_row_bookkeeping(H, rows_per_seg):
    rng = np.random.default_rng(H * 31 + rows_per_seg)
    col = rng.integers(0, 256, H)
    s5 = np.array([0.05, 0.25, 0.4, 0.25, 0.05])                     # symmetric, like the reference's smoothing kernel
    g = np.exp(-0.5 * (np.arange(-5, 6) / 1.8) ** 2)
    p11 = g / g.sum()
    got0, got1 = model_level01_column(col, s5, p11, rows_per_seg)
    want0, want1 = direct_level01_column(col, s5, p11)
    assert np.allclose(got0, want0, rtol=0, atol=1e-9)
    assert np.allclose(got1, want1, rtol=0, atol=1e-9)


def model_down2p_column(col, p11, rows_per_seg):
    """stream_down2p_kernel / down2p_rows for one column: the loop takes two row pairs per trip, an odd number of pairs
    starts one pair early, loads clamp at the last pair a segment needs (non-interior path)."""
    H = len(col)
    OH = H // 2
    out = np.full(OH, np.nan)
    for ys in range(0, OH, rows_per_seg):
        ye = min(OH, ys + rows_per_seg)
        j1 = ye + 3
        j0 = ys - 2 - ((ye - ys + 5) & 1)
        P = [0.0] * 5
        jn = [j0]

        def load_pair():
            j = min(jn[0], j1 - 1)
            jn[0] += 1
            return float(col[reflect1(2 * j, H)]), float(col[reflect1(2 * j + 1, H)])

        Y_out = [ys]

        def step(pair, j):
            nonlocal P
            he, ho = pair
            o = p11[10] * he + P[0]
            a = [p11[8] * he + P[1], p11[6] * he + P[2], p11[4] * he + P[3], p11[2] * he + P[4], p11[0] * he]
            P = [p11[9] * ho + a[0], p11[7] * ho + a[1], p11[5] * ho + a[2], p11[3] * ho + a[3], p11[1] * ho + a[4]]
            if j - 3 >= ys:
                assert Y_out[0] == j - 3 < ye
                out[Y_out[0]] = o
                Y_out[0] += 1

        A, B = load_pair(), load_pair()
        j = j0
        while j < j1:
            step(A, j); A = load_pair()
            step(B, j + 1); B = load_pair()
            j += 2
        assert Y_out[0] == ye
    return out


@pytest.mark.parametrize("H", [16, 33, 64, 97, 270])
@pytest.mark.parametrize("rows_per_seg", [8, 9, 30, 31, 1000])
def test_down2p_row_bookkeeping(H, rows_per_seg):
    rng = np.random.default_rng(H * 17 + rows_per_seg)
    col = rng.random(H) * 255
    g = np.exp(-0.5 * (np.arange(-5, 6) / 1.8) ** 2)
    p11 = g / g.sum()
    got = model_down2p_column(col, p11, rows_per_seg)
    want = np.array([sum(p11[k] * col[reflect1(2 * Y + 1 + k - 5, H)] for k in range(11)) for Y in range(H // 2)])
    assert np.allclose(got, want, rtol=0, atol=1e-9)
