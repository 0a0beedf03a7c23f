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


# ---- column bookkeeping: strips, lanes, halo lanes, mirrored blocks at the image edges ----------------------------------------
def mirror_block(c, W, n):
    """mirror_quad (n = 4) / mirror_oct (n = 8): start of the block to load and whether it is element-reversed"""
    rev = c < 0 or c >= W
    m = -c - n if c < 0 else (2 * W - c - n if c >= W else c)
    return min(max(m, 0), W - n), rev


def shfl_up(vals, lane):
    return vals[lane - 1] if lane >= 1 else vals[lane]


def shfl_down(vals, lane):
    return vals[lane + 1] if lane <= 30 else vals[lane]


def direct_row(row, s5, p11):
    W = len(row)
    ext = lambda i: float(row[reflect1(i, W)])
    sm = lambda c: sum(s5[j] * ext(c + j - 2) for j in range(5))
    L0 = np.array([sm(c) for c in range(W)])
    L1 = np.array([sum(p11[j] * sm(2 * X + 1 + j - 5) for j in range(11)) for X in range(W // 2)])
    return L0, L1


def model_level01_row(row, s5, p11):
    """stream_level01_kernel for ONE image row (horizontal passes only): lane = 8 level-0 columns, 240 per strip, lanes 0 and 31
    are halo lanes, blocks outside the image are mirrored 8-byte loads."""
    W = len(row)
    OW = W // 2
    L0, L1 = np.full(W, np.nan), np.full(OW, np.nan)
    n0, n1 = np.zeros(W, int), np.zeros(OW, int)
    for strip in range((W + 239) // 240):
        x = []
        for lane in range(32):
            m0, rev = mirror_block(strip * 240 + 8 * (lane - 1), W, 8)
            blk = [float(v) for v in row[m0:m0 + 8]]
            x.append(blk[::-1] if rev else blk)
        s = []
        for lane in range(32):
            ext = [shfl_up([b[6] for b in x], lane), shfl_up([b[7] for b in x], lane)] + x[lane] + \
                  [shfl_down([b[0] for b in x], lane), shfl_down([b[1] for b in x], lane)]
            s.append([sum(s5[j] * ext[i + j] for j in range(5)) for i in range(8)])
        for lane in range(32):
            c0, X = strip * 240 + 8 * (lane - 1), strip * 120 + 4 * (lane - 1)
            e = [shfl_up([q[4 + k] for q in s], lane) for k in range(4)] + s[lane] + \
                [shfl_down([q[k] for q in s], lane) for k in range(4)] + [shfl_down([q[4] for q in s], lane)]
            h = [sum(p11[j] * e[2 * i + j] for j in range(11)) for i in range(4)]
            inner = 1 <= lane <= 30
            if inner and c0 < W:
                L0[c0:c0 + 8] = s[lane]; n0[c0:c0 + 8] += 1
            if inner and X < OW:
                L1[X:X + 4] = h; n1[X:X + 4] += 1
    assert (n0 == 1).all() and (n1 == 1).all(), "every column is stored exactly once"
    return L0, L1


@pytest.mark.parametrize("W", [64, 72, 120, 240, 248, 320, 488, 640, 720, 960, 1000, 1280, 1920, 1928])
def test_level01_column_bookkeeping(W):
    rng = np.random.default_rng(W)
    row = rng.integers(0, 256, W)
    s5 = np.array([0.05, 0.25, 0.4, 0.25, 0.05])
    g = np.exp(-0.5 * (np.arange(-5, 6) / 1.8) ** 2)
    p11 = g / g.sum()
    got0, got1 = model_level01_row(row, s5, p11)
    want0, want1 = direct_row(row, s5, p11)
    assert np.allclose(got0, want0, rtol=0, atol=1e-9)
    assert np.allclose(got1, want1, rtol=0, atol=1e-9)


def model_smooth0x2_row(row, s5):
    """stream_smooth0x2_kernel for one row: a warp owns strips A = 2 * pair and B = A + 120 columns, lane = one quad of each."""
    W = len(row)
    out, cnt = np.full(W, np.nan), np.zeros(W, int)
    n_strips = (W + 119) // 120
    for pair in range((n_strips + 1) // 2):
        for off in (0, 120):                                   # strip A, strip B: the same lanes, independent halos
            u = []
            for lane in range(32):
                m, rev = mirror_block(pair * 240 + off + 4 * (lane - 1), W, 4)
                q = [float(v) for v in row[m:m + 4]]
                u.append(q[::-1] if rev else q)
            for lane in range(32):
                c = pair * 240 + off + 4 * (lane - 1)
                ext = [shfl_up([q[2] for q in u], lane), shfl_up([q[3] for q in u], lane)] + u[lane] + \
                      [shfl_down([q[0] for q in u], lane), shfl_down([q[1] for q in u], lane)]
                if 1 <= lane <= 30 and c < W:
                    out[c:c + 4] = [sum(s5[j] * ext[i + j] for j in range(5)) for i in range(4)]
                    cnt[c:c + 4] += 1
    assert (cnt == 1).all()
    return out


@pytest.mark.parametrize("W", [16, 120, 124, 132, 240, 244, 360, 964, 1920, 1924])
def test_smooth0x2_column_bookkeeping(W):
    rng = np.random.default_rng(W + 1)
    row = rng.integers(0, 256, W)
    s5 = np.array([0.05, 0.25, 0.4, 0.25, 0.05])
    want = np.array([sum(s5[j] * float(row[reflect1(c + j - 2, W)]) for j in range(5)) for c in range(W)])
    assert np.allclose(model_smooth0x2_row(row, s5), want, rtol=0, atol=1e-9)


def model_down2p_row(row, p11):
    """stream_down2p_kernel for one row: lane = 4 output columns X.. = 8 input columns 2X.. as two mirrored quads; the packed
    horizontal pass takes taps (0, 1) .. (8, 9) on even-aligned column pairs and the eleventh tap alone."""
    W = len(row)
    OW = W // 2
    out, cnt = np.full(OW, np.nan), np.zeros(OW, int)
    for strip in range((OW + 119) // 120):
        a, b = [], []
        for lane in range(32):
            ci = 2 * (strip * 120 + 4 * (lane - 1))
            for dst, c in ((a, ci), (b, ci + 4)):
                m, rev = mirror_block(c, W, 4)
                q = [float(v) for v in row[m:m + 4]]
                dst.append(q[::-1] if rev else q)
        for lane in range(32):
            X = strip * 120 + 4 * (lane - 1)
            E = [None] * 8
            E[2], E[3], E[4], E[5] = (a[lane][0], a[lane][1]), (a[lane][2], a[lane][3]), (b[lane][0], b[lane][1]), (b[lane][2], b[lane][3])
            E[0] = (shfl_up([q[0] for q in b], lane), shfl_up([q[1] for q in b], lane))
            E[1] = (shfl_up([q[2] for q in b], lane), shfl_up([q[3] for q in b], lane))
            r = [shfl_down([q[k] for q in a], lane) for k in range(4)] + [shfl_down([q[0] for q in b], lane)]
            E[6], E[7] = (r[0], r[1]), (r[2], r[3])
            last = [b[lane][2], r[0], r[2], r[4]]
            h = []
            for i in range(4):
                lo = sum(p11[2 * m] * E[i + m][0] for m in range(5))
                hi = sum(p11[2 * m + 1] * E[i + m][1] for m in range(5))
                h.append(p11[10] * last[i] + (lo + hi))
            if 1 <= lane <= 30 and X < OW:
                out[X:X + 4] = h; cnt[X:X + 4] += 1
    assert (cnt == 1).all()
    return out


@pytest.mark.parametrize("W", [32, 64, 240, 248, 480, 488, 960, 968, 1920])
def test_down2p_column_bookkeeping(W):
    rng = np.random.default_rng(W + 2)
    row = rng.random(W) * 255
    g = np.exp(-0.5 * (np.arange(-5, 6) / 1.8) ** 2)
    p11 = g / g.sum()
    want = np.array([sum(p11[j] * row[reflect1(2 * X + 1 + j - 5, W)] for j in range(11)) for X in range(W // 2)])
    assert np.allclose(model_down2p_row(row, p11), want, rtol=0, atol=1e-9)


def test_fused_border_argument_in_float32():
    """stream_level01 decimates the smoothed REFLECT-EXTENDED frame where the two-kernel build decimates the reflect-extended
    SMOOTHED image.  With a symmetric smoothing kernel the two are the same numbers up to the order of the float32 sums (the
    mirrored window is summed right to left); the difference must stay far below the FAST tolerance (1e-5 of 255)."""
    rng = np.random.default_rng(11)
    H, W = 64, 96
    img = rng.integers(0, 256, (H, W)).astype(np.float32)
    s5 = np.array([0.0545, 0.2442, 0.4026, 0.2442, 0.0545], np.float32)
    g = np.exp(-0.5 * (np.arange(-5, 6) / 1.8) ** 2)
    p11 = (g / g.sum()).astype(np.float32)
    pad = 8

    def conv_sep(a, taps, r):          # 'valid' separable correlation in float32, fixed left-to-right order
        out = np.zeros((a.shape[0], a.shape[1] - 2 * r), np.float32)
        for j, t in enumerate(taps):
            out += t * a[:, j:j + out.shape[1]]
        a2 = out
        out2 = np.zeros((a2.shape[0] - 2 * r, a2.shape[1]), np.float32)
        for j, t in enumerate(taps):
            out2 += t * a2[j:j + out2.shape[0], :]
        return out2

    ext = np.pad(img, pad + 2, mode="symmetric")                       # SciPy 'reflect' = NumPy 'symmetric'
    sm_of_ext = conv_sep(ext, s5, 2)                                   # smoothed image on the extended domain (fused kernel)
    sm = sm_of_ext[pad:-pad, pad:-pad]                                 # level 0 (identical in both builds)
    ext_of_sm = np.pad(sm, pad, mode="symmetric")                      # reflect-extension of level 0 (two-kernel build)
    assert np.abs(sm_of_ext - ext_of_sm).max() <= 1e-5 * 255.0
    d_a = conv_sep(sm_of_ext, p11, 5)[pad - 5 + 1::2, pad - 5 + 1::2]
    d_b = conv_sep(ext_of_sm, p11, 5)[pad - 5 + 1::2, pad - 5 + 1::2]
    assert d_a.shape == d_b.shape and d_a.shape[0] >= H // 2
    assert np.abs(d_a - d_b).max() <= 1e-5 * 255.0
