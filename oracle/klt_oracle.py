"""Python face of the CPU oracle (oracle/klt_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Parity status: PINNED against the unmodified reference (tests/test_oracle_vs_reference.py runs both in
this container; tests/golden/*.npz hold reference outputs for the GPU box, where /root/reference is absent).

The functions mirror the reference's call structure (file:line cited per function) but are written
against plain numpy arrays: images are uint8/float32 (H, W) arrays, feature lists are (x, y, val)
arrays.  Nothing here is used by pyfeaturetrack_b200.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KLT_TRACKED, KLT_NOT_FOUND, KLT_SMALL_DET, KLT_MAX_ITERATIONS, KLT_OOB, KLT_LARGE_RESIDUE = 0, -1, -2, -3, -4, -5
ORC_ASSERT = -100


def build(force=False):
    so = os.path.join(_HERE, "libkltoracle.so")
    src = os.path.join(_HERE, "klt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared",
                               "-o", so, src, "-lm"])
    return so


class _TrackParams(C.Structure):
    _fields_ = [("window_width", C.c_int), ("window_height", C.c_int), ("max_iterations", C.c_int),
                ("min_determinant", C.c_float), ("min_displacement", C.c_float), ("step_factor", C.c_float),
                ("has_max_residue", C.c_int), ("max_residue", C.c_float), ("retain_trackers", C.c_int),
                ("n_levels", C.c_int), ("subsampling", C.c_int), ("borderx", C.c_double), ("bordery", C.c_double),
                ("lighting_insensitive", C.c_int)]


class _AffineParams(C.Structure):
    _fields_ = [("affine_map", C.c_int), ("width", C.c_int), ("height", C.c_int), ("max_iterations", C.c_int),
                ("step_factor", C.c_float), ("small_det", C.c_float), ("th", C.c_float), ("th_aff", C.c_float),
                ("max_residue", C.c_float), ("mdd", C.c_float)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_conv1d.argtypes = [fp, C.c_int, C.c_int, dp, C.c_int, C.c_int, fp]
        L.orc_conv_separable.argtypes = [fp, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, fp]
        L.orc_subsample.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int]
        L.orc_subsample.restype = None
        L.orc_scan_good_features.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             fp, ip, ip]
        L.orc_select_from_scan.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_double, C.c_int, C.c_int, dp, dp, ip]
        L.orc_select_from_scan.restype = C.c_long
        pp = C.POINTER(fp)
        L.orc_track_features.argtypes = [C.POINTER(_TrackParams), pp, pp, pp, pp, pp, pp, ip, ip, C.c_int, dp, dp, ip]
        L.orc_track_features.restype = C.c_long
        L.orc_extract_patch.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, fp]
        L.orc_track_feature_level.argtypes = [C.c_float, C.c_float, dp, dp, fp, fp, fp, fp, fp, fp, C.c_int, C.c_int,
                                              C.POINTER(_TrackParams), ip]
        L.orc_affine_step.argtypes = [C.POINTER(_AffineParams), fp, fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_double,
                                      C.c_double, C.c_double, C.c_double, ip, fp, fp, fp, fp]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


# ---------------------------------------------------------------------------------------------------
# Kernel taps: _computeKernels (convolve.py:27-93).  Pure float64 scalar math; restated with the same
# operations in the same order (the taps feed bit-exact convolutions, so order matters).
# ---------------------------------------------------------------------------------------------------
MAX_KERNEL_WIDTH = 71


class KernelTooWide(Exception):
    """The reference reaches an undefined name here (convolve.py:62 -> NameError); we raise this."""


def compute_kernels(sigma):
    hw = MAX_KERNEL_WIDTH // 2
    g = [math.exp(-i * i / (2 * sigma * sigma)) for i in range(-hw, hw + 1)]
    d = [-i * g[i + hw] for i in range(-hw, hw + 1)]
    max_g, max_d = 1.0, float(sigma * math.exp(-0.5))
    gw = MAX_KERNEL_WIDTH
    i = -hw
    while abs(g[i + hw] / max_g) < 0.01:
        i += 1
        gw -= 2
    dw = MAX_KERNEL_WIDTH
    i = -hw
    while abs(d[i + hw] / max_d) < 0.01:
        i += 1
        dw -= 2
    if gw == MAX_KERNEL_WIDTH or dw == MAX_KERNEL_WIDTH:
        raise KernelTooWide(sigma)
    g = g[(MAX_KERNEL_WIDTH - gw) // 2:(MAX_KERNEL_WIDTH - gw) // 2 + gw]
    d = d[(MAX_KERNEL_WIDTH - dw) // 2:(MAX_KERNEL_WIDTH - dw) // 2 + dw]
    den = 0.0
    for v in g:
        den += v
    g = [v / den for v in g]
    dhw = dw // 2
    den = 0.0
    for i in range(-dhw, dhw + 1):
        den -= i * d[i + dhw]
    d = [v / den for v in d]
    return np.array(g, np.float64), np.array(d, np.float64)


class KernelCache:
    """The reference's single-slot, 0.05-tolerance kernel cache (convolve.py:23-25,88-91,236,258; quirk Q8)."""

    def __init__(self):
        self.sigma = None
        self.g = self.d = None

    def compute(self, sigma):
        self.g, self.d = compute_kernels(sigma)
        self.sigma = sigma
        return self.g, self.d

    def for_gradients(self, sigma):          # convolve.py:236 (raises TypeError if never primed, like the reference)
        if abs(sigma - self.sigma) > 0.05:
            return self.compute(sigma)
        return self.g, self.d

    def for_smooth(self, sigma):             # convolve.py:258
        if self.sigma is None or abs(sigma - self.sigma) > 0.05:
            return self.compute(sigma)
        return self.g, self.d


def conv1d(img, taps, axis):
    img = np.ascontiguousarray(img, np.float32)
    taps = np.ascontiguousarray(taps, np.float64)
    out = np.empty_like(img)
    rc = lib().orc_conv1d(_f(img), img.shape[0], img.shape[1], _d(taps), len(taps), axis, _f(out))
    assert rc == 0
    return out


def conv_separable(img, hk, vk):
    """_convolveSeparate (convolve.py:208-214)."""
    img = np.ascontiguousarray(img, np.float32)
    hk = np.ascontiguousarray(hk, np.float64)
    vk = np.ascontiguousarray(vk, np.float64)
    out = np.empty_like(img)
    rc = lib().orc_conv_separable(_f(img), img.shape[0], img.shape[1], _d(hk), len(hk), _d(vk), len(vk), _f(out))
    assert rc == 0
    return out


def smooth(img, sigma, cache):
    """KLTComputeSmoothedImage (convolve.py:254-264)."""
    g, _ = cache.for_smooth(sigma)
    return conv_separable(img, g, g)


def gradients(img, sigma, cache):
    """KLTComputeGradients (convolve.py:226-248): gradx = deriv_h o gauss_v, grady = gauss_h o deriv_v."""
    g, d = cache.for_gradients(sigma)
    return conv_separable(img, d, g), conv_separable(img, g, d)


def pyramid(img, subsampling, nlevels, sigma_fact, cache):
    """KLTPyramid.Compute (pyramid.py:37-77)."""
    levels = [img]
    cur = img
    sigma = subsampling * sigma_fact
    for _ in range(1, nlevels):
        sm = smooth(cur, sigma, cache)
        oh, ow = int(cur.shape[0] / subsampling), int(cur.shape[1] / subsampling)
        out = np.empty((oh, ow), np.float32)
        lib().orc_subsample(_f(sm), sm.shape[0], sm.shape[1], subsampling, _f(out), oh, ow)
        levels.append(out)
        cur = out
    return levels


class Params:
    """The fields of KLT_TrackingContext that the hot path reads (klt.py:44-73), reference defaults."""

    def __init__(self, **kw):
        self.mindist = 10
        self.window_width = 7
        self.window_height = 7
        self.sequentialMode = False
        self.retainTrackers = False
        self.smoothBeforeSelecting = True
        self.min_eigenvalue = 1
        self.min_determinant = 0.01
        self.max_iterations = 10
        self.min_displacement = 0.1
        self.max_residue = None
        self.grad_sigma = 1.0
        self.smooth_sigma_fact = 0.1
        self.pyramid_sigma_fact = 0.9
        self.step_factor = 1.0
        self.nSkippedPixels = 0
        self.lighting_insensitive = False   # the reference raises when set; restated from its commented C (UNPINNED)
        self.affineConsistencyCheck = -1
        self.affine_window_width = 15
        self.affine_window_height = 15
        self.affine_max_iterations = 10
        self.affine_max_residue = 10.0
        self.affine_min_displacement = 0.02
        self.affine_max_displacement_differ = 1.5
        self.nPyramidLevels = 2          # what KLTChangeTCPyramid(15) gives for 7x7 (klt.py:77,84-128)
        self.subsampling = 4
        self.cache = KernelCache()
        for k, v in kw.items():
            setattr(self, k, v)
        self.update_border()

    def smooth_sigma(self):              # klt_util.py:3-4
        return self.smooth_sigma_fact * max(self.window_width, self.window_height)

    def update_border(self):
        """KLTUpdateTCBorder (klt.py:137-189) with Python-3 true division (quirk Q1)."""
        window_hw = max(self.window_width, self.window_height) / 2
        gw, _ = self.cache.compute(self.smooth_sigma())
        smooth_gauss_hw = len(gw) / 2
        gw, _ = self.cache.compute(self.pyramid_sigma_fact * self.subsampling)
        pyramid_gauss_hw = len(gw) / 2
        n_invalid = smooth_gauss_hw
        for _ in range(1, self.nPyramidLevels):
            n_invalid = int((float(n_invalid) + pyramid_gauss_hw) / self.subsampling + 0.99)
        ss_power = 1
        for _ in range(1, self.nPyramidLevels):
            ss_power *= self.subsampling
        self.borderx = self.bordery = (n_invalid + window_hw) * ss_power


def image_pyramids(p, img_u8):
    """ComputeImagePyramids for one image (trackFeatures.py:165-172): float, smooth, pyramid, per-level gradients."""
    f = np.asarray(img_u8).astype(np.float32)
    sm = smooth(f, p.smooth_sigma(), p.cache)
    pyr = pyramid(sm, int(p.subsampling), p.nPyramidLevels, p.pyramid_sigma_fact, p.cache)
    gxs, gys = [], []
    for lvl in pyr:
        gx, gy = gradients(lvl, p.grad_sigma, p.cache)
        gxs.append(gx)
        gys.append(gy)
    return pyr, gxs, gys


def scan_good_features(gx, gy, bx, by, hw, hh, skip):
    """ScanImageForGoodFeatures (goodFeaturesUtils.pyx:35-73) -> (val[ny,nx], xs, ys)."""
    gx = np.ascontiguousarray(gx, np.float32)
    gy = np.ascontiguousarray(gy, np.float32)
    H, W = gx.shape
    xs = np.arange(bx, W - bx, skip + 1, dtype=np.int32)
    ys = np.arange(by, H - by, skip + 1, dtype=np.int32)
    val = np.empty((len(ys), len(xs)), np.float32)
    nx, ny = C.c_int(), C.c_int()
    rc = lib().orc_scan_good_features(_f(gx), _f(gy), H, W, bx, by, hw, hh, skip, _f(val), C.byref(nx), C.byref(ny))
    assert rc == 0, rc
    assert nx.value == len(xs) and ny.value == len(ys)
    return val, xs, ys


def select_from_gradients(p, gx, gy, n_features, existing=None):
    """_KLTSelectGoodFeatures from the scan onwards (selectGoodFeatures.py:215-246).
    existing=(x,y,val) switches to replacement mode (overwriteAllFeatures=False)."""
    H, W = gx.shape
    window_hw, window_hh = p.window_width / 2, p.window_height / 2     # true division (3.5)
    bx, by = p.borderx, p.bordery
    if bx < window_hw:
        bx = window_hw
    if by < window_hh:
        by = window_hh
    bx, by, hw, hh = int(bx), int(by), int(window_hw), int(window_hh)  # truncation at the Cython boundary (Q3)
    val, xs, ys = scan_good_features(gx, gy, bx, by, hw, hh, p.nSkippedPixels)
    mindist = max(p.mindist, 0)
    if existing is None:
        fx = np.full(n_features, -1.0)
        fy = np.full(n_features, -1.0)
        fv = np.full(n_features, KLT_NOT_FOUND, np.int32)
        overwrite = 1
    else:
        fx = np.array(existing[0], np.float64)
        fy = np.array(existing[1], np.float64)
        fv = np.array(existing[2], np.int32)
        overwrite = 0
    used = lib().orc_select_from_scan(_f(val), len(xs), len(ys), bx, by, p.nSkippedPixels, W, H, mindist,
                                      float(p.min_eigenvalue), overwrite, n_features, _d(fx), _d(fy), _i(fv))
    assert used >= 0
    return fx, fy, fv, used


def select_from_map(p, val, W, H, n_features):
    """Sort + _enforceMinimumDistance (selectGoodFeatures.py:234-246) on a caller-provided eigenvalue map val[ny, nx] laid out
    like ScanImageForGoodFeatures' output for this geometry -- used by tests to separate the effect of the map's rounding
    from the rest of the selection."""
    window_hw, window_hh = p.window_width / 2, p.window_height / 2
    bx, by = max(p.borderx, window_hw), max(p.bordery, window_hh)
    bx, by = int(bx), int(by)
    val = np.ascontiguousarray(val, np.float32)
    fx = np.full(n_features, -1.0)
    fy = np.full(n_features, -1.0)
    fv = np.full(n_features, KLT_NOT_FOUND, np.int32)
    used = lib().orc_select_from_scan(_f(val), val.shape[1], val.shape[0], bx, by, p.nSkippedPixels, W, H, max(p.mindist, 0),
                                      float(p.min_eigenvalue), 1, n_features, _d(fx), _d(fy), _i(fv))
    assert used >= 0
    return fx, fy, fv


def select_good_features(p, img_u8, n_features):
    """KLTSelectGoodFeatures (selectGoodFeatures.py:141-261, 279-294) -> (x, y, val) arrays."""
    f = np.asarray(img_u8).astype(np.float32)
    if p.smoothBeforeSelecting:
        f = smooth(f, p.smooth_sigma(), p.cache)
    gx, gy = gradients(f, p.grad_sigma, p.cache)
    fx, fy, fv, _ = select_from_gradients(p, gx, gy, n_features)
    return fx, fy, fv


def _track_params(p):
    return _TrackParams(p.window_width, p.window_height, p.max_iterations, p.min_determinant, p.min_displacement,
                        p.step_factor, 0 if p.max_residue is None else 1,
                        0.0 if p.max_residue is None else p.max_residue, 1 if p.retainTrackers else 0,
                        p.nPyramidLevels, int(p.subsampling), float(p.borderx), float(p.bordery),
                        1 if getattr(p, "lighting_insensitive", False) else 0)


def track_on_pyramids(p, pyr1, pyr2, x, y, val):
    """KLTTrackFeatures per-feature loop (trackFeatures.py:250-346). pyrN = (img levels, gx levels, gy levels)."""
    L = p.nPyramidLevels
    fp = C.POINTER(C.c_float)
    arrs = []
    keep = []
    for pyr in (pyr1, pyr2):
        for comp in pyr:
            lv = [np.ascontiguousarray(a, np.float32) for a in comp]
            keep.append(lv)
            arrs.append((fp * L)(*[_f(a) for a in lv]))
    ncols = np.array([a.shape[1] for a in keep[0]], np.int32)
    nrows = np.array([a.shape[0] for a in keep[0]], np.int32)
    x = np.array(x, np.float64)
    y = np.array(y, np.float64)
    val = np.array(val, np.int32)
    tp = _track_params(p)
    it = lib().orc_track_features(C.byref(tp), arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], arrs[5], _i(ncols),
                                  _i(nrows), len(x), _d(x), _d(y), _i(val))
    if it == ORC_ASSERT:
        raise AssertionError("patch out of bounds (trackFeaturesUtils.pyx:35)")
    return x, y, val, it


def track_features(p, img1_u8, img2_u8, x, y, val, state=None):
    """KLTTrackFeatures (trackFeatures.py:205-409).  state: dict carrying 'pyramid_last' in sequential mode."""
    if p.sequentialMode and state is not None and state.get("pyramid_last") is not None:
        pyr1 = state["pyramid_last"]
    else:
        pyr1 = image_pyramids(p, img1_u8)
    pyr2 = image_pyramids(p, img2_u8)
    out = track_on_pyramids(p, pyr1, pyr2, x, y, val)
    if p.sequentialMode and state is not None:
        state["pyramid_last"] = pyr2
    return out


def extract_patch(img, x, y, h, w):
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((h, w), np.float32)
    ok = lib().orc_extract_patch(_f(img), img.shape[1], img.shape[0], x, y, w, h, _f(out))
    if not ok:
        raise AssertionError("patch out of bounds (trackFeaturesUtils.pyx:35)")
    return out


# ---------------------------------------------------------------------------------------------------
# Affine consistency check (trackFeatures.py:347-399) -- PARITY UNPINNED BY THE REFERENCE (its callees are
# undefined there); this drives the C-KLT 1.3.4 restatement in klt_oracle.c (orc_affine_step).
# ---------------------------------------------------------------------------------------------------
class AffineState:
    """What the reference keeps on each KLT_Feature: aff_img*, aff_x, aff_y, aff_A** (selectGoodFeatures.py:120-128)."""

    def __init__(self, n, aw=15, ah=15):
        self.n, self.aw, self.ah = n, aw, ah
        self.has = np.zeros(n, np.int32)
        self.aff_x = np.full(n, -1.0, np.float32)
        self.aff_y = np.full(n, -1.0, np.float32)
        self.A = np.tile(np.array([1, 0, 0, 1], np.float32), (n, 1))
        self.tmpl = np.zeros((n, 3, ah + 2, aw + 2), np.float32)

    def reset(self, mask):
        m = np.asarray(mask, bool)
        self.has[m] = 0
        self.aff_x[m] = -1.0
        self.aff_y[m] = -1.0
        self.A[m] = (1, 0, 0, 1)


def track_features_affine(p, img1_u8, img2_u8, x, y, val, aff, state=None):
    """KLTTrackFeatures with tc.affineConsistencyCheck >= 0: translational tracking, then the affine block per feature."""
    x_in = np.array(x, np.float64)
    y_in = np.array(y, np.float64)
    val_in = np.array(val, np.int32)
    if p.sequentialMode and state is not None and state.get("pyramid_last") is not None:
        pyr1 = state["pyramid_last"]
    else:
        pyr1 = image_pyramids(p, img1_u8)
    pyr2 = image_pyramids(p, img2_u8)
    x2, y2, v2, it = track_on_pyramids(p, pyr1, pyr2, x_in, y_in, val_in)
    if p.sequentialMode and state is not None:
        state["pyramid_last"] = pyr2
    ap = _AffineParams(int(p.affineConsistencyCheck), p.affine_window_width, p.affine_window_height,
                       p.affine_max_iterations, p.step_factor, p.min_determinant, p.min_displacement,
                       p.affine_min_displacement, p.affine_max_residue, p.affine_max_displacement_differ)
    i1, g1x, g1y = (np.ascontiguousarray(a[0], np.float32) for a in pyr1)
    i2, g2x, g2y = (np.ascontiguousarray(a[0], np.float32) for a in pyr2)
    nr, nc = i1.shape
    for f in range(len(x_in)):
        if val_in[f] < 0:
            continue
        if v2[f] != KLT_TRACKED:
            aff.has[f] = 0
            continue
        has = C.c_int(int(aff.has[f]))
        ax, ay = C.c_float(float(aff.aff_x[f])), C.c_float(float(aff.aff_y[f]))
        A = np.ascontiguousarray(aff.A[f])
        t = np.ascontiguousarray(aff.tmpl[f])
        st = lib().orc_affine_step(C.byref(ap), _f(i1), _f(g1x), _f(g1y), _f(i2), _f(g2x), _f(g2y), nc, nr,
                                   float(x_in[f]), float(y_in[f]), float(x2[f]), float(y2[f]), C.byref(has),
                                   C.byref(ax), C.byref(ay), _f(A), _f(t))
        if st == ORC_ASSERT:
            raise AssertionError("affine template leaves the image")
        aff.has[f], aff.aff_x[f], aff.aff_y[f] = has.value, ax.value, ay.value
        aff.A[f] = A
        aff.tmpl[f] = t
        if st != KLT_TRACKED:
            v2[f], x2[f], y2[f] = st, -1.0, -1.0
    return x2, y2, v2, it
