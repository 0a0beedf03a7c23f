"""Drop-in for the reference's Cython module goodFeaturesUtils (goodFeaturesUtils.pyx:35-73), backed by
klt_scan_good_features (summed-area tables + min-eigenvalue map on the GPU)."""
import numpy as np

from . import _capi


def _check_f32_2d(a, name):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % a.dtype.name)


@_capi.serialized
def scan_values(gradxArr, gradyArr, borderx, bordery, window_hw, window_hh, nSkippedPixels):
    """The min-eigenvalue map as arrays: (xs int32[nx], ys int32[ny], val float32[ny, nx])."""
    _check_f32_2d(gradxArr, "gradxArr")
    _check_f32_2d(gradyArr, "gradyArr")
    # float arguments are truncated at the Cython int boundary (quirk Q3)
    borderx, bordery, window_hw, window_hh, nSkippedPixels = (int(borderx), int(bordery), int(window_hw),
                                                              int(window_hh), int(nSkippedPixels))
    nrows, ncols = gradxArr.shape
    xs = np.arange(borderx, ncols - borderx, nSkippedPixels + 1, dtype=np.int32)
    ys = np.arange(bordery, nrows - bordery, nSkippedPixels + 1, dtype=np.int32)
    val = np.empty((len(ys), len(xs)), np.float32)
    ctx = _capi.default_ctx()
    gx = np.ascontiguousarray(gradxArr)
    gy = np.ascontiguousarray(gradyArr)
    ctx.check(_capi.lib().klt_scan_good_features(ctx.handle, gx.ctypes.data, gy.ctypes.data, ncols, nrows, borderx,
                                                bordery, window_hw, window_hh, nSkippedPixels, val.ctypes.data))
    return xs, ys, val


def ScanImageForGoodFeatures(gradxArr, gradyArr, borderx, bordery, window_hw, window_hh, nSkippedPixels):
    """-> (pointlistx, pointlisty, pointlistval): Python lists in the reference's row-major candidate order."""
    xs, ys, val = scan_values(gradxArr, gradyArr, borderx, bordery, window_hw, window_hh, nSkippedPixels)
    pointlistx, pointlisty = [], []
    for y in ys:
        pointlistx.extend(xs)
        pointlisty.extend(np.ones((len(xs),), np.int32) * y)
    return pointlistx, pointlisty, val.ravel().tolist()
