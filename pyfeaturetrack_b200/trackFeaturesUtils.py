"""Drop-in for the reference's Cython module trackFeaturesUtils (trackFeaturesUtils.pyx), backed by the CUDA
library.  The batched production path is trackFeatures.KLTTrackFeatures -> klt_track_features; the
functions here are the operator-level entry points the reference exposes."""
import numpy as np

from . import _capi
from .goodFeaturesUtils import _check_f32_2d


@_capi.serialized
def extractImagePatchSlow(img, x, y, height, width):
    """Bilinear (height x width) patch centred at fractional (x, y) (trackFeaturesUtils.pyx:14-51).
    Raises AssertionError when the window leaves the image, like the reference (:35)."""
    _check_f32_2d(img, "img")
    ctx = _capi.default_ctx()
    a = np.ascontiguousarray(img)
    out = np.empty((int(height), int(width)), np.float32)
    x, y = float(np.float32(x)), float(np.float32(y))      # 'float x, float y' arguments
    ctx.check(_capi.lib().klt_extract_patch(ctx.handle, a.ctypes.data, a.shape[1], a.shape[0], x, y, int(height),
                                           int(width), out.ctypes.data))
    return out


def _params_from_tc(tc):
    from .selectGoodFeatures import make_params
    return make_params(tc)


@_capi.serialized
def trackFeatureIterateCKLT(x2, y2, img1GradxPatch, img1GradyPatch, img1Patch, img2, gradx2, grady2, tc):
    """The Newton loop of one feature on caller-provided template patches (trackFeaturesUtils.pyx:393-459).
    -> (x2, y2, status, iteration) with x2, y2 Python floats holding float32 values."""
    import ctypes as C
    for a, name in ((img1GradxPatch, "img1GradxPatch"), (img1GradyPatch, "img1GradyPatch"), (img1Patch, "img1Patch"),
                    (img2, "img2"), (gradx2, "gradx2"), (grady2, "grady2")):
        _check_f32_2d(a, name)
    if tc.lighting_insensitive:
        raise Exception("Not implemented")                       # pyx:435
    ctx = _capi.default_ctx()
    p = _params_from_tc(tc)
    gxp, gyp, ip = (np.ascontiguousarray(a) for a in (img1GradxPatch, img1GradyPatch, img1Patch))
    i2, g2x, g2y = (np.ascontiguousarray(a) for a in (img2, gradx2, grady2))
    ox, oy, st, it = C.c_float(), C.c_float(), C.c_int32(), C.c_int32()
    ctx.check(_capi.lib().klt_track_iterate(ctx.handle, C.byref(p), float(np.float32(x2)), float(np.float32(y2)),
                                           gxp.ctypes.data, gyp.ctypes.data, ip.ctypes.data, i2.ctypes.data,
                                           g2x.ctypes.data, g2y.ctypes.data, i2.shape[1], i2.shape[0], C.byref(ox),
                                           C.byref(oy), C.byref(st), C.byref(it)))
    return ox.value, oy.value, st.value, it.value


@_capi.serialized
def _patch_combine(patch1, img2, x2, y2, workingPatch, mode):
    _check_f32_2d(patch1, "img1Patch")
    _check_f32_2d(img2, "img2")
    ctx = _capi.default_ctx()
    h, w = workingPatch.shape
    p1 = np.ascontiguousarray(patch1)
    im = np.ascontiguousarray(img2)
    out = np.empty(h * w, np.float32)
    ctx.check(_capi.lib().klt_patch_combine(ctx.handle, p1.ctypes.data, im.ctypes.data, im.shape[1], im.shape[0],
                                           float(np.float32(x2)), float(np.float32(y2)), h, w, mode, out.ctypes.data))
    return out


def computeIntensityDifference(img1Patch, img2, x2, y2, workingPatch, out):
    """out[:] = img1Patch - patch(img2 at (x2, y2)), row-major (trackFeaturesUtils.pyx:61-97)."""
    out[:] = _patch_combine(img1Patch, img2, x2, y2, workingPatch, 0)
    return None


def computeGradientSum(img1GradxPatch, gradx2, x2, y2, workingPatch, out, row):
    """out[:, row] = -img1GradxPatch - patch(gradx2 at (x2, y2)) (trackFeaturesUtils.pyx:107-142; square windows only,
    the reference's row stride is wrong otherwise -- quirk Q10)."""
    out[:, row] = _patch_combine(img1GradxPatch, gradx2, x2, y2, workingPatch, 1)
    return None
