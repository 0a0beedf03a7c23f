"""Drop-in for the reference's Cython module trackFeaturesUtils (trackFeaturesUtils.pyx), backed by the CUDA
library.  The batched production path is trackFeatures.KLTTrackFeatures -> klt_track_features; the
functions here are the operator-level entry points the reference exposes."""
import numpy as np

from . import _capi
from .goodFeaturesUtils import _check_f32_2d


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
