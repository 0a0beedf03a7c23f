"""Drop-in for the reference's convolve.py: same names, same kernel-cache behaviour, arithmetic on the GPU.

The taps (_computeKernels, convolve.py:27-93) are scalar float64 math and stay on the host; the
convolutions themselves (scipy.ndimage.convolve1d in the reference, convolve.py:212-213) run as
sm_100a kernels behind klt_convolve_separable_f32 / klt_gradients_f32.
"""
from __future__ import print_function
import ctypes as C
import math

import numpy as np

from . import _capi
from . import config


class ConvolutionKernel:
    def __init__(self, maxKernelWidth=71):
        self.width = None
        self.data = [0. for i in range(maxKernelWidth)]


# the reference's single-slot cache (convolve.py:23-25); module globals so user code can inspect them
cachegauss = None
cachegaussderiv = None
cached_sigma_last = None


_kernel_memo = {}


def _computeKernels(sigma):
    """Gaussian and derivative-of-Gaussian taps, tails cut at 1 % of the peak (convolve.py:27-93).
    A pure function of sigma, so results are memoised; the reference's single-slot cache globals are still updated."""
    global cachegauss, cachegaussderiv, cached_sigma_last
    hit = _kernel_memo.get(sigma)
    if hit is not None:
        cachegauss, cachegaussderiv, cached_sigma_last = list(hit[0]), list(hit[1]), sigma
        return cachegauss, cachegaussderiv
    maxKernelWidth = 71
    factor = 0.01
    assert sigma >= 0.0
    half = maxKernelWidth // 2
    two_s2 = 2 * sigma * sigma
    gauss = [float(math.exp(-i * i / two_s2)) for i in range(-half, half + 1)]
    deriv = [-i * gauss[i + half] for i in range(-half, half + 1)]
    max_gauss = 1.0
    max_gaussderiv = float(sigma * math.exp(-0.5))

    # widths: drop symmetric pairs of taps while the outermost is below 1 % of the maximum
    gw, k = maxKernelWidth, 0
    while abs(gauss[k] / max_gauss) < factor:
        k += 1
        gw -= 2
    dw, k = maxKernelWidth, 0
    while abs(deriv[k] / max_gaussderiv) < factor:
        k += 1
        dw -= 2
    if gw == maxKernelWidth or dw == maxKernelWidth:
        # the reference calls an unimported KLTError here and dies with NameError (convolve.py:62)
        raise NameError("(_computeKernels) maxKernelWidth {0} is too small for a sigma of {1}".format(maxKernelWidth, sigma))
    g0, d0 = (maxKernelWidth - gw) // 2, (maxKernelWidth - dw) // 2
    gauss = gauss[g0:g0 + gw]
    deriv = deriv[d0:d0 + dw]

    den = 0.0
    for v in gauss:
        den += v
    gauss = [v / den for v in gauss]
    dhw = dw // 2
    den = 0.0
    for i in range(-dhw, dhw + 1):
        den -= i * deriv[i + dhw]
    deriv = [v / den for v in deriv]

    if len(_kernel_memo) < 256:
        _kernel_memo[sigma] = (tuple(gauss), tuple(deriv))
    cachegauss, cachegaussderiv, cached_sigma_last = gauss, deriv, sigma
    return gauss, deriv


def KLTGetKernelWidths(sigma):
    gauss_kernel, gaussderiv_kernel = _computeKernels(sigma)
    return len(gauss_kernel), len(gaussderiv_kernel)


def _kernels_for_gradients(sigma):
    """Cache rule of KLTComputeGradients (convolve.py:236): recompute only if sigma moved by more than 0.05."""
    if abs(sigma - cached_sigma_last) > 0.05:
        return _computeKernels(sigma)
    return cachegauss, cachegaussderiv


def _kernels_for_smoothing(sigma):
    """Cache rule of KLTComputeSmoothedImage (convolve.py:258)."""
    if cached_sigma_last is None or abs(sigma - cached_sigma_last) > 0.05:
        return _computeKernels(sigma)
    return cachegauss, cachegaussderiv


def _as_f32_image(img):
    a = np.asarray(img)
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    return np.ascontiguousarray(a, np.float32)


@_capi.serialized
def _convolveSeparate(imgin, horiz_kernel, vert_kernel, precision=None):
    """imgout = vert(horiz(imgin)), reflect borders, float32 between the passes (convolve.py:208-214)."""
    ctx = _capi.default_ctx()
    a = _as_f32_image(imgin)
    out = np.empty_like(a)
    hk, vk = _capi.Kernel1D.from_taps(horiz_kernel), _capi.Kernel1D.from_taps(vert_kernel)
    prec = config.operator_precision_code() if precision is None else precision
    ctx.check(_capi.lib().klt_convolve_separable_f32(ctx.handle, a.ctypes.data, a.shape[1], a.shape[0], C.byref(hk),
                                                    C.byref(vk), prec, out.ctypes.data))
    return out


@_capi.serialized
def KLTComputeGradients(img, sigma):
    """(gradx, grady) = (deriv_h o gauss_v, gauss_h o deriv_v) (convolve.py:226-248)."""
    gauss_kernel, gaussderiv_kernel = _kernels_for_gradients(sigma)
    ctx = _capi.default_ctx()
    a = _as_f32_image(img)
    gradx, grady = np.empty_like(a), np.empty_like(a)
    g, d = _capi.Kernel1D.from_taps(gauss_kernel), _capi.Kernel1D.from_taps(gaussderiv_kernel)
    ctx.check(_capi.lib().klt_gradients_f32(ctx.handle, a.ctypes.data, a.shape[1], a.shape[0], C.byref(g), C.byref(d),
                                           config.operator_precision_code(), gradx.ctypes.data, grady.ctypes.data))
    return gradx, grady


@_capi.serialized
def KLTComputeSmoothedImage(img, sigma):
    """gauss_h o gauss_v (convolve.py:254-264)."""
    gauss, _ = _kernels_for_smoothing(sigma)
    return _convolveSeparate(img, gauss, gauss)
