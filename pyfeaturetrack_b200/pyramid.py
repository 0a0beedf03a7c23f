"""Drop-in for the reference's pyramid.py (pyramid.py:14-77); the smoothing + decimation runs on the GPU."""
import ctypes as C

import numpy as np

from . import _capi
from . import config
from . import convolve
from .error import KLTError


class KLTPyramid:
    def __init__(self, ncols, nrows, subsampling, nlevels):
        if subsampling not in (2, 4, 8, 16, 32):
            KLTError("(_KLTCreatePyramid)  Pyramid's subsampling must " +
                     "be either 2, 4, 8, 16, or 32")
        self.subsampling = subsampling
        self.nLevels = nlevels
        self.img = []
        self.ncols = []
        self.nrows = []
        for i in range(nlevels):
            self.img.append(None)
            self.ncols.append(ncols)
            self.nrows.append(nrows)
            ncols /= subsampling      # true division: floats from level 1 on, as in the reference under Python 3
            nrows /= subsampling

    @_capi.serialized
    def Compute(self, img, sigma_fact):
        """Level 0 is `img` itself (no copy, quirk Q9); level i = smooth(level i-1, ss*sigma_fact) sampled at
        (ss*y + ss/2, ss*x + ss/2) (pyramid.py:37-77)."""
        img = np.asarray(img)
        nrows, ncols = img.shape[0], img.shape[1]
        subsampling = self.subsampling
        if subsampling not in (2, 4, 8, 16, 32):
            KLTError("(_KLTComputePyramid)  Pyramid's subsampling must " +
                     "be either 2, 4, 8, 16, or 32")
        assert self.ncols[0] == ncols
        assert self.nrows[0] == nrows
        self.img[0] = img
        if self.nLevels <= 1:
            return
        sigma = subsampling * sigma_fact
        # one cache lookup per level, exactly like the reference's KLTComputeSmoothedImage calls
        for _ in range(1, self.nLevels):
            gauss, _d = convolve._kernels_for_smoothing(sigma)
        ctx = _capi.default_ctx()
        taps = _capi.Taps()
        taps.pyramid = _capi.Kernel1D.from_taps(gauss)
        one = _capi.Kernel1D.from_taps([1.0])
        taps.smooth = taps.grad_gauss = taps.grad_deriv = one     # gradients are computed but not used here
        pyr = ctx.scratch_pyramid(ncols, nrows, self.nLevels, subsampling, 1, slot="pyramid.Compute")
        a = np.ascontiguousarray(img, np.float32)
        pyr.build_f32(a, taps, config.operator_precision_code(), already_smoothed=True)
        for i in range(1, self.nLevels):
            self.img[i] = pyr.download(0, i)


class DevicePyramid(KLTPyramid):
    """One component (0 intensity, 1 gradx, 2 grady) of a device-resident klt_pyr, presented with KLTPyramid's
    fields.  `.img[i]` downloads level i on first access, so tc.pyramid_last.img[0] keeps working."""

    class _LazyLevels(object):
        def __init__(self, owner):
            self._o = owner
            self._cache = {}

        def __len__(self):
            return self._o.nLevels

        def __getitem__(self, i):
            if isinstance(i, slice):
                return [self[j] for j in range(*i.indices(len(self)))]
            if i < 0:
                i += len(self)
            if not 0 <= i < len(self):
                raise IndexError(i)
            if i not in self._cache:
                self._cache[i] = self._o.pyr.download(self._o.which, i, self._o.image)
            return self._cache[i]

        def __iter__(self):
            return (self[i] for i in range(len(self)))

    def __init__(self, pyr, which, image=0):
        KLTPyramid.__init__(self, pyr.w, pyr.h, pyr.subsampling, pyr.n_levels)
        self.pyr, self.which, self.image = pyr, which, image
        self.img = DevicePyramid._LazyLevels(self)
