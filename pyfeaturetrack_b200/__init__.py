"""pyfeaturetrack_b200 -- PyFeatureTrack's KLT hot path on NVIDIA B200 (sm_100a).

The submodules carry the reference's module and symbol names (klt, convolve, pyramid, selectGoodFeatures,
trackFeatures, goodFeaturesUtils, trackFeaturesUtils, klt_util, error).  install_dropin() additionally
registers them under those TOP-LEVEL names, so reference user code (`from klt import *`,
`from selectGoodFeatures import *`, ...) runs unchanged on the GPU.

There is no CPU fallback: the CUDA library (libkltb200.so, built by `python -m pyfeaturetrack_b200.build`)
and a B200 are required for every compute call.
"""
import importlib
import sys

__version__ = "0.1.0"

DROPIN_MODULES = ["error", "klt_util", "convolve", "klt", "pyramid", "goodFeaturesUtils", "trackFeaturesUtils",
                  "selectGoodFeatures", "trackFeatures", "writeFeatures", "storeFeatures"]


def install_dropin():
    """Alias the drop-in modules under the reference's top-level module names."""
    for name in DROPIN_MODULES:
        sys.modules[name] = importlib.import_module(__name__ + "." + name)


def set_device(device):
    from . import _capi
    _capi.set_device(device)


def set_precision(track=None, select=None, operator=None):
    from . import config
    config.set_precision(track, select, operator)
