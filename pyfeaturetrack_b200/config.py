"""Arithmetic-mode switches of the B200 implementation (there is nothing like this in the reference).

strict : convolutions accumulate in float64 in SciPy's exact operation order -> images, selected features,
         tracked positions and status codes are bit-identical to the reference's.
fast   : float32 FMA convolutions (images within ~5e-7 relative-to-max of the reference's); tracking then
         agrees to ~1e-4 px.  Selection order is sensitive to the last bit of the gradients,
         so selection and the operator-level functions default to strict; tracking defaults to auto.
         For SELECTION, fast means the fused pass: gradients, direct window sums and the minimum eigenvalue from one read
         of the smoothed image (no gradient planes, no summed-area tables); the selected SET agrees with the reference's
         for ~99.8 % of the features, the slots do not (the reference's order carries its float32 table rounding).
windowed : (tracking only) fast arithmetic on image-only pyramids: the gradient planes are not written; the tracker
         evaluates the gradient pair inside the windows the features visit.  Same results as fast to ~1e-5 px;
         anything that asks for a gradient plane builds it on demand.
auto   : (tracking only, the default) windowed unless the feature list is dense.  Measured on B200: skipping the planes
         saves ~11 us per 1080p pair (~5.3 ps per pixel) and costs ~1.75 ns per feature, so the planes pay off above
         roughly one feature per 330 pixels; the switch sits at one per 400 (5 184 features at 1080p).
"""
import os

from . import _capi

_MODES = {"fast": _capi.PRECISION_FAST, "strict": _capi.PRECISION_STRICT}
_TRACK_MODES = dict(_MODES, windowed=_capi.PRECISION_FAST_WINDOWED, auto=None)
AUTO_PIXELS_PER_FEATURE = 400

operator_precision = os.environ.get("KLT_B200_OPERATOR_PRECISION", "strict")
select_precision = os.environ.get("KLT_B200_SELECT_PRECISION", "strict")
track_precision = os.environ.get("KLT_B200_TRACK_PRECISION", "auto")


def set_precision(track=None, select=None, operator=None):
    global track_precision, select_precision, operator_precision
    for v in (select, operator):
        if v is not None and v not in _MODES:
            raise ValueError("precision must be 'fast' or 'strict'")
    if track is not None and track not in _TRACK_MODES:
        raise ValueError("track precision must be 'auto', 'windowed', 'fast' or 'strict'")
    if track is not None:
        track_precision = track
    if select is not None:
        select_precision = select
    if operator is not None:
        operator_precision = operator


def operator_precision_code():
    return _MODES[operator_precision]


def select_precision_code():
    """Precision of the pyramid build that feeds a selection: strict planes, or an image-only fast build."""
    return _capi.PRECISION_STRICT if select_precision == "strict" else _capi.PRECISION_FAST_WINDOWED


def select_mode_code():
    return _capi.SELECT_STRICT if select_precision == "strict" else _capi.SELECT_FAST


def track_precision_code(n_features=None, n_pixels=None):
    """Precision code for a tracking pyramid build; `auto` looks at the feature density of the call."""
    if track_precision != "auto":
        return _TRACK_MODES[track_precision]
    if n_features is not None and n_pixels and n_features * AUTO_PIXELS_PER_FEATURE > n_pixels:
        return _capi.PRECISION_FAST
    return _capi.PRECISION_FAST_WINDOWED
