"""Seeded synthetic textured frames; identical for the oracle and the GPU path.

Needs only NumPy + SciPy, so it runs identically in the build container and on the GPU box.
"""
import numpy as np


def _texture(H, W, seed, pad=64):
    import scipy.ndimage as ndi
    rng = np.random.default_rng(seed)
    n = rng.standard_normal((H + 2 * pad, W + 2 * pad)).astype(np.float32)
    return ndi.gaussian_filter(n, 2.0) + 1.5 * ndi.gaussian_filter(n, 6.0)


def frames(H, W, shifts, seed=0, pad=64):
    """uint8 (H, W) frames; frame k is the texture shifted by shifts[k] = (dy, dx) (cubic, reflect),
    normalised with frame 0's min/max."""
    import scipy.ndimage as ndi
    tex = _texture(H, W, seed, pad)
    out = []
    lo = hi = None
    for (dy, dx) in shifts:
        f = ndi.shift(tex, (dy, dx), order=3, mode="reflect")[pad:-pad, pad:-pad]
        if lo is None:
            lo, hi = float(f.min()), float(f.max())
        g = np.clip((f - lo) * (255.0 / (hi - lo)), 0, 255).astype(np.uint8)
        out.append(np.ascontiguousarray(g))
    return out


def frame_pair(H, W, seed=0, shift=(1.7, -3.3)):
    """The pair configs of SURVEY 8(d): frame 0 unshifted, frame 1 shifted by (dy, dx) = (1.7, -3.3)."""
    return frames(H, W, [(0.0, 0.0), shift], seed)


def sequence_shifts(nframes):
    """Sequence configs: dy = 20 sin(2 pi k/100), dx = 20 cos(2 pi k/100) - 20."""
    k = np.arange(nframes)
    return list(zip(20.0 * np.sin(2 * np.pi * k / 100.0), 20.0 * np.cos(2 * np.pi * k / 100.0) - 20.0))


def fast_frames(H, W, count, seed=0):
    """Cheap textured uint8 frames for throughput runs where generation time matters more than realism:
    one padded texture, integer-shifted crops plus a little per-frame noise (features still trackable)."""
    import scipy.ndimage as ndi
    rng = np.random.default_rng(seed)
    pad = 32
    n = rng.standard_normal((H + 2 * pad, W + 2 * pad)).astype(np.float32)
    tex = ndi.gaussian_filter(n, 2.0) + 1.5 * ndi.gaussian_filter(n, 6.0)
    lo, hi = float(tex.min()), float(tex.max())
    tex8 = np.clip((tex - lo) * (255.0 / (hi - lo)), 0, 255).astype(np.uint8)
    out = []
    for k in range(count):
        dy, dx = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
        out.append(np.ascontiguousarray(tex8[pad + dy:pad + dy + H, pad + dx:pad + dx + W]))
    return out
