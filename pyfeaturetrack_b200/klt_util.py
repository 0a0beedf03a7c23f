"""klt_util.py of the reference (klt_util.py:3-34)."""
import numpy as np


def KLTComputeSmoothSigma(tc):
    return (tc.smooth_sigma_fact * max(tc.window_width, tc.window_height))


def KLTWriteFloatImageToPGM(img, filename):
    """Debug dump (klt_util.py:6-34): min/max-normalise a float image to 8 bits and save it.
    Accepts a PIL 'F' image (as the reference does) or a float ndarray (which the reference's callers pass)."""
    from PIL import Image
    a = np.asarray(img, np.float32)
    mmin, mmax = float(a.min()), float(a.max())
    fact = 255.0 / (mmax - mmin) if mmax != mmin else 1.0
    Image.fromarray(((a - mmin) * fact).astype(np.uint8)).save(filename)
