"""Drop-in for the reference's writeFeatures.py (off the hot path, host only).

KLTWriteFeatureListToPPM works like the reference's (writeFeatures.py:10-37).  KLTWriteFeatureList exists in the
reference only as a stub that calls undefined helpers (writeFeatures.py:53-82); here it writes C-KLT's text format
("%5.1f"-style positions, one feature per line) or a small binary format, enough for round trips in tests."""
from __future__ import print_function
import struct

import numpy as np

from . import selectGoodFeatures as _sgf
from .klt import KLTCountRemainingFeatures


def KLTWriteFeatureListToPPM(featurelist, greyimg, filename):
    """Overlay every live feature as a 3x3 red square on the grey image and save it (PPM by extension)."""
    from PIL import Image
    if isinstance(greyimg, np.ndarray):
        greyimg = Image.fromarray(greyimg)
    ncols, nrows = greyimg.size
    if _sgf.KLT_verbose:
        print("(KLT) Writing {0} features to PPM file: '{1}'".format(KLTCountRemainingFeatures(featurelist), filename))
    rgb = np.array(greyimg.convert("RGB"))
    for feat in featurelist:
        if feat.val >= 0:
            x, y = int(feat.x + 0.5), int(feat.y + 0.5)
            rgb[max(y - 1, 0):min(y + 2, nrows), max(x - 1, 0):min(x + 2, ncols)] = (255, 0, 0)
    Image.fromarray(rgb).save(filename)


_BIN_MAGIC = b"KLTFL1\n"


def KLTWriteFeatureList(fl, fname, fmt):
    """fmt like "%5.1f" or "%3d": text table 'index | (x,y)=val'; fmt None: binary (magic, int32 n, n x (f4 x, f4 y, i4 val))."""
    if _sgf.KLT_verbose >= 1 and fname is not None:
        print("(KLT) Writing feature list to {0} file: '{1}'".format("binary" if fmt is None else "text", fname))
    if fmt is not None:
        lines = ["Feature list: nFeatures = {0}".format(len(fl)), ""]
        for i, f in enumerate(fl):
            lines.append("%7d | (%s,%s)=%d" % (i, fmt % f.x, fmt % f.y, f.val))
        text = "\n".join(lines) + "\n"
        if fname is None:
            import sys
            sys.stderr.write(text)
        else:
            with open(fname, "w") as fh:
                fh.write(text)
    else:
        with open(fname, "wb") as fh:
            fh.write(_BIN_MAGIC)
            fh.write(struct.pack("<i", len(fl)))
            for f in fl:
                fh.write(struct.pack("<ffi", float(f.x), float(f.y), int(f.val)))


def KLTReadFeatureList(fname):
    """Reads the binary format written by KLTWriteFeatureList(fl, fname, None) -> list of KLT_Feature."""
    from .klt import KLT_Feature
    with open(fname, "rb") as fh:
        assert fh.read(len(_BIN_MAGIC)) == _BIN_MAGIC
        n, = struct.unpack("<i", fh.read(4))
        out = []
        for _ in range(n):
            f = KLT_Feature()
            f.x, f.y, f.val = struct.unpack("<ffi", fh.read(12))
            out.append(f)
    return out
