"""Drop-in for the reference's writeFeatures.py (off the hot path, host only).

KLTWriteFeatureListToPPM produces the reference's file byte for byte (writeFeatures.py:10-37; tests/test_host_logic.py compares
with the reference's own function).  KLTWriteFeatureList exists in the reference only as a stub that names undefined helpers
(`_printSetupTxt`, `_printHeader`, `_printFeatureTxt`, `_printSetupBin`, `binheader_fl`: writeFeatures.py:53-82); those are the
routines of C-KLT 1.3.4's writeFeatures.c, and this module writes and reads that format:

  text    "Feel free to place comments here." + warning line + banner + "nFeatures = N" + one line per feature,
          "%7d | (x,y)=%5d " with the caller's format ("%5.1f": floats; "%3d": positions rounded to the nearest integer
          unless negative)
  binary  "KLTFL1", int32 nFeatures, then per feature float32 x, float32 y, int32 val

PARITY UNPINNED: the reference cannot run these functions (NameError), so the format is restated from C-KLT, not checked
against reference output."""
from __future__ import print_function
import re
import struct
import sys

import numpy as np

from . import selectGoodFeatures as _sgf
from .klt import KLTCountRemainingFeatures

_WARNING_LINE = "!!! Warning:  This is a KLT data file.  Do not modify below this line !!!\n"
_BINHEADER_FL = b"KLTFL1"
_VAL_WIDTH = 5


def KLTWriteFeatureListToPPM(featurelist, greyimg, filename):
    """Overlay every live feature as a 3x3 red square on the grey image and save it (format by extension, like PIL)."""
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
            rgb[max(y - 1, 0):max(min(y + 2, nrows), 0), max(x - 1, 0):max(min(x + 2, ncols), 0)] = (255, 0, 0)
    Image.fromarray(rgb).save(filename)


def _parse_format(fmt):
    """C-KLT's _printSetupTxt: fmt must look like "%5.1f" or "%3d"; -> conversion type."""
    if not fmt or fmt[0] != "%" or fmt[-1] not in "fd":
        raise ValueError("(KLTWriteFeatures) Bad Format: {0}".format(fmt))
    return fmt[-1]


def _string_width(format_):
    """C-KLT's _findStringWidth: printed width of a format made of literal characters and %<width>[.prec]<type> fields."""
    width, i = 0, 0
    while i < len(format_):
        if format_[i] == "%":
            m = re.match(r"%(\d+)", format_[i:])
            if not m:
                raise ValueError("(_findStringWidth) Can't determine length of string")
            width += int(m.group(1))
            i += len(m.group(0))
            while i < len(format_) and format_[i] not in "df":
                i += 1
            i += 1
        else:
            width += 1
            i += 1
    return width


def _feature_text(feat, fmt, type_):
    x, y = float(np.float32(feat.x)), float(np.float32(feat.y))     # KLT_locType is float
    if type_ == "d":                                                # rounded to the nearest integer, unless negative
        x = int(x + 0.5) if x >= 0.0 else int(x)
        y = int(y + 0.5) if y >= 0.0 else int(y)
    return "(%s,%s)=%*d " % (fmt % x, fmt % y, _VAL_WIDTH, int(feat.val))


def KLTWriteFeatureList(fl, fname, fmt):
    """fname None: text to stderr; fmt None: binary file; otherwise a text file in C-KLT's layout."""
    if _sgf.KLT_verbose >= 1 and fname is not None:
        print("(KLT) Writing feature list to {0} file: '{1}'".format("binary" if fmt is None else "text", fname))
    if fmt is not None or fname is None:
        fmt = fmt if fmt is not None else "%5.1f"
        type_ = _parse_format(fmt)
        format_ = "(%s,%s)=%%%dd " % (fmt, fmt, _VAL_WIDTH)
        out = []
        if fname is not None:
            out += ["Feel free to place comments here.\n\n\n", "\n", _WARNING_LINE, "\n"]
        out += ["------------------------------\n", "KLT Feature List\n", "------------------------------\n\n",
                "nFeatures = %d\n\n" % len(fl), "feature | (x,y)=val\n", "--------+-" + "-" * _string_width(format_) + "\n"]
        for i, f in enumerate(fl):
            out.append("%7d | %s\n" % (i, _feature_text(f, fmt, type_)))
        text = "".join(out)
        if fname is None:
            sys.stderr.write(text)
        else:
            with open(fname, "w") as fh:
                fh.write(text)
    else:
        with open(fname, "wb") as fh:
            fh.write(_BINHEADER_FL)
            fh.write(struct.pack("<i", len(fl)))
            for f in fl:
                fh.write(struct.pack("<ffi", float(f.x), float(f.y), int(f.val)))


def KLTReadFeatureList(fname):
    """Reads either format written by KLTWriteFeatureList -> list of KLT_Feature (C-KLT's KLTReadFeatureList)."""
    from .klt import KLT_Feature
    with open(fname, "rb") as fh:
        data = fh.read()
    out = []
    if data.startswith(_BINHEADER_FL):
        n, = struct.unpack_from("<i", data, len(_BINHEADER_FL))
        for k in range(n):
            f = KLT_Feature()
            f.x, f.y, f.val = struct.unpack_from("<ffi", data, len(_BINHEADER_FL) + 4 + 12 * k)
            out.append(f)
        return out
    text = data.decode("ascii")
    body = text[text.index(_WARNING_LINE) + len(_WARNING_LINE):] if _WARNING_LINE in text else text
    m = re.search(r"nFeatures = (\d+)", body)
    if "KLT Feature List" not in body or not m:
        raise ValueError("(KLTReadFeatureList) File '{0}' does not contain a FeatureList".format(fname))
    rows = re.findall(r"^\s*(\d+) \| \(\s*(-?[\d.]+),\s*(-?[\d.]+)\)=\s*(-?\d+)", body, re.M)
    if len(rows) != int(m.group(1)):
        raise ValueError("(KLTReadFeatureList) expected {0} features, found {1}".format(m.group(1), len(rows)))
    for _, x, y, v in rows:
        f = KLT_Feature()
        f.x, f.y, f.val = float(x), float(y), int(v)
        out.append(f)
    return out
