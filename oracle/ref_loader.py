"""Imports the unmodified reference from oracle/_ref (see build_ref.py) -- TEST INFRASTRUCTURE ONLY.

oracle/_ref holds the reference's Python modules as sourceless bytecode (<name>.refbc) and its two Cython extension
modules (.so).  install() makes them importable under their own top-level names (klt, convolve, pyramid, ...),
exactly as a reference user would import them."""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import warnings

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PY_MODULES = ["error", "klt_util", "convolve", "klt", "pyramid", "selectGoodFeatures", "trackFeatures"]
OPTIONAL_MODULES = ["writeFeatures"]      # off the hot path; only the PPM overlay test reads it
EXT_MODULES = ["goodFeaturesUtils", "trackFeaturesUtils"]


def available():
    return all(os.path.exists(os.path.join(REF_DIR, m + ".refbc")) for m in PY_MODULES) and \
        all(any(f.startswith(m + ".") and f.endswith(".so") for f in os.listdir(REF_DIR)) for m in EXT_MODULES)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in PY_MODULES or fullname in OPTIONAL_MODULES:
            p = os.path.join(REF_DIR, fullname + ".refbc")
            if os.path.exists(p):
                return importlib.util.spec_from_file_location(
                    fullname, p, loader=importlib.machinery.SourcelessFileLoader(fullname, p))
        return None


_finder = None


def install():
    """Registers the reference modules' finder (idempotent).  Call in a process that does NOT also alias the product's
    drop-in modules under the same top-level names."""
    global _finder
    if not available():
        raise ImportError("oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference is mounted")
    warnings.simplefilter("ignore")          # the reference uses deprecated scipy.ndimage.filters
    if _finder is None:
        _finder = _Finder()
        sys.meta_path.insert(0, _finder)
        sys.path.insert(0, REF_DIR)          # for the two extension modules


def load():
    """-> dict name -> module, with the reference's progress prints silenced."""
    install()
    mods = {name: importlib.import_module(name) for name in PY_MODULES[:5] + EXT_MODULES + PY_MODULES[5:]}
    mods["selectGoodFeatures"].KLT_verbose = 0
    mods["trackFeatures"].KLT_verbose = 0
    for name in OPTIONAL_MODULES:
        if os.path.exists(os.path.join(REF_DIR, name + ".refbc")):
            mods[name] = importlib.import_module(name)
            if hasattr(mods[name], "KLT_verbose"):
                mods[name].KLT_verbose = 0
    return mods
