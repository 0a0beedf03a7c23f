#!/usr/bin/env python
"""Build the UNMODIFIED reference into binaries under oracle/_ref/ (test infrastructure only).

The reference (TimSC/PyFeatureTrack, mounted read-only at /root/reference) is Python plus two
Cython extension modules (setup.py:8-9).  /root/reference does not exist on the GPU box, and reference
SOURCES must never be copied into this repository, so this recipe compiles every module of the hot
path *from the sources where they lie* into extension modules (.so) and writes ONLY those binaries
into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun):

  * goodFeaturesUtils.pyx, trackFeaturesUtils.pyx  - exactly what the reference's setup.py builds;
  * klt.py, convolve.py, pyramid.py, klt_util.py, error.py, selectGoodFeatures.py, trackFeatures.py
    - byte-compiled unmodified by CPython itself (py_compile) into SOURCELESS bytecode files
      (<name>.refbc = a .pyc under another suffix, because the gpurun snapshot drops *.pyc), so the
      interpreter executes exactly the reference's bytecode at exactly the reference's speed
      ("the reference as executed under Python 3").  oracle/ref_loader.py imports them.

Intermediate C files (which quote source lines in comments) go to a temp dir and are discarded.
The two PGM fixtures are NOT copied; tests/golden/ holds what the tests need.

Only tests/, __graft_entry__.smoke()/build() and bench.py's CPU-baseline legs may use oracle/_ref.
"""
import os, shutil, subprocess, sys, sysconfig, tempfile

REF = os.environ.get("KLT_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PYX = ["goodFeaturesUtils.pyx", "trackFeaturesUtils.pyx"]
PY = ["klt.py", "convolve.py", "pyramid.py", "klt_util.py", "error.py",
      "selectGoodFeatures.py", "trackFeatures.py", "writeFeatures.py"]


def _target(f):
    base, ext = os.path.splitext(f)
    return base + (sysconfig.get_config_var("EXT_SUFFIX") if ext == ".pyx" else ".refbc")


def up_to_date():
    if not os.path.isdir(OUT):
        return False
    for f in PYX + PY:
        so = os.path.join(OUT, _target(f))
        if not os.path.exists(so):
            return False
        src = os.path.join(REF, f)
        if os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so):
            return False
    return True


def build(force=False, verbose=False):
    """Returns True if oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF):
        return up_to_date()          # GPU box: use the prebuilt binaries that travelled with the snapshot
    if up_to_date() and not force:
        return True
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (availability check)
    tmp = tempfile.mkdtemp(prefix="klt_ref_build_")
    try:
        # stage the sources OUTSIDE the repo (read-only mount cannot take build products)
        for f in PYX:
            shutil.copy(os.path.join(REF, f), os.path.join(tmp, f))
        setup_py = os.path.join(tmp, "_setup_ref.py")
        with open(setup_py, "w") as fh:
            fh.write(
                "from setuptools import setup, Extension\n"
                "from Cython.Build import cythonize\n"
                "import numpy as np\n"
                "srcs = %r\n"
                "exts = [Extension(s.rsplit('.',1)[0], [s], include_dirs=[np.get_include()],\n"
                "                  extra_compile_args=['-O2'],\n"
                "                  ) for s in srcs]\n"
                "setup(name='klt_ref', ext_modules=cythonize(exts, quiet=True),\n"
                "      script_args=['build_ext','--inplace'])\n" % (PYX,))
        r = subprocess.run([sys.executable, setup_py], cwd=tmp, capture_output=not verbose, text=True)
        if r.returncode != 0:
            if not verbose:
                sys.stderr.write(r.stdout[-4000:] + r.stderr[-4000:])
            return False
        os.makedirs(OUT, exist_ok=True)
        for f in PYX:
            shutil.copy(os.path.join(tmp, _target(f)), os.path.join(OUT, _target(f)))
        import py_compile
        for f in PY:   # sourceless bytecode next to the extension modules; imported through oracle/ref_loader.py
            py_compile.compile(os.path.join(REF, f), cfile=os.path.join(OUT, _target(f)),
                               dfile="<reference>/" + f, doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("oracle/_ref:", "ok" if ok else "UNAVAILABLE")
    sys.exit(0 if ok else 1)
