"""CPU test: libkltb200.so loads and exports every symbol include/klt_b200.h declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "klt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:int|int64_t|size_t|void \*|const char \*)\s*(klt_[a-z0-9_]+)\s*\(", src, flags=re.M)
    assert len(names) > 25
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from pyfeaturetrack_b200 import _capi
    lib = _capi.lib()
    for name in declared_functions():
        assert hasattr(lib, name), "libkltb200.so does not export " + name
        assert name in _capi.SIGNATURES, "ctypes binding lacks " + name
    assert sorted(_capi.SIGNATURES) == declared_functions()
    assert lib.klt_abi_version() == 4


def test_struct_layouts_match_header():
    """sizeof(klt_kernel1d / klt_taps / klt_params) as the C compiler sees them == the ctypes mirrors."""
    import subprocess, tempfile
    from pyfeaturetrack_b200 import _capi
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        with open(src, "w") as fh:
            fh.write('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu\\n", sizeof(klt_kernel1d), sizeof(klt_taps), sizeof(klt_params));return 0;}\n'
                     % os.path.join(ROOT, "include", "klt_b200.h"))
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-std=c11", "-o", exe, src])
        a, b, c = map(int, subprocess.check_output([exe]).split())
    assert (a, b, c) == (C.sizeof(_capi.Kernel1D), C.sizeof(_capi.Taps), C.sizeof(_capi.Params))


def test_no_cpu_fallback():
    """Without a CUDA device the context creation fails loudly (the product never routes through the oracle)."""
    from pyfeaturetrack_b200 import _capi
    try:
        ctx = _capi.Context(0)
    except _capi.KLTB200Error as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e) or "Blackwell" in str(e)
    else:
        ctx.close()       # a GPU is present: fine
    import pyfeaturetrack_b200
    pkg = os.path.dirname(pyfeaturetrack_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in text.replace("oracle/", "").lower() or fn in ("synth.py",), fn + " mentions the oracle"


def _build_c_demo(tmpdir):
    import subprocess
    pkg = os.path.join(ROOT, "pyfeaturetrack_b200")
    exe = os.path.join(str(tmpdir), "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
                           "-L" + pkg, "-lkltb200", "-Wl,-rpath," + pkg, "-lm", "-o", exe])
    return exe


def test_plain_c_program_links_against_the_abi(tmp_path):
    """examples/c_abi_demo.c compiles as C11 against include/klt_b200.h and links to libkltb200.so; without a GPU it
    stops with the library's own message (exit 3) -- the C-ABI boundary has no CPU fallback either."""
    import subprocess
    from pyfeaturetrack_b200 import _capi
    _capi.lib()                                   # builds the library if needed
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if r.returncode == 3:
        assert "no CPU fallback" in r.stderr or "Blackwell" in r.stderr or "CUDA" in r.stderr
    else:
        assert r.returncode == 0 and "tracked" in r.stdout, r.stderr
