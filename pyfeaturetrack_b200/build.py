"""Builds libkltb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkltb200.so")
SOURCES = ["klt_affine.cu", "klt_api.cu", "klt_conv.cu", "klt_probe.cu", "klt_select.cu", "klt_select_fast.cu", "klt_sequence.cu", "klt_stream.cu", "klt_track.cu", "klt_track_windowed.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=true",     # FMA contraction only where the code does not use *_rn intrinsics; never --use_fast_math
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "klt_common.cuh"), os.path.join(HERE, "..", "include", "klt_b200.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        build_featlist()
        return LIB
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= _newest(deps[len(srcs):] + [s]):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    build_featlist(force=True)
    return LIB


FEATLIST_LIB = os.path.join(HERE, "libkltfeatlist.so")


def build_featlist(force=False):
    """Host glue of the drop-in API (feature list <-> arrays, CPython C API, no compute): gcc only."""
    import sysconfig
    src = os.path.join(CSRC, "featlist.c")
    if not force and os.path.exists(FEATLIST_LIB) and os.path.getmtime(FEATLIST_LIB) >= os.path.getmtime(src):
        return FEATLIST_LIB
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-Wall", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], src,
           "-o", FEATLIST_LIB]
    subprocess.check_call(cmd)
    return FEATLIST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
