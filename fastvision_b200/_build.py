"""In-tree build of libfvb200.so with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfvb200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # parity: the reference's fp32 op order, no FMA contraction
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libfvb200.so cannot be built")
    return exe


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "fvb200.h"))
    return deps


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > lib_m for d in _deps())


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libfvb200.so next to this file."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = max(os.path.getmtime(h) for h in _deps() if h.endswith((".cuh", ".h")))
    objs, procs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
