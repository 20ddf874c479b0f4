"""Builds libmgb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m multigrid_prj_b200.build          # or: from multigrid_prj_b200.build import build_lib

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libmgb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-Xcompiler", "-fopenmp", "--use_fast_math=false",
]


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found")
    return n


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "mgb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, flags, verbose = args
    cmd = [_nvcc()] + flags + ["-c", "-o", obj, "-I", os.path.join(ROOT, "include"), src, "-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return obj


def build_lib(force=False, verbose=False):
    """one object per .cu (compiled in parallel, rebuilt only when the source or a header is newer), then one link"""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(PKG, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [os.path.join(ROOT, "include", "mgb200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj, flags, verbose))
    # the image's $CC/$CXX wrappers are not usable as nvcc host compilers for shared objects
    with ThreadPoolExecutor(max(len(jobs), 1)) as ex:
        list(ex.map(_compile_one, jobs))
    cmd = [_nvcc()] + flags[:2] + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp"]
    # NCCL is bound with dlopen at run time (csrc/nccl_dyn.h): no link-time dependency unless asked for
    cmd += ["-ldl"] + (["-lnccl"] if os.environ.get("MGB_LINK_NCCL", "0") == "1" else [])
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
