"""Build libcnsn_b200.so in-tree with nvcc for sm_100a (no torch headers, no pybind), then the host binding
_cnsn_torch.so (C++ autograd nodes above the C ABI, csrc_torch/cnsn_torch.cpp) with g++ against the torch headers.

    python crossnorm-selfnorm_b200/build.py [--force]

libcnsn_b200.so is the C-ABI boundary declared in include/cnsn_b200.h.  Both are built next to this file so that
they travel with a snapshot of the repo (git-ignored, not shipped in history).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libcnsn_b200.so")
EXT = os.path.join(PKG, "_cnsn_torch.so")
EXT_SRC = os.path.join(PKG, "csrc_torch", "cnsn_torch.cpp")
OBJ = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-Xfatbin=-compress-all",          # the line tables of ~150 heavily inlined kernel instantiations compress 4-5x
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    d = [os.path.join(ROOT, "include", "cnsn_b200.h")]
    d += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return d


def up_to_date():
    if not os.path.isfile(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcnsn_b200.so")


def ext_up_to_date():
    if not os.path.isfile(EXT):
        return False
    t = os.path.getmtime(EXT)
    return all(os.path.getmtime(p) <= t for p in (EXT_SRC, os.path.join(ROOT, "include", "cnsn_b200.h"), LIB))


def build_ext(force=False, verbose=False):
    """g++ csrc_torch/cnsn_torch.cpp -> _cnsn_torch.so (links libcnsn_b200.so, torch, cudart).  Returns the path."""
    if not force and ext_up_to_date():
        return EXT
    import sysconfig
    import torch
    from torch.utils import cpp_extension as CE
    cuda_home = os.environ.get("CUDA_HOME") or os.path.dirname(os.path.dirname(nvcc_path()))
    inc = CE.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include")]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    tmp = EXT + ".tmp"
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-DTORCH_EXTENSION_NAME=_cnsn_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch.compiled_with_cxx11_abi()),
           *["-I" + p for p in inc], EXT_SRC, "-o", tmp,
           "-L" + PKG, "-l:libcnsn_b200.so", "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python", "-L" + os.path.join(cuda_home, "lib64"), "-lcudart",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for %s:\n%s\n%s" % (EXT_SRC, r.stdout[-3000:], r.stderr[-6000:]))
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(tmp, EXT)
    return EXT


def build(force=False, verbose=False, ext=True):
    """Compile every csrc/*.cu for sm_100a and link libcnsn_b200.so, then (ext=True) the torch host binding.
    Returns the library path."""
    lib = build_lib(force, verbose)
    if ext:
        build_ext(force, verbose)
    return lib


def build_lib(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
