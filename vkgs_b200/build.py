"""Build libvkgsb.so (CUDA kernels + C ABI + C++ facade) in-tree for sm_100a.

    python -m vkgs_b200.build [--force]

nvcc cross-compiles without a GPU.  The library links the CUDA runtime statically, so it coexists with the runtime
PyTorch bundles.  project.cu and load.cu are compiled with -fmad=false: their arithmetic is pinned operation by
operation to oracle/vkgs_oracle.c (bit-exact visible set, keys and instance records).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libvkgsb.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
          "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall,-Wno-unused-function",
          "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]

# source -> extra flags
SOURCES = {
    "project.cu": ["-fmad=false"] + os.environ.get("VKGSB_PROJECT_FLAGS", "").split(),
    "load.cu": ["-fmad=false"],
    "lines.cu": ["-fmad=false"],
    "spatial.cu": [],
    "sort.cu": os.environ.get("VKGSB_SORT_FLAGS", "").split(),
    "bin.cu": os.environ.get("VKGSB_BIN_FLAGS", "").split(),
    "blend.cu": os.environ.get("VKGSB_BLEND_FLAGS", "").split(),
    "renderer.cu": [],
    "interop.cu": [],
    "ply.cc": [],
    "camera.cc": [],
    "engine.cc": [],
}
HEADERS = ["common.cuh", "kernels.h", "hostmath.h", "ply.h", os.path.join(ROOT, "include", "vkgsb.h"),
           os.path.join(ROOT, "include", "vkgs", "engine", "engine.h"),
           os.path.join(ROOT, "include", "vkgs", "scene", "camera.h")]


def _newest_header() -> float:
    t = os.path.getmtime(os.path.abspath(__file__))
    for h in HEADERS:
        p = h if os.path.isabs(h) else os.path.join(CSRC, h)
        if os.path.exists(p):
            t = max(t, os.path.getmtime(p))
    return t


def _compile(src: str, flags, force: bool, hdr_time: float, verbose: bool):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_time):
        return obj, ""
    cmd = [NVCC, *ARCH, *COMMON, *flags, "-Xptxas", "-v", "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = {s: f for s, f in SOURCES.items() if os.path.exists(os.path.join(CSRC, s))}
    hdr_time = _newest_header()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        futs = [ex.submit(_compile, s, f, force, hdr_time, verbose) for s, f in srcs.items()]
        results = [f.result() for f in futs]
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs, "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    _build_pygs(force)
    _build_viewer(force)
    return LIB


def _build_viewer(force: bool) -> str:
    """vkgs_b200/lib/vkgs_viewer: the reference's examples/vkgs_viewer.cc flow compiled against this repo's
    include/vkgs/engine/engine.h (tools/cpp/vkgs_viewer.cc) - the C++ drop-in check."""
    src = os.path.join(ROOT, "tools", "cpp", "vkgs_viewer.cc")
    out = os.path.join(HERE, "lib", "vkgs_viewer")
    if not os.path.exists(src):
        return ""
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-L", os.path.dirname(LIB), "-lvkgsb",
           "-lpthread", "-Wl,-rpath,$ORIGIN", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"viewer build failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return out


def _build_pygs(force: bool) -> str:
    """pygs/_pygs_cpp.so: the native half of the `pygs` package (pybind11 over vkgs::Engine), linked against libvkgsb."""
    import sysconfig

    import pybind11
    src = os.path.join(ROOT, "pygs", "_pygs_cpp.cc")
    out = os.path.join(ROOT, "pygs", "_pygs_cpp.so")
    if not os.path.exists(src):
        return ""
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I", pybind11.get_include(),
           "-I", sysconfig.get_paths()["include"], "-I", os.path.join(ROOT, "include"), src,
           "-L", os.path.dirname(LIB), "-lvkgsb", "-Wl,-rpath,$ORIGIN/../vkgs_b200/lib", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"pygs build failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
