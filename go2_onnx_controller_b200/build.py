"""In-tree build of the native libraries (sm_100a only).

  lib/libgo2policy.so   CUDA kernels + C ABI  (csrc/capi.cu, csrc/onnx_reader.cpp)
  lib/libonnx_actor.so  the reference-compatible C++ class (csrc/onnx_actor.cpp) over the C ABI
  lib/go2_smoke         equivalent of the reference's onnx_inference smoke executable

nvcc cross-compiles without a GPU; the .so files travel with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.environ.get("GO2P_LIB") or os.path.join(LIBDIR, "libgo2policy.so")
ACTOR_LIB = os.path.join(LIBDIR, "libonnx_actor.so")
SMOKE = os.path.join(LIBDIR, "go2_smoke")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")


def sources():
    cs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    inc = [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))]
    return cs + inc


def build(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    if force or _newer(LIB, srcs):
        flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
        _run([_nvcc(), *flags, "-shared", "-o", LIB, os.path.join(CSRC, "capi.cu"), os.path.join(CSRC, "onnx_reader.cpp")])
    gxx = shutil.which("g++") or "g++"
    if force or _newer(ACTOR_LIB, srcs + [LIB]):
        _run([gxx, "-std=c++20", "-O2", "-fPIC", "-shared", "-o", ACTOR_LIB, os.path.join(CSRC, "onnx_actor.cpp"),
              "-L" + LIBDIR, "-lgo2policy", "-Wl,-rpath,$ORIGIN"])
    if force or _newer(SMOKE, srcs + [ACTOR_LIB]):
        _run([gxx, "-std=c++20", "-O2", "-o", SMOKE, os.path.join(CSRC, "smoke_main.cpp"),
              "-L" + LIBDIR, "-lonnx_actor", "-lgo2policy", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
