"""TEST INFRASTRUCTURE ONLY: test-only sm_100a kernels (tests/csrc/*.cu) -> tests/lib/libfumi_test_kernels.so.

These kernels exercise device-side building blocks of the product (e.g. the fp16-plane warp GEMM of
fumi_b200/csrc/warp_mma.cuh) in isolation.  They are not part of libfumi_b200.so nor of include/fumi_b200.h.
The library resolves fumi_set_error / fumi_cuda_fail from libfumi_b200.so, which must be loaded RTLD_GLOBAL first
(`load()` below does both).  Built here without a GPU (nvcc cross-compiles); the .so travels to the GPU box.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libfumi_test_kernels.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    srcs = sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".cu"))
    csrc = os.path.join(ROOT, "fumi_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", "-o", OUT, *srcs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("test-kernel build failed:\n" + r.stdout + r.stderr)
    return OUT


def load():
    from fumi_b200 import build as product_build
    ctypes.CDLL(product_build.LIB, mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(build())
    P, I = ctypes.c_void_p, ctypes.c_int32
    L.fumi_debug_gemm_f16.restype = ctypes.c_int
    L.fumi_debug_gemm_f16.argtypes = [P, P, I, I, I, I, P, P]
    return L


if __name__ == "__main__":
    print(build(force=True))
