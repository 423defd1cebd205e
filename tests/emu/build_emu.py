"""TEST INFRASTRUCTURE ONLY: compile the product's kernel sources for the host with the CUDA
emulation shim (tests/emu/cuda_emu.h) into tests/emu/libfumi_emu.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fumi_b200", "csrc")
OUT = os.path.join(HERE, "libfumi_emu.so")
SRCS = ["episode.cu", "episode_fwd_f16.cu", "episode_bwd_f16.cu", "gram.cu", "dense.cu", "optim.cu", "am3.cu", "sampler.cpp", "sampler_expand.cu"]
TEST_SRCS = [os.path.join(ROOT, "tests", "csrc", "debug_gemm.cu")]


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in SRCS] + TEST_SRCS + [os.path.join(HERE, "cuda_emu.cpp")]
    deps = srcs + [os.path.join(HERE, "cuda_emu.h"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "launch.cuh"), os.path.join(CSRC, "warp_mma.cuh"), os.path.join(CSRC, "episode_common.cuh"),
                   os.path.join(ROOT, "include", "fumi_b200.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DFUMI_EMU", "-I", HERE, "-I", os.path.join(ROOT, "include"),
           "-Wno-unknown-pragmas", "-o", OUT]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
