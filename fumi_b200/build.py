"""In-tree build of libfumi_b200.so (sm_100a only).

`python -m fumi_b200.build` or fumi_b200.build.build().  nvcc cross-compiles without a GPU; the
built library lands in fumi_b200/lib/ (git-ignored, shipped to the GPU box with the snapshot).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfumi_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
          "--expt-relaxed-constexpr", "-I", INCLUDE]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hdrs.append(os.path.join(INCLUDE, "fumi_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    dep_m = _deps_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_m):
            cmd = [NVCC, *ARCH, *COMMON, "-c", src, "-o", obj]
            if src.endswith(".cpp"):
                cmd = [NVCC, *COMMON, "-x", "c++", "-c", src, "-o", obj]
            if ptxas_info and src.endswith(".cu"):
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        logs = list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        logs.append(run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lpthread"]))
    if verbose or ptxas_info:
        for l in logs:
            if l.strip():
                print(l)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
