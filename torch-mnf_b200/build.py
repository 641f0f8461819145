"""Build libmnf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python torch-mnf_b200/build.py [--force] [-v]

The .so lands in torch-mnf_b200/lib/ (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  Object files are cached per source in torch-mnf_b200/build/ and rebuilt when
the source or any header is newer.  ptxas -v output per source is kept next to the objects.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SUFFIX = os.environ.get("MNF_LIB_SUFFIX", "")  # experiment builds: separate objects and .so name
OBJ = os.path.join(HERE, "build" + SUFFIX)
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, f"libmnf_b200{SUFFIX}.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "mnf_b200.h"))
    extra = os.environ.get("MNF_NVCC_FLAGS", "").split()

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs):
            cmd = [NVCC, *ARCH, *FLAGS, *extra, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(srcs)))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
