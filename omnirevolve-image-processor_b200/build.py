"""Build libomni_b200.so (hand-written sm_100a CUDA + the C ABI of include/omni_b200.h) in-tree.

    python omnirevolve-image-processor_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libomni_b200.so")
SOURCES = ["capi.cu", "generic_kernels.cu", "fast_kernels.cu", "edges3.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false",                      # float paths must round every product and sum
              "-Xcompiler", "-fPIC,-fvisibility=default", "-shared", "-cudart", "static"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(os.path.dirname(HERE), "include", "omni_b200.h"))
    return d


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """out / defines: build an experimental variant (e.g. -DE3V_MINBLOCKS=4) next to the product library."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if out is not None:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
        return out
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(p) for p in _deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libomni_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
