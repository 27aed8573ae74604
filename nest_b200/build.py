"""In-tree build of the CUDA library (nvcc cross-compiles sm_100a without a GPU).

-fmad=false: the deterministic yield functions are held to 1e-12 against a reference
compiled for x86-64 without FMA contraction, and LArNEST's exact `Lindhard == 1.`
comparison needs every IEEE operation rounded on its own; the hot per-hit loop uses
explicit fma() where fusion is wanted.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libnest_b200.so")
SOURCES = ["nb_capi.cu"]
HEADERS = ["nb_rng.cuh", "nb_params.h", "nb_yields.cuh", "nb_chain.cuh", "nb_lar.cuh", "nb_kernels.cuh",
           "nb_host.h", os.path.join("..", "..", "include", "nest_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-ldl"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building %s" % LIB)
    if verbose:
        sys.stderr.write(r.stdout)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
