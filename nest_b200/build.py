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


# TEST AID: the same library with the warp-cooperative evaluation of the BTRS bound, the cooperative box
# rejection of the spectra and the table bounds of the photo-electron test switched off -- the plain per-lane
# algorithms.  tests/test_gpu_variants.py requires the product library to reproduce it bit for bit.
LIB_PLAIN = os.path.join(CSRC, "libnest_b200_plain.so")
PLAIN_FLAGS = ["-DNB_COOP_REJECT=0", "-DNB_PE_NO_BOUNDS"]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, plain=True):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    jobs = []
    for lib, extra in ((LIB, []), (LIB_PLAIN, PLAIN_FLAGS)):
        if lib == LIB_PLAIN and not plain:
            continue
        if not force and not needs_build(lib):
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
        jobs.append((lib, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for lib, p in jobs:  # the two compilations run side by side
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed building %s" % lib)
        if verbose:
            sys.stderr.write(out)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
