"""Build csrc/liblfr_b200.so in-tree with nvcc for sm_100a (cross-compiles
without a GPU).  The .so is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "liblfr_b200.so")
SOURCES = ["lfr_capi.cu", "lfr_wire.cc", "lfr_host.cc"]
DEPS = SOURCES + ["lfr_solve_warp.cuh", "lfr_solve_warp2.cuh", "lfr_solve_cta.cuh", "lfr_solve_tile.cuh", "lfr_math.cuh", os.path.join("..", "..", "include", "lfr.h"),
                  os.path.join("..", "..", "include", "lfr_wire.h"),
                  os.path.join("..", "..", "include", "lfr_host.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build_variant(out: str, defines) -> str:
    """Diagnostic variant of the library (e.g. -DLFR_POLY_PROF) next to the product .so."""
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"] + ["-D" + d for d in defines] + \
          ["-o", os.path.join(HERE, out)] + [os.path.join(HERE, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return os.path.join(HERE, out)


def build(force: bool = False, verbose: bool = False) -> str:
    newest = max(os.path.getmtime(os.path.join(HERE, d)) for d in DEPS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
           "-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    if "--poly-prof" in sys.argv:
        print(build_variant("liblfr_b200_polyprof.so", ["LFR_POLY_PROF"]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
