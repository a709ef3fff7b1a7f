"""Build the CPU oracle (oracle/liblfr_ref.so) — test infrastructure only.

g++ -O2 -ffp-contract=off: no FMA contraction, so the interpolator rounds like
the reference's unoptimised build (CMakeLists.txt:4 sets only `-std=c++11 -g`).

oracle/_ref (the reference's own sources compiled here) is NOT buildable:
solve.cc:20-31 needs Ceres, Eigen, COLMAP, Boost.program_options and generated
protobuf code, none of which exist in this image.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lfr_oracle.cc")
# the literal polynomial.cc restatement lives with the Ceres shim (oracle/ref_shims/)
SHIMS = os.path.join(HERE, "ref_shims")
MINI_CERES = os.path.join(SHIMS, "mini_ceres.cc")
OUT = os.path.join(HERE, "liblfr_ref.so")


def build(force: bool = False) -> str:
    deps = [SRC, MINI_CERES, os.path.join(SHIMS, "ceres", "ceres.h"), os.path.join(HERE, "..", "include", "lfr.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
           "-Wall", "-Wextra", "-I", SHIMS, "-o", OUT, SRC, MINI_CERES]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
