"""Build the CPU oracle (test infrastructure only) into oracle/_build/libvo_oracle.so.

gcc -O2 -ffp-contract=off: every double operation is individually rounded so that
the integer / IEEE-exact parts of the restatement are bit-reproducible on the GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libvo_oracle.so")
SOURCES = ["harris.c", "klt.c", "p3p.c", "triangulation.c", "gftt.c"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in SOURCES if os.path.exists(os.path.join(HERE, s)))


def build(force: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-fPIC", "-shared",
           "-fopenmp", "-o", LIB] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
