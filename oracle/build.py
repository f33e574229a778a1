"""Build recipe for the CPU oracle (test infrastructure).

`python -m oracle.build` compiles oracle/vf_oracle.c with gcc into
oracle/libvf_oracle.so.  The reference itself is pure Python (no C/C++ sources
under /root/reference), so there is nothing to compile into oracle/_ref/; the
Python reference is instead imported in the build container by
tests/golden/make_*_golden.py to pin this oracle.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "vf_oracle.c")
LIB = os.path.join(HERE, "libvf_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", LIB, SRC]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
