"""Build the CPU oracle shared library (test infrastructure): gcc on oracle/snp_oracle.c -> oracle/_build/.

The reference is pure Python (no C sources), so there is no `oracle/_ref` binary to build: the
reference itself is exercised live by tests/golden/make_golden.py in the build container and its
recorded outputs are the pin (tests/golden/*.npz).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "snp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsnp_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-Wall",
           "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
