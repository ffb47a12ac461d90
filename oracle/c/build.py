"""Builds oracle/c/libpgbp_oracle.so (gcc -O2 -fopenmp).  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pgbp_oracle.c")
LIB = os.path.join(HERE, "libpgbp_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", LIB, SRC, "-lm"],
                       check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
