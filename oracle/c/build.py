"""Builds oracle/c/libpgbp_oracle.so (gcc -O2 -march=native -fopenmp).  Test infrastructure.
The library is rebuilt when the source is newer OR when it was built on a different CPU model
(-march=native code must not travel between hosts)."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pgbp_oracle.c")
LIB = os.path.join(HERE, "libpgbp_oracle.so")
TAG = LIB + ".host"


def _host():
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    host = _host()
    stale = (force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC)
             or not os.path.exists(TAG) or open(TAG).read().strip() != host)
    if stale:
        subprocess.run(["gcc", "-O2", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", LIB, SRC, "-lm"],
                       check=True)
        open(TAG, "w").write(host)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
