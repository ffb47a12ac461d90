"""Builds oracle/c/libpgbp_oracle.so (gcc -O2 -march=native -fopenmp).  Test infrastructure.
The library is rebuilt when the source is newer OR when it was built on a different CPU model
(-march=native code must not travel between hosts)."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pgbp_oracle.c")
LIB = os.path.join(HERE, "libpgbp_oracle.so")
LIBQ = os.path.join(HERE, "libpgbp_oracle_quad.so")  # same source in IEEE binary128 (-DPGBPO_QUAD, libquadmath)
TAG = LIB + ".host"


def _host():
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False, quad=False):
    host = _host()
    lib = LIBQ if quad else LIB
    tag = lib + ".host"
    stale = (force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(SRC)
             or not os.path.exists(tag) or open(tag).read().strip() != host)
    if stale:
        extra = ["-DPGBPO_QUAD"] if quad else []
        subprocess.run(["gcc", "-O2", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off"]
                       + extra + ["-o", lib, SRC, "-lm"] + (["-lquadmath"] if quad else []), check=True)
        open(tag, "w").write(host)
    return lib


if __name__ == "__main__":
    print(build(force=True))
    print(build(force=True, quad=True))
