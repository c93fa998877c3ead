"""Build the C oracle (test infrastructure) into oracle/liboracle_dsf.so with gcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle_dsf.so")
SRCS = ["raster_oracle.c", "raster_oracle_impl.h", "pointface_oracle_impl.h", "intersect_oracle.c"]
UNITS = ["raster_oracle.c", "intersect_oracle.c"]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, s) for s in SRCS]
    if (not force and os.path.exists(SO)
            and os.path.getmtime(SO) >= max(os.path.getmtime(s) for s in srcs)):
        return SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", SO] + [os.path.join(HERE, u) for u in UNITS] + ["-lm"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True))
