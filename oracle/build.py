"""Build the C part of the oracle: oracle/pointnet2_cpu.c -> oracle/liboracle_pointnet2.so.

Test infrastructure only.  `-ffp-contract=off` keeps gcc from fusing anything the
source does not fuse explicitly with fmaf() (see the header of pointnet2_cpu.c).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pointnet2_cpu.c")
LIB = os.path.join(HERE, "liboracle_pointnet2.so")


def build(force: bool = False) -> str:
    if (not force) and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fvisibility=hidden",
           "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
