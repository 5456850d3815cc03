"""Compile the C part of the oracle (gcc, no CUDA) into oracle/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libgccb_oracle.so")
SRC = os.path.join(HERE, "gccb_oracle.c")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
