"""Compile oracle/tetra_oracle.c into oracle/_build/liboracle.so (TEST INFRASTRUCTURE)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tetra_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + f".{os.getpid()}.tmp"
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC",
           SRC, "-o", tmp, "-lm"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
