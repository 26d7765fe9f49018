"""Builds the oracle's C part (the host copy of the portable math routines) into oracle/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdie_portable_math.so")
SRC = os.path.join(HERE, "portable_math.c")
HDR = os.path.join(os.path.dirname(HERE), "die_b200", "csrc", "die_math.h")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        return LIB
    cmd = ["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
