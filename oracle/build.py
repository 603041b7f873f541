"""Build the plain-C oracle (TEST INFRASTRUCTURE ONLY) into oracle/_build/liboracle.so.

The reference itself is pure Python over OpenCV/NumPy (nothing to compile), so there is no
oracle/_ref; the "reference" leg is oracle/refport.py, which replays the reference's own
cv2/NumPy call sites.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "omni_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(HERE, "omni_tables.inc")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-shared", "-fPIC",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd, cwd=HERE)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
