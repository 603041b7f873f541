#!/usr/bin/env python3
"""Freeze outputs of the UNMODIFIED reference's 04_find_contours.thinning_zhangsuen as tests/golden/thinning.npz.
Run in the build container only (needs /root/reference):   PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_thinning.py
Inputs: the edges.png planes already frozen in the pipeline golden cases (i.e. what stage 04 really reads) plus a few
synthetic shapes (thick blobs, a frame touching the border, single pixels)."""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_fc", os.path.join(REF, "04_find_contours.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_fc"] = m
    spec.loader.exec_module(m)
    cases = {}
    for name in ("pipe_default_k4", "pipe_k8_2to1"):
        z = np.load(os.path.join(OUT, name + ".npz"))
        for i in (0, len(z["edges"]) - 1):
            cases[f"{name}_e{i}"] = z["edges"][i]
        cases[f"{name}_m0"] = z["masks"][0]                  # a thick input: many iterations
    rng = np.random.default_rng(3)
    blob = (rng.random((70, 90)) < 0.55).astype(np.uint8) * 255
    cases["noise55"] = blob
    frame = np.zeros((40, 64), np.uint8); frame[:, :5] = 255; frame[:5] = 255; frame[-5:] = 255; frame[:, -5:] = 255
    cases["frame_on_border"] = frame
    dots = np.zeros((9, 33), np.uint8); dots[4, 4] = 255; dots[2:5, 20:23] = 255; dots[0, 32] = 200
    cases["dots"] = dots
    cases["empty"] = np.zeros((5, 7), np.uint8)
    out = {}
    for k, img in cases.items():
        with contextlib.redirect_stdout(io.StringIO()):
            sk = m.thinning_zhangsuen(img.copy(), layer=k)
        out[k + "_in"] = img
        out[k + "_out"] = sk
        print(k, img.shape, int((img > 0).sum()), "->", int((sk > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "thinning.npz"), **out)


if __name__ == "__main__":
    main()
