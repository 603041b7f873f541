#!/usr/bin/env python3
"""Freeze outputs of the UNMODIFIED reference's 04_find_contours.trace_centerlines as tests/golden/trace.npz: for every skeleton of
tests/golden/thinning.npz (what stage 04 traces) plus a non-skeleton noise image, the polylines (concatenated points + lengths) and
the log with the timings blanked.  Build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_trace.py"""
import contextlib
import importlib.util
import io
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor"
OUT = os.path.join(ROOT, "tests", "golden")


def blank(log):
    return "\n".join(ln for ln in re.sub(r"[0-9.]+s", "Xs", log).splitlines() if "visited" not in ln)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_fc", os.path.join(REF, "04_find_contours.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_fc"] = m
    spec.loader.exec_module(m)
    z = np.load(os.path.join(OUT, "thinning.npz"))
    cases = {k[:-4]: z[k] for k in z.files if k.endswith("_out")}
    cases["noise30"] = (np.random.default_rng(5).random((48, 64)) < 0.3).astype(np.uint8) * 255      # thick: junction-rich
    out = {}
    for k, sk in cases.items():
        with contextlib.redirect_stdout(io.StringIO()) as log:
            paths = m.trace_centerlines(sk.copy(), layer=k)
        out[k + "_skel"] = sk
        out[k + "_len"] = np.array([len(p) for p in paths], np.int32)
        out[k + "_pts"] = np.concatenate([p.reshape(-1, 2) for p in paths]).astype(np.int32) if paths else np.zeros((0, 2), np.int32)
        out[k + "_log"] = np.array(blank(log.getvalue()))
        print(k, sk.shape, int((sk > 0).sum()), "px ->", len(paths), "polylines")
    np.savez_compressed(os.path.join(OUT, "trace.npz"), **out)


if __name__ == "__main__":
    main()
