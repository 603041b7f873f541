#!/usr/bin/env python3
"""Freeze outputs of the UNMODIFIED reference's swatch branch (02_color_extract.py:82-109) as tests/golden/swatch.npz.
The branch is unreachable through config.json (config.py drops `extraction_mode`), so main() is run with load_config
patched to return the reference's own Config plus the two attributes the branch reads with getattr.
Build container only (needs /root/reference):   PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_swatch.py"""
import contextlib
import importlib.util
import io
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor"
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth  # noqa: E402


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_ce", os.path.join(REF, "02_color_extract.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    import config as ref_config
    out = {}
    img = synth(96, 128, 11, cell=16)
    # swatches near colours that occur in the image, written as RGB (the usual case) and as BGR, plus an absent one
    px = img.reshape(-1, 3)
    cols = [tuple(int(v) for v in px[100][::-1]), tuple(int(v) for v in px[5000]), (int(px[9000][2]), int(px[9000][1]), int(px[9000][0])), (255, 0, 255)]
    names = ["layer_a", "layer_b", "layer_c", "layer_d"]
    for tol in (30, 60, 8):
        with tempfile.TemporaryDirectory() as d:
            cv2.imwrite(os.path.join(d, "resized.png"), img)
            cfg = ref_config.Config()
            cfg.output_dir = d
            cfg.color_names = list(names)
            cfg.colors = [list(c) for c in cols]
            cfg.extraction_mode = "swatch"
            cfg.color_tolerance = tol
            m.load_config = lambda: cfg
            with contextlib.redirect_stdout(io.StringIO()) as log:
                m.main()
            masks = np.stack([cv2.imread(os.path.join(d, n, "mask.png"), cv2.IMREAD_GRAYSCALE) for n in names])
            out[f"masks_tol{tol}"] = masks
            out[f"log_tol{tol}"] = np.array(log.getvalue())
            print(tol, [int((k > 0).sum()) for k in masks])
    out["img"] = img
    out["colors"] = np.array(cols, np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "swatch.npz"), **out)


if __name__ == "__main__":
    main()
