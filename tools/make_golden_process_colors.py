#!/usr/bin/env python3
"""Freeze what the UNMODIFIED reference CLI `process_colors.py` writes, as tests/golden/process_colors.npz (+ .json):
    python /root/reference/image_processor/process_colors.py IMG -o OUT -m adaptive -n 5
    python /root/reference/image_processor/process_colors.py IMG -o OUT -m palette --palette analyzer.json       (recommended_colors form,
                                                                                      the schema analyze_colors.py:395-406 writes)
    python /root/reference/image_processor/process_colors.py IMG -o OUT -m palette --palette eight.json -n 3 --edges-only   (the --colors warning)
each in a fresh subprocess on a 256x384 synthetic image.  Build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_process_colors.py"""
import glob
import json
import os
import subprocess
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor/process_colors.py"
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth  # noqa: E402

ANALYZER = {"recommended_colors": [          # out of order on purpose: the reader sorts by position
    {"position": 3, "name": "sky", "rgb": [70, 130, 200]},
    {"position": 1, "name": "ink", "rgb": [20, 20, 30]},
    {"position": 4, "name": "paper", "rgb": [240, 235, 220]},
    {"position": 2, "name": "rust", "rgb": [180, 80, 40]},
    {"name": "leaf", "rgb": [60, 160, 70]},                                   # no position: sorts last
    {"position": 5, "name": "lilac", "rgb": [170, 140, 210]}]}
EIGHT = {"recommended_colors": [{"position": i + 1, "name": n, "rgb": c} for i, (n, c) in enumerate([
    ("k", [0, 0, 0]), ("w", [255, 255, 255]), ("r", [200, 30, 30]), ("b", [30, 30, 200]), ("grey", [128, 128, 128]),
    ("y", [250, 220, 40]), ("g", [20, 150, 60]), ("m", [255, 0, 255])])]}
# the second JSON form of process_colors.py:60-64 ({"palette": [{rgb, name}]}) raises in the reference under Python 3: line 63 reads
# the loop variable `c` of the comprehension on line 62 outside of it.  Recorded as such; the mirror implements what the line means.
GENERIC = {"palette": [{"rgb": [0, 0, 0], "name": "k"}, {"rgb": [255, 255, 255]}, {"rgb": [200, 30, 30], "name": "r"}]}


def run(args, cwd, expect_ok=True):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, REF] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert (r.returncode == 0) == expect_ok, r.stdout
    return r.stdout


def collect(out_dir):
    labels = np.load(os.path.join(out_dir, "labels.npy"))
    assert np.array_equal(cv2.imread(os.path.join(out_dir, "labels.png"), cv2.IMREAD_UNCHANGED), labels)
    layers = sorted(os.path.basename(p) for p in glob.glob(os.path.join(out_dir, "layer_*.png")))
    for name in layers:                                    # the layers are exactly (labels == i) * 255
        i = int(name.split("_")[1]) - 1
        assert np.array_equal(cv2.imread(os.path.join(out_dir, name), cv2.IMREAD_UNCHANGED), (labels == i).astype(np.uint8) * 255)
    return labels, layers, json.load(open(os.path.join(out_dir, "palette.json")))


def main():
    img = synth(256, 384, 21, cell=16)
    out, meta = {"input": img}, {}
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "in.png")
        cv2.imwrite(src, img)
        json.dump(ANALYZER, open(os.path.join(td, "analyzer.json"), "w"))
        json.dump(GENERIC, open(os.path.join(td, "generic.json"), "w"))
        json.dump(EIGHT, open(os.path.join(td, "eight.json"), "w"))
        cases = {"adaptive5": ["-m", "adaptive", "-n", "5"],
                 "analyzer": ["-m", "palette", "--palette", os.path.join(td, "analyzer.json"), "-n", "6"],
                 "eight": ["-m", "palette", "--palette", os.path.join(td, "eight.json"), "-n", "3", "--edges-only"]}
        for tag, extra in cases.items():
            od = os.path.join(td, tag)
            log = run([src, "-o", od] + extra, td)
            labels, layers, pal = collect(od)
            out[f"labels_{tag}"] = labels
            # log lines without the run-specific paths
            keep = [ln for ln in log.splitlines() if ln.startswith("  [") or ln.startswith("[WARN]") or "Size:" in ln or "NOTE" in ln or "Done." in ln]
            meta[tag] = {"args": [a.replace(td, "@TD@") for a in extra], "layers": layers, "palette": pal,
                         "log": [ln.replace(od, "@OUT@") for ln in keep]}
            print(tag, labels.shape, np.bincount(labels.ravel()).tolist(), layers)
        log = run([src, "-o", os.path.join(td, "generic"), "-m", "palette", "--palette", os.path.join(td, "generic.json")], td, expect_ok=False)
        meta["generic_form_reference_error"] = log.strip().splitlines()[-1]
        print("generic form:", meta["generic_form_reference_error"])
    meta["analyzer_json"], meta["eight_json"], meta["generic_json"] = ANALYZER, EIGHT, GENERIC
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "process_colors.npz"), **out)
    json.dump(meta, open(os.path.join(ROOT, "tests", "golden", "process_colors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
