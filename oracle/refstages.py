"""The reference's stages 01-03 AS SHIPPED -- own process per stage, PNG files in and out, process pool in stage 03 -- replayed
with oracle/refport.py's restatement of the stage arithmetic.  TEST / BASELINE INFRASTRUCTURE ONLY: the reference itself
(/root/reference/image_processor/{01_resize,02_color_extract,03_edge_detect}.py, launched by pipeline.py:88-111) is Python
that cannot travel to the GPU box, so bench.py's `stage_wall` record times THIS file there as the "reference CPU pipeline as a
user of pipeline.py sees it" (CPU timing (1)).  tests/test_refstages.py (build container, `reference` marker) checks
that its files equal the ones the unmodified reference writes.

    CONFIG_PATH=<out>/config.json python oracle/refstages.py 01|02|03
"""
from __future__ import annotations

import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor, as_completed

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refport as rp  # noqa: E402

DEFAULTS = {  # config.py:14-36, the keys stages 01-03 read
    "input_image": "input.png", "output_dir": "output", "n_cores": 12, "max_dimension": 2000,
    "color_names": ["layer_dark", "layer_mid", "layer_skin", "layer_light"],
    "colors": [[0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255]],
    "edge_low_threshold": 50, "edge_high_threshold": 150, "edge_kernel_size": 3,
    "edge_morph_kernel": 3, "edge_morph_open_iters": 1, "edge_morph_close_iters": 1,
}


def load_cfg() -> dict:
    cfg = dict(DEFAULTS)
    with open(os.environ["CONFIG_PATH"], "r", encoding="utf-8") as fh:
        cfg.update({k: v for k, v in json.load(fh).items() if k in DEFAULTS})
    return cfg


def stage01(cfg):
    """01_resize.py:7-31."""
    os.makedirs(cfg["output_dir"], exist_ok=True)
    for n in cfg["color_names"]:
        os.makedirs(os.path.join(cfg["output_dir"], n), exist_ok=True)
    img = cv2.imread(cfg["input_image"])
    if img is None:
        raise ValueError(f"Failed to load image: {cfg['input_image']}")
    out = rp.resize_if_needed(img, cfg["max_dimension"])
    cv2.imwrite(os.path.join(cfg["output_dir"], "resized.png"), out)


def stage02(cfg):
    """02_color_extract.py:66-175 (k-means mode)."""
    img = cv2.imread(os.path.join(cfg["output_dir"], "resized.png"), cv2.IMREAD_COLOR)
    if img is None:
        raise RuntimeError("Cannot read resized image")
    names = list(cfg["color_names"])
    K = max(2, len(names))
    centers = rp.kmeans_lab_centers(img, K, fresh_rng=False)         # a fresh process: the RNG is in its initial state
    labels = rp.assign_lab(img, centers)
    order, lut = rp.darkness_order(centers)
    centers_sorted, labels = centers[order], lut[labels]
    names_sorted = sorted(names, key=rp.darkness_rank)
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
    counts = [(labels == k).sum() for k in range(K)]
    palette = {}
    for name, k in zip(names_sorted, range(K)):
        os.makedirs(os.path.join(cfg["output_dir"], name), exist_ok=True)
        mask = (labels == k).astype(np.uint8) * 255
        mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, se, iterations=1)
        mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, se, iterations=1)
        cv2.imwrite(os.path.join(cfg["output_dir"], name, "mask.png"), mask)
        px = np.uint8([[list(centers_sorted[k].astype(np.uint8))]])
        bgr = cv2.cvtColor(px, cv2.COLOR_Lab2BGR)[0, 0]
        palette[name] = {"mode": "kmeans", "cluster_index": int(k), "cluster_lab": [int(v) for v in centers_sorted[k]],
                         "approx_bgr": [int(bgr[0]), int(bgr[1]), int(bgr[2])], "pixels": int(counts[k]),
                         "mask_nonzero": int(np.count_nonzero(mask))}
    with open(os.path.join(cfg["output_dir"], "palette_by_name.json"), "w", encoding="utf-8") as fh:
        json.dump(palette, fh, ensure_ascii=False, indent=2)


def _process_color(name, cfg):
    """03_edge_detect.py:13-40."""
    mask = cv2.imread(os.path.join(cfg["output_dir"], name, "mask.png"), cv2.IMREAD_GRAYSCALE)
    if mask is None:
        raise FileNotFoundError(name)
    edges = rp.edge_layer(mask, cfg["edge_low_threshold"], cfg["edge_high_threshold"], cfg["edge_kernel_size"],
                          cfg["edge_morph_kernel"], cfg["edge_morph_open_iters"], cfg["edge_morph_close_iters"])
    cv2.imwrite(os.path.join(cfg["output_dir"], name, "edges.png"), edges)
    return name


def stage03(cfg):
    """03_edge_detect.py:42-48 (process pool of n_cores workers) + :60-111 (composite from the files)."""
    with ProcessPoolExecutor(max_workers=cfg["n_cores"]) as ex:
        futs = [ex.submit(_process_color, n, cfg) for n in cfg["color_names"]]
        for f in as_completed(futs):
            f.result()
    resized = cv2.imread(os.path.join(cfg["output_dir"], "resized.png"))
    h, w = resized.shape[:2]
    canvas = np.full((h, w, 3), 255, np.uint8)
    for i, name in enumerate(cfg["color_names"]):
        edges = cv2.imread(os.path.join(cfg["output_dir"], name, "edges.png"), cv2.IMREAD_GRAYSCALE)
        m = edges > 0
        if not np.any(m):
            continue
        layer = np.zeros_like(canvas)
        layer[m] = tuple(cfg["colors"][i])
        canvas[m] = layer[m]
    cv2.imwrite(os.path.join(cfg["output_dir"], "edges_composite.png"), canvas)


if __name__ == "__main__":
    {"01": stage01, "02": stage02, "03": stage03}[sys.argv[1]](load_cfg())
