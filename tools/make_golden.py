#!/usr/bin/env python3
"""Freeze outputs of the UNMODIFIED reference (stages 01-03 + process_colors.assign_labels) as
golden fixtures under tests/golden/.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

Pipeline cases run the reference exactly as a user would: `python pipeline.py IMG --output DIR
--start-step 1 --end-step 3` (one fresh subprocess per stage, reference/image_processor/pipeline.py:88-111),
with an optional pre-seeded DIR/config.json that pipeline.write_config merges (pipeline.py:29-40).
A second fresh process imports 02_color_extract._kmeans_lab to record the float32 centres and the
raw label map, which the stage itself never writes out.
"""
import importlib.util
import json
import os
import shutil
import subprocess
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor"
OUT = os.path.join(ROOT, "tests", "golden")


def synth(H, W, seed, cell=32):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H // cell, W // cell, 3), np.uint8)
    img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC)
    noise = rng.integers(-12, 13, img.shape, dtype=np.int16)
    return np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def run(cmd, env):
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise SystemExit(f"reference run failed: {cmd}")
    return r.stdout


KMEANS_SNIPPET = r"""
import sys, importlib.util, numpy as np, cv2
sys.path.insert(0, %(ref)r)
spec = importlib.util.spec_from_file_location("ce", %(ref)r + "/02_color_extract.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
img = cv2.imread(%(img)r, cv2.IMREAD_COLOR)
c, l = m._kmeans_lab(img, k=%(k)d)
np.savez(%(out)r, centers=c, labels=l.astype(np.uint8))
"""


def pipeline_case(name, img, cfg_over):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OPENCV_FOR_THREADS_NUM="1")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "input.png")
        cv2.imwrite(src, img)
        outdir = os.path.join(td, "out")
        os.makedirs(outdir)
        if cfg_over:
            # pre-seeded config.json: the runner merges CLI overrides into it (pipeline.py:29-40)
            with open(os.path.join(outdir, "config.json"), "w") as f:
                json.dump(cfg_over, f)
        log = run([sys.executable, os.path.join(REF, "pipeline.py"), src, "--output", outdir,
                   "--start-step", "1", "--end-step", "3"], env)
        cfg = json.load(open(os.path.join(outdir, "config.json")))
        names = cfg.get("color_names", ["layer_dark", "layer_mid", "layer_skin", "layer_light"])
        resized = cv2.imread(os.path.join(outdir, "resized.png"), cv2.IMREAD_COLOR)
        masks = np.stack([cv2.imread(os.path.join(outdir, n, "mask.png"), cv2.IMREAD_GRAYSCALE) for n in names])
        edges = np.stack([cv2.imread(os.path.join(outdir, n, "edges.png"), cv2.IMREAD_GRAYSCALE) for n in names])
        comp = cv2.imread(os.path.join(outdir, "edges_composite.png"), cv2.IMREAD_COLOR)
        palette = json.load(open(os.path.join(outdir, "palette_by_name.json")))
        km = os.path.join(td, "km.npz")
        run([sys.executable, "-c", KMEANS_SNIPPET % dict(ref=REF, img=os.path.join(outdir, "resized.png"),
                                                          k=max(2, len(names)), out=km)], env)
        kmd = np.load(km)
        cfg_keep = {k: cfg[k] for k in cfg if k not in ("input_image", "output_dir")}
        np.savez_compressed(os.path.join(OUT, name + ".npz"), input=img, resized=resized, masks=masks,
                            edges=edges, composite=comp, centers=kmd["centers"], labels=kmd["labels"])
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump({"config": cfg_keep, "seeded_config": cfg_over, "names": names, "palette_by_name": palette,
                       "log_tail": [l for l in log.splitlines() if "nz=" in l or "Resiz" in l or "No resize" in l]},
                      f, indent=1)
        print(name, "ok:", resized.shape, [int((m > 0).sum()) for m in masks], [int((e > 0).sum()) for e in edges])


def load_ref_module(fname, modname):
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, fname))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def function_cases():
    pc = load_ref_module("process_colors.py", "ref_process_colors")
    rz = load_ref_module("01_resize.py", "ref_resize")
    cfgm = sys.modules["config"]
    rng = np.random.default_rng(7)
    out = {}
    # process_colors.assign_labels incl. int16 wrap cases (black/white palette, random palettes)
    img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    img[:4] = 0
    img[4:8] = 255
    out["al_img"] = img
    for K in (2, 4, 8, 16):
        pal = rng.integers(0, 256, (K, 3), dtype=np.uint8)
        if K == 2:
            pal = np.array([[0, 0, 0], [255, 255, 255]], np.uint8)
        out[f"al_pal{K}"] = pal
        out[f"al_lab{K}"] = pc.assign_labels(img, pal)
    # 01 resize_if_needed on its three arithmetic paths (2:1, N:1, fractional) + no-op
    with tempfile.TemporaryDirectory() as td:
        for tag, (h, w, md) in {"rz_2to1": (96, 128, 64), "rz_3to1": (96, 144, 48), "rz_frac": (150, 100, 67),
                                "rz_frac2": (101, 203, 100), "rz_noop": (40, 50, 2000)}.items():
            src = synth(h, w, seed=h + w, cell=8)
            p = os.path.join(td, tag + ".png")
            cv2.imwrite(p, src)
            cfg = cfgm.Config(max_dimension=md)
            out[tag + "_src"] = src
            out[tag + "_md"] = np.int32(md)
            out[tag + "_dst"] = rz.resize_if_needed(p, cfg)
    np.savez_compressed(os.path.join(OUT, "functions.npz"), **out)
    print("functions ok:", sorted(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. dataclass defaults, K=4, no resize (config 1 of BASELINE.json, scaled down)
    pipeline_case("pipe_default_k4", synth(192, 256, seed=0, cell=16), None)
    # 1b. the tree's image_processor/config.json values (22/70, blur 7, light..dark name order), fractional resize
    tree_cfg = json.load(open(os.path.join(REF, "config.json")))
    tree_cfg = {k: v for k, v in tree_cfg.items() if k not in ("input_image", "output_dir")}
    tree_cfg["max_dimension"] = 150
    pipeline_case("pipe_treecfg_k4", synth(300, 200, seed=1, cell=20), tree_cfg)
    # 2. eight layers, exact 2:1 resize, blur 5, thresholds 100/200
    k8 = {"color_names": [f"layer_{i:02d}" for i in range(8)],
          "colors": [[(37 * i) % 256, (91 * i) % 256, (53 * i + 40) % 256] for i in range(8)],
          "max_dimension": 192, "edge_kernel_size": 5, "edge_low_threshold": 100, "edge_high_threshold": 200}
    pipeline_case("pipe_k8_2to1", synth(256, 384, seed=2, cell=16), k8)
    function_cases()
    print("golden bytes:", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))


if __name__ == "__main__":
    main()
