"""Mirror of the reference's stand-alone `process_colors.py` (v1.1.1): strict one-hot colour layers + a label index map.

Same command line, same files (`labels.png`, `labels.npy`, `palette.json`, `layer_<i>_<name>.png`), same log lines and
exception types as process_colors.py:89-179; the per-pixel work runs on the GPU through the C ABI:

  process_colors.py:69-77  assign_labels   -> omni_assign_rgb_i16wrap   (the int16 wrap of diff*diff is reproduced)
  process_colors.py:143    class histogram -> omni_count_nonzero on the one-hot planes
  process_colors.py:169-175 (labels == i) * 255, no morphology -> omni_layer_masks with 0 / 0 iterations

The palette producers (`kmeans_palette` :31-46 -- cv2.kmeans on a RandomState(1) sample, centres truncated to u8 -- and
`palette_from_json` :49-66 with the two JSON forms, the first one being what analyze_colors.py:395-406 writes) stay on the host,
as SURVEY 8a row 9 prescribes.  There is no CPU fallback: without the library or a CUDA device the call raises.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import cv2
import numpy as np

from .ops import get_engine

VERSION = "v1.1.1"
KMEANS_SAMPLES = 200000


def load_image_rgb(path: str) -> np.ndarray:
    """process_colors.py:24-28."""
    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise ValueError(f"Cannot load image: {path}")
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)


def kmeans_palette(img_rgb: np.ndarray, k: int, samples: int = KMEANS_SAMPLES, seed: int = 1) -> np.ndarray:
    """process_colors.py:31-46 (host: cv2.kmeans is RNG-driven and cannot be reproduced bit for bit on a GPU)."""
    px = img_rgb.reshape(-1, 3)
    rs = np.random.RandomState(seed)
    if px.shape[0] > samples:
        px = px[rs.choice(px.shape[0], size=samples, replace=False)]
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 30, 1.0)
    _c, _l, centers = cv2.kmeans(px.astype(np.float32), K=k, bestLabels=None, criteria=crit, attempts=3, flags=cv2.KMEANS_PP_CENTERS)
    return centers.astype(np.uint8)


def palette_from_json(path: str):
    """process_colors.py:49-66: {"recommended_colors": [{position, name, rgb}]} (analyze_colors.py:395-406) sorted by position,
    or {"palette": [{rgb, name}]}."""
    with open(path, "r", encoding="utf-8") as fh:
        doc = json.load(fh)
    if "recommended_colors" in doc:
        entries = sorted(doc["recommended_colors"], key=lambda e: e.get("position", 1e9))
        names = [str(e.get("name", f"color_{i}")) for i, e in enumerate(entries)]
        return np.array([e["rgb"] for e in entries], dtype=np.uint8), names
    if "palette" in doc:
        entries = doc["palette"]
        names = [str(e.get("name", f"color_{i}")) for i, e in enumerate(entries)]
        return np.array([e["rgb"] for e in entries], dtype=np.uint8), names
    raise ValueError(f"Unsupported palette JSON structure: {path}")


def assign_labels(img_rgb: np.ndarray, palette_rgb: np.ndarray) -> np.ndarray:
    """process_colors.py:69-77."""
    return get_engine().host_assign_rgb_i16wrap(img_rgb, palette_rgb)


def default_color_names(k: int):
    """process_colors.py:80-82."""
    first = ("red", "green", "blue", "black")
    return [first[i] if i < len(first) else f"color_{i}" for i in range(k)]


def save_labels_png(path, labels: np.ndarray) -> None:
    """process_colors.py:85-86 writes an 8-bit "L" PNG through PIL; cv2 writes the same pixels."""
    if not cv2.imwrite(str(path), np.ascontiguousarray(labels, dtype=np.uint8)):
        raise RuntimeError(f"Failed to write labels: {path}")


def labels_and_layers(img_rgb: np.ndarray, palette_rgb: np.ndarray):
    """One upload: labels u8 [H, W], the K one-hot planes u8 [K, H, W] {0, 255} and the pixels per label."""
    import torch
    eng = get_engine()
    pal = np.ascontiguousarray(palette_rgb, dtype=np.uint8).reshape(-1, 3)
    d_img = torch.from_numpy(np.ascontiguousarray(img_rgb, dtype=np.uint8)).cuda()
    d_labels = eng.assign_rgb_i16wrap(d_img, pal)
    d_planes = eng.layer_masks(d_labels, pal.shape[0], open_iters=0, close_iters=0)
    counts = eng.count_nonzero(d_planes)
    return d_labels.cpu().numpy(), d_planes.cpu().numpy(), [int(v) for v in counts]


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="One-hot color layer generator with labels output")
    ap.add_argument("input", help="Input image")
    ap.add_argument("-o", "--output", default="layers", help="Output directory")
    ap.add_argument("-m", "--mode", choices=["adaptive", "palette"], default="adaptive", help="adaptive: KMeans; palette: load palette JSON")
    ap.add_argument("-n", "--colors", type=int, default=4, help="Number of colors for adaptive")
    ap.add_argument("--palette", help="Palette JSON (from analyze_colors.py) for mode=palette")
    ap.add_argument("--edges-only", action="store_true", help="Kept for pipeline compatibility (ignored)")
    return ap


def main(argv=None) -> None:
    """process_colors.py:89-176."""
    args = build_parser().parse_args(argv)
    out_dir = Path(args.output).absolute()
    out_dir.mkdir(parents=True, exist_ok=True)
    tag = "[process_colors]"
    print(f"{tag} {VERSION}")
    print(f"{tag} Input: {args.input}")
    print(f"{tag} Output dir: {out_dir}")
    img_rgb = load_image_rgb(args.input)
    h, w = img_rgb.shape[:2]
    print(f"{tag} Size: {w}x{h}")

    if args.mode == "palette":
        if not args.palette:
            raise ValueError("Mode 'palette' requires --palette JSON")
        palette_rgb, names = palette_from_json(args.palette)
        K = len(palette_rgb)
        if args.colors and args.colors != K:
            print(f"[WARN] --colors={args.colors} ignored; palette has {K} entries.")
    else:
        K = int(args.colors) if args.colors else 4
        palette_rgb = kmeans_palette(img_rgb, k=K)
        names = default_color_names(K)
    name_of = lambda i: names[i] if i < len(names) else f"color_{i}"          # noqa: E731

    labels, planes, counts = labels_and_layers(img_rgb, palette_rgb)

    total = labels.size
    print(f"{tag} Class distribution:")
    for i in range(K):
        share = 100.0 * counts[i] / total if total else 0.0
        rgb = tuple(int(v) for v in palette_rgb[i])
        print(f"  [{i}] {name_of(i):12s}  {rgb}  pixels={counts[i]:8d}  {share:5.1f}%")

    save_labels_png(out_dir / "labels.png", labels)
    np.save(str(out_dir / "labels.npy"), labels.astype(np.uint8))
    print(f"{tag} Saved labels PNG: {out_dir / 'labels.png'}")
    print(f"{tag} Saved labels NPY: {out_dir / 'labels.npy'}")

    with open(out_dir / "palette.json", "w", encoding="utf-8") as fh:
        json.dump({"colors": [{"index": i, "name": name_of(i), "rgb": [int(c) for c in palette_rgb[i].tolist()]} for i in range(K)]},
                  fh, indent=2)
    print(f"{tag} Saved palette JSON: {out_dir / 'palette.json'}")

    print(f"{tag} Saving one-hot masks...")
    for i in range(K):
        target = out_dir / f"layer_{i + 1}_{name_of(i)}.png"
        if not cv2.imwrite(str(target), planes[i]):
            raise RuntimeError(f"Failed to write mask: {target}")
    print(f"{tag} Done. {K} layer files written to: {out_dir}")
    if args.edges_only:
        print(f"{tag} NOTE: --edges-only is ignored here (kept for pipeline compatibility).")


if __name__ == "__main__":
    main()
