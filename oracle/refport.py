"""Replay of the reference's stage 01-03 arithmetic through the SAME library calls it makes
(cv2 + NumPy) -- TEST INFRASTRUCTURE ONLY, and the timed "reference CPU path" of bench.py.

The reference (/root/reference/image_processor/*.py) is Python that cannot travel to the GPU
box, so its call sites are restated here with the PNG/file I/O stripped.  tests/test_refport.py
(run in the build container, where /root/reference exists) imports the real stage modules and
checks these functions against them; tools/make_golden.py freezes their outputs as fixtures.

Pinned library versions (the reference pins none): opencv-python-headless 4.13.0.92, NumPy 2.3.5.
"""
from __future__ import annotations

import numpy as np
import cv2


# ---- 01_resize.py:7-23 -------------------------------------------------------------------------
def resize_dims(h: int, w: int, max_dimension: int):
    """Target (new_w, new_h) or None; identical Python float expressions as 01_resize.py:15-18."""
    max_dim = max(h, w)
    if max_dim > max_dimension:
        scale = max_dimension / max_dim
        return int(w * scale), int(h * scale)
    return None


def resize_if_needed(img: np.ndarray, max_dimension: int) -> np.ndarray:
    dims = resize_dims(img.shape[0], img.shape[1], max_dimension)
    if dims is None:
        return img
    return cv2.resize(img, dims, interpolation=cv2.INTER_AREA)


# ---- 02_color_extract.py:32-56 -----------------------------------------------------------------
def kmeans_lab_centers(img_bgr: np.ndarray, k: int, sample_limit: int = 200_000, attempts: int = 3,
                       fresh_rng: bool = True) -> np.ndarray:
    """Centres as a fresh reference process computes them (02:35-50).  cv2.kmeans draws from the
    process-global RNG; cv2.setRNGSeed(0) restores the fresh-process state."""
    lab = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB)
    data = lab.reshape(-1, 3).astype(np.float32)
    n = data.shape[0]
    if n > sample_limit:
        pick = np.random.default_rng(42).choice(n, size=sample_limit, replace=False)
        sample = data[pick]
    else:
        sample = data
    if fresh_rng:
        cv2.setRNGSeed(0)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.5)
    _c, _l, centers = cv2.kmeans(sample, k, None, crit, attempts, cv2.KMEANS_PP_CENTERS)
    return centers.astype(np.float32)


def assign_lab(img_bgr: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """02:35-36,53-55: BGR->Lab (8-bit), float32 broadcast distance, argmin.  int32 [H,W]."""
    h, w = img_bgr.shape[:2]
    lab = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB)
    data = lab.reshape(-1, 3).astype(np.float32)
    diffs = data[:, None, :] - centers[None, :, :]
    d2 = np.sum(diffs * diffs, axis=2)
    return np.argmin(d2, axis=1).astype(np.int32).reshape(h, w)


def assign_lab_chunked(img_bgr: np.ndarray, centers: np.ndarray, rows: int = 256) -> np.ndarray:
    """Same values as assign_lab, evaluated in row bands so 8192^2 x K=16 fits in RAM."""
    out = np.empty(img_bgr.shape[:2], np.int32)
    for y in range(0, img_bgr.shape[0], rows):
        out[y:y + rows] = assign_lab(np.ascontiguousarray(img_bgr[y:y + rows]), centers)
    return out


def darkness_order(centers: np.ndarray):
    """02:121-127: order = argsort(L); lut[order] = arange.  Returns (order, lut)."""
    order = np.argsort(centers[:, 0])
    lut = np.zeros_like(order)
    lut[order] = np.arange(len(order))
    return order, lut


def darkness_rank(name: str) -> int:
    """02:17-23."""
    s = name.lower()
    for key, rank in (("dark", 0), ("mid", 1), ("skin", 2), ("light", 3)):
        if key in s:
            return rank
    return 2


def layer_masks(labels: np.ndarray, K: int, open_iters: int = 1, close_iters: int = 1) -> np.ndarray:
    """02:136-154: per cluster (labels==k)*255, RECT-3 OPEN then CLOSE.  u8 [K,H,W]."""
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
    out = np.empty((K,) + labels.shape, np.uint8)
    for k in range(K):
        m = (labels == k).astype(np.uint8) * 255
        if open_iters > 0:
            m = cv2.morphologyEx(m, cv2.MORPH_OPEN, se, iterations=open_iters)
        if close_iters > 0:
            m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, se, iterations=close_iters)
        out[k] = m
    return out


def color_extract(img_bgr: np.ndarray, K: int, centers: np.ndarray | None = None):
    """02:111-154 without files: (centers_sorted, labels_sorted, masks[K,H,W])."""
    if centers is None:
        centers = kmeans_lab_centers(img_bgr, K)
    labels = assign_lab(img_bgr, centers)
    order, lut = darkness_order(centers)
    return centers[order], lut[labels], layer_masks(lut[labels], K)


def swatch_masks(img_bgr: np.ndarray, colors, tol: int = 30) -> np.ndarray:
    """02:82-109 (swatch mode) without files: per swatch the better of inRange(RGB->BGR) / inRange(as-is), RECT-3 open/close."""
    se = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
    out = []
    for rgb in colors:
        rgb = tuple(int(v) for v in rgb)
        cands = []
        for c in ((rgb[2], rgb[1], rgb[0]), rgb):
            lower = np.array([max(0, c[0] - tol), max(0, c[1] - tol), max(0, c[2] - tol)], np.uint8)
            upper = np.array([min(255, c[0] + tol), min(255, c[1] + tol), min(255, c[2] + tol)], np.uint8)
            cands.append(cv2.inRange(img_bgr, lower, upper))
        m = cands[0] if int(np.count_nonzero(cands[0])) >= int(np.count_nonzero(cands[1])) else cands[1]
        m = cv2.morphologyEx(m, cv2.MORPH_OPEN, se, iterations=1)
        m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, se, iterations=1)
        out.append(m)
    return np.stack(out)


# ---- 03_edge_detect.py:9-34 --------------------------------------------------------------------
def ensure_odd(n) -> int:
    n = max(3, int(n))
    return n if n % 2 == 1 else n + 1


def edge_layer(mask: np.ndarray, low=50, high=150, ksize=3, morph_k=3, open_iters=1, close_iters=1) -> np.ndarray:
    """03:23-34: ELLIPSE open/close, GaussianBlur(sigma 0), Canny."""
    k_m = max(1, int(morph_k))
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k_m, k_m))
    if open_iters > 0:
        mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, se, iterations=int(open_iters))
    if close_iters > 0:
        mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, se, iterations=int(close_iters))
    k = ensure_odd(ksize)
    blurred = cv2.GaussianBlur(mask, (k, k), 0)
    return cv2.Canny(blurred, low, high)


def edges_all(masks: np.ndarray, **kw) -> np.ndarray:
    return np.stack([edge_layer(m, **kw) for m in masks])


def edges_composite(edges: np.ndarray, colors) -> np.ndarray:
    """03:60-111 paint: white canvas, later layers overwrite earlier ones."""
    canvas = np.full(edges.shape[1:] + (3,), 255, np.uint8)
    for e, col in zip(edges, colors):
        canvas[e > 0] = tuple(int(v) for v in col)
    return canvas


# ---- process_colors.py:69-77 -------------------------------------------------------------------
def assign_labels_rgb(img_rgb: np.ndarray, palette_rgb: np.ndarray) -> np.ndarray:
    """int16 broadcast (diff*diff wraps), wide sum, argmin -> u8 [H,W]."""
    h, w, _ = img_rgb.shape
    flat = img_rgb.reshape(-1, 3).astype(np.int16)
    pal = palette_rgb.astype(np.int16)
    diff = flat[:, None, :] - pal[None, :, :]
    dist2 = np.sum(diff * diff, axis=2)
    return np.argmin(dist2, axis=1).astype(np.uint8).reshape(h, w)


def assign_labels_rgb_chunked(img_rgb, palette_rgb, rows: int = 256):
    out = np.empty(img_rgb.shape[:2], np.uint8)
    for y in range(0, img_rgb.shape[0], rows):
        out[y:y + rows] = assign_labels_rgb(np.ascontiguousarray(img_rgb[y:y + rows]), palette_rgb)
    return out


# ---- 04_find_contours.py:14-22,35-99 (row "next" of SURVEY 8f) ------------------------------------
def _shift(img, dy, dx):
    h, w = img.shape
    out = np.zeros_like(img)
    out[max(0, dy):min(h, h + dy), max(0, dx):min(w, w + dx)] = img[max(0, -dy):min(h, h - dy), max(0, -dx):min(w, w - dx)]
    return out


def thinning_zhangsuen(bin_0_255: np.ndarray, max_iter: int = 120) -> np.ndarray:
    """The reference's NumPy Zhang-Suen (same shifts, same conditions, same stop rule), without the bounding-box
    crop (zero fill outside makes it equivalent) and without the progress prints."""
    roi = (bin_0_255 > 0).astype(np.uint8)
    changed, it = True, 0
    while changed and it < max_iter:
        it += 1
        changed = False
        for step in (1, 2):
            P2 = _shift(roi, -1, 0); P3 = _shift(roi, -1, 1); P4 = _shift(roi, 0, 1); P5 = _shift(roi, 1, 1)
            P6 = _shift(roi, 1, 0); P7 = _shift(roi, 1, -1); P8 = _shift(roi, 0, -1); P9 = _shift(roi, -1, -1)
            B = P2 + P3 + P4 + P5 + P6 + P7 + P8 + P9
            seq = [P2, P3, P4, P5, P6, P7, P8, P9, P2]
            A = sum(((a == 0) & (b == 1)).astype(np.uint8) for a, b in zip(seq[:-1], seq[1:]))
            if step == 1:
                cond = ((P2 * P4 * P6) == 0) & ((P4 * P6 * P8) == 0)
            else:
                cond = ((P2 * P4 * P8) == 0) & ((P2 * P6 * P8) == 0)
            dele = (roi == 1) & (A == 1) & (B >= 2) & (B <= 6) & cond
            if dele.any():
                roi[dele] = 0
                changed = True
    return (roi * 255).astype(np.uint8)


def skeleton_degree(skel_0_255: np.ndarray):
    """04_find_contours.py:117-125 for the whole skeleton at once: (deg, endpoints, junctions) with the reference's own
    calls (filter2D with the 3x3 ring kernel, BORDER_CONSTANT)."""
    S = (skel_0_255 > 0).astype(np.uint8)
    kernel = np.ones((3, 3), np.uint8)
    kernel[1, 1] = 0
    deg = cv2.filter2D(S, cv2.CV_8U, kernel, borderType=cv2.BORDER_CONSTANT)
    return deg, (S == 1) & (deg == 1), (S == 1) & (deg >= 3)


def skeleton_degree_per_component(skel_0_255: np.ndarray):
    """The reference's actual loop (04:113-125): per connected component, deg / endpoints / junctions on the component mask;
    returns the union over components (what the global map must reproduce on skeleton pixels)."""
    S = (skel_0_255 > 0).astype(np.uint8)
    num, labels = cv2.connectedComponents(S, connectivity=8)
    kernel = np.ones((3, 3), np.uint8)
    kernel[1, 1] = 0
    deg_on = np.zeros(S.shape, np.uint8)
    endpoints = np.zeros(S.shape, bool)
    junctions = np.zeros(S.shape, bool)
    for comp_id in range(1, num):
        comp_mask = (labels == comp_id).astype(np.uint8)
        deg = cv2.filter2D(comp_mask, cv2.CV_8U, kernel, borderType=cv2.BORDER_CONSTANT)
        deg_on[comp_mask == 1] = deg[comp_mask == 1]
        endpoints |= (comp_mask == 1) & (deg == 1)
        junctions |= (comp_mask == 1) & (deg >= 3)
    return deg_on, endpoints, junctions
