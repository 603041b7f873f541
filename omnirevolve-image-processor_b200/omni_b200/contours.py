"""Stage 04, first step: the reference's `thinning_zhangsuen` (04_find_contours.py:35-99) on the GPU.

Only the data-parallel part of stage 04 lives here: the skeleton the reference computes with
NumPy shifts in ~10^2 s per layer comes out of the bit-plane kernel `omni_thin_zhangsuen` in well under a
millisecond.  `trace_centerlines` (04:101-205) walks the skeleton sequentially and stays the reference's own code;
it consumes the array returned here unchanged.

    thinning_zhangsuen(bin_0_255, layer)      same signature, same result, same progress lines (04:37-95)
    thin_layers(planes)                        all K edge planes in one call (what a fused stage 03 -> 04 hand-off uses)
"""
from __future__ import annotations

import time

import numpy as np

from .ops import get_engine

MAX_ITER = 120                      # 04_find_contours.py:52


def thin_layers(planes: np.ndarray, max_iter: int = MAX_ITER):
    """[K,H,W] u8 (> 0 = edge) -> (skeletons {0,255}, removed[K,max_iter], iters[K])."""
    return get_engine().host_thin_zhangsuen(planes, max_iter)


def thinning_zhangsuen(bin_0_255: np.ndarray, layer: str) -> np.ndarray:
    """Drop-in for 04_find_contours.py:35-99 (same prints; the per-iteration seconds are the kernel's share)."""
    if bin_0_255.ndim != 2:
        raise ValueError("thinning_zhangsuen expects a single-channel image")
    ys, xs = np.nonzero(bin_0_255)
    if len(xs) == 0:
        print(f"[{layer}] Thinning: empty edges.", flush=True)
        return np.zeros_like(bin_0_255)
    pad = 2                                                       # _bbox_of_nonzero, 04:24-33 (only for the log line)
    x0, x1 = max(0, xs.min() - pad), min(bin_0_255.shape[1] - 1, xs.max() + pad)
    y0, y1 = max(0, ys.min() - pad), min(bin_0_255.shape[0] - 1, ys.max() + pad)
    total0 = int(len(xs))
    print(f"[{layer}] Thinning ROI {x1 - x0 + 1}x{y1 - y0 + 1}, fg={total0} px", flush=True)
    t0 = time.perf_counter()
    src = np.ascontiguousarray(bin_0_255 if bin_0_255.dtype == np.uint8 else (bin_0_255 > 0).astype(np.uint8) * 255)
    out, removed, iters = get_engine().host_thin_zhangsuen(src[None], MAX_ITER)
    dt = time.perf_counter() - t0
    n_it = int(iters[0])
    total_removed = 0
    for it in range(1, n_it + 1):
        r = int(removed[0, it - 1])
        total_removed += r
        pct = total_removed / max(1, total0) * 100.0
        print(f"[{layer}] Thin {it:02d}: removed={r:6d} | total={total_removed:7d} ({pct:5.1f}%) | {dt / max(1, n_it):.2f}s", flush=True)
    print(f"[{layer}] Thinning done in {dt:.2f}s, fg_now={total0 - total_removed} px", flush=True)
    return out[0].astype(bin_0_255.dtype, copy=False)


def skeleton_degree(skel_0_255: np.ndarray):
    """04_find_contours.py:117-125 for ALL components at once: (deg, endpoints, junctions) of the skeleton -- what
    trace_centerlines recomputes per component with a full-image filter2D.  On the pixels of a component the maps equal the
    reference's per-component ones (a pixel's 8-neighbours belong to its own component), so inside the component loop
    `deg`, `endpoints`, `junctions` can be replaced by `deg_all`, `ep_all & (comp_mask == 1)`, `jn_all & (comp_mask == 1)`."""
    import torch
    sk = torch.from_numpy(np.ascontiguousarray(skel_0_255, dtype=np.uint8)[None]).cuda()
    deg, nodes = get_engine().skeleton_degree(sk)
    nodes = nodes[0].cpu().numpy()
    return deg[0].cpu().numpy(), nodes == 1, nodes == 2
