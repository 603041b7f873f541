"""Timing probe for the single-image host call (omni_host_color_edge_packed): bands off / mode 1 / mode 2, beside the raw copy
times of the same buffers (H2D alone, D2H alone, both at once).  Run on a B200:  python tools/e2e_probe.py [K] [h] [w]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "omnirevolve-image-processor_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import omni_b200  # noqa: E402
from helpers import synth  # noqa: E402
from oracle import refport as rp  # noqa: E402  (centres only)

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
h = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
w = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
eng = omni_b200.Engine(0)
img = synth(h, w, 0)
ctr = rp.kmeans_lab_centers(img[:1024], K)
_o, lut = rp.darkness_order(ctr)
lut = lut.astype(np.uint8)
ec = omni_b200.EdgeConfig()
rb = (w + 7) // 8
h_img = omni_b200.pinned_empty((1, h, w, 3)); h_img[0] = img
h_mb = omni_b200.pinned_empty((K, h, rb)); h_eb = omni_b200.pinned_empty((K, h, rb))
eng.set_table_cache(False)


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(t)), 1e3 * float(np.min(t))


ref = None
for mode in (0, 1, 2):
    eng.set_host_bands(mode)
    fn = lambda: eng.host_color_edge_packed(h_img, ctr, lut, ec, mask_bits=h_mb, edge_bits=h_eb, want_counts=False)
    med, mn = timeit(fn)
    cur = (h_mb.copy(), h_eb.copy())
    same = True if ref is None else (np.array_equal(cur[0], ref[0]) and np.array_equal(cur[1], ref[1]))
    ref = ref or cur
    print(f"bands={mode}: median {med:.3f} ms  min {mn:.3f} ms  -> {h * w / med / 1e3:.0f} MP/s  resends={eng.last_band_resends()}  same_bytes={same}")
eng.set_host_bands(2)
eng.profile(True)
eng.host_color_edge_packed(h_img, ctr, lut, ec, mask_bits=h_mb, edge_bits=h_eb, want_counts=False)
torch.cuda.synchronize()
ps = eng.profile_summary()
eng.profile(False)
print("kernels of one banded call (launches, total ms):", {k: (n, round(ms, 4)) for k, (n, ms) in ps.items()}, "sum", round(sum(ms for _n, ms in ps.values()), 4))

# raw copies of the same sizes
d_in = torch.empty(h * w * 3, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * K * h * rb, dtype=torch.uint8, device="cuda")
t_in = torch.from_numpy(h_img.reshape(-1)); t_mb = torch.from_numpy(h_mb.reshape(-1)); t_eb = torch.from_numpy(h_eb.reshape(-1))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(t_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        t_mb.copy_(d_out[: t_mb.numel()], non_blocking=True); t_eb.copy_(d_out[t_mb.numel():], non_blocking=True)


def both():
    h2d(); d2h()


for name, fn, nbytes in (("H2D image", h2d, t_in.numel()), ("D2H packed planes", d2h, 2 * t_mb.numel()), ("both at once", both, t_in.numel() + 2 * t_mb.numel())):
    med, mn = timeit(fn)
    print(f"{name}: median {med:.3f} ms ({nbytes / med / 1e6:.1f} GB/s)")
