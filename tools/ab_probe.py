"""A/B probe: one workload through omni_color_edge with the library named by OMNI_B200_LIB; prints one JSON line with the step
time (CUDA events, L2 flushed between steps, tables rebuilt every step), the per-kernel times and a checksum of the outputs.
    OMNI_B200_LIB=lib/variant.so python tools/ab_probe.py [config5|config2|config3|config4] [steps]"""
import json
import os
import statistics
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200          # noqa: E402
from omni_b200.synth import synth              # noqa: E402
from omni_b200 import stages                   # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "config5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h, w, K, seed, cell = {"config5": (4096, 4096, 16, 0, 32), "config2": (4096, 4096, 8, 0, 32), "config3": (8192, 8192, 16, 1, 64),
                       "config4": (1080, 1920, 8, 0, 32), "small": (1024, 1024, 4, 0, 32)}[wl]
eng = omni_b200.Engine(0)
img = synth(h, w, seed, cell)
ctr = stages.kmeans_lab_centers(img, K)
_o, lut = stages.darkness_lut(ctr)
lut = lut.astype(np.uint8)
ec = omni_b200.EdgeConfig()
d = torch.from_numpy(img).cuda()
m = torch.empty((K, h, w), dtype=torch.uint8, device="cuda"); e = torch.empty_like(m)
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
eng.set_table_cache(False)
for _ in range(5):
    eng.color_edge(d, ctr, lut, ec, masks=m, edges=e)
torch.cuda.synchronize()


def run(prof):
    eng.profile(prof)
    ts = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.color_edge(d, ctr, lut, ec, masks=m, edges=e); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    p = eng.profile_summary() if prof else None
    eng.profile(False)
    return ts, p


ts, _ = run(False)
_, prof = run(True)


def csum(t):
    v = t.view(-1).view(torch.int64)
    i = torch.arange(v.numel(), device="cuda", dtype=torch.int64) % 1000003 + 1
    return int((v * i).sum().item())


print(json.dumps({"lib": os.path.basename(omni_b200.LIB_PATH), "workload": wl, "ms_median": statistics.median(ts), "ms_min": min(ts),
                  "kernels_us": {k: round(1e3 * v[1] / steps, 1) for k, v in prof.items()},
                  "mask_csum": csum(m), "edge_csum": csum(e)}))
