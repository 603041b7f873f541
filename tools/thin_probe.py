"""Timing probe of the thinning call on the edge planes of 4096^2, K=16 (byte planes and packed planes), library from OMNI_B200_LIB."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200          # noqa: E402
from omni_b200.synth import synth              # noqa: E402
from omni_b200 import stages                   # noqa: E402
h = w = 4096; K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = omni_b200.Engine(0)
img = synth(h, w, 0, 32)
ctr = stages.kmeans_lab_centers(img, K)
_o, lut = stages.darkness_lut(ctr)
_l, m, e = eng.color_edge(torch.from_numpy(img).cuda(), ctr, lut.astype(np.uint8), omni_b200.EdgeConfig())
sk = torch.empty_like(e)
eng.profile(True)
for _ in range(3):
    eng.thin_zhangsuen(e, out=sk)
torch.cuda.synchronize(); eng.profile_summary()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    eng.thin_zhangsuen(e, out=sk)
b.record(); torch.cuda.synchronize()
ps = eng.profile_summary()
print(os.path.basename(omni_b200.LIB_PATH), "K", K, "ms/call", round(a.elapsed_time(b) / 10, 4), {k: round(v[1] / v[0] * 1e3, 1) for k, v in ps.items()},
      "csum", int(sk.view(-1).view(torch.int64).sum().item()) % 1000003)
