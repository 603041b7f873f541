"""ncu target: thinning of the K edge planes of config 2 (two calls)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200          # noqa: E402
from omni_b200.synth import synth              # noqa: E402
from omni_b200 import stages                   # noqa: E402
h = w = 4096; K = 8
eng = omni_b200.Engine(0)
img = synth(h, w, 0, 32)
ctr = stages.kmeans_lab_centers(img, K)
_o, lut = stages.darkness_lut(ctr)
_l, m, e = eng.color_edge(torch.from_numpy(img).cuda(), ctr, lut.astype(np.uint8), omni_b200.EdgeConfig())
sk = torch.empty_like(e)
for _ in range(2):
    out, removed, iters = eng.thin_zhangsuen(e, out=sk, with_log=True)
torch.cuda.synchronize()
print("iters", iters, "removed", removed.sum(axis=1))
