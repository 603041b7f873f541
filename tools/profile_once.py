"""Three fused colour+edge calls on a device-resident image (the ncu target: capture the kernels of the LAST call).
    python tools/profile_once.py [config2|config3|config4]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200          # noqa: E402
from omni_b200.synth import synth              # noqa: E402
from omni_b200 import stages                   # noqa: E402

cfg = {"config5": (4096, 4096, 16, 0, 32), "config2": (4096, 4096, 8, 0, 32), "config3": (8192, 8192, 16, 1, 64), "config4": (1080, 1920, 8, 0, 32)}[sys.argv[1] if len(sys.argv) > 1 else "config2"]
h, w, K, seed, cell = cfg
eng = omni_b200.Engine(0)
eng.set_table_cache(False)              # a single image brings its own centres: the tables are rebuilt in every call
img = synth(h, w, seed, cell)
ctr = stages.kmeans_lab_centers(img, K)
_o, lut = stages.darkness_lut(ctr)
ec = omni_b200.EdgeConfig()
d = torch.from_numpy(img).cuda()
m = torch.empty((K, h, w), dtype=torch.uint8, device="cuda"); e = torch.empty_like(m)
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.fill_(1)
    eng.color_edge(d, ctr, lut.astype(np.uint8), ec, masks=m, edges=e)
torch.cuda.synchronize()
print("edge nz", int((e > 0).sum()), "mask nz", int((m > 0).sum()))
