"""ncu target: omni_edges with edge_kernel_size 5 and 7 on the 4096^2, K=16 masks (capture fk_blur_bits / fk_edges3_simd)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200
from omni_b200.synth import synth
from omni_b200 import stages
eng = omni_b200.Engine(0)
img = synth(4096, 4096, 0); K = 16
ctr = stages.kmeans_lab_centers(img, K); _o, lut = stages.darkness_lut(ctr); lut = lut.astype(np.uint8)
_l, masks_d, _e = eng.color_edge(torch.from_numpy(img).cuda(), ctr, lut, omni_b200.EdgeConfig())
for ks in (5, 7, 5, 7):
    e = eng.edges(masks_d, omni_b200.EdgeConfig(ksize=ks))
torch.cuda.synchronize()
print("nz", int((e > 0).sum()))
