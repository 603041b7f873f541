"""Three banded host calls (omni_host_color_edge_packed, one 4096^2 image, K from argv): the ncu launch-list target for the per-band
kernel durations.    python tools/profile_banded.py [K] [bands]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200          # noqa: E402
from omni_b200.synth import synth              # noqa: E402
from omni_b200 import stages                   # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = w = 4096
eng = omni_b200.Engine(0)
eng.set_table_cache(False)
eng.set_host_bands(bands)
img = synth(h, w, 0, 32)
ctr = stages.kmeans_lab_centers(img, K)
_o, lut = stages.darkness_lut(ctr)
rb = w // 8
h_img = omni_b200.pinned_empty((1, h, w, 3)); h_img[0] = img
h_mb = omni_b200.pinned_empty((K, h, rb)); h_eb = omni_b200.pinned_empty((K, h, rb))
for _ in range(3):
    eng.host_color_edge_packed(h_img, ctr, lut.astype(np.uint8), omni_b200.EdgeConfig(), mask_bits=h_mb, edge_bits=h_eb, want_counts=False)
torch.cuda.synchronize()
print("resends", eng.last_band_resends(), "edge bytes nz", int((h_eb != 0).sum()))
