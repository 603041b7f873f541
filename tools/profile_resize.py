"""ncu target: fractional INTER_AREA 4096^2 -> 2000^2 (the default max_dimension path), three calls."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import torch, omni_b200          # noqa: E402
eng = omni_b200.Engine(0)
src = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((2000, 2000, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.resize_area(src, 2000, 2000, out=dst)
torch.cuda.synchronize()
print(int(dst.sum()))
