import sys, os
R="/root/repo"
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,"omnirevolve-image-processor_b200")); sys.path.insert(0,os.path.join(R,"tests"))
import numpy as np, torch, omni_b200
from helpers import synth
from oracle import cmodel as cm
eng=omni_b200.Engine(0)
img=synth(4096,4096,0)
for K in (4,16):
    pal=np.random.default_rng(K).integers(0,256,(K,3),dtype=np.uint8)
    d=torch.from_numpy(img).cuda()
    out=eng.assign_rgb_i16wrap(d,pal)
    want=cm.assign_i16wrap(img[:512],pal)
    print(K,"equal:",bool(np.array_equal(out[:512].cpu().numpy(),want)))
    for _ in range(3): eng.assign_rgb_i16wrap(d,pal)
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): eng.assign_rgb_i16wrap(d,pal)
    b.record(); torch.cuda.synchronize(); print(K,"ms",a.elapsed_time(b)/20)
