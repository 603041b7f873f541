import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'omnirevolve-image-processor_b200'))
import numpy as np, torch, omni_b200
from omni_b200.synth import synth
from omni_b200 import stages
eng=omni_b200.Engine(0)
img=synth(4096,4096,0); K=8
ctr=stages.kmeans_lab_centers(img,K); _o,lut=stages.darkness_lut(ctr); lut=lut.astype(np.uint8)
ec=omni_b200.EdgeConfig()
d=torch.from_numpy(img).cuda(); m=torch.empty((K,4096,4096),dtype=torch.uint8,device='cuda'); e=torch.empty_like(m)
flush=torch.empty(384<<20,dtype=torch.uint8,device='cuda')
def run(n, prof):
    eng.profile(prof)
    ts=[]
    for _ in range(n):
        flush.fill_(1)
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); eng.color_edge(d,ctr,lut,ec,masks=m,edges=e); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    if prof: eng.profile_summary()
    eng.profile(False)
    return np.median(ts), min(ts)
for _ in range(5): eng.color_edge(d,ctr,lut,ec,masks=m,edges=e)
print("prof off", run(20,False)); print("prof on", run(20,True)); print("prof off", run(20,False))
