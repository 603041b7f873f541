import time, sys, os
t0=time.perf_counter()
import numpy as np, cv2
t1=time.perf_counter()
sys.path.insert(0, os.path.join(os.getcwd(), "omnirevolve-image-processor_b200"))
import omni_b200
from omni_b200 import capi, ops
t2=time.perf_counter()
L=capi.lib()
t3=time.perf_counter()
eng=ops.Engine(0)
t4=time.perf_counter()
img=np.random.default_rng(0).integers(0,256,(1024,1024,3),dtype=np.uint8)
ctr=np.array([[40,128,128],[110,140,120],[160,120,150],[220,128,128]],np.float32)
r=eng.host_color_edge_packed(img,ctr,np.arange(4,dtype=np.uint8),None)
t5=time.perf_counter()
r=eng.host_color_edge_packed(img,ctr,np.arange(4,dtype=np.uint8),None)
t6=time.perf_counter()
m=np.zeros((4,1024,1024),np.uint8); m[:,100:500,100:700]=255
e=eng.host_edges(m, ops.EdgeConfig())
t7=time.perf_counter()
print("import np+cv2 %.3f | import omni_b200 %.3f | dlopen %.3f | ctx_create %.3f | first packed call %.3f | second %.3f | host_edges first %.3f | torch loaded: %s"%(t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5,t7-t6,'torch' in sys.modules))
t=time.perf_counter(); import torch; print("import torch %.3f"%(time.perf_counter()-t))
