mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "omnirevolve-image-processor_b200")); sys.path.insert(0, "tests")
import numpy as np, cv2, torch, omni_b200
from omni_b200.synth import synth
eng = omni_b200.Engine(0)
for (h, w, K) in ((4096, 4096, 8), (4096, 4096, 16), (1080, 1920, 8)):
    img = synth(h, w, 0)
    n = h * w
    idx = np.random.default_rng(42).choice(n, size=200000, replace=False)
    lab = cv2.cvtColor(np.ascontiguousarray(img.reshape(-1, 3)[idx]).reshape(-1, 1, 3), cv2.COLOR_BGR2LAB).reshape(-1, 3).astype(np.float32)
    t0 = time.perf_counter(); cv2.setRNGSeed(0)
    comp_cv, _l, c_cv = cv2.kmeans(lab, K, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.5), 3, cv2.KMEANS_PP_CENTERS)
    t_cv = time.perf_counter() - t0
    d = torch.from_numpy(img).cuda()
    eng.kmeans_lab(d, K, idx)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctr, comp = eng.kmeans_lab(d, K, idx)
    t_gpu = time.perf_counter() - t0
    print(f"{h}x{w} K={K}: cv2 {t_cv*1e3:.1f} ms comp {comp_cv:.4e} | gpu {t_gpu*1e3:.2f} ms comp {comp:.4e} ratio {comp/comp_cv:.4f}")
PY
