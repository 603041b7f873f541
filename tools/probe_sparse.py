"""Debug probe: the sparse generation of the fused call (fast path 1) against the dense generation (3) and the C oracle.
    python tools/probe_sparse.py [small|big]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200")); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, omni_b200
from omni_b200.synth import synth
from omni_b200 import stages
from oracle import cmodel as cm

eng = omni_b200.Engine(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "small"
cases = [(64, 96, 4, 0), (200, 333, 4, 1), (257, 1030, 8, 2), (512, 768, 16, 3), (33, 40, 2, 4), (1080, 1920, 8, 5), (8, 32, 3, 6), (7, 31, 5, 7),
         (100, 2049, 7, 8)] if mode == "small" else ([(4096, 4096, 16, 0), (4096, 4096, 8, 0)] if mode == "time" else [(4096, 4096, 16, 0), (4096, 4096, 8, 0), (2160, 3840, 12, 1)])
bad = 0
for (h, w, K, seed) in cases:
    img = synth(h, w, seed)
    ctr = stages.kmeans_lab_centers(img, K)
    _o, lut = stages.darkness_lut(ctr)
    lut = lut.astype(np.uint8)
    d = torch.from_numpy(img).cuda()
    for ec in (omni_b200.EdgeConfig(), omni_b200.EdgeConfig(open_iters=0), omni_b200.EdgeConfig(close_iters=0), omni_b200.EdgeConfig(morph_k=1), omni_b200.EdgeConfig(low=100, high=200)):
        if mode != "small" and ec != omni_b200.EdgeConfig():
            continue
        if mode == "time":
            continue
        out = {}
        for m in (1, 3):
            eng.set_fast_path(m)
            lab, mk, ed = eng.color_edge(d, ctr, lut, ec, want_labels=True)
            torch.cuda.synchronize()
            out[m] = (lab.cpu().numpy(), mk.cpu().numpy(), ed.cpu().numpy())
        tag = f"{h}x{w} K={K} ec={ec}"
        ok = True
        for name, a, b in zip(("labels", "masks", "edges"), out[1], out[3]):
            if not np.array_equal(a, b):
                ok = False
                diff = np.argwhere(a != b)
                print("MISMATCH", tag, name, "count", len(diff), "first", diff[:5].tolist(), "planes", np.unique(diff[:, 0])[:10] if diff.shape[1] == 3 else "")
        if mode == "small" and ec == omni_b200.EdgeConfig():
            raw = cm.assign_f32(cm.bgr2lab(img), ctr)
            wm = cm.layer_masks(raw, K, lut)
            we = np.stack([cm.edge_chain(mm, 3, 1, 1, 3, 50, 150) for mm in wm])
            for name, a, b in zip(("labels", "masks", "edges"), out[1], (lut[raw], wm, we)):
                if not np.array_equal(a, b):
                    ok = False
                    print("ORACLE MISMATCH", tag, name, int((a != b).sum()))
        bad += (not ok)
        print("ok " if ok else "BAD", tag, "passes", eng.last_hysteresis_passes())
# batch
if mode == "small":
    h, w, K = 270, 480, 8
    frames = np.stack([synth(h, w, 10 + i) for i in range(4)])
    ctr = stages.kmeans_lab_centers(frames[0], K); _o, lut = stages.darkness_lut(ctr); lut = lut.astype(np.uint8)
    d = torch.from_numpy(frames).cuda()
    res = {}
    for m in (1, 3):
        eng.set_fast_path(m)
        mk, ed = eng.color_edge_batch(d, ctr, lut, omni_b200.EdgeConfig())
        torch.cuda.synchronize()
        res[m] = (mk.cpu().numpy(), ed.cpu().numpy())
    okb = all(np.array_equal(a, b) for a, b in zip(res[1], res[3]))
    print("batch", "ok" if okb else "BAD")
    bad += (not okb)
else:
    eng.set_fast_path(1)
    for (h, w, K, seed) in cases:
        img = synth(h, w, seed); ctr = stages.kmeans_lab_centers(img, K); _o, lut = stages.darkness_lut(ctr); lut = lut.astype(np.uint8)
        d = torch.from_numpy(img).cuda(); m = torch.empty((K, h, w), dtype=torch.uint8, device="cuda"); e = torch.empty_like(m)
        flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
        ec = omni_b200.EdgeConfig()
        for mode_ in ((1,) if os.environ.get("PROBE_ONLY1") else (1, 3)):
            eng.set_fast_path(mode_)
            for _ in range(3): eng.color_edge(d, ctr, lut, ec, masks=m, edges=e)
            ts = []
            for _ in range(10):
                flush.fill_(1); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); eng.color_edge(d, ctr, lut, ec, masks=m, edges=e); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            eng.profile(True)
            for _ in range(5):
                flush.fill_(1); eng.color_edge(d, ctr, lut, ec, masks=m, edges=e)
            torch.cuda.synchronize()
            pr = eng.profile_summary(); eng.profile(False)
            print(f"{h}x{w} K={K} mode {mode_}: median {np.median(ts):.4f} ms min {min(ts):.4f}", {k: round(v[1] / 5, 4) for k, v in pr.items()})
print("FAILED" if bad else "ALL OK")
