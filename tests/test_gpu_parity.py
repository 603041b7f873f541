"""Parity of the CUDA path (through the C ABI, omni_b200.Engine -> libomni_b200.so) against the CPU
oracle (oracle.cmodel = plain-C restatement, oracle.refport = cv2/NumPy replay of the reference's call
sites) and against the golden vectors frozen from the unmodified reference.  Bit-exact for labels, masks
and edges; +-1 LSB for fractional resize (exact for integer ratios).  Needs a B200: `-m gpu`."""
import numpy as np
import pytest

from helpers import PIPE_CASES, load_pipe_case, plane_of_name, GOLDEN, synth, uniform_img, smooth_u8, blob_mask

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    import omni_b200
    e = omni_b200.Engine(0)         # raises (no CPU fallback) when the .so or the GPU is missing
    yield e
    e.close()


@pytest.fixture(scope="module", params=[1, 3, 2, 0], ids=["fast", "fast_dense_pipeline", "fast_dense_edges", "generic"])
def eng_mode(request, eng):
    eng.set_fast_path(request.param)
    yield eng
    eng.set_fast_path(True)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def _cm():
    from oracle import cmodel
    return cmodel


def _rp():
    from oracle import refport
    return refport


# ---- stage 01 --------------------------------------------------------------------------------------
@pytest.mark.parametrize("hs,ws,md", [(96, 128, 64), (512, 512, 256), (768, 384, 256), (96, 144, 48), (300, 200, 133),
                                      (409, 409, 200), (1000, 1500, 700), (1080, 1920, 1000), (33, 1000, 77),
                                      (2048, 2048, 1024), (2050, 1026, 1025), (1200, 1600, 400)])
def test_resize_area(eng_mode, hs, ws, md):
    cm, rp = _cm(), _rp()
    src = synth(hs, ws, hs + ws, cell=16)
    nw, nh = rp.resize_dims(hs, ws, md)
    want = cm.resize_area(src, nw, nh)
    got = host(eng_mode.resize_area(dev(src), nw, nh))
    integer_ratio = hs % nh == 0 and ws % nw == 0
    if integer_ratio:
        assert np.array_equal(got, want)
    else:
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        assert d.max() <= 1                       # contract: +-1 LSB (north_star); in practice identical
        assert (d != 0).mean() < 1e-3


def test_resize_golden(eng_mode):
    rp = _rp()
    z = np.load(f"{GOLDEN}/functions.npz")
    for tag in ("rz_2to1", "rz_3to1", "rz_frac", "rz_frac2"):
        src, md, dst = z[tag + "_src"], int(z[tag + "_md"]), z[tag + "_dst"]
        nw, nh = rp.resize_dims(src.shape[0], src.shape[1], md)
        got = host(eng_mode.resize_area(dev(src), nw, nh))
        assert np.abs(got.astype(np.int16) - dst.astype(np.int16)).max() <= (0 if "to1" in tag else 1), tag


@pytest.mark.parametrize("hs,ws,nh,nw", [(4096, 4096, 2000, 2000), (3000, 4000, 1500, 2000), (2160, 3840, 1125, 2000), (1201, 1604, 1000, 1335),
                                         (777, 1028, 300, 397), (64, 132, 9, 17), (5000, 2000, 2000, 800)])
def test_resize_tma_equals_staged_and_oracle(eng, hs, ws, nh, nw):
    """The TMA-staged separable kernel (default) against the register-tap kernel of round 1 (mode 3) and the C oracle:
    identical bytes (same float32 operation order)."""
    cm = _cm()
    src = synth(hs, ws, hs + 3 * ws, cell=24)
    d = dev(src)
    eng.set_fast_path(1)
    n0 = eng.launch_count()
    a = host(eng.resize_area(d, nw, nh))
    assert eng.launch_count() == n0 + 1
    eng.set_fast_path(3)
    b = host(eng.resize_area(d, nw, nh))
    eng.set_fast_path(1)
    assert np.array_equal(a, b)
    if hs * ws <= 4000 * 3000:
        assert np.array_equal(a, cm.resize_area(src, nw, nh))


def test_resize_strided_views(eng):
    """pitch != 3*w on both sides (a crop of a larger tensor)."""
    cm = _cm()
    src = synth(300, 420, 1)
    big = dev(src)
    crop = big[10:266, 20:404]                       # 256 x 384, pitch 1260
    out_big = torch.zeros((200, 300, 3), dtype=torch.uint8, device="cuda")
    eng.resize_area(crop, 192, 128, out=out_big[5:133, 7:199])
    assert np.array_equal(host(out_big[5:133, 7:199]), cm.resize_area(src[10:266, 20:404], 192, 128))


# ---- stage 02 --------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [2, 3, 4, 8, 16, 32])
def test_assign_lab(eng_mode, K):
    cm, rp = _cm(), _rp()
    rng = np.random.default_rng(K)
    for img in (synth(250, 333, K), uniform_img(127, 510, K + 1)):
        lab = cm.bgr2lab(img)
        ctr = rp.kmeans_lab_centers(img, K)
        assert np.array_equal(host(eng_mode.assign_lab(dev(img), ctr)), cm.assign_f32(lab, ctr))
        # near-ties: half-integer centres, a duplicated centre (first minimum must win), relabel LUT
        ctr2 = (rng.integers(0, 255, (K, 3)) + 0.5).astype(np.float32)
        ctr2[K - 1] = ctr2[0]
        lut = rng.permutation(K).astype(np.uint8)
        assert np.array_equal(host(eng_mode.assign_lab(dev(img), ctr2, lut)), lut[cm.assign_f32(lab, ctr2)])


def test_assign_lab_all_colours(eng):
    """Every one of the 2^24 BGR colours: the Lab tables + f32 argmin agree with the oracle."""
    cm = _cm()
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    ctr = (np.random.default_rng(3).random((8, 3)) * np.array([255, 120, 120]) + np.array([0, 60, 60])).astype(np.float32)
    want = cm.assign_f32(cm.bgr2lab(img), ctr)
    assert np.array_equal(host(eng.assign_lab(dev(img), ctr)), want)


@pytest.mark.parametrize("K", [2, 4, 8, 16])
def test_assign_rgb_i16wrap(eng_mode, K):
    cm = _cm()
    z = np.load(f"{GOLDEN}/functions.npz")
    got = host(eng_mode.assign_rgb_i16wrap(dev(z["al_img"]), z[f"al_pal{K}"]))
    assert np.array_equal(got, z[f"al_lab{K}"])               # golden from process_colors.assign_labels
    img = uniform_img(211, 307, K)
    pal = np.random.default_rng(K).integers(0, 256, (K, 3), dtype=np.uint8)
    assert np.array_equal(host(eng_mode.assign_rgb_i16wrap(dev(img), pal)), cm.assign_i16wrap(img, pal))
    assert np.array_equal(eng_mode.host_assign_rgb_i16wrap(img, pal), cm.assign_i16wrap(img, pal))


@pytest.mark.parametrize("hw", [(67, 93), (1, 50), (50, 1), (2, 2), (3, 3), (128, 200), (257, 1031), (64, 33)])
def test_layer_masks(eng_mode, hw):
    cm = _cm()
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    for K in (2, 5, 8, 16):
        coarse = rng.integers(0, K, ((hw[0] + 7) // 8, (hw[1] + 7) // 8), dtype=np.uint8)
        labels = np.kron(coarse, np.ones((8, 8), np.uint8))[:hw[0], :hw[1]]
        noise = rng.random(hw) < 0.15
        labels = np.where(noise, rng.integers(0, K, hw, dtype=np.uint8), labels).astype(np.uint8)
        for oi, ci in ((1, 1), (0, 0), (2, 1), (0, 3)):
            want = cm.layer_masks(labels, K, None, oi, ci)
            got = host(eng_mode.layer_masks(dev(labels), K, oi, ci))
            assert np.array_equal(got, want), (K, oi, ci)


# ---- stage 03 --------------------------------------------------------------------------------------
def _edge_want(masks, **kw):
    cm = _cm()
    return np.stack([cm.edge_chain(m, kw.get("morph_k", 3), kw.get("open_iters", 1), kw.get("close_iters", 1),
                                   kw.get("ksize", 3), kw.get("low", 50), kw.get("high", 150)) for m in masks])


@pytest.mark.parametrize("hw", [(67, 93), (128, 200), (1, 200), (200, 1), (2, 2), (3, 300), (300, 3), (513, 1027)])
def test_edges_binary_masks(eng_mode, hw):
    import omni_b200
    masks = np.stack([blob_mask(hw[0], hw[1], s, p) for s, p in ((1, 0.5), (2, 0.2), (3, 0.8))])
    for ks, lo, hi in [(3, 50, 150), (7, 22, 70), (5, 100, 200), (3, 50, 50), (3, 200, 50), (9, 30.7, 90.2)]:
        ec = omni_b200.EdgeConfig(low=lo, high=hi, ksize=ks)
        got = host(eng_mode.edges(dev(masks), ec))
        assert np.array_equal(got, _edge_want(masks, ksize=ks, low=lo, high=hi)), (ks, lo, hi)


@pytest.mark.parametrize("mk,oi,ci", [(1, 1, 1), (5, 1, 1), (3, 2, 0), (3, 0, 2), (2, 1, 1), (3, 0, 0), (7, 1, 2)])
def test_edges_morph_params(eng_mode, mk, oi, ci):
    import omni_b200
    masks = np.stack([blob_mask(150, 210, s, 0.5, k=5) for s in (4, 5)])
    ec = omni_b200.EdgeConfig(morph_k=mk, open_iters=oi, close_iters=ci)
    got = host(eng_mode.edges(dev(masks), ec))
    assert np.array_equal(got, _edge_want(masks, morph_k=mk, open_iters=oi, close_iters=ci))


def test_edges_nonbinary_long_chains(eng_mode):
    """Arbitrary u8 'masks' (the reference reads any mask.png): long weak chains crossing many tiles."""
    import omni_b200
    rp = _rp()
    masks = np.stack([smooth_u8(700, 900, 3), smooth_u8(700, 900, 4, k=11)])
    for lo, hi in [(10, 60), (5, 200), (50, 150)]:
        ec = omni_b200.EdgeConfig(low=lo, high=hi, ksize=3, open_iters=0, close_iters=0)
        got = host(eng_mode.edges(dev(masks), ec))
        want = np.stack([rp.edge_layer(m, lo, hi, 3, 3, 0, 0) for m in masks])
        assert np.array_equal(got, want), (lo, hi)
    assert eng_mode.last_hysteresis_passes() >= 1


def test_edges_spiral_worst_case(eng_mode):
    """A one-pixel weak spiral fed from a single strong seed: hysteresis must walk the whole chain."""
    import omni_b200
    cm = _cm()
    n = 384
    img = np.zeros((n, n), np.uint8)
    # concentric square spiral, 6 px pitch, value 60 (weak after blur); one bright seed at the start
    y = x = 4
    dy, dx, run = 0, 1, n - 9
    while run > 12:
        for _ in range(run):
            img[y, x] = 90
            y, x = y + dy, x + dx
        dy, dx = dx, -dy
        run -= 6
    img[4, 4:10] = 255
    ec = omni_b200.EdgeConfig(low=20, high=250, ksize=3, open_iters=0, close_iters=0)
    got = host(eng_mode.edges(dev(img[None]), ec))[0]
    want = cm.edge_chain(img, 3, 0, 0, 3, 20, 250)
    assert np.array_equal(got, want)
    assert want.any()


def _sparse_cases(h, w):
    """Masks built to stress the tile-run lists of the sparse edge kernel: uniform planes, shapes touching every
    border, features on tile seams (rows 8j, words 32c), isolated pixels, long vertical runs."""
    z = np.zeros((h, w), np.uint8)
    cases = [z.copy(), np.full((h, w), 255, np.uint8)]
    a = z.copy(); a[: max(1, h // 3)] = 255; cases.append(a)                     # top band (row seam somewhere)
    a = z.copy(); a[:, : max(1, w // 2)] = 255; cases.append(a)                  # left half: one long vertical boundary
    a = z.copy(); a[:, -3:] = 255; a[-2:] = 255; cases.append(a)                 # right / bottom border strips
    a = z.copy(); a[::16, ::64] = 255; a[7::24, 31::96] = 255; cases.append(a)   # isolated pixels on seams
    a = z.copy()
    for y0 in range(0, h, 40):
        for x0 in range(0, w, 100):
            a[y0 + 6:y0 + 18, x0 + 28:x0 + 70] = 255                             # boxes straddling tile rows and words
    cases.append(a)
    a = np.full((h, w), 255, np.uint8); a[h // 2, w // 2] = 0; a[0, 0] = 0; a[-1, -1] = 0; cases.append(a)
    return np.stack(cases)


@pytest.mark.parametrize("hw", [(8, 32), (9, 33), (63, 250), (130, 517), (257, 96), (300, 1100)])
def test_edges_sparse_tile_runs(eng_mode, hw):
    import omni_b200
    masks = _sparse_cases(*hw)
    for oi, ci in ((0, 0), (1, 1)):
        for lo, hi in ((50, 150), (0, 10), (200, 2000)):
            ec = omni_b200.EdgeConfig(low=lo, high=hi, ksize=3, open_iters=oi, close_iters=ci)
            out = torch.full((masks.shape[0],) + hw, 7, dtype=torch.uint8, device="cuda")   # stale bytes must be overwritten
            got = host(eng_mode.edges(dev(masks), ec, out=out))
            assert np.array_equal(got, _edge_want(masks, open_iters=oi, close_iters=ci, low=lo, high=hi)), (oi, ci, lo, hi)


def test_edges_strided_planes(eng):
    import omni_b200
    masks = np.stack([blob_mask(200, 300, s) for s in (7, 8, 9)])
    big = torch.zeros((3, 220, 352), dtype=torch.uint8, device="cuda")
    big[:, 10:210, 16:316] = dev(masks)
    out = torch.zeros((3, 230, 320), dtype=torch.uint8, device="cuda")
    eng.edges(big[:, 10:210, 16:316], omni_b200.EdgeConfig(), out=out[:, 3:203, 5:305])
    assert np.array_equal(host(out[:, 3:203, 5:305]), _edge_want(masks))
    assert int(out[:, :3].sum()) == 0 and int(out[:, :, :5].sum()) == 0     # nothing written outside the view


# ---- fused path + golden pipeline cases ------------------------------------------------------------------
@pytest.mark.parametrize("case", PIPE_CASES)
def test_golden_pipeline_case(eng_mode, case):
    """Frozen outputs of `pipeline.py --start-step 1 --end-step 3` of the unmodified reference."""
    import omni_b200
    rp = _rp()
    z, meta = load_pipe_case(case)
    cfg, names = meta["config"], meta["names"]
    K = len(names)
    src = z["input"]
    dims = rp.resize_dims(src.shape[0], src.shape[1], cfg["max_dimension"])
    resized = src if dims is None else host(eng_mode.resize_area(dev(src), dims[0], dims[1]))
    assert np.abs(resized.astype(np.int16) - z["resized"].astype(np.int16)).max() <= 1
    ctr = z["centers"]
    _order, lut = rp.darkness_order(ctr)
    ec = omni_b200.EdgeConfig(low=cfg["edge_low_threshold"], high=cfg["edge_high_threshold"], ksize=cfg["edge_kernel_size"],
                              morph_k=cfg["edge_morph_kernel"], open_iters=cfg["edge_morph_open_iters"],
                              close_iters=cfg["edge_morph_close_iters"])
    labels, masks, edges = eng_mode.color_edge(dev(z["resized"]), ctr, lut.astype(np.uint8), ec, want_labels=True)
    assert np.array_equal(host(labels), lut.astype(np.uint8)[z["labels"]])
    p_of = plane_of_name(names)
    for i, n in enumerate(names):
        assert np.array_equal(host(masks[p_of[n]]), z["masks"][i]), n
        assert np.array_equal(host(edges[p_of[n]]), z["edges"][i]), n
    order_planes = [p_of[n] for n in names]
    comp = host(eng_mode.edges_composite(edges[order_planes], cfg["colors"]))
    assert np.array_equal(comp, z["composite"])
    # host-buffer entry point: same result + the counts the reference logs / stores in palette_by_name.json
    r = eng_mode.host_color_edge(z["resized"], ctr, lut.astype(np.uint8), ec)
    for i, n in enumerate(names):
        p = p_of[n]
        assert np.array_equal(r["masks"][p], z["masks"][i]) and np.array_equal(r["edges"][p], z["edges"][i])
        assert r["counts"][p, 1] == meta["palette_by_name"][n]["mask_nonzero"]
        assert r["counts"][p, 0] == meta["palette_by_name"][n]["pixels"]
        assert r["counts"][p, 2] == int(np.count_nonzero(z["edges"][i]))


@pytest.mark.parametrize("hw,K", [((1, 1), 2), ((1, 37), 3), ((41, 1), 2), ((5, 7), 4), ((64, 64), 4), ((65, 129), 8),
                                  ((333, 517), 16), ((1080, 1920), 8)])
def test_fused_equals_oracle(eng_mode, hw, K):
    import omni_b200
    cm, rp = _cm(), _rp()
    img = synth(hw[0], hw[1], K + hw[1], cell=16) if min(hw) >= 16 else uniform_img(hw[0], hw[1], K)
    ctr = rp.kmeans_lab_centers(img, K) if hw[0] * hw[1] >= K else \
        np.random.default_rng(1).random((K, 3)).astype(np.float32) * 255
    _o, lut = rp.darkness_order(ctr)
    lut = lut.astype(np.uint8)
    for ks, lo, hi in [(3, 50, 150), (7, 22, 70)]:
        ec = omni_b200.EdgeConfig(low=lo, high=hi, ksize=ks)
        labels, masks, edges = eng_mode.color_edge(dev(img), ctr, lut, ec, want_labels=True)
        raw = cm.assign_f32(cm.bgr2lab(img), ctr)
        assert np.array_equal(host(labels), lut[raw])
        wm = cm.layer_masks(raw, K, lut)
        assert np.array_equal(host(masks), wm)
        assert np.array_equal(host(edges), _edge_want(wm, ksize=ks, low=lo, high=hi))
        assert np.array_equal(eng_mode.count_nonzero(masks), [(m != 0).sum() for m in wm])


def test_param_sweep_config5_scaled(eng):
    """BASELINE config 5 (blur 3/5/7 x thresholds 50..200, K=16) at 1024^2 against the cv2 chain."""
    import omni_b200
    rp = _rp()
    img = synth(1024, 1024, 0)
    K = 16
    ctr = rp.kmeans_lab_centers(img, K)
    _o, lut = rp.darkness_order(ctr)
    _l, masks_d, _e = eng.color_edge(dev(img), ctr, lut.astype(np.uint8), omni_b200.EdgeConfig())
    masks = host(masks_d)
    _c, _ls, want_masks = rp.color_extract(img, K, ctr)
    assert np.array_equal(masks, want_masks)
    for ks in (3, 5, 7):
        for lo in (50, 100, 150):
            for hi in (100, 150, 200):
                got = host(eng.edges(masks_d, omni_b200.EdgeConfig(low=lo, high=hi, ksize=ks)))
                want = rp.edges_all(masks, low=lo, high=hi, ksize=ks)
                assert np.array_equal(got, want), (ks, lo, hi)


def test_full_size_config2(eng):
    """BASELINE config 2 at full size (4096^2, K=8, 50/150, blur 3): masks and edges bit-exact vs the
    oracle (C assign + cv2 chain), plus size-independent invariants."""
    import omni_b200
    cm, rp = _cm(), _rp()
    img = synth(4096, 4096, 0)
    K = 8
    ctr = rp.kmeans_lab_centers(img, K)
    _o, lut = rp.darkness_order(ctr)
    lut = lut.astype(np.uint8)
    labels, masks, edges = eng.color_edge(dev(img), ctr, lut, omni_b200.EdgeConfig(), want_labels=True)
    want_labels = lut[cm.assign_f32(cm.bgr2lab(img), ctr)]
    assert np.array_equal(host(labels), want_labels)
    want_masks = rp.layer_masks(want_labels, K)
    assert np.array_equal(host(masks), want_masks)
    assert np.array_equal(host(edges), rp.edges_all(want_masks))
    # invariants: label histogram sums to N; edges only where the blurred mask has a gradient, i.e. never
    # deeper than 3 px inside a uniform region of the mask
    assert int(torch.bincount(labels.flatten().int(), minlength=K).sum()) == 4096 * 4096
    # unfused composition == fused call, on the device
    m2 = eng.layer_masks(eng.assign_lab(dev(img), ctr, lut), K)
    assert torch.equal(m2, masks)
    assert torch.equal(eng.edges(m2, omni_b200.EdgeConfig()), edges)


@pytest.mark.parametrize("shape", [(4096, 4096, 16, 0, 32), (1080, 1920, 8, 0, 32), (2000, 2000, 4, 2, 32)],
                         ids=["config5_4096_k16", "config4_1080p_k8", "default_2000_k4"])
def test_full_size_families_agree(eng, shape):
    """BASELINE configs 4, 5 (and the default max_dimension size) at FULL size: the three fast kernel families behind the ABI --
    label-domain pipeline (default), first generation (RGB cells + sparse tile runs), dense edges + Lab cells -- must give
    identical bytes; the oracle pins one layer here (configs[2], every layer: tests/test_gpu_packed.py)."""
    import omni_b200
    cm, rp = _cm(), _rp()
    h, w, K, seed, cell = shape
    img = synth(h, w, seed, cell)
    ctr = rp.kmeans_lab_centers(img, K)
    _o, lut = rp.darkness_order(ctr)
    lut = lut.astype(np.uint8)
    d = dev(img)
    ec = omni_b200.EdgeConfig()
    try:
        eng.set_fast_path(1)
        l1, m1, e1 = eng.color_edge(d, ctr, lut, ec, want_labels=True)
        for mode in (3, 2):
            eng.set_fast_path(mode)
            l2, m2, e2 = eng.color_edge(d, ctr, lut, ec, want_labels=True)
            assert torch.equal(l1, l2) and torch.equal(m1, m2) and torch.equal(e1, e2), mode
    finally:
        eng.set_fast_path(1)
    # size-independent property: the labels partition the image
    assert int(torch.bincount(l1.flatten().int(), minlength=K).sum()) == h * w
    k = K // 2
    want_mask = rp.layer_masks((host(l1) == k).astype(np.uint8), 2)[1]
    assert np.array_equal(host(m1[k]), want_mask)
    assert np.array_equal(host(e1[k]), rp.edge_layer(want_mask))
    # stage 04 hand-off at full size: skeleton is a subset of the edges and idempotent
    sk = eng.thin_zhangsuen(e1)
    assert bool(((sk > 0) <= (e1 > 0)).all())
    assert torch.equal(eng.thin_zhangsuen(sk), sk)
    assert np.array_equal(host(sk[k]), cm.thin_zhangsuen(host(e1[k])))


def test_no_cpu_fallback_error_paths(eng):
    import omni_b200
    with pytest.raises(omni_b200.OmniError):
        eng.resize_area(dev(synth(64, 64, 0)), 128, 128)          # INTER_AREA path is shrink-only
    with pytest.raises(omni_b200.OmniError):
        eng.edges(dev(blob_mask(64, 64, 0)[None]), omni_b200.EdgeConfig(ksize=33))
    with pytest.raises(omni_b200.OmniError):
        eng.assign_lab(dev(synth(64, 64, 0)), np.zeros((40, 3), np.float32))


def _spiral_binary(n=384, pitch=8):
    img = np.zeros((n, n), np.uint8)
    y = x = 6
    dy, dx, run = 0, 1, n - 13
    while run > 2 * pitch:
        for _ in range(run):
            img[y, x] = 255
            y, x = y + dy, x + dx
        dy, dx = dx, -dy
        run -= pitch
    img[3:10, 3:10] = 255
    return img


def test_edges_binary_spiral_long_chain(eng_mode):
    """A {0,255} mask (fast bit-plane path) whose weak chain is ~8700 pixels deep from 4-8 strong seeds: the
    cooperative hysteresis must escalate from word rounds to tile rounds and still reach the fixed point."""
    import omni_b200
    cm = _cm()
    img = _spiral_binary()
    masks = np.stack([img, np.ascontiguousarray(img.T), np.ascontiguousarray(img[::-1])])
    ec = omni_b200.EdgeConfig(low=100, high=900, ksize=3, open_iters=0, close_iters=0)
    got = host(eng_mode.edges(dev(masks), ec))
    want = np.stack([cm.edge_chain(m, 3, 0, 0, 3, 100, 900) for m in masks])
    assert want[0].sum() > 15000 * 255
    assert np.array_equal(got, want)
    assert eng_mode.last_hysteresis_passes() > 4


# ---- stage 04: thinning -----------------------------------------------------------------------
def test_thinning_golden(eng):
    """Skeletons frozen from the unmodified reference (tools/make_golden_thinning.py), incl. thick masks (many iterations)."""
    z = np.load(f"{GOLDEN}/thinning.npz")
    names = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    cm = _cm()
    for n in names:
        src, want = z[n + "_in"], z[n + "_out"]
        got, removed, iters = eng.thin_zhangsuen(dev(src[None]), with_log=True)
        assert np.array_equal(host(got)[0], want), n
        _sk, log = cm.thin_zhangsuen(src, with_log=True)
        assert int(iters[0]) == len(log) and np.array_equal(removed[0, :len(log)], log), n     # the reference's log numbers


@pytest.mark.parametrize("hw", [(1, 1), (1, 70), (70, 1), (2, 2), (31, 33), (64, 64), (97, 161), (300, 515)])
def test_thinning_shapes_vs_oracle(eng, hw):
    cm = _cm()
    h, w = hw
    rng = np.random.default_rng(h * 1000 + w)
    planes = np.stack([blob_mask(h, w, 11, 0.5, k=5) if min(h, w) > 12 else (rng.random((h, w)) < 0.6).astype(np.uint8) * 255,
                       (rng.random((h, w)) < 0.7).astype(np.uint8) * 200,                   # any value > 0 is foreground
                       np.full((h, w), 255, np.uint8), np.zeros((h, w), np.uint8)])
    got, removed, iters = eng.thin_zhangsuen(dev(planes), with_log=True)
    got = host(got)
    for k in range(planes.shape[0]):
        want, log = cm.thin_zhangsuen(planes[k], with_log=True)
        assert np.array_equal(got[k], want), k
        assert int(iters[k]) == len(log) and np.array_equal(removed[k, :len(log)], log), k
    # host-buffer form and the iteration cap
    out_h, _r, _i = eng.host_thin_zhangsuen(planes)
    assert np.array_equal(out_h, got)
    cut = host(eng.thin_zhangsuen(dev(planes), max_iter=2))
    for k in range(planes.shape[0]):
        assert np.array_equal(cut[k], cm.thin_zhangsuen(planes[k], max_iter=2)), k


def test_thinning_in_place_and_strided(eng):
    cm = _cm()
    planes = np.stack([blob_mask(120, 200, s, 0.4, k=7) for s in (3, 4)])
    big = torch.zeros((2, 140, 256), dtype=torch.uint8, device="cuda")
    big[:, 7:127, 19:219] = dev(planes)
    view = big[:, 7:127, 19:219]
    eng.thin_zhangsuen(view, out=view)                                                      # in and out alias
    want = np.stack([cm.thin_zhangsuen(p) for p in planes])
    assert np.array_equal(host(view), want)
    assert int(big[:, :7].sum()) == 0 and int(big[:, :, :19].sum()) == 0 and int(big[:, :, 219:].sum()) == 0


def test_thinning_edges_of_fused_path(eng):
    """Stage 03 -> 04 hand-off on the device: thin the edge planes of a synthetic image, compare with the oracle chain."""
    import omni_b200
    cm, rp = _cm(), _rp()
    img = synth(512, 768, 5)
    K = 4
    centers = rp.kmeans_lab_centers(img, K)
    _order, lut = rp.darkness_order(centers)
    _l, _m, edges = eng.color_edge(dev(img), centers, lut.astype(np.uint8), omni_b200.EdgeConfig())
    sk = host(eng.thin_zhangsuen(edges))
    e = host(edges)
    for k in range(K):
        assert np.array_equal(sk[k], cm.thin_zhangsuen(e[k])), k
    assert (sk > 0).sum() < (e > 0).sum()


def test_thinning_function_mirror(eng, capsys):
    from omni_b200 import contours
    cm = _cm()
    img = blob_mask(90, 130, 21, 0.3, k=7)
    out = contours.thinning_zhangsuen(img, "layer_x")
    assert np.array_equal(out, cm.thin_zhangsuen(img))
    log = capsys.readouterr().out
    assert "[layer_x] Thinning ROI" in log and "Thin 01: removed=" in log and "Thinning done" in log
    assert np.array_equal(contours.thinning_zhangsuen(np.zeros((5, 5), np.uint8), "e"), np.zeros((5, 5), np.uint8))


# ---- stage 02 swatch mode -------------------------------------------------------------------------
def test_swatch_masks_golden(eng):
    z = np.load(f"{GOLDEN}/swatch.npz")
    for tol in (30, 60, 8):
        got = host(eng.swatch_masks(dev(z["img"]), z["colors"], tol))
        assert np.array_equal(got, z[f"masks_tol{tol}"]), tol


@pytest.mark.parametrize("hw", [(1, 1), (7, 45), (130, 517), (512, 768)])
def test_swatch_masks_vs_oracle(eng, hw):
    cm = _cm()
    img = synth(hw[0], hw[1], 17, cell=8) if min(hw) >= 8 else uniform_img(hw[0], hw[1], 3)
    px = img.reshape(-1, 3)
    cols = [px[0][::-1].tolist(), px[len(px) // 2].tolist(), [0, 0, 0], [255, 255, 255], [128, 10, 240], px[-1].tolist()]
    for tol in (0, 30, 90, 255):
        want, wc = cm.swatch_masks(img, cols, tol, with_choice=True)
        got, gc = eng.swatch_masks(dev(img), cols, tol, with_choice=True)
        assert np.array_equal(host(got), want), tol
        assert np.array_equal(gc, wc), tol
    with pytest.raises(Exception):
        eng.swatch_masks(dev(img), [[300, 0, 0]], 30)


def test_swatch_stage_function(eng, tmp_path):
    """color_extract_main with a hand-built Config (extraction_mode='swatch'): files and log lines of 02:82-109."""
    import cv2
    from omni_b200 import stages, config as ocfg
    z = np.load(f"{GOLDEN}/swatch.npz")
    cfg = ocfg.Config()
    cfg.output_dir = str(tmp_path)
    cfg.color_names = ["layer_a", "layer_b", "layer_c", "layer_d"]
    cfg.colors = z["colors"].tolist()
    cfg.extraction_mode = "swatch"
    cfg.color_tolerance = 60
    cv2.imwrite(str(tmp_path / "resized.png"), z["img"])
    stages.color_extract_main(cfg)
    for i, n in enumerate(cfg.color_names):
        assert np.array_equal(cv2.imread(str(tmp_path / n / "mask.png"), cv2.IMREAD_GRAYSCALE), z["masks_tol60"][i]), n
    cfg.colors = cfg.colors[:2]
    with pytest.raises(RuntimeError):
        stages.color_extract_main(cfg)



# ---- frame batches (BASELINE config 4) ------------------------------------------------------------------------------
@pytest.mark.parametrize("n,K,hw", [(1, 4, (90, 130)), (4, 8, (135, 240)), (5, 8, (64, 100)), (3, 16, (70, 97)), (2, 32, (40, 64))])
def test_color_edge_batch_equals_per_frame(eng_mode, n, K, hw):
    """omni_color_edge_batch == n calls of omni_color_edge (incl. n*K > 32: several passes), and one frame vs the oracle."""
    import omni_b200
    rp = _rp()
    frames = np.stack([synth(hw[0], hw[1], 50 + f, cell=16) for f in range(n)])
    ctr = rp.kmeans_lab_centers(frames[0], K)
    _o, lut = rp.darkness_order(ctr)
    lut = lut.astype(np.uint8)
    ec = omni_b200.EdgeConfig()
    d = dev(frames)
    masks = torch.full((n, K) + hw, 9, dtype=torch.uint8, device="cuda")
    edges = torch.full((n, K) + hw, 9, dtype=torch.uint8, device="cuda")
    eng_mode.color_edge_batch(d, ctr, lut, ec, masks=masks, edges=edges)
    for f in range(n):
        _l, m, e = eng_mode.color_edge(d[f], ctr, lut, ec)
        assert torch.equal(masks[f], m) and torch.equal(edges[f], e), f
    _cs, _ls, want = rp.color_extract(frames[-1], K, ctr)
    assert np.array_equal(host(masks[-1]), want)
    assert np.array_equal(host(edges[-1]), rp.edges_all(want))


@pytest.mark.parametrize("hw", [(1, 1), (3, 40), (64, 64), (211, 333)])
def test_skeleton_degree(eng, hw):
    """04_find_contours.py:121-125 (degree / endpoint / junction maps) vs the reference's filter2D call on the whole skeleton."""
    rp, cm = _rp(), _cm()
    h, w = hw
    rng = np.random.default_rng(h + w)
    planes = np.stack([cm.thin_zhangsuen(blob_mask(h, w, 5, 0.5, k=5)) if min(h, w) > 12 else (rng.random((h, w)) < 0.5).astype(np.uint8) * 255,
                       (rng.random((h, w)) < 0.3).astype(np.uint8) * 7, np.full((h, w), 255, np.uint8)])
    deg, nodes = eng.skeleton_degree(dev(planes))
    deg, nodes = host(deg), host(nodes)
    for k in range(planes.shape[0]):
        wd, ep, jn = rp.skeleton_degree(planes[k])
        assert np.array_equal(deg[k], wd), k
        assert np.array_equal(nodes[k] == 1, ep) and np.array_equal(nodes[k] == 2, jn), k


@pytest.mark.parametrize("hw,K", [((2200, 300), 4), ((2048, 96), 8), ((4100, 64), 3), ((300, 500), 5)])
def test_host_color_edge_bands(eng_mode, hw, K):
    """The band-pipelined host-buffer call (short lead bands on tall images) == the device-resident call, incl. counts."""
    import omni_b200
    rp = _rp()
    img = synth(hw[0], hw[1], 7, cell=16)
    ctr = rp.kmeans_lab_centers(img, K)
    _o, lut = rp.darkness_order(ctr)
    lut = lut.astype(np.uint8)
    ec = omni_b200.EdgeConfig()
    r = eng_mode.host_color_edge(img, ctr, lut, ec, want_labels=True, want_counts=True)
    labels, masks, edges = eng_mode.color_edge(dev(img), ctr, lut, ec, want_labels=True)
    assert np.array_equal(r["labels"], host(labels))
    assert np.array_equal(r["masks"], host(masks))
    assert np.array_equal(r["edges"], host(edges))
    assert np.array_equal(r["counts"][:, 1], (host(masks) > 0).reshape(K, -1).sum(1))
    assert np.array_equal(r["counts"][:, 2], (host(edges) > 0).reshape(K, -1).sum(1))


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_fused_vs_oracle(eng, seed):
    """Randomised shapes / centres / thresholds through the fused call (sparse tile runs, RGB cells, zero fill, run lists
    from the morphology kernel) against the C oracle: odd sizes, adversarial centres (duplicates, half-integers, far away),
    float thresholds, every stage-03 morphology setting of the fast path."""
    import omni_b200
    cm = _cm()
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(1, 260)), int(rng.integers(1, 700))
    K = int(rng.integers(2, 17))
    kind = seed % 4
    if kind == 0:
        img = uniform_img(h, w, seed)
    elif kind == 1:
        img = synth(max(h, 8), max(w, 8), seed, cell=8)[:h, :w]
    elif kind == 2:                                       # few flat regions: many uniform tiles, long straight boundaries
        img = np.zeros((h, w, 3), np.uint8)
        img[:, : w // 2] = rng.integers(0, 256, 3)
        img[h // 3:, w // 3:] = rng.integers(0, 256, 3)
        img[:, -1:] = rng.integers(0, 256, 3)
    else:
        img = np.clip(synth(max(h, 8), max(w, 8), seed, cell=4)[:h, :w].astype(np.int16) + rng.integers(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(img)
    ctr = (rng.random((K, 3)) * np.array([255, 190, 190]) + np.array([0, 30, 30])).astype(np.float32)
    if seed % 3 == 0:
        ctr = np.round(ctr * 2) / 2                       # half-integer centres: exact ties between centres are possible
    if seed % 5 == 0 and K > 2:
        ctr[1] = ctr[0]                                   # duplicated centre: first minimum must win
    if seed % 7 == 0:
        ctr[-1] = [400.0, -50.0, 300.0]                   # a centre no colour is close to
    lut = rng.permutation(K).astype(np.uint8)
    low, high = float(rng.uniform(0, 120)), float(rng.uniform(20, 400))
    mk = [3, 3, 1, 3][seed % 4]
    oi, ci = [(1, 1), (1, 0), (1, 1), (0, 1)][(seed // 4) % 4]
    ec = omni_b200.EdgeConfig(low=low, high=high, ksize=3, morph_k=mk, open_iters=oi, close_iters=ci)
    labels, masks, edges = eng.color_edge(dev(img), ctr, lut, ec, want_labels=True)
    raw = cm.assign_f32(cm.bgr2lab(img), ctr)
    assert np.array_equal(host(labels), lut[raw])
    wm = cm.layer_masks(raw, K, lut)
    assert np.array_equal(host(masks), wm)
    assert np.array_equal(host(edges), _edge_want(wm, ksize=3, low=low, high=high, morph_k=mk, open_iters=oi, close_iters=ci))


def test_assign_lut_outside_k(eng):
    """A label map with values >= K (allowed by the ABI, < 32): the RGB-cell tables cannot hold them -> Lab-cell kernel; same labels."""
    cm = _cm()
    img = synth(120, 333, 4, cell=16)
    K = 5
    ctr = _rp().kmeans_lab_centers(img, K)
    lut = np.array([20, 3, 31, 0, 15], np.uint8)
    got = host(eng.assign_lab(dev(img), ctr, lut))
    assert np.array_equal(got, lut[cm.assign_f32(cm.bgr2lab(img), ctr)])
