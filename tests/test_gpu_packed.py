"""Packed (1 bit per pixel) outputs of the fused call and the parity holes the round-1 review named: every layer of the largest
config against the oracle, BASELINE config 5 at full size, host views with row gaps, the flag blocks of hysteresis / thinning.
Needs a B200: `-m gpu`."""
import numpy as np
import pytest

from helpers import synth, uniform_img, blob_mask

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    import omni_b200
    e = omni_b200.Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def _rp():
    from oracle import refport
    return refport


def _centres(img, K):
    rp = _rp()
    ctr = rp.kmeans_lab_centers(img, K)
    _o, lut = rp.darkness_order(ctr)
    return ctr, lut.astype(np.uint8)


@pytest.mark.parametrize("mode", [1, 3, 0], ids=["fast", "fast_dense_pipeline", "generic"])
@pytest.mark.parametrize("hw,K", [((64, 96), 4), ((33, 41), 3), ((257, 1030), 8), ((300, 517), 16), ((120, 250), 20), ((1, 1), 2), ((9, 8), 2)])
def test_color_edge_packed_equals_bytes(eng, mode, hw, K):
    """The packed planes are the byte planes of color_edge, 8 pixels per byte, in both bit orders; ec=None gives the masks alone;
    counts = pixels per label / non-zeros per plane."""
    import omni_b200
    h, w = hw
    img = synth(h, w, K + w, cell=16) if min(hw) >= 16 else uniform_img(h, w, K)
    ctr, lut = _centres(img, K) if h * w >= K else (np.random.default_rng(1).random((K, 3)).astype(np.float32) * 255, np.arange(K, dtype=np.uint8))
    ec = omni_b200.EdgeConfig()
    try:
        eng.set_fast_path(mode)
        labels, masks, edges = eng.color_edge(dev(img), ctr, lut, ec, want_labels=True)
        for msb in (True, False):
            order = "big" if msb else "little"
            mb, eb, counts = eng.color_edge_packed(dev(img), ctr, lut, ec, msb_first=msb, want_counts=True)
            assert np.array_equal(host(mb), np.packbits(host(masks) > 0, axis=2, bitorder=order))
            assert np.array_equal(host(eb), np.packbits(host(edges) > 0, axis=2, bitorder=order))
            assert np.array_equal(counts[:, 0], np.bincount(host(labels).ravel(), minlength=K)[:K])
            assert np.array_equal(counts[:, 1], (host(masks) > 0).reshape(K, -1).sum(1))
            assert np.array_equal(counts[:, 2], (host(edges) > 0).reshape(K, -1).sum(1))
        mb, eb = eng.color_edge_packed(dev(img), ctr, lut, None)
        assert eb is None and np.array_equal(host(mb), np.packbits(host(masks) > 0, axis=2))
    finally:
        eng.set_fast_path(1)


def test_packed_frame_batches_host_and_device(eng):
    """n frames with one centre set: device batch (n * K <= 32) and the pipelined host call (any n: 11 frames = 3 groups of 4 at
    K = 8) give the per-frame result; the 1-bit PNG written from a packed plane decodes to the byte plane."""
    import cv2
    import omni_b200
    from omni_b200 import png1
    h, w, K, n = 135, 250, 8, 11
    frames = np.stack([synth(h, w, 50 + i, cell=16) for i in range(n)])
    ctr, lut = _centres(frames[0], K)
    ec = omni_b200.EdgeConfig(low=60, high=120)
    want_m, want_e = [], []
    for f in frames:
        _l, m, e = eng.color_edge(dev(f), ctr, lut, ec)
        want_m.append(np.packbits(host(m) > 0, axis=2)); want_e.append(np.packbits(host(e) > 0, axis=2))
    want_m, want_e = np.concatenate(want_m), np.concatenate(want_e)
    mb, eb = eng.color_edge_packed(dev(frames[:4]), ctr, lut, ec)
    assert np.array_equal(host(mb), want_m[:4 * K]) and np.array_equal(host(eb), want_e[:4 * K])
    r = eng.host_color_edge_packed(frames, ctr, lut, ec)
    assert np.array_equal(r["mask_bits"], want_m) and np.array_equal(r["edge_bits"], want_e)
    assert np.array_equal(r["counts"][:, 1], np.unpackbits(want_m, axis=2).reshape(n * K, -1).sum(1))
    assert np.array_equal(r["counts"][:, 2], np.unpackbits(want_e, axis=2).reshape(n * K, -1).sum(1))
    assert int(r["counts"][:, 0].sum()) == n * h * w
    r2 = eng.host_color_edge_packed(frames[:5], ctr, lut, None, want_counts=False)          # masks only, pinned inputs
    assert r2["edge_bits"] is None and np.array_equal(r2["mask_bits"], want_m[:5 * K])
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "mask.png")
        png1.write_png1(p, r["mask_bits"][3], w)
        got = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
        assert got.dtype == np.uint8 and np.array_equal(got, np.unpackbits(want_m[3], axis=1)[:, :w] * 255)


def test_full_size_config3_every_layer(eng):
    """BASELINE configs[2] (8192^2, K=16) at full size: labels, ALL 16 masks and ALL 16 edge planes against the oracle (the review
    of round 1: one layer is not evidence), the three fast families identical."""
    import omni_b200
    from oracle import cmodel as cm
    rp = _rp()
    h = w = 8192
    K = 16
    img = synth(h, w, 1, 64)
    ctr, lut = _centres(img, K)
    d = dev(img)
    ec = omni_b200.EdgeConfig()
    try:
        eng.set_fast_path(1)
        l1, m1, e1 = eng.color_edge(d, ctr, lut, ec, want_labels=True)
        for mode in (3, 2):
            eng.set_fast_path(mode)
            l2, m2, e2 = eng.color_edge(d, ctr, lut, ec, want_labels=True)
            assert torch.equal(l1, l2) and torch.equal(m1, m2) and torch.equal(e1, e2), mode
            del l2, m2, e2
    finally:
        eng.set_fast_path(1)
    want_labels = lut[cm.assign_f32(cm.bgr2lab(img), ctr)]
    assert np.array_equal(host(l1), want_labels)
    for k in range(K):
        want_mask = rp.layer_masks((want_labels == k).astype(np.uint8), 2)[1]
        assert np.array_equal(host(m1[k]), want_mask), k
        assert np.array_equal(host(e1[k]), rp.edge_layer(want_mask)), k


def test_param_sweep_config5_full_size(eng):
    """BASELINE configs[4] as named: blur {3,5,7} x low {50,100,150} x high {100,150,200} on 4096^2, K=16; the K=16 masks against
    refport.color_extract, every edge plane of every sweep point against the cv2 chain of 03:23-34."""
    import cv2
    import omni_b200
    rp = _rp()
    img = synth(4096, 4096, 0)
    K = 16
    ctr, lut = _centres(img, K)
    _l, masks_d, _e = eng.color_edge(dev(img), ctr, lut, omni_b200.EdgeConfig())
    masks = host(masks_d)
    want_labels = lut[rp.assign_lab_chunked(img, ctr)]
    assert np.array_equal(masks, rp.layer_masks(want_labels, K))
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    m2 = [cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_OPEN, se), cv2.MORPH_CLOSE, se) for m in masks]
    for ks in (3, 5, 7):
        blurred = [cv2.GaussianBlur(m, (ks, ks), 0) for m in m2]
        for lo in (50, 100, 150):
            for hi in (100, 150, 200):
                got = host(eng.edges(masks_d, omni_b200.EdgeConfig(low=lo, high=hi, ksize=ks)))
                for k in range(K):
                    assert np.array_equal(got[k], cv2.Canny(blurred[k], lo, hi)), (ks, lo, hi, k)


def test_host_views_with_row_gaps_are_not_overwritten(eng):
    """Host outputs that are views into a wider array (pitch > w, pitch % 16 == 0): the bytes between the rows stay untouched."""
    import omni_b200
    h, w, K = 96, 160, 4
    img = synth(h, w, 5, cell=16)
    ctr, lut = _centres(img, K)
    ec = omni_b200.EdgeConfig()
    big_m = np.full((K, h, w + 32), 0xA5, np.uint8)
    big_e = np.full((K, h, w + 32), 0x5A, np.uint8)
    r = eng.host_color_edge(img, ctr, lut, ec, want_labels=False, masks=big_m[:, :, :w], edges=big_e[:, :, :w])
    _l, m, e = eng.color_edge(dev(img), ctr, lut, ec)
    assert np.array_equal(big_m[:, :, :w], host(m)) and np.array_equal(big_e[:, :, :w], host(e))
    assert (big_m[:, :, w:] == 0xA5).all() and (big_e[:, :, w:] == 0x5A).all()
    assert r["masks"].base is big_m or r["masks"] is not None
    masks = np.stack([blob_mask(h, w, s) for s in (1, 2, 3)])
    big_o = np.full((3, h, w + 48), 0x77, np.uint8)
    eng.host_edges(masks, ec, out=big_o[:, :, :w])
    assert np.array_equal(big_o[:, :, :w], host(eng.edges(dev(masks), ec)))
    assert (big_o[:, :, w:] == 0x77).all()


def test_hysteresis_pass_count_survives_thinning(eng):
    """omni_last_hysteresis_passes reports the edge pass, not the iteration count of a later thinning call."""
    import omni_b200
    img = synth(256, 384, 9, cell=16)
    ctr, lut = _centres(img, 4)
    _l, _m, e = eng.color_edge(dev(img), ctr, lut, omni_b200.EdgeConfig())
    torch.cuda.synchronize()
    p0 = eng.last_hysteresis_passes()
    _l, _m, e = eng.color_edge(dev(img), ctr, lut, omni_b200.EdgeConfig())
    _s, removed, iters = eng.thin_zhangsuen(e, with_log=True)
    assert int(iters.max()) >= 2                         # thinning ran several iterations ...
    assert eng.last_hysteresis_passes() == p0 >= 1       # ... and did not overwrite the pass count


def test_reserved_ctx_does_not_allocate_in_the_call():
    """omni_workspace_bytes / omni_ctx_reserve: after the reservation a fused call of that geometry (device-resident, byte planes or
    packed, blur 3 or 5) leaves the free device memory where it was -- no cudaMalloc inside the timed call."""
    import omni_b200
    e = omni_b200.Engine(0)
    try:
        h, w, K = 1100, 1700, 8
        img = synth(h, w, 3)
        ctr, lut = _centres(img, K)
        d = dev(img)
        masks = torch.empty((K, h, w), dtype=torch.uint8, device="cuda")
        edges = torch.empty_like(masks)
        torch.cuda.synchronize()
        free0 = torch.cuda.mem_get_info()[0]
        need = e.reserve(h, w, K, ksize=5)
        torch.cuda.synchronize()
        free1 = torch.cuda.mem_get_info()[0]
        assert need > 0 and free0 - free1 >= need * 0.9          # the driver rounds allocations up, never down
        for ec in (omni_b200.EdgeConfig(), omni_b200.EdgeConfig(ksize=5)):
            e.color_edge(d, ctr, lut, ec, masks=masks, edges=edges)
            e.color_edge_packed(d[None], ctr, lut, omni_b200.EdgeConfig())
        torch.cuda.synchronize()
        # torch's caching allocator may have grown for the packed outputs; the library's own slots must not have
        free2 = torch.cuda.mem_get_info()[0]
        assert free1 - free2 <= 64 << 20, (free1, free2)
        e.color_edge(d, ctr, lut, omni_b200.EdgeConfig(), masks=masks, edges=edges)
        torch.cuda.synchronize()
        assert torch.cuda.mem_get_info()[0] == free2
    finally:
        e.close()


def test_assume_binary_masks_skips_only_the_check(eng):
    import omni_b200
    from oracle import cmodel as cm
    masks = np.stack([blob_mask(300, 500, s, p=0.4) for s in range(3)])
    want = np.stack([cm.edge_chain(m, 3, 1, 1, 3, 50, 150) for m in masks])
    ec = omni_b200.EdgeConfig()
    eng.assume_binary_masks(True)
    try:
        assert np.array_equal(host(eng.edges(dev(masks), ec)), want)
        ec5 = omni_b200.EdgeConfig(ksize=5)
        want5 = np.stack([cm.edge_chain(m, 3, 1, 1, 5, 50, 150) for m in masks])
        assert np.array_equal(host(eng.edges(dev(masks), ec5)), want5)
    finally:
        eng.assume_binary_masks(False)
    assert np.array_equal(host(eng.edges(dev(masks), ec)), want)


@pytest.mark.parametrize("hw,K,msb", [((300, 517), 4, True), ((64, 40), 3, False), ((1080, 1920), 8, True)])
def test_thinning_on_packed_planes_equals_byte_planes(eng, hw, K, msb):
    """omni_thin_zhangsuen_packed: the packed edge planes of the fused call go straight into the thinning kernel."""
    import omni_b200
    h, w = hw
    img = synth(h, w, 5)
    ctr, lut = _centres(img, K)
    d = dev(img)
    _m, e_bits = eng.color_edge_packed(d[None], ctr, lut, omni_b200.EdgeConfig(), msb_first=msb)
    _lab, _mb, eb = eng.color_edge(d, ctr, lut, omni_b200.EdgeConfig())
    want = host(eng.thin_zhangsuen(eb))
    got_bits = host(eng.thin_zhangsuen_packed(e_bits, w, msb_first=msb))
    got = np.unpackbits(got_bits, axis=2, bitorder="big" if msb else "little")[:, :, :w] * 255
    assert np.array_equal(got, want)
    # in place
    e2 = e_bits.clone()
    eng.thin_zhangsuen_packed(e2, w, msb_first=msb, out=e2)
    assert torch.equal(e2, torch.from_numpy(got_bits).cuda())


# ---- row-band pipelining of the single-image host call (omni_set_host_bands) ------------------------------------------------
def _host_packed(eng, img, ctr, lut, ec, bands, msb=True, counts=True):
    eng.set_host_bands(bands)
    try:
        r = eng.host_color_edge_packed(img, ctr, lut, ec, msb_first=msb, want_counts=counts)
        return r, eng.last_band_resends()
    finally:
        eng.set_host_bands(2)


@pytest.mark.parametrize("hw,K,msb", [((1100, 2048), 4, True), ((1500, 1999), 5, False), ((2051, 1025), 16, True), ((4096, 4096), 16, True)])
def test_banded_host_call_equals_unbanded(eng, hw, K, msb):
    """One image through omni_host_color_edge_packed: bands off / edge planes after the last band / edge rows with their band give
    the same bytes and counts as the device-resident call (gap-free rows, odd widths, pitches that are no multiple of 16)."""
    import omni_b200
    h, w = hw
    img = synth(h, w, K + 3, cell=32)
    ctr, lut = _centres(img[: min(h, 1024)], K)
    ec = omni_b200.EdgeConfig()
    mb0, eb0, c0 = eng.color_edge_packed(dev(img), ctr, lut, ec, msb_first=msb, want_counts=True)
    mb0, eb0 = host(mb0), host(eb0)
    for bands in (0, 1, 2):
        r, resends = _host_packed(eng, img, ctr, lut, ec, bands, msb)
        assert (resends >= 0) == (bands > 0), (bands, resends)
        assert np.array_equal(r["mask_bits"], mb0), bands
        assert np.array_equal(r["edge_bits"], eb0), bands
        assert np.array_equal(r["counts"], c0), bands
    r, resends = _host_packed(eng, img, ctr, lut, None, 2, msb)           # colour layers only
    assert resends == 0 and r["edge_bits"] is None and np.array_equal(r["mask_bits"], mb0)


def test_banded_host_call_against_the_oracle(eng):
    """The banded call against the CPU oracle directly (not only against the unbanded GPU path)."""
    import omni_b200
    rp = _rp()
    h, w, K = 1300, 1700, 4
    img = synth(h, w, 11, cell=48)
    ctr, lut = _centres(img, K)
    r, resends = _host_packed(eng, img, ctr, lut, omni_b200.EdgeConfig(), 2)
    assert resends >= 0
    _c, labels, masks = rp.color_extract(img, K, ctr)
    for k in range(K):
        assert np.array_equal(np.unpackbits(r["mask_bits"][k], axis=1)[:, :w] * 255, masks[k]), k
        assert np.array_equal(np.unpackbits(r["edge_bits"][k], axis=1)[:, :w] * 255, rp.edge_layer(masks[k])), k
    assert np.array_equal(r["counts"][:, 0], np.bincount(labels.ravel(), minlength=K)[:K])


def test_banded_host_call_weak_chains_across_bands(eng):
    """Thresholds that leave most straight edges weak and only corners strong: weak chains run across band borders, so a later band
    promotes pixels of rows that have left already -- those bands must be sent again (and with more than 8192 weak words the
    worklist overflows and everything is resent).  Bytes must equal the unbanded call whatever happened."""
    import omni_b200
    seen = []
    for (h, w, K, cell) in ((1100, 2048, 3, 96), (4096, 4096, 8, 64)):
        img = synth(h, w, 5, cell=cell)
        ctr, lut = _centres(img[:1024], K)
        for high in (400.0, 700.0, 800.0, 900.0, 1000.0):
            ec = omni_b200.EdgeConfig(low=50.0, high=high)
            r0, _ = _host_packed(eng, img, ctr, lut, ec, 0, counts=True)
            for bands in (1, 2):
                r, resends = _host_packed(eng, img, ctr, lut, ec, bands, counts=True)
                assert np.array_equal(r["mask_bits"], r0["mask_bits"]), (h, high, bands)
                assert np.array_equal(r["edge_bits"], r0["edge_bits"]), (h, high, bands)
                assert np.array_equal(r["counts"], r0["counts"]), (h, high, bands)
                if bands == 2:
                    seen.append(resends)
    print("band resends:", seen)
    assert max(seen) > 0, seen          # the resend path was exercised


def test_banded_host_call_leaves_row_gaps_alone(eng):
    """Host planes whose rows are wider than ceil(w/8) (a view into a larger array): bytes outside the view stay untouched."""
    import ctypes as C
    import omni_b200
    from omni_b200 import capi
    h, w, K = 1200, 2000, 3
    rb = (w + 7) // 8
    img = synth(h, w, 2, cell=64)
    ctr, lut = _centres(img, K)
    ec = omni_b200.EdgeConfig()
    want, _ = _host_packed(eng, img, ctr, lut, ec, 0)
    for gap in (6, 22):
        big_m = np.full((K, h, rb + gap), 0xA5, np.uint8)
        big_e = np.full((K, h, rb + gap), 0x5A, np.uint8)
        p = ec.to_c()
        ctr32 = np.ascontiguousarray(ctr, np.float32)
        capi.check(eng._L.omni_host_color_edge_packed(
            eng._h, img.ctypes.data, 1, img.strides[0] * h, h, w, img.strides[0], ctr32.ctypes.data_as(C.POINTER(C.c_float)), K,
            lut.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(p), big_m.ctypes.data, big_m.strides[0], big_m.strides[1],
            big_e.ctypes.data, big_e.strides[0], big_e.strides[1], capi.BITS_MSB_FIRST, None))
        assert eng.last_band_resends() >= 0
        assert np.array_equal(big_m[:, :, :rb], want["mask_bits"]) and np.array_equal(big_e[:, :, :rb], want["edge_bits"])
        assert (big_m[:, :, rb:] == 0xA5).all() and (big_e[:, :, rb:] == 0x5A).all()


@pytest.mark.parametrize("seed", range(8))
def test_banded_host_call_randomised(eng, seed):
    """Random geometry (odd widths, heights that leave a short last band), K, thresholds, morphology switches, bit order: the
    banded call (both modes) against the unbanded one, bytes and counts."""
    import omni_b200
    rng = np.random.default_rng(1000 + seed)
    h = int(rng.integers(1024, 2600))
    w = int(rng.integers(2 * 1024 * 1024 // h + 1, 2600))
    K = int(rng.integers(2, 17))
    img = synth(h, w, seed, cell=int(rng.choice([16, 32, 64, 128])))
    ctr, lut = _centres(img[:768, :768], K)
    low, high = sorted(float(v) for v in rng.integers(0, 900, 2))
    ec = omni_b200.EdgeConfig(low=low, high=high, open_iters=int(rng.integers(0, 2)), close_iters=int(rng.integers(0, 2)))
    msb = bool(rng.integers(0, 2))
    r0, _ = _host_packed(eng, img, ctr, lut, ec, 0, msb)
    for bands in (2, 1):
        r, resends = _host_packed(eng, img, ctr, lut, ec, bands, msb)
        assert resends >= 0, (h, w, K)
        assert np.array_equal(r["mask_bits"], r0["mask_bits"]), (h, w, K, bands)
        assert np.array_equal(r["edge_bits"], r0["edge_bits"]), (h, w, K, bands, low, high, resends)
        assert np.array_equal(r["counts"], r0["counts"]), (h, w, K, bands)
