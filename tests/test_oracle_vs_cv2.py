"""Pin the plain-C oracle against the library calls the reference makes (cv2 4.13 / NumPy), on
seeded random and structured inputs incl. degenerate shapes.  CPU only."""
import cv2
import numpy as np
import pytest

from oracle import cmodel as cm
from oracle import refport as rp
from helpers import synth, uniform_img, smooth_u8, blob_mask


@pytest.fixture(autouse=True, scope="module")
def _single_thread():
    # cv2.GaussianBlur ks=5 mis-computes tiny images when rows ~ thread count (caveat i)
    n = cv2.getNumThreads()
    cv2.setNumThreads(1)
    yield
    cv2.setNumThreads(n)


def test_structuring_elements():
    for shape in (cv2.MORPH_RECT, cv2.MORPH_ELLIPSE):
        for k in range(1, 16):
            assert np.array_equal(cm.structuring_element(shape, k), cv2.getStructuringElement(shape, (k, k)))


@pytest.mark.parametrize("hw", [(67, 93), (1, 50), (50, 1), (2, 2), (128, 200)])
def test_morphology(hw):
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    for img in ((rng.random(hw) < 0.6).astype(np.uint8) * 255, rng.integers(0, 256, hw, dtype=np.uint8)):
        for shape in (cv2.MORPH_RECT, cv2.MORPH_ELLIPSE):
            for k in (1, 2, 3, 4, 5, 7):
                se = cv2.getStructuringElement(shape, (k, k))
                for op, cop in ((0, cv2.MORPH_OPEN), (1, cv2.MORPH_CLOSE)):
                    for it in (1, 2, 3):
                        assert np.array_equal(cm.morph(img, se, op, it),
                                              cv2.morphologyEx(img, cop, se, iterations=it)), (shape, k, op, it)


@pytest.mark.parametrize("hw", [(67, 93), (128, 200), (64, 5), (5, 64), (64, 1), (1, 64), (3, 3), (64, 2), (2051, 307)])
def test_gaussian_blur(hw):
    rng = np.random.default_rng(hw[0] + hw[1])
    for img in (rng.integers(0, 256, hw, dtype=np.uint8), (rng.random(hw) < 0.5).astype(np.uint8) * 255):
        for k in (3, 5, 7, 9, 13, 21, 31):
            assert np.array_equal(cm.gaussian_blur(img, k), cv2.GaussianBlur(img, (k, k), 0)), k


def test_canny_random():
    rng = np.random.default_rng(5)
    for t in range(150):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = t % 3
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = smooth_u8(h, w, t)
        else:
            img = cv2.GaussianBlur(blob_mask(h + 8, w + 8, t), (3, 3), 0)[:h, :w].copy()
        for lo, hi in [(50, 150), (22, 70), (100, 200), (50, 50), (200, 50), (30.7, 90.2), (0, 0)]:
            assert np.array_equal(cm.canny(img, lo, hi), cv2.Canny(img, lo, hi)), (h, w, kind, lo, hi)


def test_canny_large_long_chains():
    img = smooth_u8(700, 900, 3)
    for lo, hi in [(10, 60), (5, 200)]:
        assert np.array_equal(cm.canny(img, lo, hi), cv2.Canny(img, lo, hi))


def test_bgr2lab_and_assign():
    rng = np.random.default_rng(11)
    for img in (synth(256, 256, 0), uniform_img(200, 300, 1)):
        lab = cm.bgr2lab(img)
        assert np.array_equal(lab, cv2.cvtColor(img, cv2.COLOR_BGR2LAB))
        for K in (2, 3, 4, 8, 16):
            ctr = rp.kmeans_lab_centers(img, K)
            assert np.array_equal(cm.assign_f32(lab, ctr), rp.assign_lab(img, ctr))
            # adversarial near-ties: half-integer centres, duplicated centres
            ctr2 = (rng.integers(0, 255, (K, 3)) + 0.5).astype(np.float32)
            ctr2[K - 1] = ctr2[0]
            assert np.array_equal(cm.assign_f32(lab, ctr2), rp.assign_lab(img, ctr2))
            pal = rng.integers(0, 256, (K, 3), dtype=np.uint8)
            assert np.array_equal(cm.assign_i16wrap(img, pal), rp.assign_labels_rgb(img, pal))


@pytest.mark.parametrize("hs,ws,md", [(300, 200, 133), (512, 512, 256), (768, 384, 256), (409, 409, 200),
                                      (1000, 1500, 700), (640, 480, 160), (1080, 1920, 1000), (33, 1000, 77)])
def test_resize_area(hs, ws, md):
    src = synth(hs, ws, hs + ws, cell=16)
    nw, nh = rp.resize_dims(hs, ws, md)
    assert np.array_equal(cm.resize_area(src, nw, nh), cv2.resize(src, (nw, nh), interpolation=cv2.INTER_AREA))


def test_chains():
    img = synth(320, 448, 3)
    for K in (4, 8):
        ctr = rp.kmeans_lab_centers(img, K)
        cs, ls, ms = rp.color_extract(img, K, ctr)
        order, lut = rp.darkness_order(ctr)
        lbl = cm.assign_f32(cm.bgr2lab(img), ctr)
        assert np.array_equal(cm.layer_masks(lbl, K, lut.astype(np.uint8)), ms)
        for ks, lo, hi in [(3, 50, 150), (7, 22, 70), (5, 100, 200)]:
            for m in ms:
                assert np.array_equal(cm.edge_chain(m, 3, 1, 1, ks, lo, hi), rp.edge_layer(m, lo, hi, ks))
        m = ms[0]
        for mk, oi, ci in [(1, 1, 1), (5, 1, 1), (3, 2, 0), (3, 0, 2), (2, 1, 1)]:
            assert np.array_equal(cm.edge_chain(m, mk, oi, ci, 3, 50, 150), rp.edge_layer(m, 50, 150, 3, mk, oi, ci))
