"""Host-side proof obligations of the RGB-cell colour assignment (csrc/fast_kernels.cu fk_rgb_boxes / fk_build_rgbcells).

The kernels bound the Lab image of every 4x4x4 RGB cell by the exact min / max of its 64 colours, computed with the integer
Lab arithmetic and tables they are compiled with, and store the bounds as bytes.  That is sound iff (1) this arithmetic IS
cv2's 8-bit BGR2LAB for every colour and (2) its results fit a byte without the final saturate_cast (lab_noclamp).  Both
are checked exhaustively over all 2^24 colours.  CPU only."""
import os
import re

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC_TAB = os.path.join(ROOT, "omnirevolve-image-processor_b200", "csrc", "omni_tables.inc")
ORC_TAB = os.path.join(ROOT, "oracle", "omni_tables.inc")


def _tab(txt, name):
    m = re.search(name + r"\[\d+\] = \{(.*?)\};", txt, re.S)
    return np.array([int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()], np.int64)


def _tables(path):
    txt = open(path).read()
    return _tab(txt, "OMNI_LAB_GAMMA"), _tab(txt, "OMNI_LAB_CBRT")


def _lab_noclamp(gam, cb, B8, G8, R8):
    B, G, R = gam[B8], gam[G8], gam[R8]
    fX, fY, fZ = (cb[(R * 1777 + G * 1541 + B * 778 + 2048) >> 12], cb[(R * 871 + G * 2929 + B * 296 + 2048) >> 12],
                  cb[(R * 73 + G * 448 + B * 3575 + 2048) >> 12])
    return ((296 * fY - 1336934 + 16384) >> 15, (500 * (fX - fY) + 128 * 32768 + 16384) >> 15,
            (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15)


def test_tables_shared_and_in_range():
    gam, cb = _tables(CSRC_TAB)
    gam2, cb2 = _tables(ORC_TAB)
    assert np.array_equal(gam, gam2) and np.array_equal(cb, cb2)       # kernels and oracle use the same tables
    assert len(gam) == 256
    # the largest index the XYZ sums can reach stays inside the cube-root table
    g = int(gam.max())
    assert (g * (1777 + 1541 + 778) + 2048) >> 12 < len(cb) and (g * (73 + 448 + 3575) + 2048) >> 12 < len(cb)


def test_integer_lab_is_cv2_for_all_colours_and_fits_a_byte():
    gam, cb = _tables(CSRC_TAB)
    g = np.arange(256)
    for b8 in range(0, 256, 8):                                   # 32 slabs of 8 x 256 x 256 colours
        B8, G8, R8 = np.meshgrid(np.arange(b8, b8 + 8), g, g, indexing="ij")
        L, a, b = _lab_noclamp(gam, cb, B8, G8, R8)
        mine = np.stack([L, a, b], -1)
        assert mine.min() >= 0 and mine.max() <= 255              # lab_noclamp: no saturate_cast needed, boxes fit u8
        img = np.stack([B8, G8, R8], -1).astype(np.uint8).reshape(-1, 256, 3)
        want = cv2.cvtColor(img, cv2.COLOR_BGR2LAB).reshape(mine.shape)
        assert np.array_equal(mine, want), b8


def test_exact_cell_boxes_prune_well():
    """The exact boxes are ~2.7 Lab units wide per channel (the conservative corner bound they replaced was ~9 wide in a)."""
    gam, cb = _tables(CSRC_TAB)
    s = 4
    g = np.arange(0, 256, s)
    B0, G0, R0 = np.meshgrid(g, g, g, indexing="ij")
    lo = np.full(B0.shape + (3,), 999, np.int64)
    hi = np.full(B0.shape + (3,), -1, np.int64)
    for db in range(s):
        for dg in range(s):
            for dr in range(s):
                v = np.stack(_lab_noclamp(gam, cb, B0 + db, G0 + dg, R0 + dr), -1)
                lo = np.minimum(lo, v)
                hi = np.maximum(hi, v)
    width = (hi - lo).reshape(-1, 3)
    assert width.mean(0).max() < 3.5 and width.max() < 40
