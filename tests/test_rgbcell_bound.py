"""Host-side proof obligations of the RGB-cell colour assignment (csrc/fast_kernels.cu fk_build_rgbcells):
the Lab box a cell is given must contain the Lab value (cv2 8-bit BGR2LAB arithmetic) of EVERY colour of the cell.
Checked exhaustively over all 2^24 colours with the tables the kernels are compiled with.  CPU only."""
import os
import re

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


def test_tables_monotone_and_shared():
    gam, cb = _tables(CSRC_TAB)
    gam2, cb2 = _tables(ORC_TAB)
    assert np.array_equal(gam, gam2) and np.array_equal(cb, cb2)       # kernels and oracle use the same tables
    assert len(gam) == 256 and (np.diff(gam) >= 0).all()
    assert (np.diff(cb) >= 0).all()
    # the largest index the XYZ sums can reach stays inside the cube-root table
    g = int(gam[255])
    assert (g * (1777 + 1541 + 778) + 2048) >> 12 < len(cb) and (g * (73 + 448 + 3575) + 2048) >> 12 < len(cb)


def _f(gam, cb, B8, G8, R8):
    B, G, R = gam[B8], gam[G8], gam[R8]
    return (cb[(R * 1777 + G * 1541 + B * 778 + 2048) >> 12], cb[(R * 871 + G * 2929 + B * 296 + 2048) >> 12],
            cb[(R * 73 + G * 448 + B * 3575 + 2048) >> 12])


def _lab(fX, fY, fZ):
    return ((296 * fY - 1336934 + 16384) >> 15, (500 * (fX - fY) + 128 * 32768 + 16384) >> 15,
            (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15)


def test_cell_boxes_contain_every_colour():
    gam, cb = _tables(CSRC_TAB)
    s = 4
    g = np.arange(0, 256, s)
    B0, G0, R0 = np.meshgrid(g, g, g, indexing="ij")
    xl, yl, zl = _f(gam, cb, B0, G0, R0)
    xh, yh, zh = _f(gam, cb, B0 + s - 1, G0 + s - 1, R0 + s - 1)
    Llo, alo, blo = _lab(xl, yl, zl)[0], _lab(xl, yh, zl)[1], _lab(xl, yl, zh)[2]
    Lhi, ahi, bhi = _lab(xh, yh, zh)[0], _lab(xh, yl, zh)[1], _lab(xh, yh, zl)[2]
    worst = 0
    for db in range(s):
        for dg in range(s):
            for dr in range(s):
                L, a, b = _lab(*_f(gam, cb, B0 + db, G0 + dg, R0 + dr))
                assert (L >= Llo).all() and (L <= Lhi).all()
                assert (a >= alo).all() and (a <= ahi).all()
                assert (b >= blo).all() and (b <= bhi).all()
                worst = max(worst, int((ahi - alo).max()), int((bhi - blo).max()))
    assert worst < 80           # the boxes stay small enough to prune
