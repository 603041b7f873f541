"""Stage 04: the mirror of the reference's trace_centerlines (04_find_contours.py:101-205) must return the
reference's polylines in the reference's order.  Golden: tests/golden/trace.npz, frozen from the unmodified reference by
tools/make_golden_trace.py.  CPU: the degree maps come from cv2 (what the reference uses); GPU: from omni_skeleton_degree."""
import contextlib
import io
import re

import cv2
import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE_DIR

Z = np.load(f"{GOLDEN}/trace.npz")
CASES = sorted(k[:-5] for k in Z.files if k.endswith("_skel"))


def _blank(log):
    return "\n".join(ln for ln in re.sub(r"[0-9.]+s", "Xs", log).splitlines() if "visited" not in ln)


def _maps_cv2(sk):
    S = (sk > 0).astype(np.uint8)
    k = np.ones((3, 3), np.uint8)
    k[1, 1] = 0
    deg = cv2.filter2D(S, cv2.CV_8U, k, borderType=cv2.BORDER_CONSTANT)
    return deg, (S == 1) & (deg == 1), (S == 1) & (deg >= 3)


def _check(case, maps):
    from omni_b200 import contours
    sk = Z[case + "_skel"]
    with contextlib.redirect_stdout(io.StringIO()) as log:
        paths = contours.trace_centerlines(sk.copy(), case, maps=maps(sk) if maps else None)
    assert [len(p) for p in paths] == Z[case + "_len"].tolist()
    assert all(p.dtype == np.int32 and p.shape[1:] == (1, 2) for p in paths)
    pts = np.concatenate([p.reshape(-1, 2) for p in paths]) if paths else np.zeros((0, 2), np.int32)
    assert np.array_equal(pts, Z[case + "_pts"])
    assert _blank(log.getvalue()) == str(Z[case + "_log"])


@pytest.mark.parametrize("case", CASES)
def test_trace_mirror_equals_reference_golden(case):
    _check(case, _maps_cv2)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_trace_mirror_with_gpu_degree_maps(case):
    _check(case, None)


@pytest.mark.reference
def test_trace_mirror_equals_reference_function():
    """Directly against the imported reference on skeletons outside the golden set (build container only)."""
    import importlib.util
    import os
    import sys
    from helpers import blob_mask
    from oracle import cmodel as cm
    from omni_b200 import contours
    sys.dont_write_bytecode = True
    if REFERENCE_DIR not in sys.path:
        sys.path.append(REFERENCE_DIR)
    spec = importlib.util.spec_from_file_location("ref04", os.path.join(REFERENCE_DIR, "04_find_contours.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for seed, (h, w) in enumerate([(90, 130), (64, 64), (37, 41)]):
        e = cv2.Canny(cv2.GaussianBlur(blob_mask(h, w, 50 + seed, 0.4, k=7), (3, 3), 0), 50, 150)
        sk = cm.thin_zhangsuen(e)
        thick = (np.random.default_rng(seed).random((h, w)) < 0.35).astype(np.uint8) * 255
        for S in (sk, thick):
            with contextlib.redirect_stdout(io.StringIO()) as la:
                a = ref.trace_centerlines(S.copy(), "L")
            with contextlib.redirect_stdout(io.StringIO()) as lb:
                b = contours.trace_centerlines(S.copy(), "L", maps=_maps_cv2(S))
            assert len(a) == len(b) and all(x.dtype == y.dtype and np.array_equal(x, y) for x, y in zip(a, b))
            assert _blank(la.getvalue()) == _blank(lb.getvalue())
