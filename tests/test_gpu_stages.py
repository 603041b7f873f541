"""Stage-script level parity: the drop-in 01/02/03 scripts, run the way pipeline.py runs stages (one subprocess per
stage, CONFIG_PATH in the environment, files in output_dir), must produce the files the unmodified reference
produced for the same input and config (golden fixtures).  Needs a GPU."""
import json
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest

from helpers import PIPE_CASES, load_pipe_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "omnirevolve-image-processor_b200", "image_processor")


def run_stage(script, cfg_path):
    env = dict(os.environ, CONFIG_PATH=cfg_path, PYTHONUNBUFFERED="1")
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, script)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


@pytest.mark.parametrize("case", PIPE_CASES)
def test_stage_scripts_match_reference_files(case, tmp_path):
    z, meta = load_pipe_case(case)
    cfg = dict(meta["config"])
    names = meta["names"]
    src = tmp_path / "input.png"
    cv2.imwrite(str(src), z["input"])
    out = tmp_path / "out"
    out.mkdir()
    cfg.update(input_image=str(src), output_dir=str(out))
    cfg_path = out / "config.json"
    cfg_path.write_text(json.dumps(cfg))
    log = run_stage("01_resize.py", str(cfg_path))
    assert ("Resizing:" in log) or ("No resize required" in log)
    resized = cv2.imread(str(out / "resized.png"), cv2.IMREAD_COLOR)
    assert resized.shape == z["resized"].shape
    assert np.abs(resized.astype(np.int16) - z["resized"].astype(np.int16)).max() <= 1
    # stages 02/03 of the reference ran on ITS resized.png: give ours the identical input
    cv2.imwrite(str(out / "resized.png"), z["resized"])
    log = run_stage("02_color_extract.py", str(cfg_path))
    assert "Color extraction: done." in log
    pal = json.load(open(out / "palette_by_name.json"))
    assert pal == meta["palette_by_name"]
    for i, n in enumerate(names):
        m = cv2.imread(str(out / n / "mask.png"), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(m, z["masks"][i]), n
    log = run_stage("03_edge_detect.py", str(cfg_path))
    for i, n in enumerate(names):
        e = cv2.imread(str(out / n / "edges.png"), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(e, z["edges"][i]), n
        assert f"Edges extracted: {n} | nz={int(np.count_nonzero(z['edges'][i]))}" in log
    comp = cv2.imread(str(out / "edges_composite.png"), cv2.IMREAD_COLOR)
    assert np.array_equal(comp, z["composite"])
    # stage 03 again, now on its own (`--start-step 3`): nothing parked by stage 02 is left, it recomputes from mask.png
    assert not (out / ".omni_b200_handoff.json").exists()
    for n in names:
        os.remove(out / n / "edges.png")
    os.remove(out / "edges_composite.png")
    log2 = run_stage("03_edge_detect.py", str(cfg_path))
    for i, n in enumerate(names):
        e = cv2.imread(str(out / n / "edges.png"), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(e, z["edges"][i]), n
        assert f"Edges extracted: {n} | nz={int(np.count_nonzero(z['edges'][i]))}" in log2
    assert np.array_equal(cv2.imread(str(out / "edges_composite.png"), cv2.IMREAD_COLOR), z["composite"])


@pytest.mark.parametrize("case", PIPE_CASES)
def test_front_half_one_process_matches_reference_files(case, tmp_path):
    """front_half.py = stages 01-03 in one process (in-memory hand-off): the files of the unmodified reference's three stages."""
    z, meta = load_pipe_case(case)
    cfg = dict(meta["config"])
    names = meta["names"]
    src = tmp_path / "input.png"
    cv2.imwrite(str(src), z["input"])
    out = tmp_path / "out"
    out.mkdir()
    cfg.update(input_image=str(src), output_dir=str(out))
    cfg_path = out / "config.json"
    cfg_path.write_text(json.dumps(cfg))
    log = run_stage("front_half.py", str(cfg_path))
    resized = cv2.imread(str(out / "resized.png"), cv2.IMREAD_COLOR)
    assert resized.shape == z["resized"].shape
    assert np.abs(resized.astype(np.int16) - z["resized"].astype(np.int16)).max() <= 1
    assert "Color extraction: done." in log and "Edges composite saved:" in log and "Saved:" in log
    if np.array_equal(resized, z["resized"]):          # (a fractional resize may differ by 1 LSB: the later stages then see another image)
        assert json.load(open(out / "palette_by_name.json")) == meta["palette_by_name"]
        for i, n in enumerate(names):
            assert np.array_equal(cv2.imread(str(out / n / "mask.png"), cv2.IMREAD_GRAYSCALE), z["masks"][i]), n
            assert np.array_equal(cv2.imread(str(out / n / "edges.png"), cv2.IMREAD_GRAYSCALE), z["edges"][i]), n
            assert f"Edges extracted: {n} | nz={int(np.count_nonzero(z['edges'][i]))}" in log
        assert np.array_equal(cv2.imread(str(out / "edges_composite.png"), cv2.IMREAD_COLOR), z["composite"])
    else:
        pytest.fail("resized.png differs from the reference's: expected identical bytes on the golden cases")


def test_handoff_is_dropped_when_masks_or_keys_change(tmp_path):
    """The planes stage 02 parks for stage 03 are used only for unchanged masks and edge keys: a hand-edited mask.png or a changed
    threshold makes stage 03 recompute from the files (what the reference would do)."""
    z, meta = load_pipe_case(PIPE_CASES[0])
    cfg = dict(meta["config"])
    names = meta["names"]
    out = tmp_path / "out"
    out.mkdir()
    cv2.imwrite(str(out / "resized.png"), z["resized"])
    cfg.update(input_image=str(tmp_path / "unused.png"), output_dir=str(out))
    cfg_path = out / "config.json"
    cfg_path.write_text(json.dumps(cfg))
    run_stage("02_color_extract.py", str(cfg_path))
    assert (out / ".omni_b200_handoff.json").exists()
    # (a) changed edge key
    cfg2 = dict(cfg, edge_low_threshold=100, edge_high_threshold=200)
    cfg_path.write_text(json.dumps(cfg2))
    run_stage("03_edge_detect.py", str(cfg_path))
    from oracle import refport as rp
    for i, n in enumerate(names):
        want = rp.edge_layer(z["masks"][i], 100, 200, cfg["edge_kernel_size"], cfg["edge_morph_kernel"], cfg["edge_morph_open_iters"],
                             cfg["edge_morph_close_iters"])
        assert np.array_equal(cv2.imread(str(out / n / "edges.png"), cv2.IMREAD_GRAYSCALE), want), n
    # (b) hand-edited mask
    cfg_path.write_text(json.dumps(cfg))
    run_stage("02_color_extract.py", str(cfg_path))
    m = z["masks"][0].copy()
    m[10:40, 10:60] = 255
    cv2.imwrite(str(out / names[0] / "mask.png"), m)
    run_stage("03_edge_detect.py", str(cfg_path))
    want = rp.edge_layer(m, cfg["edge_low_threshold"], cfg["edge_high_threshold"], cfg["edge_kernel_size"], cfg["edge_morph_kernel"],
                         cfg["edge_morph_open_iters"], cfg["edge_morph_close_iters"])
    assert np.array_equal(cv2.imread(str(out / names[0] / "edges.png"), cv2.IMREAD_GRAYSCALE), want)


def test_stage_error_behaviour(tmp_path):
    """Same exception types / non-zero exit as the reference: unreadable input (01), missing resized.png (02),
    missing mask (03)."""
    out = tmp_path / "out"
    out.mkdir()
    cfg = {"input_image": str(tmp_path / "nope.png"), "output_dir": str(out)}
    (out / "config.json").write_text(json.dumps(cfg))
    env = dict(os.environ, CONFIG_PATH=str(out / "config.json"))
    for script, needle in (("01_resize.py", "ValueError"), ("02_color_extract.py", "RuntimeError"),
                           ("03_edge_detect.py", "FileNotFoundError")):
        r = subprocess.run([sys.executable, os.path.join(SCRIPTS, script)], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode != 0 and needle in r.stdout, (script, r.stdout[-500:])


def test_assign_labels_dropin():
    sys.path.insert(0, SCRIPTS)
    import importlib.util
    spec = importlib.util.spec_from_file_location("dropin_process_colors", os.path.join(SCRIPTS, "process_colors.py"))
    pc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pc)
    from helpers import GOLDEN
    z = np.load(f"{GOLDEN}/functions.npz")
    for K in (2, 4, 8, 16):
        assert np.array_equal(pc.assign_labels(z["al_img"], z[f"al_pal{K}"]), z[f"al_lab{K}"])


def test_stage04_shim_swaps_thinning_and_tracing(tmp_path):
    """The stage-04 shim loads `04_find_contours_ref.py` (the reference's renamed file -- here a stand-in with the same
    surface), replaces `thinning_zhangsuen` and `trace_centerlines` by the GPU-backed mirrors and runs the original
    `vectorize_all`."""
    import shutil
    import subprocess
    import sys
    import cv2
    import numpy as np
    from helpers import blob_mask
    from oracle import cmodel as cm
    src = os.path.join(ROOT, "omnirevolve-image-processor_b200", "image_processor")
    for f in ("04_find_contours.py", "_omni_path.py"):
        shutil.copy(os.path.join(src, f), tmp_path / f)
    (tmp_path / "04_find_contours_ref.py").write_text(
        "import os, cv2, numpy as np\n"
        "def load_config():\n"
        "    class C: pass\n"
        "    c = C(); c.output_dir = os.environ['OUT_DIR']; c.color_names = ['layer_a', 'layer_b']; return c\n"
        "def thinning_zhangsuen(img, layer):\n"
        "    raise RuntimeError('the CPU thinning must have been replaced')\n"
        "def trace_centerlines(skel, layer):\n"
        "    raise RuntimeError('the per-component tracing must have been replaced')\n"
        "def vectorize_layer(name, cfg):\n"
        "    import pickle\n"
        "    e = cv2.imread(os.path.join(cfg.output_dir, name, 'edges.png'), cv2.IMREAD_GRAYSCALE)\n"
        "    sk = thinning_zhangsuen(e, layer=name)\n"
        "    np.save(os.path.join(cfg.output_dir, name, 'skel.npy'), sk)\n"
        "    pickle.dump(trace_centerlines(sk, layer=name), open(os.path.join(cfg.output_dir, name, 'paths.pkl'), 'wb')); return name, []\n"
        "def vectorize_all(cfg):\n"
        "    return dict(vectorize_layer(n, cfg) for n in cfg.color_names)\n")
    out = tmp_path / "out"
    want = {}
    for i, n in enumerate(("layer_a", "layer_b")):
        os.makedirs(out / n)
        e = cv2.Canny(cv2.GaussianBlur(blob_mask(120, 160, 30 + i, 0.4, k=7), (3, 3), 0), 50, 150)
        cv2.imwrite(str(out / n / "edges.png"), e)
        want[n] = cm.thin_zhangsuen(e)
    env = dict(os.environ, OUT_DIR=str(out), OMNI_B200_HOME=os.path.join(ROOT, "omnirevolve-image-processor_b200"))
    r = subprocess.run([sys.executable, str(tmp_path / "04_find_contours.py")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "[layer_a] Thinning ROI" in r.stdout and "Thinning done" in r.stdout
    import contextlib
    import io
    import pickle
    from omni_b200 import contours
    for n in want:
        assert np.array_equal(np.load(out / n / "skel.npy"), want[n]), n
        S = (want[n] > 0).astype(np.uint8)
        k = np.ones((3, 3), np.uint8); k[1, 1] = 0
        deg = cv2.filter2D(S, cv2.CV_8U, k, borderType=cv2.BORDER_CONSTANT)
        with contextlib.redirect_stdout(io.StringIO()):
            ref_paths = contours.trace_centerlines(want[n], n, maps=(deg, (S == 1) & (deg == 1), (S == 1) & (deg >= 3)))
        got = pickle.load(open(out / n / "paths.pkl", "rb"))
        assert len(got) == len(ref_paths) > 0 and all(np.array_equal(a, b) for a, b in zip(got, ref_paths)), n
        assert f"[{n}] Trace done: {len(got)} polylines" in r.stdout
