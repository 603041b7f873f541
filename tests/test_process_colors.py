"""process_colors.py drop-in: the host-side mirror (palette JSON readers, k-means palette, names) and the
oracle's int16-wrap assignment against what the UNMODIFIED reference CLI wrote (tests/golden/process_colors.*, frozen by
tools/make_golden_process_colors.py); on the GPU box the drop-in script itself, run as the reference is run."""
import json
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest

from conftest import GOLDEN, PKG, REFERENCE_DIR
from oracle import cmodel as cm

SCRIPT = os.path.join(PKG, "image_processor", "process_colors.py")
META = json.load(open(os.path.join(GOLDEN, "process_colors.json")))
CASES = ["adaptive5", "analyzer", "eight"]


def _z():
    return np.load(os.path.join(GOLDEN, "process_colors.npz"))


def _palette(case):
    cols = META[case]["palette"]["colors"]
    return np.array([c["rgb"] for c in cols], np.uint8), [c["name"] for c in cols]


@pytest.mark.parametrize("case", CASES)
def test_oracle_assign_equals_reference_cli(case):
    """process_colors.py:69-77 with the int16 wrap: the C oracle reproduces the label map the reference CLI saved."""
    z = _z()
    rgb = cv2.cvtColor(z["input"], cv2.COLOR_BGR2RGB)
    pal, _names = _palette(case)
    assert np.array_equal(cm.assign_i16wrap(rgb, pal), z[f"labels_{case}"])


def test_palette_json_readers(tmp_path):
    from omni_b200 import colors_cli
    for key, case in (("analyzer_json", "analyzer"), ("eight_json", "eight")):
        p = tmp_path / f"{key}.json"
        p.write_text(json.dumps(META[key]))
        rgb, names = colors_cli.palette_from_json(str(p))
        want_rgb, want_names = _palette(case)
        assert rgb.dtype == np.uint8 and np.array_equal(rgb, want_rgb) and names == want_names
    # the {"palette": [...]} form: unnamed entries get color_<i>
    p = tmp_path / "generic.json"
    p.write_text(json.dumps(META["generic_json"]))
    rgb, names = colors_cli.palette_from_json(str(p))
    assert rgb.tolist() == [e["rgb"] for e in META["generic_json"]["palette"]] and names == ["k", "color_1", "r"]
    p.write_text(json.dumps({"colours": []}))
    with pytest.raises(ValueError, match="Unsupported palette JSON structure"):
        colors_cli.palette_from_json(str(p))
    assert colors_cli.default_color_names(6) == ["red", "green", "blue", "black", "color_4", "color_5"]


def test_kmeans_palette_equals_reference_cli(tmp_path):
    """cv2.kmeans draws from OpenCV's global RNG: a fresh process, as the CLI is, gives the reference's palette."""
    z = _z()
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), z["input"])
    code = ("import sys, json; sys.path.insert(0, %r); from omni_b200 import colors_cli as c; "
            "print(json.dumps(c.kmeans_palette(c.load_image_rgb(%r), k=5).tolist()))" % (PKG, str(src)))
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1]) == _palette("adaptive5")[0].tolist()


@pytest.mark.reference
def test_reference_generic_form_raises(tmp_path):
    """process_colors.py:60-64: the second JSON form dies on line 63 under Python 3 (it reads the comprehension variable of line 62);
    the mirror implements what the line evidently means (test_palette_json_readers) -- recorded in INTEGRATION.md."""
    import importlib.util
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_pc", os.path.join(REFERENCE_DIR, "process_colors.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    p = tmp_path / "generic.json"
    p.write_text(json.dumps(META["generic_json"]))
    with pytest.raises(NameError):                       # UnboundLocalError (3.12+) is a NameError
        ref.palette_from_json(str(p))
    p2 = tmp_path / "analyzer.json"
    p2.write_text(json.dumps(META["analyzer_json"]))
    from omni_b200 import colors_cli
    a, b = ref.palette_from_json(str(p2)), colors_cli.palette_from_json(str(p2))
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert ref.default_color_names(7) == colors_cli.default_color_names(7)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_dropin_cli_matches_reference_files(case, tmp_path):
    z = _z()
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), z["input"])
    (tmp_path / "analyzer.json").write_text(json.dumps(META["analyzer_json"]))
    (tmp_path / "eight.json").write_text(json.dumps(META["eight_json"]))
    out = tmp_path / "out"
    args = [a.replace("@TD@", str(tmp_path)) for a in META[case]["args"]]
    r = subprocess.run([sys.executable, SCRIPT, str(src), "-o", str(out)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    want = z[f"labels_{case}"]
    assert np.array_equal(np.load(out / "labels.npy"), want)
    assert np.array_equal(cv2.imread(str(out / "labels.png"), cv2.IMREAD_UNCHANGED), want)
    assert json.load(open(out / "palette.json")) == META[case]["palette"]
    assert sorted(p.name for p in out.glob("layer_*.png")) == META[case]["layers"]
    for name in META[case]["layers"]:
        i = int(name.split("_")[1]) - 1
        assert np.array_equal(cv2.imread(str(out / name), cv2.IMREAD_UNCHANGED), (want == i).astype(np.uint8) * 255), name
    log = r.stdout.replace(str(out), "@OUT@").splitlines()
    for line in META[case]["log"]:
        assert line in log, line


@pytest.mark.gpu
def test_dropin_cli_errors(tmp_path):
    r = subprocess.run([sys.executable, SCRIPT, str(tmp_path / "missing.png"), "-o", str(tmp_path / "o")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode != 0 and "ValueError: Cannot load image" in r.stdout
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), _z()["input"])
    r = subprocess.run([sys.executable, SCRIPT, str(src), "-o", str(tmp_path / "o"), "-m", "palette"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode != 0 and "Mode 'palette' requires --palette JSON" in r.stdout
