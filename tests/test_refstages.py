"""Build-container checks (marker `reference`: need /root/reference):
  * oracle/refstages.py -- the as-shipped CPU baseline bench.py times on the GPU box -- writes the same files as the unmodified
    reference's pipeline.py run;
  * the drop-in stage scripts copied beside a copy of the reference's pipeline.py are the files its build_steps()/module_path()
    resolve, and its own `missing_for_step` bookkeeping accepts the outputs they are expected to write."""
import importlib.util
import json
import os
import shutil
import subprocess
import sys

import cv2
import numpy as np
import pytest

from helpers import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/image_processor"
DROPIN = os.path.join(ROOT, "omnirevolve-image-processor_b200", "image_processor")
pytestmark = pytest.mark.reference


def _files(outdir, names):
    return (cv2.imread(os.path.join(outdir, "resized.png"), cv2.IMREAD_COLOR),
            [cv2.imread(os.path.join(outdir, n, "mask.png"), cv2.IMREAD_GRAYSCALE) for n in names],
            [cv2.imread(os.path.join(outdir, n, "edges.png"), cv2.IMREAD_GRAYSCALE) for n in names],
            cv2.imread(os.path.join(outdir, "edges_composite.png"), cv2.IMREAD_COLOR),
            json.load(open(os.path.join(outdir, "palette_by_name.json"))))


def test_refstages_equals_reference_pipeline(tmp_path):
    img = synth(300, 420, 3)
    src = tmp_path / "input.png"
    cv2.imwrite(str(src), img)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    out_ref = tmp_path / "ref"
    out_ref.mkdir()
    r = subprocess.run([sys.executable, os.path.join(REF, "pipeline.py"), str(src), "--output", str(out_ref), "--start-step", "1",
                        "--end-step", "3"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    cfg = json.load(open(out_ref / "config.json"))
    out_port = tmp_path / "port"
    out_port.mkdir()
    cfg2 = dict(cfg, output_dir=str(out_port))
    (out_port / "config.json").write_text(json.dumps(cfg2))
    for st in ("01", "02", "03"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "refstages.py"), st],
                           env=dict(env, CONFIG_PATH=str(out_port / "config.json")), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
    names = cfg["color_names"]
    a, b = _files(str(out_ref), names), _files(str(out_port), names)
    assert np.array_equal(a[0], b[0])
    for x, y in zip(a[1] + a[2], b[1] + b[2]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[3], b[3]) and a[4] == b[4]


def test_reference_pipeline_resolves_the_dropins(tmp_path):
    """Copy the reference's runner + config module and OUR stage scripts into one directory, as INTEGRATION.md A says, and ask
    the runner itself which files it would launch and which outputs it waits for."""
    d = tmp_path / "image_processor"
    d.mkdir()
    for f in ("pipeline.py", "config.py"):
        shutil.copy(os.path.join(REF, f), d / f)
    for f in os.listdir(DROPIN):
        if f.endswith(".py") and f != "config.py":
            shutil.copy(os.path.join(DROPIN, f), d / f)
    sys.path.insert(0, str(d))
    try:
        spec = importlib.util.spec_from_file_location("ref_pipeline_copy", str(d / "pipeline.py"))
        pl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pl)
        steps = pl.build_steps()                                    # [(title, path)] (pipeline.py:66-82)
        for i, script in enumerate(("01_resize.py", "02_color_extract.py", "03_edge_detect.py")):
            assert os.path.samefile(steps[i][1], d / script)        # the runner launches OUR file
            assert os.path.samefile(pl.module_path(script), d / script)
            assert "omni_b200" in open(steps[i][1]).read()
        # the runner's resume bookkeeping names exactly the files our stages write (pipeline.py:113-145)
        out = tmp_path / "out"
        names = ["layer_dark", "layer_mid"]
        need = pl.missing_for_step(4, str(out), names)
        assert set(need) == {str(out / "resized.png")} | {str(out / n / f) for n in names for f in ("mask.png", "edges.png")}
    finally:
        sys.path.remove(str(d))
        sys.modules.pop("config", None)
