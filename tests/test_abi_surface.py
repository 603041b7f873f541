"""The C-ABI library must load on a box without a GPU and export every symbol include/omni_b200.h
declares; compute entry points must fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "omni_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(omni_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    sys.path.insert(0, os.path.join(ROOT, "omnirevolve-image-processor_b200"))
    import build as omni_build
    return omni_build.build()


def test_header_symbols_exported(libpath):
    from omni_b200 import capi
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(libpath)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/omni_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names          # the ctypes binding covers exactly the header
    assert L.omni_version() == 1


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "omni_b200.h"\nint main(void){ omni_edge_params p; (void)p; return OMNI_MAX_K == 32 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_no_gpu_fails_loudly(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import omni_b200
    with pytest.raises(omni_b200.OmniError):
        omni_b200.Engine(0)
    L = omni_b200.capi.lib()
    h = ctypes.c_void_p()
    assert L.omni_ctx_create(0, ctypes.byref(h)) == -2          # OMNI_ERR_CUDA
    assert b"no CPU fallback" in L.omni_last_error_string()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the product package may reference it."""
    pkg = os.path.join(ROOT, "omnirevolve-image-processor_b200")
    for d, _s, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(d, f)
                assert "liboracle" not in txt, os.path.join(d, f)


def test_host_path_does_not_import_torch():
    """The stage scripts' host-buffer operators need ctypes + NumPy only: importing the package (and the stage mirrors)
    must not pull in torch -- a stage process should not pay an import the reference's stages never had."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "omnirevolve-image-processor_b200")
    code = ("import sys; sys.path.insert(0, %r); import omni_b200; from omni_b200 import stages, contours, batch; "
            "assert 'torch' not in sys.modules, 'torch imported eagerly'; print('LAZY_OK')" % pkg)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "LAZY_OK" in r.stdout, r.stdout[-1500:]
