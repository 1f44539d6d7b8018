"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/eqvio.h declares, and
fails loudly (no CPU fallback) when there is no device.  No compute calls here."""
import ctypes
import os
import re

import pytest

from eqf_vio_b200 import abi
from eqf_vio_b200.settings import Settings, default_settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "eqvio.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqvio_[a-z_A-Z0-9]+)\s*\(", text)))


def test_library_built_and_loads():
    assert os.path.exists(abi.LIB_PATH), "run `python -m eqf_vio_b200.build`"
    L = abi.lib()
    assert b"sm_100a" in L.eqvio_version()


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(abi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 28
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # and the Python binding covers the whole header
    assert sorted(abi.SIGNATURES) == declared


def test_settings_struct_matches_header_defaults():
    L = abi.lib()
    s = Settings()
    assert L.eqvio_settings_default(ctypes.byref(s)) == 0
    assert s.as_dict() == default_settings().as_dict()
    assert ctypes.sizeof(Settings) == 15 * 8 + 4 * 4 + 13 * 8


def test_sass_is_blackwell_native():
    """TMA (UTMALDG) and DMMA in the shipped cubin; no generic-PTX fallback."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTMALDG" in out and "DMMA" in out and "SYNCS" in out


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from eqf_vio_b200.filter import VIOFilter

    with pytest.raises(abi.EqvioError) as e:
        VIOFilter(default_settings())
    assert e.value.status == abi.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eqf_vio_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, fn)


def test_cpp_facade_compiles_and_runs(tmp_path):
    """include/eqvio/VIOFilter.hpp (the reference's class surface on POD types) builds against the C ABI;
    the example exits 0 both with a GPU (runs the filter) and without (reports EQVIO_ERR_NO_DEVICE)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "replay_minimal")
    csrc = os.path.join(ROOT, "eqf_vio_b200", "csrc")
    subprocess.check_call([gxx, "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "replay_minimal.cpp"),
                           "-L", csrc, "-leqvio_b200", "-Wl,-rpath," + csrc, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
