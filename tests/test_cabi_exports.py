"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/bbgpu.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

import __graft_entry__ as entry
from bboptpy_b200 import _cabi


@pytest.fixture(scope="module")
def built():
    entry.build()
    return _cabi.lib()


def test_every_declared_symbol_is_exported(built):
    names = _cabi.symbols_in_header()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    for n in names:
        assert getattr(built, n) is not None


def test_version_and_error_string(built):
    assert built.bbgpu_version() == 100
    assert isinstance(built.bbgpu_last_error(), bytes)


def test_no_cpu_fallback_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = built.bbgpu_ctx_create(0, None, ctypes.byref(h))
    assert rc == 3  # BBGPU_ERR_NODEVICE
    assert b"no CPU fallback" in built.bbgpu_last_error()
    with pytest.raises(_cabi.BbgpuError):
        _cabi.Context(0)


def test_product_package_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hits = []
    for dirpath, _, files in os.walk(os.path.join(root, "bboptpy_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if "oracle/" in text or "from oracle" in text or "import oracle" in text or "bbo_port" in text or "libbbo_" in text:
                    hits.append(os.path.join(dirpath, f))
    assert not hits, hits
