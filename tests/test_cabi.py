"""CPU: libsfb200.so loads and exports exactly the symbols include/sfb200.h declares (no compute calls)."""
import ctypes
import os
import re

from shapeformer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert getattr(lib, n) is not None, n
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table and header disagree"


def test_version_and_errors():
    lib = _lib.load()
    assert lib.sfb200_version() == 100
    assert lib.sfb200_error_string(0) == b"ok"
    assert b"argument" in lib.sfb200_error_string(-1)


def test_layout_queries_need_no_gpu():
    from shapeformer_b200 import ar, synth
    lib = _lib.load()
    cfg = ar._cfg_struct(synth.SHIPPED_GPT, (4096, 4096), 64, 769, 512, 16, 256, 1)
    n = lib.sfb200_ar_weight_floats(ctypes.byref(cfg))
    # 324.95 M parameters
    assert abs(n - 324.95e6) < 0.05e6
    assert lib.sfb200_ar_weight_offset(ctypes.byref(cfg), _lib.W_POS_EMB, 0, 0) == 0
    assert lib.sfb200_ar_weight_offset(ctypes.byref(cfg), _lib.W_FC2_B, 1, 3) + 1024 == n
    assert lib.sfb200_ar_weight_offset(ctypes.byref(cfg), _lib.W_FC2_B, 1, 4) < 0
    assert lib.sfb200_ar_kv_bytes(ctypes.byref(cfg)) == 24 * 2 * 64 * 16 * 769 * 64 * 4
    assert lib.sfb200_ar_history_floats(ctypes.byref(cfg)) == 64 * 512 * 2 * 4097
    bad = ar._cfg_struct(dict(synth.SHIPPED_GPT, n_head=12), (4096, 4096), 1, 8, 1, 1, 1, 0)
    assert lib.sfb200_ar_weight_floats(ctypes.byref(bad)) < 0


def test_cpu_tensor_is_rejected():
    import pytest
    import torch
    with pytest.raises(_lib.Sfb200Error):
        _lib.ptr(torch.zeros(4))
