"""The C-ABI library loads without a GPU and exports every symbol include/fumi_b200.h declares."""
import ctypes
import os
import re

import pytest

from fumi_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fumi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fumi_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    path = build.build()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/fumi_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTED), set(syms) ^ set(_lib.EXPORTED)
    assert L.fumi_abi_version() == 1


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fumi_b200.engine import EpisodeEngine
    with pytest.raises(_lib.FumiError, match="CUDA"):
        EpisodeEngine("cpu")
    L = ctypes.CDLL(build.build())
    L.fumi_last_error.restype = ctypes.c_char_p
    assert L.fumi_device_sm_count() < 0          # FUMI_ERR_CUDA: no device
    assert b"cuda" in L.fumi_last_error().lower()
