"""CPU checks of the C-ABI boundary: struct layouts, exported symbols, loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from okvis_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_numpy_layouts_match_ctypes():
    assert abi.camera_dtype.itemsize == C.sizeof(abi.Camera) == 112
    assert abi.observation_dtype.itemsize == 40
    assert abi.imu_sample_dtype.itemsize == 56
    assert abi.imu_term_dtype.itemsize == 40
    assert abi.pose_prior_dtype.itemsize == 8 + 7 * 8 + 36 * 8
    assert abi.sb_prior_dtype.itemsize == 8 + 9 * 8 + 81 * 8
    assert abi.relpose_dtype.itemsize == 8 + 36 * 8
    assert abi.keypoint_dtype.itemsize == 24
    assert abi.pair_dtype.itemsize == 8
    assert C.sizeof(abi.ImuParams) == 14 * 8
    assert C.sizeof(abi.Summary) == 48
    assert C.sizeof(abi.SolveOptions) == 24


def test_library_exports_every_declared_symbol():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(abi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "okvis_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(okb_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (OKB_ERR_NO_DEVICE), never compute."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from okvis_b200 import capi
    with pytest.raises(capi.OkbError) as e:
        capi.Context(0, 1)
    assert e.value.code == abi.OKB_ERR_NO_DEVICE


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "okvis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "/oracle/" not in txt, os.path.join(dirpath, f)
