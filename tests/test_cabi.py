"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/b200match.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lidar_global_registration_b200 import build as b200_build
from lidar_global_registration_b200 import matcher as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    b200_build.build()
    return M.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200match.h")).read()
    return sorted(set(re.findall(r"B200M_API\s+[\w\s\*]+?\b(b200m_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 16
    assert sorted(names) == sorted(M.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None, n


def test_struct_layouts_match_header():
    # b200m_corr == reference Correspondence (include/common.h:120-131): 16 bytes
    assert M.CORR_DTYPE.itemsize == 16
    assert C.sizeof(M._Params) == 32     # k, mode, ratio_thr, distance_thr, precision, cand_cap, n_gpus, shard
    assert C.sizeof(M.Stats) == 6 * 8 + 7 * 8


def test_version(lib):
    assert lib.b200m_version() >= 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(M.B200MatchError) as e:
        M.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(M.B200MatchError):
        M.match_bf(np.zeros((4, 33), np.float32), np.zeros((4, 33), np.float32), M.AlignmentParameters())


def test_product_package_does_not_import_oracle():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "lidar_global_registration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
