"""The C-ABI library loads on a CPU-only box, exports every symbol include/sodso_pr.h declares,
and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import re

import pytest

from so_dso_place_recognition_b200 import _native as N


def _declared():
    txt = open(N.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sodso_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 25
    L = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sodso_pr.h but not exported"
    assert set(names) == set(N.exported_names()), "python prototypes out of sync with the header"


def test_sizes_and_version():
    L = N.lib()
    assert L.sodso_sc_signature_size() == 1200      # SC.cpp:10
    assert L.sodso_m2dp_signature_size() == 192     # M2DP.cpp:36
    assert b"sm_100a" in L.sodso_version()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from so_dso_place_recognition_b200 import api

    with pytest.raises(N.SodsoError, match="no CPU fallback|no CUDA device"):
        api.Context(0)


def test_topk_merge_host_helper():
    import numpy as np

    from so_dso_place_recognition_b200 import api

    idx = np.array([[[5, 9, -1]], [[2, 7, 30]]], dtype=np.int64)          # R=2, m=1, k=3
    sc = np.array([[[0.1, 0.5, np.nan]], [[0.1, 0.2, 0.9]]])
    oi, os_, op, od = api.topk_merge(idx, sc, sc * 2, sc * 3)
    assert oi.tolist() == [[2, 5, 7]]
    np.testing.assert_allclose(os_, [[0.1, 0.1, 0.2]])
    np.testing.assert_allclose(op, [[0.2, 0.2, 0.4]])
