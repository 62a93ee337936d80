"""CPU suite: the C-ABI library builds in-tree, loads, and exports every symbol include/bamm_b200.h declares.
No compute call is made (there is no GPU here and the product has no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from bammmotif2_b200 import build
    return build.build_lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bamm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bamm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("bamm_seqset_create", "bamm_em_estep", "bamm_em_mstep", "bamm_em_optimize", "bamm_score_logodds"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_every_declared_symbol(lib_path):
    from bammmotif2_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_symbols()
    capi.load()


def test_no_device_is_reported_not_faked(lib_path):
    """Without a GPU the library says so; nothing silently computes on the CPU."""
    import torch
    from bammmotif2_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.device_count() == 0
    import numpy as np
    with pytest.raises(capi.BammError):
        capi.SeqSet(np.array([1, 2, 3, 4], np.uint8), np.array([0, 4], np.uint64), 4)
