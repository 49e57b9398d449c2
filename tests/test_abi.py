"""CPU: the C-ABI library loads, exports every symbol include/tpq_b200.h declares, and rejects bad
arguments with the reference's error behaviour -- without any compute call (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tpq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tpq_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    import torchpq_b200 as T
    names = declared_symbols()
    assert len(names) >= 15
    lib = C.CDLL(T._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tpq_b200.h but not exported"
        assert n in T._lib.SIGNATURES, f"{n} has no ctypes signature in torchpq_b200/_lib.py"
    assert sorted(T._lib.SIGNATURES) == names


def test_version_and_error_channel():
    import torchpq_b200 as T
    assert T._lib.lib.tpq_version() >= 1
    # k outside (0, 1024] is the reference's `assert 0 < k <= 1024` (fn/IVFPQTopk.py:64): AssertionError
    rc = T._lib.lib.tpq_ivfpq_topk(None, None, None, None, None, None, 0, 8, 1, 1, 0, None, None, None, 0, None)
    assert rc == T._lib.TPQ_ERR_BAD_ARG
    assert "k must be in (0, 1024]" in T._lib.last_error()
    with pytest.raises(AssertionError):
        T._lib.check(rc)
    rc = T._lib.lib.tpq_ivfpq_topk(None, None, None, None, None, None, 0, 6, 1, 1, 5, None, None, None, 0, None)
    assert rc == T._lib.TPQ_ERR_BAD_ARG and "multiple of 4" in T._lib.last_error()
    rc = T._lib.lib.tpq_build_lut(None, None, 10, 3, 1, 0, None, None)
    assert rc == T._lib.TPQ_ERR_BAD_ARG


def test_struct_layout_matches_header():
    import torchpq_b200 as T
    # 4 int32 + int64 + 7 ptr + 4 int32 + int64 + 5 ptr + ptr + 2 int32
    assert C.sizeof(T._lib.TpqIndex) == 16 + 8 + 7 * 8 + 16 + 8 + 5 * 8 + 8 + 8


def test_no_cpu_path():
    """The product must fail loudly rather than fall back when there is no CUDA device."""
    import torch
    import torchpq_b200 as T
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    ix = T.IVFPQIndex(32, 8, 4, initial_size=4, device="cpu")
    ix.vq_codec.set_codebook(torch.zeros(32, 4))
    ix.pq_codec.set_codebook(torch.zeros(8, 4, 256))
    with pytest.raises(AssertionError):
        ix.search(torch.zeros(32, 2), k=1)
    with pytest.raises(AssertionError):
        T.fn.precompute_adc(torch.zeros(32, 2), torch.zeros(8, 4, 256))


def test_state_dict_names_match_reference():
    """Checkpoint keys of torchpq.index.IVFPQIndex."""
    import torch
    import torchpq_b200 as T
    ix = T.IVFPQIndex(32, 8, 4, initial_size=4, device="cpu")
    ix.vq_codec.set_codebook(torch.zeros(32, 4))
    ix.pq_codec.set_codebook(torch.zeros(8, 4, 256))
    keys = set(ix.state_dict().keys())
    assert {"_address2id", "_storage", "_cell_start", "_cell_size", "_cell_capacity", "_is_empty",
            "vq_codec._is_trained", "vq_codec.kmeans.centroids", "pq_codec._is_trained",
            "pq_codec.kmeans.centroids"} <= keys
    sd = ix.state_dict()
    sd["_storage"] = torch.zeros(2, 40, 4, dtype=torch.uint8)          # shape change on load is allowed
    sd["_address2id"] = -torch.ones(40, dtype=torch.long)
    sd["_is_empty"] = torch.ones(40, dtype=torch.uint8)
    ix2 = T.IVFPQIndex(32, 8, 4, initial_size=4, device="cpu")
    ix2.load_state_dict(sd)
    assert ix2._storage.shape == (2, 40, 4) and ix2.capacity == 40 and ix2.vq_codec.is_trained
