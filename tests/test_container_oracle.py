"""CPU: the oracle's container write side against the behaviours the reference's own container tests
assert (reference tests/CellContainerTestCase.py:92-241: add with / without ids, data round trip,
is_empty, expand doubling, add after holes) -- the only tests the reference ships near this path."""
import numpy as np
import torch

from oracle import ivfpq_oracle as O, build_state as B


def fresh(M=8, C=8, initial=128):
    return O.empty_state(M * 2, M, C, initial)


def test_empty():                                              # CellContainerTestCase.py:238-241
    st = fresh()
    assert st.cell_size.sum() == 0 and (st.is_empty == 1).all() and (st.address2id == -1).all()


def test_add_with_ids_roundtrip():                             # :104-125
    rng = np.random.default_rng(0)
    st = fresh()
    n = 10000
    codes = rng.integers(0, 256, (8, n)).astype(np.uint8)
    cells = rng.integers(0, 8, n)
    ids = rng.permutation(2 ** 40)[:n] if False else rng.choice(2 ** 40, n, replace=False)
    rid, adr = O.container_add(st, codes, cells, ids)
    assert np.array_equal(rid, ids)
    assert np.array_equal(O.get_id_by_address(st.address2id, adr), ids)
    assert np.array_equal(O.codes_at(st, adr), codes)
    assert (st.is_empty[adr] == 0).all()
    assert st.cell_size.sum() == n and np.array_equal(st.cell_size, np.bincount(cells, minlength=8))
    # cells are disjoint contiguous ranges and every item sits inside its own cell
    assert (st.cell_start[1:] == st.cell_start[:-1] + st.cell_capacity[:-1]).all()
    assert ((adr >= st.cell_start[cells]) & (adr < st.cell_start[cells] + st.cell_capacity[cells])).all()


def test_add_without_ids_arange():                             # :127-146
    rng = np.random.default_rng(1)
    st = fresh()
    codes = rng.integers(0, 256, (8, 500)).astype(np.uint8)
    rid, adr = O.container_add(st, codes, rng.integers(0, 8, 500))
    assert np.array_equal(rid, np.arange(500))
    rid2, _ = O.container_add(st, codes, rng.integers(0, 8, 500))
    assert np.array_equal(rid2, np.arange(500, 1000)) and st.max_id == 999


def test_expand_doubles():                                     # :95-102
    st = fresh(initial=4)
    codes = np.zeros((8, 9), np.uint8)
    O.container_add(st, codes, np.zeros(9, np.int64))          # 9 items into a 4-slot cell -> 4 -> 8 -> 16
    assert st.cell_capacity[0] == 16 and st.cell_capacity[1] == 4
    assert st.cell_start[1] == 16 and st.capacity == 16 + 7 * 4


def test_placement_is_input_order_and_fills_holes():           # :196-236 (add / remove interaction)
    st = fresh(C=2, initial=8)
    codes = np.arange(6, dtype=np.uint8)[None].repeat(8, 0)
    cells = np.array([1, 0, 1, 1, 0, 1])
    _, adr = O.container_add(st, codes, cells)
    assert adr.tolist() == [8, 0, 9, 10, 1, 11]                # ioa-th empty slot of the cell, input order
    st.is_empty[9] = 1; st.address2id[9] = -1                  # a hole (what a working remove would leave)
    _, adr2 = O.container_add(st, codes[:, :2], np.array([1, 1]))
    assert adr2.tolist() == [9, 12]


def test_search_padding_and_ids():
    st, queries = B.integer_state(16, 8, 4, 6, seed=0)
    st.n_probe, st.use_smart_probing = 4, False
    v, i = O.search(st, queries(3), k=10)
    assert (i[:, 6:] == -1).all() and np.isinf(v[:, 6:]).all() and (i[:, :6] >= 0).all()


def test_effective_probes_quirks():
    # (int) truncation of INT64_MIN -> 0 -> cell 0 only; values above n_probe are clamped
    npl = np.array([O.INT64_MIN, 0, 1, 5, 9], np.int64)
    assert O.effective_probes(npl, 8).tolist() == [1, 1, 1, 5, 8]
    s = torch.tensor([[-1e9, -4e9]])                            # softmax underflow -> p = 0 -> NaN entropy
    assert O.smart_probing(s, 2, 30.0).item() == O.INT64_MIN
