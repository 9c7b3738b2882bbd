"""Property tests (hypothesis) of the native host code against the Python restatements it replaces."""
import os
import random

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from selfrec_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    from selfrec_b200 import build
    build.build()
    return _lib.load()


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(0, 3000), frac=st.floats(0.0, 1.0), seed=st.integers(0, 2 ** 31 - 1))
def test_sample_range_equals_random_sample(lib, n, frac, seed):
    from selfrec_b200.data.augmentor import sample_range
    k = int(n * frac)
    random.seed(seed)
    want = random.sample(range(n), k)
    state = random.getstate()
    random.seed(seed)
    assert sample_range(n, k).tolist() == want and random.getstate() == state


_name = st.text(alphabet="abcdefXYZ0123456789_-", min_size=1, max_size=6)
_line = st.tuples(_name, _name, st.sampled_from(["1", "1.0", "3", "0.5", "2e0", "4.25"]))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(train=st.lists(_line, min_size=1, max_size=120), test=st.lists(_line, min_size=0, max_size=40),
       eol=st.sampled_from(["\n", "\r\n", "\r"]), final_newline=st.booleans())
def test_native_dataset_builder_equals_python_route(lib, tmp_path_factory, train, test, eol, final_newline):
    from selfrec_b200.data.loader import FileIO
    from selfrec_b200.data.native import load_interaction
    from selfrec_b200.data.ui_graph import Interaction
    d = tmp_path_factory.mktemp("ds")
    tr, te = d / "train.txt", d / "test.txt"
    tr.write_bytes((eol.join(" ".join(x) for x in train) + (eol if final_newline else "")).encode())
    te.write_bytes((eol.join(" ".join(x) for x in test) + (eol if (final_newline and test) else "")).encode())
    a = load_interaction(None, tr, te if test else None)
    b = Interaction(None, list(FileIO.load_data_set(str(tr))), list(FileIO.load_data_set(str(te))) if test else [])
    assert a.user == b.user and a.item == b.item
    assert (a.pair_users == b.pair_users).all() and (a.pair_items == b.pair_items).all()
    for x, y in ((a.norm_adj, b.norm_adj), (a.ui_adj, b.ui_adj), (a.interaction_mat, b.interaction_mat)):
        x, y = x.tocsr().copy(), y.tocsr().copy()
        x.sort_indices()
        y.sort_indices()
        assert (x.indptr == y.indptr).all() and (x.indices == y.indices).all() and (x.data.view(np.uint32) == y.data.view(np.uint32)).all()
    assert dict(a.test_set) == dict(b.test_set) and list(a.test_set) == list(b.test_set)
    assert a.training_data == b.training_data


# ------------------------------------------------------------------------------------------
# config-5 host logic that is plain tensor code (runs on CPU tensors): generator, row classes, split-row segments
# ------------------------------------------------------------------------------------------
def test_config5_pair_generator_properties_on_cpu():
    """synth.make_pairs_device (SURVEY 8d recipe) on the CPU device: exactly nnz DISTINCT pairs, every user and item has
    an edge, ids follow first appearance in the (shuffled) pair list like ui_graph.py:29-40, degrees are heavy-tailed."""
    import torch
    from selfrec_b200 import synth
    U, I, nnz = 3000, 800, 40000
    pu, pi = synth.make_pairs_device(U, I, nnz, seed=3, alpha=1.1, device="cpu")
    pu, pi = pu.numpy().astype(np.int64), pi.numpy().astype(np.int64)
    assert len(pu) == nnz and len(np.unique(pu * I + pi)) == nnz
    assert np.array_equal(np.unique(pu), np.arange(U)) and np.array_equal(np.unique(pi), np.arange(I))
    for ids, n in ((pu, U), (pi, I)):
        _, first = np.unique(ids, return_index=True)       # position of each id's first occurrence
        assert np.all(np.diff(first) > 0)                    # id k appears for the first time before id k + 1
    deg = np.bincount(pi, minlength=I)
    assert deg.max() > 20 * np.median(deg)                   # Zipf(1.1): hubs
    again = synth.make_pairs_device(U, I, nnz, seed=3, alpha=1.1, device="cpu")
    assert torch.equal(again[0], torch.from_numpy(pu.astype(np.int32))) and torch.equal(again[1], torch.from_numpy(pi.astype(np.int32)))


def test_row_classes_and_split_row_segments_on_cpu():
    """ops.classify_rows / ops.column_blocked_segments on CPU tensors: the processing order is degree-descending with the
    class sizes the kernels assume; the chunk list covers every split row exactly once; the column-blocked segments tile
    each split row's CSR range, never cross a column block or exceed a chunk, and are ordered (column block, row)."""
    import torch
    from selfrec_b200 import _lib, ops
    rng = np.random.default_rng(5)
    n_rows, n_cols, d = 400, 300000, 128          # wide enough for several 32 MB column blocks at d = 128
    deg = np.concatenate([[20000, 9000, 5000, 4096], rng.integers(0, 300, n_rows - 4)])
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    colidx = np.concatenate([np.sort(rng.choice(n_cols, k, replace=False)) for k in deg]).astype(np.int32)
    rp, ci = torch.from_numpy(rowptr.astype(np.int32)), torch.from_numpy(colidx)
    c = ops.classify_rows(rp)
    order = c["row_order"].numpy()
    assert np.all(np.diff(deg[order]) <= 0) and sorted(order.tolist()) == list(range(n_rows))
    n_huge, n_vlong, n_long = c["n_huge"], c["n_vlong"], c["n_long"]
    assert n_huge == 4 and np.all(deg[order[:n_huge]] >= _lib.HUB_MIN_NNZ)
    assert np.all(deg[order[n_huge:n_huge + n_vlong]] >= ops.VLONG_ROW_NNZ) and np.all(deg[order[n_huge + n_vlong:n_huge + n_vlong + n_long]] >= ops.LONG_ROW_NNZ)
    assert np.all(deg[order[n_huge + n_vlong + n_long:]] < ops.LONG_ROW_NNZ)
    work = c["hub_work"].numpy()
    first = c["hub_first"].numpy()
    for k, r in enumerate(order[:n_huge]):
        nch = -(-deg[r] // _lib.HUB_CHUNK)
        mine = work[first[k]:first[k] + nch]
        assert np.all(mine[:, 0] == r) and np.array_equal(mine[:, 1], np.arange(nch))
    assert c["n_work"] == sum(-(-deg[r] // _lib.HUB_CHUNK) for r in order[:n_huge])
    segs = ops.column_blocked_segments(rp, ci, c["row_order"][:n_huge].to(torch.int64), n_cols, d)
    assert segs is not None
    W = segs["block_cols"]
    seg, cnt, sfirst = segs["seg"].numpy(), segs["cnt"].numpy(), segs["first"].numpy()
    for k, r in enumerate(order[:n_huge]):
        mine = seg[sfirst[k]:sfirst[k] + cnt[k]]
        assert mine[0, 0] == rowptr[r] and mine[-1, 1] == rowptr[r + 1] and np.array_equal(mine[1:, 0], mine[:-1, 1])  # tiles the row
        for b, e in mine:
            assert 0 < e - b <= _lib.HUB_CHUNK and colidx[b] // W == colidx[e - 1] // W                                 # one block, one chunk
    proc = np.concatenate([segs["order_cta"].numpy(), segs["order_warp"].numpy()])
    assert sorted(proc.tolist()) == list(range(segs["n_seg"]))
    for lst, long_ in ((segs["order_cta"].numpy(), True), (segs["order_warp"].numpy(), False)):
        ln = seg[lst, 1] - seg[lst, 0]
        assert np.all(ln > _lib.HUB_WARP_SEG) if long_ else np.all(ln <= _lib.HUB_WARP_SEG)
        blk = colidx[seg[lst, 0]] // W
        assert np.all(np.diff(blk) >= 0)                      # column block by column block
    assert ops.column_blocked_segments(rp, ci, c["row_order"][:n_huge].to(torch.int64), 2 * W, d) is None  # a narrow matrix keeps plain chunks
