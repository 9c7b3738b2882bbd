"""Device-side graph assembly (srb_graph_assemble), split ("huge") rows of the SpMM, and the engine on a
Zipf(1.1) graph built entirely on the GPU (the config-5 recipe at a size the float64 oracle finishes in seconds).

Bit-exact: CSR structure and fp32 values against the scipy route of the reference (data/ui_graph.py:47-65,
data/graph.py:10-24, data/augmentor.py:30-40).  1e-4 relative: embeddings / losses / parameters."""
import random

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    assert torch.cuda.is_available()
    from selfrec_b200 import _lib
    _lib.require_device()
    return torch


def _same_csr(adj, ref):
    ref = sp.csr_matrix(ref)
    ref.sort_indices()
    assert adj.shape == ref.shape
    np.testing.assert_array_equal(adj.rowptr.cpu().numpy(), ref.indptr)
    np.testing.assert_array_equal(adj.colidx.cpu().numpy(), ref.indices)
    got = adj.vals.cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.data.astype(np.float32).view(np.uint32))


def test_graph_assemble_bit_exact_vs_scipy_route(torch_cuda):
    """Full graph, incl. duplicate training lines (summed to 2.0 like amazon-kindle's 2 822), empty-ish tails."""
    from selfrec_b200 import synth
    from selfrec_b200.data.device_graph import DeviceBipartite
    pu, pi = synth.make_pairs(3000, 2000, 60000, seed=3)
    rng = np.random.default_rng(0)
    dup = rng.choice(len(pu), 500, replace=False)  # duplicate lines
    pu2, pi2 = np.concatenate([pu, pu[dup]]), np.concatenate([pi, pi[dup]])
    data = synth.ArrayInteraction(pu2, pi2, 3000, 2000)
    assert data.interaction_mat.data.max() == 2.0
    bip = DeviceBipartite.from_interaction_mat(data.interaction_mat, "cuda")
    _same_csr(bip.assemble(), data.norm_adj)
    # flags route == index route == host route for an arbitrary subset, weights reset to 1
    keep = np.sort(rng.choice(bip.nnz, bip.nnz // 3, replace=False)).astype(np.int64)
    m = sp.csr_matrix(data.interaction_mat)
    rows, cols = m.nonzero()
    sub = sp.csr_matrix((np.ones(len(keep), np.float32), (rows[keep], cols[keep])), shape=m.shape)
    ref = data.convert_to_laplacian_mat(sub)
    _same_csr(bip.assemble(keep_idx=keep, reset_weights=True), ref)
    flags = torch_cuda.zeros(bip.nnz, dtype=torch_cuda.uint8, device="cuda")
    flags[torch_cuda.from_numpy(keep).cuda()] = 1
    _same_csr(bip.assemble(keep_flags=flags, reset_weights=True), ref)


def test_sgl_edge_dropout_view_on_device_matches_reference_route(torch_cuda):
    """SGL.py:89-96: random.sample keep-list (native, CPython-exact) -> device assembly == scipy route, same draws."""
    from selfrec_b200 import synth
    from selfrec_b200.data.augmentor import GraphAugmentor, sample_range
    from selfrec_b200.data.device_graph import DeviceBipartite
    data = synth.make_interaction((4000, 3000, 90000), seed=5)
    bip = DeviceBipartite.from_interaction_mat(data.interaction_mat, "cuda")
    random.seed(11)
    ref = data.convert_to_laplacian_mat(GraphAugmentor.edge_dropout(data.interaction_mat, 0.1))
    random.seed(11)
    keep = sample_range(bip.nnz, int(bip.nnz * (1 - 0.1)))
    _same_csr(bip.assemble(keep_idx=keep, reset_weights=True), ref)


@pytest.mark.parametrize("d", [64, 128])
def test_spmm_split_rows_vs_oracle(torch_cuda, orc, d):
    """Rows above SRB_HUB_MIN_NNZ are cut into chunks (spmm_hub_kernel + spmm_hub_finish_kernel)."""
    torch = torch_cuda
    from selfrec_b200 import _lib, ops
    rng = np.random.default_rng(d)
    n_rows, n_cols = 600, 30000
    deg = rng.integers(0, 40, n_rows)
    deg[[3, 77, 500]] = [_lib.HUB_MIN_NNZ, 3 * _lib.HUB_CHUNK + 17, 20000]
    deg[10] = _lib.HUB_MIN_NNZ - 1
    rows = np.repeat(np.arange(n_rows), deg)
    cols = np.concatenate([rng.choice(n_cols, k, replace=False) for k in deg])
    A = sp.csr_matrix((rng.standard_normal(len(rows)).astype(np.float32), (rows, cols)), shape=(n_rows, n_cols))
    X = rng.standard_normal((n_cols, d)).astype(np.float32)
    h = ops.SparseAdj(A).cuda()
    assert h.n_huge == 3 and h.n_work == 2 + 4 + 10
    y = torch.sparse.mm(h, torch.from_numpy(X).cuda()).cpu().numpy()
    ref = orc.spmm(A, X)
    scale = np.abs(A).dot(np.abs(X))
    assert (np.abs(y - ref) <= 4e-6 * scale + 1e-30).all()
    # masked product (row-sparse X): same answer as the plain one on the masked input
    mask_rows = rng.random(n_cols) < 0.05
    Xm = X * mask_rows[:, None]
    bits = np.packbits(mask_rows, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.int32)
    ym = torch.empty(n_rows, d, device="cuda")
    ops._spmm_raw(h, torch.from_numpy(Xm).cuda(), ym, col_mask=torch.from_numpy(bits).cuda())
    refm = orc.spmm(A, Xm)
    assert (np.abs(ym.cpu().numpy() - refm) <= 4e-6 * np.abs(A).dot(np.abs(Xm)) + 1e-30).all()


def _batch(data, b, cap, rng):
    from selfrec_b200 import _lib
    pu, pi = data.pair_users, data.pair_items
    sel = rng.choice(len(pu), b, replace=False)
    u, i = pu[sel].astype(np.int32), pi[sel].astype(np.int32)
    rp, ri = data.rated_csr()
    j = np.empty(b, np.int32)
    for k, uu in enumerate(u):
        rated = set(ri[rp[uu]:rp[uu + 1]].tolist())
        while True:
            c = int(rng.integers(0, data.item_num))
            if c not in rated:
                j[k] = c
                break
    w = np.zeros(_lib.BATCH_HEADER + 5 * cap, np.int32)
    uq, iq = np.unique(u), np.unique(i)
    w[0], w[1], w[2] = b, len(uq), len(iq)
    H = _lib.BATCH_HEADER
    w[H:H + b], w[H + cap:H + cap + b], w[H + 2 * cap:H + 2 * cap + b] = u, i, j
    w[H + 3 * cap:H + 3 * cap + len(uq)], w[H + 4 * cap:H + 4 * cap + len(iq)] = uq, iq
    return w, u, i, j


@pytest.mark.parametrize("model,d", [("SimGCL", 128), ("XSimGCL", 64)])
def test_engine_step_on_device_built_zipf_graph_vs_oracle(torch_cuda, orc, model, d):
    """config-5 recipe (Zipf 1.1 both sides, built and normalised on the GPU) at 30 k x 8 k x 1.2 M: split rows in
    the full products AND among the batch rows (hub users sit in every batch); one step vs the float64 oracle."""
    torch = torch_cuda
    from selfrec_b200 import _lib, synth
    from selfrec_b200.engine import TrainEngine
    data = synth.make_device_interaction((30000, 8000, 1200000), seed=2, alpha=1.1)
    adj = data.norm_adj
    assert adj.n_huge > 0, "the graph must exercise the split-row path"
    A = sp.csr_matrix((adj.vals.cpu().numpy(), adj.colidx.cpu().numpy(), adj.rowptr.cpu().numpy()), shape=adj.shape)
    # the device-built matrix is the scipy route's matrix
    host = synth.ArrayInteraction(data.pair_users, data.pair_items, data.user_num, data.item_num)
    _same_csr(adj, host.norm_adj)
    torch.manual_seed(0)
    L, B = 2, 512
    views = 2 if model == "SimGCL" else 1
    eng = TrainEngine(model, data, d, L, B, 1e-3, 1e-4, eps=0.1, tau=0.2, cl_rate=0.5, layer_cl=1)
    rng = np.random.default_rng(1)
    noise = rng.random((views, L, eng.N, d), dtype=np.float32)
    eng.set_noise_tensor(torch.from_numpy(noise).cuda())
    E0 = eng.params.cpu().numpy().copy()
    w, u, i, j = _batch(data, 500, B, rng)
    deg = np.diff(A.indptr)
    assert (deg[u] >= _lib.HUB_MIN_NNZ).any(), "a hub user must be in the batch"
    eng.step(w)
    torch.cuda.synchronize()
    ref = orc.train_step(model, A, E0, eng.U, u, i, j, n_layers=L, reg=1e-4, batch_size=B, eps=0.1, tau=0.2, cl_rate=0.5,
                         layer_cl=1, noise=noise)
    los = eng.losses.cpu().numpy()
    for got, want in ((los[0], ref["rec"]), (los[1], ref["l2"]), (los[2], ref["cl"])):
        assert abs(got - want) <= 1e-4 * abs(want), (got, want)
    P, _, _ = orc.adam_step(E0, ref["grad"].astype(np.float32), np.zeros_like(E0), np.zeros_like(E0), 1, 1e-3)
    got = eng.params.cpu().numpy()
    # Adam's first step moves every touched entry by ~lr; entries whose gradient is within rounding of zero may flip
    bad = np.abs(got - P) > 1e-4 * np.abs(P) + 1e-6
    assert bad.mean() < 1e-3, bad.mean()
    g = ref["grad"]
    big = np.abs(g) > 1e-3 * np.abs(g).max()
    assert not (bad & big).any()


@pytest.mark.parametrize("aug_type", [0, 1])
def test_sgl_model_views_device_route_equals_reference_route(torch_cuda, tiny_conf, tiny_triples, in_tmp_cwd, aug_type):
    """SGL.random_graph_augment (fused model): node dropout (aug_type 0) and edge dropout (1) built on the device equal
    GraphAugmentor + convert_to_laplacian_mat on the host, draw for draw (same `random` stream, same final state)."""
    from selfrec_b200.data.augmentor import GraphAugmentor
    from selfrec_b200.model.graph.SGL import SGL
    train, test = tiny_triples
    m = SGL(tiny_conf("SGL", {"n_layer": 2, "lambda": 0.1, "drop_rate": 0.2, "aug_type": aug_type, "temp": 0.2}),
            [list(t) for t in train], [list(t) for t in test])
    random.seed(77)
    drop = GraphAugmentor.node_dropout if aug_type == 0 else GraphAugmentor.edge_dropout
    ref = [m.data.convert_to_laplacian_mat(drop(m.data.interaction_mat, 0.2)) for _ in range(2)]
    state = random.getstate()
    random.seed(77)
    got = [m.random_graph_augment() for _ in range(2)]
    assert random.getstate() == state
    for a, r in zip(got, ref):
        r = sp.csr_matrix(r)
        r.eliminate_zeros()
        _same_csr(a, r)
    m._epoch_prologue(0)  # and the engine accepts the device handles
    assert m.engine.view_adj[0].rowptr.is_cuda


@pytest.mark.parametrize("d", [64, 128])
def test_spmm_column_blocked_split_rows_vs_oracle(torch_cuda, orc, d, monkeypatch):
    """Column-blocked split-row lists (srb_hub_split.seg): segments cut at column-block boundaries, processed in
    (block, row) order by a CTA or a warp each, summed per row in segment order."""
    torch = torch_cuda
    from selfrec_b200 import _lib, ops
    monkeypatch.setattr(ops, "HUB_BLOCK_BYTES", 2048 * 4 * d)  # blocks of 4096 columns (the floor): many blocks at test size
    rng = np.random.default_rng(d + 1)
    n_rows, n_cols = 400, 60000
    deg = rng.integers(0, 40, n_rows)
    deg[[5, 99, 300, 301]] = [_lib.HUB_MIN_NNZ, 25000, 7000, 59000]
    rows = np.repeat(np.arange(n_rows), deg)
    cols = np.concatenate([rng.choice(n_cols, k, replace=False) for k in deg])
    A = sp.csr_matrix((rng.standard_normal(len(rows)).astype(np.float32), (rows, cols)), shape=(n_rows, n_cols))
    X = rng.standard_normal((n_cols, d)).astype(np.float32)
    h = ops.SparseAdj(A).cuda()
    hs = h.hub_struct(d)
    assert h.n_huge == 4 and hs.seg and hs.n_cta > 0 and hs.n_warp > 0
    y = torch.sparse.mm(h, torch.from_numpy(X).cuda()).cpu().numpy()
    ref = orc.spmm(A, X)
    assert (np.abs(y - ref) <= 4e-6 * np.abs(A).dot(np.abs(X)) + 1e-30).all()
    mask_rows = rng.random(n_cols) < 0.05
    Xm = X * mask_rows[:, None]
    bits = np.packbits(mask_rows, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.int32)
    ym = torch.empty(n_rows, d, device="cuda")
    ops._spmm_raw(h, torch.from_numpy(Xm).cuda(), ym, col_mask=torch.from_numpy(bits).cuda())
    assert (np.abs(ym.cpu().numpy() - orc.spmm(A, Xm)) <= 4e-6 * np.abs(A).dot(np.abs(Xm)) + 1e-30).all()
