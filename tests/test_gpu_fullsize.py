"""BASELINE.json's full sizes (yelp2018 shape, synthetic graph): direct oracle comparison where
the oracle finishes in seconds, size-independent properties elsewhere."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def yelp(built_lib):
    import torch
    assert torch.cuda.is_available()
    from selfrec_b200 import synth
    return synth.make_interaction("yelp2018", seed=0)


def test_spmm_full_size_vs_oracle_and_linearity(yelp, orc):
    import torch
    from selfrec_b200 import ops
    A = yelp.norm_adj.tocsr()
    n = A.shape[0]
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 64)).astype(np.float32)
    Y2 = rng.standard_normal((n, 64)).astype(np.float32)
    h = ops.SparseAdj(A).cuda()
    tx, ty = torch.from_numpy(X).cuda(), torch.from_numpy(Y2).cuda()
    out = torch.sparse.mm(h, tx)
    ref = orc.spmm(A, X)
    scale = abs(A).dot(np.abs(X))
    assert (np.abs(out.cpu().numpy() - ref) <= 4e-6 * scale + 1e-30).all()
    # linearity and the row-sum identity A 1 = rowsum(A)
    lin = torch.sparse.mm(h, tx + 2.0 * ty) - (out + 2.0 * torch.sparse.mm(h, ty))
    assert lin.abs().max().item() <= 2e-5
    ones = torch.sparse.mm(h, torch.ones(n, 64, device="cuda"))
    np.testing.assert_allclose(ones[:, 0].cpu().numpy(), np.asarray(A.sum(1)).ravel(), rtol=1e-5)
    # symmetric normalised adjacency: <Ax, y> == <x, Ay>
    lhs = (out * ty).sum().item()
    rhs = (tx * torch.sparse.mm(h, ty)).sum().item()
    assert abs(lhs - rhs) <= 1e-3 * max(abs(lhs), 1.0)


def test_xsimgcl_step_full_size_vs_oracle(yelp, orc, in_tmp_cwd):
    """One north-star step (XSimGCL, yelp2018 shape, L=3, d=64, B=2048) against the float64 oracle."""
    import torch
    from selfrec_b200.engine import TrainEngine
    torch.manual_seed(0)
    eng = TrainEngine("XSimGCL", yelp, 64, 3, 2048, 1e-3, 1e-4, eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1)
    U, N = eng.U, eng.N
    rng = np.random.default_rng(1)
    noise = rng.random((1, 3, N, 64), dtype=np.float32)
    eng.set_noise_tensor(torch.from_numpy(noise).cuda())
    E0 = eng.params.cpu().numpy().copy()
    b = 2048
    u = yelp.pair_users[:b].copy()
    i = yelp.pair_items[:b].copy()
    rp, ri = yelp.rated_csr()
    j = np.array([next(x for x in rng.integers(0, eng.I, 64) if x not in ri[rp[uu]:rp[uu + 1]]) for uu in u], dtype=np.int32)
    w = np.zeros(eng.words, dtype=np.int32)
    uq, iq = np.unique(u), np.unique(i)
    w[0], w[1], w[2] = b, len(uq), len(iq)
    for s, arr in enumerate((u, i, j, uq, iq)):
        w[4 + s * b:4 + s * b + len(arr)] = arr
    eng.step(w)
    torch.cuda.synchronize()
    out = orc.train_step("XSimGCL", yelp.norm_adj.tocsr(), E0, U, u, i, j, n_layers=3, reg=1e-4, batch_size=2048, eps=0.2,
                         tau=0.2, cl_rate=0.2, layer_cl=1, noise=noise)
    los = eng.losses.cpu().numpy()
    assert abs(los[0] - out["rec"]) <= 1e-4 * abs(out["rec"])
    assert abs(los[1] - out["l2"]) <= 1e-4 * abs(out["l2"])
    assert abs(los[2] - out["cl"]) <= 1e-4 * abs(out["cl"])
    P, _, _ = orc.adam_step(E0, out["grad"].astype(np.float32), np.zeros_like(E0), np.zeros_like(E0), 1, 1e-3)
    got = eng.params.cpu().numpy()
    # Step 1 of Adam moves a weight by lr * g / (|g| + eps).  Where |g| >> eps = 1e-8 this is +-lr and the
    # parameters must agree to the embedding tolerance; where |g| ~ eps the quotient amplifies the fp32
    # rounding of g by 1/(|g| + eps): bound that explicitly.
    g = np.abs(out["grad"])
    dg = 2e-6 * g + 5e-10  # measured: GPU vs oracle gradient differs by <= 3.5e-10 absolute (tools/grad_diag.py)
    tol = 1e-4 * np.abs(P) + 1e-7 + 1e-3 * dg / (g + 1e-8)
    bad = np.abs(got - P) > tol
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - P)[bad].max()), float(g[bad].min()), float(g[bad].max()))
    big = g > 1e-6
    np.testing.assert_allclose(got[big], P[big], rtol=1e-4, atol=3e-7)


def test_full_catalog_ranking_properties(yelp, orc):
    import torch
    from selfrec_b200 import ops
    rng = np.random.default_rng(2)
    U, I = yelp.user_num, yelp.item_num
    ue = (rng.standard_normal((U, 64)) * 0.1).astype(np.float32)
    ie = (rng.standard_normal((I, 64)) * 0.1).astype(np.float32)
    tu, ti = torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda()
    rp, ri = yelp.rated_csr()
    users = np.arange(U, dtype=np.int32)
    ids, sc = ops.score_topk(tu, ti, users, rp, ri, 20)
    ids_c, sc_c = ids.cpu().numpy(), sc.cpu().numpy()
    assert (np.diff(sc_c, axis=1) <= 0).all()                      # score-descending
    assert (ids_c >= 0).all() and (ids_c < I).all()
    assert all(len(set(r)) == 20 for r in ids_c[::97])              # no duplicates
    # no rated item is ever recommended (every user has >= 20 unrated items here)
    rated_mask = np.zeros(I, dtype=bool)
    for q in range(0, U, 53):
        rated_mask[:] = False
        rated_mask[ri[rp[q]:rp[q + 1]]] = True
        assert not rated_mask[ids_c[q]].any()
    # bit-exact against the oracle on a sample of users
    sample = rng.choice(U, 96, replace=False).astype(np.int32)
    oi, os_ = orc.score_topk(ue, ie, sample, rp, ri, 20)
    assert np.array_equal(ids_c[sample], oi) and np.array_equal(sc_c[sample], os_)
    # idempotence: re-ranking only the winners (everything else masked) returns the same list
    for q in sample[:8]:
        keep = np.setdiff1d(np.arange(I, dtype=np.int32), ids_c[q])
        i2, s2 = ops.score_topk(tu, ti, [int(q)], np.array([0] * (q + 1) + [len(keep)] * (U - q), dtype=np.int32), keep, 20)
        assert np.array_equal(i2[0].cpu().numpy(), ids_c[q]) and np.array_equal(s2[0].cpu().numpy(), sc_c[q])
    # independent kernel path: dense score rows + row top-k
    sub = sample[:32]
    dense = ops.score_rows(tu, ti, sub)
    for r, q in enumerate(sub):
        dense[r, torch.from_numpy(ri[rp[q]:rp[q + 1]].astype(np.int64)).cuda()] = -10e8
    i3, s3 = ops.topk_rows(dense, 20)
    assert np.array_equal(i3.cpu().numpy(), ids_c[sub]) and np.array_equal(s3.cpu().numpy(), sc_c[sub])


def test_epoch_of_native_batches_trains(yelp, in_tmp_cwd):
    """A few hundred fused steps from the native sampler: loss decreases, parameters stay finite,
    negatives are never rated items (sampler.py:24-27 invariant at full size)."""
    import random
    import torch
    from selfrec_b200.engine import TrainEngine
    random.seed(3)
    torch.manual_seed(3)
    eng = TrainEngine("XSimGCL", yelp, 64, 3, 2048, 1e-3, 1e-4, eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1)
    rp, ri = yelp.rated_csr()
    rated = set(zip(np.repeat(np.arange(yelp.user_num), np.diff(rp)).tolist(), ri.tolist()))
    first = last = None
    for n, w in enumerate(eng.batches()):
        if n < 3:
            b = w[0]
            assert b == 2048 and not any((int(u), int(j)) in rated for u, j in zip(w[4:4 + b], w[4 + 2 * 2048:4 + 2 * 2048 + b]))
        eng.step(w)
        if n == 0:
            first = eng.losses.cpu().numpy().copy()
        if n == 200:
            break
    torch.cuda.synchronize()
    last = eng.losses.cpu().numpy()
    assert np.isfinite(eng.params.cpu().numpy()).all()
    assert last[0] < first[0]  # BPR loss goes down from log(2)
    assert abs(first[0] - np.log(2)) < 0.01


# ------------------------------------------------------------------------------------------
# the other configs of BASELINE.json at full shape: one step against the float64 oracle
# ------------------------------------------------------------------------------------------
def _words(u, i, j, cap):
    w = np.zeros(4 + 5 * cap, dtype=np.int32)
    uq, iq = np.unique(u), np.unique(i)
    w[0], w[1], w[2] = len(u), len(uq), len(iq)
    for s, arr in enumerate((u, i, j, uq, iq)):
        w[4 + s * cap:4 + s * cap + len(arr)] = arr
    return w


def _negatives(data, u, rng):
    rp, ri = data.rated_csr()
    return np.array([next(x for x in rng.integers(0, data.item_num, 64) if x not in ri[rp[uu]:rp[uu + 1]]) for uu in u], dtype=np.int32)


def _check_step(eng, out, E0, orc, lr=1e-3):
    """losses to 1e-4; Adam's first moment m = 0.1 g is linear in the gradient, so it is the well-conditioned check
    of the whole backward; parameters where the gradient is not within rounding of zero."""
    import torch
    torch.cuda.synchronize()
    los = eng.losses.cpu().numpy()
    for got, want in ((los[0], out["rec"]), (los[1], out["l2"]), (los[2], out["cl"])):
        assert abs(got - want) <= 1e-4 * abs(want) + 1e-12, (got, want)
    g = out["grad"]
    m = eng.m.cpu().numpy()
    assert np.abs(m - 0.1 * g).max() <= 1e-4 * np.abs(0.1 * g).max()
    big = np.abs(g) > 1e-3 * np.abs(g).max()
    P, _, _ = orc.adam_step(E0, g.astype(np.float32), np.zeros_like(E0), np.zeros_like(E0), 1, lr)
    np.testing.assert_allclose(eng.params.cpu().numpy()[big], P[big], rtol=1e-4, atol=2e-6)


def test_lightgcn_step_yelp_shape_vs_oracle(yelp, orc, in_tmp_cwd):
    """configs[1]: LightGCN, yelp2018 shape, 3 layers, d=64, B=2048 (LightGCN.py:21-29)."""
    import torch
    from selfrec_b200.engine import TrainEngine
    torch.manual_seed(1)
    eng = TrainEngine("LightGCN", yelp, 64, 3, 2048, 1e-3, 1e-4, l2_div=2048.0)
    E0 = eng.params.cpu().numpy().copy()
    rng = np.random.default_rng(3)
    u, i = yelp.pair_users[5000:7048].copy(), yelp.pair_items[5000:7048].copy()
    j = _negatives(yelp, u, rng)
    eng.step(_words(u, i, j, 2048))
    out = orc.train_step("LightGCN", yelp.norm_adj.tocsr(), E0, eng.U, u, i, j, n_layers=3, reg=1e-4, batch_size=2048)
    _check_step(eng, out, E0, orc)


def test_sgl_step_kindle_shape_device_views_vs_oracle(orc, built_lib, in_tmp_cwd):
    """configs[3]: SGL edge-drop at amazon-kindle shape (138 333 x 98 572 x 1 525 091 + 2 822 duplicate lines -> 2.0
    entries, dropped graphs reset them to 1): the two view graphs are drawn like the reference draws them and BUILT ON
    THE DEVICE; the step (three encoders on three graphs, one InfoNCE over cat(users, items), SGL.py:30-41,98-125)
    must match the oracle run on the scipy-route view graphs."""
    import random
    import scipy.sparse as sp
    import torch
    from selfrec_b200 import synth
    from selfrec_b200.data.augmentor import GraphAugmentor, sample_range
    from selfrec_b200.data.device_graph import DeviceBipartite
    from selfrec_b200.engine import TrainEngine
    U, I, nnz = synth.SHAPES["amazon-kindle"]
    pu, pi = synth.make_pairs(U, I, nnz, seed=4)
    dup = np.random.default_rng(4).choice(nnz, 2822, replace=False)
    data = synth.ArrayInteraction(np.concatenate([pu, pu[dup]]), np.concatenate([pi, pi[dup]]), U, I)
    assert data.interaction_mat.nnz == nnz and data.interaction_mat.data.max() == 2.0
    bip = DeviceBipartite.from_interaction_mat(data.interaction_mat, "cuda")
    random.seed(21)
    host_views = [data.convert_to_laplacian_mat(GraphAugmentor.edge_dropout(data.interaction_mat, 0.1)) for _ in range(2)]
    random.seed(21)
    dev_views = [bip.assemble(keep_idx=sample_range(bip.nnz, int(bip.nnz * 0.9)), reset_weights=True) for _ in range(2)]
    for a, h in zip(dev_views, host_views):  # bit-identical graphs
        h = sp.csr_matrix(h)
        h.sort_indices()
        assert np.array_equal(a.rowptr.cpu().numpy(), h.indptr) and np.array_equal(a.colidx.cpu().numpy(), h.indices)
        assert np.array_equal(a.vals.cpu().numpy().view(np.uint32), h.data.astype(np.float32).view(np.uint32))
    torch.manual_seed(2)
    eng = TrainEngine("SGL", data, 64, 3, 2048, 1e-3, 1e-4, tau=0.2, cl_rate=0.1)
    eng.set_view_graphs(*dev_views)
    E0 = eng.params.cpu().numpy().copy()
    rng = np.random.default_rng(5)
    sel = rng.choice(len(pu), 2048, replace=False)
    u, i = pu[sel].copy(), pi[sel].copy()
    j = _negatives(data, u, rng)
    eng.step(_words(u, i, j, 2048))
    out = orc.train_step("SGL", data.norm_adj.tocsr(), E0, eng.U, u, i, j, n_layers=3, reg=1e-4, batch_size=2048, tau=0.2, cl_rate=0.1,
                         view_csr=[sp.csr_matrix(h) for h in host_views])
    _check_step(eng, out, E0, orc)


def test_simgcl_step_one_million_nodes_vs_oracle(orc, built_lib, in_tmp_cwd):
    """configs[4] recipe (Zipf 1.1, generated + normalised on the GPU, split rows) on a graph of 1.0 M nodes
    (800 k x 200 k x 16 M), SimGCL: three encoders, merged Horner chain (SimGCL.py:25-36,81-93), noise as an input."""
    import scipy.sparse as sp
    import torch
    from selfrec_b200 import synth
    from selfrec_b200.engine import TrainEngine
    data = synth.make_device_interaction((800_000, 200_000, 16_000_000), seed=3, alpha=1.1)
    adj = data.norm_adj
    assert adj.shape[0] == 1_000_000 and adj.n_huge > 0
    d, L, B = 64, 2, 2048
    torch.manual_seed(4)
    eng = TrainEngine("SimGCL", data, d, L, B, 1e-3, 1e-4, eps=0.1, tau=0.2, cl_rate=0.5)
    rng = np.random.default_rng(6)
    noise = rng.random((2, L, eng.N, d), dtype=np.float32)
    eng.set_noise_tensor(torch.from_numpy(noise).cuda())
    E0 = eng.params.cpu().numpy().copy()
    pu, pi = data.pair_users, data.pair_items
    sel = rng.choice(len(pu), B, replace=False)
    u, i = pu[sel].copy(), pi[sel].copy()
    j = _negatives(data, u, rng)
    eng.step(_words(u, i, j, B))
    A = sp.csr_matrix((adj.vals.cpu().numpy(), adj.colidx.cpu().numpy(), adj.rowptr.cpu().numpy()), shape=adj.shape)
    out = orc.train_step("SimGCL", A, E0, eng.U, u, i, j, n_layers=L, reg=1e-4, batch_size=B, eps=0.1, tau=0.2, cl_rate=0.5, noise=noise)
    _check_step(eng, out, E0, orc)
