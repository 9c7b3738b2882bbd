"""Pins the oracle (oracle/oracle.py + oracle.c) against fixtures produced by the reference
itself (oracle/gen_golden.py).  CPU only."""
import random

import numpy as np
import scipy.sparse as sp


def _csr(g, prefix, shape):
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=shape)


def _graph(golden):
    g = golden("graph.npz")
    U, I = int(g["user_num"]), int(g["item_num"])
    return g, U, I, _csr(g, "norm", (U + I, U + I))


def test_graph_construction_bit_exact(orc, golden, tiny_triples):
    train, _ = tiny_triples
    g, U, I, norm = _graph(golden)
    user, item, pu, pi = orc.assign_ids(train)
    assert len(user) == U and len(item) == I
    assert [k for k, _ in sorted(user.items(), key=lambda kv: kv[1])] == list(g["user_names"])
    assert [k for k, _ in sorted(item.items(), key=lambda kv: kv[1])] == list(g["item_names"])
    ui, na, im = orc.build_graph(pu, pi, U, I)
    for mat, prefix, shape in ((ui, "ui", (U + I, U + I)), (na, "norm", (U + I, U + I)), (im, "im", (U, I))):
        mat = sp.csr_matrix(mat)
        mat.sort_indices()
        ref = _csr(g, prefix, shape)
        assert np.array_equal(mat.indptr, ref.indptr) and np.array_equal(mat.indices, ref.indices)
        assert np.array_equal(mat.data.astype(np.float32), ref.data)  # bit-exact fp32
    assert im.data.max() == 2.0  # duplicate lines are summed (ui_graph.py:52-53)


def test_sampler_stream_bit_exact(orc, golden, tiny_triples):
    train, _ = tiny_triples
    s = golden("sampler.npz")
    g, U, I, _ = _graph(golden)
    _, _, pu, pi = orc.assign_ids(train)
    im = _csr(g, "im", (U, I))
    random.seed(int(s["seed"]))
    state = orc.mt_state_from_python(random.getstate())
    pu, pi = pu.copy(), pi.copy()
    for epoch in range(2):
        us, is_, js, sizes = [], [], [], []
        for u, i, j in orc.next_batch_pairwise(state, pu, pi, I, im.indptr, im.indices, 100):
            us.append(u), is_.append(i), js.append(j), sizes.append(len(u))
        assert np.array_equal(np.concatenate(us), s[f"e{epoch}_u"])
        assert np.array_equal(np.concatenate(is_), s[f"e{epoch}_i"])
        assert np.array_equal(np.concatenate(js), s[f"e{epoch}_j"])
        assert sizes == list(s[f"e{epoch}_sizes"])  # short last batch (sampler.py:11-14)
    us, is_, js = [], [], []
    for u, i, j in orc.next_batch_pairwise(state, pu, pi, I, im.indptr, im.indices, 64, n_negs=3):
        us.append(u), is_.append(i), js.append(j)
    assert np.array_equal(np.concatenate(js), s["n3_j"])
    assert np.array_equal(state, s["final_state"])  # the MT19937 stream position matches too
    assert np.array_equal(pu, s["final_order_users"]) and np.array_equal(pi, s["final_order_items"])


def test_losses_match_reference_autograd(orc, golden):
    lo = golden("losses.npz")
    for tag in ("a", "b"):
        u, p, n = lo[f"bpr_{tag}_u"], lo[f"bpr_{tag}_p"], lo[f"bpr_{tag}_n"]
        loss, gu, gp, gn = orc.bpr_loss(u, p, n)
        assert abs(loss - lo[f"bpr_{tag}_loss"]) <= 2e-6 * abs(loss)
        for mine, ref in ((gu, "gu"), (gp, "gp"), (gn, "gn")):
            np.testing.assert_allclose(mine, lo[f"bpr_{tag}_{ref}"], rtol=2e-5, atol=1e-8)
        l2, grads = orc.l2_reg_loss(1e-2, u, p, n)
        assert abs(l2 - lo[f"l2_{tag}_loss"]) <= 2e-6 * abs(l2)
        for mine, ref in zip(grads, ("gu", "gp", "gn")):
            np.testing.assert_allclose(mine, lo[f"l2_{tag}_{ref}"], rtol=2e-5, atol=1e-10)
    for tag in ("a", "b", "c", "d"):
        loss, g1, g2 = orc.infonce(lo[f"nce_{tag}_v1"], lo[f"nce_{tag}_v2"], float(lo[f"nce_{tag}_tau"]), bool(lo[f"nce_{tag}_cos"]))
        assert abs(loss - lo[f"nce_{tag}_loss"]) <= 3e-6 * max(abs(loss), 1e-3)
        scale = max(np.abs(lo[f"nce_{tag}_g1"]).max(), 1e-12)
        np.testing.assert_allclose(g1, lo[f"nce_{tag}_g1"], rtol=1e-4, atol=2e-6 * scale)
        np.testing.assert_allclose(g2, lo[f"nce_{tag}_g2"], rtol=1e-4, atol=2e-6 * scale)


MODELS = {
    "MF": dict(n_layers=0),
    "LightGCN": dict(n_layers=3),
    "SimGCL": dict(n_layers=2, cl_rate=0.5, eps=0.1, tau=0.2),
    "XSimGCL": dict(n_layers=3, layer_cl=1, cl_rate=0.2, eps=0.2, tau=0.2),
    "SGL": dict(n_layers=2, cl_rate=0.1, tau=0.2),
}


def replay_noise(name, fx, L):
    """Golden noise records -> [steps][views, L, N, d] in the order the reference drew them."""
    if "noise" not in fx.files:
        return None
    nz = fx["noise"]
    views = 2 if name == "SimGCL" else 1
    per_step = views * L
    return [nz[k * per_step:(k + 1) * per_step].reshape(views, L, *nz.shape[1:]) for k in range(int(fx["n_steps"]))]


def test_train_steps_match_reference(orc, golden):
    g, U, I, norm = _graph(golden)
    for name, kw in MODELS.items():
        fx = golden(f"train_{name}.npz")
        L = kw["n_layers"]
        E = np.concatenate([fx["init_user"], fx["init_item"]]).astype(np.float32)
        if name != "MF":
            final, _, _ = orc.encoder_forward(norm, E, L, include_ego=name in ("LightGCN", "SGL"))
            np.testing.assert_allclose(final[:U], fx["clean_user"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(final[U:], fx["clean_item"], rtol=1e-5, atol=1e-7)
        noise = replay_noise(name, fx, L)
        views = None
        if name == "SGL":
            views = [_csr(fx, f"view{k}", (U + I, U + I)) for k in range(2)]
        m = np.zeros_like(E)
        v = np.zeros_like(E)
        tags, vals = list(fx["loss_tags"]), list(fx["loss_vals"])
        per = len(tags) // int(fx["n_steps"])
        for k in range(int(fx["n_steps"])):
            out = orc.train_step(name, norm if name != "MF" else None, E, U, fx[f"b{k}_u"], fx[f"b{k}_i"], fx[f"b{k}_j"],
                                 reg=float(fx["reg"]), batch_size=int(fx["batch_size"]), noise=None if noise is None else noise[k],
                                 view_csr=views, **kw)
            rec = dict()
            for t, val in zip(tags[k * per:(k + 1) * per], vals[k * per:(k + 1) * per]):
                rec.setdefault(t, []).append(val)
            assert abs(out["rec"] - rec["bpr_loss"][0]) <= 1e-5 * abs(out["rec"]), name
            div = float(fx["batch_size"]) if name in ("MF", "LightGCN") else 1.0  # the model divides after the call
            assert abs(out["l2"] - rec["l2_reg_loss"][0] / div) <= 1e-5 * abs(out["l2"]), name
            if "InfoNCE" in rec:
                assert abs(out["cl"] - kw["cl_rate"] * sum(rec["InfoNCE"])) <= 2e-5 * abs(out["cl"]), name
            E, m, v = orc.adam_step(E, out["grad"].astype(np.float32), m, v, k + 1, float(fx["lr"]))
            ref = fx[f"params_after_{k}"]
            # Adam's first steps move every weight by ~lr; 1e-4 relative on the embeddings is
            # the north-star tolerance, the oracle sits well inside it
            np.testing.assert_allclose(E, ref, rtol=2e-5, atol=2e-7, err_msg=f"{name} step {k}")


def test_find_k_largest_matches_numba_reference(orc, golden):
    tk = golden("topk.npz")
    for tag in ("rand", "ties", "survey", "const", "masked", "asc", "desc"):
        v = tk[f"{tag}_in"]
        for K in (3, 10, 20):
            ids, sc = orc.find_k_largest(K, v)
            assert np.array_equal(ids, tk[f"{tag}_K{K}_ids"]), (tag, K)  # ties included: heap + numba quicksort order
            assert np.array_equal(sc, tk[f"{tag}_K{K}_scores"])
    ids, _ = orc.find_k_largest(3, np.array([1, 3, 3, 3, 2, 3, 3, 0, 3, 3], dtype=np.float32))
    assert list(ids) == [1, 3, 2]  # SURVEY R9 measured example


def test_ranking_matches_reference_test(orc, golden):
    r = golden("rank.npz")
    g, U, I, _ = _graph(golden)
    im = _csr(g, "im", (U, I))
    uid = {n: k for k, n in enumerate(g["user_names"])}
    users = np.array([uid[u] for u in r["users"]], dtype=np.int32)
    ids, sc = orc.score_topk(r["user_emb"], r["item_emb"], users, im.indptr, im.indices, 10)
    names = g["item_names"][ids]
    assert np.array_equal(names, r["items"])  # bit-exact top-k item ids
    np.testing.assert_allclose(sc, r["scores"], rtol=2e-5, atol=1e-7)


def test_torch_cpu_port_matches_reference(golden):
    """oracle/torch_port.py (the CPU baseline / reference arm of bench.py) reproduces the
    reference's own XSimGCL train() parameters step for step."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import torch
    import torch_port
    g, U, I, norm = _graph(golden)
    fx = golden("train_XSimGCL.npz")
    m = torch_port.XSimGCLCpu(norm, U, I, 64, 3, 0.2, 0.2, 0.2, 1, 0.001, 0.0001, fx["init_user"], fx["init_item"])
    tags, vals = list(fx["loss_tags"]), list(fx["loss_vals"])
    for k in range(int(fx["n_steps"])):
        nz = [torch.from_numpy(fx["noise"][k * 3 + l]) for l in range(3)]
        rec, l2, cl = m.step(fx[f"b{k}_u"].tolist(), fx[f"b{k}_i"].tolist(), fx[f"b{k}_j"].tolist(), nz)
        assert abs(rec - vals[k * 4]) <= 1e-6 * abs(rec)
        got = torch.cat([m.ue, m.ie]).detach().numpy()
        np.testing.assert_allclose(got, fx[f"params_after_{k}"], rtol=1e-6, atol=1e-8)
