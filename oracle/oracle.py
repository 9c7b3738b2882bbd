"""oracle.py -- CPU restatement (numpy + oracle.c) of the reference's hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under selfrec_b200/ does, and the
product path has no CPU fallback.

Parity is PINNED: tests/test_oracle_golden.py compares every function here with fixtures
produced by running the reference itself (oracle/gen_golden.py, torch 2.11.0 CPU, seeds
recorded in the fixture files).  Floating-point functions work in float64 (the centre of
every fp32 evaluation order); integer functions are bit-exact restatements.

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build():
    """Compile oracle.c with gcc (plain C, no CUDA)."""
    src = os.path.join(HERE, "oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_find_k_largest.restype = C.c_int64
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---------------------------------------------------------------------------------------
# R2: graph construction   data/ui_graph.py:47-71, data/graph.py:10-24
# ---------------------------------------------------------------------------------------
def normalize_graph_mat(adj):
    """graph.py:13-23 restated with the same scipy calls (diags().dot chains)."""
    import scipy.sparse as sp

    shape = adj.get_shape()
    rowsum = np.array(adj.sum(1))
    with np.errstate(divide="ignore"):
        if shape[0] == shape[1]:
            d_inv = np.power(rowsum, -0.5).flatten()
            d_inv[np.isinf(d_inv)] = 0.0
            d = sp.diags(d_inv)
            return d.dot(adj).dot(d)
        d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.0
        return sp.diags(d_inv).dot(adj)


def build_graph(pair_users, pair_items, n_users, n_items):
    """ui_adj / norm_adj / interaction_mat from id pairs (ui_graph.py:47-56, 67-71)."""
    import scipy.sparse as sp

    n = n_users + n_items
    ones = np.ones(len(pair_users), dtype=np.float32)
    half = sp.csr_matrix((ones, (np.asarray(pair_users), np.asarray(pair_items) + n_users)), shape=(n, n), dtype=np.float32)
    ui_adj = half + half.T
    inter = sp.csr_matrix((ones, (np.asarray(pair_users), np.asarray(pair_items))), shape=(n_users, n_items), dtype=np.float32)
    return ui_adj, normalize_graph_mat(ui_adj), inter


def assign_ids(triples):
    """First-appearance ids of users and items over the training triples (ui_graph.py:29-38)."""
    user, item = {}, {}
    pu, pi = [], []
    for u, i, _ in triples:
        if u not in user:
            user[u] = len(user)
        if i not in item:
            item[i] = len(item)
        pu.append(user[u])
        pi.append(item[i])
    return user, item, np.asarray(pu, dtype=np.int32), np.asarray(pi, dtype=np.int32)


# ---------------------------------------------------------------------------------------
# R3/R4: propagation and encoders
# ---------------------------------------------------------------------------------------
def spmm(csr, X, f32seq=False):
    """torch.sparse.mm(A, X)  LightGCN.py:72 (A given as scipy CSR)."""
    csr = csr.tocsr()
    X = np.ascontiguousarray(X, dtype=np.float32)
    rowptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
    colidx = np.ascontiguousarray(csr.indices, dtype=np.int32)
    vals = np.ascontiguousarray(csr.data, dtype=np.float32)
    Y = np.empty((csr.shape[0], X.shape[1]), dtype=np.float32)
    fn = lib().orc_spmm_csr_f32seq if f32seq else lib().orc_spmm_csr
    fn(_ptr(rowptr, C.c_int32), _ptr(colidx, C.c_int32), _ptr(vals, C.c_float), _ptr(X, C.c_float), _ptr(Y, C.c_float),
       C.c_int64(csr.shape[0]), C.c_int64(X.shape[1]))
    return Y


def _spmm64(csr, X):
    return csr.astype(np.float64) @ X


def perturb(E, noise, eps):
    """E + sign(E) * F.normalize(noise, dim=-1) * eps   XSimGCL.py:90-91 (F.normalize eps 1e-12)."""
    nrm = np.maximum(np.sqrt((noise.astype(np.float64) ** 2).sum(-1, keepdims=True)), 1e-12)
    return E + np.sign(E) * (noise / nrm) * eps


def encoder_forward(csr, E0, n_layers, include_ego, noise=None, eps=0.0, layer_cl=0):
    """LGCN_Encoder.forward LightGCN.py:68-78 / SGL_Encoder.forward SGL.py:98-113 (include_ego)
    SimGCL_Encoder.forward SimGCL.py:81-93 / XSimGCL_Encoder.forward XSimGCL.py:83-101.
    float64; returns (final, cl_view, layers)."""
    E = np.asarray(E0, dtype=np.float64)
    layers = [E] if include_ego else []
    cl = E
    for k in range(n_layers):
        E = _spmm64(csr, E)
        if noise is not None:
            E = perturb(E, np.asarray(noise[k], dtype=np.float64), eps)
        layers.append(E)
        if k == layer_cl - 1:
            cl = E
    final = np.mean(np.stack(layers, 1), 1)
    return final, cl, layers


# ---------------------------------------------------------------------------------------
# R6-R8: losses with analytic gradients (float64)
# ---------------------------------------------------------------------------------------
def bpr_loss(u, p, n):
    """util/loss_torch.py:6-10.  Returns (loss, du, dp, dn)."""
    u, p, n = (np.asarray(a, dtype=np.float64) for a in (u, p, n))
    x = (u * p).sum(1) - (u * n).sum(1)
    sig = 1.0 / (1.0 + np.exp(-x))
    loss = np.mean(-np.log(10e-6 + sig))
    c = (-(sig * (1 - sig)) / (10e-6 + sig)) / len(x)
    return loss, c[:, None] * (p - n), c[:, None] * u, -c[:, None] * u


def l2_reg_loss(reg, *embs):
    """util/loss_torch.py:18-22: reg * sum_e ||e||_F / e.shape[0].  Returns (loss, [grads])."""
    loss, grads = 0.0, []
    for e in embs:
        e = np.asarray(e, dtype=np.float64)
        nrm = np.sqrt((e ** 2).sum())
        loss += nrm / e.shape[0]
        grads.append(reg * e / (nrm * e.shape[0]) if nrm > 0 else np.zeros_like(e))
    return loss * reg, grads


def infonce(v1, v2, temperature, b_cos=True):
    """util/loss_torch.py:35-50.  Returns (loss, dv1, dv2)."""
    v1, v2 = np.asarray(v1, dtype=np.float64), np.asarray(v2, dtype=np.float64)
    if b_cos:
        n1 = np.maximum(np.sqrt((v1 ** 2).sum(1, keepdims=True)), 1e-12)
        n2 = np.maximum(np.sqrt((v2 ** 2).sum(1, keepdims=True)), 1e-12)
        a, b = v1 / n1, v2 / n2
    else:
        a, b = v1, v2
    S = a @ b.T / temperature
    m = S.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(S - m).sum(1))
    n = S.shape[0]
    loss = np.mean(lse - np.diag(S))
    G = (np.exp(S - lse[:, None]) - np.eye(n)) / n / temperature
    da, db = G @ b, G.T @ a
    if b_cos:
        da = (da - a * (a * da).sum(1, keepdims=True)) / n1
        db = (db - b * (b * db).sum(1, keepdims=True)) / n2
    return loss, da, db


# ---------------------------------------------------------------------------------------
# R10: torch.optim.Adam defaults (MF.py:15 ...), single-tensor fp32 arithmetic order
# ---------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    p, g, m, v = (np.asarray(a, dtype=np.float32) for a in (p, g, m, v))
    f = np.float32
    m = m + (g - m) * f(1 - beta1)
    v = v * f(beta2) + (f(1 - beta2) * g) * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


# ---------------------------------------------------------------------------------------
# whole training step (float64 autograd-style backward, layer by layer -- deliberately NOT
# the Horner form the CUDA engine uses)
# ---------------------------------------------------------------------------------------
def train_step(model, csr, E0, U, u_idx, i_idx, j_idx, *, n_layers, reg, batch_size, eps=0.0, tau=0.2, cl_rate=0.0,
               layer_cl=0, noise=None, view_csr=None):
    """One batch of <Model>.train(): MF.py:17-25, LightGCN.py:21-29, SimGCL.py:25-36,
    XSimGCL.py:27-37, SGL.py:30-41.  noise: [views, L, N, d].  Returns dict(rec, l2, cl, total, grad)."""
    E0 = np.asarray(E0, dtype=np.float64)
    N, d = E0.shape
    u_idx, i_idx, j_idx = (np.asarray(a, dtype=np.int64) for a in (u_idx, i_idx, j_idx))
    uu, ui = np.unique(u_idx), np.unique(i_idx)  # torch.unique: sorted (XSimGCL.py:46-47)
    A64 = csr.astype(np.float64) if csr is not None else None

    def enc(mat, include_ego, nz, lcl=0):
        return encoder_forward(mat, E0, n_layers, include_ego, nz, eps, lcl)

    def enc_backward(mat, include_ego, g_final, g_cl=None, lcl=0):
        # d final / d layer_k = 1/len(layers); propagate layer by layer with A^T (= A)
        cnt = n_layers + 1 if include_ego else n_layers
        g = np.zeros((N, d))
        for k in range(n_layers, 0, -1):
            g = g + g_final / cnt
            if g_cl is not None and lcl == k:
                g = g + g_cl
            g = mat.T @ g
        if include_ego:
            g = g + g_final / cnt
        if g_cl is not None and not (1 <= lcl <= n_layers):
            g = g + g_cl
        return g

    grad = np.zeros((N, d))
    cl = 0.0
    if model == "MF":
        final = E0
    elif model in ("LightGCN", "SGL"):
        final, _, _ = enc(A64, True, None)
    elif model == "SimGCL":
        final, _, _ = enc(A64, False, None)
    elif model == "XSimGCL":
        final, clv, _ = enc(A64, False, noise[0], layer_cl)
    else:
        raise ValueError(model)
    ue, pe, ne = final[u_idx], final[U + i_idx], final[U + j_idx]
    rec, du, dp, dn = bpr_loss(ue, pe, ne)
    g_final = np.zeros((N, d))
    np.add.at(g_final, u_idx, du)
    np.add.at(g_final, U + i_idx, dp)
    np.add.at(g_final, U + j_idx, dn)
    g_ego = np.zeros((N, d))
    if model == "MF":
        l2, gl = l2_reg_loss(reg, ue, pe, ne)
        l2 /= batch_size
        for idx, gg in zip((u_idx, U + i_idx, U + j_idx), gl):
            np.add.at(g_final, idx, gg / batch_size)
    elif model == "LightGCN":
        l2, gl = l2_reg_loss(reg, E0[u_idx], E0[U + i_idx], E0[U + j_idx])  # raw params, LightGCN.py:25
        l2 /= batch_size
        for idx, gg in zip((u_idx, U + i_idx, U + j_idx), gl):
            np.add.at(g_ego, idx, gg / batch_size)
    elif model in ("SimGCL", "XSimGCL"):
        l2, gl = l2_reg_loss(reg, ue, pe)
        for idx, gg in zip((u_idx, U + i_idx), gl):
            np.add.at(g_final, idx, gg)
    else:  # SGL.py:36
        l2, gl = l2_reg_loss(reg, ue, pe, ne)
        for idx, gg in zip((u_idx, U + i_idx, U + j_idx), gl):
            np.add.at(g_final, idx, gg)

    if model == "MF":
        grad = g_final
    elif model == "LightGCN":
        grad = enc_backward(A64, True, g_final) + g_ego
    elif model == "XSimGCL":
        lu, d1u, d2u = infonce(final[uu], clv[uu], tau)
        li, d1i, d2i = infonce(final[U + ui], clv[U + ui], tau)
        cl = cl_rate * (lu + li)
        g_cl = np.zeros((N, d))
        np.add.at(g_final, uu, cl_rate * d1u)
        np.add.at(g_final, U + ui, cl_rate * d1i)
        np.add.at(g_cl, uu, cl_rate * d2u)
        np.add.at(g_cl, U + ui, cl_rate * d2i)
        grad = enc_backward(A64, False, g_final, g_cl, layer_cl)
    elif model == "SimGCL":
        v1, _, _ = enc(A64, False, noise[0])
        v2, _, _ = enc(A64, False, noise[1])
        lu, d1u, d2u = infonce(v1[uu], v2[uu], tau)
        li, d1i, d2i = infonce(v1[U + ui], v2[U + ui], tau)
        cl = cl_rate * (lu + li)
        g1, g2 = np.zeros((N, d)), np.zeros((N, d))
        np.add.at(g1, uu, cl_rate * d1u)
        np.add.at(g1, U + ui, cl_rate * d1i)
        np.add.at(g2, uu, cl_rate * d2u)
        np.add.at(g2, U + ui, cl_rate * d2i)
        grad = enc_backward(A64, False, g_final) + enc_backward(A64, False, g1) + enc_backward(A64, False, g2)
    elif model == "SGL":
        m1, m2 = (m.astype(np.float64) for m in view_csr)
        v1, _, _ = enc(m1, True, None)
        v2, _, _ = enc(m2, True, None)
        cat = np.concatenate([uu, U + ui])  # SGL.py:120-121
        lc, d1, d2 = infonce(v1[cat], v2[cat], tau)
        cl = cl_rate * lc
        g1, g2 = np.zeros((N, d)), np.zeros((N, d))
        np.add.at(g1, cat, cl_rate * d1)
        np.add.at(g2, cat, cl_rate * d2)
        grad = enc_backward(A64, True, g_final) + enc_backward(m1, True, g1) + enc_backward(m2, True, g2)
    return dict(rec=rec, l2=l2, cl=cl, total=rec + l2 + cl, grad=grad, final=final)


# ---------------------------------------------------------------------------------------
# R1: sampler (CPython MT19937 stream)   util/sampler.py:5-28
# ---------------------------------------------------------------------------------------
def mt_state_from_python(rng_state):
    return np.asarray(rng_state[1], dtype=np.uint32).copy()


def shuffle_order(state625, n):
    """Permutation random.shuffle would apply to a list of n elements; advances state625."""
    order = np.arange(n, dtype=np.int64)
    lib().orc_shuffle(_ptr(state625, C.c_uint32), _ptr(order, C.c_int64), C.c_int64(n))
    return order


def sample_negatives(state625, users, n_items, rated_ptr, rated_idx, n_negs=1):
    users = np.ascontiguousarray(users, dtype=np.int32)
    rp = np.ascontiguousarray(rated_ptr, dtype=np.int64)
    ri = np.ascontiguousarray(rated_idx, dtype=np.int32)
    out = np.empty(len(users) * n_negs, dtype=np.int32)
    lib().orc_sample_negatives(_ptr(state625, C.c_uint32), _ptr(users, C.c_int32), C.c_int64(len(users)), C.c_int32(n_negs),
                               C.c_int32(n_items), _ptr(rp, C.c_int64), _ptr(ri, C.c_int32), _ptr(out, C.c_int32))
    return out


def next_batch_pairwise(state625, pair_users, pair_items, n_items, rated_ptr, rated_idx, batch_size, n_negs=1):
    """Generator restating sampler.py:5-28 over id arrays.  Mutates pair_users/pair_items
    in place (the persistent shuffle) and state625 (the MT19937 stream)."""
    n = len(pair_users)
    order = shuffle_order(state625, n)
    pair_users[:] = pair_users[order]
    pair_items[:] = pair_items[order]
    ptr = 0
    while ptr < n:
        end = ptr + batch_size if ptr + batch_size < n else n
        u = pair_users[ptr:end].copy()
        i = pair_items[ptr:end].copy()
        j = sample_negatives(state625, u, n_items, rated_ptr, rated_idx, n_negs)
        ptr = end
        yield u, i, j


# ---------------------------------------------------------------------------------------
# R9: ranking   base/graph_recommender.py:38-58, util/algorithm.py:144-156
# ---------------------------------------------------------------------------------------
def find_k_largest(K, candidates):
    cand = np.ascontiguousarray(candidates, dtype=np.float32)
    k = min(K, len(cand))
    ids = np.empty(k, dtype=np.int64)
    sc = np.empty(k, dtype=np.float32)
    lib().orc_find_k_largest(C.c_int64(K), _ptr(cand, C.c_float), C.c_int64(len(cand)), _ptr(ids, C.c_int64), _ptr(sc, C.c_float))
    return ids, sc


def score_topk(user_emb, item_emb, users, rated_ptr, rated_idx, K, want_scores=False):
    ue = np.ascontiguousarray(user_emb, dtype=np.float32)
    ie = np.ascontiguousarray(item_emb, dtype=np.float32)
    users = np.ascontiguousarray(users, dtype=np.int32)
    n_q, n_items, d = len(users), ie.shape[0], ie.shape[1]
    ids = np.empty((n_q, K), dtype=np.int64)
    sc = np.empty((n_q, K), dtype=np.float32)
    full = np.empty((n_q, n_items), dtype=np.float32) if want_scores else None
    rp = np.ascontiguousarray(rated_ptr, dtype=np.int32) if rated_ptr is not None else None
    ri = np.ascontiguousarray(rated_idx, dtype=np.int32) if rated_ptr is not None else None
    lib().orc_score_topk(_ptr(ue, C.c_float), _ptr(ie, C.c_float), C.c_int64(n_items), C.c_int64(d), _ptr(users, C.c_int32),
                         C.c_int64(n_q), _ptr(rp, C.c_int32) if rp is not None else None,
                         _ptr(ri, C.c_int32) if ri is not None else None, C.c_int64(K), _ptr(ids, C.c_int64), _ptr(sc, C.c_float),
                         _ptr(full, C.c_float) if full is not None else None)
    return (ids, sc, full) if want_scores else (ids, sc)
