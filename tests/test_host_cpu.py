"""Host-side logic of the product, on CPU: the drop-in data/sampler/metrics modules against
the reference-generated golden fixtures, and the C-ABI library surface."""
import os
import random
import re

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _csr(g, prefix, shape):
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=shape)


def test_capi_exports_every_declared_symbol(built_lib):
    from selfrec_b200 import _lib
    header = open(os.path.join(ROOT, "include", "selfrec_b200.h")).read()
    declared = set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", header))
    declared -= {"srb_batch_words"}  # static inline helper
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(built_lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes binding"
    assert built_lib.srb_version() >= 100


def test_struct_layouts_match_header(built_lib):
    """ctypes mirrors must have the C sizes (compile a probe with gcc against the header)."""
    import subprocess, tempfile
    from selfrec_b200 import _lib
    src = r'''
#include <stdio.h>
#include "selfrec_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(srb_spmm_desc), sizeof(srb_encoder_desc), sizeof(srb_bpr_desc),
 sizeof(srb_infonce_problem), sizeof(srb_infonce_desc), sizeof(srb_topk_desc), sizeof(srb_step_desc), sizeof(srb_spmm_sharded_desc));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes as C
    mine = [C.sizeof(t) for t in (_lib.SpmmDesc, _lib.EncoderDesc, _lib.BprDesc, _lib.InfoNceProblem, _lib.InfoNceDesc,
                                  _lib.TopkDesc, _lib.StepDesc, _lib.SpmmShardedDesc)]
    assert mine == sizes


def test_no_cpu_fallback(built_lib):
    """Without a GPU every device entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from selfrec_b200 import _lib, ops
    with pytest.raises(_lib.SrbError):
        _lib.require_device()
    with pytest.raises(_lib.SrbError):
        ops.bpr_loss(torch.zeros(4, 64), torch.zeros(4, 64), torch.zeros(4, 64))
    with pytest.raises(_lib.SrbError):
        ops.SparseAdj(sp.eye(4, format="csr")).cuda()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under selfrec_b200/ may import, load or call it
    (comments may mention it)."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle\.(py|c|so)|oracle/|torch_port|orc_[a-z_]+\s*\()")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "selfrec_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), f"{f} references the oracle"


def test_interaction_matches_reference(golden, tiny_triples, tiny_conf, in_tmp_cwd):
    from selfrec_b200.data.ui_graph import Interaction
    train, test = tiny_triples
    g = golden("graph.npz")
    d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    U, I = int(g["user_num"]), int(g["item_num"])
    assert (d.user_num, d.item_num) == (U, I)
    assert [d.id2user[k] for k in range(U)] == list(g["user_names"])
    assert [d.id2item[k] for k in range(I)] == list(g["item_names"])
    for mat, prefix, shape in ((d.ui_adj, "ui", (U + I,) * 2), (d.norm_adj, "norm", (U + I,) * 2), (d.interaction_mat, "im", (U, I))):
        m = sp.csr_matrix(mat)
        m.sort_indices()
        ref = _csr(g, prefix, shape)
        assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
        assert np.array_equal(m.data.astype(np.float32), ref.data)
    assert list(d.test_set) == list(g["test_users"])  # dict order = test-file first appearance
    assert [len(d.test_set[u]) for u in d.test_set] == list(g["test_sizes"])
    assert list(d.training_size()) + list(d.test_size()) == list(g["sizes"])
    assert "ghost" not in d.test_set
    u0 = g["user_names"][0]
    assert d.contain(u0, d.user_rated(u0)[0][0]) and not d.contain("nobody", "x")
    ptr, idx = d.rated_csr()
    assert len(ptr) == U + 1 and all(np.all(np.diff(idx[ptr[u]:ptr[u + 1]]) > 0) for u in range(U))
    # convert_to_laplacian_mat == normalising the re-embedded interaction matrix (ui_graph.py:58-65)
    lap = d.convert_to_laplacian_mat(d.interaction_mat)
    inter = sp.csr_matrix(d.interaction_mat)
    big = sp.bmat([[None, inter], [inter.T, None]], format="csr", dtype=np.float32)
    from selfrec_b200.data.graph import Graph
    assert abs(sp.csr_matrix(lap) - Graph.normalize_graph_mat(big)).max() == 0


def test_normalize_isolated_nodes_and_rectangular():
    from selfrec_b200.data.graph import Graph
    a = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]], dtype=np.float32))
    n = Graph.normalize_graph_mat(a).toarray()
    assert np.isfinite(n).all() and n[2].sum() == 0 and n[0, 1] == 1.0  # inf -> 0 (graph.py:15)
    r = sp.csr_matrix(np.array([[1, 1, 0, 0], [0, 0, 0, 0]], dtype=np.float32))
    assert np.allclose(Graph.normalize_graph_mat(r).toarray(), [[0.5, 0.5, 0, 0], [0, 0, 0, 0]])


def test_native_sampler_bit_exact_with_reference(built_lib, golden, tiny_triples, tiny_conf, in_tmp_cwd):
    from selfrec_b200.data.ui_graph import Interaction
    from selfrec_b200.util.sampler import next_batch_pairwise
    train, test = tiny_triples
    s = golden("sampler.npz")
    d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    random.seed(int(s["seed"]))
    for epoch in range(2):
        us, is_, js = [], [], []
        for u, i, j in next_batch_pairwise(d, 100):
            assert isinstance(u, list) and isinstance(u[0], int)
            us += u
            is_ += i
            js += j
        assert us == list(s[f"e{epoch}_u"]) and is_ == list(s[f"e{epoch}_i"]) and js == list(s[f"e{epoch}_j"])
    js = []
    for u, i, j in next_batch_pairwise(d, 64, n_negs=3):
        js += j
    assert js == list(s["n3_j"])
    # Python's global stream ends where the reference's would, and training_data was permuted in place
    assert np.array_equal(np.array(random.getstate()[1], dtype=np.uint32), s["final_state"])
    assert [d.user[p[0]] for p in d.training_data] == list(s["final_order_users"])
    assert [d.item[p[1]] for p in d.training_data] == list(s["final_order_items"])


def test_native_sampler_fixed_layout_and_unique(built_lib, golden, tiny_triples, tiny_conf, in_tmp_cwd):
    """srb_sampler_next_batch: header + sections, sorted-unique lists (torch.unique, XSimGCL.py:46-47)."""
    from selfrec_b200.data.ui_graph import Interaction
    from selfrec_b200.util.sampler import NativePairSampler
    train, test = tiny_triples
    s = golden("sampler.npz")
    d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    ns = NativePairSampler(d)
    random.seed(int(s["seed"]))
    ns.pull_state()
    ns.begin_epoch()
    cap = 128
    allb = ns.epoch(100, cap)
    ns.push_state()
    assert allb.shape == (len(s["e0_sizes"]), 4 + 5 * cap)
    u = np.concatenate([b[4:4 + b[0]] for b in allb])
    j = np.concatenate([b[4 + 2 * cap:4 + 2 * cap + b[0]] for b in allb])
    assert np.array_equal(u, s["e0_u"]) and np.array_equal(j, s["e0_j"])
    for b in allb:
        n, nu, ni = b[0], b[1], b[2]
        assert np.array_equal(b[4 + 3 * cap:4 + 3 * cap + nu], np.unique(b[4:4 + n]))
        assert np.array_equal(b[4 + 4 * cap:4 + 4 * cap + ni], np.unique(b[4 + cap:4 + cap + n]))
    assert [int(b[0]) for b in allb] == list(s["e0_sizes"])


def test_sampler_rejects_saturated_user(built_lib):
    from selfrec_b200 import _lib
    import ctypes as C
    u = np.array([0, 0], dtype=np.int32)
    i = np.array([0, 1], dtype=np.int32)
    h = built_lib.srb_sampler_create(u.ctypes.data_as(_lib.c_i32p), i.ctypes.data_as(_lib.c_i32p), 2, 1, 2)
    assert h
    built_lib.srb_sampler_begin_epoch(h, None)
    out = np.zeros(4 + 5 * 4, dtype=np.int32)
    assert built_lib.srb_sampler_next_batch(h, 4, 4, out.ctypes.data_as(_lib.c_i32p)) == _lib.C.c_int(-3).value
    assert b"rated every item" in built_lib.srb_last_error()
    built_lib.srb_sampler_destroy(h)
    assert not built_lib.srb_sampler_create(None, None, 0, 1, 1)


def test_ranking_evaluation_format(golden, tiny_triples, tiny_conf, in_tmp_cwd):
    from selfrec_b200.data.ui_graph import Interaction
    from selfrec_b200.util.evaluation import ranking_evaluation
    train, test = tiny_triples
    r = golden("rank.npz")
    d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    rec = {u: list(zip(r["items"][k], r["scores"][k].tolist())) for k, u in enumerate(r["users"])}
    assert ranking_evaluation(d.test_set, rec, [5, 10]) == list(r["measure"])


def test_install_aliases_boundary_modules(built_lib):
    import sys
    import selfrec_b200
    names = selfrec_b200.install()
    try:
        from util.loss_torch import InfoNCE, bpr_loss, l2_reg_loss  # noqa: F401
        from util.sampler import next_batch_pairwise  # noqa: F401
        from base.torch_interface import TorchGraphInterface  # noqa: F401
        from base.graph_recommender import GraphRecommender  # noqa: F401
        from data.ui_graph import Interaction  # noqa: F401
        assert sys.modules["model.graph.XSimGCL"].XSimGCL.MODEL == "XSimGCL"
    finally:
        for n in names:
            sys.modules.pop(n, None)


def test_sparse_adj_routes_torch_sparse_mm(monkeypatch):
    """torch.sparse.mm(handle, X) must dispatch to our spmm via __torch_function__."""
    import torch
    from selfrec_b200 import ops
    called = {}
    monkeypatch.setattr(ops, "spmm", lambda a, x: called.setdefault("ok", (a, x)) and x)
    h = ops.SparseAdj(sp.eye(5, format="csr"))
    x = torch.ones(5, 64)
    out = torch.sparse.mm(h, x)
    assert called["ok"][0] is h and out is x
