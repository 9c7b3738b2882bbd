"""Host-side logic of the product, on CPU: the drop-in data/sampler/metrics modules against
the reference-generated golden fixtures, and the C-ABI library surface."""
import os
import random
import re

import numpy as np
import pytest
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

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
 sizeof(srb_infonce_problem), sizeof(srb_infonce_desc), sizeof(srb_topk_desc), sizeof(srb_step_desc), sizeof(srb_shard_desc));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes as C
    mine = [C.sizeof(t) for t in (_lib.SpmmDesc, _lib.EncoderDesc, _lib.BprDesc, _lib.InfoNceProblem, _lib.InfoNceDesc,
                                  _lib.TopkDesc, _lib.StepDesc, _lib.ShardDesc)]
    assert mine == sizes
    # the descriptors of the sharded step and the fields added last (a mirror that is one field short still has the
    # right size when padding absorbs it, so the tail offsets are compared too)
    src2 = r'''
#include <stdio.h>
#include <stddef.h>
#include "selfrec_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(srb_shard_desc), sizeof(srb_graph_csr), sizeof(srb_hub_split),
 sizeof(srb_shard_layout), offsetof(srb_shard_desc, nvls), offsetof(srb_shard_desc, fork_stream), offsetof(srb_shard_desc, Rt),
 offsetof(srb_step_desc, fork_stream), offsetof(srb_encoder_desc, x1));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe2.c")
        open(c, "w").write(src2)
        exe = os.path.join(td, "probe2")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        want = [int(x) for x in subprocess.check_output([exe]).split()]
    got = [C.sizeof(_lib.ShardDesc), C.sizeof(_lib.GraphCsr), C.sizeof(_lib.HubSplit), C.sizeof(_lib.ShardLayout), _lib.ShardDesc.nvls.offset,
           _lib.ShardDesc.fork_stream.offset, _lib.ShardDesc.Rt.offset, _lib.StepDesc.fork_stream.offset, _lib.EncoderDesc.x1.offset]
    assert got == want


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


@pytest.mark.parametrize("route", ["python_lists", "native_builder"])
def test_native_sampler_bit_exact_with_reference(built_lib, golden, tiny_triples, tiny_conf, in_tmp_cwd, route):
    from selfrec_b200.data.loader import FileIO
    from selfrec_b200.data.ui_graph import Interaction
    from selfrec_b200.util.sampler import next_batch_pairwise
    train, test = tiny_triples
    s = golden("sampler.npz")
    if route == "python_lists":
        d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    else:  # the object the native dataset builder returns must drive the sampler identically
        d = Interaction(tiny_conf("MF"), FileIO.load_data_set(os.path.join(GOLDEN, "tiny_train.txt")),
                        FileIO.load_data_set(os.path.join(GOLDEN, "tiny_test.txt")))
        assert type(d) is not Interaction
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


def test_ranking_evaluation_from_masks_matches_reference_strings(golden, tiny_triples, tiny_conf, in_tmp_cwd):
    """The id-space metric path (hit masks -> reference float expressions) reproduces the reference's
    ranking_evaluation() output on the golden full test() run, string for string."""
    from selfrec_b200.data.ui_graph import Interaction
    from selfrec_b200.util.evaluation import ranking_evaluation_from_masks
    train, test = tiny_triples
    r = golden("rank.npz")
    d = Interaction(tiny_conf("MF"), [list(t) for t in train], [list(t) for t in test])
    test_ptr, test_idx, n_test = d.test_csr()
    uids = np.array([d.user[u] for u in r["users"]], dtype=np.int32)
    assert list(r["users"]) == list(d.test_set)
    masks = []
    for k, u in enumerate(uids):
        mine = set(test_idx[test_ptr[u]:test_ptr[u + 1]].tolist())
        masks.append(sum(1 << rk for rk, name in enumerate(r["items"][k]) if d.item[name] in mine))
    for u, name in zip(uids, r["users"]):
        assert n_test[u] == len(d.test_set[name])
    assert ranking_evaluation_from_masks(n_test[uids], masks, [5, 10]) == list(r["measure"])


def _same_csr(x, y):
    x, y = x.tocsr().copy(), y.tocsr().copy()
    x.sort_indices()
    y.sort_indices()
    return (x.shape == y.shape and (x.indptr == y.indptr).all() and (x.indices == y.indices).all()
            and (x.data.view(np.uint32) == y.data.view(np.uint32)).all())


def _assert_same_interaction(a, b):
    assert a.user == b.user and a.item == b.item and a.id2user == b.id2user and a.id2item == b.id2item
    assert (a.pair_users == b.pair_users).all() and (a.pair_items == b.pair_items).all()
    assert _same_csr(a.norm_adj, b.norm_adj) and _same_csr(a.ui_adj, b.ui_adj) and _same_csr(a.interaction_mat, b.interaction_mat)
    assert dict(a.test_set) == dict(b.test_set) and list(a.test_set) == list(b.test_set) and a.test_set_item == b.test_set_item
    assert a.training_size() == b.training_size() and a.test_size() == b.test_size()
    assert a.training_data == b.training_data
    assert dict(a.training_set_u) == dict(b.training_set_u) and dict(a.training_set_i) == dict(b.training_set_i)
    for x, y in zip(a.rated_csr(), b.rated_csr()):
        assert (x == y).all()


def test_native_dataset_builder_matches_python_interaction(built_lib, golden, tmp_path):
    """srb_dataset_* (C++) against the Python restatement of loader + Interaction (itself pinned to the
    reference by graph.npz): ids, pairs, all three matrices bit for bit, test filtering, sizes."""
    from selfrec_b200.data.loader import FileIO
    from selfrec_b200.data.native import load_interaction
    from selfrec_b200.data.ui_graph import Interaction
    tr, te = os.path.join(GOLDEN, "tiny_train.txt"), os.path.join(GOLDEN, "tiny_test.txt")
    a = load_interaction(None, tr, te)
    py = Interaction(None, list(FileIO.load_data_set(tr)), list(FileIO.load_data_set(te)))  # plain lists: the Python route
    assert type(py) is Interaction and type(a) is not Interaction
    _assert_same_interaction(a, py)
    g = golden("graph.npz")  # the reference's own matrices
    ref = sp.csr_matrix((g["norm_data"], g["norm_indices"], g["norm_indptr"]), shape=a.norm_adj.shape)
    assert _same_csr(a.norm_adj, ref)
    # a harder file: duplicates, hubs, users/items unseen in training, trailing blanks, \r\n, no final newline
    rng = np.random.default_rng(3)
    n_u, n_i = 400, 300
    lines = [f"user{rng.integers(n_u)} it{int(rng.zipf(1.3)) % n_i} {rng.integers(1, 6)}" for _ in range(20000)]
    lines[5] += "   "
    lines[7] = "  " + lines[7] + "\r"
    lines[11] += " extra column"
    tr2, te2 = tmp_path / "train.txt", tmp_path / "test.txt"
    tr2.write_text("\n".join(lines))
    tl = [f"user{rng.integers(n_u + 50)} it{rng.integers(n_i + 80)} 1.5" for _ in range(3000)]
    te2.write_text("\n".join(tl) + "\n")
    a2 = load_interaction(None, tr2, te2)
    _assert_same_interaction(a2, Interaction(None, list(FileIO.load_data_set(str(tr2))), list(FileIO.load_data_set(str(te2)))))
    # FileIO's return value keeps its path, so the reference's own call sequence lands in the native builder
    via_fileio = Interaction(None, FileIO.load_data_set(str(tr2)), FileIO.load_data_set(str(te2)))
    assert type(via_fileio) is not Interaction
    _assert_same_interaction(via_fileio, a2)
    assert a2.ui_adj.data.max() > 1.0  # duplicate lines were summed
    # no test file
    a3 = load_interaction(None, tr2)
    assert len(a3.test_set) == 0 and a3.test_size() == (0, 0, 0) and _same_csr(a3.norm_adj, a2.norm_adj)


def test_native_dataset_builder_errors(built_lib, tmp_path):
    from selfrec_b200 import _lib
    from selfrec_b200.data.native import load_interaction
    with pytest.raises(_lib.SrbError, match="cannot open"):
        load_interaction(None, tmp_path / "missing.txt")
    bad = tmp_path / "bad.txt"
    bad.write_text("u1 i1 1\nu2 i2\n")
    with pytest.raises(_lib.SrbError, match="line 2"):
        load_interaction(None, bad)
    bad.write_text("u1 i1 abc\n")
    with pytest.raises(_lib.SrbError, match="line 1"):
        load_interaction(None, bad)
    bad.write_text("u1 i1 1\n\nu2 i2 1\n")  # the reference dies on an empty line (IndexError)
    with pytest.raises(_lib.SrbError, match="line 2"):
        load_interaction(None, bad)


@pytest.mark.parametrize("n,k", [(10, 0), (10, 10), (10, 3), (100, 6), (100, 90), (5000, 4500), (5000, 7), (22, 21), (21, 5), (87, 6),
                                 (300000, 270000), (300000, 50)])
def test_native_random_sample_is_cpython_exact(built_lib, n, k):
    """srb_random_sample_range == random.sample(range(n), k): same draws in the same order (both CPython
    strategies: pool list and selected set) and the same generator state afterwards."""
    from selfrec_b200.data.augmentor import sample_range
    random.seed(n * 7 + k)
    want = random.sample(range(n), k)
    st_want = random.getstate()
    random.seed(n * 7 + k)
    got = sample_range(n, k)
    assert got.tolist() == want and random.getstate() == st_want
    with pytest.raises(ValueError):
        sample_range(5, 6)


def test_sgl_views_match_reference_graphs(built_lib, golden, tiny_triples, tiny_conf, in_tmp_cwd):
    """R11: the two edge-dropped, re-normalised graphs SGL builds for an epoch (SGL.py:80-96,
    data/augmentor.py:23-32, ui_graph.py:58-65) equal the reference's own, bit for bit, when Python's
    `random` starts from the same seed -- through the native sampler of random.sample."""
    from selfrec_b200.data.augmentor import GraphAugmentor
    from selfrec_b200.data.ui_graph import Interaction
    train, test = tiny_triples
    fx = golden("train_SGL.npz")
    d = Interaction(tiny_conf("SGL"), [list(t) for t in train], [list(t) for t in test])
    random.seed(1000 + len("SGL"))  # oracle/gen_golden.py seeds the reference run this way
    n = d.user_num + d.item_num
    for k in range(2):
        dropped = GraphAugmentor.edge_dropout(d.interaction_mat, 0.1)
        lap = d.convert_to_laplacian_mat(dropped)
        ref = sp.csr_matrix((fx[f"view{k}_data"], fx[f"view{k}_indices"], fx[f"view{k}_indptr"]), shape=(n, n))
        assert _same_csr(lap, ref), k


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every ctypes.Structure in _lib.py must have the size and field offsets of the C struct it mirrors in
    include/selfrec_b200.h (a drifted field would silently shift every pointer after it): gcc prints
    sizeof / offsetof for each field, ctypes must agree."""
    import ctypes as C
    import shutil
    import subprocess
    from selfrec_b200 import _lib
    pairs = {"srb_spmm_desc": _lib.SpmmDesc, "srb_encoder_desc": _lib.EncoderDesc, "srb_scatter_seg": _lib.ScatterSeg,
             "srb_bpr_desc": _lib.BprDesc, "srb_infonce_problem": _lib.InfoNceProblem, "srb_infonce_desc": _lib.InfoNceDesc,
             "srb_topk_desc": _lib.TopkDesc, "srb_graph_csr": _lib.GraphCsr, "srb_step_desc": _lib.StepDesc,
             "srb_hub_split": _lib.HubSplit, "srb_shard_desc": _lib.ShardDesc, "srb_shard_layout": _lib.ShardLayout}
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "a C compiler is part of the toolchain"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "selfrec_b200.h"', 'int main(void) {']
    for cname, st in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _t in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(os.path.dirname(GOLDEN), "..", "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, st in pairs.items():
        assert int(out[cname]) == C.sizeof(st), cname
        for fname, _t in st._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_graph_augmentor_matches_reference_drops(built_lib, golden):
    """R11: node_dropout / edge_dropout against what the unmodified reference dropped for the same `random`
    seed (oracle/gen_golden_augment.py), including where the generator stands afterwards."""
    from selfrec_b200.data.augmentor import GraphAugmentor
    fx = golden("augment.npz")
    shape = tuple(int(x) for x in fx["in_shape"])
    mat = sp.csr_matrix((np.ones(len(fx["in_indices"]), np.float32), fx["in_indices"], fx["in_indptr"]), shape=shape)
    for case in fx["cases"]:
        kind, rate, seed, tag = str(case).split(":")
        random.seed(int(seed))
        got = sp.csr_matrix(getattr(GraphAugmentor, kind)(mat, float(rate)))
        got.sum_duplicates()
        got.eliminate_zeros()
        want = sp.csr_matrix((fx[tag + "_data"], fx[tag + "_indices"], fx[tag + "_indptr"]), shape=shape)
        assert _same_csr(got, want), case
        assert random.random() == float(fx[tag + "_next_random"][0]), case


def test_stream_epoch_equals_epoch_array(built_lib):
    """The streaming generator TrainEngine.batches() uses yields exactly the batches of srb_sampler_epoch for the
    same `random` state, leaves the same state behind (also when abandoned half way) and shuffles training_data."""
    from selfrec_b200 import synth
    from selfrec_b200.util.sampler import NativePairSampler, stream_epoch
    d1 = synth.make_interaction((300, 400, 6000), seed=2)
    d2 = synth.make_interaction((300, 400, 6000), seed=2)
    s1, s2 = NativePairSampler(d1), NativePairSampler(d2)
    for ep in range(2):
        random.seed(77 + ep)
        s1.pull_state()
        perm = s1.begin_epoch(want_perm=True)
        want = s1.epoch(128, 128)
        s1.push_state()
        st_want = random.getstate()
        random.seed(77 + ep)
        got = np.stack([w.copy() for w in stream_epoch(s2, d2, 128, 128)])
        assert random.getstate() == st_want and np.array_equal(got, want)
        first = list(d1.training_data)
        d1.shuffle_training_data(perm)
        assert d1.training_data == d2.training_data and d1.training_data != first
    # closing the generator early hands the state back at that point of the stream.  (Fresh samplers: a sampler
    # keeps its pairs in the order the previous epoch left them, like the reference's list.)
    def fresh():
        d = synth.make_interaction((300, 400, 6000), seed=2)
        return NativePairSampler(d), d

    sm, dd = fresh()
    random.seed(5)
    g = stream_epoch(sm, dd, 128, 128)
    for _ in range(3):
        next(g)
    g.close()
    after_three = random.getstate()
    sm, _ = fresh()
    random.seed(5)
    sm.pull_state()
    sm.begin_epoch(want_perm=False)
    buf = np.empty(4 + 5 * 128, dtype=np.int32)
    for _ in range(3):
        sm.next_batch(128, 128, buf)
    sm.push_state()
    assert random.getstate() == after_three


def test_lazy_training_data_shuffles_compose(built_lib):
    """The sampler records its epoch shuffles as a pending permutation; reading data.training_data applies them
    to the same list object, in order (util/sampler.py:7 shuffles that list in place every epoch)."""
    from selfrec_b200 import synth
    from selfrec_b200.data.native import load_interaction
    rng = np.random.default_rng(0)
    d = synth.make_interaction((50, 60, 400), seed=0)
    objs = [d, load_interaction(None, os.path.join(GOLDEN, "tiny_train.txt"))]
    for obj in objs:
        n = len(obj.pair_users)
        first = list(obj.training_data) if obj is d else None  # the native object builds its list on first access
        perms = [rng.permutation(n) for _ in range(11)]  # more than the fold-into-one threshold
        for p in perms:
            obj.shuffle_training_data(p)
        got = obj.training_data
        if first is None:
            un, inn = obj._unames, obj._inames
            first = [[un[u], inn[i], w] for u, i, w in zip(obj.pair_users.tolist(), obj.pair_items.tolist(), obj.pair_weights.tolist())]
        want = first
        for p in perms:
            want = [want[k] for k in p]
        assert got == want and obj.training_data is got  # applied once, same list afterwards


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


def test_sampler_ring_equals_sequential_calls(built_lib):
    """The sample-ahead ring (native producer thread) hands out exactly the batches -- and leaves exactly the MT19937
    state -- of one srb_sampler_next_batch call per batch, for several depths incl. a short last batch."""
    import random
    from selfrec_b200 import _lib, synth
    from selfrec_b200.util.sampler import NativePairSampler, stream_epoch
    data = synth.make_interaction((300, 400, 5000), seed=9)
    B = 512
    ref_batches, ref_state = None, None
    for depth in (0, 1, 3, 16):
        random.seed(123)
        smp = NativePairSampler(data)
        got = [w.copy() for w in stream_epoch(smp, data, B, B, ring_depth=depth)]
        state = random.getstate()
        assert len(got) == -(-5000 // B) and got[-1][0] == 5000 % B
        if ref_batches is None:
            ref_batches, ref_state = got, state
        else:
            assert all(np.array_equal(a, b) for a, b in zip(got, ref_batches)) and state == ref_state
    # a ring stopped early joins cleanly and can be restarted in the next epoch
    random.seed(5)
    smp = NativePairSampler(data)
    gen = stream_epoch(smp, data, B, B, ring_depth=4)
    next(gen)
    gen.close()
    assert len(list(stream_epoch(smp, data, B, B, ring_depth=4))) == -(-5000 // B)


def test_abandoned_epoch_generator_does_not_disturb_the_next_one(built_lib):
    """zip(range(n), engine.batches()) leaves the epoch generator un-closed: starting the next epoch must retire it
    (stop its ring, hand the state back at the point it was read to) and its late `finally` must be a no-op."""
    import gc
    import random
    from selfrec_b200 import synth
    from selfrec_b200.util.sampler import NativePairSampler, stream_epoch
    data = synth.make_interaction((300, 400, 5000), seed=9)
    B = 256
    random.seed(42)
    smp = NativePairSampler(data)
    g1 = stream_epoch(smp, data, B, B)
    first = [w.copy() for _, w in zip(range(3), g1)]     # abandoned after 3 batches, not closed
    g2 = stream_epoch(smp, data, B, B)
    second = [w.copy() for w in g2]
    state = random.getstate()
    del g1
    gc.collect()
    assert random.getstate() == state                     # the stale generator's finally changed nothing
    # reference: the same sequence with explicit close()
    random.seed(42)
    smp2 = NativePairSampler(synth.make_interaction((300, 400, 5000), seed=9))
    h1 = stream_epoch(smp2, data, B, B)
    f2 = [w.copy() for _, w in zip(range(3), h1)]
    h1.close()
    s2 = [w.copy() for w in stream_epoch(smp2, data, B, B)]
    assert all(np.array_equal(a, b) for a, b in zip(first + second, f2 + s2)) and len(second) == len(s2)
    assert random.getstate() == state
