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
