"""Drop-in for util/sampler.py:5-28 (next_batch_pairwise) on the native sampler.

Bit-exact with the reference given the same `random` state: the MT19937 state is pulled
from random.getstate() before every native call and pushed back afterwards, so the global
stream advances exactly as with the Python implementation; data.training_data is permuted
in place like random.shuffle would.
"""
import ctypes as C
import random

import numpy as np

from .. import _lib


class NativePairSampler:
    """Owns a srb_sampler for one Interaction object."""

    def __init__(self, data):
        lib = _lib.load()
        pu = np.ascontiguousarray(data.pair_users if hasattr(data, "pair_users") else [data.user[p[0]] for p in data.training_data], dtype=np.int32)
        pi = np.ascontiguousarray(data.pair_items if hasattr(data, "pair_items") else [data.item[p[1]] for p in data.training_data], dtype=np.int32)
        self._lib = lib
        self.n_pairs = len(pu)
        self.handle = lib.srb_sampler_create(pu.ctypes.data_as(_lib.c_i32p), pi.ctypes.data_as(_lib.c_i32p), self.n_pairs,
                                             int(data.user_num), int(data.item_num))
        if not self.handle:
            raise _lib.SrbError("srb_sampler_create failed: " + _lib.last_error())
        self._state = (C.c_uint32 * 625)()

    def __del__(self):
        if getattr(self, "handle", None):
            self._lib.srb_sampler_destroy(self.handle)
            self.handle = None

    def pull_state(self):
        st = random.getstate()
        self._gauss = st[2]
        self._state[:] = st[1]
        _lib.check(self._lib.srb_sampler_set_state(self.handle, self._state), "srb_sampler_set_state")

    def push_state(self):
        _lib.check(self._lib.srb_sampler_get_state(self.handle, self._state), "srb_sampler_get_state")
        random.setstate((3, tuple(self._state), self._gauss))

    def begin_epoch(self, want_perm=True):
        perm = np.empty(self.n_pairs, dtype=np.int64) if want_perm else None
        ptr = perm.ctypes.data_as(_lib.c_i64p) if want_perm else None
        _lib.check(self._lib.srb_sampler_begin_epoch(self.handle, ptr), "srb_sampler_begin_epoch")
        return perm

    def next_batch_negs(self, batch_size, n_negs, u, i, j):
        b = self._lib.srb_sampler_next_batch_negs(self.handle, batch_size, n_negs, u.ctypes.data_as(_lib.c_i32p),
                                                  i.ctypes.data_as(_lib.c_i32p), j.ctypes.data_as(_lib.c_i32p))
        if b < 0:
            _lib.check(b, "srb_sampler_next_batch_negs")
        return b

    def next_batch(self, batch_size, batch_cap, out):
        """Fixed layout (header + u,i,j,uniq_u,uniq_i) into the int32 array `out`."""
        b = self._lib.srb_sampler_next_batch(self.handle, batch_size, batch_cap, out.ctypes.data_as(_lib.c_i32p))
        if b < 0:
            _lib.check(b, "srb_sampler_next_batch")
        return b

    def ring_start(self, batch_size, batch_cap, depth=16):
        _lib.check(self._lib.srb_sampler_ring_start(self.handle, batch_size, batch_cap, depth), "srb_sampler_ring_start")

    def ring_pop(self, out):
        b = self._lib.srb_sampler_ring_pop(self.handle, out.ctypes.data_as(_lib.c_i32p))
        if b < 0:
            _lib.check(b, "srb_sampler_ring_pop")
        return b

    def ring_stop(self):
        _lib.check(self._lib.srb_sampler_ring_stop(self.handle), "srb_sampler_ring_stop")

    def epoch(self, batch_size, batch_cap):
        """All batches of one epoch (after begin_epoch): int32 array [n_batches, words]."""
        words = _lib.BATCH_HEADER + 5 * batch_cap
        nb = (self.n_pairs + batch_size - 1) // batch_size
        out = np.empty((nb, words), dtype=np.int32)
        got = self._lib.srb_sampler_epoch(self.handle, batch_size, batch_cap, out.ctypes.data_as(_lib.c_i32p), out.size)
        if got < 0:
            _lib.check(int(got), "srb_sampler_epoch")
        return out[:got]


def stream_epoch(sampler, data, batch_size, batch_cap, ring_depth=16):
    """One epoch of batch buffers (srb_sampler_next_batch layout) in ONE reused int32 buffer (the consumer copies
    it, e.g. into a pinned slot).  The batches come from the native sample-ahead ring: a C++ thread samples up to
    `ring_depth` batches ahead (0.2 ms each on one core) while Python enqueues GPU work, so the sampler stops being
    the ceiling of a step that is faster than that (a Python producer thread was tried in round 1 and dropped: its
    queue hand-offs under the GIL cost more than they hid).  ring_depth=0: one native call per batch.  Python's
    `random` state is taken at the start and handed back when the epoch ends or the generator is closed (batches the
    ring sampled ahead but nobody read are un-drawn: the state is the reference's at that point of the stream);
    data.training_data gets the epoch's shuffle."""
    # an epoch generator that was abandoned without close() (e.g. zip(range(n), gen)) still owns the ring and a
    # pending state hand-back: retire it now -- its own `finally`, whenever the garbage collector gets to it, must
    # neither stop the new ring nor overwrite Python's `random` state with a stale one
    if getattr(sampler, "_open_epoch", None) is not None:
        sampler.ring_stop()
        sampler.push_state()
    token = object()
    sampler._open_epoch = token
    sampler.pull_state()
    try:
        perm = sampler.begin_epoch(want_perm=True)
        permute_training_data(data, perm)
        buf = np.empty(_lib.BATCH_HEADER + 5 * batch_cap, dtype=np.int32)
        if ring_depth > 0:
            sampler.ring_start(batch_size, batch_cap, ring_depth)
            while sampler._open_epoch is token and sampler.ring_pop(buf) > 0:
                yield buf
        else:
            while sampler._open_epoch is token and sampler.next_batch(batch_size, batch_cap, buf) > 0:
                yield buf
    finally:
        if sampler._open_epoch is token:
            sampler._open_epoch = None
            sampler.ring_stop()
            sampler.push_state()


def _sampler_for(data):
    s = getattr(data, "_srb_sampler", None)
    n = len(data.pair_users) if hasattr(data, "pair_users") else len(data.training_data)
    if s is None or s.n_pairs != n:
        s = NativePairSampler(data)
        data._srb_sampler = s
    return s


def permute_training_data(data, perm):
    """data.training_data[k] <- data.training_data[perm[k]], lazily when the data object supports it."""
    if hasattr(data, "shuffle_training_data"):
        data.shuffle_training_data(perm)
    else:
        td = data.training_data
        td[:] = [td[k] for k in perm]


def next_batch_pairwise(data, batch_size, n_negs=1):
    s = _sampler_for(data)
    s.pull_state()
    perm = s.begin_epoch(want_perm=True)
    s.push_state()
    permute_training_data(data, perm)  # the in-place shuffle side effect (sampler.py:7)
    u = np.empty(batch_size, dtype=np.int32)
    i = np.empty(batch_size, dtype=np.int32)
    j = np.empty(batch_size * n_negs, dtype=np.int32)
    while True:
        s.pull_state()
        b = s.next_batch_negs(batch_size, n_negs, u, i, j)
        s.push_state()
        if b == 0:
            return
        yield u[:b].tolist(), i[:b].tolist(), j[: b * n_negs].tolist()
