"""Base of Interaction (data/data.py): the run's configuration plus the raw training / test triples.

`training_data` is the list the reference's sampler shuffles in place every epoch (util/sampler.py:7).  The native
sampler shuffles ids, not Python lists, so the list order is kept as a pending permutation and applied (to the
same list object) when somebody reads `training_data` -- 0.3 s of list copying per epoch at yelp2018 otherwise."""
import numpy as np


class Data(object):
    def __init__(self, conf, training, test):
        self.config, self.test_data = conf, test
        self._td, self._td_perm = training, None

    @property
    def training_data(self):
        if self._td_perm is not None:
            order, self._td_perm = self._td_perm.tolist(), None
            src = self._td
            src[:] = [src[k] for k in order]
        return self._td

    @training_data.setter
    def training_data(self, rows):
        self._td, self._td_perm = rows, None

    def shuffle_training_data(self, perm):
        """Record an in-place shuffle: new[k] = old[perm[k]] (composes with shuffles not yet applied)."""
        perm = np.asarray(perm, dtype=np.int64)
        self._td_perm = perm if self._td_perm is None else self._td_perm[perm]
