"""Stand-in compute provider for CPU-only tests of the HOST logic (sharding, all-reduce, gather, file naming).
It answers the backend interface of epilogos_b200/backend.py with the oracle.  Test infrastructure only: the
product never imports this, and epilogos_b200 itself has no CPU path."""
import numpy as np
import torch

from oracle import epilogos_oracle as orc


class OracleBackend:
    name = "oracle-test-double"

    def counts(self, states0, num_states):
        return torch.from_numpy(orc.bin_counts(states0, num_states).astype(np.uint16).view(np.int16))

    @staticmethod
    def _cnt(cnt):
        return cnt.numpy().view(np.uint16).astype(np.int64)

    def expected_table(self, cnt, width, saliency):
        c = self._cnt(cnt)
        if saliency == 1:
            return torch.from_numpy(c.sum(axis=0))
        n2 = c.T @ c
        n2[np.diag_indices(c.shape[1])] -= c.sum(axis=0)
        return torch.from_numpy(n2)

    def normalize(self, counts):
        return torch.from_numpy(orc.normalize_expected(counts.numpy()))

    def to_device(self, array):
        return torch.as_tensor(array)

    def scores(self, cnt, width, saliency, exp, perms=None):
        c = self._cnt(cnt)
        if saliency == 1:
            return torch.from_numpy(orc.s1_scores_from_counts(c, width, exp.numpy()))
        return torch.from_numpy(orc.s2_scores_from_counts(c, perms or width * (width - 1), exp.numpy()))
