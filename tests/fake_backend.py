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

    def add_counts(self, cnt_a, cnt_b):
        return cnt_a + cnt_b

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

    def scores(self, cnt, width, saliency, exp, perms=None, exact=False):
        c = self._cnt(cnt)
        if saliency == 1:
            return torch.from_numpy(orc.s1_scores_from_counts(c, width, exp.numpy()))
        return torch.from_numpy(orc.s2_scores_from_counts(c, perms or width * (width - 1), exp.numpy()))

    # paired mode
    def shuffled_counts_perm(self, states_a, states_b, perm, num_states, size_a, size_b):
        gs = -1 if (size_a, size_b) == (states_a.shape[1], states_b.shape[1]) else size_a
        sa, sb = orc.paired_split(states_a, states_b, perm, gs)
        return self.counts(sa, num_states), self.counts(sb, num_states)

    def shuffled_counts_device(self, cnt_a, cnt_b, size_a, size_b, seed, nperm=1, width=None, bin_offset=0):
        """numpy multivariate hypergeometric draws keyed like the device kernel: (seed, permutation, GLOBAL bin)."""
        a, b = self._cnt(cnt_a), self._cnt(cnt_b)
        rows, k = a.shape
        oa = np.zeros((nperm, rows, k), dtype=np.int64)
        ob = np.zeros_like(oa)
        for p in range(nperm):
            for r in range(rows):
                rng = np.random.default_rng([int(seed) & (2 ** 63 - 1), p, bin_offset + r])
                comb = a[r] + b[r]
                na = min(size_a, int(comb.sum()))
                ga = rng.multivariate_hypergeometric(comb, na) if na else np.zeros(k, dtype=np.int64)
                rest = comb - ga
                nb = min(size_b, int(rest.sum()))
                gb = rng.multivariate_hypergeometric(rest, nb) if nb else np.zeros(k, dtype=np.int64)
                oa[p, r], ob[p, r] = ga, gb
        ta = torch.from_numpy(oa.astype(np.uint16).view(np.int16))
        tb = torch.from_numpy(ob.astype(np.uint16).view(np.int16))
        return (ta[0], tb[0]) if nperm == 1 else (ta, tb)

    def pairwise_combine(self, score_a, score_b, null_a, null_b):
        delta = None if score_a is None else score_a - score_b
        dist = None
        if null_a is not None:
            d = (null_a - null_b).numpy()
            dist = torch.from_numpy(np.sum(np.square(d), axis=1) * np.sign(np.sum(d, axis=1)))
        return delta, dist

    def quiescent_mask(self, cnt_a, cols_a, cnt_b, cols_b, q):
        a, b = self._cnt(cnt_a), self._cnt(cnt_b)
        if q == -1:
            return torch.zeros(a.shape[0], dtype=torch.uint8)
        return torch.from_numpy(((a[:, q] == cols_a) & (b[:, q] == cols_b)).astype(np.uint8))
