"""CPU restatement (numpy) of the Epilogos scoring hot path.   *** TEST INFRASTRUCTURE ***

This module is the parity oracle for the CUDA path in epilogos_b200/.  It is imported only by tests/,
by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The product package
(epilogos_b200/) never imports it and has no CPU fallback.

Parity status: PINNED.  Every function below is checked in tests/test_oracle_golden.py against fixtures
under tests/golden/ that were produced by running the unmodified reference (/root/reference, v0.1.2)
in the authoring container with tests/golden/make_golden.py (the reference itself ships no tests and no
golden vectors for this path).

Two flavours are provided for every quantity:
  *_rowloop : follows the reference's per-row numpy algorithm (np.unique per row, Python double loop over
              present states, fancy-index scatter for S3).  Same asymptotic cost as the reference; this is
              what bench.py times as the "port" CPU baseline.
  (plain)   : whole-array numpy formulation of the same arithmetic, used by the parity tests at sizes where
              the row loop is too slow.  tests/test_oracle_golden.py asserts both flavours agree bit-for-bit.

All state matrices are 0-based integer arrays x[bins, biosamples] (helpers.py:154-155 subtracts 1).
Citations are file:line in /root/reference/epilogos/.
"""
import itertools

import numpy as np

# ------------------------------------------------------------------------------------------------
# per-bin state counts (the quantity every S1/S2 formula is a function of)
# ------------------------------------------------------------------------------------------------

def bin_counts(x, num_states):
    """cnt[b, s] = #{j : x[b, j] == s}.  Restates np.unique(dataArr[row], return_counts=True) of
    scores.py:341 / expected.py:152 for all rows at once."""
    x = np.asarray(x)
    bins, _ = x.shape
    flat = (np.arange(bins, dtype=np.int64)[:, None] * num_states + x.astype(np.int64)).ravel()
    return np.bincount(flat, minlength=bins * num_states).reshape(bins, num_states).astype(np.int64)


# ------------------------------------------------------------------------------------------------
# expected tables (integer counts) : expected.py
# ------------------------------------------------------------------------------------------------

def s1_expected_counts(x, num_states):
    """expected.py:106-113 -- histogram of every label in the chunk.  int64[K]."""
    return np.bincount(np.asarray(x).ravel().astype(np.int64), minlength=num_states).astype(np.int64)


def s1_expected_counts_rowloop(x, num_states):
    out = np.zeros(num_states, dtype=np.int64)
    states, counts = np.unique(np.asarray(x), return_counts=True)          # expected.py:111
    out[states] += counts
    return out


def s2_expected_counts(x, num_states):
    """expected.py:146-158 -- sum over bins of c_s*c_t (s != t) and c_s*(c_s-1) (s == t).  int64[K,K]."""
    cnt = bin_counts(x, num_states)
    n2 = cnt.T @ cnt
    n2[np.diag_indices(num_states)] -= cnt.sum(axis=0)
    return n2


def s2_expected_counts_rowloop(x, num_states):
    x = np.asarray(x)
    out = np.zeros((num_states, num_states), dtype=np.int64)
    for row in x:
        states, counts = np.unique(row, return_counts=True)                  # expected.py:152
        for a, ca in zip(states, counts):
            for b, cb in zip(states, counts):
                out[a, b] += ca * (cb - 1) if a == b else ca * cb         # expected.py:155-158
    return out


def s3_expected_counts(x, num_states):
    """expected.py:183-200 -- N3[i, j, x[b,i], x[b,j]] += 1 for all ordered i != j.  The reference builds an
    int32 table per worker and np.sum()s the list, which yields int64; diagonal blocks i == j
    stay zero because itertools.permutations never pairs a column with itself (expected.py:183)."""
    x = np.asarray(x)
    bins, cols = x.shape
    onehot = np.zeros((bins, cols * num_states), dtype=np.float64)
    onehot[np.arange(bins)[:, None], np.arange(cols)[None, :] * num_states + x] = 1.0
    gram = np.rint(onehot.T @ onehot).astype(np.int64)                      # exact: entries <= bins < 2^53
    n3 = gram.reshape(cols, num_states, cols, num_states).transpose(0, 2, 1, 3).copy()
    n3[np.arange(cols), np.arange(cols)] = 0
    return n3


def s3_expected_counts_rowloop(x, num_states):
    x = np.asarray(x)
    bins, cols = x.shape
    pairs = np.array(list(itertools.permutations(range(cols), 2)), dtype=np.int64).T   # expected.py:183
    out = np.zeros((cols, cols, num_states, num_states), dtype=np.int32)
    for row in x:
        out[pairs[0], pairs[1], row[pairs[0]], row[pairs[1]]] += 1           # expected.py:199-200
    return out.astype(np.int64)


def normalize_expected(counts):
    """expectedCombination.py:42 -- float64 true divide by the grand total, then round to float32."""
    counts = np.asarray(counts)
    return (counts / np.sum(counts)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# KL terms : scores.py:539-550
# ------------------------------------------------------------------------------------------------

_TINY = np.finfo(float).tiny


def kl_terms(obs, exp):
    """obs * log2(obs / exp) with the reference's masking (scores.py:550):
    numpy.ma.divide masks where |obs| * tiny >= |exp| (in practice exp == 0) and fills 0;
    numpy.ma.log2 masks arguments <= 0 and fills 0.  Result dtype = result_type(obs, exp):
    float64 for S1/S2 (obs float64), float32 for S3 (both float32)."""
    obs = np.asarray(obs)
    exp = np.asarray(exp)
    dt = np.result_type(obs, exp)
    shape = np.broadcast(obs, exp).shape
    ratio = np.zeros(shape, dtype=dt)
    with np.errstate(all="ignore"):
        ok = ~(np.abs(obs) * _TINY >= np.abs(exp))
        np.divide(obs, exp, out=ratio, where=ok, dtype=dt)
        lg = np.zeros(shape, dtype=dt)
        np.log2(ratio, out=lg, where=ratio > 0)
        return (obs * lg).astype(dt, copy=False)


# ------------------------------------------------------------------------------------------------
# scores : scores.py
# ------------------------------------------------------------------------------------------------

def s1_scores_from_counts(cnt, width, exp1, dtype=np.float32):
    """scores.py:339-344 + :317 -- o = cnt / C (float64), kl per state, stored into a float32 row.
    `width` is dataArr.shape[1] (scores.py:343).  dtype=np.float64 returns the unrounded values."""
    obs = np.asarray(cnt, dtype=np.int64) / width
    return kl_terms(obs, np.asarray(exp1)[None, :]).astype(dtype)


def s1_scores(x, num_states, exp1, dtype=np.float32):
    x = np.asarray(x)
    return s1_scores_from_counts(bin_counts(x, num_states), x.shape[1], exp1, dtype)


def s1_scores_rowloop(x, num_states, exp1):
    x = np.asarray(x)
    out = np.zeros((x.shape[0], num_states), dtype=np.float32)
    for r, row in enumerate(x):
        obs = np.zeros(num_states)
        states, counts = np.unique(row, return_counts=True)                  # scores.py:341
        obs[states] = counts / x.shape[1]                                    # scores.py:343
        out[r] = kl_terms(obs, exp1)                                         # scores.py:317
    return out


def s2_obs_from_counts(cnt, perms):
    """scores.py:443-451 -- exact integer c_s*c_t (c_s*(c_s-1) on the diagonal), ONE float64 divide by
    `perms` = C*(C-1)."""
    cnt = np.asarray(cnt, dtype=np.int64)
    prod = cnt[:, :, None] * cnt[:, None, :]
    k = cnt.shape[1]
    prod[:, np.arange(k), np.arange(k)] -= cnt
    return prod / perms


def s2_scores_from_counts(cnt, perms, exp2, dtype=np.float32, chunk=4096):
    """scores.py:412 -- klScoreND(obs, E2).sum(axis=0): for every target state t the K terms are added
    in order s = 0..K-1 (numpy reduces the outer axis of a C-contiguous [K,K] array row by row), float64,
    then the row is stored into float32."""
    cnt = np.asarray(cnt, dtype=np.int64)
    exp2 = np.asarray(exp2)
    bins, k = cnt.shape
    out = np.zeros((bins, k), dtype=dtype)
    for lo in range(0, bins, chunk):
        kl = kl_terms(s2_obs_from_counts(cnt[lo:lo + chunk], perms), exp2[None, :, :])
        acc = np.zeros((kl.shape[0], k), dtype=np.float64)
        for s in range(k):
            acc += kl[:, s, :]
        out[lo:lo + chunk] = acc.astype(dtype)
    return out


def s2_scores(x, num_states, exp2, dtype=np.float32):
    x = np.asarray(x)
    c = x.shape[1]
    return s2_scores_from_counts(bin_counts(x, num_states), c * (c - 1), exp2, dtype)


def s2_scores_rowloop(x, num_states, exp2, perms=None):
    x = np.asarray(x)
    c = x.shape[1]
    if perms is None:
        perms = c * (c - 1)                                                  # scores.py:371
    out = np.zeros((x.shape[0], num_states), dtype=np.float32)
    for r, row in enumerate(x):
        obs = np.zeros((num_states, num_states))
        states, counts = np.unique(row, return_counts=True)                  # scores.py:444
        for a, ca in zip(states, counts):
            for b, cb in zip(states, counts):
                obs[a, b] = ca * (cb - 1) / perms if a == b else ca * cb / perms   # scores.py:447-451
        out[r] = kl_terms(obs, exp2).sum(axis=0)                             # scores.py:412
    return out


def s3_pair_terms(num_cols, exp3, dtype=np.float32):
    """scores.py:479-480 -- T = klScoreND(ones/(C(C-1)), E3).  dtype float32 reproduces the reference
    (all-float32 arithmetic); dtype float64 is the exact-arithmetic variant the CUDA kernel is held to."""
    q = (np.ones((), dtype=dtype) / (num_cols * (num_cols - 1))).astype(dtype)
    return kl_terms(np.full(np.asarray(exp3).shape, q, dtype=dtype), np.asarray(exp3).astype(dtype))


def s3_scores_rowloop(x, num_states, exp3):
    """scores.py:474-504, reference-faithful: float32 terms, np.add.at in itertools.permutations order
    (i-major), bucket = state of the SECOND biosample of the pair."""
    x = np.asarray(x)
    bins, cols = x.shape
    pairs = np.array(list(itertools.permutations(range(cols), 2)), dtype=np.int64).T
    terms = s3_pair_terms(cols, exp3, np.float32)
    out = np.zeros((bins, num_states), dtype=np.float32)
    acc = np.zeros(num_states, dtype=np.float32)
    for r, row in enumerate(x):
        np.add.at(acc, row[pairs[1]], terms[pairs[0], pairs[1], row[pairs[0]], row[pairs[1]]])
        out[r] = acc
        acc.fill(0)
    return out


def s3_scores_f64(x, num_states, exp3, terms=None):
    """Same sum as s3_scores_rowloop evaluated in float64 (terms and accumulation).  This is the
    restatement the CUDA S3 score kernel is compared with at 1e-9 (SURVEY.md 8c tolerances); the
    reference's own float32 result differs from it by its accumulation noise (<= ~2e-2 abs at C=833)."""
    x = np.asarray(x)
    bins, cols = x.shape
    if terms is None:
        terms = s3_pair_terms(cols, exp3, np.float64)
    out = np.zeros((bins, num_states), dtype=np.float64)
    ii, jj = np.nonzero(~np.eye(cols, dtype=bool))
    for r, row in enumerate(x):
        v = terms[ii, jj, row[ii], row[jj]]
        out[r] = np.bincount(row[jj], weights=v, minlength=num_states)
    return out


# ------------------------------------------------------------------------------------------------
# paired mode : helpers.py:162-194, scores.py:172-256, 282-303, 319-322, 373-398, 414-421
# ------------------------------------------------------------------------------------------------

def reference_shuffle_indices(seed, rows, width):
    """helpers.py:183 -- argsort(np.random.rand(rows, width), axis=1) with the legacy global RNG seeded
    by np.random.seed(seed) in the parent; a forked worker for chunk r draws RandomState(seed).rand(rows_r,
    width) (every chunk starts from the same inherited state)."""
    return np.argsort(np.random.RandomState(seed).rand(rows, width), axis=1)


def paired_split(xa, xb, perm, group_size=-1):
    """helpers.py:173-194 -- shuffled = take_along_axis(concat(A,B), perm); halves A', B'."""
    comb = np.concatenate((np.asarray(xa), np.asarray(xb)), axis=1)
    sh = np.take_along_axis(comb, np.asarray(perm), axis=1)
    if group_size == -1:
        return sh[:, :xa.shape[1]], sh[:, xa.shape[1]:]
    return sh[:, :group_size], sh[:, group_size:2 * group_size]


def quiescent_mask(xa, xb, quiescent_state):
    """scores.py:294-303 -- True where every label of both groups equals quiescent_state (0-based here;
    the CLI converts with run.py:113-115); all False when quiescent_state == -1."""
    xa = np.asarray(xa); xb = np.asarray(xb)
    if quiescent_state == -1:
        return np.zeros(xa.shape[0], dtype=bool)
    return np.all(xa == quiescent_state, axis=1) & np.all(xb == quiescent_state, axis=1)


def paired_scores(xa, xb, perm, num_states, saliency, exp, quiescent_state, group_size=-1):
    """Returns dict(delta=f32[B,K], null_distances=f32[B], quiescence=bool[B], scoreA, scoreB, nullA, nullB).
    S1: observation width is the actual slice width (scores.py:343).  S2 quirk: the shuffled halves are
    normalised with C1(C1-1) / C2(C2-1) even if -g changed their width (scores.py:397-398, 418-421)."""
    xa = np.asarray(xa); xb = np.asarray(xb)
    sa, sb = paired_split(xa, xb, perm, group_size)
    if saliency == 1:
        f = lambda m, ref: s1_scores(m, num_states, exp)
    elif saliency == 2:
        f = lambda m, ref: s2_scores_from_counts(bin_counts(m, num_states), ref.shape[1] * (ref.shape[1] - 1), exp)
    else:
        raise ValueError("Please ensure that saliency metric is either 1 or 2 for Pairwise Epilogos")
    score_a, score_b = f(xa, xa), f(xb, xb)
    null_a, null_b = f(sa, xa), f(sb, xb)
    delta = score_a - score_b                                                 # scores.py:223
    null_diff = null_a - null_b                                               # scores.py:224-225
    sign = np.sign(np.sum(null_diff, axis=1))                                 # scores.py:231
    dist = np.sum(np.square(null_diff), axis=1) * sign                        # scores.py:232
    return dict(delta=delta, null_distances=dist, quiescence=quiescent_mask(xa, xb, quiescent_state),
                scoreA=score_a, scoreB=score_b, nullA=null_a, nullB=null_b)


# ------------------------------------------------------------------------------------------------
# text output : scores.py:509-536
# ------------------------------------------------------------------------------------------------

def format_scores_text(scores32, chrom, starts, ends):
    """One line per bin: chr \t start \t end \t K values "{:.5f}" (float32 -> Python float -> 5 dp)."""
    rows = []
    for i in range(scores32.shape[0]):
        rows.append("%s\t%d\t%d\t%s\n" % (chrom, starts[i], ends[i],
                                         "\t".join("%.5f" % float(v) for v in scores32[i])))
    return "".join(rows).encode()


# ------------------------------------------------------------------------------------------------
# whole-path drivers (what bench.py times as the CPU "port" baseline)
# ------------------------------------------------------------------------------------------------

def expected_and_scores_rowloop(x, num_states, saliency):
    """expected.main -> expectedCombination.main -> scores.main for one in-memory chunk, without the
    TSV/gzip I/O, following the reference's per-row algorithm."""
    if saliency == 1:
        exp = normalize_expected(s1_expected_counts_rowloop(x, num_states))
        return exp, s1_scores_rowloop(x, num_states, exp)
    if saliency == 2:
        exp = normalize_expected(s2_expected_counts_rowloop(x, num_states))
        return exp, s2_scores_rowloop(x, num_states, exp)
    if saliency == 3:
        exp = normalize_expected(s3_expected_counts_rowloop(x, num_states))
        return exp, s3_scores_rowloop(x, num_states, exp)
    raise ValueError("Please ensure that saliency metric is either 1, 2, or 3")


def expected_and_scores(x, num_states, saliency):
    if saliency == 1:
        exp = normalize_expected(s1_expected_counts(x, num_states))
        return exp, s1_scores(x, num_states, exp)
    if saliency == 2:
        exp = normalize_expected(s2_expected_counts(x, num_states))
        return exp, s2_scores(x, num_states, exp)
    if saliency == 3:
        exp = normalize_expected(s3_expected_counts(x, num_states))
        return exp, s3_scores_f64(x, num_states, exp)
    raise ValueError("Please ensure that saliency metric is either 1, 2, or 3")


# ------------------------------------------------------------------------------------------------
# synthetic state matrices -- numpy twin of the device generator used by bench.py
# ------------------------------------------------------------------------------------------------

def state_prior(num_states):
    """Skewed prior: last (quiescent) state ~62 %, then a geometric tail over the rest (real chr1 data:
    state 18 ~62 %, state 6 ~16 %, state 17 ~9 %)."""
    w = np.array([0.5 ** (i * 0.6) for i in range(num_states - 1)], dtype=np.float64)
    w = 0.38 * w / w.sum()
    return np.concatenate([w, [0.62]])


def synth_states(bins, cols, num_states, seed, kind="realistic"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, num_states, size=(bins, cols), dtype=np.int8)
    prior = state_prior(num_states)
    dom = rng.choice(num_states, size=bins, p=prior)
    other = rng.choice(num_states, size=(bins, cols), p=prior)
    keep = rng.random((bins, cols)) < 0.6
    return np.where(keep, dom[:, None], other).astype(np.int8)


# ------------------------------------------------------------------------------------------------
# paired ROI stage reductions : roiAndVisualPairwise.py:339-354
# ------------------------------------------------------------------------------------------------

def text_round_trip(delta32):
    """What readInData sees after re-reading pairwiseDelta_*.txt.gz: "%.5f" text -> float64 -> float32."""
    flat = np.asarray(delta32, dtype=np.float32).ravel()
    back = np.array([float("%.5f" % float(v)) for v in flat], dtype=np.float64)
    return back.astype(np.float32).reshape(np.shape(delta32))


def paired_real_reductions(delta32, round_trip=True):
    """(distanceArrReal, maxDiffArr) of roiAndVisualPairwise.py:347-354."""
    d = text_round_trip(delta32) if round_trip else np.asarray(delta32, dtype=np.float32)
    dist = np.sum(np.square(d), axis=1) * np.sign(np.sum(d, axis=1))
    max_diff = np.abs(np.argmax(np.abs(np.flip(d, axis=1)), axis=1) - d.shape[1]).astype(np.int32)
    return dist, max_diff
