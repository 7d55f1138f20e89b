"""CPU restatement of the region-of-interest selection that consumes the single-mode scores.  TEST INFRASTRUCTURE.

Follows helpers.maxMean (helpers.py:253-274) -> filter_regions.Filter.read/filter/maxmean (filter_regions.py:105-183,
203-272, 375-448) with method='maxmean', input_type='bedgraph', aggregation_method='max', preserve_cols=True, and
roiSingle.createTopScoresTxt (roiSingle.py:95-142).  pandas is not used: the rolling mean reproduces pandas'
sliding-window arithmetic (Kahan-compensated add / remove with separate compensation terms, result snapped to the
repeated value when the whole window holds one value and clamped to 0 on a sign contradiction), because windows that
share their maximum are ranked by their mean.

Parity status: pinned by tests/golden/roi_*.npz, produced by running the reference's helpers.maxMean in the
authoring container (tests/golden/make_golden.py roi).
"""
import numpy as np


def rolling_center(values, window):
    """(rolling max, rolling mean) of a float64 vector, centered window, NaN where the window is incomplete:
    pandas Series.rolling(window, center=True).max() / .mean()."""
    v = np.asarray(values, dtype=np.float64)
    n = len(v)
    off = (window - 1) // 2
    rmax = np.full(n, np.nan)
    rmean = np.full(n, np.nan)
    sum_x = comp_add = comp_rem = 0.0
    nobs = neg = same = 0
    prev = v[0] if n else 0.0
    prev_s = prev_e = 0
    for i in range(n):
        e = min(i + 1 + off, n)
        s = max(i + 1 + off - window, 0)
        if i == 0 or s >= prev_e:
            sum_x = comp_add = comp_rem = 0.0
            nobs = neg = same = 0
            prev = v[s]
            lo = s
        else:
            for j in range(prev_s, s):
                val = v[j]
                nobs -= 1
                y = -val - comp_rem
                t = sum_x + y
                comp_rem = t - sum_x - y
                sum_x = t
                if np.signbit(val):
                    neg -= 1
            lo = prev_e
        for j in range(lo, e):
            val = v[j]
            nobs += 1
            y = val - comp_add
            t = sum_x + y
            comp_add = t - sum_x - y
            sum_x = t
            if np.signbit(val):
                neg += 1
            same = same + 1 if val == prev else 1
            prev = val
        prev_s, prev_e = s, e
        if nobs >= window:
            r = sum_x / nobs
            if same >= nobs:
                r = prev
            if neg == 0 and r < 0:
                r = 0.0
            elif neg == nobs and r > 0:
                r = 0.0
            rmean[i] = r
            rmax[i] = v[s:e].max()
    return rmax, rmean


def max_mean(starts, ends, score, window, max_regions=100):
    """Returns dict(original_idx, start, end, rolling_max, rolling_mean) of the selected regions in the order
    helpers.maxMean returns them (sorted by RollingMax, RollingMean, Score descending)."""
    starts = np.asarray(starts, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    score = np.asarray(score, dtype=np.float64)
    n = len(score)
    half = window // 2
    tail = half if window % 2 else half - 1
    # shift + dropna (filter_regions.py:380-389)
    orig = np.arange(half, n - tail, dtype=np.int64)
    if len(orig) == 0:
        return dict(original_idx=orig, start=orig, end=orig, rolling_max=np.zeros(0), rolling_mean=np.zeros(0))
    st = starts[orig - half]
    en = ends[orig + tail]
    sc = score[orig]
    rmax, rmean = rolling_center(sc, window)                                  # :391-392
    ok = ~np.isnan(rmax)                                                      # :400
    orig, st, en, sc, rmax, rmean = [a[ok] for a in (orig, st, en, sc, rmax, rmean)]
    ok = st < en                                                              # :402-405 regions over two chromosomes
    orig, st, en, sc, rmax, rmean = [a[ok] for a in (orig, st, en, sc, rmax, rmean)]
    p = len(orig)
    order = np.lexsort((np.arange(p), -sc, -rmean, -rmax))                    # :411-415, stable, all descending
    hits = np.zeros(p, dtype=bool)
    chosen = []
    for m in order:                                                           # :427-436 greedy non-overlap
        if len(chosen) >= max_regions:
            break
        lo = max(m - half, 0)
        hi = min(m + half + 1 if window % 2 else m + half, p)
        if not hits[lo:hi].any():
            hits[lo:hi] = True
            chosen.append(m)
    chosen = np.array(sorted(chosen), dtype=np.int64)                         # :437-439 back to original order
    # helpers.maxMean: Score := RollingMax, sort by (RollingMax, RollingMean, Score) descending (helpers.py:271)
    final = chosen[np.lexsort((np.arange(len(chosen)), -rmean[chosen], -rmax[chosen]))]
    return dict(original_idx=orig[final], start=st[final], end=en[final], rolling_max=rmax[final],
                rolling_mean=rmean[final])


def max_states(score_arr, original_idx, window):
    """roiSingle.py:122-129: per region the state with the largest single-bin score, ties -> higher state; 1-based."""
    half = window // 2
    out = []
    for idx in original_idx:
        lo, hi = idx - half, idx + half + (1 if window % 2 else 0)
        col_max = score_arr[lo:hi].max(axis=0)
        k = score_arr.shape[1]
        out.append(k - int(np.argmax(col_max[::-1])))
    return np.array(out, dtype=np.int32)


def roi_text(chrom_of_row, sel, states, state_names):
    """Lines of regionsOfInterest_<tag>.txt (roiSingle.py:137-140): chr, start, end, state name, |score| with 5
    decimals (score = RollingMax cast to float32), sign."""
    lines = []
    for i in range(len(sel["original_idx"])):
        s32 = float(np.float32(sel["rolling_max"][i]))
        lines.append("%s\t%d\t%d\t%s\t%.5f\t%s\n" % (chrom_of_row[sel["original_idx"][i]], sel["start"][i], sel["end"][i],
                                                     state_names[states[i] - 1], abs(s32), "+" if s32 >= 0 else "-"))
    return "".join(lines)
