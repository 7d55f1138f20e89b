// Region-of-interest selection over the per-bin score sums.  Host code.
//
// Reference: helpers.maxMean (helpers.py:253-274) -> filter_regions.Filter.maxmean (filter_regions.py:375-448),
// method 'maxmean', bedgraph input, aggregation 'max': shift the coordinates to window extents, centered rolling
// max / mean of the score, drop incomplete windows and windows that straddle two chromosomes, rank by
// (RollingMax, RollingMean, Score) descending, greedily keep non-overlapping windows until max_regions.
// The rolling mean reproduces pandas' sliding arithmetic (Kahan-compensated add / remove, separate compensation
// terms, snap to the repeated value, sign clamp) because windows sharing their maximum are ranked by their mean.
// Ranking uses a strict total order (max, mean, score descending, then position ascending) = pandas' stable
// multi-key sort; only the leading candidates are ordered (nth_element + sort of a growing prefix).
#include <math.h>

#include <algorithm>
#include <deque>
#include <vector>

#include "common.cuh"

namespace epi {

struct Cand {
    double rmax, rmean, score;
    int64_t orig;        // index into the input arrays
    int64_t start, end;  // window extents
};

}  // namespace epi

using namespace epi;

extern "C" int epi_roi_maxmean(const double* score, const int64_t* starts, const int64_t* ends, int64_t n,
                               int32_t window, int32_t max_regions, int64_t* out_original_idx, int64_t* out_start,
                               int64_t* out_end, double* out_rolling_max, double* out_rolling_mean, int32_t* n_out) {
    EPI_REQUIRE(score && starts && ends && n_out, "null pointer argument");
    EPI_REQUIRE(n >= 0 && window >= 1 && max_regions >= 0, "bad arguments");
    *n_out = 0;
    const int64_t half = window / 2;
    const int64_t tail = (window % 2) ? half : half - 1;
    const int64_t m = n - half - tail;                 // rows that survive the coordinate shift (filter_regions.py:380-389)
    if (m <= 0 || max_regions == 0) return 0;
    // ---- centered rolling max / mean over the surviving rows (trimmed index t <-> original index t + half) ----
    const int64_t off = (window - 1) / 2;
    std::vector<Cand> cand;
    cand.reserve((size_t)std::max<int64_t>(m - window + 1, 0));
    {
        const double* v = score + half;
        double sum_x = 0.0, comp_add = 0.0, comp_rem = 0.0, prev = v[0];
        int64_t nobs = 0, neg = 0, same = 0, prev_s = 0, prev_e = 0;
        std::deque<int64_t> mono;                      // indices of a decreasing deque for the sliding maximum
        for (int64_t i = 0; i < m; ++i) {
            const int64_t e = std::min<int64_t>(i + 1 + off, m);
            const int64_t s = std::max<int64_t>(i + 1 + off - window, 0);
            int64_t lo;
            if (i == 0 || s >= prev_e) {
                sum_x = comp_add = comp_rem = 0.0;
                nobs = neg = same = 0;
                prev = v[s];
                lo = s;
                mono.clear();
            } else {
                for (int64_t j = prev_s; j < s; ++j) {
                    const double val = v[j];
                    --nobs;
                    const double y = -val - comp_rem;
                    const double t = sum_x + y;
                    comp_rem = t - sum_x - y;
                    sum_x = t;
                    if (signbit(val)) --neg;
                }
                lo = prev_e;
            }
            for (int64_t j = lo; j < e; ++j) {
                const double val = v[j];
                ++nobs;
                const double y = val - comp_add;
                const double t = sum_x + y;
                comp_add = t - sum_x - y;
                sum_x = t;
                if (signbit(val)) ++neg;
                same = (val == prev) ? same + 1 : 1;
                prev = val;
                while (!mono.empty() && v[mono.back()] <= val) mono.pop_back();
                mono.push_back(j);
            }
            while (!mono.empty() && mono.front() < s) mono.pop_front();
            prev_s = s;
            prev_e = e;
            if (nobs >= window) {
                double r = sum_x / (double)nobs;
                if (same >= nobs) r = prev;
                if (neg == 0 && r < 0) r = 0.0;
                else if (neg == nobs && r > 0) r = 0.0;
                const int64_t orig = i + half;
                const int64_t st = starts[orig - half], en = ends[orig + tail];
                if (st < en) {                           // windows over two chromosomes are dropped (:402-405)
                    Cand c;
                    c.rmax = v[mono.front()];
                    c.rmean = r;
                    c.score = v[i];
                    c.orig = orig;
                    c.start = st;
                    c.end = en;
                    cand.push_back(c);
                }
            }
        }
    }
    const int64_t p = (int64_t)cand.size();
    if (p == 0) return 0;
    // ---- rank + greedy non-overlap in candidate-index ("MethodIdx") space ----
    std::vector<int64_t> order((size_t)p);
    for (int64_t i = 0; i < p; ++i) order[(size_t)i] = i;
    auto better = [&](int64_t a, int64_t b) {
        const Cand &x = cand[(size_t)a], &y = cand[(size_t)b];
        if (x.rmax != y.rmax) return x.rmax > y.rmax;
        if (x.rmean != y.rmean) return x.rmean > y.rmean;
        if (x.score != y.score) return x.score > y.score;
        return a < b;
    };
    std::vector<char> hits((size_t)p, 0);
    std::vector<int64_t> chosen;
    int64_t sorted_upto = 0;
    int64_t want = std::min<int64_t>(p, std::max<int64_t>(4096, (int64_t)max_regions * window * 8));
    while ((int64_t)chosen.size() < max_regions && sorted_upto < p) {
        if (want < p) std::nth_element(order.begin() + sorted_upto, order.begin() + want, order.end(), better);
        std::sort(order.begin() + sorted_upto, order.begin() + want, better);
        for (int64_t k = sorted_upto; k < want && (int64_t)chosen.size() < max_regions; ++k) {
            const int64_t mi = order[(size_t)k];
            const int64_t lo = std::max<int64_t>(mi - half, 0);
            const int64_t hi = std::min<int64_t>((window % 2) ? mi + half + 1 : mi + half, p);
            bool free_ = true;
            for (int64_t j = lo; j < hi; ++j)
                if (hits[(size_t)j]) {
                    free_ = false;
                    break;
                }
            if (free_) {
                for (int64_t j = lo; j < hi; ++j) hits[(size_t)j] = 1;
                chosen.push_back(mi);
            }
        }
        sorted_upto = want;
        want = std::min<int64_t>(p, want * 4);
    }
    // ---- back to input order, then helpers.maxMean's final stable ranking by (RollingMax, RollingMean, Score) descending;
    //      Score was overwritten with RollingMax by the 'max' aggregation (filter_regions.py:215-216), so two keys suffice ----
    std::sort(chosen.begin(), chosen.end());
    std::stable_sort(chosen.begin(), chosen.end(), [&](int64_t a, int64_t b) {
        const Cand &x = cand[(size_t)a], &y = cand[(size_t)b];
        if (x.rmax != y.rmax) return x.rmax > y.rmax;
        return x.rmean > y.rmean;
    });
    for (size_t i = 0; i < chosen.size(); ++i) {
        const Cand& c = cand[(size_t)chosen[i]];
        if (out_original_idx) out_original_idx[i] = c.orig;
        if (out_start) out_start[i] = c.start;
        if (out_end) out_end[i] = c.end;
        if (out_rolling_max) out_rolling_max[i] = c.rmax;
        if (out_rolling_mean) out_rolling_mean[i] = c.rmean;
    }
    *n_out = (int32_t)chosen.size();
    return 0;
}
