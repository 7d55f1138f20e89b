"""Tensor-core K2 / K5 (csrc/tc_tables.cu) against the integer-pipe / fp64-pipe kernels they replace and the oracle,
plus timings at the benchmark shape.  EPI_K2_ALU / EPI_K5_ALU select the older kernels (read on every call)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402
from oracle import epilogos_oracle as orc  # noqa: E402  (checker only)


def use(alu2, alu5):
    for name, on in (("EPI_K2_ALU", alu2), ("EPI_K5_ALU", alu5)):
        if on:
            os.environ[name] = "1"
        else:
            os.environ.pop(name, None)


def check(bins, cols, k, seed, kind="realistic"):
    x = synth.synth_states_device(bins, cols, k, seed=seed, kind=kind)
    cnt = engine.bin_counts(x, cols, k)
    use(True, True)
    n1a, n2a = engine.expected_tables(cnt, cols)
    use(False, False)
    n1t, n2t = engine.expected_tables(cnt, cols)
    torch.cuda.synchronize()
    ok2 = bool(torch.equal(n1a, n1t) and torch.equal(n2a, n2t))
    if not ok2:
        d = (n2t - n2a).cpu().numpy()
        print("K2 MISMATCH", bins, cols, k, "n1 equal", bool(torch.equal(n1a, n1t)), "\n", d[:6, :6], "\nref\n",
              n2a.cpu().numpy()[:4, :4])
    e2 = engine.normalize(n2a)
    use(True, True)
    _, s_alu = engine.scores_s2(cnt, cols, e2, want64=True)
    use(False, False)
    s32, s_tc = engine.scores_s2(cnt, cols, e2, want64=True)
    torch.cuda.synchronize()
    a, b = s_alu.cpu().numpy(), s_tc.cpu().numpy()
    err = np.abs(a - b)
    ok5 = bool(np.all(err <= 1e-9 * np.abs(a) + 1e-12))
    ok32 = bool(np.array_equal(s32.cpu().numpy(), b.astype(np.float32)))
    if not ok5:
        bad = np.argwhere(err > 1e-9 * np.abs(a) + 1e-12)
        rows = np.unique(bad[:, 0])
        runs = np.split(rows, np.where(np.diff(rows) != 1)[0] + 1)
        print("K5 MISMATCH", bins, cols, k, "values", len(bad), "rows", len(rows), "runs",
              [(int(r[0]), int(r[-1]), int(r[0]) // 128, int(r[0]) % 128) for r in runs][:12])
        sub = rows[:64]
        ref = orc.s2_scores_from_counts(engine.counts_to_numpy(cnt[torch.as_tensor(sub, device="cuda")].contiguous()),
                                        cols * (cols - 1), e2.cpu().numpy(), dtype=np.float64)
        print("  alu matches oracle on bad rows:", bool(np.allclose(a[sub], ref, rtol=1e-9, atol=1e-12)),
              " tc matches oracle:", bool(np.allclose(b[sub], ref, rtol=1e-9, atol=1e-12)))
    if bins <= 20000:
        ref = orc.s2_scores_from_counts(engine.counts_to_numpy(cnt), cols * (cols - 1), e2.cpu().numpy(), dtype=np.float64)
        ok5 = ok5 and bool(np.all(np.abs(b - ref) <= 1e-9 * np.abs(ref) + 1e-12))
    print(json.dumps({"bins": bins, "cols": cols, "k": k, "kind": kind, "k2_ok": ok2, "k5_ok": ok5, "k5_f32_consistent": ok32,
                      "k5_max_abs_vs_alu": float(err.max()) if err.size else 0.0}), flush=True)
    return ok2 and ok5 and ok32


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    engine.device_info()
    ok = True
    if "--stress" in sys.argv:
        for rep in range(int(os.environ.get("STRESS_REPS", "3"))):
            ok = check(9_000_000, 127, 15, 12 + rep) and ok
        print("TC STRESS", "ALL OK" if ok else "FAILED", flush=True)
        return 0 if ok else 1
    for args in [(1000, 833, 18, 1), (128, 10, 18, 2), (129, 127, 15, 3), (5, 3, 2, 4), (40000, 300, 17, 5), (70000, 200, 32, 6),
                 (33000, 255, 16, 7), (33000, 256, 25, 8), (300000, 2047, 18, 9), (100000, 833, 18, 10, "uniform"),
                 (4_200_000, 833, 18, 11), (9_000_000, 127, 15, 12)]:
        ok = check(*args) and ok
    print("TC CHECK", "ALL OK" if ok else "FAILED", flush=True)
    if "--time" in sys.argv:
        for bins, cols, k in ((15_500_000, 833, 18), (15_500_000, 127, 15)):
            x = synth.synth_states_device(bins, cols, k, seed=1)
            cnt = engine.bin_counts(x, cols, k)
            res = {"bins": bins, "cols": cols, "k": k}
            res["k1_ms"] = timeit(lambda: engine.bin_counts(x, cols, k, out=cnt))
            del x
            out = torch.empty((bins, k), dtype=torch.float32, device="cuda")
            use(True, True)
            n1, n2 = engine.expected_tables(cnt, cols)
            e2 = engine.normalize(n2)
            res["k2_alu_ms"] = timeit(lambda: engine.expected_tables(cnt, cols, want_s1=False))
            res["k5_alu_ms"] = timeit(lambda: engine.scores_s2(cnt, cols, e2, out32=out))
            use(False, False)
            res["k2_tc_ms"] = timeit(lambda: engine.expected_tables(cnt, cols, want_s1=False))
            res["k5_tc_ms"] = timeit(lambda: engine.scores_s2(cnt, cols, e2, out32=out))
            res["k5_tc_GBps"] = bins * k * 6 / (res["k5_tc_ms"] * 1e-3) / 1e9
            res["k2_tc_GBps"] = bins * k * 2 / (res["k2_tc_ms"] * 1e-3) / 1e9
            print(json.dumps(res), flush=True)
            del cnt, out
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
