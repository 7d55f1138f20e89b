#!/usr/bin/env python
"""bench.py -- bins/sec for expected + scores on B200 (BASELINE.json metric), plus the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--saliency 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic matrix that is already resident in HBM:
K1 counts -> K2 expected table (tensor-core Gram of the count bytes) -> (allreduce over ranks) -> K4 normalise ->
K5 scores (S2: table preparation + tensor-core mat-vec kernel + the gated DIRECT fallback = 6 launches per step).
Workload at every N: BASELINE.json configs[1], S2 on 15.5 M bins x 833 biosamples x 18 states PER GPU
(weak scaling; the bins shard with no data-path collective, only the 18x18 table is all-reduced).
The matrix (13 GB) is far larger than L2 (126 MB), so no L2 flush is needed between steps.

`e2e` is the same metric through epi_single_host (the reference-facing C-ABI call) with the matrix in
pinned HOST memory: H2D of the matrix and D2H of tables + scores are inside the timed region.

`--impl reference` times the CPU restatement of the reference's row-loop algorithm (oracle/, kind "port":
the reference is pure Python and cannot travel to the GPU box) with all host cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GENOME_BINS = 15_500_000
CONFIGS = {
    # name: (bins per GPU, biosamples, states, description)
    "s2_genome_833": (GENOME_BINS, 833, 18, "S2 whole-genome 15.5M bins x 833 biosamples, 18-state (BASELINE configs[1])"),
    "s2_genome_127": (GENOME_BINS, 127, 15, "S2 whole-genome 15.5M bins x 127 biosamples, 15-state (BASELINE configs[4])"),
    "s1_chr1_833": (1_246_253, 833, 18, "S1 chr1 1,246,253 bins x 833 biosamples, 18-state (shape of BASELINE configs[0])"),
    "s3_chr1_833": (1_250_000, 833, 18, "S3 chr1 1.25M bins x 833 biosamples, 18-state (BASELINE configs[2])"),
    "paired_s1_chr1": (1_250_000, 833, 18, "paired S1 chr1 1.25M bins, groups of 400 + 433 biosamples, 18-state, 1000 null "
                                            "permutations per bin (BASELINE configs[3])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="s2_genome_833", choices=sorted(CONFIGS))
    ap.add_argument("--bins", type=int, default=0, help="override bins per GPU (debug)")
    ap.add_argument("--kind", default="realistic", choices=["realistic", "uniform"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's row-loop port of expected+scores, all host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    from oracle import epilogos_oracle as orc
    bins, cols, k, saliency, seed, phase, exp = args
    x = orc.synth_states(bins, cols, k, seed)
    if phase == "expected":
        if saliency == 1:
            return orc.s1_expected_counts_rowloop(x, k)
        return orc.s2_expected_counts_rowloop(x, k)
    if saliency == 1:
        return orc.s1_scores_rowloop(x, k, exp).sum()
    return orc.s2_scores_rowloop(x, k, exp).sum()


def cpu_port_pass(pool, cores, bins_per_core, cols, k, saliency, seed0):
    """expected.main -> expectedCombination.main -> scores.main on cores*bins_per_core bins, one row range
    per worker process (the reference's multiprocessing.Pool fan-out, expected.py:71-79, scores.py:147-154).
    Returns seconds spent in the two compute phases (synthetic chunk generation included in the workers is
    measured separately and subtracted)."""
    from oracle import epilogos_oracle as orc
    jobs = [(bins_per_core, cols, k, saliency, seed0 + i, "expected", None) for i in range(cores)]
    t0 = time.perf_counter()
    parts = pool.map(_cpu_worker, jobs)
    exp = orc.normalize_expected(sum(parts))
    jobs = [(bins_per_core, cols, k, saliency, seed0 + i, "scores", exp) for i in range(cores)]
    pool.map(_cpu_worker, jobs)
    return time.perf_counter() - t0


def _gen_worker(args):
    from oracle import epilogos_oracle as orc
    bins, cols, k, seed = args
    return int(orc.synth_states(bins, cols, k, seed).sum())


def cpu_baseline(cols, k, saliency, budget_s=12.0, steps=1, warmup=0):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # calibrate on a small pass, then size the sample for ~budget_s
        probe = 64
        t = cpu_port_pass(pool, cores, probe, cols, k, saliency, 1000)
        tg0 = time.perf_counter()
        pool.map(_gen_worker, [(probe, cols, k, 1000 + i) for i in range(cores)])
        tgen = 2 * (time.perf_counter() - tg0)
        rate = probe / max(t - tgen, 1e-3)
        per_core = int(max(64, min(20000, rate * budget_s)))
        times = []
        for it in range(warmup + steps):
            tt = cpu_port_pass(pool, cores, per_core, cols, k, saliency, 2000 + 100 * it)
            tg0 = time.perf_counter()
            pool.map(_gen_worker, [(per_core, cols, k, 2000 + 100 * it + i) for i in range(cores)])
            tgen = 2 * (time.perf_counter() - tg0)
            if it >= warmup:
                times.append(max(tt - tgen, 1e-6))
    total = per_core * cores
    sec = sum(times) / len(times)
    return dict(value=total / sec, unit="bins/s", cores=cores, kind="port",
                sample="%d bins x %d biosamples x %d states (S%d expected+scores, %d worker processes x %d bins, "
                       "row-loop port of expected.py/scores.py, I/O excluded)" % (total, cols, k, saliency, cores,
                                                                                  per_core)), sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bins, cols, k, desc = CONFIGS[args.config]
    if args.config.startswith("paired"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm times the headline S1/S2 configurations only"}),
              flush=True)
        return
    saliency = int(args.config[1])
    if saliency == 3:
        print(json.dumps({"impl": "reference", "unavailable": "S3 row-loop port needs ~1 s per bin at 833 biosamples; "
                          "see profiles/ for the sampled figure"}), flush=True)
        return
    base, sec = cpu_baseline(cols, k, saliency, budget_s=6.0, steps=max(1, args.steps), warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "bins/sec for expected+scores (S%d)" % saliency, "value": base["value"],
        "unit": "bins/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic",
        "config": {"workload": desc, "biosamples": cols, "states": k, "saliency": saliency,
                   "note": "each step is a bounded sample of the workload on the host CPU"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling through NVML while the timed region runs
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons of one GPU while the timed region runs, taken by a SEPARATE `nvidia-smi -lms` process
    (the recipe's clocks line).  An in-process NVML poller was measured to cost 0.25 ms per 3 ms step at 2 GPUs: its
    driver calls serialise with rank 0's kernel launches and the other rank then waits at the all-reduce."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_power_brake_slowdown")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap", "hw_power_brake_slowdown"]

    def __init__(self, index):
        import subprocess
        import tempfile
        self.t_lo = self.t_hi = None
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        period = os.environ.get("EPI_BENCH_CLOCK_PERIOD_MS", "20")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", period],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def start(self):
        """Call right before the timed region (the process is already sampling)."""
        self.t_lo = time.time()

    def stop(self):
        self.t_hi = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "power_w_max": None}
        time.sleep(0.03)                     # let the sample that straddles the end land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        import datetime
        rows = []
        for line in self.out.read().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) != 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, int(float(f[1])), int(float(f[2])), float(f[3]), [x.lower().startswith("active") for x in f[4:]]))
            except ValueError:
                continue
        self.out.close()
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        inside = [r for r in rows if self.t_lo - 0.005 <= r[0] <= self.t_hi + 0.005]
        if not inside and rows:              # region shorter than the sampling period: take the nearest sample
            mid = 0.5 * (self.t_lo + self.t_hi)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
        clk = sorted(r[1] for r in inside)
        reasons = sorted({n for r in inside for n, on in zip(self.NAMES, r[4]) if on})
        return {"sm_mhz": clk[len(clk) // 2] if clk else None, "sm_max_mhz": inside[0][2] if inside else None,
                "reasons": reasons, "samples": len(inside), "power_w_max": max((r[3] for r in inside), default=None),
                "how": "nvidia-smi -lms %s in a separate process, samples inside the timed region" %
                       os.environ.get("EPI_BENCH_CLOCK_PERIOD_MS", "20")}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from epilogos_b200 import engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        from epilogos_b200 import dist as edist
        edist.bind_to_gpu_numa(local)          # pinned host buffers and copies on the GPU's own socket
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bins, cols, k, desc = CONFIGS[args.config]
    if args.bins:
        bins = args.bins
    engine.device_info()
    if args.config.startswith("paired"):
        return run_ours_paired(args, engine, synth, dist, world, rank, local, bins, cols, k, desc)
    saliency = int(args.config[1])
    if saliency == 3:
        return run_ours_s3(args, engine, synth, dist, world, rank, local, bins, cols, k, desc)

    x = synth.synth_states_device(bins, cols, k, seed=1234 + rank, kind=args.kind)
    cnt = torch.empty((bins, k), dtype=torch.int16, device="cuda")
    scores = torch.empty((bins, k), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    k1_ev = []

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        engine.bin_counts(x, cols, k, out=cnt)                               # K1
        if timed:
            e1.record(stream)
            k1_ev.append((e0, e1))
        n1, n2 = engine.expected_tables(cnt, cols, want_s1=saliency == 1, want_s2=saliency == 2)   # K2
        n = n1 if saliency == 1 else n2
        if world > 1 and not os.environ.get("EPI_BENCH_SKIP_ALLREDUCE"):
            dist.all_reduce(n)                                               # the path's only exchange
        e = engine.normalize(n)                                              # K4 (2 kernels)
        if saliency == 1:
            engine.scores_s1(cnt, cols, e, out32=scores)                     # K5
        else:
            engine.scores_s2(cnt, cols, e, out32=scores)
        return e

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("EPI_BENCH_NO_CLOCKS") else None
    for _ in range(max(3, args.warmup)):
        step(False)
    torch.cuda.synchronize()
    if rank == 0:
        time.sleep(0.3)                      # nvidia-smi needs a moment to start sampling
    if world > 1:
        dist.barrier()
    for _ in range(3):                       # every rank: back under load after the pause (collectives stay matched)
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step(not os.environ.get("EPI_BENCH_NO_K1_EVENTS"))
    t1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    k1_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in k1_ev) / max(1, len(k1_ev)) if k1_ev else 1.0], device="cuda",
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = bins * world / (ms_per_step * 1e-3)
    k1_ms = float(k1_ms.item())

    # ---- end to end through the host-buffer C-ABI call ----
    e2e = None
    if not args.no_e2e:
        host = None
        e2e_bins = bins
        while host is None and e2e_bins >= 1024:
            try:
                host = torch.empty((e2e_bins, x.shape[1]), dtype=torch.int8, pin_memory=True)
            except RuntimeError:
                e2e_bins //= 2
        host.copy_(x[:e2e_bins])
        torch.cuda.synchronize()
        engine.single_host(host, cols, k, saliency)                          # warm-up (allocations)
        if world > 1:
            dist.barrier()
        times = []
        for _ in range(args.e2e_steps):
            tic = time.perf_counter()
            engine.single_host(host, cols, k, saliency)
            times.append(time.perf_counter() - tic)
        t = torch.tensor([sum(times) / len(times)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ntab = k if saliency == 1 else k * k
        e2e = {"value": e2e_bins * world / float(t.item()), "unit": "bins/s",
               "h2d_bytes_per_step": e2e_bins * int(x.shape[1]), "d2h_bytes_per_step": e2e_bins * k * 4 + ntab * 12,
               "bins_per_gpu": e2e_bins, "steps": args.e2e_steps,
               "api": "epi_single_host (C ABI, pinned host matrix in, tables + float32 scores out)"}
        del host

    if rank == 0:
        peak, peak_src = peaks()
        achieved = bins * cols / (k1_ms * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists():
            try:
                per_bin = json.loads(tf.read_text()).get(args.config, {}).get("k1_dram_bytes_per_bin")
                traffic = per_bin * bins if per_bin else None
            except Exception:
                traffic = None
        step_alg = bins * (cols + 4 * k)
        line = {
            "metric": "bins/sec for expected+scores (S%d)" % saliency, "value": value, "unit": "bins/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
            "data": "synthetic",
            "config": {"workload": desc, "bins_per_gpu": bins, "biosamples": cols, "states": k, "saliency": saliency,
                       "distribution": args.kind, "parallelism": "bins sharded x%d, int64 table allreduce" % world,
                       "l2": "inputs (%.1f GB/GPU) larger than L2, no flush needed" % (bins * x.shape[1] / 1e9)},
            "roofline": {"bound": "hbm", "kernel": "k1_counts_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bins * cols, "ms_per_launch": k1_ms},
            "step_roofline": {"algorithmic_bytes_per_step": step_alg, "achieved": step_alg / (ms_per_step * 1e-3) / 1e9,
                              "frac": step_alg / (ms_per_step * 1e-3) / 1e9 / peak, "unit": "GB/s"},
            "e2e": e2e, "gpu_launches": 6 * args.steps, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_baseline(cols, k, saliency)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours_s3(args, engine, synth, dist, world, rank, local, bins, cols, k, desc):
    """S3 (BASELINE configs[2]): one-hot expansion + tcgen05 int8 Gram (chunked) + finalise + pair terms + scores.
    Parity-test configuration; reported with the tensor-pipe roofline of the Gram kernel."""
    import torch
    if args.bins:
        bins = args.bins
    x = synth.synth_states_device(bins, cols, k, seed=4321 + rank, kind=args.kind)
    plan = engine.s3_plan(bins, cols, k)
    tiles = torch.empty(plan["tile_bytes"] // 4, dtype=torch.int32, device="cuda")
    scores = torch.empty((bins, k), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    gram_ev = []

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        engine.s3_expected_tiles(x, cols, k, tiles=tiles)
        if timed:
            e1.record(stream)
            gram_ev.append((e0, e1))
        if world > 1:
            dist.all_reduce(tiles)
        _, exp3 = engine.s3_finalize(tiles, cols, k, plan["mp"], bins * world, want_counts=False, want_exp=True)
        terms = engine.s3_terms(exp3.reshape(-1), cols, k)
        engine.scores_s3(x, cols, k, terms, out32=scores)

    steps = min(args.steps, 3)
    sampler = ClockSampler(local) if rank == 0 else None
    step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(steps):
        step(True)
    t1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1) / steps, sum(a.elapsed_time(b) for a, b in gram_ev) / len(gram_ev)],
                      device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step, gram_ms = float(ms[0]), float(ms[1])
    if rank == 0:
        ck = cols * k
        useful = bins * ck * (ck + 1)
        peak_tops = 4500.0
        pf = ROOT / "profiles" / "int8_tensor_peak.json"
        src = "datasheet dense int8 4.5 POP/s (B200)"
        if pf.exists():
            try:
                peak_tops = float(json.loads(pf.read_text())["probe_issued_TOPS"])
                src = "measured: same kernel with operand loads disabled (profiles/int8_tensor_peak.json)"
            except Exception:
                pass
        line = {
            "metric": "bins/sec for expected+scores (S3)", "value": bins * world / (ms_per_step * 1e-3), "unit": "bins/s",
            "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 (tensor), f64 scores", "data": "synthetic",
            "config": {"workload": desc, "bins_per_gpu": bins, "biosamples": cols, "states": k, "saliency": 3,
                       "distribution": args.kind, "l2": "one-hot operand panels streamed per 131072-bin chunk"},
            "roofline": {"bound": "tensor", "kernel": "s3_onehot_kernel + s3_gram_kernel (chunked)",
                         "achieved": useful / (gram_ms * 1e-3) / 1e12, "peak": peak_tops, "unit": "TOP/s",
                         "frac": useful / (gram_ms * 1e-3) / 1e12 / peak_tops, "traffic": None, "peak_source": src,
                         "algorithmic_ops_per_launch": useful, "ms_per_launch": gram_ms},
            "e2e": None, "gpu_launches": steps * (2 * ((bins + engine.S3_CHUNK_BINS - 1) // engine.S3_CHUNK_BINS) + 4),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours_paired(args, engine, synth, dist, world, rank, local, bins, cols, k, desc):
    """Paired mode (BASELINE configs[3]): counts of both groups, expected table of the union (all-reduced), scores A / B,
    delta, quiescence mask, real distances, then P null shuffles per bin (multivariate hypergeometric draws from the group
    counts, scores of the shuffled groups, signed squared null distance).  The reference draws ONE shuffle per bin;
    P = 1000 is the north-star configuration.  Parity-test configuration; reported without a roofline fraction (the null
    stage is bound by the random draws, not by a memory or tensor pipe)."""
    import torch
    if args.bins:
        bins = args.bins
    c1, c2, nperm, batch = 400, 433, 1000, 20
    xa = synth.synth_states_device(bins, c1, k, seed=5 + 10 * rank)
    xb = synth.synth_states_device(bins, c2, k, seed=6 + 10 * rank)
    stream = torch.cuda.current_stream()

    def step():
        ca = engine.bin_counts(xa, c1, k)
        cb = engine.bin_counts(xb, c2, k)
        comb = (ca + cb).contiguous()
        n1, _ = engine.expected_tables(comb, c1 + c2, want_s2=False)
        if world > 1:
            dist.all_reduce(n1)
        e = engine.normalize(n1)
        sa, sb = engine.scores_s1(ca, c1, e), engine.scores_s1(cb, c2, e)
        delta, _ = engine.pairwise_combine(sa, sb, None, None)
        engine.quiescent_mask(ca, c1, cb, c2, k - 1)
        engine.pairwise_real_reduce(delta)
        done = 0
        while done < nperm:
            n = min(batch, nperm - done)
            oa, ob = engine.shuffled_counts_philox(ca, cb, c1, c2, 100 + done, n)
            na = engine.scores_s1(oa.reshape(-1, k), c1, e)
            nb = engine.scores_s1(ob.reshape(-1, k), c2, e)
            engine.pairwise_combine(None, None, na, nb)
            done += n

    steps = min(args.steps, 3)
    sampler = ClockSampler(local) if rank == 0 else None
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(steps):
        step()
    t1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_per_step = float(ms[0])
        line = {
            "metric": "bins/sec for paired expected+scores (S1) with %d null permutations per bin" % nperm,
            "value": bins * world / (ms_per_step * 1e-3), "unit": "bins/s", "n_gpus": world, "steps": steps, "warmup": 1,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64+f64", "data": "synthetic",
            "config": {"workload": desc, "bins_per_gpu": bins, "group_sizes": [c1, c2], "states": k, "saliency": 1,
                       "null_permutations": nperm, "bin_permutation_pairs_per_s": bins * world * nperm / (ms_per_step * 1e-3)},
            "roofline": None, "e2e": None, "gpu_launches": steps * (12 + (nperm // batch) * 6), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
