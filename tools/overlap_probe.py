"""Does the expected-table kernel (K2, issue/IMAD bound) hide under the count kernel (K1, HBM bound) when the bins are
processed in chunks on two streams?"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

bins, cols, k = 15_500_000, 833, 18
x = synth.synth_states_device(bins, cols, k, seed=1)
cnt = torch.empty((bins, k), dtype=torch.int16, device="cuda")
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def sequential():
    engine.bin_counts(x, cols, k, out=cnt)
    return engine.expected_tables(cnt, cols, want_s1=False)[1]


def overlapped(nchunks):
    import ctypes
    from epilogos_b200 import _lib
    n2 = torch.zeros((k, k), dtype=torch.int64, device="cuda")
    step = (bins // nchunks + 4095) // 4096 * 4096
    side.wait_stream(main)
    for lo in range(0, bins, step):
        hi = min(bins, lo + step)
        engine.bin_counts(x[lo:hi], cols, k, out=cnt[lo:hi])
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            _lib.call("epi_expected_s1s2", ctypes.c_void_p(cnt[lo:hi].data_ptr()), hi - lo, k, cols, ctypes.c_void_p(0),
                      ctypes.c_void_p(n2.data_ptr()), ctypes.c_void_p(side.cuda_stream))
    main.wait_stream(side)
    return n2


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


ref = sequential()
print("sequential K1+K2: %.3f ms" % timeit(sequential))
for nch in (4, 8, 16, 32):
    assert torch.equal(overlapped(nch), ref)
    print("overlapped %2d chunks: %.3f ms" % (nch, timeit(lambda: overlapped(nch))))
