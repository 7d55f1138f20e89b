"""Optional wall-clock accounting of the host-side stages of a CLI run (EPILOGOS_B200_TIMING=1): where a file-to-file run
spends its time either side of the kernels -- inflate + parse, host -> device, kernels (synchronised when timing is on),
device -> host, text formatting + deflate, npz hand-over files, region-of-interest selection.  Off by default: the
context manager is a no-op and nothing synchronises."""
import os
import time
from contextlib import contextmanager

ENABLED = os.environ.get("EPILOGOS_B200_TIMING", "") not in ("", "0")
totals = {}


def enable(on=True):
    global ENABLED
    ENABLED = bool(on)


def reset():
    totals.clear()


@contextmanager
def stage(name, sync_cuda=False):
    if not ENABLED:
        yield
        return
    if sync_cuda:
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if sync_cuda:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        totals[name] = totals.get(name, 0.0) + time.perf_counter() - t0


def report():
    return {k: round(v, 4) for k, v in sorted(totals.items(), key=lambda kv: -kv[1])}
