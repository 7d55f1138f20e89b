import sys, os, json, torch
sys.path.insert(0, '/root/repo')
from epilogos_b200 import engine, synth
bins, cols, k = 15_500_000, 833, 18
x = synth.synth_states_device(bins, cols, k, seed=1)
cnt = torch.empty((bins, k), dtype=torch.int16, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]
res = {}
for st, ct in ((3, 2), (4, 2), (5, 2), (2, 3), (3, 3), (2, 4), (6, 1)):
    os.environ["EPI_K1_STAGES"] = str(st); os.environ["EPI_K1_CTAS"] = str(ct)
    try:
        res["s%d_c%d" % (st, ct)] = round(timeit(lambda: engine.bin_counts(x, cols, k, out=cnt)), 4)
    except Exception as e:
        res["s%d_c%d" % (st, ct)] = str(e)[:60]
print(json.dumps(res))
