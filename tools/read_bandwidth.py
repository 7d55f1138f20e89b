import torch, json
n = 13_144_000_000 // 16 * 16
x = torch.empty(n, dtype=torch.uint8, device="cuda"); x.random_(0, 18)
res = {}
for name, v in (("int32", x.view(torch.int32)), ("int64", x.view(torch.int64)), ("float32", x.view(torch.float32))):
    for _ in range(3): v.sum()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); v.sum(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); res["sum_%s_GBps" % name] = n / (ts[len(ts)//2] * 1e-3) / 1e9
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); res["copy_GBps_read_plus_write"] = 2 * n / (ts[len(ts)//2] * 1e-3) / 1e9
print(json.dumps(res))
