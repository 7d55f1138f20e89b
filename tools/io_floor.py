"""Traffic floor of the score kernels: read uint16 [bins, K], write float32 [bins, K] with a library elementwise kernel
(the same bytes K5 moves), plus a pure fill of the output -- context for the K5 roofline."""
import sys
import torch

bins = int(sys.argv[1]) if len(sys.argv) > 1 else 15_500_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 18
x = torch.randint(0, 800, (bins, k), dtype=torch.int16, device="cuda")
y = torch.empty((bins, k), dtype=torch.float32, device="cuda")


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


gb = bins * k * 6 / 1e9
ms = t(lambda: torch.add(x, 1, out=y) if False else y.copy_(x))
print("convert int16->f32 (read %d B + write %d B per bin): %.3f ms, %.0f GB/s" % (2 * k, 4 * k, ms, gb / ms * 1e3))
ms = t(lambda: y.zero_())
print("fill f32 output (write %d B per bin): %.3f ms, %.0f GB/s" % (4 * k, ms, bins * k * 4 / 1e9 / ms * 1e3))
z = torch.empty_like(y)
ms = t(lambda: z.copy_(y))
print("copy f32 (read+write %d B per bin): %.3f ms, %.0f GB/s" % (8 * k, ms, bins * k * 8 / 1e9 / ms * 1e3))
