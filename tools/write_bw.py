"""Write-only HBM bandwidth on this GPU (the batched mode's second roofline): torch fill_ / cudaMemsetAsync on a buffer
far larger than L2, and a strided-run pattern like the packed-triangle stores (64-byte .. 512-byte runs)."""
import sys, torch
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4 << 30     # doubles
x = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print("fill_ %d GB: %.2f ms  %.0f GB/s" % (n * 8 >> 30, ms, n * 8 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_ %d GB: %.2f ms  %.0f GB/s" % (n * 8 >> 30, ms, n * 8 / ms / 1e6))
y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
ms = t(lambda: y.copy_(x[: n // 2])); print("copy %d GB -> %d GB: %.2f ms  %.0f GB/s (read+write)" % (n * 4 >> 30, n * 4 >> 30, ms, n * 8 / ms / 1e6))
