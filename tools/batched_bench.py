"""Batched regeneration (BASELINE config 4 shape: polarized Nside=16, lmax=47): time B matrices per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl
nside, lmax = 16, 47
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = cb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_pixels(nside)
n = ctx.npix; pairs = n * (n + 1) // 2
f = capi.window_beam(lmax, 10.0)
a = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lmax, seed=12345 + b, pol=True), f, f)) for b in range(B)])
stride = capi.packed_size(3 * n)
kstride = 0 if os.environ.get('STRIDE0') else stride     # STRIDE0=1: every element overwrites matrix 0 (no DRAM streaming: SM-side time only)
out = torch.empty(B * stride, dtype=torch.float64, device="cuda")
peak = ctx.measure_fp64_peak()
for v in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 42, 901)):
    ctx.set_kernel_variant(v)
    ctx.tqu_batched(a, out, kstride); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.tqu_batched(a, out, kstride); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts); tf = B * pairs * (lmax - 1) * 20 / (ms * 1e-3) / 1e12
    print("batched TQU nside16 lmax47 B=%d variant %d: %.2f ms (%.3f ms/matrix) %.2f TFLOP/s alg(20/unit) %.1f%% of peak; HBM write %.0f GB/s" %
          (B, v, ms, ms / B, tf, 100 * tf / peak, B * stride * 8 / (ms * 1e-3) / 1e9))

# DMMA path, slab output
slabs = out[: ((B + 15) // 16) * capi.slab_doubles(3 * n)]
ctx.set_kernel_variant(0)
ctx.tqu_batched_slab(a, slabs); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.tqu_batched_slab(a, slabs); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts); tf8 = B * pairs * (lmax - 1) * 8 / (ms * 1e-3) / 1e12
print("batched TQU nside16 lmax47 B=%d slab/DMMA: %.2f ms (%.3f ms/matrix) %.2f TFLOP/s alg(8/unit, shared basis) %.1f%% of peak; HBM write %.0f GB/s" %
      (B, ms, ms / B, tf8, 100 * tf8 / peak, B * stride * 8 / (ms * 1e-3) / 1e9))
