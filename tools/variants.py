"""Time the T,Q,U kernel variants (scratch tool): variants.py NSIDE LMAX"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl

nside, lmax = int(sys.argv[1]), int(sys.argv[2])
ctx = cb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_pixels(nside)
n = ctx.npix
f = capi.window_beam(lmax, 10.0)
a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
lay = ctx.tqu_layout_single(out)
peak = ctx.measure_fp64_peak()
ref = None
pairs = n * (n + 1) // 2
for v in ([int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (22, 42, 81, 114, 122, 123, 124, 142)):
    ctx.set_kernel_variant(v)
    out.fill_(float("nan"))
    ctx.tqu(*a, lay); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.tqu(*a, lay); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    samp = out[:: max(1, out.numel() // 4000000)].clone()
    if ref is None:
        ref = samp
        err = 0.0
    else:
        err = float((samp - ref).abs().max() / ref[0])
    ms = min(ts)
    tf = pairs * (lmax - 1) * 20 / (ms * 1e-3) / 1e12
    print("variant %3d: %.3f ms  %.2f TFLOP/s alg  %.1f%% of measured peak %.2f   max rel diff vs v1 %.2e nan=%d" % (v, ms, tf, 100 * tf / peak, peak, err, int(torch.isnan(samp).sum())))

# TT at the same geometry
a = capi.tt_weights(synthetic_cl(lmax), f)
del out
out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
ctx.legendre_series(a, out); torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.legendre_series(a, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts); tf = pairs * (lmax - 1) * 4 / (ms * 1e-3) / 1e12
print("TT: %.3f ms  %.2f TFLOP/s alg  %.1f%% of measured peak" % (ms, tf, 100 * tf / peak))
