"""Time the TT kernel (static vs shared-memory table): tt_bench.py NSIDE LMAX"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl
nside, lmax = int(sys.argv[1]), int(sys.argv[2])
ctx = cb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_pixels(nside)
n = ctx.npix; pairs = n * (n + 1) // 2
a = capi.tt_weights(synthetic_cl(lmax), capi.window_beam(lmax, 10.0))
out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
peak = ctx.measure_fp64_peak()
ref = None
VARS = [(int(x), "variant " + x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [(0, "static table"), (1, "shared table")]
for v, name in VARS:
    ctx.set_kernel_variant(v)
    ctx.legendre_series(a, out); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.legendre_series(a, out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts); tf = pairs * (lmax - 1) * 4 / (ms * 1e-3) / 1e12
    if ref is None: ref = out.clone(); d = 0.0
    else: d = float((out - ref).abs().max() / ref[0])
    print("TT nside %d lmax %d %-13s %.3f ms  %.2f TFLOP/s alg  %.1f%% of %.2f  (max rel diff %.1e)" % (nside, lmax, name, ms, tf, 100 * tf / peak, peak, d))
