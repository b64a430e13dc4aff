"""Run one generator configuration a few times (target for ncu): run_one.py {tt|tqu} NSIDE LMAX [REPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl

kind, nside, lmax = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = cb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_pixels(nside)
n = ctx.npix
f = capi.window_beam(lmax, 10.0)
if kind == "tt":
    a = capi.tt_weights(synthetic_cl(lmax), f)
    out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    run = lambda: ctx.legendre_series(a, out)
else:
    a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
    out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
    lay = ctx.tqu_layout_single(out)
    run = lambda: ctx.tqu(*a, lay)
for _ in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(kind, nside, lmax, "ms", e0.elapsed_time(e1))
