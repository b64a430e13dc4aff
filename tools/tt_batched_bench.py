"""Batched TT (temperature-only chains): one z-batched launch (shared-memory table) vs per-element static-table launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl
nside, lmax = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 47)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ctx = cb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_pixels(nside)
n = ctx.npix; packed = capi.packed_size(n)
f = capi.window_beam(lmax, 10.0)
a = np.stack([capi.tt_weights(synthetic_cl(lmax, seed=s), f) for s in range(B)])
out = torch.empty(B * packed, dtype=torch.float64, device="cuda")
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: ctx.legendre_series_batched(a, out, packed))
print("TT nside %d lmax %d B=%d  batched call: %.3f ms (%.4f ms/matrix)" % (nside, lmax, B, ms, ms / B))
ref = out.clone()
def loop():
    for b in range(B):
        ctx.legendre_series(a[b], out[b * packed:(b + 1) * packed])
ms = t(loop)
print("TT nside %d lmax %d B=%d  %d single static launches: %.3f ms (%.4f ms/matrix), max rel diff %.1e" % (nside, lmax, B, B, ms, ms / B, float((out - ref).abs().max() / ref[0])))
