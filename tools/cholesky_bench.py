"""Packed in-place Cholesky (cmg_packed_cholesky) timed next to cusolverDnDpotrf on the same matrix (unpacked for cuSOLVER).
    python tools/cholesky_bench.py [nside ...]        T,Q,U matrix of each Nside (lmax = 3 nside) + white noise
The 147456-dimensional matrix of Nside = 64 (87 GB packed) is factorised in place; cuSOLVER cannot take it (174 GB unpacked,
n > 46340)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl

ctx = cb.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
peak = ctx.measure_fp64_peak()
groups = [int(g) for a in sys.argv[1:] if a.startswith("--groups=") for g in a.split("=")[1].split(",")] or [0]
ctx.set_cholesky_lookahead("--no-lookahead" not in sys.argv)
for nside in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [16, 32]:
    lmax = 3 * nside
    ctx.set_pixels(nside)
    n = 3 * ctx.npix
    f = capi.window_beam(lmax, 10.0)
    w = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
    d = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    ctx.tqu_orbit(*w, d, 0)
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    diag = idx * (idx + 1) // 2 + idx
    d[diag] += torch.where(idx < n // 3, 4.0, 0.09).double()           # white noise: the signal matrix alone is rank deficient
    torch.cuda.synchronize()
    dense_ms = None
    if n <= 40000:
        full = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        iu = torch.triu_indices(n, n, device="cuda")
        full[iu[0], iu[1]] = d[iu[1] * (iu[1] + 1) // 2 + iu[0]]
        full = full + full.T - torch.diag(torch.diagonal(full))
        torch.linalg.cholesky(full)                                    # warm-up (cuSOLVER handle, workspace)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L = torch.linalg.cholesky(full); e1.record(); torch.cuda.synchronize()
        dense_ms = e0.elapsed_time(e1)
        want_logdet = 2.0 * float(torch.log(torch.diagonal(L)).sum())
        del full, L, iu
        torch.cuda.empty_cache()
    for group in groups:
        ctx.set_cholesky_group(group)
        work = d.clone() if n <= 40000 else d
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if n <= 40000:
            ctx.packed_cholesky(work, n); work.copy_(d)                    # warm-up
        torch.cuda.synchronize()
        e0.record(stream); info = ctx.packed_cholesky(work, n); e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        logdet = ctx.packed_cholesky_logdet(work, n)
        flop = n ** 3 / 3.0
        line = {"nside": nside, "n": n, "group": group, "lookahead": "--no-lookahead" not in sys.argv, "packed_gb": capi.packed_size(n) * 8e-9, "info": info, "packed_cholesky_ms": ms,
                "tflops": flop / (ms * 1e-3) / 1e12, "fp64_peak_tflops": peak, "frac_of_peak": flop / (ms * 1e-3) / 1e12 / peak,
                "cusolver_potrf_ms_on_unpacked": dense_ms, "logdet": logdet,
                "logdet_rel_diff_vs_cusolver": (abs(logdet - want_logdet) / abs(want_logdet)) if dense_ms is not None else None}
        print(json.dumps(line), flush=True)
    del d, work
    torch.cuda.empty_cache()
