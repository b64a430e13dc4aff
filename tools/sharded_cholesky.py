"""The consumer's factorisation over the GPUs of one box: the [T;Q;U] matrix is generated over symmetry-orbit shards, the exchange
completes every rank's 36 strips, and multigpu.ShardedCholesky factorises it where it lies (nothing gathered).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_cholesky.py [nside ...] [--whole] [--group=S]
(also plain `python tools/sharded_cholesky.py 16` on one GPU).  Per Nside one JSON line on rank 0:
  generate / exchange / factorise / solve times (CUDA events, max over ranks), TFLOP/s of n^3 / 3 against the FP64 peak measured in
  the run, log det, and the checks:
    utu_max_rel_err      |(U^T U)[i, j] - A[i, j]| / sqrt(A_ii A_jj) over sampled entries (i, j) of this rank's columns, A saved
                         before the factorisation (max over ranks)
    whole_*              with --whole (n <= 36864): every rank also factorises the whole matrix on its own GPU with
                         cmg_packed_cholesky and compares its strips, log det and chi^2 with it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cosmopp_b200 as cb
from cosmopp_b200 import capi, multigpu, partition
from cosmopp_b200.synthetic import synthetic_cl


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    peak = ctx.measure_fp64_peak()
    whole_check = "--whole" in sys.argv
    nsides = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [16]
    group = ([int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--group=")] or [4])[0]

    def timed(fn):
        """ms on the device, max over ranks; barriers on both sides"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    for nside in nsides:
        lmax = 3 * nside
        ctx.set_pixels(nside)
        npix = ctx.npix
        n = 3 * npix
        f = capi.window_beam(lmax, 10.0)
        w = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
        bounds = partition.orbit_partition_blocks(nside, world)
        sh = multigpu.OrbitShardedTQU(ctx, nside, rank, world, bounds=bounds)
        gen_ms, _ = timed(lambda: sh.generate(w))
        exch_ms, _ = timed(sh.exchange)
        all_runs, ptrs = sh.chol_runs()
        mine = all_runs[rank]
        strips = sh.strips.tensor()
        # where column j of this rank starts inside the strips buffer
        run_off, at = [], 0
        for b, e in mine:
            run_off.append(at)
            at += partition.packed_size(e) - partition.packed_size(b)
        cols = torch.cat([torch.arange(b, e, device="cuda", dtype=torch.int64) for b, e in mine])
        col_start = torch.cat([run_off[k] + (torch.arange(b, e, device="cuda", dtype=torch.int64) * (torch.arange(b, e, device="cuda", dtype=torch.int64) + 1) // 2
                                             - partition.packed_size(b)) for k, (b, e) in enumerate(mine)])
        # white noise on the diagonal: the signal matrix alone is rank deficient (4 muK^2 in T, 0.09 in Q and U)
        strips[col_start + cols] += torch.where(cols < npix, 4.0, 0.09).double()
        # sampled entries (i <= j), both columns this rank's, kept for the U^T U check
        g = torch.Generator(device="cuda")
        g.manual_seed(1234 + rank)
        ns = 256
        a = torch.randint(0, cols.numel(), (ns,), device="cuda", generator=g)
        b_ = torch.randint(0, cols.numel(), (ns,), device="cuda", generator=g)
        ia, ib = torch.minimum(a, b_), torch.maximum(a, b_)
        si, sj = cols[ia], cols[ib]                                      # cols ascends, so si <= sj
        a_ij = strips[col_start[ib] + si].clone()
        a_ii = strips[col_start[ia] + si].clone()
        a_jj = strips[col_start[ib] + sj].clone()
        whole = None
        if whole_check and n <= 36864:
            whole = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
            ctx.tqu_orbit(*w, whole, 0)
            idx = torch.arange(n, device="cuda", dtype=torch.int64)
            whole[idx * (idx + 1) // 2 + idx] += torch.where(idx < npix, 4.0, 0.09).double()
            mine_before = torch.cat([whole[partition.packed_size(b):partition.packed_size(e)] for b, e in mine])
            strips_match = float((strips - mine_before).abs().max())
            del mine_before
        ch = multigpu.ShardedCholesky(ctx, n, all_runs, rank, ptrs, group=group)
        launches0 = ctx.launches
        ahead = False if "--no-lookahead" in sys.argv else None
        fact_ms, info = timed(lambda: ch.factorise(lookahead=ahead))
        launches = ctx.launches - launches0
        logdet = ch.logdet()
        rhs = torch.from_numpy(np.random.RandomState(77).normal(size=(2, n))).cuda()
        rhs0 = rhs.clone()
        solve_ms, _ = timed(lambda: ch.solve(rhs))
        chi2 = (rhs * rhs).sum(dim=1)
        # (U^T U)[i, j] = sum_{r <= i} U[r, i] U[r, j]
        err = 0.0
        ia_h, ib_h, si_h = ia.tolist(), ib.tolist(), si.tolist()
        cs = col_start.tolist() if cols.numel() <= 40000 else None
        for k in range(ns):
            ci = int(col_start[ia_h[k]]) if cs is None else cs[ia_h[k]]
            cj = int(col_start[ib_h[k]]) if cs is None else cs[ib_h[k]]
            m = si_h[k] + 1
            v = torch.dot(strips[ci:ci + m], strips[cj:cj + m])
            err = max(err, float(abs(v - a_ij[k]) / torch.sqrt(a_ii[k] * a_jj[k])))
        t = torch.tensor([err], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line = {"nside": nside, "n": n, "n_gpus": world, "group": group, "lookahead": "--no-lookahead" not in sys.argv, "packed_gb": capi.packed_size(n) * 8e-9, "strips_gb_this_rank": strips.numel() * 8e-9,
                "blocks_per_rank": int(np.bincount(ch.owners, minlength=world)[rank]), "info": info,
                "generate_ms": gen_ms, "exchange_ms": exch_ms, "factorise_ms": fact_ms, "solve_2rhs_ms": solve_ms,
                "kernel_launches_this_rank": launches, "tflops_all_gpus": n ** 3 / 3.0 / (fact_ms * 1e-3) / 1e12, "fp64_peak_tflops_per_gpu": peak,
                "frac_of_peak": n ** 3 / 3.0 / (fact_ms * 1e-3) / 1e12 / (peak * world), "logdet": logdet, "chi2": chi2.tolist(),
                "utu_max_rel_err": float(t.item()), "utu_samples_per_rank": ns}
        if whole is not None:
            info_w = ctx.packed_cholesky(whole, n)
            mine_after = torch.cat([whole[partition.packed_size(b):partition.packed_size(e)] for b, e in mine])
            d = torch.tensor([float((strips - mine_after).abs().max() / mine_after.abs().max())], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(d, op=dist.ReduceOp.MAX)
            logdet_w = ctx.packed_cholesky_logdet(whole, n)
            ctx.packed_cholesky_solve(whole, n, rhs0, 2)
            chi2_w = (rhs0 * rhs0).sum(dim=1)
            line.update({"whole_info": info_w, "whole_strips_before_max_abs_diff": strips_match, "whole_factor_max_rel_diff": float(d.item()),
                         "whole_logdet_rel_diff": abs(logdet - logdet_w) / abs(logdet_w),
                         "whole_chi2_rel_diff": float(((chi2 - chi2_w).abs() / chi2_w).max())})
            del whole, mine_after
        if rank == 0:
            print(json.dumps(line), flush=True)
        del ch, strips, rhs, rhs0
        sh.close()
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
