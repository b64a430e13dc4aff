#!/usr/bin/env python
"""Benchmark of the C-matrix hot path (BASELINE.json): pixel-pair*l/s, FP64 roofline fraction, ms per matrix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one generation of the whole covariance matrix of the workload from a C_l set.  Default
workload is BASELINE config 5 (polarized T,Q,U matrix, HEALPix Nside=64, lmax=192: 147456 x 147456,
87 GB packed); it fits one B200, and for N > 1 the SAME matrix is split over the ranks (strong scaling, no
data-path collective: every entry depends on replicated inputs).  Full-sky T,Q,U workloads run over the symmetry
orbits of the HEALPix grid (cmg_tqu_orbit: the four sums of a pixel pair are evaluated once per orbit under the
pi/2 rotation about the pole and stored at all four images -- a quarter of the recurrence work for the same matrix);
a rank then owns an in-face column range of all twelve base faces.  --no-orbit runs every pair (cmg_tqu, equal-area
pixel-column blocks).

Numbers printed (one JSON line from rank 0):
  value     pixel-pair*l/s over the K timed steps, inputs resident in HBM (geometry + C_l weights), CUDA events
            on the launch stream, max over ranks;
  e2e       the same metric through the reference-facing call with HOST buffers: C_l from pinned host memory
            in, packed matrix (this rank's shard) copied back to pinned host memory, copies inside the timing;
  roofline  algorithmic FLOP (20 per EVALUATED pixel-pair*l for T,Q,U; 4 for TT) / kernel time vs the FP64 peak
            measured by cmg_measure_fp64_peak on this GPU (MEASURED_PEAKS.json has no FP64 entry); on the orbit path only
            a quarter of the pairs are evaluated, `value` counts all pairs of the matrix that was produced;
  cpu_baseline  the reference's own object code (oracle/_ref, TT generator) on a bounded sample, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, nside, lmax, masked)
    "tqu_nside64_lmax192": ("tqu", 64, 192, False),     # BASELINE configs[4]  (default; the metric's target config)
    "tt_nside32_lmax96": ("tt", 32, 96, False),         # configs[2]
    "tqu_nside16_lmax47_masked": ("tqu", 16, 47, True), # configs[1]
    "tt_nside16_lmax47": ("tt", 16, 47, False),         # configs[0]
    "tqu_nside32_lmax96": ("tqu", 32, 96, False),
    "batched_x1024_tqu_nside16_lmax47": ("batched", 16, 47, 1024),   # configs[3]: kind, nside, lmax, batch
}
FLOP_PER_UNIT = {"tt": 4.0, "tqu": 20.0}        # algorithmic FLOP per pixel-pair*l
FWHM = 10.0
UNIT = "pixel-pair*l/s"


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, w in zip(sm, power) if w > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_geometry(name):
    kind, nside, lmax, masked = WORKLOADS[name]
    if kind == "batched":
        masked = False                    # the fourth field is the batch size there
    good = None
    if masked:
        # the reference test's deterministic mask (source/test_like_low.cpp:99-118), committed as a golden list
        path = os.path.join(ROOT, "tests", "golden", "like_low_good_pixels_nside%d.npy" % nside)
        good = np.load(path)
    npix = 12 * nside * nside if good is None else len(good)
    return kind, nside, lmax, good, npix


# ------------------------------------------------------------------------------------------ reference arm / CPU baseline

def reference_sample(kind, nside, lmax, budget_s=12.0, threads=None):
    """Reference object code (oracle/_ref: CMatrixGenerator::clToCMatrix compiled from /root/reference, TT path) on a
    bounded sample: `threads` concurrent instances, each one full small matrix over its own subset of the workload's
    pixels, same lmax / beam / synthetic C_l.  Returns pixel-pair*l/s aggregate."""
    from oracle import api
    from cosmopp_b200.synthetic import synthetic_cl
    threads = threads or os.cpu_count() or 1
    cl = synthetic_cl(lmax)
    have_ref = api.have_ref()
    # per-pair cost of the reference grows like lmax^2 (Legendre restarted for every l): calibrate on 40 pixels
    probe = np.arange(40, dtype=np.int32)
    t0 = time.perf_counter()
    (api.ref_cl_to_cmatrix if have_ref else lambda c, n, f, good: api.cl_to_cmatrix(c, n, f, good=good, literal=True))(cl, nside, FWHM, good=probe)
    per_pair = (time.perf_counter() - t0) / (40 * 41 / 2)
    n_sub = int(max(48, min(12 * nside * nside // threads, np.sqrt(2 * budget_s / per_pair))))
    subsets = [np.arange(t * n_sub, (t + 1) * n_sub, dtype=np.int32) % (12 * nside * nside) for t in range(threads)]
    done = [0.0] * threads

    def work(t):
        if have_ref:
            api.ref_cl_to_cmatrix(cl, nside, FWHM, good=subsets[t])
        else:
            api.cl_to_cmatrix(cl, nside, FWHM, good=subsets[t], literal=True)
        done[t] = time.perf_counter()

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = max(done) - t0
    pairs = threads * n_sub * (n_sub + 1) // 2
    value = pairs * (lmax - 1) / wall
    note = ""
    if kind == "tqu":
        note = "; TT block only - the reference has no TE/EE/BB pixel generator, so its cost per pixel-pair*l is a lower bound"
    if kind == "batched":
        note = ("; the reference has no batched mode and no TE/EE/BB pixel generator: its cost per pixel-pair*l*element is that of "
                "generating every matrix of the batch on its own (TT block, a lower bound)")
    return {
        "value": value, "unit": UNIT, "cores": threads, "kind": "reference" if have_ref else "port",
        "sample": "%d concurrent instances of %s, each the full TT matrix of %d of the workload's pixels (%d pixel pairs in all), "
                  "lmax=%d, fwhm=%g deg, %.1f s wall%s" % (threads, "the reference's clToCMatrix object code (oracle/_ref)" if have_ref else
                                                          "the oracle's literal port of clToCMatrix", n_sub, pairs, lmax, FWHM, wall, note),
        "wall_s": wall,
    }


def run_reference_arm(args):
    rank, world, _ = dist_info()
    if rank != 0:
        return
    kind, nside, lmax, good, npix = workload_geometry(args.workload)
    for _ in range(min(args.warmup, 1)):
        reference_sample(kind, nside, lmax, budget_s=2.0)
    vals, walls = [], []
    samples = None
    for _ in range(args.steps):
        r = reference_sample(kind, nside, lmax, budget_s=min(12.0, max(1.5, 90.0 / max(args.steps, 1))))
        vals.append(r["value"]); walls.append(r["wall_s"]); samples = r
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "pixel_pair_ell_per_s", "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, kind, nside, lmax, npix, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": samples["cores"], "kind": samples["kind"], "sample": samples["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def batched_config(name, n_gpus):
    _, nside, lmax, n_batch = WORKLOADS[name]
    n = 12 * nside * nside
    return {"workload": name, "kind": "batched tqu", "nside": nside, "lmax": lmax, "npix": n, "matrix_dim": 3 * n, "n_batch": n_batch,
            "packed_bytes": n_batch * (3 * n * (3 * n + 1) // 2) * 8, "fwhm_deg": FWHM, "pixel_window": "1 (HEALPix window file unavailable offline)",
            "sharding": "batch axis over %d rank(s), no collective" % n_gpus,
            "l2": "every sub-batch of 256 matrices writes 87 GB (>> 126 MB L2) with streaming stores; nothing is re-read between steps"}


def config_dict(name, kind, nside, lmax, npix, n_gpus, shard_mode="outbox"):
    """the WORKLOAD (identical for the GPU arm and the reference arm); how the GPU arm goes about it is the line's `path`"""
    if kind == "batched":
        return batched_config(name, n_gpus)
    dim = npix * (3 if kind == "tqu" else 1)
    return {
        "workload": name, "kind": kind, "nside": nside, "lmax": lmax, "npix": npix, "matrix_dim": dim,
        "packed_bytes": 8 * dim * (dim + 1) // 2, "fwhm_deg": FWHM, "pixel_window": "1 (HEALPix window file unavailable offline)",
        "sharding": "one matrix split over %d rank(s), no collective" % n_gpus, "shard_mode": shard_mode if n_gpus > 1 else "single",
        "l2": "each step writes its whole output (>> 126 MB L2) with streaming stores; nothing is re-read between steps",
    }


def path_dict(kind, orbit, n_gpus):
    if orbit:
        return {"method": "symmetry orbits (cmg_tqu_orbit / cmg_legendre_series_orbit[_sharded]): one evaluation per orbit of pixel pairs under "
                          "the pi/2 rotation of the HEALPix grid, every image stored -- the same packed matrix",
                "sharding": "in-face column ranges of all 12 base faces (orbit-closed sets of pixel columns) over %d rank(s)" % n_gpus}
    return {"method": "every pixel pair evaluated (%s)" % ("cmg_tqu" if kind == "tqu" else "cmg_legendre_series"),
            "sharding": "equal-area pixel-column blocks over %d rank(s)" % n_gpus}


# ------------------------------------------------------------------------------------------ GPU arm

class SharedHostMatrix:
    """One packed matrix in host memory that every rank of the box maps (POSIX shared memory): the N>1 e2e leg delivers a real
    host-side matrix -- each rank writes its own packed columns of it.  The pages a rank's copies land in are page-locked
    (cmg_host_register), everything else is written by that rank's host threads."""

    def __init__(self, n_doubles, rank, world, dist, tag):
        import mmap
        self.path = "/dev/shm/cmg_bench_%s" % tag
        self.rank, self.world, self.dist = rank, world, dist
        nbytes = 8 * n_doubles
        ok = 1
        if rank == 0:
            try:
                fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
                os.posix_fallocate(fd, 0, nbytes)             # fails here, not with SIGBUS later, when /dev/shm is too small
                os.close(fd)
            except OSError as e:
                ok = 0
                self.error = str(e)
        ok = self._all_min(ok)
        self.array = None
        self.registered = []
        if not ok:
            if rank == 0 and os.path.exists(self.path):
                os.unlink(self.path)
            return
        fd = os.open(self.path, os.O_RDWR | os.O_CREAT, 0o600)    # exists already: rank 0 made it before the all-reduce above
        if os.fstat(fd).st_size < nbytes:
            os.ftruncate(fd, nbytes)
        self.map = mmap.mmap(fd, nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        os.close(fd)
        self.array = np.frombuffer(self.map, dtype=np.float64)
        self._barrier()
        if rank == 0:
            os.unlink(self.path)                              # the mappings keep it alive; nothing is left behind on a crash

    def _all_min(self, v):
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([v], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())

    def _barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def register(self, first, count):
        """page-lock the pages holding elements [first, first + count)"""
        from cosmopp_b200 import capi
        lo = (8 * first) // 4096 * 4096
        hi = min(self.array.nbytes, (8 * (first + count) + 4095) // 4096 * 4096)
        view = self.array[lo // 8:hi // 8]
        capi.host_register(view)
        self.registered.append(view)

    def close(self):
        from cosmopp_b200 import capi
        for v in self.registered:
            capi.host_unregister(v)
        self.registered = []


def spot_check(ctx, sharded, spectra, nside, n_samples, seed):
    """Untimed checker: n_samples entries of this rank's (complete) packed columns against the CPU oracle (oracle/api.py,
    tqu_pairs), error relative to the diagonal of the entry's block (TT for the T columns, QQ otherwise) -- the gate of the
    parity tests (1e-11), applied to what THIS run produced on THIS rank."""
    import torch
    from oracle import api
    from cosmopp_b200 import capi, partition
    F = nside * nside
    n = 12 * F
    rs = np.random.RandomState(seed)
    q0, q1 = sharded.q0, sharded.q1
    s = rs.randint(0, 3, n_samples)
    f = rs.randint(0, 12, n_samples)
    q = rs.randint(q0, q1, n_samples)
    col = s * n + f * F + q
    row = (rs.random_sample(n_samples) * (col + 1)).astype(np.int64)
    sizes = partition.orbit_strip_sizes(nside, q0, q1)
    starts = np.concatenate([[0], np.cumsum([sizes[a][b] for a in range(3) for b in range(12)])])
    first_col = s * n + f * F + q0
    idx = starts[s * 12 + f] + (col * (col + 1) // 2 - first_col * (first_col + 1) // 2) + row
    got = sharded.strips.tensor()[torch.from_numpy(idx).cuda()].cpu().numpy()
    X, a = row // n, row % n
    Y, b = col // n, col % n
    blocks = api.tqu_pairs(*spectra, nside, FWHM, a, b)
    want = blocks[np.arange(n_samples), X, Y]
    diag = api.tqu_pairs(*spectra, nside, FWHM, np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64))[0]
    scale = np.where(Y == 0, diag[0, 0], diag[1, 1])
    return float((np.abs(got - want) / scale).max())


def cholesky_leg(ctx, torch, dist, nside, weights, rank, world, peak_tflops):
    """--cholesky: the step behind the generator on the matrix where it lies.  The [T;Q;U] matrix is
    generated over orbit shards whose in-face ranges sit on the 128-column block grid, exchanged, given white noise on the
    diagonal (the signal matrix alone is rank deficient), and factorised IN PLACE on every rank's 36 strips by
    multigpu.ShardedCholesky (one GPU: the same kernels on one run of columns).  Device time, max over ranks; U^T U = A is
    checked on sampled entries of every rank's own columns (saved before the factorisation)."""
    from cosmopp_b200 import capi, multigpu, partition
    stream = torch.cuda.current_stream()
    npix = 12 * nside * nside
    n = 3 * npix
    sh = multigpu.OrbitShardedTQU(ctx, nside, rank, world, bounds=partition.orbit_partition_blocks(nside, world))
    sh.generate(weights)
    sh.exchange()
    all_runs, ptrs = sh.chol_runs()
    mine = all_runs[rank]
    strips = sh.strips.tensor()
    ar = lambda b, e: torch.arange(b, e, device="cuda", dtype=torch.int64)
    run_off, at = [], 0
    for b, e in mine:
        run_off.append(at)
        at += partition.packed_size(e) - partition.packed_size(b)
    cols = torch.cat([ar(b, e) for b, e in mine])
    col_start = torch.cat([run_off[k] + (ar(b, e) * (ar(b, e) + 1) // 2 - partition.packed_size(b)) for k, (b, e) in enumerate(mine)])
    strips[col_start + cols] += torch.where(cols < npix, 4.0, 0.09).double()
    g = torch.Generator(device="cuda")
    g.manual_seed(4321 + rank)
    ns = 128
    a = torch.randint(0, cols.numel(), (ns,), device="cuda", generator=g)
    b_ = torch.randint(0, cols.numel(), (ns,), device="cuda", generator=g)
    ia, ib = torch.minimum(a, b_), torch.maximum(a, b_)
    si, sj = cols[ia], cols[ib]
    a_ij, a_ii, a_jj = strips[col_start[ib] + si].clone(), strips[col_start[ia] + si].clone(), strips[col_start[ib] + sj].clone()
    ch = multigpu.ShardedCholesky(ctx, n, all_runs, rank, ptrs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launches
    e0.record(stream)
    info = ch.factorise()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    err = 0.0
    for k in range(ns):
        ci, cj, m = int(col_start[ia[k]]), int(col_start[ib[k]]), int(si[k]) + 1
        err = max(err, float(abs(torch.dot(strips[ci:ci + m], strips[cj:cj + m]) - a_ij[k]) / torch.sqrt(a_ii[k] * a_jj[k])))
    e = torch.tensor([err], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out = {"matrix_dim": n, "packed_bytes": 8 * capi.packed_size(n), "strip_bytes_this_rank": 8 * strips.numel(), "info": info, "ms": ms,
           "tflops_all_gpus": n ** 3 / 3.0 / (ms * 1e-3) / 1e12, "frac_of_fp64_peak": n ** 3 / 3.0 / (ms * 1e-3) / 1e12 / (peak_tflops * world),
           "kernel_launches_this_rank": ctx.launches - launches0, "log_det": ch.logdet(), "utu_max_rel_err": float(e.item()),
           "utu_samples_per_rank": ns,
           "how": "multigpu.ShardedCholesky on the rank's 36 orbit strips in place (cmg_chol_*: per block a 67 KB broadcast of U_kk and an all-reduce "
                  "of one plane of the dense panel; FP64 tensor-core trailing update per group of 4 blocks, the next group's blocks and collectives on a second stream beside it); nothing gathered"}
    del ch, strips
    sh.close()
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import cosmopp_b200 as cb
    from cosmopp_b200 import capi, partition
    from cosmopp_b200.synthetic import synthetic_cl

    rank, world, local = dist_info()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the generator has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if WORKLOADS[args.workload][0] == "batched":
        return run_batched_arm(args, torch, dist, cb, capi, partition, synthetic_cl)

    kind, nside, lmax, good, npix = workload_geometry(args.workload)
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pixels(nside, good)

    # full-sky T,Q,U: the symmetry-orbit path (needs whole 64 x 32 tiles inside a base face and at least one column tile per rank)
    use_orbit = (kind == "tqu" and good is None and nside >= 8 and 2 <= lmax <= 441 and not args.no_orbit
                 and nside * nside // 32 >= world and args.shard_mode == "outbox")
    # full-sky TT: the same orbits (cmg_legendre_series_orbit with transposed images on one GPU; several ranks take the plan
    # without them, cmg_legendre_series_orbit_sharded: no entry leaves the rank that evaluated it)
    if kind == "tt" and good is None and nside >= 16 and lmax <= 1023 and not args.no_orbit and nside * nside // 16 >= world:
        use_orbit = True
    orbit_mode = 0 if (kind == "tqu" or (world == 1 and nside >= 32)) else 1
    orbit_pairs_all = partition.orbit_pairs_in_range(0, nside * nside, nside * nside, orbit_mode) if use_orbit else None
    if kind == "tqu" and not use_orbit and 2 <= lmax <= 441:
        ctx.set_kernel_variant(142)          # pins the every-pair kernel for the whole-call e2e leg as well (0 = automatic routing)
    if kind == "tt" and not use_orbit and args.no_orbit:
        ctx.set_kernel_variant(142)          # any pinned variant switches the whole-call routing off; TT keeps its default static kernel
    f = capi.window_beam(lmax, FWHM)
    bounds = partition.column_partition(npix, world, align=32)
    a0, a1 = bounds[rank], bounds[rank + 1]
    my_pairs = partition.pairs_in_block(a0, a1)
    total_pairs = npix * (npix + 1) // 2
    units_total = total_pairs * (lmax - 1)
    orbit_sharded = use_orbit and kind == "tqu"
    spectra = None

    if kind == "tt":
        weights = capi.tt_weights(synthetic_cl(lmax), f)
        if not use_orbit:
            shard = torch.empty(partition.tt_shard_size(a0, a1), dtype=torch.float64, device="cuda")
            launch = lambda: ctx.legendre_series(weights, shard, a0, a1)
            pieces = [shard]
        else:
            from cosmopp_b200 import multigpu
            tt_sharded = multigpu.OrbitShardedTT(ctx, nside, rank, world)
            launch = lambda: tt_sharded.generate(weights)
            my_pairs = tt_sharded.pairs                  # pixel pairs EVALUATED (a quarter of those stored on one GPU, 1 / 3.2 on several)
            shard = tt_sharded.strips.tensor()
            pieces = [shard]
    else:
        from cosmopp_b200 import multigpu
        spectra = synthetic_cl(lmax, pol=True)
        weights = capi.tqu_weights(*spectra, f, f)
        if use_orbit:
            # --orbit-mode 3 (one GPU): the meridian mirror on top of the rotation, 13 instead of 18 face-pair units evaluated
            orbit_mode = args.orbit_mode if world == 1 else 0
            orbit_pairs_all = partition.orbit_pairs_in_range(0, nside * nside, nside * nside, orbit_mode)
            sharded = multigpu.OrbitShardedTQU(ctx, nside, rank, world, mode=orbit_mode, exchange=args.exchange)
            launch = lambda: sharded.generate(weights)
            lay = None
            my_pairs = sharded.pairs                      # pixel pairs this rank EVALUATES (a quarter of those it stores)
        else:
            mode = args.shard_mode if world > 1 else "outbox"
            sharded = multigpu.ShardedTQU(ctx, npix, rank, world, mode=mode)
            lay = sharded.layout
            launch = lambda: ctx.tqu(*weights, lay)
        pieces = [b.tensor() for b in sharded.pieces()]
    strip_bytes = 8 * (sharded.strips.n if orbit_sharded else sum(p.numel() for p in pieces))
    h2d_bytes = (len(weights) if kind == "tt" else 4 * (lmax + 1)) * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak_tflops = ctx.measure_fp64_peak()

    # ---- value: device-resident, K timed steps.  The clock sampler runs from the warm-up on, so that short timed
    # regions (a few ms at N=8) still get samples taken under this very load.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        launch()
    barrier()
    launches0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        launch()
    e1.record(stream)
    barrier()
    ms_local = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    if rank == 0 and len(sampler.lines) < 8:
        # timed region shorter than the sampling period: keep the same kernel running (untimed) until there are samples
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end and len(sampler.lines) < 12:
            launch()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
    ms_total = max_over_ranks(ms_local)
    ms_per_step = ms_total / args.steps
    value = units_total / (ms_per_step * 1e-3)

    # roofline of the dominant kernel on this rank: its own pairs over its own time
    kernel_ms = ms_local / args.steps
    flop = FLOP_PER_UNIT[kind] * my_pairs * (lmax - 1)
    achieved = flop / (kernel_ms * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1:
        entry = json.load(open(prof)).get(args.workload + ("_orbit" if use_orbit else ""))
        traffic = entry["bytes"] if entry else None
    hbm_peak = None
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm_peak = json.load(open(mp)).get("hbm_gbs")
    written = sum(p.numel() for p in pieces) * 8
    if orbit_sharded and world > 1:
        # strips have holes where another rank's outbox holds the entry: what the kernel writes is this rank's share of the
        # 9 entries per stored pair
        written = int(8 * 9 * total_pairs * (my_pairs / max(orbit_pairs_all, 1)))
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": traffic,
        "peak_source": "cmg_measure_fp64_peak: dependent-free DFMA chains on this GPU in this run (MEASURED_PEAKS.json holds no FP64 figure)",
        "algorithmic_flop_per_unit": FLOP_PER_UNIT[kind],
        "evaluated_pixel_pairs": my_pairs, "stored_pixel_pairs_all_ranks": total_pairs,
        "note": ("orbit path: FLOP counted for the pixel pairs actually evaluated (one per orbit); the matrix holds %.2fx as many"
                 % (total_pairs / max(orbit_pairs_all, 1))) if use_orbit else None,
        "hbm_write_gbs": written / (kernel_ms * 1e-3) / 1e9, "hbm_peak_gbs_measured": hbm_peak,
        "hbm_frac": (written / (kernel_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
    }

    # ---- N>1, orbit shards: the exchange that completes every rank's packed columns (block(r -> d) of the compact outboxes
    # over NVLink + orbitInboxScatterKernel), timed on its own after a generation
    exchange = None
    if orbit_sharded and world > 1:
        sharded.exchange()                         # warm-up: inbox allocation / IPC mapping, NCCL channels
        times = []
        for _ in range(min(args.steps, 5)):
            launch()
            barrier()
            x0 = torch.cuda.Event(enable_timing=True)
            x1 = torch.cuda.Event(enable_timing=True)
            x0.record(stream)
            sharded.exchange()
            x1.record(stream)
            barrier()
            times.append(max_over_ranks(x0.elapsed_time(x1)))
        sent = 8 * sum(sharded.send_counts)
        recv = 8 * sum(sharded.recv_counts)
        exchange = {"ms": float(np.median(times)), "bytes_sent_this_rank": sent, "bytes_received_this_rank": recv,
                    "gbs_in_per_gpu": recv / (float(np.median(times)) * 1e-3) / 1e9, "mode": sharded.exchange_mode,
                    "how": ("ncclAllToAll (torch all_to_all_single, uneven splits) of the destination blocks, then orbitInboxScatterKernel per sender"
                            if sharded.exchange_mode == "nccl" else
                            "orbitInboxScatterKernel reading every sender's outbox through CUDA-IPC mapped peer memory (NVLink loads), no staging")}

    # ---- untimed checker: sampled entries of this rank's complete columns against the CPU oracle
    parity = None
    if orbit_sharded and args.spot_check > 0:
        if world == 1:
            launch()
        torch.cuda.synchronize()
        parity = max_over_ranks(spot_check(ctx, sharded, spectra, nside, args.spot_check, 777 + rank))

    # ---- whole matrix resident on every GPU (NCCL broadcasts of the strips over NVLink): by default for N>1 where it fits
    gather = None
    gather_ok = (args.gather or (orbit_sharded and not args.no_gather)) and kind == "tqu" and world > 1
    if gather_ok:
        need = 8 * capi.packed_size(3 * npix)
        ok = torch.tensor([1 if torch.cuda.mem_get_info()[0] > need + (2 << 30) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            gather_ok = False
            gather = {"skipped": "the whole matrix does not fit next to this rank's shard (%.0f GB needed)" % (need / 1e9)}
    if gather_ok:
        full = torch.empty(capi.packed_size(3 * npix), dtype=torch.float64, device="cuda")
        sharded.gather_full(full)
        barrier()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        sharded.gather_full(full)
        g1.record(stream)
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1))
        inbound = 8 * capi.packed_size(3 * npix) - strip_bytes
        gather_err = None
        if orbit_sharded and args.spot_check > 0:
            # the gathered matrix on this GPU, anywhere in the triangle (mostly other ranks' columns), against the CPU oracle
            from oracle import api
            rs = np.random.RandomState(4242 + rank)
            m = max(1, args.spot_check // 5)
            col = rs.randint(0, 3 * npix, m)
            row = (rs.random_sample(m) * (col + 1)).astype(np.int64)
            got = full[torch.from_numpy(col * (col + 1) // 2 + row).cuda()].cpu().numpy()
            blocks = api.tqu_pairs(*spectra, nside, FWHM, row % npix, col % npix)
            diag = api.tqu_pairs(*spectra, nside, FWHM, np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64))[0]
            want = blocks[np.arange(m), row // npix, col // npix]
            gather_err = max_over_ranks(float((np.abs(got - want) / np.where(col // npix == 0, diag[0, 0], diag[1, 1])).max()))
        if orbit_sharded:
            gather = {"ms": gms, "bytes_per_gpu_in": inbound, "gbs_in_per_gpu": inbound / (gms * 1e-3) / 1e9, "nvlink_gbs_per_direction": 900.0,
                      "parity_max_err": gather_err,
                      "how": "after the exchange every strip is a complete contiguous piece of the packed triangle: one ncclBroadcast each, straight into place"}
        else:
            gather = {"ms": gms, "bytes_per_gpu_in": 8 * (capi.packed_size(3 * npix) - sum(sharded.plan["strips"])),
                      "how": "ncclBroadcast of every strip straight into place; outbox blocks via a scratch buffer + cmg_tqu_scatter_block"}
        del full
        torch.cuda.empty_cache()

    # ---- e2e: host C_l in, packed matrix in HOST memory out, copies inside the timed region
    e2e = None
    if world == 1:
        ctx.set_host_expand(args.host_expand)          # -1 = the library's default (automatic), 0 = one plain copy
    if not args.no_e2e:
        shared = None
        d2h_bytes = strip_bytes
        note = "C_l from pinned host memory, the packed matrix back to pinned host memory through the reference-facing whole call"
        if kind == "tqu" and world == 1:
            del pieces, lay
            launch = None
            sharded.close()
            torch.cuda.empty_cache()
            host = torch.empty(capi.packed_size(3 * npix), dtype=torch.float64, pin_memory=True)
            spectra_pinned = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in spectra]
            step = lambda: ctx.cl_to_cmatrix_pol(*[s.numpy() for s in spectra_pinned], FWHM, host)      # reference-facing whole call
            if good is None and args.host_expand != 0 and host.numel() * 8 >= (1 << 30):
                F = nside * nside
                d2h_bytes = 8 * sum(partition.packed_size(s * npix + (fc + 1) * F) - partition.packed_size(s * npix + fc * F)
                                    for s in range(3) for fc in (3, 7, 11))
                note += "; only the columns of base faces 3, 7, 11 cross PCIe, host threads fill in the rotated images (cmg_set_host_expand)"
        elif kind == "tt" and world == 1:
            del shard, pieces
            torch.cuda.empty_cache()
            host = torch.empty(capi.packed_size(npix), dtype=torch.float64, pin_memory=True)
            cl_pinned = torch.from_numpy(synthetic_cl(lmax)).pin_memory()
            step = lambda: ctx.cl_to_cmatrix(cl_pinned.numpy(), FWHM, host)
            if good is None and args.host_expand != 0 and host.numel() * 8 >= (1 << 30):
                d2h_bytes = 8 * sum(partition.packed_size((fc + 1) * nside * nside) - partition.packed_size(fc * nside * nside) for fc in (3, 7, 11))
        elif orbit_sharded:
            # every rank writes its own packed columns of ONE host matrix in shared memory: generation, exchange, the columns of
            # base faces 3, 7, 11 over PCIe, the rotated images filled in by this rank's share of the host cores
            threads = args.host_expand if args.host_expand >= 0 else max(1, (os.cpu_count() or 1) // world)
            shared = SharedHostMatrix(capi.packed_size(3 * npix), rank, world, dist, "%d" % os.getppid())
            if shared.array is None:
                note = "no shared host matrix (/dev/shm too small): each rank copies its strips into private pinned memory"
                host_priv = torch.empty(sharded.strips.n, dtype=torch.float64, pin_memory=True)

                def step():
                    launch()
                    sharded.exchange()
                    host_priv.copy_(sharded.strips.tensor(), non_blocking=True)
                    torch.cuda.synchronize()
            else:
                F = nside * nside
                direct = max(args.direct_mask, 0) if threads > 0 else 0
                # (strip, face) runs of this rank's columns that cross PCIe: the last face of every ring, plus the direct images
                runs = [(s, fc) for s in range(3) for fc in range(12)
                        if threads == 0 or (fc & 3) == 3 or ((direct >> (3 * s + (3 - (fc & 3)) - 1)) & 1)]
                d2h_bytes = 0
                for s, fc in runs:
                    first = partition.packed_size(s * npix + fc * F + sharded.q0)
                    count = partition.packed_size(s * npix + fc * F + sharded.q1) - first
                    shared.register(first, count)
                    d2h_bytes += 8 * count
                note = ("every rank: generation + exchange, then its own packed columns of ONE host matrix in shared memory (%s)"
                        % ("columns of base faces 3, 7, 11 and the images of mask 0x%x over PCIe, the other rotated images filled in by %d host "
                           "threads per rank" % (direct, threads) if threads > 0 else "all 36 runs of packed columns copied"))

                def step():
                    launch()
                    sharded.exchange()
                    sharded.to_host(shared.array, threads, direct)
        else:
            hosts = [torch.empty(p.numel(), dtype=torch.float64, pin_memory=True) for p in pieces]

            def step():
                launch()
                for h, p in zip(hosts, pieces):
                    h.copy_(p, non_blocking=True)
                torch.cuda.synchronize()
        step()
        barrier()
        e2e_steps = min(args.steps, 10)          # each step moves the matrix into host memory (87 GB at Nside = 64): keep the run bounded
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": units_total * e2e_steps / wall, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * wall / e2e_steps, "steps": e2e_steps,
               "pcie_gbs_per_gpu": d2h_bytes / (wall / e2e_steps) / 1e9,
               "host_expand_threads": ((args.host_expand if args.host_expand >= 0 else "auto") if good is None else 0), "note": "per rank: " + note}
        if shared is not None and shared.array is not None:
            # the delivered host matrix against this rank's device strips (sampled), and rank 0 reads columns other ranks wrote
            rs = np.random.RandomState(99 + rank)
            F = nside * nside
            s_, f_, q_ = rs.randint(0, 3, 2000), rs.randint(0, 12, 2000), rs.randint(sharded.q0, sharded.q1, 2000)
            col = s_ * npix + f_ * F + q_
            row = (rs.random_sample(2000) * (col + 1)).astype(np.int64)
            sizes = partition.orbit_strip_sizes(nside, sharded.q0, sharded.q1)
            starts = np.concatenate([[0], np.cumsum([sizes[a][b] for a in range(3) for b in range(12)])])
            fcol = s_ * npix + f_ * F + sharded.q0
            idx = starts[s_ * 12 + f_] + (col * (col + 1) // 2 - fcol * (fcol + 1) // 2) + row
            dev = sharded.strips.tensor()[torch.from_numpy(idx).cuda()].cpu().numpy()
            same = float(np.abs(shared.array[col * (col + 1) // 2 + row] - dev).max())
            e2e["host_matrix_max_abs_diff_vs_device"] = max_over_ranks(same)
            shared.close()

    # ---- one GPU, matrices up to 2 GB: the whole chain a sampler step runs with the consumer on the device -- C_l from host
    # memory in, the matrix generated into device memory, + white noise, packed in-place Cholesky, chi^2 and log det of one map
    # back (16 bytes); the matrix never crosses PCIe
    consumer = None
    dim = npix * (3 if kind == "tqu" else 1)
    if world == 1 and not args.no_e2e and 8 * capi.packed_size(dim) <= (2 << 30):
        from cosmopp_b200.likelihood import Likelihood
        d_mat = torch.empty(capi.packed_size(dim), dtype=torch.float64, device="cuda")
        noise = np.zeros(capi.packed_size(dim))
        sig = np.where(np.arange(dim) < npix, 2.0, 0.3)
        noise[np.arange(dim) * (np.arange(dim) + 1) // 2 + np.arange(dim)] = sig ** 2
        d_noise = torch.from_numpy(noise).cuda()
        one_map = np.random.RandomState(3).normal(size=dim) * 5.0
        cl_in = synthetic_cl(lmax, pol=(kind == "tqu"))

        def chain():
            if kind == "tqu":
                ctx.cl_to_cmatrix_pol_dev(*cl_in, FWHM, d_mat)
            else:
                ctx.cl_to_cmatrix_dev(cl_in, FWHM, d_mat)
            like = Likelihood(ctx, d_mat, None, d_noise, dim)
            out = like.calculate(one_map)
            like.close()
            return out
        chain()
        torch.cuda.synchronize()
        reps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(reps):
            like_val = chain()
        torch.cuda.synchronize()
        ms_chain = 1e3 * (time.perf_counter() - t0) / reps
        consumer = {"ms_per_step": ms_chain, "value": units_total / (ms_chain * 1e-3), "unit": UNIT, "matrix_dim": dim, "h2d_bytes_per_step": h2d_bytes + 8 * dim,
                    "d2h_bytes_per_step": 16, "minus_two_log_like": like_val[0],
                    "note": "C_l and one map from host memory in; matrix generated on the device, + white noise, packed in-place Cholesky "
                            "(cmg_packed_cholesky), chi^2 and log det back: the matrix never crosses PCIe"}
        del d_mat, d_noise

    chol = None
    if args.cholesky and orbit_sharded and (nside * nside) % 128 == 0 and nside * nside // 128 >= world:
        if not (world == 1 and not args.no_e2e):      # (the one-GPU e2e leg has closed the shards already)
            try:
                del pieces
                sharded.close()
            except Exception:
                pass
        torch.cuda.empty_cache()
        chol = cholesky_leg(ctx, torch, dist, nside, weights, rank, world, peak_tflops)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = reference_sample(kind, nside, lmax, budget_s=12.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "pixel_pair_ell_per_s", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "ms_per_matrix": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args.workload, kind, nside, lmax, npix, world, args.shard_mode),
            "path": path_dict(kind, use_orbit, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "fp64_frac_of_peak": achieved / peak_tflops, "exchange": exchange, "gather": gather, "e2e_device_consumer": consumer, "consumer_cholesky": chol,
            "parity_max_err": parity,
            "parity_note": ("max over ranks of |entry - oracle| / diagonal of the block over %d sampled entries of each rank's packed columns "
                            "(oracle/api.py tqu_pairs, untimed; gate 1e-11)" % args.spot_check) if parity is not None else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_batched_arm(args, torch, dist, cb, capi, partition, synthetic_cl):
    """BASELINE configs[3]: one MCMC step = B synthetic C_l sets -> B polarized Nside=16, lmax=47 matrices (348 GB of output for
    B = 1024: produced in sub-batches of 256 per GPU into one slab buffer, as a consumer working through the proposals would).
    Batch axis sharded over the ranks, no collective.  FP64 tensor path (DMMA) over a basis shared by the batch: 8 FLOP per
    pixel-pair*l*element; the kernel sits on the ridge, so both bounds are reported."""
    rank, world, local = dist_info()
    _, nside, lmax, n_batch = WORKLOADS[args.workload]
    SUB = 256
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pixels(nside)
    n = ctx.npix
    pairs = n * (n + 1) // 2
    f = capi.window_beam(lmax, FWHM)
    bounds = partition.batch_partition(n_batch, world)
    b0, b1 = bounds[rank], bounds[rank + 1]
    a = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lmax, seed=12345 + b, pol=True), f, f)) for b in range(b0, b1)]) if b1 > b0 else None
    slabs = torch.empty(((min(SUB, max(b1 - b0, 1)) + 15) // 16) * capi.slab_doubles(3 * n), dtype=torch.float64, device="cuda")
    peak = ctx.measure_fp64_peak()

    def step():
        for s in range(0, b1 - b0, SUB):
            ctx.tqu_batched_slab(a[s:s + SUB], slabs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_local = e0.elapsed_time(e1) / args.steps
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms_local)
    units = n_batch * pairs * (lmax - 1)
    per_gpu_units = (b1 - b0) * pairs * (lmax - 1)
    out_bytes = (b1 - b0) * capi.packed_size(3 * n) * 8
    hbm_peak = None
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm_peak = json.load(open(mp)).get("hbm_gbs")
    achieved = 8.0 * per_gpu_units / (ms_local * 1e-3) / 1e12
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "algorithmic_flop_per_unit": 8.0,
                "note": "ridge point: 368 FLOP per 72 B written; FP64 (DMMA at the DFMA rate) and HBM write bound at once",
                "hbm_write_gbs": out_bytes / (ms_local * 1e-3) / 1e9, "hbm_peak_gbs_measured": hbm_peak,
                "hbm_frac": (out_bytes / (ms_local * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                "peak_source": "cmg_measure_fp64_peak on this GPU in this run; HBM: MEASURED_PEAKS.json copy bandwidth"}

    # e2e: the C_l weights of the whole step from pinned host memory in, one number per matrix back (the sum of its diagonal, what
    # a consumer on the device would hand on): nobody copies 348 GB of proposals' matrices to the host per MCMC step
    e2e = None
    if not args.no_e2e and b1 > b0:
        a_pinned = torch.from_numpy(a).pin_memory()
        diag_idx = torch.from_numpy(np.array([capi.packed_index(i, i) for i in range(3 * n)], dtype=np.int64)).cuda()
        sums_host = torch.empty(b1 - b0, dtype=torch.float64, pin_memory=True)
        slab_d = capi.slab_doubles(3 * n)

        def step_e2e():
            for s in range(0, b1 - b0, SUB):
                nb = min(SUB, b1 - b0 - s)
                ctx.tqu_batched_slab(a_pinned[s:s + nb].numpy(), slabs)
                view = slabs[:((nb + 15) // 16) * slab_d].view(-1, slab_d // 16, 16)          # [slab][entry][element]
                tr = view[:, diag_idx, :].sum(dim=1).reshape(-1)[:nb]
                sums_host[s:s + nb].copy_(tr, non_blocking=True)
            torch.cuda.synchronize()
        step_e2e()
        barrier()
        k = min(args.steps, 5)
        t0 = time.perf_counter()
        for _ in range(k):
            step_e2e()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": units * k / wall, "unit": UNIT.replace("l/s", "l*element/s"), "h2d_bytes_per_step": int(a.nbytes), "d2h_bytes_per_step": 8 * (b1 - b0),
               "ms_per_step": 1e3 * wall / k, "steps": k,
               "note": "per rank: its share of the step's C_l weights from pinned host memory, every matrix generated on the device, the trace of "
                       "each matrix back to pinned host memory (the matrices are consumed on the device)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = reference_sample("batched", nside, lmax, budget_s=12.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps({
            "metric": "pixel_pair_ell_per_s", "value": units / (ms * 1e-3), "unit": UNIT.replace("l/s", "l*element/s"), "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "ms_per_matrix": ms / n_batch, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": batched_config(args.workload, world),
            "path": {"method": "cmg_tqu_batched_slab: recurrences once per pixel pair, the l-sum over the batch as FP64 tensor-core (DMMA) "
                               "contractions, slab output (16 batch elements interleaved entry by entry)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "fp64_frac_of_peak": achieved / peak}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="tqu_nside64_lmax192", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shard-mode", default="outbox", choices=["outbox", "peer"],
                    help="N>1, T,Q,U: keep entries owned by another rank in local blocks (outbox) or write them into the owner's strip "
                         "through CUDA-IPC peer memory over NVLink (peer)")
    ap.add_argument("--gather", action="store_true", help="N>1, T,Q,U, every-pair shards: also time the NCCL gather of the whole matrix onto every GPU "
                                                          "(orbit shards: on by default where the matrix fits)")
    ap.add_argument("--no-gather", action="store_true", help="N>1, orbit shards: skip the gather of the whole matrix onto every GPU")
    ap.add_argument("--direct-mask", type=int, default=0, metavar="MASK",
                    help="N>1 e2e leg: images that cross PCIe next to the last-face columns instead of being filled in by host threads "
                         "(bit 3 strip + k - 1); default 0 (the host's memory write bandwidth bounds the delivery, whoever writes)")
    ap.add_argument("--exchange", default="pull", choices=["nccl", "pull"],
                    help="N>1, orbit shards: how block(r -> d) of the outboxes reaches rank d -- one NCCL all-to-all + a local scatter kernel, "
                         "or the scatter kernel reading the sender's outbox through CUDA-IPC peer memory")
    ap.add_argument("--spot-check", type=int, default=10000, metavar="N",
                    help="full-sky T,Q,U: entries per rank compared with the CPU oracle after the run (untimed checker; 0 = off)")
    ap.add_argument("--no-orbit", action="store_true", help="full-sky T,Q,U: evaluate every pixel pair (cmg_tqu) instead of one per symmetry orbit")
    ap.add_argument("--host-expand", type=int, default=-1, metavar="THREADS",
                    help="N=1, full sky, e2e leg: copy back only the last-face columns (27 %% of the matrix) and fill in the rotated images "
                         "on THREADS host threads (cmg_set_host_expand); -1 = the library's default (automatic), 0 = one plain copy of the whole matrix")
    ap.add_argument("--cholesky", action="store_true", help="full-sky T,Q,U over orbits: also factorise the matrix in place on the shards "
                    "(multigpu.ShardedCholesky; key consumer_cholesky; ~31 s on one GPU, ~4.2 s on eight for the Nside = 64 matrix)")
    ap.add_argument("--orbit-mode", type=int, default=0, choices=[0, 1, 2, 3], help="full-sky T,Q,U on one GPU: mode of cmg_tqu_orbit "
                    "(3 = with the meridian mirror: fewer evaluations, store-pattern bound; DESIGN.md)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
