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


def config_dict(name, kind, nside, lmax, npix, n_gpus, shard_mode="outbox"):
    """the WORKLOAD (identical for the GPU arm and the reference arm); how the GPU arm goes about it is the line's `path`"""
    dim = npix * (3 if kind == "tqu" else 1)
    return {
        "workload": name, "kind": kind, "nside": nside, "lmax": lmax, "npix": npix, "matrix_dim": dim,
        "packed_bytes": 8 * dim * (dim + 1) // 2, "fwhm_deg": FWHM, "pixel_window": "1 (HEALPix window file unavailable offline)",
        "sharding": "one matrix split over %d rank(s), no collective" % n_gpus, "shard_mode": shard_mode if n_gpus > 1 else "single",
        "l2": "each step writes its whole output (>> 126 MB L2) with streaming stores; nothing is re-read between steps",
    }


def path_dict(kind, orbit, n_gpus):
    if orbit:
        return {"method": "symmetry orbits (cmg_tqu_orbit / cmg_legendre_series_orbit): one evaluation per orbit of pixel pairs under the "
                          "pi/2 rotation of the HEALPix grid, every image stored -- the same packed matrix",
                "sharding": "in-face column ranges of all 12 base faces (orbit-closed sets of pixel columns) over %d rank(s)" % n_gpus}
    return {"method": "every pixel pair evaluated (%s)" % ("cmg_tqu" if kind == "tqu" else "cmg_legendre_series"),
            "sharding": "equal-area pixel-column blocks over %d rank(s)" % n_gpus}


# ------------------------------------------------------------------------------------------ GPU arm

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

    kind, nside, lmax, good, npix = workload_geometry(args.workload)
    ctx = cb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pixels(nside, good)

    # full-sky T,Q,U: the symmetry-orbit path (needs whole 64 x 32 tiles inside a base face and at least one column tile per rank)
    use_orbit = (kind == "tqu" and good is None and nside >= 8 and 2 <= lmax <= 441 and not args.no_orbit
                 and nside * nside // 32 >= world and args.shard_mode == "outbox")
    # full-sky TT on one GPU: the same orbits without transposed images (cmg_legendre_series_orbit; no sharded form yet)
    if kind == "tt" and good is None and nside >= 16 and lmax <= 1023 and not args.no_orbit and world == 1:
        use_orbit = True
    orbit_mode = 0 if kind == "tqu" else 1
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

    if kind == "tt":
        weights = capi.tt_weights(synthetic_cl(lmax), f)
        shard = torch.empty(partition.tt_shard_size(a0, a1), dtype=torch.float64, device="cuda")
        launch = lambda: ctx.legendre_series(weights, shard, a0, a1)
        if use_orbit:
            launch = lambda: ctx.legendre_series_orbit(weights, shard)
            my_pairs = orbit_pairs_all                   # pixel pairs EVALUATED (1 / 3.2 of those stored)
        pieces = [shard]
    else:
        from cosmopp_b200 import multigpu
        spectra = synthetic_cl(lmax, pol=True)
        weights = capi.tqu_weights(*spectra, f, f)
        if use_orbit:
            sharded = multigpu.OrbitShardedTQU(ctx, nside, rank, world, mode=0)
            launch = lambda: sharded.generate(weights)
            lay = None
            my_pairs = sharded.pairs                      # pixel pairs this rank EVALUATES (a quarter of those it stores)
        else:
            mode = args.shard_mode if world > 1 else "outbox"
            sharded = multigpu.ShardedTQU(ctx, npix, rank, world, mode=mode)
            lay = sharded.layout
            launch = lambda: ctx.tqu(*weights, lay)
        pieces = [b.tensor() for b in sharded.pieces()]
    d2h_bytes = sum(p.numel() for p in pieces) * 8
    h2d_bytes = (len(weights) if kind == "tt" else 4 * (lmax + 1)) * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
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
    written = d2h_bytes
    if use_orbit and world > 1:
        # the outbox blocks are allocated dense; what the kernel writes is this rank's share of the 9 entries per stored pair
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

    # ---- optional: whole matrix resident on every GPU (NCCL broadcasts of the strips over NVLink)
    gather = None
    gather_ok = args.gather and kind == "tqu" and world > 1
    if gather_ok:
        need = 8 * capi.packed_size(3 * npix)
        if use_orbit:
            need += 8 * max(sharded.sizes_of(r)[1] for r in range(world))       # scratch for a peer's outbox buffer
        ok = torch.tensor([1 if torch.cuda.mem_get_info()[0] > need + (2 << 30) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            gather_ok = False
            gather = {"skipped": "the whole matrix plus the gather scratch do not fit next to this rank's shard (%.0f GB needed)" % (need / 1e9)}
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
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        if use_orbit:
            own = sharded.sizes_of(rank)
            gather = {"ms": float(gt.item()),
                      "bytes_per_gpu_in": 8 * (capi.packed_size(3 * npix) - own[0] + sum(sharded.sizes_of(r)[1] for r in range(world) if r != rank)),
                      "how": "ncclBroadcast of every strip straight into place; each rank's (dense) outbox buffer via a scratch buffer + "
                             "orbitOutboxScatterKernel (cmg_tqu_orbit_assemble)"}
        else:
            gather = {"ms": float(gt.item()), "bytes_per_gpu_in": 8 * (capi.packed_size(3 * npix) - sum(sharded.plan["strips"])),
                      "how": "ncclBroadcast of every strip straight into place; outbox blocks via a scratch buffer + cmg_tqu_scatter_block"}
        del full
        torch.cuda.empty_cache()

    # ---- e2e: host C_l in, host packed shard out, copies inside the timed region
    e2e = None
    if args.host_expand > 0 and use_orbit and world == 1:
        ctx.set_host_expand(args.host_expand)
    if not args.no_e2e:
        if kind == "tqu" and world == 1:
            del pieces, lay
            launch = None
            sharded.close()
            torch.cuda.empty_cache()
            host = torch.empty(capi.packed_size(3 * npix), dtype=torch.float64, pin_memory=True)
            spectra_pinned = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in spectra]
            step = lambda: ctx.cl_to_cmatrix_pol(*[s.numpy() for s in spectra_pinned], FWHM, host)      # reference-facing whole call
        elif kind == "tt" and world == 1:
            del shard, pieces
            torch.cuda.empty_cache()
            host = torch.empty(capi.packed_size(npix), dtype=torch.float64, pin_memory=True)
            cl_pinned = torch.from_numpy(synthetic_cl(lmax)).pin_memory()
            step = lambda: ctx.cl_to_cmatrix(cl_pinned.numpy(), FWHM, host)
        elif use_orbit:
            # orbit shards (strips + dense outbox blocks, tens of GB per rank) stream to the host through a small pinned ring, the
            # way a consumer such as the CMatrix file writer (cmg_cmatrix_file_write_device) takes them: every byte the rank
            # holds crosses PCIe inside the timed region, but host memory stays bounded for any number of ranks on the box
            ring = [torch.empty(1 << 27, dtype=torch.float64, pin_memory=True) for _ in range(2)]      # 2 x 1 GiB

            def step():
                launch()
                k = 0
                for p in pieces:
                    for off in range(0, p.numel(), ring[0].numel()):
                        n = min(ring[0].numel(), p.numel() - off)
                        ring[k & 1][:n].copy_(p[off:off + n], non_blocking=True)
                        k += 1
                torch.cuda.synchronize()
        else:
            hosts = [torch.empty(p.numel(), dtype=torch.float64, pin_memory=True) for p in pieces]

            def step():
                launch()
                for h, p in zip(hosts, pieces):
                    h.copy_(p, non_blocking=True)
                torch.cuda.synchronize()
        step()
        barrier()
        e2e_steps = min(args.steps, 10)          # each step moves the whole shard over PCIe (87 GB at N=1): keep the run bounded
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step()
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e = {"value": units_total * e2e_steps / float(wall.item()), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * float(wall.item()) / e2e_steps, "steps": e2e_steps,
               "host_expand_threads": args.host_expand if (use_orbit and world == 1) else 0,
               "note": "per rank: C_l from pinned host memory, this rank's shard of the packed matrix back to pinned host memory"
                       + (" (strips + outbox blocks, through a 2 x 1 GiB pinned ring)" if (use_orbit and world > 1) else "")}

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
            "fp64_frac_of_peak": achieved / peak_tflops, "gather": gather,
        }
        print(json.dumps(line))
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
    ap.add_argument("--gather", action="store_true", help="N>1, T,Q,U: also time the NCCL gather of the whole matrix onto every GPU")
    ap.add_argument("--no-orbit", action="store_true", help="full-sky T,Q,U: evaluate every pixel pair (cmg_tqu) instead of one per symmetry orbit")
    ap.add_argument("--host-expand", type=int, default=0, metavar="THREADS",
                    help="N=1, full sky, e2e leg: copy back only the last-face columns (27 %% of the matrix) and fill in the rotated images "
                         "on THREADS host threads (cmg_set_host_expand; opt-in until it has been timed on the GPU box)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
