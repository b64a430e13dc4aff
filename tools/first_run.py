"""Box probe + first timings (scratch tool; numbers go to gpurun_out/first_run.json)."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi, partition
from cosmopp_b200.synthetic import synthetic_cl

res = {}
res["nproc"] = os.cpu_count()
res["free"] = subprocess.run("free -g | head -2", shell=True, capture_output=True, text=True).stdout
res["smi"] = subprocess.run("nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.limit --format=csv", shell=True, capture_output=True, text=True).stdout
res["cpu"] = subprocess.run("lscpu | grep -E 'Model name|Socket|Thread|Core'", shell=True, capture_output=True, text=True).stdout
ctx = cb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
res["fp64_peak_tflops"] = [ctx.measure_fp64_peak() for _ in range(3)]

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

for nside, lmax in [(16, 47), (32, 96), (64, 192)]:
    ctx.set_pixels(nside); n = ctx.npix
    a = capi.tt_weights(synthetic_cl(lmax), capi.window_beam(lmax, 10.0))
    out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    mn, med = timeit(lambda: ctx.legendre_series(a, out))
    pairs = n * (n + 1) // 2
    res["tt_nside%d" % nside] = dict(ms_min=mn, ms_med=med, pair_l_per_s=pairs * (lmax - 1) / (mn * 1e-3), tflops_alg=pairs * (lmax - 1) * 4 / (mn * 1e-3) / 1e12)
    del out
for nside, lmax in [(16, 47), (32, 96), (64, 192)]:
    ctx.set_pixels(nside); n = ctx.npix
    sp = synthetic_cl(lmax, pol=True); f = capi.window_beam(lmax, 10.0)
    a = capi.tqu_weights(*sp, f, f)
    out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
    lay = ctx.tqu_layout_single(out)
    mn, med = timeit(lambda: ctx.tqu(*a, lay), reps=3)
    pairs = n * (n + 1) // 2
    res["tqu_nside%d" % nside] = dict(ms_min=mn, ms_med=med, pair_l_per_s=pairs * (lmax - 1) / (mn * 1e-3), tflops_alg=pairs * (lmax - 1) * 20 / (mn * 1e-3) / 1e12)
    del out
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/first_run.json", "w"), indent=1)
print(json.dumps(res, indent=1))
