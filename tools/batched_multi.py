"""BASELINE config 4: one MCMC step = 1024 synthetic C_l sets -> 1024 polarized Nside=16, lmax=47 matrices (348 GB of
output: consumed / overwritten in sub-batches of 256 per GPU).  Batch axis sharded over the ranks (no collective in the
data path); run alone or under torchrun:
    python tools/batched_multi.py [B]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/batched_multi.py [B]
Prints one JSON line from rank 0 (device-resident timing, CUDA events, max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import cosmopp_b200 as cb
from cosmopp_b200 import capi, partition
from cosmopp_b200.synthetic import synthetic_cl

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
SUB = 256
nside, lmax = 16, 47
ctx = cb.Context(local); stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream); ctx.set_pixels(nside)
n = ctx.npix; pairs = n * (n + 1) // 2
f = capi.window_beam(lmax, 10.0)
bounds = partition.batch_partition(B, world)
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

for _ in range(3):
    step()
barrier()
steps = 5
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    step()
e1.record(stream)
barrier()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
if rank == 0:
    units = B * pairs * (lmax - 1)
    per_gpu = units / world
    print(json.dumps({
        "workload": "batched x%d tqu_nside16_lmax47 (BASELINE configs[3]), slab output, FP64 tensor path" % B, "n_gpus": world,
        "ms_per_step": ms, "ms_per_matrix": ms / B, "pixel_pair_ell_elements_per_s": units / (ms * 1e-3), "scaling": "strong (batch axis)",
        "fp64": {"algorithmic_flop_per_unit": 8.0, "achieved_tflops_per_gpu": 8.0 * per_gpu / (ms * 1e-3) / 1e12, "peak_tflops": peak,
                 "frac": 8.0 * per_gpu / (ms * 1e-3) / 1e12 / peak},
        "hbm_write": {"bytes_per_step": B * capi.packed_size(3 * n) * 8, "gbs_per_gpu": B * capi.packed_size(3 * n) * 8 / world / (ms * 1e-3) / 1e9,
                      "fill_peak_gbs": 7500.0},
        "sub_batch": SUB, "launches_per_step_per_gpu": 2 * ((b1 - b0 + SUB - 1) // SUB)}))
if world > 1:
    dist.destroy_process_group()
