"""Multi-GPU check (run under torchrun, >= 2 GPUs): the peer-written sharded matrix, gathered with NCCL, equals the
one-piece matrix bit for bit; the outbox layout holds the same values.  Usage:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py [NSIDE LMAX]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cosmopp_b200 as cb
from cosmopp_b200 import capi, multigpu
from cosmopp_b200.synthetic import synthetic_cl

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 47
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cb.Context(local); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_pixels(nside)
n = ctx.npix
f = capi.window_beam(lmax, 10.0)
a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
whole = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
ctx.tqu(*a, ctx.tqu_layout_single(whole)); torch.cuda.synchronize()

sh = multigpu.ShardedTQU(ctx, n, rank, world, mode="peer")
for b in sh.strips:
    b.tensor().fill_(float("nan"))
torch.cuda.synchronize(); dist.barrier()
ctx.tqu(*a, sh.layout)
torch.cuda.synchronize(); dist.barrier()
full = torch.full_like(whole, float("nan"))
sh.gather_full(full)
torch.cuda.synchronize()
ok_peer = bool(torch.equal(full, whole))
print("rank %d: peer-written + NCCL-gathered matrix identical to the one-piece matrix: %s (nan left: %d)" % (rank, ok_peer, int(torch.isnan(full).sum())))
sh.close()

so = multigpu.ShardedTQU(ctx, n, rank, world, mode="outbox")
for b in so.pieces():
    b.tensor().fill_(float("nan"))
torch.cuda.synchronize(); dist.barrier()
ctx.tqu(*a, so.layout); torch.cuda.synchronize(); dist.barrier()
full2 = torch.full_like(whole, float("nan"))
so.gather_full(full2)
torch.cuda.synchronize()
ok_gather = bool(torch.equal(full2, whole))
print("rank %d: outbox shards gathered with NCCL + scatter identical to the one-piece matrix: %s" % (rank, ok_gather))
from cosmopp_b200 import partition
offs = partition.tqu_strip_offsets(n, so.bounds[rank])
ok_out = True
for s in range(3):
    got = so.strips[s].tensor()
    ref = whole[offs[s]:offs[s] + got.numel()]
    written = ~torch.isnan(got)
    ok_out &= bool(torch.equal(got[written], ref[written]))
print("rank %d: outbox-mode strips agree where written: %s" % (rank, ok_out))
so.close()
flag = torch.tensor([int(ok_peer and ok_out and ok_gather)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
