"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cosmopp_b200 as cb
from cosmopp_b200 import capi, partition
from cosmopp_b200.synthetic import synthetic_cl

ctx = cb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
nside, lmax = 4, 11
good = np.array([p for p in range(192) if p % 7 != 3], dtype=np.int32)      # 165 pixels: ragged tiles
ctx.set_pixels(nside, good); n = ctx.npix
f = capi.window_beam(lmax, 10.0)
a = capi.tt_weights(synthetic_cl(lmax), f)
out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
for v in (0, 1):
    ctx.set_kernel_variant(v); ctx.legendre_series(a, out)
ctx.set_kernel_variant(0)
ctx.legendre_series_batched(np.stack([a, 2 * a, 3 * a]), torch.empty(3 * capi.packed_size(n), dtype=torch.float64, device="cuda"), capi.packed_size(n))
w = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
outp = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
for v in (22, 42, 81, 114, 122, 123, 124, 142):
    ctx.set_kernel_variant(v); ctx.tqu(*w, ctx.tqu_layout_single(outp))
ab = np.stack([np.stack(w)] * 5)
outb = torch.empty(5 * capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
for v in (0, 42, 900, 901):
    ctx.set_kernel_variant(v); ctx.tqu_batched(ab, outb, capi.packed_size(3 * n))
ctx.set_kernel_variant(0)
b = partition.column_partition(n, 3, align=32)
for r in range(3):
    plan = partition.tqu_rank_plan(n, b, r)
    strips = [torch.empty(s, dtype=torch.float64, device="cuda") for s in plan["strips"]]
    outbox = {o: [torch.empty(nc * ld, dtype=torch.float64, device="cuda") for _ in range(3)] for o, nc, ld, _ in plan["outbox"]}
    ctx.tqu(*w, capi.make_tqu_layout(b, r, [t.data_ptr() for t in strips], {k: [t.data_ptr() for t in v] for k, v in outbox.items()}))
    for o, nc, ld, row0 in plan["outbox"]:
        for t in range(3):
            ctx.tqu_scatter_block(outbox[o][t], b[o], nc, ld, row0, t, outp)
# slab path (TMA bulk copies + mbarriers), three fragment shapes, several chunks; unpack; likelihood kernels
from cosmopp_b200.likelihood import Likelihood
for lm in (9, 40, 60):
    fl = capi.window_beam(lm, 10.0)
    abl = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lm, seed=s, pol=True), fl, fl)) for s in range(37)])
    slabs = torch.empty(3 * capi.slab_doubles(3 * n), dtype=torch.float64, device="cuda")
    ctx.tqu_batched_slab(abl, slabs)
ctx.slab_unpack(slabs, 3 * n, torch.empty(16 * capi.packed_size(3 * n), dtype=torch.float64, device="cuda"), capi.packed_size(3 * n))
ctx.slab_unpack(slabs, 3 * n, outp, only_b=3)
noise = np.zeros(capi.packed_size(3 * n)); noise[[capi.packed_index(i, i) for i in range(3 * n)]] = 5.0
lk = Likelihood(ctx, slabs[2:], None, torch.from_numpy(noise).cuda(), 3 * n, foreground=np.ones(3 * n), c_stride=capi.SLAB)
lk.calculate(np.random.RandomState(1).standard_normal((4, 3 * n)))
lk.close()
ctx.tqu_batched_slab_dev(torch.from_numpy(np.ascontiguousarray(abl)).cuda(), 60, 37, slabs)
# TT variants and long series (peeled / chunk-aligned entry of the rolled loop), CMatrix file streaming
for lm in (20, 127, 128, 150):
    al = capi.tt_weights(synthetic_cl(lm), capi.window_beam(lm, 10.0))
    for v in (0, 1, 248, 2216):
        ctx.set_kernel_variant(v); ctx.legendre_series(al, out)
ctx.set_kernel_variant(0)
import tempfile
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "c.dat")
    ctx.write_cmatrix_file(path, n, [(100, out[100:], out.numel() - 100), (0, out, 100)], comment="sanitize")
    ctx.read_cmatrix_file(path, torch.empty_like(out))
d_out = torch.empty(capi.packed_size(20), dtype=torch.float64, device="cuda")
ctx.mask_matrix(out, n, np.arange(0, 160, 8), d_out)
# symmetry-orbit kernels (full sky): single owner, both modes; three ranks with outbox blocks and assembly; TT
from cosmopp_b200 import multigpu
ctx.set_pixels(16); n16 = ctx.npix
f16 = capi.window_beam(20, 10.0)
w16 = capi.tqu_weights(*synthetic_cl(20, pol=True), f16, f16)
full16 = torch.empty(capi.packed_size(3 * n16), dtype=torch.float64, device="cuda")
for mode in (0, 1, 2):
    ctx.tqu_orbit(*w16, full16, mode)
    ranks = [multigpu.OrbitShardedTQU(ctx, 16, r, 3, mode) for r in range(3)]
    for rk in ranks:
        rk.generate(w16)
    for parts in (1, 2):
        for rk in ranks:
            rk.assemble_into(full16, parts)
    # the exchange step: block(sender -> receiver) of the compact outboxes into the receivers' strips, strips to the host
    host16 = torch.empty(capi.packed_size(3 * n16), dtype=torch.float64).pin_memory()
    for d in ranks:
        for sd in ranks:
            if sd.rank != d.rank and d.recv_counts[sd.rank]:
                ctx.tqu_orbit_scatter_inbox(d.shard, sd.rank, sd.outbox.ptr + 8 * sd.layouts[sd.rank][d.rank], mode)
        d.to_host(host16, 2)
    torch.cuda.synchronize()
    for rk in ranks:
        rk.close()
ctx.legendre_series_orbit(capi.tt_weights(synthetic_cl(20), f16), torch.empty(capi.packed_size(n16), dtype=torch.float64, device="cuda"))
ctx.tqu_orbit(*w16, full16, 3)                      # meridian mirror: eight images per evaluated pair of five classes
torch.cuda.synchronize()
# packed Cholesky: groups of blocks with and without look-ahead, ragged last block; solves, log det; the step functions with
# three ranks emulated in lock step (runs of columns, shared U_kk buffer and dense panel)
def spd_packed(m, seed):
    rs = np.random.RandomState(seed)
    B = rs.normal(size=(m, m // 2 + 3))
    A = B @ B.T + 0.5 * m * np.eye(m)
    iu = np.triu_indices(m)
    p_ = np.empty(m * (m + 1) // 2)
    p_[iu[1] * (iu[1] + 1) // 2 + iu[0]] = A[iu]
    return p_
for m, group, ahead in ((300, 1, False), (700, 2, True), (1153, 4, True), (1153, 3, False)):
    ctx.set_cholesky_group(group); ctx.set_cholesky_lookahead(ahead)
    d_a = torch.from_numpy(spd_packed(m, m)).cuda()
    assert ctx.packed_cholesky(d_a, m) == 0
    ctx.packed_cholesky_logdet(d_a, m)
    ctx.packed_cholesky_solve(d_a, m, torch.ones((3, m), dtype=torch.float64, device="cuda"), 3)
ctx.set_cholesky_group(0); ctx.set_cholesky_lookahead(True)
m = 128 * 7 + 50
packed = spd_packed(m, 5)
edges = [0, 128, 384, 512, 768, 896, m]
all_runs = [[(edges[k], edges[k + 1]) for k in range(r, 6, 3)] for r in range(3)]
po = lambda c: c * (c + 1) // 2
bufs = [[torch.from_numpy(packed[po(b):po(e)].copy()).cuda() for b, e in all_runs[r]] for r in range(3)]
ukk = torch.zeros(capi.CHOL_NB * (capi.CHOL_NB + 1) // 2 + capi.CHOL_NB, dtype=torch.float64, device="cuda")
panel = torch.zeros(2 * (m + capi.CHOL_PLANE_SLACK) * capi.CHOL_NB, dtype=torch.float64, device="cuda")
class _NoComm:
    def broadcast(self, t, src): pass
    def all_reduce(self, t): pass
rk = [multigpu.ShardedCholesky(ctx, m, all_runs, r, [t.data_ptr() for t in bufs[r]], comm=_NoComm(), ukk=ukk, panel=panel, group=2) for r in range(3)]
ctx.chol_begin()
for ph in rk[0].schedule():
    for r_ in rk:
        r_.run_phase(ph)
assert ctx.chol_end() == 0
rhs = torch.ones((2, m), dtype=torch.float64, device="cuda")
for k0, kb in rk[0].blocks():
    ctx.chol_solve_diag(rk[rk[0].owners[k0 // capi.CHOL_NB]].runs, k0, kb, m, rhs, 2)
    if k0 + kb < m:
        for r_ in rk:
            ctx.chol_solve_update(r_.runs, k0, kb, m, rhs, 2)
for r_ in rk:
    ctx.chol_logdet_runs(r_.runs)
torch.cuda.synchronize()
print("sanitize_run: all kernels launched, no error reported by the runtime")
