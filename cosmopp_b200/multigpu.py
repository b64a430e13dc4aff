"""One-process-per-GPU plumbing for a sharded [T;Q;U] matrix on one box (torch.distributed for the exchange of IPC
handles, barriers and the optional NCCL gather; no data-path collective is needed to produce the shards).

Two ways to hold the entries a rank computes for another owner's columns (include/cmg.h, cmg_tqu_layout):
  "outbox"  keep them in local dense blocks (kind 1): nothing crosses NVLink, shards are block-distributed;
  "peer"    write them straight into the owner's packed strip through CUDA-IPC mapped peer memory (kind 0): the kernel's
            stores go over NVLink while the recurrences run, and afterwards every strip is complete in place.
"""
import numpy as np

from . import capi, partition


class DeviceBuffer:
    """cmg_device_malloc'ed buffer of doubles exposing __cuda_array_interface__ (zero-copy torch.as_tensor)."""

    def __init__(self, ctx, n_doubles):
        self.ctx = ctx
        self.n = int(n_doubles)
        self.ptr = ctx.device_malloc(8 * max(self.n, 1))

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n,), "typestr": "<f8", "data": (self.ptr, False), "version": 2}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda")

    def free(self):
        if self.ptr:
            self.ctx.device_free(self.ptr)
            self.ptr = 0


class ShardedTQU:
    """Rank-local storage + layout of one sharded polarized matrix."""

    def __init__(self, ctx, npix, rank, world, mode="outbox", align=32):
        import torch.distributed as dist
        self.ctx, self.npix, self.rank, self.world, self.mode = ctx, npix, rank, world, mode
        self.bounds = partition.column_partition(npix, world, align=align)
        plan = partition.tqu_rank_plan(npix, self.bounds, rank)
        self.plan = plan
        self.strips = [DeviceBuffer(ctx, s) for s in plan["strips"]]
        self.outbox = {}
        self.peer_ptrs = {}
        if mode == "outbox" or world == 1:
            for owner, ncols, ld, _ in plan["outbox"]:
                self.outbox[owner] = [DeviceBuffer(ctx, ncols * ld) for _ in range(3)]
            self.layout = capi.make_tqu_layout(self.bounds, rank, [b.ptr for b in self.strips],
                                               {k: [b.ptr for b in v] for k, v in self.outbox.items()})
        elif mode == "peer":
            mine = [ctx.ipc_export(b.ptr) for b in self.strips]
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            lay = capi.TquLayout()
            lay.n_parts, lay.own = world, rank
            for k in range(world + 1):
                lay.begin[k] = self.bounds[k]
            for k in range(world):
                if k == rank:
                    ptrs = [b.ptr for b in self.strips]
                elif k < rank and self.bounds[k + 1] > self.bounds[k]:
                    ptrs = [ctx.ipc_open(h) for h in everyone[k]]       # only owners to the left receive entries from this rank
                    self.peer_ptrs[k] = ptrs
                else:
                    continue
                for s in range(3):
                    lay.ptr[k][s] = ptrs[s]
                lay.kind[k] = 0
            self.layout = lay
        else:
            raise ValueError("mode must be 'outbox' or 'peer'")

    def pieces(self):
        """device buffers this rank holds (for copies to the host)"""
        return self.strips + [b for v in self.outbox.values() for b in v]

    def gather_full(self, full):
        """NCCL: every rank ends up with the whole packed triangle in `full` (a cuda tensor of 3N(3N+1)/2 doubles).
        Strips are contiguous pieces of the packed triangle, so each is one broadcast straight into place.  In 'outbox'
        mode the dense blocks are broadcast into a scratch buffer and placed by cmg_tqu_scatter_block (NCCL moves large
        contiguous buffers at NVLink speed; fine-grained remote stores do not)."""
        import torch
        import torch.distributed as dist
        scratch = None
        # torch's collectives run on torch's current stream, the generator and cmg_tqu_scatter_block on the context's: unless
        # the two are one stream, every phase is fenced (generation -> broadcasts -> scatter -> next broadcast into scratch)
        fence = self.ctx.stream_handle != torch.cuda.current_stream().cuda_stream

        def sync():
            if fence:
                self.ctx.synchronize()
                torch.cuda.current_stream().synchronize()
        sync()
        for k in range(self.world):
            sizes = partition.tqu_shard_sizes(self.npix, self.bounds[k], self.bounds[k + 1])
            offs = partition.tqu_strip_offsets(self.npix, self.bounds[k])
            for s in range(3):
                dst = full[offs[s]:offs[s] + sizes[s]]
                if k == self.rank:
                    dst.copy_(self.strips[s].tensor())
                if self.world > 1 and sizes[s]:
                    dist.broadcast(dst, src=k)
        if self.mode == "outbox" and self.world > 1:
            for r in range(self.world):                     # rank r's blocks for every owner to its left
                plan = partition.tqu_rank_plan(self.npix, self.bounds, r)
                for owner, ncols, ld, row0 in plan["outbox"]:
                    n = ncols * ld
                    if scratch is None or scratch.numel() < n:
                        scratch = torch.empty(n, dtype=torch.float64, device="cuda")
                    for t in range(3):
                        buf = self.outbox[owner][t].tensor() if r == self.rank else scratch[:n]
                        dist.broadcast(buf, src=r)
                        sync()
                        self.ctx.tqu_scatter_block(buf, self.bounds[owner], ncols, ld, row0, t, full)
                        sync()

    def close(self):
        for ptrs in self.peer_ptrs.values():
            for p in ptrs:
                self.ctx.ipc_close(p)
        self.peer_ptrs = {}
        for b in self.pieces():
            b.free()



class OrbitShardedTQU:
    """Rank-local storage of a full-sky [T;Q;U] matrix generated over symmetry orbits (cmg_tqu_orbit_sharded): the rank owns
    the in-face column range [q0, q1) of all twelve base faces -- 36 contiguous runs of packed columns -- plus a compact outbox,
    ordered by destination rank, for the entries whose packed column is another rank's (a third of the nine entries of a pair
    whose row pixel lies outside the range).  Producing the entries needs no collective; `exchange()` completes the strips:
    block(r -> d) of every outbox travels to rank d (one NCCL all-to-all over NVLink, or the receiver's scatter kernel reads
    the sender's outbox through CUDA IPC) and is placed into d's packed columns.  After that a rank holds exactly its columns
    of the matrix: N ranks hold 87 GB between them for the Nside = 64 matrix."""

    def __init__(self, ctx, nside, rank, world, mode=0, exchange="pull", bounds=None):
        self.ctx, self.nside, self.rank, self.world, self.mode = ctx, nside, rank, world, mode
        self.face_pix = nside * nside
        self.npix = 12 * self.face_pix
        # bounds: the in-face ranges of the ranks when not the cost-balanced ones -- partition.orbit_partition_blocks when the
        # strips go on to ShardedCholesky (its run boundaries lie on the 128-column block grid)
        self.bounds = list(bounds) if bounds is not None else partition.orbit_partition(nside, world, mode)
        if len(self.bounds) != world + 1 or self.bounds[0] != 0 or self.bounds[-1] != self.face_pix or any(
                b % 32 or b < a for a, b in zip(self.bounds, self.bounds[1:])):
            raise ValueError("bounds: world + 1 ascending multiples of 32 from 0 to nside^2")
        self.q0, self.q1 = self.bounds[rank], self.bounds[rank + 1]
        self.layouts = [capi.orbit_outbox_layout(nside, mode, self.bounds, r) for r in range(world)] if world > 1 else [[0, 0]]
        n_strips, n_outbox = self.sizes_of(rank)
        # one allocation for the 36 strips: the pieces stay individually contiguous pieces of the packed triangle
        self.strips = DeviceBuffer(ctx, n_strips)
        self.outbox = DeviceBuffer(ctx, n_outbox) if n_outbox else None
        self.shard = self.shard_of(rank, self.strips.ptr, self.outbox.ptr if self.outbox is not None else 0)
        self.pairs = partition.orbit_pairs_in_range(self.q0, self.q1, self.face_pix, mode)
        # "pull": the receiver's kernels read the sender's buffers through CUDA-IPC mapped peer memory (measured on 2 B200s:
        # exchange 13.7 ms against 19.5 ms for the NCCL all-to-all + local scatter); "nccl": NCCL collectives only.  All ranks
        # fall back to "nccl" together when a peer mapping cannot be made.
        self.exchange_mode = exchange
        self.inbox = None
        self.peer_outbox = {}
        self.peer_strips = {}
        self._mapped = False
        # doubles this rank receives from every sender / sends to every destination
        self.recv_counts = [self.block_size(r, rank) for r in range(world)]
        self.send_counts = [self.block_size(rank, d) for d in range(world)]

    def block_size(self, sender, dest):
        if sender == dest or self.world == 1:
            return 0
        lay = self.layouts[sender]
        return lay[dest + 1] - lay[dest]

    def sizes_of(self, rank):
        """doubles in the strips buffer and in the outbox buffer of rank `rank`"""
        q0, q1 = self.bounds[rank], self.bounds[rank + 1]
        n_strips = sum(sum(r) for r in partition.orbit_strip_sizes(self.nside, q0, q1))
        return n_strips, (self.layouts[rank][-1] if self.world > 1 else 0)

    def shard_of(self, rank, strips_ptr, outbox_ptr):
        """cmg_orbit_shard of rank `rank` over a strips buffer and an outbox buffer laid out like this class's own"""
        q0, q1 = self.bounds[rank], self.bounds[rank + 1]
        sizes = partition.orbit_strip_sizes(self.nside, q0, q1)
        shard = capi.OrbitShard()
        shard.n_ranks, shard.rank = self.world, rank
        for k, b in enumerate(self.bounds):
            shard.bounds[k] = b
        off = 0
        for s in range(3):
            for f in range(12):
                shard.strip[s][f] = strips_ptr + 8 * off
                off += sizes[s][f]
        shard.outbox = outbox_ptr or None
        return shard

    def generate(self, weights):
        self.ctx.tqu_orbit_sharded(*weights, self.shard, self.mode)

    def chol_runs(self):
        """(all_runs, run_ptrs) for ShardedCholesky: the 36 runs of packed columns of every rank, and where this rank's lie"""
        all_runs = [partition.orbit_column_runs(self.nside, self.bounds[r], self.bounds[r + 1]) for r in range(self.world)]
        sizes = partition.orbit_strip_sizes(self.nside, self.q0, self.q1)
        ptrs, off = [], 0
        for s in range(3):
            for f in range(12):
                ptrs.append(self.strips.ptr + 8 * off)
                off += sizes[s][f]
        return all_runs, ptrs

    def exchange(self):
        """complete this rank's strips with the entries the other ranks computed for its columns.  Collective."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        if self.exchange_mode == "pull" and self._map_peers():
            # the receiver's scatter kernel reads block(r -> this rank) straight out of rank r's outbox over NVLink
            self._sync_streams()
            dist.barrier()                         # every outbox is complete before anyone reads it
            # every rank starts with a different peer (rank + 1, rank + 2, ...): with one common order all ranks would read the same
            # sender at the same time (measured on 8 GPUs: 17.1 ms that way)
            for k in range(1, self.world):
                r = (self.rank + k) % self.world
                if r in self.peer_outbox and self.recv_counts[r]:
                    self.ctx.tqu_orbit_scatter_inbox(self.shard, r, self.peer_outbox[r] + 8 * self.layouts[r][self.rank], self.mode)
            self._sync_streams()
            dist.barrier()                         # nobody overwrites an outbox a peer is still reading
            return
        if self.inbox is None:
            self.inbox = torch.empty(max(sum(self.recv_counts), 1), dtype=torch.float64, device="cuda")
        self._sync_streams()
        out = self.outbox.tensor()
        dist.all_to_all_single(self.inbox[:sum(self.recv_counts)], out, self.recv_counts, self.send_counts)
        self._sync_streams()
        off = 0
        for r in range(self.world):
            if self.recv_counts[r]:
                self.ctx.tqu_orbit_scatter_inbox(self.shard, r, self.inbox[off:], self.mode)
                off += self.recv_counts[r]

    def _map_peers(self):
        """CUDA-IPC mappings of every peer's outbox and strips buffers (once); False -> every rank uses NCCL instead"""
        if self._mapped:
            return self.exchange_mode == "pull"
        import torch
        import torch.distributed as dist
        self._mapped = True
        ok = 1
        try:
            mine = (self.ctx.ipc_export(self.outbox.ptr) if self.outbox is not None else None, self.ctx.ipc_export(self.strips.ptr))
        except Exception:
            mine, ok = (None, None), 0
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        for r in range(self.world):
            if r == self.rank or not ok:
                continue
            try:
                if everyone[r][0] is not None:
                    self.peer_outbox[r] = self.ctx.ipc_open(everyone[r][0])
                self.peer_strips[r] = self.ctx.ipc_open(everyone[r][1])
            except Exception:
                ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            self.exchange_mode = "nccl"
        return self.exchange_mode == "pull"

    def _sync_streams(self):
        """torch's collectives run on torch's current stream, the generator on the context's: order the two unless they are one"""
        import torch
        if self.ctx.stream_handle != torch.cuda.current_stream().cuda_stream:
            self.ctx.synchronize()
            torch.cuda.current_stream().synchronize()

    def pieces(self):
        return [self.strips] + ([self.outbox] if self.outbox is not None else [])

    def to_host(self, host_packed, threads=0, direct_mask=0):
        """this rank's (complete) strips as its columns of the whole packed HOST matrix (a shared mapping for several ranks);
        direct_mask: images that come over PCIe as well instead of from the host threads (cmg_orbit_strips_to_host)"""
        self.ctx.orbit_strips_to_host(self.shard, host_packed, threads, direct_mask)

    def assemble_into(self, full, parts=3):
        """place this rank's pieces into a whole packed triangle on this GPU (parts: 1 strips, 2 outbox; the strips of all
        ranks go in before any outbox, a strip has holes where another rank's outbox holds the entry)"""
        self.ctx.tqu_orbit_assemble(self.shard, full, self.mode, parts)

    def gather_full(self, full):
        """Every rank ends up with the whole packed triangle in `full` (only when a consumer needs it).  Call after exchange():
        the strips are then complete contiguous pieces of the packed triangle.  "pull": each rank copies every peer's 36 runs
        out of the peer's strips buffer (CUDA IPC) straight into place, all ranks at once over NVSwitch; "nccl": one broadcast
        per run."""
        import torch.distributed as dist
        self._sync_streams()
        n = self.npix
        if self.world > 1 and self.exchange_mode == "pull" and self._map_peers():
            dist.barrier()                         # the peers' strips are complete
            self.assemble_into(full, 1)
            for k in range(1, self.world):           # a different first peer on every rank (see exchange)
                r = (self.rank + k) % self.world
                base = self.peer_strips[r]
                q0, q1 = self.bounds[r], self.bounds[r + 1]
                sizes = partition.orbit_strip_sizes(self.nside, q0, q1)
                off = 0
                for s in range(3):
                    for f in range(12):
                        if sizes[s][f]:
                            first = partition.packed_size(s * n + f * self.face_pix + q0)
                            self.ctx.copy_on_device(full[first:first + sizes[s][f]], base + 8 * off, 8 * sizes[s][f])
                        off += sizes[s][f]
            self._sync_streams()
            dist.barrier()                         # nobody regenerates while a peer still reads its strips
            return
        for r in range(self.world):
            q0, q1 = self.bounds[r], self.bounds[r + 1]
            sizes = partition.orbit_strip_sizes(self.nside, q0, q1)
            if r == self.rank:
                self.assemble_into(full, 1)
                self._sync_streams()
            for s in range(3):
                for f in range(12):
                    first = partition.packed_size(s * n + f * self.face_pix + q0)
                    if self.world > 1 and sizes[s][f]:
                        dist.broadcast(full[first:first + sizes[s][f]], src=r)

    def close(self):
        for p in list(self.peer_outbox.values()) + list(self.peer_strips.values()):
            self.ctx.ipc_close(p)
        self.peer_outbox = {}
        self.peer_strips = {}
        self.inbox = None
        for b in self.pieces():
            b.free()


class TorchComm:
    """the two exchanges of a sharded Cholesky step over torch.distributed (NCCL on GPUs; gloo in the CPU tests of the plan)"""

    def broadcast(self, tensor, src):
        import torch.distributed as dist
        dist.broadcast(tensor, src=src)

    def all_reduce(self, tensor):
        import torch.distributed as dist
        dist.all_reduce(tensor)


def chol_block_owners(n, all_runs):
    """owner rank of every 128-row block of an n-dimensional matrix whose columns are spread as `all_runs`:
    all_runs[rank] = [(col_begin, col_end), ...].  Every boundary must be a multiple of 128 (the end of the matrix excepted)
    and the runs of all ranks together must tile [0, n) exactly."""
    nb = capi.CHOL_NB
    owners = [-1] * ((n + nb - 1) // nb)
    covered = 0
    for rank, runs in enumerate(all_runs):
        for b, e in runs:
            if e <= b:
                continue
            if b % nb or (e % nb and e != n) or b < 0 or e > n:
                raise ValueError("run [%d, %d) of rank %d: boundaries must be multiples of %d inside [0, %d]" % (b, e, rank, nb, n))
            for k in range(b // nb, (e + nb - 1) // nb):
                if owners[k] != -1:
                    raise ValueError("block %d is in the runs of ranks %d and %d" % (k, owners[k], rank))
                owners[k] = rank
            covered += e - b
    if covered != n or -1 in owners:
        raise ValueError("the runs do not tile the %d columns of the matrix" % n)
    return owners


class ShardedCholesky:
    """A = U^T U of a packed symmetric positive definite matrix whose COLUMNS are spread over the GPUs of one box, in place
    (include/cmg.h, cmg_chol_*; kernels in csrc/cholesky.cuh).  A rank holds runs of whole packed columns -- after
    OrbitShardedTQU.exchange() its 36 strips are exactly that, so the Nside = 64 matrix (147456-dimensional, 87 GB) is
    factorised where the generator left it: 11 GB per GPU on eight of them, nothing gathered, nothing redistributed.  The runs
    of one rank interleave with everyone else's 36 times over the matrix, so the shrinking trailing matrix stays balanced
    (a block-cyclic distribution with 36 cycles).

    Right-looking, blocks of 128 rows taken in groups of `group` (the reference runs LAPACK dpptrf on the host,
    source/matrix_impl.cpp:236-263).  Per block:
      every rank: the block's 128 rows of its own columns catch up with the group's earlier blocks (cmg_chol_syrk, strip)
      owner of the block: U_kk (cmg_chol_diag)                      -> broadcast of 66 KB
      every rank: the 128 rows of its own columns (cmg_chol_panel)  -> all-reduce of that plane of the dense panel (<= 151 MB,
                  entries of other ranks' columns are zero: the sum is a gather that needs no layout agreement between ranks)
    and per group
      every rank: trailing update of its own columns with all 128 * group rows (cmg_chol_syrk, FP64 tensor-core tiles,
                  operands from the dense panel)
    Over the whole factorisation every rank receives the factor once (87 GB over NVLink against n^3 / 3 / G of arithmetic).

    `comm`: broadcast(tensor, src) / all_reduce(tensor) (TorchComm).  The context must run on torch's current stream
    (ctx.set_stream(torch.cuda.current_stream().cuda_stream)): kernels and collectives alternate 1152 times, fences would
    serialise the host with the device every time.  `ukk` / `panel`: torch buffers to use (ranks emulated on ONE GPU share
    them -- tests/test_gpu_cholesky.py); allocated when None."""

    def __init__(self, ctx, n, all_runs, rank, run_ptrs, comm=None, ukk=None, panel=None, group=4):
        import torch
        self.ctx, self.n, self.rank, self.world = ctx, int(n), rank, len(all_runs)
        self.owners = chol_block_owners(self.n, all_runs)
        mine = [(b, e) for b, e in all_runs[rank]]
        if len(run_ptrs) != len(mine):
            raise ValueError("one device pointer per run of this rank")
        if not 1 <= group <= capi.CHOL_MAX_GROUP:
            raise ValueError("group: 1 .. %d blocks" % capi.CHOL_MAX_GROUP)
        live = [(b, e, p) for (b, e), p in zip(mine, run_ptrs) if e > b]
        self.runs = capi.make_chol_runs(live) if live else None
        self.comm = comm if comm is not None else TorchComm()
        nb = capi.CHOL_NB
        self.group = group
        self.plane_stride = (self.n + capi.CHOL_PLANE_SLACK) * nb
        self.ukk = ukk if ukk is not None else torch.empty(nb * (nb + 1) // 2 + nb, dtype=torch.float64, device="cuda")
        # zeros: the spare rows behind the last column are read (never used)
        # two sets of planes: the look-ahead writes the next group's rows while the update still reads this group's
        self.panel = panel if panel is not None else torch.zeros(2 * group * self.plane_stride, dtype=torch.float64, device="cuda")
        if self.panel.numel() < group * self.plane_stride:
            raise ValueError("panel: group * (n + %d) * %d doubles" % (capi.CHOL_PLANE_SLACK, nb))
        self.device = self.panel.device
        self.info = None

    def _check_stream(self):
        import torch
        if self.world > 1 and self.device.type == "cuda" and self.ctx.stream_handle != torch.cuda.current_stream().cuda_stream:
            raise RuntimeError("ShardedCholesky: the context must share torch's current stream (ctx.set_stream(torch.cuda.current_stream().cuda_stream))")

    # ---- the phases (a test drives emulated ranks through them in lock step)
    def blocks(self):
        nb = capi.CHOL_NB
        return [(k0, min(nb, self.n - k0)) for k0 in range(0, self.n, nb)]

    def schedule(self):
        """the order of phases, the same on every rank: ("strip", group start, sub) | ("diag", k0, kb, owner) |
        ("panel", k0, group start, sub) | ("syrk", group start, blocks in the group)"""
        nb, out = capi.CHOL_NB, []
        for base in range(0, self.n, nb * self.group):
            subs = 0
            for sub in range(self.group):
                k0 = base + sub * nb
                if k0 >= self.n:
                    break
                kb = min(nb, self.n - k0)
                if sub:
                    out.append(("strip", base, sub))
                out.append(("diag", k0, kb, self.owners[k0 // nb]))
                if k0 + kb >= self.n:
                    return out
                out.append(("panel", k0, base, sub))
                subs = sub + 1
            if base + subs * nb < self.n and subs == self.group:
                out.append(("syrk", base, subs))
        return out

    def plane_region(self, k0, base, sub, buf=0):
        """the part of plane `sub` (of plane set `buf`) the panel step of block k0 fills: the 128 rows of every column behind the block"""
        nb = capi.CHOL_NB
        first = (buf * self.group + sub) * self.plane_stride + (k0 + nb - base) * nb
        return self.panel[first:first + (self.n - k0 - nb) * nb]

    def run_phase(self, ph, buf=0):
        """this rank's kernel of a phase (no exchange); buf: which of the two sets of planes (look-ahead alternates)"""
        nb = capi.CHOL_NB
        if self.runs is None:
            return
        planes = self.panel[buf * self.group * self.plane_stride:]
        if ph[0] == "strip":
            self.ctx.chol_syrk(self.runs, ph[1], ph[2] * nb, planes, self.plane_stride, ph[1], True)
        elif ph[0] == "diag":
            if ph[3] == self.rank:
                self.ctx.chol_diag(self.runs, ph[1], ph[2], self.ukk)
        elif ph[0] == "panel":
            self.ctx.chol_panel(self.runs, ph[1], nb, self.ukk, planes[ph[3] * self.plane_stride:], ph[2])
        else:
            self.ctx.chol_syrk(self.runs, ph[1], ph[2] * nb, planes, self.plane_stride, ph[1], False)

    def _run_with_exchanges(self, ph, buf=0):
        """a phase with the exchange that belongs to it (on torch's current stream)"""
        if ph[0] == "panel" and self.world > 1:
            region = self.plane_region(ph[1], ph[2], ph[3], buf)
            region.zero_()
            self.run_phase(ph, buf)
            self.comm.all_reduce(region)
            return
        self.run_phase(ph, buf)
        if ph[0] == "diag" and self.world > 1 and ph[1] + ph[2] < self.n:
            self.comm.broadcast(self.ukk, ph[3])

    def factorise(self, lookahead=None):
        """Collective.  Returns LAPACK's info (0: positive definite; k: the leading minor of order k is not), the same on every rank.
        lookahead (default: on a GPU when the panel buffer holds two sets of planes): the trailing update of a group is issued
        in pieces -- first, strip by strip, the rows the next group will factorise, then the rest -- and the next group's blocks
        (diagonal block, broadcast, panel, all-reduce, ...) run on a second, high-priority stream beside the rest."""
        import torch
        self._check_stream()
        two_sets = self.panel.numel() >= 2 * self.group * self.plane_stride
        if lookahead is None:
            lookahead = self.device.type == "cuda" and two_sets and self.n > 2 * self.group * capi.CHOL_NB
        if lookahead and not two_sets:
            raise ValueError("look-ahead needs a panel buffer of 2 * group * (n + %d) * %d doubles" % (capi.CHOL_PLANE_SLACK, capi.CHOL_NB))
        self.ctx.chol_begin()
        if lookahead:
            self._factorise_lookahead()
        else:
            for ph in self.schedule():
                self._run_with_exchanges(ph)
        info = self.ctx.chol_end()
        if self.world > 1:                       # the first failing block wins on every rank
            big = 1 << 62
            t = torch.tensor([info if info else big], dtype=torch.int64, device=self.device)
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            info = int(t.item())
            info = 0 if info == big else info
        self.info = info
        return info

    def _factorise_lookahead(self):
        import torch
        nb = capi.CHOL_NB
        rows = nb * self.group
        # the phases of the schedule, group by group (without the trailing updates: those are issued in pieces here)
        groups, cur = [], []
        for ph in self.schedule():
            if ph[0] == "syrk":
                groups.append(cur)
                cur = []
            else:
                cur.append(ph)
        groups.append(cur)
        # (host buffers -- the tests of the order of calls -- have no streams: the same order, one call after the other)
        streams = self.device.type == "cuda"
        if streams:
            main = torch.cuda.current_stream()
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream(priority=-1)
            side = self._side
            ev_a, ev_f = torch.cuda.Event(), torch.cuda.Event()
        for ph in groups[0]:
            self._run_with_exchanges(ph, 0)
        for g in range(len(groups) - 1):
            base, planes = g * rows, self.panel[(g & 1) * self.group * self.plane_stride:]
            if self.runs is not None:
                for r in range(self.group):          # the rows of the next group, strip by strip
                    self.ctx.chol_syrk(self.runs, base, rows, planes, self.plane_stride, base, True, shift=r * nb)
            if streams:
                ev_a.record(main)
                side.wait_event(ev_a)
                self.ctx.set_stream(side.cuda_stream)
                try:
                    with torch.cuda.stream(side):
                        for ph in groups[g + 1]:
                            self._run_with_exchanges(ph, (g + 1) & 1)
                        ev_f.record(side)
                finally:
                    self.ctx.set_stream(main.cuda_stream)
            else:
                for ph in groups[g + 1]:
                    self._run_with_exchanges(ph, (g + 1) & 1)
            if self.runs is not None:                # the rest of the update, beside the next group's blocks
                self.ctx.chol_syrk(self.runs, base, rows, planes, self.plane_stride, base, False, shift=rows)
            if streams:
                main.wait_event(ev_f)

    def logdet(self):
        """log det A = 2 sum log U_jj.  Collective."""
        import torch
        share = self.ctx.chol_logdet_runs(self.runs) if self.runs is not None else 0.0
        if self.world == 1:
            return share
        t = torch.tensor([share], dtype=torch.float64, device=self.device)
        self.comm.all_reduce(t)
        return float(t.item())

    def solve(self, rhs):
        """y = U^-T t in place: rhs is a (n_rhs, n) float64 cuda tensor (n x n_rhs column-major), the same on every rank going in
        and coming out; chi^2 of right-hand side m = |y_m|^2.  Collective."""
        self._check_stream()
        nb = capi.CHOL_NB
        if rhs.dim() != 2 or rhs.shape[1] != self.n or not rhs.is_contiguous():
            raise ValueError("rhs must be a contiguous (n_rhs, n) tensor")
        n_rhs = rhs.shape[0]
        for k0, kb in self.blocks():
            owner = self.owners[k0 // nb]
            if owner == self.rank:
                self.ctx.chol_solve_diag(self.runs, k0, kb, self.n, rhs, n_rhs)
            if self.world > 1:
                view = rhs[:, k0:k0 + kb]
                if view.is_contiguous():
                    self.comm.broadcast(view, owner)
                else:
                    tmp = view.contiguous()
                    self.comm.broadcast(tmp, owner)
                    view.copy_(tmp)
            if k0 + kb < self.n and self.runs is not None:
                self.ctx.chol_solve_update(self.runs, k0, kb, self.n, rhs, n_rhs)
        return rhs


class OrbitShardedTT:
    """Rank-local storage of a full-sky TT matrix generated over symmetry orbits (cmg_legendre_series_orbit_sharded): the rank
    owns the in-face column range [q0, q1) of all twelve base faces -- 12 contiguous runs of packed columns.  The sharded form
    uses the plan without transposed images, so every entry lands in a column of the rank that evaluated it: the strips are
    complete when the kernel ends, nothing is exchanged."""

    def __init__(self, ctx, nside, rank, world):
        self.ctx, self.nside, self.rank, self.world = ctx, nside, rank, world
        self.face_pix = nside * nside
        self.npix = 12 * self.face_pix
        self.mode = 0 if (world == 1 and nside >= 32) else 1      # what cmg_legendre_series_orbit_sharded chooses
        self.bounds = partition.orbit_partition(nside, world, self.mode, align=16)
        self.q0, self.q1 = self.bounds[rank], self.bounds[rank + 1]
        self.sizes = [partition.packed_size(f * self.face_pix + self.q1) - partition.packed_size(f * self.face_pix + self.q0) for f in range(12)]
        self.strips = DeviceBuffer(ctx, sum(self.sizes))
        self.ptrs, off = [], 0
        for f in range(12):
            self.ptrs.append(self.strips.ptr + 8 * off)
            off += self.sizes[f]
        self.pairs = partition.orbit_pairs_in_range(self.q0, self.q1, self.face_pix, self.mode)

    def generate(self, weights):
        self.ctx.legendre_series_orbit_sharded(weights, self.q0, self.q1, self.ptrs)

    def pieces(self):
        return [self.strips]

    def place_into(self, full):
        """this rank's 12 runs into a whole packed triangle on this GPU"""
        off = 0
        for f in range(12):
            if self.sizes[f]:
                first = partition.packed_size(f * self.face_pix + self.q0)
                self.ctx.copy_on_device(full[first:first + self.sizes[f]], self.strips.ptr + 8 * off, 8 * self.sizes[f])
            off += self.sizes[f]

    def close(self):
        self.strips.free()
