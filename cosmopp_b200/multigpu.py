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
                        self.ctx.tqu_scatter_block(buf, self.bounds[owner], ncols, ld, row0, t, full)

    def close(self):
        for ptrs in self.peer_ptrs.values():
            for p in ptrs:
                self.ctx.ipc_close(p)
        self.peer_ptrs = {}
        for b in self.pieces():
            b.free()



class OrbitShardedTQU:
    """Rank-local storage of a full-sky [T;Q;U] matrix generated over symmetry orbits (cmg_tqu_orbit_sharded): the rank owns
    the in-face column range [q0, q1) of all twelve base faces -- 36 contiguous runs of packed columns -- plus dense outbox
    blocks for the entries whose packed column is another rank's.  No data-path collective."""

    def __init__(self, ctx, nside, rank, world, mode=0):
        self.ctx, self.nside, self.rank, self.world, self.mode = ctx, nside, rank, world, mode
        self.face_pix = nside * nside
        self.npix = 12 * self.face_pix
        self.bounds = partition.orbit_partition(nside, world, mode)
        self.q0, self.q1 = self.bounds[rank], self.bounds[rank + 1]
        # (kind, column face, first row face, last row face + 1) of every outbox block the plan writes
        self.outbox_blocks = partition.orbit_outbox_blocks(capi.orbit_plan(nside, mode)) if world > 1 else []
        n_strips, n_outbox = self.sizes_of(rank)
        # one allocation for the 36 strips: the pieces stay individually contiguous pieces of the packed triangle
        self.strips = DeviceBuffer(ctx, n_strips)
        self.outbox = DeviceBuffer(ctx, n_outbox) if n_outbox else None
        self.shard = self.shard_of(rank, self.strips.ptr, self.outbox.ptr if self.outbox is not None else 0)
        self.pairs = partition.orbit_pairs_in_range(self.q0, self.q1, self.face_pix, mode)

    def sizes_of(self, rank):
        """doubles in the strips buffer and in the outbox buffer of rank `rank`"""
        q0, q1 = self.bounds[rank], self.bounds[rank + 1]
        n_strips = sum(sum(r) for r in partition.orbit_strip_sizes(self.nside, q0, q1))
        n_outbox = sum(hi - lo for _, _, lo, hi in self.outbox_blocks) * self.face_pix * (q1 - q0) if q1 > q0 else 0
        return n_strips, n_outbox

    def shard_of(self, rank, strips_ptr, outbox_ptr):
        """cmg_orbit_shard of rank `rank` over a strips buffer and an outbox buffer laid out like this class's own"""
        q0, q1 = self.bounds[rank], self.bounds[rank + 1]
        sizes = partition.orbit_strip_sizes(self.nside, q0, q1)
        shard = capi.OrbitShard()
        shard.q_begin, shard.q_end = q0, q1
        off = 0
        for s in range(3):
            for f in range(12):
                shard.strip[s][f] = strips_ptr + 8 * off
                off += sizes[s][f]
        if outbox_ptr:
            # only the row-pixel faces a block is ever addressed with are allocated; the pointer handed to the kernel is that of
            # (virtual) row pixel 0, i.e. moved back by first_face x nside^2 rows of ld elements
            per_face = self.face_pix * (q1 - q0)
            off = 0
            for t, f, lo, hi in self.outbox_blocks:
                shard.outbox[t][f] = outbox_ptr + 8 * (off - lo * per_face)
                off += (hi - lo) * per_face
        return shard

    def generate(self, weights):
        self.ctx.tqu_orbit_sharded(*weights, self.shard, self.mode)

    def pieces(self):
        return [self.strips] + ([self.outbox] if self.outbox is not None else [])

    def assemble_into(self, full, parts=3):
        """place this rank's pieces into a whole packed triangle on this GPU (parts: 1 strips, 2 outbox; the strips of all
        ranks go in before any outbox, a strip has holes where another rank's outbox holds the entry)"""
        self.ctx.tqu_orbit_assemble(self.shard, full, self.mode, parts)

    def gather_full(self, full):
        """NCCL: every rank ends up with the whole packed triangle in `full` (only when a consumer needs it).  A strip is a
        contiguous piece of the packed triangle, so each one is broadcast straight into place; a rank's outbox buffer is
        broadcast whole into a scratch buffer and placed by orbitOutboxScatterKernel.  (The outbox is allocated dense, so
        this moves about twice the bytes of the matrix; a tile-major outbox would halve it.)"""
        import torch
        import torch.distributed as dist
        n = self.npix
        for r in range(self.world):                       # strips of all ranks first
            q0, q1 = self.bounds[r], self.bounds[r + 1]
            sizes = partition.orbit_strip_sizes(self.nside, q0, q1)
            if r == self.rank:
                self.assemble_into(full, 1)
            for s in range(3):
                for f in range(12):
                    first = partition.packed_size(s * n + f * self.face_pix + q0)
                    if self.world > 1 and sizes[s][f]:
                        dist.broadcast(full[first:first + sizes[s][f]], src=r)
        if self.world == 1:
            return
        scratch = torch.empty(max(self.sizes_of(r)[1] for r in range(self.world)), dtype=torch.float64, device="cuda")
        for r in range(self.world):
            n_outbox = self.sizes_of(r)[1]
            if not n_outbox:
                continue
            buf = self.outbox.tensor() if r == self.rank else scratch[:n_outbox]
            dist.broadcast(buf, src=r)
            # only the outbox pointers of this descriptor are read (parts = 2)
            shard = self.shard_of(r, self.strips.ptr, buf.data_ptr())
            self.ctx.tqu_orbit_assemble(shard, full, self.mode, 2)

    def close(self):
        for b in self.pieces():
            b.free()
