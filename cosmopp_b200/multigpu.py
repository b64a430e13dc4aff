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

    def __init__(self, ctx, nside, rank, world, mode=0, exchange="pull"):
        self.ctx, self.nside, self.rank, self.world, self.mode = ctx, nside, rank, world, mode
        self.face_pix = nside * nside
        self.npix = 12 * self.face_pix
        self.bounds = partition.orbit_partition(nside, world, mode)
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
