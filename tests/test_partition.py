"""Host-side sharding logic (no GPU): equal-area column blocks, shard sizes, and a world_size-2 gloo run of the
per-rank planning the multi-GPU bench does."""
import os
import socket

import numpy as np
import pytest

from cosmopp_b200 import partition


@pytest.mark.parametrize("npix,parts", [(3072, 1), (3072, 2), (12288, 4), (49152, 8), (2548, 3), (100, 8)])
def test_column_partition_tiles_and_balances(npix, parts):
    b = partition.column_partition(npix, parts, align=32)
    assert b[0] == 0 and b[-1] == npix and len(b) == parts + 1 and all(x <= y for x, y in zip(b, b[1:]))
    assert all(x % 32 == 0 for x in b[1:-1])
    pairs = [partition.pairs_in_block(b[k], b[k + 1]) for k in range(parts)]
    assert sum(pairs) == npix * (npix + 1) // 2
    if npix >= 3072:
        assert max(pairs) <= 1.05 * sum(pairs) / parts        # balanced to the column-tile granularity


def test_shards_tile_the_packed_triangle():
    npix, parts = 1000, 4
    b = partition.column_partition(npix, parts, align=32)
    assert sum(partition.tt_shard_size(b[k], b[k + 1]) for k in range(parts)) == partition.packed_size(npix)
    strips = sum(sum(partition.tqu_shard_sizes(npix, b[k], b[k + 1])) for k in range(parts))
    assert strips == partition.packed_size(3 * npix)
    # every entry is written exactly once: strips hold the natural entries, outboxes the transposed ones of cross pairs
    held = sum(partition.tqu_entries_held(npix, b, r) for r in range(parts))
    assert held == partition.packed_size(3 * npix)
    for r in range(parts):
        plan = partition.tqu_rank_plan(npix, b, r)
        assert [o[0] for o in plan["outbox"]] == list(range(r))
        assert all(ld == b[r + 1] - b[r] and row0 == b[r] for _, _, ld, row0 in plan["outbox"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, npix, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bounds = partition.column_partition(npix, world, align=32)
    plan = partition.tqu_rank_plan(npix, bounds, rank)
    mine = torch.tensor([plan["pairs"], partition.tqu_entries_held(npix, bounds, rank), sum(plan["strips"]),
                         sum(3 * n * ld for _, n, ld, _ in plan["outbox"])], dtype=torch.int64)
    allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    # the timing rule of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((torch.stack(allr).tolist(), float(t)))
    dist.destroy_process_group()


def test_two_rank_planning_over_gloo():
    import torch.multiprocessing as mp
    npix, world = 3072, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, npix, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = np.array(rows)
    assert rows[:, 0].sum() == npix * (npix + 1) // 2                       # all pixel pairs, once
    assert rows[:, 1].sum() == partition.packed_size(3 * npix)               # all matrix entries, once
    assert rows[:, 2].sum() == partition.packed_size(3 * npix)               # strips tile the packed triangle
    assert rows[0, 3] == 0 and rows[1, 3] > 0                                # only the right-hand rank keeps an outbox
    assert tmax == 2.0


@pytest.mark.parametrize("n_batch,parts", [(1024, 8), (1000, 8), (37, 2), (5, 4), (0, 3), (16, 1)])
def test_batch_partition_is_slab_aligned_and_complete(n_batch, parts):
    b = partition.batch_partition(n_batch, parts)
    assert b[0] == 0 and b[-1] == n_batch and len(b) == parts + 1
    assert all(x <= y for x, y in zip(b, b[1:]))
    assert all(x % 16 == 0 for x in b[1:-1])
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert max(sizes) - min(sizes) <= 16 + 15          # balanced to one slab (plus the ragged last one)


# ---------------------------------------------------------------- symmetry-orbit sharding (cmg_tqu_orbit_sharded)

@pytest.mark.parametrize("nside,world,mode", [(64, 1, 0), (64, 2, 0), (64, 4, 0), (64, 8, 0), (64, 8, 1), (16, 3, 0), (8, 2, 1)])
def test_orbit_partition_balances_evaluated_pairs(nside, world, mode):
    f = nside * nside
    b = partition.orbit_partition(nside, world, mode)
    assert b[0] == 0 and b[-1] == f and len(b) == world + 1
    assert all(x <= y for x, y in zip(b, b[1:])) and all(x % 32 == 0 for x in b)
    pairs = [partition.orbit_pairs_in_range(b[r], b[r + 1], f, mode) for r in range(world)]
    assert sum(pairs) == partition.orbit_pairs_in_range(0, f, f, mode)
    assert sum(pairs) == sum(partition.orbit_column_cost(q, f, mode) for q in range(f))
    if nside == 64:
        assert max(pairs) <= 1.05 * sum(pairs) / world                       # to one 32-column tile
    # the 36 packed column runs of all ranks tile the packed triangle
    sizes = sum(sum(sum(row) for row in partition.orbit_strip_sizes(nside, b[r], b[r + 1])) for r in range(world))
    assert sizes == partition.packed_size(3 * 12 * f)


def _orbit_worker(rank, world, port, nside, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = nside * nside
    b = partition.orbit_partition(nside, world, 0)
    mine = torch.tensor([b[rank], b[rank + 1], partition.orbit_pairs_in_range(b[rank], b[rank + 1], f, 0),
                         sum(sum(r) for r in partition.orbit_strip_sizes(nside, b[rank], b[rank + 1]))], dtype=torch.int64)
    allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    if rank == 0:
        out.put(torch.stack(allr).tolist())
    dist.destroy_process_group()


def test_two_rank_orbit_planning_over_gloo():
    import torch.multiprocessing as mp
    nside, world = 16, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_orbit_worker, args=(r, world, port, nside, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows = np.array(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = nside * nside
    assert rows[0, 0] == 0 and rows[0, 1] == rows[1, 0] and rows[1, 1] == f      # the ranks' in-face ranges tile [0, nside^2)
    assert rows[:, 2].sum() == partition.orbit_pairs_in_range(0, f, f, 0)        # every source pair evaluated once
    assert rows[:, 3].sum() == partition.packed_size(3 * 12 * f)                 # strips tile the packed triangle


class _FakeCtx:
    """stands in for capi.Context where only device_malloc / device_free are needed (no GPU)"""

    def __init__(self):
        self.next = 0x10000000
        self.live = {}

    def device_malloc(self, nbytes):
        p = self.next
        self.next += (nbytes + 255) // 256 * 256
        self.live[p] = nbytes
        return p

    def device_free(self, ptr):
        del self.live[ptr]


@pytest.mark.parametrize("world,mode", [(1, 0), (2, 0), (3, 1)])
def test_orbit_shard_descriptor_layout(world, mode):
    """OrbitShardedTQU: the 36 strips are laid out back to back in (strip, face) order = ascending packed columns, the outbox
    is one compact buffer of the size cmg_orbit_outbox_layout states, and the send / receive counts of the exchange agree
    between every pair of ranks; with one rank the strips buffer IS the packed triangle."""
    from cosmopp_b200 import capi, multigpu
    nside = 16
    f, n = nside * nside, 12 * nside * nside
    for rank in range(world):
        ctx = _FakeCtx()
        sh = multigpu.OrbitShardedTQU(ctx, nside, rank, world, mode)
        q0, q1 = sh.q0, sh.q1
        assert (sh.shard.n_ranks, sh.shard.rank) == (world, rank) and [sh.shard.bounds[k] for k in range(world + 1)] == sh.bounds
        sizes = partition.orbit_strip_sizes(nside, q0, q1)
        off = 0
        for s in range(3):
            for face in range(12):
                assert sh.shard.strip[s][face] == sh.strips.ptr + 8 * off
                off += sizes[s][face]
        assert off == sh.strips.n and ctx.live[sh.strips.ptr] >= 8 * off
        if world == 1:
            assert sh.outbox is None and off == partition.packed_size(3 * n)
            for s in range(3):                     # single owner: strip (s, face) starts at its packed offset
                for face in range(12):
                    assert sh.shard.strip[s][face] == sh.strips.ptr + 8 * partition.packed_size(s * n + face * f)
            assert sh.shard.outbox is None and sh.send_counts == [0] and sh.recv_counts == [0]
        else:
            plan = capi.orbit_plan(nside, mode)
            offsets = partition.orbit_outbox_offsets(plan, sh.bounds, rank)
            assert sh.layouts[rank] == offsets and sh.outbox.n == offsets[-1] and sh.shard.outbox == sh.outbox.ptr
            assert sh.send_counts == [offsets[d + 1] - offsets[d] for d in range(world)] and sh.send_counts[rank] == 0
            # what this rank expects from r is what r's layout says it sends here
            for r in range(world):
                other = partition.orbit_outbox_offsets(plan, sh.bounds, r)
                assert sh.recv_counts[r] == (0 if r == rank else other[rank + 1] - other[rank])
            # all ranks together: strips (with holes) + outboxes = the matrix's entries once, i.e. the outboxes are exactly the
            # entries missing from the strips: a third of the entries of the pairs whose row pixel lies outside the column's range
            n_a, n_b = partition.orbit_combo_counts(plan)
            assert (n_a, n_b) == ((189, 90) if mode == 0 else (198, 36))
        assert sh.sizes_of(rank) == (sh.strips.n, sh.outbox.n if sh.outbox is not None else 0)
        other = sh.shard_of(rank, 0x5000, 0x9000 if world > 1 else 0)
        assert other.strip[0][0] == 0x5000 and (world == 1 or other.outbox == 0x9000)
        sh.close()
        assert not ctx.live
