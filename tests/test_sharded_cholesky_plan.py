"""No GPU: the host side of the sharded Cholesky factorisation (cosmopp_b200/multigpu.py: ShardedCholesky, chol_block_owners;
include/cmg.h: cmg_chol_*).  The step functions are stood in for by numpy statements of what the kernels do (a test double, not
a product path), so that the order of steps, the ownership of blocks and the two exchanges of a step run for real over gloo
with two ranks.  The kernels themselves are checked on the GPU in tests/test_gpu_cholesky.py."""
import socket

import numpy as np
import pytest

from cosmopp_b200 import capi, multigpu, partition

NB = capi.CHOL_NB


def off(c):
    return c * (c + 1) // 2


def test_block_owners_of_orbit_runs():
    nside, world = 16, 2
    b = partition.orbit_partition_blocks(nside, world)
    assert b == [0, 128, 256]
    all_runs = [partition.orbit_column_runs(nside, b[r], b[r + 1]) for r in range(world)]
    n = 3 * 12 * nside * nside
    owners = multigpu.chol_block_owners(n, all_runs)
    assert len(owners) == n // NB and owners == [k % 2 for k in range(n // NB)]      # 36 cycles of (rank 0, rank 1)
    assert all(len(r) == 36 for r in all_runs)
    # Nside = 64 over 8 ranks: every boundary on a block edge, every block owned once, the same number of blocks everywhere
    b = partition.orbit_partition_blocks(64, 8)
    all_runs = [partition.orbit_column_runs(64, b[r], b[r + 1]) for r in range(8)]
    counts = np.bincount(multigpu.chol_block_owners(147456, all_runs), minlength=8)
    assert counts.tolist() == [144] * 8
    assert partition.orbit_partition_blocks(32, 3) == [0, 384, 768, 1024]
    pairs = [partition.orbit_pairs_in_range(b[r], b[r + 1], 4096) for r in range(8)]
    assert max(pairs) <= 1.15 * sum(pairs) / 8                                     # generation stays balanced to ~ 12 %


def test_block_owners_reject_bad_runs():
    with pytest.raises(ValueError):
        multigpu.chol_block_owners(512, [[(0, 200)], [(200, 512)]])              # boundary off the block grid
    with pytest.raises(ValueError):
        multigpu.chol_block_owners(512, [[(0, 256)], [(128, 512)]])              # overlap
    with pytest.raises(ValueError):
        multigpu.chol_block_owners(512, [[(0, 128)], [(256, 512)]])              # hole
    assert multigpu.chol_block_owners(300, [[(0, 128), (256, 300)], [(128, 256)]]) == [0, 1, 0]


class NumpyStepCtx:
    """what cmg_chol_* do, in numpy, on host arrays addressed by fake pointers (the position in `arrays`)"""
    stream_handle = None

    def __init__(self, arrays):
        self.arrays = arrays          # fake pointer -> 1-d array holding a run of packed columns
        self.info = 0

    def _cols(self, runs):
        for r in range(runs.n_runs):
            b, e = runs.col_begin[r], runs.col_end[r]
            yield b, e, self.arrays[runs.d_run[r]]

    def _column(self, runs, j):
        for b, e, a in self._cols(runs):
            if b <= j < e:
                return a[off(j) - off(b):off(j + 1) - off(b)]
        raise KeyError(j)

    def chol_begin(self):
        self.info = 0

    def chol_end(self):
        return self.info

    def chol_diag(self, runs, k0, kb, ukk):
        S = np.zeros((kb, kb))
        for c in range(kb):
            S[:c + 1, c] = self._column(runs, k0 + c)[k0:k0 + c + 1]
        try:
            U = np.linalg.cholesky(S + np.triu(S, 1).T).T
        except np.linalg.LinAlgError:
            self.info = self.info or k0 + 1
            return
        u = ukk.numpy()
        for c in range(kb):
            self._column(runs, k0 + c)[k0:k0 + c + 1] = U[:c + 1, c]
            u[off(c):off(c + 1)] = U[:c + 1, c]
        u[off(kb):off(kb) + kb] = 1.0 / np.diag(U)

    def _ukk(self, ukk, kb):
        u = ukk.numpy()
        U = np.zeros((kb, kb))
        for c in range(kb):
            U[:c + 1, c] = u[off(c):off(c + 1)]
        return U

    def chol_panel(self, runs, k0, kb, ukk, plane, panel_col0):
        U = self._ukk(ukk, kb)
        P = plane.numpy()
        for b, e, _ in self._cols(runs):
            for j in range(max(b, k0 + kb), e):
                col = self._column(runs, j)
                x = np.linalg.solve(U.T, col[k0:k0 + kb])
                col[k0:k0 + kb] = x
                P[(j - panel_col0) * NB:(j - panel_col0) * NB + kb] = x

    def chol_syrk(self, runs, k0, kb, panel, plane_stride, panel_col0, strip_only, shift=0):
        k1 = k0 + kb + shift
        planes = [panel.numpy()[s * plane_stride:(s + 1) * plane_stride].reshape(-1, NB) for s in range(kb // NB)]
        for b, e, _ in self._cols(runs):
            for j in range(max(b, k1), e):
                col = self._column(runs, j)
                i1 = min(j + 1, k1 + NB) if strip_only else j + 1
                for P in planes:
                    col[k1:i1] -= P[k1 - panel_col0:i1 - panel_col0] @ P[j - panel_col0]

    def chol_logdet_runs(self, runs):
        return 2.0 * sum(np.log(self._column(runs, j)[j]) for b, e, _ in self._cols(runs) for j in range(b, e))

    def chol_solve_diag(self, runs, k0, kb, n, rhs, n_rhs):
        U = np.zeros((kb, kb))
        for c in range(kb):
            U[:c + 1, c] = self._column(runs, k0 + c)[k0:k0 + c + 1]
        T = rhs.numpy()
        T[:, k0:k0 + kb] = np.linalg.solve(U.T, T[:, k0:k0 + kb].T).T

    def chol_solve_update(self, runs, k0, kb, n, rhs, n_rhs):
        T = rhs.numpy()
        for b, e, _ in self._cols(runs):
            for j in range(max(b, k0 + kb), e):
                T[:, j] -= T[:, k0:k0 + kb] @ self._column(runs, j)[k0:k0 + kb]


def spd(n, seed):
    rs = np.random.RandomState(seed)
    B = rs.normal(size=(n, n + 5))
    return B @ B.T + n * np.eye(n)


def packed_of(A):
    n = A.shape[0]
    out = np.empty(off(n))
    for j in range(n):
        out[off(j):off(j + 1)] = A[:j + 1, j]
    return out


def _worker(rank, world, port, n, all_runs, bad, group, ahead, out):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    A = spd(n, 7)
    if bad:
        A[bad, bad] = -5.0
    packed = packed_of(A)
    arrays = {1000 + k: packed[off(b):off(e)].copy() for k, (b, e) in enumerate(all_runs[rank])}
    ctx = NumpyStepCtx(arrays)
    ukk = torch.zeros(off(NB) + NB, dtype=torch.float64)
    panel = torch.zeros((2 if ahead else 1) * group * (n + capi.CHOL_PLANE_SLACK) * NB, dtype=torch.float64)
    ch = multigpu.ShardedCholesky(ctx, n, all_runs, rank, list(arrays.keys()), ukk=ukk, panel=panel, group=group)
    info = ch.factorise(lookahead=ahead)
    res = {"rank": rank, "info": info}
    if not bad:
        res["logdet"] = ch.logdet()
        T = torch.from_numpy(np.random.RandomState(1).normal(size=(2, n)))
        ch.solve(T)
        res["y"] = T.numpy().copy()
        res["runs"] = {k - 1000: v for k, v in arrays.items()}
    out.put(res)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_schedule_of_phases():
    n = 5 * NB + 40
    all_runs = [[(0, n)]]
    ch = multigpu.ShardedCholesky(NumpyStepCtx({}), n, all_runs, 0, [1], ukk=np.zeros(1), panel=_zeros(3 * (n + capi.CHOL_PLANE_SLACK) * NB), group=3)
    kinds = [p[0] for p in ch.schedule()]
    assert kinds == ["diag", "panel", "strip", "diag", "panel", "strip", "diag", "panel", "syrk",
                     "diag", "panel", "strip", "diag", "panel", "strip", "diag"]
    one = multigpu.ShardedCholesky(NumpyStepCtx({}), n, all_runs, 0, [1], ukk=np.zeros(1), panel=_zeros((n + capi.CHOL_PLANE_SLACK) * NB), group=1)
    assert [p[0] for p in one.schedule()] == ["diag", "panel", "syrk"] * 5 + ["diag"]
    assert [p[0] for p in multigpu.ShardedCholesky(NumpyStepCtx({}), 256, [[(0, 256)]], 0, [1], ukk=np.zeros(1),
                                                   panel=_zeros(2 * (256 + capi.CHOL_PLANE_SLACK) * NB), group=2).schedule()] == ["diag", "panel", "strip", "diag"]


def _zeros(k):
    import torch
    return torch.zeros(k, dtype=torch.float64)


@pytest.mark.parametrize("bad, group, ahead", [(0, 1, False), (0, 2, False), (0, 3, False), (300, 2, False), (0, 1, True), (0, 2, True), (300, 2, True)])
def test_two_ranks_factorise_over_gloo(bad, group, ahead):
    import torch.multiprocessing as mp
    n, world = 5 * NB + 40, 2
    all_runs = [[(0, 128), (384, 512)], [(128, 384), (512, n)]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, all_runs, bad, group, ahead, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if bad:
        assert [d["info"] for d in res] == [bad + 1 - bad % NB] * 2      # the numpy stand-in reports the failing BLOCK; same on both ranks
        return
    assert [d["info"] for d in res] == [0, 0]
    A = spd(n, 7)
    want = np.linalg.cholesky(A).T
    got = np.zeros(off(n))
    for r in range(world):
        for k, (b, e) in enumerate(all_runs[r]):
            got[off(b):off(e)] = res[r]["runs"][k]
    assert np.abs(got - packed_of(want)).max() <= 1e-12 * np.abs(want).max()
    for d in res:
        assert abs(d["logdet"] - np.linalg.slogdet(A)[1]) <= 1e-12 * abs(np.linalg.slogdet(A)[1])
    T = np.random.RandomState(1).normal(size=(2, n))
    want_y = np.linalg.solve(want.T, T.T).T
    assert np.abs(res[0]["y"] - want_y).max() <= 1e-11 * np.abs(want_y).max()
    assert np.array_equal(res[0]["y"], res[1]["y"])                      # every rank ends with the same y
