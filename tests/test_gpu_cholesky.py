"""GPU: the packed in-place Cholesky factorisation and the consumer built on it (cmg_packed_cholesky, cmg_like_*; reference
source/matrix_impl.cpp:236-263 = LAPACK dpptrf 'U' on the packed storage of include/matrix_impl.hpp:495-502, called from
Likelihood::construct, source/likelihood.cpp:100-133).  Checked against numpy on the same matrices; tolerance 1e-9 relative for
chi^2 and log det (the reference's own test compares likelihoods statistically only)."""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu


def pack_upper(M):
    n = M.shape[0]
    iu = np.triu_indices(n)
    out = np.empty(n * (n + 1) // 2)
    out[iu[1] * (iu[1] + 1) // 2 + iu[0]] = M[iu]
    return out


def unpack_upper(p, n):
    U = np.zeros((n, n))
    iu = np.triu_indices(n)
    U[iu] = p[iu[1] * (iu[1] + 1) // 2 + iu[0]]
    return U


def random_spd(n, seed):
    rs = np.random.RandomState(seed)
    B = rs.normal(size=(n, n // 2 + 3))
    return B @ B.T + 0.5 * n * np.diag(rs.uniform(0.5, 1.5, n))


@pytest.mark.parametrize("n", [1, 7, 128, 129, 300, 1000, 1537])
def test_packed_cholesky_matches_numpy(gpu_ctx, n):
    import torch
    A = random_spd(n, 100 + n)
    d = torch.from_numpy(pack_upper(A)).cuda()
    assert gpu_ctx.packed_cholesky(d, n) == 0
    U = unpack_upper(d.cpu().numpy(), n)
    want = np.linalg.cholesky(A).T
    assert np.abs(U - want).max() <= 1e-12 * np.abs(want).max()
    assert np.abs(U.T @ U - A).max() <= 1e-13 * np.abs(A).max() * n
    assert abs(gpu_ctx.packed_cholesky_logdet(d, n) - np.linalg.slogdet(A)[1]) <= 1e-12 * abs(np.linalg.slogdet(A)[1]) + 1e-12
    # y = U^-T t for several right-hand sides (more than one pass of eight)
    rs = np.random.RandomState(5)
    T = rs.normal(size=(11, n))
    t = torch.from_numpy(T.copy()).cuda()                    # row k = right-hand side k = column-major n x 11
    gpu_ctx.packed_cholesky_solve(d, n, t, 11)
    Y = t.cpu().numpy()
    want_y = np.linalg.solve(want.T, T.T).T
    assert np.abs(Y - want_y).max() <= 1e-11 * np.abs(want_y).max()


def test_packed_cholesky_reports_the_failing_minor(gpu_ctx):
    import torch
    n = 400
    A = random_spd(n, 3)
    A[250, 250] = -1.0                                       # leading minors up to 250 are fine, the 251st is not
    d = torch.from_numpy(pack_upper(A)).cuda()
    assert gpu_ctx.packed_cholesky(d, n) == 251


@pytest.mark.parametrize("masked", [True, False])
def test_tqu_likelihood_consumer_on_the_packed_factor(gpu_ctx, oracle_api, masked):
    """the polarized matrix this library's headline produces, consumed where it lies: [T;Q;U] at Nside = 16 (BASELINE configs[1];
    full sky: dimension 9216) + noise -> packed Cholesky -> chi^2, log det for a few maps, against numpy on the oracle's matrix;
    and against the cuSOLVER route on the same device buffers"""
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 16, 47
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside, good)
    n = 3 * gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    w = capi.tqu_weights(*spectra, f, f)
    d_c = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    if good is None:
        gpu_ctx.tqu_orbit(*w, d_c, 0)
    else:
        gpu_ctx.tqu(*w, gpu_ctx.tqu_layout_single(d_c))
    noise = np.zeros(capi.packed_size(n))
    sig = np.where(np.arange(n) < n // 3, 2.0, 0.3)         # white noise: 2 muK in T, 0.3 in Q and U
    noise[np.arange(n) * (np.arange(n) + 1) // 2 + np.arange(n)] = sig ** 2
    d_n = torch.from_numpy(noise).cuda()
    rs = np.random.RandomState(11)
    maps = rs.normal(size=(5, n)) * 10.0
    from cosmopp_b200.likelihood import Likelihood
    like = Likelihood(gpu_ctx, d_c, None, d_n, n)
    _, chi2, logdet = like.calculate(maps)
    like.close()
    gpu_ctx.set_like_method(1)
    try:
        dense = Likelihood(gpu_ctx, d_c, None, d_n, n)
        _, chi2_d, logdet_d = dense.calculate(maps)
        dense.close()
    finally:
        gpu_ctx.set_like_method(0)
    S = oracle_api.unpack_symmetric(oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good), n) + np.diag(sig ** 2)
    want_logdet = np.linalg.slogdet(S)[1] + 29677.0566
    want_chi2 = np.einsum("kn,nk->k", maps, np.linalg.solve(S, maps.T))
    assert np.abs(chi2 - want_chi2).max() <= 1e-9 * np.abs(want_chi2).max()
    assert abs(logdet - want_logdet) <= 1e-9 * abs(want_logdet)
    assert np.abs(chi2 - chi2_d).max() <= 1e-9 * np.abs(chi2_d).max() and abs(logdet - logdet_d) <= 1e-9 * abs(logdet_d)


class _NoComm:
    """ranks emulated on one GPU share the U_kk buffer and the dense panel: nothing to exchange"""

    def broadcast(self, tensor, src):
        pass

    def all_reduce(self, tensor):
        pass


@pytest.mark.parametrize("lookahead", [True, False])
@pytest.mark.parametrize("group", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [128, 129, 500, 1153, 2100])
def test_packed_cholesky_groups_of_blocks(gpu_ctx, n, group, lookahead):
    """cmg_set_cholesky_group / cmg_set_cholesky_lookahead: every group size gives the same factor (strip updates + one trailing
    update per group), with the next group factorised beside the trailing update or after it"""
    import torch
    A = random_spd(n, 500 + n)
    d = torch.from_numpy(pack_upper(A)).cuda()
    gpu_ctx.set_cholesky_group(group)
    gpu_ctx.set_cholesky_lookahead(lookahead)
    try:
        assert gpu_ctx.packed_cholesky(d, n) == 0
    finally:
        gpu_ctx.set_cholesky_lookahead(True)
        gpu_ctx.set_cholesky_group(0)
    want = np.linalg.cholesky(A).T
    assert np.abs(unpack_upper(d.cpu().numpy(), n) - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("n, world, cycles, group", [(1280, 2, 2, 1), (1536 + 77, 3, 2, 2), (2048, 4, 4, 2), (2048 + 300, 2, 3, 4)])
def test_sharded_cholesky_ranks_in_lock_step(gpu_ctx, n, world, cycles, group):
    """the step functions of the sharded factorisation (cmg_chol_*; multigpu.ShardedCholesky) with `world` ranks emulated on one
    GPU: the columns are dealt out in runs of 128 ... 384 columns, `cycles` runs per rank interleaved like the strips of
    cmg_orbit_shard; every rank factorises its own copy of its columns, sharing only U_kk and the dense panel.  Factor, log det
    and the solves against numpy and against cmg_packed_cholesky on the whole triangle."""
    import torch
    from cosmopp_b200 import capi, multigpu
    A = random_spd(n, 900 + n)
    packed = pack_upper(A)
    # runs: cut [0, n) at multiples of 128 into world * cycles pieces of uneven width, dealt round robin
    nb = capi.CHOL_NB
    blocks = (n + nb - 1) // nb
    rs = np.random.RandomState(n)
    cuts = np.sort(rs.choice(np.arange(1, blocks), size=world * cycles - 1, replace=False)) * nb
    edges = [0] + [int(c) for c in cuts] + [n]
    all_runs = [[] for _ in range(world)]
    for k in range(world * cycles):
        all_runs[k % world].append((edges[k], edges[k + 1]))
    off = lambda c: c * (c + 1) // 2
    bufs = [[torch.from_numpy(packed[off(b):off(e)].copy()).cuda() for b, e in all_runs[r]] for r in range(world)]
    ukk = torch.zeros(nb * (nb + 1) // 2 + nb, dtype=torch.float64, device="cuda")
    panel = torch.zeros(group * (n + capi.CHOL_PLANE_SLACK) * nb, dtype=torch.float64, device="cuda")
    ranks = [multigpu.ShardedCholesky(gpu_ctx, n, all_runs, r, [t.data_ptr() for t in bufs[r]], comm=_NoComm(), ukk=ukk, panel=panel, group=group)
             for r in range(world)]
    assert ranks[0].owners == multigpu.chol_block_owners(n, all_runs)
    gpu_ctx.chol_begin()
    for ph in ranks[0].schedule():
        for r in ranks:                                     # "diag" runs on the owner only (run_phase checks)
            r.run_phase(ph)
    assert gpu_ctx.chol_end() == 0
    got = np.empty_like(packed)
    for r in range(world):
        for (b, e), t in zip(all_runs[r], bufs[r]):
            got[off(b):off(e)] = t.cpu().numpy()
    whole = torch.from_numpy(packed.copy()).cuda()
    assert gpu_ctx.packed_cholesky(whole, n) == 0
    want = np.linalg.cholesky(A).T
    assert np.abs(unpack_upper(got, n) - want).max() <= 1e-12 * np.abs(want).max()
    assert np.abs(got - whole.cpu().numpy()).max() <= 1e-13 * np.abs(want).max()
    logdet = sum(r.logdet() for r in ranks)                  # world == 1 inside each emulated rank: its own share
    assert abs(logdet - np.linalg.slogdet(A)[1]) <= 1e-12 * abs(np.linalg.slogdet(A)[1])
    # solves: the right-hand sides are one shared buffer here (replicated + broadcast on real ranks)
    T = np.random.RandomState(3).normal(size=(3, n))
    t = torch.from_numpy(T.copy()).cuda()
    for k0, kb in ranks[0].blocks():
        gpu_ctx.chol_solve_diag(ranks[ranks[0].owners[k0 // nb]].runs, k0, kb, n, t, 3)
        if k0 + kb < n:
            for r in ranks:
                gpu_ctx.chol_solve_update(r.runs, k0, kb, n, t, 3)
    want_y = np.linalg.solve(want.T, T.T).T
    assert np.abs(t.cpu().numpy() - want_y).max() <= 1e-11 * np.abs(want_y).max()


@pytest.mark.parametrize("n, group, ahead", [(700, 4, None), (2100, 2, True), (2100, 3, True), (1700, 4, False)])
def test_sharded_cholesky_single_rank_is_the_whole_call(gpu_ctx, n, group, ahead):
    """world = 1 through the class: factorise / logdet / solve without any exchange; with the look-ahead the next group's
    blocks run on a second stream beside the rest of the trailing update (two sets of panel planes)"""
    import torch
    from cosmopp_b200 import multigpu
    A = random_spd(n, 41)
    d = torch.from_numpy(pack_upper(A)).cuda()
    gpu_ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ch = multigpu.ShardedCholesky(gpu_ctx, n, [[(0, n)]], 0, [d.data_ptr()], group=group)
    assert ch.factorise(lookahead=ahead) == 0
    want = np.linalg.cholesky(A).T
    assert np.abs(unpack_upper(d.cpu().numpy(), n) - want).max() <= 1e-12 * np.abs(want).max()
    assert abs(ch.logdet() - np.linalg.slogdet(A)[1]) <= 1e-12 * abs(np.linalg.slogdet(A)[1])
    t = torch.from_numpy(np.random.RandomState(2).normal(size=(2, n))).cuda()
    T = t.cpu().numpy().copy()
    ch.solve(t)
    want_y = np.linalg.solve(want.T, T.T).T
    assert np.abs(t.cpu().numpy() - want_y).max() <= 1e-11 * np.abs(want_y).max()
