"""GPU parity of the symmetry-orbit path (cmg_tqu_orbit, cmg_tqu_orbit_sharded; cosmopp_b200/csrc/orbit.cuh) against the
CPU oracle: same gate as the every-pair kernel, 1e-11 of the diagonal of the block, layout position by position."""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu

REL_TOL = 1e-11


def _scale(n, want):
    """per-entry scale: TT diagonal for the columns of the T strip, QQ diagonal elsewhere"""
    from cosmopp_b200 import capi
    dim = 3 * n
    scale = np.full(capi.packed_size(dim), want[capi.packed_index(n, n)])
    scale[:capi.packed_size(n)] = want[0]
    return scale


def _inputs(ctx, nside, lmax):
    from cosmopp_b200 import capi
    ctx.set_kernel_variant(0)
    ctx.set_pixels(nside)
    spectra = synthetic_cl(lmax, pol=True)
    f = capi.window_beam(lmax, 10.0)
    return spectra, capi.tqu_weights(*spectra, f, f)


@pytest.mark.parametrize("nside,lmax", [(8, 20), (16, 47)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_orbit_matches_oracle(gpu_ctx, oracle_api, nside, lmax, mode):
    import torch
    from cosmopp_b200 import capi
    spectra, w = _inputs(gpu_ctx, nside, lmax)
    n = gpu_ctx.npix
    out = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    gpu_ctx.tqu_orbit(*w, out, mode)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert not np.isnan(got).any()                      # every entry of the triangle is written
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0)
    assert (np.abs(got - want) / _scale(n, want)).max() <= REL_TOL
    # and against the every-pair kernel: differences are roundings of n_i.n_j only
    ref = torch.empty_like(out)
    gpu_ctx.tqu(*w, gpu_ctx.tqu_layout_single(ref))
    torch.cuda.synchronize()
    assert (np.abs(got - ref.cpu().numpy()) / _scale(n, want)).max() <= 1e-13


@pytest.mark.parametrize("world,mode", [(2, 0), (3, 0), (3, 1)])
def test_orbit_shards_assemble_to_the_whole_matrix(gpu_ctx, oracle_api, world, mode):
    """every rank's pieces (36 packed column runs + outbox blocks) generated on this GPU one after the other"""
    import torch
    from cosmopp_b200 import capi, multigpu
    nside, lmax = 16, 30
    spectra, w = _inputs(gpu_ctx, nside, lmax)
    n = gpu_ctx.npix
    ranks = [multigpu.OrbitShardedTQU(gpu_ctx, nside, r, world, mode) for r in range(world)]
    assert [r.q0 for r in ranks][1:] == [r.q1 for r in ranks][:-1] and ranks[-1].q1 == nside * nside
    for r in ranks:
        for b in r.pieces():
            b.tensor().fill_(float("nan"))
        r.generate(w)
    full = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    for parts in (1, 2):                                # strips of all ranks first, then the outboxes
        for r in ranks:
            r.assemble_into(full, parts)
    torch.cuda.synchronize()
    got = full.cpu().numpy()
    for r in ranks:
        r.close()
    assert not np.isnan(got).any()
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0)
    assert (np.abs(got - want) / _scale(n, want)).max() <= REL_TOL


@pytest.mark.parametrize("world,mode,threads,direct", [(2, 0, 0, 0), (3, 0, 3, 0), (3, 1, 2, 0x49), (4, 2, 0, 0), (2, 0, 2, 0x1a5)])
def test_orbit_exchange_completes_the_strips(gpu_ctx, oracle_api, world, mode, threads, direct):
    """The exchange step with every rank on this one GPU: block(r -> d) of r's outbox is placed into d's strips by
    cmg_tqu_orbit_scatter_inbox, after which the strips alone are the matrix -- on the device (assembly of the strips only) and
    on the host (cmg_orbit_strips_to_host into one whole packed matrix, plain or with the host filling in the rotated images)."""
    import torch
    from cosmopp_b200 import capi, multigpu
    nside, lmax = 16, 30
    spectra, w = _inputs(gpu_ctx, nside, lmax)
    n = gpu_ctx.npix
    ranks = [multigpu.OrbitShardedTQU(gpu_ctx, nside, r, world, mode) for r in range(world)]
    for r in ranks:
        for b in r.pieces():
            b.tensor().fill_(float("nan"))
        r.generate(w)
    for d in ranks:
        for s in ranks:
            if s.rank != d.rank and d.recv_counts[s.rank]:
                assert d.recv_counts[s.rank] == s.send_counts[d.rank]
                gpu_ctx.tqu_orbit_scatter_inbox(d.shard, s.rank, s.outbox.ptr + 8 * s.layouts[s.rank][d.rank], mode)
    full = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    host = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64).pin_memory()
    for r in ranks:
        r.assemble_into(full, 1)                        # strips only
        r.to_host(host, threads, direct)
    torch.cuda.synchronize()
    got = full.cpu().numpy()
    for r in ranks:
        r.close()
    assert not np.isnan(got).any()
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0)
    assert (np.abs(got - want) / _scale(n, want)).max() <= REL_TOL
    got_host = host.numpy()
    assert not np.isnan(got_host).any()
    if threads == 0:
        assert np.array_equal(got_host, got)
    assert (np.abs(got_host - want) / _scale(n, want)).max() <= REL_TOL


def test_whole_call_takes_the_orbit_path_on_the_full_sky(gpu_ctx, oracle_api):
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 8, 16
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    spectra = synthetic_cl(lmax, pol=True)
    out = torch.empty(capi.packed_size(3 * gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    before = gpu_ctx.launches
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out)
    assert gpu_ctx.launches - before == 3               # the launches of cmg_tqu_orbit: one per transposed-image mask (0, 8, 12)
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0)
    assert (np.abs(out.numpy() - want) / _scale(gpu_ctx.npix, want)).max() <= REL_TOL
    gpu_ctx.set_kernel_variant(142)                     # a pinned variant switches the routing off
    before = gpu_ctx.launches
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out)
    assert gpu_ctx.launches - before == 1
    gpu_ctx.set_kernel_variant(0)
    assert (np.abs(out.numpy() - want) / _scale(gpu_ctx.npix, want)).max() <= REL_TOL


def test_orbit_path_refuses_what_it_cannot_do(gpu_ctx, oracle_api):
    import torch
    from cosmopp_b200 import capi
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(8))
    gpu_ctx.set_pixels(8, good)
    spectra = synthetic_cl(12, pol=True)
    f = capi.window_beam(12, 10.0)
    w = capi.tqu_weights(*spectra, f, f)
    out = torch.empty(capi.packed_size(3 * gpu_ctx.npix), dtype=torch.float64, device="cuda")
    with pytest.raises(capi.CmgError):
        gpu_ctx.tqu_orbit(*w, out)                      # masked sky: no rotation symmetry
    gpu_ctx.set_pixels(4)
    out = torch.empty(capi.packed_size(3 * gpu_ctx.npix), dtype=torch.float64, device="cuda")
    with pytest.raises(capi.CmgError):
        gpu_ctx.tqu_orbit(*w, out)                      # nside < 8: a 64 x 32 tile does not fit a base face


@pytest.mark.parametrize("nside,lmax", [(16, 47), (32, 96), (32, 40)])
def test_tt_orbit_matches_the_every_pair_kernel(gpu_ctx, oracle_api, nside, lmax):
    import torch
    from cosmopp_b200 import capi
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    a = capi.tt_weights(synthetic_cl(lmax), capi.window_beam(lmax, 10.0))
    ref = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series(a, ref)
    out = torch.full_like(ref, float("nan"))
    gpu_ctx.legendre_series_orbit(a, out)
    torch.cuda.synchronize()
    got, want = out.cpu().numpy(), ref.cpu().numpy()
    assert not np.isnan(got).any()
    assert np.abs(got - want).max() <= 1e-13 * want[0]


@pytest.mark.parametrize("world", [1, 2, 3])
def test_tt_orbit_shards_fill_the_matrix_without_exchange(gpu_ctx, oracle_api, world):
    """cmg_legendre_series_orbit_sharded: every rank's 12 runs of packed columns (generated one after the other on this GPU)
    placed into one triangle = the oracle's matrix; nothing but the strips exists"""
    import torch
    from cosmopp_b200 import capi, multigpu
    nside, lmax = 16, 47
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    cl = synthetic_cl(lmax)
    a = capi.tt_weights(cl, capi.window_beam(lmax, 10.0))
    full = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64, device="cuda")
    pairs = 0
    for r in range(world):
        sh = multigpu.OrbitShardedTT(gpu_ctx, nside, r, world)
        sh.strips.tensor().fill_(float("nan"))
        sh.generate(a)
        sh.place_into(full)
        torch.cuda.synchronize()
        pairs += sh.pairs
        sh.close()
    got = full.cpu().numpy()
    assert not np.isnan(got).any()
    want = oracle_api.cl_to_cmatrix(cl, nside, 10.0)
    assert np.abs(got - want).max() <= REL_TOL * want[0]
    f = nside * nside
    assert abs(pairs - 22.5 * f * f) <= 6 * f            # Nside = 16: the single-launch plan without transposed images even for one owner


def test_tt_whole_call_takes_the_orbit_path_on_the_full_sky(gpu_ctx, oracle_api):
    """cl_to_cmatrix and fiducial_matrix (series weights a_0 = a_1 != 0 as well) against the oracle"""
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 16, 30
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    cl = synthetic_cl(4 * nside)
    out = torch.empty(capi.packed_size(gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix(cl[:lmax + 1], 10.0, out)
    want = oracle_api.cl_to_cmatrix(cl[:lmax + 1], nside, 10.0)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]
    gpu_ctx.fiducial_matrix(cl, lmax, 10.0, out)
    want = oracle_api.fiducial_matrix(cl, nside, lmax, 10.0)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]
    gpu_ctx.set_kernel_variant(142)                     # any pinned variant: every-pair kernel, same numbers to rounding
    again = torch.empty_like(out)
    gpu_ctx.fiducial_matrix(cl, lmax, 10.0, again)
    gpu_ctx.set_kernel_variant(0)
    assert np.abs(again.numpy() - out.numpy()).max() <= 1e-13 * want[0]


def test_orbit_mode2_is_bit_identical_to_mode0(gpu_ctx):
    import torch
    from cosmopp_b200 import capi
    spectra, w = _inputs(gpu_ctx, 16, 40)
    n = gpu_ctx.npix
    a = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    b = torch.full_like(a, float("nan"))
    gpu_ctx.tqu_orbit(*w, a, 0)
    gpu_ctx.tqu_orbit(*w, b, 2)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_orbit_mode3_refuses_several_owners(gpu_ctx):
    """the meridian mirror puts images into columns of other ranks: single owner only"""
    from cosmopp_b200 import capi, multigpu
    gpu_ctx.set_pixels(16)
    f = capi.window_beam(20, 10.0)
    w = capi.tqu_weights(*synthetic_cl(20, pol=True), f, f)
    sh = multigpu.OrbitShardedTQU(gpu_ctx, 16, 0, 2)
    try:
        with pytest.raises(capi.CmgError):
            gpu_ctx.tqu_orbit_sharded(*w, sh.shard, 3)
    finally:
        sh.close()
