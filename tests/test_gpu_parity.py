"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): every entry within 1e-11 relative to the diagonal of its block;
index layout bit-exact (checked by comparing whole packed arrays position by position).
"""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu

REL_TOL = 1e-11


def _torch():
    import torch
    return torch


def _tt_gpu(ctx, cl, nside, fwhm, good=None):
    torch = _torch()
    from cosmopp_b200 import capi
    ctx.set_pixels(nside, good)
    n = ctx.npix
    f = capi.window_beam(len(cl) - 1, fwhm)
    a = capi.tt_weights(cl, f)
    out = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64, device="cuda")
    ctx.legendre_series(a, out)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("nside,lmax,masked", [(1, 5, False), (2, 8, False), (4, 12, False), (8, 24, True),
                                                (16, 47, False), (16, 47, True), (16, 30, True)])
def test_tt_matches_oracle(gpu_ctx, oracle_api, nside, lmax, masked):
    cl = synthetic_cl(lmax)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0, good)
    want = oracle_api.cl_to_cmatrix(cl, nside, 10.0, good=good)
    assert got.shape == want.shape
    assert not np.isnan(got).any()
    diag = want[0]
    assert np.abs(got - want).max() <= REL_TOL * diag


@pytest.mark.skipif("not __import__('oracle.api').api.have_ref()")
def test_tt_matches_reference_object_code(gpu_ctx, oracle_api):
    nside, lmax = 8, 20
    cl = synthetic_cl(lmax)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0, good)
    want = oracle_api.ref_cl_to_cmatrix(cl, nside, 10.0, good=good)
    assert np.abs(got - want).max() <= REL_TOL * want[0]


def test_tt_host_api_and_fiducial(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 8, 20
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    cl = synthetic_cl(4 * nside)
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    out = torch.empty(capi.packed_size(n), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix(cl[:lmax + 1], 10.0, out)
    want = oracle_api.cl_to_cmatrix(cl[:lmax + 1], nside, 10.0, good=good)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]
    gpu_ctx.fiducial_matrix(cl, lmax, 10.0, out)
    want = oracle_api.fiducial_matrix(cl, nside, lmax, 10.0, good=good)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]


def test_tt_column_shards_tile_the_triangle(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi, partition
    nside, lmax = 8, 16
    cl = synthetic_cl(lmax)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    a = capi.tt_weights(cl, capi.window_beam(lmax, 10.0))
    whole = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series(a, whole)
    for parts in (2, 3, 8):
        b = partition.column_partition(n, parts, align=16)
        pieces = []
        for k in range(parts):
            piece = torch.full((partition.tt_shard_size(b[k], b[k + 1]),), float("nan"), dtype=torch.float64, device="cuda")
            gpu_ctx.legendre_series(a, piece, b[k], b[k + 1])
            pieces.append(piece)
        assert torch.equal(torch.cat(pieces), whole)


def _tqu_gpu(ctx, spectra, nside, fwhm, good=None):
    torch = _torch()
    from cosmopp_b200 import capi
    ctx.set_pixels(nside, good)
    n = ctx.npix
    lmax = len(spectra[0]) - 1
    f = capi.window_beam(lmax, fwhm)
    a = capi.tqu_weights(*spectra, f, f)
    out = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    ctx.tqu(*a, ctx.tqu_layout_single(out))
    torch.cuda.synchronize()
    return out.cpu().numpy(), n


def _block_diag_scales(packed, n):
    from oracle import api
    tt = packed[api.packed_size(1) - 1]
    qq = packed[(n) * (n + 1) // 2 + n]
    return tt, qq


def _assert_tqu_close(got, want, n):
    assert not np.isnan(got).any(), "entries never written: %d" % np.isnan(got).sum()
    from oracle import api
    M = api.unpack_symmetric(want, 3 * n)
    G = api.unpack_symmetric(got, 3 * n)
    dT = M[0, 0]
    dP = M[n, n]
    scale = np.empty((3 * n, 3 * n))
    s = np.array([dT] * n + [dP] * 2 * n)
    scale = np.sqrt(np.outer(s, s))       # TT block: dT, pol blocks: dP, cross: geometric mean
    assert (np.abs(G - M) / scale).max() <= REL_TOL


@pytest.mark.parametrize("nside,lmax,masked", [(1, 6, False), (2, 8, False), (4, 12, False), (4, 12, True), (8, 24, True)])
def test_tqu_matches_oracle(gpu_ctx, oracle_api, nside, lmax, masked):
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 10.0, good)
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good)
    _assert_tqu_close(got, want, n)


def test_tqu_matches_bruteforce_definition(gpu_ctx, oracle_api):
    from oracle import pol_bruteforce as bf
    nside, lmax = 1, 6
    spectra = synthetic_cl(lmax, pol=True)
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 0.0)
    ang = [oracle_api.pix2ang_nest(nside, i) for i in range(n)]
    C = bf.covariance(ang, *spectra, lmax)
    G = oracle_api.unpack_symmetric(got, 3 * n)
    assert np.abs(G - C).max() <= 1e-12 * C[0, 0]


def test_tqu_nside16_masked_sampled(gpu_ctx, oracle_api):
    """BASELINE config 2 (polarized Nside=16 lmax=47 with the test_like_low mask): all diagonal blocks and a
    random sample of pairs against the oracle."""
    nside, lmax = 16, 47
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 10.0, good)
    assert not np.isnan(got).any()
    rs = np.random.RandomState(7)
    pj = rs.randint(0, n, 20000)
    pi = (rs.uniform(size=20000) * (pj + 1)).astype(np.int64)
    pi = np.concatenate([pi, np.arange(n)])
    pj = np.concatenate([pj, np.arange(n)])
    blocks = oracle_api.tqu_pairs(*spectra, nside, 10.0, pi, pj, good=good)
    dT = blocks[-1, 0, 0]
    dP = blocks[-1, 1, 1]
    s = np.array([dT, dP, dP])
    scale = np.sqrt(np.outer(s, s))
    idx = lambda r, c: np.where(r <= c, c * (c + 1) // 2 + r, r * (r + 1) // 2 + c)
    worst = 0.0
    for a in range(3):
        for b in range(3):
            r = a * n + pi
            c = b * n + pj
            worst = max(worst, (np.abs(got[idx(r, c)] - blocks[:, a, b]) / scale[a, b]).max())
    assert worst <= REL_TOL
