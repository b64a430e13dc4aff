"""Pixel-window path on the GPU with windows that are NOT identically one, and the direct comparison with
the reference's own object code at BASELINE configs[0] in full.

The reference multiplies the Gaussian beam by the HEALPix pixel window read from pixel_window_nNNNN.fits
(source/utils.cpp:66-170: f[l] = pixwin[l] * beam(l)) and squares the product in the generator
(source/c_matrix_generator.cpp:222).  The HEALPix data files are not available offline, so the window here is a synthetic table
with the shape of a real one (1 at l = 0, falling smoothly); what is tested is that the table reaches every entry of the matrix
the way the reference applies it -- temperature table for TT, temperature x polarization for TE, polarization for EE / BB."""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu

REL_TOL = 1e-11


def window_tables(lmax, nside):
    """shaped like HEALPix's: w_T ~ exp(-l(l+1) s^2 / 2) with s of the order of the pixel size; the polarization window differs"""
    l = np.arange(lmax + 1, dtype=np.float64)
    s = np.sqrt(4 * np.pi / (12.0 * nside * nside)) / 2.2
    w_t = np.exp(-0.5 * l * (l + 1) * s * s)
    w_p = w_t * (1.0 - 0.35 * (l / (4.0 * nside)) ** 2)
    w_p[:2] = 0.0                                   # HEALPix stores 0 for the polarization window below l = 2
    return w_t, w_p


@pytest.mark.skipif("not __import__('oracle.api').api.have_ref()")
def test_tt_config1_full_size_against_reference_object_code_with_a_window(gpu_ctx, oracle_api):
    """BASELINE configs[0] (Nside = 16, lmax = 47, full sky: 4.7 M pixel pairs) entry by entry against the reference's
    clToCMatrix compiled from its own sources (about 20 s of CPU), with a non-unit pixel window injected on both sides."""
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 16, 47
    cl = synthetic_cl(lmax)
    w_t, _ = window_tables(lmax, nside)
    oracle_api.ref_set_pixel_window(w_t, None)
    try:
        want = oracle_api.ref_cl_to_cmatrix(cl, nside, 10.0)
    finally:
        oracle_api.ref_set_pixel_window(None, None)
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    out = torch.empty(capi.packed_size(gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix(cl, 10.0, out, pixwin=w_t)           # whole call: the orbit path on the full sky
    got = out.numpy()
    assert got.shape == want.shape and np.abs(got - want).max() <= REL_TOL * want[0]
    # the window matters at this tolerance: without it the diagonal differs by several per cent
    gpu_ctx.cl_to_cmatrix(cl, 10.0, out)
    assert abs(out.numpy()[0] - want[0]) > 1e-3 * want[0]
    # the C restatement agrees with both
    assert np.abs(oracle_api.cl_to_cmatrix(cl, nside, 10.0, pixwin=w_t) - want).max() <= 1e-13 * want[0]


@pytest.mark.skipif("not __import__('oracle.api').api.have_ref()")
def test_fiducial_matrix_with_a_window_against_reference_object_code(gpu_ctx, oracle_api):
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 8, 20
    cl = synthetic_cl(4 * nside)
    w_t, _ = window_tables(4 * nside, nside)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    oracle_api.ref_set_pixel_window(w_t, None)
    try:
        want = oracle_api.ref_fiducial_matrix(cl, nside, lmax, 10.0, good=good)
    finally:
        oracle_api.ref_set_pixel_window(None, None)
    gpu_ctx.set_pixels(nside, good)
    out = torch.empty(capi.packed_size(gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    gpu_ctx.fiducial_matrix(cl, lmax, 10.0, out, pixwin=w_t)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]


@pytest.mark.parametrize("masked", [True, False])
def test_tqu_temperature_and_polarization_windows(gpu_ctx, oracle_api, masked):
    """T,Q,U at BASELINE configs[1] (Nside = 16, lmax = 47, the test_like_low mask) with DIFFERENT temperature and polarization
    windows: TT carries w_T^2, TE w_T w_P, EE and BB w_P^2 (cmg_tqu_weights).  The full-sky case runs the orbit path."""
    import torch
    from cosmopp_b200 import capi
    nside, lmax = 16, 47
    spectra = synthetic_cl(lmax, pol=True)
    w_t, w_p = window_tables(lmax, nside)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out, pixwinT=w_t, pixwinP=w_p)
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good, pixwinT=w_t, pixwinP=w_p)
    scale = np.full(want.shape, want[capi.packed_index(n, n)])
    scale[:capi.packed_size(n)] = want[0]
    assert (np.abs(out.numpy() - want) / scale).max() <= REL_TOL
    # what the reference's own polarization routine does (source/c_matrix_generator.cpp:534: the TEMPERATURE table for the
    # polarization beam as well) is the same call with pixwinP = pixwinT -- and it is a different matrix
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out, pixwinT=w_t, pixwinP=w_t)
    want_ref_style = oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good, pixwinT=w_t, pixwinP=w_t)
    assert (np.abs(out.numpy() - want_ref_style) / scale).max() <= REL_TOL
    assert (np.abs(want_ref_style - want) / scale).max() > 1e-4


def test_negative_fwhm_and_short_windows_are_errors(gpu_ctx):
    """the reference has check(fwhm >= 0); an all-zero matrix with status OK would be a silent wrong answer"""
    import torch
    from cosmopp_b200 import capi
    gpu_ctx.set_pixels(4)
    cl = synthetic_cl(12)
    out = torch.empty(capi.packed_size(gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    with pytest.raises(capi.CmgError):
        gpu_ctx.cl_to_cmatrix(cl, -1.0, out)
    with pytest.raises(capi.CmgError):
        gpu_ctx.cl_to_cmatrix(cl, 10.0, out, pixwin=np.ones(5))                # shorter than cl
    with pytest.raises(capi.CmgError):
        gpu_ctx.cl_to_cmatrix(cl, 10.0, out[:10])                              # output too small
    spectra = synthetic_cl(12, pol=True)
    outp = torch.empty(capi.packed_size(3 * gpu_ctx.npix), dtype=torch.float64, pin_memory=True)
    with pytest.raises(capi.CmgError):
        gpu_ctx.cl_to_cmatrix_pol(spectra[0], spectra[1][:-1], spectra[2], spectra[3], 10.0, outp)
    with pytest.raises(capi.CmgError):
        gpu_ctx.cl_to_cmatrix_pol(*spectra, -0.5, outp)
