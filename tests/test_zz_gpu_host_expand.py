"""GPU: the opt-in whole-call mode that copies back only the last-face columns and fills in the rotated images on the host
(cmg_set_host_expand).  Kept in a file of its own, sorted last: the host half is pinned on the CPU
(tests/test_orbit_plan.py), the device half is nine plain copies, but this combination was written after the round's GPU time
was spent."""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nside,lmax", [(8, 16), (16, 30)])
def test_whole_calls_with_host_expansion_match_the_oracle(gpu_ctx, oracle_api, nside, lmax):
    import torch
    from cosmopp_b200 import capi
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    spectra = synthetic_cl(lmax, pol=True)
    plain = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, plain)
    gpu_ctx.set_host_expand(3)
    try:
        out = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64).pin_memory()
        gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out)
        tt = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64).pin_memory()
        gpu_ctx.cl_to_cmatrix(spectra[0], 10.0, tt)
    finally:
        gpu_ctx.set_host_expand(-1)
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0)
    scale = np.full(want.shape, want[capi.packed_index(n, n)])
    scale[:capi.packed_size(n)] = want[0]
    got = out.numpy()
    assert not np.isnan(got).any()
    assert (np.abs(got - want) / scale).max() <= 1e-11
    assert (np.abs(got - plain.numpy()) / scale).max() <= 1e-13         # images of an orbit are bit-identical on the device
    want_tt = oracle_api.cl_to_cmatrix(spectra[0], nside, 10.0)
    assert not np.isnan(tt.numpy()).any()
    assert np.abs(tt.numpy() - want_tt).max() <= 1e-11 * want_tt[0]
