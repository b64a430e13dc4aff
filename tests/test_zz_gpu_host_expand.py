"""GPU: the whole-call mode that copies back only the last-face columns (plus the images selected by
cmg_set_host_expand_direct) and fills in the other rotated images on the host (cmg_set_host_expand; automatic for matrices of
1 GiB and more, forced here at small sizes).  The host half alone is pinned on the CPU (tests/test_orbit_plan.py)."""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nside,lmax,direct", [(8, 16, 0x40), (16, 30, 0x40), (16, 30, 0), (8, 16, 0x1ff), (16, 20, 0x0b2)])
def test_whole_calls_with_host_expansion_match_the_oracle(gpu_ctx, oracle_api, nside, lmax, direct):
    import torch
    from cosmopp_b200 import capi
    gpu_ctx.set_kernel_variant(0)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    spectra = synthetic_cl(lmax, pol=True)
    plain = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, plain)
    gpu_ctx.set_host_expand(3)
    gpu_ctx.set_host_expand_direct(direct)
    d2h0 = gpu_ctx.transfer_counters()[1]
    try:
        out = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64).pin_memory()
        gpu_ctx.cl_to_cmatrix_pol(*spectra, 10.0, out)
        moved = gpu_ctx.transfer_counters()[1] - d2h0
        tt = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64).pin_memory()
        gpu_ctx.cl_to_cmatrix(spectra[0], 10.0, tt)
    finally:
        gpu_ctx.set_host_expand(-1)
        gpu_ctx.set_host_expand_direct(0)
    # what crossed PCIe: the last face of every ring and the direct images, nothing else
    F = nside * nside
    expect = 8 * sum(capi.packed_size(s * n + (fc + 1) * F) - capi.packed_size(s * n + fc * F) for s in range(3) for fc in range(12)
                     if (fc & 3) == 3 or ((direct >> (3 * s + (3 - (fc & 3)) - 1)) & 1))
    assert moved == expect and (direct == 0x1ff) == (moved == 8 * capi.packed_size(3 * n))
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
