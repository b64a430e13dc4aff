"""Pixel centres: chealpix is not available, and the reference has no direct test of it, so both implementations
(the oracle's C and the product's C++) are pinned by HEALPix invariants (Gorski et al. 2005) and against each other."""
import os

import numpy as np
import pytest

from conftest import ROOT


def _vec(theta, phi):
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def test_nside1_closed_form(oracle_api):
    # faces 0-3: z = 2/3, phi = pi/4 + k pi/2; faces 4-7: z = 0, phi = k pi/2; faces 8-11: z = -2/3
    for ipix in range(12):
        theta, phi = oracle_api.pix2ang_nest(1, ipix)
        face, k = divmod(ipix, 4)
        assert abs(np.cos(theta) - (2 / 3, 0.0, -2 / 3)[face]) < 1e-15
        want = (np.pi / 4 + k * np.pi / 2) if face != 1 else (k * np.pi / 2)
        assert abs(phi - want) < 1e-15


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 64])
def test_ring_structure_and_bijection(oracle_api, nside):
    npix = 12 * nside * nside
    ang = np.array([oracle_api.pix2ang_nest(nside, i) for i in range(npix)])
    z = np.cos(ang[:, 0])
    rings, counts = np.unique(np.round(z, 12), return_counts=True)
    assert len(rings) == 4 * nside - 1
    # ring r (from the north) holds 4 min(r, nside, 4nside - r) pixels
    want = [4 * min(r, nside, 4 * nside - r) for r in range(1, 4 * nside)]
    assert list(counts[::-1]) == want
    v = np.stack([np.sin(ang[:, 0]) * np.cos(ang[:, 1]), np.sin(ang[:, 0]) * np.sin(ang[:, 1]), z], 1)
    assert np.abs(v.sum(0)).max() < 1e-9
    assert (ang[:, 1] >= 0).all() and (ang[:, 1] < 2 * np.pi).all()
    # NEST and RING enumerate the same set of centres
    ring = np.array([oracle_api.pix2ang_ring(nside, i) for i in range(npix)])
    key = lambda a: sorted((round(t, 11), round(p, 11)) for t, p in a)
    assert key(ang) == key(ring)


def test_nested_hierarchy(oracle_api):
    """The four children (nside 2N) of a NESTED pixel surround its centre: their mean direction is the parent's."""
    for nside in (1, 2, 8):
        for ipix in range(0, 12 * nside * nside, 5):
            parent = _vec(*oracle_api.pix2ang_nest(nside, ipix))
            kids = np.mean([_vec(*oracle_api.pix2ang_nest(2 * nside, 4 * ipix + k)) for k in range(4)], axis=0)
            assert np.dot(parent, kids / np.linalg.norm(kids)) > 1 - 0.2 / (nside * nside)


def test_product_pix2ang_is_bit_identical_to_oracle(oracle_api):
    from cosmopp_b200 import capi
    if not os.path.exists(capi.library_path()):
        pytest.skip("library not built")
    for nside in (1, 2, 4, 8, 32, 64):
        npix = 12 * nside * nside
        step = max(1, npix // 3000)
        for ipix in list(range(0, npix, step)) + [npix - 1]:
            assert capi.pix2ang_nest(nside, ipix) == oracle_api.pix2ang_nest(nside, ipix)
    with pytest.raises(capi.CmgError):
        capi.pix2ang_nest(12, 0)
    with pytest.raises(capi.CmgError):
        capi.pix2ang_nest(4, 192)
