"""The polarized oracles against each other.  PARITY UNPINNED BY THE REFERENCE (it has no TE/BB pixel generator and
its EE routine cannot run without HEALPix): the fast oracle (oracle/pol_oracle.c, Wigner-d recurrences + rotation
angles) is pinned by the definition-level sum over spin-weighted harmonics (oracle/pol_bruteforce.py)."""
import numpy as np
import pytest
from scipy.special import sph_harm_y

from conftest import synthetic_cl
from oracle import pol_bruteforce as bf


def test_spin_weighted_harmonics_definition():
    rs = np.random.RandomState(1)
    for _ in range(20):
        l = rs.randint(0, 9)
        m = rs.randint(-l, l + 1)
        th, ph = rs.uniform(0.05, np.pi - 0.05), rs.uniform(0, 2 * np.pi)
        assert abs(bf.sYlm(0, l, m, th, ph) - sph_harm_y(l, m, th, ph)) < 1e-13
        if l >= 2:
            # conj({s}Y_lm) = (-1)^(s+m) {-s}Y_l,-m
            a = np.conj(bf.sYlm(2, l, m, th, ph))
            b = (-1) ** (2 + m) * bf.sYlm(-2, l, -m, th, ph)
            assert abs(a - b) < 1e-13
    th, ph = 0.7, 0.3
    assert abs(bf.sYlm(2, 2, 2, th, ph) - 0.5 * np.sqrt(5 / np.pi) * np.sin(th / 2) ** 4 * np.exp(2j * ph)) < 1e-15
    assert abs(bf.sYlm(-2, 2, 2, th, ph) - 0.5 * np.sqrt(5 / np.pi) * np.cos(th / 2) ** 4 * np.exp(2j * ph)) < 1e-15


def test_bruteforce_identities_against_literal_mode_sum(oracle_api):
    lmax = 5
    spectra = synthetic_cl(lmax, pol=True)
    ang = [oracle_api.pix2ang_nest(1, i) for i in (0, 3, 5, 10)]
    assert np.abs(bf.covariance(ang, *spectra, lmax) - bf.covariance_from_modes(ang, *spectra, lmax)).max() < 1e-12 * spectra[0][2]


@pytest.mark.parametrize("nside,lmax,fwhm", [(1, 6, 0.0), (1, 9, 20.0), (2, 7, 10.0)])
def test_fast_oracle_matches_bruteforce(oracle_api, nside, lmax, fwhm):
    """Full sky: contains identical pixels (diagonal), antipodal pairs and same-meridian pairs."""
    spectra = synthetic_cl(lmax, pol=True)
    n = 12 * nside * nside
    f = oracle_api.window_beam(lmax, fwhm)
    ang = [oracle_api.pix2ang_nest(nside, i) for i in range(n)]
    C = bf.covariance(ang, *spectra, lmax, bT=f, bP=f)
    M = oracle_api.unpack_symmetric(oracle_api.tqu_matrix(*spectra, nside, fwhm), 3 * n)
    assert np.abs(M - C).max() <= 2e-14 * C[0, 0]
    assert np.linalg.eigvalsh(C).min() > 0


def test_fast_oracle_masked_and_pairs_interface(oracle_api):
    nside, lmax = 4, 10
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    n = len(good)
    M = oracle_api.unpack_symmetric(oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good), 3 * n)
    rs = np.random.RandomState(2)
    pj = rs.randint(0, n, 200)
    pi = (rs.uniform(size=200) * (pj + 1)).astype(int)
    blocks = oracle_api.tqu_pairs(*spectra, nside, 10.0, pi, pj, good=good)
    for a in range(3):
        for b in range(3):
            assert np.array_equal(blocks[:, a, b], M[a * n + pi, b * n + pj])
    # block structure on the diagonal: QQ_ii = UU_ii, QU_ii = 0, TQ_ii = TU_ii = 0
    d = np.arange(n)
    assert np.abs(M[n + d, n + d] - M[2 * n + d, 2 * n + d]).max() < 1e-15 * M[n, n]
    assert np.abs(M[n + d, 2 * n + d]).max() < 1e-15 * M[n, n] and np.abs(M[d, n + d]).max() < 1e-15 * M[0, 0]
