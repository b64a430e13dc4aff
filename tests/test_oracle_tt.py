"""The CPU oracle of the TT path against what pins it: the reference's Legendre known answers, committed vectors
produced by the reference's own object code, and (in the build container) that object code live."""
import os

import numpy as np
import pytest

from conftest import ROOT, synthetic_cl

GOLD = os.path.join(ROOT, "tests", "golden")

# reference source/test_legendre.cpp:27-51 -- (l, x, expected); the harness compares to relative 1e-5
LEGENDRE_KAT = [(4, 2.0, 55.375), (10, 0.5, -0.188228607177734375), (64, 0.1, 0.098026402863),
                (1000, -0.25, 0.0023444296560), (10000, -0.5, -0.006062503808317)]


@pytest.mark.parametrize("l,x,expected", LEGENDRE_KAT)
def test_legendre_known_answers(oracle_api, l, x, expected):
    got = oracle_api.lib().orc_legendre(l, x)
    assert abs(got - expected) <= 1e-9 * abs(expected)          # printed digits of the reference test


def test_legendre_matches_reference_object_code_vectors(oracle_api):
    g = np.load(os.path.join(GOLD, "ref_legendre.npz"))
    for a, l in enumerate(g["l"]):
        for b, x in enumerate(g["x"]):
            got = oracle_api.lib().orc_legendre(int(l), float(x))
            want = g["value"][a, b]
            if not np.isfinite(want):          # P_l(2) overflows for large l, in the reference too
                assert not np.isfinite(got)
                continue
            assert abs(got - want) <= 4e-15 * max(1.0, abs(want)) * max(1, int(l)), (l, x)


def test_beam_function(oracle_api):
    # reference source/utils.cpp:54-64
    assert oracle_api.lib().orc_beam_function(17, 0.0) == 1.0
    sigma = np.sqrt(8 * np.log(2.0)) / (10.0 * 3.141592653589793 / 180)
    for l in (0, 2, 30, 64, 192):
        assert abs(oracle_api.lib().orc_beam_function(l, 10.0) - np.exp(-l * (l + 1) / (2 * sigma * sigma))) < 1e-16


def test_tt_full_sky_matches_reference_vectors(oracle_api):
    g = np.load(os.path.join(GOLD, "ref_tt_nside4.npz"))
    got = oracle_api.cl_to_cmatrix(g["cl"], int(g["nside"]), float(g["fwhm"]))
    assert got.shape == g["packed"].shape
    assert np.abs(got - g["packed"]).max() <= 1e-14 * g["packed"][0]


def test_tt_masked_matches_reference_vectors(oracle_api):
    g = np.load(os.path.join(GOLD, "ref_tt_nside8_masked.npz"))
    got = oracle_api.cl_to_cmatrix(g["cl"], int(g["nside"]), float(g["fwhm"]), good=g["good"])
    assert np.abs(got - g["packed"]).max() <= 1e-14 * g["packed"][0]
    # the literal O(lmax^2) form of the reference loop gives the same numbers
    sub = oracle_api.cl_to_cmatrix(g["cl"], int(g["nside"]), float(g["fwhm"]), good=g["good"], cols=(40, 60), literal=True)
    lo, hi = 40 * 41 // 2, 60 * 61 // 2
    assert np.abs(sub - g["packed"][lo:hi]).max() <= 1e-14 * g["packed"][0]


def test_fiducial_matches_reference_vectors(oracle_api):
    g = np.load(os.path.join(GOLD, "ref_fiducial_nside4.npz"))
    got = oracle_api.fiducial_matrix(g["cl"], int(g["nside"]), int(g["lmax"]), float(g["fwhm"]))
    assert np.abs(got - g["packed"]).max() <= 1e-14 * g["packed"][0]


def test_noise_and_mask_match_reference_vectors(oracle_api):
    g = np.load(os.path.join(GOLD, "ref_noise_masked.npz"))
    full = oracle_api.noise_matrix(int(g["nside"]), float(g["noise"]))
    assert np.array_equal(oracle_api.mask_matrix(full, g["good"]), g["packed"])


def test_structural_invariants(oracle_api):
    """S_ii = sum w_l (P_l(1) = 1); antipodal pairs give sum (-1)^l w_l."""
    nside, lmax = 2, 9
    cl = synthetic_cl(lmax)
    f = oracle_api.window_beam(lmax, 10.0)
    l = np.arange(lmax + 1)
    w = cl * (2 * l + 1) / (4 * 3.141592653589793) * f * f
    M = oracle_api.unpack_symmetric(oracle_api.cl_to_cmatrix(cl, nside, 10.0), 48)
    assert np.abs(np.diag(M) - w[2:].sum()).max() < 1e-12 * w[2:].sum()
    v = oracle_api.unit_vectors(nside)
    anti = np.argmin(v @ v.T, axis=1)
    assert np.allclose((v * v[anti]).sum(1), -1.0, atol=1e-14)
    want = ((-1.0) ** l[2:] * w[2:]).sum()
    assert np.abs(M[np.arange(48), anti] - want).max() < 1e-12 * w[2:].sum()


@pytest.mark.skipif("not __import__('oracle.api').api.have_ref()")
class TestAgainstLiveReferenceObjects:
    """Only where oracle/_ref was built (the reference's own c_matrix.cpp / c_matrix_generator.cpp)."""

    def test_tt(self, oracle_api):
        good = np.load(os.path.join(GOLD, "like_low_good_pixels_nside4.npy"))
        cl = synthetic_cl(14, seed=3)
        for g in (None, good):
            a = oracle_api.cl_to_cmatrix(cl, 4, 7.5, good=g)
            b = oracle_api.ref_cl_to_cmatrix(cl, 4, 7.5, good=g)
            assert np.abs(a - b).max() <= 1e-14 * b[0]
        # through the reference's LegendrePolynomialContainer: same values
        c = oracle_api.ref_cl_to_cmatrix(cl, 4, 7.5, good=good, use_lp=True)
        assert np.abs(c - b).max() <= 1e-14 * b[0]

    def test_packed_index_layout(self, oracle_api):
        n = 37
        for i, j in [(0, 0), (0, 36), (36, 0), (5, 9), (9, 5), (36, 36)]:
            assert oracle_api.lib().orc_packed_index(i, j) == oracle_api.ref().ref_packed_index(n, i, j)

    def test_mask_matrix(self, oracle_api):
        rs = np.random.RandomState(0)
        n = 23
        packed = rs.standard_normal(n * (n + 1) // 2)
        good = np.array([0, 2, 3, 11, 22], dtype=np.int32)
        assert np.array_equal(oracle_api.mask_matrix(packed, good), oracle_api.ref_mask_matrix(packed, n, good))

    def test_legendre_container_file(self, oracle_api, tmp_path):
        good = np.array([1, 5, 6, 20, 40], dtype=np.int32)
        path = str(tmp_path / "lp.dat")
        oracle_api.ref_write_legendre_container(5, 2, good, path)
        raw = open(path, "rb").read()
        lmax, npix = np.frombuffer(raw[:8], dtype="<i4")
        data = np.frombuffer(raw[8:], dtype="<f8").reshape(lmax + 1, npix * (npix + 1) // 2)
        assert (lmax, npix) == (5, 5)
        assert np.abs(data - oracle_api.legendre_container(5, 2, good)).max() < 1e-15
