"""The C++ drop-in classes (include/c_matrix.hpp, include/c_matrix_generator.hpp, include/utils.hpp), driven by a C++
program that mirrors the reference's own usage (source/test_like_low.cpp:181-186) and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, synthetic_cl

LIB = os.path.join(ROOT, "cosmopp_b200", "lib")


def write_healpix_mask_fits(path, mask, ordering, form="D", tscal=None, tzero=None):
    """Minimal HEALPix-style FITS file: empty primary HDU + one BINTABLE with a 1024D column (like the reference's
    slow_test_files/mask1.fits)."""
    def card(key, val, quote=False):
        v = ("'%-8s'" % val) if quote else ("%20s" % val)
        return ("%-8s= %s" % (key, v)).ljust(80)
    def block(cards):
        s = "".join(cards) + "END".ljust(80)
        return s.ljust((len(s) + 2879) // 2880 * 2880).encode()
    n = len(mask)
    rep = 1024 if n % 1024 == 0 else (n if n < 1024 else 1)
    rows = n // rep
    primary = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    ext = block([card("XTENSION", "BINTABLE", True), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", 8 * rep), card("NAXIS2", rows),
                 card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", 1), card("TTYPE1", "MASK", True), card("TFORM1", "%d%s" % (rep, form), True)]
                + ([card("TSCAL1", tscal)] if tscal is not None else []) + ([card("TZERO1", tzero)] if tzero is not None else [])
                + [card("PIXTYPE", "HEALPIX", True), card("ORDERING", ordering, True), card("NSIDE", int(round((n / 12) ** 0.5)))])
    data = np.asarray(mask, dtype=">f8" if form == "D" else ">i8").tobytes()
    data += b"\0" * ((2880 - len(data) % 2880) % 2880)
    with open(path, "wb") as f:
        f.write(primary + ext + data)


def write_pixel_window_fits(path, w_t, w_p, form="E"):
    """HEALPix's pixel_window_nNNNN.fits as the reference reads it (source/utils.cpp:82-160): empty primary HDU, then a BINTABLE
    (HDU 2) with the columns TEMPERATURE and POLARIZATION, one row per l.  HEALPix ships them as 1E (float32); 1D is tested too."""
    def card(key, val, quote=False):
        v = ("'%-8s'" % val) if quote else ("%20s" % val)
        return ("%-8s= %s" % (key, v)).ljust(80)
    def block(cards):
        s = "".join(cards) + "END".ljust(80)
        return s.ljust((len(s) + 2879) // 2880 * 2880).encode()
    n = len(w_t)
    width = 4 if form == "E" else 8
    primary = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    ext = block([card("XTENSION", "BINTABLE", True), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", 2 * width), card("NAXIS2", n),
                 card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", 2), card("TTYPE1", "TEMPERATURE", True), card("TFORM1", "1" + form, True),
                 card("TTYPE2", "POLARIZATION", True), card("TFORM2", "1" + form, True)])
    rows = np.empty((n, 2), dtype=">f4" if form == "E" else ">f8")
    rows[:, 0] = w_t
    rows[:, 1] = w_p
    data = rows.tobytes()
    data += b"\0" * ((2880 - len(data) % 2880) % 2880)
    with open(path, "wb") as f:
        f.write(primary + ext + data)


def window_tables(lmax, nside):
    l = np.arange(lmax + 1, dtype=np.float64)
    s = np.sqrt(4 * np.pi / (12.0 * nside * nside)) / 2.2
    w_t = np.exp(-0.5 * l * (l + 1) * s * s)
    w_p = w_t * (1.0 - 0.35 * (l / (4.0 * nside)) ** 2)
    w_p[:2] = 0.0
    return w_t, w_p


@pytest.fixture(scope="module")
def dropin_binary(tmp_path_factory):
    if not os.path.exists(os.path.join(LIB, "libcosmopp_b200.so")):
        pytest.skip("library not built")
    out = str(tmp_path_factory.mktemp("bin") / "test_dropin")
    cmd = ["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
           "-L" + LIB, "-lcosmopp_b200", "-Wl,-rpath," + LIB, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def read_cmatrix(path):
    with open(path, "rb") as f:
        (n,) = struct.unpack("<i", f.read(4))
        m = np.frombuffer(f.read(8 * n * (n + 1) // 2), dtype="<f8")
        (ln,) = struct.unpack("<i", f.read(4))
        return n, m, f.read(ln).decode()


def test_cpp_dropin_host_side(dropin_binary, tmp_path, oracle_api):
    mask = oracle_api.like_low_mask(4)
    write_healpix_mask_fits(str(tmp_path / "mask_nest.fits"), mask, "NESTED")
    write_healpix_mask_fits(str(tmp_path / "mask_ring.fits"), mask, "RING")
    # 64-bit integer column, and a scaled one (stored 2 m - 1 with TSCAL = 0.5, TZERO = 0.5): the same mask both times
    write_healpix_mask_fits(str(tmp_path / "mask_k.fits"), np.asarray(mask, dtype=np.int64), "NESTED", form="K")
    write_healpix_mask_fits(str(tmp_path / "mask_scaled.fits"), 2 * np.asarray(mask, dtype=np.int64) - 1, "NESTED", form="K", tscal=0.5, tzero=0.5)
    oracle_api.good_pixels_from_mask(mask).astype("<i4").tofile(str(tmp_path / "mask_good.i32"))
    # a7: the HEALPix pixel window file, single precision as HEALPix ships it (64 entries: up to l = 4 nside - 1)
    w_t, w_p = window_tables(63, 16)
    write_pixel_window_fits(str(tmp_path / "pixel_window_n0016.fits"), w_t, w_p, "E")
    r = subprocess.run([dropin_binary, "cpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    got = np.loadtxt(str(tmp_path / "window_read.txt"))
    l = np.arange(48)
    beam = np.array([oracle_api.window_beam(47, 10.0)[k] for k in l])          # exp(-l(l+1) / (2 sigma^2)), source/utils.cpp:54-64
    # the reference reads the cells as TEXT and parses them again (source/utils.cpp:139-160); cfitsio prints a float32 cell with its
    # default display format %#14.6G, i.e. SIX significant digits -- part of what the reference computes with
    as_read_t = np.array([float("%#14.6G" % np.float32(v)) for v in w_t[:48]])
    as_read_p = np.array([float("%#14.6G" % np.float32(v)) for v in w_p[:48]])
    assert np.abs(got[:, 0] - as_read_t * beam).max() <= 1e-15 and np.abs(got[:, 1] - as_read_p * beam).max() <= 1e-15
    assert np.abs(got[:, 0] / (w_t[:48] * beam) - 1).max() < 6e-6            # i.e. the table itself to those six digits
    assert np.abs(got[:21, 2] - as_read_t[:21]).max() <= 1e-15               # fwhm = 0: the window alone
    assert got[0, 1] == 0.0 and got[1, 1] == 0.0                              # the polarization table has no monopole / dipole
    # the binary layout is the reference's: int32 nPix, packed doubles, int32 length, comment
    n, m, comment = read_cmatrix(str(tmp_path / "m.dat"))
    assert n == 5 and comment == "hello matrix" and m[4 * 5 // 2 + 2] == 42.25
    lines = open(str(tmp_path / "m.txt")).read().splitlines()
    assert lines[0] == "5" and lines[1] == "hello matrix" and lines[2] == "0\t0\t0.25" and lines[4] == "1\t1\t11.25"
    if oracle_api.have_ref():
        # byte-for-byte against the reference's own writer
        packed = np.array([10 * j + i + 0.25 for j in range(5) for i in range(j + 1)])
        oracle_api.ref_write_cmatrix(packed, 5, "hello matrix", str(tmp_path / "ref.dat"), str(tmp_path / "ref.txt"))
        assert open(str(tmp_path / "ref.dat"), "rb").read() == open(str(tmp_path / "m.dat"), "rb").read()
        assert open(str(tmp_path / "ref.txt")).read() == open(str(tmp_path / "m.txt")).read()


@pytest.mark.gpu
def test_cpp_dropin_generators_match_oracle(dropin_binary, tmp_path, oracle_api):
    nside, lmax = 8, 20
    tt, te, ee, bb = synthetic_cl(4 * nside, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    tt.astype("<f8").tofile(str(tmp_path / "cl_tt.f64"))
    for name, a in (("te", te), ("ee", ee), ("bb", bb)):
        a[:lmax + 1].astype("<f8").tofile(str(tmp_path / ("cl_%s.f64" % name)))
    good.astype("<i4").tofile(str(tmp_path / "good.i32"))
    short = synthetic_cl(12)
    with open(str(tmp_path / "cl_short.txt"), "w") as f:
        f.write("".join("%.17g\n" % v for v in short))
    w_t, w_p = window_tables(4 * nside, nside)
    w_t.astype("<f8").tofile(str(tmp_path / "win_t.f64"))
    w_p.astype("<f8").tofile(str(tmp_path / "win_p.f64"))
    # LikelihoodPolarization inputs (Nside = 4, [Q;U] over all 192 pixels restricted to a ragged set of unmasked ones)
    rng_p = np.random.default_rng(21)
    good_p = np.sort(rng_p.choice(192, size=150, replace=False)).astype("<i4")
    good_p.tofile(str(tmp_path / "good_p.i32"))
    ninv_diag = rng_p.uniform(2.0, 5.0, 2 * 192)
    ninv_diag.astype("<f8").tofile(str(tmp_path / "ninv_diag.f64"))
    v_p = rng_p.normal(size=300)
    pred_p = rng_p.normal(size=300) * 0.1
    v_p.astype("<f8").tofile(str(tmp_path / "v_p.f64"))
    pred_p.astype("<f8").tofile(str(tmp_path / "pred_p.f64"))
    rng = np.random.default_rng(7)
    maps = rng.normal(size=(6, len(good))) * 30.0
    fore = rng.normal(size=len(good)) * 5.0 + 20.0
    maps.astype("<f8").tofile(str(tmp_path / "maps.f64"))
    fore.astype("<f8").tofile(str(tmp_path / "fore.f64"))
    r = subprocess.run([dropin_binary, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]

    # the Likelihood class against a numpy restatement of reference source/likelihood.cpp:100-180 on the oracle's matrices
    ng = len(good)
    S = (oracle_api.unpack_symmetric(oracle_api.cl_to_cmatrix(tt[:lmax + 1], nside, 10.0, good=good), ng)
         + oracle_api.unpack_symmetric(oracle_api.fiducial_matrix(tt, nside, lmax, 10.0, good=good), ng)
         + oracle_api.unpack_symmetric(oracle_api.mask_matrix(oracle_api.noise_matrix(nside, 1e-2), good), ng))
    Sinv = np.linalg.inv(S)
    logdet = np.linalg.slogdet(S)[1] + 29677.0566
    got = np.loadtxt(str(tmp_path / "like.txt"))
    fCf = fore @ Sinv @ fore
    for k in range(maps.shape[0]):
        chi2 = maps[k] @ Sinv @ maps[k]
        assert abs(got[k, 0] - chi2) <= 1e-9 * chi2 and abs(got[k, 1] - logdet) <= 1e-9 * abs(logdet)
        tCf = maps[k] @ Sinv @ fore
        assert abs(got[k, 2] - (chi2 - tCf * tCf / fCf)) <= 1e-9 * chi2
        assert abs(got[k, 3] - (logdet + np.log(fCf / ng))) <= 1e-9 * abs(logdet)

    # the sampler plug-in's batch: C_l = A * base * ((l+1)/10)^tilt through the oracle's generator and numpy
    F = oracle_api.unpack_symmetric(oracle_api.fiducial_matrix(tt, nside, lmax, 10.0, good=good), ng)
    N = oracle_api.unpack_symmetric(oracle_api.mask_matrix(oracle_api.noise_matrix(nside, 1e-2), good), ng)
    for amp, tilt, got_like in np.loadtxt(str(tmp_path / "plug.txt")):
        cl_k = amp * tt[:lmax + 1] * ((np.arange(lmax + 1) + 1.0) / 10.0) ** tilt
        Sk = oracle_api.unpack_symmetric(oracle_api.cl_to_cmatrix(cl_k, nside, 10.0, good=good), ng) + F + N
        want_like = maps[0] @ np.linalg.solve(Sk, maps[0]) + np.linalg.slogdet(Sk)[1] + 29677.0566
        assert abs(got_like - want_like) <= 1e-9 * abs(want_like)

    n, c, _ = read_cmatrix(str(tmp_path / "c.dat"))
    want = oracle_api.cl_to_cmatrix(tt[:lmax + 1], nside, 10.0, good=good)
    assert n == len(good) and np.abs(c - want).max() <= 1e-11 * want[0]
    n, c, comment = read_cmatrix(str(tmp_path / "c_fiducial.dat"))
    want = oracle_api.fiducial_matrix(tt, nside, lmax, 10.0, good=good)
    assert comment == "fiducial matrix" and np.abs(c - want).max() <= 1e-11 * want[0]
    n, c, comment = read_cmatrix(str(tmp_path / "c_noise.dat"))
    assert comment == "noise matrix" and np.array_equal(c, oracle_api.mask_matrix(oracle_api.noise_matrix(nside, 1e-2), good))
    n, c, _ = read_cmatrix(str(tmp_path / "c_full.dat"))
    want = oracle_api.cl_to_cmatrix(short, 4, 10.0)
    assert n == 192 and np.abs(c - want).max() <= 1e-11 * want[0]
    n, c, _ = read_cmatrix(str(tmp_path / "c_pol.dat"))
    want = oracle_api.tqu_matrix(tt[:lmax + 1], te[:lmax + 1], ee[:lmax + 1], bb[:lmax + 1], nside, 10.0, good=good)
    ng = len(good)
    assert n == 3 * ng
    M, G = oracle_api.unpack_symmetric(want, n), oracle_api.unpack_symmetric(c, n)
    s = np.array([M[0, 0]] * ng + [M[ng, ng]] * 2 * ng)
    assert (np.abs(G - M) / np.sqrt(np.outer(s, s))).max() <= 1e-11

    # the same likelihood evaluated with every matrix resident on the GPU (before anything was copied to the host)
    res = np.loadtxt(str(tmp_path / "like_resident.txt"))
    assert abs(res[0] - maps[0] @ Sinv @ maps[0]) <= 1e-9 * res[0] and abs(res[1] - logdet) <= 1e-9 * abs(logdet)

    # polarized matrix with different temperature / polarization windows, and the reference's temperature-only choice
    for name, wp in (("c_pol_wtp.dat", w_p), ("c_pol_wtt.dat", w_t)):
        n, c, _ = read_cmatrix(str(tmp_path / name))
        want = oracle_api.tqu_matrix(tt[:lmax + 1], te[:lmax + 1], ee[:lmax + 1], bb[:lmax + 1], nside, 10.0, good=good,
                                     pixwinT=w_t[:lmax + 1], pixwinP=wp[:lmax + 1])
        M, G = oracle_api.unpack_symmetric(want, n), oracle_api.unpack_symmetric(c, n)
        s = np.array([M[0, 0]] * ng + [M[ng, ng]] * 2 * ng)
        assert (np.abs(G - M) / np.sqrt(np.outer(s, s))).max() <= 1e-11

    # LikelihoodPolarization against numpy on the oracle's [Q;U] block (reference source/likelihood.cpp:391-397, 592-606)
    n, c_qu, comment = read_cmatrix(str(tmp_path / "c_qu.dat"))
    assert n == 2 * 192 and comment == "QU covariance matrix"
    full = oracle_api.unpack_symmetric(oracle_api.tqu_matrix(tt[:13], te[:13], ee[:13], bb[:13], 4, 10.0), 3 * 192)
    QU = full[192:, 192:]
    assert np.abs(oracle_api.unpack_symmetric(c_qu, n) - QU).max() <= 1e-11 * QU[0, 0]
    idx = np.concatenate([good_p, 192 + good_p]).astype(int)
    Ninv = np.diag(ninv_diag)
    Ninv[3, 200] = Ninv[200, 3] = 0.01
    Cg, Ng = QU[np.ix_(idx, idx)], Ninv[np.ix_(idx, idx)]
    K = Ng + Ng @ Cg @ Ng
    chi2_a = v_p @ np.linalg.solve(K, v_p)
    vb = v_p - Ng @ pred_p
    chi2_b = vb @ np.linalg.solve(K, vb)
    want_logdet = np.linalg.slogdet(K)[1] - 16078.083180
    got_pol = np.loadtxt(str(tmp_path / "like_pol.txt"))
    assert abs(got_pol[0] - chi2_a) <= 1e-9 * chi2_a and abs(got_pol[1] - chi2_b) <= 1e-9 * chi2_b
    assert abs(got_pol[2] - want_logdet) <= 1e-9 * abs(want_logdet)
