"""The C++ drop-in classes (include/c_matrix.hpp, include/c_matrix_generator.hpp, include/utils.hpp), driven by a C++
program that mirrors the reference's own usage (source/test_like_low.cpp:181-186) and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, synthetic_cl

LIB = os.path.join(ROOT, "cosmopp_b200", "lib")


def write_healpix_mask_fits(path, mask, ordering):
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
                 card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", 1), card("TTYPE1", "MASK", True), card("TFORM1", "%dD" % rep, True),
                 card("PIXTYPE", "HEALPIX", True), card("ORDERING", ordering, True), card("NSIDE", int(round((n / 12) ** 0.5)))])
    data = np.asarray(mask, dtype=">f8").tobytes()
    data += b"\0" * ((2880 - len(data) % 2880) % 2880)
    with open(path, "wb") as f:
        f.write(primary + ext + data)


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
    oracle_api.good_pixels_from_mask(mask).astype("<i4").tofile(str(tmp_path / "mask_good.i32"))
    r = subprocess.run([dropin_binary, "cpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
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
