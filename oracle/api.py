"""TEST INFRASTRUCTURE ONLY: ctypes front-end for oracle/liboracle.so (the CPU restatement) and
oracle/_ref/libcosmopp_ref.so (the reference's own object code, TT path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under cosmopp_b200/ may.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = ctypes.c_void_p
_LL = ctypes.c_longlong


def build(verbose=False):
    """Compile liboracle.so and, when /root/reference is present, _ref/libcosmopp_ref.so."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("oracle build failed")


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_legendre.restype = ctypes.c_double
        L.orc_legendre.argtypes = [ctypes.c_uint, ctypes.c_double]
        L.orc_beam_function.restype = ctypes.c_double
        L.orc_beam_function.argtypes = [ctypes.c_int, ctypes.c_double]
        L.orc_packed_index.restype = _LL
        L.orc_packed_index.argtypes = [_LL, _LL]
        L.orc_good_pixels_from_mask.restype = ctypes.c_long
        L.orc_like_low_mask.restype = ctypes.c_long
        L.nside2npix.restype = ctypes.c_long
        L.nside2npix.argtypes = [ctypes.c_long]
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libcosmopp_ref.so"))


def ref():
    """The reference's own object code (TT path); raises if it was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libcosmopp_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libcosmopp_ref.so missing: run `make -C oracle` where /root/reference exists")
        R = ctypes.CDLL(path)
        R.ref_legendre.restype = ctypes.c_double
        R.ref_legendre.argtypes = [ctypes.c_uint, ctypes.c_double]
        R.ref_beam_function.restype = ctypes.c_double
        R.ref_beam_function.argtypes = [ctypes.c_int, ctypes.c_double]
        R.ref_last_error.restype = ctypes.c_char_p
        R.ref_packed_index.restype = ctypes.c_long
        R.ref_packed_index.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _ref = R
    return _ref


# ------------------------------------------------------------------ helpers over liboracle.so

def packed_size(n):
    return n * (n + 1) // 2


def pix2ang_nest(nside, ipix):
    t = ctypes.c_double()
    p = ctypes.c_double()
    lib().pix2ang_nest(ctypes.c_long(nside), ctypes.c_long(int(ipix)), ctypes.byref(t), ctypes.byref(p))
    return t.value, p.value


def pix2ang_ring(nside, ipix):
    t = ctypes.c_double()
    p = ctypes.c_double()
    lib().pix2ang_ring(ctypes.c_long(nside), ctypes.c_long(int(ipix)), ctypes.byref(t), ctypes.byref(p))
    return t.value, p.value


def unit_vectors(nside, good=None):
    good = _i32(good)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty((n, 3))
    lib().orc_unit_vectors(ctypes.c_long(nside), _ptr(good), ctypes.c_long(n), _ptr(out))
    return out


def window_beam(lmax, fwhm, pixwin=None):
    f = np.empty(lmax + 1)
    pixwin = _f64(pixwin)
    lib().orc_window_beam(_ptr(f), ctypes.c_int(lmax), ctypes.c_double(fwhm), _ptr(pixwin))
    return f


def good_pixels_from_mask(mask):
    mask = _f64(mask)
    good = np.empty(len(mask), dtype=np.int32)
    n = lib().orc_good_pixels_from_mask(_ptr(mask), ctypes.c_long(len(mask)), _ptr(good))
    return good[:n].copy()


def like_low_mask(nside):
    mask = np.empty(12 * nside * nside)
    lib().orc_like_low_mask(ctypes.c_long(nside), _ptr(mask))
    return mask


def like_low_discs(n=25, seed=1000000):
    d = np.empty((n, 3))
    lib().orc_like_low_discs(ctypes.c_int(n), ctypes.c_uint32(seed), _ptr(d))
    return d


def cl_to_cmatrix(cl, nside, fwhm, good=None, pixwin=None, cols=None, literal=False):
    """Packed TT matrix (or columns [j0,j1) of it) from the CPU restatement of clToCMatrix."""
    cl = _f64(cl)
    good = _i32(good)
    pixwin = _f64(pixwin)
    lmax = len(cl) - 1
    n = len(good) if good is not None else 12 * nside * nside
    j0, j1 = (0, n) if cols is None else cols
    out = np.empty(packed_size(j1) - packed_size(j0))
    fn = lib().orc_cl_to_cmatrix_cols_literal if literal else lib().orc_cl_to_cmatrix_cols
    rc = fn(_ptr(cl), ctypes.c_int(lmax), ctypes.c_long(nside), ctypes.c_double(fwhm), _ptr(pixwin),
            _ptr(good), ctypes.c_long(n), ctypes.c_long(j0), ctypes.c_long(j1), _ptr(out))
    assert rc == 0
    return out


def cl_to_cmatrix_pairs(cl, nside, fwhm, pi, pj, good=None, pixwin=None):
    cl = _f64(cl)
    good = _i32(good)
    pixwin = _f64(pixwin)
    pi = np.ascontiguousarray(pi, dtype=np.int64)
    pj = np.ascontiguousarray(pj, dtype=np.int64)
    n = len(good) if good is not None else 0
    out = np.empty(len(pi))
    rc = lib().orc_cl_to_cmatrix_pairs(_ptr(cl), ctypes.c_int(len(cl) - 1), ctypes.c_long(nside), ctypes.c_double(fwhm),
                                       _ptr(pixwin), _ptr(good), ctypes.c_long(n), _ptr(pi), _ptr(pj),
                                       ctypes.c_long(len(pi)), _ptr(out))
    assert rc == 0
    return out


def fiducial_matrix(cl, nside, lmax, fwhm, good=None, pixwin=None):
    cl = _f64(cl)
    assert len(cl) >= 4 * nside + 1
    good = _i32(good)
    pixwin = _f64(pixwin)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty(packed_size(n))
    rc = lib().orc_fiducial_matrix(_ptr(cl), ctypes.c_long(nside), ctypes.c_int(lmax), ctypes.c_double(fwhm),
                                   _ptr(pixwin), _ptr(good), ctypes.c_long(n), _ptr(out))
    assert rc == 0
    return out


def noise_matrix(nside, noise):
    n = 12 * nside * nside
    out = np.empty(packed_size(n))
    lib().orc_noise_matrix(ctypes.c_long(nside), ctypes.c_double(noise), _ptr(out))
    return out


def mask_matrix(packed, good):
    packed = _f64(packed)
    good = _i32(good)
    out = np.empty(packed_size(len(good)))
    lib().orc_mask_matrix(_ptr(packed), _ptr(good), ctypes.c_long(len(good)), _ptr(out))
    return out


def legendre_container(lmax, nside, good=None):
    good = _i32(good)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty((lmax + 1, packed_size(n)))
    rc = lib().orc_legendre_container(ctypes.c_int(lmax), ctypes.c_long(nside), _ptr(good), ctypes.c_long(n), _ptr(out))
    assert rc == 0
    return out


def tqu_matrix(ctt, cte, cee, cbb, nside, fwhm, good=None, pixwinT=None, pixwinP=None):
    ctt, cte, cee, cbb = map(_f64, (ctt, cte, cee, cbb))
    good = _i32(good)
    pixwinT = _f64(pixwinT)
    pixwinP = _f64(pixwinP)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.zeros(packed_size(3 * n))
    rc = lib().orc_tqu_matrix(_ptr(ctt), _ptr(cte), _ptr(cee), _ptr(cbb), ctypes.c_int(len(ctt) - 1), ctypes.c_long(nside),
                              ctypes.c_double(fwhm), _ptr(pixwinT), _ptr(pixwinP), _ptr(good), ctypes.c_long(n), _ptr(out))
    assert rc == 0
    return out


def tqu_pairs(ctt, cte, cee, cbb, nside, fwhm, pi, pj, good=None, pixwinT=None, pixwinP=None):
    """(npairs, 3, 3) blocks <X_a(i) X_b(j)>, X = (T, Q, U)."""
    ctt, cte, cee, cbb = map(_f64, (ctt, cte, cee, cbb))
    good = _i32(good)
    pixwinT = _f64(pixwinT)
    pixwinP = _f64(pixwinP)
    pi = np.ascontiguousarray(pi, dtype=np.int64)
    pj = np.ascontiguousarray(pj, dtype=np.int64)
    n = len(good) if good is not None else 0
    out = np.empty((len(pi), 3, 3))
    rc = lib().orc_tqu_pairs(_ptr(ctt), _ptr(cte), _ptr(cee), _ptr(cbb), ctypes.c_int(len(ctt) - 1), ctypes.c_long(nside),
                             ctypes.c_double(fwhm), _ptr(pixwinT), _ptr(pixwinP), _ptr(good), ctypes.c_long(n),
                             _ptr(pi), _ptr(pj), ctypes.c_long(len(pi)), _ptr(out))
    assert rc == 0
    return out


def unpack_symmetric(packed, n):
    """Dense symmetric matrix from the packed upper triangle (column j at j(j+1)/2)."""
    M = np.zeros((n, n))
    iu = np.triu_indices(n)
    # packed order is column-major over the upper triangle: (i<=j) sorted by j then i
    order = np.lexsort((iu[0], iu[1]))
    M[iu[0][order], iu[1][order]] = packed
    M = M + M.T - np.diag(np.diag(M))
    return M


# ------------------------------------------------------------------ helpers over the reference objects

def ref_set_pixel_window(wT=None, wP=None):
    """HEALPix pixel window the reference object code sees through its Utils::readPixelWindowFunction shim (oracle/ref_shim.cpp):
    wT / wP = temperature / polarization table from l = 0; None = window 1 (what every other helper here assumes: reset it)."""
    wT = _f64(wT)
    wP = _f64(wP)
    ref().ref_set_pixel_window(_ptr(wT), ctypes.c_int(0 if wT is None else len(wT)), _ptr(wP), ctypes.c_int(0 if wP is None else len(wP)))


def ref_cl_to_cmatrix(cl, nside, fwhm, good=None, use_lp=False):
    cl = _f64(cl)
    good = _i32(good)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty(packed_size(n))
    fn = ref().ref_cl_to_cmatrix_lp if use_lp else ref().ref_cl_to_cmatrix
    rc = fn(_ptr(cl), ctypes.c_int(len(cl) - 1), ctypes.c_long(nside), ctypes.c_double(fwhm), _ptr(good),
            ctypes.c_int(0 if good is None else len(good)), _ptr(out))
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
    return out


def ref_fiducial_matrix(cl, nside, lmax, fwhm, good=None):
    cl = _f64(cl)
    good = _i32(good)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty(packed_size(n))
    rc = ref().ref_fiducial_matrix(_ptr(cl), ctypes.c_int(len(cl)), ctypes.c_long(nside), ctypes.c_int(lmax),
                                   ctypes.c_double(fwhm), _ptr(good), ctypes.c_int(0 if good is None else len(good)), _ptr(out))
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
    return out


def ref_noise_matrix_masked(nside, noise, good=None):
    good = _i32(good)
    n = len(good) if good is not None else 12 * nside * nside
    out = np.empty(packed_size(n))
    rc = ref().ref_noise_matrix_masked(ctypes.c_long(nside), ctypes.c_double(noise), _ptr(good),
                                       ctypes.c_int(0 if good is None else len(good)), _ptr(out))
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
    return out


def ref_mask_matrix(packed, npix, good):
    packed = _f64(packed)
    good = _i32(good)
    out = np.empty(packed_size(len(good)))
    rc = ref().ref_mask_matrix(ctypes.c_int(npix), _ptr(packed), _ptr(good), ctypes.c_int(len(good)), _ptr(out))
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
    return out


def ref_write_cmatrix(packed, npix, comment, bin_file=None, text_file=None):
    packed = _f64(packed)
    rc = ref().ref_write_cmatrix(ctypes.c_int(npix), _ptr(packed), comment.encode(),
                                 None if bin_file is None else bin_file.encode(),
                                 None if text_file is None else text_file.encode())
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())


def ref_read_cmatrix(bin_file, capacity):
    out = np.empty(capacity)
    n = ctypes.c_int()
    comment = ctypes.create_string_buffer(4096)
    rc = ref().ref_read_cmatrix(bin_file.encode(), ctypes.byref(n), _ptr(out), ctypes.c_long(capacity), comment, 4096)
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
    return n.value, out[:packed_size(n.value)].copy(), comment.value.decode()


def ref_write_legendre_container(lmax, nside, good, file):
    good = _i32(good)
    rc = ref().ref_write_legendre_container(ctypes.c_int(lmax), ctypes.c_long(nside), _ptr(good),
                                            ctypes.c_int(0 if good is None else len(good)), file.encode())
    if rc:
        raise RuntimeError(ref().ref_last_error().decode())
