"""ctypes binding of the C ABI in include/cmg.h (the reference-side stub of INTEGRATION.md, in Python)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_PARTS = 16
LMAX_LIMIT = 1023

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p


class CmgError(RuntimeError):
    """Non-zero cmg_status (the C++ wrappers raise StandardException in the same places)."""

    def __init__(self, status, text):
        super().__init__("cmg status %d: %s" % (status, text))
        self.status = status


class TquLayout(ctypes.Structure):
    _fields_ = [
        ("n_parts", ctypes.c_int32),
        ("own", ctypes.c_int32),
        ("begin", _i64 * (MAX_PARTS + 1)),
        ("ptr", (_vp * 3) * MAX_PARTS),
        ("kind", ctypes.c_int32 * MAX_PARTS),
        ("ld", _i64 * MAX_PARTS),
        ("row0", _i64 * MAX_PARTS),
    ]


class OrbitShard(ctypes.Structure):
    """cmg_orbit_shard: a rank's pieces of a [T;Q;U] matrix generated over symmetry orbits (include/cmg.h)."""
    _fields_ = [
        ("n_ranks", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("bounds", _i64 * (MAX_PARTS + 1)),
        ("strip", (_vp * 12) * 3),
        ("outbox", _vp),
    ]


CHOL_MAX_RUNS = 36
CHOL_NB = 128
CHOL_MAX_GROUP = 4
CHOL_PLANE_SLACK = 192          # rows a plane of the dense panel has to spare behind the last column


class CholRuns(ctypes.Structure):
    """cmg_chol_runs: the runs of whole packed columns a rank of the sharded Cholesky factorisation holds (include/cmg.h)."""
    _fields_ = [
        ("n_runs", ctypes.c_int32),
        ("col_begin", _i64 * CHOL_MAX_RUNS),
        ("col_end", _i64 * CHOL_MAX_RUNS),
        ("d_run", _vp * CHOL_MAX_RUNS),
    ]


def make_chol_runs(runs):
    """runs: [(col_begin, col_end, device pointer of entry (0, col_begin))], ascending"""
    if not 1 <= len(runs) <= CHOL_MAX_RUNS:
        raise ValueError("1 .. %d runs of columns" % CHOL_MAX_RUNS)
    c = CholRuns()
    c.n_runs = len(runs)
    for k, (b, e, ptr) in enumerate(runs):
        c.col_begin[k], c.col_end[k], c.d_run[k] = int(b), int(e), int(ptr)
    return c


def library_path():
    return os.path.join(HERE, "lib", "libcosmopp_b200.so")


def build_library(verbose=False):
    """nvcc build of the CUDA library for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-6000:], r.stderr[-6000:])
    if r.returncode:
        raise RuntimeError("building libcosmopp_b200.so failed")


_SIGNATURES = {
    "cmg_version": (ctypes.c_int, []),
    "cmg_device_count": (ctypes.c_int, []),
    "cmg_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    "cmg_destroy": (None, [_vp]),
    "cmg_last_error": (ctypes.c_char_p, [_vp]),
    "cmg_set_stream": (ctypes.c_int, [_vp, _vp]),
    "cmg_use_own_stream": (ctypes.c_int, [_vp]),
    "cmg_synchronize": (ctypes.c_int, [_vp]),
    "cmg_launch_count": (_i64, [_vp]),
    "cmg_device_malloc": (ctypes.c_int, [_vp, _i64, ctypes.POINTER(_vp)]),
    "cmg_device_free": (ctypes.c_int, [_vp, _vp]),
    "cmg_host_malloc_pinned": (ctypes.c_int, [_i64, ctypes.POINTER(_vp)]),
    "cmg_host_free_pinned": (ctypes.c_int, [_vp]),
    "cmg_ipc_export": (ctypes.c_int, [_vp, _vp, _vp]),
    "cmg_ipc_open": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "cmg_ipc_close": (ctypes.c_int, [_vp, _vp]),
    "cmg_copy_to_host": (ctypes.c_int, [_vp, _vp, _vp, _i64]),
    "cmg_copy_to_device": (ctypes.c_int, [_vp, _vp, _vp, _i64]),
    "cmg_nside2npix": (_i64, [_i64]),
    "cmg_pix2ang_nest": (ctypes.c_int, [_i64, _i64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "cmg_packed_size": (_i64, [_i64]),
    "cmg_packed_index": (_i64, [_i64, _i64]),
    "cmg_good_pixels_from_mask": (ctypes.c_int, [_vp, _i64, _vp, ctypes.POINTER(_i64)]),
    "cmg_beam_function": (ctypes.c_double, [ctypes.c_int, ctypes.c_double]),
    "cmg_window_beam": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, _vp]),
    "cmg_set_pixels": (ctypes.c_int, [_vp, _i64, _vp, _i64]),
    "cmg_npix": (_i64, [_vp]),
    "cmg_get_geometry": (ctypes.c_int, [_vp, _vp]),
    "cmg_legendre_series": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _i64, _vp]),
    "cmg_legendre_series_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _i64, _vp]),
    "cmg_tt_weights": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "cmg_fiducial_weights": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, _vp]),
    "cmg_cl_to_cmatrix": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    "cmg_fiducial_matrix": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    "cmg_noise_matrix": (ctypes.c_int, [_i64, ctypes.c_double, _vp]),
    "cmg_mask_matrix": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "cmg_tqu_layout_single": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(TquLayout)]),
    "cmg_tqu": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.POINTER(TquLayout)]),
    "cmg_tqu_scatter_block": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, ctypes.c_int, _vp]),
    "cmg_tqu_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.POINTER(TquLayout)]),
    "cmg_tqu_orbit": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int]),
    "cmg_tqu_orbit_sharded": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.POINTER(OrbitShard), ctypes.c_int]),
    "cmg_tqu_orbit_assemble": (ctypes.c_int, [_vp, ctypes.POINTER(OrbitShard), ctypes.c_int, ctypes.c_int, _vp]),
    "cmg_legendre_series_orbit": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "cmg_legendre_series_orbit_sharded": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _i64, _vp]),
    "cmg_orbit_outbox_layout": (ctypes.c_int, [_i64, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, _vp]),
    "cmg_tqu_orbit_scatter_inbox": (ctypes.c_int, [_vp, ctypes.POINTER(OrbitShard), ctypes.c_int, ctypes.c_int, _vp]),
    "cmg_orbit_strips_to_host": (ctypes.c_int, [_vp, ctypes.POINTER(OrbitShard), _vp, ctypes.c_int, ctypes.c_int]),
    "cmg_copy_on_device": (ctypes.c_int, [_vp, _vp, _vp, _i64]),
    "cmg_transfer_counters": (ctypes.c_int, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "cmg_cl_to_cmatrix_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    "cmg_fiducial_matrix_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    "cmg_cl_to_cmatrix_pol_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp, _vp]),
    "cmg_matrix_to_host": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, _vp]),
    "cmg_tqu_orbit_plan_mirror": (ctypes.c_int, [_i64, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int32)]),
    "cmg_packed_cholesky": (ctypes.c_int, [_vp, _vp, _i64, ctypes.POINTER(_i64)]),
    "cmg_packed_cholesky_logdet": (ctypes.c_int, [_vp, _vp, _i64, ctypes.POINTER(ctypes.c_double)]),
    "cmg_packed_cholesky_solve": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64]),
    "cmg_chol_begin": (ctypes.c_int, [_vp]),
    "cmg_chol_end": (ctypes.c_int, [_vp, ctypes.POINTER(_i64)]),
    "cmg_chol_diag": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), _i64, ctypes.c_int, _vp]),
    "cmg_chol_panel": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), _i64, ctypes.c_int, _vp, _vp, _i64]),
    "cmg_chol_syrk": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), _i64, ctypes.c_int, _i64, _vp, _i64, _i64, ctypes.c_int]),
    "cmg_set_cholesky_group": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_set_cholesky_lookahead": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_chol_logdet_runs": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), ctypes.POINTER(ctypes.c_double)]),
    "cmg_chol_solve_diag": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), _i64, ctypes.c_int, _i64, _vp, _i64]),
    "cmg_chol_solve_update": (ctypes.c_int, [_vp, ctypes.POINTER(CholRuns), _i64, ctypes.c_int, _i64, _vp, _i64]),
    "cmg_packed_sum": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "cmg_set_like_method": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_like_create_ninv": (ctypes.c_int, [_vp, _vp, _vp, _i64, ctypes.c_double, ctypes.POINTER(_vp)]),
    "cmg_host_register": (ctypes.c_int, [_vp, _i64]),
    "cmg_host_unregister": (ctypes.c_int, [_vp]),
    "cmg_host_expand_rotations": (ctypes.c_int, [_vp, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "cmg_set_host_expand": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_set_host_expand_direct": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_tqu_orbit_plan": (ctypes.c_int, [_i64, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int32)]),
    "cmg_tqu_weights": (ctypes.c_int, [_vp] * 6 + [ctypes.c_int] + [_vp] * 4),
    "cmg_cl_to_cmatrix_pol": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_double, _vp, _vp, _vp]),
    "cmg_legendre_series_batched": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _i64, _i64, _vp, _i64]),
    "cmg_tqu_batched": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _vp, _i64]),
    "cmg_slab_doubles": (_i64, [_i64]),
    "cmg_tqu_batched_slab": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _vp]),
    "cmg_tqu_batched_slab_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _vp]),
    "cmg_slab_unpack": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, ctypes.c_int, _vp, _i64]),
    "cmg_sum_unpack": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "cmg_sum_unpack_strided": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "cmg_like_create": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, ctypes.POINTER(_vp)]),
    "cmg_like_calculate": (ctypes.c_int, [_vp, _vp, _i64, _vp, ctypes.POINTER(ctypes.c_double)]),
    "cmg_like_destroy": (None, [_vp]),
    "cmg_cmatrix_file_create": (ctypes.c_int, [_vp, ctypes.c_char_p, _i64, ctypes.POINTER(_vp)]),
    "cmg_cmatrix_file_write_device": (ctypes.c_int, [_vp, _i64, _vp, _i64]),
    "cmg_cmatrix_file_open": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_i64), ctypes.POINTER(_vp)]),
    "cmg_cmatrix_file_read_device": (ctypes.c_int, [_vp, _i64, _vp, _i64]),
    "cmg_cmatrix_file_comment": (ctypes.c_int, [_vp, ctypes.c_char_p, _i64]),
    "cmg_cmatrix_file_close": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "cmg_measure_fp64_peak": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "cmg_last_kernel_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "cmg_set_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "cmg_set_kernel_variant": (ctypes.c_int, [_vp, ctypes.c_int]),
}

_lib = None


def library():
    """Load lib/libcosmopp_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the generator has no CPU fallback)" % path)
        L = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_names():
    return sorted(_SIGNATURES)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return _vp(a.ctypes.data)
    if isinstance(a, int):
        return _vp(a)
    return _vp(a.data_ptr())        # torch tensor


def host_expand_rotations(packed, nside, strips, threads=1, faces=(0, 12)):
    """cmg_host_expand_rotations on a numpy array in place (pure CPU); strips = 1 (TT) or 3 ([T;Q;U])"""
    st = library().cmg_host_expand_rotations(_p(packed), int(nside), 0, int(strips), int(faces[0]), int(faces[1]), int(threads))
    if st:
        raise CmgError(st, "cmg_host_expand_rotations: bad arguments")


def orbit_plan(nside, mode=0):
    """Classes of base-face pairs of cmg_tqu_orbit: list of dicts (host only)."""
    out = np.zeros((24, 21), dtype=np.int32)
    n = ctypes.c_int32()
    st = library().cmg_tqu_orbit_plan(int(nside), int(mode), _p(out), ctypes.byref(n))
    if st:
        raise CmgError(st, "cmg_tqu_orbit_plan: bad nside / mode")
    mir = np.zeros((24, 5), dtype=np.int32)
    st = library().cmg_tqu_orbit_plan_mirror(int(nside), int(mode), _p(mir), ctypes.byref(n))
    if st:
        raise CmgError(st, "cmg_tqu_orbit_plan_mirror: bad nside / mode")
    plan = []
    for row, m in zip(out[:n.value], mir[:n.value]):
        imgs = [(int(row[5 + 3 * k]), int(row[6 + 3 * k]), bool(row[7 + 3 * k])) for k in range(int(row[4]))]
        # mirror images (mode 3): (row face, column face) of the four rotations of the pair's image under the meridian mirror
        mirror = [(int(m[1 + k]), int(row[6 + 3 * k])) for k in range(4)] if m[0] else []
        plan.append(dict(row_face=int(row[0]), col_face=int(row[1]), tri=bool(row[2]), same_face=bool(row[3]), images=imgs,
                         combo_base=[int(row[17 + k]) for k in range(int(row[4]))], mirror_images=mirror))
    return plan


def orbit_outbox_layout(nside, mode, bounds, rank):
    """cmg_orbit_outbox_layout: offsets (doubles) of the block for every destination rank inside `rank`'s outbox; [-1] = total"""
    b = np.ascontiguousarray(bounds, dtype=np.int64)
    off = np.zeros(len(b), dtype=np.int64)
    st = library().cmg_orbit_outbox_layout(int(nside), int(mode), len(b) - 1, _p(b), int(rank), _p(off))
    if st:
        raise CmgError(st, "cmg_orbit_outbox_layout: bad arguments")
    return [int(x) for x in off]


def host_register(array):
    """cudaHostRegister of a numpy array's memory (e.g. a shared mapping), for PCIe-speed copies into it"""
    st = library().cmg_host_register(_p(array), int(array.nbytes))
    if st:
        raise CmgError(st, "cmg_host_register: %s" % library().cmg_last_error(None).decode())


def host_unregister(array):
    library().cmg_host_unregister(_p(array))


SLAB = 16            # CMG_SLAB: batch elements interleaved in one slab of the DMMA batched path


def slab_doubles(dim):
    return int(library().cmg_slab_doubles(dim))


class Context:
    """One GPU + one stream (cmg_ctx).  Device buffers are torch tensors or raw pointers."""

    def __init__(self, device=0, stream=None):
        self._L = library()
        h = _vp()
        st = self._L.cmg_create(ctypes.byref(h), int(device))
        if st:
            raise CmgError(st, self._L.cmg_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.stream_handle = None          # None = the context's private stream
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._L.cmg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st:
            raise CmgError(st, self._L.cmg_last_error(self._h).decode())

    # ---- plumbing
    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream; 0 is the legacy
        default stream); None returns to the context's private stream."""
        if cuda_stream is None:
            self._check(self._L.cmg_use_own_stream(self._h))
            self.stream_handle = None
        else:
            self._check(self._L.cmg_set_stream(self._h, _vp(int(cuda_stream))))
            self.stream_handle = int(cuda_stream)

    def synchronize(self):
        self._check(self._L.cmg_synchronize(self._h))

    @property
    def launches(self):
        return int(self._L.cmg_launch_count(self._h))

    def set_timing(self, on):
        self._check(self._L.cmg_set_timing(self._h, 1 if on else 0))

    def set_host_expand(self, threads):
        """opt-in: full-sky whole calls copy back the last-face columns only and fill in the rest on `threads` host threads"""
        self._check(self._L.cmg_set_host_expand(self._h, int(threads)))

    def set_kernel_variant(self, v):
        self._check(self._L.cmg_set_kernel_variant(self._h, int(v)))

    def last_kernel_ms(self):
        v = ctypes.c_double()
        self._check(self._L.cmg_last_kernel_ms(self._h, ctypes.byref(v)))
        return v.value

    def measure_fp64_peak(self):
        v = ctypes.c_double()
        self._check(self._L.cmg_measure_fp64_peak(self._h, ctypes.byref(v)))
        return v.value

    # ---- raw device buffers (shareable between the ranks of one box)
    def device_malloc(self, nbytes):
        p = _vp()
        self._check(self._L.cmg_device_malloc(self._h, int(nbytes), ctypes.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._check(self._L.cmg_device_free(self._h, _vp(ptr)))

    def ipc_export(self, ptr):
        buf = ctypes.create_string_buffer(64)
        self._check(self._L.cmg_ipc_export(self._h, _vp(ptr), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = _vp()
        self._check(self._L.cmg_ipc_open(self._h, ctypes.c_char_p(handle), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._check(self._L.cmg_ipc_close(self._h, _vp(ptr)))

    def copy_to_host(self, dst_host, d_src, nbytes):
        self._check(self._L.cmg_copy_to_host(self._h, _p(dst_host), _p(d_src), int(nbytes)))

    # ---- geometry
    def set_pixels(self, nside, good=None):
        if good is None:
            self._check(self._L.cmg_set_pixels(self._h, int(nside), None, 0))
        else:
            g = np.ascontiguousarray(good, dtype=np.int32)
            self._check(self._L.cmg_set_pixels(self._h, int(nside), _p(g), len(g)))
        self.nside = int(nside)

    @property
    def npix(self):
        return int(self._L.cmg_npix(self._h))

    def geometry(self):
        out = np.empty((8, self.npix))
        self._check(self._L.cmg_get_geometry(self._h, _p(out)))
        return out

    # ---- TT
    def legendre_series(self, a, d_out, col_begin=0, col_end=None):
        a = _f64(a)
        col_end = self.npix if col_end is None else col_end
        self._check(self._L.cmg_legendre_series(self._h, _p(a), len(a) - 1, col_begin, col_end, _p(d_out)))

    def legendre_series_orbit(self, a, d_out):
        """full-sky TT matrix over symmetry orbits (include/cmg.h), whole packed triangle"""
        a = _f64(a)
        self._check(self._L.cmg_legendre_series_orbit(self._h, _p(a), len(a) - 1, _p(d_out)))

    def legendre_series_dev(self, d_a, lmax, d_out, col_begin=0, col_end=None):
        col_end = self.npix if col_end is None else col_end
        self._check(self._L.cmg_legendre_series_dev(self._h, _p(d_a), lmax, col_begin, col_end, _p(d_out)))

    def legendre_series_batched(self, a, d_out, stride, col_begin=0, col_end=None):
        a = _f64(a)
        col_end = self.npix if col_end is None else col_end
        self._check(self._L.cmg_legendre_series_batched(self._h, _p(a), a.shape[1] - 1, a.shape[0], col_begin, col_end,
                                                        _p(d_out), stride))

    def cl_to_cmatrix(self, cl, fwhm, out_host, pixwin=None):
        cl = _f64(cl)
        pixwin = _f64(pixwin)
        _need(len(cl) >= 3, "cl must reach l = 2")
        _need(pixwin is None or len(pixwin) >= len(cl), "pixel window shorter than cl")
        _need(_numel(out_host) >= packed_size(self.npix), "output buffer smaller than npix (npix + 1) / 2")
        self._check(self._L.cmg_cl_to_cmatrix(self._h, _p(cl), len(cl) - 1, float(fwhm), _p(pixwin), _p(out_host)))

    def fiducial_matrix(self, cl, lmax, fwhm, out_host, pixwin=None):
        cl = _f64(cl)
        pixwin = _f64(pixwin)
        n_need = 4 * self.nside + 1                   # reference check: cl.size() >= 4 nSide + 1 (source/c_matrix_generator.cpp:709)
        _need(len(cl) >= n_need, "cl must hold 4 nside + 1 entries for the fiducial matrix")
        _need(pixwin is None or len(pixwin) >= n_need, "pixel window shorter than 4 nside + 1")
        _need(_numel(out_host) >= packed_size(self.npix), "output buffer smaller than npix (npix + 1) / 2")
        self._check(self._L.cmg_fiducial_matrix(self._h, _p(cl), int(lmax), float(fwhm), _p(pixwin), _p(out_host)))

    def cl_to_cmatrix_dev(self, cl, fwhm, d_out, pixwin=None):
        """clToCMatrix with the packed result left on the device (include/cmg.h)"""
        cl = _f64(cl)
        pixwin = _f64(pixwin)
        _need(len(cl) >= 3, "cl must reach l = 2")
        _need(pixwin is None or len(pixwin) >= len(cl), "pixel window shorter than cl")
        _need(_numel(d_out) >= packed_size(self.npix), "output buffer smaller than npix (npix + 1) / 2")
        self._check(self._L.cmg_cl_to_cmatrix_dev(self._h, _p(cl), len(cl) - 1, float(fwhm), _p(pixwin), _p(d_out)))

    def cl_to_cmatrix_pol_dev(self, ctt, cte, cee, cbb, fwhm, d_out, pixwinT=None, pixwinP=None):
        """the [T;Q;U] whole call with the packed result left on the device (include/cmg.h)"""
        ctt, cte, cee, cbb = map(_f64, (ctt, cte, cee, cbb))
        pixwinT = _f64(pixwinT)
        pixwinP = _f64(pixwinP)
        _need(len(ctt) >= 3 and len(cte) == len(ctt) and len(cee) == len(ctt) and len(cbb) == len(ctt), "tt, te, ee, bb must have one length (lmax + 1 >= 3)")
        _need(all(w is None or len(w) >= len(ctt) for w in (pixwinT, pixwinP)), "pixel window shorter than the spectra")
        _need(_numel(d_out) >= packed_size(3 * self.npix), "output buffer smaller than 3 npix (3 npix + 1) / 2")
        self._check(self._L.cmg_cl_to_cmatrix_pol_dev(self._h, _p(ctt), _p(cte), _p(cee), _p(cbb), len(ctt) - 1, float(fwhm),
                                                      _p(pixwinT), _p(pixwinP), _p(d_out)))

    def mask_matrix(self, d_in, npix_in, good, d_out):
        g = np.ascontiguousarray(good, dtype=np.int32)
        self._check(self._L.cmg_mask_matrix(self._h, _p(d_in), npix_in, _p(g), len(g), _p(d_out)))

    # ---- TQU
    def tqu_layout_single(self, d_packed):
        lay = TquLayout()
        self._check(self._L.cmg_tqu_layout_single(self._h, _p(d_packed), ctypes.byref(lay)))
        return lay

    def tqu(self, a_tt, a_te, a_ee, a_bb, layout):
        a_tt, a_te, a_ee, a_bb = map(_f64, (a_tt, a_te, a_ee, a_bb))
        self._check(self._L.cmg_tqu(self._h, _p(a_tt), _p(a_te), _p(a_ee), _p(a_bb), len(a_tt) - 1, ctypes.byref(layout)))

    def tqu_dev(self, d_a, lmax, layout):
        self._check(self._L.cmg_tqu_dev(self._h, _p(d_a), int(lmax), ctypes.byref(layout)))

    def legendre_series_orbit_sharded(self, a, q_begin, q_end, strip_ptrs):
        """a rank's columns of the full-sky TT matrix over symmetry orbits: 12 strips, no exchange (include/cmg.h)"""
        a = _f64(a)
        ptrs = (_vp * 12)(*[_vp(int(p)) for p in strip_ptrs])
        self._check(self._L.cmg_legendre_series_orbit_sharded(self._h, _p(a), len(a) - 1, int(q_begin), int(q_end), ctypes.cast(ptrs, _vp)))

    def tqu_orbit(self, a_tt, a_te, a_ee, a_bb, d_packed, mode=0):
        """full-sky path: one evaluation per orbit of pixel pairs under the pi/2 rotation of the grid (include/cmg.h)"""
        a_tt, a_te, a_ee, a_bb = map(_f64, (a_tt, a_te, a_ee, a_bb))
        self._check(self._L.cmg_tqu_orbit(self._h, _p(a_tt), _p(a_te), _p(a_ee), _p(a_bb), len(a_tt) - 1, _p(d_packed), int(mode)))

    def tqu_orbit_sharded(self, a_tt, a_te, a_ee, a_bb, shard, mode=0):
        a_tt, a_te, a_ee, a_bb = map(_f64, (a_tt, a_te, a_ee, a_bb))
        self._check(self._L.cmg_tqu_orbit_sharded(self._h, _p(a_tt), _p(a_te), _p(a_ee), _p(a_bb), len(a_tt) - 1,
                                                  ctypes.byref(shard), int(mode)))

    def tqu_orbit_assemble(self, shard, d_full, mode=0, parts=3):
        """parts: 1 = strips, 2 = outbox blocks; place the strips of all ranks before any outbox (include/cmg.h)"""
        self._check(self._L.cmg_tqu_orbit_assemble(self._h, ctypes.byref(shard), int(mode), int(parts), _p(d_full)))

    def tqu_orbit_scatter_inbox(self, shard, sender, d_block, mode=0):
        """block(sender -> this rank) (local or IPC-mapped peer memory) into this rank's strips (include/cmg.h)"""
        self._check(self._L.cmg_tqu_orbit_scatter_inbox(self._h, ctypes.byref(shard), int(mode), int(sender), _p(d_block)))

    def orbit_strips_to_host(self, shard, host_packed, threads=0, direct_mask=0):
        """this rank's complete strips as its columns of one whole packed host matrix (include/cmg.h)"""
        self._check(self._L.cmg_orbit_strips_to_host(self._h, ctypes.byref(shard), _p(host_packed), int(threads), int(direct_mask)))

    def set_host_expand_direct(self, direct_mask):
        self._check(self._L.cmg_set_host_expand_direct(self._h, int(direct_mask)))

    def packed_cholesky(self, d_packed, n):
        """in-place A = U^T U on a packed upper triangle (include/cmg.h); returns LAPACK's info (0 = positive definite)"""
        info = _i64()
        self._check(self._L.cmg_packed_cholesky(self._h, _p(d_packed), int(n), ctypes.byref(info)))
        return info.value

    def packed_cholesky_logdet(self, d_factor, n):
        v = ctypes.c_double()
        self._check(self._L.cmg_packed_cholesky_logdet(self._h, _p(d_factor), int(n), ctypes.byref(v)))
        return v.value

    def packed_cholesky_solve(self, d_factor, n, d_rhs, n_rhs):
        self._check(self._L.cmg_packed_cholesky_solve(self._h, _p(d_factor), int(n), _p(d_rhs), int(n_rhs)))

    # ---- the factorisation step by step over a rank's runs of columns (include/cmg.h, cmg_chol_runs; multigpu.ShardedCholesky)
    def chol_begin(self):
        self._check(self._L.cmg_chol_begin(self._h))

    def chol_end(self):
        info = _i64()
        self._check(self._L.cmg_chol_end(self._h, ctypes.byref(info)))
        return info.value

    def chol_diag(self, runs, k0, kb, d_ukk):
        self._check(self._L.cmg_chol_diag(self._h, ctypes.byref(runs), int(k0), int(kb), _p(d_ukk)))

    def chol_panel(self, runs, k0, kb, d_ukk, d_plane, panel_col0):
        self._check(self._L.cmg_chol_panel(self._h, ctypes.byref(runs), int(k0), int(kb), _p(d_ukk), _p(d_plane), int(panel_col0)))

    def chol_syrk(self, runs, k0, kb, d_panel, plane_stride, panel_col0, strip_only, shift=0):
        self._check(self._L.cmg_chol_syrk(self._h, ctypes.byref(runs), int(k0), int(kb), int(shift), _p(d_panel), int(plane_stride), int(panel_col0),
                                          int(bool(strip_only))))

    def set_cholesky_group(self, blocks):
        """blocks of 128 rows per trailing update of cmg_packed_cholesky (1 .. 4)"""
        self._check(self._L.cmg_set_cholesky_group(self._h, int(blocks)))

    def chol_logdet_runs(self, runs):
        v = ctypes.c_double()
        self._check(self._L.cmg_chol_logdet_runs(self._h, ctypes.byref(runs), ctypes.byref(v)))
        return v.value

    def chol_solve_diag(self, runs, k0, kb, n, d_rhs, n_rhs):
        self._check(self._L.cmg_chol_solve_diag(self._h, ctypes.byref(runs), int(k0), int(kb), int(n), _p(d_rhs), int(n_rhs)))

    def chol_solve_update(self, runs, k0, kb, n, d_rhs, n_rhs):
        self._check(self._L.cmg_chol_solve_update(self._h, ctypes.byref(runs), int(k0), int(kb), int(n), _p(d_rhs), int(n_rhs)))

    def set_cholesky_lookahead(self, on):
        """factorise the next group of blocks beside the trailing update (cmg_packed_cholesky; default on)"""
        self._check(self._L.cmg_set_cholesky_lookahead(self._h, int(bool(on))))

    def packed_sum(self, d_c, d_f, d_n, n, d_out, c_stride=1):
        self._check(self._L.cmg_packed_sum(self._h, _p(d_c), int(c_stride), _p(d_f), _p(d_n), int(n), _p(d_out)))

    def set_like_method(self, method):
        self._check(self._L.cmg_set_like_method(self._h, int(method)))

    def copy_on_device(self, d_dst, d_src, nbytes):
        """stream-ordered device-to-device copy (either side may be CUDA-IPC mapped peer memory)"""
        self._check(self._L.cmg_copy_on_device(self._h, _p(d_dst), _p(d_src), int(nbytes)))

    def transfer_counters(self):
        """(host-to-device, device-to-host) bytes this context has moved so far"""
        a, b = _i64(), _i64()
        self._check(self._L.cmg_transfer_counters(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def tqu_scatter_block(self, d_block, col0, n_cols, ld, row0, kind, d_full):
        self._check(self._L.cmg_tqu_scatter_block(self._h, _p(d_block), col0, n_cols, ld, row0, kind, _p(d_full)))

    def sum_unpack(self, d_c, d_f, d_n, n, d_full):
        self._check(self._L.cmg_sum_unpack(self._h, _p(d_c), _p(d_f), _p(d_n), int(n), _p(d_full)))

    def tqu_batched(self, a, d_out, stride):
        a = _f64(a)         # [B][4][lmax+1]
        self._check(self._L.cmg_tqu_batched(self._h, _p(a), a.shape[2] - 1, a.shape[0], _p(d_out), stride))

    def tqu_batched_slab(self, a, d_slabs):
        """DMMA path: ceil(B / 16) slabs of 16 interleaved packed matrices (include/cmg.h)."""
        a = _f64(a)         # [B][4][lmax+1]
        self._check(self._L.cmg_tqu_batched_slab(self._h, _p(a), a.shape[2] - 1, a.shape[0], _p(d_slabs)))

    def tqu_batched_slab_dev(self, d_a, lmax, n_batch, d_slabs):
        """weights [B][4][lmax+1] already on the device"""
        self._check(self._L.cmg_tqu_batched_slab_dev(self._h, _p(d_a), int(lmax), int(n_batch), _p(d_slabs)))

    def slab_unpack(self, d_slab, dim, d_out, out_stride=0, n_live=SLAB, only_b=-1):
        self._check(self._L.cmg_slab_unpack(self._h, _p(d_slab), dim, n_live, only_b, _p(d_out), out_stride))

    # ---- CMatrix files straight from / to device memory (reference binary format)
    def write_cmatrix_file(self, path, n_pix, pieces, comment=""):
        """pieces: iterable of (first_element, device_buffer, count) in any order"""
        h = _vp()
        self._check(self._L.cmg_cmatrix_file_create(self._h, path.encode(), int(n_pix), ctypes.byref(h)))
        try:
            for first, d_src, count in pieces:
                self._check(self._L.cmg_cmatrix_file_write_device(h, int(first), _p(d_src), int(count)))
        finally:
            self._check(self._L.cmg_cmatrix_file_close(h, comment.encode()))

    def read_cmatrix_file(self, path, d_dst=None, first=0, count=None):
        """-> (n_pix, comment); fills d_dst with elements [first, first + count) when given"""
        h = _vp()
        n = _i64()
        self._check(self._L.cmg_cmatrix_file_open(self._h, path.encode(), ctypes.byref(n), ctypes.byref(h)))
        try:
            if d_dst is not None:
                total = n.value * (n.value + 1) // 2
                self._check(self._L.cmg_cmatrix_file_read_device(h, int(first), _p(d_dst), int(total - first if count is None else count)))
            buf = ctypes.create_string_buffer(4096)
            self._check(self._L.cmg_cmatrix_file_comment(h, buf, 4096))
        finally:
            self._L.cmg_cmatrix_file_close(h, None)
        return n.value, buf.value.decode()

    def cl_to_cmatrix_pol(self, ctt, cte, cee, cbb, fwhm, out_host, pixwinT=None, pixwinP=None):
        ctt, cte, cee, cbb = map(_f64, (ctt, cte, cee, cbb))
        pixwinT = _f64(pixwinT)
        pixwinP = _f64(pixwinP)
        _need(len(ctt) >= 3 and len(cte) == len(ctt) and len(cee) == len(ctt) and len(cbb) == len(ctt), "tt, te, ee, bb must have one length (lmax + 1 >= 3)")
        _need(all(w is None or len(w) >= len(ctt) for w in (pixwinT, pixwinP)), "pixel window shorter than the spectra")
        _need(_numel(out_host) >= packed_size(3 * self.npix), "output buffer smaller than 3 npix (3 npix + 1) / 2")
        self._check(self._L.cmg_cl_to_cmatrix_pol(self._h, _p(ctt), _p(cte), _p(cee), _p(cbb), len(ctt) - 1, float(fwhm),
                                                  _p(pixwinT), _p(pixwinP), _p(out_host)))


def make_tqu_layout(bounds, rank, strip_ptrs, outbox_ptrs):
    """cmg_tqu_layout of one rank of a sharded generation (cosmopp_b200.partition.tqu_rank_plan):
    strip_ptrs = device addresses of its T, Q, U strips; outbox_ptrs = {owner: (three addresses)} of its dense blocks."""
    from . import partition
    n_parts = len(bounds) - 1
    if n_parts > MAX_PARTS:
        raise ValueError("at most %d parts" % MAX_PARTS)
    plan = partition.tqu_rank_plan(bounds[-1], bounds, rank)
    lay = TquLayout()
    lay.n_parts = n_parts
    lay.own = rank
    for k in range(n_parts + 1):
        lay.begin[k] = bounds[k]
    for s in range(3):
        lay.ptr[rank][s] = strip_ptrs[s]
    lay.kind[rank] = 0
    for owner, _ncols, ld, row0 in plan["outbox"]:
        for s in range(3):
            lay.ptr[owner][s] = outbox_ptrs[owner][s]
        lay.kind[owner] = 1
        lay.ld[owner] = ld
        lay.row0[owner] = row0
    return lay


# ---- pure-host helpers of the ABI (no GPU needed)

def _need(ok, text):
    if not ok:
        raise CmgError(1, text)


def _numel(buf):
    if hasattr(buf, "numel"):
        return int(buf.numel())
    return int(np.asarray(buf).size) if not isinstance(buf, (int, ctypes.c_void_p)) else 1 << 62


def window_beam(lmax, fwhm, pixwin=None):
    f = np.empty(lmax + 1)
    pixwin = _f64(pixwin)
    _need(pixwin is None or len(pixwin) >= lmax + 1, "pixel window shorter than lmax + 1")
    st = library().cmg_window_beam(_p(f), lmax, float(fwhm), _p(pixwin))
    if st:
        raise CmgError(st, "cmg_window_beam")
    return f


def tt_weights(cl, f):
    cl = _f64(cl)
    f = _f64(f)
    _need(len(f) >= len(cl), "window*beam factors shorter than cl")
    a = np.empty(len(cl))
    st = library().cmg_tt_weights(_p(cl), _p(f), len(cl) - 1, _p(a))
    if st:
        raise CmgError(st, "cmg_tt_weights")
    return a


def fiducial_weights(cl, f, nside, lmax):
    cl = _f64(cl)
    f = _f64(f)
    _need(len(cl) >= 4 * nside + 1 and len(f) >= 4 * nside + 1, "cl and f must hold 4 nside + 1 entries")
    a = np.empty(4 * nside + 1)
    st = library().cmg_fiducial_weights(_p(cl), _p(f), nside, lmax, _p(a))
    if st:
        raise CmgError(st, "cmg_fiducial_weights")
    return a


def tqu_weights(ctt, cte, cee, cbb, fT, fP):
    ctt, cte, cee, cbb, fT, fP = map(_f64, (ctt, cte, cee, cbb, fT, fP))
    n = len(ctt)
    _need(all(len(x) == n for x in (cte, cee, cbb)) and len(fT) >= n and len(fP) >= n, "spectra of one length, factors at least as long")
    out = [np.empty(n) for _ in range(4)]
    st = library().cmg_tqu_weights(_p(ctt), _p(cte), _p(cee), _p(cbb), _p(fT), _p(fP), n - 1, *[_p(o) for o in out])
    if st:
        raise CmgError(st, "cmg_tqu_weights")
    return out


def pix2ang_nest(nside, ipix):
    t = ctypes.c_double()
    p = ctypes.c_double()
    st = library().cmg_pix2ang_nest(nside, int(ipix), ctypes.byref(t), ctypes.byref(p))
    if st:
        raise CmgError(st, "cmg_pix2ang_nest: invalid nside/ipix")
    return t.value, p.value


def good_pixels_from_mask(mask):
    mask = _f64(mask)
    good = np.empty(len(mask), dtype=np.int32)
    n = _i64()
    st = library().cmg_good_pixels_from_mask(_p(mask), len(mask), _p(good), ctypes.byref(n))
    if st:
        raise CmgError(st, "cmg_good_pixels_from_mask")
    return good[:n.value].copy()


def packed_size(dim):
    return int(library().cmg_packed_size(dim))


def packed_index(i, j):
    return int(library().cmg_packed_index(i, j))


def noise_matrix(npix, noise):
    out = np.empty(packed_size(npix))
    st = library().cmg_noise_matrix(npix, float(noise), _p(out))
    if st:
        raise CmgError(st, "cmg_noise_matrix")
    return out
