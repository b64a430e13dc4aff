"""Device-resident consumer of the generated matrices: the low-l pixel likelihood of reference source/likelihood.cpp
(`Likelihood::construct` :68-134, `calculate` :163-180), fed straight from GPU memory -- no round trip of the matrices
to the host.  Python mirror of include/likelihood.hpp over the C ABI (cmg_like_create / cmg_like_calculate): the packed sum
C + F + N is factorised in place by this library's packed Cholesky (cmg_packed_cholesky: FP64 tensor-core trailing updates
on the packed triangle, no unpacked copy), the GPU counterpart of the reference's LAPACK dpptrf / dpptri on the same storage;
Context.set_like_method(1) selects cuSOLVER potrf + cuBLAS trsm on the unpacked matrix instead (kept for comparison).
"""
import ctypes

import numpy as np

from . import capi

DET_OFFSET = -29677.0566       # the reference subtracts this constant from log det (source/likelihood.cpp:126-127)


class Likelihood:
    """chi2 = t^T C^-1 t (with optional foreground-template marginalisation), logDet = log det C - DET_OFFSET.

    c_packed / fiducial_packed / noise_packed: device buffers (torch tensors or raw pointers) holding packed matrices of
    dimension n; c_stride = capi.SLAB with c_packed pointing at element b % 16 of a slab reads a batched-slab element."""

    def __init__(self, ctx, c_packed, fiducial_packed, noise_packed, n, foreground=None, c_stride=1):
        self._ctx = ctx
        self._L = capi.library()
        self.n = int(n)
        f = None
        if foreground is not None and len(foreground):
            f = np.ascontiguousarray(foreground, dtype=np.float64)
            if f.size != self.n:
                raise ValueError("The foreground map must have one value per unmasked pixel.")
        h = ctypes.c_void_p()
        st = self._L.cmg_like_create(ctx._h, capi._p(c_packed), int(c_stride), capi._p(fiducial_packed), capi._p(noise_packed),
                                     self.n, capi._p(f), ctypes.byref(h))
        if st:
            text = self._L.cmg_last_error(ctx._h).decode()
            if st == 6:            # CMG_ENUMERIC: the reference throws with this text
                raise ValueError(text)
            raise capi.CmgError(st, text)
        self._h = h

    def calculate(self, t):
        """-> (chi2 + logDet, chi2, logDet) for one map (length n) or a batch (maps as rows)"""
        tt = np.ascontiguousarray(t, dtype=np.float64)
        single = tt.ndim == 1
        T = tt.reshape(1, -1) if single else tt
        if T.shape[1] != self.n:
            raise ValueError("map length does not match the number of unmasked pixels")
        chi2 = np.empty(T.shape[0])
        log_det = ctypes.c_double()
        st = self._L.cmg_like_calculate(self._h, capi._p(T), T.shape[0], capi._p(chi2), ctypes.byref(log_det))
        if st:
            raise capi.CmgError(st, self._L.cmg_last_error(self._ctx._h).decode())
        if single:
            return float(chi2[0] + log_det.value), float(chi2[0]), log_det.value
        return chi2 + log_det.value, chi2, log_det.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.cmg_like_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
