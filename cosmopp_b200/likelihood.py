"""Device-resident consumer of the generated matrices: the low-l pixel likelihood of reference source/likelihood.cpp
(`Likelihood::construct` :68-134, `calculate` :163-180), fed straight from GPU memory -- no 87 GB round trip to the host.

This is the "next row" after the hot path, provided as a Python utility: the C + F + N sum / unpack is
this repo's kernel (cmg_sum_unpack); the dense factorisation and triangular solves are plain library calls
(torch.linalg -> cuSOLVER / cuBLAS), the GPU counterpart of the reference's LAPACK dpptrf / dpptri.
"""
import math

DET_OFFSET = -29677.0566       # the reference subtracts this constant from log det (source/likelihood.cpp:126-127)


class Likelihood:
    """chi2 = t^T C^-1 t (with optional foreground-template marginalisation), logDet = log det C - DET_OFFSET."""

    def __init__(self, ctx, c_packed, fiducial_packed, noise_packed, n, foreground=None):
        import torch
        self.n = int(n)
        full = torch.empty((self.n, self.n), dtype=torch.float64, device="cuda")       # column-major == row-major: symmetric
        ctx.sum_unpack(c_packed, fiducial_packed, noise_packed, self.n, full)
        L, info = torch.linalg.cholesky_ex(full)
        if int(info.item()) != 0:
            raise ValueError("The determinant of the covariance matrix is not positive. The covariance matrix must be positive definite.")
        self.L = L
        self.logDet = 2.0 * float(torch.log(torch.diagonal(L)).sum().item()) - DET_OFFSET
        self.f = None
        if foreground is not None and len(foreground):
            f = torch.as_tensor(foreground, dtype=torch.float64, device="cuda").reshape(-1, 1)
            self.yf = torch.linalg.solve_triangular(L, f, upper=False)
            self.fCinvf = float((self.yf * self.yf).sum().item())
            self.f = f

    def calculate(self, t):
        """-> (chi2 + logDet, chi2, logDet) for one map (length n) or a batch (maps as rows)"""
        import torch
        tt = torch.as_tensor(t, dtype=torch.float64, device="cuda")
        single = tt.ndim == 1
        T = tt.reshape(1, -1) if single else tt
        if T.shape[1] != self.n:
            raise ValueError("map length does not match the number of unmasked pixels")
        y = torch.linalg.solve_triangular(self.L, T.T.contiguous(), upper=False)          # L y = t
        chi2 = (y * y).sum(0)
        logDet = self.logDet
        if self.f is not None:
            tCinvf = (y * self.yf).sum(0)
            logDet = logDet + math.log(self.fCinvf / self.n)
            chi2 = chi2 - tCinvf * tCinvf / self.fCinvf
        chi2 = chi2.cpu().numpy()
        if single:
            return float(chi2[0] + logDet), float(chi2[0]), logDet
        return chi2 + logDet, chi2, logDet
