"""Synthetic inputs of the benchmark and the parity tests: random C_l of CMB-like shape.

C_l^TT = 1000 u_l / (l (l+1)), u_l ~ U(0.5, 1.5); C^EE = 0.03 C^TT u', C^BB = 0.002 C^TT u'',
C^TE = rho_l sqrt(C^TT C^EE), rho_l ~ U(-0.5, 0.5), so every (T,E) 2x2 block is positive definite.
Entries l = 0, 1 are present and ignored by the generator (reference source/c_matrix_generator.cpp:190).
"""
import numpy as np


def synthetic_cl(lmax, seed=12345, pol=False):
    rs = np.random.RandomState(seed)
    l = np.arange(lmax + 1, dtype=np.float64)
    tt = np.zeros(lmax + 1)
    tt[2:] = 1000.0 * rs.uniform(0.5, 1.5, lmax - 1) / (l[2:] * (l[2:] + 1))
    if not pol:
        return tt
    ee = np.zeros(lmax + 1)
    bb = np.zeros(lmax + 1)
    te = np.zeros(lmax + 1)
    ee[2:] = 0.03 * tt[2:] * rs.uniform(0.5, 1.5, lmax - 1)
    bb[2:] = 0.002 * tt[2:] * rs.uniform(0.5, 1.5, lmax - 1)
    te[2:] = rs.uniform(-0.5, 0.5, lmax - 1) * np.sqrt(tt[2:] * ee[2:])
    return tt, te, ee, bb
