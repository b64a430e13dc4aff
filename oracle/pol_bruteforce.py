"""TEST INFRASTRUCTURE ONLY -- definition-level oracle for the polarized pixel covariance.

PARITY UNPINNED BY THE REFERENCE (cosmopp has no TE/BB pixel generator and its EE routine,
reference source/c_matrix_generator.cpp:485-695, needs HEALPix SHTs).  This module evaluates the
covariance straight from its definition, with no addition theorem, no Wigner-d recurrences and no
rotation angles, so that it can pin both oracle/pol_oracle.c and the CUDA kernels:

    T(n)        = sum_lm aT_lm Y_lm(n)
    (Q +- iU)(n) = sum_lm a_{+-2,lm} {+-2}Y_lm(n),      a_{+-2,lm} = -(aE_lm +- i aB_lm)
    <aX_lm aY*_l'm'> = C^XY_l delta_ll' delta_mm'        (TB = EB = 0)

(HEALPix primer convention; Q,U in the local (e_theta, e_phi) basis.)  Spin-weighted harmonics
come from the closed form of Goldberg et al. (1967); pure-Python loops, small cases only.
"""
import math
from math import comb, factorial

import numpy as np


def sYlm(s, l, m, theta, phi):
    """Spin-weighted spherical harmonic {s}Y_lm(theta, phi), Goldberg et al. 1967 eq. (3.1)."""
    if abs(m) > l or abs(s) > l:
        return 0j
    pref = (-1) ** (l + m - s) * math.sqrt(
        factorial(l + m) * factorial(l - m) * (2 * l + 1) / (4 * math.pi * factorial(l + s) * factorial(l - s)))
    st, ct = math.sin(theta / 2), math.cos(theta / 2)
    tot = 0.0
    for r in range(0, l - s + 1):
        k = r + s - m
        if k < 0 or k > l + s:
            continue
        tot += (-1) ** r * comb(l - s, r) * comb(l + s, k) * st ** (2 * l - 2 * r - s + m) * ct ** (2 * r + s - m)
    return pref * tot * complex(math.cos(m * phi), math.sin(m * phi))


def covariance(angles, ctt, cte, cee, cbb, lmax, bT=None, bP=None):
    """Dense (3n x 3n) covariance of [T_0..T_{n-1}, Q_0.., U_0..] at the (theta, phi) pairs in `angles`.

    Spectra are indexed by l (entries below l=2 ignored, as reference c_matrix_generator.cpp:190);
    bT / bP are optional per-l window*beam factors for temperature / polarization.
    """
    n = len(angles)
    bT = np.ones(lmax + 1) if bT is None else np.asarray(bT)
    bP = np.ones(lmax + 1) if bP is None else np.asarray(bP)
    TT = np.zeros((n, n))
    A = np.zeros((n, n), complex)      # <P_i P_j^*>
    B = np.zeros((n, n), complex)      # <P_i P_j>
    TP = np.zeros((n, n), complex)     # <T_i P_j^*>
    for l in range(2, lmax + 1):
        for m in range(-l, l + 1):
            y0 = np.array([sYlm(0, l, m, t, p) for t, p in angles])
            y2 = np.array([sYlm(2, l, m, t, p) for t, p in angles])
            ym2 = np.array([sYlm(-2, l, m, t, p) for t, p in angles])
            TT += ctt[l] * bT[l] ** 2 * np.real(np.outer(y0, y0.conj()))
            A += (cee[l] + cbb[l]) * bP[l] ** 2 * np.outer(y2, y2.conj())
            B += (cee[l] - cbb[l]) * bP[l] ** 2 * np.outer(y2, ym2.conj())
            TP += -cte[l] * bT[l] * bP[l] * np.outer(y0, y2.conj())
    C = np.zeros((3 * n, 3 * n))
    C[:n, :n] = TT
    C[:n, n:2 * n] = TP.real
    C[:n, 2 * n:] = -TP.imag
    C[n:2 * n, :n] = TP.real.T
    C[2 * n:, :n] = -TP.imag.T
    C[n:2 * n, n:2 * n] = (A + B).real / 2
    C[2 * n:, 2 * n:] = (A - B).real / 2
    C[n:2 * n, 2 * n:] = (B - A).imag / 2
    C[2 * n:, n:2 * n] = (A + B).imag / 2
    return C


def covariance_from_modes(angles, ctt, cte, cee, cbb, lmax):
    """The same covariance built the most literal way: the linear map J from the independent real
    Gaussian mode amplitudes to (T,Q,U), then J Sigma J^T.  Checks the complex-algebra identities
    used in covariance() (tests only; O(n * lmax^2) columns)."""
    n = len(angles)
    cols, var = [], []

    def field(aT, aE, aB, l, m):
        """(T,Q,U) at all points for a single mode pair (l, +-m) with the reality condition."""
        out = np.zeros(3 * n)
        for k, (t, p) in enumerate(angles):
            T = 0j
            Pp = 0j     # Q + iU
            Pm = 0j     # Q - iU
            for mm, cT, cE, cB in ([(0, aT, aE, aB)] if m == 0 else
                                   [(m, aT, aE, aB), (-m, (-1) ** m * np.conj(aT), (-1) ** m * np.conj(aE), (-1) ** m * np.conj(aB))]):
                T += cT * sYlm(0, l, mm, t, p)
                Pp += -(cE + 1j * cB) * sYlm(2, l, mm, t, p)
                Pm += -(cE - 1j * cB) * sYlm(-2, l, mm, t, p)
            assert abs(T.imag) < 1e-12 and abs((Pp - np.conj(Pm))) < 1e-12
            out[k] = T.real
            out[n + k] = ((Pp + Pm) / 2).real
            out[2 * n + k] = ((Pp - Pm) / 2j).real
        return out

    for l in range(2, lmax + 1):
        # correlated (T,E) pair: aT = sqrt(ctt) x1, aE = cte/sqrt(ctt) x1 + sqrt(cee - cte^2/ctt) x2 ; aB = sqrt(cbb) x3
        s1 = math.sqrt(ctt[l]); r = cte[l] / s1; s2 = math.sqrt(cee[l] - r * r); s3 = math.sqrt(cbb[l])
        for m in range(0, l + 1):
            parts = [1.0] if m == 0 else [math.sqrt(0.5), 1j * math.sqrt(0.5)]   # real / imaginary unit-variance parts
            for u in parts:
                cols.append(field(s1 * u, r * u, 0, l, m)); var.append(1.0)
                cols.append(field(0, s2 * u, 0, l, m)); var.append(1.0)
                cols.append(field(0, 0, s3 * u, l, m)); var.append(1.0)
    J = np.array(cols).T
    return J @ J.T
