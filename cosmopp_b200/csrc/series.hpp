// Three-term-recurrence tables for evaluating sums over l of Legendre / Wigner-d functions with
// Clenshaw's algorithm in a normalisation whose recurrence has a unit coefficient on 2z:
//
//      phi_{l+1}(z) = (2z - c_l) phi_l(z) - g_l phi_{l-1}(z),      d^l_{m m'}(beta) = N_l phi_l(z),  z = cos(beta)
//
// so that one Clenshaw step costs two FMAs when c_l = 0 (m m' = 0: P_l and d^l_20) and three otherwise
// (d^l_22, d^l_2-2; c_l flips sign with m').  A series sum_l a_l d^l is then
//      b_k = a_k N_k + (2z - c_k) b_{k+1} - g_{k+1} b_{k+2},   k = lmax .. l0,     sum = phi_{l0}(z) b_{l0}
// with l0 = max(|m|,|m'|) and phi_{l0} = d^{l0} in closed form.  Replaces the per-l restart of
// Math::Legendre::calculate (reference include/legendre.hpp:26-37) inside the pair loop of
// reference source/c_matrix_generator.cpp:219-223.
#pragma once
#include <vector>

namespace cmg
{

struct SeriesTable
{
    int l0;                  // first l of the family
    std::vector<double> N;   // N[l], l = 0..lmax+1 (0 below l0)
    std::vector<double> g;   // g[l], l = 0..lmax+1 (0 at and below l0)
    std::vector<double> c;   // c[l] = 2 m m' / (l (l+1)), 0 for l = 0
};

// tables for d^l_{m mp}, l up to lmax (inclusive), computed in long double and rounded once
SeriesTable makeSeriesTable(int lmax, int m, int mp);

} // namespace cmg
