#include "series.hpp"

#include <cmath>
#include <cstdlib>

namespace cmg
{

// Standard recurrence in l of the Wigner small-d functions (e.g. Varshalovich et al., 4.8.2):
//   l sqrt(((l+1)^2-m^2)((l+1)^2-m'^2)) d^{l+1} = (2l+1)(l(l+1) z - m m') d^l - (l+1) sqrt((l^2-m^2)(l^2-m'^2)) d^{l-1}
// written as d^{l+1} = alpha_l (z - mu_l) d^l - beta_l d^{l-1}.  With d^l = N_l phi_l and
// N_{l+1} = alpha_l N_l / 2 the phi recurrence gets the unit coefficient on 2z and
// g_l = beta_l N_{l-1} / N_{l+1}.
SeriesTable makeSeriesTable(int lmax, int m, int mp)
{
    SeriesTable t;
    t.l0 = std::max(std::abs(m), std::abs(mp));
    const int n = lmax + 2;
    t.N.assign(n, 0.0);
    t.g.assign(n, 0.0);
    t.c.assign(n, 0.0);

    std::vector<long double> alpha(n, 0.0L), beta(n, 0.0L), norm(n + 1, 0.0L);
    const long double m2 = static_cast<long double>(m) * m, mp2 = static_cast<long double>(mp) * mp;
    for(int l = t.l0; l < n; ++l)
    {
        const long double ll = l, l1 = l + 1;
        const long double den = std::sqrt((l1 * l1 - m2) * (l1 * l1 - mp2));
        alpha[l] = (2 * ll + 1) * l1 / den;
        if(l > 0)
        {
            beta[l] = l1 * std::sqrt((ll * ll - m2) * (ll * ll - mp2)) / (ll * den);
            t.c[l] = static_cast<double>(2.0L * m * mp / (ll * l1));
        }
    }
    norm[t.l0] = 1.0L;
    for(int l = t.l0; l < n; ++l)
        norm[l + 1] = alpha[l] * norm[l] / 2;
    for(int l = t.l0; l < n; ++l)
        t.N[l] = static_cast<double>(norm[l]);
    for(int l = t.l0 + 1; l < n; ++l)
        t.g[l] = static_cast<double>(beta[l] * norm[l - 1] / norm[l + 1]);
    return t;
}

} // namespace cmg
