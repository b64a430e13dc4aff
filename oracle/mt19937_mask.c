/* TEST INFRASTRUCTURE ONLY.
 * The deterministic mask of the reference's only C-matrix test, restated from
 * reference source/test_like_low.cpp:99-118 and source/utils.cpp:238-253 (Utils::maskRegions):
 * 25 discs with theta ~ U(pi/50, pi - pi/50) [seed 1000000], phi ~ U(0, 2pi) [seed 1000001],
 * radius ~ U(pi/60, pi/40) [seed 1000002], plus the band |theta - pi/2| < pi/20.
 * Math::UniformRealGenerator (reference include/random.hpp:10-28) is std::mt19937 feeding
 * std::uniform_real_distribution<double>; libstdc++ draws two 32-bit words per double
 * (generate_canonical<double,53>), low word first.  tests/test_oracle_mask.py checks this
 * restatement against a g++-compiled libstdc++ snippet.
 */
#include <math.h>
#include <stdint.h>

void pix2ang_nest(long nside, long ipix, double* theta, double* phi);
long nside2npix(long nside);

#define ORC_PI 3.141592653589793

typedef struct { uint32_t mt[624]; int idx; } orc_mt;

void orc_mt_seed(orc_mt* g, uint32_t seed)
{
    int i;
    g->mt[0] = seed;
    for (i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt* g)
{
    uint32_t y;
    if (g->idx >= 624) {
        int k;
        for (k = 0; k < 624; ++k) {
            uint32_t v = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
            g->mt[k] = g->mt[(k + 397) % 624] ^ (v >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ uniform_real_distribution<double>(a, b)(mt19937) */
double orc_uniform_real(orc_mt* g, double a, double b)
{
    const double r = 4294967296.0;
    double lo = (double)orc_mt_next(g);
    double hi = (double)orc_mt_next(g);
    double u = (lo + hi * r) / (r * r);
    if (u >= 1.0) u = nextafter(1.0, 0.0);
    return u * (b - a) + a;
}

/* fills draws[3*k + {0,1,2}] = (theta, phi, radius) of disc k; for tests */
void orc_like_low_discs(int nregions, uint32_t seed1, double* draws)
{
    orc_mt gt, gp, ga;
    int k;
    orc_mt_seed(&gt, seed1);
    orc_mt_seed(&gp, seed1 + 1);
    orc_mt_seed(&ga, seed1 + 2);
    for (k = 0; k < nregions; ++k) {
        draws[3 * k + 0] = orc_uniform_real(&gt, ORC_PI / 50, ORC_PI - ORC_PI / 50);
        draws[3 * k + 1] = orc_uniform_real(&gp, 0, 2 * ORC_PI);
        draws[3 * k + 2] = orc_uniform_real(&ga, ORC_PI / 60, ORC_PI / 40);
    }
}

/* mask[i] in {0,1} for the NESTED map of this nside; returns the number of good pixels */
long orc_like_low_mask(long nside, double* mask)
{
    const long npix = nside2npix(nside);
    double d[75];
    long i, ngood = 0;
    int k;
    orc_like_low_discs(25, 1000000u, d);
    for (i = 0; i < npix; ++i) {
        double theta, phi, v[3];
        mask[i] = 1;
        pix2ang_nest(nside, i, &theta, &phi);
        v[0] = sin(theta) * cos(phi); v[1] = sin(theta) * sin(phi); v[2] = cos(theta);
        for (k = 0; k < 25; ++k) {
            const double t = d[3 * k], p = d[3 * k + 1];
            const double c[3] = { sin(t) * cos(p), sin(t) * sin(p), cos(t) };
            const double dot = v[0] * c[0] + v[1] * c[1] + v[2] * c[2];
            if (dot > cos(d[3 * k + 2])) mask[i] = 0;
        }
        if (theta < ORC_PI / 2 + ORC_PI / 20 && theta > ORC_PI / 2 - ORC_PI / 20) mask[i] = 0;
        if (mask[i] > 0.5) ++ngood;
    }
    return ngood;
}
