/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Minimal HEALPix pixel-centre routines for the oracle and for the reference
 * object code built into oracle/_ref (the reference links chealpix 3.20, which
 * is not vendored under /root/reference: cmake_settings_template.txt:23).
 * Restated from the published algorithm (Gorski et al. 2005, ApJ 622, 759,
 * section 4 and appendix), not from HEALPix source.
 *
 * Call sites in the reference that these stand in for:
 *   nside2npix   source/c_matrix_generator.cpp:34,170,711,777  source/c_matrix.cpp:169
 *   pix2ang_nest source/c_matrix_generator.cpp:42,182,722
 * Parity of pixel centres is pinned by HEALPix invariants (tests/test_oracle_healpix.py),
 * because the reference holds no direct test of chealpix.
 */
#include <math.h>

static const double ORC_PI = 3.141592653589793238462643383279502884;

long nside2npix(long nside) { return 12 * nside * nside; }

/* x <- even bits of v, y <- odd bits of v (v < nside^2 <= 2^26 here, loops are fine for an oracle) */
static void deinterleave(long v, long* x, long* y)
{
    long xx = 0, yy = 0;
    int b;
    for (b = 0; b < 31; ++b) {
        xx |= ((v >> (2 * b)) & 1L) << b;
        yy |= ((v >> (2 * b + 1)) & 1L) << b;
    }
    *x = xx;
    *y = yy;
}

void pix2ang_nest(long nside, long ipix, double* theta, double* phi)
{
    static const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
    static const int jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
    const long npface = nside * nside;
    const long nl4 = 4 * nside;
    const long face = ipix / npface;
    long ix, iy, jr, nr, kshift, jp;
    double z;

    deinterleave(ipix % npface, &ix, &iy);
    jr = jrll[face] * nside - ix - iy - 1;      /* ring index counted from the north pole, 1..4nside-1 */

    if (jr < nside) {                           /* north polar cap */
        nr = jr;
        z = 1.0 - (double)(nr * nr) / (3.0 * (double)nside * (double)nside);
        kshift = 0;
    } else if (jr > 3 * nside) {                /* south polar cap */
        nr = nl4 - jr;
        z = -1.0 + (double)(nr * nr) / (3.0 * (double)nside * (double)nside);
        kshift = 0;
    } else {                                    /* equatorial belt */
        nr = nside;
        z = (double)(2 * nside - jr) * 2.0 / (3.0 * (double)nside);
        kshift = (jr - nside) & 1;
    }

    jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
    if (jp > nl4) jp -= nl4;
    if (jp < 1) jp += nl4;

    *theta = acos(z);
    *phi = ((double)jp - (double)(kshift + 1) * 0.5) * ((0.5 * ORC_PI) / (double)nr);
}

void pix2ang_ring(long nside, long ipix, double* theta, double* phi)
{
    const long npix = 12 * nside * nside;
    const long ncap = 2 * nside * (nside - 1);
    double z;
    if (ipix < ncap) {                          /* north cap: ring r has 4r pixels, first index 2r(r-1) */
        long r = (long)(0.5 * (1.0 + sqrt(1.0 + 2.0 * (double)ipix)));
        while (2 * r * (r - 1) > ipix) --r;
        while (2 * r * (r + 1) <= ipix) ++r;
        {
            long k = ipix - 2 * r * (r - 1);    /* 0-based position in ring */
            z = 1.0 - (double)(r * r) / (3.0 * (double)nside * (double)nside);
            *phi = ((double)k + 0.5) * (0.5 * ORC_PI) / (double)r;
        }
    } else if (ipix < npix - ncap) {            /* belt: 4nside pixels per ring */
        long ip = ipix - ncap;
        long r = ip / (4 * nside) + nside;      /* ring index from north, nside..3nside */
        long k = ip % (4 * nside);
        double shift = ((r - nside) & 1) ? 0.0 : 0.5;   /* kshift=1 -> whole-step centres */
        z = (double)(2 * nside - r) * 2.0 / (3.0 * (double)nside);
        *phi = ((double)k + shift) * (0.5 * ORC_PI) / (double)nside;
    } else {                                    /* south cap, mirrored */
        long ip = npix - 1 - ipix;
        long r = (long)(0.5 * (1.0 + sqrt(1.0 + 2.0 * (double)ip)));
        while (2 * r * (r - 1) > ip) --r;
        while (2 * r * (r + 1) <= ip) ++r;
        {
            long k = 4 * r - 1 - (ip - 2 * r * (r - 1));
            z = -1.0 + (double)(r * r) / (3.0 * (double)nside * (double)nside);
            *phi = ((double)k + 0.5) * (0.5 * ORC_PI) / (double)r;
        }
    }
    *theta = acos(z);
}
