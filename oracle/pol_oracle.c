/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the polarized (T,Q,U) pixel covariance.
 *
 * PARITY UNPINNED BY THE REFERENCE: cosmopp has no TE/BB pixel generator; its only polarization
 * routine (reference source/c_matrix_generator.cpp:485-695, EE only, HEALPix alm2map_pol based) is
 * never called by any test and needs libsharp, which is absent here.  This oracle therefore follows
 * the published Tegmark & de Oliveira-Costa (2001) formulation that BASELINE.json names, and is
 * itself pinned against the definition-level brute-force sum over spin-weighted harmonics in
 * oracle/pol_bruteforce.py (tests/test_oracle_pol.py).  Conventions (HEALPix primer): Q,U in the
 * local (e_theta, e_phi) basis, Q +- iU = sum a_{+-2,lm} {+-2}Y_lm, a_{+-2,lm} = -(aE_lm +- i aB_lm).
 * Layout follows the reference's [Q;U] convention (c_matrix_generator.cpp:678-681) extended to
 * [T_0..T_{N-1}, Q_0.., U_0..], packed upper triangle index j(j+1)/2+i (source/c_matrix.cpp:36).
 *
 * Deliberately written differently from the CUDA product (which uses division-free polynomial
 * rotation factors and Clenshaw sums): here the Wigner-d functions are run FORWARD in l in long
 * double, and the frame rotation uses explicit angles from atan2 and cos/sin.
 *
 *   TT      = sum w_l C^TT_l P_l(z)                               w_l = (2l+1)/4pi * window factors
 *   X_T     = sum w_l C^TE_l d^l_{20}(beta)        (= F^10)
 *   X_P     = sum w_l (C^EE_l + C^BB_l) d^l_{22}   (= F^12 - F^22)
 *   X_M     = sum w_l (C^EE_l - C^BB_l) d^l_{2,-2} (= F^12 + F^22)
 *   <P_i P_j*> = X_P e^{2i(psi_i - psi_j)},  <P_i P_j> = X_M e^{2i(psi_i + psi_j)},
 *   <T_i P_j*> = -X_T e^{-2i psi_j},   P = Q + iU,
 *   psi_i = angle of the great-circle direction towards j, measured from e_theta(i) towards e_phi(i).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

void pix2ang_nest(long nside, long ipix, double* theta, double* phi);
long nside2npix(long nside);
void orc_window_beam(double* f, int lmax, double fwhm, const double* pixwin);
long long orc_packed_index(long long i, long long j);

#define ORC_PIL 3.141592653589793238462643383279502884L

typedef struct {
    long double n[3], et[3], ep[3];
} orc_pix;

static void make_pix(long nside, long index, orc_pix* p)
{
    double theta, phi;
    long double st, ct, sp, cp;
    pix2ang_nest(nside, index, &theta, &phi);
    /* the unit vector itself is formed in double like the reference does (c_matrix_generator.cpp:183) */
    p->n[0] = (long double)(sin(theta) * cos(phi));
    p->n[1] = (long double)(sin(theta) * sin(phi));
    p->n[2] = (long double)cos(theta);
    st = sinl((long double)theta); ct = cosl((long double)theta);
    sp = sinl((long double)phi);   cp = cosl((long double)phi);
    p->et[0] = ct * cp; p->et[1] = ct * sp; p->et[2] = -st;
    p->ep[0] = -sp;     p->ep[1] = cp;      p->ep[2] = 0;
}

static long double dot3(const long double* a, const long double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* forward recurrence for d^l_{m m'}(beta), z = cos(beta), l = 2..lmax, into d[l] (d[0]=d[1]=0 for m=2):
 * l sqrt(((l+1)^2-m^2)((l+1)^2-m'^2)) d^{l+1} = (2l+1)(l(l+1) z - m m') d^l - (l+1) sqrt((l^2-m^2)(l^2-m'^2)) d^{l-1} */
static void wigner_d2(int lmax, int mp, long double z, long double* d)
{
    const int m = 2;
    int l;
    d[0] = 0; d[1] = 0;
    if (lmax < 2) return;
    if (mp == 0) d[2] = sqrtl(6.0L) / 4 * (1 - z * z);
    else if (mp == 2) d[2] = (1 + z) * (1 + z) / 4;
    else d[2] = (1 - z) * (1 - z) / 4;
    for (l = 2; l < lmax; ++l) {
        const long double ll = l, l1 = l + 1;
        const long double den = ll * sqrtl((l1 * l1 - m * m) * (l1 * l1 - (long double)(mp * mp)));
        const long double a = (2 * ll + 1) * (ll * l1 * z - (long double)(m * mp));
        const long double b = l1 * sqrtl((ll * ll - m * m) * (ll * ll - (long double)(mp * mp)));
        d[l + 1] = (a * d[l] - b * d[l - 1]) / den;
    }
}

static void legendre_all(int lmax, long double z, long double* p)
{
    int l;
    p[0] = 1;
    if (lmax >= 1) p[1] = z;
    for (l = 2; l <= lmax; ++l) p[l] = ((2 * l - 1) * z * p[l - 1] - (l - 1) * p[l - 2]) / l;
}

typedef struct {
    int lmax;
    long double *wtt, *wte, *wp, *wm;   /* l-weights incl. (2l+1)/4pi and window/beam factors */
} orc_weights;

static int make_weights(orc_weights* w, const double* ctt, const double* cte, const double* cee, const double* cbb,
                        int lmax, double fwhm, const double* pixwinT, const double* pixwinP)
{
    double* bt = (double*)malloc(sizeof(double) * (size_t)(lmax + 1));
    double* bp = (double*)malloc(sizeof(double) * (size_t)(lmax + 1));
    int l;
    w->lmax = lmax;
    w->wtt = (long double*)calloc((size_t)(lmax + 1) * 4, sizeof(long double));
    if (!bt || !bp || !w->wtt) { free(bt); free(bp); free(w->wtt); return 1; }
    w->wte = w->wtt + (lmax + 1); w->wp = w->wte + (lmax + 1); w->wm = w->wp + (lmax + 1);
    orc_window_beam(bt, lmax, fwhm, pixwinT);
    orc_window_beam(bp, lmax, fwhm, pixwinP);
    for (l = 2; l <= lmax; ++l) {
        const long double f = (2 * l + 1) / (4 * ORC_PIL);
        w->wtt[l] = f * ctt[l] * bt[l] * bt[l];
        w->wte[l] = f * cte[l] * bt[l] * bp[l];
        w->wp[l] = f * ((long double)cee[l] + cbb[l]) * bp[l] * bp[l];
        w->wm[l] = f * ((long double)cee[l] - cbb[l]) * bp[l] * bp[l];
    }
    free(bt); free(bp);
    return 0;
}

/* 3x3 block blk[a][b] = < X_a(i) X_b(j) >, X = (T,Q,U) */
static void pair_block(const orc_weights* w, const orc_pix* pi, const orc_pix* pj, int same,
                       long double* scratch, double blk[3][3])
{
    const int lmax = w->lmax;
    long double* pl = scratch;
    long double* d20 = pl + (lmax + 1);
    long double* d22 = d20 + (lmax + 1);
    long double* d2m = d22 + (lmax + 1);
    long double z = dot3(pi->n, pj->n);
    long double tt = 0, xt = 0, xp = 0, xm = 0, psi_i, psi_j;
    long double ai, bi, aj, bj;
    int l;
    if (same) z = 1;
    if (z > 1) z = 1;
    if (z < -1) z = -1;
    legendre_all(lmax, z, pl);
    wigner_d2(lmax, 0, z, d20);
    wigner_d2(lmax, 2, z, d22);
    wigner_d2(lmax, -2, z, d2m);
    for (l = 2; l <= lmax; ++l) {
        tt += w->wtt[l] * pl[l];
        xt += w->wte[l] * d20[l];
        xp += w->wp[l] * d22[l];
        xm += w->wm[l] * d2m[l];
    }
    ai = dot3(pj->n, pi->et); bi = dot3(pj->n, pi->ep);
    aj = dot3(pi->n, pj->et); bj = dot3(pi->n, pj->ep);
    if (same || ai * ai + bi * bi < 1e-24L) {
        /* coincident or antipodal centres: the great circle is arbitrary; take the meridian */
        psi_i = 0; psi_j = 0;
    } else {
        psi_i = atan2l(bi, ai);
        psi_j = atan2l(bj, aj);
    }
    {
        const long double cd = cosl(2 * (psi_i - psi_j)), sd = sinl(2 * (psi_i - psi_j));
        const long double cs = cosl(2 * (psi_i + psi_j)), ss = sinl(2 * (psi_i + psi_j));
        blk[0][0] = (double)tt;
        blk[0][1] = (double)(-xt * cosl(2 * psi_j));           /* T_i Q_j = Re <T_i P_j*> */
        blk[0][2] = (double)(-xt * sinl(2 * psi_j));           /* T_i U_j = -Im <T_i P_j*> */
        blk[1][0] = (double)(-xt * cosl(2 * psi_i));           /* Q_i T_j */
        blk[2][0] = (double)(-xt * sinl(2 * psi_i));           /* U_i T_j */
        blk[1][1] = (double)((xp * cd + xm * cs) / 2);         /* Q_i Q_j = Re(A+B)/2 */
        blk[2][2] = (double)((xp * cd - xm * cs) / 2);         /* U_i U_j = Re(A-B)/2 */
        blk[1][2] = (double)((xm * ss - xp * sd) / 2);         /* Q_i U_j = Im(B-A)/2 */
        blk[2][1] = (double)((xp * sd + xm * ss) / 2);         /* U_i Q_j = Im(A+B)/2 */
    }
}

/* blocks for an explicit list of pixel pairs; out[k][a][b] (9 doubles per pair) */
int orc_tqu_pairs(const double* ctt, const double* cte, const double* cee, const double* cbb, int lmax,
                  long nside, double fwhm, const double* pixwinT, const double* pixwinP,
                  const int* good, long ngood, const long long* pi, const long long* pj, long npairs, double* out)
{
    const long npix = good ? ngood : nside2npix(nside);
    orc_weights w;
    orc_pix* px = (orc_pix*)malloc(sizeof(orc_pix) * (size_t)npix);
    long k;
    int fail = 0;
    if (!px) return 1;
    if (make_weights(&w, ctt, cte, cee, cbb, lmax, fwhm, pixwinT, pixwinP)) { free(px); return 1; }
    for (k = 0; k < npix; ++k) make_pix(nside, good ? good[k] : k, px + k);
#pragma omp parallel
    {
        long double* scratch = (long double*)malloc(sizeof(long double) * 4 * (size_t)(lmax + 1));
        if (!scratch) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 64)
            for (k = 0; k < npairs; ++k) {
                double blk[3][3];
                pair_block(&w, px + pi[k], px + pj[k], pi[k] == pj[k], scratch, blk);
                memcpy(out + 9 * k, blk, sizeof(blk));
            }
            free(scratch);
        }
    }
    free(px); free(w.wtt);
    return fail;
}

/* whole packed [T;Q;U] matrix of dimension 3*npix */
int orc_tqu_matrix(const double* ctt, const double* cte, const double* cee, const double* cbb, int lmax,
                   long nside, double fwhm, const double* pixwinT, const double* pixwinP,
                   const int* good, long ngood, double* out_packed)
{
    const long npix = good ? ngood : nside2npix(nside);
    orc_weights w;
    orc_pix* px = (orc_pix*)malloc(sizeof(orc_pix) * (size_t)npix);
    long j;
    int fail = 0;
    if (!px) return 1;
    if (make_weights(&w, ctt, cte, cee, cbb, lmax, fwhm, pixwinT, pixwinP)) { free(px); return 1; }
    for (j = 0; j < npix; ++j) make_pix(nside, good ? good[j] : j, px + j);
#pragma omp parallel
    {
        long double* scratch = (long double*)malloc(sizeof(long double) * 4 * (size_t)(lmax + 1));
        if (!scratch) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (j = 0; j < npix; ++j) {
                long i;
                for (i = 0; i <= j; ++i) {
                    double b[3][3];
                    int a, c;
                    pair_block(&w, px + i, px + j, i == j, scratch, b);
                    for (a = 0; a < 3; ++a)
                        for (c = 0; c < 3; ++c) {
                            const long long row = a * npix + i, col = c * npix + j;
                            /* (row,col) with row<=col is stored directly; the T_jQ_i-type entries of the
                             * pair land in the upper triangle as (c*npix+j ... ) transposed partners */
                            if (row <= col) out_packed[orc_packed_index(row, col)] = b[a][c];
                            else if (i != j) out_packed[orc_packed_index(col, row)] = b[a][c];
                        }
                }
            }
            free(scratch);
        }
    }
    free(px); free(w.wtt);
    return fail;
}
