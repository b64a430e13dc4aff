/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference TT hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (cosmopp_b200/) never links or calls it.
 *
 * Each function cites the reference file:line it restates.  The arithmetic keeps the
 * reference's operation order (recurrence form, ascending-l accumulation, the
 * ((cl*leg)*beam)*beam product) so that it agrees with the reference object code in
 * oracle/_ref to rounding level; the only structural change is that P_l is carried
 * forward along l instead of being restarted from l=2 for every l (identical values,
 * O(lmax) instead of O(lmax^2) per pixel pair), and columns are spread over OpenMP threads.
 *
 * Parity pinned by: tests/test_oracle_tt.py (Legendre known answers of reference
 * source/test_legendre.cpp:27-51; whole matrices against oracle/_ref; committed goldens).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

void pix2ang_nest(long nside, long ipix, double* theta, double* phi);
long nside2npix(long nside);

#define ORC_PI 3.141592653589793   /* Math::pi, reference include/math_constants.hpp:8 */

/* Math::Legendre::calculate, reference include/legendre.hpp:26-37 */
double orc_legendre(unsigned int l, double x)
{
    double pm2 = 1.0, pm1 = x, p;
    unsigned int l1;
    if (l == 0) return 1.0;
    if (l == 1) return x;
    for (l1 = 2; l1 <= l; ++l1) {
        p = (2 - 1.0 / l1) * x * pm1 - (1 - 1.0 / l1) * pm2;
        pm2 = pm1;
        pm1 = p;
    }
    return pm1;
}

/* Utils::beamFunction, reference source/utils.cpp:54-64 (l*(l+1) is an int product there) */
double orc_beam_function(int l, double fwhm)
{
    double sigma;
    if (fwhm == 0) return 1.0;
    sigma = sqrt(8 * log(2.0)) / (fwhm * ORC_PI / 180);
    return exp(-l * (l + 1) / (2 * sigma * sigma));
}

/* Utils::readPixelWindowFunction, reference source/utils.cpp:154-160: f[l] = pixwin[l] * beam(l).
 * pixwin == NULL means window == 1 (the FITS table of HEALPix is not available offline). */
void orc_window_beam(double* f, int lmax, double fwhm, const double* pixwin)
{
    int l;
    for (l = 0; l <= lmax; ++l) {
        f[l] = pixwin ? pixwin[l] : 1.0;
        f[l] *= orc_beam_function(l, fwhm);
    }
}

/* Utils::readMask threshold, reference source/utils.cpp:45-51: ascending indices with mask > 0.5 */
long orc_good_pixels_from_mask(const double* mask, long npix, int* good)
{
    long i, n = 0;
    for (i = 0; i < npix; ++i)
        if (mask[i] > 0.5) good[n++] = (int)i;
    return n;
}

/* unit vectors as the generator forms them, reference source/c_matrix_generator.cpp:178-185 */
void orc_unit_vectors(long nside, const int* good, long n, double* xyz)
{
    long i;
    for (i = 0; i < n; ++i) {
        double theta, phi;
        long index = good ? good[i] : i;
        pix2ang_nest(nside, index, &theta, &phi);
        xyz[3 * i + 0] = sin(theta) * cos(phi);
        xyz[3 * i + 1] = sin(theta) * sin(phi);
        xyz[3 * i + 2] = cos(theta);
    }
}

/* CMatrix::getIndex, reference source/c_matrix.cpp:27-39, widened to 64 bit */
long long orc_packed_index(long long i, long long j)
{
    if (i > j) { long long t = i; i = j; j = t; }
    return j * (j + 1) / 2 + i;
}

static double clamped_dot(const double* a, const double* b)
{
    /* ThreeVector::operator*, reference include/three_vector.hpp:37; clamp: c_matrix_generator.cpp:205-215 */
    double dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    if (dot > 1) dot = 1;
    if (dot < -1) dot = -1;
    return dot;
}

/* sum_{l=lfirst}^{llast} coef[l] * P_l(dot) * beam[l] * beam[l], ascending l,
 * product order of reference source/c_matrix_generator.cpp:222 / :758 */
static double legendre_sum(const double* coef, const double* beam, int lfirst, int llast, double dot)
{
    double pm2 = 1.0, pm1 = dot, element = 0;
    int l;
    for (l = 2; l <= llast; ++l) {
        double p = (2 - 1.0 / l) * dot * pm1 - (1 - 1.0 / l) * pm2;
        pm2 = pm1;
        pm1 = p;
        if (l >= lfirst) element += coef[l] * p * beam[l] * beam[l];
    }
    return element;
}

/* CMatrixGenerator::clToCMatrix, reference source/c_matrix_generator.cpp:164-232, restricted to
 * columns [j_begin, j_end) of the packed triangle.  out points at the entry (0, j_begin). */
int orc_cl_to_cmatrix_cols(const double* cl, int lmax, long nside, double fwhm, const double* pixwin,
                           const int* good, long ngood, long j_begin, long j_end, double* out)
{
    const long npix = good ? ngood : nside2npix(nside);
    double* beam = (double*)malloc(sizeof(double) * (size_t)(lmax + 1));
    double* clcopy = (double*)calloc((size_t)(lmax + 1), sizeof(double));
    double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)npix);
    const long long base = orc_packed_index(0, j_begin);
    long j;
    int l;
    if (!beam || !clcopy || !xyz) { free(beam); free(clcopy); free(xyz); return 1; }
    orc_window_beam(beam, lmax, fwhm, pixwin);
    orc_unit_vectors(nside, good, npix, xyz);
    for (l = 2; l <= lmax; ++l) clcopy[l] = cl[l] * (2 * l + 1) / (4 * ORC_PI);   /* :190-193 */

#pragma omp parallel for schedule(dynamic, 4)
    for (j = j_begin; j < j_end; ++j) {
        long i;
        double* col = out + (orc_packed_index(0, j) - base);
        for (i = 0; i <= j; ++i)
            col[i] = legendre_sum(clcopy, beam, 2, lmax, clamped_dot(xyz + 3 * i, xyz + 3 * j));
    }
    free(beam); free(clcopy); free(xyz);
    return 0;
}

int orc_cl_to_cmatrix(const double* cl, int lmax, long nside, double fwhm, const double* pixwin,
                      const int* good, long ngood, double* out_packed)
{
    const long npix = good ? ngood : nside2npix(nside);
    return orc_cl_to_cmatrix_cols(cl, lmax, nside, fwhm, pixwin, good, ngood, 0, npix, out_packed);
}

/* the reference's literal O(lmax^2) loop (Legendre restarted for every l), single thread:
 * used only to time the reference algorithm when oracle/_ref is not available */
int orc_cl_to_cmatrix_cols_literal(const double* cl, int lmax, long nside, double fwhm, const double* pixwin,
                                   const int* good, long ngood, long j_begin, long j_end, double* out)
{
    const long npix = good ? ngood : nside2npix(nside);
    double* beam = (double*)malloc(sizeof(double) * (size_t)(lmax + 1));
    double* clcopy = (double*)calloc((size_t)(lmax + 1), sizeof(double));
    double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)npix);
    const long long base = orc_packed_index(0, j_begin);
    long j, i;
    int l;
    if (!beam || !clcopy || !xyz) { free(beam); free(clcopy); free(xyz); return 1; }
    orc_window_beam(beam, lmax, fwhm, pixwin);
    orc_unit_vectors(nside, good, npix, xyz);
    for (l = 2; l <= lmax; ++l) clcopy[l] = cl[l] * (2 * l + 1) / (4 * ORC_PI);
    for (j = j_begin; j < j_end; ++j)
        for (i = 0; i <= j; ++i) {
            const double dot = clamped_dot(xyz + 3 * i, xyz + 3 * j);
            double element = 0;
            for (l = 2; l <= lmax; ++l)
                element += clcopy[l] * orc_legendre((unsigned)l, dot) * beam[l] * beam[l];
            out[orc_packed_index(i, j) - base] = element;
        }
    free(beam); free(clcopy); free(xyz);
    return 0;
}

/* TT entries for an explicit list of pixel pairs (indices into the good-pixel list) */
int orc_cl_to_cmatrix_pairs(const double* cl, int lmax, long nside, double fwhm, const double* pixwin,
                            const int* good, long ngood, const long long* pi, const long long* pj, long npairs,
                            double* out)
{
    const long npix = good ? ngood : nside2npix(nside);
    double* beam = (double*)malloc(sizeof(double) * (size_t)(lmax + 1));
    double* clcopy = (double*)calloc((size_t)(lmax + 1), sizeof(double));
    double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)npix);
    long k;
    int l;
    if (!beam || !clcopy || !xyz) { free(beam); free(clcopy); free(xyz); return 1; }
    orc_window_beam(beam, lmax, fwhm, pixwin);
    orc_unit_vectors(nside, good, npix, xyz);
    for (l = 2; l <= lmax; ++l) clcopy[l] = cl[l] * (2 * l + 1) / (4 * ORC_PI);
#pragma omp parallel for schedule(static)
    for (k = 0; k < npairs; ++k)
        out[k] = legendre_sum(clcopy, beam, 2, lmax, clamped_dot(xyz + 3 * pi[k], xyz + 3 * pj[k]));
    free(beam); free(clcopy); free(xyz);
    return 0;
}

/* CMatrixGenerator::getFiducialMatrix, reference source/c_matrix_generator.cpp:705-772:
 * sum_{l=lmax+1}^{4 nside} cl[l]((2l+1)/4pi) P_l B_l^2  +  100 cl[2] (1+dot) B_2^2 ; cl holds 4*nside+1 values */
int orc_fiducial_matrix(const double* cl, long nside, int lmax, double fwhm, const double* pixwin,
                        const int* good, long ngood, double* out_packed)
{
    const int lmaxmax = (int)(4 * nside);
    const long npix = good ? ngood : nside2npix(nside);
    double* beam = (double*)malloc(sizeof(double) * (size_t)(lmaxmax + 1));
    double* coef = (double*)calloc((size_t)(lmaxmax + 1), sizeof(double));
    double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)npix);
    long j;
    int l;
    if (!beam || !coef || !xyz) { free(beam); free(coef); free(xyz); return 1; }
    orc_window_beam(beam, lmaxmax, fwhm, pixwin);
    orc_unit_vectors(nside, good, npix, xyz);
    for (l = lmax + 1; l <= lmaxmax; ++l) coef[l] = cl[l] * ((2 * l + 1) / (4 * ORC_PI));   /* :758 */

#pragma omp parallel for schedule(dynamic, 4)
    for (j = 0; j < npix; ++j) {
        long i;
        double* col = out_packed + orc_packed_index(0, j);
        for (i = 0; i <= j; ++i) {
            const double dot = clamped_dot(xyz + 3 * i, xyz + 3 * j);
            double element = legendre_sum(coef, beam, lmax + 1, lmaxmax, dot);
            element += 100 * cl[2] * (1 + dot) * beam[2] * beam[2];                         /* :762 */
            col[i] = element;
        }
    }
    free(beam); free(coef); free(xyz);
    return 0;
}

/* CMatrixGenerator::generateNoiseMatrix, reference source/c_matrix_generator.cpp:774-787 (full sky) */
void orc_noise_matrix(long nside, double noise, double* out_packed)
{
    const long npix = nside2npix(nside);
    long i;
    memset(out_packed, 0, sizeof(double) * (size_t)(npix * (npix + 1) / 2));
    for (i = 0; i < npix; ++i) out_packed[orc_packed_index(i, i)] = noise * noise;
}

/* CMatrix::maskMatrix(const std::vector<int>&), reference source/c_matrix.cpp:182-201 */
void orc_mask_matrix(const double* in_packed, const int* good, long ngood, double* out_packed)
{
    long i, j;
    for (j = 0; j < ngood; ++j)
        for (i = 0; i <= j; ++i)
            out_packed[orc_packed_index(i, j)] = in_packed[orc_packed_index(good[i], good[j])];
}

/* LegendrePolynomialContainer data, reference source/c_matrix_generator.cpp:30-76:
 * out[l][packed(i,j)] = P_l(clamped n_i.n_j), l = 0..lmax (file order of :149-160) */
int orc_legendre_container(int lmax, long nside, const int* good, long ngood, double* out)
{
    const long npix = good ? ngood : nside2npix(nside);
    const long long tri = (long long)npix * (npix + 1) / 2;
    double* xyz = (double*)malloc(sizeof(double) * 3 * (size_t)npix);
    long i, j;
    int l;
    if (!xyz) return 1;
    orc_unit_vectors(nside, good, npix, xyz);
    for (j = 0; j < npix; ++j)
        for (i = 0; i <= j; ++i) {
            const double dot = clamped_dot(xyz + 3 * i, xyz + 3 * j);
            const long long k = orc_packed_index(i, j);
            for (l = 0; l <= lmax; ++l) out[(long long)l * tri + k] = orc_legendre((unsigned)l, dot);
        }
    free(xyz);
    return 0;
}
