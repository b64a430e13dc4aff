/* test infrastructure: the two chealpix entry points the TT path calls
 * (reference source/c_matrix_generator.cpp:34,42,170,182,711,722; source/c_matrix.cpp:169).
 * Implemented in oracle/healpix_min.c from the published HEALPix algorithm. */
#ifndef ORACLE_STUB_CHEALPIX_H
#define ORACLE_STUB_CHEALPIX_H
#ifdef __cplusplus
extern "C" {
#endif
long nside2npix(long nside);
void pix2ang_nest(long nside, long ipix, double* theta, double* phi);
void pix2ang_ring(long nside, long ipix, double* theta, double* phi);
#ifdef __cplusplus
}
#endif
#endif
