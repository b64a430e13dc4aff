// Host-side HEALPix NESTED pixel centres for the generator's geometry set-up.
// Stands in for chealpix nside2npix / pix2ang_nest at the reference call sites
// source/c_matrix_generator.cpp:34,42,170,182,711,722 (HEALPix itself is not a dependency here).
// Algorithm: Gorski et al. 2005 (ApJ 622, 759), section 4 -- base-pixel face, (x,y) inside the face
// from the bit-interleaved index, ring number, then z and phi of the ring position.
#pragma once
#include <cstdint>

namespace cmg
{

inline int64_t nside2npix(int64_t nside) { return 12 * nside * nside; }

// true for a power of two in [1, 2^13] (NESTED needs a power of two; 8192 keeps 12 nside^2 in int32
// like the reference's pixel lists)
bool validNside(int64_t nside);

// colatitude theta in [0, pi], longitude phi in [0, 2 pi)
void pix2angNest(int64_t nside, int64_t ipix, double& theta, double& phi);

// Per-pixel geometry the kernels keep resident: unit vector n, and the local polarization basis
// e_theta = (cos t cos p, cos t sin p, -sin t), e_phi = (-sin p, cos p, 0).
// n is formed exactly as reference source/c_matrix_generator.cpp:183 (sin(theta)*cos(phi), ...).
struct PixelFrame
{
    double n[3];
    double eTheta[3];
    double ePhi[2];
};
PixelFrame pixelFrame(int64_t nside, int64_t ipixNest);

} // namespace cmg
