#include "healpix_nest.hpp"

#include <cmath>

namespace cmg
{

namespace
{
const double kPi = 3.141592653589793238462643383279502884;

// gather the even-position bits of v into the low half (Morton de-interleave by magic masks)
inline uint32_t compactEvenBits(uint64_t v)
{
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
    v = (v | (v >> 16)) & 0x00000000ffffffffull;
    return static_cast<uint32_t>(v);
}

// ring offset (in units of nside) and azimuthal offset of the twelve base faces
const int kFaceRing[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
const int kFacePhi[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
}

bool validNside(int64_t nside)
{
    return nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0;
}

void pix2angNest(int64_t nside, int64_t ipix, double& theta, double& phi)
{
    const int64_t perFace = nside * nside;
    const int face = static_cast<int>(ipix / perFace);
    const uint64_t inFace = static_cast<uint64_t>(ipix % perFace);
    const int64_t x = compactEvenBits(inFace);
    const int64_t y = compactEvenBits(inFace >> 1);

    const int64_t ring = kFaceRing[face] * nside - x - y - 1;      // 1 .. 4 nside - 1, north to south
    const double ns = static_cast<double>(nside);

    int64_t inRing;        // pixels per quarter of this ring
    int64_t halfShift;     // 1 when the ring's pixel centres sit on whole steps
    double z;
    if(ring < nside)
    {
        inRing = ring;
        halfShift = 0;
        z = 1.0 - static_cast<double>(inRing * inRing) / (3.0 * ns * ns);
    }
    else if(ring > 3 * nside)
    {
        inRing = 4 * nside - ring;
        halfShift = 0;
        z = -1.0 + static_cast<double>(inRing * inRing) / (3.0 * ns * ns);
    }
    else
    {
        inRing = nside;
        halfShift = (ring - nside) & 1;
        z = static_cast<double>(2 * nside - ring) * 2.0 / (3.0 * ns);
    }

    int64_t pos = (kFacePhi[face] * inRing + x - y + 1 + halfShift) / 2;   // 1-based position in ring
    if(pos > 4 * nside) pos -= 4 * nside;
    if(pos < 1) pos += 4 * nside;

    theta = std::acos(z);
    phi = (static_cast<double>(pos) - static_cast<double>(halfShift + 1) * 0.5) * ((0.5 * kPi) / static_cast<double>(inRing));
}

PixelFrame pixelFrame(int64_t nside, int64_t ipixNest)
{
    double theta, phi;
    pix2angNest(nside, ipixNest, theta, phi);
    const double st = std::sin(theta), ct = std::cos(theta);
    const double sp = std::sin(phi), cp = std::cos(phi);
    PixelFrame f;
    f.n[0] = st * cp;
    f.n[1] = st * sp;
    f.n[2] = ct;
    f.eTheta[0] = ct * cp;
    f.eTheta[1] = ct * sp;
    f.eTheta[2] = -st;
    f.ePhi[0] = -sp;
    f.ePhi[1] = cp;
    return f;
}

} // namespace cmg
