// Host half of the full-sky whole call when the matrix is wanted in HOST memory (cmg_cl_to_cmatrix_pol, cmg_cl_to_cmatrix with
// cmg_set_host_expand): the pi/2 rotation symmetry of the HEALPix grid (orbit.cuh) makes three quarters of the packed matrix
// copies of the rest, so only the columns of the LAST face of every ring of four (base faces 3, 7, 11) cross PCIe -- 27 % of the
// bytes, contiguous pieces of the packed triangle -- and the host fills in the other columns with block copies.
//
// Why those columns suffice: column (Y, b') with b' at position p < 3 of its ring is the image under R^-(3-p) of column
// (Y, b'' = R^(3-p) b'), which lies in the last face.  Its rows (X, a'), X <= Y, map to rows (X, R^(3-p) a') of that column, and
// these are on or above the diagonal as well: for X = Y and a' in the ring of b', a' <= b' means position(a') <= p, hence
// position(a') + 3 - p <= 3 without wrapping round.  R keeps the index inside a face, so a face-sized run of rows is a
// face-sized run of rows of the source column: the fill is memcpy of runs of nside^2 doubles.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/cmg.h"

namespace
{

inline int64_t packedOffset(int64_t col) { return col * (col + 1) / 2; }

// fills columns [colBegin, colEnd) (in-face column indices q) of (strip, face) from the matching column of the ring's last face
void fillColumns(double* packed, int64_t facePix, int strip, int face, int64_t qBegin, int64_t qEnd)
{
    const int64_t n = 12 * facePix;
    const int k = 3 - (face & 3);                                   // rotation that takes `face` to the last face of its ring
    const int lastFace = face | 3;
    for(int64_t q = qBegin; q < qEnd; ++q)
    {
        double* dst = packed + packedOffset(strip * n + face * facePix + q);
        const double* src = packed + packedOffset(strip * n + lastFace * facePix + q);
        for(int x = 0; x <= strip; ++x)
        {
            // faces of the row pixel: all twelve for an earlier strip, up to and including `face` for the column's own strip
            const int nFaces = x < strip ? 12 : face + 1;
            for(int fa = 0; fa < nFaces; ++fa)
            {
                const int fs = (fa & ~3) | ((fa + k) & 3);
                const int64_t len = (x == strip && fa == face) ? q + 1 : facePix;
                std::memcpy(dst + x * n + fa * facePix, src + x * n + fs * facePix, sizeof(double) * len);
            }
        }
    }
}

} // namespace

extern "C" cmg_status cmg_host_expand_rotations(double* packed, int64_t nside, int strip_begin, int strip_end, int face_begin, int face_end,
                                                int threads)
{
    if(!packed || nside < 1 || (nside & (nside - 1)) || strip_begin < 0 || strip_end > 3 || strip_begin > strip_end || face_begin < 0 ||
       face_end > 12 || face_begin > face_end)
        return CMG_EINVAL;
    const int64_t facePix = nside * nside;
    // work items: (strip, face not the last of its ring, chunk of columns); later strips have taller columns, start with them
    struct Item { int strip, face; int64_t q0, q1; };
    std::vector<Item> items;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(facePix, 64));
    for(int s = strip_end - 1; s >= strip_begin; --s)
        for(int f = face_end - 1; f >= face_begin; --f)
            if((f & 3) != 3)
                for(int64_t q = 0; q < facePix; q += chunk)
                    items.push_back({s, f, q, std::min(facePix, q + chunk)});
    std::atomic<size_t> next(0);
    auto work = [&]()
    {
        for(size_t i = next.fetch_add(1); i < items.size(); i = next.fetch_add(1))
            fillColumns(packed, facePix, items[i].strip, items[i].face, items[i].q0, items[i].q1);
    };
    const int nThreads = std::max(1, std::min(threads, 256));
    std::vector<std::thread> pool;
    for(int t = 1; t < nThreads; ++t)
        pool.emplace_back(work);
    work();
    for(auto& th : pool)
        th.join();
    return CMG_OK;
}
