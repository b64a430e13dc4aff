// Host half of the full-sky whole call when the matrix is wanted in HOST memory (cmg_cl_to_cmatrix_pol, cmg_cl_to_cmatrix with
// cmg_set_host_expand): the pi/2 rotation symmetry of the HEALPix grid (orbit.cuh) makes three quarters of the packed matrix
// copies of the rest, so only the columns of the LAST face of every ring of four (base faces 3, 7, 11) cross PCIe -- 27 % of the
// bytes, contiguous pieces of the packed triangle -- and the host fills in the other columns with block copies.
//
// Why those columns suffice: column (Y, b') with b' at position p < 3 of its ring is the image under R^-(3-p) of column
// (Y, b'' = R^(3-p) b'), which lies in the last face.  Its rows (X, a'), X <= Y, map to rows (X, R^(3-p) a') of that column, and
// these are on or above the diagonal as well: for X = Y and a' in the ring of b', a' <= b' means position(a') <= p, hence
// position(a') + 3 - p <= 3 without wrapping round.  R keeps the index inside a face, so a face-sized run of rows is a
// face-sized run of rows of the source column: the fill is a copy of runs of nside^2 doubles.
//
// The fill is bound by host memory traffic, so it is organised around the SOURCE run: a run of a last-face column is read once
// (it stays in the core's L2 for the second and third pass) and written to its three images with non-temporal stores (no
// read-for-ownership of the destination lines): 23.5 + 63.5 GB of DRAM traffic for the Nside = 64 [T;Q;U] matrix instead of
// 3 x 63.5 GB with memcpy per destination run.  A pool of worker threads takes column chunks as they are published, so the fill
// of a chunk overlaps the PCIe copies of the chunks behind it (ExpandPipeline, used by copyBackLastFacesAndExpand in cmg_api.cu).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>

#include "../../include/cmg.h"
#include "host_expand.hpp"

namespace
{

inline int64_t packedOffset(int64_t col) { return col * (col + 1) / 2; }

void plainCopy(double* dst, const double* src, size_t n) { std::memcpy(dst, src, sizeof(double) * n); }

// streaming copy: destination brought to a 64-byte boundary, then whole cache lines as non-temporal stores
__attribute__((target("avx2"))) void streamCopy(double* dst, const double* src, size_t n)
{
    while(n && (reinterpret_cast<uintptr_t>(dst) & 63))
    {
        *dst++ = *src++;
        --n;
    }
    size_t i = 0;
    for(; i + 16 <= n; i += 16)
    {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        const __m256d c = _mm256_loadu_pd(src + i + 8), d = _mm256_loadu_pd(src + i + 12);
        _mm256_stream_pd(dst + i, a);
        _mm256_stream_pd(dst + i + 4, b);
        _mm256_stream_pd(dst + i + 8, c);
        _mm256_stream_pd(dst + i + 12, d);
    }
    for(; i + 8 <= n; i += 8)
    {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        _mm256_stream_pd(dst + i, a);
        _mm256_stream_pd(dst + i + 4, b);
    }
    for(; i < n; ++i)
        dst[i] = src[i];
}

typedef void (*CopyFn)(double*, const double*, size_t);

CopyFn pickCopy()
{
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? streamCopy : plainCopy;
}

// images of the columns [qBegin, qEnd) of (strip, last face of `ring`): every run of rows is read once and written to the
// matching run of the same column index in the three other faces of the ring
void expandColumns(double* packed, int64_t facePix, int strip, int ring, int64_t qBegin, int64_t qEnd, CopyFn copy, int directMask)
{
    const int64_t n = 12 * facePix;
    const int lastFace = 4 * ring + 3;
    for(int64_t q = qBegin; q < qEnd; ++q)
    {
        const double* src = packed + packedOffset(strip * n + lastFace * facePix + q);
        double* dst[4];
        for(int k = 1; k <= 3; ++k)
            dst[k] = packed + packedOffset(strip * n + (lastFace - k) * facePix + q);
        const int skip = (directMask >> (3 * strip)) & 7;             // bit k - 1: image k of this strip arrives by itself
        if(skip == 7)
            continue;
        for(int x = 0; x <= strip; ++x)
        {
            // faces of the row pixel of the source column: all twelve for an earlier strip, up to the last face for its own
            const int nFaces = x < strip ? 12 : lastFace + 1;
            for(int fs = 0; fs < nFaces; ++fs)
            {
                const double* run = src + x * n + fs * facePix;
                for(int k = 1; k <= 3; ++k)
                {
                    if((skip >> (k - 1)) & 1)
                        continue;                                          // a direct copy brings this image
                    const int face = lastFace - k;                         // destination column face: rotation by -k
                    const int fa = (fs & ~3) | ((fs - k) & 3);             // destination row face
                    if(x == strip && fa > face)
                        continue;                                          // below the diagonal of the destination column
                    const int64_t len = (x == strip && fa == face) ? q + 1 : facePix;
                    copy(dst[k] + x * n + fa * facePix, run, static_cast<size_t>(len));
                }
            }
        }
    }
    _mm_sfence();
}

} // namespace

namespace cmg
{

struct ExpandPipeline
{
    struct Item { int strip, ring; int64_t q0, q1; };
    double* packed;
    int64_t facePix;
    CopyFn copy;
    int directMask;
    std::mutex m;
    std::condition_variable cv;
    std::deque<Item> items;
    bool closed;
    std::vector<std::thread> pool;

    void work()
    {
        for(;;)
        {
            Item it;
            {
                std::unique_lock<std::mutex> lock(m);
                cv.wait(lock, [&] { return closed || !items.empty(); });
                if(items.empty())
                    return;
                it = items.front();
                items.pop_front();
            }
            expandColumns(packed, facePix, it.strip, it.ring, it.q0, it.q1, copy, directMask);
        }
    }
};

ExpandPipeline* expandBegin(double* packed, int64_t nside, int threads, int directMask)
{
    ExpandPipeline* p = new ExpandPipeline;
    p->packed = packed;
    p->directMask = directMask;
    p->facePix = nside * nside;
    p->copy = pickCopy();
    p->closed = false;
    const int nThreads = std::max(1, std::min(threads, 256));
    for(int t = 0; t < nThreads; ++t)
        p->pool.emplace_back([p] { p->work(); });
    return p;
}

void expandPublish(ExpandPipeline* p, int strip, int ring, int64_t q0, int64_t q1)
{
    // items of a few columns each: a [T;Q;U] column of the third strip is 1.2 MB at Nside = 64, its three images 3.5 MB
    const int64_t step = std::max<int64_t>(1, std::min<int64_t>(16, (q1 - q0 + 63) / 64));
    {
        std::lock_guard<std::mutex> lock(p->m);
        for(int64_t q = q0; q < q1; q += step)
            p->items.push_back({strip, ring, q, std::min(q1, q + step)});
    }
    p->cv.notify_all();
}

void expandFinish(ExpandPipeline* p)
{
    {
        std::lock_guard<std::mutex> lock(p->m);
        p->closed = true;
    }
    p->cv.notify_all();
    for(auto& th : p->pool)
        th.join();
    delete p;
}

} // namespace cmg

extern "C" cmg_status cmg_host_expand_rotations(double* packed, int64_t nside, int strip_begin, int strip_end, int face_begin, int face_end,
                                                int threads)
{
    if(!packed || nside < 1 || (nside & (nside - 1)) || strip_begin < 0 || strip_end > 3 || strip_begin > strip_end || face_begin < 0 ||
       face_end > 12 || face_begin > face_end || (face_begin & 3) || (face_end & 3))
        return CMG_EINVAL;                       // whole rings of faces: a source run is written to all three images at once
    cmg::ExpandPipeline* p = cmg::expandBegin(packed, nside, threads);
    for(int s = strip_end - 1; s >= strip_begin; --s)
        for(int ring = face_end / 4 - 1; ring >= face_begin / 4; --ring)
            cmg::expandPublish(p, s, ring, 0, nside * nside);
    cmg::expandFinish(p);
    return CMG_OK;
}
