// CUDA kernels of the C-matrix generator, hand-written for sm_100a (B200).
//
// The work is an FP64 three-term recurrence per pixel pair, not a contraction, so it runs on the
// CUDA-core FP64 pipe (64 DFMA lanes per SM); tensor cores, TMEM and TMA tiles have nothing to
// offer here.  What matters on this part:
//   * the inner loop must be DFMA only: Clenshaw evaluation in the "2z" normalisation of series.hpp
//     costs 2 FMA per l for P_l and d^l_20 and 3 for d^l_2+-2, with no multiply and no divide;
//   * the per-l coefficients are warp-uniform: they are staged once per CTA in shared memory and
//     fetched with broadcast LDS.128 (the FP64 pipe issues one warp instruction every other cycle,
//     so the loads ride in the free issue slots);
//   * lanes run along the row index i, which is the contiguous direction of the packed upper
//     triangle, so every warp store is one 256-byte run; entries whose contiguous direction is j
//     (the transposed partners of a pair in the polarized layout) go through a shared-memory tile;
//   * several independent columns per thread give the ILP that hides the DFMA latency at the low
//     occupancy FP64 register pressure allows.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cmg.h"

namespace cmg
{

// resident per-pixel geometry, structure of arrays (each npix doubles)
struct Geometry
{
    const double* nx;
    const double* ny;
    const double* nz;
    const double* tx;   // e_theta
    const double* ty;
    const double* tz;
    const double* px;   // e_phi (z component is 0)
    const double* py;
    long long npix;
};

// recurrence tables resident on the device (series.hpp), each CMG_LMAX_LIMIT + 2 doubles
struct DeviceTables
{
    const double* N0;    // Legendre (m = m' = 0)
    const double* g0;
    const double* N20;   // d^l_20
    const double* g20;
    const double* N22;   // d^l_22 and d^l_2-2 share N and g; c flips sign
    const double* g22;
    const double* c22;
};

struct PartTable
{
    int n;
    int own;
    long long begin[CMG_MAX_PARTS + 1];
    double* ptr[CMG_MAX_PARTS][3];
    int kind[CMG_MAX_PARTS];
    long long ld[CMG_MAX_PARTS];
    long long row0[CMG_MAX_PARTS];
};

__host__ __device__ inline long long packedOffset(long long col) { return col * (col + 1) / 2; }

// ------------------------------------------------------------------------------------------------
// TT: S_ij = sum_l a_l P_l(n_i . n_j), columns [colBegin, colEnd) of the packed triangle.
// CTA = 128 rows x TT_COLS columns; a thread owns one row and walks the columns TT_R at a time.
// blockIdx.z selects a batch element (independent weight vectors / output matrices).
// ------------------------------------------------------------------------------------------------
constexpr int TT_ROWS = 128;
constexpr int TT_COLS = 16;
constexpr int TT_R = 4;

__global__ void __launch_bounds__(TT_ROWS)
legendreSeriesKernel(Geometry geo, const double* __restrict__ a, long long aStride,
                     const double* __restrict__ N0, const double* __restrict__ g0, int lmax,
                     long long colBegin, long long colEnd, double* __restrict__ out, long long outStride)
{
    extern __shared__ double2 ttTab[];      // [lmax + 1]  {a_k N_k, g_{k+1}}

    const long long rowBlock = static_cast<long long>(blockIdx.x) * TT_ROWS;
    const long long c0 = colBegin + static_cast<long long>(blockIdx.y) * TT_COLS;
    const long long c1 = min(c0 + static_cast<long long>(TT_COLS), colEnd);
    if(rowBlock > c1 - 1)
        return;                              // tile entirely below the diagonal (i > j)

    a += static_cast<long long>(blockIdx.z) * aStride;
    out += static_cast<long long>(blockIdx.z) * outStride;

    for(int k = threadIdx.x; k <= lmax; k += TT_ROWS)
        ttTab[k] = make_double2(a[k] * N0[k], g0[k + 1]);
    __syncthreads();

    if(rowBlock + (threadIdx.x & ~31) > c1 - 1)
        return;                              // this warp's 32 rows are all below the diagonal

    const long long i = rowBlock + threadIdx.x;
    const long long iLoad = min(i, geo.npix - 1);
    const double xi = geo.nx[iLoad], yi = geo.ny[iLoad], zi = geo.nz[iLoad];
    const long long base = packedOffset(colBegin);

    for(long long c = c0; c < c1; c += TT_R)
    {
        double x2[TT_R], b1[TT_R], b2[TT_R];
#pragma unroll
        for(int r = 0; r < TT_R; ++r)
        {
            const long long j = min(c + r, c1 - 1);
            // same association as the reference's ThreeVector product (include/three_vector.hpp:37)
            double dot = __dadd_rn(__dadd_rn(__dmul_rn(xi, __ldg(geo.nx + j)), __dmul_rn(yi, __ldg(geo.ny + j))),
                                   __dmul_rn(zi, __ldg(geo.nz + j)));
            dot = fmin(1.0, fmax(-1.0, dot));   // clamp of reference c_matrix_generator.cpp:206-215
            x2[r] = dot + dot;
            b1[r] = 0.0;
            b2[r] = 0.0;
        }
#pragma unroll 4
        for(int k = lmax; k >= 0; --k)
        {
            const double2 t = ttTab[k];
#pragma unroll
            for(int r = 0; r < TT_R; ++r)
            {
                const double b = fma(x2[r], b1[r], fma(-t.y, b2[r], t.x));
                b2[r] = b1[r];
                b1[r] = b;
            }
        }
#pragma unroll
        for(int r = 0; r < TT_R; ++r)
        {
            const long long j = c + r;
            if(j < c1 && i <= j)
                __stcs(out + (packedOffset(j) - base + i), b1[r]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// T,Q,U: all nine entries of every pixel pair (i <= j) with j in the pixel columns of part `own`.
// CTA = 64 rows x 32 columns of pixel pairs, 256 threads; a thread owns one row and two columns per
// pass (4 passes).  Four Clenshaw sums per pair: P_l (2 FMA), d_20 (2), d_22 (3), d_2-2 (3).
// The frame rotation is division-free:  with m = e_theta + i e_phi,
//     (1+z) e^{i(psi_i - psi_j)} = -(m_i . conj m_j),   (1-z) e^{i(psi_i + psi_j)} = m_i . m_j,
//     sin(beta) e^{i psi_i} = m_i . n_j,
// and d^2_22 = ((1+z)/2)^2, d^2_2-2 = ((1-z)/2)^2, d^2_20 = sqrt(6)/4 (1-z^2) are exactly the
// prefactors Clenshaw leaves outside the sum, so every block entry is (sum) x (polynomial in dot
// products of the two frames) -- no special case at z = +-1.
// ------------------------------------------------------------------------------------------------
constexpr int PQ_TI = 64;
constexpr int PQ_TJ = 32;
constexpr int PQ_THREADS = 256;
constexpr int PQ_R = 2;
constexpr int PQ_STAGE_LD = PQ_TJ + 1;

__device__ inline double* partEntry(const PartTable& P, int k, int strip, long long npix, long long pixCol, long long row)
{
    // packed strip of part k: first element is entry (0, strip*npix + begin[k])
    const long long col = strip * npix + pixCol;
    const long long first = strip * npix + P.begin[k];
    return P.ptr[k][strip] + (packedOffset(col) - packedOffset(first) + row);
}

__device__ inline int ownerOf(const PartTable& P, long long pixCol)
{
    int k = 0;
    while(k + 1 < P.n && pixCol >= P.begin[k + 1])
        ++k;
    return k;
}

__global__ void __launch_bounds__(PQ_THREADS, 2)
tquKernel(Geometry geo, const double* __restrict__ a, long long aStride, DeviceTables tab, int lmax,
          const __grid_constant__ PartTable P, long long outStride)
{
    extern __shared__ double4 pqSmem[];
    double4* tab4 = pqSmem;                                          // [2 (lmax+1)]
    double* sI = reinterpret_cast<double*>(tab4 + 2 * (lmax + 1));   // [8][PQ_TI]
    double* sJ = sI + 8 * PQ_TI;                                     // [8][PQ_TJ]
    double* stage = sJ + 8 * PQ_TJ;                                  // [3][PQ_TI][PQ_STAGE_LD]

    const long long npix = geo.npix;
    const long long colBegin = P.begin[P.own], colEnd = P.begin[P.own + 1];
    const long long rowBlock = static_cast<long long>(blockIdx.x) * PQ_TI;
    const long long c0 = colBegin + static_cast<long long>(blockIdx.y) * PQ_TJ;
    const long long c1 = min(c0 + static_cast<long long>(PQ_TJ), colEnd);
    if(rowBlock > c1 - 1)
        return;

    a += static_cast<long long>(blockIdx.z) * aStride;
    const long long batchOff = static_cast<long long>(blockIdx.z) * outStride;

    const int tid = threadIdx.x;
    for(int k = tid; k <= lmax; k += PQ_THREADS)
    {
        const double* att = a;
        const double* ate = a + (lmax + 1);
        const double* aee = a + 2 * (lmax + 1);
        const double* abb = a + 3 * (lmax + 1);
        double4 t0, t1;
        t0.x = att[k] * tab.N0[k];
        t0.y = tab.g0[k + 1];
        if(k >= 2)
        {
            t0.z = ate[k] * tab.N20[k] * 0.61237243569579452455;     // sqrt(6)/4 = d^2_20 / (1 - z^2)
            t0.w = tab.g20[k + 1];
            t1.x = (aee[k] + abb[k]) * tab.N22[k] * 0.125;            // 1/4 from d^2_2+-2, 1/2 from Re(A+-B)/2
            t1.y = tab.g22[k + 1];
            t1.z = (aee[k] - abb[k]) * tab.N22[k] * 0.125;
            t1.w = tab.c22[k];
        }
        else
        {
            t0.z = 0.0; t0.w = 0.0;
            t1 = make_double4(0.0, 0.0, 0.0, 0.0);
        }
        tab4[2 * k] = t0;
        tab4[2 * k + 1] = t1;
    }
    for(int idx = tid; idx < PQ_TI + PQ_TJ; idx += PQ_THREADS)
    {
        const bool isRow = idx < PQ_TI;
        const int loc = isRow ? idx : idx - PQ_TI;
        const long long pix = min(isRow ? rowBlock + loc : c0 + loc, npix - 1);
        double* dst = isRow ? sI : sJ;
        const int ld = isRow ? PQ_TI : PQ_TJ;
        dst[0 * ld + loc] = geo.nx[pix];
        dst[1 * ld + loc] = geo.ny[pix];
        dst[2 * ld + loc] = geo.nz[pix];
        dst[3 * ld + loc] = geo.tx[pix];
        dst[4 * ld + loc] = geo.ty[pix];
        dst[5 * ld + loc] = geo.tz[pix];
        dst[6 * ld + loc] = geo.px[pix];
        dst[7 * ld + loc] = geo.py[pix];
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int il = lane + 32 * (warp & 1);
    const int colGroup = warp >> 1;                     // 8 columns each
    const long long i = rowBlock + il;
    const bool warpLive = rowBlock + 32 * (warp & 1) <= c1 - 1;   // some row of this warp can be <= some column

    const double nix = sI[0 * PQ_TI + il], niy = sI[1 * PQ_TI + il], niz = sI[2 * PQ_TI + il];

    double* const ownT = P.ptr[P.own][0] + batchOff;
    double* const ownQ = P.ptr[P.own][1] + batchOff;
    double* const ownU = P.ptr[P.own][2] + batchOff;
    const long long firstT = packedOffset(colBegin);
    const long long firstQ = packedOffset(npix + colBegin);
    const long long firstU = packedOffset(2 * npix + colBegin);

    if(warpLive)
    {
        for(int pass = 0; pass < 8 / PQ_R; ++pass)
        {
            const int jl0 = colGroup * 8 + pass * PQ_R;
            if(c0 + jl0 >= c1)
                break;
            double x2[PQ_R];
            double tt1[PQ_R], tt2[PQ_R], te1[PQ_R], te2[PQ_R], pp1[PQ_R], pp2[PQ_R], mm1[PQ_R], mm2[PQ_R];
#pragma unroll
            for(int r = 0; r < PQ_R; ++r)
            {
                const int jl = jl0 + r;
                double dot = __dadd_rn(__dadd_rn(__dmul_rn(nix, sJ[0 * PQ_TJ + jl]), __dmul_rn(niy, sJ[1 * PQ_TJ + jl])),
                                       __dmul_rn(niz, sJ[2 * PQ_TJ + jl]));
                dot = fmin(1.0, fmax(-1.0, dot));
                x2[r] = dot + dot;
                tt1[r] = tt2[r] = te1[r] = te2[r] = pp1[r] = pp2[r] = mm1[r] = mm2[r] = 0.0;
            }
#pragma unroll 2
            for(int k = lmax; k >= 2; --k)
            {
                const double4 t0 = tab4[2 * k];
                const double4 t1 = tab4[2 * k + 1];
#pragma unroll
                for(int r = 0; r < PQ_R; ++r)
                {
                    const double tt = fma(x2[r], tt1[r], fma(-t0.y, tt2[r], t0.x));
                    const double te = fma(x2[r], te1[r], fma(-t0.w, te2[r], t0.z));
                    const double pp = fma(x2[r], pp1[r], fma(-t1.w, pp1[r], fma(-t1.y, pp2[r], t1.x)));
                    const double mm = fma(x2[r], mm1[r], fma(t1.w, mm1[r], fma(-t1.y, mm2[r], t1.z)));
                    tt2[r] = tt1[r]; tt1[r] = tt;
                    te2[r] = te1[r]; te1[r] = te;
                    pp2[r] = pp1[r]; pp1[r] = pp;
                    mm2[r] = mm1[r]; mm1[r] = mm;
                }
            }
#pragma unroll
            for(int k = 1; k >= 0; --k)
            {
                const double4 t0 = tab4[2 * k];
#pragma unroll
                for(int r = 0; r < PQ_R; ++r)
                {
                    const double tt = fma(x2[r], tt1[r], fma(-t0.y, tt2[r], t0.x));
                    tt2[r] = tt1[r]; tt1[r] = tt;
                }
            }

            // frame of pixel i (lane-contiguous shared loads)
            const double tix = sI[3 * PQ_TI + il], tiy = sI[4 * PQ_TI + il], tiz = sI[5 * PQ_TI + il];
            const double pix_ = sI[6 * PQ_TI + il], piy = sI[7 * PQ_TI + il];
#pragma unroll
            for(int r = 0; r < PQ_R; ++r)
            {
                const int jl = jl0 + r;
                const long long j = c0 + jl;
                const double njx = sJ[0 * PQ_TJ + jl], njy = sJ[1 * PQ_TJ + jl], njz = sJ[2 * PQ_TJ + jl];
                const double tjx = sJ[3 * PQ_TJ + jl], tjy = sJ[4 * PQ_TJ + jl], tjz = sJ[5 * PQ_TJ + jl];
                const double pjx = sJ[6 * PQ_TJ + jl], pjy = sJ[7 * PQ_TJ + jl];

                const double ai = fma(njx, tix, fma(njy, tiy, njz * tiz));    // n_j . e_theta(i)
                const double bi = fma(njx, pix_, njy * piy);                  // n_j . e_phi(i)
                const double aj = fma(nix, tjx, fma(niy, tjy, niz * tjz));
                const double bj = fma(nix, pjx, niy * pjy);
                const double p = fma(tix, tjx, fma(tiy, tjy, tiz * tjz));     // e_theta(i) . e_theta(j)
                const double q = fma(pix_, pjx, piy * pjy);                   // e_phi(i) . e_phi(j)
                const double rr = fma(pix_, tjx, piy * tjy);                  // e_phi(i) . e_theta(j)
                const double tq = fma(tix, pjx, tiy * pjy);                   // e_theta(i) . e_phi(j)

                const double su = p + q, du = rr - tq, sv = p - q, dv = tq + rr;
                const double aRe = pp1[r] * fma(su, su, -du * du), aIm = pp1[r] * (2.0 * su * du);
                const double bRe = mm1[r] * fma(sv, sv, -dv * dv), bIm = mm1[r] * (2.0 * sv * dv);
                const double xt = -te1[r];

                const bool valid = (i <= j) && (j < c1) && (i < npix);
                if(valid)
                {
                    double* colT = ownT + (packedOffset(j) - firstT);
                    double* colQ = ownQ + (packedOffset(npix + j) - firstQ);
                    double* colU = ownU + (packedOffset(2 * npix + j) - firstU);
                    __stcs(colT + i, tt1[r]);
                    __stcs(colQ + i, xt * fma(aj, aj, -bj * bj));       // T_i Q_j
                    __stcs(colQ + npix + i, aRe + bRe);                 // Q_i Q_j
                    __stcs(colU + i, xt * (2.0 * aj * bj));             // T_i U_j
                    __stcs(colU + npix + i, bIm - aIm);                 // Q_i U_j
                    __stcs(colU + 2 * npix + i, aRe - bRe);             // U_i U_j
                }
                stage[(0 * PQ_TI + il) * PQ_STAGE_LD + jl] = xt * fma(ai, ai, -bi * bi);   // Q_i T_j
                stage[(1 * PQ_TI + il) * PQ_STAGE_LD + jl] = xt * (2.0 * ai * bi);         // U_i T_j
                stage[(2 * PQ_TI + il) * PQ_STAGE_LD + jl] = aIm + bIm;                    // U_i Q_j
            }
        }
    }
    __syncthreads();

    // transposed partners: for fixed i the 32 columns j of this tile are contiguous rows of column
    // (N+i) / (2N+i); one warp store per (entry kind, i)
    const long long j = c0 + lane;
    for(int row = warp; row < 3 * PQ_TI; row += PQ_THREADS / 32)
    {
        const int t = row / PQ_TI;
        const int ilr = row - t * PQ_TI;
        const long long ir = rowBlock + ilr;
        if(ir >= npix || ir >= c1 - 1 + 1)
            continue;
        if(!(j < c1 && ir < j))
            continue;
        const double v = stage[(t * PQ_TI + ilr) * PQ_STAGE_LD + lane];
        const int k = ownerOf(P, ir);
        double* dst;
        if(P.kind[k] == 0)
        {
            if(t == 0) dst = partEntry(P, k, 1, npix, ir, j);               // Q_i T_j -> col N+i, row j
            else if(t == 1) dst = partEntry(P, k, 2, npix, ir, j);          // U_i T_j -> col 2N+i, row j
            else dst = partEntry(P, k, 2, npix, ir, npix + j);              // U_i Q_j -> col 2N+i, row N+j
        }
        else
        {
            dst = P.ptr[k][t] + ((ir - P.begin[k]) * P.ld[k] + (j - P.row0[k]));
        }
        __stcs(dst + batchOff, v);
    }
}

// ------------------------------------------------------------------------------------------------
// CMatrix::maskMatrix gather (reference source/c_matrix.cpp:182-201): out(a,b) = in(good[a], good[b])
// ------------------------------------------------------------------------------------------------
__global__ void maskGatherKernel(const double* __restrict__ in, const int* __restrict__ good, long long nGood,
                                 double* __restrict__ out)
{
    const long long b = blockIdx.y;
    const long long aIdx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if(aIdx > b || b >= nGood)
        return;
    long long gi = good[aIdx], gj = good[b];
    if(gi > gj)
    {
        const long long s = gi; gi = gj; gj = s;
    }
    out[packedOffset(b) + aIdx] = in[packedOffset(gj) + gi];
}

// ------------------------------------------------------------------------------------------------
// FP64 peak: independent DFMA chains, no memory traffic.  8 chains x 4096 iterations per thread.
// ------------------------------------------------------------------------------------------------
constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(256) fp64PeakKernel(double* sink, double x, double y)
{
    double acc[PEAK_CHAINS];
#pragma unroll
    for(int c = 0; c < PEAK_CHAINS; ++c)
        acc[c] = x + c + threadIdx.x;
#pragma unroll 1
    for(int it = 0; it < PEAK_ITERS / 8; ++it)
    {
#pragma unroll
        for(int u = 0; u < 8; ++u)
#pragma unroll
            for(int c = 0; c < PEAK_CHAINS; ++c)
                acc[c] = fma(acc[c], y, x);
    }
    double s = 0;
#pragma unroll
    for(int c = 0; c < PEAK_CHAINS; ++c)
        s += acc[c];
    if(s == 123.456)
        sink[0] = s;
}

} // namespace cmg
