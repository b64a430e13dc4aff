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
constexpr int TT_R = 4;                                // columns per thread and pass (default variant)
constexpr int TT_STATIC_STEPS = 1024;                 // slot i <-> k = TT_STATIC_STEPS - 1 - i
constexpr int TT_STATIC_CHUNK = 8;

// coefficients in the kernel parameter block (see the note at TquStaticTable): s[i] = { a_k N_k, -g_{k+1} }
struct TtStaticTable
{
    double2 s[TT_STATIC_STEPS];
};

template <int R>
__device__ __forceinline__ void ttStep(double (&x2)[R], double (&b1)[R], double (&b2)[R], const double2 t)
{
#pragma unroll
    for(int r = 0; r < R; ++r)
    {
        // (a + x2 b1) - g b2: one warp-uniform operand per DFMA (see tquStep)
        const double b = fma(t.y, b2[r], fma(x2[r], b1[r], t.x));
        b2[r] = b1[r];
        b1[r] = b;
    }
}

template <int R, int N = TT_STATIC_CHUNK>
__device__ __forceinline__ void ttStaticChunkAt(double (&x2)[R], double (&b1)[R], double (&b2)[R], const TtStaticTable& T, int first)
{
#pragma unroll
    for(int u = 0; u < N; ++u)
        ttStep<R>(x2, b1, b2, T.s[first + u]);
}

// A rolled loop over 8-step chunks with a warp-uniform chunk counter: ptxas addresses the table as c[0x0][UR + imm]
// (LDCU with a uniform-register base), rotates three uniform registers so that every coefficient is loaded two groups
// of DFMAs ahead of its use, and interleaves the R columns (all inner products, then all outer ones).  The earlier
// fully unrolled body entered through a 128-way switch got none of this: one uniform register for every load, issued
// right in front of its consumer, dependent DFMA pairs back to back, and 49 KB of straight-line code.
template <int R>
__device__ __forceinline__ void ttClenshawStatic(double (&x2)[R], double (&b1)[R], double (&b2)[R], const TtStaticTable& T, int entrySlot)
{
    // entrySlot = first slot with a non-zero weight (k = lmax): one unrolled head of 1..7 steps up to the next chunk boundary
    // (warp-uniform switch), then whole chunks
    int i = entrySlot;
    const int head = (TT_STATIC_CHUNK - (i & (TT_STATIC_CHUNK - 1))) & (TT_STATIC_CHUNK - 1);
    switch(head)
    {
        case 1: ttStaticChunkAt<R, 1>(x2, b1, b2, T, i); break;
        case 2: ttStaticChunkAt<R, 2>(x2, b1, b2, T, i); break;
        case 3: ttStaticChunkAt<R, 3>(x2, b1, b2, T, i); break;
        case 4: ttStaticChunkAt<R, 4>(x2, b1, b2, T, i); break;
        case 5: ttStaticChunkAt<R, 5>(x2, b1, b2, T, i); break;
        case 6: ttStaticChunkAt<R, 6>(x2, b1, b2, T, i); break;
        case 7: ttStaticChunkAt<R, 7>(x2, b1, b2, T, i); break;
        default: break;
    }
    i += head;
#pragma unroll 1
    for(; i < TT_STATIC_STEPS; i += TT_STATIC_CHUNK)
        ttStaticChunkAt<R>(x2, b1, b2, T, i);
}

template <bool STATIC, int R, int MINB>
__global__ void __launch_bounds__(TT_ROWS, MINB)
legendreSeriesKernel(const __grid_constant__ TtStaticTable T, Geometry geo, const double* __restrict__ a, long long aStride,
                     const double* __restrict__ N0, const double* __restrict__ g0, int lmax, int entrySlot,
                     long long colBegin, long long colEnd, double* __restrict__ out, long long outStride)
{
    extern __shared__ double2 ttTab[];      // [lmax + 1]  {a_k N_k, -g_{k+1}}  (dynamic variant only)

    const long long rowBlock = static_cast<long long>(blockIdx.x) * TT_ROWS;
    const long long c0 = colBegin + static_cast<long long>(blockIdx.y) * TT_COLS;
    const long long c1 = min(c0 + static_cast<long long>(TT_COLS), colEnd);
    if(rowBlock > c1 - 1)
        return;                              // tile entirely below the diagonal (i > j)

    out += static_cast<long long>(blockIdx.z) * outStride;

    if(!STATIC)
    {
        a += static_cast<long long>(blockIdx.z) * aStride;
        for(int k = threadIdx.x; k <= lmax; k += TT_ROWS)
            ttTab[k] = make_double2(a[k] * N0[k], -g0[k + 1]);
        __syncthreads();
    }

    if(rowBlock + (threadIdx.x & ~31) > c1 - 1)
        return;                              // this warp's 32 rows are all below the diagonal

    const long long i = rowBlock + threadIdx.x;
    const long long iLoad = min(i, geo.npix - 1);
    const double xi = geo.nx[iLoad], yi = geo.ny[iLoad], zi = geo.nz[iLoad];
    const long long base = packedOffset(colBegin);

    for(long long c = c0; c < c1; c += R)
    {
        double x2[R], b1[R], b2[R];
#pragma unroll
        for(int r = 0; r < R; ++r)
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
        if(STATIC)
            ttClenshawStatic<R>(x2, b1, b2, T, entrySlot);
        else
        {
#pragma unroll 4
            for(int k = lmax; k >= 0; --k)
                ttStep<R>(x2, b1, b2, ttTab[k]);
        }
        double* colPtr = out + (packedOffset(c) - base + i);
#pragma unroll
        for(int r = 0; r < R; ++r)
        {
            const long long j = c + r;
            if(j < c1 && i <= j)
                __stcs(colPtr, b1[r]);
            colPtr += j + 1;                     // next column starts j+1 entries further
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
constexpr int PQ_STAGE_LD = PQ_TJ + 1;

// Static-table variant: the per-l coefficients travel in the kernel parameter block (constant bank 0)
// at compile-time offsets, one fully unrolled Clenshaw step per table slot.  Only then does ptxas put
// the multiplier coefficients in uniform registers (LDCU) and issue DFMA R, R, UR, R -- two vector
// register reads instead of three.  Measured on B200 (tools/fp64_bank.cu): a DFMA reading three distinct
// vector registers runs at <= 73% of the FP64 pipe rate (register-file bandwidth), so this matters more
// than anything else in the loop.  The body is entered Duff-style at an 8-step boundary; slots above
// lmax hold zero weights, which leave the (zero) Clenshaw state untouched.
constexpr int PQ_STATIC_STEPS = 440;                 // 4-series steps k = PQ_STATIC_STEPS+1 .. 2
constexpr int PQ_STATIC_CHUNK = 8;
constexpr int PQ_STATIC_MAIN = 16;                  // steps per iteration of the rolled loop
constexpr int PQ_STATIC_LMAX = PQ_STATIC_STEPS + 1;  // largest lmax the static table holds

struct TquStaticTable
{
    // slot i <-> k = PQ_STATIC_STEPS + 1 - i:
    //   s[2i]   = { a_tt N0,  a_te N20 sqrt(6)/4,  (a_ee+a_bb) N22 / 8,  (a_ee-a_bb) N22 / 8 }  at k
    //   s[2i+1] = { -g0_{k+1}, -g20_{k+1}, -g22_{k+1}, c22_k }
    // tail (TT only): s[2 STEPS] = { a_tt N0 at k=1, -g0_2, a_tt N0 at k=0, -g0_1 }
    double4 s[2 * PQ_STATIC_STEPS + 1];
};

template <int R>
struct TquState
{
    double x2[R];
    double tt1[R], tt2[R], te1[R], te2[R], pp1[R], pp2[R], mm1[R], mm2[R];
};

// One Clenshaw step for R columns.  Association matters on this part: a DFMA whose three sources are
// three different vector registers (none held by the operand-reuse cache) costs ~0.9 cycle more than the
// pipe's 2-cycle rate (fitted over 20 loop variants, tools/exp/).  Written as (a + x2 b1) - g b2 every
// operation carries exactly one warp-uniform coefficient (a, c or g), which ptxas keeps in the reuse
// cache across the R columns; the textbook x2 b1 + (a - g b2) has an all-distinct outer operation.
template <int R>
__device__ __forceinline__ void tquStep(TquState<R>& s, const double4 A, const double4 G)
{
#pragma unroll
    for(int r = 0; r < R; ++r)
    {
        const double t = fma(s.x2[r], s.tt1[r], A.x);
        const double e = fma(s.x2[r], s.te1[r], A.y);
        const double p = fma(s.x2[r], s.pp1[r], A.z);
        const double m = fma(s.x2[r], s.mm1[r], A.w);
        const double pu = fma(-G.w, s.pp1[r], p);
        const double mu = fma(G.w, s.mm1[r], m);
        const double tt = fma(G.x, s.tt2[r], t);
        const double te = fma(G.y, s.te2[r], e);
        const double pp = fma(G.z, s.pp2[r], pu);
        const double mm = fma(G.z, s.mm2[r], mu);
        s.tt2[r] = s.tt1[r]; s.tt1[r] = tt;
        s.te2[r] = s.te1[r]; s.te1[r] = te;
        s.pp2[r] = s.pp1[r]; s.pp1[r] = pp;
        s.mm2[r] = s.mm1[r]; s.mm1[r] = mm;
    }
}

template <int R>
__device__ __forceinline__ void tquStepTT(TquState<R>& s, const double a, const double g)
{
#pragma unroll
    for(int r = 0; r < R; ++r)
    {
        const double tt = fma(g, s.tt2[r], fma(s.x2[r], s.tt1[r], a));
        s.tt2[r] = s.tt1[r];
        s.tt1[r] = tt;
    }
}

template <int R, int N = PQ_STATIC_CHUNK>
__device__ __forceinline__ void tquStaticChunkAt(TquState<R>& s, const TquStaticTable& T, int first)
{
#pragma unroll
    for(int u = 0; u < N; ++u)
        tquStep<R>(s, T.s[2 * (first + u)], T.s[2 * (first + u) + 1]);
}

// rolled loop over 8-step chunks with a warp-uniform counter (see ttClenshawStatic): LDCU c[0x0][UR + imm], coefficient
// loads software-pipelined by ptxas, a 6 KB loop body instead of 137 KB of straight-line code behind a 55-way switch
template <int R>
__device__ __forceinline__ void tquClenshawStatic(TquState<R>& s, const TquStaticTable& T, int entrySlot)
{
    // entrySlot = first slot with a non-zero weight (k = lmax): an unrolled head of 1..15 steps (warp-uniform switch over
    // 1..7, then one 8-step group) leaves a whole number of 16-step iterations; a rolled single-step head measured slower than
    // running the zero-weight slots, the unrolled one is faster than both.  16-step chunks: ptxas pipelines the LDCUs inside
    // the unrolled body but not across the back edge, so every iteration starts with a short bubble.
    int i = entrySlot;
    const int head = (PQ_STATIC_STEPS - i) & (PQ_STATIC_MAIN - 1);      // the rest is a whole number of 16-step iterations
    switch(head & 7)
    {
        case 1: tquStaticChunkAt<R, 1>(s, T, i); break;
        case 2: tquStaticChunkAt<R, 2>(s, T, i); break;
        case 3: tquStaticChunkAt<R, 3>(s, T, i); break;
        case 4: tquStaticChunkAt<R, 4>(s, T, i); break;
        case 5: tquStaticChunkAt<R, 5>(s, T, i); break;
        case 6: tquStaticChunkAt<R, 6>(s, T, i); break;
        case 7: tquStaticChunkAt<R, 7>(s, T, i); break;
        default: break;
    }
    i += head & 7;
    if(head & 8)
    {
        tquStaticChunkAt<R, 8>(s, T, i);
        i += 8;
    }
#pragma unroll 1
    for(; i < PQ_STATIC_STEPS; i += PQ_STATIC_MAIN)
        tquStaticChunkAt<R, PQ_STATIC_MAIN>(s, T, i);
    const double4 tail = T.s[2 * PQ_STATIC_STEPS];
    tquStepTT<R>(s, tail.x, tail.y);
    tquStepTT<R>(s, tail.z, tail.w);
}

template <int R>
__device__ __forceinline__ void tquClenshawShared(TquState<R>& s, const double4* tab4, int lmax)
{
    // tab4[2k] = A(k), tab4[2k+1] = G(k) in the same convention as the static table
#pragma unroll 2
    for(int k = lmax; k >= 2; --k)
        tquStep<R>(s, tab4[2 * k], tab4[2 * k + 1]);
    const double4 a1 = tab4[2], g1 = tab4[3], a0 = tab4[0], g0 = tab4[1];
    tquStepTT<R>(s, a1.x, g1.x);
    tquStepTT<R>(s, a0.x, g0.x);
}

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

struct TquDynamicArgs
{
    const double* a;        // [batch][4][lmax+1]  tt, te, ee, bb weights
    long long aStride;
    DeviceTables tab;
    int lmax;
};

// shared-memory footprint of tquKernel besides the (dynamic-variant) coefficient table
constexpr int PQ_SMEM_DOUBLES = 8 * PQ_TI + 8 * PQ_TJ + 3 * PQ_TI * PQ_STAGE_LD + 3 * PQ_TJ + 3 * PQ_TI;

template <int R, bool STATIC, int MINB>
__global__ void __launch_bounds__(PQ_THREADS, MINB)
tquKernel(const __grid_constant__ TquStaticTable T, Geometry geo, TquDynamicArgs dyn, int entrySlot,
          const __grid_constant__ PartTable P, long long outStride)      // T first: 128-byte aligned in the constant bank
{
    extern __shared__ double4 pqSmem[];
    const int tabSlots = STATIC ? 0 : 2 * (dyn.lmax + 1);
    double4* tab4 = pqSmem;                                          // [2 (lmax+1)] (dynamic variant only)
    double* sI = reinterpret_cast<double*>(tab4 + tabSlots);         // [8][PQ_TI]  frames of the tile's rows
    double* sJ = sI + 8 * PQ_TI;                                     // [8][PQ_TJ]  frames of the tile's columns
    double* stage = sJ + 8 * PQ_TJ;                                  // [3][PQ_TI][PQ_STAGE_LD] transposed partners
    double** sColPtr = reinterpret_cast<double**>(stage + 3 * PQ_TI * PQ_STAGE_LD);   // [3][PQ_TJ] row 0 of columns j, N+j, 2N+j
    double** sRowPtr = sColPtr + 3 * PQ_TJ;                          // [3][PQ_TI] destination of (kind t, row i) at column c0

    const long long npix = geo.npix;
    const long long colBegin = P.begin[P.own], colEnd = P.begin[P.own + 1];
    const long long rowBlock = static_cast<long long>(blockIdx.x) * PQ_TI;
    const long long c0 = colBegin + static_cast<long long>(blockIdx.y) * PQ_TJ;
    const long long c1 = min(c0 + static_cast<long long>(PQ_TJ), colEnd);
    if(rowBlock > c1 - 1)
        return;

    const long long batchOff = static_cast<long long>(blockIdx.z) * outStride;
    const int tid = threadIdx.x;

    if(!STATIC)
    {
        const int lmax = dyn.lmax;
        const double* att = dyn.a + static_cast<long long>(blockIdx.z) * dyn.aStride;
        const double* ate = att + (lmax + 1);
        const double* aee = att + 2 * (lmax + 1);
        const double* abb = att + 3 * (lmax + 1);
        for(int k = tid; k <= lmax; k += PQ_THREADS)
        {
            double4 A, G;
            A.x = att[k] * dyn.tab.N0[k];
            G.x = -dyn.tab.g0[k + 1];
            if(k >= 2)
            {
                A.y = ate[k] * dyn.tab.N20[k] * 0.61237243569579452455;     // sqrt(6)/4 = d^2_20 / (1 - z^2)
                A.z = (aee[k] + abb[k]) * dyn.tab.N22[k] * 0.125;            // 1/4 from d^2_2+-2, 1/2 from Re(A+-B)/2
                A.w = (aee[k] - abb[k]) * dyn.tab.N22[k] * 0.125;
                G.y = -dyn.tab.g20[k + 1];
                G.z = -dyn.tab.g22[k + 1];
                G.w = dyn.tab.c22[k];
            }
            else
            {
                A.y = A.z = A.w = 0.0;
                G.y = G.z = G.w = 0.0;
            }
            tab4[2 * k] = A;
            tab4[2 * k + 1] = G;
        }
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
    // destination pointers, once per tile instead of once per entry: the 64-bit packed-offset arithmetic and the
    // owner lookup are warp-uniform, and every non-FP64 instruction in the epilogue costs an FP64 issue cycle
    for(int idx = tid; idx < 3 * PQ_TJ + 3 * PQ_TI; idx += PQ_THREADS)
    {
        if(idx < 3 * PQ_TJ)
        {
            const int strip = idx / PQ_TJ;
            const long long jcol = min(c0 + (idx - strip * PQ_TJ), colEnd - 1);
            sColPtr[idx] = partEntry(P, P.own, strip, npix, jcol, 0) + batchOff;
        }
        else
        {
            const int q = idx - 3 * PQ_TJ;
            const int t = q / PQ_TI;
            const long long ir = min(rowBlock + (q - t * PQ_TI), npix - 1);
            const int k = ownerOf(P, ir);
            double* dst;
            if(P.kind[k] == 0)
            {
                if(t == 0) dst = partEntry(P, k, 1, npix, ir, c0);              // Q_i T_j -> col N+i, row j
                else if(t == 1) dst = partEntry(P, k, 2, npix, ir, c0);         // U_i T_j -> col 2N+i, row j
                else dst = partEntry(P, k, 2, npix, ir, npix + c0);             // U_i Q_j -> col 2N+i, row N+j
            }
            else
            {
                dst = P.ptr[k][t] + ((ir - P.begin[k]) * P.ld[k] + (c0 - P.row0[k]));
            }
            sRowPtr[q] = dst + batchOff;
        }
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int il = lane + 32 * (warp & 1);
    const int colGroup = warp >> 1;                     // 8 columns each
    const long long i = rowBlock + il;
    const bool warpLive = rowBlock + 32 * (warp & 1) <= c1 - 1;   // some row of this warp can be <= some column

    const double nix = sI[0 * PQ_TI + il], niy = sI[1 * PQ_TI + il], niz = sI[2 * PQ_TI + il];

    if(warpLive)
    {
        for(int pass = 0; pass < 8 / R; ++pass)
        {
            const int jl0 = colGroup * 8 + pass * R;
            if(c0 + jl0 >= c1)
                break;
            TquState<R> st;
#pragma unroll
            for(int r = 0; r < R; ++r)
            {
                const int jl = jl0 + r;
                double dot = __dadd_rn(__dadd_rn(__dmul_rn(nix, sJ[0 * PQ_TJ + jl]), __dmul_rn(niy, sJ[1 * PQ_TJ + jl])),
                                       __dmul_rn(niz, sJ[2 * PQ_TJ + jl]));
                dot = fmin(1.0, fmax(-1.0, dot));
                st.x2[r] = dot + dot;
                st.tt1[r] = st.tt2[r] = st.te1[r] = st.te2[r] = st.pp1[r] = st.pp2[r] = st.mm1[r] = st.mm2[r] = 0.0;
            }
            if(STATIC)
                tquClenshawStatic<R>(st, T, entrySlot);
            else
                tquClenshawShared<R>(st, tab4, dyn.lmax);

            // frame of pixel i (lane-contiguous shared loads)
            const double tix = sI[3 * PQ_TI + il], tiy = sI[4 * PQ_TI + il], tiz = sI[5 * PQ_TI + il];
            const double pix_ = sI[6 * PQ_TI + il], piy = sI[7 * PQ_TI + il];
#pragma unroll
            for(int r = 0; r < R; ++r)
            {
                const int jl = jl0 + r;
                const int j32 = static_cast<int>(c0 - rowBlock) + jl;          // j - rowBlock (fits int: tile-local)
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
                const double aRe = st.pp1[r] * fma(su, su, -du * du), aIm = st.pp1[r] * (2.0 * su * du);
                const double bRe = st.mm1[r] * fma(sv, sv, -dv * dv), bIm = st.mm1[r] * (2.0 * sv * dv);
                const double xt = -st.te1[r];

                // i <= j, j < c1 (i < npix follows: j < npix)
                if(il <= j32 && c0 + jl < c1)
                {
                    double* colT = sColPtr[0 * PQ_TJ + jl] + i;
                    double* colQ = sColPtr[1 * PQ_TJ + jl] + i;
                    double* colU = sColPtr[2 * PQ_TJ + jl] + i;
                    __stcs(colT, st.tt1[r]);
                    __stcs(colQ, xt * fma(aj, aj, -bj * bj));           // T_i Q_j
                    __stcs(colQ + npix, aRe + bRe);                     // Q_i Q_j
                    __stcs(colU, xt * (2.0 * aj * bj));                 // T_i U_j
                    __stcs(colU + npix, bIm - aIm);                     // Q_i U_j
                    __stcs(colU + 2 * npix, aRe - bRe);                 // U_i U_j
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
    const int jOff = static_cast<int>(c0 - rowBlock) + lane;       // j - rowBlock
    const bool colOk = c0 + lane < c1;
    for(int row = warp; row < 3 * PQ_TI; row += PQ_THREADS / 32)
    {
        const int ilr = row % PQ_TI;
        if(colOk && ilr < jOff)                                    // strictly above the diagonal: i < j
            __stcs(sRowPtr[row] + lane, stage[row * PQ_STAGE_LD + lane]);
    }
}

#undef CMG_CASE
#undef CMG_CASE5

// ------------------------------------------------------------------------------------------------
// Batched T,Q,U: many weight sets (one MCMC step's proposals) for the same pixel set.  Here the sum over l IS a
// contraction over a shared basis: the four function families are run FORWARD once per pixel pair and batch chunk
// (10 FP64 operations per l) and every batch element only adds 4 accumulate-FMAs per l, instead of a 10-FMA
// Clenshaw step per element.  CTA = 32 rows x 32 columns of pixel pairs; a thread owns one pair per pass (lane =
// row, warp = column, 4 passes of 8 columns) and accumulates PB_BT batch elements at a time (4 x PB_BT
// accumulators in registers); the basis value is the operand the reuse cache keeps across the PB_BT accumulate
// FMAs.  Weights arrive pre-folded (foldBatchedWeightsKernel) as [chunk][l][family][PB_BT], so staging a chunk is a
// straight 16-byte copy.  The frame rotation factors depend on the pair only and are computed once per pass.
// Output: matrix b at out + b * outStride, single-owner packed layout.
// ------------------------------------------------------------------------------------------------
constexpr int PB_BT = 16;                      // batch elements per accumulation chunk
constexpr int PB_T = 32;                       // tile edge (rows and columns)
constexpr int PB_CW = 8;                       // columns per pass (= warps per CTA)
constexpr int PB_STAGE_LD = PB_CW + 1;

__host__ __device__ inline size_t tquBatchedSmemBytes(int lmax)
{
    return sizeof(double4) * (lmax + 2)                                    // recurrence coefficients per l
           + sizeof(double) * (static_cast<size_t>(lmax + 1) * 4 * PB_BT    // weights of the chunk [l][family][b]
                               + 16 * PB_T                                   // frames of rows and columns
                               + PB_BT * 3 * PB_T * PB_STAGE_LD              // transposed partners of the chunk
                               + 6 * PB_T);                                  // destination pointers
}

// w[b][4][lmax+1] (tt, te, ee, bb weights) -> folded[chunk][l][family][PB_BT], zero padded to whole chunks
__global__ void foldBatchedWeightsKernel(const double* __restrict__ w, DeviceTables tab, int lmax, int nBatch, double* __restrict__ folded)
{
    const int n1 = lmax + 1;
    const long long total = static_cast<long long>((nBatch + PB_BT - 1) / PB_BT) * n1 * 4 * PB_BT;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const int bb = static_cast<int>(idx % PB_BT);
        const int fam = static_cast<int>((idx / PB_BT) & 3);
        const int l = static_cast<int>((idx / (4 * PB_BT)) % n1);
        const int chunk = static_cast<int>(idx / (static_cast<long long>(4 * PB_BT) * n1));
        const int b = chunk * PB_BT + bb;
        double v = 0.0;
        if(b < nBatch)
        {
            const double* wb = w + static_cast<long long>(b) * 4 * n1;
            if(fam == 0) v = wb[l] * tab.N0[l];
            else if(l >= 2)
            {
                if(fam == 1) v = wb[n1 + l] * tab.N20[l] * 0.61237243569579452455;
                else if(fam == 2) v = (wb[2 * n1 + l] + wb[3 * n1 + l]) * tab.N22[l] * 0.125;
                else v = (wb[2 * n1 + l] - wb[3 * n1 + l]) * tab.N22[l] * 0.125;
            }
        }
        folded[idx] = v;
    }
}

__global__ void __launch_bounds__(PB_T * PB_CW, 1)
tquBatchedKernel(Geometry geo, const double* __restrict__ folded, DeviceTables tab, int lmax, int nBatch,
                 const __grid_constant__ PartTable P, long long outStride)
{
    extern __shared__ double4 pbSmem[];
    double4* sCoef = pbSmem;                                            // [lmax+2] {g0, g20, g22, c22} at l
    double* sW = reinterpret_cast<double*>(sCoef + (lmax + 2));         // [lmax+1][4][PB_BT]
    double* sI = sW + static_cast<size_t>(lmax + 1) * 4 * PB_BT;        // [8][PB_T]
    double* sJ = sI + 8 * PB_T;                                         // [8][PB_T]
    double* stage = sJ + 8 * PB_T;                                      // [PB_BT][3][PB_T][PB_STAGE_LD]
    double** sColPtr = reinterpret_cast<double**>(stage + PB_BT * 3 * PB_T * PB_STAGE_LD);   // [3][PB_T]
    double** sRowPtr = sColPtr + 3 * PB_T;                              // [3][PB_T]

    constexpr int NT = PB_T * PB_CW;
    const long long npix = geo.npix;
    const long long rowBlock = static_cast<long long>(blockIdx.x) * PB_T;
    const long long c0 = static_cast<long long>(blockIdx.y) * PB_T;
    const long long c1 = min(c0 + static_cast<long long>(PB_T), npix);
    if(rowBlock > c1 - 1)
        return;
    const int tid = threadIdx.x;
    const int n1 = lmax + 1;

    for(int k = tid; k <= lmax + 1; k += NT)
        sCoef[k] = make_double4(tab.g0[k], tab.g20[k], tab.g22[k], tab.c22[k]);
    for(int idx = tid; idx < 2 * PB_T; idx += NT)
    {
        const bool isRow = idx < PB_T;
        const int loc = isRow ? idx : idx - PB_T;
        const long long pix = min(isRow ? rowBlock + loc : c0 + loc, npix - 1);
        double* dst = isRow ? sI : sJ;
        dst[0 * PB_T + loc] = geo.nx[pix];
        dst[1 * PB_T + loc] = geo.ny[pix];
        dst[2 * PB_T + loc] = geo.nz[pix];
        dst[3 * PB_T + loc] = geo.tx[pix];
        dst[4 * PB_T + loc] = geo.ty[pix];
        dst[5 * PB_T + loc] = geo.tz[pix];
        dst[6 * PB_T + loc] = geo.px[pix];
        dst[7 * PB_T + loc] = geo.py[pix];
    }
    for(int idx = tid; idx < 6 * PB_T; idx += NT)
    {
        if(idx < 3 * PB_T)
        {
            const int strip = idx / PB_T;
            const long long jcol = min(c0 + (idx - strip * PB_T), npix - 1);
            sColPtr[idx] = partEntry(P, 0, strip, npix, jcol, 0);
        }
        else
        {
            const int q = idx - 3 * PB_T;
            const int t = q / PB_T;
            const long long ir = min(rowBlock + (q - t * PB_T), npix - 1);
            double* dst;
            if(t == 0) dst = partEntry(P, 0, 1, npix, ir, c0);
            else if(t == 1) dst = partEntry(P, 0, 2, npix, ir, c0);
            else dst = partEntry(P, 0, 2, npix, ir, npix + c0);
            sRowPtr[q] = dst;
        }
    }
    __syncthreads();

    const int il = tid & 31, warp = tid >> 5;
    const long long i = rowBlock + il;
    const double nix = sI[0 * PB_T + il], niy = sI[1 * PB_T + il], niz = sI[2 * PB_T + il];
    const double tix = sI[3 * PB_T + il], tiy = sI[4 * PB_T + il], tiz = sI[5 * PB_T + il];
    const double pix_ = sI[6 * PB_T + il], piy = sI[7 * PB_T + il];
    const int nChunks = (nBatch + PB_BT - 1) / PB_BT;
    const int chunkDoubles = n1 * 4 * PB_BT;

    for(int pass = 0; pass < PB_T / PB_CW; ++pass)
    {
        if(c0 + pass * PB_CW >= c1)
            break;                                              // uniform over the CTA
        const int jl = pass * PB_CW + warp;
        double x2, fTQ, fTU, fQT, fUT, reU, imU, reV, imV;
        {
            const double njx = sJ[0 * PB_T + jl], njy = sJ[1 * PB_T + jl], njz = sJ[2 * PB_T + jl];
            const double tjx = sJ[3 * PB_T + jl], tjy = sJ[4 * PB_T + jl], tjz = sJ[5 * PB_T + jl];
            const double pjx = sJ[6 * PB_T + jl], pjy = sJ[7 * PB_T + jl];
            double dot = __dadd_rn(__dadd_rn(__dmul_rn(nix, njx), __dmul_rn(niy, njy)), __dmul_rn(niz, njz));
            dot = fmin(1.0, fmax(-1.0, dot));
            x2 = dot + dot;
            const double ai = fma(njx, tix, fma(njy, tiy, njz * tiz));
            const double bi = fma(njx, pix_, njy * piy);
            const double aj = fma(nix, tjx, fma(niy, tjy, niz * tjz));
            const double bj = fma(nix, pjx, niy * pjy);
            const double p = fma(tix, tjx, fma(tiy, tjy, tiz * tjz));
            const double q = fma(pix_, pjx, piy * pjy);
            const double rr = fma(pix_, tjx, piy * tjy);
            const double tq = fma(tix, pjx, tiy * pjy);
            const double su = p + q, du = rr - tq, sv = p - q, dv = tq + rr;
            reU = fma(su, su, -du * du); imU = 2.0 * su * du;
            reV = fma(sv, sv, -dv * dv); imV = 2.0 * sv * dv;
            fTQ = fma(aj, aj, -bj * bj); fTU = 2.0 * aj * bj;
            fQT = fma(ai, ai, -bi * bi); fUT = 2.0 * ai * bi;
        }
        const int j32 = static_cast<int>(c0 - rowBlock) + jl;
        const bool valid = il <= j32 && c0 + jl < c1;
        double* colT = sColPtr[0 * PB_T + jl] + i;
        double* colQ = sColPtr[1 * PB_T + jl] + i;
        double* colU = sColPtr[2 * PB_T + jl] + i;

        for(int chunk = 0; chunk < nChunks; ++chunk)
        {
            __syncthreads();                                    // previous chunk's weights and stage are free
            {
                const double2* src = reinterpret_cast<const double2*>(folded + static_cast<long long>(chunk) * chunkDoubles);
                double2* dst = reinterpret_cast<double2*>(sW);
                for(int idx = tid; idx < chunkDoubles / 2; idx += NT)
                    dst[idx] = __ldg(src + idx);
            }
            __syncthreads();

            double acc[4][PB_BT];
            // l = 0 and l = 1: temperature only (phi_0 = 1, phi_1 = x2)
#pragma unroll
            for(int b = 0; b < PB_BT; ++b)
            {
                acc[0][b] = fma(x2, sW[(1 * 4 + 0) * PB_BT + b], sW[(0 * 4 + 0) * PB_BT + b]);
                acc[1][b] = 0.0;
                acc[2][b] = 0.0;
                acc[3][b] = 0.0;
            }
            // basis values at l and l-1: P_l family continues from l = 1, spin-2 families start with phi_1 = 0, phi_2 = 1
            double t0 = x2, t1 = fma(x2, x2, -sCoef[1].x);
            double e0 = 0.0, e1 = 1.0, p0 = 0.0, p1 = 1.0, m0 = 0.0, m1 = 1.0;
#pragma unroll 2
            for(int l = 2; l <= lmax; ++l)
            {
                const double2* wl = reinterpret_cast<const double2*>(sW + static_cast<size_t>(l) * 4 * PB_BT);
#pragma unroll
                for(int b = 0; b < PB_BT; b += 2)
                {
                    const double2 wt = wl[(0 * PB_BT + b) / 2];
                    acc[0][b] = fma(t1, wt.x, acc[0][b]);
                    acc[0][b + 1] = fma(t1, wt.y, acc[0][b + 1]);
                }
#pragma unroll
                for(int b = 0; b < PB_BT; b += 2)
                {
                    const double2 we = wl[(1 * PB_BT + b) / 2];
                    acc[1][b] = fma(e1, we.x, acc[1][b]);
                    acc[1][b + 1] = fma(e1, we.y, acc[1][b + 1]);
                }
#pragma unroll
                for(int b = 0; b < PB_BT; b += 2)
                {
                    const double2 wp = wl[(2 * PB_BT + b) / 2];
                    acc[2][b] = fma(p1, wp.x, acc[2][b]);
                    acc[2][b + 1] = fma(p1, wp.y, acc[2][b + 1]);
                }
#pragma unroll
                for(int b = 0; b < PB_BT; b += 2)
                {
                    const double2 wm = wl[(3 * PB_BT + b) / 2];
                    acc[3][b] = fma(m1, wm.x, acc[3][b]);
                    acc[3][b + 1] = fma(m1, wm.y, acc[3][b + 1]);
                }
                const double4 c = sCoef[l];                     // phi_{l+1} = (x2 - c_l) phi_l - g_l phi_{l-1}
                const double tn = fma(x2, t1, -c.x * t0);
                const double en = fma(x2, e1, -c.y * e0);
                const double pn = fma(x2 - c.w, p1, -c.z * p0);
                const double mn = fma(x2 + c.w, m1, -c.z * m0);
                t0 = t1; t1 = tn;
                e0 = e1; e1 = en;
                p0 = p1; p1 = pn;
                m0 = m1; m1 = mn;
            }

            // rotation + stores, one batch element after the other; running pointers instead of 64-bit multiplies
            {
                const long long first = static_cast<long long>(chunk) * PB_BT * outStride;
                double* pT = colT + first;
                double* pQ = colQ + first;
                double* pU = colU + first;
                double* st = stage + il * PB_STAGE_LD + warp;
                const int nLive = min(PB_BT, nBatch - chunk * PB_BT);
#pragma unroll
                for(int b = 0; b < PB_BT; ++b)
                {
                    const double xt = -acc[1][b];
                    const double aRe = acc[2][b] * reU, aIm = acc[2][b] * imU;
                    const double bRe = acc[3][b] * reV, bIm = acc[3][b] * imV;
                    if(valid && b < nLive)
                    {
                        __stcs(pT, acc[0][b]);
                        __stcs(pQ, xt * fTQ);
                        __stcs(pQ + npix, aRe + bRe);
                        __stcs(pU, xt * fTU);
                        __stcs(pU + npix, bIm - aIm);
                        __stcs(pU + 2 * npix, aRe - bRe);
                    }
                    st[0] = xt * fQT;
                    st[1 * PB_T * PB_STAGE_LD] = xt * fUT;
                    st[2 * PB_T * PB_STAGE_LD] = aIm + bIm;
                    st += 3 * PB_T * PB_STAGE_LD;
                    pT += outStride;
                    pQ += outStride;
                    pU += outStride;
                }
            }
            __syncthreads();
            // transposed partners of the chunk: thread (row ilr, column c8) walks the 3 x PB_BT (kind, batch) values
            {
                const int c8 = tid & (PB_CW - 1);
                const int ilr = tid >> 3;                       // 0 .. 31
                const int jlw = pass * PB_CW + c8;
                const int j32w = static_cast<int>(c0 - rowBlock) + jlw;
                if(c0 + jlw < c1 && ilr < j32w)
                {
                    const int nLive = min(PB_BT, nBatch - chunk * PB_BT);
                    const long long first = static_cast<long long>(chunk) * PB_BT * outStride + jlw;
                    double* d0 = sRowPtr[0 * PB_T + ilr] + first;
                    double* d1 = sRowPtr[1 * PB_T + ilr] + first;
                    double* d2 = sRowPtr[2 * PB_T + ilr] + first;
                    const double* sv = stage + ilr * PB_STAGE_LD + c8;
                    for(int b = 0; b < nLive; ++b)
                    {
                        __stcs(d0, sv[0]);
                        __stcs(d1, sv[1 * PB_T * PB_STAGE_LD]);
                        __stcs(d2, sv[2 * PB_T * PB_STAGE_LD]);
                        sv += 3 * PB_T * PB_STAGE_LD;
                        d0 += outStride;
                        d1 += outStride;
                        d2 += outStride;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batched T,Q,U on the FP64 tensor path.  For a batch the sum over l is a contraction: per family s,
//     Out_s[pair, b] = sum_l Phi_s[pair, l] * W_s[l, b],
// so it runs as DMMA (mma.sync m8n8k4 f64; measured 37.0 TFLOP/s on B200, the same peak as DFMA, at 1/8 of the
// instruction count and without the register-bandwidth limit of three-operand DFMAs).
// CTA = 8 x 8 pixel pairs (64), 256 threads:
//   A) every thread runs ONE family's forward recurrence for one pair and leaves Phi[s][pair][l] in shared memory
//      (a basis value is computed once per pair for the whole batch);
//   B) each warp takes batch chunks of 8 elements: A fragments from shared memory, B fragments (weights,
//      pre-arranged by foldMmaWeightsKernel in fragment order) straight from L2, 8 m-tiles x 4 families of
//      accumulators in registers;
//   C) rotation per (pair, element) with the pair's geometric factors from shared memory, natural entries stored
//      directly (runs of 8 rows), transposed partners through a warp-private staging tile (runs of 8 columns).
// ------------------------------------------------------------------------------------------------
constexpr int MB_T = 8;                        // tile edge: MB_T rows x MB_T columns of pixel pairs
constexpr int MB_PAIRS = MB_T * MB_T;
constexpr int MB_THREADS = 256;
constexpr int MB_WARPS = MB_THREADS / 32;
constexpr int MB_BN = 8;                       // batch elements per chunk (one n-tile)
constexpr int MB_SP = MB_PAIRS + 2;            // staging row stride: (element, kind) rows land in distinct banks

__host__ __device__ inline int mmaKPad(int lmax) { return (lmax + 1 + 3) / 4 * 4; }
__host__ __device__ inline int mmaLd(int lmax)
{
    const int kp = mmaKPad(lmax);
    return kp + ((4 - kp % 16) + 16) % 16;     // row stride = 4 (mod 16) doubles: conflict-free A-fragment loads
}
__host__ __device__ inline size_t tquMmaSmemBytes(int lmax)
{
    return sizeof(double) * (static_cast<size_t>(4) * MB_PAIRS * mmaLd(lmax)     // Phi
                             + MB_PAIRS * 8                                       // rotation factors per pair
                             + 16 * MB_T                                          // frames of rows and columns
                             + MB_WARPS * MB_BN * 3 * MB_SP                       // staging of transposed partners
                             + 6 * MB_T                                           // destination pointers
                             + 4 * (mmaKPad(lmax) + 1));                          // recurrence coefficients (phase A only)
}

// w[b][4][lmax+1] -> frag[chunk][kk][family][lane] = folded weight of l = 4 kk + lane % 4, element b = 8 chunk + lane / 4
__global__ void foldMmaWeightsKernel(const double* __restrict__ w, DeviceTables tab, int lmax, int nBatch, double* __restrict__ frag)
{
    const int n1 = lmax + 1;
    const int nkk = mmaKPad(lmax) / 4;
    const long long total = static_cast<long long>((nBatch + MB_BN - 1) / MB_BN) * nkk * 4 * 32;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const int lane = static_cast<int>(idx & 31);
        const int fam = static_cast<int>((idx >> 5) & 3);
        const int kk = static_cast<int>((idx >> 7) % nkk);
        const int chunk = static_cast<int>((idx >> 7) / nkk);
        const int l = 4 * kk + (lane & 3);
        const int b = chunk * MB_BN + (lane >> 2);
        double v = 0.0;
        if(b < nBatch && l <= lmax)
        {
            const double* wb = w + static_cast<long long>(b) * 4 * n1;
            if(fam == 0) v = wb[l] * tab.N0[l];
            else if(l >= 2)
            {
                if(fam == 1) v = wb[n1 + l] * tab.N20[l] * 0.61237243569579452455;
                else if(fam == 2) v = (wb[2 * n1 + l] + wb[3 * n1 + l]) * tab.N22[l] * 0.125;
                else v = (wb[2 * n1 + l] - wb[3 * n1 + l]) * tab.N22[l] * 0.125;
            }
        }
        frag[idx] = v;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// LDC: compile-time row stride of Phi (0 = take it from lmax at run time); a constant stride turns the A-fragment
// addresses into immediates
template <int LDC>
__global__ void __launch_bounds__(MB_THREADS, 1)
tquBatchedMmaKernel(Geometry geo, const double* __restrict__ frag, DeviceTables tab, int lmax, int nBatch,
                    const __grid_constant__ PartTable P, long long outStride)
{
    extern __shared__ double mbSmem[];
    const int ld = LDC ? LDC : mmaLd(lmax);
    const int nkk = mmaKPad(lmax) / 4;
    double* sPhi = mbSmem;                                         // [4][MB_PAIRS][ld]
    double* sFac = sPhi + static_cast<size_t>(4) * MB_PAIRS * ld;  // [MB_PAIRS][8]
    double* sI = sFac + MB_PAIRS * 8;                              // [8][MB_T]
    double* sJ = sI + 8 * MB_T;                                    // [8][MB_T]
    double* sStage = sJ + 8 * MB_T;                                // [MB_WARPS][MB_BN][3][MB_SP]
    double** sColPtr = reinterpret_cast<double**>(sStage + MB_WARPS * MB_BN * 3 * MB_SP);      // [3][MB_T]
    double** sRowPtr = sColPtr + 3 * MB_T;                         // [3][MB_T]
    double* sTab = reinterpret_cast<double*>(sRowPtr + 3 * MB_T);  // [4][kp + 1]: g0, g20, g22, c22

    const long long npix = geo.npix;
    const long long rowBlock = static_cast<long long>(blockIdx.x) * MB_T;
    const long long c0 = static_cast<long long>(blockIdx.y) * MB_T;
    const long long c1 = min(c0 + static_cast<long long>(MB_T), npix);
    if(rowBlock > c1 - 1)
        return;
    const int tid = threadIdx.x;
    const int tabLd = 4 * nkk + 1;
    for(int k = tid; k < 4 * tabLd; k += MB_THREADS)
    {
        const int which = k / tabLd, l = k - which * tabLd;
        const double* src = which == 0 ? tab.g0 : (which == 1 ? tab.g20 : (which == 2 ? tab.g22 : tab.c22));
        sTab[k] = src[l];
    }

    if(tid < 2 * MB_T)
    {
        const bool isRow = tid < MB_T;
        const int loc = isRow ? tid : tid - MB_T;
        const long long pix = min(isRow ? rowBlock + loc : c0 + loc, npix - 1);
        double* dst = isRow ? sI : sJ;
        dst[0 * MB_T + loc] = geo.nx[pix];
        dst[1 * MB_T + loc] = geo.ny[pix];
        dst[2 * MB_T + loc] = geo.nz[pix];
        dst[3 * MB_T + loc] = geo.tx[pix];
        dst[4 * MB_T + loc] = geo.ty[pix];
        dst[5 * MB_T + loc] = geo.tz[pix];
        dst[6 * MB_T + loc] = geo.px[pix];
        dst[7 * MB_T + loc] = geo.py[pix];
    }
    else if(tid < 2 * MB_T + 6 * MB_T)
    {
        const int idx = tid - 2 * MB_T;
        if(idx < 3 * MB_T)
        {
            const int strip = idx / MB_T;
            const long long jcol = min(c0 + (idx - strip * MB_T), npix - 1);
            sColPtr[idx] = partEntry(P, 0, strip, npix, jcol, 0);
        }
        else
        {
            const int q = idx - 3 * MB_T;
            const int t = q / MB_T;
            const long long ir = min(rowBlock + (q - t * MB_T), npix - 1);
            double* dst;
            if(t == 0) dst = partEntry(P, 0, 1, npix, ir, c0);
            else if(t == 1) dst = partEntry(P, 0, 2, npix, ir, c0);
            else dst = partEntry(P, 0, 2, npix, ir, npix + c0);
            sRowPtr[q] = dst;
        }
    }
    __syncthreads();

    // ---- A) basis values and rotation factors.  pair p: row il = p % MB_T, column jl = p / MB_T
    {
        const int p = tid & (MB_PAIRS - 1);
        const int fam = tid >> 6;
        const int il = p % MB_T, jl = p / MB_T;
        const double nix = sI[0 * MB_T + il], niy = sI[1 * MB_T + il], niz = sI[2 * MB_T + il];
        const double njx = sJ[0 * MB_T + jl], njy = sJ[1 * MB_T + jl], njz = sJ[2 * MB_T + jl];
        double dot = __dadd_rn(__dadd_rn(__dmul_rn(nix, njx), __dmul_rn(niy, njy)), __dmul_rn(niz, njz));
        dot = fmin(1.0, fmax(-1.0, dot));
        const double x2 = dot + dot;
        double* phi = sPhi + (static_cast<size_t>(fam) * MB_PAIRS + p) * ld;
        const int kp = 4 * nkk;
        if(fam == 0)
        {
            double q0 = 1.0, q1 = x2;
            phi[0] = q0;
            if(kp > 1) phi[1] = q1;
            for(int l = 1; l + 1 < kp; ++l)
            {
                const double qn = fma(x2, q1, -sTab[l] * q0);
                q0 = q1; q1 = qn;
                phi[l + 1] = (l + 1 <= lmax) ? qn : 0.0;
            }
        }
        else
        {
            const double* g = sTab + (fam == 1 ? 1 : 2) * tabLd;
            const double* cc = sTab + 3 * tabLd;
            const double sgn = fam == 2 ? 1.0 : (fam == 3 ? -1.0 : 0.0);
            double q0 = 0.0, q1 = 1.0;
            phi[0] = 0.0;
            if(kp > 1) phi[1] = 0.0;
            if(kp > 2) phi[2] = lmax >= 2 ? 1.0 : 0.0;
            for(int l = 2; l + 1 < kp; ++l)
            {
                const double qn = fma(x2 - sgn * cc[l], q1, -g[l] * q0);
                q0 = q1; q1 = qn;
                phi[l + 1] = (l + 1 <= lmax) ? qn : 0.0;
            }
        }
        if(fam == 0)
        {
            const double tix = sI[3 * MB_T + il], tiy = sI[4 * MB_T + il], tiz = sI[5 * MB_T + il];
            const double pix_ = sI[6 * MB_T + il], piy = sI[7 * MB_T + il];
            const double tjx = sJ[3 * MB_T + jl], tjy = sJ[4 * MB_T + jl], tjz = sJ[5 * MB_T + jl];
            const double pjx = sJ[6 * MB_T + jl], pjy = sJ[7 * MB_T + jl];
            const double ai = fma(njx, tix, fma(njy, tiy, njz * tiz));
            const double bi = fma(njx, pix_, njy * piy);
            const double aj = fma(nix, tjx, fma(niy, tjy, niz * tjz));
            const double bj = fma(nix, pjx, niy * pjy);
            const double pp = fma(tix, tjx, fma(tiy, tjy, tiz * tjz));
            const double qq = fma(pix_, pjx, piy * pjy);
            const double rr = fma(pix_, tjx, piy * tjy);
            const double tq = fma(tix, pjx, tiy * pjy);
            const double su = pp + qq, du = rr - tq, sv = pp - qq, dv = tq + rr;
            double* fac = sFac + p * 8;
            fac[0] = fma(aj, aj, -bj * bj);     // fTQ
            fac[1] = 2.0 * aj * bj;             // fTU
            fac[2] = fma(ai, ai, -bi * bi);     // fQT
            fac[3] = 2.0 * ai * bi;             // fUT
            fac[4] = fma(su, su, -du * du);     // reU
            fac[5] = 2.0 * su * du;             // imU
            fac[6] = fma(sv, sv, -dv * dv);     // reV
            fac[7] = 2.0 * sv * dv;             // imV
        }
    }
    __syncthreads();

    // ---- B) + C) per warp and batch chunk
    const int lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;           // fragment row (pair within the m-tile) and column
    const int nChunks = (nBatch + MB_BN - 1) / MB_BN;
    double* stage = sStage + warp * (MB_BN * 3 * MB_SP);
    const int iOff = static_cast<int>(c0 - rowBlock);   // j - rowBlock = iOff + jl

    for(int chunk = warp; chunk < nChunks; chunk += MB_WARPS)
    {
        double acc[MB_T][4][2];
#pragma unroll
        for(int t = 0; t < MB_T; ++t)
#pragma unroll
            for(int s = 0; s < 4; ++s)
                acc[t][s][0] = acc[t][s][1] = 0.0;
        // B fragments of the whole chunk (nkk x 4 families x 32 lanes, <= the staging tile in size) are copied from L2
        // into this warp's staging region first, all loads in flight at once, so the L2 latency is paid once per chunk
        // instead of once per k-step; the region is reused for the transposed partners after the MMA loop.
        {
            const double* fb = frag + (static_cast<long long>(chunk) * nkk * 4) * 32 + lane;
            const int nFrag = nkk * 4;
            for(int q0 = 0; q0 < nFrag; q0 += 12)
            {
                double v[12];
#pragma unroll
                for(int u = 0; u < 12; ++u)
                    v[u] = (q0 + u < nFrag) ? __ldg(fb + (q0 + u) * 32) : 0.0;
#pragma unroll
                for(int u = 0; u < 12; ++u)
                    if(q0 + u < nFrag)
                        stage[(q0 + u) * 32 + lane] = v[u];
            }
        }
        __syncwarp();
        const double* paBase = sPhi + static_cast<size_t>(fr) * ld + fc;
#pragma unroll 1
        for(int kk = 0; kk < nkk; ++kk)
        {
#pragma unroll
            for(int s = 0; s < 4; ++s)
            {
                const double bfrag = stage[(kk * 4 + s) * 32 + lane];
                const double* pa = paBase + static_cast<size_t>(s) * MB_PAIRS * ld + 4 * kk;
#pragma unroll
                for(int t = 0; t < MB_T; ++t)
                    dmma884(acc[t][s][0], acc[t][s][1], pa[t * MB_T * ld], bfrag);
            }
        }
        __syncwarp();                                       // every lane is done reading the fragments

        // rotation and stores: m-tile t is column jl = t, this thread's pair is row il = fr, elements 2 fc, 2 fc + 1
        const int b0 = chunk * MB_BN + 2 * fc;
        const long long i = rowBlock + fr;
#pragma unroll
        for(int t = 0; t < MB_T; ++t)
        {
            const int p = t * MB_T + fr;
            const double4 f0 = *reinterpret_cast<const double4*>(sFac + p * 8);
            const double4 f1 = *reinterpret_cast<const double4*>(sFac + p * 8 + 4);
            const bool valid = fr <= iOff + t && c0 + t < c1;
            double* colT = sColPtr[0 * MB_T + t] + i;
            double* colQ = sColPtr[1 * MB_T + t] + i;
            double* colU = sColPtr[2 * MB_T + t] + i;
#pragma unroll
            for(int e = 0; e < 2; ++e)
            {
                const double xt = -acc[t][1][e];
                const double aRe = acc[t][2][e] * f1.x, aIm = acc[t][2][e] * f1.y;
                const double bRe = acc[t][3][e] * f1.z, bIm = acc[t][3][e] * f1.w;
                if(valid && b0 + e < nBatch)
                {
                    const long long boff = static_cast<long long>(b0 + e) * outStride;
                    __stcs(colT + boff, acc[t][0][e]);
                    __stcs(colQ + boff, xt * f0.x);
                    __stcs(colQ + boff + npix, aRe + bRe);
                    __stcs(colU + boff, xt * f0.y);
                    __stcs(colU + boff + npix, bIm - aIm);
                    __stcs(colU + boff + 2 * npix, aRe - bRe);
                }
                double* st = stage + ((2 * fc + e) * 3) * MB_SP + p;
                st[0] = xt * f0.z;                       // Q_i T_j
                st[MB_SP] = xt * f0.w;                   // U_i T_j
                st[2 * MB_SP] = aIm + bIm;               // U_i Q_j
            }
        }
        __syncwarp();
        // transposed partners: lanes along j (8 consecutive rows of column N+i / 2N+i), 4 rows i per store instruction;
        // lane -> column jl = lane & 7 and rows il = (lane >> 3) + 4 h, running pointers step from element to element
        {
            const int jl = lane & (MB_T - 1);
            const bool colOk = c0 + jl < c1;
            const int nLive = min(MB_BN, nBatch - chunk * MB_BN);
#pragma unroll
            for(int h = 0; h < 2; ++h)
            {
                const int il = (lane >> 3) + 4 * h;
                if(colOk && il < iOff + jl)
                {
                    const long long first = static_cast<long long>(chunk) * MB_BN * outStride + jl;
                    double* d0 = sRowPtr[0 * MB_T + il] + first;
                    double* d1 = sRowPtr[1 * MB_T + il] + first;
                    double* d2 = sRowPtr[2 * MB_T + il] + first;
                    const double* sv = stage + jl * MB_T + il;
                    for(int bl = 0; bl < nLive; ++bl)
                    {
                        __stcs(d0, sv[0]);
                        __stcs(d1, sv[MB_SP]);
                        __stcs(d2, sv[2 * MB_SP]);
                        sv += 3 * MB_SP;
                        d0 += outStride;
                        d1 += outStride;
                        d2 += outStride;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Batched T,Q,U on the FP64 tensor path, slab output.
//
// Why a second output format.  tquBatchedMmaKernel and two successors (register-resident basis; warp-specialised with
// TMA-fed weight ring and store warps -- profiles/r1_kernel_history.md) all ended at ~25 ms per 128 Nside=16 matrices,
// which tools/store_pattern.cu reproduces with stores alone: an m8n8k4 accumulator tile holds 8 pixel rows x 8 batch
// elements, so the packed triangles receive 64-byte runs scattered over all matrices of the batch, and HBM takes such
// writes at 1.7 TB/s (256-byte runs into 16 matrices at a time: 4.8 TB/s; fill: 7.5 TB/s).  The accumulator tile is
// contiguous along the BATCH axis, so that is the axis to make contiguous in memory:
//
//   slab = 16 consecutive batch elements interleaved entry by entry:   slab[e * 16 + (b % 16)],  e = packed index
//
// i.e. each element of a slab is an ordinary packed CMatrix with element stride 16.  A warp's store instruction then
// writes 1 KB contiguous (8 rows x 16 elements, natural entries) or 8 full 128-byte lines (transposed partners),
// always sector-aligned, 32 bytes per lane.  cmg_slab_unpack converts a slab into 16 separate packed matrices.
//
//   * two independent passes (blockIdx.z): {TT, TE} -> T_iT_j, T_iQ_j, T_iU_j, Q_iT_j, U_iT_j and
//     {EE+BB, EE-BB} -> Q_iQ_j, Q_iU_j, U_iU_j, U_iQ_j: a thread needs the basis of two families only;
//   * CTA = 8 x 8 pixel pairs, warp w = column j = c0 + w, fragment row = pixel row i; the A fragments of the warp's
//     m-tile (2 families x NKK k-steps of 4 multipoles) are computed once (phase A, through shared memory) and stay
//     in REGISTERS for the whole batch; the MMA loop reads only weight fragments (one LDS.128 per two DMMAs);
//   * weight fragments of the next chunks arrive by TMA bulk copies into a 4-stage ring shared by the 8 warps
//     (full / empty mbarriers; no CTA-wide barrier inside the batch loop);
//   * the columns of the weight fragments are permuted (element 4 (n / 2) + 2 nt + n % 2 in column n of n-tile nt) so
//     that a lane's four accumulators of a family are four consecutive batch elements: one 256-bit store per entry;
//   * block order: 16 adjacent column tiles per row tile, so neighbouring lines are written at about the same time.
// ------------------------------------------------------------------------------------------------
constexpr int M2_T = 8;
constexpr int M2_PAIRS = M2_T * M2_T;
constexpr int M2_THREADS = 256;
constexpr int M2_BC = CMG_SLAB;                // batch elements per chunk = per slab: two n-tiles
constexpr int M2_RING = 4;                    // weight-fragment stages (TMA bulk copies, full/empty mbarriers)
constexpr int M2_GROUP = 16;                   // column tiles walked together
static_assert(CMG_SLAB == 16, "the fragment permutation below assumes 16-element slabs");

template <int NKK>
struct M2Shape
{
    static constexpr int KP = 4 * NKK;                                   // padded number of multipoles
    static constexpr int LD = KP + ((4 - KP % 16) + 16) % 16;            // row stride of Phi
    static constexpr int CHUNK = NKK * 2 * 32 * 2;                       // doubles of weight fragments per chunk and pass
    static constexpr int PHI = 2 * M2_PAIRS * LD;
    static constexpr int LOOP = M2_RING * CHUNK;
    static constexpr int X = PHI + LOOP;                                 // ring and Phi side by side: the first weight
                                                                         // fragments travel while phase A computes
    static constexpr size_t BYTES = sizeof(double) * (X + M2_PAIRS * 4 + 16 * M2_T + 4 * (KP + 1) + 1 + 2 * M2_RING);
};

// w[b][4][lmax+1] -> frag[pass][chunk][kk][f][lane][nt]: folded weight of family 2 pass + f, l = 4 kk + lane % 4,
// element b = 16 chunk + 4 (n / 2) + 2 nt + n % 2 with n = lane / 4 the fragment column
__global__ void foldSlabWeightsKernel(const double* __restrict__ w, DeviceTables tab, int lmax, int nBatch, int nkk, double* __restrict__ frag)
{
    const int n1 = lmax + 1;
    const int nChunks = (nBatch + M2_BC - 1) / M2_BC;
    const long long total = 2LL * nChunks * nkk * 128;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const int nt = static_cast<int>(idx & 1);
        const int lane = static_cast<int>((idx >> 1) & 31);
        const int f = static_cast<int>((idx >> 6) & 1);
        const long long rest = idx >> 7;
        const int kk = static_cast<int>(rest % nkk);
        const int chunk = static_cast<int>((rest / nkk) % nChunks);
        const int pass = static_cast<int>(rest / nkk / nChunks);
        const int fam = 2 * pass + f;
        const int l = 4 * kk + (lane & 3);
        const int n = lane >> 2;
        const int b = chunk * M2_BC + 4 * (n >> 1) + 2 * nt + (n & 1);
        double v = 0.0;
        if(b < nBatch && l <= lmax)
        {
            const double* wb = w + static_cast<long long>(b) * 4 * n1;
            if(fam == 0) v = wb[l] * tab.N0[l];
            else if(l >= 2)
            {
                if(fam == 1) v = wb[n1 + l] * tab.N20[l] * 0.61237243569579452455;
                else if(fam == 2) v = (wb[2 * n1 + l] + wb[3 * n1 + l]) * tab.N22[l] * 0.125;
                else v = (wb[2 * n1 + l] - wb[3 * n1 + l]) * tab.N22[l] * 0.125;
            }
        }
        frag[idx] = v;
    }
}

__device__ __forceinline__ unsigned smemAddr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarArrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while(!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulkLoad(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// four consecutive batch elements of one entry: one 256-bit streaming store (STG.E.EF.ENL2.256), 32-byte aligned by
// construction; whole sectors are written, so evict-first costs no read-modify-write and keeps L2 for the weight fragments
// (16.9 vs 17.2 ms per 256 matrices against the default policy)
__device__ __forceinline__ void storeQuad(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int NKK, int PASS>
__device__ __forceinline__ void slabBody(double* smem, const Geometry& geo, const double* __restrict__ frag, const DeviceTables& tab,
                                         int lmax, int nBatch, double* __restrict__ out, long long slabDoubles,
                                         long long rowTile, long long colTile)
{
    using S = M2Shape<NKK>;
    constexpr int KP = S::KP, LD = S::LD;
    double* sX = smem;
    double* sFac = sX + S::X;                                      // [M2_PAIRS][4]
    double* sI = sFac + M2_PAIRS * 4;                              // [8][M2_T]
    double* sJ = sI + 8 * M2_T;
    double* sTab = sJ + 8 * M2_T;                                  // [4][KP + 1]: g0, g20, g22, c22
    const unsigned bFull = smemAddr(sTab + 4 * (KP + 1) + ((4 * (KP + 1)) & 1));      // [M2_RING] full, then [M2_RING] empty
    const unsigned bEmpty = bFull + 8 * M2_RING;

    const long long npix = geo.npix;
    const long long rowBlock = rowTile * M2_T;
    const long long c0 = colTile * M2_T;
    const long long c1 = min(c0 + static_cast<long long>(M2_T), npix);
    const int tid = threadIdx.x;
    constexpr int tabLd = KP + 1;
    for(int k = tid; k < 4 * tabLd; k += M2_THREADS)
    {
        const int which = k / tabLd, l = k - which * tabLd;
        const double* src = which == 0 ? tab.g0 : (which == 1 ? tab.g20 : (which == 2 ? tab.g22 : tab.c22));
        sTab[k] = src[l];
    }
    if(tid < 2 * M2_T)
    {
        const bool isRow = tid < M2_T;
        const int loc = isRow ? tid : tid - M2_T;
        const long long pix = min(isRow ? rowBlock + loc : c0 + loc, npix - 1);
        double* dst = isRow ? sI : sJ;
        dst[0 * M2_T + loc] = geo.nx[pix];
        dst[1 * M2_T + loc] = geo.ny[pix];
        dst[2 * M2_T + loc] = geo.nz[pix];
        dst[3 * M2_T + loc] = geo.tx[pix];
        dst[4 * M2_T + loc] = geo.ty[pix];
        dst[5 * M2_T + loc] = geo.tz[pix];
        dst[6 * M2_T + loc] = geo.px[pix];
        dst[7 * M2_T + loc] = geo.py[pix];
    }
    else if(tid == 2 * M2_T)
    {
        for(int k = 0; k < M2_RING; ++k)
        {
            mbarInit(bFull + 8 * k, 1);
            mbarInit(bEmpty + 8 * k, M2_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double* ring = sX + S::PHI;
    const int nChunks = (nBatch + M2_BC - 1) / M2_BC;
    const double* fragPass = frag + static_cast<long long>(PASS) * nChunks * S::CHUNK;
    constexpr unsigned chunkBytes = S::CHUNK * sizeof(double);
    if(tid == M2_THREADS - 1)                                      // a thread with no phase-A work
    {
        for(int k = 0; k < M2_RING && k < nChunks; ++k)
        {
            mbarExpectTx(bFull + 8 * k, chunkBytes);
            bulkLoad(smemAddr(ring + k * S::CHUNK), fragPass + static_cast<long long>(k) * S::CHUNK, chunkBytes, bFull + 8 * k);
        }
    }

    // ---- A) basis values of this pass's two families (threads 0..127) and the rotation factors (threads 128..191).
    //      pair p: row il = p % M2_T, column jl = p / M2_T
    if(tid < 3 * M2_PAIRS)
    {
        const int p = tid & (M2_PAIRS - 1);
        const int il = p % M2_T, jl = p / M2_T;
        const double nix = sI[0 * M2_T + il], niy = sI[1 * M2_T + il], niz = sI[2 * M2_T + il];
        const double njx = sJ[0 * M2_T + jl], njy = sJ[1 * M2_T + jl], njz = sJ[2 * M2_T + jl];
        if(tid < 2 * M2_PAIRS)
        {
            const int f = tid >> 6;
            const int fam = 2 * PASS + f;
            double dot = __dadd_rn(__dadd_rn(__dmul_rn(nix, njx), __dmul_rn(niy, njy)), __dmul_rn(niz, njz));
            dot = fmin(1.0, fmax(-1.0, dot));
            const double x2 = dot + dot;
            double* phi = sX + (static_cast<size_t>(f) * M2_PAIRS + p) * LD;
            if(fam == 0)
            {
                double q0 = 1.0, q1 = x2;
                phi[0] = q0;
                phi[1] = q1;
                for(int l0 = 1; l0 + 1 < KP; l0 += 4)
                {
                    double g[4];
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                        g[u] = sTab[min(l0 + u, KP)];
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                    {
                        const int l = l0 + u;
                        if(l + 1 < KP)
                        {
                            const double qn = fma(x2, q1, -g[u] * q0);
                            q0 = q1; q1 = qn;
                            phi[l + 1] = (l + 1 <= lmax) ? qn : 0.0;
                        }
                    }
                }
            }
            else
            {
                const double* gt = sTab + (fam == 1 ? 1 : 2) * tabLd;
                const double* ct = sTab + 3 * tabLd;
                const double sgn = fam == 2 ? 1.0 : (fam == 3 ? -1.0 : 0.0);
                double q0 = 0.0, q1 = 1.0;
                phi[0] = 0.0;
                phi[1] = 0.0;
                phi[2] = lmax >= 2 ? 1.0 : 0.0;
                for(int l0 = 2; l0 + 1 < KP; l0 += 4)
                {
                    double g[4], xc[4];
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                    {
                        g[u] = gt[min(l0 + u, KP)];
                        xc[u] = x2 - sgn * ct[min(l0 + u, KP)];
                    }
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                    {
                        const int l = l0 + u;
                        if(l + 1 < KP)
                        {
                            const double qn = fma(xc[u], q1, -g[u] * q0);
                            q0 = q1; q1 = qn;
                            phi[l + 1] = (l + 1 <= lmax) ? qn : 0.0;
                        }
                    }
                }
            }
        }
        else
        {
            const double tix = sI[3 * M2_T + il], tiy = sI[4 * M2_T + il], tiz = sI[5 * M2_T + il];
            const double pix_ = sI[6 * M2_T + il], piy = sI[7 * M2_T + il];
            const double tjx = sJ[3 * M2_T + jl], tjy = sJ[4 * M2_T + jl], tjz = sJ[5 * M2_T + jl];
            const double pjx = sJ[6 * M2_T + jl], pjy = sJ[7 * M2_T + jl];
            double* fac = sFac + p * 4;
            if(PASS == 0)
            {
                const double ai = fma(njx, tix, fma(njy, tiy, njz * tiz));
                const double bi = fma(njx, pix_, njy * piy);
                const double aj = fma(nix, tjx, fma(niy, tjy, niz * tjz));
                const double bj = fma(nix, pjx, niy * pjy);
                fac[0] = -fma(aj, aj, -bj * bj);    // -fTQ  (the TE sum enters with a minus sign)
                fac[1] = -2.0 * aj * bj;            // -fTU
                fac[2] = -fma(ai, ai, -bi * bi);    // -fQT
                fac[3] = -2.0 * ai * bi;            // -fUT
            }
            else
            {
                const double pp = fma(tix, tjx, fma(tiy, tjy, tiz * tjz));
                const double qq = fma(pix_, pjx, piy * pjy);
                const double rr = fma(pix_, tjx, piy * tjy);
                const double tq = fma(tix, pjx, tiy * pjy);
                const double su = pp + qq, du = rr - tq, sv = pp - qq, dv = tq + rr;
                fac[0] = fma(su, su, -du * du);     // reU
                fac[1] = 2.0 * su * du;             // imU
                fac[2] = fma(sv, sv, -dv * dv);     // reV
                fac[3] = 2.0 * sv * dv;             // imV
            }
        }
    }
    __syncthreads();

    // ---- fragments and per-pair constants to registers
    const int lane = tid & 31, w = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    double a[2][NKK];
    {
        const double* src = sX + static_cast<size_t>(w * M2_T + fr) * LD + fc;
#pragma unroll
        for(int f = 0; f < 2; ++f)
#pragma unroll
            for(int kk = 0; kk < NKK; ++kk)
                a[f][kk] = src[static_cast<size_t>(f) * M2_PAIRS * LD + 4 * kk];
    }
    const double2 fa = *reinterpret_cast<const double2*>(sFac + (w * M2_T + fr) * 4);
    const double2 fb = *reinterpret_cast<const double2*>(sFac + (w * M2_T + fr) * 4 + 2);
    const long long i = rowBlock + fr, j = c0 + w;
    const bool natural = j < c1 && i <= j;
    const bool transposed = j < c1 && i < j;
    // element offsets of this lane's four batch elements (4 fc .. 4 fc + 3) of each entry
    const long long eT = (packedOffset(j) + i) * CMG_SLAB + 4 * fc;                      // (T_i, T_j)
    const long long eQ = (packedOffset(npix + j) + i) * CMG_SLAB + 4 * fc;               // (T_i, Q_j); + npix rows: (Q_i, Q_j)
    const long long eU = (packedOffset(2 * npix + j) + i) * CMG_SLAB + 4 * fc;           // (T_i, U_j); + npix: (Q_i, U_j); + 2 npix: (U_i, U_j)
    const long long eQt = (packedOffset(npix + min(i, npix - 1)) + j) * CMG_SLAB + 4 * fc;        // (T_j, Q_i)
    const long long eUt = (packedOffset(2 * npix + min(i, npix - 1)) + j) * CMG_SLAB + 4 * fc;    // (T_j, U_i); + npix: (Q_j, U_i)
    const long long rowN = npix * CMG_SLAB;

    double* slab = out;
#pragma unroll 1
    for(int c = 0; c < nChunks; ++c, slab += slabDoubles)
    {
        const int st = c % M2_RING;
        mbarWait(bFull + 8 * st, (c / M2_RING) & 1);
        double acc[2][2][2];
#pragma unroll
        for(int f = 0; f < 2; ++f)
#pragma unroll
            for(int nt = 0; nt < 2; ++nt)
                acc[f][nt][0] = acc[f][nt][1] = 0.0;
        const double2* bs = reinterpret_cast<const double2*>(ring + st * S::CHUNK) + lane;
#pragma unroll
        for(int kk = 0; kk < NKK; ++kk)
#pragma unroll
            for(int f = 0; f < 2; ++f)
            {
                const double2 bv = bs[(kk * 2 + f) * 32];
                dmma884(acc[f][0][0], acc[f][0][1], a[f][kk], bv.x);
                dmma884(acc[f][1][0], acc[f][1][1], a[f][kk], bv.y);
            }
        __syncwarp();
        if(lane == 0)
        {
            mbarArrive(bEmpty + 8 * st);
            // warp 0 refills the stage of chunk c-1 with chunk c+3 once all eight warps have released it
            if(w == 0 && c >= 1 && c + M2_RING - 1 < nChunks)
            {
                const int sp = (c - 1) % M2_RING;
                mbarWait(bEmpty + 8 * sp, ((c - 1) / M2_RING) & 1);
                mbarExpectTx(bFull + 8 * sp, chunkBytes);
                bulkLoad(smemAddr(ring + sp * S::CHUNK), fragPass + static_cast<long long>(c + M2_RING - 1) * S::CHUNK, chunkBytes, bFull + 8 * sp);
            }
        }
        // acc[f][nt][e] belongs to batch element 4 fc + 2 nt + e of this slab
        if(PASS == 0)
        {
            if(natural)
            {
                storeQuad(slab + eT, acc[0][0][0], acc[0][0][1], acc[0][1][0], acc[0][1][1]);
                storeQuad(slab + eQ, acc[1][0][0] * fa.x, acc[1][0][1] * fa.x, acc[1][1][0] * fa.x, acc[1][1][1] * fa.x);
                storeQuad(slab + eU, acc[1][0][0] * fa.y, acc[1][0][1] * fa.y, acc[1][1][0] * fa.y, acc[1][1][1] * fa.y);
            }
            if(transposed)
            {
                storeQuad(slab + eQt, acc[1][0][0] * fb.x, acc[1][0][1] * fb.x, acc[1][1][0] * fb.x, acc[1][1][1] * fb.x);
                storeQuad(slab + eUt, acc[1][0][0] * fb.y, acc[1][0][1] * fb.y, acc[1][1][0] * fb.y, acc[1][1][1] * fb.y);
            }
        }
        else
        {
            double qq[4], qu[4], uu[4], uq[4];
#pragma unroll
            for(int nt = 0; nt < 2; ++nt)
#pragma unroll
                for(int e = 0; e < 2; ++e)
                {
                    const double s0 = acc[0][nt][e], s1 = acc[1][nt][e];
                    const double aRe = s0 * fa.x, aIm = s0 * fa.y;
                    qq[2 * nt + e] = fma(s1, fb.x, aRe);
                    uu[2 * nt + e] = fma(-s1, fb.x, aRe);
                    qu[2 * nt + e] = fma(s1, fb.y, -aIm);
                    uq[2 * nt + e] = fma(s1, fb.y, aIm);
                }
            if(natural)
            {
                storeQuad(slab + eQ + rowN, qq[0], qq[1], qq[2], qq[3]);
                storeQuad(slab + eU + rowN, qu[0], qu[1], qu[2], qu[3]);
                storeQuad(slab + eU + 2 * rowN, uu[0], uu[1], uu[2], uu[3]);
            }
            if(transposed)
                storeQuad(slab + eUt + rowN, uq[0], uq[1], uq[2], uq[3]);
        }
    }
}

template <int NKK>
__global__ void __launch_bounds__(M2_THREADS, 2)      // 3 CTAs/SM (80 registers, small spills) measured slower: 0.077 vs 0.068 ms/matrix
tquBatchedSlabKernel(Geometry geo, const double* __restrict__ frag, DeviceTables tab, int lmax, int nBatch,
                     double* __restrict__ out, long long slabDoubles)
{
    extern __shared__ __align__(16) double m2Smem[];
    // block order: M2_GROUP adjacent column tiles of one row tile, then the next row tile, then the next group
    const long long nTiles = (geo.npix + M2_T - 1) / M2_T;
    const long long colTile = static_cast<long long>(blockIdx.y) * M2_GROUP + (blockIdx.x % M2_GROUP);
    const long long rowTile = blockIdx.x / M2_GROUP;
    if(colTile >= nTiles || rowTile > colTile)
        return;
    if(blockIdx.z == 0)
        slabBody<NKK, 0>(m2Smem, geo, frag, tab, lmax, nBatch, out, slabDoubles, rowTile, colTile);
    else
        slabBody<NKK, 1>(m2Smem, geo, frag, tab, lmax, nBatch, out, slabDoubles, rowTile, colTile);
}

// slab -> separate packed matrices: out[b * outStride + e] = slab[e * 16 + b], b = 0 .. nLive-1 (onlyB >= 0: that element
// alone, written to out[e]).  32 entries x 16 elements per warp pass through shared memory so that both sides move
// whole lines.
__global__ void __launch_bounds__(256) slabUnpackKernel(const double* __restrict__ slab, long long packed, int nLive, int onlyB,
                                                        double* __restrict__ out, long long outStride)
{
    __shared__ double tile[8][32][CMG_SLAB + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long nBlocks = (packed + 255) / 256;
    for(long long blk = blockIdx.x; blk < nBlocks; blk += gridDim.x)
    {
        const long long e0 = blk * 256 + w * 32;
        // read 32 entries x 16 elements = 512 consecutive doubles
#pragma unroll
        for(int r = 0; r < CMG_SLAB; ++r)
        {
            const int q = r * 32 + lane;
            const long long e = e0 + q / CMG_SLAB;
            if(e < packed)
                tile[w][q / CMG_SLAB][q % CMG_SLAB] = slab[e * CMG_SLAB + q % CMG_SLAB];
        }
        __syncwarp();
        if(e0 + lane < packed)
        {
            if(onlyB >= 0)
                out[e0 + lane] = tile[w][lane][onlyB];
            else
                for(int b = 0; b < nLive; ++b)
                    out[b * outStride + e0 + lane] = tile[w][lane][b];
        }
        __syncwarp();
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
// Places one dense block of a kind-1 ("outbox") shard into a whole packed [T;Q;U] triangle: block kind t holds
// <Q_i T_j> (t=0), <U_i T_j> (t=1) or <U_i Q_j> (t=2) for owner columns i in [col0, col0+nCols) and rows j in
// [row0, row0+ld), element (i - col0) * ld + (j - row0).  Lanes run along j: both sides are contiguous.
// ------------------------------------------------------------------------------------------------
__global__ void scatterBlockKernel(const double* __restrict__ block, long long npix, long long col0, long long nCols,
                                   long long ld, long long row0, int t, double* __restrict__ full)
{
    const long long ic = blockIdx.y;
    if(ic >= nCols)
        return;
    const long long i = col0 + ic;
    const long long col = (t == 0 ? npix : 2 * npix) + i;
    const long long rowShift = (t == 2 ? npix : 0);
    double* dst = full + packedOffset(col) + rowShift + row0;
    const double* src = block + ic * ld;
    for(long long jj = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; jj < ld; jj += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[jj] = src[jj];
}

// ------------------------------------------------------------------------------------------------
// Consumer side (reference source/likelihood.cpp:100-110): C + F + N of three packed matrices (any of the last two may
// be absent) written as a full symmetric column-major n x n matrix, ready for a dense Cholesky.  One pass over HBM
// instead of three matrix reads on the host.  Tile 32 x 32 through shared memory so both triangles are written coalesced.
// ------------------------------------------------------------------------------------------------
// cStride: element stride of c (1 = packed matrix, CMG_SLAB = one element of a slab)
__global__ void __launch_bounds__(256) sumUnpackKernel(const double* __restrict__ c, long long cStride, const double* __restrict__ f, const double* __restrict__ nz,
                                                       long long n, double* __restrict__ full)
{
    __shared__ double tile[32][33];
    const long long bi = blockIdx.x, bj = blockIdx.y;          // row block, column block; upper triangle blocks only
    if(bi > bj)
        return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    for(int r = ty; r < 32; r += 8)
    {
        const long long j = bj * 32 + r, i = bi * 32 + tx;     // column j, row i: lanes along i (contiguous in the packed column)
        double v = 0.0;
        if(j < n && i <= j)
        {
            const long long k = packedOffset(j) + i;
            v = c[k * cStride];
            if(f) v += f[k];
            if(nz) v += nz[k];
            full[j * n + i] = v;                               // upper triangle entry (i, j)
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for(int r = ty; r < 32; r += 8)
    {
        const long long i = bi * 32 + r, j = bj * 32 + tx;     // write (j, i) of the lower triangle: column i, rows j, lanes along j
        if(j < n && i < j)
            full[i * n + j] = tile[tx][r];
    }
}

// chi2 pieces of the pixel likelihood (reference source/likelihood.cpp:163-180) after the triangular solve L y = t:
// yy[k] = sum_i y[i,k]^2 and, with a foreground template, yf[k] = sum_i y[i,k] yF[i].  One block per map, fixed
// summation order (deterministic).
__global__ void __launch_bounds__(256) columnDotsKernel(const double* __restrict__ y, const double* __restrict__ yF, long long n,
                                                        double* __restrict__ yy, double* __restrict__ yf)
{
    __shared__ double s0[256], s1[256];
    const double* col = y + static_cast<long long>(blockIdx.x) * n;
    double a = 0.0, b = 0.0;
    for(long long i = threadIdx.x; i < n; i += 256)
    {
        const double v = col[i];
        a = fma(v, v, a);
        if(yF) b = fma(v, yF[i], b);
    }
    s0[threadIdx.x] = a;
    s1[threadIdx.x] = b;
    __syncthreads();
    for(int h = 128; h > 0; h >>= 1)
    {
        if(threadIdx.x < h)
        {
            s0[threadIdx.x] += s0[threadIdx.x + h];
            s1[threadIdx.x] += s1[threadIdx.x + h];
        }
        __syncthreads();
    }
    if(threadIdx.x == 0)
    {
        yy[blockIdx.x] = s0[0];
        if(yf) yf[blockIdx.x] = s1[0];
    }
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
