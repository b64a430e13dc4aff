// Cholesky factorisation A = U^T U of a symmetric positive definite matrix held as the PACKED upper triangle (entry (i <= j) at
// j (j + 1) / 2 + i), in place on the device -- the storage CMatrix has and the factorisation the reference's consumer runs on
// the host (LAPACK dpptrf 'U' on Math::SymmetricMatrix, reference source/matrix_impl.cpp:236-263, include/matrix_impl.hpp:495-502;
// called from Likelihood::construct, source/likelihood.cpp:100-133).  Nothing is unpacked: the 147456-dimensional [T;Q;U] matrix
// of Nside = 64 (87 GB packed, 174 GB square) is factorised where the generator left it.
//
// Right-looking, block size CH_NB = 128.  Step k (rows k0 .. k0 + kb of U):
//   cholDiagKernel    one CTA: the kb x kb diagonal block in shared memory, U_kk
//   cholPanelKernel   U[k0.., j] = U_kk^-T A[k0.., j] for every column j behind the block: a thread per column, forward
//                     substitution; in packed storage the kb rows of a column are one contiguous run
//   cholSyrkKernel    A[i, j] -= sum_r U[r, i] U[r, j] for k1 <= i <= j: the n^3 / 3 of the work.  This IS a contraction, so it runs
//                     on the FP64 tensor path (mma.sync.m8n8k4.f64; on B200 its rate equals the DFMA rate, 37 TFLOP/s, but an
//                     instruction carries 256 FMAs and the operands come from shared memory once per 128 x 128 tile).  Both
//                     operands are runs of the panel (K contiguous), the same fragment pattern for A and B.
// Solves for chi^2 = |U^-T t|^2 use the same blocking: cholSolveDiagKernel + cholSolveUpdateKernel.
#pragma once

#include <cuda_runtime.h>

namespace cmg
{

constexpr int CH_NB = 128;                 // block size (rows of U per step)
constexpr int CH_LD = CH_NB + 1;           // leading dimension of the diagonal block in shared memory
constexpr int CH_TILE = 128;               // syrk: tile edge
constexpr int CH_KC = 32;                  // syrk: k-chunk staged in shared memory
constexpr int CH_SLD = CH_KC + 4;          // its leading dimension: (lane / 4) * 36 + lane % 4 hits 16 distinct 8-byte banks per half-warp
constexpr int CH_PANEL_COLS = 64;          // panel solve: columns (threads) per CTA

__host__ __device__ __forceinline__ long long chOff(long long col) { return col * (col + 1) / 2; }

// ------------------------------------------------------------------------------------------------ diagonal block
// A[k0 .. k0 + kb, k0 .. k0 + kb] -> U_kk, in place.  Thread t owns column c = t % CH_NB and the rows r = g, g + G, ... <= c
// (g = t / CH_NB of G = blockDim / CH_NB row groups).  Step j: every thread reads the pivot d = S[j][j] and the two row-j entries
// it needs, and updates its rows r > j: S[r][c] -= S[j][r] S[j][c] / d; group 0 then scales row j.  *info = k0 + j + 1 (first
// one wins) when a pivot is not positive (LAPACK's convention).
__global__ void __launch_bounds__(512)
cholDiagKernel(double* __restrict__ A, long long k0, int kb, long long* __restrict__ info)
{
    extern __shared__ double chS[];                      // [kb][CH_LD], S[c * CH_LD + r], r <= c
    const int tid = threadIdx.x;
    const int c = tid % CH_NB, g = tid / CH_NB, G = blockDim.x / CH_NB;
    if(*info != 0)
        return;                                          // an earlier block already failed
    for(int idx = tid; idx < kb * kb; idx += blockDim.x)
    {
        const int cc = idx / kb, r = idx - cc * kb;
        if(r <= cc)
            chS[cc * CH_LD + r] = A[chOff(k0 + cc) + k0 + r];
    }
    __syncthreads();
    for(int j = 0; j < kb; ++j)
    {
        const double d = chS[j * CH_LD + j];
        if(!(d > 0.0))
        {
            if(tid == 0)
                *info = k0 + j + 1;
            return;                                      // the same decision in every thread
        }
        const double inv = 1.0 / d;
        double ujc = 0.0;
        if(c > j && c < kb)
        {
            ujc = chS[c * CH_LD + j];
            const double s = ujc * inv;
            // rows r > j of this thread's group
            int r = j + 1 + ((g - (j + 1)) % G + G) % G;
            for(; r <= c; r += G)
                chS[c * CH_LD + r] -= chS[r * CH_LD + j] * s;
        }
        __syncthreads();
        // row j is final now and never read again by a later step, so no barrier is needed behind the scaling
        if(g == 0 && c >= j && c < kb)
            chS[c * CH_LD + j] = c == j ? sqrt(d) : ujc / sqrt(d);
    }
    __syncthreads();
    for(int idx = tid; idx < kb * kb; idx += blockDim.x)
    {
        const int cc = idx / kb, r = idx - cc * kb;
        if(r <= cc)
            A[chOff(k0 + cc) + k0 + r] = chS[cc * CH_LD + r];
    }
}

// ------------------------------------------------------------------------------------------------ panel
// Column j >= k1 = k0 + kb: x = U_kk^-T b with b = A[k0 .. k0 + kb, j] (one contiguous run of the packed column), in place.
// Forward substitution x_r = (b_r - sum_{s < r} U[s][r] x_s) / U[r][r]: U_kk is read through L1 (the same address in every
// thread, contiguous in s), x lives in shared memory, one column per thread.
__global__ void __launch_bounds__(CH_PANEL_COLS)
cholPanelKernel(double* __restrict__ A, long long k0, int kb, long long n, const long long* __restrict__ info)
{
    extern __shared__ double chX[];                      // [kb][CH_PANEL_COLS]
    if(*info != 0)
        return;
    const int tid = threadIdx.x;
    const long long j = k0 + kb + static_cast<long long>(blockIdx.x) * CH_PANEL_COLS + tid;
    if(j >= n)
        return;
    double* col = A + chOff(j) + k0;
    for(int r = 0; r < kb; ++r)
    {
        const double* u = A + chOff(k0 + r) + k0;        // U[0 .. r][r]
        double acc0 = col[r], acc1 = 0.0;
        int s = 0;
        for(; s + 1 < r; s += 2)
        {
            acc0 = fma(-__ldg(u + s), chX[s * CH_PANEL_COLS + tid], acc0);
            acc1 = fma(-__ldg(u + s + 1), chX[(s + 1) * CH_PANEL_COLS + tid], acc1);
        }
        if(s < r)
            acc0 = fma(-__ldg(u + s), chX[s * CH_PANEL_COLS + tid], acc0);
        chX[r * CH_PANEL_COLS + tid] = (acc0 + acc1) / __ldg(u + r);
    }
    for(int r = 0; r < kb; ++r)
        col[r] = chX[r * CH_PANEL_COLS + tid];
}

// ------------------------------------------------------------------------------------------------ trailing update
__device__ __forceinline__ void chDmma(double& c0, double& c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Tile (ti <= tj) of the trailing matrix, 128 x 128: C[i][j] -= sum_{r < kb} P[r][i] P[r][j], P[r][x] = A[chOff(x) + k0 + r] the
// panel just solved (x >= k1).  16 warps as 4 x 4: a warp owns 32 rows x 32 columns = 4 x 4 m8n8 accumulator tiles, which START
// as the C entries themselves (loaded while the first operand chunk is in flight); the A fragments are negated on the way in, so
// the tensor-core accumulation leaves C - P^T P and the epilogue is stores only.  The k-chunks of both operands are staged by
// cp.async as [column][k] with leading dimension 36, double buffered: an A fragment element (row i = lane / 4, k = lane % 4)
// and a B fragment element (k = lane % 4, column j = lane / 4) are the same access pattern, conflict-free.
constexpr int CH_SYRK_THREADS = 512;

__device__ __forceinline__ void chCpAsync8(double* dstShared, const double* src, bool live)
{
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(dstShared));
    const int bytes = live ? 8 : 0;                      // 0: nothing is read, the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(CH_SYRK_THREADS, 1)
cholSyrkKernel(double* __restrict__ A, long long k0, int kb, long long n, const long long* __restrict__ info)
{
    if(*info != 0)
        return;
    // tiles of the upper triangle in column-major order: t = tj (tj + 1) / 2 + ti, ti <= tj
    const long long t = blockIdx.x;
    int tj = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while(static_cast<long long>(tj + 1) * (tj + 2) / 2 <= t) ++tj;
    while(static_cast<long long>(tj) * (tj + 1) / 2 > t) --tj;
    const int ti = static_cast<int>(t - static_cast<long long>(tj) * (tj + 1) / 2);
    extern __shared__ double chSm[];                     // [2 stages][A, B][CH_TILE][CH_SLD]
    const long long k1 = k0 + kb;
    const long long i0 = k1 + static_cast<long long>(ti) * CH_TILE, j0 = k1 + static_cast<long long>(tj) * CH_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 2, wj = warp & 3;             // warp tile: rows wi * 32, columns wj * 32
    const bool diagTile = ti == tj;
    constexpr int STAGE = 2 * CH_TILE * CH_SLD;

    auto stage = [&](int buf, int kc)
    {
        double* sA = chSm + buf * STAGE;
        double* sB = sA + CH_TILE * CH_SLD;
        const bool live = kc + lane < kb;
        // a warp copies the 32 k-values of one panel column (256 contiguous bytes) at a time
        for(int x = warp; x < CH_TILE; x += CH_SYRK_THREADS / 32)
        {
            const long long ci = i0 + x, cj = j0 + x;
            const bool li = live && ci < n, lj = live && cj < n;
            chCpAsync8(sA + x * CH_SLD + lane, li ? A + chOff(ci) + k0 + kc + lane : A, li);
            if(!diagTile)
                chCpAsync8(sB + x * CH_SLD + lane, lj ? A + chOff(cj) + k0 + kc + lane : A, lj);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);

    // accumulators = the C entries: thread holds rows i = .. + lane / 4 and the column pair j = .. + 2 (lane % 4) + {0, 1}
    double acc[4][4][2];
#pragma unroll
    for(int b = 0; b < 4; ++b)
#pragma unroll
        for(int e = 0; e < 2; ++e)
        {
            const long long j = j0 + wj * 32 + b * 8 + 2 * (lane & 3) + e;
#pragma unroll
            for(int a = 0; a < 4; ++a)
            {
                const long long i = i0 + wi * 32 + a * 8 + (lane >> 2);
                acc[a][b][e] = (j < n && i <= j) ? A[chOff(j) + i] : 0.0;
            }
        }

    const int nChunks = (kb + CH_KC - 1) / CH_KC;
    for(int ch = 0; ch < nChunks; ++ch)
    {
        if(ch + 1 < nChunks)
        {
            stage((ch + 1) & 1, (ch + 1) * CH_KC);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double* sA = chSm + (ch & 1) * STAGE;
        const double* pB = diagTile ? sA : sA + CH_TILE * CH_SLD;
#pragma unroll
        for(int k4 = 0; k4 < CH_KC; k4 += 4)
        {
            double fa[4], fb[4];
#pragma unroll
            for(int a = 0; a < 4; ++a)
                fa[a] = -sA[(wi * 32 + a * 8 + (lane >> 2)) * CH_SLD + k4 + (lane & 3)];
#pragma unroll
            for(int b = 0; b < 4; ++b)
                fb[b] = pB[(wj * 32 + b * 8 + (lane >> 2)) * CH_SLD + k4 + (lane & 3)];
#pragma unroll
            for(int a = 0; a < 4; ++a)
#pragma unroll
                for(int b = 0; b < 4; ++b)
                    chDmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
        __syncthreads();                                 // the buffer is free for the chunk after next
    }

#pragma unroll
    for(int b = 0; b < 4; ++b)
#pragma unroll
        for(int e = 0; e < 2; ++e)
        {
            const long long j = j0 + wj * 32 + b * 8 + 2 * (lane & 3) + e;
#pragma unroll
            for(int a = 0; a < 4; ++a)
            {
                const long long i = i0 + wi * 32 + a * 8 + (lane >> 2);
                if(j < n && i <= j)                      // upper triangle only (matters on diagonal tiles); i < n follows
                    A[chOff(j) + i] = acc[a][b][e];
            }
        }
}

// ------------------------------------------------------------------------------------------------ log det, solves
// log det A = 2 sum_i log U[i][i]: one block, double accumulation in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
cholLogDetKernel(const double* __restrict__ U, long long n, double* __restrict__ out)
{
    __shared__ double red[256];
    double s = 0.0;
    for(long long i = threadIdx.x; i < n; i += 256)
        s += log(U[chOff(i) + i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for(int w = 128; w > 0; w >>= 1)
    {
        if(threadIdx.x < w)
            red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if(threadIdx.x == 0)
        *out = 2.0 * red[0];
}

// forward substitution U_kk^T y = t on the kb rows of block k0, all right-hand sides of the call: thread (r, m) = row r of
// right-hand side m (blockDim = (CH_NB, rhs per pass)); T is n x nRhs column-major
__global__ void __launch_bounds__(1024)
cholSolveDiagKernel(const double* __restrict__ U, long long k0, int kb, long long n, double* __restrict__ T, int nRhs)
{
    __shared__ double y[8][CH_NB];
    const int r = threadIdx.x, m = threadIdx.y;
    const int rhs = blockIdx.x * blockDim.y + m;
    const bool live = r < kb && rhs < nRhs;
    double t = live ? T[static_cast<long long>(rhs) * n + k0 + r] : 0.0;
    for(int s = 0; s < kb; ++s)
    {
        if(r == s)
            y[m][s] = t / U[chOff(k0 + s) + k0 + s];
        __syncthreads();
        if(live && r > s)
            t = fma(-U[chOff(k0 + r) + k0 + s], y[m][s], t);
        __syncthreads();
    }
    if(live)
        T[static_cast<long long>(rhs) * n + k0 + r] = y[m][r];
}

// t_j -= sum_{r < kb} U[k0 + r][j] y_r for every row j behind the block: a warp per column j (its kb entries are contiguous)
__global__ void __launch_bounds__(256)
cholSolveUpdateKernel(const double* __restrict__ U, long long k0, int kb, long long n, double* __restrict__ T, int nRhs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long j = k0 + kb + static_cast<long long>(blockIdx.x) * 8 + warp;
    if(j >= n)
        return;
    const double* u = U + chOff(j) + k0;
    double uv[CH_NB / 32];
#pragma unroll
    for(int q = 0; q < CH_NB / 32; ++q)
        uv[q] = (q * 32 + lane < kb) ? u[q * 32 + lane] : 0.0;
    for(int rhs = 0; rhs < nRhs; ++rhs)
    {
        const double* y = T + static_cast<long long>(rhs) * n + k0;
        double s = 0.0;
#pragma unroll
        for(int q = 0; q < CH_NB / 32; ++q)
            if(q * 32 + lane < kb)
                s = fma(uv[q], y[q * 32 + lane], s);
#pragma unroll
        for(int w = 16; w > 0; w >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, w);
        if(lane == 0)
            T[static_cast<long long>(rhs) * n + j] -= s;
    }
}

// packed sum C + F + N (F, N may be null; element stride on C for a slab element), the input of the factorisation
__global__ void __launch_bounds__(256)
packedSumKernel(const double* __restrict__ C, long long cStride, const double* __restrict__ F, const double* __restrict__ N,
                long long count, double* __restrict__ out)
{
    for(long long e = blockIdx.x * 256LL + threadIdx.x; e < count; e += static_cast<long long>(gridDim.x) * 256)
    {
        double v = C[e * cStride];
        if(F) v += F[e];
        if(N) v += N[e];
        out[e] = v;
    }
}

} // namespace cmg
