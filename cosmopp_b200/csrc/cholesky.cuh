// Cholesky factorisation A = U^T U of a symmetric positive definite matrix held as the PACKED upper triangle (entry (i <= j) at
// j (j + 1) / 2 + i), in place on the device -- the storage CMatrix has and the factorisation the reference's consumer runs on
// the host (LAPACK dpptrf 'U' on Math::SymmetricMatrix, reference source/matrix_impl.cpp:236-263, include/matrix_impl.hpp:495-502;
// called from Likelihood::construct, source/likelihood.cpp:100-133).  Nothing is unpacked: the 147456-dimensional [T;Q;U] matrix
// of Nside = 64 (87 GB packed, 174 GB square) is factorised where the generator left it.
//
// Right-looking, block size CH_NB = 128.  Step k (rows k0 .. k0 + kb of U):
//   cholDiagKernel    one CTA: the kb x kb diagonal block in shared memory, U_kk (sub-blocks of 16 rows)
//   cholPanelKernel   U[k0.., j] = U_kk^-T A[k0.., j] for every column j behind the block: a thread per column, forward
//                     substitution in blocks of 16 rows; in packed storage the kb rows of a column are one contiguous run
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
constexpr int CH_SLD = CH_KC;              // its leading dimension: no padding, the 16-byte chunks of a row are XOR-swizzled (chSwz)
constexpr int CH_PANEL_COLS = 64;          // panel solve: columns per CTA

__host__ __device__ __forceinline__ long long chOff(long long col) { return col * (col + 1) / 2; }

// The columns a launch works on: up to CH_MAX_RUNS runs of whole packed columns [colBegin, colEnd), each somewhere in this GPU's
// memory -- base[r] + chOff(j) + i is entry (i, j) of a column j of run r.  One GPU holding the whole triangle: one run, base = A.
// A rank of a sharded factorisation: its 36 orbit-closed runs (cmg_orbit_shard), already clipped to the columns behind the
// current block.  first[r] = CTAs of the launch in front of run r (the kernels look their run up with it), tile0[r] = tiles in
// front of the run's first column block in the numbering of chSyrkTilesBefore.
constexpr int CH_MAX_RUNS = 36;
struct CholRuns
{
    int count;
    long long colBegin[CH_MAX_RUNS], colEnd[CH_MAX_RUNS];
    double* base[CH_MAX_RUNS];
    long long first[CH_MAX_RUNS + 1];
    long long tile0[CH_MAX_RUNS];
};

__device__ __forceinline__ int chRunOf(const CholRuns& runs, long long cta)
{
    int r = 0;
    while(r + 1 < runs.count && cta >= runs.first[r + 1]) ++r;
    return r;
}

// ------------------------------------------------------------------------------------------------ diagonal block
__device__ __forceinline__ void chDmma(double& c0, double& c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void chCpAsync8(double* dstShared, const double* src)
{
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(dstShared));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void chCpAsync16(double* dstShared, const double* src)
{
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(dstShared));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// A[k0 .. k0 + kb, k0 .. k0 + kb] -> U_kk, in place: one CTA, the block in shared memory (S(r, c) at chS[c * CH_LD + r], r <= c;
// fetched with cp.async: every load of the block in flight at once), factorised in sub-blocks of CH_DB = 16 rows:
//   (a) 256 threads hold one element of the 16 x 16 diagonal sub-block each.  Pivot j: the threads of row j publish the row as it
//       stands (double-buffered, ONE 256-thread barrier per pivot); everyone updates its element with S(j, r) S(j, c) / S(j, j);
//       the row's own threads turn it into U(j, .) off the critical path (reciprocal square root + a Newton step: a double sqrt
//       followed by a division costs ~250 cycles).
//   (b) a thread per column behind it solves its 16 rows against that triangle (registers), multiplying by the reciprocal pivots;
//       the result also goes to chP[column][16], the operand layout of (c).
//   (c) the rank-16 update of the rest on the FP64 tensor path: 8 x 8 tiles of the upper triangle dealt to the 16 warps, K = 16.
// U_kk also leaves packed (entry (r <= c) at ukkOut[c (c + 1) / 2 + r]) with the reciprocals of its diagonal behind it: the panel
// kernel -- and, sharded, every other rank -- takes it from there in one stream of 16-byte copies.
// *info = k0 + j + 1 when pivot j is not positive (LAPACK's convention; the first failing block wins).
// (History: one step per row with the row in shared memory 0.169 ms; two barriers, sqrt and division per pivot, rank-16 update
// with one shared-memory load per FMA: 0.108 ms; reciprocal square roots: 0.087 ms.)
constexpr int CH_DB = 16;
constexpr int CH_DP_LD = CH_DB + 4;        // chP: 16 k-values of a column, leading dimension 20 (conflict-free DMMA fragments)
constexpr int CH_DIAG_SMEM_DOUBLES = CH_NB * CH_LD + CH_NB * CH_DP_LD;

__global__ void __launch_bounds__(512)
cholDiagKernel(double* __restrict__ A, long long k0, int kb, long long* __restrict__ info, double* __restrict__ ukkOut)
{
    extern __shared__ double chS[];                      // [CH_NB][CH_LD], then chP [CH_NB][CH_DP_LD]
    double* chP = chS + CH_NB * CH_LD;
    __shared__ int failed;
    __shared__ double sRow[2][CH_DB], sInv[CH_NB];       // sInv[r] = 1 / U(r, r): the solves multiply
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if(*info != 0)
        return;                                          // an earlier block already failed
    if(tid == 0)
        failed = 0;
    for(int idx = tid; idx < kb * kb; idx += blockDim.x)
    {
        const int cc = idx / kb, r = idx - cc * kb;
        if(r <= cc)
            chCpAsync8(chS + cc * CH_LD + r, A + chOff(k0 + cc) + k0 + r);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for(int jb = 0; jb < kb; jb += CH_DB)
    {
        const int wb = min(CH_DB, kb - jb);
        if(tid < CH_DB * CH_DB)
        {
            const int r = tid >> 4, cc = tid & (CH_DB - 1);
            const bool mine = r <= cc && cc < wb;
            double v = mine ? chS[(jb + cc) * CH_LD + jb + r] : 0.0;
            for(int j = 0; j < wb; ++j)
            {
                double* row = sRow[j & 1];
                if(r == j && cc >= j)
                    row[cc] = v;                         // row j as it stands (not yet divided by the pivot's root)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const double d = row[j];
                if(!(d > 0.0))
                {
                    if(tid == 0)
                        failed = jb + j + 1;
                    break;                               // the same in all 256 threads
                }
                if(mine && r > j)
                    v = fma(-(row[r] * __drcp_rn(d)), row[cc], v);
                else if(mine && r == j)
                {
                    double inv = rsqrt(d);
                    double root = d * inv;
                    root = fma(fma(-root, root, d), 0.5 * inv, root);
                    inv = fma(fma(-root, inv, 1.0), inv, inv);
                    v = cc == j ? root : v * inv;
                    if(cc == j)
                        sInv[jb + j] = inv;
                }
            }
            if(mine)
                chS[(jb + cc) * CH_LD + jb + r] = v;
        }
        __syncthreads();
        if(failed)
        {
            if(tid == 0)
                *info = k0 + failed;
            return;
        }
        const int first = jb + wb;
        if(first >= kb)
            break;
        // (b) rows jb .. jb + wb of the columns behind the sub-block: x = U_bb^-T b   (wb == CH_DB here: a short sub-block is the last)
        if(tid < kb && tid >= first)
        {
            double x[CH_DB];
#pragma unroll
            for(int i = 0; i < CH_DB; ++i)
                x[i] = chS[tid * CH_LD + jb + i];
#pragma unroll
            for(int i = 0; i < CH_DB; ++i)
            {
                x[i] *= sInv[jb + i];                    // every later row is updated at once: the chain is 16 x (multiply + FMA) long
#pragma unroll
                for(int i2 = i + 1; i2 < CH_DB; ++i2)
                    x[i2] = fma(-chS[(jb + i2) * CH_LD + jb + i], x[i], x[i2]);
            }
#pragma unroll
            for(int i = 0; i < CH_DB; ++i)
            {
                chS[tid * CH_LD + jb + i] = x[i];
                chP[tid * CH_DP_LD + i] = x[i];
            }
        }
        __syncthreads();
        // (c) S(r, c) -= sum_k U(jb + k, r) U(jb + k, c) for first <= r <= c < kb: 8 x 8 tiles (tr <= tc) over the warps
        {
            const int nT = (kb - first + 7) >> 3;
            const int g = lane >> 2, t = lane & 3;
            for(int p = warp; p < nT * (nT + 1) / 2; p += 16)
            {
                int tc = static_cast<int>((sqrtf(8.0f * p + 1.0f) - 1.0f) * 0.5f);
                while((tc + 1) * (tc + 2) / 2 <= p) ++tc;
                while(tc * (tc + 1) / 2 > p) --tc;
                const int tr = p - tc * (tc + 1) / 2;
                const int rA = min(first + tr * 8 + g, CH_NB - 1), cB = min(first + tc * 8 + g, CH_NB - 1);   // fragment rows (clamped: masked at the store)
                const int r = first + tr * 8 + g, c = first + tc * 8 + 2 * t;
                const bool ok0 = r <= c && c < kb, ok1 = r <= c + 1 && c + 1 < kb;
                double c0 = ok0 ? chS[c * CH_LD + r] : 0.0, c1 = ok1 ? chS[(c + 1) * CH_LD + r] : 0.0;
#pragma unroll
                for(int kk = 0; kk < CH_DB; kk += 4)
                    chDmma(c0, c1, -chP[rA * CH_DP_LD + kk + t], chP[cB * CH_DP_LD + kk + t]);
                if(ok0) chS[c * CH_LD + r] = c0;
                if(ok1) chS[(c + 1) * CH_LD + r] = c1;
            }
        }
        __syncthreads();
    }
    for(int idx = tid; idx < kb * kb; idx += blockDim.x)
    {
        const int cc = idx / kb, r = idx - cc * kb;
        if(r <= cc)
        {
            const double v = chS[cc * CH_LD + r];
            A[chOff(k0 + cc) + k0 + r] = v;
            ukkOut[cc * (cc + 1) / 2 + r] = v;
        }
    }
    if(tid < kb)
        ukkOut[kb * (kb + 1) / 2 + tid] = sInv[tid];
}

// ------------------------------------------------------------------------------------------------ panel
// Columns j >= k1 = k0 + kb: X = U_kk^-T B with B = A[k0 .. k0 + kb, j] (one contiguous run of each packed column), in place, and
// also into the dense panel plane the trailing update reads.  A CTA takes 64 columns, a warp 16 of them -- and only its own: the
// warps never wait for each other after the load.  Forward substitution in blocks of 16 rows; for block rb
//   C = B[rb .. rb + 16] - U[0 .. rb, rb .. rb + 16]^T X[0 .. rb]      16 x rb x 16 per warp on the FP64 tensor path
//                                                                      (mma.sync.m8n8k4.f64: the accumulators start as B, the
//                                                                      U fragments enter negated)
//   X[rb .. rb + 16] = T^-T C, T the 16 x 16 triangle of U_kk           a lane per column, 16 values in registers, U at a
//                                                                      warp-uniform shared-memory address, reciprocal pivots
// X lives in shared memory column by column (leading dimension 132: the B fragment -- k = lane % 4 along a column, column
// = lane / 4 -- is then conflict-free, and so are the loads and stores of whole columns); U_kk packed (66 KB) and the
// reciprocals of its diagonal next to it.
// (History: one dependent chain per row, 0.4 - 0.8 ms per step; U through L1 and divisions, 0.23 ms, long-scoreboard stalls
// -- profiles/r2_chol_panel_v2_metrics.txt; a thread per column with 16 FMA chains fed from shared memory, 0.108 ms per wave of
// 128 columns per SM, bound by one shared-memory load per FMA.)
constexpr int CH_PB = 16;
constexpr int CH_PANEL_THREADS = 128;
constexpr int CH_PX_LD = CH_NB + 4;
constexpr int CH_PANEL_SMEM_DOUBLES = CH_PANEL_COLS * CH_PX_LD + CH_NB * (CH_NB + 1) / 2 + CH_NB;
constexpr int CH_UKK_DOUBLES = CH_NB * (CH_NB + 1) / 2 + CH_NB;      // packed U_kk + reciprocal pivots

__global__ void __launch_bounds__(CH_PANEL_THREADS)
cholPanelKernel(const __grid_constant__ CholRuns runs, long long k0, int kb, const long long* __restrict__ info,
                const double* __restrict__ ukk, double* __restrict__ panel, long long panelCol0)
{
    extern __shared__ double chX[];                      // [CH_PANEL_COLS][CH_PX_LD], then U_kk packed + 1 / diagonal (as cholDiagKernel left them)
    if(*info != 0)
        return;
    double* sU = chX + CH_PANEL_COLS * CH_PX_LD;         // U(s, r) at sU[r (r + 1) / 2 + s]
    const int nU = kb * (kb + 1) / 2;
    double* sInv = sU + nU;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = chRunOf(runs, blockIdx.x);
    double* const A = runs.base[run];
    // everything this CTA reads, in flight at once: U_kk and the reciprocal pivots (one contiguous piece, 16 bytes a copy) ...
    for(int idx = 2 * tid; idx < nU + kb; idx += 2 * CH_PANEL_THREADS)       // kb = CH_NB here: an even count
        chCpAsync16(sU + idx, ukk + idx);
    const long long j0 = runs.colBegin[run] + (static_cast<long long>(blockIdx.x) - runs.first[run]) * CH_PANEL_COLS;
    const long long colEnd = runs.colEnd[run];
    constexpr int WCOLS = CH_PANEL_COLS / (CH_PANEL_THREADS / 32);      // 16 columns per warp
    const int col0 = warp * WCOLS;
    // ... and the kb rows of the warp's own 16 columns (a packed column starts at any multiple of 8 bytes)
    // kb is a multiple of CH_PB here (CH_NB in fact): only the LAST block of a matrix can be short, and nothing lies behind it
#pragma unroll 4
    for(int c = 0; c < WCOLS; ++c)
    {
        const long long j = j0 + col0 + c;
        const double* src = A + chOff(j) + k0;
#pragma unroll
        for(int q = 0; q < CH_NB / 32; ++q)
        {
            if(j < colEnd && q * 32 + lane < kb)
                chCpAsync8(chX + (col0 + c) * CH_PX_LD + q * 32 + lane, src + q * 32 + lane);
            else
                chX[(col0 + c) * CH_PX_LD + q * 32 + lane] = 0.0;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                     // U_kk and the reciprocals are in place (the columns are the warp's own)

    const int g = lane >> 2, t = lane & 3;
    for(int rb = 0; rb < kb; rb += CH_PB)
    {
        if(rb > 0)
        {
            double acc[2][2][2];
#pragma unroll
            for(int m = 0; m < 2; ++m)
#pragma unroll
                for(int nn = 0; nn < 2; ++nn)
#pragma unroll
                    for(int e = 0; e < 2; ++e)
                        acc[m][nn][e] = chX[(col0 + nn * 8 + 2 * t + e) * CH_PX_LD + rb + m * 8 + g];
            const double* uA0 = sU + (rb + g) * (rb + g + 1) / 2 + t;                  // U(s + t, rb + g)
            const double* uA1 = sU + (rb + 8 + g) * (rb + 8 + g + 1) / 2 + t;          // U(s + t, rb + 8 + g)
            const double* xB0 = chX + (col0 + g) * CH_PX_LD + t;                       // X(s + t, col0 + g)
            const double* xB1 = xB0 + 8 * CH_PX_LD;
#pragma unroll 4
            for(int s0 = 0; s0 < rb; s0 += 4)
            {
                const double a0 = -uA0[s0], a1 = -uA1[s0], b0 = xB0[s0], b1 = xB1[s0];
                chDmma(acc[0][0][0], acc[0][0][1], a0, b0);
                chDmma(acc[0][1][0], acc[0][1][1], a0, b1);
                chDmma(acc[1][0][0], acc[1][0][1], a1, b0);
                chDmma(acc[1][1][0], acc[1][1][1], a1, b1);
            }
#pragma unroll
            for(int m = 0; m < 2; ++m)
#pragma unroll
                for(int nn = 0; nn < 2; ++nn)
#pragma unroll
                    for(int e = 0; e < 2; ++e)
                        chX[(col0 + nn * 8 + 2 * t + e) * CH_PX_LD + rb + m * 8 + g] = acc[m][nn][e];
            __syncwarp();
        }
        if(lane < WCOLS)
        {
            double* xc = chX + (col0 + lane) * CH_PX_LD + rb;
            double x[CH_PB];
#pragma unroll
            for(int i = 0; i < CH_PB; ++i)
                x[i] = xc[i];
#pragma unroll
            for(int i = 0; i < CH_PB; ++i)
            {
                x[i] *= sInv[rb + i];
#pragma unroll
                for(int i2 = i + 1; i2 < CH_PB; ++i2)
                    x[i2] = fma(-sU[(rb + i2) * (rb + i2 + 1) / 2 + rb + i], x[i], x[i2]);
            }
#pragma unroll
            for(int i = 0; i < CH_PB; ++i)
                xc[i] = x[i];
        }
        __syncwarp();
    }
#pragma unroll 2
    for(int c = 0; c < WCOLS; ++c)
    {
        const long long j = j0 + col0 + c;
        if(j >= colEnd)
            break;
        double* dst = A + chOff(j) + k0;
        double* p = panel ? panel + (j - panelCol0) * CH_NB : nullptr;
#pragma unroll
        for(int q = 0; q < CH_NB / 32; ++q)
            if(q * 32 + lane < kb)
            {
                const double v = chX[(col0 + c) * CH_PX_LD + q * 32 + lane];
                dst[q * 32 + lane] = v;
                if(p)
                    p[q * 32 + lane] = v;
            }
    }
}

// ------------------------------------------------------------------------------------------------ trailing update
// Tile of the trailing matrix, 128 rows (i) x 64 columns (j): C[i][j] -= sum_{r < kb} P[r][i] P[r][j], P = the kb rows of U just
// solved (kb = CH_NB, or a whole group of blocks: up to CH_MAX_GROUP * CH_NB).  8 warps as 4 x 2: a warp owns 32 rows x 32 columns
// = 4 x 4 m8n8 accumulator tiles, which START as the C entries themselves (loaded while the first operand chunk is in flight);
// the A fragments enter negated (DMMA's operand modifier), so the tensor-core accumulation leaves C - P^T P and the epilogue is
// stores only.  The k-chunks (32 rows of U) of both operands are staged by 16-byte cp.async as [column][k], double buffered.
// Two CTAs per SM (96 KB of shared memory, 128 registers each): one CTA's loads of C, its barriers and its stores hide behind
// the other's DMMAs -- the first version (one 128 x 128 CTA of 512 threads per SM) left the DMMA pipe idle 46 % of the time
// (profiles/r2_syrk_v1_metrics.txt: stalls math-pipe throttle AND wait / lg-throttle / barrier: bursts, then drains).
constexpr int CH_TJ = 64;                  // syrk: columns of a tile
constexpr int CH_SYRK_THREADS = 256;

// Staged row of operand column x: 32 doubles = 16 chunks of 16 bytes, no padding.  A thread (g = lane / 4, t = lane % 4) reads the
// k pair (2 t, 2 t + 1) of eight consecutive k as ONE 16-byte load and feeds two DMMAs with it (the first contracts the even k
// of the eight, the second the odd ones -- A and B fragments use the same assignment, so the products pair up correctly).  A
// 16-byte load is served per quarter warp (rows g = 2 q, 2 q + 1, four chunks each): chunk c of an odd row sits at c ^ 4, which
// puts the two rows into different halves of the 128-byte bank window.  Half the shared-memory instructions of the 8-byte
// version with leading dimension 36 (MIO throttle / short scoreboard: profiles/r2_syrk_v2_metrics.txt, then _v3_).
// tiles behind k1, column block b (64 wide) outermost: it meets the row tiles ti = 0 .. b / 2 (128 high); cumulative count
__host__ __device__ __forceinline__ long long chSyrkTilesBefore(long long b)
{
    const long long m = b >> 1, r = b & 1;
    return b + m * (m - 1) + r * m;
}

// Operands: the DENSE panel the panel kernel writes next to the packed columns -- plane s (the s-th block of CH_NB rows of a
// group of blocks, see cmg_packed_cholesky) holds U[k0 + s CH_NB + r][x] at panel[s planeStride + (x - panelCol0) CH_NB + r].
// Rows of the panel are 1 KB apart and 16-byte aligned whatever the column (a packed column starts at an odd or even double),
// so a k-chunk of a tile's operands is staged with 16-byte cp.async from plain strided addresses -- no offset table, half the
// copy instructions.  The sharded factorisation needs the dense form anyway (rows of OTHER ranks' columns are not in this
// rank's memory; every rank holds the panel after the all-reduce).  The C tiles are the launch's runs of columns: a run's
// tiles are the column blocks [(colBegin - k1) / 64, (colEnd - k1) / 64) of the numbering above; columns from colEnd on are
// not this run's (the tile is cut there like at the edge of the matrix).
// stripOnly: only the first row tile (rows k1 .. k1 + CH_TILE) of every column block -- the catch-up update of the next block
// of a group before it is factorised (tile t of a run = column block t + tile0).  shift: the launch updates the trailing matrix
// from row and column k0 + kb + shift on (the update of a group is issued in pieces: the rows the NEXT group factorises first,
// strip by strip, so that its factorisation can run beside the rest -- cmg_packed_cholesky).
// The panel planes are allocated with CH_TILE + CH_TJ rows to spare: operand rows behind the edge of the matrix are read
// (whatever they hold) and the entries computed from them never stored.
__global__ void __launch_bounds__(CH_SYRK_THREADS, 2)
cholSyrkKernel(const __grid_constant__ CholRuns runs, long long k0, int kb, long long shift, const long long* __restrict__ info,
               const double* __restrict__ panel, long long panelCol0, long long planeStride, int stripOnly)
{
    if(*info != 0)
        return;
    const int run = chRunOf(runs, blockIdx.x);
    double* const A = runs.base[run];
    const long long n = runs.colEnd[run];
    const long long t = (static_cast<long long>(blockIdx.x) - runs.first[run]) + runs.tile0[run];
    long long bj = t;
    int ti = 0;
    if(!stripOnly)
    {
        bj = static_cast<long long>(2.0 * sqrt(static_cast<double>(t)));
        while(chSyrkTilesBefore(bj + 1) <= t) ++bj;
        while(chSyrkTilesBefore(bj) > t) --bj;
        ti = static_cast<int>(t - chSyrkTilesBefore(bj));
    }
    extern __shared__ double chSm[];                     // [2 stages][CH_TILE + CH_TJ][CH_SLD]
    const long long k1 = k0 + kb + shift;                // first row AND column of the part of the trailing matrix this launch updates
    const long long i0 = k1 + static_cast<long long>(ti) * CH_TILE, j0 = k1 + bj * CH_TJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 1, wj = warp & 1;             // warp tile: rows wi * 32, columns wj * 32
    constexpr int STAGE = (CH_TILE + CH_TJ) * CH_SLD;

    // staging: a warp instruction copies the 32 k-values of TWO operand columns (2 x 256 bytes, 16 bytes per lane); warp w takes
    // the row pairs w, w + 8, ... of the stage (rows 0 .. 127 = the tile's rows i, 128 .. 191 = its columns j)
    const int sRow = 2 * warp + (lane >> 4), sChunk = lane & 15;
    const double* srcA = panel + (i0 - panelCol0 + sRow) * CH_NB + 2 * sChunk;
    const double* srcB = panel + (j0 - panelCol0 + sRow) * CH_NB + 2 * sChunk;
    const int dstOff = sRow * CH_SLD + ((sChunk ^ ((sRow & 1) << 2)) << 1);
    auto stage = [&](int buf, int kc)
    {
        double* sP = chSm + buf * STAGE + dstOff;
        const long long kOff = static_cast<long long>(kc / CH_NB) * planeStride + (kc % CH_NB);
        constexpr int ROWS_PER_PASS = 2 * (CH_SYRK_THREADS / 32);
#pragma unroll
        for(int m = 0; m < CH_TILE / ROWS_PER_PASS; ++m)
            chCpAsync16(sP + m * ROWS_PER_PASS * CH_SLD, srcA + kOff + m * ROWS_PER_PASS * CH_NB);
#pragma unroll
        for(int m = 0; m < CH_TJ / ROWS_PER_PASS; ++m)
            chCpAsync16(sP + (CH_TILE + m * ROWS_PER_PASS) * CH_SLD, srcB + kOff + m * ROWS_PER_PASS * CH_NB);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);

    // accumulators = the C entries: thread holds rows i = .. + lane / 4 and the column pair j = .. + 2 (lane % 4) + {0, 1}
    double acc[4][4][2];
    const long long iBase = i0 + wi * 32 + (lane >> 2);
#pragma unroll
    for(int b = 0; b < 4; ++b)
#pragma unroll
        for(int e = 0; e < 2; ++e)
        {
            const int jl = wj * 32 + b * 8 + 2 * (lane & 3) + e;
            const long long j = j0 + jl;
            const long long cOff = j < n ? chOff(j) : -1;
#pragma unroll
            for(int a = 0; a < 4; ++a)
            {
                const long long i = iBase + a * 8;
                acc[a][b][e] = (cOff >= 0 && i <= j) ? A[cOff + i] : 0.0;
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
        const double* sB = sA + CH_TILE * CH_SLD;
        const int g = lane >> 2, sw = (g & 1) << 2;      // (row & 1) == (g & 1): the warp's row offsets are even
#pragma unroll
        for(int k8 = 0; k8 < CH_KC; k8 += 8)
        {
            const int kOff = (((k8 >> 1) + (lane & 3)) ^ sw) << 1;
            double2 fa[4], fb[4];
#pragma unroll
            for(int a = 0; a < 4; ++a)
                fa[a] = *reinterpret_cast<const double2*>(sA + (wi * 32 + a * 8 + g) * CH_SLD + kOff);
#pragma unroll
            for(int b = 0; b < 4; ++b)
                fb[b] = *reinterpret_cast<const double2*>(sB + (wj * 32 + b * 8 + g) * CH_SLD + kOff);
#pragma unroll
            for(int a = 0; a < 4; ++a)
#pragma unroll
                for(int b = 0; b < 4; ++b)
                    chDmma(acc[a][b][0], acc[a][b][1], -fa[a].x, fb[b].x);
#pragma unroll
            for(int a = 0; a < 4; ++a)
#pragma unroll
                for(int b = 0; b < 4; ++b)
                    chDmma(acc[a][b][0], acc[a][b][1], -fa[a].y, fb[b].y);
        }
        __syncthreads();                                 // the buffer is free for the chunk after next
    }

#pragma unroll
    for(int b = 0; b < 4; ++b)
#pragma unroll
        for(int e = 0; e < 2; ++e)
        {
            const int jl = wj * 32 + b * 8 + 2 * (lane & 3) + e;
            const long long j = j0 + jl;
            const long long cOff = j < n ? chOff(j) : -1;
#pragma unroll
            for(int a = 0; a < 4; ++a)
            {
                const long long i = iBase + a * 8;
                if(cOff >= 0 && i <= j)                  // upper triangle only (matters on tiles the diagonal crosses); i < n follows
                    A[cOff + i] = acc[a][b][e];
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

// the same over the diagonal entries of a rank's runs of columns (sharded factor): the rank's share of log det
__global__ void __launch_bounds__(256)
cholLogDetRunsKernel(const __grid_constant__ CholRuns runs, double* __restrict__ out)
{
    __shared__ double red[256];
    double s = 0.0;
    for(int r = 0; r < runs.count; ++r)
        for(long long i = runs.colBegin[r] + threadIdx.x; i < runs.colEnd[r]; i += 256)
            s += log(runs.base[r][chOff(i) + i]);
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
cholSolveUpdateKernel(const __grid_constant__ CholRuns runs, long long k0, int kb, long long n, double* __restrict__ T, int nRhs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = chRunOf(runs, blockIdx.x);
    const long long j = runs.colBegin[run] + (static_cast<long long>(blockIdx.x) - runs.first[run]) * 8 + warp;
    if(j >= runs.colEnd[run])
        return;
    const double* u = runs.base[run] + chOff(j) + k0;
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

// upper triangle of a dense column-major n x n matrix -> packed
__global__ void __launch_bounds__(256)
packUpperKernel(const double* __restrict__ full, long long n, double* __restrict__ packed)
{
    const long long j = blockIdx.y;
    for(long long i = blockIdx.x * 256LL + threadIdx.x; i <= j; i += static_cast<long long>(gridDim.x) * 256)
        packed[chOff(j) + i] = full[j * n + i];
}

} // namespace cmg
