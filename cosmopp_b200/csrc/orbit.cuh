// Symmetry orbits of the full-sky HEALPix pixelisation (cmg_tqu_orbit, cmg_tqu_orbit_sharded, cmg_legendre_series_orbit).
//
// The grid is invariant under the rotation R by pi/2 about the polar axis; in NESTED ordering R maps base face f to
// (f & ~3) | ((f + 1) & 3) and keeps the index inside the face.  The local frames (e_theta, e_phi) rotate with the
// pixel, so the whole 3x3 block of a pixel pair is invariant:  C[X R a, Y R b] = C[X a, Y b]  for X, Y in {T, Q, U}.
// The four Clenshaw sums and the frame rotation are therefore evaluated once per orbit of pixel pairs and the nine
// results stored at up to four places.  The work is cut into classes of base-face pairs (rows x columns of the upper
// triangle); a class holds the source face pair, whether only q_row <= q_col is needed, and the images:
//
//   faces (rows, columns)                      pairs computed          images
//   different rings of faces, column p = 0     whole face pair         4, all with row pixel < column pixel
//   the same face, p = 0                       q_row <= q_col          4
//   same ring, faces (0, 1)                    whole face pair         (0,1) (1,2) (2,3) and (3,0), which is stored
//                                                                      TRANSPOSED: its row image has the larger index
//   same ring, faces (0, 2)                    q_row <= q_col          (0,2) (1,3), and transposed (2,0) (3,1)
//
// = 18 of the 72 face-pair units of the triangle, i.e. a quarter of the recurrence work.  A transposed image needs
// six of its nine entries staged through shared memory instead of three (kernel template SWAPMASK).  Without transposed
// images (mode 1) the classes (0,1) x 3 images, (0,3) alone and the whole of (0,2) x 2 images cost 22.5 units (3.2x).
#pragma once

#include "kernels.cuh"

namespace cmg
{

constexpr int ORB_MAX_CLASSES = 24;
constexpr int ORB_MAX_IMAGES = 4;

struct OrbitClass
{
    int rowFace, colFace;                  // base faces of the source tile rows / columns
    int tri;                               // 1: only pairs with q_row <= q_col (index inside the face)
    int sameFace;                          // 1: q_row == q_col is ONE pixel (transposed partners need q_row < q_col)
    int nImg;
    int imgRowFace[ORB_MAX_IMAGES];        // image k of the rows / columns lives in these faces (image 0 = source)
    int imgColFace[ORB_MAX_IMAGES];
    int imgSwap[ORB_MAX_IMAGES];           // 1: the image of the rows has the LARGER pixel index
    int comboBase[ORB_MAX_IMAGES];         // outbox numbering of (this class, image k, staged kind 0): see OrbitShardDev
    int mirror;                            // 1 (mode 3): the class also stores its image under the meridian mirror, 4 rotations of it
    int mirRowFace[ORB_MAX_IMAGES];        // faces of mirror image k (the column faces are imgColFace[k]: the mirror fixes the column face)
};

struct OrbitPlan
{
    int n;
    int facePix;                           // nside^2
    int nComboA, nComboB;                  // (class, image, staged kind) combinations of the whole-face-pair / q_row <= q_col classes
    OrbitClass c[ORB_MAX_CLASSES];
};

__host__ __device__ inline int orbitRotateFace(int f, int k) { return (f & ~3) | ((f + k) & 3); }
// (ix, iy) -> (iy, ix) inside a face: the even and the odd bits of the NESTED in-face index change places
__host__ __device__ inline unsigned orbitSwapBits(unsigned q) { return ((q & 0x55555555u) << 1) | ((q & 0xAAAAAAAAu) >> 1); }

// mode 3 = mode 0 + the meridian mirror (single owner only).  The grid is also invariant under the reflection phi -> pi/2 - phi:
// polar face position p -> -p, equatorial p -> 1 - p, in-face (ix, iy) -> (iy, ix) = the even and odd bits of the NESTED in-face
// index swapped (orbitSwapBits); a reflection flips the sign of every entry with exactly one U index
// (checked in tests/test_orbit_plan.py).  Composed with the rotation there is, for every pair of
// rings, a reflection that FIXES the column face at position 0 and maps the row faces p -> p': north rows, equatorial columns
// p' = -p - 1; north rows, south columns p' = -p; equatorial rows, south columns p' = 1 - p.  Of the twelve cross-ring classes
// five are the mirror images of five others (N0|N3, N1|N2 against E0; E0|E1, E2|E3 against S0; N1|N3 against S0) and are not
// evaluated: 13 of 72 face-pair units instead of 18.  (N0, S0) and (N2, S0) map to themselves and are kept whole.  A mirror
// image's 64 row pixels are a permutation of an aligned 64-run and its 32 column pixels two 16-runs 32 apart: every warp
// store is still made of whole 128-byte segments.  The image's columns are in general ANOTHER rank's, which is why only the
// single owner uses it.
// mode 0: transposed images allowed (18 units); mode 1: none (22.5 units).  swapMask selects the classes to emit by
// their mask of transposed images (bit k = image k): 0 = none transposed, 8 = the (0,1) classes, 12 = the (0,2) classes,
// 16 = the classes that carry mirror images (mode 3; they have no transposed ones), -1 = all.  The mask is a template parameter of the kernel: with the flags read from the plan at run time ptxas no longer
// proves the series loop warp-uniform and drops its uniform-register coefficient operands (0.84 instead of 1.0 of the FP64
// issue rate in the loop).
inline void orbitBuildPlan(int64_t nside, int mode, int swapMask, OrbitPlan& plan)
{
    plan.n = 0;
    plan.facePix = static_cast<int>(nside * nside);
    plan.nComboA = plan.nComboB = 0;
    // the numbering of the (class, image, staged kind) combinations runs over ALL classes of the mode, whatever swapMask
    // selects: whole-face-pair classes first (0 .. nComboA), q_row <= q_col classes behind them (fixed up below)
    auto add = [&](int rowFace, int colFace, int tri, int sameFace, int nImg, const int* rot, const int* swap, int mirrorRowFace = -1)
    {
        int mask = mirrorRowFace >= 0 ? 16 : 0, base[ORB_MAX_IMAGES] = {0, 0, 0, 0};
        int& counter = tri ? plan.nComboB : plan.nComboA;
        for(int k = 0; k < nImg; ++k)
        {
            mask |= (swap[k] ? 1 : 0) << k;
            base[k] = counter;
            counter += swap[k] ? 6 : 3;
        }
        if(swapMask >= 0 && swapMask != mask)
            return;
        OrbitClass& c = plan.c[plan.n++];
        c.mirror = mirrorRowFace >= 0 ? 1 : 0;
        for(int k = 0; k < ORB_MAX_IMAGES; ++k)
            c.mirRowFace[k] = mirrorRowFace >= 0 ? orbitRotateFace(mirrorRowFace, k) : 0;
        for(int k = 0; k < ORB_MAX_IMAGES; ++k)
            c.comboBase[k] = base[k];
        c.rowFace = rowFace;
        c.colFace = colFace;
        c.tri = tri;
        c.sameFace = sameFace;
        c.nImg = nImg;
        for(int k = 0; k < ORB_MAX_IMAGES; ++k)
        {
            const int r = k < nImg ? rot[k] : 0;
            c.imgRowFace[k] = orbitRotateFace(rowFace, r);
            c.imgColFace[k] = orbitRotateFace(colFace, r);
            c.imgSwap[k] = k < nImg ? swap[k] : 0;
        }
    };
    const int rot4[4] = {0, 1, 2, 3}, none[4] = {0, 0, 0, 0};
    // different rings of faces: the column face is brought to position 0 of its ring
    for(int gRow = 0; gRow < 3; ++gRow)
        for(int gCol = gRow + 1; gCol < 3; ++gCol)
            for(int p = 0; p < 4; ++p)
            {
                if(mode != 3)
                {
                    add(4 * gRow + p, 4 * gCol, 0, 0, 4, rot4, none);
                    continue;
                }
                // the reflection that fixes the column face: partner position of the row face
                const int partner = gCol == 1 ? ((-p - 1) & 3) : (gRow == 1 ? ((1 - p) & 3) : ((-p) & 3));
                if(partner == p)
                    add(4 * gRow + p, 4 * gCol, 0, 0, 4, rot4, none);             // its own mirror image: kept whole
                else if(p < partner)
                    add(4 * gRow + p, 4 * gCol, 0, 0, 4, rot4, none, 4 * gRow + partner);
                // else: stored as the mirror image of class (partner)
            }
    for(int g = 0; g < 3; ++g)
    {
        add(4 * g, 4 * g, 1, 1, 4, rot4, none);                                  // the same face
        if(mode == 0 || mode == 3)
        {
            const int swap01[4] = {0, 0, 0, 1};                                  // (3,0) is stored as (0,3) transposed
            add(4 * g, 4 * g + 1, 0, 0, 4, rot4, swap01);
            const int swap02[4] = {0, 0, 1, 1};                                  // (2,0), (3,1)
            add(4 * g, 4 * g + 2, 1, 0, 4, rot4, swap02);
        }
        else
        {
            add(4 * g, 4 * g + 1, 0, 0, 3, rot4, none);
            add(4 * g, 4 * g + 3, 0, 0, 1, rot4, none);
            add(4 * g, 4 * g + 2, 0, 0, 2, rot4, none);
        }
    }
    for(int c = 0; c < plan.n; ++c)
        if(plan.c[c].tri)
            for(int k = 0; k < ORB_MAX_IMAGES; ++k)
                plan.c[c].comboBase[k] += plan.nComboA;
}

// Where a rank's entries live.  A rank owns the in-face column range [q0, q1) of ALL twelve base faces (an orbit-closed set
// of pixel columns): it evaluates the source pairs whose column pixel has q in that range and stores every image.
//   strip[s][f]   ADJUSTED base of the packed columns s N + f F + [q0, q1): entry (row, col) at strip[s][f] + col (col+1)/2 + row
//   outbox        entries whose packed column belongs to another rank (the row pixel a' of the pair has q outside [q0, q1)),
//                 compact and ordered by destination: the block for rank d starts at destOff[d] and holds, for every
//                 (class, image, staged kind) combination that can address d -- all nComboA + nComboB for d < rank, the nComboA
//                 of the whole-face-pair classes for d > rank (a q_row <= q_col class never pairs a column with a later row) --
//                 the 32 x 32 sub-tiles (row half-tile h of d's range, column tile ct of this rank's range), each stored
//                 row-major (32 consecutive column pixels of one row pixel = one 256-byte run of the destination column):
//                     outbox[destOff[d] + ((combo nct + ct) nh_d + (h - h0_d)) 1024 + (q_a' mod 32) 32 + (q_b' mod 32)]
//                 Every element of the outbox is written exactly once.  Staged kinds: 0 <Q T>, 1 <U T>, 2 <U Q>, and for a
//                 transposed image 3 <T T>, 4 <Q Q>, 5 <U U>.
// One rank (nRanks = 1): every strip[s][f] is the base of the whole packed triangle and no outbox is touched.
constexpr int ORB_MAX_RANKS = 16;
constexpr int ORB_SUB = 32;                // rows and columns of an outbox sub-tile

struct OrbitShardDev
{
    int q0, q1;
    int nRanks, rank;
    int boundH[ORB_MAX_RANKS + 1];         // range boundaries of all ranks in units of ORB_SUB rows
    long long destOff[ORB_MAX_RANKS];      // first element of the block for destination d
    double* strip[3][12];
    double* outbox;
};

// host + device: destination rank of row half-tile h, and the outbox offset of sub-tile (combo 0, ct, h) with the stride per combo
__host__ __device__ inline int orbitOwnerOfHalfTile(const OrbitShardDev& sh, int h)
{
    // branch-free: boundH is padded with INT_MAX behind the last rank (host)
    int d = 0;
#pragma unroll
    for(int r = 1; r < ORB_MAX_RANKS; ++r)
        d += h >= sh.boundH[r] ? 1 : 0;
    return d;
}

// shared memory of tquOrbitKernel: frames of rows and columns, staged entries, column pointers of every image
template <bool SWAP, bool ROWPTR = false, bool MIRROR = false>
constexpr int orbitSmemDoubles()
{
    constexpr int IMAGES = MIRROR ? 2 * ORB_MAX_IMAGES : ORB_MAX_IMAGES;
    return 8 * PQ_TI + 8 * PQ_TJ + (SWAP ? 6 : 3) * PQ_TI * PQ_STAGE_LD + IMAGES * 3 * PQ_TJ + (ROWPTR ? IMAGES * 3 * PQ_TI : 0);
}

// staged kind t -> (column strip X, row strip Y) of the entry <X a', Y b'>
__device__ __forceinline__ int orbitStripX(int t) { return t == 0 ? 1 : (t == 1 || t == 2 || t == 5) ? 2 : (t == 4 ? 1 : 0); }
__device__ __forceinline__ int orbitStripY(int t) { return t == 2 ? 1 : (t == 4 ? 1 : (t == 5 ? 2 : 0)); }

// one row of a tile's store plan (tquOrbitKernel): destinations of (image, staged kind)
struct OrbitStoreRow
{
    double* own;                           // entry (row pixel of tile row 0, first column pixel of the tile) in the rank's own strips
    double* box[PQ_TI / ORB_SUB];          // first element of the outbox sub-tile of each 32-row half
    unsigned c;                            // packed column of tile row 0
    int minGap;                            // a row is stored where q_col - q_row >= minGap
};

// tile of a CTA: class blockIdx.y, 64 rows x 32 columns of in-face indices; false = nothing to do (q_row > q_col everywhere)
__device__ __forceinline__ bool orbitTile(const OrbitPlan& plan, const OrbitShardDev& sh, int& qRow0, int& qCol0)
{
    const int tilesPerFaceRows = plan.facePix / PQ_TI;
    qRow0 = static_cast<int>(blockIdx.x % tilesPerFaceRows) * PQ_TI;
    qCol0 = sh.q0 + static_cast<int>(blockIdx.x / tilesPerFaceRows) * PQ_TJ;
    return !(plan.c[blockIdx.y].tri && qRow0 > qCol0 + PQ_TJ - 1);
}

// One CTA = one 64 x 32 tile of pixel pairs of a class (blockIdx.y), evaluated exactly as tquKernel does, stored at
// every image.  Entries whose contiguous direction is the column pixel go through the shared-memory stage: the three
// transposed partners <Q_a T_b>, <U_a T_b>, <U_a Q_b> for every image, and <T T>, <Q Q>, <U U> as well for a
// transposed image (there the row image a' has the larger index, so (X a', X b') is stored in column X a').
// SWAPMASK: bit k set = image k of every class of this launch is a transposed one (orbitBuildPlan).
// ROWPTR (SWAPMASK == 0 only; NOT YET RUN ON A GPU, selected by mode 2 of the API): the destination of every (image, kind, row)
// of the store phase is computed once per tile into shared memory by all threads in parallel, as tquKernel does, instead of
// ~25 instructions per warp store in the store loop.
// MIRROR (mode 3, single owner, classes of mask 16: whole face pairs of different rings, no transposed images; needs ROWPTR):
// images 4 .. 7 are the four rotations of the pair's mirror image -- row pixel (mirRowFace[k], swapped bits of q_row), column pixel
// (imgColFace[k], swapped bits of q_col), the entries with exactly one U index negated.  The thread's row offset and the lane's
// offset inside a staged row become orbitSwapBits(il) and orbitSwapBits(lane): a warp store is two 128-byte segments.
// store policy of tquOrbitKernel: streaming (evict-first) by default; -DCMG_ORBIT_STORE_WB builds it with plain write-back
// stores for the comparison
#ifdef CMG_ORBIT_STORE_WB
#define ORB_ST(p, v) (*(p) = (v))
#else
#define ORB_ST(p, v) __stcs((p), (v))
#endif

template <int R, int MINB, int SWAPMASK, bool ROWPTR = false, bool MIRROR = false>
__global__ void __launch_bounds__(PQ_THREADS, MINB)
tquOrbitKernel(const __grid_constant__ TquStaticTable T, Geometry geo, int entrySlot,
               const __grid_constant__ OrbitPlan plan, const __grid_constant__ OrbitShardDev sh)
{
    extern __shared__ double4 orbSmem[];
    constexpr bool SWAP = SWAPMASK != 0;
    constexpr int SLOTS = SWAP ? 6 : 3;
    constexpr int IMAGES = MIRROR ? 2 * ORB_MAX_IMAGES : ORB_MAX_IMAGES;
    double* sI = reinterpret_cast<double*>(orbSmem);                  // [8][PQ_TI]
    double* sJ = sI + 8 * PQ_TI;                                     // [8][PQ_TJ]
    double* stage = sJ + 8 * PQ_TJ;                                  // [SLOTS][PQ_TI][PQ_STAGE_LD]
    double** sColPtr = reinterpret_cast<double**>(stage + SLOTS * PQ_TI * PQ_STAGE_LD);   // [image][3][PQ_TJ]
    double** sRowPtr = sColPtr + IMAGES * 3 * PQ_TJ;                                      // [image][3][PQ_TI]  (ROWPTR only)
    static_assert(!ROWPTR || SWAPMASK == 0, "the row-pointer table is sized for three staged kinds");
    static_assert(!MIRROR || ROWPTR, "mirror images are stored through the row-pointer table");

    int qRow0, qCol0;
    if(!orbitTile(plan, sh, qRow0, qCol0))
        return;
    const OrbitClass& oc = plan.c[blockIdx.y];
    const int facePix = plan.facePix;
    const int tri = oc.tri;

    const long long npix = geo.npix;
    const long long rowBlock = static_cast<long long>(oc.rowFace) * facePix + qRow0;
    const long long c0 = static_cast<long long>(oc.colFace) * facePix + qCol0;
    const int nImg = oc.nImg;
    const int tid = threadIdx.x;

    // Store plan of the tile: for every (image k, staged kind t) where its rows start in the rank's own strips and in its
    // outbox (OrbitShardDev; one base per 32-row half of the tile), computed by 24 threads into shared memory.  The store
    // phase reads nothing but this table: no kernel parameter is referenced behind the series loop, so ptxas has nothing to
    // hoist above it -- uniform values kept live across the loop cost it its uniform-register coefficient operands
    // (tests/test_sass_guard.py; a third of the DFMAs then read three vector registers).
    __shared__ OrbitStoreRow sStore[IMAGES * 6];
    __shared__ int sRange[2];
    if(tid < IMAGES * 6)
    {
        const int k = tid / 6, t = tid - 6 * k;
        const bool mirrored = MIRROR && k >= ORB_MAX_IMAGES;
        const bool swapped = !mirrored && ((SWAPMASK >> k) & 1);
        if((mirrored || k < nImg) && t < (swapped ? 6 : 3))
        {
            const int rowFace = mirrored ? oc.mirRowFace[k - ORB_MAX_IMAGES] : oc.imgRowFace[k];
            const int colFace = oc.imgColFace[k & (ORB_MAX_IMAGES - 1)];
            const long long rowPix0 = static_cast<long long>(rowFace) * facePix + (mirrored ? static_cast<int>(orbitSwapBits(qRow0)) : qRow0);
            const long long colPix0 = static_cast<long long>(colFace) * facePix + (mirrored ? static_cast<int>(orbitSwapBits(qCol0)) : qCol0);
            const long long c = orbitStripX(t) * npix + rowPix0;                 // < 2^32 for every valid nside
            OrbitStoreRow e;
            e.own = sh.strip[orbitStripX(t)][rowFace] + packedOffset(c) + (orbitStripY(t) * npix + colPix0);
            e.c = static_cast<unsigned>(c);
            // q_col - q_row must be >= minGap: none for whole face pairs; 1 where q_row == q_col is one pixel (its partners
            // are the direct entries) or belongs to image 0 (transposed images of a q_row <= q_col class); else 0
            e.minGap = !tri ? -(1 << 30) : ((oc.sameFace || swapped) ? 1 : 0);
#pragma unroll
            for(int half = 0; half < PQ_TI / ORB_SUB; ++half)
            {
                const int h = qRow0 / ORB_SUB + half;
                const int d = orbitOwnerOfHalfTile(sh, h);
                const int nhD = sh.boundH[d + 1] - sh.boundH[d];
                const int nct = (sh.q1 - sh.q0) / PQ_TJ, ct = (qCol0 - sh.q0) / PQ_TJ;
                const long long combo = oc.comboBase[k] + t;
                e.box[half] = sh.outbox + sh.destOff[d] + ((combo * nct + ct) * nhD + (h - sh.boundH[d])) * (ORB_SUB * ORB_SUB);
            }
            sStore[tid] = e;
        }
        if(tid == 0)
        {
            sRange[0] = sh.q0;
            sRange[1] = sh.q1;
        }
    }
    for(int idx = tid; idx < PQ_TI + PQ_TJ; idx += PQ_THREADS)
    {
        const bool isRow = idx < PQ_TI;
        const int loc = isRow ? idx : idx - PQ_TI;
        const long long pix = isRow ? rowBlock + loc : c0 + loc;
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
    // row 0 of columns b', N + b', 2N + b' for the column pixels b' of every image (always this rank's own strips)
    for(int idx = tid; idx < IMAGES * 3 * PQ_TJ; idx += PQ_THREADS)
    {
        const int k = idx / (3 * PQ_TJ);
        const int rem = idx - k * 3 * PQ_TJ;
        const int strip = rem / PQ_TJ;
        const int face = oc.imgColFace[k & (ORB_MAX_IMAGES - 1)];
        const int qc = qCol0 + (rem - strip * PQ_TJ);
        const long long col = strip * npix + static_cast<long long>(face) * facePix + (MIRROR && k >= ORB_MAX_IMAGES ? static_cast<int>(orbitSwapBits(qc)) : qc);
        sColPtr[idx] = sh.strip[strip][face] + packedOffset(col);
    }
    if(ROWPTR)
    {
        // where row (image k, kind t, a') of the store phase starts: the rank's own strip, or its outbox sub-tile
        __syncthreads();
        for(int idx = tid; idx < IMAGES * 3 * PQ_TI; idx += PQ_THREADS)
        {
            const int k = idx / (3 * PQ_TI);
            const int rem = idx - k * 3 * PQ_TI;
            const int t = rem / PQ_TI;
            const unsigned ilr = static_cast<unsigned>(rem - t * PQ_TI);
            if(MIRROR && k >= ORB_MAX_IMAGES)
            {
                // tile row ilr is row orbitSwapBits(ilr) of the image's 64-run (single owner: always the rank's own strips)
                const OrbitStoreRow e = sStore[k * 6 + t];
                const unsigned rm = orbitSwapBits(ilr);
                sRowPtr[idx] = e.own + (static_cast<unsigned long long>(rm) * e.c + (rm * (rm + 1)) / 2);
            }
            else if(k < nImg)
            {
                const OrbitStoreRow e = sStore[k * 6 + t];
                const int qa = qRow0 + static_cast<int>(ilr);
                static_assert(PQ_TI == 2 * ORB_SUB, "two outbox halves per tile");
                sRowPtr[idx] = (qa >= sh.q0 && qa < sh.q1) ? e.own + (static_cast<unsigned long long>(ilr) * e.c + (ilr * (ilr + 1)) / 2)
                                                           : (ilr >= ORB_SUB ? e.box[1] : e.box[0]) + (ilr % ORB_SUB) * ORB_SUB;
            }
        }
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int il = lane + 32 * (warp & 1);
    const int colGroup = warp >> 1;                     // 8 columns each
    const bool warpLive = !tri || qRow0 + 32 * (warp & 1) <= qCol0 + PQ_TJ - 1;

    const double nix = sI[0 * PQ_TI + il], niy = sI[1 * PQ_TI + il], niz = sI[2 * PQ_TI + il];

    if(warpLive)
    {
        for(int pass = 0; pass < 8 / R; ++pass)
        {
            const int jl0 = colGroup * 8 + pass * R;
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
            tquClenshawStatic<R>(st, T, entrySlot);

            const double tix = sI[3 * PQ_TI + il], tiy = sI[4 * PQ_TI + il], tiz = sI[5 * PQ_TI + il];
            const double pix_ = sI[6 * PQ_TI + il], piy = sI[7 * PQ_TI + il];
#pragma unroll
            for(int r = 0; r < R; ++r)
            {
                const int jl = jl0 + r;
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

                const double vTT = st.tt1[r];
                const double vTQ = xt * fma(aj, aj, -bj * bj);                // T_a Q_b
                const double vQQ = aRe + bRe;
                const double vTU = xt * (2.0 * aj * bj);                      // T_a U_b
                const double vQU = bIm - aIm;                                 // Q_a U_b
                const double vUU = aRe - bRe;

                const int dq = (qCol0 + jl) - (qRow0 + il);                   // q_col - q_row
                const bool live = !tri || dq >= 0;
                const bool strict = !tri || dq > 0;       // the q_row == q_col pairs of a transposed image are image 0's own
#pragma unroll
                for(int k = 0; k < ORB_MAX_IMAGES; ++k)
                {
                    // two flat predicates per image (no nested branches: ptxas then still proves the series loop of the next
                    // pass warp-uniform and keeps its coefficients in uniform registers)
                    const bool sw = (SWAPMASK >> k) & 1;
                    const bool some = k < nImg && live && (!sw || strict);          // <T Q>, <T U>, <Q U>: every image
                    const bool all = some && !sw;                                   // <T T>, <Q Q>, <U U>: straight images only
                    const long long ip = static_cast<long long>(oc.imgRowFace[k]) * facePix + (qRow0 + il);
                    double* colT = sColPtr[(k * 3 + 0) * PQ_TJ + jl] + ip;
                    double* colQ = sColPtr[(k * 3 + 1) * PQ_TJ + jl] + ip;
                    double* colU = sColPtr[(k * 3 + 2) * PQ_TJ + jl] + ip;
                    if(some)
                    {
                        ORB_ST(colQ, vTQ);
                        ORB_ST(colU, vTU);
                        ORB_ST(colU + npix, vQU);
                    }
                    if(all)
                    {
                        ORB_ST(colT, vTT);
                        ORB_ST(colQ + npix, vQQ);
                        ORB_ST(colU + 2 * npix, vUU);
                    }
                }
                if(MIRROR)
                {
                    // the mirror image of the pair, four rotations of it: <T U> and <Q U> change sign (whole face pairs: every pair is live)
                    const long long ilM = static_cast<long long>(orbitSwapBits(static_cast<unsigned>(qRow0 + il)));
#pragma unroll
                    for(int k = 0; k < ORB_MAX_IMAGES; ++k)
                    {
                        const long long ip = static_cast<long long>(oc.mirRowFace[k]) * facePix + ilM;
                        double* colT = sColPtr[((ORB_MAX_IMAGES + k) * 3 + 0) * PQ_TJ + jl] + ip;
                        double* colQ = sColPtr[((ORB_MAX_IMAGES + k) * 3 + 1) * PQ_TJ + jl] + ip;
                        double* colU = sColPtr[((ORB_MAX_IMAGES + k) * 3 + 2) * PQ_TJ + jl] + ip;
                        ORB_ST(colQ, vTQ);
                        ORB_ST(colU, -vTU);
                        ORB_ST(colU + npix, -vQU);
                        ORB_ST(colT, vTT);
                        ORB_ST(colQ + npix, vQQ);
                        ORB_ST(colU + 2 * npix, vUU);
                    }
                }
                stage[(0 * PQ_TI + il) * PQ_STAGE_LD + jl] = xt * fma(ai, ai, -bi * bi);   // Q_a T_b
                stage[(1 * PQ_TI + il) * PQ_STAGE_LD + jl] = xt * (2.0 * ai * bi);         // U_a T_b
                stage[(2 * PQ_TI + il) * PQ_STAGE_LD + jl] = aIm + bIm;                    // U_a Q_b
                if(SWAP)
                {
                    stage[(3 * PQ_TI + il) * PQ_STAGE_LD + jl] = vTT;
                    stage[(4 * PQ_TI + il) * PQ_STAGE_LD + jl] = vQQ;
                    stage[(5 * PQ_TI + il) * PQ_STAGE_LD + jl] = vUU;
                }
            }
        }
    }
    __syncthreads();

    // entries whose contiguous direction is the column pixel: for fixed row pixel a' the 32 column pixels of the tile
    // are consecutive rows of column (X a'); one warp store per (image, entry kind, a').  Column (X a') is this rank's
    // when q_a lies in its range, else the run goes to the outbox block of (kind, face of b').
    // Addresses: row a' = a'_0 + r of column strip X sits at packedOffset(c + r) = packedOffset(c) + r c + r (r + 1) / 2 with
    // c = X N + a'_0 warp-uniform, so a row costs one 32 x 32 -> 64 bit multiply-add on top of a base computed once per
    // (image, kind); the outbox row is r (q1 - q0) further on.
    const int qColLane = qCol0 + lane;
    if(ROWPTR)
    {
        for(int k = 0; k < nImg; ++k)
        {
            const int minGap = sStore[k * 6].minGap;
#pragma unroll 4
            for(int row = warp; row < 3 * PQ_TI; row += PQ_THREADS / 32)
            {
                const int ilr = row & (PQ_TI - 1);
                if(qColLane - (qRow0 + ilr) >= minGap)
                    ORB_ST(sRowPtr[k * 3 * PQ_TI + row] + lane, stage[row * PQ_STAGE_LD + lane]);
            }
        }
        if(MIRROR)
        {
            // <Q_a T_b>, <U_a T_b>, <U_a Q_b> of the mirror images: the 32 column pixels of the tile are two 16-runs of rows of
            // column (X a'); kinds 1 and 2 carry exactly one U index
            const unsigned laneM = orbitSwapBits(static_cast<unsigned>(lane));
            for(int k = ORB_MAX_IMAGES; k < 2 * ORB_MAX_IMAGES; ++k)
            {
#pragma unroll 4
                for(int row = warp; row < 3 * PQ_TI; row += PQ_THREADS / 32)
                {
                    const double v = stage[row * PQ_STAGE_LD + lane];
                    ORB_ST(sRowPtr[k * 3 * PQ_TI + row] + laneM, row >= PQ_TI ? -v : v);
                }
            }
        }
        return;
    }
    const int ownQ0 = sRange[0], ownQ1 = sRange[1];
    for(int k = 0; k < nImg; ++k)
    {
        const int nKinds = ((SWAPMASK >> k) & 1) ? 6 : 3;
        for(int t = 0; t < nKinds; ++t)
        {
            const OrbitStoreRow e = sStore[k * 6 + t];
            const double* src = stage + (t * PQ_TI) * PQ_STAGE_LD + lane;
#pragma unroll
            for(int u = 0; u < PQ_TI / (PQ_THREADS / 32); ++u)
            {
                const unsigned r = static_cast<unsigned>(warp + u * (PQ_THREADS / 32));
                const int qa = qRow0 + static_cast<int>(r);
                constexpr int HALF_U = ORB_SUB / (PQ_THREADS / 32);                            // u < HALF_U: rows of the first half
                const bool local = qa >= ownQ0 && qa < ownQ1;                                  // the same for the whole warp
                const unsigned long long off = local ? static_cast<unsigned long long>(r) * e.c + (r * (r + 1)) / 2
                                                     : static_cast<unsigned long long>(r % ORB_SUB) * ORB_SUB;
                double* dst = (local ? e.own : e.box[u / HALF_U]) + off + lane;
                if(qColLane - qa >= e.minGap)
                    ORB_ST(dst, src[r * PQ_STAGE_LD]);
            }
        }
    }
}

#undef ORB_ST

// ------------------------------------------------------------------------------------------------
// TT over symmetry orbits (cmg_legendre_series_orbit, cmg_legendre_series_orbit_sharded; the whole-call TT entry points take it
// on the full sky).  legendreSeriesKernel's tile (128 rows x 16 columns, a thread owns a row and walks the columns R at a time)
// addressed by (class, tile) of the plan.
//   One owner: the plan WITH transposed images (mode 0: 18 of 72 face-pair units, a quarter of the work).  A straight image is
//   one direct store per column (lanes along the row: 256-byte runs); a transposed image puts entry (a', b') into column a' --
//   the thread's R consecutive column pixels are R consecutive rows there: one 64-byte run per thread.  Uncoalesced across the
//   warp, but a store per 2 (lmax - 1) FMAs of the series: it does not show behind the FP64 pipe.
//   Several ranks: the plan WITHOUT transposed images (mode 1, 22.5 units) -- every image then has its row pixel before its
//   column pixel, every entry lands in a column of the rank that computed it, and there is nothing to exchange.
// SWAPMASK as in tquOrbitKernel (one launch per mask: the flags must be compile-time constants for ptxas to keep the series
// coefficients in uniform registers).  strip[f] = ADJUSTED base of the packed columns f F + [q0, q1).
// ------------------------------------------------------------------------------------------------
struct OrbitTtShardDev
{
    int q0, q1;
    double* strip[12];
};

template <int R, int MINB, int SWAPMASK>
__global__ void __launch_bounds__(TT_ROWS, MINB)
legendreSeriesOrbitKernel(const __grid_constant__ TtStaticTable T, Geometry geo, int entrySlot,
                          const __grid_constant__ OrbitPlan plan, const __grid_constant__ OrbitTtShardDev sh)
{
    const OrbitClass& oc = plan.c[blockIdx.y];
    const int facePix = plan.facePix;
    const int tilesPerFaceRows = facePix / TT_ROWS;
    const int qRow0 = static_cast<int>(blockIdx.x % tilesPerFaceRows) * TT_ROWS;
    const int qCol0 = sh.q0 + static_cast<int>(blockIdx.x / tilesPerFaceRows) * TT_COLS;
    const int tri = oc.tri;
    if(tri && qRow0 > qCol0 + TT_COLS - 1)
        return;                              // q_row > q_col everywhere
    if(tri && qRow0 + static_cast<int>(threadIdx.x & ~31u) > qCol0 + TT_COLS - 1)
        return;                              // this warp's 32 rows are all beyond the last column

    const int qRow = qRow0 + static_cast<int>(threadIdx.x);
    const long long i = static_cast<long long>(oc.rowFace) * facePix + qRow;
    const double xi = geo.nx[i], yi = geo.ny[i], zi = geo.nz[i];
    const int nImg = oc.nImg;

    for(int c = 0; c < TT_COLS; c += R)
    {
        double x2[R], b1[R], b2[R];
#pragma unroll
        for(int r = 0; r < R; ++r)
        {
            const long long j = static_cast<long long>(oc.colFace) * facePix + qCol0 + c + r;
            double dot = __dadd_rn(__dadd_rn(__dmul_rn(xi, __ldg(geo.nx + j)), __dmul_rn(yi, __ldg(geo.ny + j))),
                                   __dmul_rn(zi, __ldg(geo.nz + j)));
            dot = fmin(1.0, fmax(-1.0, dot));
            x2[r] = dot + dot;
            b1[r] = 0.0;
            b2[r] = 0.0;
        }
        ttClenshawStatic<R>(x2, b1, b2, T, entrySlot);
#pragma unroll
        for(int k = 0; k < ORB_MAX_IMAGES; ++k)
        {
            if(k < nImg)
            {
                const bool sw = (SWAPMASK >> k) & 1;
                const long long ip = static_cast<long long>(oc.imgRowFace[k]) * facePix + qRow;
                const long long jp0 = static_cast<long long>(oc.imgColFace[k]) * facePix + qCol0 + c;
                if(!sw)
                {
                    double* colPtr = sh.strip[oc.imgColFace[k]] + packedOffset(jp0) + ip;
#pragma unroll
                    for(int r = 0; r < R; ++r)
                    {
                        if(!tri || qRow <= qCol0 + c + r)
                            __stcs(colPtr, b1[r]);
                        colPtr += jp0 + r + 1;                     // next column starts (column index + 1) entries further
                    }
                }
                else
                {
                    // transposed image: its row pixel a' = ip has the larger index; the q_row == q_col pairs of a q_row <= q_col
                    // class are image 0's own
                    double* rowPtr = sh.strip[oc.imgRowFace[k]] + packedOffset(ip) + jp0;
#pragma unroll
                    for(int r = 0; r < R; ++r)
                        if(!tri || qRow < qCol0 + c + r)
                            __stcs(rowPtr + r, b1[r]);
                }
            }
        }
    }
}

// One destination block of a rank's outbox (OrbitShardDev: block(sender -> d)) -> its places in the packed columns of rank d.
// `recv` holds the receiver's ADJUSTED strip bases (its own strips, or the base of a whole packed triangle 36 times when the
// unsharded matrix is assembled on one GPU); `block` may be local memory (after an all-to-all) or the sender's outbox mapped
// through CUDA IPC, in which case the loads of this kernel are the NVLink transfer.
//   grid.x = (column tile ct of the sender's range) x (row half-tile hh of the receiver's range), grid.y = class of the full plan
struct OrbitInboxArgs
{
    long long npix;
    int senderQ0, nct;                     // the sender's range starts at senderQ0 and has nct column tiles
    int h0, nh;                            // the receiver's range in half-tiles
    int triLive;                           // receiver rank < sender rank: the q_row <= q_col classes address it as well
    const double* block;
    double* strip[3][12];
};

__global__ void __launch_bounds__(PQ_THREADS)
orbitInboxScatterKernel(const __grid_constant__ OrbitPlan plan, const __grid_constant__ OrbitInboxArgs a)
{
    const OrbitClass& oc = plan.c[blockIdx.y];
    if(oc.tri && !a.triLive)
        return;
    const int facePix = plan.facePix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ct = static_cast<int>(blockIdx.x) / a.nh, hh = static_cast<int>(blockIdx.x) - ct * a.nh;
    const long long subTiles = static_cast<long long>(a.nct) * a.nh;
    const int qRow0 = (a.h0 + hh) * ORB_SUB, qCol0 = a.senderQ0 + ct * PQ_TJ;
    for(int k = 0; k < oc.nImg; ++k)
    {
        const int nKinds = oc.imgSwap[k] ? 6 : 3;
        const long long rowPix0 = static_cast<long long>(oc.imgRowFace[k]) * facePix + qRow0;
        const long long colPix0 = static_cast<long long>(oc.imgColFace[k]) * facePix + qCol0;
        for(int t = 0; t < nKinds; ++t)
        {
            const double* src = a.block + ((oc.comboBase[k] + t) * subTiles + blockIdx.x) * (ORB_SUB * ORB_SUB) + lane;
            double* const dstBase = a.strip[orbitStripX(t)][oc.imgRowFace[k]] + (orbitStripY(t) * a.npix + colPix0 + lane);
            const long long c = orbitStripX(t) * a.npix + rowPix0;
            double v[ORB_SUB / (PQ_THREADS / 32)];
#pragma unroll
            for(int u = 0; u < ORB_SUB / (PQ_THREADS / 32); ++u)
                v[u] = __ldcs(src + (warp + u * (PQ_THREADS / 32)) * ORB_SUB);
#pragma unroll
            for(int u = 0; u < ORB_SUB / (PQ_THREADS / 32); ++u)
                __stcs(dstBase + packedOffset(c + warp + u * (PQ_THREADS / 32)), v[u]);
        }
    }
}

} // namespace cmg
