/* cmg.h -- C ABI of the B200-native pixel-covariance (C-matrix) generator.
 *
 * This is the drop-in boundary for ONE hot path of Cosmo++ (aslanyan/cosmopp):
 * CMatrixGenerator -> CMatrix, the pixel-space CMB signal covariance.  The reference has no
 * FFI layer for this path (it is plain C++: include/c_matrix.hpp, include/c_matrix_generator.hpp),
 * so the entry points below are what a binding of those two classes needs; each one names the
 * reference code it replaces.  The C++ classes in include/c_matrix.hpp and
 * include/c_matrix_generator.hpp of THIS repository are thin wrappers over this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; all sizes/indices are int64_t (the reference's 32-bit `int`
 *     indices overflow for nPix > 46340, reference source/c_matrix.cpp:24,36);
 *   - every function returns a cmg_status (0 = ok) unless documented otherwise and never throws;
 *     cmg_last_error() gives the text (the C++ wrappers turn it into StandardException, the
 *     reference's error convention, include/exception_handler.hpp:10-31);
 *   - "packed" = upper triangle, column-major: entry (i<=j) at j(j+1)/2+i
 *     (reference source/c_matrix.cpp:27-39; identical to LAPACK 'U' packed storage);
 *   - polarized matrices use rows/cols [T_0..T_{N-1}, Q_0..Q_{N-1}, U_0..U_{N-1}] (the reference's
 *     [Q;U] convention, source/c_matrix_generator.cpp:678-681, with T in front), Q,U in the local
 *     (e_theta, e_phi) basis, HEALPix-primer sign convention;
 *   - pointers named d_* are device pointers on the context's GPU (or peer-mapped memory of
 *     another GPU where stated); all others are host pointers;
 *   - a context is bound to one GPU and one stream; calls on different contexts are independent
 *     and may be made concurrently from several host threads (no hidden globals).  There is no
 *     CPU fallback: without a usable sm_100 device cmg_create fails.
 */
#ifndef CMG_H
#define CMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmg_ctx cmg_ctx;

typedef enum cmg_status {
    CMG_OK = 0,
    CMG_EINVAL = 1,        /* bad argument */
    CMG_ECUDA = 2,         /* CUDA runtime / launch failure (text in cmg_last_error) */
    CMG_ENOMEM = 3,        /* host or device allocation failed */
    CMG_ESTATE = 4,        /* call order: geometry not set, ... */
    CMG_EUNSUPPORTED = 5,  /* size beyond what the kernels are built for (lmax > CMG_LMAX_LIMIT) */
    CMG_ENUMERIC = 6       /* the matrix handed to the likelihood is not positive definite */
} cmg_status;

#define CMG_LMAX_LIMIT 1023   /* series length the kernels stage in shared memory */
#define CMG_MAX_PARTS 16      /* column-block owners in a sharded polarized layout */

/* ---------------------------------------------------------------- context ------------------- */

int cmg_version(void);
/* number of visible CUDA devices (0 when there is no driver/GPU) */
int cmg_device_count(void);
/* binds to `device`; fails with CMG_ECUDA when no sm_100 GPU is usable (no CPU fallback) */
cmg_status cmg_create(cmg_ctx** ctx, int device);
void cmg_destroy(cmg_ctx* ctx);
/* text of the last failure on this context (ctx == NULL: last failure of cmg_create on this thread) */
const char* cmg_last_error(const cmg_ctx* ctx);
/* launch on a caller-owned cudaStream_t (e.g. torch's current stream; NULL is the legacy default
 * stream).  A new context launches on a private non-blocking stream until this is called. */
cmg_status cmg_set_stream(cmg_ctx* ctx, void* cuda_stream);
/* go back to the context's private stream */
cmg_status cmg_use_own_stream(cmg_ctx* ctx);
cmg_status cmg_synchronize(cmg_ctx* ctx);
/* number of generator kernels this context has launched so far (bench.py's gpu_launches) */
int64_t cmg_launch_count(const cmg_ctx* ctx);

/* device memory helpers for hosts that do not bring their own allocator (the C++ drop-in) */
cmg_status cmg_device_malloc(cmg_ctx* ctx, int64_t bytes, void** d_ptr);
cmg_status cmg_device_free(cmg_ctx* ctx, void* d_ptr);
cmg_status cmg_host_malloc_pinned(int64_t bytes, void** ptr);
cmg_status cmg_host_free_pinned(void* ptr);
/* CUDA IPC export / import of a cmg_device_malloc'ed buffer, for ranks of one box that write the entries they
 * compute for another owner straight into that owner's strip over NVLink (cmg_tqu_layout kind 0 with peer
 * pointers).  handle receives CMG_IPC_HANDLE_BYTES bytes; cmg_ipc_open maps a peer rank's buffer into this process
 * (peer access is enabled on first use) and cmg_ipc_close unmaps it. */
#define CMG_IPC_HANDLE_BYTES 64
cmg_status cmg_ipc_export(cmg_ctx* ctx, void* d_ptr, void* handle);
cmg_status cmg_ipc_open(cmg_ctx* ctx, const void* handle, void** d_peer_ptr);
cmg_status cmg_ipc_close(cmg_ctx* ctx, void* d_peer_ptr);
/* stream-ordered copies on the context's stream (cmg_copy_to_host returns when the data is there) */
cmg_status cmg_copy_to_host(cmg_ctx* ctx, void* dst_host, const void* d_src, int64_t bytes);
cmg_status cmg_copy_to_device(cmg_ctx* ctx, void* d_dst, const void* src_host, int64_t bytes);
/* device to device; either side may be peer memory mapped with cmg_ipc_open (the copy then crosses NVLink) */
cmg_status cmg_copy_on_device(cmg_ctx* ctx, void* d_dst, const void* d_src, int64_t bytes);
/* bytes this context has moved over PCIe so far: cmg_copy_to_device / cmg_copy_to_host, the whole calls with host output,
 * cmg_matrix_to_host, cmg_orbit_strips_to_host (the few KB of series weights of a launch are not counted) */
cmg_status cmg_transfer_counters(const cmg_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* ------------------------------------------- host pieces of the path (pure CPU, O(N) or O(lmax)) */

/* chealpix nside2npix / pix2ang_nest, call sites reference source/c_matrix_generator.cpp:34,42,170,182 */
int64_t cmg_nside2npix(int64_t nside);
cmg_status cmg_pix2ang_nest(int64_t nside, int64_t ipix, double* theta, double* phi);
/* CMatrix::getIndex / storage size, reference source/c_matrix.cpp:19-39 (64-bit) */
int64_t cmg_packed_size(int64_t dim);
int64_t cmg_packed_index(int64_t i, int64_t j);
/* Utils::readMask selection rule, reference source/utils.cpp:45-51: ascending i with mask[i] > 0.5.
 * good must hold npix entries; *n_good receives the count. */
cmg_status cmg_good_pixels_from_mask(const double* mask, int64_t npix, int32_t* good, int64_t* n_good);
/* Utils::beamFunction, reference source/utils.cpp:54-64 (fwhm in degrees; 0 => 1) */
double cmg_beam_function(int l, double fwhm_deg);
/* Utils::readPixelWindowFunction arithmetic, reference source/utils.cpp:154-160:
 * f[l] = pixwin[l] * beam(l), l = 0..lmax; pixwin == NULL => window 1 */
cmg_status cmg_window_beam(double* f, int lmax, double fwhm_deg, const double* pixwin);

/* ---------------------------------------------------------------- geometry ------------------ */

/* Pixel set of all following generate calls: HEALPix NESTED indices good_nest[0..n_good) in the
 * caller's order (NULL => all 12 nside^2 pixels).  Computes unit vectors exactly as reference
 * source/c_matrix_generator.cpp:178-185 (host libm, so they are bit-identical to the reference's)
 * plus the local (e_theta, e_phi) bases, and keeps them resident on the GPU. */
cmg_status cmg_set_pixels(cmg_ctx* ctx, int64_t nside, const int32_t* good_nest, int64_t n_good);
int64_t cmg_npix(const cmg_ctx* ctx);
/* copy back the resident geometry: out[8][npix] = x,y,z, e_theta(x,y,z), e_phi(x,y) */
cmg_status cmg_get_geometry(cmg_ctx* ctx, double* out);

/* ---------------------------------------------------------------- TT generation ------------- */

/* Columns [col_begin, col_end) of  S_ij = sum_{l=0}^{lmax} a[l] P_l(n_i.n_j)  into d_out, whose first
 * element is entry (0, col_begin) (so a rank's shard is one contiguous piece of the packed triangle).
 * a[] is a HOST array of lmax+1 series weights.  This one entry point serves
 *   clToCMatrix        a[l] = cl[l](2l+1)/(4pi) B_l^2, l>=2          (reference c_matrix_generator.cpp:164-232)
 *   getFiducialMatrix  a[l] = that for lmax<l<=4nside, a[0]=a[1]=100 cl[2] B_2^2   (:705-772; (1+z) = P_0+P_1)
 */
cmg_status cmg_legendre_series(cmg_ctx* ctx, const double* a, int lmax,
                               int64_t col_begin, int64_t col_end, double* d_out);
/* same, with the lmax+1 weights already resident on the device (batched / graph-replayed use) */
cmg_status cmg_legendre_series_dev(cmg_ctx* ctx, const double* d_a, int lmax,
                                   int64_t col_begin, int64_t col_end, double* d_out);

/* weights of clToCMatrix from C_l and the window*beam factors f[l] (cmg_window_beam): a[0]=a[1]=0 */
cmg_status cmg_tt_weights(const double* cl, const double* f, int lmax, double* a);
/* weights of getFiducialMatrix: cl and f hold 4*nside+1 entries, a receives 4*nside+1 */
cmg_status cmg_fiducial_weights(const double* cl, const double* f, int64_t nside, int lmax, double* a);

/* Whole-call equivalents of the reference API with HOST output (device work + copy back):
 * CMatrixGenerator::clToCMatrix(cl, nSide, fwhm, goodPixels) -- geometry must have been set with the
 * same nside/goodPixels.  out_packed holds npix(npix+1)/2 doubles (pinned memory makes the copy faster). */
cmg_status cmg_cl_to_cmatrix(cmg_ctx* ctx, const double* cl, int lmax, double fwhm_deg,
                             const double* pixwin, double* out_packed);
cmg_status cmg_fiducial_matrix(cmg_ctx* ctx, const double* cl, int lmax, double fwhm_deg,
                               const double* pixwin, double* out_packed);
/* The same two with DEVICE output (d_out: npix(npix+1)/2 doubles on the context's GPU): nothing but C_l crosses PCIe.  What the
 * C++ drop-in's device-resident CMatrix is filled by. */
cmg_status cmg_cl_to_cmatrix_dev(cmg_ctx* ctx, const double* cl, int lmax, double fwhm_deg, const double* pixwin, double* d_out);
cmg_status cmg_fiducial_matrix_dev(cmg_ctx* ctx, const double* cl, int lmax, double fwhm_deg, const double* pixwin, double* d_out);
/* A packed matrix of dimension dim from device to host memory.  full_sky_strips = 1 or 3 vouches that it is the full-sky NESTED
 * [T] / [T;Q;U] matrix of the context's current geometry (dim = strips x npix), which lets the copy use the host expansion
 * (cmg_set_host_expand); 0 = any packed matrix, one plain copy. */
cmg_status cmg_matrix_to_host(cmg_ctx* ctx, const double* d_packed, int64_t dim, int full_sky_strips, double* out_packed);
/* CMatrixGenerator::generateNoiseMatrix, reference c_matrix_generator.cpp:774-787 (host, trivial) */
cmg_status cmg_noise_matrix(int64_t npix, double noise, double* out_packed);
/* CMatrix::maskMatrix gather, reference source/c_matrix.cpp:182-201; device buffers */
cmg_status cmg_mask_matrix(cmg_ctx* ctx, const double* d_in_packed, int64_t npix_in,
                           const int32_t* good, int64_t n_good, double* d_out_packed);

/* ---------------------------------------------------------------- T,Q,U generation ---------- */

/* Where the entries of a (possibly sharded) polarized matrix live.  The 3N columns are owned in
 * pixel-column blocks: part k owns pixel columns [begin[k], begin[k+1]) of each of the T, Q and U
 * strips.  kind 0 (packed): ptr[s] is the address of entry (0, s*N + begin[k]) of a contiguous run
 * of packed columns.  kind 1 (dense blocks; only for entries a rank computes on behalf of another
 * owner): ptr[0..2] are three column-major blocks holding <Q_i T_j>, <U_i T_j>, <U_i Q_j> for
 * owner-columns i in [begin[k], begin[k+1]) and rows j in [row0, row0+ld): element at
 * ptr[t][(i-begin[k])*ld + (j-row0)].  Pointers may be peer-mapped memory of another GPU. */
typedef struct cmg_tqu_layout {
    int32_t n_parts;
    int32_t own;                               /* the part whose pixel columns this call computes */
    int64_t begin[CMG_MAX_PARTS + 1];
    double* ptr[CMG_MAX_PARTS][3];
    int32_t kind[CMG_MAX_PARTS];
    int64_t ld[CMG_MAX_PARTS];
    int64_t row0[CMG_MAX_PARTS];
} cmg_tqu_layout;

/* Fills a single-owner layout over one whole packed buffer of dimension 3N (n_parts = 1). */
cmg_status cmg_tqu_layout_single(cmg_ctx* ctx, double* d_packed, cmg_tqu_layout* layout);

/* All pixel pairs (i <= j), j in the pixel columns of part `own`: the 3x3 blocks
 *   TT = sum a_tt[l] P_l,  <T Q'> = -sum a_te[l] d^l_20,  <QQ'>+<UU'> = sum (a_ee+a_bb)[l] d^l_22,
 *   <QQ'>-<UU'> = sum (a_ee-a_bb)[l] d^l_2-2   (great-circle frame; F^10, F^12-+F^22 of Tegmark &
 *   de Oliveira-Costa 2001), rotated into the local frames, scattered to all nine destinations.
 * a_xx are HOST arrays of lmax+1 weights C^XX_l (2l+1)/(4pi) x window factors; entries l<2 ignored. */
cmg_status cmg_tqu(cmg_ctx* ctx, const double* a_tt, const double* a_te, const double* a_ee,
                   const double* a_bb, int lmax, const cmg_tqu_layout* layout);
/* Places one dense kind-1 block (see cmg_tqu_layout) into a whole packed 3N triangle on this GPU: kind 0/1/2 =
 * <Q_i T_j>, <U_i T_j>, <U_i Q_j> for owner columns i in [col0, col0+n_cols) and rows j in [row0, row0+ld).  Used when
 * the unsharded matrix must be assembled (after the blocks were moved with NCCL or a copy). */
cmg_status cmg_tqu_scatter_block(cmg_ctx* ctx, const double* d_block, int64_t col0, int64_t n_cols, int64_t ld, int64_t row0,
                                 int kind, double* d_full_packed);
/* same with the 4 x (lmax+1) weights (tt, te, ee, bb, each lmax+1 doubles, contiguous) already on the device: nothing is
 * read from the host, so the call can be captured in a CUDA graph and replayed after the weights were updated in place
 * (shared-memory coefficient table kernel; the parameter-block kernel needs the weights on the host at launch) */
cmg_status cmg_tqu_dev(cmg_ctx* ctx, const double* d_a, int lmax, const cmg_tqu_layout* layout);
/* The same matrix as cmg_tqu for the FULL sky (what cmg_cl_to_cmatrix_pol and the drop-in classes run there), evaluated once per
 * orbit of pixel pairs under the pi/2 rotation of the HEALPix grid about the polar axis (NESTED: base face f -> next face of
 * its ring, index inside the face kept).  Frames rotate with the pixels, so C[X Ra, Y Rb] = C[X a, Y b] and the four sums of
 * a pair are computed once and stored at its (up to) four images: a quarter of the recurrence work of cmg_tqu (mode 0), or
 * 1/3.2 without transposed images (mode 1).  Entries of different images of one orbit are bit-identical to each other; each
 * differs from cmg_tqu's by the rounding of its own n_i.n_j only (the 1e-11 gate holds with the same margin).
 * Needs cmg_set_pixels(ctx, nside >= 8, NULL, 0), 2 <= lmax <= 441, a single-owner packed buffer d_packed of dimension 3N.
 * mode 0 computes the store destinations of a tile once into shared memory for the classes without transposed images (measured
 * 36.2 ms against 36.8 ms per Nside = 64 matrix); mode 2 = mode 0 without that table (kept for the comparison); same classes,
 * same storage, same results.
 * mode 3 = mode 0 + the meridian mirror of the grid (phi -> pi/2 - phi: in-face (ix, iy) -> (iy, ix), entries with exactly one U
 * index change sign): five of the twelve classes of face pairs of different rings are stored as mirror images of five others,
 * 13 instead of 18 of 72 face-pair units are evaluated.  Same matrix (images differ from cmg_tqu's entries by roundings only).
 * Opt-in: measured 34.3 against 35.9 ms at Nside = 64 -- the eight images per evaluated pair are bound by the scattered-store
 * rate of the memory system, not by arithmetic (open items).  Single owner only. */
cmg_status cmg_tqu_orbit(cmg_ctx* ctx, const double* a_tt, const double* a_te, const double* a_ee,
                         const double* a_bb, int lmax, double* d_packed, int mode);
/* The same over several GPUs.  Rank r of n_ranks owns the in-face index range [bounds[r], bounds[r+1]) of ALL twelve base
 * faces -- an orbit-closed set of pixel columns -- evaluates the source pairs whose column pixel lies there and stores all their
 * images (no data-path collective to PRODUCE the entries):
 *   strip[s][f]   address of entry (0, s N + f nside^2 + q_begin): the packed columns s N + f nside^2 + [q_begin, q_end), one
 *                 contiguous piece of the packed triangle each (s = T, Q, U strip; f = base face);
 *   outbox        the entries whose packed column belongs to another rank (row pixel a' of the pair outside the range; a third
 *                 of the nine entries of such a pair), compact -- every element is written exactly once -- and ordered by
 *                 destination rank: cmg_orbit_outbox_layout gives the offset of the block for each destination.  Inside the
 *                 block for rank d, for every (class, image, staged kind) combination `combo` that can address d (numbered over
 *                 the plan of cmg_tqu_orbit_plan: whole-face-pair classes first; the q_row <= q_col classes only address d < r),
 *                 the 32 x 32 sub-tiles (row half-tile h of d's range, column tile ct of r's range), row-major:
 *                     block[((combo nct_r + ct) nh_d + h) 1024 + (q_a' mod 32) 32 + (q_b' mod 32)].
 *                 NULL when n_ranks == 1.
 * bounds are multiples of 32.  To complete the strips, block(r -> d) has to reach rank d (one all-to-all over NVLink, or the
 * receiver reads the sender's outbox through CUDA IPC) and cmg_tqu_orbit_scatter_inbox places it; after that a rank holds
 * exactly its 36 runs of packed columns, complete.  cmg_tqu_orbit_assemble places a rank's pieces into a whole packed triangle on
 * this GPU (tests, single-GPU assembly): parts & 1 = its strips, parts & 2 = its outbox blocks.  A strip has holes where another
 * rank's outbox holds the entry, so place the strips of ALL ranks before any outbox. */
typedef struct cmg_orbit_shard {
    int32_t n_ranks, rank;
    int64_t bounds[CMG_MAX_PARTS + 1];
    double* strip[3][12];
    double* outbox;
} cmg_orbit_shard;
cmg_status cmg_tqu_orbit_sharded(cmg_ctx* ctx, const double* a_tt, const double* a_te, const double* a_ee,
                                 const double* a_bb, int lmax, const cmg_orbit_shard* shard, int mode);
/* offsets[d], d = 0 .. n_ranks: first element (in doubles) of the block for destination d inside rank `rank`'s outbox;
 * offsets[n_ranks] = doubles in the whole outbox (pure host arithmetic) */
cmg_status cmg_orbit_outbox_layout(int64_t nside, int mode, int n_ranks, const int64_t* bounds, int rank, int64_t* offsets);
/* places block(sender -> shard->rank), wherever it is now (device memory of this GPU, or peer memory of the sender mapped with
 * cmg_ipc_open: the kernel's loads then are the transfer), into this rank's strips */
cmg_status cmg_tqu_orbit_scatter_inbox(cmg_ctx* ctx, const cmg_orbit_shard* shard, int mode, int sender, const double* d_block);
cmg_status cmg_tqu_orbit_assemble(cmg_ctx* ctx, const cmg_orbit_shard* shard, int mode, int parts, double* d_full_packed);
/* Brings a rank's COMPLETE strips to the host as its columns of one whole packed matrix `host_packed` (dimension 3N; for several
 * ranks on one box a shared mapping every rank writes its own columns of).  threads > 0: only the columns of base faces 3, 7, 11
 * cross PCIe and `threads` host threads fill in this rank's columns of the other faces as rotated images while the copies are in
 * flight (cmg_host_expand_rotations restricted to [q_begin, q_end)); threads = 0: all 36 runs are copied.  Page-locked
 * destinations (cmg_host_register) make the copies run at PCIe speed.  direct_mask (with threads > 0), bit 3 s + k - 1: the image
 * of strip s in the k-th face below the last one of every ring is copied over PCIe as well instead of being filled in by the
 * host threads -- a balance between the copy engines and the host threads for hosts where the two do not share one memory
 * write bandwidth (0 on the measured one).  Returns when the rank's columns are complete in host memory. */
cmg_status cmg_orbit_strips_to_host(cmg_ctx* ctx, const cmg_orbit_shard* shard, double* host_packed, int threads, int direct_mask);
/* cudaHostRegister / cudaHostUnregister of caller-owned memory (e.g. a shared mapping) */
cmg_status cmg_host_register(void* ptr, int64_t bytes);
cmg_status cmg_host_unregister(void* ptr);
/* The TT matrix of cmg_legendre_series (any series weights: clToCMatrix, getFiducialMatrix) over the same orbits.  One owner:
 * the plan with transposed images, 18 of the 72 face-pair units -- a quarter of the recurrence work; a transposed image is stored
 * as one 64-byte run per thread, which does not show behind the series.  Full sky, nside >= 16, d_out = the whole packed triangle
 * of dimension N.  cmg_cl_to_cmatrix / cmg_fiducial_matrix take this path by themselves on the full sky (any non-zero
 * cmg_set_kernel_variant pins the every-pair kernels). */
cmg_status cmg_legendre_series_orbit(cmg_ctx* ctx, const double* a, int lmax, double* d_out);
/* The same over several GPUs: a rank owns the in-face range [q_begin, q_end) (multiples of 16) of all twelve base faces and
 * writes the packed columns f nside^2 + [q_begin, q_end) into d_strips[f] (entry (0, first column) first; 12 contiguous pieces of
 * the packed triangle).  It uses the plan WITHOUT transposed images (22.5 of 72 units): every image then has its row pixel before
 * its column pixel, so every entry lands in a column of the rank that evaluated it -- no outbox, no exchange, the strips are
 * complete when the kernel ends. */
cmg_status cmg_legendre_series_orbit_sharded(cmg_ctx* ctx, const double* a, int lmax, int64_t q_begin, int64_t q_end, double* const* d_strips);
/* Host half of a full-sky whole call (pure CPU): `packed` is a packed matrix of dimension N (TT) or 3N ([T;Q;U]) in HOST memory
 * whose columns of the last face of every ring of four base faces (faces 3, 7, 11, of each strip) are filled in; the columns
 * of the other faces in [face_begin, face_end), strips [strip_begin, strip_end), are written as their rotated images (runs
 * of nside^2 rows copied from the matching column of the ring's last face), `threads` host threads.  With it only 27 % of the
 * matrix has to cross PCIe; cmg_set_host_expand makes the whole calls work that way. */
cmg_status cmg_host_expand_rotations(double* packed, int64_t nside, int strip_begin, int strip_end, int face_begin, int face_end,
                                     int threads);
/* How the full-sky whole calls (cmg_cl_to_cmatrix_pol, cmg_cl_to_cmatrix, cmg_fiducial_matrix) bring the matrix to the host:
 * threads > 0 = copy back only the last-face columns and fill in the rest with the host expansion on `threads` host threads while
 * the remaining copies are in flight; 0 = one plain copy of the whole matrix; -1 (the default) = automatic: the expansion on all
 * host cores for matrices of 1 GiB and more.  Measured for the 87 GB matrix on a 16-core host: 1.75 s plain, 1.38 s expanded
 * (round 2, first version; profiles/README.md has the current figure). */
cmg_status cmg_set_host_expand(cmg_ctx* ctx, int threads);
/* which images the [T;Q;U] whole call copies over PCIe next to the last-face columns (bit 3 strip + k - 1, as direct_mask of
 * cmg_orbit_strips_to_host); default 0.  Measured with 0x40 on a 16-core host: 991 ms against 840 ms for the 87 GB matrix -- the
 * call is bound by the host's memory WRITE bandwidth (~105 GB/s there), which DMA and streaming stores share. */
cmg_status cmg_set_host_expand_direct(cmg_ctx* ctx, int direct_mask);
/* the classes of base-face pairs cmg_tqu_orbit works through (host only; for tests): out[c][CMG_ORBIT_CLASS_INTS] =
 * { row face, column face, only q_row <= q_col, same face, n_images, then for image k = 0..3: row face, column face,
 *   stored transposed }, then for image k = 0..3 the outbox number of (this class, image k, staged kind 0) (cmg_orbit_shard);
 * out must hold CMG_ORBIT_MAX_CLASSES classes */
#define CMG_ORBIT_CLASS_INTS 21
#define CMG_ORBIT_MAX_CLASSES 24
cmg_status cmg_tqu_orbit_plan(int64_t nside, int mode, int32_t* out, int32_t* n_classes);
/* the same classes' mirror images (mode 3): out[c][5] = { 1 when the class also stores its image under the meridian mirror, then the
 * row face of mirror image k = 0..3 (its column face is that of image k; both in-face indices have their even and odd bits swapped,
 * entries with exactly one U index change sign) } */
cmg_status cmg_tqu_orbit_plan_mirror(int64_t nside, int mode, int32_t* out, int32_t* n_classes);
/* weights from spectra and the temperature / polarization window*beam factors */
cmg_status cmg_tqu_weights(const double* ctt, const double* cte, const double* cee, const double* cbb,
                           const double* fT, const double* fP, int lmax,
                           double* a_tt, double* a_te, double* a_ee, double* a_bb);
/* whole call with HOST output: out_packed holds 3N(3N+1)/2 doubles.  TT carries pixwinT^2, TE pixwinT pixwinP, EE and BB
 * pixwinP^2 (each times the beam); the reference's own polarization routine uses the TEMPERATURE table throughout
 * (source/c_matrix_generator.cpp:534: readPixelWindowFunction without the polarization flag) -- pass pixwinP = pixwinT for that. */
cmg_status cmg_cl_to_cmatrix_pol(cmg_ctx* ctx, const double* ctt, const double* cte, const double* cee,
                                 const double* cbb, int lmax, double fwhm_deg,
                                 const double* pixwinT, const double* pixwinP, double* out_packed);
/* same with DEVICE output */
cmg_status cmg_cl_to_cmatrix_pol_dev(cmg_ctx* ctx, const double* ctt, const double* cte, const double* cee,
                                     const double* cbb, int lmax, double fwhm_deg,
                                     const double* pixwinT, const double* pixwinP, double* d_out);

/* ---------------------------------------------------------------- batched regeneration ------ */

/* n_batch matrices for n_batch weight sets in one launch (one MCMC step's proposals):
 * a[b][l], b < n_batch; d_out[b] receives columns [col_begin,col_end) of matrix b,
 * consecutive matrices `stride` doubles apart. */
cmg_status cmg_legendre_series_batched(cmg_ctx* ctx, const double* a, int lmax, int64_t n_batch,
                                       int64_t col_begin, int64_t col_end, double* d_out, int64_t stride);
/* a[b][4][lmax+1] in the order tt, te, ee, bb; single-owner packed layout per batch element */
cmg_status cmg_tqu_batched(cmg_ctx* ctx, const double* a, int lmax, int64_t n_batch,
                           double* d_out, int64_t stride);

/* Batched T,Q,U on the FP64 tensor path (DMMA), slab output.  The reference has no batched generator, so the layout of
 * a batch is this library's to define.  The tensor-core accumulator tile is contiguous along the batch axis, and that is
 * the axis a slab keeps contiguous in memory:
 *     slab k holds batch elements 16 k .. 16 k + 15, interleaved entry by entry:
 *     d_slabs[k * cmg_slab_doubles(3 npix) + e * CMG_SLAB + (b % CMG_SLAB)],   e = i + j (j + 1) / 2, the reference's
 *     packed index (source/c_matrix.cpp:19-39) in the [T;Q;U] ordering,
 * so every batch element is an ordinary packed CMatrix with element stride CMG_SLAB.  d_slabs must hold
 * ceil(n_batch / 16) slabs; elements beyond n_batch in the last slab are written as zeros.  2 <= lmax <= CMG_SLAB_LMAX;
 * single-owner layout only (a batch shards over GPUs along the batch axis).  a[b][4][lmax+1] host, as above. */
#define CMG_SLAB 16
#define CMG_SLAB_LMAX 63
int64_t cmg_slab_doubles(int64_t dim);
cmg_status cmg_tqu_batched_slab(cmg_ctx* ctx, const double* a, int lmax, int64_t n_batch, double* d_slabs);
/* same with the weights a[b][4][lmax+1] already on the device (e.g. written by a C_l emulator running on the GPU): nothing is
 * read from the host.  Capture in a CUDA graph only after one plain call with the same n_batch (it sizes the staging buffer). */
cmg_status cmg_tqu_batched_slab_dev(cmg_ctx* ctx, const double* d_a, int lmax, int64_t n_batch, double* d_slabs);
/* one slab -> separate packed matrices of dimension dim: d_out[b * out_stride + e] for b < n_live (only_b < 0), or the
 * single element only_b -> d_out[e] */
cmg_status cmg_slab_unpack(cmg_ctx* ctx, const double* d_slab, int64_t dim, int n_live, int only_b,
                           double* d_out, int64_t out_stride);

/* ---------------------------------------------------------------- towards the consumer ------- */

/* First step of the likelihood that consumes these matrices (reference source/likelihood.cpp:100-110): the sum of up to
 * three packed matrices of dimension n (d_f and/or d_n may be NULL) written as a FULL symmetric column-major n x n matrix
 * on the device, ready for a dense Cholesky factorisation, in one pass over HBM. */
cmg_status cmg_sum_unpack(cmg_ctx* ctx, const double* d_c, const double* d_f, const double* d_n, int64_t n, double* d_full);
/* same with an element stride on d_c: c_stride = CMG_SLAB and d_c = slab + (b % CMG_SLAB) reads element b of a slab */
cmg_status cmg_sum_unpack_strided(cmg_ctx* ctx, const double* d_c, int64_t c_stride, const double* d_f, const double* d_n,
                                  int64_t n, double* d_full);

/* The factorisation of that consumer, on the device and IN PLACE on the packed triangle: A = U^T U with U upper triangular in the
 * same packed layout (LAPACK dpptrf 'U', what the reference runs on the host: source/matrix_impl.cpp:236-263 on the storage of
 * include/matrix_impl.hpp:495-502).  Blocked right-looking; the trailing updates -- the n^3 / 3 -- are FP64 tensor-core
 * contractions (mma.sync.m8n8k4.f64) whose operands are contiguous runs of packed columns.  Nothing is unpacked: memory is the
 * n (n + 1) / 2 doubles of the matrix itself, so the 147456-dimensional matrix of Nside = 64 is factorised where the generator
 * left it.  *info = 0, or k > 0 when the leading minor of order k is not positive definite (d_packed is then partly overwritten). */
cmg_status cmg_packed_cholesky(cmg_ctx* ctx, double* d_packed, int64_t n, int64_t* info);
/* blocks of 128 rows per trailing update (1 .. 4; 0 = default: 2 below n = 16384, 4 from there on): the trailing matrix is read and written once per GROUP of blocks;
 * workspace blocks x (n + 192) x 128 doubles on the device (the rows of U of the group in dense form, the operands of the update) */
cmg_status cmg_set_cholesky_group(cmg_ctx* ctx, int blocks);
/* look-ahead (default on): the next group's diagonal blocks and panels are factorised on a high-priority side stream beside the
 * bulk of the trailing update; a second set of planes.  The call still completes in the order of the context's stream. */
cmg_status cmg_set_cholesky_lookahead(cmg_ctx* ctx, int on);
/* log det A = 2 sum_i log U_ii from the factor */
cmg_status cmg_packed_cholesky_logdet(cmg_ctx* ctx, const double* d_factor, int64_t n, double* log_det);
/* y = U^-T t for n_rhs right-hand sides, in place: d_rhs is n x n_rhs column-major on the device; t^T A^-1 t = |y|^2 */
cmg_status cmg_packed_cholesky_solve(cmg_ctx* ctx, const double* d_factor, int64_t n, double* d_rhs, int64_t n_rhs);

/* ---- The same factorisation with the columns spread over several GPUs (cosmopp_b200/multigpu.py: ShardedCholesky drives the
 * steps; the library itself links no collective library -- the two exchanges of a step are the caller's).
 * A rank holds up to CMG_CHOL_MAX_RUNS runs of WHOLE packed columns [col_begin, col_end), ascending, every boundary a multiple of
 * 128 (the end of the matrix excepted); d_run[r] = address of entry (0, col_begin[r]).  These are exactly the 36 strips a rank
 * of cmg_orbit_shard holds after the exchange (partition boundaries rounded to 128), so the Nside = 64 matrix is factorised
 * where the generator and the exchange left it: no gather, no redistribution.
 * Blocks of 128 rows are taken in groups of S <= 4 (as cmg_packed_cholesky does, cmg_set_cholesky_group).  The dense panel is
 * S planes of (n + 192) x 128 doubles, d_panel[s * plane_stride + (column - panel_col0) * 128 + row]; panel_col0 = first row of
 * the group.  Block s of a group (k0 = group start + 128 s, kb = min(128, n - k0)):
 *   cmg_chol_syrk   (every rank, s > 0, strip_only = 1, k0 = group start, kb = 128 s)  rows of block s of the rank's own columns
 *                   catch up with the blocks of the group already solved.
 *   cmg_chol_diag   (owner of the block only)  U_kk in place, and packed into d_ukk: kb (kb + 1) / 2 entries, then kb reciprocal
 *                   pivots.  The caller broadcasts d_ukk (<= 67 KB) from the owner.
 *   cmg_chol_panel  (every rank, kb = 128)  rows k0 .. k0 + 128 of the rank's own columns behind the block: solved in place and
 *                   also written to plane s (d_plane = d_panel + s * plane_stride).  The caller zeroes that part of the plane
 *                   before and all-reduces (sum) it after: every rank then holds the 128 rows of EVERY column behind the block.
 * and after the last block of the group
 *   cmg_chol_syrk   (every rank, strip_only = 0, k0 = group start, kb = 128 S)  trailing update of the rank's own columns behind
 *                   the group, operands read from the S planes.
 * shift (a multiple of 128, normally 0): the update starts at row and column k0 + kb + shift -- the update of a group can be issued
 * in pieces, first strip by strip the rows the NEXT group factorises (strip_only = 1, shift = 0, 128, ...), then the rest
 * (shift = 128 S) while the next group's blocks are factorised on another stream (multigpu.ShardedCholesky, look-ahead).
 * cmg_chol_begin before the first step, cmg_chol_end after the last (*info as cmg_packed_cholesky: the first non-positive pivot
 * of a block this rank owns, 0 otherwise; the caller takes the minimum of the non-zero values over ranks). */
#define CMG_CHOL_MAX_RUNS 36
typedef struct cmg_chol_runs
{
    int32_t n_runs;
    int64_t col_begin[CMG_CHOL_MAX_RUNS];
    int64_t col_end[CMG_CHOL_MAX_RUNS];
    double* d_run[CMG_CHOL_MAX_RUNS];
} cmg_chol_runs;
cmg_status cmg_chol_begin(cmg_ctx* ctx);
cmg_status cmg_chol_end(cmg_ctx* ctx, int64_t* info);
cmg_status cmg_chol_diag(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, double* d_ukk);
cmg_status cmg_chol_panel(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, const double* d_ukk, double* d_plane, int64_t panel_col0);
cmg_status cmg_chol_syrk(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t shift, const double* d_panel, int64_t plane_stride,
                         int64_t panel_col0, int strip_only);
/* this rank's share of log det A: 2 sum log U_jj over its columns (the caller sums over ranks) */
cmg_status cmg_chol_logdet_runs(cmg_ctx* ctx, const cmg_chol_runs* runs, double* log_det_share);
/* y = U^-T t on the sharded factor, d_rhs (n x n_rhs, column-major) replicated on every rank.  Step k: the owner of block k solves
 * its kb rows (cmg_chol_solve_diag; rows k0 .. of d_rhs then hold y) and the caller broadcasts them; every rank subtracts their
 * contribution from the rows of its OWN columns behind the block (cmg_chol_solve_update) -- the only rows it will ever solve. */
cmg_status cmg_chol_solve_diag(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t n, double* d_rhs, int64_t n_rhs);
cmg_status cmg_chol_solve_update(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t n, double* d_rhs, int64_t n_rhs);

/* d_out = C + F + N, all packed of dimension n (d_f, d_n may be NULL; element stride c_stride on d_c as in cmg_sum_unpack_strided;
 * d_out may be d_c when c_stride = 1) */
cmg_status cmg_packed_sum(cmg_ctx* ctx, const double* d_c, int64_t c_stride, const double* d_f, const double* d_n, int64_t n, double* d_out);
/* how cmg_like_create factorises: 0 (default) = cmg_packed_cholesky on the packed sum; 1 = cusolverDnDpotrf on the unpacked
 * n x n matrix (a plain library call, twice the memory, n <= 46340; kept for comparison) */
cmg_status cmg_set_like_method(cmg_ctx* ctx, int method);

/* The temperature pixel likelihood of reference source/likelihood.cpp (`Likelihood::construct` :68-134, `calculate`
 * :163-180) with everything resident on the device: C + F + N (packed, device; d_c with element stride c_stride) ->
 * Cholesky factor -> log det - offset (the reference's constant -29677.0566, :126); `foreground` (host, n values or NULL)
 * is the template marginalised over.  CMG_ENUMERIC if the sum is not positive definite (the reference throws). */
typedef struct cmg_like cmg_like;
cmg_status cmg_like_create(cmg_ctx* ctx, const double* d_c, int64_t c_stride, const double* d_f, const double* d_n,
                           int64_t n, const double* foreground, cmg_like** out);
/* n_maps maps (host, map k at t + k n): chi2[k] = t^T C^-1 t (minus the foreground projection), *log_det = log det term
 * (with the foreground term when a template was given); -2 log L = chi2[k] + *log_det */
cmg_status cmg_like_calculate(cmg_like* like, const double* t, int64_t n_maps, double* chi2, double* log_det);
void cmg_like_destroy(cmg_like* like);
/* The same consumer in the inverse-noise form of the reference's LikelihoodPolarization (source/likelihood.cpp:341-406):
 * K = N^-1 + N^-1 C N^-1 from the packed covariance d_c and the packed symmetric inverse noise matrix d_ninv (dimension m, e.g. the
 * [Q;U] block over the unmasked pixels), factorised on the device; log det K - det_offset (the reference subtracts 16078.083180).
 * cmg_like_calculate on the handle gives chi2 = v^T K^-1 v for v = N^-1 d (the reference's `v`, source/likelihood.cpp:536-612,
 * minus N^-1 times the temperature-predicted map, which needs the reference's SHT machinery and is the caller's to subtract). */
cmg_status cmg_like_create_ninv(cmg_ctx* ctx, const double* d_c, const double* d_ninv, int64_t m, double det_offset, cmg_like** out);

/* ---------------------------------------------------------------- CMatrix files from / to device memory ---- */

/* The reference's binary CMatrix file (source/c_matrix.cpp:41-104: int32 nPix, packed doubles, int32 length, comment)
 * written from and read into DEVICE buffers piece by piece, so that a matrix that lives sharded on the GPU(s) (a rank's
 * packed strip is one contiguous element range) never has to be assembled in host memory.  Pieces may arrive in any order;
 * copies go through pinned bounce buffers and overlap the file I/O.  Elements that no piece covers read back as 0. */
typedef struct cmg_file cmg_file;
cmg_status cmg_cmatrix_file_create(cmg_ctx* ctx, const char* path, int64_t n_pix, cmg_file** out);
cmg_status cmg_cmatrix_file_write_device(cmg_file* file, int64_t first_element, const double* d_src, int64_t count);
cmg_status cmg_cmatrix_file_open(cmg_ctx* ctx, const char* path, int64_t* n_pix, cmg_file** out);
cmg_status cmg_cmatrix_file_read_device(cmg_file* file, int64_t first_element, double* d_dst, int64_t count);
cmg_status cmg_cmatrix_file_comment(cmg_file* file, char* buffer, int64_t capacity);
/* writes the comment (NULL = empty) when the file was created for writing; releases the handle either way */
cmg_status cmg_cmatrix_file_close(cmg_file* file, const char* comment);

/* ---------------------------------------------------------------- measurement --------------- */

/* dependent-free DFMA microbenchmark: achieved FP64 TFLOP/s on this GPU (the roofline denominator) */
cmg_status cmg_measure_fp64_peak(cmg_ctx* ctx, double* tflops);
/* milliseconds the last generate call's kernels took on the device (CUDA events on the launch stream) */
cmg_status cmg_last_kernel_ms(cmg_ctx* ctx, double* ms);
/* tuning hook: pick the T,Q,U kernel variant.  0 = automatic; otherwise 100*S + 10*R + B with S = 1 for the
 * coefficient table in the kernel parameter block (0 = staged in shared memory), R columns per thread and
 * B resident CTAs per SM the kernel is compiled for (only the combinations built into the library) */
cmg_status cmg_set_kernel_variant(cmg_ctx* ctx, int variant);
/* switch the per-call event timing on/off (off by default: it synchronises) */
cmg_status cmg_set_timing(cmg_ctx* ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* CMG_H */
