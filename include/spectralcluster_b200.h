/*
 * spectralcluster_b200 -- C ABI of the B200-native SpectralClusterer.predict() hot path.
 *
 * The reference (wq2012/SpectralCluster v0.2.22) is pure Python and has no FFI:
 * its drop-in boundary is the Python object surface.  This header
 * is the boundary underneath our Python mirror of that surface: one entry point per
 * reference operator, bound with ctypes by spectralcluster_b200/_native.py.  Every
 * declaration cites the reference function (file:line under /root/reference) whose
 * arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from sc_last_error() (thread local).
 *   - all matrix pointers are DEVICE pointers (cudaMalloc'ed by the caller --
 *     the Python host uses torch tensors purely as device buffers), row-major,
 *     with an explicit leading dimension `ld` counted in elements.  fp32 matrices
 *     need ld % 4 == 0 and 16-byte aligned bases; fp16 planes need ld % 8 == 0.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are
 *     asynchronous unless stated otherwise.
 *   - "split fp16 planes" (hi, lo): value = hi + lo with hi = fp16(value),
 *     lo = fp16(value - hi): the operand format of the tcgen05 GEMMs (three
 *     kind::f16 MMAs hi*hi + hi*lo + lo*hi reproduce an fp32-accurate product).
 */
#ifndef SPECTRALCLUSTER_B200_H_
#define SPECTRALCLUSTER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_ABI_VERSION 2

/* enums mirror the reference's enum *names* (values are ours) */
enum sc_threshold_type { SC_THRESHOLD_ROWMAX = 0, SC_THRESHOLD_PERCENTILE = 1 }; /* refinement.py:21-27 */
enum sc_symmetrize_type { SC_SYMMETRIZE_MAX = 0, SC_SYMMETRIZE_AVERAGE = 1 };    /* refinement.py:30-36 */
enum sc_laplacian_type {                                                          /* laplacian.py:9-21 */
  SC_LAPLACIAN_AFFINITY = 0, SC_LAPLACIAN_UNNORMALIZED = 1,
  SC_LAPLACIAN_RANDOMWALK = 2, SC_LAPLACIAN_GRAPHCUT = 3
};
enum sc_gemm_engine {
  SC_GEMM_TCGEN05 = 0,   /* tcgen05.mma kind::f16 on split planes, TMA-fed (product path) */
  SC_GEMM_SIMT_F64ACC = 1 /* SIMT fp32 operands, fp64 accumulation (validation path, small N) */
};
enum sc_gemm_precision {
  SC_GEMM_SPLIT3 = 0,    /* hi*hi + hi*lo + lo*hi : ~2^-22 relative per product */
  SC_GEMM_SINGLE = 1,    /* hi*hi only            : ~2^-11 relative per product */
  SC_GEMM_SPLIT2 = 2     /* (hi+lo)*hi : B rounded to fp16, zero-mean 2^-12 per product */
};
enum sc_which_end { SC_EIG_LARGEST = 0, SC_EIG_SMALLEST = 1 };

typedef struct sc_context sc_context;

int sc_abi_version(void);
const char* sc_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
long long sc_launch_count(void);

/* One context per (process, device).  Owns nothing the caller can see except
 * cached device properties and stream-ordered scratch. */
int sc_context_create(int device, sc_context** out);
int sc_context_destroy(sc_context* ctx);
int sc_context_sm_count(const sc_context* ctx);
/* Cap the CTA count of the persistent GEMM (0 = every SM) so that communication kernels issued
 * concurrently (NCCL send/recv in the sharded pipeline) find free SMs. */
int sc_context_set_gemm_sm_limit(sc_context* ctx, int sms);

/* ---- utils.compute_affinity_matrix (utils.py:20-41) --------------------------------- */
/* Row L2-normalisation (utils.py:32-33).  x is [n,d] fp32 (x_is_f64=0) or fp64 (=1) on the
 * device.  Norms are accumulated in fp64.  Any of xn (fp32 [n,ldxn]) and hi/lo (fp16
 * [n,ldh]) may be NULL.  Zero rows give NaN like the reference. */
int sc_normalize_rows(sc_context* ctx, const void* x, int x_is_f64, int64_t n, int64_t d,
                      int64_t ldx, float* xn, int64_t ldxn, void* hi, void* lo, int64_t ldh,
                      void* stream);

/* A = (Xn Xn^T + 1) / 2 (utils.py:35-39).  engine TCGEN05 reads the split planes,
 * engine SIMT reads xn.  If rowmax_offdiag != NULL it receives max_j!=i A[i,j] clamped at 0
 * (the CropDiagonal value, refinement.py:148-150) -- it must be zero-filled by the caller. */
int sc_affinity_cosine(sc_context* ctx, int engine, int precision, const float* xn, int64_t ldxn,
                       const void* hi, const void* lo, int64_t ldh, int64_t n, int64_t d,
                       float* a, int64_t lda, float* rowmax_offdiag, void* stream);

/* ---- refinement.py operators --------------------------------------------------------- */
/* CropDiagonal.refine (refinement.py:145-151): out = a with diag[i] = max(0, max_{j!=i} a[i,j]).
 * out may alias a. */
int sc_crop_diagonal(sc_context* ctx, const float* a, int64_t n, int64_t lda, float* out,
                     int64_t ldo, void* stream);

/* Only the new diagonal of CropDiagonal (the vector the fused blur pass substitutes on read). */
int sc_crop_diagonal_values(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                            float* diag_out, void* stream);

/* GaussianBlur.refine (refinement.py:160-162) == scipy.ndimage.gaussian_filter(a, sigma):
 * separable, axis 0 then axis 1, mode='reflect', truncate=4.0 (radius int(4*sigma+0.5)),
 * weights computed in fp64.  sigma <= 1e-15 copies.  diag_override (may be NULL) replaces
 * a[i,i] on read (fused CropDiagonal).  out == NULL runs the statistics-only pass.
 * rowmax_out (may be NULL; caller zero-fills; values must be >= 0) receives max_j out[i,j]. */
int sc_gaussian_blur(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                     const float* diag_override, double sigma, float* out, int64_t ldo,
                     float* rowmax_out, void* stream);

/* Statistics-only pass of the blur: rowmax_out[i] = max_j blur(a)[i,j], with the diagonal read
 * as zero when zero_diagonal != 0 (RowWiseThreshold's preserve_diagonal, refinement.py:185-186). */
int sc_gaussian_blur_rowmax(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                            const float* diag_override, double sigma, int zero_diagonal,
                            float* rowmax_out, void* stream);

/* RowWiseThreshold.refine (refinement.py:182-210), any option combination. */
int sc_row_threshold(sc_context* ctx, const float* a, int64_t n, int64_t lda, int type, double p,
                     double mult, int binarize, int preserve_diagonal, float* out, int64_t ldo,
                     void* stream);

/* Symmetrize.refine (refinement.py:219-226).  out must not alias a. */
int sc_symmetrize(sc_context* ctx, const float* a, int64_t n, int64_t lda, int type, float* out,
                  int64_t ldo, void* stream);

/* Fused GaussianBlur -> RowWiseThreshold(RowMax) -> Symmetrize for a SYMMETRIC input
 * (SURVEY.md A.3): y[i,j] = sym(t(b, m_i), t(b, m_j)), b = blur(a)[i,j], t = the threshold rule
 * with row maxima `rowmax` (from the statistics-only pass of sc_gaussian_blur).  sigma <= 1e-15
 * skips the blur.  Writes fp32 `y` and/or split planes (either may be NULL). */
int sc_blur_threshold_symmetrize(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                 const float* diag_override, double sigma, const float* rowmax,
                                 double p, double mult, int binarize, int preserve_diagonal,
                                 int sym_type, float* y, int64_t ldy, void* hi, void* lo,
                                 int64_t ldh, void* stream);

/* The same chain for a SYMMETRIC whole matrix with the blur evaluated once, on the tiles that touch
 * the upper triangle only (blur(a) is symmetric): sc_blur_upper_rowmax stores those tiles of
 * b = blur(a) (the rest of b is left untouched) and assembles rowmax[i] = max_j b[i,j] from their row
 * and column maxima (zero-fills rowmax itself); sc_threshold_symmetrize_upper then writes the whole
 * y (fp32 and/or planes; lo may be NULL when only the hi plane is wanted, i.e. for a single-MMA
 * Diffuse), mirroring every tile.  10 B (8 B without lo) of HBM traffic per matrix element instead
 * of 12, half the filter arithmetic.  Radius-4 blur only (int(4*sigma+0.5) == 4). */
int sc_blur_upper_rowmax(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                         const float* diag_override, double sigma, int zero_diagonal, float* b_out,
                         int64_t ldb, float* rowmax_out, void* stream);
int sc_threshold_symmetrize_upper(sc_context* ctx, const float* b, int64_t n, int64_t ldb,
                                  const float* rowmax, double p, double mult, int binarize,
                                  int preserve_diagonal, int sym_type, float* y, int64_t ldy,
                                  void* hi, void* lo, int64_t ldh, void* stream);

/* fp32 [n,n] -> split fp16 planes (for matrices that did not come out of the fused pass). */
int sc_split_planes(sc_context* ctx, const float* a, int64_t n, int64_t lda, void* hi, void* lo,
                    int64_t ldh, void* stream);

/* Diffuse.refine (refinement.py:232-234): s = y y^T.  rowmax (fp32) / rowsum (fp64), both or
 * neither (tcgen05 engine only), receive max_j s[i,j] and sum_j s[i,j] from the GEMM epilogue: the
 * reductions of the following RowWiseNormalize (refinement.py:243) and of the Laplacian degree
 * (laplacian.py:41) without another pass over S.  The fused maximum assumes max_j s[i,j] >= 0,
 * which holds for every y (s[i,i] = |y_i|^2). */
int sc_diffuse(sc_context* ctx, int engine, int precision, const float* y, int64_t ldy,
               const void* hi, const void* lo, int64_t ldh, int64_t n, float* s, int64_t lds,
               float* rowmax, double* rowsum, void* stream);

/* Row maxima and row sums (fp64) of an fp32 matrix: the reductions of
 * RowWiseNormalize (refinement.py:243) and of the degree (laplacian.py:41). */
int sc_row_stats(sc_context* ctx, const float* a, int64_t n, int64_t lda, double* rowmax,
                 double* rowsum, void* stream);

/* The reductions behind fallback_clusterer.check_single_cluster (fallback_clusterer.py:127-187):
 * out_host[0] = min a, [1] = sum a, [2] = sum a^2 over all n^2 entries (np.std, :154),
 * [3] = min_i a[i][i+1] (np.diag(affinity, k=1).min(), :148-150).  SYNCHRONOUS. */
int sc_affinity_stats(sc_context* ctx, const float* a, int64_t n, int64_t lda, double* out_host,
                      void* stream);

/* RowWiseNormalize.refine (refinement.py:240-245), materialised.  out may alias a. */
int sc_row_normalize(sc_context* ctx, const float* a, int64_t n, int64_t lda, float* out,
                     int64_t ldo, void* stream);

/* ---- laplacian.compute_laplacian (laplacian.py:24-60), materialised in fp32 ----------- */
int sc_laplacian(sc_context* ctx, const float* w, int64_t n, int64_t ldw, int type, double eps,
                 float* out, int64_t ldo, void* stream);

/* ---- row-block variants for the row-sharded multi-GPU pipeline ------------- */
/* Rows [row_begin, row_begin+row_count) of the affinity (utils.py:35-39) from the split planes of
 * ALL n normalised embeddings.  rowmax_offdiag_block is indexed by local row. */
int sc_affinity_cosine_block(sc_context* ctx, int precision, const void* hi, const void* lo,
                             int64_t ldh, int64_t n, int64_t d, int64_t row_begin,
                             int64_t row_count, float* a_block, int64_t lda,
                             float* rowmax_offdiag_block, void* stream);
/* Blur statistics / blur+threshold+symmetrize for global rows [row_begin, row_end) when `a` holds
 * global rows [in_row_base, in_row_base+in_rows) (owned rows + the blur halo, recomputed locally
 * rather than exchanged).  Output buffers start at global row row_begin; diag_override, rowmax and
 * rowmax_out are indexed by GLOBAL row.  (row_end-row_begin) % 32 == 0 unless row_end == n. */
int sc_gaussian_blur_rowmax_block(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                  int64_t in_row_base, int64_t in_rows, int64_t row_begin,
                                  int64_t row_end, const float* diag_override, double sigma,
                                  int zero_diagonal, float* rowmax_out, void* stream);
int sc_blur_threshold_symmetrize_block(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                       int64_t in_row_base, int64_t in_rows, int64_t row_begin,
                                       int64_t row_end, const float* diag_override, double sigma,
                                       const float* rowmax, double p, double mult, int binarize,
                                       int preserve_diagonal, int sym_type, float* y, int64_t ldy,
                                       void* hi, void* lo, int64_t ldh, void* stream);
/* Row maxima / sums of a rectangular [rows, cols] block (local rows of a sharded matrix). */
int sc_row_stats_block(sc_context* ctx, const float* a, int64_t rows, int64_t cols, int64_t lda,
                       double* rowmax, double* rowsum, void* stream);
/* dst[cols, rows] = src[rows, cols]^T (the mirrored S blocks exchanged between ranks). */
int sc_transpose(sc_context* ctx, const float* src, int64_t rows, int64_t cols, int64_t lds,
                 float* dst, int64_t ldd, void* stream);
/* c[m,n] = (a_hi+a_lo)[m,k] (b_hi+b_lo)[n,k]^T : one (row block) x (peer row block) piece of
 * Diffuse (refinement.py:232-234) in the sharded pipeline.  c_mirror (may be NULL) additionally
 * receives the transpose, c_mirror[j*ldm + i] = c[i,j], from the same epilogue: it may point into
 * a PEER GPU's row block of S (sc_ipc_open), which is how the symmetric half of the sharded
 * product reaches its owner over NVLink without a separate exchange step. */
int sc_gemm_nt_planes(sc_context* ctx, int precision, const void* a_hi, const void* a_lo,
                      int64_t lda, int64_t m, const void* b_hi, const void* b_lo, int64_t ldb,
                      int64_t n, int64_t k, float* c, int64_t ldc, float* c_mirror, int64_t ldm,
                      void* stream);

/* ---- constraint operators (constraint.py:95-164) ---------------------------------------------- */
/* Element-wise: mode 0 max(a, q) (AffinityIntegration Max, :112-113); 1 (a + q)/2 (Average,
 * :114-115); 2 q > 0 ? 1 - (1 - q)(1 - a) : (1 + q) a (the propagation's final adjustment,
 * :156-163, q = the propagated constraint matrix).  out may alias a. */
int sc_constraint_combine(sc_context* ctx, const float* a, int64_t lda, const float* q, int64_t ldq,
                          int64_t n, int mode, float* out, int64_t ldo, void* stream);
/* out = alpha * diag(row_scale) x diag(col_scale) + beta * I (fp64 scale vectors, NULL = ones): the
 * D^-1/2 A D^-1/2 normalisation (:145-147) and the I - alpha A / 2I - P steps of the Newton-Schulz
 * inverse that replaces np.linalg.inv (:151).  out may alias x. */
int sc_scale_shift(sc_context* ctx, const float* x, int64_t ldx, int64_t n, const double* row_scale,
                   const double* col_scale, double alpha, double beta, float* out, int64_t ldo,
                   void* stream);
/* c[m,n] = a[m,k] b[n,k]^T, fp32 operands, fp64 accumulation (SIMT engine): the products of the
 * propagation (:151-153) for matrices too small for a tcgen05 tile. */
int sc_gemm_nt_f32(sc_context* ctx, const float* a, int64_t lda, const float* b, int64_t ldb,
                   int64_t m, int64_t n, int64_t k, float* c, int64_t ldc, void* stream);

/* ---- peer memory (one process per GPU, NVLink/NVSwitch) ------------------------------------ */
/* CUDA IPC: export the allocation that contains dev_ptr (64-byte handle + byte offset of dev_ptr
 * inside it); open it in another process of the same box (peer access is enabled lazily); close
 * every mapping this process opened.  No reference counterpart. */
int sc_ipc_export(sc_context* ctx, const void* dev_ptr, void* handle_out, int64_t* offset_out);
int sc_ipc_open(sc_context* ctx, const void* handle, int64_t offset, void** out);
int sc_ipc_close_all(sc_context* ctx);
/* cudaMemcpyAsync(cudaMemcpyDefault) on `stream`: with a peer-mapped source the copy engines pull
 * the bytes over NVLink and no SM is involved. */
int sc_memcpy_async(sc_context* ctx, void* dst, const void* src, int64_t bytes, void* stream);

/* ---- utils.compute_sorted_eigenvectors (utils.py:44-71) ------------------------------ */
/* The matrix decomposed is M = diag(delta) + sign * diag(left) S diag(right) with S symmetric
 * fp32 and left,right > 0; delta/left/right are fp64 device vectors, NULL
 * meaning zeros/ones/ones.  M is similar to the symmetric delta + sign * c S c, c = sqrt(left*
 * right); reference eigenvectors are v = E u / |E u|, E = sqrt(left/right).
 * Eigenvalues are returned sorted (descending for LARGEST, ascending for SMALLEST) in
 * w_host[0..n_values); unit-norm eigenvectors of M for the first n_vectors of them go to
 * v_dev, fp64 row-major [n, n_vectors].  SYNCHRONOUS (returns after the stream drains).
 *
 * sc_eigh_dense: Householder tridiagonalisation in fp64 on the device, then the full spectrum
 * (n_values <= n <= 32768) by Sturm-count bisection (one thread per eigenvalue) with inverse
 * iteration for the requested eigenvectors, or by implicit QL when (nearly) all eigenvectors are
 * wanted at small n.  sc_eigh_extremal: thick-restart block Lanczos on the implicit operator
 * (fp32 S streamed from HBM once per block of vectors, fp64 vectors), n_values <= 32 << n.
 * stats_host[4] = {matrix-vector products, restarts, converged pairs, passes over S}. */
/* sc_eigh_dense with pick != NULL: once the sorted eigenvalues are in w_host the callback decides
 * how many eigenvectors are needed (the eigengap of utils.py:74-130 runs on the host) and hands
 * back the device buffer [n, count] for them; n_vectors / v_dev are then ignored. */
typedef int64_t (*sc_pick_fn)(void* user, const double* w_sorted, int64_t n_values, void** v_dev_out);
int sc_eigh_dense(sc_context* ctx, const float* s, int64_t n, int64_t lds, const double* delta,
                  const double* left, const double* right, double sign, int which,
                  int64_t n_values, int64_t n_vectors, double* w_host, double* v_dev,
                  sc_pick_fn pick, void* user, void* stream);
int sc_eigh_extremal(sc_context* ctx, const float* s, int64_t n, int64_t lds, const double* delta,
                     const double* left, const double* right, double sign, int which,
                     int64_t n_values, int64_t n_vectors, double tol, int64_t max_matvecs,
                     double* w_host, double* v_dev, int64_t* stats_host, void* stream);

/* Block size b (vectors multiplied per pass over S) the extremal solver uses for n_values pairs. */
int sc_eigh_block_size(int64_t n_values);

/* Row-sharded variant: `s_block` holds rows [row_begin, row_begin+rows) of S, and this rank is
 * slab `slab` of a partition into slabs of `slab_len` rows (row_begin == slab * slab_len).  Every
 * block product writes b = sc_eigh_block_size(n_values) partial vectors into this rank's slab of
 * y_slabs, laid out [slab][vector][slab_len] (room for (number of slabs) * b * slab_len doubles),
 * and calls gather(user, b), which must all-gather the slabs across the ranks on `stream` (b * N
 * doubles per pass over S); all other work is replicated, so every rank returns identical results.
 * delta/left/right and the outputs are full length, as in sc_eigh_extremal. */
typedef int (*sc_gather_fn)(void* user, int count);
int sc_eigh_extremal_sharded(sc_context* ctx, const float* s_block, int64_t rows, int64_t row_begin,
                             int64_t n, int64_t lds, const double* delta, const double* left,
                             const double* right, double sign, int which, int64_t n_values,
                             int64_t n_vectors, double tol, int64_t max_matvecs, double* y_slabs,
                             int slab, int64_t slab_len, sc_gather_fn gather, void* user,
                             double* w_host, double* v_dev, int64_t* stats_host, void* stream);

/* The block product of the extremal solver on its own (the O(N^2) part of what replaces
 * np.linalg.eig at utils.py:59): y[p][0..rows) = s[0..rows, 0..n) t[p][0..n) for p < b <= 16; `s` fp32
 * row-major (16-byte aligned, lds % 4 == 0), `t` / `y` fp64 vector-major with leading dimensions
 * ldt (even, t 16-byte aligned) / ldy.  One pass over `s`; products on the fp64 tensor cores,
 * column splits summed in a fixed order (bit-reproducible).  Used by sc_eigh_extremal[_sharded];
 * exported so that the kernel has its own parity test against a float64 product. */
int sc_block_product(sc_context* ctx, const float* s, int64_t rows, int64_t n, int64_t lds,
                     const double* t, int64_t ldt, int b, double* y, int64_t ldy, void* stream);

/* ---- Krylov primitives of the general (non-symmetrisable) eigen path: utils.py:59-61 is
 * np.linalg.eig + .real; sequences such as [RowWiseThreshold] alone leave a genuinely
 * non-symmetric matrix.  The Krylov-Schur recurrence is host logic
 * (spectralcluster_b200/arnoldi.py); everything of length n runs here.  Vectors are fp64, basis
 * vectors contiguous (vector q at v + q*n).
 *   matvec        y[rows] = a[rows, n] x            (fp32 matrix streamed once, fp64 accumulate)
 *   orthogonalize w -= V (V^T w) twice; h_host[count] = coefficients, *nrm2_host = |w|^2  (SYNC)
 *   scale         out = alpha * w
 *   random        w = a reproducible pseudo-random direction
 *   combine       out[p] = sum_q z_host[q*k + p] v[q], p < k <= 64                          (SYNC)
 *   columns       v_out[n, k] row-major <- the k vectors of u, each scaled to unit norm */
int sc_krylov_matvec(sc_context* ctx, const float* a, int64_t rows, int64_t n, int64_t lda,
                     const double* x, double* y, void* stream);
int sc_krylov_orthogonalize(sc_context* ctx, const double* v, int64_t n, int64_t count, double* w,
                            double* h_host, double* nrm2_host, void* stream);
int sc_krylov_scale(sc_context* ctx, const double* w, int64_t n, double alpha, double* out,
                    void* stream);
int sc_krylov_random(sc_context* ctx, double* w, int64_t n, int64_t seed, void* stream);
int sc_krylov_combine(sc_context* ctx, const double* v, int64_t n, int64_t m, const double* z_host,
                      int64_t k, double* out, void* stream);
int sc_krylov_columns(sc_context* ctx, const double* u, int64_t n, int64_t k, double* v_out,
                      void* stream);

/* Rows of e[n,k] (fp64) scaled to unit L2 norm (spectral_clusterer.py:301-305). In place. */
int sc_row_renorm(sc_context* ctx, double* e, int64_t n, int64_t k, void* stream);

/* ---- custom_distance_kmeans.run_kmeans (custom_distance_kmeans.py:13-52, 85-141) ------ */
/* e is fp64 [n,k_dim] on the device.  Seeding = scikit-learn KMeans(init="k-means++",
 * max_iter=1, random_state=0, n_init="auto") restated on the device: `first_center` and the
 * uniform draws u[(k-1)*trials] come from the caller's np.random.RandomState(0) (host RNG
 * logic only).  metric 0 = cosine, 1 = euclidean.  labels_host int64[n].  SYNCHRONOUS.
 * iters_host (may be NULL) receives the number of assignment passes. */
int sc_kmeans(sc_context* ctx, const double* e, int64_t n, int64_t k_dim, int64_t k,
              int64_t first_center, const double* u_host, int64_t trials, int metric,
              int64_t max_iter, double tol, int64_t* labels_host, int64_t* iters_host,
              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRALCLUSTER_B200_H_ */
