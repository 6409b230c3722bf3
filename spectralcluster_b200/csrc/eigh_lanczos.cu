// Extremal symmetric eigensolver for large N: thick-restart Lanczos (symmetric Krylov-Schur)
// with full re-orthogonalisation, on the implicit operator
//
//      Op x = delta .* x + sign * c .* (S (c .* x)),   c = sqrt(left*right)
//
// (the symmetrised form of the matrix the reference hands to np.linalg.eig at utils.py:59, see
// SURVEY.md A.2 and eigh_dense.cu).  Only `n_values` eigenpairs at one end are wanted
// (utils.compute_number_of_clusters reads w[0..max_clusters], utils.py:100-102), so the
// 4/3 N^3 tridiagonalisation (>= 47 s of HBM traffic at N = 65,536) is replaced by O(100s) of
// matrix-vector products, each ONE streaming pass over the fp32 matrix S:
//
//   k_symv_f32_f64  y = S t : fp32 matrix rows streamed from HBM with 128-bit loads, fp64
//                   vector staged through shared memory, fp64 accumulation.  4 B/element of
//                   HBM traffic -- the roofline of this solver is N^2 * 4 B per matvec.
//   k_proj / k_axpy_basis   classical Gram-Schmidt (twice) against the N x m basis
//   k_combine       thick restart V <- V Z
// The m x m projected problem is solved on the host by cyclic Jacobi (m <= 128).
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

namespace sc {

constexpr int SYMV_ROWS = 8;       // rows per CTA (one warp each)
constexpr int SYMV_CHUNK = 2048;   // multiple of 256   // columns of t staged in shared memory per step

__global__ void __launch_bounds__(SYMV_ROWS * 32)
k_symv_f32_f64(const float* __restrict__ s, int64_t rows, int64_t n, int64_t lds,
               const double* __restrict__ t, double* __restrict__ y) {
  // y[0..rows) = S[0..rows, 0..n) t   (rows == n on one GPU; a row block when sharded)
  __shared__ double ts[SYMV_CHUNK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * SYMV_ROWS + warp;
  const float* r = s + (row < rows ? row : 0) * lds;
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t c0 = 0; c0 < n; c0 += SYMV_CHUNK) {
    const int64_t len = (n - c0 < SYMV_CHUNK) ? (n - c0) : SYMV_CHUNK;
    __syncthreads();
    for (int64_t j = threadIdx.x; j < SYMV_CHUNK; j += blockDim.x) ts[j] = (j < len) ? t[c0 + j] : 0.0;
    __syncthreads();
    if (row < rows) {
      // lanes walk consecutive columns (coalesced 128 B per warp load, conflict-free shared
      // reads); 8 independent loads in flight per lane
      const float* rc = r + c0;
      for (int64_t j = lane; j < len; j += 256) {
        float q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = (j + 32 * u < len) ? ld_stream1(rc + j + 32 * u) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          acc0 = fma((double)q[u], ts[j + 32 * u], acc0);          // ts is zero past len
          acc1 = fma((double)q[u + 1], ts[j + 32 * (u + 1)], acc1);
        }
      }
    }
  }
  double acc = warp_sum(acc0 + acc1);
  if (row < rows && lane == 0) y[row] = acc;
}

// t = c .* x
__global__ void k_prescale(const double* __restrict__ x, const double* __restrict__ left,
                           const double* __restrict__ right, int64_t n, double* __restrict__ t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = sqrt((left ? left[i] : 1.0) * (right ? right[i] : 1.0));
  t[i] = c * x[i];
}

// w = flip * (delta .* x + sign * c .* y)
__global__ void k_postscale(const double* __restrict__ x, const double* __restrict__ y,
                            const double* __restrict__ delta, const double* __restrict__ left,
                            const double* __restrict__ right, double sign, double flip, int64_t n,
                            double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = sqrt((left ? left[i] : 1.0) * (right ? right[i] : 1.0));
  w[i] = flip * ((delta ? delta[i] * x[i] : 0.0) + sign * c * y[i]);
}

// h[j] = <V_j, w>, one CTA per basis vector (deterministic tree reduction)
__global__ void k_proj(const double* __restrict__ v, int64_t n, const double* __restrict__ w,
                       double* __restrict__ h) {
  __shared__ double red[32];
  const double* vj = v + (int64_t)blockIdx.x * n;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(vj[i], w[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) h[blockIdx.x] = s;
}

// w -= sum_j h[j] V_j
__global__ void k_axpy_basis(const double* __restrict__ v, int64_t n, int nvec,
                             const double* __restrict__ h, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = w[i];
  for (int j = 0; j < nvec; ++j) acc = fma(-h[j], v[(int64_t)j * n + i], acc);
  w[i] = acc;
}

__global__ void k_norm2(const double* __restrict__ w, int64_t n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(w[i], w[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void k_scale_into(const double* __restrict__ w, int64_t n, double alpha,
                             double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = alpha * w[i];
}

__global__ void k_random_vec(double* __restrict__ w, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  w[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// out[p] = sum_q z[q*ldz + p] V_q  for p in [p0, p0+8) : thick restart / Ritz vectors
__global__ void k_combine(const double* __restrict__ v, int64_t n, int m,
                          const double* __restrict__ z, int ldz, int n_out,
                          double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int p0 = blockIdx.y * 8;
  if (i >= n) return;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = 0; q < m; ++q) {
    const double x = v[(int64_t)q * n + i];
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p0 + p < n_out) acc[p] = fma(z[q * ldz + p0 + p], x, acc[p]);
  }
#pragma unroll
  for (int p = 0; p < 8; ++p)
    if (p0 + p < n_out) out[(int64_t)(p0 + p) * n + i] = acc[p];
}

// v_out[i, col] = E_i u_col[i] / |E u_col|  (row-major [n, n_out]); one CTA per column
__global__ void k_mapback(const double* __restrict__ u, int64_t n, int n_out,
                          const double* __restrict__ left, const double* __restrict__ right,
                          double* __restrict__ v_out) {
  __shared__ double red[32];
  const int col = blockIdx.x;
  const double* uc = u + (int64_t)col * n;
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double e = sqrt((left ? left[i] : 1.0) / (right ? right[i] : 1.0));
    const double x = e * uc[i];
    ss = fma(x, x, ss);
  }
  ss = block_sum(ss, red);
  const double inv = 1.0 / sqrt(ss);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double e = sqrt((left ? left[i] : 1.0) / (right ? right[i] : 1.0));
    v_out[i * n_out + col] = e * uc[i] * inv;
  }
}

// Cyclic Jacobi for a dense symmetric m x m matrix (row-major, destroyed); eigenvalues in w,
// eigenvectors in the columns of z.
static void jacobi_eigh(std::vector<double>& a, int m, std::vector<double>& w,
                        std::vector<double>& z) {
  z.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) z[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < m; ++i) {
      diag += a[(size_t)i * m + i] * a[(size_t)i * m + i];
      for (int j = i + 1; j < m; ++j) off += a[(size_t)i * m + j] * a[(size_t)i * m + j];
    }
    if (off <= 1e-30 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < m - 1; ++p) {
      for (int q = p + 1; q < m; ++q) {
        const double apq = a[(size_t)p * m + q];
        if (apq == 0.0) continue;
        const double app = a[(size_t)p * m + p], aqq = a[(size_t)q * m + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < m; ++k) {          // columns p, q
          const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
          a[(size_t)k * m + p] = c * akp - s * akq;
          a[(size_t)k * m + q] = s * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {          // rows p, q
          const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
          a[(size_t)p * m + k] = c * apk - s * aqk;
          a[(size_t)q * m + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < m; ++k) {
          const double zkp = z[(size_t)k * m + p], zkq = z[(size_t)k * m + q];
          z[(size_t)k * m + p] = c * zkp - s * zkq;
          z[(size_t)k * m + q] = s * zkp + c * zkq;
        }
      }
    }
  }
  w.resize(m);
  for (int i = 0; i < m; ++i) w[i] = a[(size_t)i * m + i];
}

}  // namespace sc

using namespace sc;

typedef int (*sc_gather_fn)(void* user);

// Shared implementation.  Unsharded: rows == n, row_begin == 0, y_ext == NULL.  Row-sharded: `s`
// holds the rows [row_begin, row_begin+rows) of S; every matvec writes its slice of y_ext (a
// device fp64 vector the caller owns) and calls `gather` to all-gather y_ext across the ranks
// (N doubles of traffic per matvec); everything else runs replicated on full-length vectors, so
// every rank takes identical decisions.
static int lanczos_impl(sc_context* ctx, const float* s, int64_t rows, int64_t row_begin,
                        int64_t n, int64_t lds, const double* delta, const double* left,
                        const double* right, double sign, int which, int64_t n_values,
                        int64_t n_vectors, double tol, int64_t max_matvecs, double* y_ext,
                        sc_gather_fn gather, void* user, double* w_host, double* v_dev,
                        int64_t* stats_host, void* stream) {
  SC_REQUIRE(ctx && s && w_host && n > 0, "sc_eigh_extremal: bad arguments");
  SC_REQUIRE(n_values >= 1 && n_values <= 32 && n_vectors >= 0 && n_vectors <= n_values,
             "sc_eigh_extremal: need 1 <= n_values <= 32 and n_vectors <= n_values");
  SC_REQUIRE(n_vectors == 0 || v_dev, "sc_eigh_extremal: v_dev missing");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(s) & 15) == 0 && lds % 4 == 0,
             "sc_eigh_extremal: S needs a 16-byte aligned base and lds %% 4 == 0");
  const int nev = (int)n_values;
  const int m = std::max(2 * nev + 32, 64);          // basis size
  const int keep_extra = std::max(8, nev / 2);
  SC_REQUIRE(n >= 4 * (int64_t)m, "sc_eigh_extremal: n=%lld too small for the Lanczos basis "
             "(%d); use sc_eigh_dense", (long long)n, m);
  if (tol <= 0) tol = 1e-9;
  if (max_matvecs <= 0) max_matvecs = 20000;
  cudaStream_t st = as_stream(stream);
  const double flip = (which == SC_EIG_LARGEST) ? 1.0 : -1.0;

  Scratch vb, vb2, work, small;
  SC_CUDA(vb.alloc(sizeof(double) * (size_t)(m + 1) * n, st));
  SC_CUDA(vb2.alloc(sizeof(double) * (size_t)(m + 1) * n, st));
  SC_CUDA(work.alloc(sizeof(double) * (size_t)n * 3, st));
  SC_CUDA(small.alloc(sizeof(double) * (size_t)(2 * (m + 1) + (m + 1) * 64 + 8), st));
  double* V = vb.as<double>();
  double* V2 = vb2.as<double>();
  double* t = work.as<double>();
  double* y = y_ext ? y_ext : t + n;
  double* w = t + 2 * n;
  double* h_dev = small.as<double>();          // [m+1]
  double* h2_dev = h_dev + (m + 1);            // [m+1]
  double* nrm_dev = h2_dev + (m + 1);          // [1] (+pad)
  double* z_dev = nrm_dev + 8;                 // [(m+1) x 64]

  const unsigned gn = (unsigned)((n + 255) / 256);
  std::vector<double> T((size_t)(m + 1) * (m + 1), 0.0), hh(m + 1), hh2(m + 1);
  auto Tat = [&](int i, int j) -> double& { return T[(size_t)i * (m + 1) + j]; };

  // start vector
  k_random_vec<<<gn, 256, 0, st>>>(w, n, 0x5CB200ull); sc::launched();
  k_norm2<<<1, 1024, 0, st>>>(w, n, nrm_dev); sc::launched();
  double nrm2 = 0.0;
  SC_CUDA(cudaMemcpyAsync(&nrm2, nrm_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  k_scale_into<<<gn, 256, 0, st>>>(w, n, 1.0 / std::sqrt(nrm2), V); sc::launched();
  SC_LAUNCH_CHECK();

  int j = 0;                 // basis vectors V[0..j] valid, T[0..j-1][0..j-1] valid
  int64_t matvecs = 0, restarts = 0;
  int converged = 0;
  int mm = m;                // size of the projected problem behind theta / Z
  std::vector<double> theta, Z;
  double beta_last = 0.0;
  uint64_t reseed = 1;

  // Rayleigh-Ritz on the leading sz x sz block of T; Ritz values sorted descending.  Returns how
  // many of the first nev pairs have residual |beta * Z[sz-1, p]| <= tol * max|theta|.
  auto ritz = [&](int sz, double beta) -> int {
    std::vector<double> A((size_t)sz * sz);
    for (int r = 0; r < sz; ++r)
      for (int c = 0; c < sz; ++c) A[(size_t)r * sz + c] = 0.5 * (Tat(r, c) + Tat(c, r));
    std::vector<double> wv, zv;
    jacobi_eigh(A, sz, wv, zv);
    std::vector<int> ord(sz);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return wv[a] > wv[b]; });
    theta.assign(sz, 0.0);
    Z.assign((size_t)sz * sz, 0.0);
    for (int p = 0; p < sz; ++p) {
      theta[p] = wv[ord[p]];
      for (int q = 0; q < sz; ++q) Z[(size_t)q * sz + p] = zv[(size_t)q * sz + ord[p]];
    }
    mm = sz;
    double tmax = 0.0;
    for (int p = 0; p < sz; ++p) tmax = std::max(tmax, std::fabs(theta[p]));
    int ok = 0;
    for (int p = 0; p < nev && p < sz; ++p) {
      if (std::fabs(beta * Z[(size_t)(sz - 1) * sz + p]) <= tol * tmax) ++ok;
      else break;
    }
    return ok;
  };

  bool done = false;
  while (!done) {
    for (int i = j; i < m; ++i) {
      // w = flip * Op V_i
      k_prescale<<<gn, 256, 0, st>>>(V + (size_t)i * n, left, right, n, t); sc::launched();
      k_symv_f32_f64<<<(unsigned)((rows + SYMV_ROWS - 1) / SYMV_ROWS), SYMV_ROWS * 32, 0, st>>>(
          s, rows, n, lds, t, y + row_begin); sc::launched();
      if (gather) {
        SC_LAUNCH_CHECK();
        SC_REQUIRE(gather(user) == 0, "sc_eigh_extremal_sharded: the gather callback failed");
      }
      k_postscale<<<gn, 256, 0, st>>>(V + (size_t)i * n, y, delta, left, right, sign, flip, n, w); sc::launched();
      ++matvecs;
      // classical Gram-Schmidt twice against V_0..V_i
      k_proj<<<i + 1, 256, 0, st>>>(V, n, w, h_dev); sc::launched();
      k_axpy_basis<<<gn, 256, 0, st>>>(V, n, i + 1, h_dev, w); sc::launched();
      k_proj<<<i + 1, 256, 0, st>>>(V, n, w, h2_dev); sc::launched();
      k_axpy_basis<<<gn, 256, 0, st>>>(V, n, i + 1, h2_dev, w); sc::launched();
      k_norm2<<<1, 1024, 0, st>>>(w, n, nrm_dev); sc::launched();
      SC_LAUNCH_CHECK();
      SC_CUDA(cudaMemcpyAsync(hh.data(), h_dev, sizeof(double) * (i + 1), cudaMemcpyDeviceToHost, st));
      SC_CUDA(cudaMemcpyAsync(hh2.data(), h2_dev, sizeof(double) * (i + 1), cudaMemcpyDeviceToHost, st));
      SC_CUDA(cudaMemcpyAsync(&nrm2, nrm_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
      SC_CUDA(cudaStreamSynchronize(st));
      for (int q = 0; q <= i; ++q) {
        const double v = hh[q] + hh2[q];
        Tat(q, i) = v;
        Tat(i, q) = v;
      }
      double beta = std::sqrt(nrm2);
      const double scale = std::fabs(Tat(i, i)) + 1e-300;
      if (!(beta > 1e-13 * scale)) {
        // invariant subspace: continue with a fresh direction orthogonal to the basis
        k_random_vec<<<gn, 256, 0, st>>>(w, n, 0x5CB200ull + 7919ull * reseed++); sc::launched();
        for (int pass = 0; pass < 2; ++pass) {
          k_proj<<<i + 1, 256, 0, st>>>(V, n, w, h_dev); sc::launched();
          k_axpy_basis<<<gn, 256, 0, st>>>(V, n, i + 1, h_dev, w); sc::launched();
        }
        k_norm2<<<1, 1024, 0, st>>>(w, n, nrm_dev); sc::launched();
        SC_CUDA(cudaMemcpyAsync(&nrm2, nrm_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
        SC_CUDA(cudaStreamSynchronize(st));
        k_scale_into<<<gn, 256, 0, st>>>(w, n, 1.0 / std::sqrt(nrm2), V + (size_t)(i + 1) * n); sc::launched();
        beta = 0.0;
      } else {
        k_scale_into<<<gn, 256, 0, st>>>(w, n, 1.0 / beta, V + (size_t)(i + 1) * n); sc::launched();
      }
      Tat(i + 1, i) = beta;
      Tat(i, i + 1) = beta;
      beta_last = beta;
      // early exit: test the Ritz pairs of the growing basis every 8 steps
      const int sz = i + 1;
      if (sz < m && sz >= 2 * nev && sz % 8 == 0) {
        converged = ritz(sz, beta);
        if (converged >= nev) { done = true; break; }
      }
    }
    if (done) break;
    converged = ritz(m, beta_last);
    if (converged >= nev || matvecs >= max_matvecs) break;
    // thick restart: keep the leading `keep` Ritz vectors (all converged + a buffer)
    int keep = std::min(m - 8, nev + keep_extra);
    std::vector<double> zk((size_t)m * 64, 0.0);
    SC_REQUIRE(keep <= 64, "sc_eigh_extremal: internal (keep > 64)");
    for (int q = 0; q < m; ++q)
      for (int p = 0; p < keep; ++p) zk[(size_t)q * 64 + p] = Z[(size_t)q * m + p];
    SC_CUDA(cudaMemcpyAsync(z_dev, zk.data(), sizeof(double) * (size_t)m * 64,
                            cudaMemcpyHostToDevice, st));
    k_combine<<<dim3(gn, (unsigned)((keep + 7) / 8)), 256, 0, st>>>(V, n, m, z_dev, 64, keep, V2); sc::launched();
    SC_CUDA(cudaMemcpyAsync(V2 + (size_t)keep * n, V + (size_t)m * n, sizeof(double) * (size_t)n,
                            cudaMemcpyDeviceToDevice, st));
    SC_CUDA(cudaStreamSynchronize(st));      // zk lives on the host stack frame
    std::swap(V, V2);
    std::fill(T.begin(), T.end(), 0.0);
    for (int p = 0; p < keep; ++p) {
      Tat(p, p) = theta[p];
      const double cpl = beta_last * Z[(size_t)(m - 1) * m + p];
      Tat(keep, p) = cpl;
      Tat(p, keep) = cpl;
    }
    j = keep;
    ++restarts;
  }
  for (int p = 0; p < nev; ++p) w_host[p] = flip * theta[p];
  if (n_vectors > 0) {
    std::vector<double> zk((size_t)mm * 64, 0.0);
    for (int q = 0; q < mm; ++q)
      for (int p = 0; p < (int)n_vectors; ++p) zk[(size_t)q * 64 + p] = Z[(size_t)q * mm + p];
    SC_CUDA(cudaMemcpyAsync(z_dev, zk.data(), sizeof(double) * (size_t)mm * 64,
                            cudaMemcpyHostToDevice, st));
    k_combine<<<dim3(gn, (unsigned)((n_vectors + 7) / 8)), 256, 0, st>>>(V, n, mm, z_dev, 64,
                                                                        (int)n_vectors, V2); sc::launched();
    k_mapback<<<(unsigned)n_vectors, 512, 0, st>>>(V2, n, (int)n_vectors, left, right, v_dev); sc::launched();
    SC_LAUNCH_CHECK();
    SC_CUDA(cudaStreamSynchronize(st));
  }
  if (stats_host) {
    stats_host[0] = matvecs;
    stats_host[1] = restarts;
    stats_host[2] = converged;
    stats_host[3] = m;
  }
  SC_REQUIRE(converged >= nev, "sc_eigh_extremal: only %d of %d eigenpairs converged to %g in "
             "%lld matrix-vector products", converged, nev, tol, (long long)matvecs);
  return 0;
}

extern "C" int sc_eigh_extremal(sc_context* ctx, const float* s, int64_t n, int64_t lds,
                                const double* delta, const double* left, const double* right,
                                double sign, int which, int64_t n_values, int64_t n_vectors,
                                double tol, int64_t max_matvecs, double* w_host, double* v_dev,
                                int64_t* stats_host, void* stream) {
  return lanczos_impl(ctx, s, n, 0, n, lds, delta, left, right, sign, which, n_values, n_vectors,
                      tol, max_matvecs, nullptr, nullptr, nullptr, w_host, v_dev, stats_host,
                      stream);
}

extern "C" int sc_eigh_extremal_sharded(sc_context* ctx, const float* s_block, int64_t rows,
                                        int64_t row_begin, int64_t n, int64_t lds,
                                        const double* delta, const double* left,
                                        const double* right, double sign, int which,
                                        int64_t n_values, int64_t n_vectors, double tol,
                                        int64_t max_matvecs, double* y_full, sc_gather_fn gather,
                                        void* user, double* w_host, double* v_dev,
                                        int64_t* stats_host, void* stream) {
  SC_REQUIRE(s_block && y_full && gather && rows > 0 && row_begin >= 0 && row_begin + rows <= n,
             "sc_eigh_extremal_sharded: bad arguments");
  return lanczos_impl(ctx, s_block, rows, row_begin, n, lds, delta, left, right, sign, which,
                      n_values, n_vectors, tol, max_matvecs, y_full, gather, user, w_host, v_dev,
                      stats_host, stream);
}
