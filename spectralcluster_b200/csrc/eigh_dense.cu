// Dense symmetric eigensolver, fp64, on the device: Householder tridiagonalisation + implicit QL
// with eigenvector accumulation, then back-transformation of the requested eigenvectors and the
// similarity map-back of SURVEY.md A.2.  Replaces utils.compute_sorted_eigenvectors
// (utils.py:44-71, np.linalg.eig = LAPACK geev) for the symmetrisable matrices of the hot path:
//
//      M = diag(delta) + sign * diag(left) S diag(right)        (what the reference decomposes)
//      T = diag(delta) + sign * c S c,  c = sqrt(left*right)    (symmetric, same spectrum)
//      v = E u / |E u|,  E = sqrt(left/right)                   (reference eigenvectors)
//
// This is the full-spectrum solver (O(n^3), everything resident: 2 n^2 doubles + the rotation
// log).  It serves small/medium n and the max_clusters=None scan of utils.py:100-102 that needs
// all eigenvalues; large n with a bounded cluster count goes to the Lanczos solver in
// eigh_lanczos.cu.
//
// Kernels
//   k_build_sym       T <- delta + sign * c S c                         (n^2, HBM)
//   k_hh_reflector    Householder vector of column j (one CTA)
//   k_hh_symv         p <- T22 v          (warp per row, the HBM/L2-bound half of tridiag)
//   k_hh_w            w <- tau p - (tau^2 p.v / 2) v
//   k_hh_rank2        T22 <- T22 - v w^T - w v^T
//   k_tql_rotations   implicit QL on (d, e); logs every Givens rotation  (one thread, O(n^2))
//   k_apply_rotations Z <- Z G_1 G_2 ...  (one thread per row of Z, rotations streamed)
//   k_backtransform   u <- H_0 ... H_{n-3} z, v <- E u / |E u|  (one CTA per eigenvector)
#include "common.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

namespace sc {

__global__ void k_build_sym(const float* __restrict__ s, int64_t n, int64_t lds,
                            const double* __restrict__ delta, const double* __restrict__ left,
                            const double* __restrict__ right, double sign,
                            double* __restrict__ t) {
  const int64_t i = blockIdx.x;
  const double ci = sqrt((left ? left[i] : 1.0) * (right ? right[i] : 1.0));
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < n;
       j += (int64_t)gridDim.y * blockDim.x) {
    const double cj = sqrt((left ? left[j] : 1.0) * (right ? right[j] : 1.0));
    // symmetrise the fp32 input explicitly: S is symmetric only up to rounding
    const double sij = 0.5 * ((double)s[i * lds + j] + (double)s[j * lds + i]);
    double v = sign * ci * cj * sij;
    if (i == j && delta) v += delta[i];
    t[i * n + j] = v;
  }
}

// Row j of the (fully updated, symmetric) matrix holds column j; x = T[j, j+1:n].
__global__ void k_hh_reflector(double* __restrict__ t, int64_t n, int64_t j,
                               double* __restrict__ d, double* __restrict__ e,
                               double* __restrict__ tau, double* __restrict__ vbuf) {
  __shared__ double red[32];
  __shared__ double sh[2];
  const int64_t m = n - j - 1;
  double* x = t + j * n + j + 1;
  double ss = 0.0;
  for (int64_t i = 1 + threadIdx.x; i < m; i += blockDim.x) ss += x[i] * x[i];
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    const double alpha = x[0];
    d[j] = t[j * n + j];
    if (ss == 0.0) {
      tau[j] = 0.0;
      e[j] = alpha;
      sh[0] = 0.0;      // scale (unused)
      sh[1] = 0.0;      // tau
    } else {
      const double beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      tau[j] = (beta - alpha) / beta;
      e[j] = beta;
      sh[0] = 1.0 / (alpha - beta);
      sh[1] = tau[j];
    }
  }
  __syncthreads();
  const double scale = sh[0];
  const bool trivial = (sh[1] == 0.0);
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    double v;
    if (i == 0) v = 1.0;
    else v = trivial ? 0.0 : x[i] * scale;
    vbuf[i] = v;
    x[i] = v;            // reflector kept in row j for the back-transformation
  }
}

__global__ void k_hh_symv(const double* __restrict__ t, int64_t n, int64_t j,
                          const double* __restrict__ vbuf, double* __restrict__ pbuf) {
  const int64_t m = n - j - 1;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= m) return;
  const int lane = threadIdx.x & 31;
  const double* row = t + (j + 1 + r) * n + (j + 1);
  double acc = 0.0;
  for (int64_t k = lane; k < m; k += 32) acc = fma(row[k], vbuf[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) pbuf[r] = acc;
}

__global__ void k_hh_w(const double* __restrict__ tau, int64_t j, int64_t m,
                       const double* __restrict__ vbuf, const double* __restrict__ pbuf,
                       double* __restrict__ wbuf) {
  __shared__ double red[32];
  const double tj = tau[j];
  double pv = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) pv += pbuf[i] * vbuf[i];
  pv = block_sum(pv, red);
  const double corr = 0.5 * tj * tj * pv;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) wbuf[i] = tj * pbuf[i] - corr * vbuf[i];
}

__global__ void k_hh_rank2(double* __restrict__ t, int64_t n, int64_t j,
                           const double* __restrict__ vbuf, const double* __restrict__ wbuf) {
  const int64_t m = n - j - 1;
  const int64_t r = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const double vr = vbuf[r], wr = wbuf[r];
  t[(j + 1 + r) * n + (j + 1 + c)] -= vr * wbuf[c] + wr * vbuf[c];
}

__global__ void k_hh_tail(const double* __restrict__ t, int64_t n, double* d, double* e,
                          double* tau) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (n >= 2) {
    d[n - 2] = t[(n - 2) * n + (n - 2)];
    e[n - 2] = t[(n - 2) * n + (n - 1)];
    tau[n - 2] = 0.0;
  }
  d[n - 1] = t[(n - 1) * n + (n - 1)];
  e[n - 1] = 0.0;
  tau[n - 1] = 0.0;
}

struct Sweep {
  int first_i;      // first rotation acts on columns (first_i, first_i + 1), then first_i - 1, ...
  int count;
  long long offset; // into the rotation log
};

struct QlStatus {
  long long n_rot;
  int n_sweeps;
  int status;       // 0 ok, 1 no convergence, 2 rotation log overflow, 3 sweep log overflow
};

// Implicit QL with Wilkinson shift (the classical tql2 recurrences, restated).  d: diagonal,
// e[i]: coupling between i and i+1 (e[n-1] = 0).  One thread; the rotations are logged so that
// the O(n^3) eigenvector update can run in parallel afterwards.
__global__ void k_tql_rotations(double* __restrict__ d, double* __restrict__ e, int n,
                                double* __restrict__ rc, double* __restrict__ rs,
                                long long rot_cap, Sweep* __restrict__ sweeps, int sweep_cap,
                                QlStatus* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double eps = 2.220446049250313e-16;
  long long nrot = 0;
  int nsw = 0, status = 0;
  for (int l = 0; l < n && status == 0; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= eps * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) { status = 1; break; }
        if (nsw >= sweep_cap) { status = 3; break; }
        if (nrot + (m - l) > rot_cap) { status = 2; break; }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        const long long start = nrot;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          rc[nrot] = c;
          rs[nrot] = s;
          ++nrot;
        }
        sweeps[nsw].first_i = m - 1;
        sweeps[nsw].count = (int)(nrot - start);
        sweeps[nsw].offset = start;
        ++nsw;
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  out->n_rot = nrot;
  out->n_sweeps = nsw;
  out->status = status;
}

// zt is Z transposed: zt[col * n + row]; thread = row.  Z starts as the identity.
__global__ void k_apply_rotations(double* __restrict__ zt, int64_t n,
                                  const Sweep* __restrict__ sweeps, int n_sweeps,
                                  const double* __restrict__ rc, const double* __restrict__ rs) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  for (int sw = 0; sw < n_sweeps; ++sw) {
    const int i0 = sweeps[sw].first_i, cnt = sweeps[sw].count;
    const long long off = sweeps[sw].offset;
    if (cnt == 0) continue;
    double carry = zt[(int64_t)(i0 + 1) * n + k];
    int t = 0;
    // batches of 8: the loads are independent of the carry chain, issue them together
    for (; t + 8 <= cnt; t += 8) {
      double z[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) z[u] = zt[(int64_t)(i0 - t - u) * n + k];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double c = rc[off + t + u], s = rs[off + t + u];
        zt[(int64_t)(i0 - t - u + 1) * n + k] = s * z[u] + c * carry;
        carry = c * z[u] - s * carry;
      }
    }
    for (; t < cnt; ++t) {
      const double c = rc[off + t], s = rs[off + t];
      const double zi = zt[(int64_t)(i0 - t) * n + k];
      zt[(int64_t)(i0 - t + 1) * n + k] = s * zi + c * carry;
      carry = c * zi - s * carry;
    }
    zt[(int64_t)(i0 - cnt + 1) * n + k] = carry;
  }
}

__global__ void k_identity(double* __restrict__ zt, int64_t n) {
  const int64_t i = blockIdx.x;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) zt[i * n + j] = (i == j) ? 1.0 : 0.0;
}

// One CTA per requested eigenvector: u = H_0 H_1 ... H_{n-3} z ; v = E u / |E u|.
// t holds reflector j in row j (columns j+1..n-1).  u lives in dynamic shared memory.
__global__ void k_backtransform(const double* __restrict__ t, int64_t n,
                                const double* __restrict__ tau, const double* __restrict__ zt,
                                const int* __restrict__ sel, int64_t n_sel,
                                const double* __restrict__ left, const double* __restrict__ right,
                                double* __restrict__ v_out) {
  extern __shared__ double u[];
  __shared__ double red[32];
  const int64_t col = blockIdx.x;
  const double* z = zt + (int64_t)sel[col] * n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) u[i] = z[i];
  __syncthreads();
  for (int64_t j = n - 3; j >= 0; --j) {
    const double tj = tau[j];
    if (tj == 0.0) continue;                      // uniform across the block
    const int64_t m = n - j - 1;
    const double* v = t + j * n + j + 1;
    double dot = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) dot += v[i] * u[j + 1 + i];
    dot = block_sum(dot, red);
    const double f = tj * dot;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) u[j + 1 + i] -= f * v[i];
    __syncthreads();
  }
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double ei = sqrt((left ? left[i] : 1.0) / (right ? right[i] : 1.0));
    const double x = ei * u[i];
    u[i] = x;
    ss += x * x;
  }
  ss = block_sum(ss, red);
  const double inv = 1.0 / sqrt(ss);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v_out[i * n_sel + col] = u[i] * inv;
}

}  // namespace sc

using namespace sc;

extern "C" int sc_eigh_dense(sc_context* ctx, const float* s, int64_t n, int64_t lds,
                             const double* delta, const double* left, const double* right,
                             double sign, int which, int64_t n_values, int64_t n_vectors,
                             double* w_host, double* v_dev, void* stream) {
  SC_REQUIRE(ctx && s && n > 0 && w_host, "sc_eigh_dense: bad arguments");
  SC_REQUIRE(n_values >= 0 && n_values <= n && n_vectors >= 0 && n_vectors <= n_values,
             "sc_eigh_dense: need 0 <= n_vectors <= n_values <= n");
  SC_REQUIRE(n_vectors == 0 || v_dev, "sc_eigh_dense: v_dev missing");
  SC_REQUIRE(n <= 16384, "sc_eigh_dense: n=%lld exceeds the dense solver limit (16384); use "
             "sc_eigh_extremal", (long long)n);
  cudaStream_t st = as_stream(stream);
  const size_t nn = (size_t)n * (size_t)n;
  const long long rot_cap = 2LL * n * n + 1024;
  const int sweep_cap = (int)std::min<long long>(64LL * n + 64, 2000000000LL);

  Scratch T, Zt, vec, rc, rs, sw, stat, selbuf;
  SC_CUDA(T.alloc(sizeof(double) * nn, st));
  SC_CUDA(vec.alloc(sizeof(double) * (size_t)n * 6, st));
  double* d = vec.as<double>();
  double* e = d + n;
  double* tau = e + n;
  double* vbuf = tau + n;
  double* pbuf = vbuf + n;
  double* wbuf = pbuf + n;

  const unsigned gy = (unsigned)std::min<int64_t>((n + 255) / 256, 64);
  k_build_sym<<<dim3((unsigned)n, gy), 256, 0, st>>>(s, n, lds, delta, left, right, sign,
                                                    T.as<double>()); sc::launched();
  SC_LAUNCH_CHECK();
  for (int64_t j = 0; j + 2 < n; ++j) {
    const int64_t m = n - j - 1;
    k_hh_reflector<<<1, 512, 0, st>>>(T.as<double>(), n, j, d, e, tau, vbuf); sc::launched();
    k_hh_symv<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(T.as<double>(), n, j, vbuf, pbuf); sc::launched();
    k_hh_w<<<1, 512, 0, st>>>(tau, j, m, vbuf, pbuf, wbuf); sc::launched();
    k_hh_rank2<<<dim3((unsigned)((m + 255) / 256), (unsigned)m), 256, 0, st>>>(T.as<double>(), n,
                                                                                j, vbuf, wbuf); sc::launched();
  }
  SC_LAUNCH_CHECK();
  k_hh_tail<<<1, 32, 0, st>>>(T.as<double>(), n, d, e, tau); sc::launched();
  SC_LAUNCH_CHECK();

  SC_CUDA(rc.alloc(sizeof(double) * (size_t)rot_cap, st));
  SC_CUDA(rs.alloc(sizeof(double) * (size_t)rot_cap, st));
  SC_CUDA(sw.alloc(sizeof(Sweep) * (size_t)sweep_cap, st));
  SC_CUDA(stat.alloc(sizeof(QlStatus), st));
  k_tql_rotations<<<1, 32, 0, st>>>(d, e, (int)n, rc.as<double>(), rs.as<double>(), rot_cap,
                                    sw.as<Sweep>(), sweep_cap, stat.as<QlStatus>()); sc::launched();
  SC_LAUNCH_CHECK();
  QlStatus hs;
  std::vector<double> dh((size_t)n);
  SC_CUDA(cudaMemcpyAsync(&hs, stat.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaMemcpyAsync(dh.data(), d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  SC_REQUIRE(hs.status == 0, "sc_eigh_dense: implicit QL failed (status %d: 1=no convergence, "
             "2=rotation log overflow, 3=sweep log overflow)", hs.status);

  // host: order the spectrum (argsort of +-w, utils.py:62-67)
  std::vector<int> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  if (which == SC_EIG_LARGEST)
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dh[a] > dh[b]; });
  else
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dh[a] < dh[b]; });
  for (int64_t i = 0; i < n_values; ++i) w_host[i] = dh[(size_t)order[(size_t)i]];

  if (n_vectors > 0) {
    SC_CUDA(Zt.alloc(sizeof(double) * nn, st));
    k_identity<<<(unsigned)n, 256, 0, st>>>(Zt.as<double>(), n); sc::launched();
    SC_LAUNCH_CHECK();
    if (hs.n_sweeps > 0) {
      k_apply_rotations<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(
          Zt.as<double>(), n, sw.as<Sweep>(), hs.n_sweeps, rc.as<double>(), rs.as<double>()); sc::launched();
      SC_LAUNCH_CHECK();
    }
    SC_CUDA(selbuf.alloc(sizeof(int) * (size_t)n_vectors, st));
    SC_CUDA(cudaMemcpyAsync(selbuf.p, order.data(), sizeof(int) * (size_t)n_vectors,
                            cudaMemcpyHostToDevice, st));
    const size_t smem = sizeof(double) * (size_t)n;
    SC_REQUIRE(smem <= ctx->smem_optin, "sc_eigh_dense: n too large for back-transformation");
    SC_CUDA(cudaFuncSetAttribute(k_backtransform, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    k_backtransform<<<(unsigned)n_vectors, 256, smem, st>>>(T.as<double>(), n, tau,
                                                            Zt.as<double>(), selbuf.as<int>(),
                                                            n_vectors, left, right, v_dev); sc::launched();
    SC_LAUNCH_CHECK();
    SC_CUDA(cudaStreamSynchronize(st));   // `order` must outlive the H2D copy
  }
  return 0;
}
