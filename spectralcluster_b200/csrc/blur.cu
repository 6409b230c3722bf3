// GaussianBlur (refinement.py:160-162 == scipy.ndimage.gaussian_filter(a, sigma), separable
// correlate1d along axis 0 then axis 1, mode='reflect', truncate=4.0) as a shared-memory tiled
// stencil, with three epilogues so that the ICASSP-2018 chain
//     CropDiagonal -> GaussianBlur -> RowWiseThreshold(RowMax) -> Symmetrize
// costs 12 B per matrix element of HBM traffic:
//   pass 1  EPI_STATS : read A (diag <- crop vector), blur on the fly, emit row maxima only
//   pass 2  EPI_THRSYM: read A again, blur again, apply the threshold/symmetrize rule with
//                       (m_i, m_j), write Y as split fp16 planes (4 B/element) and/or fp32
//   EPI_STORE materialises B for the generic (unfused) operator API.
// The blur is recomputed rather than stored: 18 FMAs per element are free next to 4+4 bytes.
//
// Two kernels: k_blur_tile<R> (compile-time radius, register sliding windows; R=4 is sigma=1,
// the configuration every BASELINE config uses) and k_blur_generic (run-time radius <= 64).
#include "common.cuh"

namespace sc {

constexpr int kMaxRadius = 64;

struct BlurArgs {
  const float* a;          // rows [in_row_base, ...) of the global n x n matrix
  int64_t n, lda;
  // Row block (multi-GPU row sharding; the single-GPU case is 0, n, 0, 0): outputs are produced
  // for global rows [row_begin, row_end); output buffers start at global row out_row_base.
  // Vectors (diag, m, rowmax_out) are indexed by GLOBAL row/column.
  int64_t row_begin, row_end, in_row_base, out_row_base;
  const float* diag;       // optional replacement of a[i][i]
  int radius;
  float* out;              // EPI_STORE
  int64_t ldo;
  float* rowmax_out;       // EPI_STORE (optional) / EPI_STATS
  // threshold + symmetrize epilogue
  const float* m;          // row maxima of the blurred matrix
  float p, mult;
  int binarize, preserve_diag, sym_type, stats_zero_diag;
  float* y;
  int64_t ldy;
  __half* hi;
  __half* lo;
  int64_t ldh;
};

struct BlurWeights {
  float w[2 * kMaxRadius + 1];
};

enum { EPI_STORE = 0, EPI_STATS = 1, EPI_THRSYM = 2 };

// scipy 'reflect' (half-sample symmetric, period 2n): d c b a | a b c d | d c b a
__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t n) {
  const int64_t period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return (i < n) ? i : (period - 1 - i);
}

__device__ __forceinline__ float threshold_rule(float b, float m, float p, float mult,
                                                int binarize) {
  return (b < m * p) ? b * mult : (binarize ? 1.0f : b);
}

// One output element (i, j) with blurred value b.  Returns the value that takes part in the
// row maximum (EPI_STORE / EPI_STATS).
template <int EPI>
__device__ __forceinline__ float blur_epilogue(const BlurArgs& g, int64_t i, int64_t j, float b) {
  if (EPI == EPI_STORE) {
    g.out[(i - g.out_row_base) * g.ldo + j] = b;
    return b;
  } else if (EPI == EPI_STATS) {
    return (g.stats_zero_diag && i == j) ? 0.0f : b;   // RowWiseThreshold preserve_diagonal
  } else {
    float yv;
    if (g.preserve_diag && i == j) {
      yv = 1.0f;                                        // refinement.py:208-209
    } else {
      const float t1 = threshold_rule(b, g.m[i], g.p, g.mult, g.binarize);
      const float t2 = threshold_rule(b, g.m[j], g.p, g.mult, g.binarize);
      yv = (g.sym_type == SC_SYMMETRIZE_MAX) ? fmaxf(t1, t2) : 0.5f * (t1 + t2);
    }
    if (g.y) g.y[(i - g.out_row_base) * g.ldy + j] = yv;
    if (g.hi) {
      __half h, l;
      split_half(yv, h, l);
      g.hi[(i - g.out_row_base) * g.ldh + j] = h;
      g.lo[(i - g.out_row_base) * g.ldh + j] = l;
    }
    return yv;
  }
}

__device__ __forceinline__ float load_input(const BlurArgs& g, int64_t gr, int64_t gc) {
  if (g.diag && gr == gc) return g.diag[gr];
  return g.a[(gr - g.in_row_base) * g.lda + gc];
}

// ------------------------------------------------------------------ generic radius
constexpr int GTH = 32, GTW = 64, GTHREADS = 256;

template <int EPI>
__global__ void __launch_bounds__(GTHREADS)
k_blur_generic(const BlurArgs g, const BlurWeights bw) {
  extern __shared__ float smem[];
  const int R = g.radius;
  const int IW = GTW + 2 * R, IH = GTH + 2 * R;
  float* in = smem;                 // [IH][IW]
  float* mid = smem + IH * IW;      // [GTH][IW]
  const int64_t row0 = g.row_begin + (int64_t)blockIdx.y * GTH, col0 = (int64_t)blockIdx.x * GTW;
  for (int idx = threadIdx.x; idx < IH * IW; idx += GTHREADS) {
    const int r = idx / IW, c = idx - r * IW;
    const int64_t gr = reflect_index(row0 - R + r, g.n), gc = reflect_index(col0 - R + c, g.n);
    in[idx] = load_input(g, gr, gc);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < GTH * IW; idx += GTHREADS) {   // axis 0 (down the rows)
    const int r = idx / IW, c = idx - r * IW;
    float acc = 0.0f;
    for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bw.w[k], in[(r + k) * IW + c], acc);
    mid[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < GTH * GTW; idx += GTHREADS) {  // axis 1 + epilogue
    const int r = idx / GTW, c = idx - r * GTW;                     // a warp = half a row
    float acc = 0.0f;
    for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bw.w[k], mid[r * IW + c + k], acc);
    const int64_t i = row0 + r, j = col0 + c;
    float v = 0.0f;
    if (i < g.row_end && j < g.n) v = blur_epilogue<EPI>(g, i, j, acc);
    if (EPI != EPI_THRSYM && g.rowmax_out) {
      v = fmaxf(v, 0.0f);
      v = warp_max(v);
      if ((threadIdx.x & 31) == 0 && i < g.row_end) atomic_max_nonneg(g.rowmax_out + i, v);
    }
  }
}

// ------------------------------------------------------------------ compile-time radius
// Tile 32 x 128 outputs, 256 threads.  Vertical pass: one thread per (column, 8-row strip)
// slides a register window down the column (2 LDS per output instead of 2R+1) and stores the
// result TRANSPOSED (pitch 33) so that the horizontal pass -- one thread per (row, 8-column
// strip), lanes along rows -- is bank-conflict free as well.  The outputs are staged back
// through shared memory so that the global writes (and the epilogue) are coalesced by row.
constexpr int TTH = 32, TTW = 128, TTHREADS = 256, STRIP = 8;

// Four consecutive outputs (i, j..j+3) of the threshold/symmetrize epilogue.
__device__ __forceinline__ void thrsym4(const BlurArgs& g, int64_t i, int64_t j, const float (&b)[4],
                                        bool full) {
  const float mi = g.m[i];
  float mj[4];
  if (full) {
    const float4 q = *reinterpret_cast<const float4*>(g.m + j);
    mj[0] = q.x; mj[1] = q.y; mj[2] = q.z; mj[3] = q.w;
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) mj[t] = (j + t < g.n) ? g.m[j + t] : 0.0f;
  }
  float y[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float t1 = threshold_rule(b[t], mi, g.p, g.mult, g.binarize);
    const float t2 = threshold_rule(b[t], mj[t], g.p, g.mult, g.binarize);
    y[t] = (g.sym_type == SC_SYMMETRIZE_MAX) ? fmaxf(t1, t2) : 0.5f * (t1 + t2);
    if (g.preserve_diag && i == j + t) y[t] = 1.0f;                  // refinement.py:208-209
  }
  if (full) {
    const int64_t io = i - g.out_row_base;
    if (g.y) *reinterpret_cast<float4*>(g.y + io * g.ldy + j) = make_float4(y[0], y[1], y[2], y[3]);
    if (g.hi) {
      const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(y[0] - f01.x, y[1] - f01.y);
      const __half2 l23 = __floats2half2_rn(y[2] - f23.x, y[3] - f23.y);
      uint2 hv, lv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(g.hi + io * g.ldh + j) = hv;
      *reinterpret_cast<uint2*>(g.lo + io * g.ldh + j) = lv;
    }
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (j + t >= g.n) continue;
      const int64_t io = i - g.out_row_base;
      if (g.y) g.y[io * g.ldy + j + t] = y[t];
      if (g.hi) {
        __half h, l;
        split_half(y[t], h, l);
        g.hi[io * g.ldh + j + t] = h;
        g.lo[io * g.ldh + j + t] = l;
      }
    }
  }
}

// One 32 x 128 output tile.  `rmax` (EPI_STATS only) holds this lane's running maxima of rows
// warp, warp+8, warp+16, warp+24 of the band.
template <int R, int EPI>
__device__ __forceinline__ void blur_tile_body(const BlurArgs& g, const float (&w)[2 * R + 1],
                                               int64_t row0, int64_t col0, float* smem,
                                               float (&rmax)[4]) {
  constexpr int IW = TTW + 2 * R, IH = TTH + 2 * R;
  constexpr int MP = TTH + 1;                         // transposed pitch
  constexpr int OP = TTW + 1;                         // output staging pitch
  static_assert(IW % 4 == 0 && R % 4 == 0, "vector fill needs 16-byte aligned tile origins");
  float* in = smem;                                    // [IH][IW]; reused as out [TTH][OP]
  float* midT = smem + IH * IW;                        // [IW][MP]
  const bool interior = (row0 >= R) && (col0 >= R) && (row0 + TTH + R <= g.n) &&
                        (col0 + TTW + R <= g.n);
  const bool vec_ok = interior && ((g.lda & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(g.a) & 15) == 0);
  if (vec_ok) {
    const float* src = g.a + (row0 - R - g.in_row_base) * g.lda + (col0 - R);
    for (int idx = threadIdx.x; idx < IH * (IW / 4); idx += TTHREADS) {
      const int r = idx / (IW / 4), q = idx - r * (IW / 4);
      *reinterpret_cast<float4*>(in + r * IW + 4 * q) =
          *reinterpret_cast<const float4*>(src + (int64_t)r * g.lda + 4 * q);
    }
    if (g.diag) {
      // fused CropDiagonal: patch the diagonal elements that fall inside the halo'd tile
      const int64_t lo = (row0 > col0 ? row0 : col0) - R;
      const int64_t hi = ((row0 + TTH < col0 + TTW) ? row0 + TTH : col0 + TTW) + R;
      if (lo < hi) {                                   // block-uniform
        __syncthreads();
        for (int64_t d = lo + threadIdx.x; d < hi; d += TTHREADS)
          in[(d - (row0 - R)) * IW + (d - (col0 - R))] = g.diag[d];
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < IH * IW; idx += TTHREADS) {
      const int r = idx / IW, c = idx - r * IW;
      int64_t gr = row0 - R + r, gc = col0 - R + c;
      if (!interior) {
        gr = reflect_index(gr, g.n);
        gc = reflect_index(gc, g.n);
      }
      in[idx] = load_input(g, gr, gc);
    }
  }
  __syncthreads();
  // vertical: items = IW columns x (TTH/STRIP) strips
  for (int item = threadIdx.x; item < IW * (TTH / STRIP); item += TTHREADS) {
    const int c = item % IW, r0 = (item / IW) * STRIP;
    float win[STRIP + 2 * R];
#pragma unroll
    for (int k = 0; k < STRIP + 2 * R; ++k) win[k] = in[(r0 + k) * IW + c];
#pragma unroll
    for (int o = 0; o < STRIP; ++o) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) acc = fmaf(w[k], win[o + k], acc);
      midT[c * MP + r0 + o] = acc;
    }
  }
  __syncthreads();
  // horizontal: items = TTH rows x (TTW/STRIP) strips; lanes run along rows.  All reads of
  // `in` finished before the barrier above, so it can take the outputs.
  float* outs = in;
  for (int item = threadIdx.x; item < TTH * (TTW / STRIP); item += TTHREADS) {
    const int r = item % TTH, c0 = (item / TTH) * STRIP;
    float win[STRIP + 2 * R];
#pragma unroll
    for (int k = 0; k < STRIP + 2 * R; ++k) win[k] = midT[(c0 + k) * MP + r];
#pragma unroll
    for (int o = 0; o < STRIP; ++o) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) acc = fmaf(w[k], win[o + k], acc);
      outs[r * OP + c0 + o] = acc;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool cols_full = (col0 + TTW <= g.n);
  if (EPI == EPI_STATS) {
    // warp -> rows warp, warp+8, ...; lanes stride the 128 columns; maxima stay in registers
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = warp + 8 * q;
      const int64_t i = row0 + r;
      if (i >= g.row_end) continue;
      float v = rmax[q];
#pragma unroll
      for (int u = 0; u < TTW / 32; ++u) {
        const int c = lane + 32 * u;
        const int64_t j = col0 + c;
        float b = outs[r * OP + c];
        if (!cols_full && j >= g.n) b = 0.0f;
        if (g.stats_zero_diag && i == j) b = 0.0f;     // RowWiseThreshold preserve_diagonal
        v = fmaxf(v, b);
      }
      rmax[q] = v;
    }
  } else {
    // a warp = one row: lane handles 4 consecutive columns (512 B fp32 / 256 B fp16 per warp)
    for (int r = warp; r < TTH; r += TTHREADS / 32) {
      const int64_t i = row0 + r;
      if (i >= g.row_end) continue;
      const int c = lane * 4;
      const int64_t j = col0 + c;
      if (j >= g.n) continue;
      float b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) b[t] = outs[r * OP + c + t];
      const bool full = (j + 3 < g.n);
      if (EPI == EPI_THRSYM) {
        thrsym4(g, i, j, b, full);
      } else {   // EPI_STORE
        if (full) {
          *reinterpret_cast<float4*>(g.out + (i - g.out_row_base) * g.ldo + j) = make_float4(b[0], b[1], b[2], b[3]);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (j + t < g.n) g.out[(i - g.out_row_base) * g.ldo + j + t] = b[t];
        }
        if (g.rowmax_out) {
          float v = 0.0f;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (j + t < g.n) v = fmaxf(v, b[t]);
          rmax[0] = v;   // scratch
        }
      }
      if (EPI == EPI_STORE && g.rowmax_out) {
        // (all lanes of the row reach here together only when none `continue`d; do the
        //  reduction with the active mask)
        const unsigned mask = __activemask();
        float v = rmax[0];
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o));
        if (lane == 0) atomic_max_nonneg(g.rowmax_out + i, v);
      }
    }
  }
}

// one tile per CTA (EPI_STORE / EPI_THRSYM)
template <int R, int EPI>
__global__ void __launch_bounds__(TTHREADS)
k_blur_tile(const BlurArgs g, const BlurWeights bw) {
  extern __shared__ float smem[];
  float w[2 * R + 1];
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) w[k] = bw.w[k];
  float scratch[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  blur_tile_body<R, EPI>(g, w, g.row_begin + (int64_t)blockIdx.y * TTH, (int64_t)blockIdx.x * TTW, smem,
                         scratch);
}

// statistics pass: one CTA per 32-row band sweeps all column tiles; each lane keeps running
// maxima of its 4 rows in registers -- no atomics at all (the first version issued N^2/32 global
// atomics onto N addresses and ran at 4% of the HBM roofline).
template <int R>
__global__ void __launch_bounds__(TTHREADS)
k_blur_band_stats(const BlurArgs g, const BlurWeights bw) {
  extern __shared__ float smem[];
  float w[2 * R + 1];
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) w[k] = bw.w[k];
  float rmax[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int64_t row0 = g.row_begin + (int64_t)blockIdx.x * TTH;
  const int ntx = (int)((g.n + TTW - 1) / TTW);
  for (int tx = 0; tx < ntx; ++tx) {
    __syncthreads();                                   // previous tile's readers are done
    blur_tile_body<R, EPI_STATS>(g, w, row0, (int64_t)tx * TTW, smem, rmax);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float v = warp_max(rmax[q]);
    const int64_t i = row0 + warp + 8 * q;
    if (lane == 0 && i < g.row_end) g.rowmax_out[i] = v;
  }
}

// ------------------------------------------------------------------ no blur (sigma == 0)
template <int EPI>
__global__ void k_noblur(const BlurArgs g) {
  const int64_t i = g.row_begin + blockIdx.x;
  float vmax = 0.0f;
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < g.n;
       j += (int64_t)gridDim.y * blockDim.x) {
    const float v = blur_epilogue<EPI>(g, i, j, load_input(g, i, j));
    vmax = fmaxf(vmax, v);
  }
  if (EPI != EPI_THRSYM && g.rowmax_out) {
    vmax = warp_max(fmaxf(vmax, 0.0f));
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(g.rowmax_out + i, vmax);
  }
}

static int make_weights(double sigma, BlurWeights& bw, int& radius) {
  // scipy _gaussian_kernel1d: radius = int(truncate*sigma + 0.5), exp(-x^2/(2 sigma^2)) / sum
  radius = (int)(4.0 * sigma + 0.5);
  if (radius > kMaxRadius) return 1;
  double w[2 * kMaxRadius + 1], sum = 0.0;
  for (int k = -radius; k <= radius; ++k) {
    w[k + radius] = std::exp(-0.5 / (sigma * sigma) * (double)k * (double)k);
    sum += w[k + radius];
  }
  for (int k = 0; k <= 2 * radius; ++k) bw.w[k] = (float)(w[k] / sum);
  return 0;
}

template <int EPI>
static int launch_blur(const sc_context* ctx, BlurArgs& g, double sigma, cudaStream_t st) {
  if (sigma <= 1e-15) {     // scipy: "if sigma > 1e-15 ... else output[...] = input[...]"
    const unsigned gy = (unsigned)std::min<int64_t>((g.n + 1023) / 1024, 64);
    k_noblur<EPI><<<dim3((unsigned)(g.row_end - g.row_begin), gy), 256, 0, st>>>(g); sc::launched();
    SC_LAUNCH_CHECK();
    return 0;
  }
  BlurWeights bw;
  int radius;
  SC_REQUIRE(make_weights(sigma, bw, radius) == 0,
             "sc_gaussian_blur: sigma %g needs radius > %d", sigma, kMaxRadius);
  g.radius = radius;
  if (radius == 0) {        // a single tap of weight 1
    const unsigned gy = (unsigned)std::min<int64_t>((g.n + 1023) / 1024, 64);
    k_noblur<EPI><<<dim3((unsigned)(g.row_end - g.row_begin), gy), 256, 0, st>>>(g); sc::launched();
    SC_LAUNCH_CHECK();
    return 0;
  }
  if (radius == 4) {
    constexpr int R = 4;
    const size_t smem = sizeof(float) * ((TTH + 2 * R) * (TTW + 2 * R) + (TTW + 2 * R) * (TTH + 1));
    static_assert((TTH + 2 * R) * (TTW + 2 * R) >= TTH * (TTW + 1), "output staging must fit");
    if (EPI == EPI_STATS) {
      auto kband = k_blur_band_stats<R>;
      SC_CUDA(cudaFuncSetAttribute(kband, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kband<<<(unsigned)((g.row_end - g.row_begin + TTH - 1) / TTH), TTHREADS, smem, st>>>(g, bw);
      sc::launched();
      SC_LAUNCH_CHECK();
      return 0;
    }
    auto kern = k_blur_tile<R, EPI>;
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((g.n + TTW - 1) / TTW),
                    (unsigned)((g.row_end - g.row_begin + TTH - 1) / TTH));
    kern<<<grid, TTHREADS, smem, st>>>(g, bw); sc::launched();
    SC_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = sizeof(float) * ((size_t)(GTH + 2 * radius) * (GTW + 2 * radius) +
                                       (size_t)GTH * (GTW + 2 * radius));
  SC_REQUIRE(smem <= ctx->smem_optin, "sc_gaussian_blur: tile needs %zu B of shared memory", smem);
  auto kern = k_blur_generic<EPI>;
  SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)((g.n + GTW - 1) / GTW),
                  (unsigned)((g.row_end - g.row_begin + GTH - 1) / GTH));
  kern<<<grid, GTHREADS, smem, st>>>(g, bw); sc::launched();
  SC_LAUNCH_CHECK();
  return 0;
}

}  // namespace sc

using namespace sc;

static void whole_matrix(BlurArgs& g) {
  g.row_begin = 0; g.row_end = g.n; g.in_row_base = 0; g.out_row_base = 0;
}

static int check_block(int64_t n, int64_t in_row_base, int64_t in_rows, int64_t row_begin,
                       int64_t row_end, int radius) {
  SC_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= n, "blur block: bad row range");
  SC_REQUIRE((row_end - row_begin) % 32 == 0 || row_end == n,
             "blur block: the row count must be a multiple of 32 unless the block ends the matrix");
  // every input row the block can touch (after reflection at the matrix edges) must be resident
  const int64_t need_lo = row_begin - radius < 0 ? 0 : row_begin - radius;
  const int64_t need_hi = row_end + radius > n ? n : row_end + radius;
  SC_REQUIRE(in_row_base <= need_lo && in_row_base + in_rows >= need_hi,
             "blur block: input rows [%lld, %lld) do not cover the halo [%lld, %lld)",
             (long long)in_row_base, (long long)(in_row_base + in_rows), (long long)need_lo,
             (long long)need_hi);
  return 0;
}

extern "C" int sc_gaussian_blur(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                const float* diag_override, double sigma, float* out,
                                int64_t ldo, float* rowmax_out, void* stream) {
  SC_REQUIRE(ctx && a && n > 0, "sc_gaussian_blur: bad arguments");
  SC_REQUIRE(out || rowmax_out, "sc_gaussian_blur: nothing to compute");
  SC_REQUIRE(n <= 65535LL * 32, "sc_gaussian_blur: n too large for the tile grid");
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.out = out; g.ldo = ldo; g.rowmax_out = rowmax_out;
  whole_matrix(g);
  if (out) return launch_blur<EPI_STORE>(ctx, g, sigma, as_stream(stream));
  return launch_blur<EPI_STATS>(ctx, g, sigma, as_stream(stream));
}

// statistics pass with the diagonal read as zero (RowWiseThreshold preserve_diagonal)
extern "C" int sc_gaussian_blur_rowmax(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                       const float* diag_override, double sigma,
                                       int zero_diagonal, float* rowmax_out, void* stream) {
  SC_REQUIRE(ctx && a && rowmax_out && n > 0, "sc_gaussian_blur_rowmax: bad arguments");
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.rowmax_out = rowmax_out; g.stats_zero_diag = zero_diagonal;
  whole_matrix(g);
  return launch_blur<EPI_STATS>(ctx, g, sigma, as_stream(stream));
}

extern "C" int sc_blur_threshold_symmetrize(sc_context* ctx, const float* a, int64_t n,
                                            int64_t lda, const float* diag_override,
                                            double sigma, const float* rowmax, double p,
                                            double mult, int binarize, int preserve_diagonal,
                                            int sym_type, float* y, int64_t ldy, void* hi,
                                            void* lo, int64_t ldh, void* stream) {
  SC_REQUIRE(ctx && a && rowmax && n > 0, "sc_blur_threshold_symmetrize: bad arguments");
  SC_REQUIRE(y || (hi && lo), "sc_blur_threshold_symmetrize: no output given");
  SC_REQUIRE((hi == nullptr) == (lo == nullptr), "hi/lo must come together");
  SC_REQUIRE(sym_type == SC_SYMMETRIZE_MAX || sym_type == SC_SYMMETRIZE_AVERAGE,
             "Unsupported symmetrize_type.");
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.m = rowmax; g.p = (float)p; g.mult = (float)mult; g.binarize = binarize;
  g.preserve_diag = preserve_diagonal; g.sym_type = sym_type;
  g.y = y; g.ldy = ldy; g.hi = (__half*)hi; g.lo = (__half*)lo; g.ldh = ldh;
  whole_matrix(g);
  return launch_blur<EPI_THRSYM>(ctx, g, sigma, as_stream(stream));
}

// ---- row-block variants (multi-GPU row sharding): `a` holds global rows
// [in_row_base, in_row_base + in_rows) -- the owned rows plus the blur halo, recomputed locally
// from the embeddings instead of exchanged; outputs cover global rows [row_begin, row_end) and are
// written to buffers whose row 0 is global row `row_begin`.  diag_override / rowmax vectors are
// indexed by global row.
extern "C" int sc_gaussian_blur_rowmax_block(sc_context* ctx, const float* a, int64_t n,
                                             int64_t lda, int64_t in_row_base, int64_t in_rows,
                                             int64_t row_begin, int64_t row_end,
                                             const float* diag_override, double sigma,
                                             int zero_diagonal, float* rowmax_out,
                                             void* stream) {
  SC_REQUIRE(ctx && a && rowmax_out && n > 0, "sc_gaussian_blur_rowmax_block: bad arguments");
  const int radius = sigma > 1e-15 ? (int)(4.0 * sigma + 0.5) : 0;
  if (int rc = check_block(n, in_row_base, in_rows, row_begin, row_end, radius)) return rc;
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.rowmax_out = rowmax_out; g.stats_zero_diag = zero_diagonal;
  g.row_begin = row_begin; g.row_end = row_end; g.in_row_base = in_row_base;
  g.out_row_base = row_begin;
  return launch_blur<EPI_STATS>(ctx, g, sigma, as_stream(stream));
}

extern "C" int sc_blur_threshold_symmetrize_block(
    sc_context* ctx, const float* a, int64_t n, int64_t lda, int64_t in_row_base, int64_t in_rows,
    int64_t row_begin, int64_t row_end, const float* diag_override, double sigma,
    const float* rowmax, double p, double mult, int binarize, int preserve_diagonal, int sym_type,
    float* y, int64_t ldy, void* hi, void* lo, int64_t ldh, void* stream) {
  SC_REQUIRE(ctx && a && rowmax && n > 0, "sc_blur_threshold_symmetrize_block: bad arguments");
  SC_REQUIRE(y || (hi && lo), "sc_blur_threshold_symmetrize_block: no output given");
  SC_REQUIRE((hi == nullptr) == (lo == nullptr), "hi/lo must come together");
  SC_REQUIRE(sym_type == SC_SYMMETRIZE_MAX || sym_type == SC_SYMMETRIZE_AVERAGE,
             "Unsupported symmetrize_type.");
  const int radius = sigma > 1e-15 ? (int)(4.0 * sigma + 0.5) : 0;
  if (int rc = check_block(n, in_row_base, in_rows, row_begin, row_end, radius)) return rc;
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.m = rowmax; g.p = (float)p; g.mult = (float)mult; g.binarize = binarize;
  g.preserve_diag = preserve_diagonal; g.sym_type = sym_type;
  g.y = y; g.ldy = ldy; g.hi = (__half*)hi; g.lo = (__half*)lo; g.ldh = ldh;
  g.row_begin = row_begin; g.row_end = row_end; g.in_row_base = in_row_base;
  g.out_row_base = row_begin;
  return launch_blur<EPI_THRSYM>(ctx, g, sigma, as_stream(stream));
}
