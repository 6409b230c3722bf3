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
  const float* a;
  int64_t n, lda;
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
    g.out[i * g.ldo + j] = b;
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
    if (g.y) g.y[i * g.ldy + j] = yv;
    if (g.hi) {
      __half h, l;
      split_half(yv, h, l);
      g.hi[i * g.ldh + j] = h;
      g.lo[i * g.ldh + j] = l;
    }
    return yv;
  }
}

__device__ __forceinline__ float load_input(const BlurArgs& g, int64_t gr, int64_t gc) {
  if (g.diag && gr == gc) return g.diag[gr];
  return g.a[gr * g.lda + gc];
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
  const int64_t row0 = (int64_t)blockIdx.y * GTH, col0 = (int64_t)blockIdx.x * GTW;
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
    if (i < g.n && j < g.n) v = blur_epilogue<EPI>(g, i, j, acc);
    if (EPI != EPI_THRSYM && g.rowmax_out) {
      v = fmaxf(v, 0.0f);
      v = warp_max(v);
      if ((threadIdx.x & 31) == 0 && i < g.n) atomic_max_nonneg(g.rowmax_out + i, v);
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

template <int R, int EPI>
__device__ __forceinline__ void blur_tile_body(const BlurArgs& g, const float (&w)[2 * R + 1],
                                               int64_t row0, int64_t col0, float* smem,
                                               int* band_max) {
  constexpr int IW = TTW + 2 * R, IH = TTH + 2 * R;
  constexpr int MP = TTH + 1;                         // transposed pitch
  float* in = smem;                                    // [IH][IW]; reused as out [TTH][TTW+1]
  float* midT = smem + IH * IW;                        // [IW][MP]
  const bool interior = (row0 >= R) && (col0 >= R) && (row0 + TTH + R <= g.n) &&
                        (col0 + TTW + R <= g.n);
  for (int idx = threadIdx.x; idx < IH * IW; idx += TTHREADS) {
    const int r = idx / IW, c = idx - r * IW;
    int64_t gr = row0 - R + r, gc = col0 - R + c;
    if (!interior) {
      gr = reflect_index(gr, g.n);
      gc = reflect_index(gc, g.n);
    }
    in[idx] = load_input(g, gr, gc);
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
      outs[r * (TTW + 1) + c0 + o] = acc;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TTH * TTW; idx += TTHREADS) {
    const int r = idx / TTW, c = idx - r * TTW;        // a warp = a quarter row
    const int64_t i = row0 + r, j = col0 + c;
    float v = 0.0f;
    if (i < g.n && j < g.n) v = blur_epilogue<EPI>(g, i, j, outs[r * (TTW + 1) + c]);
    if (EPI != EPI_THRSYM && (band_max || g.rowmax_out)) {
      v = warp_max(fmaxf(v, 0.0f));
      if ((threadIdx.x & 31) == 0 && i < g.n) {
        if (band_max) atomicMax(band_max + r, __float_as_int(v));     // shared, non-negative
        else atomic_max_nonneg(g.rowmax_out + i, v);
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
  blur_tile_body<R, EPI>(g, w, (int64_t)blockIdx.y * TTH, (int64_t)blockIdx.x * TTW, smem,
                         nullptr);
}

// statistics pass: one CTA per 32-row band sweeps all column tiles and keeps the 32 running row
// maxima in shared memory -- no global atomics (the 2-D version issued N^2/32 of them onto N
// addresses and ran at 4% of the HBM roofline).
template <int R>
__global__ void __launch_bounds__(TTHREADS)
k_blur_band_stats(const BlurArgs g, const BlurWeights bw) {
  extern __shared__ float smem[];
  __shared__ int band_max[TTH];
  float w[2 * R + 1];
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) w[k] = bw.w[k];
  if (threadIdx.x < TTH) band_max[threadIdx.x] = 0;    // bit pattern of +0.0f
  const int64_t row0 = (int64_t)blockIdx.x * TTH;
  const int ntx = (int)((g.n + TTW - 1) / TTW);
  for (int tx = 0; tx < ntx; ++tx) {
    __syncthreads();                                   // previous tile's readers are done
    blur_tile_body<R, EPI_STATS>(g, w, row0, (int64_t)tx * TTW, smem, band_max);
  }
  __syncthreads();
  if (threadIdx.x < TTH && row0 + threadIdx.x < g.n)
    g.rowmax_out[row0 + threadIdx.x] = __int_as_float(band_max[threadIdx.x]);
}

// ------------------------------------------------------------------ no blur (sigma == 0)
template <int EPI>
__global__ void k_noblur(const BlurArgs g) {
  const int64_t i = blockIdx.x;
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
    k_noblur<EPI><<<dim3((unsigned)g.n, gy), 256, 0, st>>>(g); sc::launched();
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
    k_noblur<EPI><<<dim3((unsigned)g.n, gy), 256, 0, st>>>(g); sc::launched();
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
      kband<<<(unsigned)((g.n + TTH - 1) / TTH), TTHREADS, smem, st>>>(g, bw);
      sc::launched();
      SC_LAUNCH_CHECK();
      return 0;
    }
    auto kern = k_blur_tile<R, EPI>;
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((g.n + TTW - 1) / TTW), (unsigned)((g.n + TTH - 1) / TTH));
    kern<<<grid, TTHREADS, smem, st>>>(g, bw); sc::launched();
    SC_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = sizeof(float) * ((size_t)(GTH + 2 * radius) * (GTW + 2 * radius) +
                                       (size_t)GTH * (GTW + 2 * radius));
  SC_REQUIRE(smem <= ctx->smem_optin, "sc_gaussian_blur: tile needs %zu B of shared memory", smem);
  auto kern = k_blur_generic<EPI>;
  SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)((g.n + GTW - 1) / GTW), (unsigned)((g.n + GTH - 1) / GTH));
  kern<<<grid, GTHREADS, smem, st>>>(g, bw); sc::launched();
  SC_LAUNCH_CHECK();
  return 0;
}

}  // namespace sc

using namespace sc;

extern "C" int sc_gaussian_blur(sc_context* ctx, const float* a, int64_t n, int64_t lda,
                                const float* diag_override, double sigma, float* out,
                                int64_t ldo, float* rowmax_out, void* stream) {
  SC_REQUIRE(ctx && a && n > 0, "sc_gaussian_blur: bad arguments");
  SC_REQUIRE(out || rowmax_out, "sc_gaussian_blur: nothing to compute");
  SC_REQUIRE(n <= 65535LL * 32, "sc_gaussian_blur: n too large for the tile grid");
  BlurArgs g = {};
  g.a = a; g.n = n; g.lda = lda; g.diag = diag_override;
  g.out = out; g.ldo = ldo; g.rowmax_out = rowmax_out;
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
  return launch_blur<EPI_THRSYM>(ctx, g, sigma, as_stream(stream));
}
