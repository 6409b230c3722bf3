// Context, error reporting and the GEMM-backed entry points of the C ABI.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>

namespace sc {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void launched(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// gemm_simt.cu
int gemm_nt_simt(int epi, const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                 int64_t N, int64_t K, float* C, int64_t ldc, float* rowmax_offdiag,
                 cudaStream_t st);
// gemm_tcgen05.cu
int gemm_nt_tcgen05(const sc_context* ctx, int epi, int precision, const __half* a_hi,
                    const __half* a_lo, int64_t lda, const __half* b_hi, const __half* b_lo,
                    int64_t ldb, int64_t M, int64_t N, int64_t K, float* C, int64_t ldc,
                    float* rowmax_offdiag, bool symmetric, int diag_shift, float* stat_rowmax,
                    double* stat_rowsum, float* mirror, int64_t ldm, cudaStream_t st);

}  // namespace sc

using namespace sc;

extern "C" int sc_abi_version(void) { return SC_ABI_VERSION; }

extern "C" const char* sc_last_error(void) { return g_error; }

extern "C" long long sc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int sc_context_create(int device, sc_context** out) {
  SC_REQUIRE(out != nullptr, "sc_context_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("sc_context_create: no CUDA device is visible (%s). spectralcluster_b200 has no "
              "CPU fallback.", cudaGetErrorString(e));
    return 1;
  }
  SC_REQUIRE(device >= 0 && device < count, "sc_context_create: device %d out of range [0,%d)",
             device, count);
  SC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SC_CUDA(cudaGetDeviceProperties(&prop, device));
  SC_REQUIRE(prop.major == 10, "sc_context_create: device %d is sm_%d%d; this library is built "
             "for sm_100a (B200) only", device, prop.major, prop.minor);
  sc_context* ctx = new sc_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  ctx->gemm_sm_limit = 0;
  ctx->gemm_pace = nullptr;
  if (!getenv("SCB_NO_GEMM_PACING") &&
      cudaMalloc(&ctx->gemm_pace, sizeof(unsigned int) * SC_GEMM_PACE_SLOTS) != cudaSuccess)
    ctx->gemm_pace = nullptr;
  // keep stream-ordered scratch inside the pool between calls (no trim at every sync)
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = ctx;
  return 0;
}

extern "C" int sc_context_destroy(sc_context* ctx) {
  if (ctx && ctx->gemm_pace) cudaFree(ctx->gemm_pace);
  delete ctx;
  return 0;
}

extern "C" int sc_context_sm_count(const sc_context* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int sc_context_set_gemm_sm_limit(sc_context* ctx, int sms) {
  SC_REQUIRE(ctx && sms >= 0, "sc_context_set_gemm_sm_limit: bad arguments");
  ctx->gemm_sm_limit = sms;
  return 0;
}

extern "C" int sc_affinity_cosine(sc_context* ctx, int engine, int precision, const float* xn,
                                  int64_t ldxn, const void* hi, const void* lo, int64_t ldh,
                                  int64_t n, int64_t d, float* a, int64_t lda,
                                  float* rowmax_offdiag, void* stream) {
  SC_REQUIRE(ctx && a && n > 0 && d > 0, "sc_affinity_cosine: bad arguments");
  if (engine == SC_GEMM_SIMT_F64ACC) {
    SC_REQUIRE(xn, "sc_affinity_cosine: the SIMT engine needs the fp32 operand");
    return gemm_nt_simt(1, xn, ldxn, xn, ldxn, n, n, d, a, lda, rowmax_offdiag,
                        as_stream(stream));
  }
  SC_REQUIRE(engine == SC_GEMM_TCGEN05, "sc_affinity_cosine: unknown engine %d", engine);
  SC_REQUIRE(hi && (lo || precision == SC_GEMM_SINGLE),
             "sc_affinity_cosine: the tcgen05 engine needs the split fp16 planes");
  return gemm_nt_tcgen05(ctx, 1, precision, (const __half*)hi, (const __half*)lo, ldh,
                         (const __half*)hi, (const __half*)lo, ldh, n, n, d, a, lda,
                         rowmax_offdiag, /*symmetric=*/true, /*diag_shift=*/0, nullptr, nullptr,
                         nullptr, 0, as_stream(stream));
}

extern "C" int sc_diffuse(sc_context* ctx, int engine, int precision, const float* y,
                          int64_t ldy, const void* hi, const void* lo, int64_t ldh, int64_t n,
                          float* s, int64_t lds, float* rowmax, double* rowsum, void* stream) {
  SC_REQUIRE(ctx && s && n > 0, "sc_diffuse: bad arguments");
  SC_REQUIRE((rowmax == nullptr) == (rowsum == nullptr), "sc_diffuse: rowmax/rowsum come together");
  if (engine == SC_GEMM_SIMT_F64ACC) {
    SC_REQUIRE(y, "sc_diffuse: the SIMT engine needs the fp32 operand");
    SC_REQUIRE(!rowmax, "sc_diffuse: the SIMT engine has no fused row statistics");
    return gemm_nt_simt(0, y, ldy, y, ldy, n, n, n, s, lds, nullptr, as_stream(stream));
  }
  SC_REQUIRE(engine == SC_GEMM_TCGEN05, "sc_diffuse: unknown engine %d", engine);
  SC_REQUIRE(hi && (lo || precision == SC_GEMM_SINGLE),
             "sc_diffuse: the tcgen05 engine needs the split fp16 planes");
  if (rowmax) {
    SC_CUDA(cudaMemsetAsync(rowmax, 0, sizeof(float) * n, as_stream(stream)));
    SC_CUDA(cudaMemsetAsync(rowsum, 0, sizeof(double) * n, as_stream(stream)));
  }
  return gemm_nt_tcgen05(ctx, 0, precision, (const __half*)hi, (const __half*)lo, ldh,
                         (const __half*)hi, (const __half*)lo, ldh, n, n, n, s, lds, nullptr,
                         /*symmetric=*/true, /*diag_shift=*/0, rowmax, rowsum, nullptr, 0,
                         as_stream(stream));
}

// ---- row-block / general variants used by the row-sharded multi-GPU pipeline ----------------
extern "C" int sc_affinity_cosine_block(sc_context* ctx, int precision, const void* hi,
                                        const void* lo, int64_t ldh, int64_t n, int64_t d,
                                        int64_t row_begin, int64_t row_count, float* a_block,
                                        int64_t lda, float* rowmax_offdiag_block, void* stream) {
  SC_REQUIRE(ctx && hi && a_block && n > 0 && d > 0 && row_begin >= 0 && row_count > 0 &&
             row_begin + row_count <= n, "sc_affinity_cosine_block: bad arguments");
  SC_REQUIRE(lo || precision == SC_GEMM_SINGLE, "sc_affinity_cosine_block: lo plane missing");
  const __half* h = (const __half*)hi;
  const __half* l = (const __half*)lo;
  return gemm_nt_tcgen05(ctx, 1, precision, h + row_begin * ldh, l ? l + row_begin * ldh : nullptr,
                         ldh, h, l, ldh, row_count, n, d, a_block, lda, rowmax_offdiag_block,
                         /*symmetric=*/false, /*diag_shift=*/(int)row_begin, nullptr, nullptr,
                         nullptr, 0, as_stream(stream));
}

extern "C" int sc_gemm_nt_planes(sc_context* ctx, int precision, const void* a_hi,
                                 const void* a_lo, int64_t lda, int64_t m, const void* b_hi,
                                 const void* b_lo, int64_t ldb, int64_t n, int64_t k, float* c,
                                 int64_t ldc, float* c_mirror, int64_t ldm, void* stream) {
  SC_REQUIRE(ctx && a_hi && b_hi && c && m > 0 && n > 0 && k > 0, "sc_gemm_nt_planes: bad arguments");
  SC_REQUIRE((a_lo && b_lo) || precision == SC_GEMM_SINGLE || (a_lo && precision == SC_GEMM_SPLIT2),
             "sc_gemm_nt_planes: lo planes missing");
  SC_REQUIRE(c_mirror == nullptr || ldm >= m, "sc_gemm_nt_planes: bad mirror leading dimension");
  return gemm_nt_tcgen05(ctx, 0, precision, (const __half*)a_hi, (const __half*)a_lo, lda,
                         (const __half*)b_hi, (const __half*)b_lo, ldb, m, n, k, c, ldc, nullptr,
                         /*symmetric=*/true, /*diag_shift=*/0, nullptr, nullptr, c_mirror, ldm,
                         as_stream(stream));
}

// c[m,n] = a[m,k] b[n,k]^T with fp32 operands on the SIMT engine (fp64 accumulation): the small-
// matrix twin of sc_gemm_nt_planes for the constraint propagation products (constraint.py:147-153).
extern "C" int sc_gemm_nt_f32(sc_context* ctx, const float* a, int64_t lda, const float* b,
                              int64_t ldb, int64_t m, int64_t n, int64_t k, float* c, int64_t ldc,
                              void* stream) {
  SC_REQUIRE(ctx && a && b && c && m > 0 && n > 0 && k > 0, "sc_gemm_nt_f32: bad arguments");
  return gemm_nt_simt(0, a, lda, b, ldb, m, n, k, c, ldc, nullptr, as_stream(stream));
}
