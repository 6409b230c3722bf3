// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (both operands K-major),
// the only two dense contractions of the hot path:
//   utils.py:35-39        affinity = (Xn Xn^T + 1)/2        K = d      (EPI_AFFINITY)
//   refinement.py:232-234 Diffuse  = Y Y^T                  K = N      (EPI_PLAIN)
//
// Operands are "split fp16 planes" (value = hi + lo).  Per 64-wide K block the MMA warp issues
//     D += Ahi*Bhi ; D += Ahi*Blo ; D += Alo*Bhi          (tcgen05.mma kind::f16, fp32 accum)
// which reproduces the fp32 product to ~2^-22 relative (the dropped lo*lo term is 2^-22) at the
// fp16 tensor-pipe rate.  SC_GEMM_SPLIT2 drops the B_lo plane (D += Alo*Bhi ; D += Ahi*Bhi: the
// missing Ahi*Blo term is a zero-mean 2^-12 relative perturbation per product that averages down
// with sqrt(K)); SC_GEMM_SINGLE issues only hi*hi (2^-11 on both sides).  profiles/ holds the
// measured eigenvalue / label evidence for each mode.
//
// Structure (one CTA per SM, persistent, 384 threads = 3 warpgroups):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) of the 4 (or 2) planes of a
//            K block into a ring of stages, completion on an mbarrier (expect_tx); paced against
//            the other CTAs' producers through a global checkpoint counter (see below)
//   warp 1   MMA issuer: one lane issues tcgen05.mma M=128 N=256 K=16 from shared-memory
//            descriptors into one of two TMEM chain buffers (128 lanes x 256 fp32 columns each),
//            frees the stage with tcgen05.commit; a chain is only 2 K-blocks long
//   warp 2   TMEM allocator (512 columns)
//   warps 4-11 epilogue (setmaxnreg 232): tcgen05.ld 32x32b.x32 of every finished chain, added
//            round-to-nearest into a register-resident 128 x 256 fp32 tile (tcgen05 accumulates
//            with truncation -- the two-level scheme keeps the bias below 1e-6); at the end of the
//            tile: fused (x+1)/2 + off-diagonal row maximum (affinity), 128-bit stores, and the
//            mirrored (transposed, coalesced) stores of the symmetric variant
// Tiles are rasterised in groups of 16 M-blocks so that a wave of 148 CTAs shares 16 A-panels
// and ~9 B-panels through L2; the symmetric variant (C = Y Y^T) computes only the tiles that touch
// the upper triangle.
#include "common.cuh"

#include <cuda.h>
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace sc {

constexpr int BM = 128, BN = 256, BK = 64;          // BK fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_PLANE_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_PLANE_BYTES = BN * BK * 2;           // 32 KB
constexpr int GROUP_M = 16;
constexpr int PACE_KB = 64;       // K-blocks between pacing checkpoints of the producers
constexpr int PACE_SPINS = 400;   // x 100 ns: bounded wait (a hint, never a dependency)
constexpr int TMEM_COLS = 512;
constexpr uint32_t SPIN_LIMIT = 1u << 27;            // trap instead of hanging the GPU

enum { TC_EPI_PLAIN = 0, TC_EPI_AFFINITY = 1 };

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// ---- CTA pair (cta_group::2): the two CTAs of a cluster run ONE 256 x 256 MMA tile.  The leader
// (cluster rank 0) issues the MMAs; barriers that the leader waits on live in the leader's shared
// memory and are addressed from the peer by clearing the CTA-rank bit of the shared::cluster
// address (the pair occupies ranks 0/1 of the cluster, bit 24 of the window address).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];"
               ::"r"(bar & PEER_BIT_MASK) : "memory");
}
// the load lands in THIS CTA's shared memory, its bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B
// between 8-row groups | version [46,48) = 1 | layout [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=F16 [7,10)=0,
// B=F16 [10,13)=0, both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TileCoord {
  int m_blk, n_blk;
};

// Rasterisation.  Full problem: groups of GROUP_M M-blocks sweep all N-blocks.  Symmetric
// problem (C = Y Y^T): only tiles that touch the upper triangle are computed, i.e.
// 128 m_blk < 256 (n_blk + 1); the rest is filled by mirrored stores.  group_start[g] is the
// first tile index of group g (host-computed with the same counting rule).
constexpr int MAX_GROUPS = 160;
struct TileTable {
  int group_start[MAX_GROUPS + 1];
  int num_groups;
  int num_tiles;
};

__host__ __device__ inline int sym_valid_mblocks(int g, int n_blk, int tiles_m) {
  // m blocks of group g that are valid for column block n_blk: m_blk <= 2 n_blk + 1
  int lim = 2 * n_blk + 2 - g * GROUP_M;
  const int avail = tiles_m - g * GROUP_M;
  if (lim > GROUP_M) lim = GROUP_M;
  if (lim > avail) lim = avail;
  return lim < 0 ? 0 : lim;
}

template <bool SYM>
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_m, int tiles_n,
                                                const TileTable& tab) {
  TileCoord t;
  if (!SYM) {
    const int group = GROUP_M * tiles_n;
    const int g = tile / group;
    const int first_m = g * GROUP_M;
    const int gm = min(GROUP_M, tiles_m - first_m);
    const int r = tile - g * group;
    t.m_blk = first_m + r % gm;
    t.n_blk = r / gm;
  } else {
    int g = 0;
    while (g + 1 < tab.num_groups && tab.group_start[g + 1] <= tile) ++g;
    int r = tile - tab.group_start[g];
    int nb = (g * GROUP_M) / 2;            // first column block that touches the group
    for (;; ++nb) {
      const int cnt = sym_valid_mblocks(g, nb, tiles_m);
      if (r < cnt) break;
      r -= cnt;
    }
    t.m_blk = g * GROUP_M + r;
    t.n_blk = nb;
  }
  return t;
}

// Incremental form of tile_coord for a CTA that walks tile, tile + stride, tile + 2 stride, ...:
// the symmetric table search costs ~1.3 k instructions per call, which is noise at K = N but was a
// third of the epilogue at K = d (affinity).  advance() moves the cursor forward inside the
// (group, column block) enumeration instead of restarting it.
template <bool SYM>
struct TileCursor {
  int tiles_m, tiles_n;
  int g, nb, base;          // SYM: current group, column block, first tile index of (g, nb)
  __device__ __forceinline__ void init(int tm, int tn) {
    tiles_m = tm; tiles_n = tn; g = 0; nb = 0; base = 0;
  }
  __device__ __forceinline__ TileCoord at(int tile, const TileTable& tab) {
    TileCoord t;
    if (!SYM) {
      const int group = GROUP_M * tiles_n;
      const int gg = tile / group;
      const int first_m = gg * GROUP_M;
      const int gm = min(GROUP_M, tiles_m - first_m);
      const int r = tile - gg * group;
      t.m_blk = first_m + r % gm;
      t.n_blk = r / gm;
    } else {
      // tiles are visited in increasing order: move forward only
      while (g + 1 < tab.num_groups && tab.group_start[g + 1] <= tile) {
        ++g;
        nb = (g * GROUP_M) / 2;
        base = tab.group_start[g];
      }
      for (;;) {
        const int cnt = sym_valid_mblocks(g, nb, tiles_m);
        if (tile - base < cnt) break;
        base += cnt;
        ++nb;
      }
      t.m_blk = g * GROUP_M + (tile - base);
      t.n_blk = nb;
    }
    return t;
  }
};

constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS_V2 = 128 + NUM_EPI_WARPS * 32;   // 384

// C = A B^T with the accumulation split in two levels:
//   level 1 (tensor core, TMEM): a chain of only CHUNK_KB K-blocks (24 or 16 MMAs).  tcgen05
//           accumulates in fp32 with truncation, which biases long chains low by ~5e-8 per MMA
//           (measured: -1.5e-5 relative at K = 1536); short chains keep that below 1e-6.
//   level 2 (CUDA cores, registers): the epilogue warps pull each finished chain out of TMEM
//           (tcgen05.ld) and add it, round-to-nearest, into a register-resident 128 x 256 fp32
//           tile (8 warps x 32 lanes x 128 registers), while the tensor core is already working
//           on the next chain in the other TMEM buffer.
constexpr int STORE_STAGE_BYTES = NUM_EPI_WARPS * 32 * 32 * 4;   // 32 KB: one 32x32 fp32 block per epilogue warp
template <int PREC, int CTAS>
struct StageGeom {
  // planes of one K block in a stage: [A_hi][B_hi][A_lo (PREC>=2)][B_lo (PREC==3)]; in a CTA pair
  // each CTA holds its own 128 rows of A and HALF of the 256 rows of B
  static constexpr int B_BYTES = B_PLANE_BYTES / CTAS;
  static constexpr int BYTES = A_PLANE_BYTES * (PREC >= 2 ? 2 : 1) + B_BYTES * (PREC == 3 ? 2 : 1);
  static constexpr int BUDGET = 227 * 1024 - STORE_STAGE_BYTES - 1024 - 256;
  static constexpr int STAGES = (BUDGET / BYTES) > 8 ? 8 : (BUDGET / BYTES);
};

template <int PREC, int EPI, bool SYM, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS_V2, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a_hi,
               const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi,
               const __grid_constant__ CUtensorMap map_b_lo,
               const __grid_constant__ TileTable tab, int M, int N, int K,
               float* __restrict__ C, int64_t ldc, float* __restrict__ rowmax_offdiag,
               int diag_shift, unsigned int* __restrict__ pace, int pace_kb,
               float* __restrict__ stat_rowmax, double* __restrict__ stat_rowsum,
               float* __restrict__ mirror_out, int64_t ldm, int aff_chunk_kb) {
  constexpr bool SPLIT = PREC >= 2;
  constexpr int STAGES = StageGeom<PREC, CTAS>::STAGES;
  constexpr int STAGE_BYTES = StageGeom<PREC, CTAS>::BYTES;
  constexpr int B_BYTES = StageGeom<PREC, CTAS>::B_BYTES;
  // CTA pair: rank r of the cluster takes tile 2 q + r of the enumeration -- the same column block
  // and the m-blocks (2 j, 2 j + 1), i.e. one 256 x 256 MMA tile per pair (the host only picks
  // this variant when the number of m-blocks is even, which keeps every pair aligned)
  const int cta_rank = (CTAS == 2) ? (int)cluster_ctarank() : 0;
  const int tile_first = (CTAS == 2) ? 2 * ((int)blockIdx.x >> 1) + cta_rank : (int)blockIdx.x;
  const int tile_stride = (int)gridDim.x;
  // K blocks per TMEM chain; the affinity (K = d, output-bound) can afford the shortest chain
  // K blocks per TMEM chain.  Diffuse: 2 (split) / 4 (single).  Affinity (K = d): `aff_chunk_kb`
  // from the host, 1 by default.  Chains of 2 make a d = 256 tile exactly two chains, one per TMEM
  // buffer, so the tensor core runs the NEXT tile while the epilogue warps write this one out:
  // 5.3 -> 4.35 ms at N = 65,536 (profiles/r02_ab_stages_affchain_symmstages_one_box.txt) -- but
  // the 8 extra truncating accumulates push 4 of 1.8 M test elements to 4.5 fp32 ulps from the
  // float64 oracle (bar: 4), and the affinity feeds a threshold, so the faster setting stays an
  // opt-in (SCB_AFFINITY_CHUNK_KB=2).
  const int CHUNK_KB = (EPI == TC_EPI_AFFINITY) ? aff_chunk_kb : (PREC == 1 ? 4 : 2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  float* store_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STORE_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = SYM ? tab.num_tiles : tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const int num_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    if (PREC >= 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    if (PREC == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), NUM_EPI_WARPS * CTAS);   // the leader hears both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();      // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // producer / MMA warpgroup gives its registers to the epilogue warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================================================== TMA producer
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        // Pacing: the CTAs of a wave share their A/B panels through L2 only while they stream K
        // at the same position.  Left alone they drift apart (ncu, N=65,536: 2.9 TB of DRAM reads
        // for 0.05 TB of operands, L2 hit rate 39 %), so every PACE_KB K-blocks each producer
        // signs in on a global counter and waits -- bounded, it is only a hint -- until the
        // whole grid has reached the same checkpoint.  64 K-blocks x ~1.1 MB per wave and K-block
        // keep the live window (~70 MB) inside the 126 MB L2.
        const int ck_per_tile = (num_kb + pace_kb - 1) / pace_kb;
        unsigned int passed = 0;
        TileCursor<SYM> cursor;
        cursor.init(tiles_m, tiles_n);
        for (int round = 0; round * tile_stride < num_tiles; ++round) {
          const int tile = round * tile_stride + tile_first;
          if (tile >= num_tiles) {            // no tile in the last round: keep the targets reachable
            if (pace) atomicAdd(pace, (unsigned int)ck_per_tile);
            continue;
          }
          const TileCoord tc = cursor.at(tile, tab);
          const int row_a = tc.m_blk * BM, row_b = tc.n_blk * BN + cta_rank * (BN / CTAS);
          for (int kb = 0; kb < num_kb; ++kb) {
            if (pace && kb % pace_kb == 0) {
              atomicAdd(pace, 1u);
              ++passed;
              const unsigned int target = passed * gridDim.x;
              for (int spin = 0; spin < PACE_SPINS; ++spin) {
                if (*reinterpret_cast<volatile unsigned int*>(pace) >= target) break;
                __nanosleep(100);
              }
            }
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t bar = smem_u32(&full_bar[stage]);
            const uint32_t base = smem_u32(smem + stage * STAGE_BYTES);
            const int k0 = kb * BK;
            if (CTAS == 2) {
              // the leader's barrier counts the bytes of both CTAs' loads
              if (cta_rank == 0) mbar_expect_tx(bar, (uint32_t)(2 * STAGE_BYTES));
              tma_load_2d_pair(base, &map_a_hi, bar, k0, row_a);
              tma_load_2d_pair(base + A_PLANE_BYTES, &map_b_hi, bar, k0, row_b);
              if (PREC >= 2) tma_load_2d_pair(base + A_PLANE_BYTES + B_BYTES, &map_a_lo, bar, k0, row_a);
              if (PREC == 3)
                tma_load_2d_pair(base + 2 * A_PLANE_BYTES + B_BYTES, &map_b_lo, bar, k0, row_b);
            } else {
              mbar_expect_tx(bar, (uint32_t)STAGE_BYTES);
              tma_load_2d(base, &map_a_hi, bar, k0, row_a);
              tma_load_2d(base + A_PLANE_BYTES, &map_b_hi, bar, k0, row_b);
              if (PREC >= 2) tma_load_2d(base + A_PLANE_BYTES + B_BYTES, &map_a_lo, bar, k0, row_a);
              if (PREC == 3)
                tma_load_2d(base + 2 * A_PLANE_BYTES + B_BYTES, &map_b_lo, bar, k0, row_b);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer
      if (lane == 0 && cta_rank == 0) {
        constexpr uint32_t idesc = make_idesc(BM * CTAS, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        auto umma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accumulate) {
          if (CTAS == 2) umma_f16_pair(d, a, b, id, accumulate);
          else umma_f16(d, a, b, id, accumulate);
        };
        auto commit = [](uint32_t bar) {
          if (CTAS == 2) umma_commit_pair(bar);
          else umma_commit(bar);
        };
        for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
          for (int kb = 0; kb < num_kb; ++kb) {
            const int in_chunk = kb % CHUNK_KB;
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            if (in_chunk == 0) {
              mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);   // chain buffer drained
              tcgen05_fence_after();
            }
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tcgen05_fence_after();
            const uint32_t base = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_hi = make_smem_desc(base);
            const uint64_t b_hi = make_smem_desc(base + A_PLANE_BYTES);
            const uint64_t a_lo = make_smem_desc(base + A_PLANE_BYTES + B_BYTES);
            const uint64_t b_lo = make_smem_desc(base + 2 * A_PLANE_BYTES + B_BYTES);
            // The small cross products go first: tcgen05 truncates each accumulate to the
            // accumulator's current ulp, so terms added while the chain is still small cost
            // almost nothing; only the BK/16 hi*hi accumulates run at full magnitude.
            if (PREC == 3) {
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                const uint64_t adv = (uint64_t)(kk * UMMA_K * 2 / 16);
                umma(tmem_d, a_hi + adv, b_lo + adv, idesc, (in_chunk | kk) ? 1u : 0u);
                umma(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            } else if (PREC == 2) {
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                const uint64_t adv = (uint64_t)(kk * UMMA_K * 2 / 16);
                umma(tmem_d, a_lo + adv, b_hi + adv, idesc, (in_chunk | kk) ? 1u : 0u);
              }
            }
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              const uint64_t adv = (uint64_t)(kk * UMMA_K * 2 / 16);
              umma(tmem_d, a_hi + adv, b_hi + adv, idesc, (SPLIT || in_chunk || kk) ? 1u : 0u);
            }
            commit(smem_u32(&empty_bar[stage]));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            if (in_chunk == CHUNK_KB - 1 || kb == num_kb - 1) {
              commit(smem_u32(&tfull_bar[acc]));            // chain complete
              if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
          }
        }
      }
    }
  } else {
    // ===================================================== epilogue warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int quad = warp & 3;                            // TMEM lanes 32*quad .. +31
    const int half = (warp - 4) >> 2;                     // columns [128*half, +128)
    int acc = 0;
    uint32_t acc_phase = 0;
    TileCursor<SYM> cursor;
    cursor.init(tiles_m, tiles_n);
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const TileCoord tc = cursor.at(tile, tab);
      float sum[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) sum[i] = 0.0f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + half * 128) +
                               ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum[c * 32 + i] += __uint_as_float(v[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_leader(smem_u32(&tempty_bar[acc]));
          else mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      // ---- write-out.  Each thread holds one row x 128 columns; stored straight from registers
      // that is 16 B per lane into 32 different rows.  Instead every 32x32 block goes through a
      // per-warp staging buffer (float4 granules XOR-swizzled by row: conflict-free both ways) and
      // leaves as 128-byte row segments, 4 rows per store instruction.  The same staged block
      // serves the column reductions of the mirrored half (lane = column: 32 conflict-free
      // shared loads instead of 5 shuffles per column).  The instruction count matters: at
      // K = d = 256 (affinity) the epilogue IS the kernel, and a fully unrolled shuffle version
      // (13 k instructions, 200 KB of code) ran at 1 TB/s.
      const int64_t row_base = (int64_t)tc.m_blk * BM + quad * 32;
      const int64_t row = row_base + lane;
      const int64_t col0 = (int64_t)tc.n_blk * BN + half * 128;
      const bool full = (row_base + 32 <= M) && (col0 + 128 <= N);       // warp-uniform
      const int rows_valid = (int)min((int64_t)32, max((int64_t)0, (int64_t)M - row_base));
      // mirror into the tiles that were skipped: target (col, row) lies in tile (col/128, row/256),
      // which is skipped iff col/128 >= 2 (row/256) + 2.  col0 is a multiple of 128, so the
      // decision is uniform over the warp's 32 x 128 block; such a block never meets the diagonal.
      const bool mirror = SYM && (col0 / BM) >= 2 * (row_base / BN) + 2;
      if (EPI == TC_EPI_AFFINITY) {
#pragma unroll
        for (int i = 0; i < 128; ++i) sum[i] = fmaf(sum[i], 0.5f, 0.5f);   // (x + 1) / 2, utils.py:39
        if (rowmax_offdiag) {
          float rmax = 0.0f;
          const int64_t dcol = row + diag_shift;             // this row's diagonal column
          const bool no_diag = (row_base + diag_shift + 32 <= col0) || (row_base + diag_shift >= col0 + 128);
          if (full && no_diag) {
#pragma unroll
            for (int i = 0; i < 128; ++i) rmax = fmaxf(rmax, sum[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 128; ++i)
              if (col0 + i != dcol && col0 + i < N) rmax = fmaxf(rmax, sum[i]);
          }
          if (row < M) atomic_max_nonneg(rowmax_offdiag + row, rmax);
        }
      }
      if (EPI == TC_EPI_PLAIN && stat_rowmax) {
        // fused RowWiseNormalize / degree reductions (refinement.py:243, laplacian.py:41): columns
        // past N are zero-filled by TMA, so they change neither the sum nor a non-negative maximum
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, mx = 0.0f;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          s0 += sum[i]; s1 += sum[i + 1]; s2 += sum[i + 2]; s3 += sum[i + 3];
          mx = fmaxf(fmaxf(mx, fmaxf(sum[i], sum[i + 1])), fmaxf(sum[i + 2], sum[i + 3]));
        }
        if (row < M) {
          atomic_max_nonneg(stat_rowmax + row, mx);
          atomicAdd(stat_rowsum + row, (double)((s0 + s1) + (s2 + s3)));
        }
      }
      {
        float* stg = store_stage + (warp - 4) * 1024;
        const int sr = lane >> 3, sq = lane & 7;             // store phase: row sr of 4, granule sq
        const bool col_stats = mirror && ((EPI == TC_EPI_AFFINITY && rowmax_offdiag != nullptr) ||
                                          (EPI == TC_EPI_PLAIN && stat_rowmax != nullptr));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                make_float4(sum[c * 32 + 4 * q], sum[c * 32 + 4 * q + 1], sum[c * 32 + 4 * q + 2],
                            sum[c * 32 + 4 * q + 3]);
          __syncwarp();
          const int64_t col = col0 + c * 32 + sq * 4;
          float* dst = C + (row_base + sr) * ldc + col;
          if (full) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
              const int r = rr * 4 + sr;
              *reinterpret_cast<float4*>(dst + (int64_t)rr * 4 * ldc) =
                  *reinterpret_cast<const float4*>(stg + r * 32 + ((sq ^ (r & 7)) << 2));
            }
          } else {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
              const int r = rr * 4 + sr;
              const float4 v = *reinterpret_cast<const float4*>(stg + r * 32 + ((sq ^ (r & 7)) << 2));
              if (r < rows_valid) {
                float* d = dst + (int64_t)rr * 4 * ldc;
                if (col + 3 < N) {
                  *reinterpret_cast<float4*>(d) = v;
                } else {
                  if (col < N) d[0] = v.x;
                  if (col + 1 < N) d[1] = v.y;
                  if (col + 2 < N) d[2] = v.z;
                }
              }
            }
          }
          if (col_stats) {
            // lane = column c*32 + lane of the block; the mirrored elements of that column form a
            // piece of ROW col0 + c*32 + lane of C.  Rows past M hold 0 (plain) or 0.5 (affinity)
            // and are skipped.
            float mx = 0.0f, sm = 0.0f;
            const int g = lane >> 2, w4 = lane & 3;
            if (rows_valid == 32) {
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float x = stg[r * 32 + ((g ^ (r & 7)) << 2) + w4];
                mx = fmaxf(mx, x);
                if (EPI == TC_EPI_PLAIN) sm += x;
              }
            } else {
              for (int r = 0; r < rows_valid; ++r) {
                const float x = stg[r * 32 + ((g ^ (r & 7)) << 2) + w4];
                mx = fmaxf(mx, x);
                sm += x;
              }
            }
            const int64_t tr = col0 + c * 32 + lane;
            if (tr < N) {
              if (EPI == TC_EPI_AFFINITY) {
                atomic_max_nonneg(rowmax_offdiag + tr, mx);
              } else {
                atomic_max_nonneg(stat_rowmax + tr, mx);
                atomicAdd(stat_rowsum + tr, (double)sm);
              }
            }
          }
        }
      }
      if (!SYM && mirror_out != nullptr && row < M) {
        // Row-sharded Diffuse: S(g,p) = Y_g Y_p^T is also S(p,g)^T.  `mirror_out` is rank p's row
        // block of S mapped into this process (CUDA IPC over NVLink): the transposed tile is
        // stored there straight from the accumulator registers -- the exchange step of the
        // sharded product rides on the GEMM epilogue, tile by tile, instead of a separate
        // transpose + send/recv after the last product.  Lanes hold consecutive rows, so every
        // store instruction writes 128 contiguous bytes of one peer row.
        float* dst = mirror_out + col0 * ldm + row;
        if (col0 + 128 <= N) {
#pragma unroll
          for (int i = 0; i < 128; ++i) dst[(int64_t)i * ldm] = sum[i];
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (col0 + i < N) dst[(int64_t)i * ldm] = sum[i];
        }
      }
      if (mirror && row < M) {
        float* dst = C + col0 * ldc + row;                 // lanes -> consecutive rows
        if (col0 + 128 <= N) {
#pragma unroll
          for (int i = 0; i < 128; ++i) dst[(int64_t)i * ldc] = sum[i];
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (col0 + i < N) dst[(int64_t)i * ldc] = sum[i];
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();      // nobody leaves while the peer may still signal it
  if (warp == 2) {
    tcgen05_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"
                   ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                   ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp16 tensor [rows, k] with leading dimension ld (elements); box = BK x box_rows.
static int make_plane_map(CUtensorMap* map, const __half* ptr, int64_t rows, int64_t k,
                          int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  SC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0,
             "tcgen05 GEMM: fp16 planes need 16-byte aligned base and ld %% 8 == 0");
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static void build_tile_table(TileTable& tab, int tiles_m, int tiles_n) {
  tab.num_groups = (tiles_m + GROUP_M - 1) / GROUP_M;
  int total = 0;
  for (int g = 0; g < tab.num_groups; ++g) {
    tab.group_start[g] = total;
    for (int nb = (g * GROUP_M) / 2; nb < tiles_n; ++nb) total += sym_valid_mblocks(g, nb, tiles_m);
  }
  tab.group_start[tab.num_groups] = total;
  tab.num_tiles = total;
}

template <int PREC, int EPI, bool SYM, int CTAS>
static int launch(const sc_context* ctx, const CUtensorMap& ah, const CUtensorMap& al,
                  const CUtensorMap& bh, const CUtensorMap& bl, int M, int N, int K, float* C,
                  int64_t ldc, float* rowmax, int diag_shift, float* stat_rowmax,
                  double* stat_rowsum, float* mirror, int64_t ldm, cudaStream_t st) {
  constexpr int STAGES = StageGeom<PREC, CTAS>::STAGES;
  constexpr int STAGE_BYTES = StageGeom<PREC, CTAS>::BYTES;
  static_assert(STAGES >= 2, "tcgen05 GEMM: the operand ring needs at least two stages");
  const size_t smem = (size_t)STAGES * STAGE_BYTES + STORE_STAGE_BYTES + 1024 /*align*/ +
                      256 /*barriers*/;
  SC_REQUIRE(smem <= ctx->smem_optin, "tcgen05 GEMM needs %zu B of shared memory", smem);
  auto kern = k_gemm_tcgen05<PREC, EPI, SYM, CTAS>;
  SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  TileTable tab = {};
  int tiles = tiles_m * tiles_n;
  if (SYM) {
    SC_REQUIRE((tiles_m + GROUP_M - 1) / GROUP_M <= MAX_GROUPS, "tcgen05 GEMM: N too large for "
               "the symmetric tile table");
    build_tile_table(tab, tiles_m, tiles_n);
    tiles = tab.num_tiles;
  }
  int sms = ctx->sm_count;
  if (ctx->gemm_sm_limit > 0 && ctx->gemm_sm_limit < sms) sms = ctx->gemm_sm_limit;
  int grid = tiles < sms ? tiles : sms;
  if (CTAS == 2) grid &= ~1;                 // whole CTA pairs (the tile count is even)
  unsigned int* pace = nullptr;
  static int pace_kb = 0;
  if (pace_kb == 0) {
    const char* e = getenv("SCB_GEMM_PACE_KB");
    pace_kb = (e && atoi(e) > 0) ? atoi(e) : PACE_KB;
  }
  if (ctx->gemm_pace && (K + BK - 1) / BK >= 2 * pace_kb && tiles > grid) {
    static std::atomic<unsigned int> pace_turn{0};
    pace = ctx->gemm_pace + (pace_turn.fetch_add(1) % SC_GEMM_PACE_SLOTS);
    SC_CUDA(cudaMemsetAsync(pace, 0, sizeof(unsigned int), st));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS_V2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int aff_chunk_kb = 0;
  if (aff_chunk_kb == 0) {
    const char* e = getenv("SCB_AFFINITY_CHUNK_KB");
    aff_chunk_kb = (e && atoi(e) > 0) ? atoi(e) : 1;
  }
  SC_CUDA(cudaLaunchKernelEx(&cfg, kern, ah, al, bh, bl, tab, M, N, K, C, ldc, rowmax, diag_shift,
                             pace, pace_kb, stat_rowmax, stat_rowsum, mirror, ldm, aff_chunk_kb));
  sc::launched();
  SC_LAUNCH_CHECK();
  return 0;
}

// CTA pairs (cta_group::2, one 256 x 256 MMA tile per two SMs: a third less operand traffic from L2
// and half the B reads from shared memory per SM) whenever the m-blocks pair up; SCB_GEMM_2CTA=0
// forces the single-CTA kernel.
static bool want_cta_pair(int M, int K, int prec) {
  static int mode = -1;                      // 0 never, 1 whenever possible, 2 measured policy
  if (mode < 0) {
    const char* e = getenv("SCB_GEMM_2CTA");
    mode = e ? (atoi(e) == 0 ? 0 : 1) : 2;
  }
  const int tiles_m = (M + BM - 1) / BM;
  if (mode == 0 || tiles_m < 2 || tiles_m % 2 != 0) return false;
  // Measured (profiles/r02_gemm_2cta.txt, r02_ab_stages_one_box.txt): the pair wins where the four
  // operand planes of the three-MMA split crowd L2 and shared memory -- N = K = 65,536: 602 vs
  // 659 ms -- and loses a few per cent where a single MMA per stage is already power-bound
  // (single: 228 vs 223 ms in predict(); split2: 458 vs 434 ms) and on the short-K affinity
  // (K = 256: 5.7 vs 5.25 ms).
  return mode == 1 || (prec == 3 && K >= 2048);
}

int gemm_nt_tcgen05(const sc_context* ctx, int epi, int precision, const __half* a_hi,
                    const __half* a_lo, int64_t lda, const __half* b_hi, const __half* b_lo,
                    int64_t ldb, int64_t M, int64_t N, int64_t K, float* C, int64_t ldc,
                    float* rowmax_offdiag, bool symmetric, int diag_shift, float* stat_rowmax,
                    double* stat_rowsum, float* mirror, int64_t ldm, cudaStream_t st) {
  SC_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31),
             "tcgen05 GEMM: bad shape");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0,
             "tcgen05 GEMM: C needs a 16-byte aligned base and ldc %% 4 == 0");
  SC_REQUIRE(precision == SC_GEMM_SPLIT3 || precision == SC_GEMM_SPLIT2 ||
             precision == SC_GEMM_SINGLE, "tcgen05 GEMM: unknown precision %d", precision);
  SC_REQUIRE((stat_rowmax == nullptr) == (stat_rowsum == nullptr),
             "tcgen05 GEMM: the row statistics come together");
  const int prec = precision == SC_GEMM_SPLIT3 ? 3 : (precision == SC_GEMM_SPLIT2 ? 2 : 1);
  const bool pair = want_cta_pair((int)M, (int)K, prec);
  const int b_box = pair ? BN / 2 : BN;      // each CTA of a pair loads half of the B tile
  CUtensorMap ah, al, bh, bl;
  if (int r = make_plane_map(&ah, a_hi, M, K, lda, BM)) return r;
  if (int r = make_plane_map(&bh, b_hi, N, K, ldb, b_box)) return r;
  al = ah;
  bl = bh;
  if (prec >= 2) {
    SC_REQUIRE(a_lo, "tcgen05 GEMM: the lo plane of A is missing");
    if (int r = make_plane_map(&al, a_lo, M, K, lda, BM)) return r;
  }
  if (prec == 3) {
    SC_REQUIRE(b_lo, "tcgen05 GEMM: the lo plane of B is missing");
    if (int r = make_plane_map(&bl, b_lo, N, K, ldb, b_box)) return r;
  }
  const int m = (int)M, n = (int)N, k = (int)K;
  // C = Y Y^T is symmetric when both operands are the same matrix: compute the upper tiles only
  const bool sym = symmetric && a_hi == b_hi && a_lo == b_lo && M == N && lda == ldb;
  SC_REQUIRE(!(sym && mirror), "tcgen05 GEMM: a mirrored copy only makes sense for an off-diagonal block");
#define SC_TC_LAUNCH(PR, EP, SY)                                                                 \
  do {                                                                                           \
    if (pair)                                                                                    \
      return launch<PR, EP, SY, 2>(ctx, ah, al, bh, bl, m, n, k, C, ldc, rowmax_offdiag,         \
                                   diag_shift, stat_rowmax, stat_rowsum, mirror, ldm, st);       \
    return launch<PR, EP, SY, 1>(ctx, ah, al, bh, bl, m, n, k, C, ldc, rowmax_offdiag,           \
                                 diag_shift, stat_rowmax, stat_rowsum, mirror, ldm, st);         \
  } while (0)
#define SC_TC_PREC(PR)                                                                        \
  do {                                                                                        \
    if (epi == TC_EPI_AFFINITY) { if (sym) SC_TC_LAUNCH(PR, TC_EPI_AFFINITY, true);           \
                                  SC_TC_LAUNCH(PR, TC_EPI_AFFINITY, false); }                 \
    if (sym) SC_TC_LAUNCH(PR, TC_EPI_PLAIN, true);                                            \
    SC_TC_LAUNCH(PR, TC_EPI_PLAIN, false);                                                    \
  } while (0)
  if (prec == 3) SC_TC_PREC(3);
  if (prec == 2) SC_TC_PREC(2);
  SC_TC_PREC(1);
#undef SC_TC_PREC
#undef SC_TC_LAUNCH
}

}  // namespace sc
