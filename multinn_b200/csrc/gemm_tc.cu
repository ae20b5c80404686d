// fp32-accurate GEMM on the 5th-gen tensor cores (tcgen05, kind::tf32) by error-compensated splitting ("3xTF32"):
//   x = hi + lo, hi = the top 19 bits of the fp32 word (what kind::tf32 reads), lo = x - hi (exact in fp32)
//   A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi      (lo.lo ~ 2^-22 dropped), fp32 accumulation in TMEM.
// Replaces the tf.matmul / Dense / LSTMBlockCell matmuls of the reference (common/rnn.py:124,
// generators/rnn_nade.py:54-57,212, generators/rnn_rbm.py:252-253, common/rbm.py:351,370) and the GEMMs
// tf.gradients derives from them.
//
// Pipeline per CTA (persistent over output tiles, optional split-K), 384 threads:
//   warp 0      TMA producer: raw fp32 tiles of A and B -> smem ring (SWIZZLE_128B), mbarrier complete_tx
//   warps 8-11  converters: lo = x - trunc_tf32(x) for every element of the stage (same swizzled layout, so it is
//               a flat elementwise pass), fence.proxy.async, arrive
//   warp 1      MMA issuer: one thread, 2-3 tcgen05.mma per K=8 step into double-buffered TMEM accumulators (main:
//               hi.hi, aux: the cross terms), tcgen05.commit frees the smem stage / publishes the accumulators
//   warps 4-7   epilogue: tcgen05.ld 32x32b -> registers -> alpha*acc + bias + beta*C -> global (or red.add for split-K)
// Operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); all four combinations are
// expressed through the UMMA instruction descriptor's a_major/b_major bits, so no transposed copies exist.
#include <cuda_bf16.h>
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "multinn_b200.h"
#include "tc_common.cuh"

namespace mnn {
namespace tc {

constexpr int kThreads = 384;
constexpr int kConvThreads = 128;
constexpr int kEpiThreads = 128;

struct Params {
  float* C;
  const float* bias;
  long long ldc;
  float alpha, beta;
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  int n_products;  // 3, or 2 when A is exactly representable in tf32 (binary piano-roll rows)
  int atomic;      // split-K: red.global.add into C
  int tma_store;   // pair kernel: C tiles leave through shared memory + TMA (needs beta == 0, no split-K, aligned C)
  int bf16x;       // pair kernel: cross terms hi.lo + lo.hi as bf16 MMAs (kind::f16, twice the tf32 rate)
  int share_conv;  // pair kernel, long K loops: the epilogue warps convert too (256 converter threads per CTA)
  int a_pre;       // pair kernel, bf16x == 2, binary A: A arrives as ONE exact bf16 plane (no raw A stage, no A conversion)
  int b_pre;       // pair kernel, bf16x == 2: B arrives as bf16 pair planes (mnn_split_bf16_pair); TMA drops them straight into
                   // the [hi | lo] tiles: no raw B stage, no B conversion (a third less shared-memory traffic per stage)
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 4;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // raw + lo for both operands
  static constexpr int STAGES = (BN == 128 ? 3 : 4);
  // two accumulators per tile (main: hi.hi, aux: hi.lo + lo.hi), double buffered across tiles
  static constexpr int TMEM_COLS = 4 * BN;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;  // + alignment slack
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  using C_ = Cfg<BN>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 4];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = smem_u32(&bars[0]);              // [STAGES]
  const uint32_t bar_conv = smem_u32(&bars[STAGES]);         // [STAGES]
  const uint32_t bar_empty = smem_u32(&bars[2 * STAGES]);    // [STAGES]
  const uint32_t bar_tfull = smem_u32(&bars[3 * STAGES]);    // [2]
  const uint32_t bar_tempty = smem_u32(&bars[3 * STAGES + 2]);  // [2]

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvThreads);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_items = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int n_blk = w % p.tiles_n, m_blk = (w / p.tiles_n) % p.tiles_m, split = w / (p.tiles_n * p.tiles_m);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, C_::A_BYTES + C_::B_BYTES);
          const uint32_t a_dst = smem0 + stage * C_::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * C_::A_BYTES;
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(a_dst, &map_a, full, k0, m0);                 // box {32 k, 128 rows}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_dst + c * (BK * 128), &map_a, full, m0 + 32 * c, k0);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &map_b, full, k0, n0);                 // box {32 k, BN rows}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(b_dst + c * (BK * 128), &map_b, full, n0 + 32 * c, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(BN, A_MN, B_MN), idesc2 = idesc_tf32(2 * BN, A_MN, B_MN);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), LBO unused (=16 B like CUTLASS); k-step = +32 B in the row.
      // MN-major (SWIZZLE_128B_BASE32B): 32-element column chunks BK*128 B apart (LBO), 4-k-row atoms 512 B apart (SBO);
      // one K=8 instruction spans two atoms, k-step = +1024 B.
      const uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
      const uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
      const uint32_t a_lay = A_MN ? 1 : 2, b_lay = B_MN ? 1 : 2;
      const uint32_t a_kstep = A_MN ? 1024 : 32, b_kstep = B_MN ? 1024 : 32;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int split = w / (p.tiles_n * p.tiles_m);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        // main accumulator takes only the hi.hi products: the tensor core truncates its fp32 accumulator on every
        // instruction, so the error of a chain grows with its length; the small cross terms go to their own chain
        const uint32_t tmem_d = tmem_base + acc * 2 * BN, tmem_x = tmem_d + BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          mbar_wait(bar_conv + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_raw = smem0 + stage * C_::STAGE_BYTES, a_lo = a_raw + C_::A_BYTES;
          const uint32_t b_raw = a_raw + 2 * C_::A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t da = smem_desc(a_raw + j * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t db = smem_desc(b_raw + j * b_kstep, b_lbo, b_sbo, b_lay);
            const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
            // the lo tile of B sits right behind the raw tile in the same layout, so [B_hi ; B_lo] is ONE operand of
            // 2*BN rows: a single N = 2*BN instruction yields A_hi.B_hi (columns [0,BN) = main) and A_hi.B_lo
            // (columns [BN,2BN) = aux) while reading A from shared memory once
            umma_tf32(tmem_d, da, db, idesc2, first);
            if (p.n_products == 3) {
              const uint64_t dal = smem_desc(a_lo + j * a_kstep, a_lbo, a_sbo, a_lay);
              umma_tf32(tmem_x, dal, db, idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * stage);   // smem stage reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);        // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ converters: lo = x - trunc_tf32(x)
    const int t = threadIdx.x - 8 * 32;
    int stage = 0;
    uint32_t phase = 0;
    const int a_vec = (p.n_products == 3) ? C_::A_BYTES / 16 : 0;
    constexpr int b_vec = C_::B_BYTES / 16;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int split = w / (p.tiles_n * p.tiles_m);
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        uint8_t* base = smem_gen + (size_t)stage * C_::STAGE_BYTES;
        const float4* a_raw = reinterpret_cast<const float4*>(base);
        float4* a_lo = reinterpret_cast<float4*>(base + C_::A_BYTES);
        const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * C_::A_BYTES);
        float4* b_lo = reinterpret_cast<float4*>(base + 2 * C_::A_BYTES + C_::B_BYTES);
        auto split4 = [](float4 x) { return tf32_lo4(x); };
#pragma unroll 4
        for (int i = t; i < a_vec; i += kConvThreads) a_lo[i] = split4(a_raw[i]);
#pragma unroll 4
        for (int i = t; i < b_vec; i += kConvThreads) b_lo[i] = split4(b_raw[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
        mbar_arrive(bar_conv + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (warps 4..7 -> TMEM lane quadrants 0..3)
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int n_blk = w % p.tiles_n, m_blk = (w / p.tiles_n) % p.tiles_m, split = w / (p.tiles_n * p.tiles_m);
      const int row = m_blk * BM + q * 32 + lane;
      const int n0 = n_blk * BN;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const bool add_bias = p.bias != nullptr && (!p.atomic || split == 0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;   // warp-uniform
        float v[32], vx[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + c * 32), v);
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + BN + c * 32), vx);
        if (row < p.M) {
          float* dst = p.C + (size_t)row * p.ldc + col0;
          const bool full = col0 + 32 <= p.N;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float o = p.alpha * (v[i] + vx[i]);
            if (add_bias && (full || col0 + i < p.N)) o += __ldg(p.bias + col0 + i);
            v[i] = o;
          }
          if (p.atomic) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (full || col0 + i < p.N) atomicAdd(dst + i, v[i]);
          } else if (full && vec_ok) {
            float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (p.beta != 0.f) {
                const float4 old = d4[i];
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              d4[i] = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) dst[i] = (p.beta != 0.f) ? v[i] + p.beta * dst[i] : v[i];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C_::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ 2-CTA variant
// The same pipeline on a CTA PAIR (cluster of 2, tcgen05 cta_group::2): output tile 256 x 256, each CTA holds 128 rows
// of A and 128 of the 256 B rows per stage and ends with the accumulators of its own 128 output rows (main and aux,
// 256 TMEM columns each). Why: a 128x128 tf32 instruction reads 8 KB of operands from shared memory per 64 clk -- all of
// the 128 B/clk the SM has -- before the TMA writes and the hi/lo conversion are counted, so the 1-CTA kernel cannot
// exceed ~55 % tensor utilisation; a 256x256 pair instruction reads 8 KB per CTA per 128 clk.
//   * every role exists in both CTAs; only the leader's (rank 0) MMA thread issues, its commits are multicast to both CTAs
//   * converters of both CTAs arrive on the LEADER's conv barrier (256 arrivals) -> "both halves of the stage are ready"
//   * epilogue threads of both CTAs arrive on the LEADER's tempty barrier (256 arrivals) -> "both TMEM halves are drained"
constexpr int BN2 = 256;
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 4;         // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN2 / 2) * BK * 4;  // this CTA's 128 rows of B
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int TMEM_COLS = 2 * BN2;           // main + aux, single buffered
  static constexpr int CTILE = BM * 128;              // [128 rows x 32 cols] fp32 staging tile of the TMA-store epilogue
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * CTILE + 1024;
};

// one converter thread's share of a landed stage [A raw | A lo | B raw | B lo]: the tf32 residual tiles (3xTF32) or the bf16
// cross-term tiles (bf16x); t in [0, 128) or, when the epilogue warps convert too (share_conv), [0, 256)
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void convert_stage(uint8_t* base, int t, const Params& p) {
  using C_ = Cfg2;
  if (p.bf16x) {
    const bool pair = p.bf16x == 2;
    if (p.share_conv) {
      if (!p.a_pre) convert_bf16_tiles<A_MN, 256>(base, base + C_::A_BYTES, t, p.n_products == 3, pair);
      if (!p.b_pre) convert_bf16_tiles<B_MN, 256>(base + 2 * C_::A_BYTES, base + 2 * C_::A_BYTES + C_::B_BYTES, t, true, pair);
    } else {
      if (!p.a_pre) convert_bf16_tiles<A_MN, 128>(base, base + C_::A_BYTES, t, p.n_products == 3, pair);
      if (!p.b_pre) convert_bf16_tiles<B_MN, 128>(base + 2 * C_::A_BYTES, base + 2 * C_::A_BYTES + C_::B_BYTES, t, true, pair);
    }
  } else {
    const float4* a_raw = reinterpret_cast<const float4*>(base);
    float4* a_lo = reinterpret_cast<float4*>(base + C_::A_BYTES);
    const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * C_::A_BYTES);
    float4* b_lo = reinterpret_cast<float4*>(base + 2 * C_::A_BYTES + C_::B_BYTES);
    const int nt = p.share_conv ? 2 * kConvThreads : kConvThreads;
    const int a_vec = (p.n_products == 3) ? C_::A_BYTES / 16 : 0;
    constexpr int b_vec = C_::B_BYTES / 16;
#pragma unroll 4
    for (int i = t; i < a_vec; i += nt) a_lo[i] = tf32_lo4(a_raw[i]);
#pragma unroll 4
    for (int i = t; i < b_vec; i += nt) b_lo[i] = tf32_lo4(b_raw[i]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the pair's UMMA
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_c, const Params p) {
  using C_ = Cfg2;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 4];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  const uint32_t bar_full = smem_u32(&bars[0]);              // [STAGES] local TMA completion
  const uint32_t bar_conv = smem_u32(&bars[STAGES]);         // [STAGES] leader: 2 x 128 converter arrivals
  const uint32_t bar_empty = smem_u32(&bars[2 * STAGES]);    // [STAGES] multicast MMA commit
  const uint32_t bar_tfull = smem_u32(&bars[3 * STAGES]);    // [2] multicast MMA commit, one per accumulator buffer
  const uint32_t bar_tempty = smem_u32(&bars[3 * STAGES + 2]);  // [2] leader: 2 x 128 epilogue arrivals
  // bf16-pair split (bf16x == 2): all three products go into ONE accumulator, so the 512 TMEM columns hold two of them
  // and the epilogue of tile i overlaps the mainloop of tile i + 1; the other splits keep main + aux, single buffered
  const bool acc2 = p.bf16x == 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 2 * ((p.share_conv ? 2 : 1) * kConvThreads / 32));   // one elected arrive per converter warp
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * (kEpiThreads / 32));        // one elected arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_items = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (own halves of A and B)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        const int n_blk = w % p.tiles_n, m_blk = (w / p.tiles_n) % p.tiles_m, split = w / (p.tiles_n * p.tiles_m);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * 2 * BM + (int)rank * BM, n0 = n_blk * BN2 + (int)rank * (BN2 / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, (p.a_pre ? C_::A_BYTES / 2 : C_::A_BYTES) + C_::B_BYTES);
          const uint32_t a_dst = smem0 + stage * C_::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * C_::A_BYTES;
          const int k0 = kb * BK;
          if (p.a_pre) {
            // exact bf16 A (binary piano-roll rows): one plane straight into the hi tile of the "lo" region
            const uint32_t ah = a_dst + C_::A_BYTES;
            if (!A_MN) {
              tma_load_2d(ah, &map_a, full, k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c) tma_load_2d(ah + c * 4096, &map_a, full, m0 + 64 * c, k0);
            }
          } else if (!A_MN) {
            tma_load_2d(a_dst, &map_a, full, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_dst + c * (BK * 128), &map_a, full, m0 + 32 * c, k0);
          }
          if (p.b_pre) {
            // pre-split B: bf16 planes (map_b: hi, map_b2: lo) land in the [hi | lo] tiles of the "lo" region in the UMMA
            // layout the converter would have written: K-major SWIZZLE_64B box {32 k, 128 rows}; MN-major SWIZZLE_128B
            // boxes {64 n, 32 k}, the two 64-n blocks 4 KB apart. Same byte count as the raw fp32 tile.
            const uint32_t bh = b_dst + C_::B_BYTES, bl = bh + C_::B_BYTES / 2;
            if (!B_MN) {
              tma_load_2d(bh, &map_b, full, k0, n0);
              tma_load_2d(bl, &map_b2, full, k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                tma_load_2d(bh + c * 4096, &map_b, full, n0 + 64 * c, k0);
                tma_load_2d(bl + c * 4096, &map_b2, full, n0 + 64 * c, k0);
              }
            }
          } else if (!B_MN) {
            tma_load_2d(b_dst, &map_b, full, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN2 / 64; ++c) tma_load_2d(b_dst + c * (BK * 128), &map_b, full, n0 + 32 * c, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one thread of the LEADER CTA
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = idesc_tf32(BN2, A_MN, B_MN, 2 * BM);
      const uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
      const uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
      const uint32_t a_lay = A_MN ? 1 : 2, b_lay = B_MN ? 1 : 2;
      const uint32_t a_kstep = A_MN ? 1024 : 32, b_kstep = B_MN ? 1024 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int tile_no = 0;
      // bf16 cross-term tiles (convert_bf16_tiles): [hi | lo] of 8 KB each in the operand's "lo" region
      const uint32_t idesc16 = idesc_bf16(BN2, A_MN, B_MN, 2 * BM);
      const uint32_t a16_lbo = A_MN ? 4096 : 16, b16_lbo = B_MN ? 4096 : 16;
      const uint32_t a16_sbo = A_MN ? 1024 : 512, b16_sbo = B_MN ? 1024 : 512;
      const uint32_t a16_lay = A_MN ? 2 : 4, b16_lay = B_MN ? 2 : 4;
      const uint32_t a16_kstep = A_MN ? 2048 : 32, b16_kstep = B_MN ? 2048 : 32;
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        const int split = w / (p.tiles_n * p.tiles_m);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int abuf = acc2 ? (tile_no & 1) : 0, use_no = acc2 ? (tile_no >> 1) : tile_no;
        const uint32_t tmem_d = tmem_base + (uint32_t)(abuf * BN2), tmem_x = acc2 ? tmem_d : tmem_base + BN2;
        mbar_wait(bar_tempty + 8 * abuf, (uint32_t)((use_no & 1) ^ 1));
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_conv + 8 * stage, phase);   // both CTAs: TMA landed and lo tiles written
          tc_fence_after();
          const uint32_t a_raw = smem0 + stage * C_::STAGE_BYTES, a_lo = a_raw + C_::A_BYTES;
          const uint32_t b_raw = a_raw + 2 * C_::A_BYTES, b_lo = b_raw + C_::B_BYTES;
          if (p.bf16x) {
            // A.B ~ A_hi.B_hi (tf32, 4 x K=8) + bf16(A).bf16(B_lo) + bf16(A_lo).bf16(B) (kind::f16, 2 x K=16 each)
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
              const uint64_t dah = smem_desc(a_lo + j * a16_kstep, a16_lbo, a16_sbo, a16_lay);
              const uint64_t dbl = smem_desc(b_lo + C_::B_BYTES / 2 + j * b16_kstep, b16_lbo, b16_sbo, b16_lay);
              umma_bf16_2cta(tmem_x, dah, dbl, idesc16, first);
              if (p.n_products == 3) {
                const uint64_t dal = smem_desc(a_lo + C_::A_BYTES / 2 + j * a16_kstep, a16_lbo, a16_sbo, a16_lay);
                const uint64_t dbh = smem_desc(b_lo + j * b16_kstep, b16_lbo, b16_sbo, b16_lay);
                umma_bf16_2cta(tmem_x, dal, dbh, idesc16, 1u);
              }
            }
            if (p.bf16x == 2) {
              // all-bf16 pair scheme: the main product too is a kind::f16 MMA, A1.B1 (twice the tf32 rate, half the
              // operand bytes); [hi | lo] = the bf16 pair x1 + x2. tmem_x == tmem_d here: the cross terms above opened
              // the accumulator, so the main product always accumulates
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                const uint64_t dah = smem_desc(a_lo + j * a16_kstep, a16_lbo, a16_sbo, a16_lay);
                const uint64_t dbh = smem_desc(b_lo + j * b16_kstep, b16_lbo, b16_sbo, b16_lay);
                umma_bf16_2cta(tmem_d, dah, dbh, idesc16, 1u);
              }
            } else {
#pragma unroll
              for (int j = 0; j < BK / 8; ++j) {
                const uint64_t da = smem_desc(a_raw + j * a_kstep, a_lbo, a_sbo, a_lay);
                const uint64_t db = smem_desc(b_raw + j * b_kstep, b_lbo, b_sbo, b_lay);
                umma_tf32_2cta(tmem_d, da, db, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
              }
            }
            umma_commit_2cta(bar_empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t da = smem_desc(a_raw + j * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t db = smem_desc(b_raw + j * b_kstep, b_lbo, b_sbo, b_lay);
            const uint64_t dbl = smem_desc(b_lo + j * b_kstep, b_lbo, b_sbo, b_lay);
            const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
            umma_tf32_2cta(tmem_x, da, dbl, idesc, first);
            if (p.n_products == 3) {
              const uint64_t dal = smem_desc(a_lo + j * a_kstep, a_lbo, a_sbo, a_lay);
              umma_tf32_2cta(tmem_x, dal, db, idesc, 1u);
            }
            umma_tf32_2cta(tmem_d, da, db, idesc, first);
          }
          umma_commit_2cta(bar_empty + 8 * stage);   // both CTAs may refill this stage
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(bar_tfull + 8 * abuf);      // both CTAs' accumulator halves complete
        ++tile_no;
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ converters: lo = x - trunc_tf32(x)
    const int t = threadIdx.x - 8 * 32;
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t conv_leader = mapa_cluster(bar_conv, 0);
    for (int w = cluster_id; w < n_items; w += n_clusters) {
      const int split = w / (p.tiles_n * p.tiles_m);
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        convert_stage<A_MN, B_MN>(smem_gen + (size_t)stage * C_::STAGE_BYTES, t, p);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(conv_leader + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: this CTA's 128 rows x 256 columns
    const int q = warp - 4;
    int tile_no = 0;
    int store_no = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const uint32_t tempty_leader = mapa_cluster(bar_tempty, 0);
    const uint32_t conv_leader = mapa_cluster(bar_conv, 0);
    int cstage = 0;
    uint32_t cphase = 0;
    for (int w = cluster_id; w < n_items; w += n_clusters) {
      const int n_blk = w % p.tiles_n, m_blk = (w / p.tiles_n) % p.tiles_m, split = w / (p.tiles_n * p.tiles_m);
      const int row = m_blk * 2 * BM + (int)rank * BM + q * 32 + lane;
      const int n0 = n_blk * BN2;
      if (p.share_conv) {
        // long K loop: the epilogue has nothing to do until the tile is complete, so these warps convert as well
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * cstage, cphase);
          convert_stage<A_MN, B_MN>(smem_gen + (size_t)cstage * C_::STAGE_BYTES, (int)threadIdx.x, p);   // t in [128, 256)
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(conv_leader + 8 * cstage);
          if (++cstage == STAGES) { cstage = 0; cphase ^= 1; }
        }
      }
      const int abuf = acc2 ? (tile_no & 1) : 0, use_no = acc2 ? (tile_no >> 1) : tile_no;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * BN2);
      mbar_wait(bar_tfull + 8 * abuf, (uint32_t)(use_no & 1));
      tc_fence_after();
      const bool add_bias = p.bias != nullptr && (!p.atomic || split == 0);
      if (p.tma_store) {
        // row-per-thread global stores touch 32 cache lines per warp instruction (4 us per tile, as long as a K = 256
        // mainloop); instead: registers -> swizzled [128 x 32] staging tile -> TMA store, double buffered
        const int r = q * 32 + lane, sw = r & 7;
        const int row_t = m_blk * 2 * BM + (int)rank * BM;
#pragma unroll 1
        for (int c = 0; c < BN2 / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;   // warp-uniform
          float v[32], vx[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), v);
          if (acc2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) vx[i] = 0.f;
          } else {
            tmem_ld32(tacc + (uint32_t)(BN2 + c * 32), vx);
          }
          float bv = 0.f;
          if (p.bias != nullptr && col0 + lane < p.N) bv = __ldg(p.bias + col0 + lane);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(p.alpha, v[i] + vx[i], __shfl_sync(0xffffffffu, bv, i));
          const int buf = store_no & 1;
          if (threadIdx.x == 4 * 32 && store_no >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 2, 128;" ::: "memory");   // staging buffer `buf` has been read by its previous store
          uint8_t* rowp = smem_gen + (size_t)STAGES * C_::STAGE_BYTES + (size_t)buf * C_::CTILE + r * 128;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(rowp + ((jj ^ sw) << 4)) = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (threadIdx.x == 4 * 32) {
            tma_store_2d(&map_c, smem0 + STAGES * C_::STAGE_BYTES + buf * C_::CTILE, col0, row_t);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++store_no;
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < BN2 / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;   // warp-uniform
        float v[32], vx[32];
        tmem_ld32(tacc + (uint32_t)(c * 32), v);
        if (acc2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) vx[i] = 0.f;
        } else {
          tmem_ld32(tacc + (uint32_t)(BN2 + c * 32), vx);
        }
        if (row < p.M) {
          float* dst = p.C + (size_t)row * p.ldc + col0;
          const bool full = col0 + 32 <= p.N;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float o = p.alpha * (v[i] + vx[i]);
            if (add_bias && (full || col0 + i < p.N)) o += __ldg(p.bias + col0 + i);
            v[i] = o;
          }
          if (p.atomic) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (full || col0 + i < p.N) atomicAdd(dst + i, v[i]);
          } else if (full && vec_ok) {
            float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (p.beta != 0.f) {
                const float4 old = d4[i];
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              d4[i] = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) dst[i] = (p.beta != 0.f) ? v[i] + p.beta * dst[i] : v[i];
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * abuf);
      ++tile_no;
    }
    if (threadIdx.x == 4 * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C_::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr; long long ld; long long inner, outer; int box_outer;   // box_outer < 0 encodes the MN-major swizzle
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && inner == o.inner && outer == o.outer && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.ld * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= ((size_t)k.inner * 0x9E3779B1ull ^ (size_t)k.outer) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    return h ^ (size_t)k.box_outer * 0x165667B19E3779F9ull;
  }
};

// 2-D fp32 tensor map over a row-major matrix [outer][inner] with row stride ld; box {32, box_outer}; SWIZZLE_128B for
// K-major tiles, SWIZZLE_128B_ATOM_32B for MN-major tiles (matches UMMA SWIZZLE_128B_BASE32B).
static int make_map(const float* ptr, long long ld, long long inner, long long outer, int box_outer, bool mn_major,
                    CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, ld, inner, outer, mn_major ? -box_outer : box_outer};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MNN_OK; }
  }
  EncodeTiledFn enc = get_encode();
  MNN_REQUIRE(enc != nullptr, MNN_ERR_UNSUPPORTED, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mnn_set_error("gemm_tc: cuTensorMapEncodeTiled failed");
    return MNN_ERR_ARG;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return MNN_OK;
}

// 2-D fp32 tensor map with an arbitrary box and no swizzle (small elementwise tiles staged by the LSTM cell phases)
static int make_map_plain(const float* ptr, long long ld, long long inner, long long outer, int box_inner, int box_outer,
                          CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, ld, inner, outer, (box_inner << 16) | box_outer};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MNN_OK; }
  }
  EncodeTiledFn enc = get_encode();
  MNN_REQUIRE(enc != nullptr, MNN_ERR_UNSUPPORTED, "tensor map: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mnn_set_error("tensor map: cuTensorMapEncodeTiled failed");
    return MNN_ERR_ARG;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return MNN_OK;
}

// 2-D bf16 tensor map over a row-major matrix [outer][inner], box {32, box_outer}, SWIZZLE_64B: a box lands in shared memory
// in the UMMA canonical K-major SWIZZLE_64B layout (64-byte rows, 8-row atoms) that kind::f16 descriptors (layout 4) read
static int make_map_bf16(const void* ptr, long long ld_elems, long long inner, long long outer, int box_outer,
                         CUtensorMap* out) {
  EncodeTiledFn enc = get_encode();
  MNN_REQUIRE(enc != nullptr, MNN_ERR_UNSUPPORTED, "tensor map: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mnn_set_error("tensor map (bf16): cuTensorMapEncodeTiled failed");
    return MNN_ERR_ARG;
  }
  return MNN_OK;
}

// 2-D bf16 tensor map over a row-major matrix [outer = k][inner = mn], box {64 mn, 32 k}, SWIZZLE_128B: a box lands as the
// UMMA canonical MN-major SWIZZLE_128B block of 64 mn x 32 k (128-byte rows = 64 mn, 8-row atoms of 1 KB = the k-groups)
static int make_map_bf16_mn(const void* ptr, long long ld_elems, long long inner, long long outer, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode();
  MNN_REQUIRE(enc != nullptr, MNN_ERR_UNSUPPORTED, "tensor map: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {64u, 32u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mnn_set_error("tensor map (bf16, MN-major): cuTensorMapEncodeTiled failed");
    return MNN_ERR_ARG;
  }
  return MNN_OK;
}

static int device_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
// SM budget of the calling thread's persistent GEMM grids (mnn_set_sm_budget): lets a GEMM run beside co-resident
// persistent recurrence kernels of other streams instead of queueing CTAs behind them
static thread_local int g_sm_budget = 0;
static thread_local int g_gemm_split = 0;   // mnn_set_gemm_split: 0 = 2.5-product scheme, 1 = bf16 pairs (pair kernel only)
static int num_sms() {
  const int n = device_sms();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
  using C_ = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM);
    attr_set = true;
  }
  const int items = p.tiles_m * p.tiles_n * p.splits;
  const int grid = items < num_sms() ? items : num_sms();
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kThreads, C_::SMEM, stream>>>(ma, mb, p);
  return mnn_check_launch("gemm_tc");
}

static int num_clusters2(const void* fn) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * device_sms());
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg2::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = device_sms() / 2; }
  return n;
}

template <bool A_MN, bool B_MN>
static int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const CUtensorMap& mc, const Params& p,
                   cudaStream_t stream) {
  static int max_clusters = 0;
  if (!max_clusters) {
    cudaFuncSetAttribute(gemm_tc2_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2::SMEM);
    max_clusters = num_clusters2(reinterpret_cast<const void*>(gemm_tc2_kernel<A_MN, B_MN>));
  }
  const int items = p.tiles_m * p.tiles_n * p.splits;
  int clusters = items < max_clusters ? items : max_clusters;
  if (clusters > num_sms() / 2) clusters = num_sms() / 2 > 0 ? num_sms() / 2 : 1;
  gemm_tc2_kernel<A_MN, B_MN><<<2 * clusters, kThreads, Cfg2::SMEM, stream>>>(ma, mb, mb2, mc, p);
  return mnn_check_launch("gemm_tc2");
}

static int dispatch_major2(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2,
                           const CUtensorMap& mc, const Params& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch2<false, false>(ma, mb, mb2, mc, p, s);
  if (!a_mn && b_mn) return launch2<false, true>(ma, mb, mb2, mc, p, s);
  if (a_mn && !b_mn) return launch2<true, false>(ma, mb, mb2, mc, p, s);
  return launch2<true, true>(ma, mb, mb2, mc, p, s);
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const Params& p,
                          cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(ma, mb, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true>(ma, mb, p, s);
  if (a_mn && !b_mn) return launch<BN, true, false>(ma, mb, p, s);
  return launch<BN, true, true>(ma, mb, p, s);
}

}  // namespace tc
}  // namespace mnn

using namespace mnn;

int mnn_tc_make_map(const float* ptr, long long ld, long long inner, long long outer, int box_outer, bool mn_major,
                    CUtensorMap* out) {
  return mnn::tc::make_map(ptr, ld, inner, outer, box_outer, mn_major, out);
}
int mnn_tc_make_map_plain(const float* ptr, long long ld, long long inner, long long outer, int box_inner, int box_outer,
                           CUtensorMap* out) {
  return mnn::tc::make_map_plain(ptr, ld, inner, outer, box_inner, box_outer, out);
}
int mnn_tc_make_map_bf16(const void* ptr, long long ld_elems, long long inner, long long outer, int box_outer,
                         CUtensorMap* out) {
  return mnn::tc::make_map_bf16(ptr, ld_elems, inner, outer, box_outer, out);
}
int mnn_tc_num_sms() { return mnn::tc::device_sms(); }

int mnn_tc_sm_budget() { return mnn::tc::g_sm_budget; }
int mnn_tc_gemm_split() { return mnn::tc::g_gemm_split; }   // mnn_set_gemm_split of the calling thread

extern "C" int mnn_set_sm_budget(int sms) {
  mnn::tc::g_sm_budget = sms > 0 ? sms : 0;
  return MNN_OK;
}

extern "C" int mnn_set_gemm_split(int mode) {
  MNN_REQUIRE(mode == 0 || mode == 1, MNN_ERR_ARG, "set_gemm_split: mode must be 0 (2.5 products) or 1 (bf16 pairs)");
  mnn::tc::g_gemm_split = mode;
  return MNN_OK;
}

extern "C" int mnn_gemm_tc_supported(const float* A, long long lda, const float* B, long long ldb) {
  return ((lda & 3) == 0) && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
}

// bpair != nullptr: B is given as bf16 pair planes (mnn_split_bf16_pair) with row stride ldb ELEMENTS; B itself is unused
// a16 != nullptr: A is given as ONE exact bf16 plane (binary rows) with row stride lda ELEMENTS; A itself is unused
static int gemm_tc_impl(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                        long long ldc, const float* bias, float alpha, float beta, int M, int N, int K, int a_exact,
                        const void* bpair, cudaStream_t stream, const void* a16 = nullptr) {
  using namespace mnn::tc;
  const bool a_mn = transA != 0;   // A stored [K,M]: M contiguous
  const bool b_mn = transB == 0;   // B stored [K,N]: N contiguous
  static const bool force_1cta = getenv("MNN_GEMM_1CTA") != nullptr;
  static const char* kmin_env = getenv("MNN_GEMM_PAIR_KMIN");
  static const int pair_kmin = kmin_env ? atoi(kmin_env) : 64;
  // 256x256 tiles on CTA pairs; with the TMA-store epilogue they win down to K = 256 (Dense forward: 2.28 vs 3.67 ms)
  const bool pair = !force_1cta && M >= 256 && N > 128 && K >= pair_kmin;
  MNN_REQUIRE((!bpair && !a16) || pair, MNN_ERR_UNSUPPORTED,
              "gemm_tc: pre-split operands need the CTA-pair kernel (M >= 256, N > 128, K >= 64)");
  const int BN = pair ? BN2 : (N > 64 ? 128 : 64);
  const int TM = pair ? 2 * BM : BM;
  const int units = pair ? num_sms() / 2 : num_sms();

  Params p{};
  p.C = C; p.bias = bias; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.M = M; p.N = N; p.K = K;
  p.tiles_m = (M + TM - 1) / TM;
  p.tiles_n = (N + BN - 1) / BN;
  p.kb_total = (K + BK - 1) / BK;
  p.n_products = a_exact ? 2 : 3;
  const int tiles = p.tiles_m * p.tiles_n;
  int splits = 1;
  if (tiles * 2 <= units && p.kb_total >= 32) {
    splits = units / tiles;
    const int maxs = p.kb_total / 8;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.atomic = p.splits > 1;
  if (p.atomic) {
    // partial sums are red.add-ed into C: bring C to beta*C first (beta is 0 or 1 at every call site)
    MNN_REQUIRE(beta == 0.f || beta == 1.f, MNN_ERR_UNSUPPORTED, "gemm_tc: split-K needs beta in {0,1}");
    if (beta == 0.f) cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream);
  }

  CUtensorMap ma, mb;
  int rc;
  const int box_n = pair ? BN2 / 2 : BN;
  if (a16) {
    if (!a_mn) rc = make_map_bf16(a16, lda, K, M, BM, &ma);  // [M rows][K]   box {32 k, 128 rows}, SWIZZLE_64B
    else rc = make_map_bf16_mn(a16, lda, M, K, &ma);         // [K rows][M]   box {64 m, 32 k}, SWIZZLE_128B
  } else if (!a_mn) rc = make_map(A, lda, K, M, BM, false, &ma);    // [M rows][K]   box {32 k, 128 rows}
  else rc = make_map(A, lda, M, K, BK, true, &ma);           // [K rows][M]   box {32 m, 32 k}
  if (rc) return rc;
  CUtensorMap mb2;
  if (bpair) {
    // planes: hi at bpair, lo right behind it; [N rows][K] (K-major) or [K rows][N] (MN-major), row stride ldb elements
    const long long rows_b = b_mn ? K : N;
    const char* lo = static_cast<const char*>(bpair) + (size_t)rows_b * ldb * 2;
    if (!b_mn) {
      rc = make_map_bf16(bpair, ldb, K, N, box_n, &mb);
      if (!rc) rc = make_map_bf16(lo, ldb, K, N, box_n, &mb2);
    } else {
      rc = make_map_bf16_mn(bpair, ldb, N, K, &mb);
      if (!rc) rc = make_map_bf16_mn(lo, ldb, N, K, &mb2);
    }
  } else {
    if (!b_mn) rc = make_map(B, ldb, K, N, box_n, false, &mb); // [N rows][K]   box {32 k, BN (or BN/2 per CTA) rows}
    else rc = make_map(B, ldb, N, K, BK, true, &mb);           // [K rows][N]   box {32 n, 32 k}
    mb2 = mb;
  }
  if (rc) return rc;

  if (pair) {
    // cross terms as bf16 MMAs ("2.5 products"): on for general A; with a binary A (2 tf32 products) the extra bf16(A)
    // tile makes the converter warps the bottleneck (measured 3.15 -> 3.33 ms). MNN_GEMM_BF16X=0 / 1 forces it off / on.
    static const char* bf16x_env = getenv("MNN_GEMM_BF16X");
    p.bf16x = bf16x_env ? (bf16x_env[0] == '2' ? 2 : (bf16x_env[0] == '1' ? 1 : 0))
                        : (g_gemm_split == 1 ? 2 : (p.n_products == 3 ? 1 : 0));
    static const char* share_env = getenv("MNN_GEMM_SHARE_KB");   // k-blocks per tile from which the epilogue warps convert too
    static const int share_kb = share_env ? atoi(share_env) : 8;
    // only with the bf16 tiles: in the tf32 path with a binary A just B is converted and the extra arrivals cost more
    // than the shared work saves (dW1x 3.16 -> 3.71 ms)
    if (bpair) { p.bf16x = 2; p.b_pre = 1; }   // the pre-split planes ARE the bf16 pair
    if (a16) { p.bf16x = 2; p.a_pre = 1; p.n_products = 2; }
    p.share_conv = (p.bf16x && share_kb > 0 && p.kb_per_split >= share_kb) ? 1 : 0;
    CUtensorMap mc = ma;
    p.tma_store = (!p.atomic && beta == 0.f && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0;
    if (p.tma_store) {
      rc = make_map(C, ldc, N, M, BM, false, &mc);             // [M rows][N]   box {32 n, 128 rows}
      if (rc) return rc;
    }
    return dispatch_major2(a_mn, b_mn, ma, mb, mb2, mc, p, stream);
  }
  if (BN == 128) return dispatch_major<128>(a_mn, b_mn, ma, mb, p, stream);
  return dispatch_major<64>(a_mn, b_mn, ma, mb, p, stream);
}

extern "C" int mnn_gemm_tc(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                           long long ldc, const float* bias, float alpha, float beta, int M, int N, int K, int a_exact,
                           cudaStream_t stream) {
  MNN_REQUIRE(A && B && C, MNN_ERR_ARG, "gemm_tc: null pointer");
  MNN_REQUIRE(M > 0 && N > 0 && K > 0, MNN_ERR_ARG, "gemm_tc: non-positive size");
  MNN_REQUIRE(mnn_gemm_tc_supported(A, lda, B, ldb), MNN_ERR_UNSUPPORTED,
              "gemm_tc: TMA needs 16-byte aligned operand pointers and row strides that are multiples of 4 floats");
  return gemm_tc_impl(A, lda, transA, B, ldb, transB, C, ldc, bias, alpha, beta, M, N, K, a_exact, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------ pre-split weights
namespace mnn {
__global__ void split_bf16_pair_kernel(const float* __restrict__ src, long long ld, int rows, int cols, uint16_t* __restrict__ dst,
                                       long long ldp) {
  const long long n = (long long)rows * ldp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ldp), c = (int)(i - (long long)r * ldp);
    const float x = c < cols ? __ldg(src + (size_t)r * ld + c) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    dst[i] = __bfloat16_as_ushort(hi);
    dst[n + i] = __bfloat16_as_ushort(lo);
  }
}
}  // namespace mnn

extern "C" int mnn_split_bf16_pair(const float* src, long long ld, int rows, int cols, void* dst, long long ld_elems,
                                   cudaStream_t stream) {
  MNN_REQUIRE(src && dst && rows > 0 && cols > 0, MNN_ERR_ARG, "split_bf16_pair: bad argument");
  MNN_REQUIRE(ld >= cols && ld_elems >= cols && (ld_elems & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              MNN_ERR_ARG, "split_bf16_pair: ld_elems must cover cols and be a multiple of 8, dst 16-byte aligned");
  const long long n = (long long)rows * ld_elems;
  const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  mnn::split_bf16_pair_kernel<<<blocks, 256, 0, stream>>>(src, ld, rows, cols, static_cast<uint16_t*>(dst), ld_elems);
  return mnn_check_launch("split_bf16_pair");
}

extern "C" int mnn_gemm_tc_bpair(const float* A, long long lda, int transA, const void* Bpair, long long ldb_elems, int transB,
                                 float* C, long long ldc, const float* bias, float alpha, float beta, int M, int N, int K,
                                 int a_exact, cudaStream_t stream) {
  MNN_REQUIRE(A && Bpair && C, MNN_ERR_ARG, "gemm_tc_bpair: null pointer");
  MNN_REQUIRE(M > 0 && N > 0 && K > 0, MNN_ERR_ARG, "gemm_tc_bpair: non-positive size");
  MNN_REQUIRE((lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (ldb_elems & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(Bpair) & 15) == 0,
              MNN_ERR_UNSUPPORTED, "gemm_tc_bpair: TMA needs 16-byte aligned operands and row strides");
  return gemm_tc_impl(A, lda, transA, nullptr, ldb_elems, transB, C, ldc, bias, alpha, beta, M, N, K, a_exact, Bpair, stream);
}

extern "C" int mnn_gemm_tc_abf16(const void* A16, long long lda_elems, int transA, const float* B, long long ldb,
                                 const void* Bpair, long long ldb_elems, int transB, float* C, long long ldc,
                                 const float* bias, float alpha, float beta, int M, int N, int K, cudaStream_t stream) {
  MNN_REQUIRE(A16 && (B || Bpair) && C, MNN_ERR_ARG, "gemm_tc_abf16: null pointer");
  MNN_REQUIRE(M > 0 && N > 0 && K > 0, MNN_ERR_ARG, "gemm_tc_abf16: non-positive size");
  MNN_REQUIRE((lda_elems & 7) == 0 && (reinterpret_cast<uintptr_t>(A16) & 15) == 0, MNN_ERR_UNSUPPORTED,
              "gemm_tc_abf16: the bf16 A plane needs a 16-byte aligned pointer and a row stride that is a multiple of 8");
  if (Bpair) {
    MNN_REQUIRE((ldb_elems & 7) == 0 && (reinterpret_cast<uintptr_t>(Bpair) & 15) == 0, MNN_ERR_UNSUPPORTED,
                "gemm_tc_abf16: the B planes need a 16-byte aligned pointer and a row stride that is a multiple of 8");
    return gemm_tc_impl(nullptr, lda_elems, transA, nullptr, ldb_elems, transB, C, ldc, bias, alpha, beta, M, N, K, 1, Bpair,
                        stream, A16);
  }
  MNN_REQUIRE((ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, MNN_ERR_UNSUPPORTED,
              "gemm_tc_abf16: B needs a 16-byte aligned pointer and a row stride that is a multiple of 4 floats");
  return gemm_tc_impl(nullptr, lda_elems, transA, B, ldb, transB, C, ldc, bias, alpha, beta, M, N, K, 1, nullptr, stream, A16);
}

// stacked piano-roll rows as an exact bf16 plane: xin16[(t + 1) * B + b][0 .. D*M) = x[b][t][.][.] (feature d*M + m is the
// memory order of x), slot t = 0 and the pad columns zero -- the A operand of the layer-0 projection and weight-gradient
// GEMMs of the training step (core/multi_encoder_nn.py:66-76 + multinn_composer.py:73-80)
namespace mnn {
__device__ __forceinline__ uint32_t bf16x2_of(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// one thread = 8 consecutive features of one row: a 16-byte store, 8 input bytes (uint8) or two float4 loads (I % 4 == 0
// keeps them aligned; other I take the scalar tail path)
template <typename TIn>
__global__ void pack_stacked_bf16_kernel(const TIn* __restrict__ x, uint16_t* __restrict__ out, int ld16, int B, int T, int I) {
  const int chunks = ld16 >> 3;
  const long long n = (long long)(T + 1) * B * chunks;
  const bool aligned = (I & 3) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / chunks), ch = (int)(i - (long long)row * chunks);
    const int t1 = row / B, b = row - t1 * B, c0 = ch * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t1 > 0 && c0 < I) {
      const TIn* src = x + ((size_t)b * T + (t1 - 1)) * I + c0;
      if (aligned && c0 + 8 <= I) {
        if (sizeof(TIn) == 1) {
          const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(src)), w1 = __ldg(reinterpret_cast<const uint32_t*>(src) + 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[j] = (float)((w0 >> (8 * j)) & 0xffu); v[4 + j] = (float)((w1 >> (8 * j)) & 0xffu); }
        } else {
          const float4 f0 = __ldg(reinterpret_cast<const float4*>(src)), f1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w; v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < I) v[j] = (float)src[j];
      }
    }
    *reinterpret_cast<uint4*>(out + (size_t)row * ld16 + c0) =
        make_uint4(bf16x2_of(v[0], v[1]), bf16x2_of(v[2], v[3]), bf16x2_of(v[4], v[5]), bf16x2_of(v[6], v[7]));
  }
}
}  // namespace mnn

extern "C" int mnn_pack_stacked_bf16(const void* x, int x_is_u8, void* xin16, long long ld16, int B, int T, int I,
                                     cudaStream_t stream) {
  MNN_REQUIRE(x && xin16 && B > 0 && T > 0 && I > 0, MNN_ERR_ARG, "pack_stacked_bf16: bad argument");
  MNN_REQUIRE(ld16 >= I && (ld16 & 7) == 0 && (reinterpret_cast<uintptr_t>(xin16) & 15) == 0, MNN_ERR_ARG,
              "pack_stacked_bf16: ld16 must cover the row and be a multiple of 8, xin16 16-byte aligned");
  MNN_REQUIRE(ld16 < (1ll << 30) && (long long)(T + 1) * B < (1ll << 31), MNN_ERR_ARG, "pack_stacked_bf16: too many rows");
  const long long n = (long long)(T + 1) * B * (ld16 / 8);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (x_is_u8)
    mnn::pack_stacked_bf16_kernel<uint8_t><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(x), static_cast<uint16_t*>(xin16),
                                                                     (int)ld16, B, T, I);
  else
    mnn::pack_stacked_bf16_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<uint16_t*>(xin16),
                                                                   (int)ld16, B, T, I);
  return mnn_check_launch("pack_stacked_bf16");
}
