// Weight-resident persistent LSTM forward recurrence for small per-GPU batches (the data-parallel shards: B <= 512).
// Same semantics as lstm_tc.cu (CudnnCompatibleLSTMCell = LSTMBlockCell with forget_bias 0, gate blocks i, j, f, o,
// DropoutWrapper on the output; reference common/rnn.py:104-145, driven like dynamic_decode at
// generators/rnn_nade.py:204-218), other schedule -- BASELINE north_star (1): "a persistent kernel that keeps the
// recurrent weights in shared memory across all T steps".
//
// Why: at B = 256 the step of lstm_tc_kernel<.., SHARE> is a 17 us latency chain (trace in DESIGN.md): every step
// re-streams the CTA's W_h slice from L2 and re-splits both operands into hi/lo tiles in shared memory before the MMAs
// can start. Here
//   * work item = (128-row batch slab, UB = 8 or 16 hidden units = 4 UB gate columns), ONE item per CTA for the whole
//     launch, all CTAs co-resident (cooperative launch); the item's W_h slice [R x 4 UB] is split ONCE per launch into
//     bf16 w1 + w2 and stays in shared memory in the UMMA canonical K-major SWIZZLE_64B layout (64 KB at R = 512, UB = 8);
//   * h_t is PUBLISHED already split (bf16 h1 + h2, a double-buffered [2][B][R] pair in the workspace) by the epilogue
//     that produces it, so a step's mainloop is TMA (SWIZZLE_64B boxes land in the canonical layout) -> tcgen05.mma,
//     with no converter stage; the whole h slab of a step (R/32 k-blocks x 16 KB) streams through a 6-8 stage ring;
//   * h.W ~ h1.w1 (main accumulator) + h1.w2 + h2.w1 (second accumulator): fp32 accumulation in TMEM, dropped terms
//     ~2^-17 |h w| (the LSTM tests' 1e-5 bound holds);
//   * the cell epilogue keeps c_t in REGISTERS across steps (a thread owns (row, 8 units) for the whole launch) and
//     loads the step's input-projection pre-activations before it waits for the accumulator;
//   * CTAs of a slab hand h_t over through a per-slab counter (red.release.gpu / ld.acquire.gpu), nothing is grid-wide.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "multinn_b200.h"
#include "tc_common.cuh"

int mnn_tc_make_map_bf16(const void* ptr, long long ld_elems, long long inner, long long outer, int box_outer,
                         CUtensorMap* out);   // gemm_tc.cu: bf16 [outer][inner], box {32, box_outer}, SWIZZLE_64B

namespace mnn {
namespace res {

using namespace mnn::tc;

constexpr int kThreads = 384;
constexpr int kMaxStages = 8;
constexpr int A_STAGE = 2 * 128 * 64;      // h1 + h2 tiles of one k-block

struct RParams {
  float* gates; float* hbuf; float* cbuf; float* out; float* dscale; const float* u;
  const float* wh;                 // [R][4R]
  __nv_bfloat16* hs1; __nv_bfloat16* hs2;   // [2][B][R] split h, slot s & 1 = state before local step s
  float keep; unsigned long long seed; RowMap rmap;
  int T, B, R, slabs, blocks, kb_total, stages;
  unsigned int* flags;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t sw64_off(int row, int k0) {
  const int r8 = row & 7;
  return (uint32_t)((row >> 3) * 512 + r8 * 64 + (((k0 >> 3) ^ (r8 >> 1)) << 4) + ((k0 & 7) << 1));
}
__device__ __forceinline__ void split_bf16x4(float a, float b, float c, float d, uint2& hi, uint2& lo) {
  hi = pack_bf16x4(a, b, c, d);
  lo = pack_bf16x4(a - __uint_as_float(hi.x << 16), b - __uint_as_float(hi.x & 0xffff0000u),
                   c - __uint_as_float(hi.y << 16), d - __uint_as_float(hi.y & 0xffff0000u));
}

// hs1/hs2 slot 0 <- split(h0): the initial state of the launch
__global__ void split_h_kernel(const float* __restrict__ h, __nv_bfloat16* __restrict__ h1, __nv_bfloat16* __restrict__ h2,
                               size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(h)[i];
    uint2 hi, lo;
    split_bf16x4(x.x, x.y, x.z, x.w, hi, lo);
    reinterpret_cast<uint2*>(h1)[i] = hi;
    reinterpret_cast<uint2*>(h2)[i] = lo;
  }
}

template <int UB>
__global__ void __launch_bounds__(kThreads, 1)
lstm_res_fwd_kernel(const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2, const RParams p) {
  constexpr int BN = 4 * UB;                 // gate columns of the item = MMA N
  constexpr int B_TILE = BN * 64;            // [BN x 32] bf16, K-major SWIZZLE_64B
  constexpr int EPI_WARPS = 4 * (UB / 8);
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R, B = p.B, KB = p.kb_total, STAGES = p.stages;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages, bar_tempty = bar_tfull + 8;
  const uint32_t off_b2 = (uint32_t)KB * B_TILE, off_a = 2u * KB * B_TILE;   // [w1 tiles | w2 tiles | A ring]

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int n_blk = blockIdx.x % p.blocks, m_blk = blockIdx.x / p.blocks;
  const int unit0 = n_blk * UB;

  // resident B operand: element (n = g UB + ul, k) = Wh[k][g R + unit0 + ul], split into bf16 w1 + w2
  for (int idx = threadIdx.x; idx < BN * (R / 4); idx += kThreads) {
    const int n = idx % BN, k0 = (idx / BN) * 4;
    const int col = (n / UB) * R + unit0 + (n % UB);
    const float* w = p.wh + (size_t)k0 * 4 * R + col;
    uint2 hi, lo;
    split_bf16x4(__ldg(w), __ldg(w + 4 * R), __ldg(w + 8 * R), __ldg(w + 12 * R), hi, lo);
    const uint32_t off = (uint32_t)((k0 >> 5) * B_TILE) + sw64_off(n, k0 & 31);
    *reinterpret_cast<uint2*>(smem + off) = hi;
    *reinterpret_cast<uint2*>(smem + off_b2 + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: the slab's h1 / h2 tiles of every step
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.T; ++s) {
        if (s > 0) {
          const unsigned int target = (unsigned int)s * (unsigned int)p.blocks;
          while (ld_acquire_u32(p.flags + m_blk) < target) __nanosleep(20);
          asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic-proxy stores -> our TMA reads
        }
        const int row0 = (s & 1) * B + m_blk * BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, A_STAGE);
          const uint32_t dst = smem0 + off_a + stage * A_STAGE;
          tma_load_2d(dst, &map_h1, full, kb * 32, row0);
          tma_load_2d(dst + A_STAGE / 2, &map_h2, full, kb * 32, row0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BN, false, false, 128);
      const uint32_t d_main = tmem_base, d_aux = tmem_base + BN;
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.T; ++s) {
        if (s > 0) {
          mbar_wait(bar_tempty, (uint32_t)(s - 1) & 1u);   // the previous step's epilogue has drained the accumulators
          tc_fence_after();
        }
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a1 = smem0 + off_a + stage * A_STAGE, a2 = a1 + A_STAGE / 2;
          const uint32_t b1 = smem0 + kb * B_TILE, b2 = smem0 + off_b2 + kb * B_TILE;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da1 = smem_desc(a1 + j * 32, 16, 512, 4), da2 = smem_desc(a2 + j * 32, 16, 512, 4);
            const uint64_t db1 = smem_desc(b1 + j * 32, 16, 512, 4), db2 = smem_desc(b2 + j * 32, 16, 512, 4);
            const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
            umma_bf16(d_main, da1, db1, idesc, first);
            umma_bf16(d_aux, da1, db2, idesc, first);
            umma_bf16(d_aux, da2, db1, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull);
      }
    }
  } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
    // ------------------------------------------------------------------ cell epilogue: thread = (row, 8 units)
    const int we = warp - 4, q = we & 3, ug = we >> 2;
    const int b = m_blk * BM + q * 32 + lane;
    const bool row_ok = b < B;
    const int unit = unit0 + ug * 8;
    const size_t BR = (size_t)B * R;
    const size_t ridx = (size_t)(row_ok ? b : 0) * R + unit;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    const float inv_keep = 1.0f / p.keep;
    float cst[8];
    {
      const float4 c0 = *reinterpret_cast<const float4*>(p.cbuf + ridx);
      const float4 c1 = *reinterpret_cast<const float4*>(p.cbuf + ridx + 4);
      cst[0] = c0.x; cst[1] = c0.y; cst[2] = c0.z; cst[3] = c0.w; cst[4] = c1.x; cst[5] = c1.y; cst[6] = c1.z; cst[7] = c1.w;
    }
    for (int s = 0; s < p.T; ++s) {
      float* gp = p.gates + ((size_t)s * B + (row_ok ? b : 0)) * 4 * R + unit;
      // the input-projection pre-activations (written before the launch) and the uniforms do not depend on h_{t-1}
      float pre[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 p0 = *reinterpret_cast<const float4*>(gp + g * R);
        const float4 p1 = *reinterpret_cast<const float4*>(gp + g * R + 4);
        pre[g][0] = p0.x; pre[g][1] = p0.y; pre[g][2] = p0.z; pre[g][3] = p0.w;
        pre[g][4] = p1.x; pre[g][5] = p1.y; pre[g][6] = p1.z; pre[g][7] = p1.w;
      }
      float dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = 1.0f;
      if (p.out && p.keep < 1.0f) {
        float uu[8];
        if (p.u) {
          const float4 u0 = __ldg(reinterpret_cast<const float4*>(p.u + (size_t)s * BR + ridx));
          const float4 u1 = __ldg(reinterpret_cast<const float4*>(p.u + (size_t)s * BR + ridx + 4));
          uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w; uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
        } else {
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const uint4 r4 = dropout_bits4(p.seed, p.rmap, b, unit + 4 * h4, R, s);
            uu[4 * h4] = u01(r4.x); uu[4 * h4 + 1] = u01(r4.y); uu[4 * h4 + 2] = u01(r4.z); uu[4 * h4 + 3] = u01(r4.w);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) dv[i] = floorf(p.keep + uu[i]) * inv_keep;   // tf.nn.dropout: x / keep * floor(keep + u)
      }
      mbar_wait(bar_tfull, (uint32_t)s & 1u);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float m8[8], x8[8];
        tmem_ld8(tacc + (uint32_t)(g * UB + ug * 8), m8);
        tmem_ld8(tacc + (uint32_t)(BN + g * UB + ug * 8), x8);
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[g][i] += m8[i] + x8[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gi, gj, gf, go, cn;
        lstm_cell_xu(pre[0][i], pre[1][i], pre[2][i], pre[3][i], cst[i], gi, gj, gf, go, cn, hv[i]);
        pre[0][i] = gi; pre[1][i] = gj; pre[2][i] = gf; pre[3][i] = go;
        cst[i] = cn;
      }
      if (row_ok) {
        // h_t first: it gates the next step of the whole slab
        uint2 hi0, lo0, hi1, lo1;
        split_bf16x4(hv[0], hv[1], hv[2], hv[3], hi0, lo0);
        split_bf16x4(hv[4], hv[5], hv[6], hv[7], hi1, lo1);
        const size_t so = (size_t)((s + 1) & 1) * BR + ridx;
        *reinterpret_cast<uint4*>(p.hs1 + so) = make_uint4(hi0.x, hi0.y, hi1.x, hi1.y);
        *reinterpret_cast<uint4*>(p.hs2 + so) = make_uint4(lo0.x, lo0.y, lo1.x, lo1.y);
      }
      // bar.sync orders every epilogue thread's h stores before thread 0's release (cumulative at gpu scope): no
      // per-thread __threadfence on the critical path
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (we == 0 && lane == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.flags + m_blk), "r"(1u) : "memory");
      if (row_ok) {
        // everything else (saved for BPTT / read by the next layer) leaves after the flag
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<float4*>(gp + g * R) = make_float4(pre[g][0], pre[g][1], pre[g][2], pre[g][3]);
          *reinterpret_cast<float4*>(gp + g * R + 4) = make_float4(pre[g][4], pre[g][5], pre[g][6], pre[g][7]);
        }
        float* cn = p.cbuf + (size_t)(s + 1) * BR + ridx;
        *reinterpret_cast<float4*>(cn) = make_float4(cst[0], cst[1], cst[2], cst[3]);
        *reinterpret_cast<float4*>(cn + 4) = make_float4(cst[4], cst[5], cst[6], cst[7]);
        float* hn = p.hbuf + (size_t)(s + 1) * BR + ridx;
        *reinterpret_cast<float4*>(hn) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4*>(hn + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        if (p.out) {
          float* op = p.out + (size_t)s * BR + ridx;
          *reinterpret_cast<float4*>(op) = make_float4(hv[0] * dv[0], hv[1] * dv[1], hv[2] * dv[2], hv[3] * dv[3]);
          *reinterpret_cast<float4*>(op + 4) = make_float4(hv[4] * dv[4], hv[5] * dv[5], hv[6] * dv[6], hv[7] * dv[7]);
          if (p.keep < 1.0f) {
            float* dsp = p.dscale + (size_t)s * BR + ridx;
            *reinterpret_cast<float4*>(dsp) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            *reinterpret_cast<float4*>(dsp + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// smallest unit block whose items fit the SM budget and whose resident weights leave room for >= 4 ring stages
static int pick_ub(int B, int R, int sms, int* stages) {
  if (R % 32 != 0 || B > 512) return 0;
  const int slabs = (B + BM - 1) / BM;
  for (int UB : {8, 16}) {
    if (R % UB) continue;
    const int items = slabs * (R / UB);
    const size_t resident = (size_t)2 * (R / 32) * (4 * UB * 64);
    const long long room = (long long)224 * 1024 - (long long)resident;
    if (items <= sms && room >= 4LL * A_STAGE) {
      int st = (int)(room / A_STAGE);
      if (st > kMaxStages) st = kMaxStages;
      if (st > R / 32) st = R / 32;
      *stages = st;
      return UB;
    }
  }
  return 0;
}

}  // namespace res
}  // namespace mnn

using namespace mnn;

// CTAs the weight-resident kernel would occupy for this shape under `sms` SMs (0: shape not taken)
int mnn_lstm_res_ctas(int T, int B, int R, int sms) {
  static const char* env = getenv("MNN_LSTM_RES");   // "0": never
  if ((env && env[0] == '0') || T < 2) return 0;
  int st = 0;
  const int UB = res::pick_ub(B, R, sms, &st);
  return UB ? ((B + tc::BM - 1) / tc::BM) * (R / UB) : 0;
}

// forward: split h double buffer (8 B R bytes); BPTT: split dG double buffer (32 B R bytes); never live at the same time
size_t mnn_lstm_res_workspace_bytes(int B, int R) { return (size_t)32 * B * R + 256; }

// hs: mnn_lstm_res_workspace_bytes(B, R) bytes, 256-byte aligned; flags: >= slabs counters
int mnn_lstm_res_fwd(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale, const float* u,
                     float keep, unsigned long long seed, int T, int B, int R, void* hs, unsigned int* flags, int sms,
                     cudaStream_t stream) {
  int stages = 0;
  const int UB = res::pick_ub(B, R, sms, &stages);
  MNN_REQUIRE(UB != 0, MNN_ERR_UNSUPPORTED, "lstm_res_fwd: shape not taken");
  res::RParams p{};
  p.gates = gates; p.hbuf = hbuf; p.cbuf = cbuf; p.out = out; p.dscale = dscale; p.u = u; p.wh = wh;
  p.hs1 = reinterpret_cast<__nv_bfloat16*>(hs);
  p.hs2 = p.hs1 + (size_t)2 * B * R;
  p.keep = keep; p.seed = seed; p.rmap = current_row_map();
  p.T = T; p.B = B; p.R = R; p.slabs = (B + tc::BM - 1) / tc::BM; p.blocks = R / UB; p.kb_total = R / 32; p.stages = stages;
  p.flags = flags;
  CUtensorMap m1, m2;
  int rc = mnn_tc_make_map_bf16(p.hs1, R, R, 2LL * B, tc::BM, &m1);
  if (rc) return rc;
  rc = mnn_tc_make_map_bf16(p.hs2, R, R, 2LL * B, tc::BM, &m2);
  if (rc) return rc;
  cudaMemsetAsync(flags, 0, (size_t)p.slabs * sizeof(unsigned int), stream);
  const size_t n4 = (size_t)B * R / 4;
  res::split_h_kernel<<<(unsigned)((n4 + 255) / 256 < 296 ? (n4 + 255) / 256 : 296), 256, 0, stream>>>(hbuf, p.hs1, p.hs2, n4);
  rc = mnn_check_launch("lstm_res split_h");
  if (rc) return rc;
  const size_t smem = (size_t)2 * p.kb_total * (4 * UB * 64) + (size_t)stages * res::A_STAGE + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.slabs * p.blocks);
  cfg.blockDim = dim3(res::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e;
  if (UB == 8) {
    cudaFuncSetAttribute(res::lstm_res_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaLaunchKernelEx(&cfg, res::lstm_res_fwd_kernel<8>, m1, m2, p);
  } else {
    cudaFuncSetAttribute(res::lstm_res_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaLaunchKernelEx(&cfg, res::lstm_res_fwd_kernel<16>, m1, m2, p);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();   // not sticky: clear it so that the next launch check does not report it again
    mnn_set_error(cudaGetErrorString(e));
    return (int)e;
  }
  return mnn_check_launch("lstm_seq_fwd(resident)");
}

// ================================================================================================ BPTT
// Weight-resident persistent BPTT for the same small-batch regime. dh_t = dG_{t+1} . Wh^T has K = 4R: a CTA that owned
// whole K would stream 1 MB of dG per step. Instead a CLUSTER OF 4 CTAs owns (128-row slab, 32 units): CTA kq keeps the
// K-slice Wh[32 units][kq R .. (kq+1) R) resident (bf16 w1 + w2, 64 KB at R = 512), streams only its quarter of the slab's
// dG_{t+1} (published pre-split in bf16 by the step before, like h in the forward kernel) and produces a PARTIAL
// dh[128 x 32] in TMEM; the partials are exchanged through distributed shared memory (st.shared::cluster + a remote
// mbarrier arrive with release.cluster): CTA d receives the 8 unit columns it owns from its three peers, adds its own,
// and runs the cell backward for (row, 8 units) with the cell-gradient carry dc in REGISTERS across steps.
namespace mnn {
namespace res {

struct BParams {
  float* gates; const float* cbuf; const float* dout; const float* dscale; float* dc;
  const float* wh;
  __nv_bfloat16* gs1; __nv_bfloat16* gs2;   // [2][B][4R] split dG, slot s & 1 = dG_{t+1} of local step s
  int B, R, t_top, n_steps, slabs, blocks, kb_local, stages;
  unsigned int* flags;
};

constexpr int kSplit = 4;                  // cluster size = K split
constexpr int kBU = 32;                    // units per cluster = MMA N
constexpr int XBUF = 3 * 128 * 32;         // one parity: [3 peers][128 rows][8 floats]

__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void split_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ x1, __nv_bfloat16* __restrict__ x2,
                                  size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 hi, lo;
    split_bf16x4(v.x, v.y, v.z, v.w, hi, lo);
    reinterpret_cast<uint2*>(x1)[i] = hi;
    reinterpret_cast<uint2*>(x2)[i] = lo;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_res_bwd_kernel(const __grid_constant__ CUtensorMap map_g1, const __grid_constant__ CUtensorMap map_g2, const BParams p) {
  constexpr int BN = kBU, B_TILE = BN * 64, TMEM_COLS = 64;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 3];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R, B = p.B, KB = p.kb_local, STAGES = p.stages, KS = R;   // K-slice of this CTA: [kq R, (kq+1) R)
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages, bar_tempty = bar_tfull + 8, bar_x = bar_tempty + 8;
  const uint32_t off_b2 = (uint32_t)KB * B_TILE, off_x = 2u * KB * B_TILE, off_a = off_x + 2 * XBUF * 4;

  const int kq = (int)cluster_ctarank();
  const int cid = blockIdx.x / kSplit;
  const int n_blk = cid % p.blocks, m_blk = cid / p.blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    mbar_init(bar_x, 3 * 4);                 // 3 peers x 4 epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident B operand: element (n, k) = Wh[n_blk 32 + n][kq R + k]
  for (int idx = threadIdx.x; idx < BN * (KS / 4); idx += kThreads) {
    const int n = idx / (KS / 4), k0 = (idx - n * (KS / 4)) * 4;
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.wh + (size_t)(n_blk * kBU + n) * 4 * R + (size_t)kq * KS + k0));
    uint2 hi, lo;
    split_bf16x4(w.x, w.y, w.z, w.w, hi, lo);
    const uint32_t off = (uint32_t)((k0 >> 5) * B_TILE) + sw64_off(n, k0 & 31);
    *reinterpret_cast<uint2*>(smem + off) = hi;
    *reinterpret_cast<uint2*>(smem + off_b2 + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();                          // every CTA's barriers exist before a peer arrives on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const unsigned int per_step = (unsigned int)p.blocks * kSplit;   // CTAs of a slab

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.n_steps; ++s) {
        if (s > 0) {
          while (ld_acquire_u32(p.flags + m_blk) < (unsigned int)s * per_step) __nanosleep(20);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int row0 = (s & 1) * B + m_blk * BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, A_STAGE);
          const uint32_t dst = smem0 + off_a + stage * A_STAGE;
          tma_load_2d(dst, &map_g1, full, kq * KS + kb * 32, row0);
          tma_load_2d(dst + A_STAGE / 2, &map_g2, full, kq * KS + kb * 32, row0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BN, false, false, 128);
      const uint32_t d_main = tmem_base, d_aux = tmem_base + BN;
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.n_steps; ++s) {
        if (s > 0) {
          mbar_wait(bar_tempty, (uint32_t)(s - 1) & 1u);
          tc_fence_after();
        }
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a1 = smem0 + off_a + stage * A_STAGE, a2 = a1 + A_STAGE / 2;
          const uint32_t b1 = smem0 + kb * B_TILE, b2 = smem0 + off_b2 + kb * B_TILE;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da1 = smem_desc(a1 + j * 32, 16, 512, 4), da2 = smem_desc(a2 + j * 32, 16, 512, 4);
            const uint64_t db1 = smem_desc(b1 + j * 32, 16, 512, 4), db2 = smem_desc(b2 + j * 32, 16, 512, 4);
            const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
            umma_bf16(d_main, da1, db1, idesc, first);
            umma_bf16(d_aux, da1, db2, idesc, first);
            umma_bf16(d_aux, da2, db1, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ exchange + cell backward: thread = (row, 8 units)
    const int q = warp - 4, r = q * 32 + lane;
    const int b = m_blk * BM + r;
    const bool row_ok = b < B;
    const int unit = n_blk * kBU + kq * 8;
    const size_t BR = (size_t)B * R;
    const size_t ridx = (size_t)(row_ok ? b : 0) * R + unit;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t xlocal = smem0 + off_x;
    uint32_t xremote[4], xbar_remote[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) { xremote[d] = mapa_cluster(xlocal, (uint32_t)d); xbar_remote[d] = mapa_cluster(bar_x, (uint32_t)d); }
    auto ld8 = [](const float* ptr, float (&v)[8]) {
      const float4 a = *reinterpret_cast<const float4*>(ptr), c = *reinterpret_cast<const float4*>(ptr + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    };
    auto st8 = [](float* ptr, const float (&v)[8]) {
      *reinterpret_cast<float4*>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(ptr + 4) = make_float4(v[4], v[5], v[6], v[7]);
    };
    float dc[8];
    ld8(p.dc + ridx, dc);
    for (int s = 0; s < p.n_steps; ++s) {
      const int t = p.t_top - s;
      float* gp = p.gates + ((size_t)t * B + (row_ok ? b : 0)) * 4 * R + unit;
      // operands of the cell backward that do not depend on dh: in flight while the mainloop runs
      float gi[8], gj[8], gf[8], go[8], cp[8], cn[8], dh[8];
      ld8(gp, gi); ld8(gp + R, gj); ld8(gp + 2 * R, gf); ld8(gp + 3 * R, go);
      ld8(p.cbuf + (size_t)t * BR + ridx, cp);
      ld8(p.cbuf + (size_t)(t + 1) * BR + ridx, cn);
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] = 0.f;
      if (p.dout) {
        float d8[8];
        ld8(p.dout + (size_t)t * BR + ridx, d8);
        if (p.dscale) {
          float s8[8];
          ld8(p.dscale + (size_t)t * BR + ridx, s8);
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] = d8[i] * s8[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] = d8[i];
        }
      }
      mbar_wait(bar_tfull, (uint32_t)s & 1u);
      tc_fence_after();
      float v[32];
      {
        float vx[32];
        tmem_ld32(tacc, v);
        tmem_ld32(tacc + BN, vx);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += vx[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
      // send the 8 columns each peer owns; slot index of this CTA in a peer's buffer: kq - (kq > d)
      const uint32_t par = (uint32_t)(s & 1) * (XBUF * 4);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (d == kq) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] += v[8 * d + i];
        } else {
          const int slot = kq - (kq > d ? 1 : 0);
          const uint32_t a = xremote[d] + par + (uint32_t)((slot * 128 + r) * 32);
          st_cluster_f4(a, v[8 * d], v[8 * d + 1], v[8 * d + 2], v[8 * d + 3]);
          st_cluster_f4(a + 16, v[8 * d + 4], v[8 * d + 5], v[8 * d + 6], v[8 * d + 7]);
        }
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
          if (d != kq) mbar_arrive_release_cluster(xbar_remote[d]);
      }
      mbar_wait_cluster(bar_x, (uint32_t)s & 1u);
      {
        const float* xb = reinterpret_cast<const float*>(smem + off_x) + (s & 1) * XBUF;
#pragma unroll
        for (int sl = 0; sl < 3; ++sl) {
          const float4 a = *reinterpret_cast<const float4*>(xb + (sl * 128 + r) * 8);
          const float4 c = *reinterpret_cast<const float4*>(xb + (sl * 128 + r) * 8 + 4);
          dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w; dh[4] += c.x; dh[5] += c.y; dh[6] += c.z; dh[7] += c.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float tcn = tanh_fast(cn[i]);
        const float dcc = dc[i] + dh[i] * go[i] * (1.f - tcn * tcn);
        const float di = dcc * gj[i] * gi[i] * (1.f - gi[i]);
        const float dj = dcc * gi[i] * (1.f - gj[i] * gj[i]);
        const float df = dcc * cp[i] * gf[i] * (1.f - gf[i]);
        const float d_o = dh[i] * tcn * go[i] * (1.f - go[i]);
        gi[i] = di; gj[i] = dj; go[i] = d_o;
        dc[i] = dcc * gf[i];
        gf[i] = df;
      }
      if (row_ok) {
        // dG_t split first: it gates the next step of the whole slab
        const size_t so = ((size_t)((s + 1) & 1) * B + b) * 4 * R + unit;
        uint2 h0, l0, h1, l1;
        split_bf16x4(gi[0], gi[1], gi[2], gi[3], h0, l0); split_bf16x4(gi[4], gi[5], gi[6], gi[7], h1, l1);
        *reinterpret_cast<uint4*>(p.gs1 + so) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(p.gs2 + so) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        split_bf16x4(gj[0], gj[1], gj[2], gj[3], h0, l0); split_bf16x4(gj[4], gj[5], gj[6], gj[7], h1, l1);
        *reinterpret_cast<uint4*>(p.gs1 + so + R) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(p.gs2 + so + R) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        split_bf16x4(gf[0], gf[1], gf[2], gf[3], h0, l0); split_bf16x4(gf[4], gf[5], gf[6], gf[7], h1, l1);
        *reinterpret_cast<uint4*>(p.gs1 + so + 2 * R) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(p.gs2 + so + 2 * R) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        split_bf16x4(go[0], go[1], go[2], go[3], h0, l0); split_bf16x4(go[4], go[5], go[6], go[7], h1, l1);
        *reinterpret_cast<uint4*>(p.gs1 + so + 3 * R) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(p.gs2 + so + 3 * R) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 4 && lane == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.flags + m_blk), "r"(1u) : "memory");
      if (row_ok) {
        st8(gp, gi); st8(gp + R, gj); st8(gp + 2 * R, gf); st8(gp + 3 * R, go);
      }
    }
    if (row_ok) st8(p.dc + ridx, dc);
  }

  tc_fence_before();
  cluster_sync_all();                          // nobody leaves while a peer may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

static size_t bwd_smem(int R, int* stages) {
  const size_t resident = (size_t)2 * (R / 32) * (kBU * 64) + 2 * XBUF * 4;
  long long room = (long long)224 * 1024 - (long long)resident;
  int st = (int)(room / A_STAGE);
  if (st > kMaxStages) st = kMaxStages;
  if (st > R / 32) st = R / 32;
  *stages = st;
  return resident + (size_t)st * A_STAGE + 1024;
}

static int bwd_max_clusters(size_t smem) {
  static int cached = -1;
  static size_t cached_smem = 0;
  if (cached >= 0 && cached_smem == smem) return cached;
  cudaFuncSetAttribute(lstm_res_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kSplit * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kSplit; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, reinterpret_cast<const void*>(lstm_res_bwd_kernel), &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cached = n; cached_smem = smem;
  return n;
}

}  // namespace res
}  // namespace mnn

// CTAs of the weight-resident BPTT kernel for this shape under `sms` SMs (0: shape not taken)
int mnn_lstm_res_bwd_ctas(int n_steps, int B, int R, int sms) {
  static const char* env = getenv("MNN_LSTM_RES_BWD");   // "0": never
  if ((env && env[0] == '0') || n_steps < 2 || R % 32 != 0 || B > 512) return 0;
  int stages = 0;
  const size_t smem = res::bwd_smem(R, &stages);
  if (stages < 4) return 0;
  const int clusters = ((B + tc::BM - 1) / tc::BM) * (R / res::kBU);
  if (clusters * res::kSplit > sms || clusters > res::bwd_max_clusters(smem)) return 0;
  return clusters * res::kSplit;
}

// gs: 32 B R bytes + 256 (split dG double buffer); steps t = t_top .. t_top - n_steps + 1, gates slot t_top + 1 holds
// dG of the step after, dc the carried cell gradient
int mnn_lstm_res_bwd(float* gates, const float* wh, const float* cbuf, const float* dout, const float* dscale, float* dc,
                     int t_top, int n_steps, int B, int R, void* gs, unsigned int* flags, cudaStream_t stream) {
  res::BParams p{};
  p.gates = gates; p.cbuf = cbuf; p.dout = dout; p.dscale = dscale; p.dc = dc; p.wh = wh;
  p.gs1 = reinterpret_cast<__nv_bfloat16*>(gs);
  p.gs2 = p.gs1 + (size_t)2 * B * 4 * R;
  p.B = B; p.R = R; p.t_top = t_top; p.n_steps = n_steps;
  p.slabs = (B + tc::BM - 1) / tc::BM; p.blocks = R / res::kBU; p.kb_local = R / 32;
  const size_t smem = res::bwd_smem(R, &p.stages);
  p.flags = flags;
  CUtensorMap m1, m2;
  int rc = mnn_tc_make_map_bf16(p.gs1, 4LL * R, 4LL * R, 2LL * B, tc::BM, &m1);
  if (rc) return rc;
  rc = mnn_tc_make_map_bf16(p.gs2, 4LL * R, 4LL * R, 2LL * B, tc::BM, &m2);
  if (rc) return rc;
  cudaMemsetAsync(flags, 0, (size_t)p.slabs * sizeof(unsigned int), stream);
  const size_t n4 = (size_t)B * R;   // B * 4R / 4
  res::split_rows_kernel<<<(unsigned)((n4 + 255) / 256 < 296 ? (n4 + 255) / 256 : 296), 256, 0, stream>>>(
      gates + (size_t)(t_top + 1) * B * 4 * R, p.gs1, p.gs2, n4);
  rc = mnn_check_launch("lstm_res split_dG");
  if (rc) return rc;
  cudaFuncSetAttribute(res::lstm_res_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.slabs * p.blocks * res::kSplit);
  cfg.blockDim = dim3(res::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = res::kSplit; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, res::lstm_res_bwd_kernel, m1, m2, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    mnn_set_error(cudaGetErrorString(e));
    return (int)e;
  }
  return mnn_check_launch("lstm_seq_bwd(resident)");
}
