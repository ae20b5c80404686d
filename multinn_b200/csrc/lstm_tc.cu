// LSTM recurrence on the tensor cores (K2): the sequential part of the temporal unit, h_{t-1}.Wh and its BPTT
// counterpart dG_{t+1}.Wh^T, as tcgen05 3xTF32 GEMM steps with the LSTM cell fused into the epilogue, optionally
// persistent over all T steps (one cooperative launch per layer and direction; CTAs that share a 128-row batch slab
// synchronise through a per-slab counter in global memory, nothing is grid-wide).
// Semantics: CudnnCompatibleLSTMCell == LSTMBlockCell(forget_bias 0), gate blocks i, j, f, o, DropoutWrapper on the
// output (reference common/rnn.py:104-145), driven over time like dynamic_decode
// (generators/rnn_nade.py:204-218); the backward is what tf.gradients builds through that loop.
//
// Forward work item (slab m, unit block n): accumulator[128 rows, 4 gates x UB units] = h_{t-1}[slab] . WhP^T where WhP
// is Wh transposed and permuted so that the four gate columns of a unit block are adjacent (prepared once per call).
// Backward work item (slab m, unit block n): accumulator[128 rows, BN units] = dG_{t+1}[slab] . Wh^T, then the cell
// backward for step t writes dG_t in place of the saved gate activations.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "multinn_b200.h"
#include "tc_common.cuh"

// lstm_res.cu: weight-resident persistent forward kernel for small per-GPU batches
int mnn_lstm_res_ctas(int T, int B, int R, int sms);
size_t mnn_lstm_res_workspace_bytes(int B, int R);
int mnn_lstm_res_bwd_ctas(int n_steps, int B, int R, int sms);
int mnn_lstm_res_bwd(float* gates, const float* wh, const float* cbuf, const float* dout, const float* dscale, float* dc,
                     int t_top, int n_steps, int B, int R, void* gs, unsigned int* flags, cudaStream_t stream);
int mnn_lstm_res_fwd(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale, const float* u,
                     float keep, unsigned long long seed, int T, int B, int R, void* hs, unsigned int* flags, int sms,
                     cudaStream_t stream);

int mnn_tc_make_map_bf16(const void* ptr, long long ld_elems, long long inner, long long outer, int box_outer,
                         CUtensorMap* out);   // gemm_tc.cu: K-major bf16 tiles, SWIZZLE_64B
int mnn_tc_gemm_split();                       // gemm_tc.cu: mnn_set_gemm_split of the calling thread (1 = bf16 pairs)

namespace mnn {
namespace tc {

constexpr int kLThreads = 384;   // warp 0 TMA, warp 1 MMA, warps 4-7 epilogue, warps 8-11 hi/lo converters

template <int BN>
struct LCfg {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr int TMEM_COLS = (4 * BN < 32) ? 32 : 4 * BN;   // (main + aux) x double buffer
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
};

struct LstmParams {
  float* gates;         // [T][B][4R]
  float* hbuf;          // [(T+1)][B][R]    (fwd)
  float* cbuf;          // [(T+1)][B][R]
  float* out;           // [T][B][R] or null (fwd)
  float* dscale;        // [T][B][R] or null
  const float* u;       // [T][B][R] or null (fwd)
  const float* dout;    // [T][B][R] or null (bwd)
  float* dc;            // [B][R] carry (bwd)
  float keep;
  unsigned long long seed;
  RowMap rmap;          // global-row / global-time keying of the dropout Philox counter
  int T, B, R;
  int t0, t1;           // fwd: steps t0..t1-1 ascending; bwd: steps t1-1..t0 descending
  int slabs, blocks;    // work items per step
  int kb_total;         // k-blocks per item
  unsigned int* flags;  // [slabs] completed-item counters (persistent launches), zeroed by the host
  unsigned long long* trace;  // debug (MNN_LSTM_TRACE): [cta][step][16] globaltimer stamps, or null
  int bf16x;            // pair forward kernel: cross terms as bf16 MMAs (2.5-product scheme, see gemm_tc.cu)
  int b_pre;            // pair forward kernel, bf16x == 2: WhP^T arrives as bf16 pair planes (split once per launch, not per step)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr int kTraceSteps = 8, kTraceS0 = 16, kTraceEv = 16;
#define MNN_TRACE(ev)                                                                                        \
  do {                                                                                                       \
    if (p.trace && s >= kTraceS0 && s < kTraceS0 + kTraceSteps)                                              \
      p.trace[((size_t)blockIdx.x * kTraceSteps + (s - kTraceS0)) * kTraceEv + (ev)] = gtimer();             \
  } while (0)

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Shared skeleton of both directions. FWD: A = h slot t (rows t*B + m*128), B = WhP^T rows n*BN..; BWD: A = dG slot t+1,
// B = Wh rows n*BN.. (both operands K-major).
// SHARE (one item per CTA and step, the small-batch regime): the mainloop and the cell epilogue of an item cannot
// overlap anyway (the next step's h does not exist yet), so all 8 worker warps convert during the mainloop and all 8 run
// the epilogue (warps 4-7 the even 8-unit groups, warps 8-11 the odd ones; a warp reads the TMEM lanes of quadrant
// warp % 4). Measured at B=256, R=512 (globaltimer trace): conversion by 4 warps bounded the mainloop at 0.55 us per
// k-block (9.5 us per step) and the 4-warp row-per-thread epilogue took 7.4 us.
template <int BN, bool FWD, bool SHARE>
__global__ void __launch_bounds__(kLThreads, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const LstmParams p) {
  using C_ = LCfg<BN>;
  constexpr int kConvThreads = SHARE ? 256 : 128;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 4];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[STAGES]);
  const uint32_t bar_empty = smem_u32(&bars[2 * STAGES]), bar_tfull = smem_u32(&bars[3 * STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[3 * STAGES + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvThreads);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kConvThreads);
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

  const int n_items = p.slabs * p.blocks;
  const int n_steps = p.t1 - p.t0;
  const bool persistent = n_steps > 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int t = FWD ? p.t0 + s : p.t1 - 1 - s;
        const int a_slot = FWD ? t : t + 1;     // fwd reads h_{t-1} (slot t); bwd reads dG_{t+1}
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
          const int n_blk = w % p.blocks, m_blk = w / p.blocks;
          if (persistent && s > 0) {
            const unsigned int target = (unsigned int)s * (unsigned int)p.blocks;
            while (ld_acquire(p.flags + m_blk) < target) __nanosleep(32);
            asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic-proxy stores -> our TMA reads
          }
          if (w == blockIdx.x) MNN_TRACE(0);
          const int row0 = a_slot * p.B + m_blk * BM;
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, C_::A_BYTES + C_::B_BYTES);
            const uint32_t a_dst = smem0 + stage * C_::STAGE_BYTES;
            tma_load_2d(a_dst, &map_a, full, kb * BK, row0);
            tma_load_2d(a_dst + 2 * C_::A_BYTES, &map_b, full, kb * BK, n_blk * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (w == blockIdx.x) MNN_TRACE(1);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(BN, false, false), idesc2 = idesc_tf32(2 * BN, false, false);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int s = 0; s < n_steps; ++s) {
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * 2 * BN, tmem_x = tmem_d + BN;
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            mbar_wait(bar_conv + 8 * stage, phase);
            tc_fence_after();
            if (kb == 0 && w == blockIdx.x) MNN_TRACE(4);
            const uint32_t a_raw = smem0 + stage * C_::STAGE_BYTES, a_lo = a_raw + C_::A_BYTES;
            const uint32_t b_raw = a_raw + 2 * C_::A_BYTES;
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t da = smem_desc(a_raw + j * 32, 16, 1024, 2), dal = smem_desc(a_lo + j * 32, 16, 1024, 2);
              const uint64_t db = smem_desc(b_raw + j * 32, 16, 1024, 2);
              const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
              // [B_hi ; B_lo] is one contiguous 2*BN-row operand: main and aux accumulators in a single instruction
              umma_tf32(tmem_d, da, db, idesc2, first);
              umma_tf32(tmem_x, dal, db, idesc, 1u);
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_tfull + 8 * acc);
          if (w == blockIdx.x) MNN_TRACE(5);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ worker warps: hi/lo converters (lo = x - trunc_tf32(x))
    // and the epilogue = the LSTM cell (fwd) / its backward. !SHARE: warps 8-11 convert, warps 4-7 run the epilogue.
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int hf = (warp - 4) >> 2;            // 0: warps 4-7, 1: warps 8-11
    const bool conv_role = SHARE || hf == 1, epi_role = SHARE || hf == 0;
    const int tc = SHARE ? threadIdx.x - 4 * 32 : threadIdx.x - 8 * 32;
    const int ug0 = SHARE ? hf : 0;
    constexpr int UGS = SHARE ? 2 : 1;
    int acc = 0, stage = 0;
    uint32_t acc_phase = 0, phase = 0;
    const int R = p.R, B = p.B;
    const size_t BR = (size_t)B * R;
    for (int s = 0; s < n_steps; ++s) {
      const int t = FWD ? p.t0 + s : p.t1 - 1 - s;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int n_blk = w % p.blocks, m_blk = w / p.blocks;
        const int b = m_blk * BM + q * 32 + lane;
        const bool row_ok = b < B;
        // the operands of the NEXT step's epilogue for this item stream from HBM (they were written long ago): pull
        // them into L2 now, off the critical path of the recurrence
        if (hf == 0 && persistent && s + 1 < n_steps && row_ok) {
          const int tn = FWD ? t + 1 : t - 1;
          if (FWD) {
            constexpr int UBp = BN / 4;
            const float* gn = p.gates + ((size_t)tn * B + b) * 4 * R + n_blk * UBp;
            if (n_blk * UBp < R) {
#pragma unroll
              for (int g = 0; g < 4; ++g) prefetch_l2(gn + g * R);
              if (p.u) prefetch_l2(p.u + (size_t)tn * BR + (size_t)b * R + n_blk * UBp);
            }
          } else {
            const int u0 = n_blk * BN;
            if (u0 < R) {
              const float* gn = p.gates + ((size_t)tn * B + b) * 4 * R + u0;
#pragma unroll
              for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int c = 0; c < BN; c += 32) prefetch_l2(gn + g * R + c);
#pragma unroll
              for (int c = 0; c < BN; c += 32) {
                prefetch_l2(p.cbuf + (size_t)tn * BR + (size_t)b * R + u0 + c);
                if (p.dout) prefetch_l2(p.dout + (size_t)tn * BR + (size_t)b * R + u0 + c);
                if (p.dscale) prefetch_l2(p.dscale + (size_t)tn * BR + (size_t)b * R + u0 + c);
              }
            }
          }
        }
        if (conv_role) {
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            if (tc == 0 && w == blockIdx.x) {
              if (kb == 0) MNN_TRACE(2);
              if (kb == p.kb_total - 1) MNN_TRACE(3);
            }
            uint8_t* base = smem_gen + (size_t)stage * C_::STAGE_BYTES;
            const float4* a_raw = reinterpret_cast<const float4*>(base);
            float4* a_lo = reinterpret_cast<float4*>(base + C_::A_BYTES);
            const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * C_::A_BYTES);
            float4* b_lo = reinterpret_cast<float4*>(base + 2 * C_::A_BYTES + C_::B_BYTES);
#pragma unroll 4
            for (int i = tc; i < C_::A_BYTES / 16; i += kConvThreads) a_lo[i] = tf32_lo4(a_raw[i]);
#pragma unroll 4
            for (int i = tc; i < C_::B_BYTES / 16; i += kConvThreads) b_lo[i] = tf32_lo4(b_raw[i]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_conv + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (!epi_role) continue;
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        if (threadIdx.x == 128 && w == blockIdx.x) MNN_TRACE(6);
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN);
        if (FWD) {
          constexpr int UB = BN / 4;
          float* gp = p.gates + ((size_t)t * B + (row_ok ? b : 0)) * 4 * R;
          const size_t sidx = (size_t)(row_ok ? b : 0) * R;
          const float* cprev = p.cbuf + (size_t)t * BR + sidx;
          float* cnew = p.cbuf + (size_t)(t + 1) * BR + sidx;
          float* hnew = p.hbuf + (size_t)(t + 1) * BR + sidx;
#pragma unroll 1
          for (int ug = ug0; ug < UB / 8; ug += UGS) {
            const int unit = n_blk * UB + ug * 8;
            if (unit >= R) break;   // warp-uniform (R % 8 == 0)
            float pre[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float m8[8], x8[8];
              tmem_ld8(tacc + (uint32_t)(g * UB + ug * 8), m8);
              tmem_ld8(tacc + (uint32_t)(BN + g * UB + ug * 8), x8);
#pragma unroll
              for (int i = 0; i < 8; ++i) pre[g][i] = m8[i] + x8[i];
            }
            if (row_ok) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 p0 = *reinterpret_cast<const float4*>(gp + g * R + unit);
                const float4 p1 = *reinterpret_cast<const float4*>(gp + g * R + unit + 4);
                pre[g][0] += p0.x; pre[g][1] += p0.y; pre[g][2] += p0.z; pre[g][3] += p0.w;
                pre[g][4] += p1.x; pre[g][5] += p1.y; pre[g][6] += p1.z; pre[g][7] += p1.w;
              }
              const float4 c0 = *reinterpret_cast<const float4*>(cprev + unit);
              const float4 c1 = *reinterpret_cast<const float4*>(cprev + unit + 4);
              const float cp[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
              float cv[8], hv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float gi, gj, gf, go;
                lstm_cell_xu(pre[0][i], pre[1][i], pre[2][i], pre[3][i], cp[i], gi, gj, gf, go, cv[i], hv[i]);
                pre[0][i] = gi; pre[1][i] = gj; pre[2][i] = gf; pre[3][i] = go;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                *reinterpret_cast<float4*>(gp + g * R + unit) = make_float4(pre[g][0], pre[g][1], pre[g][2], pre[g][3]);
                *reinterpret_cast<float4*>(gp + g * R + unit + 4) = make_float4(pre[g][4], pre[g][5], pre[g][6], pre[g][7]);
              }
              *reinterpret_cast<float4*>(cnew + unit) = make_float4(cv[0], cv[1], cv[2], cv[3]);
              *reinterpret_cast<float4*>(cnew + unit + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
              *reinterpret_cast<float4*>(hnew + unit) = make_float4(hv[0], hv[1], hv[2], hv[3]);
              *reinterpret_cast<float4*>(hnew + unit + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
              if (p.out) {
                float ov[8], dv[8];
                if (p.keep < 1.0f) {
                  const size_t e = (size_t)t * BR + sidx + unit;
                  float uu[8];
                  if (p.u) {
                    const float4 u0 = __ldg(reinterpret_cast<const float4*>(p.u + e));
                    const float4 u1 = __ldg(reinterpret_cast<const float4*>(p.u + e + 4));
                    uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w;
                    uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
                  } else {
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                      const uint4 r4 = dropout_bits4(p.seed, p.rmap, b, unit + 4 * h4, R, t);
                      uu[4 * h4] = u01(r4.x); uu[4 * h4 + 1] = u01(r4.y);
                      uu[4 * h4 + 2] = u01(r4.z); uu[4 * h4 + 3] = u01(r4.w);
                    }
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float km = floorf(p.keep + uu[i]);   // tf.nn.dropout: x / keep * floor(keep + u)
                    ov[i] = hv[i] / p.keep * km;
                    dv[i] = km / p.keep;
                  }
                  float* dsp = p.dscale + e;
                  *reinterpret_cast<float4*>(dsp) = make_float4(dv[0], dv[1], dv[2], dv[3]);
                  *reinterpret_cast<float4*>(dsp + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) ov[i] = hv[i];
                }
                float* op = p.out + (size_t)t * BR + sidx + unit;
                *reinterpret_cast<float4*>(op) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(ov[4], ov[5], ov[6], ov[7]);
              }
            }
          }
        } else {
          float* gp = p.gates + ((size_t)t * B + (row_ok ? b : 0)) * 4 * R;
          const size_t sidx = (size_t)(row_ok ? b : 0) * R;
          const float* cprev = p.cbuf + (size_t)t * BR + sidx;
          const float* cnow = p.cbuf + (size_t)(t + 1) * BR + sidx;
#pragma unroll 1
          for (int ug = ug0; ug < BN / 8; ug += UGS) {
            const int unit = n_blk * BN + ug * 8;
            if (unit >= R) break;
            float dh[8];
            {
              float m8[8], x8[8];
              tmem_ld8(tacc + (uint32_t)(ug * 8), m8);
              tmem_ld8(tacc + (uint32_t)(BN + ug * 8), x8);
#pragma unroll
              for (int i = 0; i < 8; ++i) dh[i] = m8[i] + x8[i];
            }
            if (row_ok) {
              auto ld8 = [](const float* ptr, float (&v)[8]) {
                const float4 a = *reinterpret_cast<const float4*>(ptr), c = *reinterpret_cast<const float4*>(ptr + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
              };
              auto st8 = [](float* ptr, const float (&v)[8]) {
                *reinterpret_cast<float4*>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(ptr + 4) = make_float4(v[4], v[5], v[6], v[7]);
              };
              if (p.dout) {
                float d8[8];
                ld8(p.dout + (size_t)t * BR + sidx + unit, d8);
                if (p.dscale) {
                  float s8[8];
                  ld8(p.dscale + (size_t)t * BR + sidx + unit, s8);
#pragma unroll
                  for (int i = 0; i < 8; ++i) dh[i] = fmaf(d8[i], s8[i], dh[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) dh[i] += d8[i];
                }
              }
              float gi[8], gj[8], gf[8], go[8], cp[8], cn[8], dc[8];
              ld8(gp + unit, gi); ld8(gp + R + unit, gj); ld8(gp + 2 * R + unit, gf); ld8(gp + 3 * R + unit, go);
              ld8(cprev + unit, cp); ld8(cnow + unit, cn); ld8(p.dc + sidx + unit, dc);
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
              st8(gp + unit, gi); st8(gp + R + unit, gj); st8(gp + 2 * R + unit, gf); st8(gp + 3 * R + unit, go);
              st8(p.dc + sidx + unit, dc);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        if (threadIdx.x == 128 && w == blockIdx.x) MNN_TRACE(7);
        if (persistent) {
          // publish this item's h_t / dG_t rows to the CTAs of the same slab
          __threadfence();
          if (SHARE) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 4 * 32) {
            asm volatile("fence.proxy.async;" ::: "memory");
            atomicAdd(p.flags + m_blk, 1u);
            if (w == blockIdx.x) MNN_TRACE(8);
          }
        }
      }
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

// ------------------------------------------------------------------------------------------------ forward on CTA pairs
// The forward recurrence on CTA pairs (cluster of 2, tcgen05 cta_group::2): one work item = (256-row batch slab,
// block of 64 units = 256 gate columns); each CTA of the pair holds 128 rows of h_{t-1} and 128 of the 256 WhP^T rows
// per stage and ends with the accumulators of its own 128 batch rows (main + aux = all 512 TMEM columns). The pair
// instruction reads half as many operand bytes from shared memory per flop as the 128x128 one, which is what bounds the
// 1-CTA kernel. The converter warps have nothing to convert while the cell epilogue runs (the next step's h does not
// exist yet), so they take the second half of the item's units: 8 epilogue warps per CTA.
constexpr int kL2UB = 64;   // units per item
constexpr size_t kPairSmem = 2 * 7 * (size_t)(BM * 128) + 1024;   // epilogue tiles (224 KB) alias the 3 x 64 KB stages

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLThreads, 1)
lstm_tc2_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_g,
                    const __grid_constant__ CUtensorMap map_c,
                    const __grid_constant__ CUtensorMap map_o, const LstmParams p) {
  constexpr int BNP = 4 * kL2UB;                       // 256 gate columns per item
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = (BNP / 2) * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES, STAGES = 3;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 2];
  __shared__ __align__(8) uint64_t ebars[4];   // [0],[1]: tile loads of the two halves; [2]: tiles read out; [3]: tile stores complete
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[STAGES]);
  const uint32_t bar_empty = smem_u32(&bars[2 * STAGES]), bar_tfull = smem_u32(&bars[3 * STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[3 * STAGES + 1]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 2 * 4);      // one elected arrive per converter warp of both CTAs
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 2 * 8);
    mbar_init(smem_u32(&ebars[0]), 1);
    mbar_init(smem_u32(&ebars[1]), 1);
    mbar_init(smem_u32(&ebars[2]), 2);
    mbar_init(smem_u32(&ebars[3]), 2);              // one elected arrive per epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_items = p.slabs * p.blocks;     // slabs of 256 rows
  const int n_steps = p.t1 - p.t0;
  const int R = p.R, B = p.B;
  const size_t BR = (size_t)B * R;

  // ---- cell epilogue of one item half (eh): units [n_blk*64 + eh*32, +32) of this CTA's 128 rows.
  // A TMEM lane is a batch row, so a thread owns a row; touching global memory row-per-thread costs 32 cache lines per
  // warp instruction (measured: 22 us of a 41 us step). All tile I/O therefore goes through shared memory with TMA:
  // the mainloop stages are idle once the item's MMAs are done, and hold per half 7 tiles of [128 rows x 32 units]
  // (SWIZZLE_128B: 16-byte chunk c of row r at chunk c ^ (r & 7)): the four gate pre-activations and c_{t-1} are
  // loaded, overwritten in place with the gate activations and c_t, h_t and the dropped-out output are added, and
  // everything is stored back with TMA; only dscale (mask / keep) is written directly.
  constexpr int TILE = BM * 128;                       // 16 KB
  constexpr int HALF_BYTES = 7 * TILE;                 // gates i, j, f, o | c | h | out
  auto epilogue_half = [&](int s, int t, int n_blk, int m_blk, int q, int eh, uint32_t ein_phase, uint32_t tempty_leader) {
    const int r = q * 32 + lane;                       // row inside this CTA's 128
    const int row_g = m_blk * 2 * BM + (int)rank * BM; // first batch row of this CTA
    const int b = row_g + r;
    const int unit0 = n_blk * kL2UB + eh * 32;
    const uint32_t ebase = smem0 + (uint32_t)eh * HALF_BYTES;
    uint8_t* egen = smem_gen + (size_t)eh * HALF_BYTES;
    const uint32_t bar_ein = smem_u32(&ebars[eh]);
    const bool leader = (lane == 0 && q == 0);
    mbar_wait(bar_ein, ein_phase);   // gate pre-activation and c_{t-1} tiles (loaded by the producer warp as stages free up)
    if (eh == 0 && threadIdx.x == 128) MNN_TRACE(11);
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    const float inv_keep = 1.0f / p.keep;
    uint8_t* rowp = egen + r * 128;
    const int sw = r & 7;
#pragma unroll 1
    for (int ug = 0; ug < 4; ++ug) {
      const int ul = eh * 32 + ug * 8;                 // unit offset inside the item
      float pre[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float m8[8], x8[8];
        tmem_ld8(tacc + (uint32_t)(g * kL2UB + ul), m8);
        tmem_ld8(tacc + (uint32_t)(BNP + g * kL2UB + ul), x8);
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[g][i] = m8[i] + x8[i];
      }
      const int c0 = ((2 * ug) ^ sw) << 4, c1 = ((2 * ug + 1) ^ sw) << 4;   // swizzled chunk offsets of this row
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 p0 = *reinterpret_cast<const float4*>(rowp + g * TILE + c0);
        const float4 p1 = *reinterpret_cast<const float4*>(rowp + g * TILE + c1);
        pre[g][0] += p0.x; pre[g][1] += p0.y; pre[g][2] += p0.z; pre[g][3] += p0.w;
        pre[g][4] += p1.x; pre[g][5] += p1.y; pre[g][6] += p1.z; pre[g][7] += p1.w;
      }
      const float4 cc0 = *reinterpret_cast<const float4*>(rowp + 4 * TILE + c0);
      const float4 cc1 = *reinterpret_cast<const float4*>(rowp + 4 * TILE + c1);
      const float cp[8] = {cc0.x, cc0.y, cc0.z, cc0.w, cc1.x, cc1.y, cc1.z, cc1.w};
      float cv[8], hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gi, gj, gf, go;
        lstm_cell_xu(pre[0][i], pre[1][i], pre[2][i], pre[3][i], cp[i], gi, gj, gf, go, cv[i], hv[i]);
        pre[0][i] = gi; pre[1][i] = gj; pre[2][i] = gf; pre[3][i] = go;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<float4*>(rowp + g * TILE + c0) = make_float4(pre[g][0], pre[g][1], pre[g][2], pre[g][3]);
        *reinterpret_cast<float4*>(rowp + g * TILE + c1) = make_float4(pre[g][4], pre[g][5], pre[g][6], pre[g][7]);
      }
      *reinterpret_cast<float4*>(rowp + 4 * TILE + c0) = make_float4(cv[0], cv[1], cv[2], cv[3]);
      *reinterpret_cast<float4*>(rowp + 4 * TILE + c1) = make_float4(cv[4], cv[5], cv[6], cv[7]);
      // h_t gates the next step of the whole slab: it goes straight to global memory (8 stores per thread) so that the
      // flag can be raised without waiting for a TMA store to complete; everything else leaves through the tiles
      {
        float* hdst = p.hbuf + (size_t)(t + 1) * BR + (size_t)b * R + unit0 + ug * 8;
        *reinterpret_cast<float4*>(hdst) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4*>(hdst + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      }
      if (p.out) {
        float ov[8];
        if (p.keep < 1.0f) {
          const size_t e = (size_t)t * BR + (size_t)b * R + unit0 + ug * 8;
          float uu[8], dv[8];
          if (p.u) {
            const float4 u0 = __ldg(reinterpret_cast<const float4*>(p.u + e));
            const float4 u1 = __ldg(reinterpret_cast<const float4*>(p.u + e + 4));
            uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w;
            uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
          } else {
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
              const uint4 r4 = dropout_bits4(p.seed, p.rmap, b, unit0 + ug * 8 + 4 * h4, R, t);
              uu[4 * h4] = u01(r4.x); uu[4 * h4 + 1] = u01(r4.y);
              uu[4 * h4 + 2] = u01(r4.z); uu[4 * h4 + 3] = u01(r4.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            dv[i] = floorf(p.keep + uu[i]) * inv_keep;   // tf.nn.dropout: x / keep * floor(keep + u)
            ov[i] = hv[i] * dv[i];
          }
          float* dsp = p.dscale + e;
          *reinterpret_cast<float4*>(dsp) = make_float4(dv[0], dv[1], dv[2], dv[3]);
          *reinterpret_cast<float4*>(dsp + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) ov[i] = hv[i];
        }
        *reinterpret_cast<float4*>(rowp + 6 * TILE + c0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(rowp + 6 * TILE + c1) = make_float4(ov[4], ov[5], ov[6], ov[7]);
      }
    }
    if (eh == 0 && threadIdx.x == 128) MNN_TRACE(7);
    if (eh == 1 && threadIdx.x == 256) MNN_TRACE(10);
    // tiles complete -> TMA stores; the flag may only be raised once h_t is globally visible
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __threadfence();                                 // this thread's h_t stores are visible device-wide
    if (eh == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
    if (leader) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + m_blk) : "memory");
      if (eh == 0) MNN_TRACE(8);
#pragma unroll
      for (int g = 0; g < 4; ++g) tma_store_2d(&map_g, ebase + g * TILE, g * R + unit0, t * B + row_g);
      tma_store_2d(&map_c, ebase + 4 * TILE, unit0, (t + 1) * B + row_g);
      if (p.out) tma_store_2d(&map_o, ebase + 6 * TILE, unit0, t * B + row_g);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile stores have read shared memory
      mbar_arrive(smem_u32(&ebars[2]));                                 // -> the stages may be refilled
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // c_t (and the rest) is written: the next step's
      mbar_arrive(smem_u32(&ebars[3]));                                 // tile loads of this item may be issued
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(tempty_leader);
  };
  // pull the next step's epilogue operands of this item half into L2 (they stream from HBM)
  auto prefetch_next = [&](int t, int n_blk, int m_blk, int q, int eh) {
    const int b = m_blk * 2 * BM + (int)rank * BM + q * 32 + lane;
    const int unit = n_blk * kL2UB + eh * 32;
    if (b >= B || unit >= R) return;
    const float* gn = p.gates + ((size_t)(t + 1) * B + b) * 4 * R + unit;
#pragma unroll
    for (int g = 0; g < 4; ++g) prefetch_l2(gn + g * R);
    if (p.u) prefetch_l2(p.u + (size_t)(t + 1) * BR + (size_t)b * R + unit);
  };
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0, item_no = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int t = p.t0 + s;
        for (int w = cluster_id; w < n_items; w += n_clusters) {
          const int n_blk = w % p.blocks, m_blk = w / p.blocks;
          if (s > 0) {
            const unsigned int target = (unsigned int)s * (unsigned int)p.blocks * 4u;   // 2 CTAs x 2 halves per item
            while (ld_acquire(p.flags + m_blk) < target) __nanosleep(32);
            asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic-proxy stores -> our TMA reads
          }
          MNN_TRACE(0);
          if (item_no > 0) mbar_wait(smem_u32(&ebars[2]), (uint32_t)(item_no - 1) & 1u);   // previous epilogue's tiles drained
          const int row0 = t * p.B + m_blk * 2 * BM + (int)rank * BM;
          const int brow0 = n_blk * BNP + (int)rank * (BNP / 2);
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, A_BYTES + B_BYTES);
            const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
            tma_load_2d(a_dst, &map_a, full, kb * BK, row0);
            if (p.b_pre) {   // bf16 planes straight into the [hi | lo] tiles of the "lo" region (same bytes as the raw tile)
              tma_load_2d(a_dst + 2 * A_BYTES + B_BYTES, &map_b, full, kb * BK, brow0);
              tma_load_2d(a_dst + 2 * A_BYTES + B_BYTES + B_BYTES / 2, &map_b2, full, kb * BK, brow0);
            } else {
              tma_load_2d(a_dst + 2 * A_BYTES, &map_b, full, kb * BK, brow0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          MNN_TRACE(1);
          // epilogue inputs: the stage buffers are done for this item as their last MMAs retire; fill the tiles that
          // alias each stage the moment it frees (tile k of half eh lives at eh*7*TILE + k*TILE; stage st at st*4*TILE)
          {
            const int row_g = m_blk * 2 * BM + (int)rank * BM;
            const uint32_t ein0 = smem_u32(&ebars[0]), ein1 = smem_u32(&ebars[1]);
            if (item_no > 0) mbar_wait(smem_u32(&ebars[3]), (uint32_t)(item_no - 1) & 1u);   // c_{t-1} tile is in memory
            mbar_expect_tx(ein0, 5 * TILE);
            mbar_expect_tx(ein1, 5 * TILE);
            int st = stage;
            uint32_t ph = phase;
            for (int i = 0; i < STAGES; ++i) {
              mbar_wait(bar_empty + 8 * st, ph ^ 1);
              for (int k = 0; k < 4; ++k) {
                const int tile = st * 4 + k;                 // 16 KB slot index in the dynamic region
                const int eh = tile >= 7 ? 1 : 0, k7 = tile - 7 * eh;
                if (k7 > 4) continue;                        // h / out slots: outputs only
                const uint32_t dst = smem0 + (uint32_t)tile * TILE;
                const int unit0 = n_blk * kL2UB + eh * 32;
                if (k7 < 4) tma_load_2d(dst, &map_g, eh ? ein1 : ein0, k7 * R + unit0, t * B + row_g);
                else tma_load_2d(dst, &map_c, eh ? ein1 : ein0, unit0, t * B + row_g);
              }
              if (++st == STAGES) { st = 0; ph ^= 1; }
            }
          }
          ++item_no;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = idesc_tf32(BNP, false, false, 2 * BM);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t tmem_d = tmem_base, tmem_x = tmem_base + BNP;
      for (int s = 0; s < n_steps; ++s) {
        for (int w = cluster_id; w < n_items; w += n_clusters) {
          mbar_wait(bar_tempty, acc_phase ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_conv + 8 * stage, phase);
            tc_fence_after();
            if (kb == 0) MNN_TRACE(4);
            const uint32_t a_raw = smem0 + stage * STAGE_BYTES, a_lo = a_raw + A_BYTES;
            const uint32_t b_raw = a_raw + 2 * A_BYTES, b_lo = b_raw + B_BYTES;
            if (p.bf16x) {
              // h.Wh ~ hi.hi (tf32) + bf16(h).bf16(Wh_lo) + bf16(h_lo).bf16(Wh) (2.5-product scheme, gemm_tc.cu)
              const uint32_t idesc16 = idesc_bf16(BNP, false, false, 2 * BM);
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
                umma_bf16_2cta(tmem_x, smem_desc(a_lo + j * 32, 16, 512, 4), smem_desc(b_lo + B_BYTES / 2 + j * 32, 16, 512, 4),
                               idesc16, first);
                umma_bf16_2cta(tmem_x, smem_desc(a_lo + A_BYTES / 2 + j * 32, 16, 512, 4), smem_desc(b_lo + j * 32, 16, 512, 4),
                               idesc16, 1u);
              }
              if (p.bf16x == 2) {   // bf16 pairs: the main product too on kind::f16 (gemm_tc.cu, mnn_set_gemm_split)
#pragma unroll
                for (int j = 0; j < BK / 16; ++j)
                  umma_bf16_2cta(tmem_d, smem_desc(a_lo + j * 32, 16, 512, 4), smem_desc(b_lo + j * 32, 16, 512, 4), idesc16,
                                 (kb > 0 || j > 0) ? 1u : 0u);
              } else {
#pragma unroll
              for (int j = 0; j < BK / 8; ++j)
                umma_tf32_2cta(tmem_d, smem_desc(a_raw + j * 32, 16, 1024, 2), smem_desc(b_raw + j * 32, 16, 1024, 2), idesc,
                               (kb > 0 || j > 0) ? 1u : 0u);
              }
            } else {
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t da = smem_desc(a_raw + j * 32, 16, 1024, 2), dal = smem_desc(a_lo + j * 32, 16, 1024, 2);
              const uint64_t db = smem_desc(b_raw + j * 32, 16, 1024, 2), dbl = smem_desc(b_lo + j * 32, 16, 1024, 2);
              const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
              umma_tf32_2cta(tmem_x, da, dbl, idesc, first);
              umma_tf32_2cta(tmem_x, dal, db, idesc, 1u);
              umma_tf32_2cta(tmem_d, da, db, idesc, first);
            }
            }
            umma_commit_2cta(bar_empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit_2cta(bar_tfull);
          MNN_TRACE(5);
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ converters, then the second epilogue half
    const int tc = threadIdx.x - 8 * 32;
    const int q = warp - 8;
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t conv_leader = mapa_cluster(bar_conv, 0), tempty_leader = mapa_cluster(bar_tempty, 0);
    for (int s = 0; s < n_steps; ++s) {
      const int t = p.t0 + s;
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        const int n_blk = w % p.blocks, m_blk = w / p.blocks;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          if (kb == 0 && tc == 0) MNN_TRACE(2);
          uint8_t* base = smem_gen + (size_t)stage * STAGE_BYTES;
          const float4* a_raw = reinterpret_cast<const float4*>(base);
          float4* a_lo = reinterpret_cast<float4*>(base + A_BYTES);
          const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * A_BYTES);
          float4* b_lo = reinterpret_cast<float4*>(base + 2 * A_BYTES + B_BYTES);
          if (p.bf16x) {
            convert_bf16_tiles<false, 128>(base, base + A_BYTES, tc, true, p.bf16x == 2);
            if (!p.b_pre) convert_bf16_tiles<false, 128>(base + 2 * A_BYTES, base + 2 * A_BYTES + B_BYTES, tc, true, p.bf16x == 2);
          } else {
#pragma unroll 4
            for (int i = tc; i < A_BYTES / 16; i += 128) a_lo[i] = tf32_lo4(a_raw[i]);
#pragma unroll 4
            for (int i = tc; i < B_BYTES / 16; i += 128) b_lo[i] = tf32_lo4(b_raw[i]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(conv_leader + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc == 0) MNN_TRACE(3);
        if (s + 1 < n_steps) prefetch_next(t, n_blk, m_blk, q, 1);
        mbar_wait(bar_tfull, acc_phase);
        tc_fence_after();
        if (tc == 0) MNN_TRACE(9);
        epilogue_half(s, t, n_blk, m_blk, q, 1, acc_phase, tempty_leader);
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ first epilogue half
    const int q = warp - 4;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = mapa_cluster(bar_tempty, 0);
    for (int s = 0; s < n_steps; ++s) {
      const int t = p.t0 + s;
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        const int n_blk = w % p.blocks, m_blk = w / p.blocks;
        if (s + 1 < n_steps) prefetch_next(t, n_blk, m_blk, q, 0);
        mbar_wait(bar_tfull, acc_phase);
        tc_fence_after();
        if (threadIdx.x == 128) MNN_TRACE(6);
        epilogue_half(s, t, n_blk, m_blk, q, 0, acc_phase, tempty_leader);
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ BPTT on CTA pairs
// The backward recurrence dh_t = dG_{t+1} . Wh^T has a long reduction (K = 4R) and a narrow output (N = R), so whole
// output tiles leave most SMs idle. It runs in two phases per step, both spread over every CTA:
//   phase A  (slab of 256 rows, tile of 256 units, K split): pair MMA (cta_group::2, 3xTF32) -> partial dh in TMEM ->
//            swizzled shared tiles -> TMA reduce-add (fp32 add at L2) into dh_acc[B,R]; per-slab counter cntA
//   phase B  (16 rows x 16 units per WARP, double buffered): when the slab's partials are all in, TMA-load the dh_acc, gate, c, dout,
//            dscale and dc tiles, run the cell backward, TMA-store dG_t (in place of the gate activations), dc and a
//            zeroed dh_acc tile; per-slab counter cntB releases the next step's phase A
// No thread ever touches global memory row-per-thread; all tile traffic is TMA.
struct Lstm3Params {
  const float* dout;      // null: no gradient from above
  const float* dscale;    // null: no dropout
  int T, B, R;
  int t_top;              // first (highest) step of the launch: T-2, or T-1 when a later chunk follows (dG_T readable)
  int slabs, ntiles, splits, kb_per_split;   // phase A: slabs of 256 rows, tiles of 256 units, K = 4R in `splits` parts
  unsigned int* cntA;     // [slabs]
  unsigned int* cntB;     // [slabs]
  unsigned long long* trace;   // debug (MNN_LSTM_TRACE)
  int bf16x;              // phase A: cross terms as bf16 MMAs (2.5-product scheme, see gemm_tc.cu)
  int b_pre;              // phase A, bf16x == 2: Wh arrives as bf16 pair planes (split once per launch)
};
constexpr int kCellRows = 16, kCellUnits = 16, kCellTile = kCellRows * kCellUnits * 4;   // 1 KB tiles
constexpr size_t kBwd3Smem = 3 * 64 * 1024 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLThreads, 1)
lstm_tc3_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_b2,
                    const __grid_constant__ CUtensorMap map_dhr, const __grid_constant__ CUtensorMap map_dh,
                    const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_c,
                    const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_ds,
                    const __grid_constant__ CUtensorMap map_dc, const Lstm3Params p) {
  constexpr int NA = 256;
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = (NA / 2) * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES, STAGES = 3;
  constexpr int TILE = BM * 128;                      // 16 KB partial-out tile [128 rows x 32 units], SWIZZLE_128B
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 2];
  __shared__ __align__(8) uint64_t cbar[16];           // per cell warp and buffer: tile loads landed
  __shared__ __align__(8) uint64_t bfree;              // cell tiles drained -> stages may be refilled
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_conv = smem_u32(&bars[STAGES]);
  const uint32_t bar_empty = smem_u32(&bars[2 * STAGES]), bar_tfull = smem_u32(&bars[3 * STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[3 * STAGES + 1]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 2 * 4);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 2 * 8);
    for (int w = 0; w < 16; ++w) mbar_init(smem_u32(&cbar[w]), 1);
    mbar_init(smem_u32(&bfree), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int R = p.R, B = p.B;
  const int n_itemsA = p.slabs * p.ntiles * p.splits;      // <= n_clusters (host): at most one phase-A item per pair
  const int n_steps = p.t_top + 1;                         // t = t_top .. 0
  const int kb_total = (4 * R + BK - 1) / BK;
  const bool has_item = cluster_id < n_itemsA;
  const int ks = cluster_id % p.splits, nt = (cluster_id / p.splits) % p.ntiles, mA = cluster_id / (p.splits * p.ntiles);
  const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
  const unsigned int perA = (unsigned int)(p.ntiles * p.splits * 4);             // 2 CTAs x 2 halves per item
  const int chunks = R / kCellUnits, rblocks = B / kCellRows;
  const unsigned int perB = (unsigned int)((2 * BM / kCellRows) * chunks);       // cell items per slab
  const int n_cells = rblocks * chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (phase A operands)
    if (lane == 0 && has_item) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int t = p.t_top - s;
        if (s > 0) {
          const unsigned int target = (unsigned int)s * perB;                    // dG_{t+1} of the slab is complete
          while (ld_acquire(p.cntB + mA) < target) __nanosleep(32);
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_wait(smem_u32(&bfree), (uint32_t)(s - 1) & 1u);                  // this CTA's cell tiles are drained
        }
        MNN_TRACE(0);
        const int row0 = (t + 1) * B + mA * 2 * BM + (int)rank * BM;
        const int brow0 = nt * NA + (int)rank * (NA / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, A_BYTES + B_BYTES);
          const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
          tma_load_2d(a_dst, &map_a, full, kb * BK, row0);
          if (p.b_pre) {
            tma_load_2d(a_dst + 2 * A_BYTES + B_BYTES, &map_b, full, kb * BK, brow0);
            tma_load_2d(a_dst + 2 * A_BYTES + B_BYTES + B_BYTES / 2, &map_b2, full, kb * BK, brow0);
          } else {
            tma_load_2d(a_dst + 2 * A_BYTES, &map_b, full, kb * BK, brow0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0 && has_item) {
      const uint32_t idesc = idesc_tf32(NA, false, false, 2 * BM);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t tmem_d = tmem_base, tmem_x = tmem_base + NA;
      for (int s = 0; s < n_steps; ++s) {
        mbar_wait(bar_tempty, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_conv + 8 * stage, phase);
          tc_fence_after();
          if (kb == kb0) MNN_TRACE(1);
          const uint32_t a_raw = smem0 + stage * STAGE_BYTES, a_lo = a_raw + A_BYTES;
          const uint32_t b_raw = a_raw + 2 * A_BYTES, b_lo = b_raw + B_BYTES;
          if (p.bf16x) {
            // dG.Wh^T ~ hi.hi (tf32) + bf16(dG).bf16(Wh_lo) + bf16(dG_lo).bf16(Wh): both operands K-major, bf16 tiles
            // [hi | lo] of 8 KB each in the "lo" regions (K-major SWIZZLE_64B, convert_bf16_tiles)
            const uint32_t idesc16 = idesc_bf16(NA, false, false, 2 * BM);
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
              umma_bf16_2cta(tmem_x, smem_desc(a_lo + j * 32, 16, 512, 4), smem_desc(b_lo + B_BYTES / 2 + j * 32, 16, 512, 4),
                             idesc16, first);
              umma_bf16_2cta(tmem_x, smem_desc(a_lo + A_BYTES / 2 + j * 32, 16, 512, 4), smem_desc(b_lo + j * 32, 16, 512, 4),
                             idesc16, 1u);
            }
            if (p.bf16x == 2) {   // bf16 pairs: the main product too on kind::f16
#pragma unroll
              for (int j = 0; j < BK / 16; ++j)
                umma_bf16_2cta(tmem_d, smem_desc(a_lo + j * 32, 16, 512, 4), smem_desc(b_lo + j * 32, 16, 512, 4), idesc16,
                               (kb > kb0 || j > 0) ? 1u : 0u);
            } else {
#pragma unroll
            for (int j = 0; j < BK / 8; ++j)
              umma_tf32_2cta(tmem_d, smem_desc(a_raw + j * 32, 16, 1024, 2), smem_desc(b_raw + j * 32, 16, 1024, 2), idesc,
                             (kb > kb0 || j > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t da = smem_desc(a_raw + j * 32, 16, 1024, 2), dal = smem_desc(a_lo + j * 32, 16, 1024, 2);
            const uint64_t db = smem_desc(b_raw + j * 32, 16, 1024, 2), dbl = smem_desc(b_lo + j * 32, 16, 1024, 2);
            const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
            umma_tf32_2cta(tmem_x, da, dbl, idesc, first);
            umma_tf32_2cta(tmem_x, dal, db, idesc, 1u);
            umma_tf32_2cta(tmem_d, da, db, idesc, first);
          }
          }
          umma_commit_2cta(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(bar_tfull);
        MNN_TRACE(2);
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ converters (warps 8-11), phase-A epilogue, cells
    const int cw = warp - 4;                 // cell warp 0..7
    const int eh = cw >> 2, q = cw & 3;      // phase-A epilogue: column half, TMEM lane quadrant
    const int tc = threadIdx.x - 8 * 32;
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0, cphase[2] = {0u, 0u};
    const uint32_t conv_leader = mapa_cluster(bar_conv, 0), tempty_leader = mapa_cluster(bar_tempty, 0);
    const uint32_t my_cbar = smem_u32(&cbar[2 * cw]);
    const uint32_t cell0 = smem0 + (uint32_t)cw * 20 * kCellTile;   // two buffers of 10 tiles
    uint8_t* cellg = smem_gen + (size_t)cw * 20 * kCellTile;
    const int gw = blockIdx.x * 8 + cw, n_gw = gridDim.x * 8;
    for (int s = 0; s < n_steps; ++s) {
      const int t = p.t_top - s;
      // this step's cell tiles stream from HBM (saved activations, c, dout, dscale): pull them into L2 while phase A runs
      if (lane == 0) {
        for (int it = gw; it < n_cells; it += n_gw) {
          const int row0 = (it / chunks) * kCellRows, u0 = (it % chunks) * kCellUnits;
#pragma unroll
          for (int g = 0; g < 4; ++g) tma_prefetch_l2_2d(&map_g, g * R + u0, t * B + row0);
          tma_prefetch_l2_2d(&map_c, u0, t * B + row0);
          if (s == 0) tma_prefetch_l2_2d(&map_c, u0, (t + 1) * B + row0);
          if (p.dout) tma_prefetch_l2_2d(&map_do, u0, t * B + row0);
          if (p.dscale) tma_prefetch_l2_2d(&map_ds, u0, t * B + row0);
        }
      }
      __syncwarp();
      if (has_item) {
        if (warp >= 8) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            uint8_t* base = smem_gen + (size_t)stage * STAGE_BYTES;
            const float4* a_raw = reinterpret_cast<const float4*>(base);
            float4* a_lo = reinterpret_cast<float4*>(base + A_BYTES);
            const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * A_BYTES);
            float4* b_lo = reinterpret_cast<float4*>(base + 2 * A_BYTES + B_BYTES);
            if (p.bf16x) {
              convert_bf16_tiles<false, 128>(base, base + A_BYTES, tc, true, p.bf16x == 2);
              if (!p.b_pre) convert_bf16_tiles<false, 128>(base + 2 * A_BYTES, base + 2 * A_BYTES + B_BYTES, tc, true, p.bf16x == 2);
            } else {
#pragma unroll 4
              for (int i = tc; i < A_BYTES / 16; i += 128) a_lo[i] = tf32_lo4(a_raw[i]);
#pragma unroll 4
              for (int i = tc; i < B_BYTES / 16; i += 128) b_lo[i] = tf32_lo4(b_raw[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(conv_leader + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        // ---- phase-A epilogue: this CTA's 128 rows, columns [eh*128, +128) of the tile -> 4 swizzled tiles -> reduce-add
        mbar_wait(bar_tfull, acc_phase);
        tc_fence_after();
        if (threadIdx.x == 128) MNN_TRACE(3);
        {
          const int r = q * 32 + lane, sw = r & 7;
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float v[32], vx[32];
            tmem_ld32(tacc + (uint32_t)(eh * 128 + c * 32), v);
            tmem_ld32(tacc + (uint32_t)(NA + eh * 128 + c * 32), vx);
            uint8_t* rowp = smem_gen + (size_t)(eh * 4 + c) * TILE + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) =
                  make_float4(v[4 * j] + vx[4 * j], v[4 * j + 1] + vx[4 * j + 1], v[4 * j + 2] + vx[4 * j + 2],
                              v[4 * j + 3] + vx[4 * j + 3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          if (eh == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
          if (lane == 0 && q == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              tma_reduce_add_2d(&map_dhr, smem0 + (uint32_t)(eh * 4 + c) * TILE, nt * NA + eh * 128 + c * 32,
                                mA * 2 * BM + (int)rank * BM);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cntA + mA) : "memory");
            if (threadIdx.x == 128) MNN_TRACE(4);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
        }
        acc_phase ^= 1;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves' partial tiles are out of shared memory
      }
      // ---- phase B: cell backward on [16 rows x 16 units] tiles, one warp per tile; the loads of the warp's next tile
      // are in flight while the current one is computed, and the stores are only waited for once, at the end
      {
        auto issue_loads = [&](int it, int buf) {   // lane 0
          const int uc = it % chunks, rb = it / chunks;
          const int row0 = rb * kCellRows, u0 = uc * kCellUnits;
          const int m = row0 / (2 * BM);
          const unsigned int target = (unsigned int)(s + 1) * perA;
          while (ld_acquire(p.cntA + m) < target) __nanosleep(20);
          asm volatile("fence.proxy.async;" ::: "memory");
          if (threadIdx.x == 128 && it == gw) MNN_TRACE(5);
          const uint32_t bar = my_cbar + 8 * buf, dst = cell0 + (uint32_t)buf * 10 * kCellTile;
          const int ntl = 8 + (p.dout ? 1 : 0) + (p.dscale ? 1 : 0);
          mbar_expect_tx(bar, ntl * kCellTile);
          tma_load_2d(dst, &map_dh, bar, u0, row0);
#pragma unroll
          for (int g = 0; g < 4; ++g) tma_load_2d(dst + (1 + g) * kCellTile, &map_g, bar, g * R + u0, t * B + row0);
          tma_load_2d(dst + 5 * kCellTile, &map_c, bar, u0, t * B + row0);
          tma_load_2d(dst + 6 * kCellTile, &map_c, bar, u0, (t + 1) * B + row0);
          if (p.dout) tma_load_2d(dst + 7 * kCellTile, &map_do, bar, u0, t * B + row0);
          if (p.dscale) tma_load_2d(dst + 8 * kCellTile, &map_ds, bar, u0, t * B + row0);
          tma_load_2d(dst + 9 * kCellTile, &map_dc, bar, u0, row0);
        };
        if (lane == 0 && gw < n_cells) issue_loads(gw, 0);
        int kk = 0;
        for (int it = gw; it < n_cells; it += n_gw, ++kk) {
          const int buf = kk & 1;
          const int uc = it % chunks, rb = it / chunks;
          const int row0 = rb * kCellRows, u0 = uc * kCellUnits;
          if (lane == 0 && it + n_gw < n_cells) {
            // the other buffer was last used by tile kk-1: its stores must have read shared memory
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_loads(it + n_gw, buf ^ 1);
          }
          mbar_wait(my_cbar + 8 * buf, cphase[buf]);
          cphase[buf] ^= 1;
          if (threadIdx.x == 128 && it == gw) MNN_TRACE(6);
          // elementwise: a tile is 64 chunks of 16 bytes, lane L takes chunks L and L + 32 of every tile (conflict-free)
          uint8_t* tb = cellg + (size_t)buf * 10 * kCellTile + lane * 16;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            auto ld4 = [&](int tile) { return *reinterpret_cast<const float4*>(tb + tile * kCellTile + h * 512); };
            auto st4 = [&](int tile, float4 v) { *reinterpret_cast<float4*>(tb + tile * kCellTile + h * 512) = v; };
            const float4 Dh = ld4(0), Gi = ld4(1), Gj = ld4(2), Gf = ld4(3), Go = ld4(4), Cp = ld4(5), Cn = ld4(6), Dc = ld4(9);
            float dh[4] = {Dh.x, Dh.y, Dh.z, Dh.w};
            const float gi[4] = {Gi.x, Gi.y, Gi.z, Gi.w}, gj[4] = {Gj.x, Gj.y, Gj.z, Gj.w}, gf[4] = {Gf.x, Gf.y, Gf.z, Gf.w};
            const float go[4] = {Go.x, Go.y, Go.z, Go.w}, cp[4] = {Cp.x, Cp.y, Cp.z, Cp.w}, cn[4] = {Cn.x, Cn.y, Cn.z, Cn.w};
            const float dc[4] = {Dc.x, Dc.y, Dc.z, Dc.w};
            if (p.dout) {
              const float4 D = ld4(7);
              const float d4[4] = {D.x, D.y, D.z, D.w};
              if (p.dscale) {
                const float4 S = ld4(8);
                const float s4[4] = {S.x, S.y, S.z, S.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) dh[i] = fmaf(d4[i], s4[i], dh[i]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) dh[i] += d4[i];
              }
            }
            float di[4], dj[4], df[4], d_o[4], dcn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float tcn = tanh_mufu(cn[i]);
              const float dcc = dc[i] + dh[i] * go[i] * (1.f - tcn * tcn);
              di[i] = dcc * gj[i] * gi[i] * (1.f - gi[i]);
              dj[i] = dcc * gi[i] * (1.f - gj[i] * gj[i]);
              df[i] = dcc * cp[i] * gf[i] * (1.f - gf[i]);
              d_o[i] = dh[i] * tcn * go[i] * (1.f - go[i]);
              dcn[i] = dcc * gf[i];
            }
            st4(1, make_float4(di[0], di[1], di[2], di[3]));
            st4(2, make_float4(dj[0], dj[1], dj[2], dj[3]));
            st4(3, make_float4(df[0], df[1], df[2], df[3]));
            st4(4, make_float4(d_o[0], d_o[1], d_o[2], d_o[3]));
            st4(9, make_float4(dcn[0], dcn[1], dcn[2], dcn[3]));
            st4(0, make_float4(0.f, 0.f, 0.f, 0.f));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (threadIdx.x == 128 && it == gw) MNN_TRACE(7);
          if (lane == 0) {
            const uint32_t src = cell0 + (uint32_t)buf * 10 * kCellTile;
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_store_2d(&map_g, src + (1 + g) * kCellTile, g * R + u0, t * B + row0);
            tma_store_2d(&map_dc, src + 9 * kCellTile, u0, row0);
            tma_store_2d(&map_dh, src, u0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (lane == 0 && gw < n_cells) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every tile of this warp is written
          asm volatile("fence.proxy.async;" ::: "memory");
          for (int it = gw; it < n_cells; it += n_gw) {
            const int m = ((it / chunks) * kCellRows) / (2 * BM);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cntB + m) : "memory");
          }
          if (threadIdx.x == 128) MNN_TRACE(8);
        }
        __syncwarp();
      }
      if (threadIdx.x == 128) MNN_TRACE(9);
      if (lane == 0) mbar_arrive(smem_u32(&bfree));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// WhP^T[(n*4 + g)*UB + u][k] = Wh[k][g*R + n*UB + u]  (zero rows for units >= R)
__global__ void lstm_prep_wh_kernel(const float* __restrict__ wh, float* __restrict__ whp, int R, int UB, int blocks) {
  const size_t total = (size_t)blocks * 4 * UB * R;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % R);
    const int row = (int)(idx / R);
    const int u = row % UB, g = (row / UB) & 3, n = row / (4 * UB);
    const int unit = n * UB + u;
    whp[idx] = unit < R ? __ldg(wh + (size_t)k * 4 * R + (size_t)g * R + unit) : 0.f;
  }
}

// the same permutation written as bf16 pair planes [rows][R] (hi, then lo right behind it) for the pre-split operand path
__global__ void lstm_prep_wh_pair_kernel(const float* __restrict__ wh, uint16_t* __restrict__ planes, int R, int UB, int blocks) {
  const size_t total = (size_t)blocks * 4 * UB * R;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % R);
    const int row = (int)(idx / R);
    const int u = row % UB, g = (row / UB) & 3, n = row / (4 * UB);
    const int unit = n * UB + u;
    const float w = unit < R ? __ldg(wh + (size_t)k * 4 * R + (size_t)g * R + unit) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    planes[idx] = __bfloat16_as_ushort(hi);
    planes[total + idx] = __bfloat16_as_ushort(__float2bfloat16_rn(w - __bfloat162float(hi)));
  }
}

// dG_{T-1} from dout only (no recurrent term): plain elementwise pass before the fused steps
__global__ void lstm_last_bwd_kernel(float* gates, const float* cprev, const float* c, const float* dout,
                                     const float* dscale, float* dc, int B, int R) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * R) return;
  const int b = idx / R, r = idx - b * R;
  float* g = gates + (size_t)b * 4 * R;
  const float gi = g[r], gj = g[R + r], gf = g[2 * R + r], go = g[3 * R + r];
  float dh = 0.f;
  if (dout) dh = dscale ? dout[idx] * dscale[idx] : dout[idx];
  const float tcn = tanh_acc(c[idx]);
  const float dcc = dh * go * (1.f - tcn * tcn);
  g[r] = dcc * gj * gi * (1.f - gi);
  g[R + r] = dcc * gi * (1.f - gj * gj);
  g[2 * R + r] = dcc * cprev[idx] * gf * (1.f - gf);
  g[3 * R + r] = dh * tcn * go * (1.f - go);
  dc[idx] = dcc * gf;
}

template <int BN, bool FWD>
static int launch_lstm(const CUtensorMap& ma, const CUtensorMap& mb, LstmParams p, int persistent, cudaStream_t stream) {
  using C_ = LCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(lstm_tc_kernel<BN, FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM);
    cudaFuncSetAttribute(lstm_tc_kernel<BN, FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM);
    attr_set = true;
  }
  const int items = p.slabs * p.blocks;
  const int grid = items < mnn_tc_num_sms() ? items : mnn_tc_num_sms();
  const int t0 = p.t0, t1 = p.t1;
  static const char* share_env = getenv("MNN_LSTM_SHARE");   // "0": never (debug)
  const bool share = items <= grid && !(share_env && share_env[0] == '0');
  if (persistent && t1 - t0 > 1) {
    cudaMemsetAsync(p.flags, 0, (size_t)p.slabs * sizeof(unsigned int), stream);
    static const char* trace_path = getenv("MNN_LSTM_TRACE");   // debug only: allocates and synchronises
    const size_t trace_n = (size_t)grid * kTraceSteps * kTraceEv;
    if (trace_path) {
      cudaMalloc(&p.trace, trace_n * sizeof(unsigned long long));
      cudaMemsetAsync(p.trace, 0, trace_n * sizeof(unsigned long long), stream);
    }
    void* args[3] = {const_cast<CUtensorMap*>(&ma), const_cast<CUtensorMap*>(&mb), &p};
    cudaError_t e = cudaLaunchCooperativeKernel(
        share ? reinterpret_cast<void*>(lstm_tc_kernel<BN, FWD, true>) : reinterpret_cast<void*>(lstm_tc_kernel<BN, FWD, false>),
        dim3(grid), dim3(kLThreads), args, C_::SMEM, stream);
    if (e != cudaSuccess) {
      mnn_set_error(cudaGetErrorString(e));
      return (int)e;
    }
    if (trace_path) {
      cudaStreamSynchronize(stream);
      std::vector<unsigned long long> h(trace_n);
      cudaMemcpy(h.data(), p.trace, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      cudaFree(p.trace);
      if (FILE* f = fopen(trace_path, "a")) {
        fprintf(f, "# lstm 1cta %s B=%d R=%d T=%d ctas=%d BN=%d share=%d\n", FWD ? "fwd" : "bwd", p.B, p.R, t1 - t0, grid, BN,
                (int)share);
        for (int c = 0; c < grid; ++c)
          for (int st = 0; st < kTraceSteps; ++st) {
            fprintf(f, "%d %d", c, st);
            for (int ev = 0; ev < kTraceEv; ++ev) fprintf(f, " %llu", h[((size_t)c * kTraceSteps + st) * kTraceEv + ev]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
    return mnn_check_launch(FWD ? "lstm_seq_fwd(persistent)" : "lstm_seq_bwd(persistent)");
  }
  if (FWD) {
    for (int t = t0; t < t1; ++t) {
      p.t0 = t; p.t1 = t + 1;
      if (share) lstm_tc_kernel<BN, FWD, true><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
      else lstm_tc_kernel<BN, FWD, false><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
    }
  } else {
    for (int t = t1 - 1; t >= t0; --t) {
      p.t0 = t; p.t1 = t + 1;
      if (share) lstm_tc_kernel<BN, FWD, true><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
      else lstm_tc_kernel<BN, FWD, false><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
    }
  }
  return mnn_check_launch(FWD ? "lstm_seq_fwd" : "lstm_seq_bwd", t1 - t0);
}

static int fwd_unit_block(int B, int R) {
  const int slabs = (B + BM - 1) / BM;
  // an SM budget below the number of 16-unit items: 32-unit items, half as many CTAs (the caller wants the SMs for the
  // bulk work it runs beside the recurrence)
  const int budget = mnn_tc_sm_budget();
  if (budget > 0 && slabs * ((R + 15) / 16) > budget) return 32;
  return (slabs * ((R + 31) / 32) >= 96) ? 32 : 16;
}

}  // namespace tc
}  // namespace mnn

using namespace mnn;
using namespace mnn::tc;

// bytes reserved for the permuted recurrent weights at the head of the workspace (largest unit-block padding)
static size_t whp_region_bytes(int R) {
  size_t mx = 0;
  for (int UB : {16, 32, kL2UB}) {
    const size_t b = (size_t)((R + UB - 1) / UB) * 4 * UB * R * sizeof(float);
    if (b > mx) mx = b;
  }
  return (mx + 255) / 256 * 256;
}

// workspace: [permuted weights | 4 KB of counters: flags @0, cntA @1024, cntB @2048 | dh_acc[B,R]]
constexpr size_t kCounterBytes = 4096;
// SMs the weight-resident forward kernel may occupy: the calling thread's SM budget (wavefront / chunk pipeline) or all
static int res_sms() {
  const int n = mnn_tc_num_sms(), b = mnn_tc_sm_budget();
  return (b > 0 && b < n) ? b : n;
}
// ... | split h double buffer of the weight-resident forward kernel (lstm_res.cu)]
static size_t res_region_offset(int B, int R) {
  return (whp_region_bytes(R) + kCounterBytes + (size_t)B * R * sizeof(float) + 255) / 256 * 256;
}
extern "C" size_t mnn_lstm_workspace_bytes(int B, int R) {
  return res_region_offset(B, R) + mnn_lstm_res_workspace_bytes(B, R);
}

static int pair_bwd_clusters() {
  static int n = -1;
  if (n < 0) {
    cudaFuncSetAttribute(lstm_tc3_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwd3Smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * mnn_tc_num_sms());
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = kBwd3Smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int c = 0;
    if (cudaOccupancyMaxActiveClusters(&c, reinterpret_cast<const void*>(lstm_tc3_bwd_kernel), &cfg) != cudaSuccess) {
      cudaGetLastError();
      c = 0;
    }
    n = c;
  }
  return n;
}

static bool use_pair_bwd(int B, int R, int T, int persistent) {
  static const char* env = getenv("MNN_LSTM_PAIR_BWD");   // "0": never
  if (!persistent || T < 3 || R % 256 != 0 || B % (2 * BM) != 0) return false;
  if (env && env[0] == '0') return false;
  if ((B / (2 * BM)) * (R / 256) > pair_bwd_clusters()) return false;
  if ((B / (2 * BM)) > 256) return false;                 // counters
  return true;
}

// forward on CTA pairs: needs every cluster co-resident (the CTAs of a slab wait on one another)
static int pair_fwd_clusters() {
  static int n = -1;
  if (n < 0) {
    constexpr size_t smem = kPairSmem;
    cudaFuncSetAttribute(lstm_tc2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * mnn_tc_num_sms());
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int c = 0;
    if (cudaOccupancyMaxActiveClusters(&c, reinterpret_cast<const void*>(lstm_tc2_fwd_kernel), &cfg) != cudaSuccess) {
      cudaGetLastError();
      c = 0;
    }
    n = c;
  }
  return n;
}

static bool use_pair_fwd(int B, int R, int T, int persistent) {
  static const char* env = getenv("MNN_LSTM_PAIR");   // "1": force (tests), "0": never
  // TMA tile stores need whole 256-row slabs and whole 64-unit blocks; one item per cluster per step because the
  // epilogue tiles alias the mainloop stages
  if (!persistent || T < 2 || R % kL2UB != 0 || B % (2 * BM) != 0) return false;
  if (env && env[0] == '0') return false;
  if ((B / (2 * BM)) * (R / kL2UB) > pair_fwd_clusters()) return false;
  if (env && env[0] == '1') return true;
  // below 1024 rows the 1-CTA kernel's smaller items win, unless the caller runs the layers as a wavefront (SM budget
  // set): then the pair kernel's few CTAs leave room for the other layer
  return B >= 1024 || (mnn_tc_sm_budget() > 0 && B >= 512);
}

extern "C" int mnn_lstm_tc_supported(int B, int R) { return R % 8 == 0 && R >= 8 && B > 0; }

// CTAs (= SMs, one CTA per SM) the persistent recurrence kernels occupy under the calling thread's SM budget: lets the
// host size the budget of the work it runs beside them (time-chunk pipeline).
extern "C" int mnn_lstm_seq_fwd_ctas(int T, int B, int R) {
  if (!mnn_lstm_tc_supported(B, R)) return 0;
  const int sms = mnn_tc_num_sms();
  if (const int n = mnn_lstm_res_ctas(T, B, R, res_sms())) return n;
  if (use_pair_fwd(B, R, T, 1)) {
    const int items = (B / (2 * BM)) * (R / kL2UB);
    return 2 * (items < pair_fwd_clusters() ? items : pair_fwd_clusters());
  }
  const int UB = fwd_unit_block(B, R);
  const int items = ((B + BM - 1) / BM) * ((R + UB - 1) / UB);
  return items < sms ? items : sms;
}
extern "C" int mnn_lstm_seq_bwd_ctas(int T, int B, int R) {
  if (!mnn_lstm_tc_supported(B, R)) return 0;
  const int sms = mnn_tc_num_sms();
  if (const int n = mnn_lstm_res_bwd_ctas(T - 1, B, R, res_sms())) return n;
  if (use_pair_bwd(B, R, T, 1)) {
    const int need = (B / (2 * BM)) * (R / 256);
    int clusters = pair_bwd_clusters();
    if (mnn_tc_sm_budget() > 0 && clusters > mnn_tc_sm_budget() / 2) clusters = mnn_tc_sm_budget() / 2;
    if (clusters < need) clusters = need;
    return 2 * clusters;
  }
  const int slabs = (B + BM - 1) / BM;
  const int BN = (slabs * ((R + 63) / 64) >= 96) ? 64 : 32;
  const int items = slabs * ((R + BN - 1) / BN);
  return items < sms ? items : sms;
}

extern "C" int mnn_lstm_seq_fwd_tc(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale,
                                   const float* u, float keep, unsigned long long seed, int T, int B, int R, void* ws,
                                   int persistent, cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && hbuf && cbuf && ws, MNN_ERR_ARG, "lstm_seq_fwd_tc: null pointer");
  MNN_REQUIRE(T > 0 && mnn_lstm_tc_supported(B, R), MNN_ERR_UNSUPPORTED, "lstm_seq_fwd_tc: needs num_units % 8 == 0");
  MNN_REQUIRE(!(out && keep < 1.f && !dscale), MNN_ERR_ARG, "lstm_seq_fwd_tc: dscale required when keep < 1");
  if (persistent && mnn_lstm_res_ctas(T, B, R, res_sms())) {
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    return mnn_lstm_res_fwd(gates, wh, hbuf, cbuf, out, dscale, u, keep, seed, T, B, R, w8 + res_region_offset(B, R),
                            reinterpret_cast<unsigned int*>(w8 + whp_region_bytes(R)), res_sms(), stream);
  }
  const bool pair = use_pair_fwd(B, R, T, persistent);
  const int UB = pair ? kL2UB : fwd_unit_block(B, R);
  const int blocks = (R + UB - 1) / UB;
  float* whp = reinterpret_cast<float*>(ws);
  unsigned int* flags = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(ws) + whp_region_bytes(R));
  // MNN_LSTM_BF16X: "0" 3xTF32, "1" 2.5 products, "2" bf16 pairs, "3" bf16 pairs with the weights pre-split once per launch
  // (pair kernels only). Unset: "3" while the calling thread is in the training step's pair mode (mnn_set_gemm_split(1):
  // 26.1 -> 22.2 us per step at B = 2048, R = 512), "1" otherwise (evaluate / generate keep ~2^-20 per product)
  static const char* bf16x_env0 = getenv("MNN_LSTM_BF16X");
  const bool wh_pre = pair && (bf16x_env0 ? bf16x_env0[0] == '3' : mnn_tc_gemm_split() == 1);
  if (wh_pre) lstm_prep_wh_pair_kernel<<<296, 256, 0, stream>>>(wh, reinterpret_cast<uint16_t*>(ws), R, UB, blocks);
  else lstm_prep_wh_kernel<<<296, 256, 0, stream>>>(wh, whp, R, UB, blocks);
  int rc = mnn_check_launch("lstm_prep_wh");
  if (rc) return rc;
  if (pair) {
    LstmParams p{};
    p.gates = gates; p.hbuf = hbuf; p.cbuf = cbuf; p.out = out; p.dscale = dscale; p.u = u; p.keep = keep; p.seed = seed; p.rmap = current_row_map();
    p.T = T; p.B = B; p.R = R; p.t0 = 0; p.t1 = T;
    p.slabs = (B + 2 * BM - 1) / (2 * BM); p.blocks = blocks; p.kb_total = (R + BK - 1) / BK; p.flags = flags;
    static const char* bf16x_env = getenv("MNN_LSTM_BF16X");   // "0": 3xTF32, "1": 2.5 products, "2": bf16 pairs
    p.bf16x = bf16x_env ? (bf16x_env[0] == '0' ? 0 : (bf16x_env[0] >= '2' ? 2 : 1)) : (wh_pre ? 2 : 1);
    p.b_pre = wh_pre ? 1 : 0;
    CUtensorMap ma, mb, mb2;
    rc = mnn_tc_make_map(hbuf, R, R, (long long)(T + 1) * B, BM, false, &ma);
    if (rc) return rc;
    if (wh_pre) {
      const long long rows = (long long)blocks * 4 * UB;
      rc = mnn_tc_make_map_bf16(ws, R, R, rows, BM, &mb);
      if (!rc) rc = mnn_tc_make_map_bf16(reinterpret_cast<uint16_t*>(ws) + (size_t)rows * R, R, R, rows, BM, &mb2);
    } else {
      rc = mnn_tc_make_map(whp, R, R, (long long)blocks * 4 * UB, BM, false, &mb);
      mb2 = mb;
    }
    if (rc) return rc;
    CUtensorMap mg, mc, mo;
    rc = mnn_tc_make_map(gates, 4LL * R, 4LL * R, (long long)T * B, BM, false, &mg);
    if (rc) return rc;
    rc = mnn_tc_make_map(cbuf, R, R, (long long)(T + 1) * B, BM, false, &mc);
    if (rc) return rc;
    rc = mnn_tc_make_map(out ? out : hbuf, R, R, (long long)T * B, BM, false, &mo);
    if (rc) return rc;
    cudaMemsetAsync(flags, 0, (size_t)p.slabs * sizeof(unsigned int), stream);
    const int items = p.slabs * p.blocks;
    static const char* trace_path = getenv("MNN_LSTM_TRACE");   // debug only: allocates and synchronises
    const size_t trace_n = (size_t)2 * mnn_tc_num_sms() * kTraceSteps * kTraceEv;
    if (trace_path) {
      cudaMalloc(&p.trace, trace_n * sizeof(unsigned long long));
      cudaMemsetAsync(p.trace, 0, trace_n * sizeof(unsigned long long), stream);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (items < pair_fwd_clusters() ? items : pair_fwd_clusters()));   // all clusters co-resident
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = kPairSmem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc2_fwd_kernel, ma, mb, mb2, mg, mc, mo, p);
    if (e != cudaSuccess) {
      mnn_set_error(cudaGetErrorString(e));
      return (int)e;
    }
    if (trace_path) {
      cudaStreamSynchronize(stream);
      std::vector<unsigned long long> h(trace_n);
      cudaMemcpy(h.data(), p.trace, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      cudaFree(p.trace);
      if (FILE* f = fopen(trace_path, "a")) {
        fprintf(f, "# lstm pair fwd B=%d R=%d T=%d ctas=%d\n", B, R, T, (int)cfg.gridDim.x);
        for (unsigned c = 0; c < cfg.gridDim.x; ++c)
          for (int st = 0; st < kTraceSteps; ++st) {
            fprintf(f, "%u %d", c, st);
            for (int ev = 0; ev < kTraceEv; ++ev) fprintf(f, " %llu", h[((size_t)c * kTraceSteps + st) * kTraceEv + ev]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
    return mnn_check_launch("lstm_seq_fwd(pair)");
  }

  LstmParams p{};
  p.gates = gates; p.hbuf = hbuf; p.cbuf = cbuf; p.out = out; p.dscale = dscale; p.u = u; p.keep = keep; p.seed = seed; p.rmap = current_row_map();
  p.T = T; p.B = B; p.R = R; p.t0 = 0; p.t1 = T;
  p.slabs = (B + BM - 1) / BM; p.blocks = blocks; p.kb_total = (R + BK - 1) / BK; p.flags = flags;
  CUtensorMap ma, mb;
  rc = mnn_tc_make_map(hbuf, R, R, (long long)(T + 1) * B, BM, false, &ma);
  if (rc) return rc;
  rc = mnn_tc_make_map(whp, R, R, (long long)blocks * 4 * UB, 4 * UB, false, &mb);
  if (rc) return rc;
  if (UB == 32) return launch_lstm<128, true>(ma, mb, p, persistent, stream);
  return launch_lstm<64, true>(ma, mb, p, persistent, stream);
}

extern "C" int mnn_lstm_seq_bwd_tc_chunk(float* gates, const float* wh, const float* cbuf, const float* dout,
                                         const float* dscale, float* dc_work, int T, int B, int R, void* ws, int persistent,
                                         int has_next, cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && cbuf && dc_work && ws, MNN_ERR_ARG, "lstm_seq_bwd_tc: null pointer");
  // has_next: this is a time chunk [t0, t0+T) of a longer sequence whose later chunk has already been back-propagated:
  // gates slot T holds dG of the next step and dc_work the carried cell gradient, so step T-1 is an ordinary step
  MNN_REQUIRE(T > 0 && mnn_lstm_tc_supported(B, R), MNN_ERR_UNSUPPORTED, "lstm_seq_bwd_tc: needs num_units % 8 == 0");
  const size_t BR = (size_t)B * R;
  const int n = B * R;
  int rc = MNN_OK;
  if (!has_next) {
    lstm_last_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gates + (size_t)(T - 1) * B * 4 * R, cbuf + (size_t)(T - 1) * BR,
                                                           cbuf + (size_t)T * BR, dout ? dout + (size_t)(T - 1) * BR : nullptr,
                                                           dscale ? dscale + (size_t)(T - 1) * BR : nullptr, dc_work, B, R);
    rc = mnn_check_launch("lstm_last_bwd");
    if (rc || T == 1) return rc;
  }
  const int t_top = has_next ? T - 1 : T - 2;

  if (persistent && mnn_lstm_res_bwd_ctas(t_top + 1, B, R, res_sms())) {
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    return mnn_lstm_res_bwd(gates, wh, cbuf, dout, dscale, dc_work, t_top, t_top + 1, B, R, w8 + res_region_offset(B, R),
                            reinterpret_cast<unsigned int*>(w8 + whp_region_bytes(R)), stream);
  }
  if (use_pair_bwd(B, R, T, persistent)) {
    uint8_t* cnt = reinterpret_cast<uint8_t*>(ws) + whp_region_bytes(R);
    float* dh_acc = reinterpret_cast<float*>(cnt + kCounterBytes);
    Lstm3Params q{};
    q.dout = dout; q.dscale = dscale; q.T = T; q.B = B; q.R = R; q.t_top = t_top;
    q.slabs = B / (2 * BM); q.ntiles = R / 256;
    int clusters = pair_bwd_clusters();
    if (mnn_tc_sm_budget() > 0 && clusters > mnn_tc_sm_budget() / 2) clusters = mnn_tc_sm_budget() / 2;
    if (clusters < q.slabs * q.ntiles) clusters = q.slabs * q.ntiles;
    const int kb_total = 4 * R / BK;
    int splits = clusters / (q.slabs * q.ntiles);
    if (splits > kb_total / 2) splits = kb_total / 2;
    if (splits < 1) splits = 1;
    q.kb_per_split = (kb_total + splits - 1) / splits;
    q.splits = (kb_total + q.kb_per_split - 1) / q.kb_per_split;
    static const char* bf16x_env = getenv("MNN_LSTM_BF16X");   // see mnn_lstm_seq_fwd_tc
    const bool wh_pre = bf16x_env ? bf16x_env[0] == '3' : mnn_tc_gemm_split() == 1;
    q.bf16x = bf16x_env ? (bf16x_env[0] == '0' ? 0 : (bf16x_env[0] >= '2' ? 2 : 1)) : (wh_pre ? 2 : 1);
    q.b_pre = wh_pre ? 1 : 0;
    q.cntA = reinterpret_cast<unsigned int*>(cnt + 1024);
    q.cntB = reinterpret_cast<unsigned int*>(cnt + 2048);
    cudaMemsetAsync(cnt + 1024, 0, 2048, stream);
    cudaMemsetAsync(dh_acc, 0, BR * sizeof(float), stream);
    CUtensorMap ma, mb, mb2, mdhr, mdh, mg, mc, mdo, mds, mdc;
    const long long TB = (long long)T * B;
    if ((rc = mnn_tc_make_map(gates, 4LL * R, 4LL * R, TB + (has_next ? B : 0), BM, false, &ma))) return rc;
    if (wh_pre) {
      // Wh [R][4R] as bf16 pair planes in the (otherwise unused) permuted-weights region at the head of the workspace
      if ((rc = mnn_split_bf16_pair(wh, 4LL * R, R, 4 * R, ws, 4LL * R, stream))) return rc;
      if ((rc = mnn_tc_make_map_bf16(ws, 4LL * R, 4LL * R, R, BM, &mb))) return rc;
      if ((rc = mnn_tc_make_map_bf16(reinterpret_cast<uint16_t*>(ws) + (size_t)R * 4 * R, 4LL * R, 4LL * R, R, BM, &mb2))) return rc;
    } else {
      if ((rc = mnn_tc_make_map(wh, 4LL * R, 4LL * R, R, BM, false, &mb))) return rc;
      mb2 = mb;
    }
    if ((rc = mnn_tc_make_map(dh_acc, R, R, B, BM, false, &mdhr))) return rc;
    if ((rc = mnn_tc_make_map_plain(dh_acc, R, R, B, kCellUnits, kCellRows, &mdh))) return rc;
    if ((rc = mnn_tc_make_map_plain(gates, 4LL * R, 4LL * R, TB, kCellUnits, kCellRows, &mg))) return rc;
    if ((rc = mnn_tc_make_map_plain(cbuf, R, R, TB + B, kCellUnits, kCellRows, &mc))) return rc;
    if ((rc = mnn_tc_make_map_plain(dout ? dout : cbuf, R, R, TB, kCellUnits, kCellRows, &mdo))) return rc;
    if ((rc = mnn_tc_make_map_plain(dscale ? dscale : cbuf, R, R, TB, kCellUnits, kCellRows, &mds))) return rc;
    if ((rc = mnn_tc_make_map_plain(dc_work, R, R, B, kCellUnits, kCellRows, &mdc))) return rc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = kBwd3Smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    static const char* trace_path = getenv("MNN_LSTM_TRACE");   // debug only: allocates and synchronises
    const size_t trace_n = (size_t)2 * mnn_tc_num_sms() * kTraceSteps * kTraceEv;
    if (trace_path) {
      cudaMalloc(&q.trace, trace_n * sizeof(unsigned long long));
      cudaMemsetAsync(q.trace, 0, trace_n * sizeof(unsigned long long), stream);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc3_bwd_kernel, ma, mb, mb2, mdhr, mdh, mg, mc, mdo, mds, mdc, q);
    if (e != cudaSuccess) {
      mnn_set_error(cudaGetErrorString(e));
      return (int)e;
    }
    if (trace_path) {
      cudaStreamSynchronize(stream);
      std::vector<unsigned long long> h(trace_n);
      cudaMemcpy(h.data(), q.trace, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      cudaFree(q.trace);
      if (FILE* f = fopen(trace_path, "a")) {
        fprintf(f, "# lstm pair bwd B=%d R=%d T=%d ctas=%d splits=%d\n", B, R, T, (int)cfg.gridDim.x, q.splits);
        for (unsigned c = 0; c < cfg.gridDim.x; ++c)
          for (int st = 0; st < kTraceSteps; ++st) {
            fprintf(f, "%u %d", c, st);
            for (int ev = 0; ev < kTraceEv; ++ev) fprintf(f, " %llu", h[((size_t)c * kTraceSteps + st) * kTraceEv + ev]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
    return mnn_check_launch("lstm_seq_bwd(pair)");
  }

  const int slabs = (B + BM - 1) / BM;
  const int BN = (slabs * ((R + 63) / 64) >= 96) ? 64 : 32;
  const size_t whp_bytes = whp_region_bytes(R);
  LstmParams p{};
  p.gates = gates; p.cbuf = const_cast<float*>(cbuf); p.dout = dout; p.dscale = const_cast<float*>(dscale); p.dc = dc_work;
  p.T = T; p.B = B; p.R = R; p.t0 = 0; p.t1 = t_top + 1;
  p.slabs = slabs; p.blocks = (R + BN - 1) / BN; p.kb_total = (4 * R + BK - 1) / BK;
  p.flags = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(ws) + whp_bytes);
  CUtensorMap ma, mb;
  rc = mnn_tc_make_map(gates, 4LL * R, 4LL * R, (long long)(T + (has_next ? 1 : 0)) * B, BM, false, &ma);
  if (rc) return rc;
  rc = mnn_tc_make_map(wh, 4LL * R, 4LL * R, R, BN, false, &mb);
  if (rc) return rc;
  if (BN == 64) return launch_lstm<64, false>(ma, mb, p, persistent, stream);
  return launch_lstm<32, false>(ma, mb, p, persistent, stream);
}

extern "C" int mnn_lstm_seq_bwd_tc(float* gates, const float* wh, const float* cbuf, const float* dout,
                                   const float* dscale, float* dc_work, int T, int B, int R, void* ws, int persistent,
                                   cudaStream_t stream) {
  return mnn_lstm_seq_bwd_tc_chunk(gates, wh, cbuf, dout, dscale, dc_work, T, B, R, ws, persistent, 0, stream);
}
