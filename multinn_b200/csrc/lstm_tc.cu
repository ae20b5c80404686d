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
#include "multinn_b200.h"
#include "tc_common.cuh"

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
  int T, B, R;
  int t0, t1;           // fwd: steps t0..t1-1 ascending; bwd: steps t1-1..t0 descending
  int slabs, blocks;    // work items per step
  int kb_total;         // k-blocks per item
  unsigned int* flags;  // [slabs] completed-item counters (persistent launches), zeroed by the host
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Shared skeleton of both directions. FWD: A = h slot t (rows t*B + m*128), B = WhP^T rows n*BN..; BWD: A = dG slot t+1,
// B = Wh rows n*BN.. (both operands K-major).
template <int BN, bool FWD>
__global__ void __launch_bounds__(kLThreads, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const LstmParams p) {
  using C_ = LCfg<BN>;
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
      mbar_init(bar_conv + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 128);
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
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ converters: lo = x - trunc_tf32(x)
    const int tc = threadIdx.x - 8 * 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < n_steps; ++s) {
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          uint8_t* base = smem_gen + (size_t)stage * C_::STAGE_BYTES;
          const float4* a_raw = reinterpret_cast<const float4*>(base);
          float4* a_lo = reinterpret_cast<float4*>(base + C_::A_BYTES);
          const float4* b_raw = reinterpret_cast<const float4*>(base + 2 * C_::A_BYTES);
          float4* b_lo = reinterpret_cast<float4*>(base + 2 * C_::A_BYTES + C_::B_BYTES);
#pragma unroll 4
          for (int i = tc; i < C_::A_BYTES / 16; i += 128) a_lo[i] = tf32_lo4(a_raw[i]);
#pragma unroll 4
          for (int i = tc; i < C_::B_BYTES / 16; i += 128) b_lo[i] = tf32_lo4(b_raw[i]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(bar_conv + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: the LSTM cell (fwd) / its backward
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
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
        if (persistent && s + 1 < n_steps && row_ok) {
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
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN);
        if (FWD) {
          constexpr int UB = BN / 4;
          float* gp = p.gates + ((size_t)t * B + (row_ok ? b : 0)) * 4 * R;
          const size_t sidx = (size_t)(row_ok ? b : 0) * R;
          const float* cprev = p.cbuf + (size_t)t * BR + sidx;
          float* cnew = p.cbuf + (size_t)(t + 1) * BR + sidx;
          float* hnew = p.hbuf + (size_t)(t + 1) * BR + sidx;
#pragma unroll 1
          for (int ug = 0; ug < UB / 8; ++ug) {
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
                const float gi = sigmoid_fast(pre[0][i]), gj = tanh_fast(pre[1][i]);
                const float gf = sigmoid_fast(pre[2][i]), go = sigmoid_fast(pre[3][i]);
                pre[0][i] = gi; pre[1][i] = gj; pre[2][i] = gf; pre[3][i] = go;
                cv[i] = gj * gi + cp[i] * gf;
                hv[i] = tanh_fast(cv[i]) * go;
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
                      const unsigned long long ctr = (unsigned long long)(e >> 2) + h4;
                      const uint4 r4 = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                                     make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
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
          for (int ug = 0; ug < BN / 8; ++ug) {
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
        if (persistent) {
          // publish this item's h_t / dG_t rows to the CTAs of the same slab
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 4 * 32) {
            asm volatile("fence.proxy.async;" ::: "memory");
            atomicAdd(p.flags + m_blk, 1u);
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
    cudaFuncSetAttribute(lstm_tc_kernel<BN, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM);
    attr_set = true;
  }
  const int items = p.slabs * p.blocks;
  const int grid = items < mnn_tc_num_sms() ? items : mnn_tc_num_sms();
  const int t0 = p.t0, t1 = p.t1;
  if (persistent && t1 - t0 > 1) {
    cudaMemsetAsync(p.flags, 0, (size_t)p.slabs * sizeof(unsigned int), stream);
    void* args[3] = {const_cast<CUtensorMap*>(&ma), const_cast<CUtensorMap*>(&mb), &p};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_tc_kernel<BN, FWD>), dim3(grid),
                                                dim3(kLThreads), args, C_::SMEM, stream);
    if (e != cudaSuccess) {
      mnn_set_error(cudaGetErrorString(e));
      return (int)e;
    }
    return mnn_check_launch(FWD ? "lstm_seq_fwd(persistent)" : "lstm_seq_bwd(persistent)");
  }
  if (FWD) {
    for (int t = t0; t < t1; ++t) {
      p.t0 = t; p.t1 = t + 1;
      lstm_tc_kernel<BN, FWD><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
    }
  } else {
    for (int t = t1 - 1; t >= t0; --t) {
      p.t0 = t; p.t1 = t + 1;
      lstm_tc_kernel<BN, FWD><<<grid, kLThreads, C_::SMEM, stream>>>(ma, mb, p);
    }
  }
  return mnn_check_launch(FWD ? "lstm_seq_fwd" : "lstm_seq_bwd", t1 - t0);
}

static int fwd_unit_block(int B, int R) {
  const int slabs = (B + BM - 1) / BM;
  return (slabs * ((R + 31) / 32) >= 96) ? 32 : 16;
}

}  // namespace tc
}  // namespace mnn

using namespace mnn;
using namespace mnn::tc;

extern "C" size_t mnn_lstm_workspace_bytes(int B, int R) {
  const int UB = fwd_unit_block(B, R);
  const int blocks = (R + UB - 1) / UB;
  const size_t whp = (size_t)blocks * 4 * UB * R * sizeof(float);
  const size_t flags = ((size_t)((B + BM - 1) / BM) * sizeof(unsigned int) + 255) / 256 * 256;
  return (whp + 255) / 256 * 256 + flags;
}

extern "C" int mnn_lstm_tc_supported(int B, int R) { return R % 8 == 0 && R >= 8 && B > 0; }

extern "C" int mnn_lstm_seq_fwd_tc(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale,
                                   const float* u, float keep, unsigned long long seed, int T, int B, int R, void* ws,
                                   int persistent, cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && hbuf && cbuf && ws, MNN_ERR_ARG, "lstm_seq_fwd_tc: null pointer");
  MNN_REQUIRE(T > 0 && mnn_lstm_tc_supported(B, R), MNN_ERR_UNSUPPORTED, "lstm_seq_fwd_tc: needs num_units % 8 == 0");
  MNN_REQUIRE(!(out && keep < 1.f && !dscale), MNN_ERR_ARG, "lstm_seq_fwd_tc: dscale required when keep < 1");
  const int UB = fwd_unit_block(B, R);
  const int blocks = (R + UB - 1) / UB;
  float* whp = reinterpret_cast<float*>(ws);
  const size_t whp_bytes = ((size_t)blocks * 4 * UB * R * sizeof(float) + 255) / 256 * 256;
  unsigned int* flags = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(ws) + whp_bytes);
  lstm_prep_wh_kernel<<<296, 256, 0, stream>>>(wh, whp, R, UB, blocks);
  int rc = mnn_check_launch("lstm_prep_wh");
  if (rc) return rc;

  LstmParams p{};
  p.gates = gates; p.hbuf = hbuf; p.cbuf = cbuf; p.out = out; p.dscale = dscale; p.u = u; p.keep = keep; p.seed = seed;
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

extern "C" int mnn_lstm_seq_bwd_tc(float* gates, const float* wh, const float* cbuf, const float* dout,
                                   const float* dscale, float* dc_work, int T, int B, int R, void* ws, int persistent,
                                   cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && cbuf && dc_work && ws, MNN_ERR_ARG, "lstm_seq_bwd_tc: null pointer");
  MNN_REQUIRE(T > 0 && mnn_lstm_tc_supported(B, R), MNN_ERR_UNSUPPORTED, "lstm_seq_bwd_tc: needs num_units % 8 == 0");
  const size_t BR = (size_t)B * R;
  const int n = B * R;
  lstm_last_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gates + (size_t)(T - 1) * B * 4 * R, cbuf + (size_t)(T - 1) * BR,
                                                         cbuf + (size_t)T * BR, dout ? dout + (size_t)(T - 1) * BR : nullptr,
                                                         dscale ? dscale + (size_t)(T - 1) * BR : nullptr, dc_work, B, R);
  int rc = mnn_check_launch("lstm_last_bwd");
  if (rc || T == 1) return rc;

  const int slabs = (B + BM - 1) / BM;
  const int BN = (slabs * ((R + 63) / 64) >= 96) ? 64 : 32;
  const int UBf = fwd_unit_block(B, R);
  const size_t whp_bytes = ((size_t)((R + UBf - 1) / UBf) * 4 * UBf * R * sizeof(float) + 255) / 256 * 256;
  LstmParams p{};
  p.gates = gates; p.cbuf = const_cast<float*>(cbuf); p.dout = dout; p.dscale = const_cast<float*>(dscale); p.dc = dc_work;
  p.T = T; p.B = B; p.R = R; p.t0 = 0; p.t1 = T - 1;
  p.slabs = slabs; p.blocks = (R + BN - 1) / BN; p.kb_total = (4 * R + BK - 1) / BK;
  p.flags = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(ws) + whp_bytes);
  CUtensorMap ma, mb;
  rc = mnn_tc_make_map(gates, 4LL * R, 4LL * R, (long long)T * B, BM, false, &ma);
  if (rc) return rc;
  rc = mnn_tc_make_map(wh, 4LL * R, 4LL * R, R, BN, false, &mb);
  if (rc) return rc;
  if (BN == 64) return launch_lstm<64, false>(ma, mb, p, persistent, stream);
  return launch_lstm<32, false>(ma, mb, p, persistent, stream);
}
