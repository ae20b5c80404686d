// K6 -- fused autoregressive generation: ALL num_steps steps of RnnEstimator.generate (reference
// models/generators/rnn_estimator.py:271-323: scan of {sample_single, single_step}) in ONE cooperative launch, for the
// batch sizes sample.py really runs (num_songs x intros = 72 rows with the shipped config; one 128-row slab here).
// Per generated step the reference runs, per track, the 84-iteration NADE.sample loop (common/nade.py:231-308,
// rnn_multinade.py:292-317), then MultiRNNCell (common/rnn.py:104-145) and the Dense output layer
// (rnn_nade.py:253-277): hundreds of TF ops; the unfused path here is 8 kernel launches per step.
//
// The CTAs of the grid split into four role groups that hand the step's state over through per-group counters in global
// memory (red.release.gpu / ld.acquire.gpu; nothing is grid-wide), every group keeping ITS weights resident for the
// whole call:
//   S  samplers   CTA = (track m, row group): W_enc[m], W_dec[m] fp32 in shared memory (172 KB), one warp per (row, track):
//                 the same arithmetic as nade_sample_kernel (strict u < p, supplied uniforms or Philox keyed by
//                 (global row, track, dim; step)); writes the sampled frame to the output and, as exact bf16, into the
//                 input half of layer 0's operand buffer
//   L0, L1        CTA = 8 units (32 gate columns) of an LSTM layer: the kernel slice [(in + R) x 32] split once into bf16
//                 w1 + w2 in shared memory (UMMA K-major SWIZZLE_64B); per step TMA streams [input ; h_{t-1}] (published
//                 pre-split by its producers) and 3 tcgen05.mma per K=16 give the gate pre-activations in TMEM; the cell
//                 epilogue (c in registers) publishes h_t split into its own operand buffer and the next group's
//   D  Dense      CTA = 64 output columns of the Dense layer, same pipeline, writes fc (the next step's NADE biases)
// Chain per step: S -> L0 -> (L1) -> D -> S. Operands are bf16 pairs (16 mantissa bits): the recurrent state carries
// ~2^-17 relative rounding per step, so against an fp32 run a Bernoulli draw can flip only when its uniform falls within
// ~1e-5 of the probability (tests pick uniforms with a margin; the multi-launch path stays the fp32-accurate one).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "multinn_b200.h"
#include "tc_common.cuh"

int mnn_tc_make_map_bf16(const void* ptr, long long ld_elems, long long inner, long long outer, int box_outer,
                         CUtensorMap* out);
int mnn_tc_num_sms();

namespace mnn {
namespace gen {

using namespace mnn::tc;

constexpr int kThreads = 512;
constexpr int kMaxStages = 8;
constexpr int A_STAGE = 2 * 128 * 64;   // hi + lo tiles of one k-block of the A operand
constexpr int kSampWarps = kThreads / 32;

struct Layer {
  const float* kern; const float* bias;   // [(in + R), 4R], [4R]
  float* c; float* h;                      // [B, R] state, in/out
  __nv_bfloat16* a1; __nv_bfloat16* a2;    // [2][128][Kp] operand buffer: cols [0,in) input, [inp,inp+R) own h, zeros between
  int in, inp, R, Kp;                      // inp = in padded to 32 (16-byte aligned h stores), Kp = inp + R padded to 32
};

struct GParams {
  Layer l[2];
  int L, B, S, M, D, H, C;                 // C = dense columns = M * (H + D)
  const float* dk; const float* db;        // [R_top, C], [C]
  __nv_bfloat16* d1; __nv_bfloat16* d2;    // [2][128][R_top] Dense operand (h of the top layer)
  float* fc; long long ldfc;               // [B, ldfc]: NADE biases (in: after the intro, out: after the last step)
  const float* w_enc; const float* w_dec;  // [M, D, H]
  const float* u;                          // [S, M, B, D] or null
  int use_philox; unsigned long long seed, offset0; RowMap rmap;
  float* out; long long out_ld, out_step;  // out[b * out_ld + s * out_step + d * M + m]
  unsigned int* flags;                     // [4]: x, h0, h1, fc
  int nS, nL[2], nD;                       // CTAs per group; order in the grid: L0 | L1 | D | S
  int stages[3];                           // ring depth of L0, L1, D
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const unsigned int* f, unsigned int target) {
  while (ld_acquire_u32(f) < target) __nanosleep(20);
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
__device__ __forceinline__ void store_split8(__nv_bfloat16* p1, __nv_bfloat16* p2, const float (&v)[8]) {
  uint2 h0, l0, h1, l1;
  split_bf16x4(v[0], v[1], v[2], v[3], h0, l0);
  split_bf16x4(v[4], v[5], v[6], v[7], h1, l1);
  *reinterpret_cast<uint4*>(p1) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(p2) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// operand buffers <- the state after the intro: own-h part of every layer's slot 0
__global__ void gen_prep_kernel(GParams p) {
  for (int l = 0; l < p.L; ++l) {
    const Layer& ly = p.l[l];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.B * ly.R / 4; idx += gridDim.x * blockDim.x) {
      const int b = idx / (ly.R / 4), k = (idx - b * (ly.R / 4)) * 4;
      const float4 x = *reinterpret_cast<const float4*>(ly.h + (size_t)b * ly.R + k);
      uint2 hi, lo;
      split_bf16x4(x.x, x.y, x.z, x.w, hi, lo);
      const size_t o = (size_t)b * ly.Kp + ly.inp + k;
      *reinterpret_cast<uint2*>(ly.a1 + o) = hi;
      *reinterpret_cast<uint2*>(ly.a2 + o) = lo;
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(kThreads, 1)
gen_fused_kernel(const __grid_constant__ CUtensorMap map_a0h, const __grid_constant__ CUtensorMap map_a0l,
                 const __grid_constant__ CUtensorMap map_a1h, const __grid_constant__ CUtensorMap map_a1l,
                 const __grid_constant__ CUtensorMap map_dh, const __grid_constant__ CUtensorMap map_dl, const GParams p) {
  constexpr int H = NCH * 128;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B, M = p.M, D = p.D;

  // ---- which role
  int role, j = blockIdx.x;          // role 0 / 1: LSTM layer, 2: Dense, 3: sampler
  if (j < p.nL[0]) role = 0;
  else if ((j -= p.nL[0]) < p.nL[1]) role = 1;
  else if ((j -= p.nL[1]) < p.nD) role = 2;
  else { j -= p.nD; role = 3; }
  unsigned int* f_x = p.flags;
  unsigned int* f_h[2] = {p.flags + 1, p.flags + 2};
  unsigned int* f_fc = p.flags + 3;
  const int top = p.L - 1;

  if (role == 3) {
    // ================================================================ samplers
    float* wdec_s = reinterpret_cast<float*>(smem);
    float* wenc_s = wdec_s + (size_t)D * H;
    const int m = j % M, grp = j / M, ngrp = p.nS / M;
    {
      const float4* gd = reinterpret_cast<const float4*>(p.w_dec + (size_t)m * D * H);
      const float4* ge = reinterpret_cast<const float4*>(p.w_enc + (size_t)m * D * H);
      float4* sd = reinterpret_cast<float4*>(wdec_s);
      float4* se = reinterpret_cast<float4*>(wenc_s);
      for (int i = threadIdx.x; i < D * H / 4; i += kThreads) { sd[i] = __ldg(gd + i); se[i] = __ldg(ge + i); }
    }
    __syncthreads();
    const int enc_col = m * H, dec_col = M * H + m * D;
    const bool threshold = (p.u == nullptr) && !p.use_philox;
    for (int s = 0; s < p.S; ++s) {
      if (s > 0) {
        if (threadIdx.x == 0) wait_flag(f_fc, (unsigned int)s * (unsigned int)p.nD);
        __syncthreads();
      }
      __nv_bfloat16* xin = p.l[0].a1 + (size_t)(s & 1) * 128 * p.l[0].Kp;
      for (int row = grp * kSampWarps + warp; row < B; row += ngrp * kSampWarps) {
        float4 a[NCH], h[NCH];
        const float* be = p.fc + (size_t)row * p.ldfc + enc_col;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          a[c] = __ldcg(reinterpret_cast<const float4*>(be + c * 128 + lane * 4));   // written by other SMs: L2, not L1
          h[c] = make_float4(sigmoid_acc(a[c].x), sigmoid_acc(a[c].y), sigmoid_acc(a[c].z), sigmoid_acc(a[c].w));
        }
        float bdv[4], uv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          bdv[q] = i < D ? __ldcg(p.fc + (size_t)row * p.ldfc + dec_col + i) : 0.f;
          float uu = 0.f;
          if (i < D) {
            if (p.u) uu = __ldg(p.u + (((size_t)s * M + m) * B + row) * D + i);
            else if (p.use_philox) {
              const unsigned long long idx = (global_row(p.rmap, (unsigned long long)row) * M + m) * D + i;
              const unsigned long long off = p.offset0 + (unsigned long long)s;
              const uint4 r4 = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)off, (uint32_t)(off >> 32)),
                                             make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
              uu = u01(r4.x);
            }
          }
          uv[q] = uu;
        }
        float outv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int iend = min(32, D - q * 32);
          for (int ii = 0; ii < iend; ++ii) {
            const int i = q * 32 + ii;
            const float* wd = wdec_s + (size_t)i * H;
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              const float4 w = *reinterpret_cast<const float4*>(wd + c * 128 + lane * 4);
              part = fmaf(h[c].x, w.x, part); part = fmaf(h[c].y, w.y, part);
              part = fmaf(h[c].z, w.z, part); part = fmaf(h[c].w, w.w, part);
            }
            const float dot = warp_sum(part);
            const float bd = __shfl_sync(0xffffffffu, bdv[q], ii);
            const float uu = __shfl_sync(0xffffffffu, uv[q], ii);
            const float pr = sigmoid_acc(bd + dot);
            const bool v = threshold ? (pr >= 0.5f) : (uu < pr);
            if (lane == ii) outv[q] = v ? 1.f : 0.f;
            if (v) {
              const float* we = wenc_s + (size_t)i * H;
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(we + c * 128 + lane * 4);
                a[c].x += w.x; a[c].y += w.y; a[c].z += w.z; a[c].w += w.w;
                h[c] = make_float4(sigmoid_acc(a[c].x), sigmoid_acc(a[c].y), sigmoid_acc(a[c].z), sigmoid_acc(a[c].w));
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          if (i < D) {
            p.out[(size_t)row * p.out_ld + (size_t)s * p.out_step + (size_t)i * M + m] = outv[q];
            xin[(size_t)row * p.l[0].Kp + i * M + m] = __float2bfloat16(outv[q]);     // {0,1}: exact in bf16
          }
        }
      }
      __syncthreads();                 // every warp's stores before this CTA's release
      if (threadIdx.x == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(f_x), "r"(1u) : "memory");
    }
    return;
  }

  // ================================================================ GEMM roles: LSTM layer (32 gate columns) or Dense (64 columns)
  const bool lstm = role < 2;
  const Layer& ly = p.l[lstm ? role : top];
  const int BN = lstm ? 32 : 64;
  const int B_TILE = BN * 64;
  const int Kp = lstm ? ly.Kp : ((ly.R + 31) / 32) * 32;     // Dense: K = R_top (multiple of 8; padded to 32 by the buffer)
  const int KB = Kp / 32;
  const int STAGES = p.stages[role];
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages, bar_tempty = bar_tfull + 8;
  const uint32_t off_b2 = (uint32_t)KB * B_TILE, off_a = 2u * KB * B_TILE;
  const uint32_t tmem_cols = lstm ? 64u : 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident B operand, split into bf16 w1 + w2. LSTM: element (n = g 8 + ul, k) = kern[k][g R + 8 j + ul];
  // Dense: (n, k) = dk[k][64 j + n]; rows k >= K and columns >= C are zero
  {
    for (int idx = threadIdx.x; idx < BN * (Kp / 4); idx += kThreads) {
      const int n = idx % BN, k0 = (idx / BN) * 4;
      float w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        float v = 0.f;
        if (lstm) {
          // operand column k -> kernel row: the input part, a zero gap up to inp, then the recurrent part
          const int row = k < ly.in ? k : ((k >= ly.inp && k < ly.inp + ly.R) ? ly.in + (k - ly.inp) : -1);
          if (row >= 0) v = __ldg(ly.kern + (size_t)row * 4 * ly.R + (n >> 3) * ly.R + 8 * j + (n & 7));
        } else if (k < ly.R && 64 * j + n < p.C) {
          v = __ldg(p.dk + (size_t)k * p.C + 64 * j + n);
        }
        w[e] = v;
      }
      uint2 hi, lo;
      split_bf16x4(w[0], w[1], w[2], w[3], hi, lo);
      const uint32_t off = (uint32_t)((k0 >> 5) * B_TILE) + sw64_off(n, k0 & 31);
      *reinterpret_cast<uint2*>(smem + off) = hi;
      *reinterpret_cast<uint2*>(smem + off_b2 + off) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const CUtensorMap* mh = role == 0 ? &map_a0h : (role == 1 ? &map_a1h : &map_dh);
  const CUtensorMap* ml = role == 0 ? &map_a0l : (role == 1 ? &map_a1l : &map_dl);
  // what this role waits for before step s, and what it bumps after it
  //   L0: x_s (all samplers, s + 1 rounds) and h0_{s-1} (all L0 CTAs, s rounds)
  //   L1: h0_s (s + 1 rounds of L0) and h1_{s-1};   Dense: h_top of step s
  unsigned int* f_out = lstm ? f_h[role] : f_fc;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.S; ++s) {
        if (role == 0) {
          wait_flag(f_x, (unsigned int)(s + 1) * (unsigned int)p.nS);
          wait_flag(f_h[0], (unsigned int)s * (unsigned int)p.nL[0]);
        } else if (role == 1) {
          wait_flag(f_h[0], (unsigned int)(s + 1) * (unsigned int)p.nL[0]);
          wait_flag(f_h[1], (unsigned int)s * (unsigned int)p.nL[1]);
        } else {
          wait_flag(f_h[top], (unsigned int)(s + 1) * (unsigned int)p.nL[top]);
        }
        asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic-proxy stores -> our TMA reads
        const int row0 = (s & 1) * 128;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, A_STAGE);
          const uint32_t dst = smem0 + off_a + stage * A_STAGE;
          tma_load_2d(dst, mh, full, kb * 32, row0);
          tma_load_2d(dst + A_STAGE / 2, ml, full, kb * 32, row0);
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
      for (int s = 0; s < p.S; ++s) {
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
          for (int jj = 0; jj < 2; ++jj) {
            const uint64_t da1 = smem_desc(a1 + jj * 32, 16, 512, 4), da2 = smem_desc(a2 + jj * 32, 16, 512, 4);
            const uint64_t db1 = smem_desc(b1 + jj * 32, 16, 512, 4), db2 = smem_desc(b2 + jj * 32, 16, 512, 4);
            const uint32_t first = (kb > 0 || jj > 0) ? 1u : 0u;
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
    const int q = warp - 4, r = q * 32 + lane;
    const bool row_ok = r < B;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    if (lstm) {
      // ------------------------------------------------------------ cell epilogue: thread = (row, 8 units), c in registers
      const int R = ly.R, unit = 8 * j;
      float bias[4][8], cst[8];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) bias[g][i] = __ldg(ly.bias + g * R + unit + i);
#pragma unroll
      for (int i = 0; i < 8; ++i) cst[i] = row_ok ? ly.c[(size_t)r * R + unit + i] : 0.f;
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hv[i] = 0.f;
      // where h_s goes: own buffer's next slot (own-h part) and the consumer's current slot (its input part / Dense operand)
      const bool last_layer = role == top;
      for (int s = 0; s < p.S; ++s) {
        mbar_wait(bar_tfull, (uint32_t)s & 1u);
        tc_fence_after();
        float pre[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float m8[8], x8[8];
          tmem_ld8(tacc + (uint32_t)(g * 8), m8);
          tmem_ld8(tacc + (uint32_t)(32 + g * 8), x8);
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[g][i] = m8[i] + x8[i] + bias[g][i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float gi = sigmoid_acc(pre[0][i]), gj = tanh_acc(pre[1][i]);
          const float gf = sigmoid_acc(pre[2][i]), go = sigmoid_acc(pre[3][i]);
          cst[i] = gj * gi + cst[i] * gf;
          hv[i] = tanh_acc(cst[i]) * go;
        }
        if (row_ok) {
          const size_t own = ((size_t)((s + 1) & 1) * 128 + r) * ly.Kp + ly.inp + unit;
          store_split8(ly.a1 + own, ly.a2 + own, hv);
          if (last_layer) {
            const int Rp = ((R + 31) / 32) * 32;
            const size_t o = ((size_t)(s & 1) * 128 + r) * Rp + unit;
            store_split8(p.d1 + o, p.d2 + o, hv);
          } else {
            const Layer& nx = p.l[role + 1];
            const size_t o = ((size_t)(s & 1) * 128 + r) * nx.Kp + unit;
            store_split8(nx.a1 + o, nx.a2 + o, hv);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(f_out), "r"(1u) : "memory");
      }
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ly.c[(size_t)r * R + unit + i] = cst[i];
          ly.h[(size_t)r * R + unit + i] = hv[i];
        }
      }
    } else {
      // ------------------------------------------------------------ Dense epilogue: thread = row, 64 columns
      const int col0 = 64 * j;
      for (int s = 0; s < p.S; ++s) {
        mbar_wait(bar_tfull, (uint32_t)s & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32], vx[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), v);
          tmem_ld32(tacc + (uint32_t)(64 + c * 32), vx);
          if (row_ok) {
            float* dst = p.fc + (size_t)r * p.ldfc + col0 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = col0 + c * 32 + i;
              if (col < p.C) dst[i] = v[i] + vx[i] + __ldg(p.db + col);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(f_out), "r"(1u) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static inline int pad32(int x) { return (x + 31) / 32 * 32; }

struct Plan {
  int nL[2], nD, nS, stages[3];
  size_t smem, ws_bytes, off_a[2], off_d, off_flags;
  int Kp[2], Rp;
};

static bool make_plan(int L, int I, const int* R, int B, int M, int D, int H, Plan* pl) {
  if (L < 1 || L > 2 || B < 1 || B > 128 || D > 128 || (H != 128 && H != 256)) return false;
  const int sms = mnn_tc_num_sms();
  size_t smem = (size_t)2 * D * H * sizeof(float);          // samplers: both weight matrices of a track
  size_t off = 0;
  int in = I;
  for (int l = 0; l < 2; ++l) {
    pl->nL[l] = 0;
    pl->Kp[l] = 0;
    if (l >= L) continue;
    if (R[l] % 8) return false;
    pl->nL[l] = R[l] / 8;
    pl->Kp[l] = pad32(pad32(in) + R[l]);
    const size_t resident = (size_t)2 * (pl->Kp[l] / 32) * (32 * 64);
    long long room = (long long)224 * 1024 - (long long)resident;
    int st = (int)(room / A_STAGE);
    if (st < 2) return false;
    if (st > kMaxStages) st = kMaxStages;
    if (st > pl->Kp[l] / 32) st = pl->Kp[l] / 32;
    pl->stages[l] = st;
    if (resident + (size_t)st * A_STAGE > smem) smem = resident + (size_t)st * A_STAGE;
    pl->off_a[l] = off;
    off += (size_t)2 * 2 * 128 * pl->Kp[l] * 2;             // hi + lo, 2 slots
    in = R[l];
  }
  const int C = M * (H + D);
  pl->Rp = pad32(R[L - 1]);
  pl->nD = (C + 63) / 64;
  {
    const size_t resident = (size_t)2 * (pl->Rp / 32) * (64 * 64);
    int st = (int)(((long long)224 * 1024 - (long long)resident) / A_STAGE);
    if (st > kMaxStages) st = kMaxStages;
    if (st > pl->Rp / 32) st = pl->Rp / 32;
    pl->stages[2] = st;
    if (resident + (size_t)st * A_STAGE > smem) smem = resident + (size_t)st * A_STAGE;
  }
  pl->off_d = off;
  off += (size_t)2 * 2 * 128 * pl->Rp * 2;
  pl->off_flags = off;
  off += 64;
  pl->ws_bytes = off + 256;
  pl->smem = smem + 1024;
  if (pl->smem > 227 * 1024) return false;
  const int fixed = pl->nL[0] + pl->nL[1] + pl->nD;
  if (fixed + M > sms) return false;
  int groups = (B + kSampWarps - 1) / kSampWarps;            // row groups per track that have work
  const int room_groups = (sms - fixed) / M;
  if (groups > room_groups) groups = room_groups;
  pl->nS = groups * M;
  return true;
}

}  // namespace gen
}  // namespace mnn

using namespace mnn;

extern "C" size_t mnn_generate_fused_workspace_bytes(int num_layers, int num_inputs, int r0, int r1, int B, int M, int D, int H) {
  gen::Plan pl;
  const int R[2] = {r0, r1};
  return gen::make_plan(num_layers, num_inputs, R, B, M, D, H, &pl) ? pl.ws_bytes : 0;
}

extern "C" int mnn_generate_fused(int num_layers, int num_inputs, const float* kern0, const float* bias0, float* c0, float* h0,
                                  int r0, const float* kern1, const float* bias1, float* c1, float* h1, int r1,
                                  const float* dense_kernel, const float* dense_bias, const float* w_enc, const float* w_dec,
                                  float* fc, long long ldfc, const float* u, int use_philox, unsigned long long seed,
                                  unsigned long long offset0, float* out, long long out_ld, long long out_step, int B,
                                  int S, int M, int D, int H, void* ws, cudaStream_t stream) {
  MNN_REQUIRE(kern0 && bias0 && c0 && h0 && dense_kernel && dense_bias && w_enc && w_dec && fc && out && ws, MNN_ERR_ARG,
              "generate_fused: null pointer");
  MNN_REQUIRE(num_layers == 1 || (kern1 && bias1 && c1 && h1), MNN_ERR_ARG, "generate_fused: layer 1 pointers missing");
  MNN_REQUIRE(S > 0 && num_inputs == M * D, MNN_ERR_ARG, "generate_fused: the sampled frame (M*D values) is the layer-0 input");
  gen::Plan pl;
  const int R[2] = {r0, r1};
  MNN_REQUIRE(gen::make_plan(num_layers, num_inputs, R, B, M, D, H, &pl), MNN_ERR_UNSUPPORTED,
              "generate_fused: needs B <= 128, 1-2 layers with num_units % 8 == 0, H in {128, 256}, D <= 128, and the role "
              "groups within the SM count");
  gen::GParams p{};
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  p.L = num_layers; p.B = B; p.S = S; p.M = M; p.D = D; p.H = H; p.C = M * (H + D);
  const float* kerns[2] = {kern0, kern1}; const float* biases[2] = {bias0, bias1};
  float* cs[2] = {c0, c1}; float* hs[2] = {h0, h1};
  int in = num_inputs;
  for (int l = 0; l < num_layers; ++l) {
    gen::Layer& ly = p.l[l];
    ly.kern = kerns[l]; ly.bias = biases[l]; ly.c = cs[l]; ly.h = hs[l];
    ly.in = in; ly.inp = gen::pad32(in); ly.R = R[l]; ly.Kp = pl.Kp[l];
    ly.a1 = reinterpret_cast<__nv_bfloat16*>(w8 + pl.off_a[l]);
    ly.a2 = ly.a1 + (size_t)2 * 128 * ly.Kp;
    in = R[l];
  }
  p.dk = dense_kernel; p.db = dense_bias;
  p.d1 = reinterpret_cast<__nv_bfloat16*>(w8 + pl.off_d);
  p.d2 = p.d1 + (size_t)2 * 128 * pl.Rp;
  p.fc = fc; p.ldfc = ldfc; p.w_enc = w_enc; p.w_dec = w_dec; p.u = u; p.use_philox = use_philox; p.seed = seed;
  p.offset0 = offset0; p.rmap = current_row_map(); p.out = out; p.out_ld = out_ld; p.out_step = out_step;
  p.flags = reinterpret_cast<unsigned int*>(w8 + pl.off_flags);
  p.nS = pl.nS; p.nL[0] = pl.nL[0]; p.nL[1] = pl.nL[1]; p.nD = pl.nD;
  p.stages[0] = pl.stages[0]; p.stages[1] = pl.stages[1]; p.stages[2] = pl.stages[2];
  cudaMemsetAsync(ws, 0, pl.ws_bytes, stream);               // zero pads, zero lo halves of the binary inputs, zero flags
  gen::gen_prep_kernel<<<32, 256, 0, stream>>>(p);
  int rc = mnn_check_launch("generate_fused prep");
  if (rc) return rc;
  CUtensorMap m[6];
  for (int l = 0; l < 2; ++l) {
    const gen::Layer& ly = p.l[l < num_layers ? l : 0];
    if ((rc = mnn_tc_make_map_bf16(ly.a1, ly.Kp, ly.Kp, 256, 128, &m[2 * l]))) return rc;
    if ((rc = mnn_tc_make_map_bf16(ly.a2, ly.Kp, ly.Kp, 256, 128, &m[2 * l + 1]))) return rc;
  }
  if ((rc = mnn_tc_make_map_bf16(p.d1, pl.Rp, pl.Rp, 256, 128, &m[4]))) return rc;
  if ((rc = mnn_tc_make_map_bf16(p.d2, pl.Rp, pl.Rp, 256, 128, &m[5]))) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl.nL[0] + pl.nL[1] + pl.nD + pl.nS);
  cfg.blockDim = dim3(gen::kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e;
  if (H == 256) {
    cudaFuncSetAttribute(gen::gen_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    e = cudaLaunchKernelEx(&cfg, gen::gen_fused_kernel<2>, m[0], m[1], m[2], m[3], m[4], m[5], p);
  } else {
    cudaFuncSetAttribute(gen::gen_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    e = cudaLaunchKernelEx(&cfg, gen::gen_fused_kernel<1>, m[0], m[1], m[2], m[3], m[4], m[5], p);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    mnn_set_error(cudaGetErrorString(e));
    return (int)e;
  }
  return mnn_check_launch("generate_fused");
}
