// NADE / MultiNADE kernels (K4 forward log-prob, K5 backward, K6 ancestral sampling).
// Arithmetic follows reference multinn/models/common/nade.py:155-329 (log_prob, sample, _cond_prob)
// and multinn/utils/auxiliary.py:9-11 (safe_log). Track batching follows
// multinn/models/generators/rnn_multinade.py:231-317.
//
// The recursion a_{i+1} = a_i + v_i * w_enc[i] only changes `a` where the target bit v_i is 1, so
// h_i = sigmoid(a_i) is piecewise constant over i ("segments"); sigmoids are evaluated once per
// segment, the decode dots l_i = b_dec_i + h_i . w_dec[i] for every i. Results are identical to
// evaluating every i; only the transcendental count drops from D*H to (1 + popcount(v)) * H per row.
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

constexpr int kNW = 4;          // 32-bit mask words per (row, track): D <= 128
constexpr int kFwdThreads = 384;
constexpr int kFwdRows = 4;     // rows per warp

struct NadeArgs {
  const uint32_t* bits;  // [M][N][kNW]
  const float* fc;       // [N][ld]   b_enc(m) at col enc_col0 + m*H, b_dec(m) at dec_col0 + m*D
  long long ld;
  int enc_col0, dec_col0;
  const float* w_enc;    // [M][D][H]
  const float* w_dec;    // [M][D][H]
  float* nll;            // [M][N]
  float* cond_p;         // [M][N][D] or null
  float* dfc;            // [N][ld] or null: d b_dec columns written by fwd, d b_enc columns by bwd
  float* dw_enc;         // [M][D][H] accumulated (bwd)
  float* dw_dec;         // [M][D][H] accumulated (bwd)
  float gscale;          // d loss / d nll[n,m]
  int N, M, D;
};

__device__ __forceinline__ uint32_t pick_word(const uint32_t (&w)[kNW], int idx) {
  return idx == 0 ? w[0] : (idx == 1 ? w[1] : (idx == 2 ? w[2] : w[3]));
}

// ------------------------------------------------------------------------------------------------
// K4 forward. One warp = kFwdRows rows of one track; lane owns H/32 hidden units (float4 chunks at
// k = c*128 + lane*4). Eight output dims at a time: 32 partial dots per lane are reduce-scattered
// over the warp so lane L ends with l(row L>>3, dim 8c + (L&7)) and does that element's sigmoid/BCE.
template <int NCH>
__global__ void __launch_bounds__(kFwdThreads, 1) nade_fwd_kernel(NadeArgs p) {
  constexpr int H = NCH * 128;
  extern __shared__ __align__(16) float smem[];
  const int D = p.D;
  float* wdec_s = smem;                 // [D][H]
  float* wenc_s = smem + (size_t)D * H; // [D][H]

  const int m = blockIdx.x % p.M;
  const int cta = blockIdx.x / p.M;
  const int nctas = (gridDim.x - m + p.M - 1) / p.M;
  {
    const float4* gd = reinterpret_cast<const float4*>(p.w_dec + (size_t)m * D * H);
    const float4* ge = reinterpret_cast<const float4*>(p.w_enc + (size_t)m * D * H);
    float4* sd = reinterpret_cast<float4*>(wdec_s);
    float4* se = reinterpret_cast<float4*>(wenc_s);
    for (int i = threadIdx.x; i < D * H / 4; i += blockDim.x) {
      sd[i] = __ldg(gd + i);
      se[i] = __ldg(ge + i);
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ngroups = (p.N + kFwdRows - 1) / kFwdRows;
  const uint32_t* bits = p.bits + (size_t)m * p.N * kNW;
  const int enc_col = p.enc_col0 + m * H, dec_col = p.dec_col0 + m * D;
  const int nchunks = (D + 7) / 8;
  const int er = lane >> 3, eii = lane & 7;  // epilogue ownership

  for (int g = cta * nwarps + warp; g < ngroups; g += nctas * nwarps) {
    const int row0 = g * kFwdRows;
    uint32_t mk[kFwdRows][kNW], sm[kFwdRows][kNW];  // target bits; sm = mk << 1 (bit i = v_{i-1})
    float4 a[kFwdRows][NCH], h[kFwdRows][NCH];
#pragma unroll
    for (int r = 0; r < kFwdRows; ++r) {
      const int row = min(row0 + r, p.N - 1);
      const uint4 mm = __ldg(reinterpret_cast<const uint4*>(bits + (size_t)row * kNW));
      mk[r][0] = mm.x; mk[r][1] = mm.y; mk[r][2] = mm.z; mk[r][3] = mm.w;
      sm[r][0] = mm.x << 1;
      sm[r][1] = __funnelshift_l(mm.x, mm.y, 1);
      sm[r][2] = __funnelshift_l(mm.y, mm.z, 1);
      sm[r][3] = __funnelshift_l(mm.z, mm.w, 1);
      const float* be = p.fc + (size_t)row * p.ld + enc_col;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        a[r][c] = __ldg(reinterpret_cast<const float4*>(be + c * 128 + lane * 4));
        h[r][c] = make_float4(sigmoid_fast(a[r][c].x), sigmoid_fast(a[r][c].y),
                              sigmoid_fast(a[r][c].z), sigmoid_fast(a[r][c].w));
      }
    }
    const int erow = row0 + er;
    const bool erow_ok = erow < p.N;
    const size_t erow_c = (size_t)min(erow, p.N - 1);
    float nll_acc = 0.f;

    for (int c8 = 0; c8 < nchunks; ++c8) {
      const int i0 = c8 * 8;
      // prefetch this lane's decoder bias
      const int ei = i0 + eii;
      const bool e_ok = erow_ok && ei < D;
      float bd = 0.f;
      if (ei < D) bd = __ldg(p.fc + erow_c * p.ld + dec_col + ei);

      float vals[32];
      uint32_t upd[kFwdRows];
#pragma unroll
      for (int r = 0; r < kFwdRows; ++r) upd[r] = (pick_word(sm[r], c8 >> 2) >> ((c8 & 3) * 8)) & 0xffu;

#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = min(i0 + ii, D - 1);  // tail dims recompute dim D-1; discarded in the epilogue
#pragma unroll
        for (int r = 0; r < kFwdRows; ++r) {
          if (upd[r] & (1u << ii)) {  // warp-uniform: v_{i-1} == 1 for this row
            const float* we = wenc_s + (size_t)(i0 + ii - 1) * H;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              const float4 w = *reinterpret_cast<const float4*>(we + c * 128 + lane * 4);
              a[r][c].x += w.x; a[r][c].y += w.y; a[r][c].z += w.z; a[r][c].w += w.w;
              h[r][c] = make_float4(sigmoid_fast(a[r][c].x), sigmoid_fast(a[r][c].y),
                                    sigmoid_fast(a[r][c].z), sigmoid_fast(a[r][c].w));
            }
          }
        }
        float4 w[NCH];
        const float* wd = wdec_s + (size_t)i * H;
#pragma unroll
        for (int c = 0; c < NCH; ++c) w[c] = *reinterpret_cast<const float4*>(wd + c * 128 + lane * 4);
#pragma unroll
        for (int r = 0; r < kFwdRows; ++r) {
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            acc = ffma2(make_float2(h[r][c].x, h[r][c].y), make_float2(w[c].x, w[c].y), acc);
            acc = ffma2(make_float2(h[r][c].z, h[r][c].w), make_float2(w[c].z, w[c].w), acc);
          }
          vals[r * 8 + ii] = acc.x + acc.y;
        }
      }
      // reduce-scatter: after step s, lane bit s selects which half of the remaining values it keeps
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
          const float send = up ? vals[j] : vals[j + s];
          const float keep = up ? vals[j + s] : vals[j];
          vals[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      // epilogue: lane owns (row er, dim ei)
      {
        // select this lane's row mask word without dynamic register indexing
        uint32_t wsel = 0;
#pragma unroll
        for (int r = 0; r < kFwdRows; ++r) {
          const uint32_t wr = pick_word(mk[r], c8 >> 2);
          wsel = (er == r) ? wr : wsel;
        }
        const bool v = (wsel >> (((c8 & 3) * 8) + eii)) & 1u;
        const float l = vals[0] + bd;
        const float pr = sigmoid_acc(l);
        const float q = 1.0f - pr;
        const float lp = v ? logf(kSafeLogEps + pr) : logf(kSafeLogEps + q);
        if (e_ok) {
          nll_acc -= lp;
          if (p.cond_p) p.cond_p[((size_t)m * p.N + erow) * D + ei] = pr;
          if (p.dfc) {
            // d(-lp)/dl with dp/dl = p(1-p)   (safe_log eps kept, nade.py:210)
            const float pq = pr * q;
            const float dl = v ? -pq / (kSafeLogEps + pr) : pq / (kSafeLogEps + q);
            p.dfc[(size_t)erow * p.ld + dec_col + ei] = p.gscale * dl;
          }
        }
      }
    }
    nll_acc += __shfl_xor_sync(0xffffffffu, nll_acc, 1);
    nll_acc += __shfl_xor_sync(0xffffffffu, nll_acc, 2);
    nll_acc += __shfl_xor_sync(0xffffffffu, nll_acc, 4);
    if (eii == 0 && erow_ok) p.nll[(size_t)m * p.N + erow] = nll_acc;
  }
}

// ------------------------------------------------------------------------------------------------
// K5 backward. Thread k owns hidden unit k: its w_dec column and its dW_dec column accumulators live in registers
// (the i loop is fully unrolled), its dW_enc column accumulators in shared memory, so no reduction over k is ever
// needed. Walks i = D-1..0 and reverses the prefix (a -= w_enc[i-1]) at each set target bit. Requires the forward
// kernel to have written dl into the d b_dec columns of dfc.
// Rows are staged in batches of kBwdRows through a cp.async double buffer (b_enc row, dl row, target mask), so the
// global-load latency of batch b+1 hides behind the arithmetic of batch b; two rows are walked at a time to give
// every thread two independent dependency chains.
constexpr int kBwdRows = 8;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}

template <int H, int D>
__global__ void __launch_bounds__(H, 1) nade_bwd_kernel(NadeArgs p) {
  static_assert(D % 4 == 0, "D must be a multiple of 4");
  extern __shared__ __align__(16) float smem[];
  float* wenc_s = smem;                            // [D][H]
  float* acce_s = smem + (size_t)D * H;            // [D][H]
  float* a_s = acce_s + (size_t)D * H;             // [2][kBwdRows][H]
  float* dl_s = a_s + 2 * kBwdRows * H;            // [2][kBwdRows][D]
  uint32_t* mk_s = reinterpret_cast<uint32_t*>(dl_s + 2 * kBwdRows * D);  // [2][kBwdRows][kNW]

  const int m = blockIdx.x % p.M;
  const int cta = blockIdx.x / p.M;
  const int nctas = (gridDim.x - m + p.M - 1) / p.M;
  const int k = threadIdx.x;
  const float* gwe = p.w_enc + (size_t)m * D * H;
  const float* gwd = p.w_dec + (size_t)m * D * H;

  float wdec[D], accw[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    wdec[i] = __ldg(gwd + (size_t)i * H + k);
    accw[i] = 0.f;
    wenc_s[i * H + k] = __ldg(gwe + (size_t)i * H + k);
    acce_s[i * H + k] = 0.f;
  }
  const uint32_t* bits = p.bits + (size_t)m * p.N * kNW;
  const int enc_col = p.enc_col0 + m * H, dec_col = p.dec_col0 + m * D;
  const int nbatches = (p.N + kBwdRows - 1) / kBwdRows;

  auto prefetch = [&](int batch, int buf) {
    const int row0 = batch * kBwdRows;
    float* ab = a_s + (size_t)buf * kBwdRows * H;
    float* db = dl_s + (size_t)buf * kBwdRows * D;
    uint32_t* mb = mk_s + buf * kBwdRows * kNW;
    for (int c = k; c < kBwdRows * H / 4; c += H) {          // b_enc rows, 16 B chunks
      const int r = c / (H / 4), off = (c % (H / 4)) * 4, row = row0 + r;
      if (row < p.N) cp_async16(ab + r * H + off, p.fc + (size_t)row * p.ld + enc_col + off);
      else *reinterpret_cast<float4*>(ab + r * H + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c = k; c < kBwdRows * D; c += H) {              // dl rows
      const int r = c / D, i = c - r * D, row = row0 + r;
      if (row < p.N) cp_async4(db + c, p.dfc + (size_t)row * p.ld + dec_col + i);
      else db[c] = 0.f;
    }
    if (k < kBwdRows) {
      const int row = row0 + k;
      if (row < p.N) cp_async16(mb + k * kNW, bits + (size_t)row * kNW);
      else *reinterpret_cast<uint4*>(mb + k * kNW) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (cta < nbatches) prefetch(cta, 0);
  __syncthreads();   // weights staged
  for (int b = cta; b < nbatches; b += nctas) {
    const int nb = b + nctas;
    if (nb < nbatches) {
      prefetch(nb, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int row0 = b * kBwdRows;
    const float* ab = a_s + (size_t)buf * kBwdRows * H;
    const float* db = dl_s + (size_t)buf * kBwdRows * D;
    const uint32_t* mb = mk_s + buf * kBwdRows * kNW;
#pragma unroll 1
    for (int r = 0; r < kBwdRows; r += 2) {
      if (row0 + r >= p.N) break;
      uint32_t mk0[kNW], mk1[kNW];
#pragma unroll
      for (int w = 0; w < kNW; ++w) { mk0[w] = mb[r * kNW + w]; mk1[w] = mb[(r + 1) * kNW + w]; }
      float a0 = ab[r * H + k], a1 = ab[(r + 1) * H + k];
      // forward prefix over set bits j < D-1 (bit D-1 opens no segment that any dim reads)
#pragma unroll
      for (int w = 0; w < kNW; ++w) {
        uint32_t x0 = mk0[w], x1 = mk1[w];
        if (w * 32 + 32 > D - 1) {
          const int keep = (D - 1) - w * 32;  // number of low bits to keep (may be <= 0)
          const uint32_t msk = keep <= 0 ? 0u : (keep >= 32 ? 0xffffffffu : ((1u << keep) - 1u));
          x0 &= msk; x1 &= msk;
        }
        while (x0) { const int j = __ffs(x0) - 1; x0 &= x0 - 1; a0 += wenc_s[(w * 32 + j) * H + k]; }
        while (x1) { const int j = __ffs(x1) - 1; x1 &= x1 - 1; a1 += wenc_s[(w * 32 + j) * H + k]; }
      }
      float h0 = sigmoid_fast(a0), h1 = sigmoid_fast(a1);
      float ga0 = 0.f, ga1 = 0.f, dh0 = 0.f, dh1 = 0.f;
      const float* dl0 = db + r * D;
      const float* dl1 = db + (r + 1) * D;
      // Segment boundaries (a set target bit j ends the segment that dims > j read) are rare, so the walk over dims is
      // a jump-table entry into ONE unrolled run of FMAs (Duff's device): the hot code stays a few KB instead of one
      // boundary block per dim, which overflowed the instruction cache (ncu: stall_no_instruction dominated).
      uint32_t rm[kNW];   // boundaries still ahead: bit j set <=> stop after dim j + 1
#pragma unroll
      for (int w = 0; w < kNW; ++w) {
        rm[w] = mk0[w] | mk1[w];
        if (w * 32 + 32 > D - 1) {
          const int keep = (D - 1) - w * 32;
          rm[w] &= keep <= 0 ? 0u : (keep >= 32 ? 0xffffffffu : ((1u << keep) - 1u));
        }
      }
      int i = D - 1;
#pragma unroll 1
      for (;;) {
        int s = 0;   // stop position: process dims i..s, then handle the boundary in front of dim s (if s > 0)
#pragma unroll
        for (int w = kNW - 1; w >= 0; --w)
          if (s == 0 && rm[w]) {
            const int bpos = 31 - __clz(rm[w]);
            rm[w] &= ~(1u << bpos);
            s = w * 32 + bpos + 1;
          }
        float4 q0 = *reinterpret_cast<const float4*>(dl0 + (i & ~3));
        float4 q1 = *reinterpret_cast<const float4*>(dl1 + (i & ~3));
#define MNN_BWD_STEP(I)                                                                                   \
  case (I):                                                                                               \
    if constexpr ((I) < D) {                                                                              \
      if (((I) & 3) == 3) {                                                                               \
        q0 = *reinterpret_cast<const float4*>(dl0 + ((I) & ~3));                                          \
        q1 = *reinterpret_cast<const float4*>(dl1 + ((I) & ~3));                                          \
      }                                                                                                   \
      const float d0 = ((I) & 3) == 3 ? q0.w : (((I) & 3) == 2 ? q0.z : (((I) & 3) == 1 ? q0.y : q0.x));  \
      const float d1 = ((I) & 3) == 3 ? q1.w : (((I) & 3) == 2 ? q1.z : (((I) & 3) == 1 ? q1.y : q1.x));  \
      dh0 = fmaf(d0, wdec[(I) < D ? (I) : 0], dh0);                                                       \
      dh1 = fmaf(d1, wdec[(I) < D ? (I) : 0], dh1);                                                       \
      accw[(I) < D ? (I) : 0] = fmaf(d0, h0, fmaf(d1, h1, accw[(I) < D ? (I) : 0]));                      \
      if ((I) == s) break;                                                                                \
    }
#define MNN_BWD_STEP4(I) MNN_BWD_STEP((I) + 3) MNN_BWD_STEP((I) + 2) MNN_BWD_STEP((I) + 1) MNN_BWD_STEP(I)
#define MNN_BWD_STEP16(I) MNN_BWD_STEP4((I) + 12) MNN_BWD_STEP4((I) + 8) MNN_BWD_STEP4((I) + 4) MNN_BWD_STEP4(I)
        switch (i) {
          MNN_BWD_STEP16(112) MNN_BWD_STEP16(96) MNN_BWD_STEP16(80) MNN_BWD_STEP16(64)
          MNN_BWD_STEP16(48) MNN_BWD_STEP16(32) MNN_BWD_STEP16(16) MNN_BWD_STEP16(0)
          default: break;
        }
#undef MNN_BWD_STEP16
#undef MNN_BWD_STEP4
#undef MNN_BWD_STEP
        if (s == 0) break;
        if (pick_word(mk0, (s - 1) >> 5) & (1u << ((s - 1) & 31))) {  // block-uniform
          ga0 += dh0 * h0 * (1.f - h0);
          dh0 = 0.f;
          acce_s[(s - 1) * H + k] += ga0;
          a0 -= wenc_s[(s - 1) * H + k];
          h0 = sigmoid_fast(a0);
        }
        if (pick_word(mk1, (s - 1) >> 5) & (1u << ((s - 1) & 31))) {
          ga1 += dh1 * h1 * (1.f - h1);
          dh1 = 0.f;
          acce_s[(s - 1) * H + k] += ga1;
          a1 -= wenc_s[(s - 1) * H + k];
          h1 = sigmoid_fast(a1);
        }
        i = s - 1;
      }
      ga0 += dh0 * h0 * (1.f - h0);
      ga1 += dh1 * h1 * (1.f - h1);
      p.dfc[(size_t)(row0 + r) * p.ld + enc_col + k] = ga0;
      if (row0 + r + 1 < p.N) p.dfc[(size_t)(row0 + r + 1) * p.ld + enc_col + k] = ga1;
    }
    __syncthreads();   // everyone is done with `buf` before it is refilled two iterations from now
    buf ^= 1;
  }
  float* gdd = p.dw_dec + (size_t)m * D * H;
  float* gde = p.dw_enc + (size_t)m * D * H;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    atomicAdd(gdd + (size_t)i * H + k, accw[i]);
    atomicAdd(gde + (size_t)i * H + k, acce_s[i * H + k]);
  }
}

// ------------------------------------------------------------------------------------------------
// K6 ancestral sampling (nade.py:231-308). One warp = one (row, track); every dim needs the full
// reduction before the next can start. v_i = (u_i < p_i) strict (TFP 0.6.0 Bernoulli), or p_i >= 0.5
// when u is null (temperature=None).
struct NadeSampleArgs {
  const float* fc; long long ld; int enc_col0, dec_col0;
  const float* w_enc; const float* w_dec;  // [M][D][H]
  const float* u;                          // [M][N][D] uniforms or null
  float* out; long long out_ld; int out_dim_stride, out_track_stride;  // out[row*out_ld + i*ds + m*ts]
  float* nll;                              // [M][N] or null
  int N, M, D;
  unsigned long long seed, offset; int use_philox;
};

template <int NCH>
__global__ void __launch_bounds__(256, 1) nade_sample_kernel(NadeSampleArgs p) {
  constexpr int H = NCH * 128;
  extern __shared__ __align__(16) float smem[];
  const int D = p.D;
  float* wdec_s = smem;
  float* wenc_s = smem + (size_t)D * H;
  const int m = blockIdx.x % p.M;
  const int cta = blockIdx.x / p.M;
  const int nctas = (gridDim.x - m + p.M - 1) / p.M;
  {
    const float4* gd = reinterpret_cast<const float4*>(p.w_dec + (size_t)m * D * H);
    const float4* ge = reinterpret_cast<const float4*>(p.w_enc + (size_t)m * D * H);
    float4* sd = reinterpret_cast<float4*>(wdec_s);
    float4* se = reinterpret_cast<float4*>(wenc_s);
    for (int i = threadIdx.x; i < D * H / 4; i += blockDim.x) { sd[i] = __ldg(gd + i); se[i] = __ldg(ge + i); }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int enc_col = p.enc_col0 + m * H, dec_col = p.dec_col0 + m * D;
  for (int row = cta * nwarps + warp; row < p.N; row += nctas * nwarps) {
    float4 a[NCH], h[NCH];
    const float* be = p.fc + (size_t)row * p.ld + enc_col;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      a[c] = __ldg(reinterpret_cast<const float4*>(be + c * 128 + lane * 4));
      h[c] = make_float4(sigmoid_acc(a[c].x), sigmoid_acc(a[c].y), sigmoid_acc(a[c].z), sigmoid_acc(a[c].w));
    }
    // this lane's share of b_dec and u: lane holds dims lane, lane+32, ...
    float bdv[4], uv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = q * 32 + lane;
      bdv[q] = i < D ? __ldg(p.fc + (size_t)row * p.ld + dec_col + i) : 0.f;
      float uu = 0.f;
      if (i < D) {
        if (p.u) uu = __ldg(p.u + ((size_t)m * p.N + row) * D + i);
        else if (p.use_philox) {
          const unsigned long long idx = ((unsigned long long)m * p.N + row) * D + i;
          const uint4 r4 = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)p.offset,
                                                    (uint32_t)(p.offset >> 32)),
                                         make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          uu = u01(r4.x);
        }
      }
      uv[q] = uu;
    }
    const bool threshold = (p.u == nullptr) && !p.use_philox;
    float nll = 0.f;
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
        nll -= v ? logf(kSafeLogEps + pr) : logf(kSafeLogEps + (1.0f - pr));
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
      if (i < D) p.out[(size_t)row * p.out_ld + (size_t)i * p.out_dim_stride + (size_t)m * p.out_track_stride] = outv[q];
    }
    if (p.nll && lane == 0) p.nll[(size_t)m * p.N + row] = nll;
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace mnn

using namespace mnn;

static int check_nade_common(int N, int M, int D, int H, long long ld, int enc_col0, int dec_col0) {
  MNN_REQUIRE(N > 0 && M > 0 && D > 0 && H > 0, MNN_ERR_ARG, "nade: non-positive size");
  MNN_REQUIRE(D <= 32 * kNW, MNN_ERR_UNSUPPORTED, "nade: num_dims > 128 not instantiated");
  MNN_REQUIRE(H == 128 || H == 256, MNN_ERR_UNSUPPORTED, "nade: num_hidden must be 128 or 256");
  MNN_REQUIRE((size_t)2 * D * H * sizeof(float) <= 200 * 1024, MNN_ERR_UNSUPPORTED, "nade: weights exceed shared memory");
  MNN_REQUIRE(ld % 4 == 0 && enc_col0 % 4 == 0, MNN_ERR_ARG, "nade: fc row stride / b_enc column must be 16-byte aligned");
  (void)dec_col0;
  return MNN_OK;
}

extern "C" int mnn_nade_logprob_fwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0,
                                    const float* w_enc, const float* w_dec, float* nll, float* cond_p, float* dfc,
                                    float gscale, int N, int M, int D, int H, cudaStream_t stream) {
  MNN_REQUIRE(bits && fc && w_enc && w_dec && nll, MNN_ERR_ARG, "nade_logprob_fwd: null pointer");
  int rc = check_nade_common(N, M, D, H, ld, enc_col0, dec_col0);
  if (rc) return rc;
  NadeArgs a{bits, fc, ld, enc_col0, dec_col0, w_enc, w_dec, nll, cond_p, dfc, nullptr, nullptr, gscale, N, M, D};
  const size_t smem = (size_t)2 * D * H * sizeof(float);
  const int groups = (N + kFwdRows - 1) / kFwdRows;
  const int warps_per_cta = kFwdThreads / 32;
  int grid = num_sms();
  const int need = M * ((groups + warps_per_cta - 1) / warps_per_cta);
  if (grid > need) grid = need;
  if (grid < M) grid = M;
  if (H == 256) {
    cudaFuncSetAttribute(nade_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_fwd_kernel<2><<<grid, kFwdThreads, smem, stream>>>(a);
  } else {
    cudaFuncSetAttribute(nade_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_fwd_kernel<1><<<grid, kFwdThreads, smem, stream>>>(a);
  }
  return mnn_check_launch("nade_logprob_fwd");
}

template <int H, int D>
static int launch_bwd(const NadeArgs& a, cudaStream_t stream) {
  const size_t smem = ((size_t)2 * D * H + 2 * kBwdRows * H + 2 * kBwdRows * D) * sizeof(float) +
                      2 * kBwdRows * kNW * sizeof(uint32_t);
  cudaFuncSetAttribute(nade_bwd_kernel<H, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = num_sms();
  const int need = a.M * ((a.N + kBwdRows - 1) / kBwdRows);
  if (grid > need) grid = need;
  if (grid < a.M) grid = a.M;
  nade_bwd_kernel<H, D><<<grid, H, smem, stream>>>(a);
  return mnn_check_launch("nade_logprob_bwd");
}

extern "C" int mnn_nade_logprob_bwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0,
                                    const float* w_enc, const float* w_dec, float* dfc, float* dw_enc, float* dw_dec,
                                    int N, int M, int D, int H, cudaStream_t stream) {
  MNN_REQUIRE(bits && fc && w_enc && w_dec && dfc && dw_enc && dw_dec, MNN_ERR_ARG, "nade_logprob_bwd: null pointer");
  int rc = check_nade_common(N, M, D, H, ld, enc_col0, dec_col0);
  if (rc) return rc;
  NadeArgs a{bits, fc, ld, enc_col0, dec_col0, w_enc, w_dec, nullptr, nullptr, dfc, dw_enc, dw_dec, 0.f, N, M, D};
  if (H == 256 && D == 84) return launch_bwd<256, 84>(a, stream);
  if (H == 128 && D == 84) return launch_bwd<128, 84>(a, stream);
  if (H == 128 && D == 20) return launch_bwd<128, 20>(a, stream);
  mnn_set_error("nade_logprob_bwd: (num_dims, num_hidden) not instantiated; built: (84,256) (84,128) (20,128)");
  return MNN_ERR_UNSUPPORTED;
}

extern "C" int mnn_nade_sample(const float* fc, long long ld, int enc_col0, int dec_col0, const float* w_enc,
                               const float* w_dec, const float* u, int use_philox, unsigned long long seed,
                               unsigned long long offset, float* out, long long out_ld, int out_dim_stride,
                               int out_track_stride, float* nll, int N, int M, int D, int H, cudaStream_t stream) {
  MNN_REQUIRE(fc && w_enc && w_dec && out, MNN_ERR_ARG, "nade_sample: null pointer");
  int rc = check_nade_common(N, M, D, H, ld, enc_col0, dec_col0);
  if (rc) return rc;
  NadeSampleArgs a{fc, ld, enc_col0, dec_col0, w_enc, w_dec, u, out, out_ld, out_dim_stride, out_track_stride,
                   nll, N, M, D, seed, offset, use_philox};
  const size_t smem = (size_t)2 * D * H * sizeof(float);
  int grid = num_sms();
  const int need = M * ((N + 7) / 8);
  if (grid > need) grid = need;
  if (grid < M) grid = M;
  if (H == 256) {
    cudaFuncSetAttribute(nade_sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_sample_kernel<2><<<grid, 256, smem, stream>>>(a);
  } else {
    cudaFuncSetAttribute(nade_sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_sample_kernel<1><<<grid, 256, smem, stream>>>(a);
  }
  return mnn_check_launch("nade_sample");
}
