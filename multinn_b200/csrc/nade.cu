// NADE / MultiNADE kernels (K4 forward log-prob, K5 backward, K6 ancestral sampling).
// Arithmetic follows reference multinn/models/common/nade.py:155-329 (log_prob, sample, _cond_prob)
// and multinn/utils/auxiliary.py:9-11 (safe_log). Track batching follows
// multinn/models/generators/rnn_multinade.py:231-317.
//
// The recursion a_{i+1} = a_i + v_i * w_enc[i] only changes `a` where the target bit v_i is 1, so
// h_i = sigmoid(a_i) is piecewise constant over i ("segments"); sigmoids are evaluated once per
// segment, the decode dots l_i = b_dec_i + h_i . w_dec[i] for every i. Results are identical to
// evaluating every i; only the transcendental count drops from D*H to (1 + popcount(v)) * H per row.
#include <cstdlib>

#include "common.cuh"
#include "multinn_b200.h"

int mnn_tc_sm_budget();   // gemm_tc.cu: mnn_set_sm_budget of the calling thread (0 = whole device)
// nade_tc.cu: the same forward pass on the tcgen05 tensor cores (segment-row GEMM); nade_fwd_kernel below stays as the
// path for shapes it does not take and as the checker (mnn_set_nade_mode(1))
int mnn_nade_tc_wanted(int D, int H, long long ld, const float* w_enc);
int mnn_nade_tc_fwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0, const float* w_enc,
                    const float* w_dec, float* nll, float* cond_p, float* dfc, float gscale, int N, int M, int D, int H,
                    long long tstride, int sms, cudaStream_t stream);

namespace mnn {

constexpr int kNW = 4;          // 32-bit mask words per (row, track): D <= 128
constexpr int kFwdThreads = 384;
constexpr int kFwdRows = 4;     // rows per warp

struct NadeArgs {
  const uint32_t* bits;  // [M][TS][kNW]  (TS = tstride rows per track >= N: a row chunk of a longer buffer)
  const float* fc;       // [N][ld]   b_enc(m) at col enc_col0 + m*H, b_dec(m) at dec_col0 + m*D
  long long ld;
  int enc_col0, dec_col0;
  const float* w_enc;    // [M][D][H]
  const float* w_dec;    // [M][D][H]
  float* nll;            // [M][TS]
  float* cond_p;         // [M][TS][D] or null
  float* dfc;            // [N][ld] or null: d b_dec columns written by fwd, d b_enc columns by bwd
  float* dw_enc;         // [M][D][H] accumulated (bwd)
  float* dw_dec;         // [M][D][H] accumulated (bwd)
  float gscale;          // d loss / d nll[n,m]
  int N, M, D;
  long long tstride;     // rows between consecutive tracks in bits / nll / cond_p
};

__device__ __forceinline__ uint32_t pick_word(const uint32_t (&w)[kNW], int idx) {
  return idx == 0 ? w[0] : (idx == 1 ? w[1] : (idx == 2 ? w[2] : w[3]));
}

// ------------------------------------------------------------------------------------------------
// K4 forward. One warp = kFwdRows rows of one track; lane owns H/32 hidden units (float4 chunks at
// k = c*128 + lane*4). Eight output dims at a time: 32 partial dots per lane are reduce-scattered
// over the warp so lane L ends with l(row L>>3, dim 8c + (L&7)) and does that element's sigmoid/BCE.
__device__ __forceinline__ float4 sigmoid_fwd4(float4 a) {
  return make_float4(sigmoid_mufu(a.x), sigmoid_mufu(a.y), sigmoid_mufu(a.z), sigmoid_mufu(a.w));
}

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
  const uint32_t* bits = p.bits + (size_t)m * p.tstride * kNW;
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
        h[r][c] = sigmoid_fwd4(a[r][c]);
      }
    }
    const int erow = row0 + er;
    const bool erow_ok = erow < p.N;
    const size_t erow_c = (size_t)min(erow, p.N - 1);
    float nll_acc = 0.f;

#pragma unroll 1
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

      uint32_t upd_any = 0;   // one branch per dim for the common case "no row starts a segment here" (81 % at 5 % density)
#pragma unroll
      for (int r = 0; r < kFwdRows; ++r) upd_any |= upd[r];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = min(i0 + ii, D - 1);  // tail dims recompute dim D-1; discarded in the epilogue
        if (upd_any & (1u << ii)) {
#pragma unroll
          for (int r = 0; r < kFwdRows; ++r) {
            if (upd[r] & (1u << ii)) {  // warp-uniform: v_{i-1} == 1 for this row
              const float* we = wenc_s + (size_t)(i0 + ii - 1) * H;
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(we + c * 128 + lane * 4);
                a[r][c].x += w.x; a[r][c].y += w.y; a[r][c].z += w.z; a[r][c].w += w.w;
                h[r][c] = sigmoid_fwd4(a[r][c]);
              }
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
          if (p.cond_p) p.cond_p[((size_t)m * p.tstride + erow) * D + ei] = pr;
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
    if (eii == 0 && erow_ok) p.nll[(size_t)m * p.tstride + erow] = nll_acc;
  }
}

// ------------------------------------------------------------------------------------------------
// K5 backward. 2-D thread layout: thread (ig, kq) owns hidden units k = 4*kq..4*kq+3 and the dims of group ig
// (G = 4 groups of DG = D/4 consecutive dims): its w_dec / dW_dec tiles [DG][4] live in registers, so every step of
// the walk over dims is 8 packed FFMA2 (two rows in flight) against one uniform branch. Per row the thread walks its
// dims i = hi..lo, reversing the prefix (a -= w_enc[i-1]) at the set target bits inside its group; boundaries come
// from ONE 32-bit word (the group's slice of the row mask), the walk enters one unrolled run of DG steps through a
// DG-entry jump table. With gA_i = dh_i * h_i * (1 - h_i):  d b_enc = sum_i gA_i,  dW_enc[j] += v_j * sum_{i>j} gA_i.
// Each group produces its own total; totals are exchanged through shared memory once per batch of rows so that the
// suffix sums crossing group borders and d b_enc can be completed (fix-up pass). Row data (b_enc row, dl row, mask)
// is staged through a cp.async double buffer one batch ahead. Requires the forward kernel to have written dl into
// the d b_dec columns of dfc.
constexpr int kBwdRows = 8;
constexpr int kBwdGroups = 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
// (one shared reciprocal for the four, sigmoid4_xu, was measured and rejected here: 18.4 -> 18.9 ms backward, 7.4 -> 13.0 ms
// forward -- these kernels are latency-, not XU-bound, and the product / rcp / multiply chain is longer)
__device__ __forceinline__ float4 sigmoid_mufu4(float4 a) {
  return make_float4(sigmoid_mufu(a.x), sigmoid_mufu(a.y), sigmoid_mufu(a.z), sigmoid_mufu(a.w));
}
// acc += (d, d, d, d) * v as two FFMA2
__device__ __forceinline__ void fma4(float4& acc, float d, const float4& v) {
  const float2 dd = make_float2(d, d);
  const float2 lo = ffma2(dd, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
  const float2 hi = ffma2(dd, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void sub4(float4& a, const float4& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
// S += dh * h * (1 - h)
__device__ __forceinline__ void flush_seg(float4& S, const float4& dh, const float4& h) {
  S.x = fmaf(dh.x * h.x, 1.f - h.x, S.x); S.y = fmaf(dh.y * h.y, 1.f - h.y, S.y);
  S.z = fmaf(dh.z * h.z, 1.f - h.z, S.z); S.w = fmaf(dh.w * h.w, 1.f - h.w, S.w);
}

// packed (FADD2 / FMUL2 / FFMA2) forms of the three helpers: same roundings, half the issue slots
__device__ __forceinline__ void add4p(float4& a, const float4& b) {
  const float2 lo = fadd2(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = fadd2(make_float2(a.z, a.w), make_float2(b.z, b.w));
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void sub4p(float4& a, const float4& b) {
  const float2 lo = fsub2(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = fsub2(make_float2(a.z, a.w), make_float2(b.z, b.w));
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void flush_segp(float4& S, const float4& dh, const float4& h) {
  const float2 one = make_float2(1.f, 1.f);
  const float2 hl = make_float2(h.x, h.y), hh = make_float2(h.z, h.w);
  const float2 lo = ffma2(fmul2(make_float2(dh.x, dh.y), hl), fsub2(one, hl), make_float2(S.x, S.y));
  const float2 hi = ffma2(fmul2(make_float2(dh.z, dh.w), hh), fsub2(one, hh), make_float2(S.z, S.w));
  S = make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int H, int D>
struct BwdCfg {
  static constexpr int G = kBwdGroups, DG = D / G, DGP = (DG + 3) / 4 * 4, KQ = H / 4, R = kBwdRows;
  static constexpr size_t kFloats = (size_t)2 * D * H + H  // wenc_s (+ one all-zero row), acce_s
                                    + 2 * R * H            // a_s double buffer
                                    + 2 * R * G * DGP      // dl_s double buffer (per-group padded slices)
                                    + R * G * H;           // tot_s
  static constexpr size_t SMEM = kFloats * sizeof(float) + 2 * R * 8 * sizeof(uint32_t);  // + shifted masks (5 words, padded to 8)
};

// V = 0: the walk enters an unrolled run through a jump table and breaks out at the next boundary (round 1).
// V = 3: V = 2 walking each group's dims UPWARDS. The walk needs the pre-activation prefix a at its first dim: top-down
//        that is every set bit below the group's TOP (group g re-adds about (g + 1) / 4 of the row's bits, 2.5 / 4 on
//        average, and the profile shows this rebuild as the largest single item, 21 % of the stall samples); bottom-up
//        it is the bits below the group's BOTTOM (g / 4 of them, 1.5 / 4 on average, none for group 0). The suffix sums
//        dW_enc[j] += sum_{i > j} gA_i become "row total from this group upwards minus the running prefix inside the
//        group": the walk subtracts the running prefix at each set bit, the fix-up pass adds sum_{g' >= g} total(g').
// V = 2: V = 1 with the float4 add / sub / segment-flush helpers on FADD2 / FMUL2 / FFMA2 (same roundings).
// V = 1: straight-line walk over the group's DG dims with one warp-uniform boundary test per dim (no BRX dispatch, no
//        run loop), and the (group, unit-half) -> warp map chosen so that every SM sub-partition holds one light and one
//        heavy group (the prefix rebuild costs group g about (g + 1) / 4 of the row's set bits).
template <int H, int D, int V>
__global__ void __launch_bounds__(H, 1) nade_bwd_kernel(NadeArgs p) {
  using C_ = BwdCfg<H, D>;
  constexpr bool PK = (V >= 2);
  constexpr bool UP = (V == 3);   // walk the group's dims upwards (see the V = 3 note above the kernel)
  auto ADD4 = [](float4& x, const float4& y) { if constexpr (PK) add4p(x, y); else add4(x, y); };
  auto SUB4 = [](float4& x, const float4& y) { if constexpr (PK) sub4p(x, y); else sub4(x, y); };
  auto FLUSH = [](float4& S, const float4& dh, const float4& h) { if constexpr (PK) flush_segp(S, dh, h); else flush_seg(S, dh, h); };
  constexpr int G = C_::G, DG = C_::DG, DGP = C_::DGP, KQ = C_::KQ, R = C_::R;
  static_assert(D % G == 0 && DG <= 31, "num_dims must be a multiple of 4 and at most 124");
  static_assert(KQ * G == H, "one thread per (hidden-unit quad, dim group)");
  extern __shared__ __align__(16) float smem[];
  float* wenc_s = smem;                            // [D + 1][H], row D = 0
  float* acce_s = wenc_s + (size_t)(D + 1) * H;    // [D][H]
  float* a_s = acce_s + (size_t)D * H;             // [2][R][H]
  float* dl_s = a_s + 2 * R * H;                   // [2][R][G][DGP]
  float* tot_s = dl_s + 2 * R * G * DGP;           // [R][G][H]
  uint32_t* mk_s = reinterpret_cast<uint32_t*>(tot_s + R * G * H);  // [2][R][8]: words 0..4 = (mask << 1), bit i = v_{i-1}

  const int m = blockIdx.x % p.M;
  const int cta = blockIdx.x / p.M;
  const int nctas = (gridDim.x - m + p.M - 1) / p.M;
  const int tid = threadIdx.x;
  int kq = tid % KQ, ig = tid / KQ;                // a warp shares ig (KQ is a multiple of 32)
  if constexpr (V >= 1 && H == 256) {
    // warps 0..7 sit on sub-partitions w % 4: pair groups (0,3) on SMSP 0 / 2 and (1,2) on SMSP 1 / 3
    const int w = tid >> 5;
    ig = (w & 4) ? 3 - (w & 1) : (w & 1);
    kq = ((w >> 1) & 1) * 32 + (tid & 31);
  }
  const int k0 = 4 * kq, lo = ig * DG;
  const float* gwe = p.w_enc + (size_t)m * D * H;
  const float* gwd = p.w_dec + (size_t)m * D * H;

  float4 wd[DG], aw[DG];
#pragma unroll
  for (int ii = 0; ii < DG; ++ii) {
    wd[ii] = __ldg(reinterpret_cast<const float4*>(gwd + (size_t)(lo + ii) * H + k0));
    aw[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int c = tid; c < D * H / 4; c += H) {
    reinterpret_cast<float4*>(wenc_s)[c] = __ldg(reinterpret_cast<const float4*>(gwe) + c);
    reinterpret_cast<float4*>(acce_s)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int c = tid; c < H / 4; c += H) reinterpret_cast<float4*>(wenc_s + (size_t)D * H)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t* bits = p.bits + (size_t)m * p.tstride * kNW;
  const int enc_col = p.enc_col0 + m * H, dec_col = p.dec_col0 + m * D;
  const int nbatches = (p.N + R - 1) / R;

  // staging slots of this thread (loop-invariant: chunk c = tid + it * H of the batch)
  constexpr int A_IT = (R * H / 4 + H - 1) / H, D_IT = (R * D + H - 1) / H;
  int a_r[A_IT], a_off[A_IT], d_r[D_IT], d_i[D_IT], d_dst[D_IT];
#pragma unroll
  for (int it = 0; it < A_IT; ++it) {
    const int c = tid + it * H;
    a_r[it] = c / (H / 4);
    a_off[it] = (c % (H / 4)) * 4;
  }
#pragma unroll
  for (int it = 0; it < D_IT; ++it) {
    const int c = tid + it * H;
    d_r[it] = c < R * D ? c / D : -1;
    d_i[it] = c - (c / D) * D;
    d_dst[it] = ((c / D) * G + d_i[it] / DG) * DGP + (d_i[it] % DG);
  }
  const float* fc_enc = p.fc + enc_col;
  const float* dfc_dec = p.dfc + dec_col;
  auto prefetch = [&](int batch, int buf) {
    const int row0 = batch * R;
    float* ab = a_s + (size_t)buf * R * H;
    float* db = dl_s + (size_t)buf * R * G * DGP;
    uint32_t* mb = mk_s + buf * R * 8;
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {                      // b_enc rows, 16 B chunks
      const int row = row0 + a_r[it];
      float* dst = ab + a_r[it] * H + a_off[it];
      if (row < p.N) cp_async16(dst, fc_enc + (size_t)row * p.ld + a_off[it]);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < D_IT; ++it) {                      // dl rows -> per-group padded slices
      if (d_r[it] < 0) continue;
      const int row = row0 + d_r[it];
      if (row < p.N) cp_async4(db + d_dst[it], dfc_dec + (size_t)row * p.ld + d_i[it]);
      else db[d_dst[it]] = 0.f;
    }
    if (tid < R) {
      const int row = row0 + tid;
      uint4 mm = make_uint4(0u, 0u, 0u, 0u);
      if (row < p.N) mm = __ldg(reinterpret_cast<const uint4*>(bits + (size_t)row * kNW));
      uint32_t* d = mb + tid * 8;
      d[0] = mm.x << 1;
      d[1] = __funnelshift_l(mm.x, mm.y, 1);
      d[2] = __funnelshift_l(mm.y, mm.z, 1);
      d[3] = __funnelshift_l(mm.z, mm.w, 1);
      d[4] = mm.w >> 31;
      d[5] = 0u;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // the group's slice of a shifted row mask: bit q = v_{lo+q-1}, q = 0..DG-1
  auto group_bits = [&](const uint32_t* sm5) {
    const uint32_t w0 = sm5[lo >> 5], w1 = sm5[(lo >> 5) + 1];
    return __funnelshift_r(w0, w1, lo & 31) & ((1u << DG) - 1u);
  };

  int buf = 0;
  if (cta < nbatches) prefetch(cta, 0);
  for (int b = cta; b < nbatches; b += nctas) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // batch b staged and visible; everyone is done with batch b - nctas (fix-up included)
    if (b + nctas < nbatches) prefetch(b + nctas, buf ^ 1);
    const int row0 = b * R;
    const float* ab = a_s + (size_t)buf * R * H;
    const float* db = dl_s + (size_t)buf * R * G * DGP;
    const uint32_t* mb = mk_s + buf * R * 8;

#pragma unroll 1
    for (int r = 0; r < R; r += 2) {
      const uint32_t* sm0 = mb + r * 8;
      const uint32_t* sm1 = mb + (r + 1) * 8;
      float4 a0 = *reinterpret_cast<const float4*>(ab + r * H + k0);
      float4 a1 = *reinterpret_cast<const float4*>(ab + (r + 1) * H + k0);
      // forward prefix up to the walk's first dim: top-down hi = lo + DG - 1, bottom-up hi = lo: all set target bits
      // j < hi, i.e. shifted bits 1..hi
      {
        const int hi = UP ? lo : lo + DG - 1;
#pragma unroll
        for (int w = 0; w < kNW; ++w) {
          if (w * 32 > hi) break;    // warp-uniform
          const int nb = hi - w * 32 + 1;   // shifted bits w*32 .. hi
          const uint32_t keep = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
          uint32_t x0 = sm0[w] & keep, x1 = sm1[w] & keep;
          if (w == 0) { x0 &= ~1u; x1 &= ~1u; }
          while (x0) {
            const int j = w * 32 + __ffs(x0) - 2;
            x0 &= x0 - 1;
            ADD4(a0, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
          }
          while (x1) {
            const int j = w * 32 + __ffs(x1) - 2;
            x1 &= x1 - 1;
            ADD4(a1, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
          }
        }
      }
      float4 h0 = sigmoid_mufu4(a0), h1 = sigmoid_mufu4(a1);
      float4 dh0 = make_float4(0.f, 0.f, 0.f, 0.f), dh1 = dh0, S0 = dh0, S1 = dh0;
      // boundaries inside the group: bit q (1 <= q < DG) set <=> v_{lo+q-1} = 1 <=> new segment below local dim q
      const uint32_t bm0 = group_bits(sm0) & ~1u, bm1 = group_bits(sm1) & ~1u;
      uint32_t rm = bm0 | bm1;
      const float* dl0 = db + (r * G + ig) * DGP;
      const float* dl1 = db + ((r + 1) * G + ig) * DGP;
      auto boundary = [&](int s) {   // the set target bit j = lo + s - 1 ends the segment below local dim s
        const int j = lo + s - 1;
        if (bm0 & (1u << s)) {      // block-uniform within the warp (a warp shares ig and the row)
          FLUSH(S0, dh0, h0);
          dh0 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4* ae = reinterpret_cast<float4*>(acce_s + (size_t)j * H + k0);
          float4 acc = *ae;
          ADD4(acc, S0);
          *ae = acc;
          SUB4(a0, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
          h0 = sigmoid_mufu4(a0);
        }
        if (bm1 & (1u << s)) {
          FLUSH(S1, dh1, h1);
          dh1 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4* ae = reinterpret_cast<float4*>(acce_s + (size_t)j * H + k0);
          float4 acc = *ae;
          ADD4(acc, S1);
          *ae = acc;
          SUB4(a1, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
          h1 = sigmoid_mufu4(a1);
        }
      };
      if constexpr (UP) {
        // unshifted group bits: bit q = v_{lo+q}; a set bit ends the segment ABOVE local dim q
        const int l1 = lo + 1;
        const uint32_t um0 = __funnelshift_r(sm0[l1 >> 5], sm0[(l1 >> 5) + 1], l1 & 31) & ((1u << DG) - 1u);
        const uint32_t um1 = __funnelshift_r(sm1[l1 >> 5], sm1[(l1 >> 5) + 1], l1 & 31) & ((1u << DG) - 1u);
        const uint32_t um = um0 | um1;
        auto boundary_up = [&](int s, bool last) {   // after local dim s = target bit j = lo + s
          const int j = lo + s;
          if (um0 & (1u << s)) {
            FLUSH(S0, dh0, h0);                      // S0 = running prefix of gA over the group's dims <= j
            dh0 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* ae = reinterpret_cast<float4*>(acce_s + (size_t)j * H + k0);
            float4 acc = *ae;
            SUB4(acc, S0);
            *ae = acc;
            if (!last) {
              ADD4(a0, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
              h0 = sigmoid_mufu4(a0);
            }
          }
          if (um1 & (1u << s)) {
            FLUSH(S1, dh1, h1);
            dh1 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* ae = reinterpret_cast<float4*>(acce_s + (size_t)j * H + k0);
            float4 acc = *ae;
            SUB4(acc, S1);
            *ae = acc;
            if (!last) {
              ADD4(a1, *reinterpret_cast<const float4*>(wenc_s + (size_t)j * H + k0));
              h1 = sigmoid_mufu4(a1);
            }
          }
        };
        float4 q0, q1;
#pragma unroll
        for (int I = 0; I < DG; ++I) {
          if ((I & 3) == 0) {
            q0 = *reinterpret_cast<const float4*>(dl0 + I);
            q1 = *reinterpret_cast<const float4*>(dl1 + I);
          }
          const float d0 = (I & 3) == 3 ? q0.w : ((I & 3) == 2 ? q0.z : ((I & 3) == 1 ? q0.y : q0.x));
          const float d1 = (I & 3) == 3 ? q1.w : ((I & 3) == 2 ? q1.z : ((I & 3) == 1 ? q1.y : q1.x));
          fma4(dh0, d0, wd[I]);
          fma4(dh1, d1, wd[I]);
          fma4(aw[I], d0, h0);
          fma4(aw[I], d1, h1);
          if (um & (1u << I)) boundary_up(I, I == DG - 1);
        }
      } else if constexpr (V >= 1) {
        float4 q0, q1;
#pragma unroll
        for (int I = DG - 1; I >= 0; --I) {
          if ((I & 3) == 3 || I == DG - 1) {
            q0 = *reinterpret_cast<const float4*>(dl0 + (I & ~3));
            q1 = *reinterpret_cast<const float4*>(dl1 + (I & ~3));
          }
          const float d0 = (I & 3) == 3 ? q0.w : ((I & 3) == 2 ? q0.z : ((I & 3) == 1 ? q0.y : q0.x));
          const float d1 = (I & 3) == 3 ? q1.w : ((I & 3) == 2 ? q1.z : ((I & 3) == 1 ? q1.y : q1.x));
          fma4(dh0, d0, wd[I]);
          fma4(dh1, d1, wd[I]);
          fma4(aw[I], d0, h0);
          fma4(aw[I], d1, h1);
          if (I > 0 && (rm & (1u << I))) boundary(I);
        }
      } else {
      int i = DG - 1;
#pragma unroll 1
      for (;;) {
        int s = 0;   // process local dims i..s, then handle the boundary below dim s (if s > 0)
        if (rm) {
          s = 31 - __clz(rm);
          rm &= ~(1u << s);
        }
        float4 q0 = *reinterpret_cast<const float4*>(dl0 + (i & ~3));
        float4 q1 = *reinterpret_cast<const float4*>(dl1 + (i & ~3));
#define MNN_BWD_STEP(I)                                                                                   \
  case (I):                                                                                               \
    if constexpr ((I) < DG) {                                                                             \
      if (((I) & 3) == 3) {                                                                               \
        q0 = *reinterpret_cast<const float4*>(dl0 + ((I) & ~3));                                          \
        q1 = *reinterpret_cast<const float4*>(dl1 + ((I) & ~3));                                          \
      }                                                                                                   \
      const float d0 = ((I) & 3) == 3 ? q0.w : (((I) & 3) == 2 ? q0.z : (((I) & 3) == 1 ? q0.y : q0.x));  \
      const float d1 = ((I) & 3) == 3 ? q1.w : (((I) & 3) == 2 ? q1.z : (((I) & 3) == 1 ? q1.y : q1.x));  \
      fma4(dh0, d0, wd[(I) < DG ? (I) : 0]);                                                              \
      fma4(dh1, d1, wd[(I) < DG ? (I) : 0]);                                                              \
      fma4(aw[(I) < DG ? (I) : 0], d0, h0);                                                               \
      fma4(aw[(I) < DG ? (I) : 0], d1, h1);                                                               \
      if ((I) == s) break;                                                                                \
    }
#define MNN_BWD_STEP4(I) MNN_BWD_STEP((I) + 3) MNN_BWD_STEP((I) + 2) MNN_BWD_STEP((I) + 1) MNN_BWD_STEP(I)
        switch (i) {
          MNN_BWD_STEP4(28) MNN_BWD_STEP4(24) MNN_BWD_STEP4(20) MNN_BWD_STEP4(16)
          MNN_BWD_STEP4(12) MNN_BWD_STEP4(8) MNN_BWD_STEP4(4) MNN_BWD_STEP4(0)
          default: break;
        }
#undef MNN_BWD_STEP4
#undef MNN_BWD_STEP
        if (s == 0) break;
        boundary(s);
        i = s - 1;
      }
      }
      FLUSH(S0, dh0, h0);
      FLUSH(S1, dh1, h1);
      *reinterpret_cast<float4*>(tot_s + ((size_t)r * G + ig) * H + k0) = S0;
      *reinterpret_cast<float4*>(tot_s + ((size_t)(r + 1) * G + ig) * H + k0) = S1;
    }
    __syncthreads();   // group totals of the batch are complete

    // fix-up: suffix sums that cross group borders, then d b_enc
    if (UP || ig < G - 1) {
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        // set target bits j in [lo, lo + DG): shifted bits lo+1 .. lo+DG
        const uint32_t* sm5 = mb + r * 8;
        const int l1 = lo + 1;
        uint32_t x = __funnelshift_r(sm5[l1 >> 5], sm5[(l1 >> 5) + 1], l1 & 31) & ((1u << DG) - 1u);
        if (!x) continue;
        float4 higher = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = UP ? ig : ig + 1; g < G; ++g)   // bottom-up: own total included (the walk subtracted the running prefix)
          ADD4(higher, *reinterpret_cast<const float4*>(tot_s + ((size_t)r * G + g) * H + k0));
        while (x) {
          const int j = lo + __ffs(x) - 1;
          x &= x - 1;
          float4* ae = reinterpret_cast<float4*>(acce_s + (size_t)j * H + k0);
          float4 acc = *ae;
          ADD4(acc, higher);
          *ae = acc;
        }
      }
    }
    for (int r = ig; r < R; r += G) {
      const int row = row0 + r;
      if (row < p.N) {
        float4 t = *reinterpret_cast<const float4*>(tot_s + ((size_t)r * G) * H + k0);
#pragma unroll
        for (int g = 1; g < G; ++g) ADD4(t, *reinterpret_cast<const float4*>(tot_s + ((size_t)r * G + g) * H + k0));
        *reinterpret_cast<float4*>(p.dfc + (size_t)row * p.ld + enc_col + k0) = t;
      }
    }
    buf ^= 1;
  }
  __syncthreads();
  float* gdd = p.dw_dec + (size_t)m * D * H;
  float* gde = p.dw_enc + (size_t)m * D * H;
#pragma unroll
  for (int ii = 0; ii < DG; ++ii) atomicAdd(reinterpret_cast<float4*>(gdd + (size_t)(lo + ii) * H + k0), aw[ii]);
  for (int c = tid; c < D * H / 4; c += H)
    atomicAdd(reinterpret_cast<float4*>(gde) + c, reinterpret_cast<const float4*>(acce_s)[c]);
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
  RowMap rmap;   // Philox counter = ((global_row(row) * M + m) * D + i, offset)
};

// 1024 threads: every dim of a row is a dependent chain (dot -> reduce -> sigmoid -> compare -> maybe 8 more sigmoids), so
// the kernel lives on warps in flight: 32 warps per SM instead of 8 (weights of one track fill shared memory: 1 CTA/SM).
constexpr int kSampleThreads = 1024;
template <int NCH>
__global__ void __launch_bounds__(kSampleThreads, 1) nade_sample_kernel(NadeSampleArgs p) {
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
          const unsigned long long idx = (global_row(p.rmap, (unsigned long long)row) * p.M + m) * D + i;
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
static int device_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}
// grids of the persistent NADE kernels honour the caller's SM budget so that they can run beside the recurrence
// kernels of other streams (time-chunk pipeline at small per-GPU batches)
static int num_sms() {
  const int n = device_sms(), b = ::mnn_tc_sm_budget();
  return (b > 0 && b < n) ? b : n;
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
                                    float gscale, int N, int M, int D, int H, long long track_stride,
                                    cudaStream_t stream) {
  MNN_REQUIRE(bits && fc && w_enc && w_dec && nll, MNN_ERR_ARG, "nade_logprob_fwd: null pointer");
  int rc = check_nade_common(N, M, D, H, ld, enc_col0, dec_col0);
  if (rc) return rc;
  MNN_REQUIRE(track_stride == 0 || track_stride >= N, MNN_ERR_ARG, "nade_logprob_fwd: track_stride < N");
  if (mnn_nade_tc_wanted(D, H, ld, w_enc))
    return mnn_nade_tc_fwd(bits, fc, ld, enc_col0, dec_col0, w_enc, w_dec, nll, cond_p, dfc, gscale, N, M, D, H,
                           track_stride ? track_stride : N, num_sms(), stream);
  NadeArgs a{bits, fc, ld, enc_col0, dec_col0, w_enc, w_dec, nll, cond_p, dfc, nullptr, nullptr, gscale, N, M, D,
             track_stride ? track_stride : N};
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

static int nade_bwd_variant() {   // MNN_NADE_BWD_V: 0 = round-1 jump-table walk, 1 = straight-line walk, 2 = 1 + packed helpers, 3 (default) = 2 walking upwards
  static const int v = [] { const char* e = getenv("MNN_NADE_BWD_V"); return e ? atoi(e) : 3; }();
  return v;
}

template <int H, int D, int V>
static int launch_bwd_v(const NadeArgs& a, cudaStream_t stream) {
  const size_t smem = BwdCfg<H, D>::SMEM;
  cudaFuncSetAttribute(nade_bwd_kernel<H, D, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = num_sms();
  const int need = a.M * ((a.N + kBwdRows - 1) / kBwdRows);
  if (grid > need) grid = need;
  if (grid < a.M) grid = a.M;
  nade_bwd_kernel<H, D, V><<<grid, H, smem, stream>>>(a);
  return mnn_check_launch("nade_logprob_bwd");
}

template <int H, int D>
static int launch_bwd(const NadeArgs& a, cudaStream_t stream) {
  switch (nade_bwd_variant()) {
    case 0: return launch_bwd_v<H, D, 0>(a, stream);
    case 1: return launch_bwd_v<H, D, 1>(a, stream);
    case 2: return launch_bwd_v<H, D, 2>(a, stream);
    default: return launch_bwd_v<H, D, 3>(a, stream);
  }
}

extern "C" int mnn_nade_logprob_bwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0,
                                    const float* w_enc, const float* w_dec, float* dfc, float* dw_enc, float* dw_dec,
                                    int N, int M, int D, int H, long long track_stride, cudaStream_t stream) {
  MNN_REQUIRE(bits && fc && w_enc && w_dec && dfc && dw_enc && dw_dec, MNN_ERR_ARG, "nade_logprob_bwd: null pointer");
  int rc = check_nade_common(N, M, D, H, ld, enc_col0, dec_col0);
  if (rc) return rc;
  MNN_REQUIRE(track_stride == 0 || track_stride >= N, MNN_ERR_ARG, "nade_logprob_bwd: track_stride < N");
  NadeArgs a{bits, fc, ld, enc_col0, dec_col0, w_enc, w_dec, nullptr, nullptr, dfc, dw_enc, dw_dec, 0.f, N, M, D,
             track_stride ? track_stride : N};
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
                   nll, N, M, D, seed, offset, use_philox, current_row_map()};
  const size_t smem = (size_t)2 * D * H * sizeof(float);
  int grid = num_sms();
  const int wpc = kSampleThreads / 32;
  const int need = M * ((N + wpc - 1) / wpc);
  if (grid > need) grid = need;
  if (grid < M) grid = M;
  if (H == 256) {
    cudaFuncSetAttribute(nade_sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_sample_kernel<2><<<grid, kSampleThreads, smem, stream>>>(a);
  } else {
    cudaFuncSetAttribute(nade_sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nade_sample_kernel<1><<<grid, kSampleThreads, smem, stream>>>(a);
  }
  return mnn_check_launch("nade_sample");
}
