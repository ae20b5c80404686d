// RBM / DBN kernels (K7). Follows reference common/rbm.py:148-231 (forward, reconstruct, k-step Gibbs
// chain), :233-263 (free energy), :337-387 (_cond_prob_h/_v, Bernoulli _sample) and common/dbn.py:136-180.
// Bernoulli sampling contract (TFP 0.6.0): sample = float(u < p), strict.
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

struct SigSampleArgs {
  const float* pre; long long ld_pre;     // [N][C] pre-activations without bias (GEMM output)
  const float* bias; long long ld_bias;   // per-row bias [N][C] (ld_bias > 0) or broadcast row (ld_bias == 0) or null
  const float* u; long long ld_u;         // uniforms [N][C] or null
  float* p; long long ld_p;               // probabilities out or null
  float* s; long long ld_s;               // samples out or null
  int N, C;
  int mode;                               // 0: no sampling, 1: u supplied, 2: philox
  unsigned long long seed, offset;
  RowMap rmap;                            // Philox counter = offset + global_row(r) * C + c
};

__global__ void bias_sigmoid_sample_kernel(SigSampleArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.N * a.C) return;
  const int r = (int)(idx / a.C), c = (int)(idx - (size_t)r * a.C);
  float x = a.pre[(size_t)r * a.ld_pre + c];
  if (a.bias) x += a.bias[(size_t)r * a.ld_bias + c];
  const float pr = sigmoid_acc(x);
  if (a.p) a.p[(size_t)r * a.ld_p + c] = pr;
  if (a.s) {
    float uu = 0.f;
    if (a.mode == 1) uu = a.u[(size_t)r * a.ld_u + c];
    else if (a.mode == 2) {
      const unsigned long long e = a.offset + global_row(a.rmap, (unsigned long long)r) * a.C + c;
      const uint4 r4 = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), 0u, 0u),
                                     make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
      uu = u01(r4.x);
    }
    a.s[(size_t)r * a.ld_s + c] = (uu < pr) ? 1.f : 0.f;
  }
}

// F(v)[n] = -sum_j softplus(pre[n][j] + bh[j]) - v[n] . bv     (one warp per row)
__global__ void free_energy_kernel(const float* pre, long long ld_pre, const float* bh, long long ld_bh,
                                   const float* v, long long ld_v, const float* bv, long long ld_bv, float* F, int N,
                                   int H, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float s = 0.f;
  for (int j = lane; j < H; j += 32) {
    const float x = pre[(size_t)row * ld_pre + j] + bh[(size_t)row * ld_bh + j];
    s += fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));  // log(1 + exp(x)), overflow-safe
  }
  float d = 0.f;
  for (int i = lane; i < D; i += 32) d += v[(size_t)row * ld_v + i] * bv[(size_t)row * ld_bv + i];
  s = warp_sum(s);
  d = warp_sum(d);
  if (lane == 0) F[row] = -s - d;
}

// d(pre) = dy * y * (1 - y)  (backward of the sigmoid Dense feedback layers, common/dnn.py:56-60)
__global__ void sigmoid_bwd_kernel(const float* __restrict__ y, long long ld_y, const float* __restrict__ dy, long long ld_dy,
                                   float* __restrict__ dpre, long long ld_d, int N, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int r = (int)(idx / C), c = (int)(idx - (size_t)r * C);
  const float yy = y[(size_t)r * ld_y + c];
  dpre[(size_t)r * ld_d + c] = dy[(size_t)r * ld_dy + c] * yy * (1.f - yy);
}


// ---------------------------------------------------------------------------------------------------------------
// Fused k-step Gibbs chain (reference common/rbm.py:192-231: tf.while_loop over k x {_cond_prob_h + _sample,
// _cond_prob_v + _sample}, :337-387). ONE launch for the whole chain:
//   * W[D][H] and its transpose Wt[H][D] are staged once per CTA in shared memory (2*D*H*4 bytes: 172 KB at 84 x 256),
//     so both half-steps are the same "input element broadcast x weight row" loop with conflict-free 128-bit reads;
//   * a warp owns 4 rows at a time; the binary states v_s / h_s live in a per-warp shared buffer laid out [dim][row] so
//     that one 128-bit broadcast read feeds the 4 rows (a per-dim "all 4 rows are 0" skip was measured and dropped: it
//     only pays in the first half-step and its branch kept the loop from being unrolled - 4.75 ms per C3 chain);
//   * lane l owns output columns 4*l + 128*q: bias (per-row or broadcast) and the uniforms come in as 128-bit loads,
//     sigmoid + strict `u < p` Bernoulli (TFP 0.6.0 contract) run in registers, FFMA2 on row pairs;
//   * HBM traffic per row: v0, bh_t, bv_t in, p_v and v_k (and optionally h_k) out - nothing between the 2k half-steps.
// Philox mode: counter = (row64 = offset + row, half-step index 2*s + {0,1}, column group c/4), key = seed; the 4 words
// of one call serve the 4 columns of a group. The stream is a function of the global row index only (independent of
// how the batch is sharded over GPUs); it differs from the stream of the unfused half-step kernel above.
struct GibbsArgs {
  const float* v0; long long ld_v;
  const float* W;                          // [D][H] contiguous
  const float* bh; long long ld_bh;        // [N][H] (ld_bh > 0) or one row (ld_bh == 0)
  const float* bv; long long ld_bv;
  const float* uh;                         // [k][N][H] contiguous or null
  const float* uv;                         // [k][N][D] contiguous or null
  float* p_v; long long ld_p;              // last step's p(v | h_k)            (nullable)
  float* v_k; long long ld_vk;             // last step's visible sample        (nullable)
  float* h_k; long long ld_hk;             // last step's hidden sample         (nullable)
  int N, D, H, k;
  int philox;
  unsigned long long seed, offset;
  RowMap rmap;                             // Philox counter row = offset + global_row(row)
};

constexpr int kGibbsThreads = 256;
constexpr int kGibbsRows = 4;              // rows per warp

// acc[q][c][0] = rows (0,1), acc[q][c][1] = rows (2,3) of output column 4*lane + 128*q + c.
// Four input dims per trip (in_dim % 4 == 0): 4 broadcast reads of the rows' inputs and 4*NQ weight reads are in flight
// before the 32*NQ FFMA2 that consume them; the CTA's second warp per scheduler covers the shared-memory latency.
template <int NQ>
__device__ __forceinline__ void gibbs_matvec(const float* __restrict__ wsm, int in_dim, int out_dim,
                                             const float* __restrict__ inbuf, int lane, float2 (&acc)[NQ][4][2]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[q][c][0] = acc[q][c][1] = make_float2(0.f, 0.f);
  bool act[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) act[q] = 4 * lane + 128 * q < out_dim;
  for (int i0 = 0; i0 < in_dim; i0 += 4) {
    float4 x[4], w[NQ][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(inbuf + 4 * (i0 + u));   // 4 rows' input i0+u
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[q][u] = act[q] ? *reinterpret_cast<const float4*>(wsm + (size_t)(i0 + u) * out_dim + 4 * lane + 128 * q)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 x01 = make_float2(x[u].x, x[u].y), x23 = make_float2(x[u].z, x[u].w);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        acc[q][0][0] = ffma2(x01, make_float2(w[q][u].x, w[q][u].x), acc[q][0][0]);
        acc[q][0][1] = ffma2(x23, make_float2(w[q][u].x, w[q][u].x), acc[q][0][1]);
        acc[q][1][0] = ffma2(x01, make_float2(w[q][u].y, w[q][u].y), acc[q][1][0]);
        acc[q][1][1] = ffma2(x23, make_float2(w[q][u].y, w[q][u].y), acc[q][1][1]);
        acc[q][2][0] = ffma2(x01, make_float2(w[q][u].z, w[q][u].z), acc[q][2][0]);
        acc[q][2][1] = ffma2(x23, make_float2(w[q][u].z, w[q][u].z), acc[q][2][1]);
        acc[q][3][0] = ffma2(x01, make_float2(w[q][u].w, w[q][u].w), acc[q][3][0]);
        acc[q][3][1] = ffma2(x23, make_float2(w[q][u].w, w[q][u].w), acc[q][3][1]);
      }
    }
  }
}

// bias + sigmoid + Bernoulli on the accumulators; the samples go to outbuf[col][row] (the next half-step's input) and,
// when asked, probabilities / samples to global memory as 128-bit stores.
template <int NQ>
__device__ __forceinline__ void gibbs_epilogue(float2 (&acc)[NQ][4][2], int out_dim, int lane, long long row0, int N,
                                               const float* __restrict__ bias, long long ld_bias,
                                               const float* __restrict__ u /* [N][out_dim] of this half-step or null */,
                                               int philox, unsigned long long seed, unsigned long long offset,
                                               const RowMap& rmap, unsigned half_step, float* __restrict__ outbuf, float* __restrict__ p_out,
                                               long long ld_p, float* __restrict__ s_out, long long ld_s) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int col = 4 * lane + 128 * q;
    if (col >= out_dim) continue;
    float smp[4][4];   // [row][c]
#pragma unroll
    for (int r = 0; r < kGibbsRows; ++r) {
      const long long row = row0 + r;
      float pre[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 a = acc[q][c][r >> 1];
        pre[c] = (r & 1) ? a.y : a.x;
      }
      if (row < N) {
        if (bias) {
          const float4 b = *reinterpret_cast<const float4*>(bias + (size_t)row * ld_bias + col);
          pre[0] += b.x; pre[1] += b.y; pre[2] += b.z; pre[3] += b.w;
        }
        float4 uu = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u) {
          uu = *reinterpret_cast<const float4*>(u + (size_t)row * out_dim + col);
        } else if (philox) {
          const unsigned long long e = offset + global_row(rmap, (unsigned long long)row);
          const uint4 r4 = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), half_step, (uint32_t)(col >> 2)),
                                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
          uu = make_float4(u01(r4.x), u01(r4.y), u01(r4.z), u01(r4.w));
        }
        const float4 pr = make_float4(sigmoid_acc(pre[0]), sigmoid_acc(pre[1]), sigmoid_acc(pre[2]), sigmoid_acc(pre[3]));
        smp[r][0] = uu.x < pr.x ? 1.f : 0.f;
        smp[r][1] = uu.y < pr.y ? 1.f : 0.f;
        smp[r][2] = uu.z < pr.z ? 1.f : 0.f;
        smp[r][3] = uu.w < pr.w ? 1.f : 0.f;
        if (p_out) *reinterpret_cast<float4*>(p_out + (size_t)row * ld_p + col) = pr;
        if (s_out)
          *reinterpret_cast<float4*>(s_out + (size_t)row * ld_s + col) = make_float4(smp[r][0], smp[r][1], smp[r][2], smp[r][3]);
      } else {
        smp[r][0] = smp[r][1] = smp[r][2] = smp[r][3] = 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(outbuf + 4 * (col + c)) = make_float4(smp[0][c], smp[1][c], smp[2][c], smp[3][c]);
  }
}

template <int NQH, int NQD>
__global__ void __launch_bounds__(kGibbsThreads, 1) rbm_gibbs_kernel(GibbsArgs a) {
  extern __shared__ __align__(16) float gsm[];
  const int D = a.D, H = a.H;
  float* wsm = gsm;                          // [D][H]
  float* wtsm = gsm + (size_t)D * H;         // [H][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* vbuf = wtsm + (size_t)D * H + (size_t)warp * 4 * (D + H);   // [D][4]
  float* hbuf = vbuf + 4 * D;                                          // [H][4]
  for (int idx = threadIdx.x; idx < D * H / 4; idx += kGibbsThreads) {
    const float4 w = reinterpret_cast<const float4*>(a.W)[idx];        // 128-bit loads of W
    reinterpret_cast<float4*>(wsm)[idx] = w;
    const int i = (4 * idx) / H, j = (4 * idx) - i * H;                // H % 4 == 0: the 4 values share row i
    wtsm[(size_t)(j + 0) * D + i] = w.x;
    wtsm[(size_t)(j + 1) * D + i] = w.y;
    wtsm[(size_t)(j + 2) * D + i] = w.z;
    wtsm[(size_t)(j + 3) * D + i] = w.w;
  }
  __syncthreads();
  const long long groups = ((long long)a.N + kGibbsRows - 1) / kGibbsRows;
  const int wpc = kGibbsThreads / 32;
  for (long long g = (long long)blockIdx.x * wpc + warp; g < groups; g += (long long)gridDim.x * wpc) {
    const long long row0 = g * kGibbsRows;
    __syncwarp();
    for (int i = lane; i < D; i += 32) {
      float4 x;
      x.x = row0 + 0 < a.N ? a.v0[(size_t)(row0 + 0) * a.ld_v + i] : 0.f;
      x.y = row0 + 1 < a.N ? a.v0[(size_t)(row0 + 1) * a.ld_v + i] : 0.f;
      x.z = row0 + 2 < a.N ? a.v0[(size_t)(row0 + 2) * a.ld_v + i] : 0.f;
      x.w = row0 + 3 < a.N ? a.v0[(size_t)(row0 + 3) * a.ld_v + i] : 0.f;
      *reinterpret_cast<float4*>(vbuf + 4 * i) = x;
    }
    __syncwarp();
    for (int s = 0; s < a.k; ++s) {
      const bool last = s == a.k - 1;
      {
        float2 acc[NQH][4][2];
        gibbs_matvec<NQH>(wsm, D, H, vbuf, lane, acc);
        gibbs_epilogue<NQH>(acc, H, lane, row0, a.N, a.bh, a.ld_bh,
                            a.uh ? a.uh + (size_t)s * a.N * H : nullptr, a.philox, a.seed, a.offset, a.rmap, 2u * s, hbuf,
                            nullptr, 0, last ? a.h_k : nullptr, a.ld_hk);
      }
      __syncwarp();
      {
        float2 acc[NQD][4][2];
        gibbs_matvec<NQD>(wtsm, H, D, hbuf, lane, acc);
        __syncwarp();   // every lane has read vbuf's predecessor state before it is overwritten (it was: matvec above)
        gibbs_epilogue<NQD>(acc, D, lane, row0, a.N, a.bv, a.ld_bv,
                            a.uv ? a.uv + (size_t)s * a.N * D : nullptr, a.philox, a.seed, a.offset, a.rmap, 2u * s + 1u, vbuf,
                            last ? a.p_v : nullptr, a.ld_p, last ? a.v_k : nullptr, a.ld_vk);
      }
      __syncwarp();
    }
  }
}

}  // namespace mnn

using namespace mnn;

extern "C" int mnn_sigmoid_bwd(const float* y, long long ld_y, const float* dy, long long ld_dy, float* dpre,
                               long long ld_d, int N, int C, cudaStream_t stream) {
  MNN_REQUIRE(y && dy && dpre && N > 0 && C > 0, MNN_ERR_ARG, "sigmoid_bwd: bad argument");
  const size_t n = (size_t)N * C;
  sigmoid_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(y, ld_y, dy, ld_dy, dpre, ld_d, N, C);
  return mnn_check_launch("sigmoid_bwd");
}

extern "C" int mnn_bias_sigmoid_sample(const float* pre, long long ld_pre, const float* bias, long long ld_bias,
                                       const float* u, long long ld_u, int use_philox, unsigned long long seed,
                                       unsigned long long offset, float* p, long long ld_p, float* s, long long ld_s,
                                       int N, int C, cudaStream_t stream) {
  MNN_REQUIRE(pre && (p || s) && N > 0 && C > 0, MNN_ERR_ARG, "bias_sigmoid_sample: bad argument");
  MNN_REQUIRE(!(s && !u && !use_philox), MNN_ERR_ARG, "bias_sigmoid_sample: sampling needs uniforms or philox");
  SigSampleArgs a{pre, ld_pre, bias, ld_bias, u, ld_u, p, ld_p, s, ld_s, N, C, u ? 1 : (use_philox ? 2 : 0), seed, offset, current_row_map()};
  const size_t n = (size_t)N * C;
  bias_sigmoid_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a);
  return mnn_check_launch("bias_sigmoid_sample");
}

extern "C" int mnn_rbm_free_energy(const float* pre, long long ld_pre, const float* bh, long long ld_bh,
                                   const float* v, long long ld_v, const float* bv, long long ld_bv, float* F, int N,
                                   int H, int D, cudaStream_t stream) {
  MNN_REQUIRE(pre && bh && v && bv && F && N > 0 && H > 0 && D > 0, MNN_ERR_ARG, "rbm_free_energy: bad argument");
  free_energy_kernel<<<(N + 7) / 8, 256, 0, stream>>>(pre, ld_pre, bh, ld_bh, v, ld_v, bv, ld_bv, F, N, H, D);
  return mnn_check_launch("rbm_free_energy");
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" size_t mnn_rbm_gibbs_smem_bytes(int D, int H) {
  if (D <= 0 || H <= 0 || D % 4 || H % 4 || D > 256 || H > 256) return 0;
  const size_t b = ((size_t)2 * D * H + (size_t)(kGibbsThreads / 32) * 4 * (D + H)) * sizeof(float);
  return b <= 227 * 1024 ? b : 0;
}

extern "C" int mnn_rbm_gibbs(const float* v0, long long ld_v, const float* W, const float* bh, long long ld_bh,
                             const float* bv, long long ld_bv, const float* uh, const float* uv, int use_philox,
                             unsigned long long seed, unsigned long long offset, float* p_v, long long ld_p, float* v_k,
                             long long ld_vk, float* h_k, long long ld_hk, int N, int D, int H, int k,
                             cudaStream_t stream) {
  MNN_REQUIRE(v0 && W && (p_v || v_k || h_k) && N > 0 && D > 0 && H > 0 && k > 0, MNN_ERR_ARG, "rbm_gibbs: bad argument");
  MNN_REQUIRE((uh && uv) || (!uh && !uv && use_philox), MNN_ERR_ARG,
              "rbm_gibbs: needs both uniform tensors uh[k,N,H], uv[k,N,D] or philox");
  const size_t smem = mnn_rbm_gibbs_smem_bytes(D, H);
  MNN_REQUIRE(smem > 0, MNN_ERR_UNSUPPORTED,
              "rbm_gibbs: needs D, H multiples of 4, <= 256, with W and its transpose fitting in shared memory");
  MNN_REQUIRE(aligned16(W) && aligned16(bh) && aligned16(bv) && aligned16(uh) && aligned16(uv) && aligned16(p_v) &&
                  aligned16(v_k) && aligned16(h_k) && ld_bh % 4 == 0 && ld_bv % 4 == 0 && ld_p % 4 == 0 &&
                  ld_vk % 4 == 0 && ld_hk % 4 == 0,
              MNN_ERR_ARG, "rbm_gibbs: pointers must be 16-byte aligned and row strides multiples of 4 floats");
  GibbsArgs a{v0, ld_v, W, bh, ld_bh, bv, ld_bv, uh, uv, p_v, ld_p, v_k, ld_vk, h_k, ld_hk, N, D, H, k,
              use_philox ? 1 : 0, seed, offset, current_row_map()};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const long long groups = ((long long)N + kGibbsRows - 1) / kGibbsRows;
  const long long need = (groups + kGibbsThreads / 32 - 1) / (kGibbsThreads / 32);
  const int grid = (int)(need < sms ? need : sms);
#define MNN_GIBBS_LAUNCH(QH, QD)                                                                              \
  do {                                                                                                        \
    cudaFuncSetAttribute(rbm_gibbs_kernel<QH, QD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    rbm_gibbs_kernel<QH, QD><<<grid, kGibbsThreads, smem, stream>>>(a);                                       \
  } while (0)
  const int qh = (H + 127) / 128, qd = (D + 127) / 128;
  if (qh == 1 && qd == 1) MNN_GIBBS_LAUNCH(1, 1);
  else if (qh == 2 && qd == 1) MNN_GIBBS_LAUNCH(2, 1);
  else if (qh == 1 && qd == 2) MNN_GIBBS_LAUNCH(1, 2);
  else MNN_GIBBS_LAUNCH(2, 2);
#undef MNN_GIBBS_LAUNCH
  return mnn_check_launch("rbm_gibbs");
}
