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
      const unsigned long long e = a.offset + idx;
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
  SigSampleArgs a{pre, ld_pre, bias, ld_bias, u, ld_u, p, ld_p, s, ld_s, N, C, u ? 1 : (use_philox ? 2 : 0), seed, offset};
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
