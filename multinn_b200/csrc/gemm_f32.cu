// fp32 CUDA-core GEMM (exact-fp32 path): C = alpha * op(A) * op(B) + beta * C (+ bias[n]).
// Used where bit-for-bit fp32 arithmetic is wanted (parity mode) and as the checker for the
// tcgen05 split-bf16 GEMM in gemm_tc.cu. Replaces the tf.matmul / Dense / LSTMBlockCell matmuls at
// reference common/rnn.py:124, generators/rnn_nade.py:54-57,212, generators/rnn_rbm.py:252-253.
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

constexpr int BM = 128, BN = 128, BK = 8, GT = 256;

struct GemmArgs {
  const float* A; long long lda; int ta;  // ta=0: A is [M,K] row-major; ta=1: A is stored [K,M]
  const float* B; long long ldb; int tb;  // tb=0: B is [K,N] row-major; tb=1: B is stored [N,K]
  float* C; long long ldc;
  const float* bias;
  float alpha, beta;
  int M, N, K;
  int kchunk;  // K range per blockIdx.z (split-K; atomics when gridDim.z > 1)
};

__device__ __forceinline__ float ld_guard(const float* p, bool ok) { return ok ? __ldg(p) : 0.f; }

__global__ void __launch_bounds__(GT) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.kchunk;
  const int kend = min(g.K, kbeg + g.kchunk);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8x8 outputs each

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int k0) {
    // A tile: BM x BK
    if (!g.ta) {
      const int r = tid >> 1, c = (tid & 1) * 4;
      const int gm = m0 + r;
      const float* src = g.A + (size_t)gm * g.lda + k0 + c;
#pragma unroll
      for (int q = 0; q < 4; ++q) ra[q] = ld_guard(src + q, gm < g.M && k0 + c + q < kend);
    } else {
      const int kk = tid >> 5, c = (tid & 31) * 4;
      const float* src = g.A + (size_t)(k0 + kk) * g.lda + m0 + c;
#pragma unroll
      for (int q = 0; q < 4; ++q) ra[q] = ld_guard(src + q, k0 + kk < kend && m0 + c + q < g.M);
    }
    if (!g.tb) {
      const int kk = tid >> 5, c = (tid & 31) * 4;
      const float* src = g.B + (size_t)(k0 + kk) * g.ldb + n0 + c;
#pragma unroll
      for (int q = 0; q < 4; ++q) rb[q] = ld_guard(src + q, k0 + kk < kend && n0 + c + q < g.N);
    } else {
      const int r = tid >> 1, c = (tid & 1) * 4;
      const int gn = n0 + r;
      const float* src = g.B + (size_t)gn * g.ldb + k0 + c;
#pragma unroll
      for (int q = 0; q < 4; ++q) rb[q] = ld_guard(src + q, gn < g.N && k0 + c + q < kend);
    }
  };
  auto sstore = [&](int buf) {
    if (!g.ta) {
      const int r = tid >> 1, c = (tid & 1) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) As[buf][c + q][r] = ra[q];
    } else {
      const int kk = tid >> 5, c = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][kk][c]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    }
    if (!g.tb) {
      const int kk = tid >> 5, c = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][c]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
      const int r = tid >> 1, c = (tid & 1) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[buf][c + q][r] = rb[q];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  const bool atomic = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= g.N) continue;
      float v = g.alpha * acc[i][j];
      float* dst = g.C + (size_t)gm * g.ldc + gn;
      if (atomic) {
        if (g.bias && blockIdx.z == 0) v += g.bias[gn];
        atomicAdd(dst, v);
      } else {
        if (g.bias) v += g.bias[gn];
        if (g.beta != 0.f) v += g.beta * *dst;
        *dst = v;
      }
    }
  }
}

__global__ void scale_rows_kernel(float* C, long long ldc, int M, int N, float beta) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  const int r = idx / N, c = idx % N;
  float* p = C + (size_t)r * ldc + c;
  *p = beta == 0.f ? 0.f : *p * beta;
}

}  // namespace mnn

using namespace mnn;

extern "C" int mnn_gemm_f32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                            float* C, long long ldc, const float* bias, float alpha, float beta, int M, int N, int K,
                            cudaStream_t stream) {
  MNN_REQUIRE(A && B && C, MNN_ERR_ARG, "gemm_f32: null pointer");
  MNN_REQUIRE(M > 0 && N > 0 && K > 0, MNN_ERR_ARG, "gemm_f32: non-positive size");
  GemmArgs g{A, lda, transA, B, ldb, transB, C, ldc, bias, alpha, beta, M, N, K, K};
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  // split-K when the output grid cannot fill the chip and K is long (weight-gradient GEMMs)
  const int tiles = grid.x * grid.y;
  int splits = 1;
  if (tiles < 148 && K >= 4096) {
    splits = (2 * 148 + tiles - 1) / tiles;
    const int maxs = K / 1024;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    int kchunk = (K + splits - 1) / splits;
    kchunk = (kchunk + BK - 1) / BK * BK;
    splits = (K + kchunk - 1) / kchunk;
    g.kchunk = kchunk;
    grid.z = splits;
    if (beta != 1.f) {
      const size_t n = (size_t)M * N;
      scale_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(C, ldc, M, N, beta);
    }
  }
  gemm_f32_kernel<<<grid, GT, 0, stream>>>(g);
  return mnn_check_launch("gemm_f32", (splits > 1 && beta != 1.f) ? 2 : 1);
}
