// K9: gradient global norm + clip + TF-style Adam / SGD on one flat fp32 parameter bucket.
// Follows reference utils/training.py:151-177 (clip_by_global_norm(5.) over the union of variables)
// and train.py:61-64 (tf.train.AdamOptimizer(lr, epsilon=1e-4) / GradientDescentOptimizer):
//   scale = clip / max(||g||, clip);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   theta -= lr_t * m / (sqrt(v) + eps),  lr_t = lr sqrt(1-b2^t)/(1-b1^t)   (eps on the uncorrected sqrt(v))
// Reductions are two-stage with a fixed order: deterministic run to run.
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

constexpr int kRedBlocks = 592;  // 4 x 148
constexpr int kRedThreads = 256;

template <bool SQ>
__global__ void reduce_stage1(const float* __restrict__ x, size_t n, double* __restrict__ partial) {
  __shared__ double red[kRedThreads / 32];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += SQ ? (double)v * (double)v : (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kRedThreads / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void reduce_stage2(const double* __restrict__ partial, int nb, float* __restrict__ out, float scale,
                              int accumulate) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float r = (float)(t * (double)scale);
    out[0] = accumulate ? out[0] + r : r;
  }
}

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, const float* __restrict__ sqnorm, float gscale,
                                 float clip, float lr_t, float b1, float b2, float eps) {
  const float gn = sqrtf(sqnorm[0]) * gscale;
  const float s = gscale * (clip > 0.f ? clip / fmaxf(gn, clip) : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * s;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void clip_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n,
                                const float* __restrict__ sqnorm, float gscale, float clip, float lr) {
  const float gn = sqrtf(sqnorm[0]) * gscale;
  const float s = gscale * (clip > 0.f ? clip / fmaxf(gn, clip) : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] -= lr * g[i] * s;
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

}  // namespace mnn

using namespace mnn;

// y += alpha * x  (CD-k assign_add of reference common/rbm.py:322-330; gradient accumulation)
extern "C" int mnn_axpy(float* y, const float* x, float alpha, size_t n, cudaStream_t stream) {
  MNN_REQUIRE(y && x && n > 0, MNN_ERR_ARG, "axpy: bad argument");
  int nb = (int)((n + 255) / 256);
  if (nb > 148 * 8) nb = 148 * 8;
  axpy_kernel<<<nb, 256, 0, stream>>>(y, x, alpha, n);
  return mnn_check_launch("axpy");
}

extern "C" size_t mnn_reduce_workspace_bytes(void) { return kRedBlocks * sizeof(double); }

static int reduce_impl(const float* x, size_t n, void* ws, float* out, float scale, int accumulate, bool sq,
                       cudaStream_t stream) {
  MNN_REQUIRE(x && ws && out && n > 0, MNN_ERR_ARG, "reduce: bad argument");
  int nb = (int)((n + kRedThreads - 1) / kRedThreads);
  if (nb > kRedBlocks) nb = kRedBlocks;
  double* partial = reinterpret_cast<double*>(ws);
  if (sq) reduce_stage1<true><<<nb, kRedThreads, 0, stream>>>(x, n, partial);
  else reduce_stage1<false><<<nb, kRedThreads, 0, stream>>>(x, n, partial);
  reduce_stage2<<<1, 256, 0, stream>>>(partial, nb, out, scale, accumulate);
  return mnn_check_launch("reduce", 2);
}

extern "C" int mnn_sum(const float* x, size_t n, void* ws, float* out, float scale, int accumulate,
                       cudaStream_t stream) {
  return reduce_impl(x, n, ws, out, scale, accumulate, false, stream);
}

extern "C" int mnn_sqnorm(const float* x, size_t n, void* ws, float* out, cudaStream_t stream) {
  return reduce_impl(x, n, ws, out, 1.f, 0, true, stream);
}

extern "C" int mnn_clip_adam(float* p, const float* g, float* m, float* v, size_t n, const float* sqnorm,
                             float grad_scale, float clip_norm, float lr, float beta1, float beta2, float eps,
                             int step, cudaStream_t stream) {
  MNN_REQUIRE(p && g && m && v && sqnorm && n > 0 && step > 0, MNN_ERR_ARG, "clip_adam: bad argument");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  int nb = (int)((n + 255) / 256);
  if (nb > 148 * 8) nb = 148 * 8;
  clip_adam_kernel<<<nb, 256, 0, stream>>>(p, g, m, v, n, sqnorm, grad_scale, clip_norm, (float)lr_t, beta1, beta2, eps);
  return mnn_check_launch("clip_adam");
}

extern "C" int mnn_clip_sgd(float* p, const float* g, size_t n, const float* sqnorm, float grad_scale,
                            float clip_norm, float lr, cudaStream_t stream) {
  MNN_REQUIRE(p && g && sqnorm && n > 0, MNN_ERR_ARG, "clip_sgd: bad argument");
  int nb = (int)((n + 255) / 256);
  if (nb > 148 * 8) nb = 148 * 8;
  clip_sgd_kernel<<<nb, 256, 0, stream>>>(p, g, n, sqnorm, grad_scale, clip_norm, lr);
  return mnn_check_launch("clip_sgd");
}


// ------------------------------------------------------------------------------------------------ padded rows
// x[r*ld + c] *= w[r % period] for c < ncols: drops the rows past a sequence's length from per-row results and
// gradients (flatten_maybe_padded_sequences keeps only rows t < lengths[b], reference utils/sequences.py:6-37).
namespace mnn {
__global__ void scale_rows_kernel(float* x, long long ld, int ncols, const float* w, long long rows, int period) {
  const long long total = rows * ncols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ncols;
    const int c = (int)(i - r * ncols);
    x[r * ld + c] *= __ldg(w + (r % period));
  }
}
}  // namespace mnn

extern "C" int mnn_scale_rows(float* x, long long ld, int ncols, const float* w, long long rows, int period,
                              cudaStream_t stream) {
  MNN_REQUIRE(x && w, MNN_ERR_ARG, "scale_rows: null pointer");
  MNN_REQUIRE(rows > 0 && ncols > 0 && period > 0, MNN_ERR_ARG, "scale_rows: non-positive size");
  const long long total = rows * ncols;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  mnn::scale_rows_kernel<<<grid, 256, 0, stream>>>(x, ld, ncols, w, rows, period);
  return mnn_check_launch("scale_rows");
}
