// K0: input staging. Restates the pad / unstack / stack / shift plumbing of reference
// core/multi_encoder_nn.py:66-87 and multinn_composer.py:73-87 (Composer) / multinn_jamming.py:60-68
// (per-track) as one pass over x[B,T,D,M]:
//   xin  [(T+1)][B][D*M]  time-major, slot 0 = zero frame, slot t+1 = x[:,t] (feature = d*M + m)
//        -> generator inputs = slots 0..T-1, generator targets = slots 1..T
//   xtr  [M][(T+1)][B][D] per-track time-major copy (Jamming / Feedback), optional
//   bits [M][T*B][4]      target bit masks, row n' = t*B + b, bit d of track m = x[b,t,d,m] != 0
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

template <typename TIn>
__global__ void pack_kernel(const TIn* __restrict__ x, float* __restrict__ xin, float* __restrict__ xtr,
                            uint32_t* __restrict__ bits, int B, int T, int D, int M) {
  extern __shared__ uint32_t msk[];  // [warps][M][4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t* mw = msk + (size_t)warp * M * 4;
  const int I = D * M;
  const long long rows = (long long)B * T;
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
    const int b = (int)(row / T), t = (int)(row % T);
    for (int i = lane; i < M * 4; i += 32) mw[i] = 0u;
    __syncwarp();
    const TIn* src = x + (size_t)row * I;
    float* dst = xin ? xin + ((size_t)(t + 1) * B + b) * I : nullptr;
    for (int e = lane; e < I; e += 32) {
      const float v = (float)__ldg(src + e);
      if (dst) dst[e] = v;
      const int d = e / M, m = e - d * M;
      if (xtr) xtr[(((size_t)m * (T + 1) + (t + 1)) * B + b) * D + d] = v;
      if (v != 0.f) atomicOr(&mw[m * 4 + (d >> 5)], 1u << (d & 31));
    }
    __syncwarp();
    if (bits)
      for (int i = lane; i < M * 4; i += 32) {
        const int m = i >> 2;
        bits[((size_t)m * rows + (size_t)t * B + b) * 4 + (i & 3)] = mw[i];
      }
    __syncwarp();
  }
}

// bits for an already flattened per-track target matrix v[N][D] (ld floats per row) -> bits[N][4]
__global__ void pack_rows_kernel(const float* __restrict__ v, long long ld, int dim_stride, uint32_t* __restrict__ bits,
                                 int N, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int d = w * 32 + lane;
    const bool on = d < D && __ldg(v + (size_t)row * ld + (size_t)d * dim_stride) != 0.f;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[(size_t)row * 4 + w] = word;
  }
}

}  // namespace mnn

using namespace mnn;

template <typename TIn>
static int pack_impl(const TIn* x, float* xin, float* xtr, uint32_t* bits, int B, int T, int D, int M, cudaStream_t stream) {
  MNN_REQUIRE(x && (xin || xtr || bits), MNN_ERR_ARG, "pack_pianoroll: null pointer");
  MNN_REQUIRE(B > 0 && T > 0 && D > 0 && M > 0, MNN_ERR_ARG, "pack_pianoroll: non-positive size");
  MNN_REQUIRE(D <= 128, MNN_ERR_UNSUPPORTED, "pack_pianoroll: num_dims > 128 not instantiated");
  // zero frames (slot 0)
  if (xin) cudaMemsetAsync(xin, 0, (size_t)B * D * M * sizeof(float), stream);
  if (xtr)
    for (int m = 0; m < M; ++m)
      cudaMemsetAsync(xtr + (size_t)m * (T + 1) * B * D, 0, (size_t)B * D * sizeof(float), stream);
  const int warps = 8;
  const long long rows = (long long)B * T;
  long long grid = (rows + warps - 1) / warps;
  if (grid > 148 * 8) grid = 148 * 8;
  pack_kernel<TIn><<<(unsigned)grid, warps * 32, (size_t)warps * M * 4 * sizeof(uint32_t), stream>>>(x, xin, xtr, bits, B,
                                                                                                 T, D, M);
  return mnn_check_launch("pack_pianoroll");
}

extern "C" int mnn_pack_pianoroll(const float* x, float* xin, float* xtr, uint32_t* bits, int B, int T, int D, int M,
                                  cudaStream_t stream) {
  return pack_impl<float>(x, xin, xtr, bits, B, T, D, M, stream);
}

// Same for piano-rolls kept as bytes (the reference's .npy files are bool arrays, prepare_data.py:56; they are fed to the
// float32 placeholder as they are): 4x less host->device traffic.
extern "C" int mnn_pack_pianoroll_u8(const uint8_t* x, float* xin, float* xtr, uint32_t* bits, int B, int T, int D, int M,
                                     cudaStream_t stream) {
  return pack_impl<uint8_t>(x, xin, xtr, bits, B, T, D, M, stream);
}

extern "C" int mnn_pack_rows(const float* v, long long ld, int dim_stride, uint32_t* bits, int N, int D,
                             cudaStream_t stream) {
  MNN_REQUIRE(v && bits && N > 0 && D > 0, MNN_ERR_ARG, "pack_rows: bad argument");
  MNN_REQUIRE(D <= 128, MNN_ERR_UNSUPPORTED, "pack_rows: num_dims > 128 not instantiated");
  pack_rows_kernel<<<(N + 7) / 8, 256, 0, stream>>>(v, ld, dim_stride, bits, N, D);
  return mnn_check_launch("pack_rows");
}
