// LSTM temporal unit (K2), time-major. Semantics: CudnnCompatibleLSTMCell == LSTMBlockCell with
// forget_bias 0, gate column blocks i, j (cell input), f, o, wrapped in DropoutWrapper(output_keep_prob)
// and MultiRNNCell (reference common/rnn.py:104-145).
//   c' = tanh(j) * sigmoid(i) + c * sigmoid(f);  h' = tanh(c') * sigmoid(o);  out = h'/keep * floor(keep + u)
// Sequence driver: per step, gates[t] += h[t-1] . Wh (GEMM) then the fused cell kernel below.
#include "common.cuh"
#include "multinn_b200.h"

namespace mnn {

struct CellFwdArgs {
  float* gates;        // [B][4R] in: pre-activations, out: activations (sig i, tanh j, sig f, sig o)
  const float* c_prev; // [B][R]
  float* c;            // [B][R]
  float* h;            // [B][R]
  float* out;          // [B][R] dropout(h) (may alias nothing; null -> skip)
  float* dscale;       // [B][R] 0 or 1/keep (null when keep == 1)
  const float* u;      // [B][R] uniforms or null
  float keep;
  unsigned long long seed, offset;  // philox when u == null and keep < 1
  int B, R;
  RowMap rmap;                      // global-row keying of the Philox counter (offset must be a multiple of R)
};

__global__ void lstm_cell_fwd_kernel(CellFwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.R) return;
  const int b = idx / a.R, r = idx - b * a.R;
  float* g = a.gates + (size_t)b * 4 * a.R;
  const float gi = sigmoid_acc(g[r]);
  const float gj = tanh_acc(g[a.R + r]);
  const float gf = sigmoid_acc(g[2 * a.R + r]);
  const float go = sigmoid_acc(g[3 * a.R + r]);
  const float c = gj * gi + a.c_prev[idx] * gf;
  const float h = tanh_acc(c) * go;
  g[r] = gi; g[a.R + r] = gj; g[2 * a.R + r] = gf; g[3 * a.R + r] = go;
  a.c[idx] = c;
  a.h[idx] = h;
  if (a.out) {
    float o = h;
    if (a.keep < 1.0f) {
      float uu;
      if (a.u) uu = a.u[idx];
      else {
        // offset = t * B * R of the sequence loop (mnn_lstm_seq_fwd); the word of unit r is r % 4 of its group of 4
        const int t = (int)(a.offset / ((unsigned long long)a.B * a.R));
        const uint4 r4 = dropout_bits4(a.seed, a.rmap, b, r & ~3, a.R, t);
        const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
        uu = u01(w4[r & 3]);
      }
      // tf.nn.dropout: x / keep * floor(keep + u)
      const float keepmask = floorf(a.keep + uu);
      o = h / a.keep * keepmask;
      a.dscale[idx] = keepmask / a.keep;
    }
    a.out[idx] = o;
  }
}

struct CellBwdArgs {
  float* gates;          // [B][4R] in: activations, out: d pre-activations
  const float* c_prev;   // [B][R]
  const float* c;        // [B][R]
  const float* dout;     // [B][R] grad wrt dropout(h) output, or null
  const float* dscale;   // [B][R] or null
  const float* dh_rec;   // [B][R] grad wrt h from step t+1's recurrent GEMM (null at t = T-1)
  float* dc;             // [B][R] in: dc carried from t+1, out: dc for t-1
  int B, R;
};

__global__ void lstm_cell_bwd_kernel(CellBwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.R) return;
  const int b = idx / a.R, r = idx - b * a.R;
  float* g = a.gates + (size_t)b * 4 * a.R;
  const float gi = g[r], gj = g[a.R + r], gf = g[2 * a.R + r], go = g[3 * a.R + r];
  float dh = a.dh_rec ? a.dh_rec[idx] : 0.f;
  if (a.dout) dh += a.dscale ? a.dout[idx] * a.dscale[idx] : a.dout[idx];
  const float tc = tanh_acc(a.c[idx]);
  const float dc = a.dc[idx] + dh * go * (1.f - tc * tc);
  g[r] = dc * gj * gi * (1.f - gi);
  g[a.R + r] = dc * gi * (1.f - gj * gj);
  g[2 * a.R + r] = dc * a.c_prev[idx] * gf * (1.f - gf);
  g[3 * a.R + r] = dh * tc * go * (1.f - go);
  a.dc[idx] = dc * gf;
}

// out[c] (+)= sum_r A[r][c]. Two deterministic stages: grid (cols/32, RB) blocks of 32x8 threads, every thread reduces
// its column over every 8th row of the block's row range into ws[rb*8 + ty][c]; a second pass adds the 8*RB partials in
// a fixed order. HBM-bound: every element of A is read once. No shared memory: the blocks fit beside a resident
// tensor-core GEMM CTA, so the bias-gradient sums can run under the weight-gradient GEMMs on a side stream.
__global__ void colsum_partial_kernel(const float* __restrict__ A, long long ld, int rows, int cols, int rows_per_block,
                                      float* __restrict__ ws) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      s0 += __ldg(A + (size_t)r * ld + c);
      s1 += __ldg(A + (size_t)(r + 8) * ld + c);
      s2 += __ldg(A + (size_t)(r + 16) * ld + c);
      s3 += __ldg(A + (size_t)(r + 24) * ld + c);
    }
    for (; r < r1; r += 8) s0 += __ldg(A + (size_t)r * ld + c);
    ws[((size_t)blockIdx.y * 8 + threadIdx.y) * cols + c] = (s0 + s1) + (s2 + s3);
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ ws, int rb, int cols, float* __restrict__ out,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = 0.f;
  for (int i = 0; i < rb; ++i) t += ws[(size_t)i * cols + c];
  out[c] = accumulate ? out[c] + t : t;
}

}  // namespace mnn

using namespace mnn;

extern "C" int mnn_lstm_cell_fwd(float* gates, const float* c_prev, float* c, float* h, float* out, float* dscale,
                                 const float* u, float keep, unsigned long long seed, unsigned long long offset,
                                 int B, int R, cudaStream_t stream) {
  MNN_REQUIRE(gates && c_prev && c && h, MNN_ERR_ARG, "lstm_cell_fwd: null pointer");
  MNN_REQUIRE(B > 0 && R > 0, MNN_ERR_ARG, "lstm_cell_fwd: non-positive size");
  MNN_REQUIRE(!(out && keep < 1.f && !dscale), MNN_ERR_ARG, "lstm_cell_fwd: dscale required when keep < 1");
  CellFwdArgs a{gates, c_prev, c, h, out, dscale, u, keep, seed, offset, B, R, current_row_map()};
  const int n = B * R;
  lstm_cell_fwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a);
  return mnn_check_launch("lstm_cell_fwd");
}

extern "C" int mnn_lstm_seq_fwd(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale,
                                const float* u, float keep, unsigned long long seed, int T, int B, int R,
                                cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && hbuf && cbuf, MNN_ERR_ARG, "lstm_seq_fwd: null pointer");
  MNN_REQUIRE(T > 0 && B > 0 && R > 0, MNN_ERR_ARG, "lstm_seq_fwd: non-positive size");
  const size_t BR = (size_t)B * R;
  for (int t = 0; t < T; ++t) {
    float* g = gates + (size_t)t * B * 4 * R;
    // gates[t] += h[t-1] . Wh     (hbuf slot t holds h[t-1]; slot 0 is the initial state)
    int rc = mnn_gemm_f32(hbuf + (size_t)t * BR, R, 0, wh, 4 * R, 0, g, 4 * R, nullptr, 1.f, 1.f, B, 4 * R, R, stream);
    if (rc) return rc;
    rc = mnn_lstm_cell_fwd(g, cbuf + (size_t)t * BR, cbuf + (size_t)(t + 1) * BR, hbuf + (size_t)(t + 1) * BR,
                           out ? out + (size_t)t * BR : nullptr, (dscale && keep < 1.f) ? dscale + (size_t)t * BR : nullptr,
                           u ? u + (size_t)t * BR : nullptr, keep, seed, (unsigned long long)t * BR, B, R, stream);
    if (rc) return rc;
  }
  return MNN_OK;
}

extern "C" int mnn_lstm_seq_bwd(float* gates, const float* wh, const float* cbuf, const float* dout,
                                const float* dscale, float* dh_work, float* dc_work, int T, int B, int R,
                                cudaStream_t stream) {
  MNN_REQUIRE(gates && wh && cbuf && dh_work && dc_work, MNN_ERR_ARG, "lstm_seq_bwd: null pointer");
  MNN_REQUIRE(T > 0 && B > 0 && R > 0, MNN_ERR_ARG, "lstm_seq_bwd: non-positive size");
  const size_t BR = (size_t)B * R;
  cudaMemsetAsync(dc_work, 0, BR * sizeof(float), stream);
  const int n = B * R;
  for (int t = T - 1; t >= 0; --t) {
    float* g = gates + (size_t)t * B * 4 * R;
    CellBwdArgs a{g, cbuf + (size_t)t * BR, cbuf + (size_t)(t + 1) * BR, dout ? dout + (size_t)t * BR : nullptr,
                  dscale ? dscale + (size_t)t * BR : nullptr, t == T - 1 ? nullptr : dh_work, dc_work, B, R};
    lstm_cell_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a);
    int rc = mnn_check_launch("lstm_cell_bwd");
    if (rc) return rc;
    if (t > 0) {
      // dh_rec[t-1] = dG[t] . Wh^T
      rc = mnn_gemm_f32(g, 4 * R, 0, wh, 4 * R, 1, dh_work, R, nullptr, 1.f, 0.f, B, R, 4 * R, stream);
      if (rc) return rc;
    }
  }
  return MNN_OK;
}

extern "C" size_t mnn_colsum_workspace_bytes(int cols) { return (size_t)64 * 8 * (size_t)cols * sizeof(float); }

extern "C" int mnn_colsum(const float* A, long long ld, int rows, int cols, float* out, int accumulate, void* ws,
                          cudaStream_t stream) {
  MNN_REQUIRE(A && out && ws && rows > 0 && cols > 0, MNN_ERR_ARG, "colsum: bad argument");
  int rb = (rows + 255) / 256;
  if (rb > 64) rb = 64;
  int rpb = (rows + rb - 1) / rb;
  rpb = (rpb + 7) / 8 * 8;
  rb = (rows + rpb - 1) / rpb;
  colsum_partial_kernel<<<dim3((cols + 31) / 32, rb), dim3(32, 8), 0, stream>>>(A, ld, rows, cols, rpb,
                                                                              reinterpret_cast<float*>(ws));
  colsum_final_kernel<<<(cols + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const float*>(ws), rb * 8, cols, out, accumulate);
  return mnn_check_launch("colsum", 2);
}
