// Shared device helpers for libmultinn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MNN_OK 0
#define MNN_ERR_ARG (-1)          // bad argument (null pointer, non-positive size)
#define MNN_ERR_UNSUPPORTED (-2)  // shape outside what the sm_100a kernels are instantiated for
#define MNN_ERR_WORKSPACE (-3)

void mnn_set_error(const char* msg);
int mnn_check_launch(const char* what, int kernels = 1);  // also counts kernel launches

#define MNN_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) {                   \
      mnn_set_error(msg);            \
      return (code);                 \
    }                                \
  } while (0)

namespace mnn {

// Data-parallel row keying of the in-kernel Philox streams: a kernel that draws noise for its LOCAL row r
// keys the counter by the GLOBAL row  (r / rows_local) * rows_global + row_base + r % rows_local  -- local rows come in
// groups of rows_local (the per-GPU batch of one time step of a time-major tensor) that sit rows_global apart in the
// global tensor, this rank's group starting at row_base. rows_local == 0: identity. Set per host thread with
// mnn_set_row_map(); every Philox-capable launch copies the calling thread's map into its kernel arguments.
struct RowMap {
  long long rows_local, rows_global, row_base;
  long long t_base;   // index of a sequence launch's first time step inside the whole sequence (chunked launches)
};
__host__ __device__ __forceinline__ unsigned long long global_row(const RowMap& m, unsigned long long r) {
  if (m.rows_local <= 0) return r;
  const unsigned long long g = r / (unsigned long long)m.rows_local;
  return g * (unsigned long long)m.rows_global + (unsigned long long)m.row_base + (r - g * (unsigned long long)m.rows_local);
}
RowMap current_row_map();   // host: the calling thread's map (util.cu)

constexpr float kSafeLogEps = 1e-6f;  // reference utils/auxiliary.py:11

__device__ __forceinline__ float sigmoid_fast(float x) {
  // ex2.approx + rcp.approx: ~2 ulp, 2 MUFU ops.
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  // tanh via exp: accurate to ~2 ulp over the LSTM's range; tanhf would also do.
  return tanhf(x);
}

// sigmoid / tanh straight from MUFU.EX2 + MUFU.RCP (4-5 instructions, absolute error ~1e-7; ex2.approx.ftz underflows to
// 0 and overflows to +inf cleanly, so the saturated ends come out as exactly 0 / 1 / -1)
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_mufu(float x) {   // 2 * sigmoid(2x) - 1
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(2.0f, r, -1.0f);
}

// tanh from the same two MUFU ops: 1 - 2 / (1 + e^{2x}) away from 0 (absolute error ~1e-7), odd Taylor polynomial
// near 0 where that form cancels (relative error < 1e-7 for |x| < 0.04).
__device__ __forceinline__ float tanh_fast(float x) {
  const float x2 = x * x;
  const float poly = x * fmaf(x2, fmaf(x2, 0.13333333f, -0.33333334f), 1.0f);
  const float big = fmaf(-2.0f, __fdividef(1.0f, 1.0f + __expf(2.0f * x)), 1.0f);
  return fabsf(x) < 0.04f ? poly : big;
}
// One LSTM cell from the four gate pre-activations with 5 MUFU.EX2 + 2 MUFU.RCP instead of 5 + 5. The XU pipe retires
// rcp.approx at a third of the ex2.approx rate (mnn_probe_mufu: 5.4 against 15.8 per clock per SM), and the cell epilogue
// of the recurrence kernels is XU-bound, so the four gate reciprocals share ONE rcp (Montgomery's trick:
// 1/a = b*c*d * rcp(a*b*c*d)); the exponentials are clamped to 1e9 so that the product of four denominators stays finite
// (a gate below 1e-9 comes out as 1e-9: absolute error < 1e-9). tanh keeps the odd polynomial near 0 (tanh_fast).
// c' = tanh(j) * sig(i) + c * sig(f);  h' = tanh(c') * sig(o)   (common/rnn.py:124)
__device__ __forceinline__ float ex2_clamped(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return fminf(e, 1.0e9f);
}
__device__ __forceinline__ float tanh_from_rcp(float x, float r) {   // r = 1 / (1 + e^{-2x})
  const float x2 = x * x;
  const float poly = x * fmaf(x2, fmaf(x2, 0.13333333f, -0.33333334f), 1.0f);
  return fabsf(x) < 0.04f ? poly : fmaf(2.0f, r, -1.0f);
}
__device__ __forceinline__ void lstm_cell_xu(float pi, float pj, float pf, float po, float cprev, float& gi, float& gj,
                                             float& gf, float& go, float& c, float& h) {
  const float di = 1.0f + ex2_clamped(pi * -1.4426950408889634f);
  const float dj = 1.0f + ex2_clamped(pj * -2.8853900817779268f);
  const float df = 1.0f + ex2_clamped(pf * -1.4426950408889634f);
  const float dO = 1.0f + ex2_clamped(po * -1.4426950408889634f);
  const float pij = di * dj, pfo = df * dO;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pij * pfo));
  const float rij = r * pfo, rfo = r * pij;          // 1 / (di dj), 1 / (df do)
  gi = rij * dj;
  gj = tanh_from_rcp(pj, rij * di);
  gf = rfo * dO;
  go = rfo * df;
  c = fmaf(gj, gi, cprev * gf);
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex2_clamped(c * -2.8853900817779268f)));
  h = tanh_from_rcp(c, rc) * go;
}
// Four sigmoids with 4 MUFU.EX2 + ONE MUFU.RCP (see lstm_cell_xu: rcp.approx retires at a third of the ex2 rate).
__device__ __forceinline__ float4 sigmoid4_xu(float4 a) {
  const float dx = 1.0f + ex2_clamped(a.x * -1.4426950408889634f), dy = 1.0f + ex2_clamped(a.y * -1.4426950408889634f);
  const float dz = 1.0f + ex2_clamped(a.z * -1.4426950408889634f), dw = 1.0f + ex2_clamped(a.w * -1.4426950408889634f);
  const float pxy = dx * dy, pzw = dz * dw;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pxy * pzw));
  const float rxy = r * pzw, rzw = r * pxy;
  return make_float4(rxy * dy, rxy * dx, rzw * dw, rzw * dz);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Packed fp32x2 FMA (Blackwell FFMA2): d = a*b + c on both halves.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ua, ub, uc, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

// Packed fp32x2 add / sub / mul (Blackwell FADD2 / FMUL2): both halves in one issue slot.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011). One call -> 4 x 32 random bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// Dropout noise of the LSTM kernels (SIMT cell, tcgen05 sequence kernels, any chunking, any GPU count draw the SAME mask):
// the 4 uniforms of units [4q, 4q+4) of batch row b at time step t are the 4 words of
// Philox(ctr = ((global b) * R/4 + q as 64 bits, global t, 0), key = seed).
__device__ __forceinline__ uint4 dropout_bits4(unsigned long long seed, const RowMap& m, int b, int unit4, int R, int t) {
  const unsigned long long e = global_row(m, (unsigned long long)b) * (unsigned long long)(R >> 2) + (unsigned)(unit4 >> 2);
  return philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)(t + (int)m.t_base), 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// 24-bit uniform in [0,1): exactly representable in fp32, strict-< comparisons behave like TF's.
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace mnn
