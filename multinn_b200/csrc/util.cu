// Library-level plumbing: version, last-error string, launch checks. No global mutable state
// beyond the thread-local error message.
#include <atomic>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "multinn_b200.h"

static thread_local char g_err[512] = "";

void mnn_set_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

static std::atomic<unsigned long long> g_launches{0};

int mnn_check_launch(const char* what, int kernels) {
  g_launches.fetch_add((unsigned long long)kernels, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return MNN_OK;
}

extern "C" int mnn_version(void) { return 100; }
extern "C" unsigned long long mnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* mnn_last_error_string(void) { return g_err; }

static thread_local mnn::RowMap g_row_map{0, 0, 0, 0};
mnn::RowMap mnn::current_row_map() { return g_row_map; }
extern "C" int mnn_set_row_map(long long rows_local, long long rows_global, long long row_base) {
  MNN_REQUIRE(rows_local >= 0 && row_base >= 0, MNN_ERR_ARG, "set_row_map: negative sizes");
  MNN_REQUIRE(rows_local == 0 || rows_global >= rows_local + 0, MNN_ERR_ARG, "set_row_map: rows_global < rows_local");
  MNN_REQUIRE(rows_local == 0 || row_base + rows_local <= rows_global, MNN_ERR_ARG,
              "set_row_map: the local group does not fit into the global one");
  g_row_map = mnn::RowMap{rows_local, rows_global, row_base, g_row_map.t_base};
  return MNN_OK;
}
extern "C" int mnn_set_time_base(long long t_base) {
  MNN_REQUIRE(t_base >= 0, MNN_ERR_ARG, "set_time_base: negative");
  g_row_map.t_base = t_base;
  return MNN_OK;
}

// ------------------------------------------------------------------------------------------------ roofline probes
// MUFU (XU pipe) peak: the NADE kernels are bounded by sigmoids = ex2.approx + rcp.approx, and MEASURED_PEAKS.json has no
// figure for that pipe (SURVEY 7 asks for one). Each thread runs `iters` rounds of 8 independent ex2.approx chains
// (kind 0), rcp.approx chains (kind 1) or full sigmoid_mufu chains (kind 2); out[block] keeps the compiler honest.
// ops = grid * block * iters * 8 MUFU instructions (x2 for kind 2).
__global__ void mufu_probe_kernel(float* out, int iters, int kind, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + 0.001f * (threadIdx.x + 32 * i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kind == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (kind == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else x[i] = mnn::sigmoid_mufu(x[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) out[blockIdx.x] = s;
}
extern "C" int mnn_probe_mufu(float* out, int blocks, int threads, int iters, int kind, cudaStream_t stream) {
  MNN_REQUIRE(out && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && kind >= 0 && kind <= 2, MNN_ERR_ARG,
              "probe_mufu: bad argument");
  mufu_probe_kernel<<<blocks, threads, 0, stream>>>(out, iters, kind, 0.25f);
  return mnn_check_launch("probe_mufu");
}
