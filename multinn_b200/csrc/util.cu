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
