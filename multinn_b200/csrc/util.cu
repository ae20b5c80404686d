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
