// Shared helpers for the extern "C" entry points (error string, stream cast, argument checks).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace vb {
extern thread_local char g_err[512];
inline int fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return -2;
  }
  return 0;
}
}  // namespace vb
#define VB_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) return vb::fail(msg); \
  } while (0)
