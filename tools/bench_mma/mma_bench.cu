// Microbenchmark: throughput of back-to-back tcgen05.mma (kind::f16, bf16 in, fp32 accumulate, cta_group::1, M=128)
// for SWIZZLE_NONE vs SWIZZLE_128B shared-memory operand layouts and several N.  Data content is irrelevant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vame_b200/csrc/common.cuh"
using namespace vb;

__device__ __forceinline__ uint64_t desc_generic(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int N>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int iters, int mode, int same_k, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: 128 rows x 256 k bf16 = 64 KB, B: 256 rows x 256 k = 128 KB (zero-filled; only timing matters)
  for (int i = threadIdx.x; i < (64 + 128) * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
      const uint32_t idesc = make_idesc_bf16(128, N);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int ks = same_k ? 0 : (i & 15);
        uint64_t da, db;
        if (mode == 0) {          // SWIZZLE_NONE: atoms 128 B, LBO 128, SBO 1024 (tile = 64 k wide: chunk stride 16/32 KB)
          const uint32_t ko = (ks & 3) * 256, kc = ks >> 2;
          da = desc_generic(sa + kc * 16384 + ko, 128, 1024, 0);
          db = desc_generic(sb + kc * 32768 + ko, 128, 1024, 0);
        } else {                  // SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart, K advance = 32 B
          const uint32_t ko = (ks & 3) * 32, kc = ks >> 2;
          da = desc_generic(sa + kc * 16384 + ko, 16, 1024, 2);
          db = desc_generic(sb + kc * 32768 + ko, 16, 1024, 2);
        }
        umma_bf16(tmem + (uint32_t)((i % nacc) * N), da, db, idesc, i >= nacc);
      }
      t1 = clock64();
      umma_commit(&done);
      mbar_wait(&done, 0);
      t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
void run(int mode, int same_k, long long* d_out, int nacc = 1) {
  const int iters = 512, smem = (64 + 128) * 1024;
  cudaFuncSetAttribute(mma_bench_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    mma_bench_kernel<N><<<1, 128, smem>>>(iters, mode, same_k, nacc, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d mode=%d: %s\n", N, mode, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("M=128 N=%3d K=16 %-13s %s nacc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (ideal %d)\n", N, mode ? "SWIZZLE_128B" : "SWIZZLE_NONE",
         same_k ? "same-k " : "k-sweep", nacc, (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int mode = 0; mode < 2; ++mode) {
    run<48>(mode, 0, d_out);
    run<96>(mode, 0, d_out);
    run<128>(mode, 0, d_out);
    run<192>(mode, 0, d_out);
    run<256>(mode, 0, d_out);
  }
  // independent accumulators (round-robin over nacc TMEM regions): is the ~100-cycle floor a dependency latency?
  for (int nacc = 2; nacc <= 4; nacc *= 2) {
    run<48>(0, 0, d_out, nacc);
    run<96>(0, 0, d_out, nacc);
    run<128>(0, 0, d_out, nacc);
    if (nacc == 2) run<192>(0, 0, d_out, nacc);
  }
  return 0;
}
