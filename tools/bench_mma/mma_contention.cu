// Microbenchmark: does activity of the other warps of the CTA slow down a chain of tcgen05.mma that streams its A operand
// from shared memory?  Shape of the resident-weight GRU sweep (gru_rw.cu): 48 MMAs (M=128, N=32, K=16) per "step" whose A
// operand is 192 KB of shared memory (12 tiles of 128 rows x 64 k), B = 4 chunks of [32 rows x 64 k].
//   mode 0: the 4 other warps sleep in a hardware barrier (bar.sync) while the MMAs run
//   mode 1: all 128 threads poll the completion mbarrier with mbarrier.try_wait (what the sweep kernels do)
//   mode 2: one lane per warp polls, the others are parked at __syncwarp
//   mode 3: all threads poll with mbarrier.test_wait + nanosleep(20) back-off
//   mode 4: the 4 warps stream 16-byte global stores (the sweep's saved-gate stores) and then poll as in mode 1
//   mode 5: like mode 2, and the polling lane uses test_wait + nanosleep(20)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vame_b200/csrc/common.cuh"
using namespace vb;

constexpr int ATILE = 128 * 64 * 2, BTILE = 32 * 64 * 2;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(160, 1) contention_kernel(int steps, int mode, float* sink, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_issue0[64], t_issue1[64], t_seen[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (12 * ATILE + 4 * BTILE) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(&tmem_slot, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, 32);
  for (int s = 0; s < steps; ++s) {
    const uint32_t ph = s & 1;
    asm volatile("bar.sync 2, 160;" ::: "memory");          // common start of the step
    if (warp == 4) {
      if (lane == 0) {
        const uint64_t dA = make_desc(smem_u32(smem)), dB = make_desc(smem_u32(smem) + 12 * ATILE);
        const long long t0 = clock64();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ao = (g * 4 + kc) * ATILE + ks * 2 * ATOM_BYTES, bo = kc * BTILE + ks * 2 * ATOM_BYTES;
              if (kc == 0 && ks == 0) umma_bf16_c<0>(tmem + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
              else umma_bf16_c<1>(tmem + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
            }
          }
        }
        umma_commit(&done);
        const long long t1 = clock64();
        t_issue0[s] = t0;
        t_issue1[s] = t1;
        if (mode == 0) {
          while (!mbar_try_wait(&done, ph)) {}
          t_seen[s] = clock64();
        }
      }
      __syncwarp();
      if (mode == 0) asm volatile("bar.sync 1, 160;" ::: "memory");
    } else {
      if (mode == 0) {
        asm volatile("bar.sync 1, 160;" ::: "memory");
      } else {
        if (mode == 4) {
          float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            reinterpret_cast<float4*>(sink)[((size_t)(blockIdx.x * 64 + s * 16 + i) * 128 + threadIdx.x)] = v;
        }
        if (mode == 1 || mode == 4) {
          while (!mbar_try_wait(&done, ph)) {}
        } else if (mode == 2) {
          if (lane == 0) while (!mbar_try_wait(&done, ph)) {}
          __syncwarp();
        } else if (mode == 3) {
          while (!mbar_test_wait(&done, ph)) __nanosleep(20);
        } else if (mode == 5) {
          if (lane == 0) while (!mbar_test_wait(&done, ph)) __nanosleep(20);
          __syncwarp();
        }
        if (threadIdx.x == 0) t_seen[s] = clock64();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long issue = 0, total = 0;
    for (int s = 2; s < steps; ++s) {
      issue += t_issue1[s] - t_issue0[s];
      total += t_seen[s] - t_issue0[s];
    }
    out[0] = issue / (steps - 2);
    out[1] = total / (steps - 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

int main() {
  long long* d_out;
  float* sink;
  cudaMalloc(&d_out, 16);
  cudaMalloc(&sink, (size_t)148 * 64 * 128 * 16);
  const int smem = 12 * ATILE + 4 * BTILE;
  cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"others in bar.sync", "128 threads try_wait", "1 lane/warp try_wait", "128 threads test_wait+nanosleep(20)",
                         "global stores, then 128 threads try_wait", "1 lane/warp test_wait+nanosleep(20)"};
  for (int grid = 1; grid <= 128; grid *= 128)
    for (int mode = 0; mode < 6; ++mode) {
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        contention_kernel<<<grid, 160, smem>>>(32, mode, sink, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("grid %3d mode %d (%s): 48 MMAs issued in %lld cycles (%.1f / MMA), completion seen after %lld cycles (%.1f / MMA)\n", grid, mode,
             names[mode], h[0], h[0] / 48.0, h[1], h[1] / 48.0);
    }
  return 0;
}
