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

// A operand from tensor memory (K = 16 bf16 = 8 columns per k-step)
template <int ACC>
__device__ __forceinline__ void umma_bf16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "n"(ACC)
      : "memory");
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
      // tight issue loop: descriptors precomputed, 4 k-steps unrolled, constant predicates (like the production kernels)
      uint64_t da[4], db[4];
      for (int ks = 0; ks < 4; ++ks) {
        if (mode == 0) { da[ks] = desc_generic(sa + ks * 256, 128, 1024, 0); db[ks] = desc_generic(sb + ks * 256, 128, 1024, 0); }
        else           { da[ks] = desc_generic(sa + ks * 32, 16, 1024, 2);   db[ks] = desc_generic(sb + ks * 32, 16, 1024, 2); }
      }
      const uint32_t d1 = tmem + (nacc > 1 ? N : 0);
      t0 = clock64();
      if (mode == 2) {                      // A from TMEM columns [256, 256 + 32), B no-swizzle from shared memory
        const uint32_t ta = tmem + 256;
        for (int ks = 0; ks < 4; ++ks) db[ks] = desc_generic(sb + ks * 256, 128, 1024, 0);
        umma_bf16_ta<0>(tmem, ta, db[0], idesc);
        umma_bf16_ta<0>(d1, ta, db[0], idesc);
        for (int i = 0; i < iters; i += 4) {
          umma_bf16_ta<1>(tmem, ta, db[0], idesc);
          umma_bf16_ta<1>(d1, ta + 8, db[1], idesc);
          umma_bf16_ta<1>(tmem, ta + 16, db[2], idesc);
          umma_bf16_ta<1>(d1, ta + 24, db[3], idesc);
        }
      } else {
      umma_bf16_c<0>(tmem, da[0], db[0], idesc);
      umma_bf16_c<0>(d1, da[0], db[0], idesc);
      for (int i = 0; i < iters; i += 4) {
        umma_bf16_c<1>(tmem, da[0], db[0], idesc);
        umma_bf16_c<1>(d1, da[1], db[1], idesc);
        umma_bf16_c<1>(tmem, da[2], db[2], idesc);
        umma_bf16_c<1>(d1, da[3], db[3], idesc);
      }
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
  printf("M=128 N=%3d K=16 %-13s %s nacc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (ideal %d)\n", N, mode == 2 ? "A-from-TMEM" : mode ? "SWIZZLE_128B" : "SWIZZLE_NONE",
         same_k ? "same-k " : "k-sweep", nacc, (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int nacc = 1; nacc <= 2; ++nacc)
    for (int mode = 0; mode < 3; ++mode) {
      run<16>(mode, 0, d_out, nacc);
      run<32>(mode, 0, d_out, nacc);
      run<64>(mode, 0, d_out, nacc);
      run<48>(mode, 0, d_out, nacc);
      run<96>(mode, 0, d_out, nacc);
      run<128>(mode, 0, d_out, nacc);
      run<192>(mode, 0, d_out, nacc);
      run<256>(mode, 0, d_out, nacc);
    }
  return 0;
}
