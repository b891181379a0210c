// Microbenchmarks behind the round-2 sweep design (results: profiles/r2_tmem_dsmem_microbench.log):
//   1. tcgen05.ld throughput: cycles per 32x32b.x16 load (2 KB per warp instruction) with 1 / 4 / 8 warps reading (one or two warps
//      per tensor-memory lane quarter) - is the accumulator drain of the recurrent sweeps a bandwidth term?
//   2. DSMEM push bandwidth inside a 4-CTA cluster: every CTA sends X bytes to each of its 3 peers with (a) 16-byte st.async
//      from 128 / 256 threads, (b) one cp.async.bulk (shared::cta -> shared::cluster) per peer issued by one thread; the
//      receiver's mbarrier counts the bytes.  Reported: cycles from "start pushing" to "everything addressed to me has landed".
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vame_b200/csrc/common.cuh"
using namespace vb;

__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async16(uint32_t dst, const uint4& v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---- 1. tensor-memory read throughput -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) tmem_ld_kernel(int nwarps, int iters, int wait_every, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)((i * 16) & 255) + (warp >= 4 ? 256u : 0u), v);
      if ((i + 1) % wait_every == 0) tmem_ld_wait();
      acc += v[0] + v[15];
    }
    tmem_ld_wait();
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0 && warp < nwarps) out[warp] = t1 - t0;
  if (acc == 123.456f) out[31] = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

// ---- 2. DSMEM push ---------------------------------------------------------------------------------------------------
// mode 0: st.async from `nthreads` threads; mode 1: one bulk copy per peer; mode 2: bulk copies split in 4 pieces per peer
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int mode, int bytes, int nthreads, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* src = smem;                 // [bytes]
  uint8_t* dst = smem + 32768;         // [3][bytes]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 3 * 32768);
  const uint32_t c = blockIdx.x & 3u;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(src)[i] = make_uint4(i, c, 0, 0);
  fence_proxy_async_smem();
  __syncthreads();
  long long total = 0, worst = 0;
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) mbar_expect_tx(bar, 3 * bytes);
    cluster_arrive_release();
    cluster_wait_acquire();
    const long long t0 = clock64();
    if (mode == 0) {
      if (tid < nthreads) {
        for (int i = tid; i < bytes / 16; i += nthreads) {
          const uint4 v = reinterpret_cast<const uint4*>(src)[i];
#pragma unroll
          for (uint32_t r = 1; r < 4; ++r) {
            const uint32_t peer = (c + r) & 3, slot = 3 - r;          // my position among the peer's three sources
            st_async16(mapa(smem_u32(dst) + slot * 32768 + i * 16, peer), v, mapa(smem_u32(bar), peer));
          }
        }
      }
    } else if (tid == 0) {
      const int pieces = mode == 2 ? 4 : 1;
      for (uint32_t r = 1; r < 4; ++r) {
        const uint32_t peer = (c + r) & 3, slot = 3 - r;
        for (int p = 0; p < pieces; ++p)
          bulk_s2c(mapa(smem_u32(dst) + slot * 32768 + p * (bytes / pieces), peer), smem_u32(src) + p * (bytes / pieces), bytes / pieces,
                   mapa(smem_u32(bar), peer));
      }
    }
    const long long t1 = clock64();
    while (!try_wait_cluster(bar, it & 1)) {
    }
    const long long t2 = clock64();
    if (tid == 0) {
      total += t2 - t0;
      if (t1 - t0 > worst) worst = t1 - t0;
    }
  }
  cluster_arrive_release();
  cluster_wait_acquire();
  if (tid == 0 && blockIdx.x == 0) {
    out[0] = total / iters;
    out[1] = worst;
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  long long h[32];
  for (int wait_every : {1, 4}) {
    for (int nw : {1, 4, 8}) {
      const int iters = 1024;
      for (int rep = 0; rep < 2; ++rep) tmem_ld_kernel<<<1, 256>>>(nw, iters, wait_every, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("tmem_ld: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("tcgen05.ld 32x32b.x16 (2 KB): %d warps, wait every %d: %.1f cycles per load per warp -> %.1f B/cycle/SM\n", nw, wait_every,
             (double)mx / iters, 2048.0 * nw * iters / (double)mx);
    }
  }
  const int smem = 32768 + 3 * 32768 + 64;
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {4, 128}) {
    for (int bytes : {4096, 8192, 16384}) {
      for (int cfgi = 0; cfgi < 4; ++cfgi) {
        const int mode = cfgi < 2 ? 0 : cfgi - 1, nth = cfgi == 1 ? 256 : 128;
        cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, dsmem_kernel, mode, bytes, nth, 200, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("dsmem: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
        printf("DSMEM push grid %3d, %5d B to each of 3 peers, %-28s: %5lld cycles until all 3 x %d B have landed (%.1f B/cycle in), issue %lld\n",
               grid, bytes, mode == 0 ? (nth == 256 ? "st.async x 256 threads" : "st.async x 128 threads") : mode == 1 ? "1 bulk copy per peer" : "4 bulk copies per peer",
               h[0], bytes, 3.0 * bytes / (double)h[0], h[1]);
      }
    }
  }
  return 0;
}
