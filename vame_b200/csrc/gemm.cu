// Generic fp32-accurate GEMM on tcgen05 tensor cores:  C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias)
// Operands are P16 (bf16 hi/lo split, UMMA canonical K-major tiles, see common.cuh); every product is evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM.
//   tile 128 x 128, K pipelined in 64-wide chunks through a 3-stage cp.async.bulk (TMA engine) + mbarrier ring,
//   one producer thread, one MMA-issuing thread, 4 epilogue warps reading TMEM with tcgen05.ld.
// Used for the hoisted (non-recurrent) contractions of the GRU layers: input projections of encoder layer 1, dx of
// layer 1 -> layer 0, and all large weight-gradient GEMMs (split-K with red.add).
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int G_BM = 128, G_BN = 128, G_STAGES = 3;
constexpr int G_TILE_BYTES = 128 * KCHUNK * 2 * 2;       // 32 KB (hi + lo)
constexpr int G_STAGE_BYTES = 2 * G_TILE_BYTES;          // A + B
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + 256;

__device__ __forceinline__ const __nv_bfloat16* seg_tile(const GemmSeg* s, int rb, int kc) {
  // segment i covers the next s[i].nkc chunks
  int i = 0;
  while (i < 3 && kc >= s[i].nkc) {
    kc -= s[i].nkc;
    ++i;
  }
  return reinterpret_cast<const __nv_bfloat16*>(s[i].p) + ((size_t)rb * s[i].rb_stride + kc) * p16_tile_elems(128);
}

__global__ void __launch_bounds__(128, 1) gemm_p16_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);
  uint64_t* empty = full + G_STAGES;
  uint64_t* done = empty + G_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = blockIdx.x, mb = blockIdx.y, split = blockIdx.z;
  const int nkc_total = g.a[0].nkc + g.a[1].nkc + g.a[2].nkc + g.a[3].nkc;
  const int per = (nkc_total + g.splits - 1) / g.splits;
  const int kc_begin = split * per;
  const int kc_end = min(nkc_total, kc_begin + per);
  const int nk = kc_end - kc_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (nk > 0) {
    // warp 0 = producer, warp 1 = MMA issuer: one elected lane works, its siblings are parked at __syncwarp (a sibling
    // spinning on an mbarrier would put the whole warp to sleep and starve the working lane)
    if (warp == 0) {
     if (lane == 0) {
      // ===== producer: TMA-engine bulk copies of whole P16 tiles =====
      for (int i = 0; i < nk; ++i) {
        const int s = i % G_STAGES;
        const uint32_t ph = (i / G_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], G_STAGE_BYTES);
        uint8_t* sa = smem + s * G_STAGE_BYTES;
        bulk_g2s(sa, seg_tile(g.a, mb, kc_begin + i), G_TILE_BYTES, &full[s]);
        bulk_g2s(sa + G_TILE_BYTES, seg_tile(g.b, nb, kc_begin + i), G_TILE_BYTES, &full[s]);
      }
     }
     __syncwarp();
    } else if (warp == 1) {
     if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_bf16(G_BM, 2 * G_BN);          // one descriptor spans [B_hi ; B_lo]
      const uint64_t dA0 = make_desc(smem_u32(smem));
      for (int i = 0; i < nk; ++i) {
        const int s = i % G_STAGES;
        const uint32_t ph = (i / G_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t dA = desc_advance(dA0, s * G_STAGE_BYTES), dB = desc_advance(dA0, s * G_STAGE_BYTES + G_TILE_BYTES);
        constexpr uint32_t plane = 128 * KCHUNK * 2;                  // bytes between hi and lo planes
#pragma unroll
        for (int ks = 0; ks < KCHUNK / 16; ++ks) {
          // D[:, 0:128] += a * b_hi, D[:, 128:256] += a * b_lo (summed in the epilogue)
          const uint32_t ko = ks * 2 * ATOM_BYTES;                   // 16 k-elements = 2 atoms
          if (ks == 0 && i == 0) umma_bf16_c<0>(tmem, desc_advance(dA, plane + ko), desc_advance(dB, ko), idesc);
          else umma_bf16_c<1>(tmem, desc_advance(dA, plane + ko), desc_advance(dB, ko), idesc);
          umma_bf16_c<1>(tmem, desc_advance(dA, ko), desc_advance(dB, ko), idesc);
        }
        umma_commit(&empty[s]);          // frees the smem stage once these MMAs retire
      }
      umma_commit(done);
     }
     __syncwarp();
    }
    // ===== epilogue: all 4 warps, warp w owns TMEM lanes 32w..32w+31 (= rows of the tile) =====
    mbar_wait(done, 0);
    __syncwarp();
    tc_fence_after();
    const int row = mb * G_BM + warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < G_BN; c0 += 16) {
      float v[16], v2[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld16(taddr + G_BN + c0, v2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += v2[j];
      const int col = nb * G_BN + c0;
      if (g.c_fm) {                       // feature-major output: C[n*ldc + m]; lanes (= rows m) are contiguous -> coalesced
        if (row < g.M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (col + j < g.N) {
              float x = v[j];
              if (g.bias && split == 0) x += g.bias[col + j];
              float* dst = g.C + (long)(col + j) * g.ldc + row;
              if (g.atomic) atomicAdd(dst, x);
              else *dst = x;
            }
          }
        }
      } else if (row < g.M && col < g.N) {
        float* crow = g.C + (long)row * g.ldc + col;
        if (g.bias && split == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < g.N) v[j] += g.bias[col + j];
        }
        if (g.atomic) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < g.N) atomicAdd(crow + j, v[j]);
        } else if (col + 16 <= g.N && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < g.N) crow[j] = v[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

void launch_gemm_p16(const GemmArgs& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_p16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    attr_set = true;
  }
  dim3 grid((g.N + G_BN - 1) / G_BN, (g.M + G_BM - 1) / G_BM, g.splits > 0 ? g.splits : 1);
  GemmArgs a = g;
  if (a.splits < 1) a.splits = 1;
  count_launch();
  gemm_p16_kernel<<<grid, 128, G_SMEM, st>>>(a);
}

}  // namespace vb
