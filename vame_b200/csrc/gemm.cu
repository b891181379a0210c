// Generic fp32-accurate GEMM on tcgen05 tensor cores:  C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias)
// Operands are P16 (bf16 hi/lo split, UMMA canonical K-major tiles, see common.cuh).  A 128 x 128 output tile is
// accumulated as  D[:, 0:128] += a_hi * b_hi + a_lo * b_hi ,  D[:, 128:256] += a_hi * b_lo : the hi and lo planes of a
// B tile are contiguous row groups, so ONE N = 256 descriptor covers both for the a_hi pass (128 cycles per 16-wide k-step)
// and the a_lo pass uses the first 128 rows only (64 cycles); a tcgen05.mma costs max(~45, N/2) cycles (tools/bench_mma).
// The epilogue adds the two column halves.  (The a_lo * b_lo products are below the split's own 2^-17 representation error.)
//
// Persistent, warp-specialised (the canonical Blackwell GEMM shape):
//   warp 0      producer : whole P16 tiles with the TMA engine (cp.async.bulk) into a 3-stage smem ring (full/empty mbarriers)
//   warp 1      MMA issuer (one elected lane), owns the TMEM allocation (512 columns = two 256-column accumulators)
//   warps 2..9  epilogue : TMEM -> registers -> global (row-major, feature-major or red.add), two warps per TMEM lane quarter
// The accumulator is double buffered (tmem_full / tmem_empty mbarriers), so the epilogue of work item i overlaps the
// main loop of item i+1; every CTA walks the (m tile, n tile, k split) work list with stride gridDim.x.
//
// Used for all hoisted (non-recurrent) contractions of the GRU layers: input projections, dx of layer 1 -> layer 0, the
// small linears and all weight-gradient GEMMs (split-K with red.add).
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int G_BM = 128, G_BN = 128, G_STAGES = 3;
constexpr bool g_gemm_lolo = false;                      // true: also accumulate a_lo x B_lo (4 products instead of 3)
constexpr int G_TILE_BYTES = 128 * KCHUNK * 2 * 2;       // 32 KB (hi + lo)
constexpr int G_STAGE_BYTES = 2 * G_TILE_BYTES;          // A + B
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + 256;
constexpr int G_THREADS = 320;                           // 10 warps
constexpr int G_EPI_WARPS = 8;

__device__ __forceinline__ const __nv_bfloat16* seg_tile(const GemmSeg* s, int rb, int kc) {
  // segment i covers the next s[i].nkc chunks
  int i = 0;
  while (i < 3 && kc >= s[i].nkc) {
    kc -= s[i].nkc;
    ++i;
  }
  return reinterpret_cast<const __nv_bfloat16*>(s[i].p) + ((size_t)rb * s[i].rb_stride + kc) * p16_tile_elems(128);
}

struct WorkItem {
  int mb, nb, split, kc_begin, nk;
};
__device__ __forceinline__ WorkItem work_item(const GemmArgs& g, int w, int n_nb, int nkc_total, int per) {
  WorkItem it;
  it.split = w % g.splits;
  const int t = w / g.splits;
  it.nb = t % n_nb;
  it.mb = t / n_nb;
  it.kc_begin = it.split * per;
  const int kc_end = min(nkc_total, it.kc_begin + per);
  it.nk = max(0, kc_end - it.kc_begin);
  return it;
}

__global__ void __launch_bounds__(G_THREADS, 1) gemm_p16_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);
  uint64_t* empty = full + G_STAGES;
  uint64_t* tfull = empty + G_STAGES;        // [2] accumulator ready
  uint64_t* tempty = tfull + 2;              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int n_nb = (g.N + G_BN - 1) / G_BN, n_mb = (g.M + G_BM - 1) / G_BM;
  const int nkc_total = g.a[0].nkc + g.a[1].nkc + g.a[2].nkc + g.a[3].nkc;
  const int per = (nkc_total + g.splits - 1) / g.splits;
  const int n_work = n_mb * n_nb * g.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], G_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    if (elect_one()) {
      uint32_t it_global = 0;                                  // running stage counter across work items
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const WorkItem it = work_item(g, w, n_nb, nkc_total, per);
        for (int i = 0; i < it.nk; ++i, ++it_global) {
          const int s = it_global % G_STAGES;
          const uint32_t ph = (it_global / G_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], G_STAGE_BYTES);
          uint8_t* sa = smem + s * G_STAGE_BYTES;
          bulk_g2s(sa, seg_tile(g.a, it.mb, it.kc_begin + i), G_TILE_BYTES, &full[s]);
          bulk_g2s(sa + G_TILE_BYTES, seg_tile(g.b, it.nb, it.kc_begin + i), G_TILE_BYTES, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(G_BM, 2 * G_BN);          // one descriptor spans [B_hi ; B_lo]
      // the a_lo pass multiplies B_hi only (N = 128: same descriptor, half the columns): the lo x lo products are ~2^-18 of a
      // term, below the 2^-17 representation error of the hi/lo split itself, and cost 25 % of the tensor time
      const uint32_t idesc_lo = g_gemm_lolo ? idesc : make_idesc_bf16(G_BM, G_BN);
      const uint64_t dA0 = make_desc(smem_u32(smem));
      constexpr uint32_t plane = 128 * KCHUNK * 2;                     // bytes between hi and lo planes
      uint32_t it_global = 0, n_items = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const WorkItem it = work_item(g, w, n_nb, nkc_total, per);
        if (it.nk == 0) continue;                                     // (the epilogue skips it as well)
        const uint32_t buf = n_items & 1, tph = (n_items >> 1) & 1;
        ++n_items;
        mbar_wait(&tempty[buf], tph ^ 1);                             // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t dcol = tmem + buf * 256;
        for (int i = 0; i < it.nk; ++i, ++it_global) {
          const int s = it_global % G_STAGES;
          const uint32_t ph = (it_global / G_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t dA = desc_advance(dA0, s * G_STAGE_BYTES), dB = desc_advance(dA0, s * G_STAGE_BYTES + G_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < KCHUNK / 16; ++ks) {
            const uint32_t ko = ks * 2 * ATOM_BYTES;                 // 16 k-elements = 2 atoms
            // a_hi x [B_hi ; B_lo] first (it initialises all 256 columns), then a_lo x B_hi into columns 0-127
            if (ks == 0 && i == 0) umma_bf16_c<0>(dcol, desc_advance(dA, ko), desc_advance(dB, ko), idesc);
            else umma_bf16_c<1>(dcol, desc_advance(dA, ko), desc_advance(dB, ko), idesc);
            umma_bf16_c<1>(dcol, desc_advance(dA, plane + ko), desc_advance(dB, ko), idesc_lo);
          }
          umma_commit(&empty[s]);                                    // frees the smem stage once these MMAs retire
        }
        umma_commit(&tfull[buf]);                                    // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..9: lane quarter q = warp % 4 (hardware restriction of tcgen05.ld), column half = (warp - 2) / 4 =====
    const int q = warp & 3, chalf = (warp - 2) >> 2;
    uint32_t n_items = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const WorkItem it = work_item(g, w, n_nb, nkc_total, per);
      if (it.nk == 0) continue;
      const uint32_t buf = n_items & 1, tph = (n_items >> 1) & 1;
      ++n_items;
      mbar_wait(&tfull[buf], tph);
      __syncwarp();
      tc_fence_after();
      const int row = it.mb * G_BM + q * 32 + lane;
      const uint32_t taddr = tmem + buf * 256 + ((uint32_t)(q * 32) << 16);
      const bool fm_vec = g.c_fm && !g.atomic && (g.M & 3) == 0 && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll 1
      for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 16) {
        float v[16], v2[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld16(taddr + G_BN + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += v2[j];
        const int col = it.nb * G_BN + c0;
        if (g.c_fm && fm_vec) {
          // feature-major output C[n*ldc + m], 16 bytes per lane: an SM retires one warp-level store instruction per ~25 cycles
          // whatever its width (measured: 512 scalar stores per tile made the K = 64 input projection 43 us instead of ~12),
          // so 4 x 4 blocks are transposed across 4 adjacent lanes first: lane 4g + r then holds rows 4g .. 4g+3 of column j4 + r
          if (g.bias && it.split == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col + j < g.N) v[j] += g.bias[col + j];
          }
          const int r = lane & 3;
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            float a0 = v[j4], a1 = v[j4 + 1], a2 = v[j4 + 2], a3 = v[j4 + 3];
            {   // exchange 2 x 2 blocks with lane ^ 2
              const bool up = (r & 2) != 0;
              const float s0 = up ? a0 : a2, s1 = up ? a1 : a3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
              if (up) { a0 = r0; a1 = r1; } else { a2 = r0; a3 = r1; }
            }
            {   // exchange single elements with lane ^ 1
              const bool up = (r & 1) != 0;
              const float s0 = up ? a0 : a1, s1 = up ? a2 : a3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              if (up) { a0 = r0; a2 = r1; } else { a1 = r0; a3 = r1; }
            }
            // now (a0, a1, a2, a3) = column j4 + r of rows 4g, 4g+1, 4g+2, 4g+3
            const int cj = col + j4 + r;
            const int row4 = it.mb * G_BM + q * 32 + (lane & ~3);
            if (cj < g.N && row4 < g.M) {
              long off = (long)cj * g.ldc + row4;
              if (g.c_fm == 2) {          // same plane, permuted into the rw sweeps' lane-major blocks
                const int u = cj & 255, t = row4 / g.pv_bp, b = row4 - t * g.pv_bp;
                const long blk = ((((long)t * (g.pv_bp >> 4) + (b >> 4)) * 4 + (u >> 6)) * 4 + ((u & 63) >> 4)) * 256;
                off = (long)(cj >> 8) * 256 * g.ldc + blk + ((u & 15) + 16 * ((b & 15) >> 3)) * 8 + (b & 7);
              }
              *reinterpret_cast<float4*>(g.C + off) = make_float4(a0, a1, a2, a3);
            }
          }
        } else if (g.c_fm) {                // feature-major output: C[n*ldc + m]; lanes (= rows m) are contiguous -> coalesced
          if (row < g.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (col + j < g.N) {
                float x = v[j];
                if (g.bias && it.split == 0) x += g.bias[col + j];
                float* dst = g.C + (long)(col + j) * g.ldc + row;
                if (g.atomic) atomicAdd(dst, x);
                else *dst = x;
              }
            }
          }
        } else if (row < g.M && col < g.N) {
          float* crow = g.C + (long)row * g.ldc + col;
          if (g.bias && it.split == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col + j < g.N) v[j] += g.bias[col + j];
          }
          if (g.atomic) {
            // split-K accumulation: vectorised reductions (red.global.add.v4.f32, 16 B per lane) - the scalar form costs
            // ~1.3 cycles per lane-op on the SM's LSU path and made the epilogue (16K ops per item) longer than the main loop
            if (col + 16 <= g.N && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) red_add_v4(crow + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col + j < g.N) atomicAdd(crow + j, v[j]);
            }
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);                       // this warp's TMEM reads of the buffer are complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

void launch_gemm_p16(const GemmArgs& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_p16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    attr_set = true;
  }
  GemmArgs a = g;
  if (a.splits < 1) a.splits = 1;
  const int n_work = ((a.N + G_BN - 1) / G_BN) * ((a.M + G_BM - 1) / G_BM) * a.splits;
  const int cap = (a.max_ctas > 0 && a.max_ctas < 148) ? a.max_ctas : 148;
  int grid = n_work < cap ? n_work : cap;
  if (grid < 1) grid = 1;
  count_launch();
  gemm_p16_kernel<<<grid, G_THREADS, G_SMEM, st>>>(a);
}

}  // namespace vb
