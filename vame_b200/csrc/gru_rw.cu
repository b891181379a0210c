// Resident-weight recurrent sweeps ("rw" kernels): the whole forward / BPTT sweep of one bi-GRU layer in one
// thread-block-cluster kernel whose recurrence never leaves the cluster.
//
// Why: in the slice kernels of gru.cu a time step hands h_t (or the BPTT partial sums) from the producing SMs to the
// consuming SMs through L2 (64 KB per CTA and step) - a 1.7 us round trip that dominates the 5 / 7 us step at B = 256.
// Here the GEMM is transposed (swap-AB): the WEIGHTS are the M side of tcgen05.mma and stay in shared memory for the
// whole sweep, the batch rows are the (small) N side.
//   * a cluster of 4 CTAs owns 16 batch rows of one direction; CTA c owns the gate rows of H/4 hidden units:
//     3 (gates) x H/64 (k chunks) A tiles of 128 rows x 64 k bf16 = 3 H^2 bytes (192 KB at H = 256).  A tile row (= TMEM lane)
//     32 q + l holds the bf16 HI plane of unit 16 q + l for l < 16 and the LO plane of unit 16 q + l - 16 for l >= 16, so the
//     hi / lo partial products of one unit sit in the two half-warps of the warp that owns TMEM lane quarter q.
//   * the B operand is h_{t-1} of the 16 rows as [16 hi rows ; 16 lo rows] x H (K-major, 4 KB per k chunk): ONE
//     MMA (M = 128, N = 32, K = 16) yields all four hi/lo partial products.  3 H/16 = 48 MMAs per step and CTA.
//   * after the gate math every CTA pushes its 16 x H/4 slice of h_t (bf16 hi/lo, already in UMMA operand layout) into the
//     B operand buffer of all 4 CTAs with 16-byte st.shared::cluster stores (12 KB out per CTA and step) and the cluster
//     synchronises with ONE barrier.cluster per step (double-buffered operand).  No L2 round trip, no TMA re-fetch.
//   * BPTT: dh_{t-1}^T = W_hh^T dgh^T as a K-split: CTA c holds W_hh^T[:, gate rows of its units] (H/64 A tiles of 128 x 3H/4),
//     builds the dgh^T operand of its own units locally and pushes the fp32 partial sums of the units owned by CTA o into
//     o's receive buffer (4 source slots); the operand buffer aliases the receive slot that was consumed at the top of the step.
// Numerics are those of the slice kernels (bf16 hi/lo split of both operands, fp32 accumulation, all four partial
// products); cell equations: torch.nn.GRU as used at vame/model/rnn_model.py:41,106,141, BPTT: SURVEY.md section 3.5.
#include "common.cuh"
#include "kernels.h"

namespace vb {

__device__ unsigned int g_rw_timeouts = 0;     // bounded mbarrier waits that gave up (must stay 0; read by vame_debug_rw_timeouts)
__device__ int g_rw_timeout_site[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // first time-out: site id, parity, blockIdx.x/y/z, threadIdx.x

#ifdef VAME_ACCURATE_MATH
__device__ __forceinline__ float rw_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float rw_tanh(float x) { return tanhf(x); }
#else
__device__ __forceinline__ float rw_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float rw_tanh(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
#endif

constexpr int RW_ATILE = 128 * KCHUNK * 2;     // 16384 B: 128 weight rows x 64 k (bf16)
constexpr int RW_BTILE = 32 * KCHUNK * 2;      //  4096 B: [16 hi rows ; 16 lo rows] x 64 k (bf16)
constexpr int RW_THREADS = 160;                // warps 0-3: element-wise work (one TMEM lane quarter each), warp 4: TMA + MMA issue

// Rank of the CTA in its cluster: the rw grids are (4, groups, directions) with cluster dims (4, 1, 1), so the rank is
// blockIdx.x (which the compiler knows to be warp-uniform, unlike a special register read through inline asm).
__device__ __forceinline__ uint32_t cluster_ctarank() { return blockIdx.x & 3u; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// a wait that cannot hang the GPU: a protocol bug shows up as a counted time-out (and wrong numbers), not as a dead box
__device__ __noinline__ void rw_timeout(int site, uint32_t parity) {
  if (atomicAdd(&g_rw_timeouts, 1u) == 0) {
    g_rw_timeout_site[0] = site; g_rw_timeout_site[1] = (int)parity;
    g_rw_timeout_site[2] = blockIdx.x; g_rw_timeout_site[3] = blockIdx.y; g_rw_timeout_site[4] = blockIdx.z;
    g_rw_timeout_site[5] = threadIdx.x;
  }
}
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int site = 0) {
  for (int i = 0; i < (1 << 20); ++i) {
    if (mbar_try_wait(bar, parity)) return;
    if ((i & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(&g_rw_timeouts) != 0) return;   // someone already gave up
  }
  rw_timeout(site, parity);
}
// all lanes of a warp wait, then reconverge (tcgen05.ld is .sync.aligned)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int site = 0) {
  mbar_wait_b(bar, parity, site);
  __syncwarp();
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 element-wise warps

// 8 consecutive floats of one thread as ONE 256-bit global access (sm_100: LDG/STG.E.ENL2.256): an SM retires about one
// warp-level memory instruction per ~25 cycles whatever its width, so the instruction count is what the sweeps pay for.
// All call sites are 32-byte aligned (batch offsets are multiples of 8 rows, leading dimensions multiples of 128).
__device__ __forceinline__ void ld8(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}
// the same from shared memory (generic address)
__device__ __forceinline__ void ld8s(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// One accumulator block of 32 columns: this lane's row holds (plane of lane) x [h_hi rows 0-15 | h_lo rows 0-15].
// Lane l < 16 ends up with rows 0-7, lane l + 16 with rows 8-15 of the same unit, hi and lo planes summed.
__device__ __forceinline__ void rw_reduce32(uint32_t taddr, int half, float* out8) {
  float v[32];
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16, v + 16);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float lo = v[i] + v[16 + i], hi = v[8 + i] + v[24 + i];
    const float send = half ? lo : hi, keep = half ? hi : lo;
    out8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
}

// v[i] = value(unit of this lane, row 8 half + i).  8 x 8 transpose inside each group of 8 lanes (3 butterfly stages on
// packed bf16 hi|lo words): afterwards this lane holds row 8 half + (lane & 7) for the 8 consecutive units of its octet,
// as one 16-byte vector per plane - an atom row of the UMMA K-major operand layout.
__device__ __forceinline__ void rw_transpose_pack(const float* v, int lane, uint4& hi, uint4& lo) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat16 h, l;
    split_bf16(v[i], h, l);
    w[i] = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
  }
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    const bool bit = (lane & d) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j & d) continue;
      const uint32_t x = bit ? w[j] : w[j | d];
      const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
      if (bit) w[j] = y;
      else w[j | d] = y;
    }
  }
  hi.x = (w[0] & 0xffffu) | (w[1] << 16); hi.y = (w[2] & 0xffffu) | (w[3] << 16);
  hi.z = (w[4] & 0xffffu) | (w[5] << 16); hi.w = (w[6] & 0xffffu) | (w[7] << 16);
  lo.x = (w[0] >> 16) | (w[1] & 0xffff0000u); lo.y = (w[2] >> 16) | (w[3] & 0xffff0000u);
  lo.z = (w[4] >> 16) | (w[5] & 0xffff0000u); lo.w = (w[6] >> 16) | (w[7] & 0xffff0000u);
}
// byte offset of (row n of the 32-row operand, k) inside an operand buffer of RW_BTILE-sized k chunks; k % 8 == 0
__device__ __forceinline__ uint32_t rw_b_off(int n, int k) { return (uint32_t)(k >> 6) * RW_BTILE + 2u * (uint32_t)p16_in_tile(n, k & 63); }

// MN-major B operand of the H = 256 kernels (tcgen05 instruction-descriptor bit 16): inside every 8 x 8 core matrix the 8 batch
// rows of ONE k (hidden unit / gate row) are the contiguous 16 bytes, core matrices of consecutive k-groups 128 B apart (LBO),
// of consecutive 8-row groups 1024 B apart (SBO) - the K-major tile with its core matrices transposed.  A thread owns one unit
// and 8 consecutive batch rows, so it writes its operand rows as ONE 16-byte vector per plane: no 8 x 8 shuffle transpose on
// the recurrence (24 shuffles per operand in the K-major form).
#ifndef RW_MN_FWD
#define RW_MN_FWD 0      // the forward kernel keeps the K-major operand: with MN-major the store-warp variant faults on the B200
#endif                   // (the variants without store warps pass; not understood yet - see DESIGN.md, open items)
#ifndef RW_MN_BWD
#define RW_MN_BWD 1
#endif
__host__ __device__ constexpr uint32_t rw_idesc(bool mn) { return make_idesc_bf16(128, 32) | (mn ? (1u << 16) : 0u); }
// byte offset of (rows n0 .. n0+7 of the 32-row operand, k) inside an operand buffer of RW_BTILE-sized 64-k chunks
__device__ __forceinline__ uint32_t rw_mn_off(int n0, int k) {
  return (uint32_t)(k >> 6) * RW_BTILE + (uint32_t)((n0 >> 3) * 1024 + ((k & 63) >> 3) * 128 + (k & 7) * 16);
}

#define RW_STAMP(i)                                                                                          \
  do {                                                                                                        \
    if (a.dbg && s == 10 && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {      \
      unsigned long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
      a.dbg[(i)] = t_;                                                                                        \
    }                                                                                                         \
  } while (0)
// per-CTA stamps of cluster (y = 0, z = 0), step 10: dbg[16 + 4 * rank + i]
#define RW_STAMP_CTA(i)                                                                                      \
  do {                                                                                                        \
    if (a.dbg && (a.exp & 32) && s == 10 && blockIdx.y == 0 && blockIdx.z == 0) {                             \
      unsigned long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
      a.dbg[16 + 4 * blockIdx.x + (i)] = t_;                                                                  \
    }                                                                                                         \
  } while (0)
#define RW_STAMP_CTA8(i)                                                                                     \
  do {                                                                                                        \
    if (a.dbg && (a.exp & 32) && s == 10 && blockIdx.y == 0 && blockIdx.z == 0) {                             \
      unsigned long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
      a.dbg[16 + 8 * blockIdx.x + (i)] = t_;                                                                  \
    }                                                                                                         \
  } while (0)
#define RW_STAMP_MMA(i)                                                                                      \
  do {                                                                                                        \
    if (a.dbg && s == 10 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {                          \
      unsigned long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
      a.dbg[(i)] = t_;                                                                                        \
    }                                                                                                         \
  } while (0)

// =================================================================================================
// forward sweep
// =================================================================================================
// grid = (4, B_pad / 16, directions), cluster = (4, 1, 1), 160 threads.  NKC = H / 64.
template <int NKC>
__global__ void __launch_bounds__(RW_THREADS, 1) gru_rw_fwd_kernel(const GruSeqFwdArgs a) {
  constexpr int H = 64 * NKC, UC = 16 * NKC;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                                             // [3 gates][NKC][RW_ATILE]
  uint8_t* sH = smem + 3 * NKC * RW_ATILE;                        // [2][NKC][RW_BTILE]
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sH + 2 * NKC * RW_BTILE);
  uint64_t* done = wbar + 1;                                      // [3]: accumulator of gate g complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 3);

  const uint32_t c = cluster_ctarank();
  const GruSeqDirFwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp, half = lane >> 4, oct = (lane >> 3) & 1;
  const bool epi = warp < 4 && q < NKC;                           // this warp owns 16 units of the slice
  const long Bp = (long)a.tiles * 128;
  const int steps = a.steps;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int g = 0; g < 3; ++g) mbar_init(&done[g], 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.w_rw) + (size_t)c * 3 * NKC * (RW_ATILE / 2);
      mbar_expect_tx(wbar, 3 * NKC * RW_ATILE);
      for (int i = 0; i < 3 * NKC; ++i) bulk_g2s(sW + (size_t)i * RW_ATILE, wp + (size_t)i * (RW_ATILE / 2), RW_ATILE, wbar);
    }
    __syncwarp();
  }

  const int j = 16 * q + (lane & 15);                             // unit inside the slice
  const int u = (int)c * UC + j;                                  // hidden unit
  const long b0 = (long)blockIdx.y * 16 + 8 * half;               // first of this lane's 8 batch rows
  const int nrow = 8 * half + (lane & 7);                         // operand row this lane writes after the transpose
  float hprev[8], bhn = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) hprev[i] = 0.f;
  if (epi) {
    bhn = d.b_hn[u];
    // initial state -> operand buffer 0 (every CTA builds the complete 16 x H operand from global memory)
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float v[8];
      ld8(d.h0 + (long)(cc * UC + j) * d.h0_ld + b0, v);
      if (cc == (int)c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hprev[i] = v[i];
      }
      uint4 hi, lo;
      rw_transpose_pack(v, lane, hi, lo);
      const int k0 = cc * UC + 16 * q + 8 * oct;
      *reinterpret_cast<uint4*>(sH + rw_b_off(nrow, k0)) = hi;
      *reinterpret_cast<uint4*>(sH + rw_b_off(16 + nrow, k0)) = lo;
    }
    fence_proxy_async_smem();
  }
  // every CTA of the cluster is running and initialised before anyone writes into a peer's shared memory
  cluster_arrive_release();
  cluster_wait_acquire();

  const uint32_t idesc = make_idesc_bf16(128, 32);
  const uint32_t taddr = tmem + ((uint32_t)((q & 3) * 32) << 16);
  float gir[8], giz[8], gin[8];
  if (epi) {
    const int t0 = d.reverse ? steps - 1 : 0;
    const float* gi_row = d.gi + (b0 * d.gi_bs + (long)t0 * d.gi_ts);
    ld8(gi_row + (long)u * d.gi_ld, gir);
    ld8(gi_row + (long)(H + u) * d.gi_ld, giz);
    ld8(gi_row + (long)(2 * H + u) * d.gi_ld, gin);
  }

  for (int s = 0; s < steps; ++s) {
    const int t = d.reverse ? steps - 1 - s : s;
    const int so = (d.out_slots == steps) ? t : (s & 1);
    const int sp = (d.out_p_slots == steps) ? t : (s & 1);
    const uint32_t ph = s & 1;
    RW_STAMP(0);
    if (s > 0) cluster_wait_acquire();                            // every CTA's slice of h_{t-1} has landed in buffer s & 1
    RW_STAMP(1);
    if (warp == 4) {
      if (elect_one()) {
        if (s == 0) mbar_wait_b(wbar, 0, 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint64_t dA = make_desc(smem_u32(sW)), dB = make_desc(smem_u32(sH) + ph * NKC * RW_BTILE);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ao = (g * NKC + kc) * RW_ATILE + ks * 2 * ATOM_BYTES, bo = kc * RW_BTILE + ks * 2 * ATOM_BYTES;
              if (kc == 0 && ks == 0) umma_bf16_c<0>(tmem + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
              else umma_bf16_c<1>(tmem + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
            }
          }
          umma_commit(&done[g]);
        }
        RW_STAMP_MMA(2);
      }
      __syncwarp();
    }
    float hn[8], sr[8], sz[8], sn[8], sg[8];
    uint4 phi = make_uint4(0, 0, 0, 0), plo = make_uint4(0, 0, 0, 0);
    if (epi) {
      float ar[8], az[8], an[8];
      mbar_wait_warp(&done[0], ph, 2);
      tc_fence_after();
      rw_reduce32(taddr, half, ar);
#pragma unroll
      for (int i = 0; i < 8; ++i) sr[i] = rw_sigmoid(gir[i] + ar[i]);
      mbar_wait_warp(&done[1], ph, 3);
      tc_fence_after();
      rw_reduce32(taddr + 32, half, az);
#pragma unroll
      for (int i = 0; i < 8; ++i) sz[i] = rw_sigmoid(giz[i] + az[i]);
      mbar_wait_warp(&done[2], ph, 4);
      tc_fence_after();
      RW_STAMP(3);
      rw_reduce32(taddr + 64, half, an);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sg[i] = an[i] + bhn;
        sn[i] = rw_tanh(gin[i] + sr[i] * sg[i]);
        hn[i] = (1.0f - sz[i]) * sn[i] + sz[i] * hprev[i];
        hprev[i] = hn[i];
      }
      RW_STAMP(4);
      rw_transpose_pack(hn, lane, phi, plo);
      if (s + 1 < steps) {                                        // h_t -> operand buffer (s + 1) & 1 of all 4 CTAs
        const int k0 = (int)c * UC + 16 * q + 8 * oct;
        const uint32_t base = smem_u32(sH) + (ph ^ 1u) * NKC * RW_BTILE;
        const uint32_t ohi = base + rw_b_off(nrow, k0), olo = base + rw_b_off(16 + nrow, k0);
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
          st_cluster_v4(mapa_u32(ohi, r), phi);
          st_cluster_v4(mapa_u32(olo, r), plo);
        }
        fence_proxy_async_all();
      }
      tc_fence_before();
    }
    RW_STAMP(5);
    if (s + 1 < steps) cluster_arrive_release();
    RW_STAMP(6);
    // ---- off the recurrence: sequence outputs, saved gates, next step's input projections ----
    if (epi) {
      if (s + 1 < steps) {                                        // loads first: the LSU works in order and the stores below are many
        const int tn = d.reverse ? t - 1 : t + 1;
        const float* gi_row = d.gi + (b0 * d.gi_bs + (long)tn * d.gi_ts);
        ld8(gi_row + (long)u * d.gi_ld, gir);
        ld8(gi_row + (long)(H + u) * d.gi_ld, giz);
        ld8(gi_row + (long)(2 * H + u) * d.gi_ld, gin);
      }
      st8(d.out + (long)u * d.out_ld + (long)so * Bp + b0, hn);
      {
        const long row = (long)blockIdx.y * 16 + nrow;            // batch row of the transposed vectors
        const int k0 = (int)c * UC + 16 * q + 8 * oct;
        __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(d.out_p) + (size_t)sp * d.out_p_slot_elems +
                            ((size_t)(row >> 7) * NKC + (k0 >> 6)) * p16_tile_elems(128);
        const int off = p16_in_tile((int)(row & 127), k0 & 63);
        *reinterpret_cast<uint4*>(tl + off) = phi;
        *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = plo;
      }
      if (d.sv[0]) {
        const long o = (long)u * d.sv_ld + (long)t * Bp + b0;
        st8(d.sv[0] + o, sr);
        st8(d.sv[1] + o, sz);
        st8(d.sv[2] + o, sn);
        st8(d.sv[3] + o, sg);
      }
    }
    RW_STAMP(7);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

// =================================================================================================
// backward sweep (BPTT)
// =================================================================================================
// SUM: also accumulate the time sums of the input-gate gradients (decoders: the GRU input is z at every step).
template <int NKC, bool SUM>
__global__ void __launch_bounds__(RW_THREADS, 1) gru_rw_bwd_kernel(const GruSeqBwdArgs a) {
  constexpr int H = 64 * NKC, UC = 16 * NKC;
  constexpr int MT = NKC;                          // A tiles: 64 input units (hi + lo rows) each
  constexpr int KS = 3 * NKC;                      // k-steps of 16 over the CTA's 3 UC gate rows
  constexpr int NKB = (3 * UC + KCHUNK - 1) / KCHUNK;
  constexpr int RSLOT = (4 * UC * 16 * 4 > NKB * RW_BTILE) ? 4 * UC * 16 * 4 : NKB * RW_BTILE;   // receive slot (also the operand)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                                             // [MT][NKB][RW_ATILE]
  uint8_t* sR = smem + MT * NKB * RW_ATILE;                       // [2][RSLOT]
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sR + 2 * RSLOT);
  uint64_t* done = wbar + 1;                                      // [MT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 4);

  const uint32_t c = cluster_ctarank();
  const GruSeqDirBwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp, half = lane >> 4, oct = (lane >> 3) & 1;
  const bool ew = warp < 4;                                       // element-wise warp (drains TMEM lane quarter q)
  const bool epi = ew && q < NKC;                                 // ... that also owns 16 units of the slice
  const long bpad = (long)a.tiles * 128;
  const int steps = a.steps;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int m = 0; m < 4; ++m) mbar_init(&done[m], 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.wT_rw) + (size_t)c * MT * NKB * (RW_ATILE / 2);
      mbar_expect_tx(wbar, MT * NKB * RW_ATILE);
      for (int i = 0; i < MT * NKB; ++i) bulk_g2s(sW + (size_t)i * RW_ATILE, wp + (size_t)i * (RW_ATILE / 2), RW_ATILE, wbar);
    }
    __syncwarp();
  }
  cluster_arrive_release();
  cluster_wait_acquire();

  const int j = 16 * q + (lane & 15);
  const int u = (int)c * UC + j;
  const long b0 = (long)blockIdx.y * 16 + 8 * half;
  const int nrow = 8 * half + (lane & 7);
  const uint32_t idesc = make_idesc_bf16(128, 32);
  const uint32_t taddr = tmem + ((uint32_t)((q & 3) * 32) << 16);

  float carry[8];
  float sum_r[SUM ? 8 : 1], sum_z[SUM ? 8 : 1], sum_n[SUM ? 8 : 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) carry[i] = 0.f;
  if constexpr (SUM) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sum_r[i] = sum_z[i] = sum_n[i] = 0.f;
  }
  float r[8], z[8], n[8], ghn[8], hp[8], dh[8];
  auto load_step = [&](int s) {
    const int t = d.reverse ? s : steps - 1 - s;
    const bool first_fwd = d.reverse ? (t == steps - 1) : (t == 0);
    const int tprev = d.reverse ? t + 1 : t - 1;
    const long so = (long)u * d.sv_ld + (long)t * bpad + b0;
    ld8(d.sv[0] + so, r);
    ld8(d.sv[1] + so, z);
    ld8(d.sv[2] + so, n);
    ld8(d.sv[3] + so, ghn);
    if (first_fwd) ld8(d.h0 + (long)u * d.h0_ld + b0, hp);
    else ld8(d.out + (long)u * d.out_ld + (long)tprev * bpad + b0, hp);
    if (d.dout) ld8(d.dout + (long)u * d.dout_ld + (long)t * bpad + b0, dh);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] = 0.f;
    }
  };
  if (epi) {
    if (d.dh_last) ld8(d.dh_last + (long)u * d.dh_last_ld + b0, carry);
    load_step(0);
  }

  for (int s = 0; s < steps; ++s) {
    const int t = d.reverse ? s : steps - 1 - s;
    const uint32_t ph = s & 1;
    uint8_t* rprev = sR + (ph ^ 1u) * RSLOT;                      // partial sums of step s - 1; then this step's operand
    RW_STAMP(0);
    if (s > 0) cluster_wait_acquire();                            // the partial sums of step s - 1 have landed
    RW_STAMP(1);
    float dar[8], daz[8], dan[8], dgn[8];
    if (epi) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] += carry[i];
      if (s > 0) {
#pragma unroll
        for (int src = 0; src < 4; ++src) {
          float v[8];
          ld8s(reinterpret_cast<const float*>(rprev) + ((src * UC + j) * 16 + 8 * half), v);
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] += v[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dn = dh[i] * (1.0f - z[i]);
        const float dz = dh[i] * (hp[i] - n[i]);
        dan[i] = dn * (1.0f - n[i] * n[i]);
        daz[i] = dz * z[i] * (1.0f - z[i]);
        dar[i] = dan[i] * ghn[i] * r[i] * (1.0f - r[i]);
        dgn[i] = dan[i] * r[i];
        carry[i] = dh[i] * z[i];
        if constexpr (SUM) { sum_r[i] += dar[i]; sum_z[i] += daz[i]; sum_n[i] += dan[i]; }
      }
    }
    if (ew) epi_bar_sync();                                       // everyone has consumed rprev before it becomes the operand
    uint4 rhi, rlo, zhi, zlo;
    if (epi) {
      uint4 ghi, glo;
      rw_transpose_pack(dar, lane, rhi, rlo);
      rw_transpose_pack(daz, lane, zhi, zlo);
      rw_transpose_pack(dgn, lane, ghi, glo);
      const int kq = 16 * q + 8 * oct;                            // operand k = gate * UC + unit inside the slice
      *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, kq)) = rhi;
      *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, kq)) = rlo;
      *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, UC + kq)) = zhi;
      *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, UC + kq)) = zlo;
      *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, 2 * UC + kq)) = ghi;
      *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, 2 * UC + kq)) = glo;
      fence_proxy_async_smem();
    }
    RW_STAMP(2);
    __syncthreads();
    if (warp == 4) {
      if (elect_one()) {
        if (s == 0) mbar_wait_b(wbar, 0, 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint64_t dA = make_desc(smem_u32(sW)), dB = make_desc(smem_u32(rprev));
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t ao = (m * NKB + (ks >> 2)) * RW_ATILE + (ks & 3) * 2 * ATOM_BYTES;
            const uint32_t bo = (ks >> 2) * RW_BTILE + (ks & 3) * 2 * ATOM_BYTES;
            if (ks == 0) umma_bf16_c<0>(tmem + m * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
            else umma_bf16_c<1>(tmem + m * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
          }
          umma_commit(&done[m]);
        }
        RW_STAMP_MMA(3);
      }
      __syncwarp();
    }
    // ---- outputs that do not need the MMA (weight-gradient GEMMs, dx of the layer below) ----
    if (epi) {
      const long o = (long)t * bpad + b0;
      st8(d.dgi + (long)u * d.dg_ld + o, dar);
      st8(d.dgi + (long)(H + u) * d.dg_ld + o, daz);
      st8(d.dgi + (long)(2 * H + u) * d.dg_ld + o, dan);
      st8(d.dgh + (long)u * d.dg_ld + o, dar);
      st8(d.dgh + (long)(H + u) * d.dg_ld + o, daz);
      st8(d.dgh + (long)(2 * H + u) * d.dg_ld + o, dgn);
      if (d.dgi_p) {
        uint4 nhi, nlo;
        rw_transpose_pack(dan, lane, nhi, nlo);
        constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
        const long row = (long)blockIdx.y * 16 + nrow;
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_p) + (size_t)t * d.dgi_p_slot_elems +
                              (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
        const int rr = (int)(row & 127);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int k = g * H + (int)c * UC + 16 * q + 8 * oct;
          __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
          const int off = p16_in_tile(rr, k & 63);
          *reinterpret_cast<uint4*>(tl + off) = g == 0 ? rhi : (g == 1 ? zhi : nhi);
          *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = g == 0 ? rlo : (g == 1 ? zlo : nlo);
        }
      }
    }
    // ---- partial sums of dh_{t-1}: tile m / lane quarter q holds input units 64 m + 16 q .. + 15 -> push to their owner ----
    if (ew) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mbar_wait_warp(&done[m], ph, 5);
        tc_fence_after();
        float part[8];
        rw_reduce32(taddr + m * 32, half, part);
        const int ui = 64 * m + 16 * q;                           // first input unit of this warp's lanes
        const uint32_t owner = (uint32_t)(ui / UC);
        const int ul = ui % UC + (lane & 15);
        const uint32_t off = smem_u32(sR) + ph * RSLOT + (uint32_t)((((int)c * UC + ul) * 16 + 8 * half) * 4);
        const uint32_t ra = mapa_u32(off, owner);
        st_cluster_f4(ra, part[0], part[1], part[2], part[3]);
        st_cluster_f4(ra + 16, part[4], part[5], part[6], part[7]);
      }
      tc_fence_before();
    }
    RW_STAMP(4);
    cluster_arrive_release();
    RW_STAMP(5);
    if (epi && s + 1 < steps) load_step(s + 1);
    RW_STAMP(6);
  }
  // gradient of the initial state: carry + the partial sums of the last step
  cluster_wait_acquire();
  if (epi) {
    const uint8_t* rl = sR + ((steps - 1) & 1) * RSLOT;
    float g0[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g0[i] = carry[i];
#pragma unroll
    for (int src = 0; src < 4; ++src) {
      float v[8];
      ld8s(reinterpret_cast<const float*>(rl) + ((src * UC + j) * 16 + 8 * half), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) g0[i] += v[i];
    }
    st8(d.dh0_out + (long)u * bpad + b0, g0);
    if constexpr (SUM) {
      st8(d.dgi_sum + (long)u * bpad + b0, sum_r);
      st8(d.dgi_sum + (long)(H + u) * bpad + b0, sum_z);
      st8(d.dgi_sum + (long)(2 * H + u) * bpad + b0, sum_n);
      constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
      const long row = (long)blockIdx.y * 16 + nrow;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_sum_p) + (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
      const int rr = (int)(row & 127);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        uint4 hi, lo;
        rw_transpose_pack(g == 0 ? sum_r : (g == 1 ? sum_z : sum_n), lane, hi, lo);
        const int k = g * H + (int)c * UC + 16 * q + 8 * oct;
        __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
        const int off = p16_in_tile(rr, k & 63);
        *reinterpret_cast<uint4*>(tl + off) = hi;
        *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = lo;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

// =================================================================================================
// version 2 (H = 256 only: the CTA's unit slice is exactly one 64-wide k chunk of the operand)
// =================================================================================================
// Same decomposition, but the cluster never executes barrier.cluster inside the sweep:
//  * forward: a CTA writes its slice of h_t (one 4 KB k chunk, hi and lo rows) into its OWN operand buffer with plain
//    st.shared and into the three peers' buffers with st.async (16-byte asynchronous DSMEM stores that complete tx-bytes on
//    the peer's mbarrier: no release fence, no acknowledgement on the push path); the MMA lane starts the MMAs of its own
//    chunk at once and issues the MMAs over the peers' chunks when their 12 KB have landed.  The element-wise warps do
//    the global stores of step t while the exchange and the MMAs of step t + 1 run.
//  * backward: accumulator tiles are issued peers-first; each tile's partial sums are pushed (st.shared::cluster) as soon as
//    that tile completes, again with st.async + complete_tx on the owner's mbarrier; the CTA's own tile stays in registers.
// Ordering without the cluster barrier: a CTA can only be one step ahead of a peer (it needs the peer's data of step t
// to finish step t + 1), which together with the double-buffered operand / receive slots rules out overwrites of live data.
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// asynchronous 16-byte store into a peer's shared memory that completes 16 tx-bytes on the peer's mbarrier: data and
// signal travel together, the issuing thread never waits for an acknowledgement (no release fence on the push path)
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, const uint4& v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst_cluster),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void st_async_f4(uint32_t dst_cluster, float a, float b, float c, float d, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst_cluster),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar_cluster)
               : "memory");
}
// private lane-major interchange layout of the H = 256 kernels: per (t, 16-row group g, CTA c, warp q) one 1 KB block
// [lane: 32][8 floats] -> one 256-bit access per thread, a warp instruction moves 1 KB of contiguous memory
__device__ __forceinline__ long pv_block(long t, long groups, long g, uint32_t c, int q) { return (((t * groups + g) * 4 + c) * 4 + (q & 3)) * 256; }
__device__ __forceinline__ void ld8p(const float* blk, int lane, float* v) { ld8(blk + lane * 8, v); }
__device__ __forceinline__ void st8p(float* blk, int lane, const float* v) { st8(blk + lane * 8, v); }
// 8 consecutive k (= batch rows) of row `row` of a transposed P16 operand [rows, K]: one atom row per plane
__device__ __forceinline__ void st8_T_p16(void* base, long nk, int row, long k, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(base) + ((size_t)(row >> 7) * nk + (k >> 6)) * p16_tile_elems(128);
  const int off = p16_in_tile(row & 127, (int)(k & 63));
  *reinterpret_cast<uint4*>(tl + off) = hi;
  *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = lo;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// relaxed variant: a pure "I am done reading" signal that publishes no data of this thread.  The release form made the warp
// wait for all of its earlier stores (st.async pushes, global stores): ~1.3 us per step on the slowest CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster_b(uint64_t* bar, uint32_t parity, int site = 0) {
  for (int i = 0; i < (1 << 20); ++i) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    if ((i & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(&g_rw_timeouts) != 0) return;
  }
  rw_timeout(site, parity);
}

// Warp-specialised register allocation (setmaxnreg, per warpgroup of 4 warps): the kernel is launched with the register count
// that its thread count allows; the MMA / store warpgroups then give registers back and the gate-math warpgroups take them.
// Why it matters beyond speed: tcgen05.ld writes its destination registers asynchronously; the variants that ptxas could only
// fit with spills (13 warps = 128 registers per thread) raised illegal-address faults on the B200 that vanished under
// compute-sanitizer - no spills, no faults (profiles/r2_ng2_sw_fault.md).
// a value the compiler cannot see through (blocks loop-invariant hoisting of everything derived from it)
__device__ __forceinline__ uint64_t opaque64(uint64_t v) {
  asm volatile("" : "+l"(v));
  return v;
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// TMEM <-> registers, 8 consecutive 32-bit columns of this thread's lane (staging of the store warps, see SW below)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// SW ("store warps", training sweeps in the private layouts): an SM drains its global stores at only ~18 B/cycle, so the 28 KB a
// CTA writes per time step kept the four gate-math warps stuck in their store instructions for ~1 us of a 3.2 us step
// (measured with the stores removed: 2.3 us).  With SW the gate-math warps hand h_t and the saved gates to four extra warps
// through TENSOR MEMORY (tcgen05.st into spare columns, double-buffered, mbarrier handshake; warp 5 + i owns the TMEM lane
// quarter (5 + i) % 4) and go straight back to the recurrence; the store warps do the bf16 splits / transposes and all global
// stores of step t while step t + 1 is being computed.
// NG: 16-row groups per cluster.  With NG = 2 a cluster serves two INDEPENDENT groups of 16 batch rows that are interleaved in
// time: while the gate-math warps of group A drain the accumulators, do the gate math and exchange h_t (per group and step:
// ~0.4 us of tensor-memory reads at 32 B/cycle per lane quarter, ~0.5 us for 12 KB of incoming DSMEM at ~15 B/cycle per SM -
// tools/bench_mma/tmem_dsmem_bench.cu), the MMA lane issues the 48 MMAs of group B (1.1 us at the 46-cycle floor).  A 512-window
// batch then runs in ONE wave of 32 clusters.  (Batching both groups into N = 64 MMAs was measured first: 5.6 us per step,
// because everything after the MMAs scales with the rows - profiles/r2_probe_ng2_batched.log.)  Every group has its own
// gate-math warps, accumulator columns, operand buffer and barriers; with NG = 2 the operand buffers are single (2 x 2 x 16 KB
// do not fit beside the 192 KB of weights), so a CTA pushes h_t to a peer only after that peer has signalled that its MMAs of
// step t are complete (ofree).
__host__ __device__ constexpr int rw2_fwd_threads(bool sw, int ng) { return ng == 2 ? (sw ? 512 : 384) : (sw ? 288 : 160); }
template <bool PRIV, bool SW, int NG>
__global__ void __launch_bounds__(rw2_fwd_threads(SW, NG), 1) gru_rw2_fwd_kernel(const GruSeqFwdArgs a) {
  static_assert(PRIV || !SW, "store warps serve the private layouts only");
  static_assert(NG == 1 || NG == 2, "one or two 16-row groups per cluster");
  constexpr bool RW_MN = RW_MN_FWD != 0;
  constexpr int NKC = 4, H = 256, UC = 64;
  constexpr int NGW = 4 * NG;                                      // gate-math warps
  constexpr int NBUF = NG == 1 ? 2 : 1;                            // operand buffers per group
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                                             // [3 gates][NKC][RW_ATILE]
  uint8_t* sH = smem + 3 * NKC * RW_ATILE;                        // [NG][NBUF][NKC][RW_BTILE]
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sH + NG * NBUF * NKC * RW_BTILE);
  uint64_t* done = wbar + 1;                                      // [NG][3]: accumulator of gate g complete
  uint64_t* lfull = done + 3 * NG;                                // [NG][2]: own chunk of operand buffer b written (128 arrivals)
  uint64_t* hfull = lfull + 2 * NG;                               // [NG][2]: the three peers' chunks of buffer b have landed (tx bytes)
  uint64_t* sfull = hfull + 2 * NG;                               // [NG][2] SW: staging buffer b written by the 128 gate-math threads
  uint64_t* sempty = sfull + 2 * NG;                              // [NG][2] SW: staging buffer b read back by the 128 store threads
  uint64_t* gfull = sempty + 2 * NG;                              // [NG][3] SW: gi of step s (TMEM buffer s % 3) written by the store threads
  uint64_t* gempty = gfull + 3 * NG;                              // [NG][3] SW: ... consumed by the gate-math threads
  uint64_t* ofree = gempty + 3 * NG;                              // [NG] NBUF == 1: the 3 peers' MMAs of step s are complete (phase s)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofree + NG);

  const uint32_t c = cluster_ctarank();
  const GruSeqDirFwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp: provably uniform
  const int q = warp & 3, half = lane >> 4, oct = (lane >> 3) & 1;  // q: TMEM lane quarter of this warp = its 16 units of the slice
  const bool epi = warp < NGW;                                     // gate-math warp of group warp >> 2
  // NG = 2: warp-specialised register allocation, whole warpgroups per role: warps 0-7 gate math, 8 MMA issue (9-11 idle),
  // SW: 12-15 store warps
  constexpr bool WS = NG == 2;
  constexpr int STW0 = WS ? 12 : NGW + 1;
  const bool stw = SW && warp >= STW0;                             // store warp of lane quarter q (serves every group in turn)
  const int grp = epi ? (warp >> 2) : 0;
  const long gidx = (long)blockIdx.y * NG + grp;                   // 16-row group of the batch (gate-math warps)
  const long Bp = (long)a.tiles * 128;
  const int steps = a.steps;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int p = 0; p < NG; ++p) {
      for (int g = 0; g < 3; ++g) mbar_init(&done[3 * p + g], 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&lfull[2 * p + b], 128);
        mbar_init(&hfull[2 * p + b], 1);
        mbar_init(&sfull[2 * p + b], 128);
        mbar_init(&sempty[2 * p + b], 128);
      }
      for (int b = 0; b < 3; ++b) {
        mbar_init(&gfull[3 * p + b], 128);
        mbar_init(&gempty[3 * p + b], 128);
      }
      mbar_init(&ofree[p], 3);
    }
    mbar_fence_init();
  }
  // tensor-memory columns of group p at 248 p: accumulators of gate g at 32 g; SW staging (h, r, z, n, gh_n: 40 columns), buffer b
  // at 96 + 40 b; SW input projections (24 columns), buffer b at 176 + 24 b
  constexpr uint32_t GCOLS = 248, SBASE = 96, GBASE = 176;
  constexpr uint32_t TCOLS = SW ? 256 * NG : 128 * NG;
  if (warp == NGW) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == NGW) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.w_rw) + (size_t)c * 3 * NKC * (RW_ATILE / 2);
      mbar_expect_tx(wbar, 3 * NKC * RW_ATILE);
      for (int i = 0; i < 3 * NKC; ++i) bulk_g2s(sW + (size_t)i * RW_ATILE, wp + (size_t)i * (RW_ATILE / 2), RW_ATILE, wbar);
    }
    __syncwarp();
  }

  const int j = 16 * q + (lane & 15);
  const int u = (int)c * UC + j;
  const long b0 = gidx * 16 + 8 * half;
  const int nrow = 8 * half + (lane & 7);                         // row inside the group after the transpose
  const int kloc = 16 * q + 8 * oct;                              // k inside the CTA's own chunk after the transpose
  uint8_t* sHg = sH + (size_t)grp * NBUF * NKC * RW_BTILE;         // this group's operand buffers
  const uint32_t tcol = (SW ? GCOLS : 96u) * (uint32_t)grp;        // this group's tensor-memory columns
  float hprev[8], bhn = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) hprev[i] = 0.f;
  if (epi) {
    bhn = d.b_hn[u];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {                               // initial state: every CTA builds the whole operand itself
      float v[8];
      ld8(d.h0 + (long)(cc * UC + j) * d.h0_ld + b0, v);
      if (cc == (int)c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hprev[i] = v[i];
      }
      uint4 hi, lo;
      if constexpr (RW_MN) {
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(sHg + rw_mn_off(8 * half, cc * UC + j)) = hi;
        *reinterpret_cast<uint4*>(sHg + rw_mn_off(16 + 8 * half, cc * UC + j)) = lo;
      } else {
        rw_transpose_pack(v, lane, hi, lo);
        *reinterpret_cast<uint4*>(sHg + rw_b_off(nrow, cc * UC + kloc)) = hi;
        *reinterpret_cast<uint4*>(sHg + rw_b_off(16 + nrow, cc * UC + kloc)) = lo;
      }
    }
    fence_proxy_async_smem();
  }
  // arm the receive barriers of the first exchanges (two buffers: buffer 1 is filled for step 1, buffer 0 for step 2; one buffer:
  // for step 1); later phases are armed right after the previous phase has been waited for, so a peer's complete_tx never
  // precedes the expect_tx
  const uint32_t hx = (a.exp & 8) ? 3 * RW_BTILE / 2 : 3 * RW_BTILE;   // (experiment 8: only the hi plane is pushed)
  if (tid == 0) {
    for (int p = 0; p < NG; ++p) {
      if constexpr (NBUF == 2) {
        if (steps > 1) mbar_expect_tx(&hfull[2 * p + 1], hx);
        if (steps > 2) mbar_expect_tx(&hfull[2 * p + 0], hx);
      } else {
        if (steps > 1) mbar_expect_tx(&hfull[2 * p], hx);
      }
    }
  }
  // all mbarriers of the cluster are initialised (and the h0 operands written) before any remote signal is sent
  cluster_arrive_release();
  cluster_wait_acquire();

  const uint32_t idesc = rw_idesc(RW_MN);
  const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);        // this warp's lane quarter, column 0
  const uint32_t taddr = tlane + tcol;                             // ... at this group's columns (gate-math warps)
  const long groups = Bp / 16;
  const bool gi_const = d.gi_ts == 0;                              // decoders: the input projection does not depend on t
  // input projections are prefetched TWO steps ahead (registers): under the write bursts of the sweep's own stores a load
  // issued one step ahead came back too late for the next gate math (measured: 3.25 -> 2.5 us per step)
  float gir[8], giz[8], gin[8], gir2[8], giz2[8], gin2[8];
  auto load_gi = [&](int g_, int s_, float* r_, float* z_, float* n_) {
    const int t_ = d.reverse ? steps - 1 - s_ : s_;
    const float* gi_row = d.gi + ((((long)blockIdx.y * NG + g_) * 16 + 8 * half) * d.gi_bs + (long)t_ * d.gi_ts);
    ld8(gi_row + (long)u * d.gi_ld, r_);          // (ld.global.cg instead: 3.9 instead of 3.3 us per step - the second
    ld8(gi_row + (long)(H + u) * d.gi_ld, z_);    //  16 bytes of each sector are L1 hits with the default policy)
    ld8(gi_row + (long)(2 * H + u) * d.gi_ld, n_);
  };
  // SW && !gi_const: the store warps fetch gi(s + 2) during step s and pass it on through tensor memory, so that the gate-math
  // warps issue no global memory instruction at all
  const bool gi_tmem = SW && !gi_const;
  auto gi_col = [&](int g_, int s_) -> uint32_t { return tlane + GCOLS * (uint32_t)g_ + GBASE + 24u * (uint32_t)(s_ % 3); };
  auto gi_publish = [&](int g_, int s_, const float* r_, const float* z_, const float* n_) {   // store warps
    const uint32_t gc = gi_col(g_, s_);
    __syncwarp();
    tmem_st8(gc, r_);
    tmem_st8(gc + 8, z_);
    tmem_st8(gc + 16, n_);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&gfull[3 * g_ + s_ % 3]);
  };
  if (epi && !gi_tmem) {
    load_gi(grp, 0, gir, giz, gin);
    if constexpr (!SW) {
      if (steps > 1 && !gi_const) load_gi(grp, 1, gir2, giz2, gin2);
    }
  }
  if (stw && gi_tmem) {
#pragma unroll
    for (int g_ = 0; g_ < NG; ++g_) {
      load_gi(g_, 0, gir, giz, gin);
      if (steps > 1) load_gi(g_, 1, gir2, giz2, gin2);
      gi_publish(g_, 0, gir, giz, gin);
      if (steps > 1) gi_publish(g_, 1, gir2, giz2, gin2);
    }
  }

  // one loop per warp role: the register budget of a role is set by setmaxnreg before its loop (warp-specialised allocation)
  if (warp >= NGW && !stw) {
    if constexpr (WS) reg_dec<72>();
    if (warp == NGW)
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? steps - 1 - s : s;
      const int so = (d.out_slots == steps) ? t : (s & 1);
      const int sp = (d.out_p_slots == steps) ? t : (s & 1);
      const uint32_t ph = s & 1;
      const uint32_t b = NBUF == 2 ? (uint32_t)(s & 1) : 0u;         // operand buffer of this step
      const uint32_t nb = NBUF == 2 ? (b ^ 1u) : 0u;                 // ... of the next step
      // phase parity of the barriers of buffer b: two buffers - used every other step; one buffer - every step
      const uint32_t bpar = NBUF == 2 ? (uint32_t)(((s - 1) >> 1) & 1) : (uint32_t)((s - 1) & 1);
      RW_STAMP(0);
      if (elect_one()) {
#pragma unroll
        for (int p = 0; p < NG; ++p) {                              // the groups take turns on the tensor pipe
          const uint32_t hb = smem_u32(sH) + ((uint32_t)p * NBUF + b) * NKC * RW_BTILE;
          const uint32_t dcol = tmem + (SW ? GCOLS : 96u) * (uint32_t)p;
          if (s == 0) {
            if (p == 0) mbar_wait_b(wbar, 0, 10);
          } else {
            mbar_wait_b(&lfull[2 * p + b], bpar, 11);               // own chunk of h_{t-1} written (and TMEM drained) by the gate warps
            if (p == 0) RW_STAMP_MMA(8);
          }
          fence_proxy_async_smem();
          tc_fence_after();
          // (opaque copies: without them the compiler hoists all 64 operand descriptors of a step out of the time loop - 128 registers)
          const uint64_t dA = opaque64(make_desc(smem_u32(sW))), dB = opaque64(make_desc(hb));
          // own chunk first (all three gates), then the peers' chunks gate by gate
#pragma unroll
          for (int g = 0; g < 3; ++g) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ao = (g * NKC + c) * RW_ATILE + ks * 2 * ATOM_BYTES, bo = c * RW_BTILE + ks * 2 * ATOM_BYTES;
              if (ks == 0) umma_bf16_c<0>(dcol + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
              else umma_bf16_c<1>(dcol + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
            }
          }
          if (s > 0) {
            mbar_wait_b(&hfull[2 * p + b], bpar, 12);
            if (s + NBUF < steps) mbar_expect_tx(&hfull[2 * p + b], hx);   // arm the next use of this buffer (step s + NBUF)
            fence_proxy_async_smem();                               // the peers' chunks were written through the generic proxy
          }
          if (p == 0) {
            RW_STAMP_MMA(9);
            RW_STAMP_CTA(0);
          }
#pragma unroll
          for (int g = 0; g < 3; ++g) {
#pragma unroll
            for (uint32_t r = 1; r < 4; ++r) {
              const uint32_t kc = (c + r) & 3;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ao = (g * NKC + kc) * RW_ATILE + ks * 2 * ATOM_BYTES, bo = kc * RW_BTILE + ks * 2 * ATOM_BYTES;
                umma_bf16_c<1>(dcol + g * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
              }
            }
            umma_commit(&done[3 * p + g]);
          }
          if constexpr (NBUF == 1) {
            // every MMA of this group's step has finished reading the operand buffer once this commit fires: the tensor pipe tells
            // the three peers that they may push h_t into it (no gate-math thread has to notice the completion first)
            if (s + 1 < steps) umma_commit_multicast(&ofree[p], (uint16_t)(0xFu & ~(1u << c)));
          }
          if (p == 0) RW_STAMP_MMA(2);
        }
      }
      __syncwarp();
    }
  } else if (stw) {
    if constexpr (WS) reg_dec<104>();
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? steps - 1 - s : s;
      const int so = (d.out_slots == steps) ? t : (s & 1);
      const int sp = (d.out_p_slots == steps) ? t : (s & 1);
      const uint32_t ph = s & 1;
      const uint32_t b = NBUF == 2 ? (uint32_t)(s & 1) : 0u;         // operand buffer of this step
      const uint32_t nb = NBUF == 2 ? (b ^ 1u) : 0u;                 // ... of the next step
      // phase parity of the barriers of buffer b: two buffers - used every other step; one buffer - every step
      const uint32_t bpar = NBUF == 2 ? (uint32_t)(((s - 1) >> 1) & 1) : (uint32_t)((s - 1) & 1);
      RW_STAMP(0);
      const uint32_t sb = s & 1;
      const bool fetch = gi_tmem && s + 2 < steps;
#pragma unroll 1
      for (int g_ = 0; g_ < NG; ++g_) {
        float hn[8], sr[8], sz[8], sn[8], sg[8];
        const uint32_t stg = tlane + GCOLS * (uint32_t)g_ + SBASE + 40u * sb;
        if (fetch) load_gi(g_, s + 2, gir, giz, gin);              // loads first: they are in flight while the warp waits below
        mbar_wait_warp(&sfull[2 * g_ + sb], (s >> 1) & 1, 17);
        tc_fence_after();
        tmem_ld8(stg, hn);
        tmem_ld8(stg + 8, sr);
        tmem_ld8(stg + 16, sz);
        tmem_ld8(stg + 24, sn);
        tmem_ld8(stg + 32, sg);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sempty[2 * g_ + sb]);
        if (!(a.exp & 16)) {
          // (pacing these stores - one per warp every 256-384 cycles - changed nothing: 3.04 us per step either way)
          uint4 phi, plo;
          rw_transpose_pack(hn, lane, phi, plo);
          const long gi_ = (long)blockIdx.y * NG + g_, bb = gi_ * 16 + 8 * half, row = gi_ * 16 + nrow;
          const int k0 = (int)c * UC + kloc;
          __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(d.out_p) + (size_t)sp * d.out_p_slot_elems +
                              ((size_t)(row >> 7) * NKC + (k0 >> 6)) * p16_tile_elems(128);
          const int off = p16_in_tile((int)(row & 127), k0 & 63);
          const long blk = pv_block(t, groups, gi_, c, q);
          st8p(d.out + blk, lane, hn);
          st8p(d.sv[0] + blk, lane, sr);
          st8p(d.sv[1] + blk, lane, sz);
          st8p(d.sv[2] + blk, lane, sn);
          st8p(d.sv[3] + blk, lane, sg);
          st8_T_p16(d.outT_p, d.outT_nk, u, (long)t * Bp + bb, hn);
          if (s + 1 == steps) st8(d.hfin + (long)u * Bp + bb, hn);
          *reinterpret_cast<uint4*>(tl + off) = phi;
          *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = plo;
        }
        if (fetch) {                                                // buffer (s + 2) % 3 was last read in step s - 1
          if (s >= 1) mbar_wait_warp(&gempty[3 * g_ + (s + 2) % 3], (((s + 2) / 3) - 1) & 1, 18);
          tc_fence_after();
          gi_publish(g_, s + 2, gir, giz, gin);
        }
      }
    }
  } else if (epi) {
    if constexpr (WS) reg_inc<SW ? 168 : 216>();
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? steps - 1 - s : s;
      const int so = (d.out_slots == steps) ? t : (s & 1);
      const int sp = (d.out_p_slots == steps) ? t : (s & 1);
      const uint32_t ph = s & 1;
      const uint32_t b = NBUF == 2 ? (uint32_t)(s & 1) : 0u;         // operand buffer of this step
      const uint32_t nb = NBUF == 2 ? (b ^ 1u) : 0u;                 // ... of the next step
      // phase parity of the barriers of buffer b: two buffers - used every other step; one buffer - every step
      const uint32_t bpar = NBUF == 2 ? (uint32_t)(((s - 1) >> 1) & 1) : (uint32_t)((s - 1) & 1);
      RW_STAMP(0);
      float hn[8], sr[8], sz[8], sn[8], sg[8], ar[8], az[8], an[8];
      uint4 phi = make_uint4(0, 0, 0, 0), plo = make_uint4(0, 0, 0, 0);
      if (gi_tmem) {                                                // this step's input projections (published >= 1 step ago)
        mbar_wait_warp(&gfull[3 * grp + s % 3], (s / 3) & 1, 19);
        tc_fence_after();
        const uint32_t gc = gi_col(grp, s);
        tmem_ld8(gc, gir);
        tmem_ld8(gc + 8, giz);
        tmem_ld8(gc + 16, gin);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&gempty[3 * grp + s % 3]);
      }
      mbar_wait_warp(&done[3 * grp + 0], ph, 13);
      tc_fence_after();
      rw_reduce32(taddr, half, ar);
#pragma unroll
      for (int i = 0; i < 8; ++i) sr[i] = rw_sigmoid(gir[i] + ar[i]);
      mbar_wait_warp(&done[3 * grp + 1], ph, 14);
      tc_fence_after();
      rw_reduce32(taddr + 32, half, az);
#pragma unroll
      for (int i = 0; i < 8; ++i) sz[i] = rw_sigmoid(giz[i] + az[i]);
      mbar_wait_warp(&done[3 * grp + 2], ph, 15);
      tc_fence_after();
      RW_STAMP(3);
      if (tid == 0) RW_STAMP_CTA(1);
      rw_reduce32(taddr + 64, half, an);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sg[i] = an[i] + bhn;
        sn[i] = rw_tanh(gin[i] + sr[i] * sg[i]);
        hn[i] = (1.0f - sz[i]) * sn[i] + sz[i] * hprev[i];
        hprev[i] = hn[i];
      }
      RW_STAMP(4);
      uint4 ohv, olv;                                               // this thread's operand vectors (hi / lo plane)
      if constexpr (RW_MN) split8(hn, ohv, olv);
      else {
        rw_transpose_pack(hn, lane, phi, plo);
        ohv = phi; olv = plo;
      }
      tc_fence_before();
      if (s + 1 < steps) {                                        // own chunk of the next operand buffer, then tell the MMA lane
        uint8_t* own = sHg + nb * NKC * RW_BTILE + c * RW_BTILE;
        const uint32_t o_hi = RW_MN ? rw_mn_off(8 * half, j) : 2u * p16_in_tile(nrow, kloc);
        const uint32_t o_lo = RW_MN ? rw_mn_off(16 + 8 * half, j) : 2u * p16_in_tile(16 + nrow, kloc);
        const uint32_t ohi = smem_u32(own) + o_hi, olo = smem_u32(own) + o_lo;
        const uint32_t hbar = smem_u32(&hfull[2 * grp + nb]);
        if constexpr (NBUF == 1) {                                  // the peers' MMAs of this step are complete
          mbar_wait_cluster_b(&ofree[grp], ph, 30);
          __syncwarp();
        }
#pragma unroll
        for (uint32_t r = 1; r < 4; ++r) {                          // the same chunk position in the three peers' buffers
          const uint32_t peer = (c + r) & 3, pbar = mapa_u32(hbar, peer);
          st_async_v4(mapa_u32(ohi, peer), ohv, pbar);
          if (!(a.exp & 8)) st_async_v4(mapa_u32(olo, peer), olv, pbar);
        }
        *reinterpret_cast<uint4*>(own + o_hi) = ohv;
        *reinterpret_cast<uint4*>(own + o_lo) = olv;
        fence_proxy_async_smem();
        mbar_arrive(&lfull[2 * grp + nb]);
      }
      if constexpr (RW_MN && !SW) rw_transpose_pack(hn, lane, phi, plo);   // K-major P16 copy for the GEMMs (off the recurrence)
      RW_STAMP(5);
      if (tid == 0) RW_STAMP_CTA(2);
      if (tid == 96) RW_STAMP_CTA(3);
      // ---- off the recurrence (overlaps the exchange and the next step's MMAs); loads first: the LSU works in order ----
      if constexpr (!SW) {                                          // (SW: the store warps deliver gi through tensor memory)
        if (s + 1 < steps && !gi_const && !(a.exp & 2)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { gir[i] = gir2[i]; giz[i] = giz2[i]; gin[i] = gin2[i]; }
          if (s + 2 < steps) load_gi(grp, s + 2, gir2, giz2, gin2);
        }
      }
      if constexpr (SW) {                                           // hand h_t and the saved gates to the store warps
        const uint32_t sb = s & 1, stg = taddr + SBASE + 40u * sb;
        if (s >= 2) mbar_wait_warp(&sempty[2 * grp + sb], ((s - 2) >> 1) & 1, 16);
        __syncwarp();                                               // tcgen05.st is .sync.aligned: the warp must be converged
        tc_fence_after();
        tmem_st8(stg, hn);
        tmem_st8(stg + 8, sr);
        tmem_st8(stg + 16, sz);
        tmem_st8(stg + 24, sn);
        tmem_st8(stg + 32, sg);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&sfull[2 * grp + sb]);
        continue;
      }
      if constexpr (PRIV) {
        if (a.exp & 16) continue;                                   // (experiment 16: no global stores)
        const long blk = pv_block(t, groups, gidx, c, q);
        st8p(d.out + blk, lane, hn);
        st8p(d.sv[0] + blk, lane, sr);
        st8p(d.sv[1] + blk, lane, sz);
        st8p(d.sv[2] + blk, lane, sn);
        st8p(d.sv[3] + blk, lane, sg);
        st8_T_p16(d.outT_p, d.outT_nk, u, (long)t * Bp + b0, hn);
        if (s + 1 == steps) st8(d.hfin + (long)u * Bp + b0, hn);
      } else if (!(a.exp & 1)) {
        st8(d.out + (long)u * d.out_ld + (long)so * Bp + b0, hn);
      }
      if (!(a.exp & 4)) {
        const long row = gidx * 16 + nrow;
        const int k0 = (int)c * UC + kloc;
        __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(d.out_p) + (size_t)sp * d.out_p_slot_elems +
                            ((size_t)(row >> 7) * NKC + (k0 >> 6)) * p16_tile_elems(128);
        const int off = p16_in_tile((int)(row & 127), k0 & 63);
        *reinterpret_cast<uint4*>(tl + off) = phi;
        *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = plo;
      }
      if (!PRIV && d.sv[0] && !(a.exp & 1)) {
        const long o = (long)u * d.sv_ld + (long)t * Bp + b0;
        st8(d.sv[0] + o, sr);
        st8(d.sv[1] + o, sz);
        st8(d.sv[2] + o, sn);
        st8(d.sv[3] + o, sg);
      }
      RW_STAMP(7);
    }
  }
  // no CTA leaves while a peer may still be reading its shared memory (outgoing bulk copies) or signalling its barriers
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_before();
  __syncthreads();
  if (warp == NGW) tmem_dealloc(tmem, TCOLS);
}

// SW: as in the forward kernel, four extra warps take over everything that is not on the recurrence - here the gate gradients
// of step t travel through tensor memory (32 columns, double-buffered) and the store warps write the transposed P16 operands of
// the weight-gradient GEMMs, the P16 copy for dx, and keep the bias-gradient / time sums.
// NG = 2 (see the forward kernel): two independent 16-row groups per cluster, interleaved in time - group B's 48 MMAs run while
// group A waits for its partial sums, does the gate-gradient math and builds its operand.  Four 12 KB receive slots no longer
// fit beside 192 KB of weights, so the accumulator tile that is issued last (the CTA's own input units) takes its A operand
// from TENSOR MEMORY: its 48 KB of W_hh^T (hi + lo rows x 192 k) are copied global -> registers -> tcgen05.st once before the
// sweep (96 columns, column 8 ks + i of lane r = k elements 16 ks + 2 i, + 1 of tile row r) and only three tiles stay in
// shared memory.
__host__ __device__ constexpr int rw2_bwd_threads(bool sw, int ng) { return ng == 2 ? (sw ? 512 : 384) : (sw ? 288 : 160); }
template <bool SUM, bool PRIV, bool SW, int NG>
__global__ void __launch_bounds__(rw2_bwd_threads(SW, NG), 1) gru_rw2_bwd_kernel(const GruSeqBwdArgs a) {
  static_assert(PRIV || !SW, "store warps serve the private layouts only");
  static_assert(NG == 1 || NG == 2, "one or two 16-row groups per cluster");
  static_assert(!(SW && NG == 2 && SUM), "two-group store warps do not keep the time sums (decoders use the variant without store warps)");
  constexpr bool RW_MN = RW_MN_BWD != 0;
  constexpr int H = 256, UC = 64, MT = 4, KS = 12, NKB = 3;
  constexpr int NGW = 4 * NG;                                      // gate-math warps
  constexpr int RSLOT = 3 * UC * 16 * 4;                           // 12288 B: 3 source slots = the 3 k chunks of the operand
  static_assert(RSLOT == NKB * RW_BTILE, "receive slot and operand must have the same size");
  constexpr bool WT = NG == 2;                                     // the own tile's A operand lives in tensor memory
  constexpr int MS = WT ? 3 : 4;                                   // A tiles in shared memory
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                                             // [MS][NKB][RW_ATILE]; WT: tile i = input units of CTA (c + 1 + i) & 3
  uint8_t* sR = smem + MS * NKB * RW_ATILE;                       // [NG][2][RSLOT]
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sR + NG * 2 * RSLOT);
  uint64_t* done = wbar + 1;                                      // [NG][4] accumulator tiles, in issue order
  uint64_t* ofull = done + 4 * NG;                                // [NG] operand of this step written (128 arrivals)
  uint64_t* pfull = ofull + NG;                                   // [NG][2]: partial sums of 3 peers x 4 warps have landed in slot b
  uint64_t* mfree = pfull + 2 * NG;                               // [NG][2]: all MMAs of the 3 peers' step s are complete (3 x 4 warps), slot s & 1
  uint64_t* sfull = mfree + 2 * NG;                               // [NG][2] SW: staging buffer b written by the 128 gate-math threads
  uint64_t* sempty = sfull + 2 * NG;                              // [NG][2] SW: ... read back by the 128 store threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 2 * NG);   // (mfree: a peer may signal step s + 1 before this CTA has looked at step s)

  const uint32_t c = cluster_ctarank();
  const GruSeqDirBwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp: provably uniform
  const int q = warp & 3, half = lane >> 4, oct = (lane >> 3) & 1;
  const bool epi = warp < NGW;
  constexpr bool WS = NG == 2;                                     // warps 0-7 gate math, 8 MMA issue (9-11 idle), SW: 12-15 store warps
  const bool stw = SW && warp >= (WS ? 12 : NGW + 1);               // (see the forward kernel)
  const int grp = epi ? (warp >> 2) : (stw ? ((warp - NGW - 1) >> 2) : 0);
  const long gidx = (long)blockIdx.y * NG + grp;
  const long bpad = (long)a.tiles * 128;
  const int steps = a.steps;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int p = 0; p < NG; ++p) {
      for (int m = 0; m < 4; ++m) mbar_init(&done[4 * p + m], 1);
      mbar_init(&ofull[p], 128);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&pfull[2 * p + b], 1);
        mbar_init(&mfree[2 * p + b], 12);
      }
    }
    for (int b = 0; b < 2 * NG; ++b) {
      mbar_init(&sfull[b], 128);
      mbar_init(&sempty[b], 128);
    }
    mbar_fence_init();
  }
  // tensor-memory columns: accumulator tile i of group p at 128 p + 32 i; WT: own A tile at WCOL + [0, 96); SW staging (32 columns)
  // of group p, buffer b at SBASE + 32 (2 p + b)
  constexpr uint32_t WCOL = 128 * NG, SBASE = WT ? WCOL + 96 : 128;
  constexpr uint32_t TCOLS = WT ? 512 : (SW ? 256 : 128);
  static_assert(!SW || SBASE + 64 * NG <= TCOLS, "tensor-memory budget");
  if (warp == NGW) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);

  const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.wT_rw) + (size_t)c * MT * NKB * (RW_ATILE / 2);
  if (warp == NGW) {
    if (elect_one()) {
      mbar_expect_tx(wbar, MS * NKB * RW_ATILE);
      if constexpr (WT) {
        for (uint32_t i = 0; i < 3; ++i) {
          const uint32_t m = (c + 1 + i) & 3;
          for (int kc = 0; kc < NKB; ++kc)
            bulk_g2s(sW + (size_t)(i * NKB + kc) * RW_ATILE, wp + (size_t)(m * NKB + kc) * (RW_ATILE / 2), RW_ATILE, wbar);
        }
      } else {
        for (int i = 0; i < MT * NKB; ++i) bulk_g2s(sW + (size_t)i * RW_ATILE, wp + (size_t)i * (RW_ATILE / 2), RW_ATILE, wbar);
      }
    }
    __syncwarp();
  }
  if constexpr (WT) {
    if (warp < 4) {                                               // tile row 32 q + lane -> the same tensor-memory lane
      const int r = 32 * q + lane;
      const uint8_t* tile = reinterpret_cast<const uint8_t*>(wp + (size_t)(c * NKB) * (RW_ATILE / 2));
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint8_t* src = tile + (size_t)(ks >> 2) * RW_ATILE + ((r >> 3) * 8 + (ks & 3) * 2) * 128 + (r & 7) * 16;
        const uint4 v0 = *reinterpret_cast<const uint4*>(src), v1 = *reinterpret_cast<const uint4*>(src + 128);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr + WCOL + ks * 8), "r"(v0.x),
                     "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w)
                     : "memory");
      }
      tmem_st_wait();
      tc_fence_before();
    }
  }
  // arm the receive barriers of the first two steps' pushes (3 peers x 64 units x 16 rows x 4 B); later phases are armed by
  // the group's first thread right after it has consumed the previous phase of the same slot, i.e. before any peer can push
  // into it again
  if (tid == 0) {
    for (int p = 0; p < NG; ++p) {
      mbar_expect_tx(&pfull[2 * p], RSLOT);
      if (steps > 1) mbar_expect_tx(&pfull[2 * p + 1], RSLOT);
    }
  }
  if constexpr (WT) __syncthreads();                              // the tensor-memory A tile is complete before the MMA lane starts
  cluster_arrive_release();
  cluster_wait_acquire();

  const int j = 16 * q + (lane & 15);
  const int u = (int)c * UC + j;
  const long b0 = gidx * 16 + 8 * half;
  const int nrow = 8 * half + (lane & 7);
  const int kq = 16 * q + 8 * oct;
  const uint32_t idesc = rw_idesc(RW_MN);
  uint8_t* sRg = sR + (size_t)grp * 2 * RSLOT;                     // this group's two receive slots
  uint64_t* doneg = done + 4 * grp;
  uint64_t* pfullg = pfull + 2 * grp;
  uint64_t* mfreeg = mfree + 2 * grp;
  const uint32_t tacc = taddr + 128u * (uint32_t)grp;             // this group's accumulator tiles (gate-math warps)
  // fp32 partial sums in a receive slot: [source slot: 3][unit: 64][row: 16]
  auto part_off = [&](int slot, int unit) -> uint32_t { return (uint32_t)(((slot * UC + unit) * 16 + 8 * half) * 4); };

  float carry[8], own[8];
  float sum_r[SUM ? 8 : 1], sum_z[SUM ? 8 : 1], sum_n[SUM ? 8 : 1];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};                            // PRIV: bias gradients = sums over t and rows of dar, daz, dan, dgn
#pragma unroll
  for (int i = 0; i < 8; ++i) carry[i] = own[i] = 0.f;
  if constexpr (SUM) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sum_r[i] = sum_z[i] = sum_n[i] = 0.f;
  }
  float r[8], z[8], n[8], ghn[8], hp[8], dh[8];
  auto load_step = [&](int s) {
    const int t = d.reverse ? s : steps - 1 - s;
    const bool first_fwd = d.reverse ? (t == steps - 1) : (t == 0);
    const int tprev = d.reverse ? t + 1 : t - 1;
    if constexpr (PRIV) {
      const long blk = pv_block(t, bpad / 16, gidx, c, q);
      ld8p(d.sv[0] + blk, lane, r);
      ld8p(d.sv[1] + blk, lane, z);
      ld8p(d.sv[2] + blk, lane, n);
      ld8p(d.sv[3] + blk, lane, ghn);
      if (first_fwd) ld8(d.h0 + (long)u * d.h0_ld + b0, hp);
      else ld8p(d.out + pv_block(tprev, bpad / 16, gidx, c, q), lane, hp);
    } else {
      const long so = (long)u * d.sv_ld + (long)t * bpad + b0;
      ld8(d.sv[0] + so, r);
      ld8(d.sv[1] + so, z);
      ld8(d.sv[2] + so, n);
      ld8(d.sv[3] + so, ghn);
      if (first_fwd) ld8(d.h0 + (long)u * d.h0_ld + b0, hp);
      else ld8(d.out + (long)u * d.out_ld + (long)tprev * bpad + b0, hp);
    }
    if (d.dout) {
      if (PRIV && d.dout_pv) ld8p(d.dout + pv_block(t, bpad / 16, gidx, c, q), lane, dh);
      else ld8(d.dout + (long)u * d.dout_ld + (long)t * bpad + b0, dh);
    }
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] = 0.f;
    }
  };
  // one loop per warp role: the register budget of a role is set by setmaxnreg before its loop (warp-specialised allocation)
  // (everything a role does after the common set-up lives inside its branch: ptxas applies a setmaxnreg budget only to code
  //  that the instruction dominates)
  if (warp >= NGW && !stw) {
    // the pool that setmaxnreg.inc draws from holds only RELEASED registers: 4 x (168 - this) must cover 8 x (gate - 168)
    if constexpr (WS) reg_dec<SW ? 56 : (SUM ? 56 : 72)>();        // (SW: 512 threads start with 128; + 104 for the store warps -> gate 176)
    if (warp == NGW)
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? s : steps - 1 - s;
      const uint32_t ph = s & 1;
      uint8_t* rprev = sRg + (ph ^ 1u) * RSLOT;                     // partial sums of step s - 1; then this step's operand
      RW_STAMP(0);
      if (elect_one()) {
#pragma unroll
        for (int p = 0; p < NG; ++p) {                            // the groups take turns on the tensor pipe
          if (s == 0 && p == 0) mbar_wait_b(wbar, 0, 20);
          mbar_wait_b(&ofull[p], ph, 21);
          fence_proxy_async_smem();
          tc_fence_after();
          const uint64_t dA = opaque64(make_desc(smem_u32(sW))), dB = opaque64(make_desc(smem_u32(sR) + ((uint32_t)p * 2 + (ph ^ 1u)) * RSLOT));
          const uint32_t dcol = tmem + 128u * (uint32_t)p;
#pragma unroll
          for (uint32_t i = 0; i < 4; ++i) {                      // peers' tiles first, the own tile last
            const uint32_t m = WT ? i : ((c + 1 + i) & 3);        // WT: shared memory holds the tiles in issue order
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const uint32_t bo = (ks >> 2) * RW_BTILE + (ks & 3) * 2 * ATOM_BYTES;
              if (WT && i == 3) {
                if (ks == 0) umma_bf16_ta<0>(dcol + i * 32, tmem + WCOL + ks * 8, desc_advance(dB, bo), idesc);
                else umma_bf16_ta<1>(dcol + i * 32, tmem + WCOL + ks * 8, desc_advance(dB, bo), idesc);
              } else {
                const uint32_t ao = (m * NKB + (ks >> 2)) * RW_ATILE + (ks & 3) * 2 * ATOM_BYTES;
                if (ks == 0) umma_bf16_c<0>(dcol + i * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
                else umma_bf16_c<1>(dcol + i * 32, desc_advance(dA, ao), desc_advance(dB, bo), idesc);
              }
            }
            umma_commit(&done[4 * p + i]);
          }
          if (p == 0) RW_STAMP_MMA(3);
        }
      }
      __syncwarp();
    }
  } else if (stw) {
    if constexpr (WS) reg_dec<104>();
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? s : steps - 1 - s;
      const uint32_t sb = s & 1;
#pragma unroll 1
      for (int g_ = 0; g_ < NG; ++g_) {                             // the store warps serve the groups in turn
        float dar[8], daz[8], dan[8], dgn[8];
        const uint32_t stg = taddr + SBASE + 32u * (uint32_t)(2 * g_ + (int)sb);
        mbar_wait_warp(&sfull[2 * g_ + sb], (s >> 1) & 1, 26);
        tc_fence_after();
        tmem_ld8(stg, dar);
        tmem_ld8(stg + 8, daz);
        tmem_ld8(stg + 16, dan);
        tmem_ld8(stg + 24, dgn);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sempty[2 * g_ + sb]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if constexpr (SUM) { sum_r[i] += dar[i]; sum_z[i] += daz[i]; sum_n[i] += dan[i]; }
          bsum[0] += dar[i]; bsum[1] += daz[i]; bsum[2] += dan[i]; bsum[3] += dgn[i];
        }
        if (a.exp & 16) continue;
        const long gi_ = (long)blockIdx.y * NG + g_;
        const long o = (long)t * bpad + gi_ * 16 + 8 * half;
        st8_T_p16(d.dghT_p, d.gT_nk, u, o, dar);
        st8_T_p16(d.dghT_p, d.gT_nk, H + u, o, daz);
        st8_T_p16(d.dghT_p, d.gT_nk, 2 * H + u, o, dgn);
        if (d.dgiT_p) {
          st8_T_p16(d.dgiT_p, d.gT_nk, u, o, dar);
          st8_T_p16(d.dgiT_p, d.gT_nk, H + u, o, daz);
          st8_T_p16(d.dgiT_p, d.gT_nk, 2 * H + u, o, dan);
        }
        if (d.dgi_p) {
          constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
          const long row = gi_ * 16 + nrow;
          __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_p) + (size_t)t * d.dgi_p_slot_elems +
                                (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
          const int rr = (int)(row & 127);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            uint4 hi, lo;
            rw_transpose_pack(g == 0 ? dar : (g == 1 ? daz : dan), lane, hi, lo);
            const int k = g * H + (int)c * UC + kq;
            __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
            const int off = p16_in_tile(rr, k & 63);
            *reinterpret_cast<uint4*>(tl + off) = hi;
            *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = lo;
          }
        }
      }
    }
    if constexpr (SW) {
      if constexpr (PRIV) {                                         // lanes l and l + 16 own the same unit
#pragma unroll
        for (int k = 0; k < 4; ++k) bsum[k] += __shfl_xor_sync(0xffffffffu, bsum[k], 16);
        if (half == 0) {
          atomicAdd(d.db_ih + u, bsum[0]); atomicAdd(d.db_ih + H + u, bsum[1]); atomicAdd(d.db_ih + 2 * H + u, bsum[2]);
          atomicAdd(d.db_hh + u, bsum[0]); atomicAdd(d.db_hh + H + u, bsum[1]); atomicAdd(d.db_hh + 2 * H + u, bsum[3]);
        }
      }
      if constexpr (SUM) {                                          // (NG = 1 only: b0 / gidx are those of group 0)
        st8(d.dgi_sum + (long)u * bpad + b0, sum_r);
        st8(d.dgi_sum + (long)(H + u) * bpad + b0, sum_z);
        st8(d.dgi_sum + (long)(2 * H + u) * bpad + b0, sum_n);
        constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
        const long row = gidx * 16 + nrow;
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_sum_p) + (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
        const int rr = (int)(row & 127);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint4 hi, lo;
          rw_transpose_pack(g == 0 ? sum_r : (g == 1 ? sum_z : sum_n), lane, hi, lo);
          const int k = g * H + (int)c * UC + kq;
          __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
          const int off = p16_in_tile(rr, k & 63);
          *reinterpret_cast<uint4*>(tl + off) = hi;
          *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = lo;
        }
      }
    }
  } else if (epi) {
    if constexpr (WS) reg_inc<SW ? 176 : (SUM ? 224 : 216)>();
    if (d.dh_last) ld8(d.dh_last + (long)u * d.dh_last_ld + b0, carry);
    load_step(0);
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? s : steps - 1 - s;
      const uint32_t ph = s & 1;
      uint8_t* rprev = sRg + (ph ^ 1u) * RSLOT;                     // partial sums of step s - 1; then this step's operand
      RW_STAMP(0);
      float dar[8], daz[8], dan[8], dgn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] += carry[i] + own[i];
      if (s > 0) {
        mbar_wait_cluster_b(&pfullg[ph ^ 1u], ((s - 1) >> 1) & 1, 22);   // the peers' partial sums of step s - 1 have landed
        if (q == 0 && lane == 0 && s + 1 < steps) mbar_expect_tx(&pfullg[ph ^ 1u], RSLOT);   // arm it for the pushes of step s + 1
        __syncwarp();
        RW_STAMP(1);
        if (tid == 0) RW_STAMP_CTA8(0);
        if (tid == 96) RW_STAMP_CTA8(1);
#pragma unroll
        for (int src = 0; src < 3; ++src) {
          float v[8];
          ld8s(reinterpret_cast<const float*>(rprev + part_off(src, j)), v);
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] += v[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dn = dh[i] * (1.0f - z[i]);
        const float dz = dh[i] * (hp[i] - n[i]);
        dan[i] = dn * (1.0f - n[i] * n[i]);
        daz[i] = dz * z[i] * (1.0f - z[i]);
        dar[i] = dan[i] * ghn[i] * r[i] * (1.0f - r[i]);
        dgn[i] = dan[i] * r[i];
        carry[i] = dh[i] * z[i];
        if constexpr (SUM && !SW) { sum_r[i] += dar[i]; sum_z[i] += daz[i]; sum_n[i] += dan[i]; }
        if constexpr (PRIV && !SW) { bsum[0] += dar[i]; bsum[1] += daz[i]; bsum[2] += dan[i]; bsum[3] += dgn[i]; }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // the group has consumed rprev before it becomes the operand
      uint4 rhi, rlo, zhi, zlo, ghi, glo;
      if constexpr (RW_MN) {
        split8(dar, rhi, rlo);
        split8(daz, zhi, zlo);
        split8(dgn, ghi, glo);
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(8 * half, j)) = rhi;
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(16 + 8 * half, j)) = rlo;
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(8 * half, UC + j)) = zhi;
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(16 + 8 * half, UC + j)) = zlo;
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(8 * half, 2 * UC + j)) = ghi;
        *reinterpret_cast<uint4*>(rprev + rw_mn_off(16 + 8 * half, 2 * UC + j)) = glo;
      } else {
        rw_transpose_pack(dar, lane, rhi, rlo);
        rw_transpose_pack(daz, lane, zhi, zlo);
        rw_transpose_pack(dgn, lane, ghi, glo);
        *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, kq)) = rhi;
        *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, kq)) = rlo;
        *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, UC + kq)) = zhi;
        *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, UC + kq)) = zlo;
        *reinterpret_cast<uint4*>(rprev + rw_b_off(nrow, 2 * UC + kq)) = ghi;
        *reinterpret_cast<uint4*>(rprev + rw_b_off(16 + nrow, 2 * UC + kq)) = glo;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&ofull[grp]);
      RW_STAMP(2);
      if (tid == 0) RW_STAMP_CTA8(2);
      if (tid == 96) RW_STAMP_CTA8(3);
      // ---- while the MMAs run: next step's inputs ----
      if (s + 1 < steps && !(a.exp & 64)) load_step(s + 1);          // (experiment 64: no global loads)
      // ---- partial sums of dh_{t-1}: accumulator i holds the input units owned by CTA (c + 1 + i) & 3 ----
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) {
        mbar_wait_warp(&doneg[i], ph, 24);
        tc_fence_after();
        RW_STAMP(6 + i);
        float part[8];
        rw_reduce32(tacc + i * 32, half, part);
        if (i == 0 && s > 0) {
          // the receive slot written below is the peers' operand of step s - 1: their MMAs of that step must be complete
          // (signalled long ago - this wait only makes the ordering a guarantee instead of a timing margin)
          mbar_wait_cluster_b(&mfreeg[ph ^ 1u], ((s - 1) >> 1) & 1, 23);
          __syncwarp();
        }
        if (i < 3) {
          const uint32_t owner = (c + 1 + i) & 3;
          const uint32_t slot = (c - owner - 1) & 3;             // = 2 - i: position of this CTA among the owner's three peers
          const uint32_t off = smem_u32(sRg) + ph * RSLOT + part_off((int)slot, j);
          const uint32_t ra = mapa_u32(off, owner), pbar = mapa_u32(smem_u32(&pfullg[ph]), owner);
          st_async_f4(ra, part[0], part[1], part[2], part[3], pbar);
          st_async_f4(ra + 16, part[4], part[5], part[6], part[7], pbar);
          if (i == 2 && tid == 0) RW_STAMP_CTA8(4);
          if (i == 2 && tid == 96) RW_STAMP_CTA8(5);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) own[k] = part[k];
          // commits complete in order: every MMA of this step has finished reading the operand
          if (lane == 0 && s + 1 < steps) {
#pragma unroll
            for (uint32_t r = 1; r < 4; ++r) {
              if (a.exp & 128) mbar_arrive_remote(mapa_u32(smem_u32(&mfreeg[ph]), (c + r) & 3));
              else mbar_arrive_remote_relaxed(mapa_u32(smem_u32(&mfreeg[ph]), (c + r) & 3));
            }
          }
        }
      }
      tc_fence_before();
      RW_STAMP(4);
      if (tid == 0) RW_STAMP_CTA8(6);
      if (tid == 96) RW_STAMP_CTA8(7);
      // ---- this step's outputs (weight-gradient GEMMs, dx of the layer below): after the pushes, because a release at
      //      cluster scope waits for every earlier store of the thread ----
      if constexpr (SW) {                                           // hand the gate gradients to the store warps
        const uint32_t sb = s & 1, stg = taddr + SBASE + 32u * (uint32_t)(2 * grp + (int)sb);
        if (s >= 2) mbar_wait_warp(&sempty[2 * grp + sb], ((s - 2) >> 1) & 1, 27);
        __syncwarp();                                               // tcgen05.st is .sync.aligned: the warp must be converged
        tc_fence_after();
        tmem_st8(stg, dar);
        tmem_st8(stg + 8, daz);
        tmem_st8(stg + 16, dan);
        tmem_st8(stg + 24, dgn);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&sfull[2 * grp + sb]);
        continue;
      }
      if (!(a.exp & 16)) {                                         // (experiment 16: no global stores)
        const long o = (long)t * bpad + b0;
        if constexpr (PRIV) {
          st8_T_p16(d.dghT_p, d.gT_nk, u, o, dar);
          st8_T_p16(d.dghT_p, d.gT_nk, H + u, o, daz);
          st8_T_p16(d.dghT_p, d.gT_nk, 2 * H + u, o, dgn);
          if (d.dgiT_p) {
            st8_T_p16(d.dgiT_p, d.gT_nk, u, o, dar);
            st8_T_p16(d.dgiT_p, d.gT_nk, H + u, o, daz);
            st8_T_p16(d.dgiT_p, d.gT_nk, 2 * H + u, o, dan);
          }
        } else {
          st8(d.dgi + (long)u * d.dg_ld + o, dar);
          st8(d.dgi + (long)(H + u) * d.dg_ld + o, daz);
          st8(d.dgi + (long)(2 * H + u) * d.dg_ld + o, dan);
          st8(d.dgh + (long)u * d.dg_ld + o, dar);
          st8(d.dgh + (long)(H + u) * d.dg_ld + o, daz);
          st8(d.dgh + (long)(2 * H + u) * d.dg_ld + o, dgn);
        }
        if (d.dgi_p) {
          uint4 nhi, nlo;
          rw_transpose_pack(dan, lane, nhi, nlo);
          if constexpr (RW_MN) {                                    // the K-major P16 copy for the dx GEMM (off the recurrence)
            rw_transpose_pack(dar, lane, rhi, rlo);
            rw_transpose_pack(daz, lane, zhi, zlo);
          }
          constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
          const long row = gidx * 16 + nrow;
          __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_p) + (size_t)t * d.dgi_p_slot_elems +
                                (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
          const int rr = (int)(row & 127);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int k = g * H + (int)c * UC + kq;
            __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
            const int off = p16_in_tile(rr, k & 63);
            *reinterpret_cast<uint4*>(tl + off) = g == 0 ? rhi : (g == 1 ? zhi : nhi);
            *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = g == 0 ? rlo : (g == 1 ? zlo : nlo);
          }
        }
      }
      RW_STAMP(5);
    }
  // gradient of the initial state: carry + own tile + the peers' partial sums of the last step
      mbar_wait_cluster_b(&pfullg[(steps - 1) & 1], ((steps - 1) >> 1) & 1, 25);
      __syncwarp();
      const uint8_t* rl = sRg + ((steps - 1) & 1) * RSLOT;
      float g0[8];
  #pragma unroll
      for (int i = 0; i < 8; ++i) g0[i] = carry[i] + own[i];
  #pragma unroll
      for (int src = 0; src < 3; ++src) {
        float v[8];
        ld8s(reinterpret_cast<const float*>(rl + part_off(src, j)), v);
  #pragma unroll
        for (int i = 0; i < 8; ++i) g0[i] += v[i];
      }
      st8(d.dh0_out + (long)u * bpad + b0, g0);
    if constexpr (!SW) {
      if constexpr (PRIV) {                                         // lanes l and l + 16 own the same unit
  #pragma unroll
        for (int k = 0; k < 4; ++k) bsum[k] += __shfl_xor_sync(0xffffffffu, bsum[k], 16);
        if (half == 0) {
          atomicAdd(d.db_ih + u, bsum[0]); atomicAdd(d.db_ih + H + u, bsum[1]); atomicAdd(d.db_ih + 2 * H + u, bsum[2]);
          atomicAdd(d.db_hh + u, bsum[0]); atomicAdd(d.db_hh + H + u, bsum[1]); atomicAdd(d.db_hh + 2 * H + u, bsum[3]);
        }
      }
      if constexpr (SUM) {
        st8(d.dgi_sum + (long)u * bpad + b0, sum_r);
        st8(d.dgi_sum + (long)(H + u) * bpad + b0, sum_z);
        st8(d.dgi_sum + (long)(2 * H + u) * bpad + b0, sum_n);
        constexpr int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
        const long row = gidx * 16 + nrow;
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_sum_p) + (size_t)(row >> 7) * nkc3 * p16_tile_elems(128);
        const int rr = (int)(row & 127);
  #pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint4 hi, lo;
          rw_transpose_pack(g == 0 ? sum_r : (g == 1 ? sum_z : sum_n), lane, hi, lo);
          const int k = g * H + (int)c * UC + kq;
          __nv_bfloat16* tl = base + (size_t)(k >> 6) * p16_tile_elems(128);
          const int off = p16_in_tile(rr, k & 63);
          *reinterpret_cast<uint4*>(tl + off) = hi;
          *reinterpret_cast<uint4*>(tl + 128 * KCHUNK + off) = lo;
        }
      }
    }
  }
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_before();
  __syncthreads();
  if (warp == NGW) tmem_dealloc(tmem, TCOLS);
}

// =================================================================================================
// launchers
// =================================================================================================
// private interchange layouts: both sweeps of the layer must run on the H = 256 barrier-free kernels
bool rw_priv_mode(int H, int tiles) { return g_opt_rw_priv && g_opt_rw == 3 && g_opt_rw2 && H == 256 && rw_applicable(H, tiles); }
// 16-row groups per cluster of the H = 256 kernels: one while the sweep fits one wave of the SMs a cluster-of-4 launch can use
// (132 of 148), otherwise two (32 rows per cluster, N = 64 MMAs at the cost of N = 32 ones); option rw_ng forces 1 or 2
int rw_groups_per_cluster(int H, int tiles) {
  if (H != 256 || !g_opt_rw2) return 1;
  if (g_opt_rw_ng == 1 || g_opt_rw_ng == 2) return g_opt_rw_ng;
  return tiles * 8 * 2 * 4 <= 132 ? 1 : 2;
}
bool rw_applicable(int H, int tiles) {
  // 2 directions x (B_pad / 16 / groups per cluster) clusters x 4 CTAs must fit the allowed number of waves
  return H >= 64 && H <= 256 && H % 64 == 0 && tiles >= 1 && tiles * 8 / rw_groups_per_cluster(H, tiles) * 2 * 4 <= 132 * g_opt_rw_waves;
}
size_t rw_whh_bytes(int H) { return (size_t)4 * 3 * (H / 64) * RW_ATILE; }
size_t rw_whhT_bytes(int H) {
  const int UC = H / 4, nkb = (3 * UC + KCHUNK - 1) / KCHUNK;
  return (size_t)4 * (H / 64) * nkb * RW_ATILE;
}

template <typename K, typename A>
static void rw_launch(K kernel, const A& a, size_t smem, int groups, int ndir, cudaStream_t st, int threads = RW_THREADS) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(4, groups, ndir);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, a);
}

template <int NG>
static void rw2_launch_fwd(const GruSeqFwdArgs& a, cudaStream_t st) {
  const int groups = a.tiles * 8 / NG;
  const size_t smem = (size_t)12 * RW_ATILE + (size_t)(NG == 1 ? 2 : 1) * 4 * RW_BTILE * NG + 512;
  if (a.d[0].priv && (g_opt_rw_sw & 1)) rw_launch(gru_rw2_fwd_kernel<true, true, NG>, a, smem, groups, a.ndir, st, rw2_fwd_threads(true, NG));
  else if (a.d[0].priv) rw_launch(gru_rw2_fwd_kernel<true, false, NG>, a, smem, groups, a.ndir, st, rw2_fwd_threads(false, NG));
  else rw_launch(gru_rw2_fwd_kernel<false, false, NG>, a, smem, groups, a.ndir, st, rw2_fwd_threads(false, NG));
}

void launch_gru_rw_fwd(const GruSeqFwdArgs& a_in, cudaStream_t st) {
  GruSeqFwdArgs a = a_in;
  a.dbg = g_dbg_buffer;
  a.exp = g_opt_rw_exp;
  const int nkc = a.H / 64, groups = a.tiles * 8;
  const size_t smem = (size_t)3 * nkc * RW_ATILE + (size_t)2 * nkc * RW_BTILE + 256;
  count_launch();
  if (nkc == 4 && g_opt_rw2) {
    if (rw_groups_per_cluster(a.H, a.tiles) == 2) rw2_launch_fwd<2>(a, st);
    else rw2_launch_fwd<1>(a, st);
    return;
  }
  switch (nkc) {
    case 1: rw_launch(gru_rw_fwd_kernel<1>, a, smem, groups, a.ndir, st); break;
    case 2: rw_launch(gru_rw_fwd_kernel<2>, a, smem, groups, a.ndir, st); break;
    case 3: rw_launch(gru_rw_fwd_kernel<3>, a, smem, groups, a.ndir, st); break;
    default: rw_launch(gru_rw_fwd_kernel<4>, a, smem, groups, a.ndir, st); break;
  }
}

template <bool SUM, int NG>
static void rw2_launch_bwd(const GruSeqBwdArgs& a, cudaStream_t st) {
  const int groups = a.tiles * 8 / NG;
  const size_t sm2 = (size_t)(NG == 2 ? 9 : 12) * RW_ATILE + (size_t)2 * 12288 * NG + 512;
  if constexpr (NG == 1 || !SUM) {
    if (a.d[0].priv && (g_opt_rw_sw & (NG == 1 ? 2 : 4))) {         // rw_sw bit 1: store warps at 16 rows per cluster, bit 2: at 2 x 16
      rw_launch(gru_rw2_bwd_kernel<SUM, true, true, NG>, a, sm2, groups, a.ndir, st, rw2_bwd_threads(true, NG));
      return;
    }
  }
  if (a.d[0].priv) rw_launch(gru_rw2_bwd_kernel<SUM, true, false, NG>, a, sm2, groups, a.ndir, st, rw2_bwd_threads(false, NG));
  else rw_launch(gru_rw2_bwd_kernel<SUM, false, false, NG>, a, sm2, groups, a.ndir, st, rw2_bwd_threads(false, NG));
}

template <bool SUM>
static void rw_launch_bwd(const GruSeqBwdArgs& a, cudaStream_t st) {
  const int nkc = a.H / 64, groups = a.tiles * 8, UC = a.H / 4;
  const int nkb = (3 * UC + KCHUNK - 1) / KCHUNK;
  const size_t rslot = (size_t)(4 * UC * 16 * 4 > nkb * RW_BTILE ? 4 * UC * 16 * 4 : nkb * RW_BTILE);
  const size_t smem = (size_t)nkc * nkb * RW_ATILE + 2 * rslot + 256;
  if (nkc == 4 && g_opt_rw2) {
    if (rw_groups_per_cluster(a.H, a.tiles) == 2) rw2_launch_bwd<SUM, 2>(a, st);
    else rw2_launch_bwd<SUM, 1>(a, st);
    return;
  }
  switch (nkc) {
    case 1: rw_launch(gru_rw_bwd_kernel<1, SUM>, a, smem, groups, a.ndir, st); break;
    case 2: rw_launch(gru_rw_bwd_kernel<2, SUM>, a, smem, groups, a.ndir, st); break;
    case 3: rw_launch(gru_rw_bwd_kernel<3, SUM>, a, smem, groups, a.ndir, st); break;
    default: rw_launch(gru_rw_bwd_kernel<4, SUM>, a, smem, groups, a.ndir, st); break;
  }
}
void launch_gru_rw_bwd(const GruSeqBwdArgs& a_in, cudaStream_t st) {
  GruSeqBwdArgs a = a_in;
  a.dbg = g_dbg_buffer;
  a.exp = g_opt_rw_exp;
  count_launch();
  if (a.d[0].dgi_sum) rw_launch_bwd<true>(a, st);
  else rw_launch_bwd<false>(a, st);
}

unsigned int rw_timeouts() {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_rw_timeouts, sizeof(v));
  return v;
}
int rw_timeout_info(int i) {
  int v[8];
  cudaMemcpyFromSymbol(v, g_rw_timeout_site, sizeof(v));
  return v[i & 7];
}
void rw_timeouts_reset() {
  unsigned int z = 0;
  cudaMemcpyToSymbol(g_rw_timeouts, &z, sizeof(z));
}

}  // namespace vb
