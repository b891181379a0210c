// Row-resident forward sweep for LARGE inference batches (sliding-window embedding, vame/analysis/pose_segmentation.py:87-98:
// 1e6 independent windows, no backward pass): one persistent CTA owns 128 batch rows of one direction for the whole sweep.
//
// The resident-weight kernels of gru_rw.cu are built for the latency-bound training step (16-32 rows per cluster, weights in
// shared memory, h exchanged through DSMEM); with thousands of rows per step the roles flip:
//   * the 128 rows of h_{t-1} are the M side of tcgen05.mma and stay on the SM for all T steps: the bf16 hi / lo operand tiles in
//     shared memory (128 KB, exactly the P16 tile format - so the per-step / final h that the next GEMM consumes is ONE bulk copy
//     of this buffer), the exact fp32 state in tensor memory (256 columns, lane = row);
//   * W_hh (hi + lo, 786 KB per direction) is L2-resident and streamed by a TMA producer warp, one 24 KB (32-unit slice, 64-k
//     chunk) tile per pipeline stage; N = 96 = the r, z, n gate rows of the slice;
//   * the three split products a_hi w_hi + a_lo w_hi + a_hi w_lo accumulate into the SAME accumulator cell (the split is on the
//     K side here), so the epilogue reads one fp32 per output - a quarter of the tensor-memory traffic of the swap-AB kernels;
//   * two 96-column accumulators: the gate math of slice c (8 warps, thread = row, 16 units each) overlaps the MMAs of slice
//     c + 1.  New h values are parked in the fp32 tensor-memory state (the MMAs read shared memory, so there is no hazard) and
//     the operand tiles are rebuilt once per step, after the last slice.
// Per step and CTA: 151 MFLOP issued (9.4 us at the tensor peak), 786 KB of weights from L2, 393 KB of input projections.
// Cell equations: torch.nn.GRU as used at vame/model/rnn_model.py:41; numerics as everywhere in this library.
#include "common.cuh"
#include "kernels.h"

namespace vb {

__device__ unsigned int g_rows_timeouts = 0;   // bounded mbarrier waits that gave up (must stay 0; vame_get_option("rows_timeouts"))

namespace {
// a wait that cannot hang the GPU: a protocol bug shows up as a counted time-out (and wrong numbers), not as a dead box
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 20); ++i) {
    if (mbar_try_wait(bar, parity)) return;
    if ((i & 4095) == 4095 && *reinterpret_cast<volatile unsigned int*>(&g_rows_timeouts) != 0) return;   // someone already gave up
  }
  atomicAdd(&g_rows_timeouts, 1u);
}
constexpr int GR_H = 256, GR_NKC = 4, GR_NSL = 8;              // hidden units, 64-k chunks, 32-unit slices
constexpr int GR_WTILE = 96 * KCHUNK * 2 * 2;                   // 24576 B: one (slice, k chunk) W tile, hi + lo planes
constexpr int GR_HTILE = 128 * KCHUNK * 2 * 2;                  // 32768 B: one k chunk of the h operand, hi + lo planes
constexpr int GR_NST = 4;                                       // W pipeline stages (96 KB in flight per SM)
// EW gate-math warps per tensor-memory lane quarter (2 or 4): warps 0 .. 4 EW - 1 gate math (32 / EW units of a slice per thread),
// then one warpgroup: TMA producer (+ TMEM alloc), MMA issue, two idle warps (whole warpgroups, so that setmaxnreg can move
// registers to the gate-math warps)
__host__ __device__ constexpr int gr_threads(int ew) { return (4 * ew + 4) * 32; }
constexpr uint32_t GR_HCOL = 192;                               // tensor memory: accumulators [0, 96), [96, 192); fp32 h [192, 448)
constexpr size_t GR_SMEM = (size_t)GR_NKC * GR_HTILE + (size_t)GR_NST * GR_WTILE + 1024 + 256;

#ifdef VAME_ACCURATE_MATH
__device__ __forceinline__ float gr_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gr_tanh(float x) { return tanhf(x); }
#else
__device__ __forceinline__ float gr_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gr_tanh(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
#endif

__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, float* v) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// shared -> global bulk copy (TMA engine), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gr_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void gr_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int NT>
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }   // the gate-math warps
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
template <int N>
__device__ __forceinline__ void tm_ld(uint32_t taddr, float* v) {
  if constexpr (N == 16) tmem_ld16u(taddr, v);
  else tmem_ld8u(taddr, v);
}
template <int N>
__device__ __forceinline__ void tm_st(uint32_t taddr, const float* v) {
  if constexpr (N == 16) tmem_st16(taddr, v);
  else tmem_st8u(taddr, v);
}
}  // namespace

// grid = (tiles, directions), 1 CTA per SM
template <int EW>
__global__ void __launch_bounds__(gr_threads(EW), 1) gru_rows_fwd_kernel(const GruSeqFwdArgs a) {
  constexpr int NEW = 4 * EW, NET = 128 * EW, UPT = 32 / EW;      // gate-math warps / threads, units of a slice per thread
  constexpr bool PF = EW == 2;                                   // input projections fetched one slice ahead (second register set)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sH = smem;                                            // [NKC][hi 128 x 64 | lo 128 x 64]  (P16 tiles, RB = 128)
  uint8_t* sW = smem + GR_NKC * GR_HTILE;                         // [NST][hi 96 x 64 | lo 96 x 64]     (P16 tiles, RB = 96)
  float* sBhn = reinterpret_cast<float*>(sW + GR_NST * GR_WTILE); // [256]
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sBhn + GR_H);     // [NST]
  uint64_t* wempty = wfull + GR_NST;                              // [NST]
  uint64_t* accfull = wempty + GR_NST;                            // [2]
  uint64_t* accempty = accfull + 2;                               // [2]
  uint64_t* hready = accempty + 2;                                // operand tiles of this step built (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready + 1);

  const GruSeqDirFwd& d = a.d[blockIdx.y];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int tile = blockIdx.x, steps = a.steps;
  const long Bp = (long)a.tiles * 128;

  if (tid == 0) {
    for (int i = 0; i < GR_NST; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], NET); }
    mbar_init(hready, NET);
    mbar_fence_init();
  }
  for (int i = tid; i < GR_H; i += gr_threads(EW)) sBhn[i] = d.b_hn[i];
  if (warp == NEW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // EW = 2: 384 threads start with 168 registers each; the producer / MMA warpgroup gives 4 x 96 back, the two gate-math
  // warpgroups take 8 x 48.  EW = 4: 640 threads start with 96; 4 x 56 given back, 16 x 8 taken.  (No spills around the
  // asynchronous tcgen05.ld - see gru_rw.cu / profiles/r2_ng2_sw_fault.md; each role's code sits inside the branch of its
  // setmaxnreg: ptxas applies a budget only to code the instruction dominates.)
  if (warp >= NEW) {
  gr_reg_dec<EW == 2 ? 72 : 40>();
  if (warp == NEW) {
    // ---------------- TMA producer: the W tiles of (slice, k chunk), the same sequence every step ----------------
    if (elect_one()) {
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.w_rows);
      uint32_t it = 0;
      for (int s = 0; s < steps; ++s) {
        for (int i = 0; i < GR_NSL * GR_NKC; ++i, ++it) {
          const uint32_t st = it % GR_NST, use = it / GR_NST;
          if (use > 0) mbar_wait_bounded(&wempty[st], (use - 1) & 1);
          mbar_expect_tx(&wfull[st], GR_WTILE);
          bulk_g2s(sW + (size_t)st * GR_WTILE, wp + (size_t)i * GR_WTILE, GR_WTILE, &wfull[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp == NEW + 1) {
    // ---------------- MMA issue ----------------
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 96);
      uint32_t it = 0, acc_use = 0;
      for (int s = 0; s < steps; ++s) {
        mbar_wait_bounded(hready, s & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        for (int c = 0; c < GR_NSL; ++c, ++acc_use) {
          const uint32_t ab = acc_use & 1;
          if (acc_use >= 2) mbar_wait_bounded(&accempty[ab], ((acc_use >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t dcol = tmem + ab * 96;
#pragma unroll 1
          for (int kc = 0; kc < GR_NKC; ++kc, ++it) {
            const uint32_t st = it % GR_NST;
            mbar_wait_bounded(&wfull[st], (it / GR_NST) & 1);
            tc_fence_after();
            const uint64_t dAhi = make_desc(smem_u32(sH) + kc * GR_HTILE), dAlo = make_desc(smem_u32(sH) + kc * GR_HTILE + GR_HTILE / 2);
            const uint64_t dBhi = make_desc(smem_u32(sW) + st * GR_WTILE), dBlo = make_desc(smem_u32(sW) + st * GR_WTILE + GR_WTILE / 2);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = ks * 2 * ATOM_BYTES;
              if (kc == 0 && ks == 0) umma_bf16_c<0>(dcol, desc_advance(dAhi, o), desc_advance(dBhi, o), idesc);
              else umma_bf16_c<1>(dcol, desc_advance(dAhi, o), desc_advance(dBhi, o), idesc);
              umma_bf16_c<1>(dcol, desc_advance(dAlo, o), desc_advance(dBhi, o), idesc);
              umma_bf16_c<1>(dcol, desc_advance(dAhi, o), desc_advance(dBlo, o), idesc);
            }
            umma_commit(&wempty[st]);                             // the stage is free once these MMAs have read it
          }
          umma_commit(&accfull[ab]);
        }
      }
    }
    __syncwarp();
  }
  } else {
    gr_reg_inc<EW == 2 ? 216 : 104>();
    // ---------------- gate math: warp w owns tensor-memory lane quarter w & 3 (32 rows) and units UPT (w >> 2) .. + UPT - 1 of a slice ----
    const int q = warp & 3, hh = warp >> 2;
    const int r = 32 * q + lane;                                  // row inside the tile = tensor-memory lane
    const long row = (long)tile * 128 + r;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float* gi_row = d.gi + row * d.gi_bs;

    // rebuild of the operand tiles from the fp32 state: this warp converts units [256 / EW * hh, + 256 / EW) of its 32 rows
    auto rebuild = [&]() {
#pragma unroll 1
      for (int j = 0; j < 16 / EW; ++j) {                         // 16 units per iteration
        float v[16];
        const int k0 = (256 / EW) * hh + 16 * j;
        tmem_ld16u(tlane + GR_HCOL + k0, v);
        tmem_ld_wait();
        uint4 hi0, lo0, hi1, lo1;
        split8(v, hi0, lo0);
        split8(v + 8, hi1, lo1);
        uint8_t* t = sH + (size_t)(k0 >> 6) * GR_HTILE;
        const uint32_t o0 = 2u * (uint32_t)p16_in_tile(r, k0 & 63), o1 = 2u * (uint32_t)p16_in_tile(r, (k0 & 63) + 8);
        *reinterpret_cast<uint4*>(t + o0) = hi0;
        *reinterpret_cast<uint4*>(t + o1) = hi1;
        *reinterpret_cast<uint4*>(t + GR_HTILE / 2 + o0) = lo0;
        *reinterpret_cast<uint4*>(t + GR_HTILE / 2 + o1) = lo1;
      }
    };
    // initial state -> fp32 tensor-memory state (every warp: its share of each slice), then the operand tiles
#pragma unroll 1
    for (int c = 0; c < GR_NSL; ++c) {
      float v[UPT];
#pragma unroll
      for (int i = 0; i < UPT; ++i) v[i] = d.h0[(long)(32 * c + UPT * hh + i) * d.h0_ld + row];
      tm_st<UPT>(tlane + GR_HCOL + 32 * c + UPT * hh, v);
    }
    tmem_st_wait_();
    tc_fence_before();
    epi_sync<NET>();
    tc_fence_after();
    rebuild();
    fence_proxy_async_smem();
    mbar_arrive(hready);

    uint32_t acc_use = 0;
    // input projections are fetched ONE SLICE AHEAD into a second register set: they stream from HBM (393 KB per CTA and step),
    // and a load issued at the top of its own slice left ~1.5 us of latency exposed 8 times per step (measured: 25.7 us per step)
    float gA[3 * UPT], gB[PF ? 3 * UPT : 1];
    auto load_gi = [&](int s_, int c_, float* g_) {
      const int t_ = d.reverse ? steps - 1 - s_ : s_;
      const float* p_ = gi_row + (long)t_ * d.gi_ts + (long)(32 * c_ + UPT * hh) * d.gi_ld;
#pragma unroll
      for (int i = 0; i < UPT; ++i) {
        g_[i] = p_[(long)i * d.gi_ld];
        g_[UPT + i] = p_[(long)(GR_H + i) * d.gi_ld];
        g_[2 * UPT + i] = p_[(long)(2 * GR_H + i) * d.gi_ld];
      }
    };
    auto slice = [&](int c, const float* g_) {
      const uint32_t ab = acc_use & 1;
      const int u0 = 32 * c + UPT * hh;                           // first of this thread's units
      mbar_wait_bounded(&accfull[ab], (acc_use >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      float ar[UPT], az[UPT], an[UPT], hp[UPT];
      const uint32_t acol = tlane + ab * 96 + UPT * hh;
      tm_ld<UPT>(acol, ar);
      tm_ld<UPT>(acol + 32, az);
      tm_ld<UPT>(acol + 64, an);
      tm_ld<UPT>(tlane + GR_HCOL + u0, hp);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&accempty[ab]);                                 // the accumulator is in registers: the MMAs of slice c + 2 may start
      float hn[UPT];
#ifdef VAME_ACCURATE_MATH
#pragma unroll
      for (int i = 0; i < UPT; ++i) {
        const float rg = gr_sigmoid(g_[i] + ar[i]);
        const float zg = gr_sigmoid(g_[UPT + i] + az[i]);
        const float ng = gr_tanh(g_[2 * UPT + i] + rg * (an[i] + sBhn[u0 + i]));
        hn[i] = (1.0f - zg) * ng + zg * hp[i];
      }
#else
      // The kernel is bound by the MUFU pipe (ncu: XU 72 % of peak, profiles/r2_ncu_full_v13_rows.md): 3 ex2 + 3 rcp per unit.
      // Reciprocals are shared instead: 1 / d_r and 1 / d_z of one unit come from ONE rcp of d_r d_z, the tanh reciprocals of two
      // units from one rcp of their product - 3 ex2 + 1 rcp per unit.  Arguments are clamped to [-20, 20] (sigmoid / tanh are
      // saturated to 2e-9 there) so that the products stay far below the fp32 range.
#pragma unroll
      for (int i = 0; i < UPT; i += 2) {
        float rg[2], zg[2], dn[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float xr = fminf(fmaxf(g_[i + k] + ar[i + k], -20.f), 20.f);
          const float xz = fminf(fmaxf(g_[UPT + i + k] + az[i + k], -20.f), 20.f);
          const float dr = 1.0f + __expf(-xr), dz = 1.0f + __expf(-xz);
          const float inv = __fdividef(1.0f, dr * dz);
          rg[k] = inv * dz;
          zg[k] = inv * dr;
          const float y = fminf(fmaxf(g_[2 * UPT + i + k] + rg[k] * (an[i + k] + sBhn[u0 + i + k]), -10.f), 10.f);
          dn[k] = 1.0f + __expf(-2.0f * y);
        }
        const float invn = __fdividef(1.0f, dn[0] * dn[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float ng = 2.0f * (invn * dn[k ^ 1]) - 1.0f;       // tanh(y) = 2 / (1 + exp(-2 y)) - 1
          hn[i + k] = (1.0f - zg[k]) * ng + zg[k] * hp[i + k];
        }
      }
#endif
      tm_st<UPT>(tlane + GR_HCOL + u0, hn);
      ++acc_use;
    };
    if constexpr (PF) load_gi(0, 0, gA);
    for (int s = 0; s < steps; ++s) {
      const int t = d.reverse ? steps - 1 - s : s;
      if constexpr (PF) {
#pragma unroll 1
        for (int c = 0; c < GR_NSL; c += 2) {
          load_gi(s, c + 1, gB);
          slice(c, gA);
          if (c + 2 < GR_NSL) load_gi(s, c + 2, gA);
          else if (s + 1 < steps) load_gi(s + 1, 0, gA);
          slice(c + 1, gB);
        }
      } else {                                                    // 16 gate-math warps hide the load latency among themselves
#pragma unroll 1
        for (int c = 0; c < GR_NSL; ++c) {
          load_gi(s, c, gA);
          slice(c, gA);
        }
      }
      // ---- end of the step: every MMA has completed (accfull of the last slice), the new state is complete in tensor memory ----
      tmem_st_wait_();
      tc_fence_before();
      if (tid == 0) bulk_wait_read0();                            // the previous step's bulk store has finished reading sH
      epi_sync<NET>();
      tc_fence_after();
      rebuild();
      fence_proxy_async_smem();
      const bool last = s + 1 == steps;
      if (!last) mbar_arrive(hready);
      const bool store = d.out_p && (d.out_p_slots == steps || last);
      if (store) {
        epi_sync<NET>();                                          // all gate-math warps have written their part of the tiles
        if (tid == 0) {
          const int sp = (d.out_p_slots == steps) ? t : (s & 1);
          uint8_t* dst = reinterpret_cast<uint8_t*>(d.out_p) + ((size_t)sp * d.out_p_slot_elems + (size_t)tile * GR_NKC * (GR_HTILE / 2)) * 2;
          fence_proxy_async_smem();
          bulk_s2g(dst, sH, GR_NKC * GR_HTILE);
          bulk_commit();
        }
      }
      if (last && d.out) {                                        // fp32 final state, feature-major (coalesced over the rows)
        const int so = (d.out_slots == steps) ? t : (s & 1);
#pragma unroll 1
        for (int c = 0; c < GR_NSL; ++c) {
          float v[UPT];
          tm_ld<UPT>(tlane + GR_HCOL + 32 * c + UPT * hh, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < UPT; ++i) d.out[(long)(32 * c + UPT * hh + i) * d.out_ld + (long)so * Bp + row] = v[i];
        }
      }
    }
    if (tid == 0) bulk_wait0();                                   // the last bulk store has left shared memory and is visible
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NEW) tmem_dealloc(tmem, 512);
}

bool rows_fwd_applicable(int H, int tiles) { return g_opt_rows && H == GR_H && tiles >= 1; }
unsigned int rows_timeouts() {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_rows_timeouts, sizeof(v));
  return v;
}

void launch_gru_rows_fwd(const GruSeqFwdArgs& a, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gru_rows_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GR_SMEM);
    cudaFuncSetAttribute(gru_rows_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GR_SMEM);
    attr = true;
  }
  count_launch();
  if (g_opt_rows >= 4) gru_rows_fwd_kernel<4><<<dim3(a.tiles, a.ndir), gr_threads(4), GR_SMEM, st>>>(a);
  else gru_rows_fwd_kernel<2><<<dim3(a.tiles, a.ndir), gr_threads(2), GR_SMEM, st>>>(a);
}

}  // namespace vb
