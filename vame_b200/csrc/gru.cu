// Per-timestep bi-GRU kernels (forward gate update and BPTT step) on tcgen05 tensor cores.
//
// Forward step (replaces one time step of nn.GRU as used at vame/model/rnn_model.py:41,106,141; cell equations from
// the torch.nn.GRU documentation, gate order r,z,n):
//     gh = h_{t-1} W_hh^T            -> tcgen05.mma, M = 128 batch rows, N = 96 (r,z,n of 32 hidden units), K = H
//     r = s(gi_r + gh_r) ; z = s(gi_z + gh_z) ; n = tanh(gi_n + r (gh_n + b_hn)) ; h = (1-z) n + z h_{t-1}
//   grid = (H/32 unit slices, batch tiles of 128, directions).  Each CTA keeps its 96 x H slice of W_hh (bf16 hi/lo)
//   and the 128 x H tile of h_{t-1} in shared memory (both fetched by the TMA engine as whole P16 tiles), issues
//   3 x H/16 MMAs into 96 TMEM columns and runs the gate math in the TMEM->register epilogue; the new h is written
//   as fp32 (sequence output / next step's elementwise operand) and as P16 hi/lo (next step's MMA operand).
//   Steps are chained with programmatic dependent launch: the W slice and the input-projection rows are fetched
//   before griddepcontrol.wait, so only the h tile load + MMA + epilogue are on the recurrence's critical path.
//
// Backward step (BPTT equations of SURVEY.md §3.5, verified against autograd in oracle/gru_numpy.py):
//     dh = sum(parts) + dout_t ; gate gradients ; dgh = [da_r, da_z, da_n r]
//     partial_c = dgh[:, gates of slice c] W_hh[gates of slice c, :]   -> tcgen05.mma, M = 128, N = H, K = 96
//   The K-split keeps the A operand local to the CTA (written to shared memory by its own threads); the H/32 partial
//   sums plus the carry dh*z are summed by the next step's prologue.
#include "common.cuh"
#include "kernels.h"

namespace vb {

#ifdef VAME_ACCURATE_MATH
__device__ __forceinline__ float gate_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) { return tanhf(x); }
#else
// default: MUFU-based gates (ex2.approx + rcp.approx, ~2 ulp each).  The absolute error (<3e-7) is far inside the
// stated parity tolerances and keeps the serial gate epilogue short; -DVAME_ACCURATE_MATH selects expf/tanhf.
__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
#endif

__device__ __forceinline__ void ld16(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void st16(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// feature-major access: 16 consecutive features of one row (lane = row -> each access is coalesced across the warp)
__device__ __forceinline__ void ldf16(const float* p, long ld, float* v) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = p[i * ld];
}
__device__ __forceinline__ void stf16(float* p, long ld, const float* v) {
#pragma unroll
  for (int i = 0; i < 16; ++i) p[i * ld] = v[i];
}
// write 16 consecutive k-values of one row into a P16 tile (two atoms), hi and lo planes; generic pointer (smem or global)
__device__ __forceinline__ void st16_p16(__nv_bfloat16* tile, int rows_in_tile, int r, int k, const float* v) {
  uint4 hi0, lo0, hi1, lo1;
  split8(v, hi0, lo0);
  split8(v + 8, hi1, lo1);
  const int off = p16_in_tile(r, k);
  __nv_bfloat16* lo = tile + (size_t)rows_in_tile * KCHUNK;
  *reinterpret_cast<uint4*>(tile + off) = hi0;
  *reinterpret_cast<uint4*>(tile + off + 64) = hi1;
  *reinterpret_cast<uint4*>(lo + off) = lo0;
  *reinterpret_cast<uint4*>(lo + off + 64) = lo1;
}

// ---- width-generic versions (N = values per thread: 16 with 8 warps per CTA, 8 with 16 warps per CTA) ----
template <int N>
__device__ __forceinline__ void ldN(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
template <int N>
__device__ __forceinline__ void ldfN(const float* p, long ld, float* v) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = p[i * ld];
}
template <int N>
__device__ __forceinline__ void stfN(float* p, long ld, const float* v) {
#pragma unroll
  for (int i = 0; i < N; ++i) p[i * ld] = v[i];
}
template <int N>
__device__ __forceinline__ void stN_p16(__nv_bfloat16* tile, int rows_in_tile, int r, int k, const float* v) {
  __nv_bfloat16* lo = tile + (size_t)rows_in_tile * KCHUNK;
  const int off = p16_in_tile(r, k);
  if constexpr (N == 4) {                 // half an atom: k % 4 == 0
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(v[i], h[i], l[i]);
    *reinterpret_cast<uint2*>(tile + off) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<uint2*>(l);
  } else {
#pragma unroll
    for (int a = 0; a < N / 8; ++a) {
      uint4 h, l;
      split8(v + 8 * a, h, l);
      *reinterpret_cast<uint4*>(tile + off + 64 * a) = h;
      *reinterpret_cast<uint4*>(lo + off + 64 * a) = l;
    }
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, float* v) {
  if constexpr (N == 16) tmem_ld16(taddr, v);
  else if constexpr (N == 8) tmem_ld8(taddr, v);
  else tmem_ld4(taddr, v);
}

// optional in-kernel timeline (vame_set_debug_buffer): CTA (0,0,0) thread 0 records %globaltimer at fixed points
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_STAMP(i)                                                                              \
  do {                                                                                            \
    if (a.dbg && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.dbg[(i)] = gtime(); \
  } while (0)
#define DBG_STAMP0(i)                                                                             \
  do {                                                                                            \
    if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.dbg[(i)] = gtime();     \
  } while (0)

// =================================================================================================
// forward step
// =================================================================================================
constexpr int F_WSUB = 48 * KCHUNK * 2 * 2;      // 12288 B : one K chunk of a 16-unit W slice (hi + lo planes of 48 rows)
constexpr int F_WTILE = 2 * F_WSUB;              // 24576 B : one K chunk of a 32-unit W slice
constexpr int F_ATILE = 128 * KCHUNK * 2 * 2;    // 32768 B : one K chunk of the h tile (hi + lo)

// UPC = hidden units per CTA (16 or 32), UPT = units per thread; 128 * UPC / UPT threads.
// MT = batch rows per CTA: 128, or 64 (grid.y = 2 x tiles; tcgen05.mma M = 64 keeps D in lanes 0-15 of each TMEM lane quarter,
// so lanes 16-31 of every warp idle through the element-wise part).  Halving the rows per CTA while doubling the CTAs
// shortens the step as long as the grid fits the GPU in one wave.
// TMEM columns of sub-slice s (16 units): [96 s, 96 s + 48) = a * w_hi for (r, z, n) x 16, [96 s + 48, 96 s + 96) = a * w_lo.
template <int UPC, int UPT, int MT>
__global__ void __launch_bounds__(128 * (UPC / UPT), 1) gru_step_fwd_kernel(const GruFwdArgs a) {
  constexpr int WT = (UPC / 16) * F_WSUB;
  constexpr int APL = MT * KCHUNK * 2;          // one plane (hi or lo) of a K chunk of the h tile
  constexpr int AT = 2 * APL;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, nkc = (H + KCHUNK - 1) / KCHUNK;
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)nkc * WT;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sA + (size_t)nkc * AT);
  uint64_t* abar = wbar + 4;
  uint64_t* done = abar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int c = blockIdx.x, tile = blockIdx.y / (128 / MT), hf = blockIdx.y % (128 / MT);
  const GruDirFwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;                 // half = which UPT-wide group of the slice's 32 units

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&wbar[i], 1);
      mbar_init(&abar[i], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- prologue: everything that does not depend on the previous time step ----
  // NOTE on warp roles: the TMA / MMA issuing thread is lane 0 of warp 0 and its 31 sibling lanes are parked at a
  // __syncwarp() — if they spun on an mbarrier instead, the divergent try_wait path would put the whole warp to sleep
  // and starve the issuing lane.
  DBG_STAMP(0);
  if (warp == 0) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.w_p);
      for (int kc = 0; kc < nkc; ++kc) {
        mbar_expect_tx(&wbar[kc], WT);
#pragma unroll
        for (int sub = 0; sub < UPC / 16; ++sub)
          bulk_g2s(sW + (size_t)kc * WT + sub * F_WSUB, wp + ((size_t)(c * (UPC / 16) + sub) * nkc + kc) * p16_tile_elems(48), F_WSUB,
                   &wbar[kc]);
      }
    }
    __syncwarp();
  }
  const bool act = (MT == 128) || lane < 16;                                      // this lane owns a row
  const int r_in = (MT == 128) ? q * 32 + lane : hf * 64 + q * 16 + (lane & 15);  // row inside the 128-row tile
  const long b = (long)tile * 128 + r_in;
  const int j0 = half * UPT;           // first unit inside the slice
  const int u0 = c * UPC + j0;         // first hidden unit handled by this thread
  float gir[UPT], giz[UPT], gin[UPT], bhn[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) gir[i] = giz[i] = gin[i] = bhn[i] = 0.f;
  if (act) {
    const float* gi_row = d.gi + (b * d.gi_bs + (long)d.t * d.gi_ts);
    ldfN<UPT>(gi_row + (long)u0 * d.gi_ld, d.gi_ld, gir);
    ldfN<UPT>(gi_row + (long)(H + u0) * d.gi_ld, d.gi_ld, giz);
    ldfN<UPT>(gi_row + (long)(2 * H + u0) * d.gi_ld, d.gi_ld, gin);
    ldN<UPT>(d.b_hn + u0, bhn);
  }

  DBG_STAMP(1);
  if (a.pdl && !a.flags) pdl_wait();   // previous step's h is now complete and visible
  if (a.pdl && !a.flags) pdl_launch_dependents();  // let the next step start its prologue
  DBG_STAMP(2);

  if (warp == 0) {
    if (elect_one()) {
      if (a.flags) {
        // flag hand-over: the previous step's CTAs (a still-running kernel) publish h and then bump the counter
        if (d.flag_in) {
          while (flag_acquire_load(d.flag_in + tile) < a.flag_expected) {
          }
        }
        fence_proxy_async_all();
        pdl_launch_dependents();       // at most one successor is resident (prologue + spin) while this step works
      }
      const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(d.h_in_p) + (size_t)tile * nkc * p16_tile_elems(128);
      for (int kc = 0; kc < nkc; ++kc) {
        mbar_expect_tx(&abar[kc], AT);
        if constexpr (MT == 128) {
          bulk_g2s(sA + (size_t)kc * AT, hp + (size_t)kc * p16_tile_elems(128), AT, &abar[kc]);
        } else {       // rows [64 hf, 64 hf + 64) of the hi plane and of the lo plane: two contiguous 8 KB pieces
          const __nv_bfloat16* t = hp + (size_t)kc * p16_tile_elems(128) + (size_t)hf * 64 * KCHUNK;
          bulk_g2s(sA + (size_t)kc * AT, t, APL, &abar[kc]);
          bulk_g2s(sA + (size_t)kc * AT + APL, t + 128 * KCHUNK, APL, &abar[kc]);
        }
      }
      const uint32_t idesc = make_idesc_bf16(MT, 6 * UPC);
      // the issue loop is the critical resource (one thread): descriptors are precomputed, the k-steps fully unrolled
      const uint64_t dA = make_desc(smem_u32(sA)), dAl = make_desc(smem_u32(sA) + APL), dW = make_desc(smem_u32(sW));
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        if (kc < nkc) {
          mbar_wait(&wbar[kc], 0);
          mbar_wait(&abar[kc], 0);
          if (kc == 0) DBG_STAMP0(3);
          tc_fence_after();
          const int ksteps = min(KCHUNK, H - kc * KCHUNK) / 16;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              // one descriptor covers the [W_hi ; W_lo] planes of all sub-slices (6 UPC rows)
              const uint32_t ao = kc * AT + ks * 2 * ATOM_BYTES, wo = kc * WT + ks * 2 * ATOM_BYTES;
              if (kc == 0 && ks == 0) umma_bf16_c<0>(tmem, desc_advance(dAl, ao), desc_advance(dW, wo), idesc);
              else umma_bf16_c<1>(tmem, desc_advance(dAl, ao), desc_advance(dW, wo), idesc);
              umma_bf16_c<1>(tmem, desc_advance(dA, ao), desc_advance(dW, wo), idesc);
            }
          }
        }
      }
      umma_commit(done);
      DBG_STAMP0(4);
    }
    __syncwarp();
  }

  float hprev[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) hprev[i] = 0.f;
  if (!a.flags && act) ldfN<UPT>(d.h_in + (long)u0 * d.h_in_ld + b, d.h_in_ld, hprev);     // complete + visible after griddepcontrol.wait
  mbar_wait(done, 0);                   // (with flags this also orders the h_in read below after the producer's release)
  __syncwarp();
  if (a.flags && act) {
#pragma unroll
    for (int i = 0; i < UPT; ++i) hprev[i] = ld_cg(d.h_in + (long)(u0 + i) * d.h_in_ld + b);
  }
  DBG_STAMP(5);
  tc_fence_after();
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  float ar[UPT], az[UPT], an[UPT];
  {
    float br[UPT], bz[UPT], bn[UPT];                 // columns 96.. hold the (* w_lo) products
    const uint32_t cb = taddr + (j0 / 16) * 96 + (j0 % 16);
    tmem_ldN<UPT>(cb, ar);
    tmem_ldN<UPT>(cb + 16, az);
    tmem_ldN<UPT>(cb + 32, an);
    tmem_ldN<UPT>(cb + 48, br);
    tmem_ldN<UPT>(cb + 64, bz);
    tmem_ldN<UPT>(cb + 80, bn);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < UPT; ++i) { ar[i] += br[i]; az[i] += bz[i]; an[i] += bn[i]; }
  }
  DBG_STAMP(8);

  float hn[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) {
    const float r = gate_sigmoid(gir[i] + ar[i]);
    const float z = gate_sigmoid(giz[i] + az[i]);
    const float ghn = an[i] + bhn[i];
    const float n = gate_tanh(gin[i] + r * ghn);
    hn[i] = (1.0f - z) * n + z * hprev[i];
    ar[i] = r; az[i] = z; an[i] = n; bhn[i] = ghn;
  }
  DBG_STAMP(9);
  if (act) stfN<UPT>(d.h_out + (long)u0 * d.h_out_ld + b, d.h_out_ld, hn);
  if (act) {
    const int kc = u0 / KCHUNK, kk = u0 % KCHUNK;
    __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(d.h_out_p) + ((size_t)tile * nkc + kc) * p16_tile_elems(128);
    stN_p16<UPT>(t, 128, r_in, kk, hn);
  }
  if (a.flags) {                        // publish: everything the next step reads is written above
    fence_proxy_async_all();
    __threadfence();
    __syncwarp();
    if (lane == 0) flag_release_add(d.flag_out + tile, 1u);
  }
  DBG_STAMP(10);
  if (d.sv_r && act) {
    const long so = (long)u0 * d.sv_ld + b;
    stfN<UPT>(d.sv_r + so, d.sv_ld, ar);
    stfN<UPT>(d.sv_z + so, d.sv_ld, az);
    stfN<UPT>(d.sv_n + so, d.sv_ld, an);
    stfN<UPT>(d.sv_ghn + so, d.sv_ld, bhn);
  }
  DBG_STAMP(6);
  // flag mode: kernels complete in launch order (each waits here for its predecessor), so the completion of the last step
  // kernel implies that every step's stores have landed — what a following normal launch / graph edge relies on
  if (a.pdl && a.flags) pdl_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
  DBG_STAMP(7);
}

static void launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, dim3 grid, int threads, size_t smem, cudaStream_t st,
                       int pdl) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.numAttrs = 0;
  if (pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
}

template <int UPC, int UPT, int MT>
static void launch_fwd_variant(const GruFwdArgs& a, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(gru_step_fwd_kernel<UPC, UPT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch_cfg(cfg, attr, dim3(a.H / UPC, a.tiles * (128 / MT), a.ndir), 128 * (UPC / UPT), smem, st, a.pdl);
  cudaLaunchKernelEx(&cfg, gru_step_fwd_kernel<UPC, UPT, MT>, a);
}

void launch_gru_step_fwd(const GruFwdArgs& a_in, cudaStream_t st) {
  GruFwdArgs a = a_in;
  a.dbg = g_dbg_buffer;
  const int nkc = (a.H + KCHUNK - 1) / KCHUNK;
  const int upc = a.upc == 16 ? 16 : 32;
  const int mt = (a.mt == 64 && g_opt_warps16 && !a.flags) ? 64 : 128;      // (the flag arrays are per 128-row tile)
  const size_t smem = (size_t)nkc * ((upc / 16) * F_WSUB + mt * KCHUNK * 4) + 256;
  count_launch();
  // 16 warps: the serial gate epilogue is twice as parallel (8 or 4 units per thread instead of 16 or 8)
  if (mt == 64) {
    if (upc == 32) launch_fwd_variant<32, 8, 64>(a, smem, st);
    else launch_fwd_variant<16, 4, 64>(a, smem, st);
  } else if (upc == 32) {
    if (g_opt_warps16) launch_fwd_variant<32, 8, 128>(a, smem, st);
    else launch_fwd_variant<32, 16, 128>(a, smem, st);
  } else {
    if (g_opt_warps16) launch_fwd_variant<16, 4, 128>(a, smem, st);
    else launch_fwd_variant<16, 8, 128>(a, smem, st);
  }
}

// =================================================================================================
// persistent forward sweep: all `steps` time steps of one (batch tile, direction) in ONE kernel.
// A thread-block cluster of H/32 CTAs owns the tile; CTA c keeps its 96 x H slice of W_hh in shared memory for the
// whole sweep, TMEM is allocated once, h_{t-1} of the CTA's own units stays in registers.  After each step every CTA
// publishes its 128 x 32 slice of h_t (fp32 + P16) to global memory, the cluster synchronises (barrier.cluster with
// release/acquire + generic->async proxy fences), and each CTA re-fetches the full 128 x H tile with the TMA engine.
// This removes the per-step kernel boundary (launch gap + prologue + teardown, ~4 us of the ~11 us step).
// =================================================================================================
__global__ void __launch_bounds__(256, 1) gru_seq_fwd_kernel(const GruSeqFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, nkc = (H + KCHUNK - 1) / KCHUNK;
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)nkc * F_WTILE;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sA + (size_t)nkc * F_ATILE);
  uint64_t* abar = wbar + 4;
  uint64_t* done = abar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int c = blockIdx.x, tile = blockIdx.y;          // cluster = all CTAs with the same (y, z): c is the cluster rank
  const GruSeqDirFwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const long Bp = (long)a.tiles * 128;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&wbar[i], 1);
      mbar_init(&abar[i], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.w_p);
      for (int kc = 0; kc < nkc; ++kc) {
        mbar_expect_tx(&wbar[kc], F_WTILE);
        for (int sub = 0; sub < 2; ++sub)
          bulk_g2s(sW + (size_t)kc * F_WTILE + sub * F_WSUB, wp + ((size_t)(2 * c + sub) * nkc + kc) * p16_tile_elems(48), F_WSUB, &wbar[kc]);
      }
    }
    __syncwarp();
  }
  const int r_in = q * 32 + lane;
  const long b = (long)tile * 128 + r_in;
  const int j0 = half * 16, u0 = c * 32 + j0;
  float bhn[16], hprev[16];
  ld16(d.b_hn + u0, bhn);
  ldf16(d.h0 + (long)u0 * d.h0_ld + b, d.h0_ld, hprev);
  const uint32_t idesc = make_idesc_bf16(128, 192);
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  const uint64_t dA = make_desc(smem_u32(sA)), dAl = make_desc(smem_u32(sA) + 128 * KCHUNK * 2), dW = make_desc(smem_u32(sW));

  for (int s = 0; s < a.steps; ++s) {
    const int t = d.reverse ? a.steps - 1 - s : s;
    const int so = (d.out_slots == a.steps) ? t : (s & 1);
    const int sp = (d.out_p_slots == a.steps) ? t : (s & 1);
    const int tprev = d.reverse ? t + 1 : t - 1;
    const int sp_prev = (d.out_p_slots == a.steps) ? tprev : ((s - 1) & 1);
    // input projections of this step (independent of the recurrence): issue before waiting for the cluster
    float gir[16], giz[16], gin[16];
    {
      const float* gi_row = d.gi + (b * d.gi_bs + (long)t * d.gi_ts);
      ldf16(gi_row + (long)u0 * d.gi_ld, d.gi_ld, gir);
      ldf16(gi_row + (long)(H + u0) * d.gi_ld, d.gi_ld, giz);
      ldf16(gi_row + (long)(2 * H + u0) * d.gi_ld, d.gi_ld, gin);
    }
    if (s > 0) cluster_wait_acquire();          // every CTA of the cluster has published its slice of h_{t-1}
    const uint32_t ph = s & 1;
    if (warp == 0) {
      if (elect_one()) {
        fence_proxy_async_all();
        const __nv_bfloat16* hp = (s == 0 ? reinterpret_cast<const __nv_bfloat16*>(d.h0_p)
                                          : reinterpret_cast<const __nv_bfloat16*>(d.out_p) + (size_t)sp_prev * d.out_p_slot_elems) +
                                  (size_t)tile * nkc * p16_tile_elems(128);
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_expect_tx(&abar[kc], F_ATILE);
          bulk_g2s(sA + (size_t)kc * F_ATILE, hp + (size_t)kc * p16_tile_elems(128), F_ATILE, &abar[kc]);
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          if (kc < nkc) {
            if (s == 0) mbar_wait(&wbar[kc], 0);
            mbar_wait(&abar[kc], ph);
            tc_fence_after();
            const int ksteps = min(KCHUNK, H - kc * KCHUNK) / 16;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                const uint32_t ao = kc * F_ATILE + ks * 2 * ATOM_BYTES, wo = kc * F_WTILE + ks * 2 * ATOM_BYTES;
                if (kc == 0 && ks == 0) umma_bf16_c<0>(tmem, desc_advance(dAl, ao), desc_advance(dW, wo), idesc);
                else umma_bf16_c<1>(tmem, desc_advance(dAl, ao), desc_advance(dW, wo), idesc);
                umma_bf16_c<1>(tmem, desc_advance(dA, ao), desc_advance(dW, wo), idesc);
              }
            }
          }
        }
        umma_commit(done);
      }
      __syncwarp();
    }
    mbar_wait(done, ph);
    __syncwarp();
    tc_fence_after();
    float ar[16], az[16], an[16];
    {
      float br[16], bz[16], bn[16];
      const uint32_t cb = taddr + half * 96;       // sub-slice `half`: (r, z, n) x 16 hi-products, then the lo-products
      tmem_ld16(cb, ar);
      tmem_ld16(cb + 16, az);
      tmem_ld16(cb + 32, an);
      tmem_ld16(cb + 48, br);
      tmem_ld16(cb + 64, bz);
      tmem_ld16(cb + 80, bn);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) { ar[i] += br[i]; az[i] += bz[i]; an[i] += bn[i]; }
    }
    float ghn[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float r = gate_sigmoid(gir[i] + ar[i]);
      const float z = gate_sigmoid(giz[i] + az[i]);
      ghn[i] = an[i] + bhn[i];
      const float n = gate_tanh(gin[i] + r * ghn[i]);
      hprev[i] = (1.0f - z) * n + z * hprev[i];
      ar[i] = r; az[i] = z; an[i] = n;
    }
    // publish: P16 slice first (it is what the other CTAs of the cluster wait for), then the fp32 / BPTT copies
    {
      const int kc = u0 / KCHUNK, kk = u0 % KCHUNK;
      __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(d.out_p) + (size_t)sp * d.out_p_slot_elems +
                          ((size_t)tile * nkc + kc) * p16_tile_elems(128);
      st16_p16(tl, 128, r_in, kk, hprev);
    }
    tc_fence_before();                           // TMEM reads of this step are ordered before the next step's MMAs
    if (s + 1 < a.steps) {
      fence_proxy_async_all();                   // generic-proxy stores above -> visible to the peers' TMA (async proxy) reads
      cluster_arrive_release();                  // release at cluster scope publishes them to the other CTAs
    }
    stf16(d.out + (long)u0 * d.out_ld + (long)so * Bp + b, d.out_ld, hprev);
    if (d.sv[0]) {
      const long o = (long)u0 * d.sv_ld + (long)t * Bp + b;
      stf16(d.sv[0] + o, d.sv_ld, ar);
      stf16(d.sv[1] + o, d.sv_ld, az);
      stf16(d.sv[2] + o, d.sv_ld, an);
      stf16(d.sv[3] + o, d.sv_ld, ghn);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

void launch_gru_seq_fwd(const GruSeqFwdArgs& a, cudaStream_t st) {
  const int nkc = (a.H + KCHUNK - 1) / KCHUNK;
  const size_t smem = (size_t)nkc * (F_WTILE + F_ATILE) + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(gru_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(a.H / 32, a.tiles, a.ndir);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.H / 32;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  cudaLaunchKernelEx(&cfg, gru_seq_fwd_kernel, a);
}

// =================================================================================================
// backward step
// =================================================================================================
constexpr int B_APLANE = 128 * KCHUNK * 2;        // 16 KB: one plane of a 128-row K chunk

__global__ void __launch_bounds__(256, 1) gru_step_bwd_kernel(const GruBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, nrb = (H + 127) / 128, nsl = H / 32;
  // B operand (W_hh^T slice): per K chunk kc (2 chunks): hi plane [nrb*128 rows][64], lo plane likewise
  const size_t wchunk = (size_t)nrb * 2 * B_APLANE;
  uint8_t* sW = smem;
  uint8_t* sA = smem + 2 * wchunk;                 // A operand: 2 K chunks x (hi, lo) x 16 KB
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sA + 4 * B_APLANE);
  uint64_t* done = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int c = blockIdx.x, tile = blockIdx.y;
  const GruDirBwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const uint32_t tmem_cols = (H <= 32) ? 32 : (H <= 64) ? 64 : (H <= 128) ? 128 : 256;

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- prologue independent of the previous BPTT step: weights + saved forward activations ----
  if (warp == 0) {
    if (elect_one()) {
      // global layout: [slice c][rb][kc(2)][plane(2)][128x64]; shared layout per kc: hi[rb0..], lo[rb0..]
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.wT_p) + (size_t)c * nrb * 2 * p16_tile_elems(128);
      mbar_expect_tx(wbar, (uint32_t)(2 * wchunk));
      for (int kc = 0; kc < 2; ++kc)
        for (int rb = 0; rb < nrb; ++rb) {
          const __nv_bfloat16* t = wp + ((size_t)rb * 2 + kc) * p16_tile_elems(128);
          bulk_g2s(sW + kc * wchunk + (size_t)rb * B_APLANE, t, B_APLANE, wbar);                                    // hi
          bulk_g2s(sW + kc * wchunk + (size_t)(nrb + rb) * B_APLANE, t + 128 * KCHUNK, B_APLANE, wbar);             // lo
        }
    }
    __syncwarp();
  }
  const int r_in = q * 32 + lane;
  const long b = (long)tile * 128 + r_in;
  const int j0 = half * 16, u0 = c * 32 + j0;
  float r[16], z[16], n[16], ghn[16], hp[16], dh[16];
  {
    const long so = (long)u0 * d.sv_ld + b;
    ldf16(d.sv_r + so, d.sv_ld, r);
    ldf16(d.sv_z + so, d.sv_ld, z);
    ldf16(d.sv_n + so, d.sv_ld, n);
    ldf16(d.sv_ghn + so, d.sv_ld, ghn);
  }
  ldf16(d.h_prev + (long)u0 * d.h_prev_ld + b, d.h_prev_ld, hp);
  if (d.dout) ldf16(d.dout + (long)u0 * d.dout_ld + b, d.dout_ld, dh);
  else {
#pragma unroll
    for (int i = 0; i < 16; ++i) dh[i] = 0.f;
  }

  if (a.pdl && !a.flags) pdl_wait();
  if (a.pdl && !a.flags) pdl_launch_dependents();
  if (a.flags) {
    if (d.flag_in && lane == 0) {
      while (flag_acquire_load(d.flag_in + tile) < a.flag_expected) {
      }
    }
    __syncwarp();
    if (tid == 0) pdl_launch_dependents();
  }

  const long bpad = (long)a.tiles * 128;
  for (int p = 0; p < d.n_parts; ++p) {
    float t[16];
    if (a.flags) {
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = ld_cg(d.parts + (long)p * d.parts_stride + (long)(u0 + i) * d.parts_ld + b);
    } else {
      ldf16(d.parts + (long)p * d.parts_stride + (long)u0 * d.parts_ld + b, d.parts_ld, t);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) dh[i] += t[i];
  }

  float dar[16], daz[16], dan[16], dgn[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float dn = dh[i] * (1.0f - z[i]);
    const float dz = dh[i] * (hp[i] - n[i]);
    dan[i] = dn * (1.0f - n[i] * n[i]);
    daz[i] = dz * z[i] * (1.0f - z[i]);
    dar[i] = dan[i] * ghn[i] * r[i] * (1.0f - r[i]);
    dgn[i] = dan[i] * r[i];
    dh[i] = dh[i] * z[i];                                   // carry
  }
  // A operand (dgh slice) into shared memory: k = g*32 + j ; chunk 0 = gates r,z ; chunk 1 = gate n (k 64..95)
  {
    __nv_bfloat16* a0 = reinterpret_cast<__nv_bfloat16*>(sA);
    __nv_bfloat16* a1 = reinterpret_cast<__nv_bfloat16*>(sA + 2 * B_APLANE);
    st16_p16(a0, 128, r_in, j0, dar);
    st16_p16(a0, 128, r_in, 32 + j0, daz);
    st16_p16(a1, 128, r_in, j0, dgn);
  }
  fence_proxy_async_smem();
  __syncthreads();

  if (warp == 0) {
    if (elect_one()) {
      mbar_wait(wbar, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, H);
      {
        const uint64_t dA0 = make_desc(smem_u32(sA)), dW0 = make_desc(smem_u32(sW));
        const uint32_t wplane = (uint32_t)nrb * B_APLANE, wch = (uint32_t)wchunk;
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) {                 // k-steps 0..3 in chunk 0, 4..5 in chunk 1
          const uint32_t kc = kk >> 2, ko = (kk & 3) * 2 * ATOM_BYTES;
          const uint32_t ao = kc * 2 * B_APLANE + ko, wo = kc * wch + ko;
          if (kk == 0) umma_bf16_c<0>(tmem, desc_advance(dA0, ao + B_APLANE), desc_advance(dW0, wo), idesc);
          else umma_bf16_c<1>(tmem, desc_advance(dA0, ao + B_APLANE), desc_advance(dW0, wo), idesc);
          umma_bf16_c<1>(tmem, desc_advance(dA0, ao), desc_advance(dW0, wo + wplane), idesc);
          umma_bf16_c<1>(tmem, desc_advance(dA0, ao), desc_advance(dW0, wo), idesc);
        }
      }
      umma_commit(done);
    }
    __syncwarp();
  }

  // the carry is part of what the next step reads; the gate gradients are not (they feed the GEMMs after the sweep)
  stf16(d.parts_out + ((long)nsl * H + u0) * bpad + b, bpad, dh);     // carry slot
  auto store_gate_grads = [&]() {
    stf16(d.dgi + (long)u0 * d.dg_ld + b, d.dg_ld, dar);
    stf16(d.dgi + (long)(H + u0) * d.dg_ld + b, d.dg_ld, daz);
    stf16(d.dgi + (long)(2 * H + u0) * d.dg_ld + b, d.dg_ld, dan);
    stf16(d.dgh + (long)u0 * d.dg_ld + b, d.dg_ld, dar);
    stf16(d.dgh + (long)(H + u0) * d.dg_ld + b, d.dg_ld, daz);
    stf16(d.dgh + (long)(2 * H + u0) * d.dg_ld + b, d.dg_ld, dgn);
    if (d.dgi_p) {
      const int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_p) + (size_t)tile * nkc3 * p16_tile_elems(128);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int k = g * H + u0;
        st16_p16(base + (size_t)(k / KCHUNK) * p16_tile_elems(128), 128, r_in, k % KCHUNK, g == 0 ? dar : (g == 1 ? daz : dan));
      }
    }
  };
  if (!a.flags) store_gate_grads();     // overlap with the MMA when the next step waits for kernel completion anyway

  mbar_wait(done, 0);
  __syncwarp();
  tc_fence_after();
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  float* pbase = d.parts_out + (long)c * H * bpad + b;                 // part c, feature-major [H][B_pad]
  const int hh = H / 2;
  for (int c0 = half * hh; c0 < (half + 1) * hh; c0 += 16) {
    float v[16];
    tmem_ld16(taddr + c0, v);
    tmem_ld_wait();
    stf16(pbase + (long)c0 * bpad, bpad, v);
  }
  if (a.flags) {                        // publish the partial sums + carry, then write what only later kernels read
    __threadfence();
    __syncwarp();
    if (lane == 0) flag_release_add(d.flag_out + tile, 1u);
    store_gate_grads();
    if (a.pdl) pdl_wait();              // in-order completion of the chain (see the forward kernel)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, tmem_cols);
}

// =================================================================================================
// persistent backward sweep (BPTT over all steps in one cluster kernel).  CTA c owns hidden units [32c, 32c+32): it keeps
// its 96 x H slice of W_hh (as the [H, 96] B operand) in shared memory, the carry dh*z of its own units in registers, and
// exchanges the H/32 partial products dgh_c W_hh[c-rows, :] with the other CTAs of the cluster through global memory
// (L2-coherent loads) between steps.
// =================================================================================================
// MT = batch rows per CTA (128, or 64 with grid.y = 2 x tiles; see the forward step kernel).
// With M = 64 the accumulator rows live in lanes 0-15 of each TMEM lane quarter: lane l and lane l + 16 of a warp then share
// row l and split the warp group's UPT units between them (U = UPT / 2 units per thread); only the TMEM drain is done by the
// lower half-warp, which hands half of each 16-column batch to its partner lanes with shuffles.
template <int UPT, bool SUM, int MT>
__global__ void __launch_bounds__(32 * 4 * (32 / UPT), 1) gru_seq_bwd_kernel(const GruSeqBwdArgs a) {
  constexpr int APL = MT * KCHUNK * 2;      // one plane of one K chunk of the A operand (dgh slice of this CTA)
  constexpr int U = (MT == 64) ? UPT / 2 : UPT;   // units per thread
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, nrb = (H + 127) / 128, nsl = H / 32;
  const size_t wchunk = (size_t)nrb * 2 * B_APLANE;
  uint8_t* sW = smem;
  uint8_t* sA = smem + 2 * wchunk;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sA + 4 * APL);
  uint64_t* done = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int c = blockIdx.x, tile = blockIdx.y / (128 / MT), hf = blockIdx.y % (128 / MT);
  const GruSeqDirBwd& d = a.d[blockIdx.z];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;                 // half = which UPT-wide group of the slice's 32 units
  const uint32_t tmem_cols = (H <= 32) ? 32 : (H <= 64) ? 64 : (H <= 128) ? 128 : 256;
  const long bpad = (long)a.tiles * 128;
  const size_t slotf = (size_t)bpad * H, pslot = (size_t)(nsl + 1) * slotf;

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(d.wT_p) + (size_t)c * nrb * 2 * p16_tile_elems(128);
      mbar_expect_tx(wbar, (uint32_t)(2 * wchunk));
      for (int kc = 0; kc < 2; ++kc)
        for (int rb = 0; rb < nrb; ++rb) {
          const __nv_bfloat16* t = wp + ((size_t)rb * 2 + kc) * p16_tile_elems(128);
          bulk_g2s(sW + kc * wchunk + (size_t)rb * B_APLANE, t, B_APLANE, wbar);
          bulk_g2s(sW + kc * wchunk + (size_t)(nrb + rb) * B_APLANE, t + 128 * KCHUNK, B_APLANE, wbar);
        }
    }
    __syncwarp();
  }
  const int hl = (MT == 64) ? (lane >> 4) : 0;                                    // which half of the warp group's units
  const bool tm = (MT == 128) || lane < 16;                                       // this lane reads accumulator rows
  const int r_loc = (MT == 128) ? q * 32 + lane : q * 16 + (lane & 15);           // row inside this CTA's M tile
  const int r_in = (MT == 128) ? r_loc : hf * 64 + r_loc;                         // row inside the 128-row tile
  const long b = (long)tile * 128 + r_in;
  const int j0 = half * UPT + hl * U, u0 = c * 32 + j0;
  const uint32_t idesc = make_idesc_bf16(MT, H);
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  const int hh = H / (32 / UPT);            // TMEM columns drained by each warp group
  float carry[U];
  float sum_r[SUM ? U : 1], sum_z[SUM ? U : 1], sum_n[SUM ? U : 1];   // time sums of the input-gate gradients (decoders)
#pragma unroll
  for (int i = 0; i < U; ++i) carry[i] = 0.f;
  if constexpr (SUM) {
#pragma unroll
    for (int i = 0; i < U; ++i) sum_r[i] = sum_z[i] = sum_n[i] = 0.f;
  }
  if (d.dh_last) ldfN<U>(d.dh_last + (long)u0 * d.dh_last_ld + b, d.dh_last_ld, carry);

  for (int s = 0; s < a.steps; ++s) {
    const int t = d.reverse ? s : a.steps - 1 - s;                       // BPTT order = reverse of the forward order
    const bool first_fwd = d.reverse ? (t == a.steps - 1) : (t == 0);
    const int tprev = d.reverse ? t + 1 : t - 1;
#define SEQ_STAMP(i)                   \
  do {                                 \
    if (s == 10) DBG_STAMP(i);         \
  } while (0)
    SEQ_STAMP(0);
    float r[U], z[U], n[U], ghn[U], hp[U], dh[U];
    {
      const long so = (long)u0 * d.sv_ld + (long)t * bpad + b;
      ldfN<U>(d.sv[0] + so, d.sv_ld, r);
      ldfN<U>(d.sv[1] + so, d.sv_ld, z);
      ldfN<U>(d.sv[2] + so, d.sv_ld, n);
      ldfN<U>(d.sv[3] + so, d.sv_ld, ghn);
    }
    if (first_fwd) ldfN<U>(d.h0 + (long)u0 * d.h0_ld + b, d.h0_ld, hp);
    else ldfN<U>(d.out + (long)u0 * d.out_ld + (long)tprev * bpad + b, d.out_ld, hp);
    if (d.dout) ldfN<U>(d.dout + (long)u0 * d.dout_ld + (long)t * bpad + b, d.dout_ld, dh);
    else {
#pragma unroll
      for (int i = 0; i < U; ++i) dh[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) dh[i] += carry[i];
    if (s > 0) {
      cluster_wait_acquire();                                            // all partial products of the previous step are published
      SEQ_STAMP(1);
      const float* pin = d.parts + ((s - 1) & 1) * pslot + (long)u0 * bpad + b;
      for (int p = 0; p < nsl; ++p) {
#pragma unroll
        for (int i = 0; i < U; ++i) dh[i] += ld_cg(pin + (long)p * slotf + (long)i * bpad);
      }
    }
    float dar[U], daz[U], dan[U], dgn[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const float dn = dh[i] * (1.0f - z[i]);
      const float dz = dh[i] * (hp[i] - n[i]);
      dan[i] = dn * (1.0f - n[i] * n[i]);
      daz[i] = dz * z[i] * (1.0f - z[i]);
      dar[i] = dan[i] * ghn[i] * r[i] * (1.0f - r[i]);
      dgn[i] = dan[i] * r[i];
      carry[i] = dh[i] * z[i];
      if constexpr (SUM) { sum_r[i] += dar[i]; sum_z[i] += daz[i]; sum_n[i] += dan[i]; }
    }
    {
      __nv_bfloat16* a0 = reinterpret_cast<__nv_bfloat16*>(sA);
      __nv_bfloat16* a1 = reinterpret_cast<__nv_bfloat16*>(sA + 2 * APL);
      stN_p16<U>(a0, MT, r_loc, j0, dar);
      stN_p16<U>(a0, MT, r_loc, 32 + j0, daz);
      stN_p16<U>(a1, MT, r_loc, j0, dgn);
    }
    SEQ_STAMP(2);
    fence_proxy_async_smem();
    __syncthreads();
    SEQ_STAMP(3);
    const uint32_t ph = s & 1;
    if (warp == 0) {
      if (elect_one()) {
        if (s == 0) mbar_wait(wbar, 0);
        tc_fence_after();
        {
          const uint64_t dA0 = make_desc(smem_u32(sA)), dW0 = make_desc(smem_u32(sW));
          const uint32_t wplane = (uint32_t)nrb * B_APLANE, wch = (uint32_t)wchunk;
#pragma unroll
          for (int kk = 0; kk < 6; ++kk) {                 // k-steps 0..3 in chunk 0, 4..5 in chunk 1
            const uint32_t kc = kk >> 2, ko = (kk & 3) * 2 * ATOM_BYTES;
            const uint32_t ao = kc * 2 * APL + ko, wo = kc * wch + ko;
            if (kk == 0) umma_bf16_c<0>(tmem, desc_advance(dA0, ao + APL), desc_advance(dW0, wo), idesc);
            else umma_bf16_c<1>(tmem, desc_advance(dA0, ao + APL), desc_advance(dW0, wo), idesc);
            umma_bf16_c<1>(tmem, desc_advance(dA0, ao), desc_advance(dW0, wo + wplane), idesc);
            umma_bf16_c<1>(tmem, desc_advance(dA0, ao), desc_advance(dW0, wo), idesc);
          }
        }
        umma_commit(done);
        SEQ_STAMP(4);
      }
      __syncwarp();
    }
    // outputs that do not need the MMA (read by later kernels: weight-gradient GEMMs, dx of the layer below)
    {
      const long o = (long)t * bpad + b;
      stfN<U>(d.dgi + (long)u0 * d.dg_ld + o, d.dg_ld, dar);
      stfN<U>(d.dgi + (long)(H + u0) * d.dg_ld + o, d.dg_ld, daz);
      stfN<U>(d.dgi + (long)(2 * H + u0) * d.dg_ld + o, d.dg_ld, dan);
      stfN<U>(d.dgh + (long)u0 * d.dg_ld + o, d.dg_ld, dar);
      stfN<U>(d.dgh + (long)(H + u0) * d.dg_ld + o, d.dg_ld, daz);
      stfN<U>(d.dgh + (long)(2 * H + u0) * d.dg_ld + o, d.dg_ld, dgn);
    }
    if (d.dgi_p) {
      const int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_p) + (size_t)t * d.dgi_p_slot_elems +
                            (size_t)tile * nkc3 * p16_tile_elems(128);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int k = g * H + u0;
        stN_p16<U>(base + (size_t)(k / KCHUNK) * p16_tile_elems(128), 128, r_in, k % KCHUNK, g == 0 ? dar : (g == 1 ? daz : dan));
      }
    }
    float* pout = d.parts + (s & 1) * pslot;
    if (s == a.steps - 1) stfN<U>(pout + ((long)nsl * H + u0) * bpad + b, bpad, carry);   // final carry -> dh0 reduction

    SEQ_STAMP(5);
    mbar_wait(done, ph);
    __syncwarp();
    SEQ_STAMP(6);
    tc_fence_after();
    float* pbase = pout + (long)c * H * bpad + b;
    for (int c0 = half * hh; c0 < (half + 1) * hh; c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if constexpr (MT == 128) {
        stfN<16>(pbase + (long)c0 * bpad, bpad, v);
      } else {
        // rows live in lanes 0-15; lane l + 16 takes over columns c0 + 8 .. c0 + 15 of row l
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float hi8 = __shfl_sync(0xffffffffu, v[8 + i], lane & 15);
          w[i] = tm ? v[i] : hi8;
        }
        stfN<8>(pbase + (long)(c0 + hl * 8) * bpad, bpad, w);
      }
    }
    tc_fence_before();
    SEQ_STAMP(7);
    if (s + 1 < a.steps) cluster_arrive_release();   // release at cluster scope publishes the partial sums to the peers
    SEQ_STAMP(8);
  }
  if constexpr (SUM) {
    stfN<U>(d.dgi_sum + (long)u0 * bpad + b, bpad, sum_r);
    stfN<U>(d.dgi_sum + (long)(H + u0) * bpad + b, bpad, sum_z);
    stfN<U>(d.dgi_sum + (long)(2 * H + u0) * bpad + b, bpad, sum_n);
    const int nkc3 = (3 * H + KCHUNK - 1) / KCHUNK;
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(d.dgi_sum_p) + (size_t)tile * nkc3 * p16_tile_elems(128);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int k = g * H + u0;
      stN_p16<U>(base + (size_t)(k / KCHUNK) * p16_tile_elems(128), 128, r_in, k % KCHUNK, g == 0 ? sum_r : (g == 1 ? sum_z : sum_n));
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, tmem_cols);
}

template <int UPT, bool SUM, int MT>
static void launch_seq_bwd_variant(const GruSeqBwdArgs& a, cudaStream_t st) {
  const int nrb = (a.H + 127) / 128;
  const size_t smem = (size_t)2 * nrb * 2 * B_APLANE + 4 * (MT * KCHUNK * 2) + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(gru_seq_bwd_kernel<UPT, SUM, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(a.H / 32, a.tiles * (128 / MT), a.ndir);
  cfg.blockDim = dim3(128 * (32 / UPT));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.H / 32;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, gru_seq_bwd_kernel<UPT, SUM, MT>, a);
}

void launch_gru_seq_bwd(const GruSeqBwdArgs& a_in, cudaStream_t st) {
  GruSeqBwdArgs a = a_in;
  a.dbg = g_dbg_buffer;
  const bool w16 = g_opt_warps16 && (a.H % 64 == 0);       // each of the 4 warp groups drains H/4 (multiple of 16) columns
  const bool m64 = a.mt == 64 && w16;
  const bool sum = a.d[0].dgi_sum != nullptr;              // (set for both directions or for none)
  count_launch();
  if (m64) {
    if (sum) launch_seq_bwd_variant<8, true, 64>(a, st);
    else launch_seq_bwd_variant<8, false, 64>(a, st);
  } else if (w16) {
    if (sum) launch_seq_bwd_variant<8, true, 128>(a, st);
    else launch_seq_bwd_variant<8, false, 128>(a, st);
  } else {
    if (sum) launch_seq_bwd_variant<16, true, 128>(a, st);
    else launch_seq_bwd_variant<16, false, 128>(a, st);
  }
}

void launch_gru_step_bwd(const GruBwdArgs& a, cudaStream_t st) {
  const int nrb = (a.H + 127) / 128;
  const size_t smem = (size_t)2 * nrb * 2 * B_APLANE + 4 * B_APLANE + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(gru_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch_cfg(cfg, attr, dim3(a.H / 32, a.tiles, a.ndir), 256, smem, st, a.pdl);
  count_launch();
  cudaLaunchKernelEx(&cfg, gru_step_bwd_kernel, a);
}

}  // namespace vb
