// Common device helpers for the B200 (sm_100a) RNN-VAE hot path.
//   * raw PTX wrappers: mbarrier, cp.async.bulk (TMA engine, 1-D bulk form), tcgen05 (alloc / mma / commit / ld),
//     programmatic dependent launch (griddepcontrol)
//   * the "P16" packed-operand layout shared by every tensor-core kernel in this library
//
// P16 layout (bf16 hi/lo split operands, UMMA canonical K-major, SWIZZLE_NONE):
//   logical matrix X[R, K] (K = contraction dim); x = hi + lo with hi = bf16(x), lo = bf16(x - hi).
//   Stored as tiles of RB rows x 64 k-elements; tile (rb, kc) is one contiguous block
//       [plane: hi, lo][r8: RB/8][k8: 8][row-in-atom: 8][elt-in-row: 8]      (bf16)
//   i.e. an "atom" (UMMA core matrix) is 8 rows x 16 bytes = 128 contiguous bytes; atoms adjacent along K are
//   128 B apart (descriptor LBO), 8-row groups are 1024 B apart (descriptor SBO).  One tile = RB*256 bytes and is
//   fetched with a single cp.async.bulk.  Tile (rb, kc) starts at ((rb * KC) + kc) * RB * 128 elements.
//   A product a*w is evaluated as a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM
//   (error ~2^-16 relative per product; measured end-to-end error vs fp32 reference is ~3e-6).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vb {

constexpr int KCHUNK = 64;                       // k-elements per tile
constexpr int ATOM_BYTES = 128;                  // 8 rows x 8 bf16
constexpr uint32_t DESC_LBO = 128;               // bytes between K-adjacent atoms
constexpr uint32_t DESC_SBO = 1024;              // bytes between 8-row groups

__host__ __device__ inline size_t p16_tile_elems(int RB) { return (size_t)RB * KCHUNK * 2; }      // hi + lo
__host__ __device__ inline size_t p16_tile_bytes(int RB) { return (size_t)RB * KCHUNK * 2 * 2; }
__host__ __device__ inline size_t p16_bytes(int R, int K, int RB) {
  size_t nrb = (R + RB - 1) / RB, nkc = (K + KCHUNK - 1) / KCHUNK;
  return nrb * nkc * p16_tile_bytes(RB);
}
// element offset (in bf16 elements) of (r, k) inside one plane of a tile; r < RB, k < 64
__host__ __device__ inline int p16_in_tile(int r, int k) { return ((r >> 3) * 8 + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7); }

// ------------------------------------------------------------------------------------------------
// shared-address helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------------------------------------
// TMA engine, 1-D bulk copy global -> shared with mbarrier transaction accounting (SASS: UBLKCP)
// size must be a multiple of 16, both addresses 16-byte aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// thread-block cluster barrier (split arrive / wait) and cross-proxy fences for data exchanged through global memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// orders this thread's generic-proxy accesses (st.global / ld.global) against async-proxy accesses (TMA bulk copies)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// vectorised split-K accumulation: one 16-byte reduction per lane (PTX ISA 8.1, sm_90+); p must be 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }   // L2-coherent load (bypasses the non-coherent L1)
// inter-kernel hand-over flags (PDL chains without griddepcontrol.wait): release-increment / acquire-poll at gpu scope
__device__ __forceinline__ void flag_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int flag_acquire_load(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle, LBO/SBO as in the P16 layout, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);              // start address      bits [0,14)
  d |= (uint64_t)(DESC_LBO >> 4) << 16;                  // leading byte off   bits [16,30)
  d |= (uint64_t)(DESC_SBO >> 4) << 32;                  // stride byte off    bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version bits [46,48)
  return d;                                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}
// instruction descriptor for kind::f16, A/B = BF16, D = F32, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with a compile-time accumulate flag (no run-time predicate set-up in the issue loop)
template <int ACC>
__device__ __forceinline__ void umma_bf16_c(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC)
      : "memory");
}
// A operand from tensor memory (lane r = row r of the M = 128 tile, 8 columns per k-step of 16: column i holds the bf16 pair
// k = 2 i, 2 i + 1), B from shared memory; compile-time accumulate flag
template <int ACC>
__device__ __forceinline__ void umma_bf16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "n"(ACC)
      : "memory");
}
// descriptor of the same tile `bytes` further into shared memory (the 14-bit start-address field cannot overflow:
// shared memory is < 256 KB)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }



// One lane of a converged warp (elect.sync).  MMA-issuing code is guarded by this instead of `lane == 0`: the compiler then
// treats the region as single-lane code and keeps warp-uniform values (UMMA descriptors, barrier addresses) in uniform
// registers.  Under `if (lane == 0)` every descriptor that depended on a loop variable was computed in vector registers and
// each tcgen05.mma was preceded by an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: 73 instead of 46 cycles per MMA in the
// resident-weight GRU sweeps (measured, tools/bench_mma/mma_contention.cu gives the 46-cycle floor).
__device__ __forceinline__ bool elect_one() {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok));
  return ok != 0;
}

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the same signal delivered to the mbarrier at the same shared-memory offset in every CTA of the cluster whose rank bit is set in
// cta_mask: the tensor pipe itself tells the peers "my MMAs up to here are complete" - no thread has to observe the completion first
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// TMEM -> registers: warp reads its own 32 lanes, 16 consecutive 32-bit columns per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// bf16 hi/lo split
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// split 8 consecutive floats into two 16-byte vectors (hi, lo)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(x[i], h[i], l[i]);
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vb
