// fp32 -> P16 (bf16 hi/lo, UMMA canonical tiles) packing kernels and small layout helpers.
#include "common.cuh"
#include "kernels.h"

namespace vb {

// One thread produces one 16-byte atom row (8 consecutive k) of both planes.
// value(r, k) = src[rm * ld + km]  (or src[km * ld + rm] when transposed), rm = row_map ? row_map[r] : r, idem km;
// out-of-range / negative map entries give 0.
// rowsum != nullptr (non-transposed sources only): additionally accumulates rowsum[r] += sum_k value(r, k) — the bias
// gradient of a feature-major gate-gradient matrix comes for free with the pack that reads it anyway.
__device__ __forceinline__ void pack_p16_body(const float* __restrict__ src, long ld, int transposed, int R, int K, int R_src,
                                              int K_src, const int* __restrict__ row_map, const int* __restrict__ col_map, int RB,
                                              __nv_bfloat16* __restrict__ out, long bid, long nb, float* __restrict__ rowsum = nullptr) {
  const int nkc = (K + KCHUNK - 1) / KCHUNK;
  const int nrb = (R + RB - 1) / RB;
  const long total = (long)nrb * RB * nkc * 8;           // atom rows
  for (long idx = bid * (long)blockDim.x + threadIdx.x; idx < total; idx += nb * blockDim.x) {
    // transposed sources are contiguous along r, others along k: pick the thread order that coalesces the reads
    long r, k8g;
    if (transposed) {
      r = idx % ((long)nrb * RB);
      k8g = idx / ((long)nrb * RB);
    } else {
      k8g = idx % ((long)nkc * 8);
      r = idx / ((long)nkc * 8);
    }
    const int kbase = (int)k8g * 8;
    float v[8];
    const int rm = (r < R) ? (row_map ? row_map[r] : (int)r) : -1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kbase + i;
      const int km = (k < K) ? (col_map ? col_map[k] : k) : -1;
      float x = 0.f;
      if (rm >= 0 && km >= 0 && rm < R_src && km < K_src) x = transposed ? src[(long)km * ld + rm] : src[(long)rm * ld + km];
      v[i] = x;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int rb = (int)(r / RB), rr = (int)(r % RB);
    const int kc = kbase / KCHUNK, kk = kbase % KCHUNK;
    __nv_bfloat16* tile = out + ((size_t)rb * nkc + kc) * p16_tile_elems(RB);
    const int off = p16_in_tile(rr, kk);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + (size_t)RB * KCHUNK + off) = lo;
    if (rowsum) {
      // non-transposed order: the lanes of a warp hold consecutive 8-wide k groups; a row has nkc*8 groups.  Reduce over the
      // lanes that share this thread's row, one atomic per (warp, row).
      float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      const unsigned mask = __activemask();
      const long r0 = __shfl_sync(mask, r, 0);
      const bool uniform = __all_sync(mask, r == r0);
      if (uniform && mask == 0xffffffffu) {
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0 && r < R) atomicAdd(rowsum + r, s);
      } else if (r < R) {
        atomicAdd(rowsum + r, s);
      }
    }
  }
}
__global__ void pack_p16_rowsum_kernel(const float* __restrict__ src, long ld, int R, int K, int R_src, int RB,
                                       __nv_bfloat16* __restrict__ out, float* __restrict__ rowsum) {
  pack_p16_body(src, ld, 0, R, K, R_src, K, nullptr, nullptr, RB, out, blockIdx.x, gridDim.x, rowsum);
}
__global__ void pack_p16_kernel(const float* __restrict__ src, long ld, int transposed, int R, int K, int R_src, int K_src,
                                const int* __restrict__ row_map, const int* __restrict__ col_map, int RB,
                                __nv_bfloat16* __restrict__ out) {
  pack_p16_body(src, ld, transposed, R, K, R_src, K_src, row_map, col_map, RB, out, blockIdx.x, gridDim.x);
}

// Fast path of the plain (row-major source, no maps) pack for K % 64 == 0, 16-byte aligned rows: a warp owns 8 rows x `kspan`
// columns; lane = (row % 8) * 4 + (8-wide k group % 4), so every load instruction reads 8 x 128 contiguous bytes and every store
// instruction writes 512 contiguous bytes of a tile plane ([r8][k8][8 rows][8 elts]); the optional row sums are kept in
// registers over the whole span (one atomic per row and span).
__global__ void __launch_bounds__(256) pack_rows_fast_kernel(const float* __restrict__ src, long ld, int R, int K, int R_src,
                                                             __nv_bfloat16* __restrict__ out, float* __restrict__ rowsum, int kspan) {
  const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r7 = lane >> 2, k8l = lane & 3;
  const int nkc = K / KCHUNK, nrg = ((R + 127) / 128) * 16, nks = (K + kspan - 1) / kspan;
  for (long job = (long)blockIdx.x * wpb + warp; job < (long)nrg * nks; job += (long)gridDim.x * wpb) {
    const int rg = (int)(job / nks), ks = (int)(job % nks);
    const int r = rg * 8 + r7;
    const bool valid = r < R_src;
    const float* row = src + (long)r * ld;
    __nv_bfloat16* tile_row = out + (size_t)(r >> 7) * nkc * p16_tile_elems(128);
    const int rr = r & 127;
    const int kend = min(K, (ks + 1) * kspan);
    float s = 0.f;
#pragma unroll 4
    for (int k = ks * kspan + k8l * 8; k < kend; k += 32) {
      float v[8];
      if (valid) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + k + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      __nv_bfloat16* tile = tile_row + (size_t)(k >> 6) * p16_tile_elems(128);
      const int off = p16_in_tile(rr, k & 63);
      *reinterpret_cast<uint4*>(tile + off) = hi;
      *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
      s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    if (rowsum) {
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (k8l == 0 && r < R && valid) atomicAdd(rowsum + r, s);
    }
  }
}
static bool pack_fast_ok(const float* src, long ld, int K) {
  return (K % KCHUNK) == 0 && (ld % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}
static void launch_pack_rows_fast(const float* src, long ld, int R, int K, int R_src, void* out, float* rowsum, cudaStream_t st) {
  const int kspan = K >= 4096 ? 512 : (K >= 512 ? 256 : 64);
  const long jobs = (long)((R + 127) / 128) * 16 * ((K + kspan - 1) / kspan);
  long blocks = (jobs + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_launch();
  pack_rows_fast_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld, R, K, R_src, (__nv_bfloat16*)out, rowsum, kspan);
}

void launch_pack_p16(const float* src, long ld, int transposed, int R, int K, int R_src, int K_src, const int* row_map,
                     const int* col_map, int RB, void* out, cudaStream_t st) {
  if (!transposed && !row_map && !col_map && RB == 128 && K_src == K && pack_fast_ok(src, ld, K)) {
    launch_pack_rows_fast(src, ld, R, K, R_src, out, nullptr, st);
    return;
  }
  const int nkc = (K + KCHUNK - 1) / KCHUNK, nrb = (R + RB - 1) / RB;
  long total = (long)nrb * RB * nkc * 8;
  int threads = 256;
  long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  count_launch();
  pack_p16_kernel<<<(unsigned)blocks, threads, 0, st>>>(src, ld, transposed, R, K, R_src, K_src, row_map, col_map, RB,
                                                        (__nv_bfloat16*)out);
}

void launch_pack_p16_rowsum(const float* src, long ld, int R, int K, int R_src, void* out, float* rowsum, cudaStream_t st) {
  if (pack_fast_ok(src, ld, K)) {
    launch_pack_rows_fast(src, ld, R, K, R_src, out, rowsum, st);
    return;
  }
  const int nkc = (K + KCHUNK - 1) / KCHUNK, nrb = (R + 127) / 128;
  long total = (long)nrb * 128 * nkc * 8;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  count_launch();
  pack_p16_rowsum_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld, R, K, R_src, 128, (__nv_bfloat16*)out, rowsum);
}

// ------------------------------------------------------------------------------------------------
// W_hh slices for the recurrent step kernels (see kernels.h)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_whh_body(const float* __restrict__ w, int H, int mode, __nv_bfloat16* __restrict__ out, long bid,
                                              long nb) {
  const int nsl = H / 32;
  if (mode == 0) {
    // forward operand: slices of 16 hidden units = 48 rows [r(16) z(16) n(16)], one P16 tile (hi plane, lo plane) per
    // (slice, K chunk).  A step CTA that owns 16 (32) units loads 1 (2) consecutive slices per chunk; one UMMA descriptor
    // then spans [hi48 ; lo48] (N = 96) or [hi48 ; lo48 ; hi48 ; lo48] (N = 192).
    const int nkc = (H + KCHUNK - 1) / KCHUNK;
    const long total = (long)(H / 16) * 48 * nkc * 8;
    for (long idx = bid * (long)blockDim.x + threadIdx.x; idx < total; idx += nb * blockDim.x) {
      const int k8g = (int)(idx % (nkc * 8));
      const int p = (int)(idx / (nkc * 8));            // packed row
      const int c = p / 48, g = (p % 48) / 16, j = p % 16;
      const int srow = g * H + 16 * c + j, kbase = k8g * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (kbase + i < H) ? w[(long)srow * H + kbase + i] : 0.f;
      uint4 hi, lo;
      split8(v, hi, lo);
      __nv_bfloat16* tile = out + ((size_t)c * nkc + kbase / KCHUNK) * p16_tile_elems(48);
      const int off = p16_in_tile(p % 48, kbase % KCHUNK);
      *reinterpret_cast<uint4*>(tile + off) = hi;
      *reinterpret_cast<uint4*>(tile + 48 * KCHUNK + off) = lo;
    }
  } else {
    const int nrb = (H + 127) / 128;
    const long total = (long)nsl * nrb * 128 * 2 * 8;
    for (long idx = bid * (long)blockDim.x + threadIdx.x; idx < total; idx += nb * blockDim.x) {
      const int u = (int)(idx % (nrb * 128));          // coalesced along u (source columns)
      const long rest = idx / (nrb * 128);
      const int k8g = (int)(rest % 16), c = (int)(rest / 16);
      const int kbase = k8g * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kbase + i, g = k / 32, j = k % 32;
        v[i] = (k < 96 && u < H) ? w[(long)(g * H + 32 * c + j) * H + u] : 0.f;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      __nv_bfloat16* tile = out + (((size_t)c * nrb + u / 128) * 2 + kbase / KCHUNK) * p16_tile_elems(128);
      const int off = p16_in_tile(u % 128, kbase % KCHUNK);
      *reinterpret_cast<uint4*>(tile + off) = hi;
      *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
    }
  }
}
// W_hh [3H, H] -> resident-weight A tiles of gru_rw.cu (H % 64 == 0, UC = H/4 units per cluster CTA c).  A tile is 128 rows x
// 64 k bf16 in the UMMA canonical K-major layout; tile row (= TMEM lane) L = 32 q + l carries the HI plane (l < 16) or the LO
// plane (l >= 16) of one weight row.
//   mode 0 (forward):  tiles [c][gate g][kc]; lane L <-> W_hh[g*H + c*UC + 16 q + (l & 15), :], zero if 16 q >= UC
//   mode 1 (backward): tiles [c][m][kc], K = 3 UC (k = g*UC + j); lane L <-> input unit ui = 64 m + 16 q + (l & 15):
//                      value(ui, k) = W_hh[g*H + c*UC + j, ui]
__device__ __forceinline__ void pack_whh_rw_body(const float* __restrict__ w, int H, int mode, __nv_bfloat16* __restrict__ out, long bid,
                                                 long nb) {
  const int NKC = H / 64, UC = H / 4;
  const int NKB = (3 * UC + KCHUNK - 1) / KCHUNK;
  const long ntiles = mode == 0 ? 4L * 3 * NKC : 4L * NKC * NKB;
  const long total = ntiles * 128 * 8;
  for (long idx = bid * (long)blockDim.x + threadIdx.x; idx < total; idx += nb * blockDim.x) {
    const int k8 = (int)(idx & 7), L = (int)((idx >> 3) & 127);
    const long tile = idx >> 10;
    const int qq = L >> 5, l = L & 31, plane = l >> 4, jj = l & 15;
    float v[8];
    if (mode == 0) {
      const int kc = (int)(tile % NKC), g = (int)((tile / NKC) % 3), c = (int)(tile / (3 * NKC));
      const int j = 16 * qq + jj;
      const long row = (long)g * H + c * UC + j;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (j < UC) ? w[row * H + kc * KCHUNK + k8 * 8 + i] : 0.f;
    } else {
      const int kc = (int)(tile % NKB), m = (int)((tile / NKB) % NKC), c = (int)(tile / ((long)NKB * NKC));
      const int ui = 64 * m + 16 * qq + jj;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * KCHUNK + k8 * 8 + i;
        v[i] = (k < 3 * UC) ? w[((long)(k / UC) * H + c * UC + (k % UC)) * H + ui] : 0.f;
      }
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(out + (size_t)tile * (128 * KCHUNK) + p16_in_tile(L, k8 * 8)) = plane ? lo : hi;
  }
}
// Row-resident inference format (gru_rows.cu): 32-unit slices, one P16 tile (RB = 96: hi plane, lo plane) per (slice c, 64-k chunk);
// tile row p = 32 g + j <-> W_hh[g*H + 32 c + j, :] (r, z, n gate rows of the slice's units).  H % 64 == 0.
__device__ __forceinline__ void pack_whh_rows_body(const float* __restrict__ w, int H, __nv_bfloat16* __restrict__ out, long bid, long nb) {
  const int nkc = H / KCHUNK;
  const long total = (long)(H / 32) * 96 * nkc * 8;
  for (long idx = bid * (long)blockDim.x + threadIdx.x; idx < total; idx += nb * blockDim.x) {
    const int k8g = (int)(idx % (nkc * 8));
    const int p = (int)(idx / (nkc * 8));              // packed row
    const int c = p / 96, g = (p % 96) / 32, j = p % 32;
    const int srow = g * H + 32 * c + j, kbase = k8g * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w[(long)srow * H + kbase + i];
    uint4 hi, lo;
    split8(v, hi, lo);
    __nv_bfloat16* tile = out + ((size_t)c * nkc + kbase / KCHUNK) * p16_tile_elems(96);
    const int off = p16_in_tile(p % 96, kbase % KCHUNK);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + 96 * KCHUNK + off) = lo;
  }
}
__global__ void pack_whh_kernel(const float* __restrict__ w, int H, int mode, __nv_bfloat16* __restrict__ out) {
  pack_whh_body(w, H, mode, out, blockIdx.x, gridDim.x);
}
void launch_pack_whh(const float* w_hh, int H, int mode, void* out, cudaStream_t st) {
  count_launch();
  pack_whh_kernel<<<148, 256, 0, st>>>(w_hh, H, mode, (__nv_bfloat16*)out);
}
__global__ void bias_fuse_kernel(const float* __restrict__ bi0, const float* __restrict__ bh0, const float* __restrict__ bi1,
                                 const float* __restrict__ bh1, int H, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * H) return;
  const int d = i / (3 * H), r = i % (3 * H);
  const float* bi = d ? bi1 : bi0;
  const float* bh = d ? bh1 : bh0;
  out[i] = bi[r] + (r < 2 * H ? bh[r] : 0.f);
}
void launch_bias_fuse(const float* b_ih0, const float* b_hh0, const float* b_ih1, const float* b_hh1, int H, float* out, cudaStream_t st) {
  count_launch();
  bias_fuse_kernel<<<(6 * H + 255) / 256, 256, 0, st>>>(b_ih0, b_hh0, b_ih1, b_hh1, H, out);
}

// ------------------------------------------------------------------------------------------------
// batched packing: one launch executes a table of independent pack jobs (the ~50 weight re-packs after every optimizer
// step would otherwise be ~50 serial 3-us launches)
// ------------------------------------------------------------------------------------------------
__global__ void pack_jobs_kernel(const PackJobs jobs) {
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.j[j + 1].block_begin) ++j;
  const PackJob& J = jobs.j[j];
  const long bid = blockIdx.x - J.block_begin;
  const long nb = (j + 1 < jobs.n ? jobs.j[j + 1].block_begin : (int)gridDim.x) - J.block_begin;
  if (J.kind == 0) {
    pack_p16_body(J.src, J.ld, J.transposed, J.R, J.K, J.R_src, J.K_src, nullptr, nullptr, J.RB, (__nv_bfloat16*)J.out, bid, nb);
  } else if (J.kind == 1 || J.kind == 2) {
    pack_whh_body(J.src, J.R, J.kind - 1, (__nv_bfloat16*)J.out, bid, nb);
  } else if (J.kind == 4 || J.kind == 5) {
    pack_whh_rw_body(J.src, J.R, J.kind - 4, (__nv_bfloat16*)J.out, bid, nb);
  } else if (J.kind == 6) {
    pack_whh_rows_body(J.src, J.R, (__nv_bfloat16*)J.out, bid, nb);
  } else {                                         // fused projection bias: out[i] = b_ih[i] + (i < 2H ? b_hh[i] : 0), R = H
    const int H = J.R;
    for (long i = bid * blockDim.x + threadIdx.x; i < 3L * H; i += nb * blockDim.x)
      ((float*)J.out)[i] = J.src[i] + (i < 2 * H ? J.src2[i] : 0.f);
  }
}
void launch_pack_jobs(PackJobs& jobs, cudaStream_t st) {
  if (jobs.n == 0) return;
  int total = 0;
  for (int i = 0; i < jobs.n; ++i) {
    PackJob& J = jobs.j[i];
    long work;
    if (J.kind == 0) work = (long)((J.R + J.RB - 1) / J.RB) * J.RB * ((J.K + KCHUNK - 1) / KCHUNK) * 8;
    else if (J.kind == 6) work = (long)(J.R / 32) * 96 * (J.R / KCHUNK) * 8;
    else if (J.kind == 1) work = (long)(J.R / 32) * 96 * ((J.R + KCHUNK - 1) / KCHUNK) * 8;
    else if (J.kind == 2) work = (long)(J.R / 32) * ((J.R + 127) / 128) * 128 * 16;
    else if (J.kind == 4 || J.kind == 5) work = 4L * 3 * (J.R / 64) * 128 * 8;
    else work = 3L * J.R;
    long nb = (work + 255) / 256;
    if (nb > 96) nb = 96;
    if (nb < 1) nb = 1;
    J.block_begin = total;
    total += (int)nb;
  }
  count_launch();
  pack_jobs_kernel<<<total, 256, 0, st>>>(jobs);
  jobs.n = 0;
}

// ------------------------------------------------------------------------------------------------
// h0 preparation: src is a flat fp32 buffer viewed as [D][B][H] (for the decoder this is the raw
// reinterpretation of the (B, 2H) latent_to_hidden output, vame/model/rnn_model.py:104,137); src == nullptr -> zeros.
// Writes fp32 feature-major [D][H][B_pad] and packed P16 [D][tiles][KC][2][128x64].
// ------------------------------------------------------------------------------------------------
__global__ void h0_prepare_kernel(const float* __restrict__ src, int D, int B, int B_pad, int H, float* __restrict__ h32,
                                  __nv_bfloat16* __restrict__ hp) {
  const int nkc = (H + KCHUNK - 1) / KCHUNK;
  const long total = (long)D * B_pad * nkc * 8;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k8g = (int)(idx % (nkc * 8));
    const long rest = idx / (nkc * 8);
    const int b = (int)(rest % B_pad), d = (int)(rest / B_pad);
    const int kbase = k8g * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kbase + i;
      v[i] = (src && b < B && k < H) ? src[((long)d * B + b) * H + k] : 0.f;
    }
    if (h32 && kbase < H) {              // feature-major [D][H][B_pad]
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (kbase + i < H) h32[((long)d * H + kbase + i) * B_pad + b] = v[i];
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int tile_i = b / 128, rr = b % 128;
    const int kc = kbase / KCHUNK, kk = kbase % KCHUNK;
    __nv_bfloat16* tile = hp + (((size_t)d * (B_pad / 128) + tile_i) * nkc + kc) * p16_tile_elems(128);
    const int off = p16_in_tile(rr, kk);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
  }
}

void launch_h0_prepare(const float* src, int D, int B, int B_pad, int H, float* h32, void* hp, cudaStream_t st) {
  const int nkc = (H + KCHUNK - 1) / KCHUNK;
  long total = (long)D * B_pad * nkc * 8;
  int threads = 256;
  long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  h0_prepare_kernel<<<(unsigned)blocks, threads, 0, st>>>(src, D, B, B_pad, H, h32, (__nv_bfloat16*)hp);
}

// ------------------------------------------------------------------------------------------------
// (B, T, C) <-> time-major (T, B_pad, C) transposes (tiny tensors: x, fut, pred)
// ------------------------------------------------------------------------------------------------
__global__ void bt_to_tb_kernel(const float* __restrict__ src, int B, int T, int C, long src_bstride, long src_tstride,
                                int B_pad, float* __restrict__ dst) {
  const long total = (long)T * B_pad * C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long r = idx / C;
    const int b = (int)(r % B_pad), t = (int)(r / B_pad);
    dst[idx] = (b < B) ? src[b * src_bstride + t * src_tstride + c] : 0.f;
  }
}
__global__ void tb_to_bt_kernel(const float* __restrict__ src, int B, int T, int C, int B_pad, float* __restrict__ dst) {
  const long total = (long)B * T * C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long r = idx / C;
    const int t = (int)(r % T), b = (int)(r / T);
    dst[idx] = src[((long)t * B_pad + b) * C + c];
  }
}
// the same + the result as P16 [T*B_pad rows, K = C -> padded]: one thread per 8-wide k atom of one (t, b) row
__global__ void bt_to_tb_p16_kernel(const float* __restrict__ src, int B, int T, int C, long src_bstride, long src_tstride, int B_pad,
                                    float* __restrict__ dst, __nv_bfloat16* __restrict__ dst_p) {
  const int nkc = (C + KCHUNK - 1) / KCHUNK, apr = nkc * 8;
  const long total = (long)T * B_pad * apr;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % apr) * 8;
    const long r = idx / apr;
    const int b = (int)(r % B_pad), t = (int)(r / B_pad);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = k0 + i;
      v[i] = 0.f;
      if (c < C) {
        v[i] = (b < B) ? src[b * src_bstride + t * src_tstride + c] : 0.f;
        dst[r * C + c] = v[i];
      }
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    __nv_bfloat16* tile = dst_p + ((size_t)(r >> 7) * nkc + (k0 >> 6)) * p16_tile_elems(128);
    const int off = p16_in_tile((int)(r & 127), k0 & 63);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
  }
}
void launch_bt_to_tb_p16(const float* src, int B, int T, int C, long bs, long ts, int B_pad, float* dst, void* dst_p, cudaStream_t st) {
  const long total = (long)T * B_pad * ((C + KCHUNK - 1) / KCHUNK) * 8;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_launch();
  bt_to_tb_p16_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, B, T, C, bs, ts, B_pad, dst, (__nv_bfloat16*)dst_p);
}
void launch_bt_to_tb(const float* src, int B, int T, int C, long bs, long ts, int B_pad, float* dst, cudaStream_t st) {
  long total = (long)T * B_pad * C;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_launch();
  bt_to_tb_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, B, T, C, bs, ts, B_pad, dst);
}
void launch_tb_to_bt(const float* src, int B, int T, int C, int B_pad, float* dst, cudaStream_t st) {
  long total = (long)B * T * C;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_launch();
  tb_to_bt_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, B, T, C, B_pad, dst);
}

}  // namespace vb
