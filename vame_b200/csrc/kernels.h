// Internal launch prototypes shared between the .cu translation units (not part of the public C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

// number of kernels this library has enqueued (all streams); bench.py reports the delta over its timed region
extern long g_launch_count;
extern int g_opt_pdl;       // 1: chain the recurrent steps with programmatic dependent launch
extern unsigned long long* g_dbg_buffer;   // device buffer for kernel timeline stamps or nullptr
extern int g_opt_flags;     // 1: PDL-chained step kernels hand h over through release/acquire flags (tail of step t overlaps t+1);
                            //    default 0: measured equal to griddepcontrol.wait (tools/gpu_probe_graph.py, profiles/)
extern int g_opt_warps16;   // 1: 16 warps per CTA (8 hidden units per thread) in the default recurrent kernels
extern int g_opt_streams;   // 1: run independent branches of a step on internal side streams
extern int g_opt_m64;       // 1: recurrent kernels take 64 batch rows per CTA (twice the CTAs) when the sweep fits the GPU
extern int g_opt_slice16;   // 1: forward step kernels own 16 hidden units per CTA (twice the CTAs, N = 96 MMAs) when the sweep fits the GPU
extern int g_opt_persistent; // bit 0 / bit 1: run the forward / backward recurrent sweeps as one persistent cluster kernel
                             // instead of one PDL-chained kernel per time step (measured: per-step wins forward, persistent backward)   // 1: run independent branches of a step on internal side streams
extern int g_opt_rw;         // bit 0 / bit 1: forward / backward sweeps by the resident-weight cluster kernels (gru_rw.cu) when applicable
extern int g_opt_rw2;        // 1: H = 256 sweeps use the barrier-free rw kernels (bulk-copy / mbarrier exchange)
extern int g_opt_rw_sw;     // bit 0 / bit 1: forward / backward private-mode sweeps hand their global stores to extra store warps through tensor memory
extern int g_opt_rw_priv;   // 1: training sweeps of H = 256 layers use the private interchange layouts (needs rw = 3 and rw2 = 1)
extern int g_opt_rw_waves;
extern int g_opt_rw_ng;      // 16-row groups per cluster of the H = 256 rw kernels (0 = automatic)
extern int g_opt_rw_exp;      // measurement experiments (wrong results), see GruSeqFwdArgs::exp   // rw kernels are used while their grid fits this many waves of the 132 cluster-schedulable SMs
inline void count_launch(int n = 1) { g_launch_count += n; }

// ---- pack.cu -------------------------------------------------------------------------------------
// value(r, k) = src[i * ld + j] with (i, j) = transposed ? (k, r) : (r, k); zero outside i < nrows_src, j < ncols_src
void launch_pack_p16(const float* src, long ld, int transposed, int R, int K, int R_src, int K_src, const int* row_map,
                     const int* col_map, int RB, void* out, cudaStream_t st);
// plain pack of a row-major [R_src, K] matrix that also accumulates rowsum[r] += sum_k src[r, k]
void launch_pack_p16_rowsum(const float* src, long ld, int R, int K, int R_src, void* out, float* rowsum, cudaStream_t st);

// batched pack jobs (one launch): kind 0 = generic pack_p16, 1 / 2 = W_hh forward / backward slices (R = H), 3 = fused bias (R = H),
// 4 / 5 = W_hh in the resident-weight forward / backward format of gru_rw.cu (R = H, H % 64 == 0),
// 6 = W_hh in the row-resident inference format of gru_rows.cu (R = H, H % 64 == 0)
struct PackJob {
  const float* src; const float* src2; void* out;
  long ld;
  int kind, transposed, R, K, R_src, K_src, RB, block_begin;
};
struct PackJobs {
  PackJob j[40];
  int n;
};
void launch_pack_jobs(PackJobs& jobs, cudaStream_t st);   // resets jobs.n to 0

// W_hh [3H, H] -> forward-step slices (mode 0: P16 RB=96, rows (slice c, gate g, unit j) = W_hh[g*H + 32c + j, :], K = H)
//               or backward-step slices (mode 1: per slice c a P16 RB=128 matrix [rows = unit u, K = 128]:
//                                        value(u, k = g*32 + j) = W_hh[g*H + 32c + j, u], zero for k >= 96)
void launch_pack_whh(const float* w_hh, int H, int mode, void* out, cudaStream_t st);
// fused input-projection bias: out[d*3H + g*H + u] = b_ih_d[g*H+u] + (g < 2 ? b_hh_d[g*H+u] : 0)
void launch_bias_fuse(const float* b_ih0, const float* b_hh0, const float* b_ih1, const float* b_hh1, int H, float* out, cudaStream_t st);
void launch_h0_prepare(const float* src, int D, int B, int B_pad, int H, float* h32, void* hp, cudaStream_t st);
void launch_bt_to_tb(const float* src, int B, int T, int C, long bs, long ts, int B_pad, float* dst, cudaStream_t st);
void launch_bt_to_tb_p16(const float* src, int B, int T, int C, long bs, long ts, int B_pad, float* dst, void* dst_p, cudaStream_t st);
void launch_tb_to_bt(const float* src, int B, int T, int C, int B_pad, float* dst, cudaStream_t st);

// ---- gemm.cu -------------------------------------------------------------------------------------
struct GemmSeg {
  const void* p;        // P16 tiles (RB = 128): tile(rb, kc) = p + (rb * rb_stride + kc) * tile_elems
  long rb_stride;       // in tiles
  int nkc;              // number of 64-wide K chunks in this segment
};
struct GemmArgs {
  GemmSeg a[4], b[4];   // A: [M, K] , B: [N, K]; K = concatenation of up to 4 segments (A and B split independently)
  int M, N;             // logical output extent (rows >= M / cols >= N are masked)
  float* C;             // fp32 row-major (C[m*ldc + n]) or, with c_fm, feature-major (C[n*ldc + m])
  long ldc;
  int c_fm;             // 1: feature-major; 2: feature-major planes of 256 features in the lane-major blocks of the H = 256 rw
                        //    sweeps ([t][16-row group][64-unit CTA][16-unit warp][lane][8 rows], gru_rw.cu pv_block), rows = t*pv_bp + b
  int pv_bp;            // padded batch (c_fm == 2)
  const float* bias;    // [N] or nullptr (added only by split 0)
  int atomic;           // 1: red.add into C (split-K / accumulation), 0: plain store
  int splits;           // k splits (work items = tiles x splits)
  int max_ctas;         // upper bound of the persistent grid (0 = all SMs): SMs held by a concurrently running sweep are not
                        // available, and a persistent grid larger than the free SMs would serialise into a second wave
};
void launch_gemm_p16(const GemmArgs& g, cudaStream_t st);

// ---- gru.cu --------------------------------------------------------------------------------------
// fp32 interchange arrays of the step kernels are FEATURE-MAJOR: element (feature f, row r) lives at base + f*ld + r with
// r = slot*B_pad + b.  A TMEM lane is a batch row, so for a fixed feature the 32 lanes of a warp touch 128 contiguous bytes.
struct GruDirFwd {
  const void* w_p;        // P16 (RB=96) W_hh slices: [H/32][KC][2][96x64], slice rows = r,z,n gates of 32 units
  const float* b_hn;      // [H] hidden bias of the n gate (b_hr, b_hz are folded into gi)
  const float* gi;        // input projections incl. biases, feature-major [3H][gi_ld]; row of (b, t) = b*gi_bs + t*gi_ts
  long gi_ld, gi_bs, gi_ts;
  int t;                  // time index of this step for this direction
  const void* h_in_p;     // P16 (RB=128) [tiles][KC][2][128x64]
  const float* h_in;      // fp32 feature-major [H][h_in_ld], already offset to the slot
  long h_in_ld;
  void* h_out_p;          // P16 like h_in_p
  float* h_out;           // fp32 feature-major [H][h_out_ld], already offset to the slot
  long h_out_ld;
  float* sv_r; float* sv_z; float* sv_n; float* sv_ghn;   // saved gates for BPTT, feature-major [H][sv_ld] at this t's slot (or nullptr)
  long sv_ld;
  // hand-over flags (a.flags != 0): per (tile) counters; this step waits until flag_in[tile] == flag_expected (nullptr: no wait)
  // and adds 1 per warp to flag_out[tile] once its slice of h is published
  const unsigned int* flag_in; unsigned int* flag_out;
};
struct GruFwdArgs {
  GruDirFwd d[2];
  int ndir, H, tiles;
  int upc;                // hidden units per CTA: 16 or 32 (0 = 32)
  int mt;                 // batch rows per CTA: 64 or 128 (0 = 128)
  int pdl;                // launch with programmatic stream serialization
  int flags;              // 1: the recurrence dependency is carried by flag_in/flag_out instead of griddepcontrol.wait, so the
                          //    tail of step t (saved-gate stores, teardown) overlaps step t+1
  unsigned int flag_expected;
  unsigned long long* dbg;   // optional device buffer for %globaltimer stamps (filled in by the launcher)
};
void launch_gru_step_fwd(const GruFwdArgs& a, cudaStream_t st);

// ---- persistent (whole-sweep) forward kernel: one thread-block cluster of H/32 CTAs per (batch tile, direction) ----
struct GruSeqDirFwd {
  const void* w_p; const float* b_hn;
  const void* w_rw;                                     // resident-weight format (pack kind 4), used by launch_gru_rw_fwd
  const void* w_rows;                                   // row-resident inference format (pack kind 6), used by launch_gru_rows_fwd
  const float* gi; long gi_ld, gi_bs, gi_ts;
  const float* h0; long h0_ld; const void* h0_p;       // initial state: fp32 feature-major + P16
  float* out; long out_ld; int out_slots;               // fp32 h sequence, slot offset = slot*B_pad (slots = steps or 2)
  void* out_p; int out_p_slots; long out_p_slot_elems;  // P16 h sequence
  float* sv[4]; long sv_ld;                             // saved gates r,z,n,ghn (slot t) or nullptr
  int reverse;                                          // 1: processes t = steps-1 .. 0
  // "private" training mode of the H = 256 rw kernels (both sweeps of a layer run on them): sv[] and out hold the
  // lane-major layout shared by the forward and backward kernels (fully coalesced 512-byte accesses per warp instruction),
  // the transposed P16 operand of the weight-gradient GEMMs is written directly, the final state also as fp32 feature-major
  int priv;
  void* outT_p; long outT_nk;                           // P16 [H rows, K = steps * B_pad], nk = K / 64
  float* hfin;                                          // [H][B_pad]
};
struct GruSeqFwdArgs {
  GruSeqDirFwd d[2];
  int ndir, H, tiles, steps;
  unsigned long long* dbg;   // optional %globaltimer stamps of step 10 (filled in by the launcher)
  // Measurement experiments of the H = 256 rw kernels (option "rw_exp", tools/gpu_probe_sweeps.py; results become WRONG unless
  // noted): 1 skip out / saved-gate stores (non-private path), 2 skip the input-projection loads, 4 skip the P16 h stores
  // (non-private path), 8 push only the hi plane, 16 skip every global store, 32 per-CTA %globaltimer stamps (results
  // stay right), 64 (BPTT) skip the saved-activation loads, 128 (BPTT) release instead of relaxed remote arrive (results
  // stay right).  What they showed is summarised in DESIGN.md section 4.1.
  int exp;
};
void launch_gru_seq_fwd(const GruSeqFwdArgs& a, cudaStream_t st);

// ---- persistent (whole-sweep) backward kernel, same cluster decomposition ----
struct GruSeqDirBwd {
  const void* wT_p;
  const void* wT_rw;                           // resident-weight format (pack kind 5), used by launch_gru_rw_bwd
  float* dh0_out;                              // rw kernel only: gradient of the initial state, feature-major [H][B_pad]
  const float* dh_last; long dh_last_ld;       // gradient of the final hidden state, feature-major [H][ld], or nullptr
  const float* dout; long dout_ld;             // per-step output gradients, feature-major [H][dout_ld] (slot t at + t*B_pad) or nullptr
  const float* sv[4]; long sv_ld;              // saved gates r,z,n,ghn, [H][sv_ld]
  const float* out; long out_ld;               // forward h sequence [H][out_ld] (h_prev of step t = slot t -/+ 1)
  const float* h0; long h0_ld;                 // initial state [H][h0_ld]
  float* parts;                                // [2 slots][H/32 + 1][H][B_pad] partial sums exchanged between the CTAs
  float* dgi; float* dgh; long dg_ld;          // outputs, feature-major [3H][dg_ld] (slot t at + t*B_pad)
  void* dgi_p; long dgi_p_slot_elems;          // optional P16 copy of dgi per t
  float* dgi_sum; void* dgi_sum_p;             // optional: sum over t of dgi, feature-major [3H][B_pad] and P16 [B_pad rows, K = 3H]
                                               // (decoders: the GRU input is z at every step, so dz needs only the time sum)
  // private mode (see GruSeqDirFwd): sv / out are read in the lane-major layout; instead of fp32 dgi / dgh the kernel writes the
  // transposed P16 operands of the weight-gradient GEMMs and adds the bias gradients (sums over t and b) into db_ih / db_hh
  int priv;
  int dout_pv;                                 // dout is in the lane-major block layout (GemmArgs::c_fm == 2): one coalesced 256-bit load
  void* dghT_p; void* dgiT_p; long gT_nk;      // P16 [3H rows, K = steps * B_pad] (dgiT_p may be nullptr), nk = K / 64
  float* db_ih; float* db_hh;                  // [3H] each, accumulated with atomicAdd
  int reverse;                                 // direction of the FORWARD recurrence (0: t ascending) -> BPTT runs the other way
};
struct GruSeqBwdArgs {
  GruSeqDirBwd d[2];
  int ndir, H, tiles, steps;
  int mt;                    // batch rows per CTA: 64 or 128 (0 = 128)
  unsigned long long* dbg;   // optional %globaltimer stamps of step 10 (filled in by the launcher)
  int exp;                   // measurement experiments only (option "rw_exp", results become wrong)
};
void launch_gru_seq_bwd(const GruSeqBwdArgs& a, cudaStream_t st);

// ---- gru_rw.cu: resident-weight sweeps (4-CTA clusters, swap-AB, h / partial sums exchanged through DSMEM) ----
bool rw_applicable(int H, int tiles);
int rw_groups_per_cluster(int H, int tiles);   // 1 or 2 (H = 256: 32 rows per cluster when one wave of 16-row clusters does not fit)
bool rw_priv_mode(int H, int tiles);   // training sweeps use the private interchange layouts (see GruSeqDirFwd::priv)
size_t rw_whh_bytes(int H);       // packed W_hh of one direction, forward format
size_t rw_whhT_bytes(int H);      // ... backward format
void launch_gru_rw_fwd(const GruSeqFwdArgs& a, cudaStream_t st);
void launch_gru_rw_bwd(const GruSeqBwdArgs& a, cudaStream_t st);
// ---- gru_rows.cu: row-resident forward sweep for large inference batches (128 rows per persistent CTA, W_hh streamed from L2) ----
extern int g_opt_side_split;       // 2: the two directions' weight-gradient GEMMs get half of the persistent-grid cap each; 1: the whole cap each
extern int g_opt_side_sms;         // > 0: persistent-grid cap of the weight-gradient GEMMs that run beside a sweep (0 = built-in rule)
extern int g_opt_rows;             // 1: inference sweeps of H = 256 layers that do not fit the rw kernels use gru_rows_fwd_kernel
bool rows_fwd_applicable(int H, int tiles);
void launch_gru_rows_fwd(const GruSeqFwdArgs& a, cudaStream_t st);
unsigned int rows_timeouts();      // bounded waits that gave up since the library was loaded (0 unless there is a protocol bug)
unsigned int rw_timeouts();
int rw_timeout_info(int i);        // first time-out: 0 site id, 1 parity, 2-4 blockIdx, 5 threadIdx.x
void rw_timeouts_reset();       // bounded waits that gave up since the library was loaded (0 unless there is a protocol bug)

struct GruDirBwd {
  const void* wT_p;       // P16 (RB=128) [H/32 slices][rb: H_pad/128][KC=2][2][128x64]: B[n=u, k=g*32+j] = W_hh[g*H+32c+j, u]
  const float* parts;     // incoming dh pieces, feature-major: part p, unit u, row b at parts + p*parts_stride + u*parts_ld + b;
  int n_parts;            //   their sum is the dh flowing into this step
  long parts_stride, parts_ld;
  const float* dout;      // upstream gradient of this step's output, feature-major [H][dout_ld] at this t's slot, or nullptr
  long dout_ld;
  const float* sv_r; const float* sv_z; const float* sv_n; const float* sv_ghn;   // [H][sv_ld] at this t's slot
  long sv_ld;
  const float* h_prev;    // [H][h_prev_ld]
  long h_prev_ld;
  float* parts_out;       // [H/32 + 1][H][B_pad]: slice partial sums of dgh @ W_hh, last slot = carry dh*z
  float* dgi;             // feature-major [3H][dg_ld] at this t's slot (dgi_r, dgi_z, dgi_n)
  float* dgh;             // same layout (dgi_r, dgi_z, dgi_n * r)
  long dg_ld;
  void* dgi_p;            // optional P16 (RB=128) A-operand copy of dgi for the dx GEMM: [tiles][KC3H][2][128x64] slot of this t
  const unsigned int* flag_in; unsigned int* flag_out;   // as in GruDirFwd
};
struct GruBwdArgs {
  GruDirBwd d[2];
  int ndir, H, tiles;
  int pdl;
  int flags;
  unsigned int flag_expected;
};
void launch_gru_step_bwd(const GruBwdArgs& a, cudaStream_t st);

// ---- simt.cu -------------------------------------------------------------------------------------
// C[m,n] (=|+=) sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]); generic strides cover NT / NN / TN forms.
struct SgemmArgs {
  const float* A; long sam, sak;
  const float* B; long sbk, sbn;
  float* C; long scm, scn;
  const float* bias;
  int M, N, K;
  int accumulate;        // 1: atomicAdd into C (also enables split-K), 0: store
  int splitk;            // >= 1
  float alpha;
};
void launch_sgemm(const SgemmArgs& g, cudaStream_t st);

}  // namespace vb
