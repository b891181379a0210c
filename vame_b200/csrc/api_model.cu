// Model-level C-ABI: forward / loss / backward / optimizer / embedding of the RNN-VAE, expressed as sequences of
// kernel launches on the caller's stream.  Host code only: all arithmetic happens in gemm.cu / gru.cu / simt.cu / pack.cu.
#include "../../include/vame_b200.h"
#include "api_common.h"
#include "common.cuh"
#include "kernels.h"
#include "model_layout.h"
#include "simt.h"

namespace vb {

typedef __nv_bfloat16 bf16;

// ================================================================================================
// workspace
// ================================================================================================
struct GruBuf {
  int steps, H, In;
  float* gi; long gi_bs, gi_ts, gi_ld;   // feature-major [6H][gi_ld]; row of (b,t) = b*gi_bs + t*gi_ts
  float* out[2];  int out_slots;        // fp32 h sequence, feature-major [H][slots*B_pad]; slots = steps (training) or 2
  void* out_p[2]; int out_p_slots;      // P16 h sequence [slots][tiles][nkc]
  float* sv[2][4];
  float* h0[2]; void* h0_p[2];          // initial state (zeros for the encoder)
  bool dout_pv = false;                 // the upstream gradient of the per-step outputs was written in the lane-major block layout
  float* dgi[2]; float* dgh[2]; void* dgi_p[2];
  float* parts[2];
  int parts_n;                          // partial-sum slots that hold the final dh0 after the last backward sweep (1 after an rw sweep)
  bool priv;                            // training sweeps use the private interchange layouts of the H = 256 rw kernels (kernels.h)
  float* hfin[2];                       // priv: final hidden state, fp32 feature-major [H][B_pad]
  void* dghT_p[2]; void* dgiT_p[2]; void* outT_p[2]; void* h0T_p[2];
  unsigned int* flags;                  // [2 dirs][steps][tiles] hand-over counters of the per-step kernels
};

struct DecBuf {
  GruBuf g;
  float* hid;        // [B, 2H] latent_to_hidden output (flat buffer aliased as [2][B][H] by the .view quirk)
  float* pred_tb;    // [steps*B_pad, F]
  float* dpred_tb;   // [steps*B_pad, F]
  void* dpred_p;     // P16 [steps*B_pad, K=F]
  void* dpredT_p;    // P16 [F rows, K = steps*B_pad]
  float* ddec;       // feature-major [2H][steps*B_pad]
  float* dgi_sum[2]; // feature-major [3H][B_pad]
  void* dgi_sum_p[2]; void* dgi_sumT_p[2];
  float* dhid;       // [B, 2H]
  void* dhid_p; void* dhidT_p;
  float* dz;         // [B, Z]
  float* target_tb;  // [steps*B_pad, F]  future target / optional clean reconstruction target (default: x_tb)
};

struct Ws {
  int B, B_pad, tiles;
  float* x_tb; void* x_p; void* xT_p;
  float* zeros_f32; void* zeros_p; size_t zeros_p_bytes; int Hmax;   // [Hmax][B_pad] fp32 zeros, P16 zero tiles
  void* hid_p[4];                       // P16 copies of an externally supplied hidden (vame_lambda_forward)
  GruBuf e0, e1;
  float* lin;        // [B_pad, 2Z]
  float* z; float* mu; float* logvar; float* eps;   // [B, Z]
  void* z_p; void* zT_p;
  DecBuf dec[2];
  float* dz_km;      // [B, Z]
  float* dlin; void* dlin_p; void* dlinT_p;
  void* hidT_p[4];   // [H rows, K = B_pad] per hidden piece
  float* dhidden;    // feature-major [4H][B_pad]
  float* dx1;        // feature-major [2H][T*B_pad]
  double* acc;       // [8]
  double* prior_state;   // [CP_STATE_DOUBLES] eigenvector basis of the previous k-means-prior call (training workspaces only)
  size_t bytes;
};

static void carve_gru(Arena& A, GruBuf& g, int steps, int H, int In, int B_pad, bool training, bool per_t_p16, bool gi_full,
                      bool want_dgi_p, bool want_dgiT, bool own_h0) {
  const int tiles = B_pad / 128, nkc = nkc_of(H), nkc3 = nkc_of(3 * H);
  const size_t slotf = (size_t)B_pad * H;
  const size_t slotp = (size_t)tiles * nkc * p16_tile_bytes(128);
  const long rows = (long)steps * B_pad;
  g.steps = steps; g.H = H; g.In = In;
  g.priv = training && rw_priv_mode(H, tiles);
  if (gi_full) {
    g.gi = A.f32((size_t)rows * 6 * H);
    g.gi_bs = 1; g.gi_ts = B_pad; g.gi_ld = rows;
  } else {
    g.gi = A.f32((size_t)B_pad * 6 * H);            // time-invariant input (decoder): one row per sample
    g.gi_bs = 1; g.gi_ts = 0; g.gi_ld = B_pad;
  }
  g.out_slots = training ? steps : 2;
  g.out_p_slots = per_t_p16 ? steps : 2;
  g.flags = (unsigned int*)A.raw((size_t)2 * steps * tiles * sizeof(unsigned int));
  if (own_h0) {                               // both directions adjacent: written by ONE h0_prepare launch
    g.h0[0] = A.f32(2 * slotf);
    g.h0_p[0] = A.raw(2 * slotp);
    g.h0[1] = g.h0[0] ? g.h0[0] + slotf : nullptr;
    g.h0_p[1] = g.h0_p[0] ? (char*)g.h0_p[0] + slotp : nullptr;
  }
  for (int d = 0; d < 2; ++d) {
    g.out[d] = A.f32(slotf * g.out_slots);
    g.out_p[d] = A.raw(slotp * g.out_p_slots);
    if (training) {
      g.hfin[d] = A.f32(slotf);
      for (int i = 0; i < 4; ++i) g.sv[d][i] = A.f32(slotf * steps);
      g.dgi[d] = A.f32((size_t)rows * 3 * H);
      g.dgh[d] = A.f32((size_t)rows * 3 * H);
      g.dgi_p[d] = want_dgi_p ? A.raw((size_t)steps * tiles * nkc3 * p16_tile_bytes(128)) : nullptr;
      g.parts[d] = A.f32((size_t)2 * (H / 32 + 1) * slotf);
      g.dghT_p[d] = A.raw(p16_bytes(3 * H, (int)rows, 128));
      g.dgiT_p[d] = want_dgiT ? A.raw(p16_bytes(3 * H, (int)rows, 128)) : nullptr;
      g.outT_p[d] = A.raw(p16_bytes(H, (int)rows, 128));
      g.h0T_p[d] = own_h0 ? A.raw(p16_bytes(H, B_pad, 128)) : nullptr;
    }
  }
}

static Ws carve_ws(const vame_dims& d, int B, bool training, void* base) {
  Ws w{};
  Arena A(base);
  const int T = d.time_window, F = d.num_features, Z = d.zdims, H = d.hidden_enc;
  w.B = B; w.B_pad = pad128(B); w.tiles = w.B_pad / 128;
  const int Bp = w.B_pad;
  const long rows = (long)T * Bp;
  int Hmax = H;
  if (d.hidden_rec > Hmax) Hmax = d.hidden_rec;
  if (d.future_decoder && d.hidden_pred > Hmax) Hmax = d.hidden_pred;
  w.x_tb = A.f32((size_t)rows * F);
  w.x_p = A.raw(p16_bytes((int)rows, F, 128));
  w.xT_p = training ? A.raw(p16_bytes(F, (int)rows, 128)) : nullptr;
  w.zeros_f32 = A.f32((size_t)Bp * Hmax);
  {
    size_t zp = (size_t)w.tiles * nkc_of(Hmax) * p16_tile_bytes(128);
    size_t zt = p16_bytes(Hmax, Bp, 128);                 // also used as the transposed h0 (= 0) of the encoder
    w.zeros_p_bytes = zp > zt ? zp : zt;
    w.zeros_p = A.raw(w.zeros_p_bytes);
    w.Hmax = Hmax;
  }
  for (int i = 0; i < 4; ++i) w.hid_p[i] = A.raw(p16_bytes(Bp, H, 128));
  carve_gru(A, w.e0, T, H, F, Bp, training, /*per_t_p16=*/true, /*gi_full=*/true, /*dgi_p=*/false, /*dgiT=*/true, /*own_h0=*/false);
  carve_gru(A, w.e1, T, H, 2 * H, Bp, training, false, true, /*dgi_p=*/true, /*dgiT=*/true, false);
  w.lin = A.f32((size_t)Bp * 2 * Z);
  w.z = A.f32((size_t)B * Z); w.mu = A.f32((size_t)B * Z); w.logvar = A.f32((size_t)B * Z); w.eps = A.f32((size_t)B * Z);
  w.z_p = A.raw(p16_bytes(Bp, Z, 128));
  w.zT_p = training ? A.raw(p16_bytes(Z, Bp, 128)) : nullptr;
  for (int i = 0; i < (d.future_decoder ? 2 : 1); ++i) {
    DecBuf& D = w.dec[i];
    const int Hd = i == 0 ? d.hidden_rec : d.hidden_pred;
    const int steps = i == 0 ? T : d.future_steps;
    const long r = (long)steps * Bp;
    carve_gru(A, D.g, steps, Hd, Z, Bp, training, true, /*gi_full=*/false, false, false, /*own_h0=*/true);
    D.hid = A.f32((size_t)B * 2 * Hd);
    D.pred_tb = A.f32((size_t)r * F);
    D.target_tb = A.f32((size_t)r * F);
    if (training) {
      D.dpred_tb = A.f32((size_t)r * F);
      D.dpred_p = A.raw(p16_bytes((int)r, F, 128));
      D.dpredT_p = A.raw(p16_bytes(F, (int)r, 128));
      D.ddec = A.f32((size_t)r * 2 * Hd);
      for (int dd = 0; dd < 2; ++dd) {
        D.dgi_sum[dd] = A.f32((size_t)Bp * 3 * Hd);
        D.dgi_sum_p[dd] = A.raw(p16_bytes(Bp, 3 * Hd, 128));
        D.dgi_sumT_p[dd] = A.raw(p16_bytes(3 * Hd, Bp, 128));
      }
      D.dhid = A.f32((size_t)B * 2 * Hd);
      D.dhid_p = A.raw(p16_bytes(Bp, 2 * Hd, 128));
      D.dhidT_p = A.raw(p16_bytes(2 * Hd, Bp, 128));
      D.dz = A.f32((size_t)B * Z);
    }
  }
  if (training) {
    w.dz_km = A.f32((size_t)B * Z);
    w.dlin = A.f32((size_t)Bp * 2 * Z);
    w.dlin_p = A.raw(p16_bytes(Bp, 2 * Z, 128));
    w.dlinT_p = A.raw(p16_bytes(2 * Z, Bp, 128));
    for (int i = 0; i < 4; ++i) w.hidT_p[i] = A.raw(p16_bytes(H, Bp, 128));
    w.dhidden = A.f32((size_t)Bp * 4 * H);
    w.dx1 = A.f32((size_t)rows * 2 * H);
  }
  w.acc = (double*)A.raw(8 * sizeof(double));
  w.prior_state = training ? (double*)A.raw((size_t)CP_STATE_DOUBLES * sizeof(double)) : nullptr;
  w.bytes = (A.off + 1023) & ~(size_t)1023;
  return w;
}

// ================================================================================================
// side streams: independent branches of a step (future decoder, k-means prior, weight-gradient GEMMs) run concurrently
// with the latency-bound recurrent sweeps; fork/join are event edges, so the whole thing stays CUDA-graph capturable.
// ================================================================================================
struct Side {
  cudaStream_t s[4];       // [0] future decoder, [1] weight-gradient work, [2] k-means prior, [3] second weight-gradient stream
  cudaEvent_t ev[64];
  int nev;
};
static Side& side() {
  static Side S{};
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 4; ++i) cudaStreamCreateWithFlags(&S.s[i], cudaStreamNonBlocking);
    for (int i = 0; i < 64; ++i) cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming);
    init = true;
  }
  return S;
}
// ---- optional section timeline (measurement hook): cudaEvents on the caller's stream at section boundaries -------------
struct Timeline {
  bool on = false;
  int n = 0;
  cudaEvent_t ev[64];
  const char* name[64];
};
static Timeline& timeline() {
  static Timeline T;
  return T;
}
static inline void mark(cudaStream_t st, const char* name) {
  Timeline& T = timeline();
  if (!T.on || T.n >= 64) return;
  if (!T.ev[T.n]) cudaEventCreate(&T.ev[T.n]);
  cudaEventRecord(T.ev[T.n], st);
  T.name[T.n++] = name;
}

// make `to` wait for everything enqueued on `from` so far
static inline void edge(cudaStream_t from, cudaStream_t to) {
  if (from == to) return;
  Side& S = side();
  cudaEvent_t e = S.ev[S.nev++ & 63];
  cudaEventRecord(e, from);
  cudaStreamWaitEvent(to, e, 0);
}

// ================================================================================================
// small host helpers
// ================================================================================================
static inline GemmSeg seg(const void* p, long rb_stride, int nkc) { return GemmSeg{p, rb_stride, nkc}; }

// SMs that the side-stream GEMMs can count on while a recurrent sweep runs on the main stream (set by vame_backward)
static int g_side_sms = 148;
struct GemmB {
  GemmArgs g{};
  int na = 0, nb = 0;
  GemmB& A(const void* p, long rbs, int nkc) { g.a[na++] = seg(p, rbs, nkc); return *this; }
  GemmB& Bm(const void* p, long rbs, int nkc) { g.b[nb++] = seg(p, rbs, nkc); return *this; }
  void run(int M, int N, float* C, long ldc, const float* bias, int atomic, int splits, cudaStream_t st) {
    g.M = M; g.N = N; g.C = C; g.ldc = ldc; g.bias = bias; g.atomic = atomic; g.splits = splits;
    g.max_ctas = atomic ? g_side_sms : 0;      // accumulating (weight-gradient) GEMMs run beside a sweep
    launch_gemm_p16(g, st);
  }
  // feature-major output: C[n*ldc + m]
  void run_fm(int M, int N, float* C, long ldc, const float* bias, cudaStream_t st) {
    g.c_fm = 1;
    run(M, N, C, ldc, bias, 0, 1, st);
  }
  // feature-major planes of 256 features, each permuted into the lane-major blocks of the H = 256 rw sweeps (rows = t*Bp + b)
  void run_pv(int M, int N, float* C, long ldc, int Bp, cudaStream_t st) {
    g.c_fm = 2; g.pv_bp = Bp;
    run(M, N, C, ldc, nullptr, 0, 1, st);
  }
};
// plain (non-transposed) pack of a row-major [R_src, K] matrix into P16 with R rows (zero padded)
static inline void pack_rows(const float* src, long ld, int R, int K, int R_src, void* out, cudaStream_t st) {
  launch_pack_p16(src, ld, 0, R, K, R_src, K, nullptr, nullptr, 128, out, st);
}
// transposed pack: source [K_src rows, R cols] row-major -> P16 [R rows, K] with K = padded K_src
static inline void pack_T(const float* src, long ld, int R, int K, int K_src, void* out, cudaStream_t st) {
  launch_pack_p16(src, ld, 1, R, K, R, K_src, nullptr, nullptr, 128, out, st);
}
static inline int splits_for(int M, int N, int nkc) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  int s = g_side_sms / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > nkc) s = nkc;
  if (s > 32) s = 32;
  return s;
}

// ================================================================================================
// weights
// ================================================================================================
// All weight re-packs of one optimizer step are queued as jobs and executed by 2-3 launches of pack_jobs_kernel.
struct JobQ {
  PackJobs jobs{};
  cudaStream_t st;
  explicit JobQ(cudaStream_t s) : st(s) { jobs.n = 0; }
  void push(const PackJob& j) {
    if (jobs.n == 40) launch_pack_jobs(jobs, st);
    jobs.j[jobs.n++] = j;
  }
  void rows(const float* src, long ld, int R, int K, int R_src, void* out) {           // like pack_rows
    PackJob j{}; j.src = src; j.out = out; j.ld = ld; j.kind = 0; j.transposed = 0; j.R = R; j.K = K; j.R_src = R_src; j.K_src = K; j.RB = 128;
    push(j);
  }
  void T(const float* src, long ld, int R, int K, int K_src, void* out) {               // like pack_T
    PackJob j{}; j.src = src; j.out = out; j.ld = ld; j.kind = 0; j.transposed = 1; j.R = R; j.K = K; j.R_src = R; j.K_src = K_src; j.RB = 128;
    push(j);
  }
  void whh(const float* w, int H, int mode, void* out) {
    PackJob j{}; j.src = w; j.out = out; j.kind = 1 + mode; j.R = H;
    push(j);
  }
  void whh_rw(const float* w, int H, int mode, void* out) {
    PackJob j{}; j.src = w; j.out = out; j.kind = 4 + mode; j.R = H;
    push(j);
  }
  void whh_rows(const float* w, int H, void* out) {
    PackJob j{}; j.src = w; j.out = out; j.kind = 6; j.R = H;
    push(j);
  }
  void bias(const float* b_ih, const float* b_hh, int H, float* out) {
    PackJob j{}; j.src = b_ih; j.src2 = b_hh; j.out = out; j.kind = 3; j.R = H;
    push(j);
  }
  void flush() { launch_pack_jobs(jobs, st); }
};

// formats the FORWARD pass of a bi-GRU layer reads / formats only the backward pass reads
// slices = false: skip the W_hh formats of the slice kernels (gru.cu) - the caller knows that the sweeps of this batch size run
// on the resident-weight kernels (vame_pack_weights_train)
static void pack_gru_fwd(const float* P, const GruOff& o, const GruPacked& W, JobQ& q, bool slices = true) {
  const int H = o.H, In = o.In;
  for (int d = 0; d < 2; ++d) {
    if (slices || !W.whh_rw[d]) q.whh(P + o.whh[d], H, 0, W.whh_p[d]);
    if (W.whh_rw[d]) q.whh_rw(P + o.whh[d], H, 0, W.whh_rw[d]);
    if (slices && W.whh_rows[d]) q.whh_rows(P + o.whh[d], H, W.whh_rows[d]);     // large-batch inference only: not in the train-loop re-pack
    q.bias(P + o.bih[d], P + o.bhh[d], H, W.bias_gi + (size_t)d * 3 * H);
  }
  // both directions' W_ih are adjacent in the flat buffer -> one [6H, In] matrix; K split in wih_nseg column blocks
  const int Ks = In / W.wih_nseg;
  for (int s = 0; s < W.wih_nseg; ++s) q.rows(P + o.wih[0] + (long)s * Ks, In, 6 * H, Ks, 6 * H, W.wih_p[s]);
}
static void pack_gru_bwd(const float* P, const GruOff& o, const GruPacked& W, JobQ& q, bool slices = true) {
  const int H = o.H, In = o.In;
  for (int d = 0; d < 2; ++d) {
    if (slices || !W.whhT_rw[d]) q.whh(P + o.whh[d], H, 1, W.whhT_p[d]);
    if (W.whh_rw[d]) q.whh_rw(P + o.whh[d], H, 1, W.whhT_rw[d]);
    q.T(P + o.wih[d], In, In, 3 * H, 3 * H, W.wihT_p[d]);                 // [In rows, K = 3H]
  }
}
static void pack_gru_weights(const float* P, const GruOff& o, const GruPacked& W, JobQ& q, bool slices = true) {
  pack_gru_fwd(P, o, W, q, slices);
  pack_gru_bwd(P, o, W, q, slices);
}

// everything except the forward formats of encoder layer 0 (first = false), or only those (first = true), or all (both)
static void pack_weights_part(const vame_dims& d, const float* P, const PackedWeights& W, bool first, bool rest, cudaStream_t st,
                              int train_batch = 0, bool skip_e0_bwd = false) {
  const ParamLayout L = param_layout(d);
  const int H = d.hidden_enc, F = d.num_features, Z = d.zdims;
  // train_batch > 0: formats of the slice kernels are skipped for the layers whose sweeps run on the rw kernels at that batch
  const int tiles = train_batch > 0 ? pad128(train_batch) / 128 : 0;
  auto slices_of = [&](int Hl) { return !(tiles > 0 && g_opt_rw == 3 && rw_applicable(Hl, tiles)); };
  JobQ q(st);
  if (first) pack_gru_fwd(P, L.e0, W.e0, q, slices_of(H));
  if (rest) {
    if (!skip_e0_bwd) pack_gru_bwd(P, L.e0, W.e0, q, slices_of(H));
    pack_gru_weights(P, L.e1, W.e1, q, slices_of(H));
    for (int i = 0; i < 4; ++i) q.rows(P + L.lam_w + (long)i * H, 4 * H, 2 * Z, H, 2 * Z, W.lam_p[i]);
    q.T(P + L.lam_w, 4 * H, 4 * H, 2 * Z, 2 * Z, W.lamT_p);
    for (int i = 0; i < (d.future_decoder ? 2 : 1); ++i) {
      const int Hd = i == 0 ? d.hidden_rec : d.hidden_pred;
      pack_gru_weights(P, i == 0 ? L.dec : L.fut, i == 0 ? W.dec : W.fut, q, slices_of(Hd));
      q.rows(P + L.l2h_w[i], Z, 2 * Hd, Z, 2 * Hd, W.l2h_p[i]);
      q.T(P + L.l2h_w[i], Z, Z, 2 * Hd, 2 * Hd, W.l2hT_p[i]);
      for (int dd = 0; dd < 2; ++dd) q.rows(P + L.h2o_w[i] + (long)dd * Hd, 2 * Hd, F, Hd, F, W.h2o_p[i][dd]);
      q.T(P + L.h2o_w[i], 2 * Hd, 2 * Hd, F, F, W.h2oT_p[i]);
    }
  }
  q.flush();
}
// split by LAYER (early optimizer step of the train loop): part 0 = every format of encoder layer 0, part 1 = everything else
static void pack_weights_layers(const vame_dims& d, const float* P, const PackedWeights& W, int part, cudaStream_t st, int batch) {
  const ParamLayout L = param_layout(d);
  const int H = d.hidden_enc;
  const int tiles = batch > 0 ? pad128(batch) / 128 : 0;
  if (part == 0) {
    const bool slices = !(tiles > 0 && g_opt_rw == 3 && rw_applicable(H, tiles));
    JobQ q(st);
    pack_gru_fwd(P, L.e0, W.e0, q, slices);
    pack_gru_bwd(P, L.e0, W.e0, q, slices);
    q.flush();
  } else {
    pack_weights_part(d, P, W, false, true, st, batch, /*skip_e0_bwd=*/true);
  }
}
static void pack_all_weights(const vame_dims& d, const float* P, const PackedWeights& W, cudaStream_t st) {
  pack_weights_part(d, P, W, true, true, st);
}

// Deferred re-pack (vame_pack_weights_deferred): the formats the forward pass needs first are packed on the caller's stream,
// the other ~95 % on a low-priority internal stream while the first encoder sweep (latency-bound, 128 of 148 SMs) is running;
// the next vame_forward on that stream joins it right after it has enqueued that sweep.
struct DeferredPack {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, done = nullptr;
  bool pending = false;
  vame_dims dims{};
  const float* params = nullptr;
  void* packed = nullptr;
};
static DeferredPack& deferred_pack() {
  static DeferredPack D;
  if (!D.s) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = least priority (numerically greatest)
    cudaStreamCreateWithPriority(&D.s, cudaStreamNonBlocking, lo);
    cudaEventCreateWithFlags(&D.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&D.done, cudaEventDisableTiming);
  }
  return D;
}
// Call order in encoder_forward: deferred_pack_fork(st) BEFORE the first sweep is enqueued (the rest of the re-pack becomes
// runnable together with the sweep, not earlier - it would only compete with the input-projection GEMM), then the sweep, then
// deferred_pack_run(st): the pack kernels are submitted behind the sweep on the low-priority stream and joined back into st.
static inline void deferred_pack_fork(cudaStream_t st) {
  DeferredPack& D = deferred_pack();
  if (D.pending) cudaEventRecord(D.fork, st);
}
static inline void deferred_pack_run(cudaStream_t st) {
  DeferredPack& D = deferred_pack();
  if (!D.pending) return;
  D.pending = false;
  cudaStreamWaitEvent(D.s, D.fork, 0);
  pack_weights_part(D.dims, D.params, packed_layout(D.dims, D.packed), false, true, D.s);
  cudaEventRecord(D.done, D.s);
  cudaStreamWaitEvent(st, D.done, 0);
}
// any other consumer of the packed weights: finish a pending deferred re-pack on its own stream first
static inline void join_deferred_pack(cudaStream_t st) {
  DeferredPack& D = deferred_pack();
  if (!D.pending) return;
  D.pending = false;
  pack_weights_part(D.dims, D.params, packed_layout(D.dims, D.packed), false, true, st);
}

// ================================================================================================
// GRU sweeps
// ================================================================================================
// The step kernels write only the valid k-range of their P16 outputs; when the range does not fill whole 64-wide chunks
// (H or 3H not a multiple of 64) the padding must be zero because the generic GEMM always consumes full chunks.
static void zero_p16_padding(GruBuf& L, int tiles, bool fwd, cudaStream_t st) {
  const int H = L.H;
  if (fwd && (H % KCHUNK) != 0) {
    const size_t slotp = (size_t)tiles * nkc_of(H) * p16_tile_bytes(128);
    for (int d = 0; d < 2; ++d) cudaMemsetAsync(L.out_p[d], 0, slotp * L.out_p_slots, st);
  }
  if (!fwd && ((3 * H) % KCHUNK) != 0) {
    for (int d = 0; d < 2; ++d)
      if (L.dgi_p[d]) cudaMemsetAsync(L.dgi_p[d], 0, (size_t)L.steps * tiles * nkc_of(3 * H) * p16_tile_bytes(128), st);
  }
}

static int g_fwd_concurrent = 1;     // forward sweeps that run side by side (2 while the decoder and the future decoder overlap)
static void gru_sweep_fwd(const GruPacked& W, const float* b_hn0, const float* b_hn1, GruBuf& L, int tiles, bool save, bool pdl,
                          cudaStream_t st) {
  const int H = L.H;
  zero_p16_padding(L, tiles, true, st);
  // resident-weight cluster kernel (gru_rw.cu): needs 8 consecutive batch rows of gi per vector load
  const bool rw = (g_opt_rw & 1) && W.whh_rw[0] && rw_applicable(H, tiles) && L.gi_bs == 1 && (L.gi_ts % 4) == 0;
  // large inference batches (embedding): row-resident kernel, 128 rows per persistent CTA (gru_rows.cu)
  const bool rows = !rw && !save && W.whh_rows[0] && rows_fwd_applicable(H, tiles);
  if (rw || rows || (g_opt_persistent & 1)) {   // one kernel for the whole sweep
    GruSeqFwdArgs a{};
    a.ndir = 2; a.H = H; a.tiles = tiles; a.steps = L.steps;
    const long Bp = (long)tiles * 128;
    for (int d = 0; d < 2; ++d) {
      GruSeqDirFwd& D = a.d[d];
      D.w_p = W.whh_p[d]; D.b_hn = d == 0 ? b_hn0 : b_hn1;
      D.w_rw = W.whh_rw[d];
      D.w_rows = W.whh_rows[d];
      D.gi = L.gi + (size_t)d * 3 * H * L.gi_ld; D.gi_ld = L.gi_ld; D.gi_bs = L.gi_bs; D.gi_ts = L.gi_ts;
      D.h0 = L.h0[d]; D.h0_ld = Bp; D.h0_p = L.h0_p[d];
      D.out = L.out[d]; D.out_ld = (long)L.out_slots * Bp; D.out_slots = L.out_slots;
      D.out_p = L.out_p[d]; D.out_p_slots = L.out_p_slots;
      D.out_p_slot_elems = (long)tiles * nkc_of(H) * (long)p16_tile_elems(128);
      for (int i = 0; i < 4; ++i) D.sv[i] = save ? L.sv[d][i] : nullptr;
      D.sv_ld = (long)L.steps * Bp;
      D.reverse = d;
      D.priv = (rw && save && L.priv) ? 1 : 0;
      D.outT_p = L.outT_p[d]; D.outT_nk = (long)L.steps * Bp / KCHUNK;
      D.hfin = L.hfin[d];
    }
    if (rw) launch_gru_rw_fwd(a, st);
    else if (rows) launch_gru_rows_fwd(a, st);
    else launch_gru_seq_fwd(a, st);
    return;
  }
  const size_t slotp = (size_t)tiles * nkc_of(H) * p16_tile_elems(128);
  const bool use_flags = pdl && g_opt_pdl && g_opt_flags;
  if (use_flags) cudaMemsetAsync(L.flags, 0, (size_t)2 * L.steps * tiles * sizeof(unsigned int), st);
  for (int s = 0; s < L.steps; ++s) {
    GruFwdArgs a{};
    a.ndir = 2; a.H = H; a.tiles = tiles; a.pdl = (pdl && g_opt_pdl && s > 0) ? 1 : 0;
    a.flags = use_flags ? 1 : 0;
    // 16-unit slices double the CTAs (shorter MMA chain and epilogue per step) as long as two concurrent sweeps
    // (decoder + future decoder) still fit the 148 SMs in one wave
    // CTA shape: use as many SMs as one wave allows: first halve the rows per CTA (M = 64), then the units per CTA (16) -
    // as long as the concurrent sweeps still fit the 148 SMs
    int ctas = g_fwd_concurrent * tiles * (H / 32) * 2;
    a.mt = 128; a.upc = 32;
    if (g_opt_m64 && g_opt_warps16 && 2 * ctas <= 148) { a.mt = 64; ctas *= 2; }
    if (g_opt_slice16 && 2 * ctas <= 148) { a.upc = 16; ctas *= 2; }
    a.flag_expected = (unsigned int)(H / a.upc) * (g_opt_warps16 ? 16u : 8u);   // every warp of every slice CTA signals once
    for (int d = 0; d < 2; ++d) {
      const int t = d == 0 ? s : L.steps - 1 - s;
      const int tprev = d == 0 ? t - 1 : t + 1;
      GruDirFwd& D = a.d[d];
      if (use_flags) {
        unsigned int* f = L.flags + (size_t)d * L.steps * tiles;
        D.flag_in = s == 0 ? nullptr : f + (size_t)(s - 1) * tiles;
        D.flag_out = f + (size_t)s * tiles;
      }
      D.w_p = W.whh_p[d];
      D.b_hn = d == 0 ? b_hn0 : b_hn1;
      D.gi = L.gi + (size_t)d * 3 * H * L.gi_ld;
      D.gi_bs = L.gi_bs; D.gi_ts = L.gi_ts; D.gi_ld = L.gi_ld; D.t = t;
      const int so = (L.out_slots == L.steps) ? t : (s & 1), so_prev = (L.out_slots == L.steps) ? tprev : ((s - 1) & 1);
      const int sp = (L.out_p_slots == L.steps) ? t : (s & 1), sp_prev = (L.out_p_slots == L.steps) ? tprev : ((s - 1) & 1);
      const long Bp = (long)tiles * 128, out_ld = (long)L.out_slots * Bp;
      D.h_in = s == 0 ? L.h0[d] : L.out[d] + so_prev * Bp;
      D.h_in_ld = s == 0 ? Bp : out_ld;
      D.h_in_p = s == 0 ? L.h0_p[d] : (void*)((bf16*)L.out_p[d] + sp_prev * slotp);
      D.h_out = L.out[d] + so * Bp;
      D.h_out_ld = out_ld;
      D.h_out_p = (bf16*)L.out_p[d] + sp * slotp;
      if (save) {
        D.sv_ld = (long)L.steps * Bp;
        D.sv_r = L.sv[d][0] + t * Bp; D.sv_z = L.sv[d][1] + t * Bp;
        D.sv_n = L.sv[d][2] + t * Bp; D.sv_ghn = L.sv[d][3] + t * Bp;
      }
    }
    launch_gru_step_fwd(a, st);
  }
}
// location of the final hidden state of direction d after a forward sweep
static inline const float* final_h(const GruBuf& L, int d, int tiles) {      // feature-major, ld = final_h_ld()
  if (L.priv) return L.hfin[d];
  const int t = d == 0 ? L.steps - 1 : 0;
  const int so = (L.out_slots == L.steps) ? t : ((L.steps - 1) & 1);
  return L.out[d] + (size_t)so * tiles * 128;
}
static inline long final_h_ld(const GruBuf& L, int tiles) { return L.priv ? (long)tiles * 128 : (long)L.out_slots * tiles * 128; }
static inline const void* final_h_p(const GruBuf& L, int d, int tiles) {
  const size_t slotp = (size_t)tiles * nkc_of(L.H) * p16_tile_elems(128);
  const int t = d == 0 ? L.steps - 1 : 0;
  const int sp = (L.out_p_slots == L.steps) ? t : ((L.steps - 1) & 1);
  return (bf16*)L.out_p[d] + sp * slotp;
}

// dout0/1: feature-major [H][dout_ld] upstream gradients of the per-step outputs (slot t at + t*B_pad) or nullptr;
// dhl0/1: feature-major [H][dhl_ld] gradient of the final hidden state or nullptr
static int g_bwd_concurrent = 1;     // backward sweeps that run side by side (2 with a future decoder)
// dgi_sum / dgi_sum_p (optional, per direction): time sums of dgi written by the persistent kernel; returns true if they were produced
static bool gru_sweep_bwd(const GruPacked& W, GruBuf& L, int tiles, const float* dout0, const float* dout1, long dout_ld,
                          const float* dhl0, const float* dhl1, long dhl_ld, bool pdl, cudaStream_t st,
                          float* const* dgi_sum = nullptr, void* const* dgi_sum_p = nullptr, const GruOff* goff = nullptr,
                          float* G = nullptr) {
  const int H = L.H, nsl = H / 32, nkc3 = nkc_of(3 * H);
  const long Bp = (long)tiles * 128;
  const size_t slotf = (size_t)Bp * H;
  const size_t pslot = (size_t)(nsl + 1) * slotf;
  zero_p16_padding(L, tiles, false, st);
  const bool rw = (g_opt_rw & 2) && W.whhT_rw[0] && rw_applicable(H, tiles);
  L.parts_n = rw ? 1 : nsl + 1;
  if (rw || (g_opt_persistent & 2)) {           // one cluster kernel for the whole BPTT sweep
    GruSeqBwdArgs a{};
    a.ndir = 2; a.H = H; a.tiles = tiles; a.steps = L.steps;
    a.mt = (g_opt_m64 && g_bwd_concurrent * tiles * (H / 32) * 2 * 2 <= 148) ? 64 : 128;
    const long seq_ld = (long)L.steps * Bp;
    for (int d = 0; d < 2; ++d) {
      GruSeqDirBwd& D = a.d[d];
      D.wT_p = W.whhT_p[d];
      D.wT_rw = W.whhT_rw[d];
      D.dh0_out = L.parts[d] + ((L.steps - 1) & 1) * pslot;          // = final_parts(L, d, tiles), slot 0
      D.dh_last = d == 0 ? dhl0 : dhl1; D.dh_last_ld = dhl_ld;
      D.dout = d == 0 ? dout0 : dout1; D.dout_ld = dout_ld;
      for (int i = 0; i < 4; ++i) D.sv[i] = L.sv[d][i];
      D.sv_ld = seq_ld;
      D.out = L.out[d]; D.out_ld = seq_ld;
      D.h0 = L.h0[d]; D.h0_ld = Bp;
      D.parts = L.parts[d];
      D.dgi = L.dgi[d]; D.dgh = L.dgh[d]; D.dg_ld = seq_ld;
      D.dgi_p = L.dgi_p[d]; D.dgi_p_slot_elems = (long)tiles * nkc3 * (long)p16_tile_elems(128);
      D.priv = (rw && L.priv && goff && G) ? 1 : 0;
      D.dout_pv = (D.priv && L.dout_pv) ? 1 : 0;
      D.dghT_p = L.dghT_p[d]; D.dgiT_p = L.dgiT_p[d]; D.gT_nk = seq_ld / KCHUNK;
      D.db_ih = G ? G + goff->bih[d] : nullptr; D.db_hh = G ? G + goff->bhh[d] : nullptr;
      D.dgi_sum = dgi_sum ? dgi_sum[d] : nullptr;
      D.dgi_sum_p = dgi_sum ? dgi_sum_p[d] : nullptr;
      if (dgi_sum && ((3 * H) % KCHUNK) != 0) cudaMemsetAsync(dgi_sum_p[d], 0, (size_t)tiles * nkc3 * p16_tile_bytes(128), st);
      D.reverse = d;
    }
    if (rw) launch_gru_rw_bwd(a, st);
    else launch_gru_seq_bwd(a, st);
    return dgi_sum != nullptr;
  }
  const bool use_flags = pdl && g_opt_pdl && g_opt_flags;
  if (use_flags) cudaMemsetAsync(L.flags, 0, (size_t)2 * L.steps * tiles * sizeof(unsigned int), st);
  for (int s = 0; s < L.steps; ++s) {
    GruBwdArgs a{};
    a.ndir = 2; a.H = H; a.tiles = tiles; a.pdl = (pdl && g_opt_pdl && s > 0) ? 1 : 0;
    a.flags = use_flags ? 1 : 0;
    a.flag_expected = (unsigned int)(H / 32) * 8u;
    for (int d = 0; d < 2; ++d) {
      const int t = d == 0 ? L.steps - 1 - s : s;           // reverse of the forward order
      const bool first_fwd = d == 0 ? (t == 0) : (t == L.steps - 1);
      const int tprev = d == 0 ? t - 1 : t + 1;
      GruDirBwd& D = a.d[d];
      if (use_flags) {
        unsigned int* f = L.flags + (size_t)d * L.steps * tiles;
        D.flag_in = s == 0 ? nullptr : f + (size_t)(s - 1) * tiles;
        D.flag_out = f + (size_t)s * tiles;
      }
      D.wT_p = W.whhT_p[d];
      const float* dhl = d == 0 ? dhl0 : dhl1;
      if (s == 0) {
        D.parts = dhl; D.n_parts = dhl ? 1 : 0; D.parts_stride = 0; D.parts_ld = dhl_ld;
      } else {
        D.parts = L.parts[d] + ((s - 1) & 1) * pslot; D.n_parts = nsl + 1; D.parts_stride = (long)slotf; D.parts_ld = Bp;
      }
      const float* dout = d == 0 ? dout0 : dout1;
      D.dout = dout ? dout + (long)t * Bp : nullptr;
      D.dout_ld = dout_ld;
      const long seq_ld = (long)L.steps * Bp;
      D.sv_ld = seq_ld;
      D.sv_r = L.sv[d][0] + t * Bp; D.sv_z = L.sv[d][1] + t * Bp;
      D.sv_n = L.sv[d][2] + t * Bp; D.sv_ghn = L.sv[d][3] + t * Bp;
      D.h_prev = first_fwd ? L.h0[d] : L.out[d] + tprev * Bp;
      D.h_prev_ld = first_fwd ? Bp : seq_ld;
      D.parts_out = L.parts[d] + (s & 1) * pslot;
      D.dg_ld = seq_ld;
      D.dgi = L.dgi[d] + (size_t)t * Bp;
      D.dgh = L.dgh[d] + (size_t)t * Bp;
      D.dgi_p = L.dgi_p[d] ? (void*)((bf16*)L.dgi_p[d] + (size_t)t * tiles * nkc3 * p16_tile_elems(128)) : nullptr;
    }
    launch_gru_step_bwd(a, st);
  }
  return false;
}
static inline const float* final_parts(const GruBuf& L, int d, int tiles) {
  const size_t pslot = (size_t)(L.H / 32 + 1) * tiles * 128 * L.H;
  return L.parts[d] + ((L.steps - 1) & 1) * pslot;
}

// weight / bias gradients of the recurrent part of one bi-GRU layer after its backward sweep:
//   dW_hh[d] = dgh[d]^T hprev[d],  db_hh[d] = colsum(dgh[d]),  db_ih[d] = colsum(dgi[d])
static void pack_outT(GruBuf& L, int Bp, cudaStream_t st) {      // forward activations only: can run any time after the forward
  const long rows = (long)L.steps * Bp;
  if (L.priv) return;                                                 // written by the forward sweep itself
  for (int d = 0; d < 2; ++d) pack_rows(L.out[d], rows, L.H, (int)rows, L.H, L.outT_p[d], st);   // out is already [H][rows]
}
// dgi_rowsum_fused: the caller packs dgi (encoder layers) and lets that pass produce db_ih
static void gru_recurrent_grads(const GruOff& o, GruBuf& L, int Bp, const void* h0T0, const void* h0T1, float* G, cudaStream_t st0,
                                bool dgi_rowsum_fused = false, cudaStream_t st1 = nullptr) {
  const int H = L.H;
  const long rows = (long)L.steps * Bp;
  const int cB = Bp / KCHUNK, nk = (int)(rows / KCHUNK);
  for (int d = 0; d < 2; ++d) {
    cudaStream_t st = (d == 1 && st1) ? st1 : st0;           // the two directions are independent
    // dgh is already [3H][rows]; its row sums (= db_hh) are accumulated by the same pass
    if (!L.priv) launch_pack_p16_rowsum(L.dgh[d], rows, 3 * H, (int)rows, 3 * H, L.dghT_p[d], G + o.bhh[d], st);   // priv: by the sweep
    const void* h0T = d == 0 ? h0T0 : h0T1;
    GemmB gb;
    gb.A(L.dghT_p[d], nk, nk);
    const bf16* oT = (const bf16*)L.outT_p[d];
    if (d == 0) {                                   // hprev(t) = out(t-1), hprev(0) = h0
      gb.Bm(h0T, cB, cB);
      if (nk > cB) gb.Bm(oT, nk, nk - cB);
    } else {                                        // hprev(t) = out(t+1), hprev(T-1) = h0
      if (nk > cB) gb.Bm(oT + (size_t)cB * p16_tile_elems(128), nk, nk - cB);
      gb.Bm(h0T, cB, cB);
    }
    gb.run(3 * H, H, G + o.whh[d], H, nullptr, 1, splits_for(3 * H, H, nk), st);
    if (!dgi_rowsum_fused && !L.priv) launch_rowsum_fm(L.dgi[d], rows, rows, 3 * H, G + o.bih[d], st);
  }
}

// ================================================================================================
// forward
// ================================================================================================
static void encoder_forward(const vame_dims& d, const float* P, const ParamLayout& L, const PackedWeights& W, Ws& w, const float* x,
                            long x_bs, long x_ts, bool save, cudaStream_t st) {
  const int T = d.time_window, F = d.num_features, H = d.hidden_enc, Bp = w.B_pad;
  const int rows = T * Bp, nkcH = nkc_of(H);
  cudaMemsetAsync(w.zeros_f32, 0, (size_t)Bp * w.Hmax * 4, st);     // feature-major [Hmax][B_pad] zeros
  cudaMemsetAsync(w.zeros_p, 0, w.zeros_p_bytes, st);
  for (int dd = 0; dd < 2; ++dd) {
    w.e0.h0[dd] = w.zeros_f32; w.e0.h0_p[dd] = w.zeros_p;
    w.e1.h0[dd] = w.zeros_f32; w.e1.h0_p[dd] = w.zeros_p;
  }
  launch_bt_to_tb_p16(x, w.B, T, F, x_bs, x_ts, Bp, w.x_tb, w.x_p, st);     // time-major copy + its P16 operand in one pass
  // layer 0: gi = x W_ih^T + (b_ih + [b_hr, b_hz, 0]) for both directions at once
  GemmB().A(w.x_p, nkc_of(F), nkc_of(F)).Bm(W.e0.wih_p[0], nkc_of(F), nkc_of(F))
      .run_fm(rows, 6 * H, w.e0.gi, rows, W.e0.bias_gi, st);
  mark(st, "enc:x pack + gi0 gemm");
  deferred_pack_fork(st);
  gru_sweep_fwd(W.e0, P + L.e0.bhh[0] + 2 * H, P + L.e0.bhh[1] + 2 * H, w.e0, w.tiles, save, true, st);
  deferred_pack_run(st);                     // the rest of a deferred weight re-pack runs beside the sweep
  mark(st, "enc:L0 sweep");
  // layer 1: input = [out_f(t), out_b(t)] (rnn_model.py:41, inter-layer dropout is 0 by default)
  GemmB().A(w.e0.out_p[0], nkcH, nkcH).A(w.e0.out_p[1], nkcH, nkcH)
      .Bm(W.e1.wih_p[0], nkcH, nkcH).Bm(W.e1.wih_p[1], nkcH, nkcH)
      .run_fm(rows, 6 * H, w.e1.gi, rows, W.e1.bias_gi, st);
  mark(st, "enc:gi1 gemm");
  gru_sweep_fwd(W.e1, P + L.e1.bhh[0] + 2 * H, P + L.e1.bhh[1] + 2 * H, w.e1, w.tiles, save, true, st);
}

// mu/logvar linear on hidden = cat(h_n[0..3]) (rnn_model.py:43,65-69); hidden pieces given as P16 [tiles][nkc]
static void lambda_linear(const vame_dims& d, const float* P, const ParamLayout& L, const PackedWeights& W, Ws& w, const void* hp[4],
                          cudaStream_t st) {
  const int H = d.hidden_enc, Z = d.zdims, nkcH = nkc_of(H);
  GemmB gb;
  for (int i = 0; i < 4; ++i) gb.A(hp[i], nkcH, nkcH);
  for (int i = 0; i < 4; ++i) gb.Bm(W.lam_p[i], nkcH, nkcH);
  // 2 output tiles x K = 4H: split-K over 8 CTAs per tile (red.add into the zeroed output) instead of a 16-chunk serial loop
  cudaMemsetAsync(w.lin, 0, (size_t)w.B * 2 * Z * sizeof(float), st);
  const int keep = g_side_sms;
  g_side_sms = 148;
  gb.run(w.B, 2 * Z, w.lin, 2 * Z, P + L.lam_b, 1, 8, st);
  g_side_sms = keep;
}

static void decoder_forward(const vame_dims& d, int which, const float* P, const ParamLayout& L, const PackedWeights& W, Ws& w,
                            bool save, cudaStream_t st) {
  const int F = d.num_features, Z = d.zdims, Bp = w.B_pad;
  const int Hd = which == 0 ? d.hidden_rec : d.hidden_pred;
  const GruOff& o = which == 0 ? L.dec : L.fut;
  const GruPacked& Wg = which == 0 ? W.dec : W.fut;
  DecBuf& D = w.dec[which];
  const int steps = D.g.steps, nkcZ = nkc_of(Z), nkcH = nkc_of(Hd);
  // hidden = latent_to_hidden(z); h0 = hidden.view(2, B, H)  (raw reinterpretation, rnn_model.py:102-104)
  // the decoder input is z at every time step (rnn_model.py:169-170): one projection per sample - independent of the h0
  // chain, so it runs beside it on a side stream that is idle during the forward pass
  cudaStream_t sg = g_opt_streams ? side().s[which == 0 ? 1 : 3] : st;
  edge(st, sg);
  GemmB().A(w.z_p, nkcZ, nkcZ).Bm(Wg.wih_p[0], nkcZ, nkcZ).run_fm(Bp, 6 * Hd, D.g.gi, Bp, Wg.bias_gi, sg);
  GemmB().A(w.z_p, nkcZ, nkcZ).Bm(W.l2h_p[which], nkcZ, nkcZ).run(w.B, 2 * Hd, D.hid, 2 * Hd, P + L.l2h_b[which], 0, 1, st);
  launch_h0_prepare(D.hid, 2, w.B, Bp, Hd, D.g.h0[0], D.g.h0_p[0], st);
  edge(sg, st);
  g_fwd_concurrent = (d.future_decoder && g_opt_streams) ? 2 : 1;
  gru_sweep_fwd(Wg, P + o.bhh[0] + 2 * Hd, P + o.bhh[1] + 2 * Hd, D.g, w.tiles, save, true, st);
  g_fwd_concurrent = 1;
  // prediction = hidden_to_output([out_f, out_b])
  GemmB().A(D.g.out_p[0], nkcH, nkcH).A(D.g.out_p[1], nkcH, nkcH).Bm(W.h2o_p[which][0], nkcH, nkcH).Bm(W.h2o_p[which][1], nkcH, nkcH)
      .run(steps * Bp, F, D.pred_tb, F, P + L.h2o_b[which], 0, 1, st);
}

static int check_dims(const vame_dims* d) {
  VB_REQUIRE(d, "null dims");
  VB_REQUIRE(d->num_features > 0 && d->time_window > 0 && d->zdims > 0 && d->zdims <= 64, "dims: need F>0, T>0, 0<Z<=64");
  VB_REQUIRE(d->hidden_enc % 32 == 0 && d->hidden_enc >= 32 && d->hidden_enc <= 256, "dims: hidden_enc must be a multiple of 32 in [32,256]");
  VB_REQUIRE(d->hidden_rec % 32 == 0 && d->hidden_rec >= 32 && d->hidden_rec <= 256, "dims: hidden_rec must be a multiple of 32 in [32,256]");
  if (d->future_decoder) {
    VB_REQUIRE(d->hidden_pred % 32 == 0 && d->hidden_pred >= 32 && d->hidden_pred <= 256, "dims: hidden_pred must be a multiple of 32 in [32,256]");
    VB_REQUIRE(d->future_steps > 0 && d->future_steps <= d->time_window, "dims: need 0 < future_steps <= time_window");
  }
  return 0;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vame_param_tensors(const vame_dims* d) { return d && d->future_decoder ? 44 : 32; }

long vame_param_layout(const vame_dims* d, long* offsets, long* sizes) {
  if (check_dims(d)) return -1;
  const ParamLayout L = param_layout(*d);
  if (offsets && sizes) {
    int n = 0;
    auto gru = [&](const GruOff& g) {
      for (int dir = 0; dir < 2; ++dir) {      // state_dict order: w_ih, w_hh, b_ih, b_hh, then the *_reverse four
        offsets[n] = g.wih[dir]; sizes[n++] = 3L * g.H * g.In;
        offsets[n] = g.whh[dir]; sizes[n++] = 3L * g.H * g.H;
        offsets[n] = g.bih[dir]; sizes[n++] = 3L * g.H;
        offsets[n] = g.bhh[dir]; sizes[n++] = 3L * g.H;
      }
    };
    const int H = d->hidden_enc, Z = d->zdims, F = d->num_features;
    // encoder.encoder_rnn: l0, l0_reverse, l1, l1_reverse
    gru(L.e0);
    gru(L.e1);
    offsets[n] = L.lam_w; sizes[n++] = (long)Z * 4 * H;                       // hidden_to_mean.weight
    offsets[n] = L.lam_b; sizes[n++] = Z;                                     // hidden_to_mean.bias
    offsets[n] = L.lam_w + (long)Z * 4 * H; sizes[n++] = (long)Z * 4 * H;     // hidden_to_logvar.weight
    offsets[n] = L.lam_b + Z; sizes[n++] = Z;                                 // hidden_to_logvar.bias
    for (int i = 0; i < (d->future_decoder ? 2 : 1); ++i) {
      const int Hd = i == 0 ? d->hidden_rec : d->hidden_pred;
      gru(i == 0 ? L.dec : L.fut);
      offsets[n] = L.l2h_w[i]; sizes[n++] = 2L * Hd * Z;
      offsets[n] = L.l2h_b[i]; sizes[n++] = 2L * Hd;
      offsets[n] = L.h2o_w[i]; sizes[n++] = (long)F * 2 * Hd;
      offsets[n] = L.h2o_b[i]; sizes[n++] = F;
    }
  }
  return L.total;
}

size_t vame_packed_weights_bytes(const vame_dims* d) {
  if (check_dims(d)) return 0;
  return packed_layout(*d, nullptr).bytes;
}

int vame_pack_weights(const vame_dims* d, const float* params, void* packed, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(params && packed, "vame_pack_weights: null pointer");
  pack_all_weights(*d, params, packed_layout(*d, packed), (cudaStream_t)stream);
  return check_launch("vame_pack_weights");
}

int vame_pack_weights_train(const vame_dims* d, const float* params, void* packed, int batch, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(params && packed && batch > 0, "vame_pack_weights_train: null pointer / empty batch");
  pack_weights_part(*d, params, packed_layout(*d, packed), true, true, (cudaStream_t)stream, batch);
  return check_launch("vame_pack_weights_train");
}

int vame_pack_weights_train_part(const vame_dims* d, const float* params, void* packed, int batch, int part, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(params && packed && batch > 0 && (part == 0 || part == 1), "vame_pack_weights_train_part: bad arguments");
  pack_weights_layers(*d, params, packed_layout(*d, packed), part, (cudaStream_t)stream, batch);
  return check_launch("vame_pack_weights_train_part");
}

int vame_pack_weights_deferred(const vame_dims* d, const float* params, void* packed, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(params && packed, "vame_pack_weights_deferred: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DeferredPack& D = deferred_pack();
  pack_weights_part(*d, params, packed_layout(*d, packed), true, false, st);
  D.dims = *d; D.params = params; D.packed = packed;
  D.pending = true;
  return check_launch("vame_pack_weights_deferred");
}

size_t vame_workspace_bytes(const vame_dims* d, int batch, int training) {
  if (check_dims(d) || batch <= 0) return 0;
  return carve_ws(*d, batch, training != 0, nullptr).bytes;
}

// Early start of the k-means prior (vame_arm_prior): the prior only needs z, which exists ~100 us into the forward pass, but
// its single-CTA Jacobi solve takes ~300 us - started from vame_loss it kept the Lambda backward waiting for ~130 us of a
// 1.08 ms step.  An armed vame_forward(save_for_backward = 1) launches it on the prior's side stream right after the
// reparameterisation; the following vame_loss sees it in flight and does not launch it again.
struct ArmedPrior {
  bool armed = false, inflight = false;
  vame_loss_cfg cfg{};
  const float* hyper = nullptr;
};
static ArmedPrior& armed_prior() {
  static ArmedPrior A;
  return A;
}
int vame_arm_prior(const vame_loss_cfg* cfg, const float* hyper) {
  ArmedPrior& A = armed_prior();
  A.armed = cfg != nullptr;
  if (cfg) A.cfg = *cfg;
  A.hyper = hyper;
  return 0;
}

int vame_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* x, long x_bs, long x_ts,
                 const float* eps, int save_for_backward, float* pred, float* future, float* z, float* mu, float* logvar, void* ws,
                 size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && x && ws, "vame_forward: null pointer / empty batch");
  const bool save = save_for_backward != 0;
  Ws w = carve_ws(*d, batch, save, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_forward: workspace too small (see vame_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int T = d->time_window, F = d->num_features, Z = d->zdims;
  cudaMemsetAsync(w.acc, 0, 8 * sizeof(double), st);
  mark(st, "fwd:start");
  encoder_forward(*d, params, L, W, w, x, x_bs, x_ts, save, st);
  mark(st, "fwd:encoder done");
  const void* hp[4] = {final_h_p(w.e0, 0, w.tiles), final_h_p(w.e0, 1, w.tiles), final_h_p(w.e1, 0, w.tiles), final_h_p(w.e1, 1, w.tiles)};
  lambda_linear(*d, params, L, W, w, hp, st);
  if (eps) cudaMemcpyAsync(w.eps, eps, (size_t)batch * Z * 4, cudaMemcpyDeviceToDevice, st);
  launch_lambda_fwd_p16(w.lin, 2 * Z, eps ? w.eps : nullptr, batch, w.B_pad, Z, d->softplus, w.z, w.mu, w.logvar, w.acc, w.z_p, st);
  {
    ArmedPrior& A = armed_prior();
    A.inflight = false;
    if (A.armed && save && g_opt_streams) {
      cudaStream_t sp = side().s[2];
      edge(st, sp);
      launch_cluster_prior(w.z, batch, Z, A.cfg.kmeans_loss, A.cfg.kmeans_lambda, A.cfg.bsize > 0 ? A.cfg.bsize : (float)batch,
                           A.cfg.kl_weight, A.hyper, w.dz_km, w.acc, sp, w.prior_state);
      A.inflight = true;
    }
  }
  mark(st, "fwd:lambda done");
  cudaStream_t sA = (d->future_decoder && g_opt_streams) ? side().s[0] : st;
  if (d->future_decoder) {
    edge(st, sA);
    decoder_forward(*d, 1, params, L, W, w, save, sA);
  }
  decoder_forward(*d, 0, params, L, W, w, save, st);
  if (d->future_decoder) edge(sA, st);
  mark(st, "fwd:decoders done");
  if (pred) launch_tb_to_bt(w.dec[0].pred_tb, batch, T, F, w.B_pad, pred, st);
  if (future && d->future_decoder) launch_tb_to_bt(w.dec[1].pred_tb, batch, d->future_steps, F, w.B_pad, future, st);
  if (z) cudaMemcpyAsync(z, w.z, (size_t)batch * Z * 4, cudaMemcpyDeviceToDevice, st);
  if (mu) cudaMemcpyAsync(mu, w.mu, (size_t)batch * Z * 4, cudaMemcpyDeviceToDevice, st);
  if (logvar) cudaMemcpyAsync(logvar, w.logvar, (size_t)batch * Z * 4, cudaMemcpyDeviceToDevice, st);
  // remember whether eps was given (backward needs it): all-ones / all-zero bytes in acc[7]
  cudaMemsetAsync(w.acc + 7, eps ? 0xFF : 0, sizeof(double), st);
  return check_launch("vame_forward");
}

int vame_loss(const vame_dims* d, int batch, const vame_loss_cfg* cfg, const float* fut, long f_bs, long f_ts, const float* target,
              long t_bs, long t_ts, const float* hyper, float* losses_out, int want_grads, void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(cfg && ws && losses_out && batch > 0, "vame_loss: null pointer");
  const bool training = want_grads != 0;
  Ws w = carve_ws(*d, batch, training, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int T = d->time_window, F = d->num_features, Z = d->zdims, Bp = w.B_pad;
  const bool with_fut = d->future_decoder && cfg->with_future;
  VB_REQUIRE(!with_fut || fut, "vame_loss: future target missing");
  const double nrec = (double)batch * T * F;
  cudaStream_t sA = g_opt_streams ? side().s[2] : st;   // the prior's own side stream
  edge(st, sA);
  ArmedPrior& AP = armed_prior();
  if (AP.inflight && training && g_opt_streams) AP.inflight = false;       // already running since the forward pass (vame_arm_prior)
  else
    launch_cluster_prior(w.z, batch, Z, cfg->kmeans_loss, cfg->kmeans_lambda, cfg->bsize, cfg->kl_weight, hyper,
                         training ? w.dz_km : nullptr, w.acc, sA, training ? w.prior_state : nullptr);
  const float* rec_target = w.x_tb;
  if (target) {                              // clean target of a noisy forward input
    launch_bt_to_tb(target, batch, T, F, t_bs, t_ts, Bp, w.dec[0].target_tb, st);
    rec_target = w.dec[0].target_tb;
  }
  // training: the gradient is also written as the P16 operand of the hidden_to_output backward GEMM (vame_backward)
  if (training)
    launch_mse_p16(w.dec[0].pred_tb, F, rec_target, T * Bp, batch, Bp, F, cfg->mse_red_mean ? (float)(2.0 / nrec) : 2.0f,
                   w.dec[0].dpred_tb, w.acc, ACC_REC, w.dec[0].dpred_p, st);
  else
    launch_mse(w.dec[0].pred_tb, F, rec_target, T * Bp, batch, Bp, F, cfg->mse_red_mean ? (float)(2.0 / nrec) : 2.0f, nullptr, w.acc,
               ACC_REC, st);
  double nfut = 1.0;
  if (with_fut) {
    const int S = d->future_steps;
    nfut = (double)batch * S * F;
    launch_bt_to_tb(fut, batch, S, F, f_bs, f_ts, Bp, w.dec[1].target_tb, st);
    if (training)
      launch_mse_p16(w.dec[1].pred_tb, F, w.dec[1].target_tb, S * Bp, batch, Bp, F, cfg->mse_pred_mean ? (float)(2.0 / nfut) : 2.0f,
                     w.dec[1].dpred_tb, w.acc, ACC_FUT, w.dec[1].dpred_p, st);
    else
      launch_mse(w.dec[1].pred_tb, F, w.dec[1].target_tb, S * Bp, batch, Bp, F, cfg->mse_pred_mean ? (float)(2.0 / nfut) : 2.0f, nullptr,
                 w.acc, ACC_FUT, st);
  }
  if (training && cfg->defer_prior_join && g_opt_streams) {
    // the prior keeps running on sA while the caller's stream proceeds into vame_backward (which joins sA before the
    // Lambda backward); the loss vector is finalised on sA once the MSE sums of the main stream are in
    edge(st, sA);
    launch_finalize_losses(w.acc, losses_out, cfg->mse_red_mean ? nrec : 1.0, cfg->mse_pred_mean ? nfut : 1.0, (double)batch * Z,
                           cfg->beta, cfg->kl_weight, hyper, with_fut ? 1 : 0, sA);
  } else {
    edge(sA, st);
    launch_finalize_losses(w.acc, losses_out, cfg->mse_red_mean ? nrec : 1.0, cfg->mse_pred_mean ? nfut : 1.0, (double)batch * Z,
                           cfg->beta, cfg->kl_weight, hyper, with_fut ? 1 : 0, st);
  }
  return check_launch("vame_loss");
}

// ---- gradient-bucket hand-over for the data-parallel step -------------------------------------------------------------------
// Every gradient except encoder layer 0's is final well before vame_backward ends (the layer-0 BPTT sweep and its weight-gradient
// GEMMs are the last ~200 us).  With the overlap enabled, vame_backward records an EXTERNAL event (an event-record node when the
// call is captured into a CUDA graph) at that point; the host makes its communication stream wait for it and all-reduces the
// first bucket while the last sweep runs.
struct GradOverlap {
  bool on = false;
  bool internal = false;     // mode 2: the consumer is captured into the SAME graph (plain event = dependency edge)
  cudaEvent_t ev = nullptr;
};
static GradOverlap& grad_overlap() {
  static GradOverlap G;
  return G;
}
int vame_grad_overlap(int enable) {
  GradOverlap& G = grad_overlap();
  if (enable && !G.ev) {
    if (cudaEventCreateWithFlags(&G.ev, cudaEventDisableTiming) != cudaSuccess) return fail("vame_grad_overlap: cudaEventCreate failed");
  }
  G.on = enable != 0;
  G.internal = enable == 2;
  return 0;
}
long vame_grad_bucket_split(const vame_dims* d) {
  if (!d || check_dims(d)) return -1;
  return param_layout(*d).e1.wih[0];     // [0, split): encoder layer 0 (final last); [split, total): everything else
}
int vame_wait_grads_ready(void* stream) {
  GradOverlap& G = grad_overlap();
  VB_REQUIRE(G.ev, "vame_wait_grads_ready: no vame_backward has run with vame_grad_overlap(1) yet");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing((cudaStream_t)stream, &cap);
  if (cudaStreamWaitEvent((cudaStream_t)stream, G.ev,
                          (cap == cudaStreamCaptureStatusActive && !G.internal) ? cudaEventWaitExternal : cudaEventWaitDefault) != cudaSuccess)
    return fail("vame_wait_grads_ready: cudaStreamWaitEvent failed");
  return 0;
}

int vame_backward(const vame_dims* d, int batch, const float* params, const void* packed, int use_loss_grads, const vame_loss_cfg* cfg,
                  const float* hyper, const float* dpred, const float* dfuture, const float* dz_ext, const float* dmu_ext,
                  const float* dlv_ext, float* grads, void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && grads && ws, "vame_backward: null pointer");
  VB_REQUIRE(!use_loss_grads || cfg, "vame_backward: cfg required with use_loss_grads");
  Ws w = carve_ws(*d, batch, true, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_backward: workspace too small (forward must have used save_for_backward)");
  cudaStream_t st = (cudaStream_t)stream;
  join_deferred_pack(st);
  // sA: the future decoder's backward; sB: weight-gradient work that is not on the data-gradient chain
  cudaStream_t sA = g_opt_streams ? side().s[0] : st;
  cudaStream_t sB = g_opt_streams ? side().s[1] : st;
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int T = d->time_window, F = d->num_features, Z = d->zdims, H = d->hidden_enc, Bp = w.B_pad, B = batch;
  {   // Persistent-grid cap of the weight-gradient GEMMs that run on the side streams beside the sweeps.  Round 1 capped them to
      // the SMs a sweep leaves free (20-48 CTAs); measured in round 2 (option "side_sms", profiles/r2_bench_v16_*): NO cap is
      // better - the CTAs that do not fit wait in the hardware queue and take an SM the moment a sweep CTA exits, instead of a
      // few resident CTAs working through the whole backlog (C2 296.1 -> 310.3 k, C5 363.0 -> 410.1 k windows/s; 20 CTAs: 219.5 k).
    g_side_sms = 148;
    if (g_opt_side_sms > 0) g_side_sms = g_opt_side_sms;      // (experiment knob: option "side_sms")
  }
  const int nkcB = Bp / KCHUNK;
  float* G = grads;
  cudaMemsetAsync(G, 0, (size_t)L.total * 4, st);
  for (int dd = 0; dd < 2; ++dd) {           // encoder buffers alias the zero regions (not stored in the workspace)
    w.e0.h0[dd] = w.zeros_f32; w.e0.h0_p[dd] = w.zeros_p;
    w.e1.h0[dd] = w.zeros_f32; w.e1.h0_p[dd] = w.zeros_p;
  }
  mark(st, "bwd:start");
  pack_T(w.z, Z, Z, Bp, B, w.zT_p, st);                                   // [Z rows, K = B_pad]
  edge(st, sB);
  // operands that only depend on the forward pass: transposed activations for the weight-gradient GEMMs
  pack_outT(w.e0, Bp, sB);
  pack_outT(w.e1, Bp, sB);
  pack_T(w.x_tb, F, F, T * Bp, T * Bp, w.xT_p, sB);
  mark(sB, "side:early packs done");
  if (g_opt_streams) edge(sB, side().s[3]);   // the second weight-gradient stream also reads these operands

  const int ndec = d->future_decoder ? 2 : 1;
  const float* dz_dec[2] = {nullptr, nullptr};
  bool used_sA = false;
  for (int i = ndec - 1; i >= 0; --i) {      // the future decoder (i = 1) is enqueued first, on its own stream
    DecBuf& D = w.dec[i];
    const float* ext = i == 0 ? dpred : dfuture;
    const bool have_grad = use_loss_grads ? (i == 0 || cfg->with_future) : (ext != nullptr);
    if (!have_grad) continue;                                             // this decoder received no gradient
    cudaStream_t sd = i == 0 ? st : sA;      // data-gradient chain of this decoder
    cudaStream_t sw = i == 0 ? sB : sA;      // its weight-gradient work
    if (i == 1) {
      edge(st, sA);
      used_sA = true;
    }
    const int Hd = D.g.H, steps = D.g.steps;
    const long rows = (long)steps * Bp;
    const int nk = (int)(rows / KCHUNK), nkcF = nkc_of(F), nkc3 = nkc_of(3 * Hd), nkc2H = nkc_of(2 * Hd);
    const GruOff& o = i == 0 ? L.dec : L.fut;
    const GruPacked& Wg = i == 0 ? W.dec : W.fut;
    if (!use_loss_grads) launch_bt_to_tb(ext, B, steps, F, (long)steps * F, F, Bp, D.dpred_tb, sd);
    // ---- data-gradient chain: hidden_to_output backward, BPTT, dz
    cudaMemsetAsync(D.dz, 0, (size_t)B * Z * 4, sd);                      // accumulated by the split-K dz GEMM after the sweep
    if (!use_loss_grads) pack_rows(D.dpred_tb, F, (int)rows, F, (int)rows, D.dpred_p, sd);   // (vame_loss wrote it already)
    D.g.dout_pv = D.g.priv && Hd == 256 && rw_priv_mode(Hd, w.tiles);
    if (D.g.dout_pv) GemmB().A(D.dpred_p, nkcF, nkcF).Bm(W.h2oT_p[i], nkcF, nkcF).run_pv((int)rows, 2 * Hd, D.ddec, rows, Bp, sd);
    else GemmB().A(D.dpred_p, nkcF, nkcF).Bm(W.h2oT_p[i], nkcF, nkcF).run_fm((int)rows, 2 * Hd, D.ddec, rows, nullptr, sd);
    if (i == 0) mark(st, "bwd:dec dpred pack + ddec gemm");
    {   // weight-gradient operands that only need the forward pass and dpred: packed on sw while the sweep runs
      edge(sd, sw);
      pack_T(D.dpred_tb, F, F, (int)rows, (int)rows, D.dpredT_p, sw);
      launch_colsum(D.dpred_tb, F, rows, F, G + L.h2o_b[i], sw);
      for (int dd = 0; dd < 2; ++dd) pack_rows(D.g.h0[dd], Bp, Hd, Bp, Hd, D.g.h0T_p[dd], sw);       // h0 is [H][B_pad]
      pack_outT(D.g, Bp, sw);
      for (int dd = 0; dd < 2; ++dd)                                        // dW_out[:, dd*H:(dd+1)*H] = dpred^T out_dd
        GemmB().A(D.dpredT_p, nk, nk).Bm(D.g.outT_p[dd], nk, nk)
            .run(F, Hd, G + L.h2o_w[i] + (long)dd * Hd, 2 * Hd, nullptr, 1, splits_for(F, Hd, nk), sw);
    }
    g_bwd_concurrent = (ndec == 2 && g_opt_streams) ? 2 : 1;
    const bool have_sums = gru_sweep_bwd(Wg, D.g, w.tiles, D.ddec, D.ddec + (size_t)Hd * rows, rows, nullptr, nullptr, 0, true, sd,
                                         D.dgi_sum, D.dgi_sum_p, &o, G);
    g_bwd_concurrent = 1;
    if (i == 0) mark(st, "bwd:dec sweep");
    edge(sd, sw);                            // recurrent weight gradients start as soon as the sweep is done (beside the dz chain)
    gru_recurrent_grads(o, D.g, Bp, D.g.h0T_p[0], D.g.h0T_p[1], G, sw);
    if (!have_sums) {
      for (int dd = 0; dd < 2; ++dd) {       // the input is z at every step -> reduce dgi over time first
        launch_timesum_fm(D.g.dgi[dd], rows, steps, Bp, 3 * Hd, D.dgi_sum[dd], sd);              // [3H][B_pad]
        pack_T(D.dgi_sum[dd], Bp, Bp, 3 * Hd, 3 * Hd, D.dgi_sum_p[dd], sd);                      // -> [B_pad rows, K = 3H]
      }
    }
    // latent_to_hidden backward through the inverse of the .view(2,B,H) quirk
    launch_parts_reduce_pack(final_parts(D.g, 0, w.tiles), D.g.parts_n,
                             (long)(final_parts(D.g, 1, w.tiles) - final_parts(D.g, 0, w.tiles)), 2, B, Bp, Hd, D.dhid, D.dhid_p, sd);
    // dz = [dgi_sum_f, dgi_sum_b, dhid] [W_ih_f ; W_ih_b ; W_l2h]: one split-K GEMM over the concatenated K
    GemmB().A(D.dgi_sum_p[0], nkc3, nkc3).A(D.dgi_sum_p[1], nkc3, nkc3).A(D.dhid_p, nkc2H, nkc2H)
        .Bm(Wg.wihT_p[0], nkc3, nkc3).Bm(Wg.wihT_p[1], nkc3, nkc3).Bm(W.l2hT_p[i], nkc2H, nkc2H)
        .run(B, Z, D.dz, Z, nullptr, 1, 8, sd);
    dz_dec[i] = D.dz;
    // ---- remaining weight gradients of this decoder (they need dgi_sum / dhid of the dz chain)
    edge(sd, sw);
    if (i == 1) edge(sA, st);                // main needs dz of the future decoder, not its weight gradients
    for (int dd = 0; dd < 2; ++dd) {
      pack_rows(D.dgi_sum[dd], Bp, 3 * Hd, Bp, 3 * Hd, D.dgi_sumT_p[dd], sw);
      GemmB().A(D.dgi_sumT_p[dd], nkcB, nkcB).Bm(w.zT_p, nkcB, nkcB)
          .run(3 * Hd, Z, G + o.wih[dd], Z, nullptr, 1, splits_for(3 * Hd, Z, nkcB), sw);
    }
    pack_T(D.dhid, 2 * Hd, 2 * Hd, Bp, B, D.dhidT_p, sw);
    GemmB().A(D.dhidT_p, nkcB, nkcB).Bm(w.zT_p, nkcB, nkcB).run(2 * Hd, Z, G + L.l2h_w[i], Z, nullptr, 1, splits_for(2 * Hd, Z, nkcB), sw);
    launch_colsum(D.dhid, 2 * Hd, B, 2 * Hd, G + L.l2h_b[i], sw);
    if (i == 0) mark(sw, "side:decoder weight grads done");
  }

  mark(st, "bwd:dec dz chain");
  // ---- Lambda backward (main chain: dlin -> dhidden)
  if (use_loss_grads && g_opt_streams) edge(side().s[2], st);   // k-means prior gradient (vame_loss may have left it running)
  {
    LambdaBwdArgs a{};
    a.dz[0] = use_loss_grads ? w.dz_km : nullptr;
    a.dz[1] = dz_dec[0]; a.dz[2] = dz_dec[1];
    a.dz[3] = use_loss_grads ? nullptr : dz_ext;
    a.dmu_ext = use_loss_grads ? nullptr : dmu_ext;
    a.dlv_ext = use_loss_grads ? nullptr : dlv_ext;
    a.mu = w.mu; a.logvar = w.logvar; a.eps = w.eps; a.use_eps_flag = reinterpret_cast<const long long*>(w.acc + 7);
    a.lin = w.lin; a.ldl = 2 * Z;
    a.hyper = use_loss_grads ? hyper : nullptr;
    a.c_kl = use_loss_grads ? cfg->beta * cfg->kl_weight / (float)((long)B * Z) : 0.f;
    a.B = B; a.B_pad = Bp; a.Z = Z; a.softplus = d->softplus;
    a.dlin = w.dlin; a.ldd = 2 * Z;
    launch_lambda_bwd_p16(a, w.dlin_p, st);            // dlin and its P16 copy for the dhidden GEMM
  }
  const int nkc2Z = nkc_of(2 * Z);
  edge(st, sB);
  GemmB().A(w.dlin_p, nkc2Z, nkc2Z).Bm(W.lamT_p, nkc2Z, nkc2Z).run_fm(Bp, 4 * H, w.dhidden, Bp, nullptr, st);
  {   // Lambda weight gradients on sB
    pack_T(w.dlin, 2 * Z, 2 * Z, Bp, Bp, w.dlinT_p, sB);
    launch_colsum(w.dlin, 2 * Z, Bp, 2 * Z, G + L.lam_b, sB);
    const float* hf[4] = {final_h(w.e0, 0, w.tiles), final_h(w.e0, 1, w.tiles), final_h(w.e1, 0, w.tiles), final_h(w.e1, 1, w.tiles)};
    const long hld[4] = {final_h_ld(w.e0, w.tiles), final_h_ld(w.e0, w.tiles), final_h_ld(w.e1, w.tiles), final_h_ld(w.e1, w.tiles)};
    for (int i = 0; i < 4; ++i) {
      pack_rows(hf[i], hld[i], H, Bp, H, w.hidT_p[i], sB);                 // final h is [H][.] feature-major
      GemmB().A(w.dlinT_p, nkcB, nkcB).Bm(w.hidT_p[i], nkcB, nkcB)
          .run(2 * Z, H, G + L.lam_w + (long)i * H, 4 * H, nullptr, 1, splits_for(2 * Z, H, nkcB), sB);
    }
    mark(sB, "side:lambda weight grads done");
  }

  mark(st, "bwd:lambda bwd + dhidden gemm");
  // ---- encoder layer 1 (only h_n is used downstream, rnn_model.py:41-43: no per-step output gradient)
  const long rows = (long)T * Bp;
  const int nk = (int)(rows / KCHUNK), nkc3 = nkc_of(3 * H);
  gru_sweep_bwd(W.e1, w.e1, w.tiles, nullptr, nullptr, 0, w.dhidden + (size_t)2 * H * Bp, w.dhidden + (size_t)3 * H * Bp, Bp, true, st,
                nullptr, nullptr, &L.e1, G);
  mark(st, "bwd:L1 sweep");
  // the weight gradients of the two directions are independent: one side stream each (their fixed per-kernel costs overlap)
  cudaStream_t sC = g_opt_streams ? side().s[3] : st;
  cudaStream_t sdir[2] = {sB, sC};
  const int side_total = g_side_sms;
  if (g_opt_streams) g_side_sms = side_total / g_opt_side_split;
  edge(st, sB);
  edge(st, sC);
  w.e0.dout_pv = w.e0.priv && H == 256 && rw_priv_mode(H, w.tiles);
  {
    GemmB gb;
    gb.A(w.e1.dgi_p[0], nkc3, nkc3).A(w.e1.dgi_p[1], nkc3, nkc3).Bm(W.e1.wihT_p[0], nkc3, nkc3).Bm(W.e1.wihT_p[1], nkc3, nkc3);
    if (w.e0.dout_pv) gb.run_pv((int)rows, 2 * H, w.dx1, rows, Bp, st);
    else gb.run_fm((int)rows, 2 * H, w.dx1, rows, nullptr, st);
  }
  gru_recurrent_grads(L.e1, w.e1, Bp, w.zeros_p, w.zeros_p, G, sB, true, sC);
  for (int dd = 0; dd < 2; ++dd) {           // dW_ih(l1)[dd][:, e*H:(e+1)*H] = dgi1[dd]^T out0[e]
    if (!w.e1.priv) launch_pack_p16_rowsum(w.e1.dgi[dd], rows, 3 * H, (int)rows, 3 * H, w.e1.dgiT_p[dd], G + L.e1.bih[dd], sdir[dd]);
    for (int e = 0; e < 2; ++e)
      GemmB().A(w.e1.dgiT_p[dd], nk, nk).Bm(w.e0.outT_p[e], nk, nk)
          .run(3 * H, H, G + L.e1.wih[dd] + (long)e * H, 2 * H, nullptr, 1, splits_for(3 * H, H, nk), sdir[dd]);
  }
  mark(sB, "side:L1 weight grads done");
  if (grad_overlap().on) {
    // every gradient outside encoder layer 0 has been enqueued: decoder(s), Lambda and layer-1 weight gradients live on sB / sC /
    // sA; once they are complete the first bucket [vame_grad_bucket_split, total) may be all-reduced
    edge(sC, sB);
    if (used_sA) edge(sA, sB);
    edge(st, sB);          // ... and the dx1 GEMM (main stream) has read the layer-1 weights: the consumer may re-pack them
    // (the external flag - an event-record NODE that code outside the graph can wait for - is only legal during stream capture;
    //  mode 2: the consumer is captured into the same graph, a plain record is the dependency edge)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(sB, &cap);
    cudaEventRecordWithFlags(grad_overlap().ev, sB,
                             (cap == cudaStreamCaptureStatusActive && !grad_overlap().internal) ? cudaEventRecordExternal : cudaEventRecordDefault);
  }
  // ---- encoder layer 0
  mark(st, "bwd:dx1 gemm");
  gru_sweep_bwd(W.e0, w.e0, w.tiles, w.dx1, w.dx1 + (size_t)H * rows, rows, w.dhidden, w.dhidden + (size_t)H * Bp, Bp, true, st,
                nullptr, nullptr, &L.e0, G);
  edge(st, sB);
  edge(st, sC);
  g_side_sms = g_opt_streams ? 148 / g_opt_side_split : 148;     // nothing else runs beside the last block: half of the GPU per direction
  gru_recurrent_grads(L.e0, w.e0, Bp, w.zeros_p, w.zeros_p, G, sB, true, sC);
  for (int dd = 0; dd < 2; ++dd) {           // dW_ih(l0)[dd] = dgi0[dd]^T x
    if (!w.e0.priv) launch_pack_p16_rowsum(w.e0.dgi[dd], rows, 3 * H, (int)rows, 3 * H, w.e0.dgiT_p[dd], G + L.e0.bih[dd], sdir[dd]);
    GemmB().A(w.e0.dgiT_p[dd], nk, nk).Bm(w.xT_p, nk, nk)
        .run(3 * H, F, G + L.e0.wih[dd], F, nullptr, 1, splits_for(3 * H, F, nk), sdir[dd]);
  }
  mark(st, "bwd:L0 sweep");
  edge(sB, st);
  edge(sC, st);
  mark(st, "bwd:side-stream tail (weight gradients)");
  if (used_sA) edge(sA, st);                 // only streams that were forked into this call may be joined (graph capture isolation)
  return check_launch("vame_backward");
}

int vame_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long n, float lr,
                   const float* hyper, int* step_dev, float* scratch, float beta1, float beta2, float eps, float grad_scale,
                   void* stream) {
  VB_REQUIRE(params && grads && exp_avg && exp_avg_sq && max_exp_avg_sq && step_dev && scratch, "vame_adam_step: null pointer");
  VB_REQUIRE(n > 0, "vame_adam_step: empty");
  launch_adam(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr, hyper, step_dev, scratch, beta1, beta2, eps, grad_scale,
              (cudaStream_t)stream);
  return check_launch("vame_adam_step");
}

/* Re-run one recurrent sweep of encoder layer 1 on the buffers of the last vame_forward(save=1) [+ vame_backward]:
 * which = 0 forward steps, 1 backward steps.  Used by bench.py to time the dominant kernel with CUDA events. */
int vame_adam_prepare(float lr, const float* hyper, int* step_dev, float* scratch, float beta1, float beta2, void* stream) {
  VB_REQUIRE(step_dev && scratch, "vame_adam_prepare: null pointer");
  launch_adam_prepare(lr, hyper, step_dev, scratch, beta1, beta2, (cudaStream_t)stream);
  return check_launch("vame_adam_prepare");
}

int vame_adam_apply(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long n,
                    const float* scratch, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  VB_REQUIRE(params && grads && exp_avg && exp_avg_sq && max_exp_avg_sq && scratch, "vame_adam_apply: null pointer");
  VB_REQUIRE(n > 0 && n % 4 == 0, "vame_adam_apply: the range must be a positive multiple of 4 floats");
  launch_adam_apply(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, n, scratch, beta1, beta2, eps, grad_scale, (cudaStream_t)stream);
  return check_launch("vame_adam_apply");
}

int vame_debug_gru_sweep(const vame_dims* d, int batch, int which, const float* params, const void* packed, void* ws, size_t ws_bytes,
                         void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && ws, "vame_debug_gru_sweep: null pointer");
  Ws w = carve_ws(*d, batch, true, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_debug_gru_sweep: workspace too small");
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int H = d->hidden_enc;
  for (int dd = 0; dd < 2; ++dd) { w.e1.h0[dd] = w.zeros_f32; w.e1.h0_p[dd] = w.zeros_p; }
  cudaStream_t st = (cudaStream_t)stream;
  if (which == 0) gru_sweep_fwd(W.e1, params + L.e1.bhh[0] + 2 * H, params + L.e1.bhh[1] + 2 * H, w.e1, w.tiles, true, true, st);
  else gru_sweep_bwd(W.e1, w.e1, w.tiles, nullptr, nullptr, 0, w.dhidden + (size_t)2 * H * w.B_pad, w.dhidden + (size_t)3 * H * w.B_pad, w.B_pad, true, st,
                     nullptr, nullptr, &L.e1, w.dx1 /* scratch target of the bias-gradient sums (timing only) */);
  return check_launch("vame_debug_gru_sweep");
}

/* measurement hook: enable = 1 starts recording section boundaries (cudaEvents on the caller's stream) in the next
 * vame_forward / vame_backward calls; vame_debug_timeline_read synchronises and returns the number of sections, their
 * names (static strings) and durations in ms. */
int vame_debug_timeline(int enable) {
  Timeline& T = timeline();
  T.on = enable != 0;
  T.n = 0;
  return 0;
}
int vame_debug_timeline_read(const char** names, float* ms, int max) {
  Timeline& T = timeline();
  cudaDeviceSynchronize();
  int n = 0;
  for (int i = 1; i < T.n && n < max; ++i) {     // absolute time of every mark since the first one (marks may be on side streams)
    float t = 0.f;
    cudaEventElapsedTime(&t, T.ev[0], T.ev[i]);
    names[n] = T.name[i];
    ms[n++] = t;
  }
  T.n = 0;
  return n;
}

int vame_cluster_loss(const float* latent, int batch, int zdims, int kloss, float lmbda, float bsize, float grad_coef, double* loss_out,
                      float* dlatent, void* stream) {
  VB_REQUIRE(latent && loss_out && batch > 0 && zdims > 0 && zdims <= 64, "vame_cluster_loss: bad arguments (Z <= 64)");
  // the kernel writes acc[ACC_KMEANS]; give it a pointer such that this slot is loss_out[0]
  launch_cluster_prior(latent, batch, zdims, kloss, lmbda, bsize, grad_coef, nullptr, dlatent, loss_out - ACC_KMEANS,
                       (cudaStream_t)stream);
  return check_launch("vame_cluster_loss");
}

}  // extern "C"

// ================================================================================================
// inference entry points: sub-module forwards and the sliding-window embedding
// ================================================================================================
namespace vb {

struct EmbedWs {
  int Bc, Bc_pad, tiles;
  void* series_p;
  float* G; long G_rows;
  float* zeros_f32; void* zeros_p; size_t zeros_p_bytes;
  GruBuf e0, e1;
  float* lin; float* scratch_z; float* scratch_lv;
  double* acc;
  size_t bytes;
};

static EmbedWs carve_embed(const vame_dims& d, long n_frames, int chunk, void* base) {
  EmbedWs w{};
  Arena A(base);
  const int T = d.time_window, F = d.num_features, Z = d.zdims, H = d.hidden_enc;
  w.Bc = chunk; w.Bc_pad = pad128(chunk); w.tiles = w.Bc_pad / 128;
  const long Np = (n_frames + 127) / 128 * 128;
  w.series_p = A.raw(p16_bytes((int)Np, F, 128));
  w.G_rows = Np + w.Bc_pad + T + 128;
  w.G = A.f32((size_t)w.G_rows * 6 * H);                 // feature-major [6H][G_rows]
  w.zeros_f32 = A.f32((size_t)w.Bc_pad * H);
  w.zeros_p_bytes = (size_t)w.tiles * nkc_of(H) * p16_tile_bytes(128);
  w.zeros_p = A.raw(w.zeros_p_bytes);
  carve_gru(A, w.e0, T, H, F, w.Bc_pad, false, true, /*gi_full=*/false, false, false, false);
  carve_gru(A, w.e1, T, H, 2 * H, w.Bc_pad, false, false, /*gi_full=*/true, false, false, false);
  w.lin = A.f32((size_t)w.Bc_pad * 2 * Z);
  w.scratch_z = A.f32((size_t)w.Bc_pad * Z);
  w.scratch_lv = A.f32((size_t)w.Bc_pad * Z);
  w.acc = (double*)A.raw(8 * sizeof(double));
  w.bytes = (A.off + 1023) & ~(size_t)1023;
  return w;
}

}  // namespace vb

extern "C" {

size_t vame_embed_workspace_bytes(const vame_dims* d, long n_frames, int chunk) {
  if (check_dims(d) || n_frames <= 0 || chunk <= 0) return 0;
  return carve_embed(*d, n_frames, chunk, nullptr).bytes;
}

int vame_embed_windows(const vame_dims* d, const float* params, const void* packed, const float* series, long n_frames,
                       long first_window, long n_windows, int chunk, float* mu_out, void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(params && packed && series && mu_out && ws, "vame_embed_windows: null pointer");
  const int T = d->time_window, F = d->num_features, Z = d->zdims, H = d->hidden_enc;
  VB_REQUIRE(chunk > 0 && n_windows >= 0 && first_window >= 0, "vame_embed_windows: bad range");
  VB_REQUIRE(first_window + n_windows + T - 1 <= n_frames, "vame_embed_windows: window range exceeds the series (a window needs T frames)");
  VB_REQUIRE(n_frames < (1L << 31) - 1024, "vame_embed_windows: series too long for 32-bit row indices");
  EmbedWs w = carve_embed(*d, n_frames, chunk, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_embed_windows: workspace too small (see vame_embed_workspace_bytes)");
  if (n_windows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  join_deferred_pack(st);
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int nkcF = nkc_of(F), nkcH = nkc_of(H);
  const long Np = (n_frames + 127) / 128 * 128;
  // once per call: the layer-0 input projection of every FRAME (a window's step t reads row i + t)
  pack_rows(series, F, (int)Np, F, (int)n_frames, w.series_p, st);
  cudaMemsetAsync(w.G, 0, (size_t)w.G_rows * 6 * H * 4, st);       // rows beyond n_frames are read by padding windows only
  GemmB().A(w.series_p, nkcF, nkcF).Bm(W.e0.wih_p[0], nkcF, nkcF).run_fm((int)n_frames, 6 * H, w.G, w.G_rows, W.e0.bias_gi, st);
  cudaMemsetAsync(w.zeros_f32, 0, (size_t)w.Bc_pad * H * 4, st);
  cudaMemsetAsync(w.zeros_p, 0, w.zeros_p_bytes, st);
  for (int dd = 0; dd < 2; ++dd) {
    w.e0.h0[dd] = w.zeros_f32; w.e0.h0_p[dd] = w.zeros_p;
    w.e1.h0[dd] = w.zeros_f32; w.e1.h0_p[dd] = w.zeros_p;
  }
  for (long i0 = 0; i0 < n_windows; i0 += w.Bc) {
    const int Bc = (int)((n_windows - i0 < w.Bc) ? (n_windows - i0) : w.Bc);
    const int Bp = pad128(Bc), tiles = Bp / 128;
    // layer 0 reads its projections straight out of G: row(b, t) = first_window + i0 + b + t
    w.e0.gi = w.G + (size_t)(first_window + i0);
    w.e0.gi_bs = 1; w.e0.gi_ts = 1; w.e0.gi_ld = w.G_rows;
    gru_sweep_fwd(W.e0, params + L.e0.bhh[0] + 2 * H, params + L.e0.bhh[1] + 2 * H, w.e0, tiles, false, true, st);
    // NOTE: buffers are laid out for Bc_pad rows per time step; a short last chunk uses the first `tiles` tiles of each slot
    // only if the slot stride matches, so the sweep above is run with the chunk's own tile count and slot strides.
    w.e1.gi_bs = 1; w.e1.gi_ts = Bp; w.e1.gi_ld = (long)T * Bp;
    GemmB().A(w.e0.out_p[0], nkcH, nkcH).A(w.e0.out_p[1], nkcH, nkcH).Bm(W.e1.wih_p[0], nkcH, nkcH).Bm(W.e1.wih_p[1], nkcH, nkcH)
        .run_fm(T * Bp, 6 * H, w.e1.gi, (long)T * Bp, W.e1.bias_gi, st);
    gru_sweep_fwd(W.e1, params + L.e1.bhh[0] + 2 * H, params + L.e1.bhh[1] + 2 * H, w.e1, tiles, false, true, st);
    GemmB gb;
    gb.A(final_h_p(w.e0, 0, tiles), nkcH, nkcH).A(final_h_p(w.e0, 1, tiles), nkcH, nkcH)
        .A(final_h_p(w.e1, 0, tiles), nkcH, nkcH).A(final_h_p(w.e1, 1, tiles), nkcH, nkcH);
    for (int i = 0; i < 4; ++i) gb.Bm(W.lam_p[i], nkcH, nkcH);
    gb.run(Bc, 2 * Z, w.lin, 2 * Z, params + L.lam_b, 0, 1, st);
    launch_lambda_fwd(w.lin, 2 * Z, nullptr, Bc, Z, d->softplus, w.scratch_z, mu_out + (size_t)i0 * Z, w.scratch_lv, nullptr, st);
  }
  return check_launch("vame_embed_windows");
}

int vame_encoder_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* x, long x_bs, long x_ts,
                         float* hidden, void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && x && hidden && ws, "vame_encoder_forward: null pointer");
  Ws w = carve_ws(*d, batch, false, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_encoder_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int H = d->hidden_enc;
  encoder_forward(*d, params, L, W, w, x, x_bs, x_ts, false, st);
  const float* hf[4] = {final_h(w.e0, 0, w.tiles), final_h(w.e0, 1, w.tiles), final_h(w.e1, 0, w.tiles), final_h(w.e1, 1, w.tiles)};
  const long hld[4] = {final_h_ld(w.e0, w.tiles), final_h_ld(w.e0, w.tiles), final_h_ld(w.e1, w.tiles), final_h_ld(w.e1, w.tiles)};
  for (int i = 0; i < 4; ++i)      // torch.cat((h_n[0], h_n[1], h_n[2], h_n[3]), 1)  (rnn_model.py:43)
    launch_fm_to_rows(hf[i], hld[i], H, batch, hidden + (size_t)i * H, 4 * H, st);
  return check_launch("vame_encoder_forward");
}

int vame_lambda_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* hidden, const float* eps,
                        float* z, float* mu, float* logvar, void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && hidden && z && mu && logvar && ws, "vame_lambda_forward: null pointer");
  Ws w = carve_ws(*d, batch, false, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_lambda_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  join_deferred_pack(st);
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  const int H = d->hidden_enc, Z = d->zdims;
  const void* hp[4];
  for (int i = 0; i < 4; ++i) {
    pack_rows(hidden + (size_t)i * H, 4 * H, w.B_pad, H, batch, w.hid_p[i], st);
    hp[i] = w.hid_p[i];
  }
  lambda_linear(*d, params, L, W, w, hp, st);
  launch_lambda_fwd(w.lin, 2 * Z, eps, batch, Z, d->softplus, z, mu, logvar, nullptr, st);
  return check_launch("vame_lambda_forward");
}

int vame_decoder_forward(const vame_dims* d, int batch, int which, const float* params, const void* packed, const float* z, float* pred,
                         void* ws, size_t ws_bytes, void* stream) {
  if (check_dims(d)) return -1;
  VB_REQUIRE(batch > 0 && params && packed && z && pred && ws, "vame_decoder_forward: null pointer");
  VB_REQUIRE(which == 0 || (which == 1 && d->future_decoder), "vame_decoder_forward: which must be 0 (decoder) or 1 (decoder_future)");
  Ws w = carve_ws(*d, batch, false, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_decoder_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  join_deferred_pack(st);
  const ParamLayout L = param_layout(*d);
  const PackedWeights W = packed_layout(*d, const_cast<void*>(packed));
  pack_rows(z, d->zdims, w.B_pad, d->zdims, batch, w.z_p, st);
  decoder_forward(*d, which, params, L, W, w, false, st);
  launch_tb_to_bt(w.dec[which].pred_tb, batch, w.dec[which].g.steps, d->num_features, w.B_pad, pred, st);
  return check_launch("vame_decoder_forward");
}

}  // extern "C"
