// Host-side layout of the RNN-VAE: parameter offsets in the flat buffer, packed-weight cache and per-batch workspace.
// All three are computed by running the same "carve" code over an Arena, once without a base pointer to get the size
// (vame_*_bytes) and once with the caller's buffer to get the pointers — the caller (PyTorch) owns every allocation.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/vame_b200.h"
#include "common.cuh"

namespace vb {

struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* b) : base((char*)b), off(0) {}
  void* raw(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;          // 1 KB alignment (>= 16 B needed by the TMA engine)
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f32(size_t n) { return (float*)raw(n * 4); }
};

inline int pad128(int b) { return (b + 127) / 128 * 128; }
inline int nkc_of(int k) { return (k + KCHUNK - 1) / KCHUNK; }

// ---- parameters -------------------------------------------------------------------------------------
struct GruOff {          // offsets (in floats) into the flat parameter / gradient buffer
  long wih[2], whh[2], bih[2], bhh[2];
  int In, H;
};
struct ParamLayout {
  GruOff e0, e1, dec, fut;
  long lam_w, lam_b;                 // [2Z, 4H] = [W_mu; W_lv], [2Z]
  long l2h_w[2], l2h_b[2], h2o_w[2], h2o_b[2];   // index 0 = decoder, 1 = decoder_future
  long total;
};

inline long carve_gru(long& off, GruOff& g, int In, int H) {
  g.In = In;
  g.H = H;
  off = (off + 3) & ~3L;
  g.wih[0] = off; off += 3L * H * In;
  g.wih[1] = off; off += 3L * H * In;
  g.whh[0] = off; off += 3L * H * H;
  g.whh[1] = off; off += 3L * H * H;
  g.bih[0] = off; off += 3L * H;
  g.bih[1] = off; off += 3L * H;
  g.bhh[0] = off; off += 3L * H;
  g.bhh[1] = off; off += 3L * H;
  return off;
}

inline ParamLayout param_layout(const vame_dims& d) {
  ParamLayout L{};
  long off = 0;
  const int H = d.hidden_enc, F = d.num_features, Z = d.zdims;
  carve_gru(off, L.e0, F, H);
  carve_gru(off, L.e1, 2 * H, H);
  off = (off + 3) & ~3L;
  L.lam_w = off; off += 2L * Z * 4 * H;
  L.lam_b = off; off += 2L * Z;
  for (int i = 0; i < (d.future_decoder ? 2 : 1); ++i) {
    const int Hd = i == 0 ? d.hidden_rec : d.hidden_pred;
    carve_gru(off, i == 0 ? L.dec : L.fut, Z, Hd);
    off = (off + 3) & ~3L;
    L.l2h_w[i] = off; off += 2L * Hd * Z;
    L.l2h_b[i] = off; off += 2L * Hd;
    off = (off + 3) & ~3L;
    L.h2o_w[i] = off; off += (long)F * 2 * Hd;
    L.h2o_b[i] = off; off += F;
  }
  L.total = (off + 3) & ~3L;
  return L;
}

// ---- packed weights ------------------------------------------------------------------------------------
struct GruPacked {
  void* whh_p[2];      // forward-step slices
  void* whhT_p[2];     // backward-step slices
  void* whh_rw[2];     // resident-weight forward / backward formats of gru_rw.cu (nullptr when H % 64 != 0)
  void* whhT_rw[2];
  void* whh_rows[2];   // row-resident inference format of gru_rows.cu: [H/32 slices][KC][hi 96x64 | lo 96x64] (nullptr unless H == 256)
  void* wih_p[2];      // input projection B operand, K segments (encoder layer 1 has two: fwd / bwd halves of its input)
  int wih_nseg;
  void* wihT_p[2];     // per direction: [In rows, K = 3H] (for dx / dz)
  float* bias_gi;      // [6H]
};
struct PackedWeights {
  GruPacked e0, e1, dec, fut;
  void* lam_p[4];      // [2Z, K = H] per hidden piece
  void* lamT_p;        // [4H rows, K = 2Z]
  void* l2h_p[2];      // [2H, K = Z]
  void* l2hT_p[2];     // [Z rows, K = 2H]
  void* h2o_p[2][2];   // per decoder, per direction: [F, K = H]
  void* h2oT_p[2];     // [2H rows, K = F]
  size_t bytes;
};

inline void carve_gru_packed(Arena& A, GruPacked& g, int In, int H, int nseg) {
  for (int d = 0; d < 2; ++d) {
    g.whh_p[d] = A.raw(p16_bytes(3 * H, H, 96));
    g.whhT_p[d] = A.raw((size_t)(H / 32) * p16_bytes(H, 128, 128));
    g.wihT_p[d] = A.raw(p16_bytes(In, 3 * H, 128));
    if (H % 64 == 0) {     // [4 CTAs][3 gates][H/64] and [4 CTAs][H/64][ceil(3H/4 / 64)] A tiles of 128 x 64 bf16
      g.whh_rw[d] = A.raw((size_t)4 * 3 * (H / 64) * 16384);
      g.whhT_rw[d] = A.raw((size_t)4 * (H / 64) * ((3 * (H / 4) + KCHUNK - 1) / KCHUNK) * 16384);
    } else {
      g.whh_rw[d] = g.whhT_rw[d] = nullptr;
    }
    g.whh_rows[d] = H == 256 ? A.raw((size_t)(H / 32) * (H / KCHUNK) * p16_tile_bytes(96)) : nullptr;
  }
  g.wih_nseg = nseg;
  for (int s = 0; s < nseg; ++s) g.wih_p[s] = A.raw(p16_bytes(6 * H, In / nseg, 128));
  g.bias_gi = A.f32(6 * H);
}

inline PackedWeights packed_layout(const vame_dims& d, void* base) {
  PackedWeights P{};
  Arena A(base);
  const int H = d.hidden_enc, F = d.num_features, Z = d.zdims;
  carve_gru_packed(A, P.e0, F, H, 1);
  carve_gru_packed(A, P.e1, 2 * H, H, 2);
  for (int i = 0; i < 4; ++i) P.lam_p[i] = A.raw(p16_bytes(2 * Z, H, 128));
  P.lamT_p = A.raw(p16_bytes(4 * H, 2 * Z, 128));
  for (int i = 0; i < (d.future_decoder ? 2 : 1); ++i) {
    const int Hd = i == 0 ? d.hidden_rec : d.hidden_pred;
    carve_gru_packed(A, i == 0 ? P.dec : P.fut, Z, Hd, 1);
    P.l2h_p[i] = A.raw(p16_bytes(2 * Hd, Z, 128));
    P.l2hT_p[i] = A.raw(p16_bytes(Z, 2 * Hd, 128));
    for (int dd = 0; dd < 2; ++dd) P.h2o_p[i][dd] = A.raw(p16_bytes(F, Hd, 128));
    P.h2oT_p[i] = A.raw(p16_bytes(2 * Hd, F, 128));
  }
  P.bytes = (A.off + 1023) & ~(size_t)1023;
  return P;
}

}  // namespace vb
