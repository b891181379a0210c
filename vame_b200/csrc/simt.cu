// HBM-bound elementwise / reduction kernels of the RNN-VAE hot path: Lambda reparameterisation + KL, MSE loss with its
// gradient, the k-means prior (cluster_loss) via a Z x Z Gram matrix and an in-kernel fp64 Jacobi eigensolver,
// bias-gradient column sums, AMSGrad Adam.  Vectorised global accesses, warp-shuffle reductions, one atomic per warp/CTA.
#include "common.cuh"
#include "kernels.h"
#include "simt.h"

namespace vb {

// -------------------------------------------------------------------------------------------------
// Lambda forward (vame/model/rnn_model.py:63-76) on the fused [mu | logvar] linear output + KL partial sum
// (vame/model/rnn_vae.py:53-60: KLD = -0.5 * mean(1 + logvar - mu^2 - exp(logvar)))
//   lin: [B_pad, ldl] with mu in cols [0,Z), pre-activation logvar in [Z,2Z)
//   eps == nullptr -> eval mode, z = mu.   acc[ACC_KL] += sum(1 + lv - mu^2 - e^lv)
// -------------------------------------------------------------------------------------------------
__global__ void lambda_fwd_kernel(const float* __restrict__ lin, long ldl, const float* __restrict__ eps, int B, int Z,
                                  int softplus, float* __restrict__ z, float* __restrict__ mu, float* __restrict__ logvar,
                                  double* __restrict__ acc) {
  const long total = (long)B * Z;
  float part = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Z);
    const long b = idx / Z;
    const float m = lin[b * ldl + j];
    float lv = lin[b * ldl + Z + j];
    if (softplus) lv = (lv > 20.f) ? lv : log1pf(expf(lv));
    mu[idx] = m;
    logvar[idx] = lv;
    z[idx] = eps ? eps[idx] * expf(0.5f * lv) + m : m;
    part += 1.f + lv - m * m - expf(lv);
  }
  part = warp_sum(part);
  if (acc && (threadIdx.x & 31) == 0) atomicAdd(acc + ACC_KL, (double)part);
}

// Lambda backward: dlin = [dmu | dlv_lin] for valid rows, 0 for padded rows.
//   dz = sum of up to 4 gradient pieces (cluster prior, decoder, future decoder, external)
//   dmu = dz + dmu_ext + c_kl * mu ;  dlv = dz * eps * 0.5 * exp(0.5 lv) + dlv_ext + c_kl * 0.5 * (exp(lv) - 1)
__global__ void lambda_bwd_kernel(LambdaBwdArgs a) {
  const long total = (long)a.B_pad * a.Z;
  const float c_kl = a.hyper ? a.hyper[HY_BETA] * a.hyper[HY_KLW] / (float)((long)a.B * a.Z) : a.c_kl;
  const bool use_eps = a.eps && (!a.use_eps_flag || *a.use_eps_flag != 0);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % a.Z);
    const long b = idx / a.Z;
    float dmu = 0.f, dlv = 0.f;
    if (b < a.B) {
      const long i = b * a.Z + j;
      float dz = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (a.dz[p]) dz += a.dz[p][i];
      const float m = a.mu[i], lv = a.logvar[i];
      dmu = dz + c_kl * m + (a.dmu_ext ? a.dmu_ext[i] : 0.f);
      dlv = c_kl * 0.5f * (expf(lv) - 1.f) + (a.dlv_ext ? a.dlv_ext[i] : 0.f);
      if (use_eps) dlv += dz * a.eps[i] * 0.5f * expf(0.5f * lv);
      if (a.softplus) {
        const float x = a.lin[b * a.ldl + a.Z + j];
        dlv *= 1.f / (1.f + expf(-x));
      }
    }
    a.dlin[b * a.ldd + j] = dmu;
    a.dlin[b * a.ldd + a.Z + j] = dlv;
  }
}

// -------------------------------------------------------------------------------------------------
// MSE (vame/model/rnn_vae.py:35-43, nn.MSELoss(reduction)) on time-major buffers + gradient
//   pred, target: [T * B_pad, F] ; rows with (row % B_pad) >= B are padding.  acc[slot] += sum (pred - target)^2
//   dpred = gscale * (pred - target) on valid rows, 0 on padding (gscale = 2 for 'sum', 2/(B*T*F) for 'mean')
// -------------------------------------------------------------------------------------------------
__global__ void mse_kernel(const float* __restrict__ pred, long ldp, const float* __restrict__ target, int rows, int B,
                           int B_pad, int F, float gscale, float* __restrict__ dpred, double* __restrict__ acc, int slot) {
  const long total = (long)rows * F;
  float part = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int f = (int)(idx % F);
    const long r = idx / F;
    float d = 0.f;
    if ((int)(r % B_pad) < B) d = pred[r * ldp + f] - target[idx];
    part += d * d;
    if (dpred) dpred[idx] = gscale * d;
  }
  // one fp64 atomic per block (thousands of same-address atomics from every warp serialised in L2: 12 us for 184 K elements)
  __shared__ float wsum[8];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0 && acc) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)wsum[i];
    atomicAdd(acc + slot, t);
  }
}

// -------------------------------------------------------------------------------------------------
// k-means prior (vame/model/rnn_vae.py:45-50 called as cluster_loss(latent.T, k, lambda, bsize) at :126,:137):
//   loss = lambda * sum_{i<k} sqrt(sv_i(latent latent^T / bsize)).  The B x B matrix's non-zero singular values are
//   the eigenvalues of G = latent^T latent / bsize (Z x Z), so G is built with a coalesced batch reduction and
//   diagonalised in-kernel by a parallel cyclic Jacobi sweep in fp64 (Z <= 64).
//   grad wrt latent = coef * 2 * latent * (V_k diag(0.5 / sqrt(ev)) V_k^T) / bsize.
// One CTA of 1024 threads.
// -------------------------------------------------------------------------------------------------
constexpr int CP_MAXZ = 64;
constexpr size_t CP_SMEM = 5 * CP_MAXZ * (CP_MAXZ + 1) * 8 + (3 * CP_MAXZ + 1) * 8 + (CP_MAXZ * 2) * 4 + 64 * (CP_MAXZ + 1) * 4 + 64;

// ------------------------------------------------------------------------------------------------
// Variants that also write the P16 operand of the GEMM that follows (one thread per 8-wide k atom of one row), so that the
// train step's main chain needs no separate pack launch after them.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_p16_atom(__nv_bfloat16* out, int nkc, long row, int k, const float* v) {   // k % 8 == 0
  uint4 hi, lo;
  split8(v, hi, lo);
  __nv_bfloat16* tile = out + ((size_t)(row >> 7) * nkc + (k >> 6)) * p16_tile_elems(128);
  const int off = p16_in_tile((int)(row & 127), k & 63);
  *reinterpret_cast<uint4*>(tile + off) = hi;
  *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
}
// Lambda forward + z as P16 [B_pad rows, K = Z -> padded to 64-chunks]
__global__ void lambda_fwd_p16_kernel(const float* __restrict__ lin, long ldl, const float* __restrict__ eps, int B, int B_pad, int Z,
                                      int softplus, float* __restrict__ z, float* __restrict__ mu, float* __restrict__ logvar,
                                      double* __restrict__ acc, __nv_bfloat16* __restrict__ z_p) {
  const int nkc = (Z + KCHUNK - 1) / KCHUNK, apr = nkc * 8;
  const long total = (long)B_pad * apr;
  float part = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % apr) * 8;
    const long b = idx / apr;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = k0 + i;
      v[i] = 0.f;
      if (b < B && j < Z) {
        const float m = lin[b * ldl + j];
        float lv = lin[b * ldl + Z + j];
        if (softplus) lv = (lv > 20.f) ? lv : log1pf(expf(lv));
        const long o = b * Z + j;
        mu[o] = m;
        logvar[o] = lv;
        v[i] = eps ? eps[o] * expf(0.5f * lv) + m : m;
        z[o] = v[i];
        part += 1.f + lv - m * m - expf(lv);
      }
    }
    store_p16_atom(z_p, nkc, b, k0, v);
  }
  part = warp_sum(part);
  if (acc && (threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(acc + ACC_KL, (double)part);
}
// Lambda backward + dlin as P16 [B_pad rows, K = 2Z -> padded]
__global__ void lambda_bwd_p16_kernel(LambdaBwdArgs a, __nv_bfloat16* __restrict__ dlin_p) {
  const int nkc = (2 * a.Z + KCHUNK - 1) / KCHUNK, apr = nkc * 8;
  const long total = (long)a.B_pad * apr;
  const float c_kl = a.hyper ? a.hyper[HY_BETA] * a.hyper[HY_KLW] / (float)((long)a.B * a.Z) : a.c_kl;
  const bool use_eps = a.eps && (!a.use_eps_flag || *a.use_eps_flag != 0);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % apr) * 8;
    const long b = idx / apr;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      v[i] = 0.f;
      if (k < 2 * a.Z) {
        const bool is_lv = k >= a.Z;
        const int j = is_lv ? k - a.Z : k;
        float g = 0.f;
        if (b < a.B) {
          const long o = b * a.Z + j;
          float dz = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (a.dz[p]) dz += a.dz[p][o];
          const float m = a.mu[o], lv = a.logvar[o];
          if (!is_lv) {
            g = dz + c_kl * m + (a.dmu_ext ? a.dmu_ext[o] : 0.f);
          } else {
            g = c_kl * 0.5f * (expf(lv) - 1.f) + (a.dlv_ext ? a.dlv_ext[o] : 0.f);
            if (use_eps) g += dz * a.eps[o] * 0.5f * expf(0.5f * lv);
            if (a.softplus) {
              const float x = a.lin[b * a.ldl + a.Z + j];
              g *= 1.f / (1.f + expf(-x));
            }
          }
        }
        a.dlin[b * a.ldd + k] = g;
        v[i] = g;
      }
    }
    store_p16_atom(dlin_p, nkc, b, k0, v);
  }
}
// MSE + gradient + the gradient as P16 [rows, K = F -> padded]
__global__ void mse_p16_kernel(const float* __restrict__ pred, long ldp, const float* __restrict__ target, int rows, int B, int B_pad,
                               int F, float gscale, float* __restrict__ dpred, double* __restrict__ acc, int slot,
                               __nv_bfloat16* __restrict__ dpred_p) {
  const int nkc = (F + KCHUNK - 1) / KCHUNK, apr = nkc * 8;
  const long total = (long)rows * apr;
  float part = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % apr) * 8;
    const long r = idx / apr;
    const bool live = (int)(r % B_pad) < B;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = k0 + i;
      v[i] = 0.f;
      if (f < F) {
        const float d = live ? pred[r * ldp + f] - target[r * F + f] : 0.f;
        part += d * d;
        v[i] = gscale * d;
        dpred[r * F + f] = v[i];
      }
    }
    store_p16_atom(dpred_p, nkc, r, k0, v);
  }
  __shared__ float wsum[8];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0 && acc) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)wsum[i];
    atomicAdd(acc + slot, t);
  }
}

// state (optional, CP_STATE_DOUBLES doubles that persist between calls, e.g. in the training workspace): the eigenvector basis of
// the previous call.  Consecutive training batches are samples of the same latent distribution, so their Gram matrices are
// close and V_prev^T G V_prev is already almost diagonal: the Jacobi iteration then needs 2-3 sweeps instead of 8-9 (~35 us
// each at Z = 30).  The result does not depend on the starting basis (same off-diagonal stopping criterion); the basis is
// re-started from the identity every 256 calls so that rounding drift of V's orthonormality cannot accumulate, and whenever
// the state does not carry this kernel's tag for the same Z (first call, other model).
constexpr unsigned long long CP_STATE_TAG = 0x5641'4D45'5052'494FULL;     // "VAMEPRIO"
__global__ void __launch_bounds__(1024, 1) cluster_prior_kernel(const float* __restrict__ z, int B, int Z, int kloss,
                                                                double lmbda, double bsize, double gcoef,
                                                                const float* __restrict__ hyper,
                                                                float* __restrict__ dz, double* __restrict__ acc,
                                                                double* __restrict__ state) {
  if (hyper) {
    lmbda = (double)hyper[HY_KMLAMBDA];
    gcoef = (double)hyper[HY_KLW];
  }
  extern __shared__ __align__(16) uint8_t cp_smem[];
  typedef double Row[CP_MAXZ + 1];
  Row* A = reinterpret_cast<Row*>(cp_smem);
  Row* V = A + CP_MAXZ;
  Row* A2 = V + CP_MAXZ;
  Row* V2 = A2 + CP_MAXZ;
  Row* T5 = V2 + CP_MAXZ;
  double* own = reinterpret_cast<double*>(T5 + CP_MAXZ);
  double* oth = own + CP_MAXZ;
  double* ev = oth + CP_MAXZ;
  double* offmax_p = ev + CP_MAXZ;
  int* part = reinterpret_cast<int*>(offmax_p + 1);
  int* order = part + CP_MAXZ;
  typedef float ZRow[CP_MAXZ + 1];
  ZRow* zs = reinterpret_cast<ZRow*>(order + CP_MAXZ);
#define offmax (*offmax_p)
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = (Z + 1) & ~1;                       // even size for the round-robin pairing (pad row/col is zero)

  // ---- Gram matrix in fp64: each thread owns entries (i,j) = tid, tid+nt, ... of the n x n matrix
  double g[4] = {0, 0, 0, 0};                        // n*n <= 4096 = 4 * 1024
  for (int b0 = 0; b0 < B; b0 += 64) {
    const int nb = min(64, B - b0);
    for (int e = tid; e < nb * Z; e += nt) zs[e / Z][e % Z] = z[(long)(b0 + e / Z) * Z + e % Z];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int e = tid + s * nt;
      if (e < Z * Z) {
        const int i = e / Z, j = e % Z;
        double t = 0;
        for (int b = 0; b < nb; ++b) t += (double)zs[b][i] * (double)zs[b][j];
        g[s] += t;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < CP_MAXZ * CP_MAXZ; e += nt) {
    A[e / CP_MAXZ][e % CP_MAXZ] = 0;
    V[e / CP_MAXZ][e % CP_MAXZ] = (e / CP_MAXZ == e % CP_MAXZ) ? 1.0 : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int e = tid + s * nt;
    if (e < Z * Z) A[e / Z][e % Z] = g[s] / bsize;
  }
  __syncthreads();
  // ---- full-spectrum case (kmeans_loss >= zdims, the reference's default: cfg['kmeans_loss'] = cfg['zdims']) with a full-rank
  // Gram matrix: sum_i sqrt(ev_i) = trace(G^{1/2}) and the gradient matrix V diag(0.5 / sqrt(ev)) V^T = 0.5 G^{-1/2} need no
  // eigen-decomposition at all.  The coupled Newton-Schulz iteration  Y <- Y (3I - ZY) / 2,  Z <- (3I - ZY) Z / 2  (Y0 = G / tr G,
  // Z0 = I) converges quadratically to (G / tr G)^{1/2} and its inverse with three Z x Z products and two barriers per iteration:
  // ~25-40 cheap iterations instead of 8-9 Jacobi sweeps of ~35 us each (C2, ncu launch list: 315 -> ~120 us per launch).  If it has not
  // converged after 64 iterations (numerically singular Gram) the Jacobi path below takes over from the saved Gram entries.
  bool ns_done = false;
  if (kloss >= Z && B >= Z) {
    double* resid = own;                                        // [2] residual max|ZY - I| of the last two iterations (own/oth are free here)
    double tr = 0;
    for (int i = 0; i < Z; ++i) tr += A[i][i];
    if (tr > 0 && tr == tr) {
      for (int e = tid; e < Z * Z; e += nt) {
        const int i = e / Z, j = e % Z;
        A2[i][j] = A[i][j] / tr;                                 // Y
        V[i][j] = (i == j) ? 1.0 : 0.0;                          // Zm
      }
      if (tid < 2) resid[tid] = 0;
      __syncthreads();
      Row* Y = A2; Row* Zm = V; Row* Yn = A; Row* Zn = V2;
      bool conv = false;
      for (int it = 0; it < 64; ++it) {
        double m = 0;
        for (int e = tid; e < Z * Z; e += nt) {                   // E = 3I - Zm Y, residual of Zm Y against I
          const int i = e / Z, j = e % Z;
          double t = 0;
          for (int q = 0; q < Z; ++q) t += Zm[i][q] * Y[q][j];
          const double dl = (i == j) ? 1.0 : 0.0;
          m = fmax(m, fabs(t - dl));
          T5[i][j] = 3.0 * dl - t;
        }
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((tid & 31) == 0 && m > 0) atomicMax(reinterpret_cast<unsigned long long*>(&resid[it & 1]), (unsigned long long)__double_as_longlong(m));
        __syncthreads();
        const double r = resid[it & 1];
        for (int e = tid; e < Z * Z; e += nt) {
          const int i = e / Z, j = e % Z;
          double a = 0, b = 0;
          for (int q = 0; q < Z; ++q) {
            a += Y[i][q] * T5[q][j];
            b += T5[i][q] * Zm[q][j];
          }
          Yn[i][j] = 0.5 * a;
          Zn[i][j] = 0.5 * b;
        }
        if (tid == 0) resid[(it + 1) & 1] = 0;
        __syncthreads();
        Row* t1 = Y; Y = Yn; Yn = t1;
        Row* t2 = Zm; Zm = Zn; Zn = t2;
        if (!(r == r) || r > 1e6) break;                          // diverging (singular Gram): Jacobi fallback
        if (r < 1e-8) { conv = true; break; }                     // the update just applied squares the error: < 1e-15
      }
      if (conv) {
        const double sq = sqrt(tr);
        if (tid == 0 && acc) {
          double t = 0;
          for (int i = 0; i < Z; ++i) t += Y[i][i];
          acc[ACC_KMEANS] = lmbda * sq * t;
        }
        if (dz) {
          __syncthreads();
          // S = 0.5 G^{-1/2} = 0.5 Zm / sqrt(tr) -> A (read by the gradient pass below); Zm may BE A's buffer: go through T5
          for (int e = tid; e < Z * Z; e += nt) T5[e / Z][e % Z] = 0.5 * Zm[e / Z][e % Z] / sq;
          __syncthreads();
          for (int e = tid; e < Z * Z; e += nt) A[e / Z][e % Z] = T5[e / Z][e % Z];
          __syncthreads();
        }
        ns_done = true;
      } else {                                                    // restore the Gram matrix for the Jacobi path
        __syncthreads();
        for (int e = tid; e < CP_MAXZ * CP_MAXZ; e += nt) {
          A[e / CP_MAXZ][e % CP_MAXZ] = 0;
          V[e / CP_MAXZ][e % CP_MAXZ] = (e / CP_MAXZ == e % CP_MAXZ) ? 1.0 : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int e = tid + s * nt;
          if (e < Z * Z) A[e / Z][e % Z] = g[s] / bsize;
        }
        __syncthreads();
      }
    }
  }
  if (!ns_done) {
  // ---- warm start: A <- V0^T A V0, V <- V0 with the previous call's eigenvector basis
  unsigned long long calls = 0;
  if (state) {
    const unsigned long long* su = reinterpret_cast<const unsigned long long*>(state);
    const bool valid = su[0] == CP_STATE_TAG && su[1] == (unsigned long long)Z;
    calls = valid ? su[2] : 0;
    if (valid && (calls & 255ull) != 0) {                      // (block-uniform: every thread reads the same words)
      const double* v0 = state + 4;
      for (int e = tid; e < Z * Z; e += nt) V[e / Z][e % Z] = v0[e];
      __syncthreads();
      for (int e = tid; e < Z * Z; e += nt) {                   // A2 = A V0
        const int i = e / Z, j = e % Z;
        double t = 0;
        for (int m = 0; m < Z; ++m) t += A[i][m] * V[m][j];
        A2[i][j] = t;
      }
      __syncthreads();
      for (int e = tid; e < Z * Z; e += nt) {                   // A = V0^T A2 (symmetrised: the two triangles differ by rounding)
        const int i = e / Z, j = e % Z;
        if (i <= j) {
          double t = 0;
          for (int m = 0; m < Z; ++m) t += V[m][i] * A2[m][j];
          A[i][j] = t;
          A[j][i] = t;
        }
      }
      __syncthreads();
    }
  }

  // ---- parallel cyclic Jacobi (round-robin tournament ordering: n-1 rounds of n/2 disjoint rotations).
  // Each round is two phases: (A) the n/2 rotations are computed from the current matrix, (B) every element of
  // A' = J^T A J and V' = V J is produced from the OLD matrices (double buffering), so a round costs two barriers.
  const int half = n / 2;
  Row* Acur = A; Row* Anew = A2;
  Row* Vcur = V; Row* Vnew = V2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    if (tid == 0) offmax = 0;
    __syncthreads();
    {
      double m = 0;
      for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        if (i != j) m = fmax(m, fabs(Acur[i][j]));
      }
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((tid & 31) == 0 && m > 0) atomicMax(reinterpret_cast<unsigned long long*>(&offmax), (unsigned long long)__double_as_longlong(m));
    }
    __syncthreads();
    double dmax = 0;
    for (int i = 0; i < n; ++i) dmax = fmax(dmax, fabs(Acur[i][i]));
    // eigenvalue error ~ off^2 / gap, eigenvector error ~ off / gap: 1e-11 is far below fp32 resolution for both, and the
    // quadratically convergent tail (... 1e-6, 1e-12, 1e-24) then stops one sweep (~34 us) earlier than with 1e-14
    // (a rank-deficient Gram, B < Z, keeps the tight threshold: its null-space eigenvalues enter the gradient as 1/sqrt)
    if (offmax <= (B < Z ? 1e-14 : 1e-11) * dmax || offmax == 0) break;
    for (int round = 0; round < n - 1; ++round) {
      if (tid < half) {
        // circle method: player n-1 fixed, the others rotate
        const int a0 = (tid == 0) ? n - 1 : (round + tid) % (n - 1);
        const int b0 = (tid == 0) ? round : (round - tid + (n - 1)) % (n - 1);
        const int p = min(a0, b0), q = max(a0, b0);
        const double apq = Acur[p][q];
        double c = 1.0, sn_ = 0.0;
        if (fabs(apq) > 1e-300) {
          // the rotation ANGLE is computed in fp32 with MUFU ops (it only steers convergence: an angle that is off by
          // 1e-7 leaves a 1e-7 residue that the next sweep removes), the rotation itself (c, s) is orthonormal in fp64
          const float tau = __fdividef((float)(Acur[q][q] - Acur[p][p]), (float)(2.0 * apq));
          const float tf = __fdividef(tau >= 0.f ? 1.f : -1.f, fabsf(tau) + sqrtf(1.f + tau * tau));
          const double t = (tf == tf) ? (double)tf : 0.0;
          c = rsqrt(1.0 + t * t);
          sn_ = t * c;
        }
        // new_p = c*old_p - s*old_q ; new_q = s*old_p + c*old_q
        part[p] = q; part[q] = p;
        own[p] = c; own[q] = c;
        oth[p] = -sn_; oth[q] = sn_;
      }
      __syncthreads();
      for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        const int pi = part[i], pj = part[j];
        const double oi = own[i], xi = oth[i], oj = own[j], xj = oth[j];
        Anew[i][j] = oi * (oj * Acur[i][j] + xj * Acur[i][pj]) + xi * (oj * Acur[pi][j] + xj * Acur[pi][pj]);
        Vnew[i][j] = oj * Vcur[i][j] + xj * Vcur[i][pj];
      }
      __syncthreads();
      Row* tA = Acur; Acur = Anew; Anew = tA;
      Row* tV = Vcur; Vcur = Vnew; Vnew = tV;
    }
  }
  A = Acur; V = Vcur;
  if (state) {                                                   // this call's basis is the next call's starting point
    for (int e = tid; e < Z * Z; e += nt) state[4 + e] = V[e / Z][e % Z];
    if (tid == 0) {
      unsigned long long* su = reinterpret_cast<unsigned long long*>(state);
      su[0] = CP_STATE_TAG; su[1] = (unsigned long long)Z; su[2] = calls + 1;
    }
  }
  // ---- eigenvalues, top-k selection (rank bound min(k, B, Z)), loss
  if (tid < n) {
    const double e = (tid < Z) ? A[tid][tid] : -1e300;
    ev[tid] = (e == e) ? e : -1e300;               // NaN-safe: a total order is needed for the ranking below
  }
  __syncthreads();
  if (tid < n) {
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (ev[j] > ev[tid]) || (ev[j] == ev[tid] && j < tid);
    order[rank] = tid;
  }
  __syncthreads();
  const int k = min(kloss, min(B, Z));
  if (tid == 0 && acc) {
    double s = 0;
    for (int i = 0; i < k; ++i) s += sqrt(fmax(ev[order[i]], 0.0));
    acc[ACC_KMEANS] = lmbda * s;
  }
  if (dz) {
  // S = V_k diag(0.5 / sqrt(ev)) V_k^T  -> reuse A
  __syncthreads();
  for (int e = tid; e < Z * Z; e += nt) {
    const int i = e / Z, j = e % Z;
    double t = 0;
    for (int m = 0; m < k; ++m) {
      const int col = order[m];
      const double l = ev[col];
      if (l > 0) t += V[i][col] * V[j][col] * (0.5 / sqrt(l));
    }
    A[i][j] = t;
  }
  __syncthreads();
  }
  }   // !ns_done
  if (!dz) return;
  const double sc = gcoef * lmbda * 2.0 / bsize;
  for (int b0 = 0; b0 < B; b0 += 64) {
    const int nb = min(64, B - b0);
    for (int e = tid; e < nb * Z; e += nt) zs[e / Z][e % Z] = z[(long)(b0 + e / Z) * Z + e % Z];
    __syncthreads();
    for (int e = tid; e < nb * Z; e += nt) {
      const int b = e / Z, j = e % Z;
      double t = 0;
      for (int i = 0; i < Z; ++i) t += (double)zs[b][i] * A[i][j];
      dz[(long)(b0 + b) * Z + j] = (float)(sc * t);
    }
    __syncthreads();
  }
}

#undef offmax
// -------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] (+)= sum_m X[m*ld + n]
// -------------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ X, long ld, long rows, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long per = (rows + gridDim.y - 1) / gridDim.y;
  const long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long r = r0;
  for (; r + 3 < r1; r += 4) {                      // four independent loads in flight per thread
    s0 += X[r * ld + n];
    s1 += X[(r + 1) * ld + n];
    s2 += X[(r + 2) * ld + n];
    s3 += X[(r + 3) * ld + n];
  }
  for (; r < r1; ++r) s0 += X[r * ld + n];
  atomicAdd(out + n, (s0 + s1) + (s2 + s3));
}

// feature-major time reduction: out[c*Bp + b] = sum_t X[c*ld + t*Bp + b]
__global__ void timesum_fm_kernel(const float* __restrict__ X, long ld, int T, int Bp, int C, float* __restrict__ out) {
  const long total = (long)C * Bp;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % Bp);
    const long c = idx / Bp;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += X[c * ld + (long)t * Bp + b];
    out[idx] = s;
  }
}

// row sums of a feature-major matrix (bias gradients): out[f] += sum_r X[f*ld + r]; one warp per feature
__global__ void rowsum_fm_kernel(const float* __restrict__ X, long ld, long ncols, int nfeat, float* __restrict__ out) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= nfeat) return;
  const long per = ((ncols + gridDim.y - 1) / gridDim.y + 3) & ~3L;       // column split (gridDim.y) for parallelism
  const long c0 = blockIdx.y * per, c1 = min(ncols, c0 + per);
  const float* row = X + (long)f * ld;
  float s = 0.f;
  if ((((uintptr_t)(row + c0)) & 15) == 0) {
    long r = c0 + (threadIdx.x & 31) * 4;
    for (; r + 3 < c1; r += 128) {
      const float4 v = *reinterpret_cast<const float4*>(row + r);
      s += (v.x + v.y) + (v.z + v.w);
    }
    for (; r < c1; ++r) s += row[r];
  } else {
    for (long r = c0 + (threadIdx.x & 31); r < c1; r += 32) s += row[r];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + f, s);
}

// feature-major [H][ld] -> row-major dst[b*dst_ld + u] (b < B)
__global__ void fm_to_rows_kernel(const float* __restrict__ src, long ld, int H, int B, float* __restrict__ dst, long dst_ld) {
  __shared__ float tile[32][33];
  const int u0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
  for (int i = ty; i < 32; i += 8) {
    const int u = u0 + i, b = b0 + tx;
    tile[i][tx] = (u < H && b < B) ? src[(long)u * ld + b] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int b = b0 + i, u = u0 + tx;
    if (b < B && u < H) dst[(long)b * dst_ld + u] = tile[tx][i];
  }
}

// sum of the dh pieces left by the last BPTT step (feature-major parts [p][H][B_pad]) -> dh0 written in the UNPADDED
// [D][B][H] order, i.e. the flat buffer that the reference's hidden.view(2,B,H) aliases (rnn_model.py:104)
__global__ void parts_reduce_kernel(const float* __restrict__ parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H,
                                    float* __restrict__ out) {
  const long total = (long)D * H * B_pad;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % B_pad);
    const long r = idx / B_pad;
    const int u = (int)(r % H), d = (int)(r / H);
    if (b >= B) continue;
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += parts[d * dir_stride + ((long)p * H + u) * B_pad + b];
    out[((long)d * B + b) * H + u] = s;
  }
}

// same reduction, 8 consecutive units per thread, that also emits the P16 operand of the latent_to_hidden backward GEMM: the
// flat [D][B][H] result viewed as [B rows][D*H] (the inverse of the reference's view quirk), rows >= B zero
__global__ void parts_reduce_pack_kernel(const float* __restrict__ parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H,
                                         float* __restrict__ out, __nv_bfloat16* __restrict__ out_p) {
  const int K = D * H, k8n = K / 8, nkc = (K + KCHUNK - 1) / KCHUNK;
  const long total = (long)B_pad * k8n;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % B_pad), col = (int)(idx / B_pad) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (r < B) {
      const long f = (long)r * K + col;
      const int d = (int)(f / ((long)B * H));
      const long rem = f % ((long)B * H);
      const int b = (int)(rem / H), u = (int)(rem % H);
      const float* p0 = parts + d * dir_stride + (long)u * B_pad + b;
      for (int p = 0; p < n_parts; ++p) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += p0[((long)p * H + i) * B_pad];
      }
      *reinterpret_cast<float4*>(out + f) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + f + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    __nv_bfloat16* tile = out_p + ((size_t)(r / 128) * nkc + col / KCHUNK) * p16_tile_elems(128);
    const int off = p16_in_tile(r % 128, col % KCHUNK);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + 128 * KCHUNK + off) = lo;
  }
}

// -------------------------------------------------------------------------------------------------
// AMSGrad Adam, flat multi-tensor (torch.optim.Adam(amsgrad=True), vame/model/rnn_vae.py:332,143)
//   36 B/param of HBM traffic (p, g, m, v, vmax read; p, m, v, vmax written); grad_scale folds the DP 1/world average.
// -------------------------------------------------------------------------------------------------
// prepare: increments the device step counter and derives the bias corrections (1 thread)
__global__ void adam_prepare_kernel(int* __restrict__ step_dev, float* __restrict__ sc, float lr, const float* __restrict__ hyper,
                                    float beta1, float beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int step = *step_dev + 1;
    *step_dev = step;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const double l = hyper ? (double)hyper[HY_LR] : (double)lr;
    sc[0] = (float)(l / bc1);
    sc[1] = (float)sqrt(bc2);
  }
}
__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, float* __restrict__ vmax, long n, const float* __restrict__ sc,
                                    float beta1, float beta2, float eps, float grad_scale) {
  const long n4 = n / 4;
  const float step_size = sc[0], bc2_sqrt = sc[1];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i], X = reinterpret_cast<float4*>(vmax)[i];
    float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &Vv.x; float* xx = &X.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * grad_scale;
      mm[j] = beta1 * mm[j] + (1.f - beta1) * gr;
      vv[j] = beta2 * vv[j] + (1.f - beta2) * gr * gr;
      xx[j] = fmaxf(xx[j], vv[j]);
      const float denom = sqrtf(xx[j]) / bc2_sqrt + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = Vv;
    reinterpret_cast<float4*>(vmax)[i] = X;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = n4 * 4 + threadIdx.x;
    const float gr = g[i] * grad_scale;
    m[i] = beta1 * m[i] + (1.f - beta1) * gr;
    v[i] = beta2 * v[i] + (1.f - beta2) * gr * gr;
    vmax[i] = fmaxf(vmax[i], v[i]);
    p[i] -= step_size * (m[i] / (sqrtf(vmax[i]) / bc2_sqrt + eps));
  }
}

// losses: acc (double[8]) -> out (float[8]) following rnn_vae.py:124-129
__global__ void finalize_losses_kernel(const double* __restrict__ acc, float* __restrict__ out, double rec_div, double fut_div,
                                       double kl_n, double beta, double klw, const float* __restrict__ hyper, int future) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (hyper) {
      beta = (double)hyper[HY_BETA];
      klw = (double)hyper[HY_KLW];
    }
    const double rec = acc[ACC_REC] / rec_div;
    const double fut = future ? acc[ACC_FUT] / fut_div : 0.0;
    const double kl = -0.5 * acc[ACC_KL] / kl_n;
    const double km = acc[ACC_KMEANS];
    out[0] = (float)rec; out[1] = (float)fut; out[2] = (float)kl; out[3] = (float)km;
    out[4] = (float)(rec + fut + beta * klw * kl + klw * km);
  }
}

// ---- launchers ----------------------------------------------------------------------------------
static inline unsigned grid_for(long total, int threads, int cap = 148 * 8) {
  long b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}
void launch_lambda_fwd(const float* lin, long ldl, const float* eps, int B, int Z, int softplus, float* z, float* mu, float* logvar,
                       double* acc, cudaStream_t st) {
  count_launch();
  lambda_fwd_kernel<<<grid_for((long)B * Z, 256), 256, 0, st>>>(lin, ldl, eps, B, Z, softplus, z, mu, logvar, acc);
}
void launch_lambda_fwd_p16(const float* lin, long ldl, const float* eps, int B, int B_pad, int Z, int softplus, float* z, float* mu,
                           float* logvar, double* acc, void* z_p, cudaStream_t st) {
  count_launch();
  const long total = (long)B_pad * ((Z + KCHUNK - 1) / KCHUNK) * 8;
  lambda_fwd_p16_kernel<<<grid_for(total, 128), 128, 0, st>>>(lin, ldl, eps, B, B_pad, Z, softplus, z, mu, logvar, acc,
                                                              (__nv_bfloat16*)z_p);
}
void launch_lambda_bwd_p16(const LambdaBwdArgs& a, void* dlin_p, cudaStream_t st) {
  count_launch();
  const long total = (long)a.B_pad * ((2 * a.Z + KCHUNK - 1) / KCHUNK) * 8;
  lambda_bwd_p16_kernel<<<grid_for(total, 128), 128, 0, st>>>(a, (__nv_bfloat16*)dlin_p);
}
void launch_mse_p16(const float* pred, long ldp, const float* target, int rows, int B, int B_pad, int F, float gscale, float* dpred,
                    double* acc, int slot, void* dpred_p, cudaStream_t st) {
  count_launch();
  const long total = (long)rows * ((F + KCHUNK - 1) / KCHUNK) * 8;
  mse_p16_kernel<<<grid_for(total, 256, 296), 256, 0, st>>>(pred, ldp, target, rows, B, B_pad, F, gscale, dpred, acc, slot,
                                                           (__nv_bfloat16*)dpred_p);
}
void launch_lambda_bwd(const LambdaBwdArgs& a, cudaStream_t st) {
  count_launch();
  lambda_bwd_kernel<<<grid_for((long)a.B_pad * a.Z, 256), 256, 0, st>>>(a);
}
void launch_mse(const float* pred, long ldp, const float* target, int rows, int B, int B_pad, int F, float gscale, float* dpred,
                double* acc, int slot, cudaStream_t st) {
  count_launch();
  mse_kernel<<<grid_for((long)rows * F, 256, 296), 256, 0, st>>>(pred, ldp, target, rows, B, B_pad, F, gscale, dpred, acc, slot);   // grid-stride
}
void launch_cluster_prior(const float* z, int B, int Z, int kloss, double lmbda, double bsize, double gcoef, const float* hyper,
                          float* dz, double* acc, cudaStream_t st, double* state) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(cluster_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CP_SMEM);
    attr = true;
  }
  count_launch();
  cluster_prior_kernel<<<1, 1024, CP_SMEM, st>>>(z, B, Z, kloss, lmbda, bsize, gcoef, hyper, dz, acc, state);
}
void launch_colsum(const float* X, long ld, long rows, int N, float* out, cudaStream_t st) {
  int ysplit = (int)min((long)256, (rows + 31) / 32);
  if (ysplit < 1) ysplit = 1;
  count_launch();
  colsum_kernel<<<dim3((N + 127) / 128, ysplit), 128, 0, st>>>(X, ld, rows, N, out);
}
void launch_timesum_fm(const float* X, long ld, int T, int Bp, int C, float* out, cudaStream_t st) {
  count_launch();
  timesum_fm_kernel<<<grid_for((long)C * Bp, 256), 256, 0, st>>>(X, ld, T, Bp, C, out);
}
void launch_rowsum_fm(const float* X, long ld, long ncols, int nfeat, float* out, cudaStream_t st) {
  count_launch();
  int ysplit = (int)((ncols + 2047) / 2048);
  if (ysplit > 8) ysplit = 8;
  if (ysplit < 1) ysplit = 1;
  rowsum_fm_kernel<<<dim3((nfeat + 7) / 8, ysplit), 256, 0, st>>>(X, ld, ncols, nfeat, out);
}
void launch_fm_to_rows(const float* src, long ld, int H, int B, float* dst, long dst_ld, cudaStream_t st) {
  count_launch();
  fm_to_rows_kernel<<<dim3((B + 31) / 32, (H + 31) / 32), 256, 0, st>>>(src, ld, H, B, dst, dst_ld);
}
void launch_parts_reduce(const float* parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H, float* out, cudaStream_t st) {
  count_launch();
  parts_reduce_kernel<<<grid_for((long)D * B_pad * H, 256), 256, 0, st>>>(parts, n_parts, dir_stride, D, B, B_pad, H, out);
}
void launch_parts_reduce_pack(const float* parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H, float* out, void* out_p,
                              cudaStream_t st) {
  count_launch();
  parts_reduce_pack_kernel<<<grid_for((long)B_pad * (D * H / 8), 256), 256, 0, st>>>(parts, n_parts, dir_stride, D, B, B_pad, H, out,
                                                                                  (__nv_bfloat16*)out_p);
}
void launch_adam(float* p, const float* g, float* m, float* v, float* vmax, long n, float lr, const float* hyper, int* step_dev,
                 float* scratch2, float b1, float b2, float eps, float grad_scale, cudaStream_t st) {
  count_launch();
  adam_prepare_kernel<<<1, 32, 0, st>>>(step_dev, scratch2, lr, hyper, b1, b2);
  count_launch();
  adam_amsgrad_kernel<<<grid_for(n / 4 + 1, 256, 148 * 4), 256, 0, st>>>(p, g, m, v, vmax, n, scratch2, b1, b2, eps, grad_scale);
}
void launch_adam_prepare(float lr, const float* hyper, int* step_dev, float* scratch2, float b1, float b2, cudaStream_t st) {
  count_launch();
  adam_prepare_kernel<<<1, 32, 0, st>>>(step_dev, scratch2, lr, hyper, b1, b2);
}
void launch_adam_apply(float* p, const float* g, float* m, float* v, float* vmax, long n, const float* scratch2, float b1, float b2,
                       float eps, float grad_scale, cudaStream_t st) {
  count_launch();
  adam_amsgrad_kernel<<<grid_for(n / 4 + 1, 256, 148 * 4), 256, 0, st>>>(p, g, m, v, vmax, n, scratch2, b1, b2, eps, grad_scale);
}
void launch_finalize_losses(const double* acc, float* out, double rec_div, double fut_div, double kl_n, double beta, double klw,
                            const float* hyper, int future, cudaStream_t st) {
  count_launch();
  finalize_losses_kernel<<<1, 32, 0, st>>>(acc, out, rec_div, fut_div, kl_n, beta, klw, hyper, future);
}

}  // namespace vb
