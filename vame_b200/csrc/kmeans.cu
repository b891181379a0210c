// k-means on the latent vectors: replaces the sklearn.cluster.KMeans(init='k-means++') calls of
// vame/analysis/pose_segmentation.py:141-143 (same_parameterization) and :183-185 (individual_parameterization).
// scikit-learn is a third-party dependency of the reference (not vendored); the algorithm restated here is the published
// one: greedy k-means++ seeding (Arthur & Vassilvitskii 2007, 2 + ln k local trials) and Lloyd iterations on the mean-
// centred data with tol scaled by the mean feature variance, strict-convergence / centre-shift stopping and a final
// consistent E-step (oracle/kmeans_numpy.py states the same steps on the CPU and is pinned to sklearn's outputs).
//
// All kernels are HBM-bound passes over X [n, dim] fp32 (dim <= 64, k <= 128): a block stages 256 points in shared memory
// with coalesced loads, every thread then owns one point; per-block partial sums (fp32 in shared memory) are flushed to fp64
// global accumulators, so one Lloyd iteration reads X exactly once (4*dim bytes per point) and writes 4 bytes per point.
#include <cub/device/device_scan.cuh>
#include <stdlib.h>

#include "../../include/vame_b200.h"
#include "api_common.h"
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int KM_T = 256;          // threads = points per block iteration
constexpr int KM_MAXD = 64, KM_MAXK = 128;

struct KmWs {                      // workspace carve (all 256-byte aligned)
  double* colsum;                  // [dim]       column sums of X
  double* colsq;                   // [dim]       column sums of squares
  float* mean;                     // [dim]
  double* sums;                    // [k, dim]    M-step accumulators
  int* counts;                     // [k]
  double* inertia;                 // [1]
  int* changed;                    // [1]         labels that changed in the last E-step
  double* shift;                   // [1]         squared centre shift of the last M-step
  float* cen[2];                   // [k, dim]    centres (mean-centred space), ping-pong
  double* scan;                    // [n]         k-means++: inclusive cumsum of the closest squared distances
  void* cub_tmp; size_t cub_bytes;
  size_t bytes;
};
static KmWs km_carve(long n, int dim, int k, void* base) {
  KmWs w{};
  size_t off = 0;
  auto take = [&](size_t b) {
    void* p = base ? (char*)base + off : nullptr;
    off += (b + 255) & ~(size_t)255;
    return p;
  };
  w.colsum = (double*)take(sizeof(double) * dim);
  w.colsq = (double*)take(sizeof(double) * dim);
  w.mean = (float*)take(sizeof(float) * dim);
  w.sums = (double*)take(sizeof(double) * k * dim);
  w.counts = (int*)take(sizeof(int) * k);
  w.inertia = (double*)take(sizeof(double));
  w.changed = (int*)take(sizeof(int));
  w.shift = (double*)take(sizeof(double));
  w.cen[0] = (float*)take(sizeof(float) * k * dim);
  w.cen[1] = (float*)take(sizeof(float) * k * dim);
  w.scan = (double*)take(sizeof(double) * n);
  size_t cb = 0;
  cub::DeviceScan::InclusiveSum((void*)nullptr, cb, (const double*)nullptr, (double*)nullptr, (int)n);
  w.cub_bytes = cb;
  w.cub_tmp = take(cb);
  w.bytes = off;
  return w;
}

// stage `cnt` points starting at row p0 into xs[KM_T][dim + 1] (coalesced global reads, conflict-free row reads afterwards)
__device__ __forceinline__ void km_stage(const float* __restrict__ x, long p0, int cnt, int dim, const float* __restrict__ mean,
                                         float* xs) {
  const int total = cnt * dim;
  const float* src = x + p0 * dim;
  for (int i = threadIdx.x; i < total; i += KM_T) {
    const int r = i / dim, c = i - r * dim;
    xs[r * (dim + 1) + c] = src[i] - (mean ? mean[c] : 0.f);
  }
}

// column sums / sums of squares (fp64) -> mean and the variance used for sklearn's tol scaling
__global__ void __launch_bounds__(KM_T) km_stats_kernel(const float* __restrict__ x, long n, int dim, double* colsum, double* colsq) {
  extern __shared__ float sm[];
  float* xs = sm;                               // [KM_T][dim + 1]
  __shared__ double s1[KM_MAXD], s2[KM_MAXD];
  if (threadIdx.x < dim) s1[threadIdx.x] = s2[threadIdx.x] = 0.0;
  const int G = KM_T / dim;                     // row groups: thread t sums column t % dim over rows t / dim, t / dim + G, ...
  const int c = threadIdx.x % dim, g = threadIdx.x / dim;
  double a = 0.0, b = 0.0;
  const long per = ((n + gridDim.x - 1) / gridDim.x + KM_T - 1) / KM_T * KM_T;
  const long b0 = (long)blockIdx.x * per, b1 = min(n, b0 + per);
  for (long p0 = b0; p0 < b1; p0 += KM_T) {
    const int c_rows = (int)min((long)KM_T, b1 - p0);
    __syncthreads();
    km_stage(x, p0, c_rows, dim, nullptr, xs);
    __syncthreads();
    if (g < G) {
      for (int r = g; r < c_rows; r += G) {
        const double v = xs[r * (dim + 1) + c];
        a += v;
        b += v * v;
      }
    }
  }
  __syncthreads();
  if (g < G) {
    atomicAdd(&s1[c], a);
    atomicAdd(&s2[c], b);
  }
  __syncthreads();
  if (threadIdx.x < dim) {
    atomicAdd(colsum + threadIdx.x, s1[threadIdx.x]);
    atomicAdd(colsq + threadIdx.x, s2[threadIdx.x]);
  }
}
__global__ void km_mean_kernel(const double* colsum, long n, int dim, float* mean) {
  if (threadIdx.x < dim) mean[threadIdx.x] = (float)(colsum[threadIdx.x] / (double)n);
}
// centres given in the original space -> mean-centred working copy (and back)
__global__ void km_shift_centers_kernel(const float* in, const float* mean, int k, int dim, float sign, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k * dim) out[i] = in[i] + sign * mean[i % dim];
}

// E-step (+ optional M-step accumulation): label = argmin_c ||x - c||^2 (ties -> lowest index)
__global__ void __launch_bounds__(KM_T) km_assign_kernel(const float* __restrict__ x, long n, int dim, int k,
                                                         const float* __restrict__ mean, const float* __restrict__ centers,
                                                         int* __restrict__ labels, int accumulate, double* __restrict__ sums,
                                                         int* __restrict__ counts, double* __restrict__ inertia,
                                                         int* __restrict__ changed) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* acc = reinterpret_cast<double*>(smraw);          // [k][dim] block partial sums (fp64: order-independent after rounding)
  float* cs = reinterpret_cast<float*>(acc + k * dim);     // [k][dim]
  float* xs = cs + k * dim;                                // [KM_T][dim + 1]
  int* cnt = reinterpret_cast<int*>(xs + KM_T * (dim + 1));   // [k]
  for (int i = threadIdx.x; i < k * dim; i += KM_T) {
    cs[i] = centers[i];
    acc[i] = 0.0;
  }
  for (int i = threadIdx.x; i < k; i += KM_T) cnt[i] = 0;
  double my_inertia = 0.0;
  int my_changed = 0;
  const long per = ((n + gridDim.x - 1) / gridDim.x + KM_T - 1) / KM_T * KM_T;     // rows per block, multiple of KM_T
  const long b0 = (long)blockIdx.x * per, b1 = min(n, b0 + per);
  for (long p0 = b0; p0 < b1; p0 += KM_T) {
    const int c_rows = (int)min((long)KM_T, b1 - p0);
    __syncthreads();
    km_stage(x, p0, c_rows, dim, mean, xs);
    __syncthreads();
    if (threadIdx.x < c_rows) {
      const float* xr = xs + threadIdx.x * (dim + 1);
      // fp64 distances: the assignment follows exact arithmetic on the fp32 data / centres, so that the Lloyd trajectory does
      // not depend on summation order (a flipped borderline point can send an overlapping mixture to another optimum)
      double best = 1.0e300;
      int bi = 0;
      for (int c = 0; c < k; ++c) {
        const float* cr = cs + c * dim;
        double d = 0.0;
        for (int j = 0; j < dim; ++j) {
          const double t = (double)xr[j] - (double)cr[j];
          d = fma(t, t, d);
        }
        if (d < best) {
          best = d;
          bi = c;
        }
      }
      const long p = p0 + threadIdx.x;
      if (labels[p] != bi) ++my_changed;
      labels[p] = bi;
      my_inertia += best;
      if (accumulate) {
        atomicAdd(&cnt[bi], 1);
        double* ar = acc + bi * dim;
        for (int j = 0; j < dim; ++j) atomicAdd(ar + j, (double)xr[j]);
      }
    }
  }
  __syncthreads();
  if (accumulate) {
    for (int i = threadIdx.x; i < k * dim; i += KM_T)
      if (acc[i] != 0.0) atomicAdd(sums + i, acc[i]);
    for (int i = threadIdx.x; i < k; i += KM_T)
      if (cnt[i]) atomicAdd(counts + i, cnt[i]);
  }
  my_inertia = warp_sum_d(my_inertia);
  for (int o = 16; o > 0; o >>= 1) my_changed += __shfl_xor_sync(0xffffffffu, my_changed, o);
  if ((threadIdx.x & 31) == 0) {
    if (my_inertia != 0.0) atomicAdd(inertia, my_inertia);
    if (my_changed) atomicAdd(changed, my_changed);
  }
}

// M-step finalisation (one block): new centres, squared shift; empty clusters keep their centre
__global__ void km_update_kernel(const float* __restrict__ old_c, const double* __restrict__ sums, const int* __restrict__ counts, int k,
                                 int dim, float* __restrict__ new_c, double* __restrict__ shift) {
  __shared__ double sh;
  if (threadIdx.x == 0) sh = 0.0;
  __syncthreads();
  double s = 0.0;
  for (int i = threadIdx.x; i < k * dim; i += blockDim.x) {
    const int c = i / dim;
    const float o = old_c[i];
    const float nv = counts[c] > 0 ? (float)(sums[i] / (double)counts[c]) : o;
    new_c[i] = nv;
    const double df = (double)nv - (double)o;
    s += df * df;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sh, s);
  __syncthreads();
  if (threadIdx.x == 0) *shift = sh;
}

// k-means++ helpers ----------------------------------------------------------------------------------
// newmin[j][i] = min(closest[i], ||x_i - x_cand_j||^2) ; pot[j] = sum_i newmin[j][i]   (closest == nullptr: no min)
__global__ void __launch_bounds__(KM_T) km_cand_dist_kernel(const float* __restrict__ x, long n, int dim, const long* __restrict__ cand,
                                                            int m, const float* __restrict__ closest, float* __restrict__ newmin,
                                                            double* __restrict__ pot) {
  extern __shared__ float sm[];
  float* cs = sm;                               // [m][dim]
  float* xs = cs + m * dim;                     // [KM_T][dim + 1]
  for (int i = threadIdx.x; i < m * dim; i += KM_T) cs[i] = x[cand[i / dim] * dim + i % dim];
  double my[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) my[j] = 0.0;
  const long per = ((n + gridDim.x - 1) / gridDim.x + KM_T - 1) / KM_T * KM_T;
  const long b0 = (long)blockIdx.x * per, b1 = min(n, b0 + per);
  for (long p0 = b0; p0 < b1; p0 += KM_T) {
    const int c_rows = (int)min((long)KM_T, b1 - p0);
    __syncthreads();
    km_stage(x, p0, c_rows, dim, nullptr, xs);
    __syncthreads();
    if (threadIdx.x < c_rows) {
      const float* xr = xs + threadIdx.x * (dim + 1);
      const long p = p0 + threadIdx.x;
      const float cl = closest ? closest[p] : 3.4e38f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < m) {
          const float* cr = cs + j * dim;
          double dd = 0.0;
          for (int q = 0; q < dim; ++q) {
            const double t = (double)xr[q] - (double)cr[q];
            dd = fma(t, t, dd);
          }
          const float d = fminf((float)dd, cl);
          newmin[(long)j * n + p] = d;
          my[j] += (double)d;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < m) {
      const double s = warp_sum_d(my[j]);
      if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(pot + j, s);
    }
  }
}
__global__ void km_to_double_kernel(const float* in, long n, double* out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
// out[j] = first index with cum[idx] >= vals[j]  (numpy.searchsorted side='left'), clipped to n - 1
__global__ void km_searchsorted_kernel(const double* __restrict__ cum, long n, const double* __restrict__ vals, int m, long* out) {
  const int j = threadIdx.x;
  if (j >= m) return;
  const double v = vals[j];
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (cum[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  out[j] = lo < n ? lo : n - 1;
}

static inline unsigned km_grid(long n) {
  long b = (n + KM_T - 1) / KM_T;
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace vb

using namespace vb;

extern "C" {

size_t vame_kmeans_workspace_bytes(long n, int dim, int k) {
  if (n <= 0 || dim <= 0 || k <= 0) return 0;
  return km_carve(n, dim, k, nullptr).bytes;
}

/* Lloyd iterations from given initial centres.  Blocking call (reads the convergence status once per iteration). */
int vame_kmeans_lloyd(const float* x, long n, int dim, int k, const float* centers_init, int max_iter, float tol, int* labels,
                      float* centers_out, double* inertia_out, int* n_iter_out, void* ws, size_t ws_bytes, void* stream) {
  VB_REQUIRE(x && centers_init && labels && centers_out && ws, "vame_kmeans_lloyd: null pointer");
  VB_REQUIRE(n > 0 && dim > 0 && dim <= KM_MAXD && k > 0 && k <= KM_MAXK, "vame_kmeans_lloyd: need 0 < dim <= 64, 0 < k <= 128");
  VB_REQUIRE(n < 2147483647L, "vame_kmeans_lloyd: n must fit int32");
  KmWs w = km_carve(n, dim, k, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_kmeans_lloyd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(double) * k * dim + sizeof(float) * ((size_t)k * dim + (size_t)KM_T * (dim + 1)) + sizeof(int) * k;
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(km_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  // mean / variance of the data (the tolerance is relative to the mean feature variance)
  cudaMemsetAsync(w.colsum, 0, sizeof(double) * dim, st);
  cudaMemsetAsync(w.colsq, 0, sizeof(double) * dim, st);
  count_launch(3);
  km_stats_kernel<<<km_grid(n), KM_T, sizeof(float) * KM_T * (dim + 1), st>>>(x, n, dim, w.colsum, w.colsq);
  km_mean_kernel<<<1, 64, 0, st>>>(w.colsum, n, dim, w.mean);
  km_shift_centers_kernel<<<(k * dim + 255) / 256, 256, 0, st>>>(centers_init, w.mean, k, dim, -1.f, w.cen[0]);
  double h_sum[KM_MAXD], h_sq[KM_MAXD];
  cudaMemcpyAsync(h_sum, w.colsum, sizeof(double) * dim, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(h_sq, w.colsq, sizeof(double) * dim, cudaMemcpyDeviceToHost, st);
  cudaMemsetAsync(labels, 0xff, sizeof(int) * n, st);                  // -1: every label "changes" in the first E-step
  cudaStreamSynchronize(st);
  double var_mean = 0.0;
  for (int j = 0; j < dim; ++j) {
    const double mu = h_sum[j] / (double)n;
    var_mean += h_sq[j] / (double)n - mu * mu;
  }
  var_mean /= dim;
  const double tol_abs = (double)tol * var_mean;

  int cur = 0, it = 0;
  for (it = 0; it < max_iter; ++it) {
    cudaMemsetAsync(w.sums, 0, sizeof(double) * k * dim, st);
    cudaMemsetAsync(w.counts, 0, sizeof(int) * k, st);
    cudaMemsetAsync(w.inertia, 0, sizeof(double), st);
    cudaMemsetAsync(w.changed, 0, sizeof(int), st);
    count_launch(2);
    km_assign_kernel<<<km_grid(n), KM_T, smem, st>>>(x, n, dim, k, w.mean, w.cen[cur], labels, 1, w.sums, w.counts, w.inertia, w.changed);
    km_update_kernel<<<1, 256, 0, st>>>(w.cen[cur], w.sums, w.counts, k, dim, w.cen[cur ^ 1], w.shift);
    int h_changed = 0;
    double h_shift = 0.0;
    cudaMemcpyAsync(&h_changed, w.changed, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_shift, w.shift, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cur ^= 1;                                                          // the new centres become current
    if (getenv("VAME_B200_KM_DEBUG")) fprintf(stderr, "[kmeans] it %d changed %d shift %.6e tol %.6e var %.6e\n", it, h_changed, h_shift, tol_abs, var_mean);
    if (h_changed == 0) {                                              // strict convergence: labels unchanged
      ++it;
      break;
    }
    if (h_shift <= tol_abs) {
      ++it;
      break;
    }
  }
  // labels consistent with the final centres (+ the inertia of the result)
  cudaMemsetAsync(w.inertia, 0, sizeof(double), st);
  cudaMemsetAsync(w.changed, 0, sizeof(int), st);
  count_launch(2);
  km_assign_kernel<<<km_grid(n), KM_T, smem, st>>>(x, n, dim, k, w.mean, w.cen[cur], labels, 0, w.sums, w.counts, w.inertia, w.changed);
  km_shift_centers_kernel<<<(k * dim + 255) / 256, 256, 0, st>>>(w.cen[cur], w.mean, k, dim, 1.f, centers_out);
  if (inertia_out) cudaMemcpyAsync(inertia_out, w.inertia, sizeof(double), cudaMemcpyDeviceToDevice, st);
  if (n_iter_out) *n_iter_out = it;
  return check_launch("vame_kmeans_lloyd");
}

/* KMeans.predict: labels (and optionally the inertia, device double[1]) for given centres */
int vame_kmeans_assign(const float* x, long n, int dim, int k, const float* centers, int* labels, double* inertia_out, void* ws,
                       size_t ws_bytes, void* stream) {
  VB_REQUIRE(x && centers && labels && ws, "vame_kmeans_assign: null pointer");
  VB_REQUIRE(n > 0 && dim > 0 && dim <= KM_MAXD && k > 0 && k <= KM_MAXK, "vame_kmeans_assign: need 0 < dim <= 64, 0 < k <= 128");
  KmWs w = km_carve(n, dim, k, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_kmeans_assign: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(double) * k * dim + sizeof(float) * ((size_t)k * dim + (size_t)KM_T * (dim + 1)) + sizeof(int) * k;
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(km_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  cudaMemsetAsync(w.inertia, 0, sizeof(double), st);
  cudaMemsetAsync(w.changed, 0, sizeof(int), st);
  count_launch();
  km_assign_kernel<<<km_grid(n), KM_T, smem, st>>>(x, n, dim, k, nullptr, centers, labels, 0, w.sums, w.counts, w.inertia, w.changed);
  if (inertia_out) cudaMemcpyAsync(inertia_out, w.inertia, sizeof(double), cudaMemcpyDeviceToDevice, st);
  return check_launch("vame_kmeans_assign");
}

/* One greedy k-means++ round: for the m (<= 8) candidate rows cand[] (device int64), newmin[j] = min(closest, d^2(x, x_cand_j))
 * (closest == NULL for the first centre) and pot[j] (device double[m], zeroed here) = sum of newmin[j]. */
int vame_kmeans_candidates(const float* x, long n, int dim, const long* cand, int m, const float* closest, float* newmin, double* pot,
                           void* stream) {
  VB_REQUIRE(x && cand && newmin && pot, "vame_kmeans_candidates: null pointer");
  VB_REQUIRE(n > 0 && dim > 0 && dim <= KM_MAXD && m > 0 && m <= 8, "vame_kmeans_candidates: need dim <= 64, m <= 8");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * ((size_t)m * dim + (size_t)KM_T * (dim + 1));
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(km_cand_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  cudaMemsetAsync(pot, 0, sizeof(double) * m, st);
  count_launch();
  km_cand_dist_kernel<<<km_grid(n), KM_T, smem, st>>>(x, n, dim, cand, m, closest, newmin, pot);
  return check_launch("vame_kmeans_candidates");
}

/* Sampling step of k-means++: idx[j] = searchsorted(cumsum(closest) in fp64, vals[j]) clipped to n-1; vals device double[m] */
int vame_kmeans_sample(const float* closest, long n, const double* vals, int m, long* idx, void* ws, size_t ws_bytes, void* stream) {
  VB_REQUIRE(closest && vals && idx && ws, "vame_kmeans_sample: null pointer");
  VB_REQUIRE(n > 0 && n < 2147483647L && m > 0 && m <= 32, "vame_kmeans_sample: bad sizes");
  KmWs w = km_carve(n, 1, 1, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_kmeans_sample: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  count_launch(3);
  // (the scan buffer doubles as the fp64 copy of the input: in-place inclusive sum)
  km_to_double_kernel<<<km_grid(n), KM_T, 0, st>>>(closest, n, w.scan);
  size_t cb = w.cub_bytes;
  cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.scan, w.scan, (int)n, st);
  km_searchsorted_kernel<<<1, 32, 0, st>>>(w.scan, n, vals, m, idx);
  return check_launch("vame_kmeans_sample");
}

}  // extern "C"
