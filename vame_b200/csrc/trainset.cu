// Training-set preparation: replaces the per-file z-scoring / outlier removal / interpolation and the
// Savitzky-Golay smoothing of vame/model/create_training.py:94-264 (traindata_aligned, traindata_fixed), whose outlier pass
// is a Python double loop over every frame and marker.  Everything is float64, like the reference (numpy), and streaming:
// the series stays in the reference's (F, N) layout (marker-major, time contiguous), every kernel is one or two passes over it.
//   z-score  (:112-114, :210-212)  global mean / population std over all entries (two-pass, fp64 block partials)
//   IQR      (:126, :218)          scipy.stats.iqr = numpy 'linear' percentiles 75 - 25 of all entries: device radix sort
//   outliers (:128-135, :220-227)  |x| > iqr_factor * iqr -> NaN
//   fixed    (:231)                every FRAME interpolated over the marker index (np.interp: linear, clamped at the ends)
//   aligned  (:137)                interpol() on the 2-D array: a NaN of marker f becomes the LAST valid time sample of marker f
//                                  (np.interp on repeated sample points, see oracle/trainset_numpy.py); markers without any
//                                  valid entry are interpolated between their neighbours' last / first valid samples
//   savgol   (:175-178)            FIR along time + the least-squares polynomial edges of scipy's mode='interp'; the tables
//                                  (window coefficients, edge matrices) are tiny and computed by the host mirror
#include <cub/device/device_radix_sort.cuh>
#include <math.h>

#include "../../include/vame_b200.h"
#include "api_common.h"
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int TS_T = 256;
constexpr int TS_MAXF = 128;

struct TsWs {
  double* acc;        // [8]: 0 sum, 1 sum of squared deviations, 2 iqr, 3 #NaN after the outlier pass, 4 #entries left NaN
  double* first;      // [F] first valid sample per marker (aligned path)
  double* last;       // [F] last valid sample
  double* fill;       // [F] replacement value per marker
  int* has;           // [F]
  double* keys[2];    // [F * N] sort buffers
  void* cub_tmp; size_t cub_bytes;
  size_t bytes;
};
static TsWs ts_carve(long n_frames, int F, void* base) {
  TsWs w{};
  size_t off = 0;
  auto take = [&](size_t b) {
    void* p = base ? (char*)base + off : nullptr;
    off += (b + 255) & ~(size_t)255;
    return p;
  };
  const size_t n = (size_t)n_frames * F;
  w.acc = (double*)take(8 * sizeof(double));
  w.first = (double*)take(sizeof(double) * F);
  w.last = (double*)take(sizeof(double) * F);
  w.fill = (double*)take(sizeof(double) * F);
  w.has = (int*)take(sizeof(int) * F);
  w.keys[0] = (double*)take(sizeof(double) * n);
  w.keys[1] = (double*)take(sizeof(double) * n);
  size_t cb = 0;
  cub::DeviceRadixSort::SortKeys((void*)nullptr, cb, (const double*)nullptr, (double*)nullptr, (int)n);
  w.cub_bytes = cb;
  w.cub_tmp = take(cb);
  w.bytes = off;
  return w;
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double s[TS_T / 32];
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < TS_T / 32; ++i) t += s[i];
  __syncthreads();
  return t;       // valid in thread 0
}

// pass 1: sum; pass 2 (mean given through acc[0] / n): sum of squared deviations
__global__ void __launch_bounds__(TS_T) ts_moment_kernel(const double* __restrict__ x, long n, int pass, double* __restrict__ acc) {
  const double mean = pass ? acc[0] / (double)n : 0.0;
  double p = 0;
  for (long i = blockIdx.x * (long)TS_T + threadIdx.x; i < n; i += (long)gridDim.x * TS_T) {
    const double d = x[i] - mean;
    p += pass ? d * d : d;
  }
  const double t = block_sum(p);
  if (threadIdx.x == 0) atomicAdd(acc + pass, t);
}
__global__ void __launch_bounds__(TS_T) ts_zscore_kernel(const double* __restrict__ x, long n, const double* __restrict__ acc,
                                                         double* __restrict__ out, double* __restrict__ keys) {
  const double mean = acc[0] / (double)n, sd = sqrt(acc[1] / (double)n);
  for (long i = blockIdx.x * (long)TS_T + threadIdx.x; i < n; i += (long)gridDim.x * TS_T) {
    const double z = (x[i] - mean) / sd;
    out[i] = z;
    if (keys) keys[i] = z;
  }
}
// numpy's 'linear' percentile incl. its lerp form (a + (b-a) t, and b - (b-a)(1-t) where t >= 0.5)
__device__ double ts_percentile(const double* __restrict__ s, long n, double q) {
  const double v = (double)(n - 1) * q;
  const long lo = (long)floor(v);
  const long hi = lo + 1 < n ? lo + 1 : n - 1;
  const double t = v - (double)lo, a = s[lo], b = s[hi], d = b - a;
  return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}
__global__ void ts_iqr_kernel(const double* __restrict__ sorted, long n, double* __restrict__ acc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) acc[2] = ts_percentile(sorted, n, 0.75) - ts_percentile(sorted, n, 0.25);
}
__global__ void __launch_bounds__(TS_T) ts_mark_kernel(double* __restrict__ x, long n, double factor, double* __restrict__ acc) {
  const double cut = factor * acc[2];
  double cnt = 0;
  for (long i = blockIdx.x * (long)TS_T + threadIdx.x; i < n; i += (long)gridDim.x * TS_T) {
    const double v = x[i];
    if (v > cut || v < -cut) {
      x[i] = nan("");
      cnt += 1;
    }
  }
  const double t = block_sum(cnt);
  if (threadIdx.x == 0 && t != 0) atomicAdd(acc + 3, t);
}
// fixed path: one thread per frame, np.interp over the marker index (values live at x[f * N + frame])
__global__ void __launch_bounds__(TS_T) ts_interp_frames_kernel(double* __restrict__ x, long N, int F, double* __restrict__ acc) {
  for (long fr = blockIdx.x * (long)TS_T + threadIdx.x; fr < N; fr += (long)gridDim.x * TS_T) {
    int prev = -1;                       // last valid marker index seen so far
    double vprev = 0;
    int f = 0;
    while (f < F) {
      const double v = x[(long)f * N + fr];
      if (v == v) {
        prev = f; vprev = v;
        ++f;
        continue;
      }
      int nx = f + 1;                    // run of NaNs [f, nx)
      while (nx < F && !(x[(long)nx * N + fr] == x[(long)nx * N + fr])) ++nx;
      if (nx >= F && prev < 0) {         // the whole frame is NaN: np.interp would raise; counted, left as NaN
        atomicAdd(acc + 4, (double)(F));
        break;
      }
      const double vnext = nx < F ? x[(long)nx * N + fr] : 0.0;
      for (int g = f; g < nx; ++g) {
        double r;
        if (prev < 0) r = vnext;                          // left of the first sample point
        else if (nx >= F) r = vprev;                      // right of the last one
        else {
          const double slope = (vnext - vprev) / (double)(nx - prev);
          r = slope * (double)(g - prev) + vprev;         // numpy: slope * (x - xp[j]) + fp[j]
        }
        x[(long)g * N + fr] = r;
      }
      f = nx;
    }
  }
}
// aligned path: first / last valid sample of every marker (one block per marker)
__global__ void __launch_bounds__(TS_T) ts_first_last_kernel(const double* __restrict__ x, long N, double* __restrict__ first,
                                                             double* __restrict__ last, int* __restrict__ has) {
  __shared__ long long s_lo, s_hi;
  const int f = blockIdx.x;
  if (threadIdx.x == 0) { s_lo = (long long)N; s_hi = -1; }
  __syncthreads();
  long long lo = N, hi = -1;
  for (long i = threadIdx.x; i < N; i += TS_T) {
    const double v = x[(long)f * N + i];
    if (v == v) {
      if (i < lo) lo = i;
      if (i > hi) hi = i;
    }
  }
  atomicMin(&s_lo, lo);
  atomicMax(&s_hi, hi);
  __syncthreads();
  if (threadIdx.x == 0) {
    has[f] = s_hi >= 0;
    first[f] = s_hi >= 0 ? x[(long)f * N + s_lo] : 0.0;
    last[f] = s_hi >= 0 ? x[(long)f * N + s_hi] : 0.0;
  }
}
__global__ void ts_fill_values_kernel(int F, const double* __restrict__ first, const double* __restrict__ last, const int* __restrict__ has,
                                      double* __restrict__ fill, double* __restrict__ acc) {
  const int f = threadIdx.x;
  if (f >= F) return;
  if (has[f]) {
    fill[f] = last[f];
    return;
  }
  int a = -1, b = -1;
  for (int g = f - 1; g >= 0; --g)
    if (has[g]) { a = g; break; }
  for (int g = f + 1; g < F; ++g)
    if (has[g]) { b = g; break; }
  if (a < 0 && b < 0) {
    fill[f] = nan("");
    atomicAdd(acc + 4, 1.0);
  } else if (a < 0) fill[f] = first[b];
  else if (b < 0) fill[f] = last[a];
  else {
    const double slope = (first[b] - last[a]) / (double)(b - a);
    fill[f] = slope * (double)(f - a) + last[a];
  }
}
__global__ void __launch_bounds__(TS_T) ts_fill_kernel(double* __restrict__ x, long N, int F, const double* __restrict__ fill) {
  const long n = N * F;
  for (long i = blockIdx.x * (long)TS_T + threadIdx.x; i < n; i += (long)gridDim.x * TS_T) {
    const double v = x[i];
    if (!(v == v)) x[i] = fill[i / N];
  }
}
// population standard deviation of every row (marker) over time: one block per row, two passes
__global__ void __launch_bounds__(TS_T) ts_row_std_kernel(const double* __restrict__ x, long N, double* __restrict__ out) {
  __shared__ double s_mean;
  const double* row = x + (long)blockIdx.x * N;
  double p = 0;
  for (long i = threadIdx.x; i < N; i += TS_T) p += row[i];
  double t = block_sum(p);
  if (threadIdx.x == 0) s_mean = t / (double)N;
  __syncthreads();
  const double m = s_mean;
  p = 0;
  for (long i = threadIdx.x; i < N; i += TS_T) {
    const double d = row[i] - m;
    p += d * d;
  }
  t = block_sum(p);
  if (threadIdx.x == 0) out[blockIdx.x] = sqrt(t / (double)N);
}
// Savitzky-Golay along time: interior = correlation with coeffs[window]; the first / last `half` outputs are head / tail
// [half][window] applied to the first / last `window` samples (scipy mode='interp')
__global__ void __launch_bounds__(TS_T) ts_savgol_kernel(const double* __restrict__ x, long N, int F, int window,
                                                         const double* __restrict__ coeffs, const double* __restrict__ head,
                                                         const double* __restrict__ tail, double* __restrict__ out) {
  const int half = window / 2;
  const long n = N * F;
  for (long i = blockIdx.x * (long)TS_T + threadIdx.x; i < n; i += (long)gridDim.x * TS_T) {
    const long f = i / N, t = i - f * N;
    const double* row = x + f * N;
    double a = 0;
    if (t < half) {
      for (int k = 0; k < window; ++k) a += head[t * window + k] * row[k];
    } else if (t >= N - half) {
      const long e = t - (N - half);
      for (int k = 0; k < window; ++k) a += tail[e * window + k] * row[N - window + k];
    } else {
      for (int k = 0; k < window; ++k) a += coeffs[k] * row[t - half + k];
    }
    out[i] = a;
  }
}

static inline unsigned ts_grid(long n) {
  long b = (n + TS_T - 1) / TS_T;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace vb

using namespace vb;

extern "C" {

size_t vame_trainset_workspace_bytes(long n_frames, int num_features) {
  if (n_frames <= 0 || num_features <= 0) return 0;
  return ts_carve(n_frames, num_features, nullptr).bytes;
}

int vame_trainset_zscore_clean(const double* data_fn, long n_frames, int num_features, int robust, double iqr_factor, int fixed,
                               double* xz_fn, double* stats_out, void* ws, size_t ws_bytes, void* stream) {
  VB_REQUIRE(data_fn && xz_fn && ws, "vame_trainset_zscore_clean: null pointer");
  VB_REQUIRE(n_frames > 0 && num_features > 0 && num_features <= TS_MAXF, "vame_trainset_zscore_clean: 1 <= num_features <= 128, n_frames > 0");
  const long n = n_frames * num_features;
  VB_REQUIRE(n < (1L << 31), "vame_trainset_zscore_clean: more than 2^31 entries");
  TsWs w = ts_carve(n_frames, num_features, ws);
  VB_REQUIRE(ws_bytes >= w.bytes, "vame_trainset_zscore_clean: workspace too small (see vame_trainset_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(w.acc, 0, 8 * sizeof(double), st);
  count_launch(3);
  ts_moment_kernel<<<ts_grid(n), TS_T, 0, st>>>(data_fn, n, 0, w.acc);
  ts_moment_kernel<<<ts_grid(n), TS_T, 0, st>>>(data_fn, n, 1, w.acc);
  ts_zscore_kernel<<<ts_grid(n), TS_T, 0, st>>>(data_fn, n, w.acc, xz_fn, robust ? w.keys[0] : nullptr);
  if (robust) {
    size_t cb = w.cub_bytes;
    cub::DeviceRadixSort::SortKeys(w.cub_tmp, cb, (const double*)w.keys[0], w.keys[1], (int)n, 0, 64, st);
    count_launch(3);
    ts_iqr_kernel<<<1, 32, 0, st>>>(w.keys[1], n, w.acc);
    ts_mark_kernel<<<ts_grid(n), TS_T, 0, st>>>(xz_fn, n, iqr_factor, w.acc);
    if (fixed) {
      ts_interp_frames_kernel<<<ts_grid(n_frames), TS_T, 0, st>>>(xz_fn, n_frames, num_features, w.acc);
    } else {
      count_launch(2);
      ts_first_last_kernel<<<num_features, TS_T, 0, st>>>(xz_fn, n_frames, w.first, w.last, w.has);
      ts_fill_values_kernel<<<1, TS_MAXF, 0, st>>>(num_features, w.first, w.last, w.has, w.fill, w.acc);
      ts_fill_kernel<<<ts_grid(n), TS_T, 0, st>>>(xz_fn, n_frames, num_features, w.fill);
    }
  }
  // stats: [mean, std, iqr, #outliers, #entries left NaN]
  if (stats_out) cudaMemcpyAsync(stats_out, w.acc, 5 * sizeof(double), cudaMemcpyDeviceToDevice, st);
  return check_launch("vame_trainset_zscore_clean");
}

int vame_trainset_row_std(const double* x_fn, long n_frames, int num_features, double* std_out, void* stream) {
  VB_REQUIRE(x_fn && std_out && n_frames > 0 && num_features > 0, "vame_trainset_row_std: bad arguments");
  count_launch();
  ts_row_std_kernel<<<num_features, TS_T, 0, (cudaStream_t)stream>>>(x_fn, n_frames, std_out);
  return check_launch("vame_trainset_row_std");
}

int vame_trainset_savgol(const double* x_fn, long n_frames, int num_features, int window, const double* coeffs, const double* head,
                         const double* tail, double* out_fn, void* stream) {
  VB_REQUIRE(x_fn && out_fn && coeffs && head && tail, "vame_trainset_savgol: null pointer");
  VB_REQUIRE(window >= 3 && (window & 1) && n_frames >= window && num_features > 0, "vame_trainset_savgol: odd window <= n_frames required");
  count_launch();
  ts_savgol_kernel<<<ts_grid(n_frames * num_features), TS_T, 0, (cudaStream_t)stream>>>(x_fn, n_frames, num_features, window, coeffs, head,
                                                                                        tail, out_fn);
  return check_launch("vame_trainset_savgol");
}

}  // extern "C"
