// On-device window sampler: SEQUENCE_DATASET.__getitem__ (vame/model/dataloader.py:45-56) for a whole
// batch, the DataLoader collation and the first lines of train() (vame/model/rnn_vae.py:107-112: permute, split into data /
// future, cast to float32) as ONE kernel that writes straight into the static input buffers of the captured train step - plus
// the reparameterisation noise the reference draws with randn_like (rnn_model.py:73).
//   * the series stays resident in HBM in the reference's own layout and dtype ((F, N) float64, <file>.npy as saved by
//     create_trainset), so (x - mean) / std is evaluated in float64 exactly as dataloader.py:54 does and then rounded to float32
//     exactly as rnn_vae.py:109-110 does: for a given start index the output is bit-identical to the reference's
//   * window starts: uniform in [0, N - window) like np.random.choice(N - window) (the reference's stream is unseeded numpy, so
//     only the distribution can match); here a counter-based Philox4x32-10 stream keyed by (seed, draw counter, window index):
//     reproducible, graph-replay safe (the draw counter lives in device memory and is advanced by the kernel chain itself)
//   * HBM-bound by construction: 2T x F x 8 B read + (T + S) x F x 4 B written per window (C2: 11.5 KB + 4.3 KB), staged through
//     shared memory so that both the float64 reads (along time) and the float32 writes (along features) are coalesced
#include "../../include/vame_b200.h"
#include "api_common.h"
#include "kernels.h"

namespace vb {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}
// Philox4x32-10 (Salmon et al., SC'11): 4 x 32 random bits for counter (c0..c3) under key (k0, k1)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct SampleArgs {
  const double* series;       // (F, N) float64
  long n_frames;
  int F, window, T, S, Z, batch;
  double mean, std;
  const long long* starts_in; // optional explicit starts [batch]
  unsigned long long seed;
  const unsigned long long* counter;   // device: number of batches drawn so far
  float* x; float* fut; float* eps;
  long long* starts_out;
};

// grid = batch, block = 256; dynamic shared memory = F * window floats
__global__ void __launch_bounds__(256) sample_windows_kernel(const SampleArgs a) {
  extern __shared__ float sm_win[];                       // [t][f]: the (B, T, F) order of the outputs
  const int b = blockIdx.x, tid = threadIdx.x;
  const unsigned long long draw = a.counter ? *a.counter : 0ull;
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
  long long start;
  if (a.starts_in) {
    start = a.starts_in[b];
  } else {
    const uint4 r = philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)b, 0u, k0, k1);
    const unsigned long long n_start = (unsigned long long)(a.n_frames - a.window);       // np.random.choice(nf - temp_window)
    const unsigned long long u = ((unsigned long long)r.x << 32) | r.y;
    start = (long long)__umul64hi(u, n_start);                                             // floor(u / 2^64 * n_start)
  }
  if (tid == 0 && a.starts_out) a.starts_out[b] = start;
  const int W = a.window, F = a.F;
  for (int i = tid; i < F * W; i += blockDim.x) {         // coalesced along time
    const int f = i / W, t = i - f * W;
    const double v = (a.series[(long)f * a.n_frames + start + t] - a.mean) / a.std;        // dataloader.py:54, float64
    sm_win[t * F + f] = (float)v;                                                           // rnn_vae.py:109-110
  }
  __syncthreads();
  float* xb = a.x + (long)b * a.T * F;
  for (int i = tid; i < a.T * F; i += blockDim.x) xb[i] = sm_win[i];
  if (a.fut) {
    float* fb = a.fut + (long)b * a.S * F;
    for (int i = tid; i < a.S * F; i += blockDim.x) fb[i] = sm_win[a.T * F + i];
  }
  if (a.eps) {                                            // standard normal noise: Box-Muller on two Philox words per element
    for (int z = tid; z < a.Z; z += blockDim.x) {
      const uint32_t e = (uint32_t)(b * a.Z + z);
      const uint4 r = philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), e, 1u, k0, k1);
      const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);                   // (0, 1)
      const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
      a.eps[(long)b * a.Z + z] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
  }
}
__global__ void sample_advance_kernel(unsigned long long* counter) { *counter += 1ull; }

}  // namespace vb

extern "C" int vame_sample_windows(const double* series_fn, long n_frames, int num_features, int window, double mean, double std,
                                   int batch, int t_data, int t_future, int zdims, const long long* starts, unsigned long long seed,
                                   unsigned long long* counter, float* x, float* fut, float* eps, long long* starts_out, void* stream) {
  VB_REQUIRE(series_fn && x, "vame_sample_windows: null pointer");
  VB_REQUIRE(num_features > 0 && window > 0 && batch > 0, "vame_sample_windows: empty batch");
  VB_REQUIRE(t_data > 0 && t_future >= 0 && t_data + t_future <= window, "vame_sample_windows: need 0 < t_data, t_data + t_future <= window");
  VB_REQUIRE(n_frames > window, "vame_sample_windows: the series is shorter than one window");
  VB_REQUIRE(std != 0.0, "vame_sample_windows: std must not be 0");
  VB_REQUIRE(!eps || zdims > 0, "vame_sample_windows: zdims must be positive when eps is requested");
  VB_REQUIRE((size_t)num_features * window * sizeof(float) <= 160 * 1024, "vame_sample_windows: window x features exceeds 160 KB of shared memory");
  vb::SampleArgs a{};
  a.series = series_fn; a.n_frames = n_frames; a.F = num_features; a.window = window; a.T = t_data; a.S = fut ? t_future : 0; a.Z = zdims;
  a.batch = batch; a.mean = mean; a.std = std; a.starts_in = starts; a.seed = seed; a.counter = counter;
  a.x = x; a.fut = fut; a.eps = eps; a.starts_out = starts_out;
  const size_t smem = (size_t)num_features * window * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(vb::sample_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  vb::count_launch();
  vb::sample_windows_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(a);
  if (counter) {                                          // every call consumes one draw (also with explicit starts: eps)
    vb::count_launch();
    vb::sample_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
  }
  return vb::check_launch("vame_sample_windows");
}
