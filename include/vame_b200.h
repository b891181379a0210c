/* vame_b200 — C-ABI of the B200-native RNN-VAE hot path (drop-in for the compute behind
 * LINCellularNeuroscience/VAME's vame.train_model / vame.pose_segmentation).
 *
 * Conventions
 *   - every entry point returns 0 on success, <0 on error; vame_last_error() gives a thread-local message
 *   - all pointers are DEVICE pointers unless the name says host; the caller owns every buffer (incl. workspaces)
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit synchronisation
 *   - fp32 in / fp32 out, tensors contiguous row-major exactly as PyTorch holds them
 */
#ifndef VAME_B200_H
#define VAME_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VAME_B200_ABI_VERSION 1

const char* vame_last_error(void);
int vame_abi_version(void);
/* number of kernels enqueued by this library so far (host counter) */
long vame_launch_count(void);
/* runtime options: "pdl" (default 1) chains the per-step recurrent kernels with programmatic dependent launch;
 * "streams" (default 1) runs independent branches on internal side streams; "persistent" (default 2) bit0/bit1 select the
 * persistent cluster kernel for the forward/backward sweeps */
int vame_set_option(const char* name, int value);
int vame_get_option(const char* name);   /* -1 for unknown names */
/* measurement hook: device buffer of 16 uint64 that gru_step_fwd_kernel's CTA 0 fills with %globaltimer stamps (NULL = off) */
int vame_set_debug_buffer(void* device_u64x16);

/* ---- building blocks ------------------------------------------------------------------------ */
/* bytes of a P16 (bf16 hi/lo split, tensor-core tiled) copy of a [rows, k] matrix */
size_t vame_p16_bytes(int rows, int k, int row_block);
/* fp32 -> P16.  value(r,c) = src[row_map[r]*ld + col_map[c]] (maps optional; transposed swaps the roles) */
int vame_pack_p16(const float* src, long ld, int transposed, int rows, int k, int rows_src, int k_src,
                  const int* row_map, const int* col_map, int row_block, void* out, void* stream);
/* C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N]) on tcgen05 tensor cores, 3-pass bf16 split (fp32-accurate).
 * Replaces the torch.nn.functional.linear / addmm calls behind nn.GRU's input projections
 * (vame/model/rnn_model.py:34-35,41) and their backward. */
int vame_gemm_p16(const void* a_p, int a_nkc, const void* b_p, int b_nkc, int M, int N, float* C, long ldc,
                  const float* bias, int accumulate, int splits, void* stream);

/* ---- model level ----------------------------------------------------------------------------- */
/* Mirrors the constructor arguments of RNN_VAE (vame/model/rnn_model.py:148-161); time_window is the model's
 * seq_len = TEMPORAL_WINDOW/2 (rnn_model.py:154).  hidden sizes must be multiples of 32 and <= 256. */
typedef struct {
  int num_features;    /* NUM_FEATURES */
  int time_window;     /* seq_len */
  int zdims;           /* ZDIMS (<= 64) */
  int hidden_enc;      /* hidden_size_layer_1 (the reference ignores hidden_size_layer_2, rnn_model.py:34) */
  int hidden_rec;      /* hidden_size_rec */
  int hidden_pred;     /* hidden_size_pred */
  int future_decoder;  /* FUTURE_DECODER */
  int future_steps;    /* FUTURE_STEPS */
  int softplus;        /* Lambda softplus flag */
} vame_dims;

/* loss configuration of train()/test() (vame/model/rnn_vae.py:94-210) */
typedef struct {
  int mse_red_mean;      /* mse_reconstruction_reduction: 0 = 'sum', 1 = 'mean' */
  int mse_pred_mean;     /* mse_prediction_reduction */
  int kmeans_loss;       /* number of singular values kept (cfg['kmeans_loss']) */
  float kmeans_lambda;   /* used when hyper == NULL */
  float bsize;           /* the batch_size argument of cluster_loss (cfg['batch_size']) */
  float beta;            /* used when hyper == NULL */
  float kl_weight;       /* used when hyper == NULL */
  int with_future;       /* include the future-reconstruction term (train) or not (test, rnn_vae.py:186-190) */
  int defer_prior_join;  /* 1: the k-means-prior branch (internal side stream) is joined by the following vame_backward
                            instead of by vame_loss; losses_out is then valid only after that vame_backward */
} vame_loss_cfg;

/* Number of parameter tensors (44 with the future decoder, 32 without) and their placement in the flat fp32 buffer.
 * offsets[i]/sizes[i] follow the reference's state_dict order; returns the total float count. */
int vame_param_tensors(const vame_dims* d);
long vame_param_layout(const vame_dims* d, long* offsets, long* sizes);

/* bf16 hi/lo tensor-core copies of the weights; must be refreshed after every parameter update */
size_t vame_packed_weights_bytes(const vame_dims* d);
int vame_pack_weights(const vame_dims* d, const float* params, void* packed, void* stream);
/* Re-pack for a train loop with a fixed batch size: the W_hh formats that only the kernels for OTHER batch sizes read are
 * skipped (at B <= 512, H = 256 the sweeps run on the resident-weight cluster kernels; the slice-kernel formats are ~40 % of
 * the re-pack).  `packed` is then valid for vame_forward / vame_backward at that batch size only - call vame_pack_weights
 * before anything else uses it (vame_b200.engine marks the copies dirty). */
int vame_pack_weights_train(const vame_dims* d, const float* params, void* packed, int batch, void* stream);
/* Same result, for the train loop: only the formats the forward pass reads first (encoder layer 0) are packed now, on
 * `stream`; the rest is packed by the NEXT vame_forward (same params / packed pointers, which must stay valid until then) on a
 * low-priority internal stream beside its first recurrent sweep.  Any other entry point that reads `packed` completes the
 * re-pack on its own stream first.  Graph-capturable together with that vame_forward. */
int vame_pack_weights_deferred(const vame_dims* d, const float* params, void* packed, void* stream);

size_t vame_workspace_bytes(const vame_dims* d, int batch, int training);

/* RNN_VAE.forward (vame/model/rnn_model.py:162-179).  x: [B, T, F] (strides in floats, feature stride 1).
 * eps: [B, Z] reparameterisation noise (the reference draws it with randn_like, rnn_model.py:73) or NULL for eval
 * mode (z = mu).  save_for_backward keeps the activations BPTT needs in the workspace.  Outputs may be NULL. */
int vame_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* x, long x_bs,
                 long x_ts, const float* eps, int save_for_backward, float* pred, float* future, float* z, float* mu,
                 float* logvar, void* ws, size_t ws_bytes, void* stream);

/* The four loss terms of train()/test() on the last vame_forward (rnn_vae.py:124-129,135-138,188-197) and their
 * gradients wrt pred / future / z (kept in the workspace for vame_backward).
 * hyper: device float[8] {lr, kl_weight, beta, kmeans_lambda, ...} or NULL to use the cfg scalars.
 * losses_out: device float[8] = {rec, fut, kl, kmeans, total, 0, 0, 0}. */
/* target: optional [B, T, F] reconstruction target when it differs from the forward input (cfg['noise'], rnn_vae.py:116-124:
 * the model sees x + noise, the loss compares against the clean x); NULL = the forward input. */
/* Optional, for the train loop: with a cfg armed, the next vame_forward(save_for_backward = 1) calls start the k-means prior
 * (cluster_loss, rnn_vae.py:45-50) of THAT cfg / hyper on an internal stream as soon as z exists, and the vame_loss that
 * follows (same cfg, want_grads = 1) uses it instead of launching its own.  cfg = NULL disarms.  Host-side state only. */
int vame_arm_prior(const vame_loss_cfg* cfg, const float* hyper);

int vame_loss(const vame_dims* d, int batch, const vame_loss_cfg* cfg, const float* fut, long f_bs, long f_ts,
              const float* target, long t_bs, long t_ts, const float* hyper, float* losses_out, int want_grads, void* ws,
              size_t ws_bytes, void* stream);

/* Backward of the last vame_forward (replaces loss.backward(), rnn_vae.py:142).  If use_loss_grads != 0 the upstream
 * gradients are the ones vame_loss left in the workspace (d pred, d future, d z from the k-means prior, KL through
 * hyper / cfg); otherwise they are the external tensors (autograd path), any of which may be NULL.
 * grads: flat fp32 buffer in the vame_param_layout order, OVERWRITTEN (zero_grad + backward). */
int vame_backward(const vame_dims* d, int batch, const float* params, const void* packed, int use_loss_grads,
                  const vame_loss_cfg* cfg, const float* hyper, const float* dpred, const float* dfuture,
                  const float* dz, const float* dmu, const float* dlogvar, float* grads, void* ws,
                  size_t ws_bytes, void* stream);

/* Data-parallel overlap of the gradient allreduce with the tail of the backward pass (no counterpart in the reference, which is
 * single-device).  Every gradient except encoder layer 0's - the flat range [vame_grad_bucket_split(d), total) - is final ~200 us
 * before vame_backward ends.  With vame_grad_overlap(1), vame_backward records an external CUDA event at that point (an
 * event-record node when the call is captured into a CUDA graph); vame_wait_grads_ready(stream) makes `stream` wait for the
 * event of the most recently enqueued vame_backward, so the caller can all-reduce that range on a communication stream while
 * the last BPTT sweep still runs, and only the small encoder-layer-0 range [0, split) after the call. */
int vame_grad_overlap(int enable);   /* 0 off, 1 = consumer outside the captured graph (external event), 2 = consumer captured into the same graph */
long vame_grad_bucket_split(const vame_dims* d);
int vame_wait_grads_ready(void* stream);

/* torch.optim.Adam(amsgrad=True) (rnn_vae.py:332,143) over the flat buffers.  lr is read from hyper[0] when hyper is
 * not NULL.  step_dev: device int32 step counter (incremented here).  scratch: device float[2].
 * grad_scale multiplies the gradient first (1/world_size after a sum-allreduce). */
int vame_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long n,
                   float lr, const float* hyper, int* step_dev, float* scratch, float beta1, float beta2, float eps,
                   float grad_scale, void* stream);

/* The same optimizer step in two pieces, so that a train loop can update the parameters whose gradients are final early (everything
 * but encoder layer 0, see vame_grad_bucket_split) on a side stream while the backward pass still runs: vame_adam_prepare
 * increments the step counter and writes the bias-corrected step size into scratch (once per step), vame_adam_apply updates one
 * range of the flat buffers (n a multiple of 4, pointers already offset).  vame_pack_weights_train_part re-packs the formats of
 * encoder layer 0 (part 0) or of everything else (part 1). */
int vame_adam_prepare(float lr, const float* hyper, int* step_dev, float* scratch, float beta1, float beta2, void* stream);
int vame_adam_apply(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long n,
                    const float* scratch, float beta1, float beta2, float eps, float grad_scale, void* stream);
int vame_pack_weights_train_part(const vame_dims* d, const float* params, void* packed, int batch, int part, void* stream);

/* embedd_latent_vectors (vame/analysis/pose_segmentation.py:87-98): mu of every stride-1 window.
 * series: [n_frames, F] fp32 (frame-major, i.e. the transpose of the reference's (F, N) array);
 * windows first_window .. first_window + n_windows - 1 (window i covers frames i .. i+T-1) -> mu_out [n_windows, Z]. */
size_t vame_embed_workspace_bytes(const vame_dims* d, long n_frames, int chunk);
int vame_embed_windows(const vame_dims* d, const float* params, const void* packed, const float* series, long n_frames,
                       long first_window, long n_windows, int chunk, float* mu_out, void* ws, size_t ws_bytes,
                       void* stream);

/* standalone sub-module forwards (inference): Encoder.forward (rnn_model.py:40-45) -> hidden [B, 4H];
 * Lambda.forward (:63-76) ; Decoder.forward / Decoder_Future.forward (:99-109 / :133-144) with the input being the
 * broadcast of z (rnn_model.py:169-170). */
int vame_encoder_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* x, long x_bs,
                         long x_ts, float* hidden, void* ws, size_t ws_bytes, void* stream);
int vame_lambda_forward(const vame_dims* d, int batch, const float* params, const void* packed, const float* hidden,
                        const float* eps, float* z, float* mu, float* logvar, void* ws, size_t ws_bytes, void* stream);
int vame_decoder_forward(const vame_dims* d, int batch, int which, const float* params, const void* packed, const float* z,
                         float* pred, void* ws, size_t ws_bytes, void* stream);

/* measurement hook: re-run the encoder layer-1 forward (which=0) or backward (which=1) sweep on the buffers of the last
 * vame_forward(save)/vame_backward so that bench.py can time the recurrent step kernels with CUDA events */
int vame_debug_gru_sweep(const vame_dims* d, int batch, int which, const float* params, const void* packed, void* ws,
                         size_t ws_bytes, void* stream);

/* measurement hook: per-section timeline of vame_forward / vame_backward (eager mode only, not graph-capturable) */
int vame_debug_timeline(int enable);
int vame_debug_timeline_read(const char** names, float* ms, int max);

/* k-means prior on its own (cluster_loss, rnn_vae.py:45-50): loss_out device double[1], dlatent [B, Z] or NULL */
int vame_cluster_loss(const float* latent, int batch, int zdims, int kloss, float lmbda, float bsize, float grad_coef,
                      double* loss_out, float* dlatent, void* stream);

/* Training-set preparation (vame/model/create_training.py:94-264, traindata_aligned / traindata_fixed), float64 like the
 * reference; series are (num_features, n_frames) row-major = the reference's <file>-PE-seq.npy layout.
 * vame_trainset_zscore_clean: per-file stage (:106-137 / :204-232): z-score with the global mean / std of the file, and with
 *   robust != 0 the IQR outlier removal (|x| > iqr_factor * scipy.stats.iqr -> NaN) followed by the reference's interpolation:
 *   fixed != 0 every frame over the marker index (:231), else the 2-D interpol() of :137 (a NaN of marker f becomes the last
 *   valid time sample of marker f).  stats_out (device, 5 doubles, may be NULL): mean, std, iqr, #outliers, #entries left NaN.
 * vame_trainset_row_std: population std of every marker over time (anchor detection, :148).
 * vame_trainset_savgol: scipy.signal.savgol_filter(X, window, order) along time with mode='interp' (:175-178); coeffs[window],
 *   head / tail [window/2][window] are device arrays computed by the host mirror (vame_b200/create_training.py). */
size_t vame_trainset_workspace_bytes(long n_frames, int num_features);
int vame_trainset_zscore_clean(const double* data_fn, long n_frames, int num_features, int robust, double iqr_factor, int fixed,
                               double* xz_fn, double* stats_out, void* ws, size_t ws_bytes, void* stream);
int vame_trainset_row_std(const double* x_fn, long n_frames, int num_features, double* std_out, void* stream);
int vame_trainset_savgol(const double* x_fn, long n_frames, int num_features, int window, const double* coeffs, const double* head,
                         const double* tail, double* out_fn, void* stream);

/* ---- on-device window sampler ------------------------------------------------------------------------------
 * Replaces SEQUENCE_DATASET.__getitem__ (vame/model/dataloader.py:45-56) + torch DataLoader collation + the permute / split /
 * float32 cast at the top of train() (vame/model/rnn_vae.py:107-112) for one batch.  series_fn: the reference's (num_features,
 * n_frames) float64 array (train_seq.npy / test_seq.npy), resident on the device; mean / std: the scalars of seq_mean.npy /
 * seq_std.npy.  Window b covers frames start_b .. start_b + window - 1 with start_b uniform in [0, n_frames - window)
 * (np.random.choice(nf - temp_window), dataloader.py:49) drawn from a Philox4x32-10 stream keyed by (seed, *counter, b), or
 * taken from `starts` (device int64[batch]) when given.  Outputs: x [batch, t_data, F] = frames 0 .. t_data-1 of the z-scored
 * window as float32, fut [batch, t_future, F] = frames t_data .. t_data+t_future-1 (may be NULL), eps [batch, zdims] standard
 * normal (may be NULL; the reference's randn_like, rnn_model.py:73), starts_out int64[batch] (may be NULL).
 * counter: device uint64 draw counter (may be NULL = draw 0); incremented on the stream after the batch, so CUDA-graph replays
 * of the call produce fresh batches. */
int vame_sample_windows(const double* series_fn, long n_frames, int num_features, int window, double mean, double std,
                        int batch, int t_data, int t_future, int zdims, const long long* starts, unsigned long long seed,
                        unsigned long long* counter, float* x, float* fut, float* eps, long long* starts_out, void* stream);

/* ---- k-means on the latent vectors ---------------------------------------------------------------------
 * Replaces sklearn.cluster.KMeans(init='k-means++', n_clusters, random_state, n_init).fit / .predict as called at
 * vame/analysis/pose_segmentation.py:141-143 and :183-185.  x is [n, dim] fp32 row-major on the device (dim <= 64,
 * k <= 128); the host side (vame_b200/kmeans.py) draws the random numbers with the same numpy RandomState stream as
 * sklearn and drives these entry points. */
size_t vame_kmeans_workspace_bytes(long n, int dim, int k);
/* Lloyd iterations from centers_init [k, dim]: mean-centred data, tol relative to the mean feature variance, stops on
 * unchanged labels or centre shift <= tol, final E-step with the final centres.  Blocking (one status read per iteration).
 * labels int32 [n], centers_out [k, dim], inertia_out device double[1] (may be NULL), n_iter_out host int (may be NULL). */
int vame_kmeans_lloyd(const float* x, long n, int dim, int k, const float* centers_init, int max_iter, float tol, int* labels,
                      float* centers_out, double* inertia_out, int* n_iter_out, void* ws, size_t ws_bytes, void* stream);
/* KMeans.predict */
int vame_kmeans_assign(const float* x, long n, int dim, int k, const float* centers, int* labels, double* inertia_out, void* ws,
                       size_t ws_bytes, void* stream);
/* k-means++ round: newmin[j][i] = min(closest[i], |x_i - x_cand[j]|^2) (closest NULL: no min), pot[j] = sum_i newmin[j][i];
 * cand device int64[m], m <= 8, newmin [m, n], pot device double[m] */
int vame_kmeans_candidates(const float* x, long n, int dim, const long* cand, int m, const float* closest, float* newmin, double* pot,
                           void* stream);
/* k-means++ sampling: idx[j] = searchsorted(cumsum_fp64(closest), vals[j]) clipped to n - 1; vals device double[m], idx device int64[m] */
int vame_kmeans_sample(const float* closest, long n, const double* vals, int m, long* idx, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
