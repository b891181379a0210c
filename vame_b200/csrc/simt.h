// Launch prototypes of the SIMT (HBM-bound) kernels in simt.cu.
#pragma once
#include <cuda_runtime.h>

namespace vb {

enum { ACC_REC = 0, ACC_FUT = 1, ACC_KL = 2, ACC_KMEANS = 3, ACC_COUNT = 8 };
// device-resident hyper-parameters that change between steps (kept out of kernel arguments so that a captured
// CUDA graph of the whole train step stays valid): hyper[HY_*]
enum { HY_LR = 0, HY_KLW = 1, HY_BETA = 2, HY_KMLAMBDA = 3, HY_COUNT = 8 };

struct LambdaBwdArgs {
  const float* dz[4];          // gradient pieces wrt z, each [B, Z] or nullptr
  const float* dmu_ext;        // optional external gradient wrt mu / logvar outputs ([B, Z])
  const float* dlv_ext;
  const float* mu; const float* logvar; const float* eps;   // [B, Z]; eps nullptr -> eval (no reparam term)
  const long long* use_eps_flag;   // optional device flag: 0 -> the forward ran in eval mode, ignore eps
  const float* lin; long ldl;  // forward linear output (needed for softplus')
  const float* hyper;          // device hyper-parameters or nullptr
  float c_kl;                  // used when hyper == nullptr: beta * kl_weight / (B * Z)
  int B, B_pad, Z, softplus;
  float* dlin; long ldd;       // [B_pad, 2Z]
};

void launch_lambda_fwd(const float* lin, long ldl, const float* eps, int B, int Z, int softplus, float* z, float* mu, float* logvar,
                       double* acc, cudaStream_t st);
void launch_lambda_bwd(const LambdaBwdArgs& a, cudaStream_t st);
// variants that also write the P16 operand of the GEMM that follows (z_p [B_pad, Z], dlin_p [B_pad, 2Z], dpred_p [rows, F])
void launch_lambda_fwd_p16(const float* lin, long ldl, const float* eps, int B, int B_pad, int Z, int softplus, float* z, float* mu,
                           float* logvar, double* acc, void* z_p, cudaStream_t st);
void launch_lambda_bwd_p16(const LambdaBwdArgs& a, void* dlin_p, cudaStream_t st);
void launch_mse_p16(const float* pred, long ldp, const float* target, int rows, int B, int B_pad, int F, float gscale, float* dpred,
                    double* acc, int slot, void* dpred_p, cudaStream_t st);
void launch_mse(const float* pred, long ldp, const float* target, int rows, int B, int B_pad, int F, float gscale, float* dpred,
                double* acc, int slot, cudaStream_t st);
// hyper != nullptr: lambda = hyper[HY_KMLAMBDA], gradient coefficient = hyper[HY_KLW]; else the scalar arguments
// state: optional CP_STATE_DOUBLES doubles that persist between calls (warm start of the eigen-solver, see simt.cu)
constexpr int CP_STATE_DOUBLES = 4 + 64 * 64;
void launch_cluster_prior(const float* z, int B, int Z, int kloss, double lmbda, double bsize, double gcoef, const float* hyper,
                          float* dz, double* acc, cudaStream_t st, double* state = nullptr);
void launch_adam_prepare(float lr, const float* hyper, int* step_dev, float* scratch2, float b1, float b2, cudaStream_t st);
void launch_adam_apply(float* p, const float* g, float* m, float* v, float* vmax, long n, const float* scratch2, float b1, float b2,
                       float eps, float grad_scale, cudaStream_t st);
void launch_colsum(const float* X, long ld, long rows, int N, float* out, cudaStream_t st);
void launch_timesum_fm(const float* X, long ld, int T, int Bp, int C, float* out, cudaStream_t st);
void launch_rowsum_fm(const float* X, long ld, long ncols, int nfeat, float* out, cudaStream_t st);
void launch_fm_to_rows(const float* src, long ld, int H, int B, float* dst, long dst_ld, cudaStream_t st);
void launch_parts_reduce(const float* parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H, float* out, cudaStream_t st);
// + P16 [B_pad rows, K = D*H] of the result viewed as [B][D*H]
void launch_parts_reduce_pack(const float* parts, int n_parts, long dir_stride, int D, int B, int B_pad, int H, float* out, void* out_p,
                              cudaStream_t st);
// step_dev: device int32 step counter (incremented by the kernel chain); lr from hyper[HY_LR] when hyper != nullptr
void launch_adam(float* p, const float* g, float* m, float* v, float* vmax, long n, float lr, const float* hyper, int* step_dev,
                 float* scratch2, float b1, float b2, float eps, float grad_scale, cudaStream_t st);
void launch_finalize_losses(const double* acc, float* out, double rec_div, double fut_div, double kl_n, double beta, double klw,
                            const float* hyper, int future, cudaStream_t st);

}  // namespace vb
