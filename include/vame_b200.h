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

#ifdef __cplusplus
}
#endif
#endif
