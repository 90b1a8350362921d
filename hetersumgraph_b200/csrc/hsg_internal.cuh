// Internal (non-ABI) variants of the C entry points that the whole-loop sequencer (hsg_loop.cu) needs:
// the same kernels with an `accumulate` switch on the outputs that are summed over the applications of the
// update loop (weights are shared by the 1 + 2*n_iter applications, HiGraph.py:98-106), so gradient
// accumulation happens in the producing kernel's last stage instead of separate add launches.
#pragma once
#include "hsg_common.cuh"

namespace hsg {

// C (+)= A^T B, colsum (+)= column sums of A
int gemm_tn_ex(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
               float* colsum, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s);

// dgamma / dbeta (+)= ...
int layernorm_bwd_ex(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma, float* dr,
                     float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s);

// dq (+)= ...
int edge_bwd_ex(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q, const float* g,
                const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, int accumulate_dq, cudaStream_t s);

// dW, dWf, dbf, da, dT (+)= ...   (acc_params: the four layer parameters; acc_T: the shared TF-IDF table)
int attn_prep_bwd_ex(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                     const float* bf, const float* a, const float* T, const float* dW_aug, const float* dq, float* dW,
                     float* dWf, float* dbf, float* da, float* dT, int acc_params, int acc_T, cudaStream_t s);

// hsg_gemm_small.cu: 64x64 FFMA tiles, reduction split over a thread-block cluster (DSMEM reduce)
int gemm_small_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* bias, const float* R, int ldr, int epi, cudaStream_t s);
int gemm_small_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* R, int ldr, int epi, cudaStream_t s);
int gemm_small_tn(int Rows, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  float* colsum, int accumulate, cudaStream_t s);

}  // namespace hsg
