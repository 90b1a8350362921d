// Internal (non-ABI) variants of the C entry points that the whole-loop sequencer (hsg_loop.cu) needs:
// the same kernels with an `accumulate` switch on the outputs that are summed over the applications of the
// update loop (weights are shared by the 1 + 2*n_iter applications, HiGraph.py:98-106), so gradient
// accumulation happens in the producing kernel's last stage instead of separate add launches.
#pragma once
#include "hsg_common.cuh"

namespace hsg {

// ---- counter-based dropout masks --------------------------------------------------------------------------------
// keep(seed, stream, idx): splitmix64 of the element counter, top 24 bits against p * 2^24.  A pure function of
// its arguments: forward and backward regenerate the same mask, nothing is stored.
struct DropCfg {
  unsigned long long key;   // seed mixed with the stream id
  unsigned int thresh;      // drop when the 24-bit draw < thresh
  float scale;              // 1 / (1 - p)
  // optional device-resident step counter mixed into the key at run time: a CUDA graph that replays the same kernel
  // arguments every step still draws fresh masks (the counter is advanced by hsg_adam_step_dev).  NULL: key as is.
  const unsigned long long* step;
};

inline DropCfg make_drop(float p, unsigned long long seed, unsigned int stream_id,
                         const unsigned long long* step = nullptr) {
  DropCfg c;
  c.step = step;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(stream_id + 1u);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  c.key = z ^ (z >> 31);
  c.thresh = (unsigned int)((double)p * 16777216.0);
  c.scale = 1.f / (1.f - p);
  return c;
}

__device__ __forceinline__ unsigned long long drop_key(const DropCfg& c) {
  return c.step ? c.key + 0xD1B54A32D192ED03ull * (__ldg(c.step) + 1ull) : c.key;
}

__device__ __forceinline__ bool drop_keep(const DropCfg& c, unsigned long long idx) {
  unsigned long long z = drop_key(c) + 0x9E3779B97F4A7C15ull * (idx + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 40) >= c.thresh;
}

// hsg_dropout.cu
int dropout_expand(int n, int in_dim, int H, const float* h, float* out, DropCfg dc, cudaStream_t s);
int dropout_reduce(int n, int in_dim, int H, const float* dA, const float* add, float* out, DropCfg dc, cudaStream_t s);
int wblk_build(int H, int d, int in_dim, int ld_rows, const float* W_aug, float* W_blk, cudaStream_t s);
int wblk_gather(int H, int d, int in_dim, int ld_rows, const float* dW_blk, float* dW_aug, int accumulate, cudaStream_t s);
int dropout_mul(size_t n, const float* x, float* out, DropCfg dc, cudaStream_t s);
// hsg_ffn.cu: r <- dropout(r) + resid (in place), then LayerNorm(r)
int layernorm_fwd_dropres(int N, int D, float* r, const float* resid, DropCfg dc, const float* gamma, const float* beta,
                          float* y, float* stats, cudaStream_t s);

// C (+)= A^T B, colsum (+)= column sums of A
int gemm_tn_ex(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
               float* colsum, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s, int cta_budget = 0);

// dgamma / dbeta (+)= ...
// defer_blocks != NULL (here, in ffn_rows_bwd and in edge_bwd_ex): the kernel leaves its per-block partials in `ws`,
// the fixed-order reduce is NOT launched and *defer_blocks tells how many partial rows there are - the update loop
// reduces the partials of all applications of a layer with ONE launch, off the critical chain.
int layernorm_bwd_ex(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma, float* dr,
                     float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s,
                     int* defer_blocks = nullptr);
// dgamma / dbeta (+)= column sums of nseg x nblocks partial rows (segment j at part + j * seg_stride floats)
int ln_partials_reduce(int nseg, int nblocks, size_t seg_stride, int D, const float* part, float* dgamma, float* dbeta,
                       int accumulate, cudaStream_t s);

// hsg_ffn.cu: whole position-wise FFN of a small node set in one launch each way (exact fp32); ffn_rows_ok tells
// whether the shape qualifies.  Same outputs as gemm_nt + gemm_nt + layernorm_fwd / layernorm_bwd + gemm_nn + gemm_nn.
bool ffn_rows_ok(int n, int F, int d_hid);
int ffn_rows_fwd(int n, int F, int d_hid, const float* x, const float* w1, const float* b1, const float* w2,
                 const float* b2, const float* gamma, const float* beta, float* hdn, float* r, float* y, float* stats,
                 cudaStream_t s);
int ffn_rows_bwd(int n, int F, int d_hid, const float* dy, const float* r, const float* stats, const float* gamma,
                 const float* hdn, const float* w1, const float* w2, float* dr, float* dhp, float* dx, float* dgamma,
                 float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s, int* defer_blocks = nullptr);
// hsg_gemm.cu: products below the small-product threshold (hsg_set_gemm_small_flops) run on FFMA tiles
bool gemm_is_small(int M, int N, int K);

// dq (+)= ...
int edge_bwd_ex(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q, const float* g,
                const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, int accumulate_dq, cudaStream_t s,
                int* defer_blocks = nullptr);
// hsg_edge.cu: dq (+)= sum of nseg x nblocks partial rows (segment j at dq_part + j * seg_stride floats), fixed order
int edge_dq_reduce(int nseg, int nblocks, size_t seg_stride, int nq, const float* dq_part, float* dq, int accumulate,
                   cudaStream_t s);

// hsg_edge_rc.cu: whether the update loop recomputes sh in the backward prep (then the forward does not store it)
bool edge_recompute_use(const hsg_csc* csc_fwd, int H, int d, int ldz);

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
