// Training-mode dropout of the WSWGAT path.
//
// The reference draws an independent nn.Dropout mask of the layer INPUT for every attention head
// (`attn_head(g, self.dropout(h))`, GATStackLayer.py:56) and one mask on the FFN output before the residual
// (GATLayer.py:41-42).  Masks here come from a counter-based generator - keep(seed, stream, element index) is a
// pure function - so nothing is stored: the backward pass regenerates the same mask from the same three numbers.
//
// Per-head input dropout as ONE product on the existing GEMM kernels:
//     z_k = W_k (h * m_k) / (1-p)    for all heads k    ==    zp = A' . W_blk^T
//     A'   [N_src, H*in] = [ h*m_0 | h*m_1 | ... ] / (1-p)                       (dropout_expand)
//     W_blk[ldz,   H*in] : row c carries W_aug[c, :] in the column block of ITS head, zeros elsewhere  (wblk_build)
// and backward  dA' = dzp . W_blk,  d h = sum_k m_k * dA'_k / (1-p)  (dropout_reduce),
//               dW_blk = dzp^T . A',  dW_aug[c, :] = dW_blk[c, block head(c)]       (wblk_gather).
#include "hsg_common.cuh"
#include "hsg_edge_layout.cuh"
#include "hsg_internal.cuh"

namespace hsg {

__device__ __forceinline__ int wblk_head(const EdgeLayout& L, int row) {
  if (row < L.fp) {
    const int c = edge_unperm(L, row);
    return c < 0 ? -1 : c / L.D;
  }
  return row < L.fp + L.H ? row - L.fp : -1;
}

// grid.x over rows u, threads over (k, i)
__global__ void __launch_bounds__(256)
dropout_expand_kernel(int n, int in_dim, int H, const float* __restrict__ h, float* __restrict__ out, DropCfg dc) {
  pdl_prologue();
  const int W = H * in_dim;
  for (int u = blockIdx.x; u < n; u += gridDim.x) {
    const float* hr = h + (size_t)u * in_dim;
    float* o = out + (size_t)u * W;
    for (int t = threadIdx.x; t < W; t += blockDim.x) {
      const int k = t / in_dim, i = t - k * in_dim;
      const uint64_t idx = ((uint64_t)k * (uint64_t)n + (uint64_t)u) * (uint64_t)in_dim + (uint64_t)i;
      o[t] = drop_keep(dc, idx) ? hr[i] * dc.scale : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
dropout_reduce_kernel(int n, int in_dim, int H, const float* __restrict__ dA, const float* __restrict__ add,
                      float* __restrict__ out, DropCfg dc) {
  pdl_prologue();
  const int W = H * in_dim;
  for (int u = blockIdx.x; u < n; u += gridDim.x) {
    const float* dr = dA + (size_t)u * W;
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < H; ++k) {                       // fixed head order
        const uint64_t idx = ((uint64_t)k * (uint64_t)n + (uint64_t)u) * (uint64_t)in_dim + (uint64_t)i;
        if (drop_keep(dc, idx)) s += dr[k * in_dim + i];
      }
      s *= dc.scale;
      if (add) s += add[(size_t)u * in_dim + i];
      out[(size_t)u * in_dim + i] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
wblk_build_kernel(EdgeLayout L, int in_dim, const float* __restrict__ W_aug, float* __restrict__ W_blk) {
  pdl_prologue();
  const int c = blockIdx.x, hd = wblk_head(L, c), W = L.H * in_dim;
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    const int k = t / in_dim, i = t - k * in_dim;
    W_blk[(size_t)c * W + t] = (k == hd) ? W_aug[(size_t)c * in_dim + i] : 0.f;
  }
}

__global__ void __launch_bounds__(256)
wblk_gather_kernel(EdgeLayout L, int in_dim, const float* __restrict__ dW_blk, float* __restrict__ dW_aug,
                   int accumulate) {
  pdl_prologue();
  const int c = blockIdx.x, hd = wblk_head(L, c), W = L.H * in_dim;
  for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
    const float v = hd >= 0 ? dW_blk[(size_t)c * W + hd * in_dim + i] : 0.f;
    float* o = dW_aug + (size_t)c * in_dim + i;
    *o = accumulate ? *o + v : v;
  }
}

// out = x * keep / (1-p)  over a flat [n] array (element index = position)
__global__ void __launch_bounds__(256)
dropout_mul_kernel(size_t n, const float* __restrict__ x, float* __restrict__ out, DropCfg dc) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = drop_keep(dc, i) ? x[i] * dc.scale : 0.f;
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(size_t n, unsigned char* __restrict__ out, DropCfg dc) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = drop_keep(dc, i) ? 1 : 0;
}

static int grid_rows(int n) { return n < 1 ? 1 : (n > 148 * 16 ? 148 * 16 : n); }
static int grid_flat(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

int dropout_expand(int n, int in_dim, int H, const float* h, float* out, DropCfg dc, cudaStream_t s) {
  if (n <= 0) return HSG_OK;
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(dropout_expand_kernel, dim3(grid_rows(n)), dim3(256), 0, s, n, in_dim, H, h, out, dc);
  return check_launch();
}

int dropout_reduce(int n, int in_dim, int H, const float* dA, const float* add, float* out, DropCfg dc,
                   cudaStream_t s) {
  if (n <= 0) return HSG_OK;
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(dropout_reduce_kernel, dim3(grid_rows(n)), dim3(256), 0, s, n, in_dim, H, dA, add, out, dc);
  return check_launch();
}

int wblk_build(int H, int d, int in_dim, int ld_rows, const float* W_aug, float* W_blk, cudaStream_t s) {
  const EdgeLayout L = make_edge_layout(H, d);
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(wblk_build_kernel, dim3(ld_rows), dim3(256), 0, s, L, in_dim, W_aug, W_blk);
  return check_launch();
}

int wblk_gather(int H, int d, int in_dim, int ld_rows, const float* dW_blk, float* dW_aug, int accumulate,
                cudaStream_t s) {
  const EdgeLayout L = make_edge_layout(H, d);
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(wblk_gather_kernel, dim3(ld_rows), dim3(256), 0, s, L, in_dim, dW_blk, dW_aug, accumulate);
  return check_launch();
}

int dropout_mul(size_t n, const float* x, float* out, DropCfg dc, cudaStream_t s) {
  if (n == 0) return HSG_OK;
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(dropout_mul_kernel, dim3(grid_flat(n)), dim3(256), 0, s, n, x, out, dc);
  return check_launch();
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_dropout_mask(size_t n, float p, unsigned long long seed, unsigned int stream_id, unsigned char* out,
                     void* stream) {
  if (!out || !(p >= 0.f) || !(p < 1.f)) return HSG_ERR_ARG;
  if (n == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_DROPOUT, s);
  launch_k(dropout_mask_kernel, dim3(grid_flat(n)), dim3(256), 0, s, n, out, make_drop(p, seed, stream_id));
  return check_launch();
}

}  // extern "C"
