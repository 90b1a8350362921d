// K2: attention prep.  Folds attn_fc (GATLayer.py:87,126) into the node projection
// and into a 10 x H table over the TF-IDF boxes, so the edge kernel only adds two
// scalars per (edge, head):
//     p_u   = a_k[0:d] . z_u          -> extra rows of the projection weight (W_aug[F+k])
//     q[b,k]= a_k[2d:3d] . (Wf_k T[b] + bf_k)       (feat_fc of the tfidfembed of box b, HiGraph.py:150-151)
// The middle third a_k[d:2d] multiplies the destination's `z`, which DGL zero-fills
// (GATLayer.py:111,147 write z on the source node type only), so it is dead: its
// gradient is exactly zero, but it stays in the state_dict.
#include "hsg_common.cuh"

namespace hsg {

// grid = ld_rows + 1 blocks.  Block r < ld_rows writes row r of W_aug; the last block writes q.
__global__ void __launch_bounds__(256)
attn_prep_fwd_kernel(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* __restrict__ W,
                     const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ a,
                     const float* __restrict__ T, float* __restrict__ W_aug, float* __restrict__ q) {
  extern __shared__ float dfeat_s[];  // [10, F] (last block only)
  const int F = H * d;
  const int r = blockIdx.x;
  if (r < ld_rows) {
    float* out = W_aug + (size_t)r * in_dim;
    if (r < F) {
      for (int i = threadIdx.x; i < in_dim; i += blockDim.x) out[i] = W[(size_t)r * in_dim + i];
    } else if (r < F + H) {
      const int k = r - F;
      for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < d; ++j) s = fmaf(a[k * 3 * d + j], W[(size_t)(k * d + j) * in_dim + i], s);
        out[i] = s;
      }
    } else {
      for (int i = threadIdx.x; i < in_dim; i += blockDim.x) out[i] = 0.f;
    }
    return;
  }
  for (int o = threadIdx.x; o < HSG_N_BINS * F; o += blockDim.x) {
    const int b = o / F, c = o % F;
    float s = bf ? bf[c] : 0.f;
    for (int f = 0; f < feat_dim; ++f) s = fmaf(Wf[(size_t)c * feat_dim + f], T[b * feat_dim + f], s);
    dfeat_s[o] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < HSG_N_BINS * H; o += blockDim.x) {
    const int b = o / H, k = o % H;
    float s = 0.f;
    for (int j = 0; j < d; ++j) s = fmaf(a[k * 3 * d + 2 * d + j], dfeat_s[b * F + k * d + j], s);
    q[o] = s;
  }
}

// grid = F + 1 blocks.  Block r < F: dW row r and da[k, j] (a_src part).  Last block: feat path.
__global__ void __launch_bounds__(256)
attn_prep_bwd_kernel(int H, int d, int in_dim, int feat_dim, const float* __restrict__ W,
                     const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ a,
                     const float* __restrict__ T, const float* __restrict__ dW_aug, const float* __restrict__ dq,
                     float* __restrict__ dW, float* __restrict__ dWf, float* __restrict__ dbf,
                     float* __restrict__ da, float* __restrict__ dT) {
  extern __shared__ float sm[];
  const int F = H * d;
  const int r = blockIdx.x;
  if (r < F) {
    const int k = r / d, j = r % d;
    const float aj = a[k * 3 * d + j];
    const float* dwp = dW_aug + (size_t)(F + k) * in_dim;
    float part = 0.f;
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
      const float g = dwp[i];
      dW[(size_t)r * in_dim + i] = fmaf(aj, g, dW_aug[(size_t)r * in_dim + i]);
      part = fmaf(g, W[(size_t)r * in_dim + i], part);
    }
    // fixed-order block reduction
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
      da[k * 3 * d + j] = s;
      da[k * 3 * d + d + j] = 0.f;   // a_dst multiplies DGL's zero-filled destination z
    }
    return;
  }
  float* dfeat_s = sm;                       // [10, F]
  float* ddfeat_s = sm + HSG_N_BINS * F;     // [10, F]
  for (int o = threadIdx.x; o < HSG_N_BINS * F; o += blockDim.x) {
    const int b = o / F, c = o % F;
    float s = bf ? bf[c] : 0.f;
    for (int f = 0; f < feat_dim; ++f) s = fmaf(Wf[(size_t)c * feat_dim + f], T[b * feat_dim + f], s);
    dfeat_s[o] = s;
    const int k = c / d, j = c % d;
    ddfeat_s[o] = dq[b * H + k] * a[k * 3 * d + 2 * d + j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    const int k = c / d, j = c % d;
    float s = 0.f, sb = 0.f;
    for (int b = 0; b < HSG_N_BINS; ++b) {
      s = fmaf(dq[b * H + k], dfeat_s[b * F + c], s);
      sb += ddfeat_s[b * F + c];
    }
    da[k * 3 * d + 2 * d + j] = s;
    if (dbf) dbf[c] = sb;
  }
  for (int o = threadIdx.x; o < F * feat_dim; o += blockDim.x) {
    const int c = o / feat_dim, f = o % feat_dim;
    float s = 0.f;
    for (int b = 0; b < HSG_N_BINS; ++b) s = fmaf(ddfeat_s[b * F + c], T[b * feat_dim + f], s);
    dWf[o] = s;
  }
  for (int o = threadIdx.x; o < HSG_N_BINS * feat_dim; o += blockDim.x) {
    const int b = o / feat_dim, f = o % feat_dim;
    float s = 0.f;
    for (int c = 0; c < F; ++c) s = fmaf(ddfeat_s[b * F + c], Wf[(size_t)c * feat_dim + f], s);
    dT[o] = s;
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_attn_prep_fwd(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                      const float* bf, const float* a, const float* T, float* W_aug, float* q, void* stream) {
  if (H <= 0 || d <= 0 || in_dim <= 0 || feat_dim <= 0 || !W || !Wf || !a || !T || !W_aug || !q) return HSG_ERR_ARG;
  if (ld_rows < H * d + H) return HSG_ERR_SHAPE;
  const size_t smem = (size_t)HSG_N_BINS * H * d * sizeof(float);
  if (smem > 48 * 1024) return HSG_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_PREP_FWD, s);
  attn_prep_fwd_kernel<<<ld_rows + 1, 256, smem, s>>>(H, d, in_dim, feat_dim, ld_rows, W, Wf, bf, a, T, W_aug, q);
  return check_launch();
}

int hsg_attn_prep_bwd(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                      const float* bf, const float* a, const float* T, const float* dW_aug, const float* dq,
                      float* dW, float* dWf, float* dbf, float* da, float* dT, void* stream) {
  if (H <= 0 || d <= 0 || in_dim <= 0 || feat_dim <= 0 || !W || !Wf || !a || !T || !dW_aug || !dq || !dW || !dWf ||
      !da || !dT)
    return HSG_ERR_ARG;
  if (ld_rows < H * d + H) return HSG_ERR_SHAPE;
  const size_t smem = (size_t)2 * HSG_N_BINS * H * d * sizeof(float);
  if (smem > 48 * 1024) return HSG_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_PREP_BWD, s);
  attn_prep_bwd_kernel<<<H * d + 1, 256, smem < 64 ? 64 : smem, s>>>(H, d, in_dim, feat_dim, W, Wf, bf, a, T, dW_aug,
                                                                     dq, dW, dWf, dbf, da, dT);
  return check_launch();
}

}  // extern "C"
