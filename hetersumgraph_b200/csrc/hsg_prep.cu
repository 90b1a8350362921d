// K2: attention prep.  Folds attn_fc (GATLayer.py:87,126) into the node projection
// and into a 10 x H table over the TF-IDF boxes, so the edge kernel only adds two
// scalars per (edge, head):
//     p_u   = a_k[0:d] . z_u          -> extra rows of the projection weight (W_aug[FP+k])
//     q[b,k]= a_k[2d:3d] . (Wf_k T[b] + bf_k)       (feat_fc of the tfidfembed of box b, HiGraph.py:150-151)
// The middle third a_k[d:2d] multiplies the destination's `z`, which DGL zero-fills
// (GATLayer.py:111,147 write z on the source node type only), so it is dead: its
// gradient is exactly zero, but it stays in the state_dict.
//
// W_aug rows are stored in the lane-interleaved order of hsg_edge_layout.cuh, so the
// projection product writes z rows directly in the layout the edge kernels gather:
//     W_aug[perm(c)] = W[c]   (c < F),   W_aug[FP + k] = sum_j a_k[j] W[k d + j],   other rows 0.
#include "hsg_common.cuh"
#include "hsg_internal.cuh"
#include "hsg_edge_layout.cuh"

namespace hsg {

// grid = ld_rows + HSG_N_BINS blocks.  Block r < ld_rows writes row r of W_aug; block ld_rows + b writes q[b, :].
__global__ void __launch_bounds__(256)
attn_prep_fwd_kernel(EdgeLayout L, int in_dim, int feat_dim, int ld_rows, const float* __restrict__ W,
                     const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ a,
                     const float* __restrict__ T, float* __restrict__ W_aug, float* __restrict__ q) {
  pdl_prologue();
  extern __shared__ float dfeat_s[];  // [F] (q blocks only)
  const int H = L.H, d = L.D, F = H * d;
  const int r = blockIdx.x;
  if (r < ld_rows) {
    float* out = W_aug + (size_t)r * in_dim;
    const int c = edge_unperm(L, r);
    if (c >= 0) {
      for (int i = threadIdx.x; i < in_dim; i += blockDim.x) out[i] = W[(size_t)c * in_dim + i];
    } else if (r >= L.fp && r < L.fp + H) {
      const int k = r - L.fp;
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
  const int b = r - ld_rows;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float s = bf ? bf[c] : 0.f;
    for (int f = 0; f < feat_dim; ++f) s = fmaf(Wf[(size_t)c * feat_dim + f], T[b * feat_dim + f], s);
    dfeat_s[c] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < d; ++j) s = fmaf(a[k * 3 * d + 2 * d + j], dfeat_s[k * d + j], s);
    q[b * H + k] = s;
  }
}

__device__ __forceinline__ void put(float* p, float v, int accumulate) { *p = accumulate ? *p + v : v; }

constexpr int PREP_CCH = 32;   // feature columns per block in the feat-path backward

// grid = F + ceil(F/PREP_CCH) + 10 blocks:
//   block r < F                   : dW row r and da[k, j] (a_src part), da[k, d + j] = 0
//   block F + cb                  : columns [cb*32, cb*32+32): da[k, 2d+j], dbf, dWf rows
//   last 10 blocks                : dT[b, :], one block per TF-IDF box
__global__ void __launch_bounds__(256)
attn_prep_bwd_kernel(EdgeLayout L, int in_dim, int feat_dim, const float* __restrict__ W,
                     const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ a,
                     const float* __restrict__ T, const float* __restrict__ dW_aug, const float* __restrict__ dq,
                     float* __restrict__ dW, float* __restrict__ dWf, float* __restrict__ dbf,
                     float* __restrict__ da, float* __restrict__ dT, int acc_p, int acc_T) {
  pdl_prologue();
  __shared__ float sm[2 * PREP_CCH * HSG_N_BINS + 32];   // 672 floats (>= 512 for the dT blocks)
  const int H = L.H, d = L.D, F = H * d;
  const int r = blockIdx.x;
  const int ncb = ceil_div(F, PREP_CCH);
  if (r < F) {
    const int k = r / d, j = r % d;
    const float aj = a[k * 3 * d + j];
    const float* dwp = dW_aug + (size_t)(L.fp + k) * in_dim;
    const float* dwr = dW_aug + (size_t)edge_perm(L, r) * in_dim;
    float part = 0.f;
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
      const float g = dwp[i];
      put(dW + (size_t)r * in_dim + i, fmaf(aj, g, dwr[i]), acc_p);
      part = fmaf(g, W[(size_t)r * in_dim + i], part);
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];   // fixed order
      put(da + k * 3 * d + j, s, acc_p);
      put(da + k * 3 * d + d + j, 0.f, acc_p);   // a_dst multiplies DGL's zero-filled destination z
    }
    return;
  }
  if (r < F + ncb) {
    const int c0 = (r - F) * PREP_CCH;
    float* ddfeat_s = sm;                               // [PREP_CCH][10]  dq[b,k] * a_feat[c]
    float* prod_s = sm + PREP_CCH * HSG_N_BINS;         // [PREP_CCH][10]  dq[b,k] * dfeat[b,c]
    for (int o = threadIdx.x; o < PREP_CCH * HSG_N_BINS; o += blockDim.x) {
      const int cl = o / HSG_N_BINS, b = o % HSG_N_BINS, c = c0 + cl;
      float dd = 0.f, prod = 0.f;
      if (c < F) {
        float s = bf ? bf[c] : 0.f;
        for (int f = 0; f < feat_dim; ++f) s = fmaf(Wf[(size_t)c * feat_dim + f], T[b * feat_dim + f], s);
        const int k = c / d, j = c % d;
        const float dqv = dq[b * H + k];
        dd = dqv * a[k * 3 * d + 2 * d + j];
        prod = dqv * s;
      }
      ddfeat_s[o] = dd;
      prod_s[o] = prod;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < PREP_CCH; cl += blockDim.x) {
      const int c = c0 + cl;
      if (c >= F) continue;
      const int k = c / d, j = c % d;
      float sb = 0.f, sa = 0.f;
      for (int b = 0; b < HSG_N_BINS; ++b) {            // fixed order
        sa += prod_s[cl * HSG_N_BINS + b];
        sb += ddfeat_s[cl * HSG_N_BINS + b];
      }
      put(da + k * 3 * d + 2 * d + j, sa, acc_p);
      if (dbf) put(dbf + c, sb, acc_p);
    }
    for (int o = threadIdx.x; o < PREP_CCH * feat_dim; o += blockDim.x) {
      const int cl = o / feat_dim, f = o % feat_dim, c = c0 + cl;
      if (c >= F) continue;
      float s = 0.f;
      for (int b = 0; b < HSG_N_BINS; ++b) s = fmaf(ddfeat_s[cl * HSG_N_BINS + b], T[b * feat_dim + f], s);
      put(dWf + (size_t)c * feat_dim + f, s, acc_p);
    }
    return;
  }
  // dT[b, :] = sum_c ddfeat[b, c] Wf[c, :]   one block per TF-IDF box b, 4 column chunks x 64 features
  const int b = r - (F + ncb);
  float* dd_s = sm;                                      // [F] ddfeat[b, :]  (F <= 512 fits: 2*320+32 floats? no -> loop)
  __shared__ float part_s[4][64];
  const int f = threadIdx.x & 63, ch = threadIdx.x >> 6;   // 256 threads
  float acc = 0.f;
  for (int c0 = 0; c0 < F; c0 += 512) {
    const int cn = min(512, F - c0);
    __syncthreads();
    for (int c = threadIdx.x; c < cn; c += blockDim.x) {
      const int cc = c0 + c, k = cc / d, j = cc % d;
      dd_s[c] = dq[b * H + k] * a[k * 3 * d + 2 * d + j];
    }
    __syncthreads();
    if (f < feat_dim)
      for (int c = ch; c < cn; c += 4) acc = fmaf(dd_s[c], Wf[(size_t)(c0 + c) * feat_dim + f], acc);
  }
  part_s[ch][f] = acc;
  __syncthreads();
  if (ch == 0 && f < feat_dim)
    put(dT + b * feat_dim + f, (part_s[0][f] + part_s[1][f]) + (part_s[2][f] + part_s[3][f]), acc_T);
  for (int f2 = 64 + threadIdx.x; f2 < feat_dim; f2 += blockDim.x) {   // feat_dim > 64: plain loop
    float s2 = 0.f;
    for (int c = 0; c < F; ++c) {
      const int k = c / d, j = c % d;
      s2 = fmaf(dq[b * H + k] * a[k * 3 * d + 2 * d + j], Wf[(size_t)c * feat_dim + f2], s2);
    }
    put(dT + b * feat_dim + f2, s2, acc_T);
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_attn_prep_fwd(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                      const float* bf, const float* a, const float* T, float* W_aug, float* q, void* stream) {
  if (H <= 0 || d <= 0 || H > 32 || in_dim <= 0 || feat_dim <= 0 || !W || !Wf || !a || !T || !W_aug || !q) return HSG_ERR_ARG;
  const EdgeLayout L = make_edge_layout(H, d);
  if (ld_rows < L.fp + H) return HSG_ERR_SHAPE;
  const size_t smem = (size_t)H * d * sizeof(float);
  if (smem > 48 * 1024) return HSG_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_PREP_FWD, s);
  launch_k(attn_prep_fwd_kernel, dim3(ld_rows + HSG_N_BINS), dim3(256), smem, s, L, in_dim, feat_dim, ld_rows, W, Wf, bf, a, T, W_aug, q);
  return check_launch();
}

int hsg_attn_prep_bwd(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                      const float* bf, const float* a, const float* T, const float* dW_aug, const float* dq,
                      float* dW, float* dWf, float* dbf, float* da, float* dT, void* stream) {
  return attn_prep_bwd_ex(H, d, in_dim, feat_dim, ld_rows, W, Wf, bf, a, T, dW_aug, dq, dW, dWf, dbf, da, dT, 0, 0,
                          (cudaStream_t)stream);
}

}  // extern "C"

namespace hsg {
int attn_prep_bwd_ex(int H, int d, int in_dim, int feat_dim, int ld_rows, const float* W, const float* Wf,
                     const float* bf, const float* a, const float* T, const float* dW_aug, const float* dq, float* dW,
                     float* dWf, float* dbf, float* da, float* dT, int acc_params, int acc_T, cudaStream_t s) {
  if (H <= 0 || d <= 0 || H > 32 || in_dim <= 0 || feat_dim <= 0 || !W || !Wf || !a || !T || !dW_aug || !dq || !dW ||
      !dWf || !da || !dT)
    return HSG_ERR_ARG;
  const EdgeLayout L = make_edge_layout(H, d);
  if (ld_rows < L.fp + H) return HSG_ERR_SHAPE;
  LaunchScope ls(SLOT_PREP_BWD, s);
  const int F = H * d;
  launch_k(attn_prep_bwd_kernel, dim3(F + ceil_div(F, PREP_CCH) + HSG_N_BINS), dim3(256), 0, s, L, in_dim, feat_dim, W, Wf, bf, a, T, dW_aug, dq,
                                                                    dW, dWf, dbf, da, dT, acc_params, acc_T);
  return check_launch();
}

}  // namespace hsg
