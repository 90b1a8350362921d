// Sentence encoder in front of the WSWGAT path: the n-gram CNN of module/Encoder.py:56-76.
//
//   x[s, t, :]  = embed[tok[s, t]] + pos_table[t < len_s ? t + 1 : 0]                  Encoder.py:58-69
//   y_h[s, t, c] = b_h[c] + sum_{j < h} W_h[c, j, :] . x[s, t + j, :]   t = 0 .. L - h    six Conv2d(1, 50, (h, D)), :71
//   out[s, (h - 2) * 50 + c] = max_t relu(y_h[s, t, c])                                   :71-73
//
// B200 formulation
//   * the reference pads every sentence to L = 100 rows and convolves all of them (78 GFLOP forward at 1 009
//     sentences, ~3/4 of it on padding, plus a host loop with one device sync per sentence for the position ids,
//     Encoder.py:61-66).  Every window that lies entirely behind the last real token sees the same input row
//     (embed[0] + pos[0]) h times, so all those windows produce the same value: a sentence only needs its
//     n_s = min(tail_s + 7, L) first rows (tail_s = index after its last non-zero id).  The rows of all sentences are
//     stored back to back in ONE compact matrix Xc [R, D] (R = sum n_s, ~1/3 of S * L on CNN/DM-shaped text).
//   * a window of height h starting at compact row r is the contiguous slab Xc[r*D : (r+h)*D].  With every kernel
//     zero-padded to height 7, all six convolutions are the product
//         Y [R, 312] = A [R, 7D] . Wpad [312, 7D]^T ,   A = Xc viewed with row pitch D (overlapping rows)
//     on the tcgen05 GEMM (hsg_gemm_nt, TMA reads the overlapping rows straight from Xc: no im2col buffer).  Rows
//     r + h .. r + 6 meet zero weights; they may belong to the next sentence or to the zeroed 8-row tail of Xc.
//     The 50 channels of a height are padded to 52 columns so that every height group starts 16-byte aligned: the
//     caller issues the product as a few K-CHUNKS (kernel rows j = 0-1, 2-3, 4-5, 6) accumulated in place, each
//     over only the column groups whose kernels reach that row (31 % fewer flops than the dense [312, 7D] product,
//     and the tensor core's truncating fp32 accumulation chain stays at 75 updates: measured 5e-6 -> 1.5e-6).
//   * max over time runs over the VALID windows of a sentence only (t <= n_s - h); bias and ReLU commute with the max.
//   * backward: the embedding is frozen in the reference default (train.py:340-342), so only dW_h / db_h are needed.
//     d(out)/dY is one-hot per (sentence, channel), so dW_h[c] = sum_s g[s, c] * slab(argmax row) is a sparse
//     accumulation of <= S * 300 slabs (0.4 GFLOP) instead of a 40 GFLOP dense product; fixed summation order.
#include "hsg_common.cuh"

namespace hsg {

constexpr int ENC_KH = 6;          // kernel heights 2..7 (Encoder.py:36-37)
constexpr int ENC_CH = 50;         // channels per height (Encoder.py:35)
constexpr int ENC_F = ENC_KH * ENC_CH;
constexpr int ENC_CHP = 52;        // channels per height in the product's column layout (16-byte aligned groups)
constexpr int ENC_FP = ENC_KH * ENC_CHP;
constexpr int ENC_HMAX = 7;
constexpr int ENC_TAIL_ROWS = 8;   // zeroed rows behind Xc

// ---- compact input rows: one CTA per sentence -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
enc_gather_kernel(int n_sent, int L, int D, int n_rows, const int32_t* __restrict__ tokens,
                  const int32_t* __restrict__ sent_len, const int32_t* __restrict__ row_ptr,
                  const float* __restrict__ embed, const float* __restrict__ pos_table, float* __restrict__ xc) {
  pdl_prologue();
  const int s = blockIdx.x;
  const int d4 = D >> 2;
  if (s == n_sent) {                                       // zero tail: windows of the last rows read into it
    float4* t = reinterpret_cast<float4*>(xc + (size_t)n_rows * D);
    for (int i = threadIdx.x; i < ENC_TAIL_ROWS * d4; i += blockDim.x) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int r0 = row_ptr[s], n = row_ptr[s + 1] - r0;
  const int len = min(sent_len[s], L);
  for (int i = threadIdx.x; i < n * d4; i += blockDim.x) {
    const int t = i / d4, c = i - t * d4;
    const int tok = __ldg(tokens + (size_t)s * L + t);
    const float4 e = __ldg(reinterpret_cast<const float4*>(embed + (size_t)tok * D) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos_table + (size_t)(t < len ? t + 1 : 0) * D) + c);
    reinterpret_cast<float4*>(xc + (size_t)(r0 + t) * D)[c] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

// ---- Wpad [312, 7D]: row (h-2)*52 + c = [ W_h[c, 0, :, :] flattened (h*D) | zeros ], rows c = 50, 51 zero ----------
struct ConvPtrs {
  const float* w[ENC_KH];
};
struct ConvGradPtrs {
  float* w[ENC_KH];
  float* b[ENC_KH];
};

__global__ void __launch_bounds__(256) enc_pack_w_kernel(int D, ConvPtrs p, float* __restrict__ wpad) {
  pdl_prologue();
  const int K = ENC_HMAX * D;
  const int row = blockIdx.x;
  const int hi = row / ENC_CHP, c = row - hi * ENC_CHP, h = hi + 2;
  const float* src = p.w[hi] + (size_t)(c < ENC_CH ? c : 0) * h * D;
  for (int e = threadIdx.x; e < K; e += blockDim.x)
    wpad[(size_t)row * K + e] = (c < ENC_CH && e < h * D) ? __ldg(src + e) : 0.f;
}

// ---- max over the valid windows + bias + ReLU: one CTA per sentence, one thread per channel ------------------------
__global__ void __launch_bounds__(320)
enc_pool_fwd_kernel(int n_sent, const int32_t* __restrict__ row_ptr, const float* __restrict__ y, int ldy,
                    ConvPtrs bias, float* __restrict__ out, int ldo, int32_t* __restrict__ arg_t) {
  pdl_prologue();
  const int s = blockIdx.x, col = threadIdx.x;
  if (col >= ENC_F) return;
  const int hi = col / ENC_CH, h = hi + 2;
  const int r0 = row_ptr[s], n = row_ptr[s + 1] - r0;
  const int nwin = n - h + 1;                                // >= 1: n >= 7
  const float* p = y + (size_t)r0 * ldy + hi * ENC_CHP + (col - hi * ENC_CH);
  float best = -INFINITY;
  int at = 0;
  int t = 0;
  for (; t + 4 <= nwin; t += 4) {                            // 4 independent loads in flight
    const float v0 = __ldg(p + (size_t)t * ldy), v1 = __ldg(p + (size_t)(t + 1) * ldy);
    const float v2 = __ldg(p + (size_t)(t + 2) * ldy), v3 = __ldg(p + (size_t)(t + 3) * ldy);
    if (v0 > best) { best = v0; at = t; }
    if (v1 > best) { best = v1; at = t + 1; }
    if (v2 > best) { best = v2; at = t + 2; }
    if (v3 > best) { best = v3; at = t + 3; }
  }
  for (; t < nwin; ++t) {
    const float v = __ldg(p + (size_t)t * ldy);
    if (v > best) { best = v; at = t; }
  }
  const float r = best + __ldg(bias.w[hi] + (col - hi * ENC_CH));
  out[(size_t)s * ldo + col] = r > 0.f ? r : 0.f;
  arg_t[(size_t)col * n_sent + s] = r > 0.f ? r0 + at : -1;   // -1: ReLU inactive, no gradient (threshold backward)
}

// ---- weight / bias gradients: CTA (column, sentence split) accumulates g * slab in registers -----------------------
constexpr int WG_THREADS = 256;
constexpr int WG_NV = 3;                                     // float4 per thread: 7 * D / 4 <= 768  (D <= 438)

__global__ void __launch_bounds__(WG_THREADS)
enc_conv_wgrad_kernel(int n_sent, int D, const float* __restrict__ xc, const float* __restrict__ d_out, int ldo,
                      const int32_t* __restrict__ arg_t, int per_split, float* __restrict__ part_w,
                      float* __restrict__ part_b) {
  pdl_prologue();
  const int col = blockIdx.x, split = blockIdx.y;
  const int hi = col / ENC_CH, h = hi + 2;
  const int n4 = (h * D) >> 2;
  const int K4 = (ENC_HMAX * D) >> 2;
  __shared__ float g_s[WG_THREADS];
  __shared__ int r_s[WG_THREADS];
  float4 acc[WG_NV];
#pragma unroll
  for (int v = 0; v < WG_NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;
  const int s_beg = split * per_split, s_end = min(n_sent, s_beg + per_split);
  for (int c0 = s_beg; c0 < s_end; c0 += WG_THREADS) {
    const int cnt = min(WG_THREADS, s_end - c0);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const int r = __ldg(arg_t + (size_t)col * n_sent + c0 + threadIdx.x);
      r_s[threadIdx.x] = r < 0 ? 0 : r;
      g_s[threadIdx.x] = r < 0 ? 0.f : __ldg(d_out + (size_t)(c0 + threadIdx.x) * ldo + col);
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < cnt; ++i) bsum += g_s[i];
    int i = 0;
    for (; i + 4 <= cnt; i += 4) {                           // 4 sentences x WG_NV slabs in flight per thread
      float4 x[4][WG_NV];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4* slab = reinterpret_cast<const float4*>(xc + (size_t)r_s[i + u] * D);
#pragma unroll
        for (int v = 0; v < WG_NV; ++v) {
          const int idx = threadIdx.x + v * WG_THREADS;
          x[u][v] = idx < n4 ? __ldg(slab + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float g = g_s[i + u];
#pragma unroll
        for (int v = 0; v < WG_NV; ++v) {
          acc[v].x = fmaf(g, x[u][v].x, acc[v].x);
          acc[v].y = fmaf(g, x[u][v].y, acc[v].y);
          acc[v].z = fmaf(g, x[u][v].z, acc[v].z);
          acc[v].w = fmaf(g, x[u][v].w, acc[v].w);
        }
      }
    }
    for (; i < cnt; ++i) {
      const float g = g_s[i];
      const float4* slab = reinterpret_cast<const float4*>(xc + (size_t)r_s[i] * D);
#pragma unroll
      for (int v = 0; v < WG_NV; ++v) {
        const int idx = threadIdx.x + v * WG_THREADS;
        if (idx < n4) {
          const float4 xv = __ldg(slab + idx);
          acc[v].x = fmaf(g, xv.x, acc[v].x);
          acc[v].y = fmaf(g, xv.y, acc[v].y);
          acc[v].z = fmaf(g, xv.z, acc[v].z);
          acc[v].w = fmaf(g, xv.w, acc[v].w);
        }
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(part_w) + ((size_t)split * ENC_F + col) * K4;
#pragma unroll
  for (int v = 0; v < WG_NV; ++v) {
    const int idx = threadIdx.x + v * WG_THREADS;
    if (idx < n4) dst[idx] = acc[v];
  }
  if (threadIdx.x == 0) part_b[split * ENC_F + col] = bsum;
}

// fixed-order sum over the sentence splits, written (or accumulated) into the six dW_h / db_h tensors
__global__ void __launch_bounds__(256)
enc_conv_wgrad_reduce_kernel(int D, int nsplit, const float* __restrict__ part_w, const float* __restrict__ part_b,
                             ConvGradPtrs out, int accumulate) {
  pdl_prologue();
  const int col = blockIdx.x;
  const int hi = col / ENC_CH, c = col - hi * ENC_CH, h = hi + 2;
  const int K = ENC_HMAX * D;
  float* dw = out.w[hi] + (size_t)c * h * D;
  for (int e = threadIdx.x; e < h * D; e += blockDim.x) {
    float sres = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) sres += part_w[((size_t)sp * ENC_F + col) * K + e];
    dw[e] = accumulate ? dw[e] + sres : sres;
  }
  if (threadIdx.x == 0) {
    float sres = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) sres += part_b[sp * ENC_F + col];
    out.b[hi][c] = accumulate ? out.b[hi][c] + sres : sres;
  }
}

// out[r, :] = x[r, :] + table[idx[r], :]      (ngram_feature + sent_pos_embed(position), HiGraph.py:130-132)
__global__ void __launch_bounds__(256)
add_rows_kernel(int n, int D, const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx,
                const float* __restrict__ table, float* __restrict__ out, int ldo) {
  pdl_prologue();
  const int d4 = D >> 2;
  const long long total = (long long)n * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d4), c = (int)(i - (long long)r * d4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx) + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(table + (size_t)__ldg(idx + r) * D) + c);
    reinterpret_cast<float4*>(out + (size_t)r * ldo)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

static int wgrad_splits(int n_sent) {
  int sp = ceil_div(n_sent, 128);
  return sp < 1 ? 1 : (sp > 16 ? 16 : sp);
}

}  // namespace hsg

using namespace hsg;

extern "C" {

// Host-side plan of a batch (plain C loop over the host token matrix; no device work, no synchronisation).
int hsg_enc_plan_host(int n_sent, int L, const int32_t* tokens, int n_graphs, const int32_t* graph_sent_ptr,
                      int32_t* sent_len, int32_t* row_ptr, int32_t* sent_pos) {
  if (n_sent < 0 || L < ENC_HMAX || n_graphs < 0 || !row_ptr) return HSG_ERR_ARG;
  if (n_sent > 0 && (!tokens || !sent_len || !sent_pos)) return HSG_ERR_ARG;
  if (n_graphs > 0 && (!graph_sent_ptr || graph_sent_ptr[0] != 0 || graph_sent_ptr[n_graphs] != n_sent)) return HSG_ERR_ARG;
  long long rows = 0;
  row_ptr[0] = 0;
  for (int s = 0; s < n_sent; ++s) {
    const int32_t* t = tokens + (size_t)s * L;
    int len = 0, tail = 0;
    for (int i = 0; i < L; ++i) {
      if (t[i] != 0) {
        ++len;
        tail = i + 1;
      }
    }
    sent_len[s] = len;
    const int n = tail + ENC_HMAX < L ? tail + ENC_HMAX : L;
    rows += n;
    if (rows > 0x7fffffffLL) return HSG_ERR_CAPACITY;
    row_ptr[s + 1] = (int32_t)rows;
  }
  for (int g = 0; g < n_graphs; ++g) {
    const int b = graph_sent_ptr[g], e = graph_sent_ptr[g + 1];
    if (e < b || e > n_sent) return HSG_ERR_ARG;
    for (int s = b; s < e; ++s) sent_pos[s] = s - b + 1;       // dataloader.py:241
  }
  return HSG_OK;
}

int hsg_enc_gather(int n_sent, int L, int D, int n_rows, const int32_t* tokens, const int32_t* sent_len,
                   const int32_t* row_ptr, const float* embed, const float* pos_table, float* xc, void* stream) {
  if (n_sent < 0 || L < ENC_HMAX || D <= 0 || n_rows < 0) return HSG_ERR_ARG;
  if (D % 4 != 0) return HSG_ERR_SHAPE;
  if (!xc || !embed || !pos_table || (n_sent > 0 && (!tokens || !sent_len || !row_ptr))) return HSG_ERR_ARG;
  if (!aligned16(xc) || !aligned16(embed) || !aligned16(pos_table)) return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_ENCODER, s);
  launch_k(enc_gather_kernel, dim3(n_sent + 1), dim3(256), 0, s, n_sent, L, D, n_rows, tokens, sent_len, row_ptr, embed,
           pos_table, xc);
  return check_launch();
}

int hsg_enc_pack_weights(int D, const float* const* conv_w, float* wpad, void* stream) {
  if (D <= 0 || !conv_w || !wpad) return HSG_ERR_ARG;
  ConvPtrs p;
  for (int i = 0; i < ENC_KH; ++i) {
    if (!conv_w[i]) return HSG_ERR_ARG;
    p.w[i] = conv_w[i];
  }
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_ENCODER, s);
  launch_k(enc_pack_w_kernel, dim3(ENC_FP), dim3(256), 0, s, D, p, wpad);
  return check_launch();
}

int hsg_enc_pool_fwd(int n_sent, const int32_t* row_ptr, const float* y, int ldy, const float* const* conv_b,
                     float* out, int ldo, int32_t* arg_t, void* stream) {
  if (n_sent < 0 || ldy < ENC_FP || ldo < ENC_F) return HSG_ERR_ARG;
  if (n_sent == 0) return HSG_OK;
  if (!row_ptr || !y || !conv_b || !out || !arg_t) return HSG_ERR_ARG;
  ConvPtrs p;
  for (int i = 0; i < ENC_KH; ++i) {
    if (!conv_b[i]) return HSG_ERR_ARG;
    p.w[i] = conv_b[i];
  }
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_ENCODER, s);
  launch_k(enc_pool_fwd_kernel, dim3(n_sent), dim3(320), 0, s, n_sent, row_ptr, y, ldy, p, out, ldo, arg_t);
  return check_launch();
}

size_t hsg_enc_conv_wgrad_workspace_bytes(int n_sent, int D) {
  return (size_t)wgrad_splits(n_sent) * ENC_F * ((size_t)ENC_HMAX * D + 1) * sizeof(float) + 256;
}

int hsg_enc_conv_wgrad(int n_sent, int D, const float* xc, const float* d_out, int ldo, const int32_t* arg_t,
                       float* const* d_conv_w, float* const* d_conv_b, int accumulate, void* ws, size_t ws_bytes,
                       void* stream) {
  if (n_sent < 0 || D <= 0 || !d_conv_w || !d_conv_b) return HSG_ERR_ARG;
  if (D % 4 != 0 || ENC_HMAX * D > 4 * WG_THREADS * WG_NV) return HSG_ERR_SHAPE;
  if (n_sent > 0 && (!xc || !d_out || !arg_t)) return HSG_ERR_ARG;
  if (!ws || ws_bytes < hsg_enc_conv_wgrad_workspace_bytes(n_sent, D)) return HSG_ERR_WORKSPACE;
  if (!aligned16(ws) || (n_sent > 0 && !aligned16(xc))) return HSG_ERR_ALIGN;
  ConvGradPtrs g;
  for (int i = 0; i < ENC_KH; ++i) {
    if (!d_conv_w[i] || !d_conv_b[i]) return HSG_ERR_ARG;
    g.w[i] = d_conv_w[i];
    g.b[i] = d_conv_b[i];
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int nsplit = wgrad_splits(n_sent);
  const int per_split = ceil_div(n_sent > 0 ? n_sent : 1, nsplit);
  float* part_w = static_cast<float*>(ws);
  float* part_b = part_w + (size_t)nsplit * ENC_F * ENC_HMAX * D;
  {
    LaunchScope ls(SLOT_ENCODER, s);
    launch_k(enc_conv_wgrad_kernel, dim3(ENC_F, nsplit), dim3(WG_THREADS), 0, s, n_sent, D, xc, d_out, ldo, arg_t,
             per_split, part_w, part_b);
  }
  {
    LaunchScope ls(SLOT_ENCODER, s);
    launch_k(enc_conv_wgrad_reduce_kernel, dim3(ENC_F), dim3(256), 0, s, D, nsplit, (const float*)part_w,
             (const float*)part_b, g, accumulate);
  }
  return check_launch();
}

int hsg_add_rows(int n, int D, const float* x, int ldx, const int32_t* idx, const float* table, float* out, int ldo,
                 void* stream) {
  if (n < 0 || D <= 0) return HSG_ERR_ARG;
  if (n == 0) return HSG_OK;
  if (!x || !idx || !table || !out) return HSG_ERR_ARG;
  if (D % 4 != 0 || ldx % 4 != 0 || ldo % 4 != 0) return HSG_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(table) || !aligned16(out)) return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_ENCODER, s);
  const long long total = (long long)n * (D / 4);
  const int blocks = (int)(ceil_div_ll(total, 256) < 148 * 8 ? ceil_div_ll(total, 256) : 148 * 8);
  launch_k(add_rows_kernel, dim3(blocks), dim3(256), 0, s, n, D, x, ldx, idx, table, out, ldo);
  return check_launch();
}

}  // extern "C"
