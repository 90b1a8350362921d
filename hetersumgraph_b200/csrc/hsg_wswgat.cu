// Coarse-grained entry points: one C call launches every kernel of one WSWGAT application
// (module/GAT.py:45-59) forward or backward.  Host-side sequencing only - the kernels are the ones of
// hsg_gemm*.cu / hsg_edge.cu / hsg_ffn.cu.  This removes the per-kernel Python/ctypes round trips that
// dominate small batches (about 10 us per call against 5-60 us of GPU work per kernel).
#include "hsg_common.cuh"
#include "hsg_internal.cuh"

using namespace hsg;

#define HSG_TRY(expr)      \
  do {                     \
    int rc_ = (expr);      \
    if (rc_) return rc_;   \
  } while (0)

extern "C" {

int hsg_wswgat_fwd(const hsg_wswgat_fwd_args* a, void* stream) {
  if (!a || !a->csc || !a->neighbor || !a->origin || !a->W_aug || !a->q || !a->w1 || !a->b1 || !a->w2 || !a->b2 ||
      !a->gamma || !a->beta || !a->zp || !a->sh || !a->x || !a->stat || !a->hdn || !a->r || !a->ln_stats || !a->out)
    return HSG_ERR_ARG;
  const int F = a->H * a->d;
  // zp = neighbor . W_aug^T                          (fc of all heads + p = a_src . z, GATLayer.py:110,146)
  HSG_TRY(hsg_gemm_nt(a->n_src, a->ldz, a->in_dim, a->neighbor, a->in_dim, a->W_aug, a->in_dim, a->zp, a->ldz,
                      nullptr, nullptr, 0, 0, stream));
  // sh, x = elu(sh) + origin, stat                   (GATLayer.py:88-102,112-113; GAT.py:56-57)
  HSG_TRY(hsg_edge_fwd(a->csc, a->H, a->d, a->zp, a->ldz, a->q, a->origin, a->sh, a->x, a->stat, stream));
  // FFN                                              (GATLayer.py:35-44)
  if (ffn_rows_ok(a->n_dst, F, a->d_hid))            // small destination set: one launch (same choice as hsg_loop.cu)
    return ffn_rows_fwd(a->n_dst, F, a->d_hid, a->x, a->w1, a->b1, a->w2, a->b2, a->gamma, a->beta, a->hdn, a->r, a->out,
                        a->ln_stats, (cudaStream_t)stream);
  HSG_TRY(hsg_gemm_nt(a->n_dst, a->d_hid, F, a->x, F, a->w1, F, a->hdn, a->d_hid, a->b1, nullptr, 0,
                      HSG_EPI_BIAS | HSG_EPI_RELU, stream));
  HSG_TRY(hsg_gemm_nt(a->n_dst, F, a->d_hid, a->hdn, a->d_hid, a->w2, a->d_hid, a->r, F, a->b2, a->x, F,
                      HSG_EPI_BIAS | HSG_EPI_ADD, stream));
  return hsg_layernorm_fwd(a->n_dst, F, a->r, a->gamma, a->beta, a->out, a->ln_stats, stream);
}

size_t hsg_wswgat_bwd_workspace_bytes(int H, int d, int in_dim, int d_hid, int n_src, int n_dst) {
  const int F = H * d;
  int fp = 0, ldz = 0;
  if (hsg_edge_layout(H, d, &fp, &ldz) != HSG_OK) return 0;
  size_t w = hsg_layernorm_bwd_workspace_bytes(n_dst, F);
  size_t t = hsg_gemm_tn_workspace_bytes(n_dst, F, d_hid);
  if (t > w) w = t;
  t = hsg_gemm_tn_workspace_bytes(n_dst, d_hid, F);
  if (t > w) w = t;
  t = hsg_gemm_tn_workspace_bytes(n_src, ldz, in_dim);
  if (t > w) w = t;
  t = hsg_edge_bwd_workspace_bytes(H);
  if (t > w) w = t;
  return w;
}

int hsg_wswgat_bwd(const hsg_wswgat_bwd_args* a, void* stream) {
  if (!a || !a->csc_t || !a->dout || !a->neighbor || !a->W_aug || !a->q || !a->w1 || !a->w2 || !a->gamma || !a->zp ||
      !a->sh || !a->x || !a->stat || !a->hdn || !a->r || !a->ln_stats || !a->dr || !a->dhp || !a->dx || !a->g ||
      !a->dzp || !a->dq || !a->d_neighbor || !a->dW_aug || !a->dw1 || !a->db1 || !a->dw2 || !a->db2 || !a->dgamma ||
      !a->dbeta || !a->ws)
    return HSG_ERR_ARG;
  const int F = a->H * a->d;
  if (a->ws_bytes < hsg_wswgat_bwd_workspace_bytes(a->H, a->d, a->in_dim, a->d_hid, a->n_src, a->n_dst))
    return HSG_ERR_WORKSPACE;
  const bool rows_kernel = ffn_rows_ok(a->n_dst, F, a->d_hid);
  if (rows_kernel) {
    // small destination set: LayerNorm backward, dhp and dx in one launch
    HSG_TRY(ffn_rows_bwd(a->n_dst, F, a->d_hid, a->dout, a->r, a->ln_stats, a->gamma, a->hdn, a->w1, a->w2, a->dr, a->dhp,
                         a->dx, a->dgamma, a->dbeta, a->ws, a->ws_bytes, 0, (cudaStream_t)stream));
  } else {
    // LayerNorm
    HSG_TRY(hsg_layernorm_bwd(a->n_dst, F, a->dout, a->r, a->ln_stats, a->gamma, a->dr, a->dgamma, a->dbeta, a->ws,
                              a->ws_bytes, stream));
    // FFN: dhp = (dr . W2) * relu', dW2 = dr^T hdn, dW1 = dhp^T x, dx = dhp . W1 + dr
    HSG_TRY(hsg_gemm_nn(a->n_dst, a->d_hid, F, a->dr, F, a->w2, a->d_hid, a->dhp, a->d_hid, a->hdn, a->d_hid,
                        HSG_EPI_RELU_MASK, stream));
  }
  HSG_TRY(hsg_gemm_tn(a->n_dst, F, a->d_hid, a->dr, F, a->hdn, a->d_hid, a->dw2, a->d_hid, a->db2, a->ws, a->ws_bytes,
                      stream));
  HSG_TRY(hsg_gemm_tn(a->n_dst, a->d_hid, F, a->dhp, a->d_hid, a->x, F, a->dw1, F, a->db1, a->ws, a->ws_bytes, stream));
  if (!rows_kernel)
    HSG_TRY(hsg_gemm_nn(a->n_dst, F, a->d_hid, a->dhp, a->d_hid, a->w1, F, a->dx, F, a->dr, F, HSG_EPI_ADD, stream));
  // edge backward (d origin = dx, GAT.py:57)
  HSG_TRY(hsg_edge_bwd_prep(a->n_dst, a->H, a->d, a->dx, nullptr, a->sh, a->g, a->stat, stream));
  HSG_TRY(hsg_edge_bwd(a->csc_t, a->H, a->d, a->zp, a->ldz, a->q, a->g, a->stat, a->dzp, a->dq, a->ws, a->ws_bytes,
                       stream));
  // projection backward
  HSG_TRY(hsg_gemm_nn(a->n_src, a->in_dim, a->ldz, a->dzp, a->ldz, a->W_aug, a->in_dim, a->d_neighbor, a->in_dim,
                      nullptr, 0, 0, stream));
  return hsg_gemm_tn(a->n_src, a->ldz, a->in_dim, a->dzp, a->ldz, a->neighbor, a->in_dim, a->dW_aug, a->in_dim, nullptr,
                     a->ws, a->ws_bytes, stream);
}

}  // extern "C"
