// Whole update loop of HSumGraph.forward / HSumDocGraph.forward (HiGraph.py:98-106, 205-214) from ONE C call
// each way.  Host-side sequencing only; the kernels are those of hsg_prep.cu / hsg_gemm*.cu / hsg_edge.cu /
// hsg_ffn.cu.  What it buys over per-application calls (hsg_wswgat_fwd/bwd):
//   * ~45 launches are enqueued back to back from C (2-3 us each) instead of through Python/ctypes/autograd;
//   * gradients of the shared weights are accumulated by the last stage of the producing kernels
//     (gemm_tn reduce, LayerNorm reduce, dq reduce, attention-prep backward) - no add launches;
//   * the two gradient paths that meet at a node state (GAT.py:47-57: `origin` of the later application,
//     `neighbor` of the earlier one) are summed by the projection-backward epilogue (HSG_EPI_ADD).
//
// Applications: app 0 = W2S(word_feature, super_feature); then for it < n_iter
//   app 2it+1 = S2W(origin = word state, neighbor = super state),  app 2it+2 = W2S(neighbor = new word state,
//   origin = super state).  Even apps are W2S (destinations = supernodes), odd apps S2W (destinations = words).
#include "hsg_common.cuh"
#include "hsg_edge_layout.cuh"
#include "hsg_internal.cuh"

using namespace hsg;

#define HSG_TRY(expr)    \
  do {                   \
    int rc_ = (expr);    \
    if (rc_) return rc_; \
  } while (0)

namespace {

inline size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }

struct AppOff {
  size_t out, zp, sh, x, stat, hdn, r, ln, total;
};

// per-application block of the forward arena for n_src sources / n_dst destinations
AppOff app_layout(const hsg_layer_params& P, int ldz, int n_src, int n_dst) {
  const size_t F = (size_t)P.H * P.d;
  AppOff o;
  size_t off = 0;
  o.out = off;  off += r4((size_t)n_dst * F);
  o.zp = off;   off += r4((size_t)n_src * ldz);
  o.sh = off;   off += r4((size_t)n_dst * F);
  o.x = off;    off += r4((size_t)n_dst * F);
  o.stat = off; off += r4((size_t)n_dst * 3 * P.H);
  o.hdn = off;  off += r4((size_t)n_dst * P.d_hid);
  o.r = off;    off += r4((size_t)n_dst * F);
  o.ln = off;   off += r4((size_t)n_dst * 2);
  o.total = off;
  return o;
}

struct Layout {
  int n_apps;
  int ldz_ws, ldz_sw, fp_ws, fp_sw;
  size_t waug_ws, q_ws, waug_sw, q_sw, apps_base;
  AppOff ws_app, sw_app;   // block shapes (offsets relative to the block start)
  size_t state_total;
  // backward scratch
  size_t dWaug_ws, dq_ws, dWaug_sw, dq_sw;
  size_t s_zero;
  size_t s_dr, s_dhp, s_g, s_dx, s_dzp, s_gstate;   // W2S backward: dst = supernodes, dzp / gstate over words
  size_t w_dr, w_dhp, w_g, w_dx, w_dzp, w_gstate;   // S2W backward: dst = words, dzp / gstate over supernodes
  size_t scratch_total;
  size_t ws_bytes;

  size_t app_start(int i) const {   // i-th application block
    const int n_ws_before = (i + 1) / 2, n_sw_before = i / 2;
    return apps_base + (size_t)n_ws_before * ws_app.total + (size_t)n_sw_before * sw_app.total;
  }
};

int make_layout(const hsg_loop_args* a, Layout* L) {
  if (!a || a->n_iter < 0 || a->n_word < 0 || a->n_super < 0) return HSG_ERR_ARG;
  const hsg_layer_params& ws = a->w2s;
  const hsg_layer_params& sw = a->s2w;
  int fp = 0, ldz = 0;
  HSG_TRY(hsg_edge_layout(ws.H, ws.d, &fp, &ldz));
  L->fp_ws = fp;
  L->ldz_ws = ldz;
  if (ws.in_dim <= 0 || ws.feat_dim <= 0 || ws.d_hid <= 0) return HSG_ERR_ARG;
  const bool has_sw = a->n_iter > 0;
  if (has_sw) {
    HSG_TRY(hsg_edge_layout(sw.H, sw.d, &fp, &ldz));
    L->fp_sw = fp;
    L->ldz_sw = ldz;
    if (sw.in_dim != ws.H * ws.d || ws.in_dim != sw.H * sw.d || sw.feat_dim != ws.feat_dim || sw.d_hid <= 0)
      return HSG_ERR_SHAPE;   // the two layers map word dim <-> hidden dim and share the TF-IDF table
  } else {
    L->fp_sw = L->ldz_sw = 0;
  }
  L->n_apps = 1 + 2 * a->n_iter;
  size_t off = 0;
  L->waug_ws = off; off += r4((size_t)L->ldz_ws * ws.in_dim);
  L->q_ws = off;    off += r4((size_t)HSG_N_BINS * ws.H);
  L->waug_sw = off; off += has_sw ? r4((size_t)L->ldz_sw * sw.in_dim) : 0;
  L->q_sw = off;    off += has_sw ? r4((size_t)HSG_N_BINS * sw.H) : 0;
  L->apps_base = off;
  L->ws_app = app_layout(ws, L->ldz_ws, a->n_word, a->n_super);
  if (has_sw) L->sw_app = app_layout(sw, L->ldz_sw, a->n_super, a->n_word);
  else L->sw_app = AppOff{0, 0, 0, 0, 0, 0, 0, 0, 0};
  L->state_total = L->app_start(L->n_apps);
  if (L->state_total < 4) L->state_total = 4;

  const size_t Fs = (size_t)ws.H * ws.d, Fw = has_sw ? (size_t)sw.H * sw.d : (size_t)ws.in_dim;
  const size_t Ns = (size_t)a->n_super, Nw = (size_t)a->n_word;
  off = 0;
  L->dWaug_ws = off; off += r4((size_t)L->ldz_ws * ws.in_dim);
  L->dq_ws = off;    off += r4((size_t)HSG_N_BINS * ws.H);
  L->dWaug_sw = off; off += has_sw ? r4((size_t)L->ldz_sw * sw.in_dim) : 0;
  L->dq_sw = off;    off += has_sw ? r4((size_t)HSG_N_BINS * sw.H) : 0;
  L->s_zero = off;   off += r4(Ns * Fs);
  L->s_dr = off;     off += r4(Ns * Fs);
  L->s_dhp = off;    off += r4(Ns * ws.d_hid);
  L->s_g = off;      off += r4(Ns * L->fp_ws);
  L->s_dx = off;     off += r4(Ns * Fs);
  L->s_dzp = off;    off += r4(Nw * L->ldz_ws);
  L->s_gstate = off; off += r4(Nw * Fw);
  if (has_sw) {
    L->w_dr = off;     off += r4(Nw * Fw);
    L->w_dhp = off;    off += r4(Nw * sw.d_hid);
    L->w_g = off;      off += r4(Nw * L->fp_sw);
    L->w_dx = off;     off += r4(Nw * Fw);
    L->w_dzp = off;    off += r4(Ns * L->ldz_sw);
    L->w_gstate = off; off += r4(Ns * Fs);
  } else {
    L->w_dr = L->w_dhp = L->w_g = L->w_dx = L->w_dzp = L->w_gstate = off;
  }
  L->scratch_total = off < 4 ? 4 : off;

  size_t w = hsg_wswgat_bwd_workspace_bytes(ws.H, ws.d, ws.in_dim, ws.d_hid, a->n_word, a->n_super);
  if (has_sw) {
    const size_t t = hsg_wswgat_bwd_workspace_bytes(sw.H, sw.d, sw.in_dim, sw.d_hid, a->n_super, a->n_word);
    if (t > w) w = t;
  }
  L->ws_bytes = w;
  return HSG_OK;
}

bool params_ok(const hsg_layer_params& P) {
  return P.W && P.Wf && P.a && P.w1 && P.b1 && P.w2 && P.b2 && P.gamma && P.beta;
}

bool grads_ok(const hsg_layer_grads& G, bool need_bf) {
  return G.dW && G.dWf && G.da && G.dw1 && G.db1 && G.dw2 && G.db2 && G.dgamma && G.dbeta && (!need_bf || G.dbf);
}

// backward of one application.  dnb == nullptr skips the d_neighbor product; dnb_add (may be nullptr) is summed into it.
int app_bwd(const hsg_layer_params& P, int ldz, const hsg_csc* csc_t, int n_src, int n_dst, const float* neighbor,
            const float* W_aug, const float* q, float* blk, const AppOff& o, const float* dout, float* dr, float* dhp,
            float* g, float* dzp, float* dx, float* dnb, const float* dnb_add, float* dW_aug, float* dq, int acc_aug,
            const hsg_layer_grads& G, int acc_ffn, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int F = P.H * P.d;
  const float *zp = blk + o.zp, *sh = blk + o.sh, *x = blk + o.x, *hdn = blk + o.hdn, *r = blk + o.r,
              *ln = blk + o.ln;
  float* stat = blk + o.stat;
  // LayerNorm
  HSG_TRY(layernorm_bwd_ex(n_dst, F, dout, r, ln, P.gamma, dr, G.dgamma, G.dbeta, ws, ws_bytes, acc_ffn, s));
  // FFN: dhp = (dr . W2) * relu', dW2 = dr^T hdn, dW1 = dhp^T x, dx = dhp . W1 + dr
  HSG_TRY(hsg_gemm_nn(n_dst, P.d_hid, F, dr, F, P.w2, P.d_hid, dhp, P.d_hid, hdn, P.d_hid, HSG_EPI_RELU_MASK, s));
  HSG_TRY(gemm_tn_ex(n_dst, F, P.d_hid, dr, F, hdn, P.d_hid, G.dw2, P.d_hid, G.db2, ws, ws_bytes, acc_ffn, s));
  HSG_TRY(gemm_tn_ex(n_dst, P.d_hid, F, dhp, P.d_hid, x, F, G.dw1, F, G.db1, ws, ws_bytes, acc_ffn, s));
  HSG_TRY(hsg_gemm_nn(n_dst, F, P.d_hid, dhp, P.d_hid, P.w1, F, dx, F, dr, F, HSG_EPI_ADD, s));
  // edge backward (d origin = dx, GAT.py:57)
  HSG_TRY(hsg_edge_bwd_prep(n_dst, P.H, P.d, dx, nullptr, sh, g, stat, s));
  HSG_TRY(edge_bwd_ex(csc_t, P.H, P.d, zp, ldz, q, g, stat, dzp, dq, ws, ws_bytes, acc_aug, s));
  // projection backward
  if (dnb)
    HSG_TRY(hsg_gemm_nn(n_src, P.in_dim, ldz, dzp, ldz, W_aug, P.in_dim, dnb, P.in_dim, dnb_add, P.in_dim,
                        dnb_add ? HSG_EPI_ADD : 0, s));
  return gemm_tn_ex(n_src, ldz, P.in_dim, dzp, ldz, neighbor, P.in_dim, dW_aug, P.in_dim, nullptr, ws, ws_bytes,
                    acc_aug, s);
}

}  // namespace

extern "C" {

int hsg_update_loop_plan(const hsg_loop_args* a, hsg_loop_plan* plan) {
  if (!plan) return HSG_ERR_ARG;
  Layout L;
  HSG_TRY(make_layout(a, &L));
  plan->state_floats = L.state_total;
  plan->scratch_floats = L.scratch_total;
  plan->ws_bytes = L.ws_bytes;
  plan->super_state_off = L.app_start(L.n_apps - 1) + L.ws_app.out;
  plan->word_state_off = a->n_iter > 0 ? L.app_start(L.n_apps - 2) + L.sw_app.out : (size_t)-1;
  plan->hdn_off[0] = L.app_start(0) + L.ws_app.hdn;
  plan->hdn_off[1] = a->n_iter > 0 ? L.app_start(1) + L.sw_app.hdn : (size_t)-1;
  plan->pair_stride = L.ws_app.total + L.sw_app.total;
  return HSG_OK;
}

int hsg_update_loop_fwd(const hsg_loop_args* a, void* stream) {
  Layout L;
  HSG_TRY(make_layout(a, &L));
  if (!a->csc_super || !a->csc_word || !a->T || !a->word_feature || !a->super_feature || !a->state ||
      !params_ok(a->w2s) || (a->n_iter > 0 && !params_ok(a->s2w)))
    return HSG_ERR_ARG;
  if (a->state_floats < L.state_total) return HSG_ERR_WORKSPACE;
  if (!aligned16(a->state) || !aligned16(a->word_feature) || !aligned16(a->super_feature)) return HSG_ERR_ALIGN;
  if (a->csc_super->n_dst != a->n_super || a->csc_super->n_src != a->n_word || a->csc_word->n_dst != a->n_word ||
      a->csc_word->n_src != a->n_super)
    return HSG_ERR_SHAPE;
  float* st = a->state;
  const hsg_layer_params& ws = a->w2s;
  const hsg_layer_params& sw = a->s2w;
  // attention prep: once per layer (parameters only)
  HSG_TRY(hsg_attn_prep_fwd(ws.H, ws.d, ws.in_dim, ws.feat_dim, L.ldz_ws, ws.W, ws.Wf, ws.bf, ws.a, a->T,
                            st + L.waug_ws, st + L.q_ws, stream));
  if (a->n_iter > 0)
    HSG_TRY(hsg_attn_prep_fwd(sw.H, sw.d, sw.in_dim, sw.feat_dim, L.ldz_sw, sw.W, sw.Wf, sw.bf, sw.a, a->T,
                              st + L.waug_sw, st + L.q_sw, stream));
  const float* word = a->word_feature;
  const float* sup = a->super_feature;
  for (int i = 0; i < L.n_apps; ++i) {
    const bool w2s = (i & 1) == 0;
    const hsg_layer_params& P = w2s ? ws : sw;
    const AppOff& o = w2s ? L.ws_app : L.sw_app;
    float* blk = st + L.app_start(i);
    hsg_wswgat_fwd_args f;
    f.H = P.H; f.d = P.d; f.in_dim = P.in_dim; f.d_hid = P.d_hid;
    f.n_src = w2s ? a->n_word : a->n_super;
    f.n_dst = w2s ? a->n_super : a->n_word;
    f.ldz = w2s ? L.ldz_ws : L.ldz_sw;
    f.reserved = 0;
    f.csc = w2s ? a->csc_super : a->csc_word;
    f.neighbor = w2s ? word : sup;
    f.origin = w2s ? sup : word;
    f.W_aug = st + (w2s ? L.waug_ws : L.waug_sw);
    f.q = st + (w2s ? L.q_ws : L.q_sw);
    f.w1 = P.w1; f.b1 = P.b1; f.w2 = P.w2; f.b2 = P.b2; f.gamma = P.gamma; f.beta = P.beta;
    f.zp = blk + o.zp; f.sh = blk + o.sh; f.x = blk + o.x; f.stat = blk + o.stat; f.hdn = blk + o.hdn;
    f.r = blk + o.r; f.ln_stats = blk + o.ln; f.out = blk + o.out;
    HSG_TRY(hsg_wswgat_fwd(&f, stream));
    if (w2s) sup = f.out; else word = f.out;
  }
  return HSG_OK;
}

int hsg_update_loop_bwd(const hsg_loop_args* a, const hsg_loop_bwd_args* b, void* stream) {
  Layout L;
  HSG_TRY(make_layout(a, &L));
  if (!b || !a->csc_super || !a->csc_word || !a->T || !a->word_feature || !a->super_feature || !a->state ||
      !params_ok(a->w2s) || (a->n_iter > 0 && !params_ok(a->s2w)))
    return HSG_ERR_ARG;
  if (!b->d_super_feature || !b->dT || !b->scratch || !b->ws || !grads_ok(b->w2s, false) ||
      (a->n_iter > 0 && !grads_ok(b->s2w, a->s2w.bf != nullptr)) || (a->w2s.bf && !b->w2s.dbf))
    return HSG_ERR_ARG;
  if (!b->d_super_state && !b->d_word_state) return HSG_ERR_ARG;
  if (a->state_floats < L.state_total || b->scratch_floats < L.scratch_total || b->ws_bytes < L.ws_bytes)
    return HSG_ERR_WORKSPACE;
  if (!aligned16(b->scratch) || !aligned16(b->d_super_feature) || (b->d_word_feature && !aligned16(b->d_word_feature)) ||
      (b->d_super_state && !aligned16(b->d_super_state)) || (b->d_word_state && !aligned16(b->d_word_state)))
    return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  float* st = a->state;
  float* sc = b->scratch;
  const hsg_layer_params& ws = a->w2s;
  const hsg_layer_params& sw = a->s2w;
  const int acc = b->accumulate ? 1 : 0;
  const size_t Fs = (size_t)ws.H * ws.d, Fw = (size_t)ws.in_dim;
  const size_t ns_fs = (size_t)a->n_super * Fs;
  (void)Fw;

  // gradient w.r.t. the final supernode state; a missing one (only d_word_state given) is a zero buffer
  const float* g_sup = b->d_super_state;
  if (!g_sup) {
    if (cudaMemsetAsync(sc + L.s_zero, 0, ns_fs * sizeof(float), s) != cudaSuccess) return HSG_ERR_CUDA;
    g_sup = sc + L.s_zero;
  }
  const float* g_word_ext = b->d_word_state;   // gradient w.r.t. the final word state from outside (may be NULL)

  int n_ws_done = 0, n_sw_done = 0;
  const float* dx_w_prev = nullptr;   // origin-path gradient of the word state from the later S2W application
  const float* dx_s_prev = nullptr;   // origin-path gradient of the supernode state from the later W2S application
  for (int i = L.n_apps - 1; i >= 0; --i) {
    const bool w2s = (i & 1) == 0;
    float* blk = st + L.app_start(i);
    if (w2s) {
      // neighbor = word state produced by app i-1 (or word_feature for app 0), origin = supernode state before it
      const float* neighbor = i == 0 ? a->word_feature : st + L.app_start(i - 1) + L.sw_app.out;
      // dout: upstream gradient for the last application, else written by the S2W backward of app i+1
      // (its d_neighbor + the dx of app i+2)
      const float* dout = (i == L.n_apps - 1) ? g_sup : sc + L.w_gstate;
      // d_neighbor: gradient of the word state (app i-1's output) = this + dx of the S2W app i+1 (origin path), or,
      // for the last application, + the external d_word_state
      float* dnb;
      const float* dnb_add;
      if (i == 0) {
        dnb = b->d_word_feature;                              // may be NULL: frozen embedding
        dnb_add = a->n_iter > 0 ? dx_w_prev : g_word_ext;     // origin path of app 1 / pass-through when n_iter == 0
      } else {
        dnb = sc + L.s_gstate;
        dnb_add = (i == L.n_apps - 1) ? g_word_ext : dx_w_prev;
      }
      float* dx = (i == 0) ? b->d_super_feature : sc + L.s_dx;
      HSG_TRY(app_bwd(ws, L.ldz_ws, a->csc_word, a->n_word, a->n_super, neighbor, st + L.waug_ws, st + L.q_ws, blk,
                      L.ws_app, dout, sc + L.s_dr, sc + L.s_dhp, sc + L.s_g, sc + L.s_dzp, dx, dnb, dnb_add,
                      sc + L.dWaug_ws, sc + L.dq_ws, n_ws_done > 0, b->w2s, acc || n_ws_done > 0, b->ws, b->ws_bytes, s));
      ++n_ws_done;
      dx_s_prev = dx;
    } else {
      // S2W app i: neighbor = supernode state from app i-1, origin = word state from app i-2 (or word_feature)
      const float* neighbor = st + L.app_start(i - 1) + L.ws_app.out;
      const float* dout = sc + L.s_gstate;                     // written by the W2S backward of app i+1
      HSG_TRY(app_bwd(sw, L.ldz_sw, a->csc_super, a->n_super, a->n_word, neighbor, st + L.waug_sw, st + L.q_sw, blk,
                      L.sw_app, dout, sc + L.w_dr, sc + L.w_dhp, sc + L.w_g, sc + L.w_dzp, sc + L.w_dx,
                      sc + L.w_gstate, dx_s_prev, sc + L.dWaug_sw, sc + L.dq_sw, n_sw_done > 0, b->s2w,
                      acc || n_sw_done > 0, b->ws, b->ws_bytes, s));
      ++n_sw_done;
      dx_w_prev = sc + L.w_dx;
    }
  }
  // attention-prep backward: (dW_aug, dq) summed over the applications -> fc / feat_fc / attn_fc / TF-IDF table
  HSG_TRY(attn_prep_bwd_ex(ws.H, ws.d, ws.in_dim, ws.feat_dim, L.ldz_ws, ws.W, ws.Wf, ws.bf, ws.a, a->T,
                           sc + L.dWaug_ws, sc + L.dq_ws, b->w2s.dW, b->w2s.dWf, b->w2s.dbf, b->w2s.da, b->dT, acc, acc,
                           s));
  if (a->n_iter > 0)
    HSG_TRY(attn_prep_bwd_ex(sw.H, sw.d, sw.in_dim, sw.feat_dim, L.ldz_sw, sw.W, sw.Wf, sw.bf, sw.a, a->T,
                             sc + L.dWaug_sw, sc + L.dq_sw, b->s2w.dW, b->s2w.dWf, b->s2w.dbf, b->s2w.da, b->dT, acc, 1,
                             s));
  return HSG_OK;
}

}  // extern "C"
