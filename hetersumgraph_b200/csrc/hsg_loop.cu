// Whole update loop of HSumGraph.forward / HSumDocGraph.forward (HiGraph.py:98-106, 205-214) from ONE C call
// each way.  Host-side sequencing only; the kernels are those of hsg_prep.cu / hsg_gemm*.cu / hsg_edge.cu /
// hsg_ffn.cu / hsg_dropout.cu.  What it buys over per-application calls (hsg_wswgat_fwd/bwd):
//   * ~45 launches are enqueued back to back from C (2-3 us each) instead of through Python/ctypes/autograd;
//   * gradients of the shared weights are accumulated by the last stage of the producing kernels
//     (gemm_tn reduce, LayerNorm reduce, dq reduce, attention-prep backward) - no add launches;
//   * the two gradient paths that meet at a node state (GAT.py:47-57: `origin` of the later application,
//     `neighbor` of the earlier one) are summed by the projection-backward epilogue (HSG_EPI_ADD).
//
// A chain of n_apps applications whose kinds alternate from start_kind: kind 0 = W2S (destinations = supernodes,
// neighbor = current word state, origin = current supernode state -> new supernode state), kind 1 = S2W
// (destinations = words, neighbor = supernode state, origin = word state -> new word state).
// HSG / HDSG: start_kind 0, n_apps = 1 + 2 n_iter.  (1, kind) is one stand-alone WSWGAT.forward.
#include <atomic>
#include <cstdlib>
#include <mutex>

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

// ---- side stream for the weight-gradient products of the backward pass -------------------------------------------
// In every application's backward the three weight-gradient products (dW2 = dr^T hdn, dW1 = dhp^T x,
// dW_aug = dzp^T neighbor) feed nothing downstream except the final attention-prep backward, while the chain
// dhp -> dx -> edge backward -> d_neighbor is strictly serial and made of kernels that leave most SMs idle at
// batch 32 (64-127 CTAs).  They are forked onto a second stream (own reduction workspace) and joined by events:
// same kernels, same per-buffer order, bitwise identical results.
struct SideRes {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;                   // finish_kind of the layer whose applications are done first
  cudaEvent_t fork = nullptr, dzp = nullptr, done[2] = {nullptr, nullptr}, fin = nullptr;
  bool ok = false;
};
static SideRes g_side[32];
static std::mutex g_side_mu;
static std::atomic<int> g_overlap{-1};
// SMs the side-stream weight-gradient products may occupy (see tn_plan_tc in hsg_gemm.cu)
static std::atomic<int> g_side_ctas{-1};

int side_ctas() {
  int v = g_side_ctas.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HSG_SIDE_CTAS");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
    g_side_ctas.store(v);
  }
  return v;
}

// HSG_TAIL_ON_MAIN=0: round-2 session-3 placement (last dzp product and the first finish_kind on the one side stream)
bool tail_on_main() {
  static const int v = [] {
    const char* e = getenv("HSG_TAIL_ON_MAIN");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v != 0;
}

// A weight-gradient product whose launch is postponed to the tail of the step (see hsg_update_loop_bwd)
struct DeferredTN {
  bool set = false;
  int M = 0, N1 = 0, N2 = 0, lda = 0, ldb = 0, ldc = 0, acc = 0;
  const float *A = nullptr, *B = nullptr;
  float *C = nullptr, *colsum = nullptr;
};
bool tail_defer_dw1() {
  static const int v = [] {
    const char* e = getenv("HSG_TAIL_DEFER_DW1");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v != 0;
}

// tuning knobs of the tail (measured defaults): SMs the last projection product may occupy on the caller's
// stream (0 = all), and whether dW1 of the last application follows it there
int tail_tn_ctas() {
  static const int v = [] {
    const char* e = getenv("HSG_TAIL_TN_CTAS");
    const int x = e ? atoi(e) : 0;
    return x < 0 ? 0 : x;
  }();
  return v;
}
bool tail_dw1_main() {
  static const int v = [] {
    const char* e = getenv("HSG_TAIL_DW1_MAIN");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v != 0;
}

bool overlap_enabled() {
  int v = g_overlap.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HSG_BWD_OVERLAP");
    v = (e && e[0] == '0') ? 0 : 1;
    g_overlap.store(v);
  }
  return v != 0;
}

SideRes* side_res() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  std::lock_guard<std::mutex> lk(g_side_mu);
  SideRes& r = g_side[dev];
  if (!r.ok) {
    // lowest priority: when SMs free up, the critical-path kernels of the caller's stream (dx, edge backward,
    // d_neighbor) are scheduled before the weight-gradient products
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (cudaStreamCreateWithPriority(&r.stream, cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&r.stream2, cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
    bool good = cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.fin, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.dzp, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.done[0], cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.done[1], cudaEventDisableTiming) == cudaSuccess;
    if (!good) return nullptr;
    r.ok = true;
  }
  return &r;
}

inline size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }
inline size_t mx(size_t a, size_t b) { return a > b ? a : b; }

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

struct KindBuf {            // backward scratch of one kind (dst = its destination node type)
  size_t dr, dhp, g, dx, dzp, gstate, drm, dA, dWblk;
  size_t ln_part, dq_part;  // per-block partials of (dgamma, dbeta) / dq, one slot per application of the kind
  size_t ln_stride, dq_stride;
};

struct Layout {
  int n_apps, start;
  bool has[2];              // kind used at all
  bool drop_attn, drop_ffn;
  int ldz[2], fp[2], n_src[2], n_dst[2];
  size_t waug[2], q[2], wblk[2], aexp, apps_base;
  AppOff app[2];
  size_t state_total;
  size_t dWaug[2], dq[2], zero_out;
  KindBuf kb[2];
  size_t aexp_b;            // backward: regenerated A'
  size_t scratch_total, ws_bytes, ws_half;

  int kind(int i) const { return (i + start) & 1; }
  size_t app_start(int i) const {
    const int n_other = i / 2;                          // applications of the non-starting kind among indices < i
    const int n_start = i - n_other;
    const int n0 = start == 0 ? n_start : n_other, n1 = start == 0 ? n_other : n_start;
    return apps_base + (size_t)n0 * app[0].total + (size_t)n1 * app[1].total;
  }
  int last_of_kind(int k) const {                      // index of the last application of kind k, -1 if none
    for (int i = n_apps - 1; i >= 0; --i)
      if (kind(i) == k) return i;
    return -1;
  }
};

const hsg_layer_params& layer(const hsg_loop_args* a, int k) { return k == 0 ? a->w2s : a->s2w; }

int make_layout(const hsg_loop_args* a, Layout* L) {
  if (!a || a->n_apps < 1 || a->n_word < 0 || a->n_super < 0 || (a->start_kind != 0 && a->start_kind != 1))
    return HSG_ERR_ARG;
  if (!(a->attn_p >= 0.f) || !(a->attn_p < 1.f) || !(a->ffn_p >= 0.f) || !(a->ffn_p < 1.f)) return HSG_ERR_ARG;
  L->n_apps = a->n_apps;
  L->start = a->start_kind;
  L->has[a->start_kind] = true;
  L->has[a->start_kind ^ 1] = a->n_apps > 1;
  L->drop_attn = a->attn_p > 0.f;
  L->drop_ffn = a->ffn_p > 0.f;
  L->n_src[0] = a->n_word;  L->n_dst[0] = a->n_super;
  L->n_src[1] = a->n_super; L->n_dst[1] = a->n_word;
  for (int k = 0; k < 2; ++k) {
    L->ldz[k] = L->fp[k] = 0;
    if (!L->has[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    HSG_TRY(hsg_edge_layout(P.H, P.d, &L->fp[k], &L->ldz[k]));
    if (P.in_dim <= 0 || P.feat_dim <= 0 || P.d_hid <= 0) return HSG_ERR_ARG;
  }
  if (L->has[0] && L->has[1]) {
    const hsg_layer_params &ws = a->w2s, &sw = a->s2w;
    if (sw.in_dim != ws.H * ws.d || ws.in_dim != sw.H * sw.d || sw.feat_dim != ws.feat_dim)
      return HSG_ERR_SHAPE;   // the two layers map word dim <-> hidden dim and share the TF-IDF table
  }
  size_t off = 0;
  size_t aexp = 0;
  for (int k = 0; k < 2; ++k) {
    L->waug[k] = L->q[k] = L->wblk[k] = off;
    if (!L->has[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    L->waug[k] = off; off += r4((size_t)L->ldz[k] * P.in_dim);
    L->q[k] = off;    off += r4((size_t)HSG_N_BINS * P.H);
    if (L->drop_attn) {
      L->wblk[k] = off; off += r4((size_t)L->ldz[k] * P.H * P.in_dim);
      aexp = mx(aexp, r4((size_t)L->n_src[k] * P.H * P.in_dim));
    }
  }
  L->aexp = off; off += aexp;
  L->apps_base = off;
  for (int k = 0; k < 2; ++k)
    L->app[k] = L->has[k] ? app_layout(layer(a, k), L->ldz[k], L->n_src[k], L->n_dst[k])
                          : AppOff{0, 0, 0, 0, 0, 0, 0, 0, 0};
  L->state_total = mx(L->app_start(L->n_apps), 4);

  // backward scratch
  off = 0;
  size_t zero_out = 0;
  for (int k = 0; k < 2; ++k) {
    KindBuf& b = L->kb[k];
    L->dWaug[k] = L->dq[k] = off;
    b = KindBuf{off, off, off, off, off, off, off, off, off, off, off, 0, 0};
    if (!L->has[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    const size_t F = (size_t)P.H * P.d, nd = (size_t)L->n_dst[k], ns = (size_t)L->n_src[k];
    L->dWaug[k] = off; off += r4((size_t)L->ldz[k] * P.in_dim);
    L->dq[k] = off;    off += r4((size_t)HSG_N_BINS * P.H);
    b.dr = off;     off += r4(nd * F);
    b.dhp = off;    off += r4(nd * P.d_hid);
    b.g = off;      off += r4(nd * L->fp[k]);
    b.dx = off;     off += r4(nd * F);
    b.dzp = off;    off += r4(ns * L->ldz[k]);
    b.gstate = off; off += r4(ns * P.in_dim);
    if (L->drop_ffn) { b.drm = off; off += r4(nd * F); }
    if (L->drop_attn) {
      b.dA = off;    off += r4(ns * P.H * P.in_dim);
      b.dWblk = off; off += r4((size_t)L->ldz[k] * P.H * P.in_dim);
    }
    // deferred reduces: every application of the kind leaves its partial rows in its own slot
    const int n_k = (L->n_apps + (L->start == k ? 1 : 0)) / 2;           // applications of this kind
    b.ln_stride = r4(hsg_layernorm_bwd_workspace_bytes((int)nd, (int)F) / sizeof(float) + 1);
    b.dq_stride = r4(hsg_edge_bwd_workspace_bytes(P.H) / sizeof(float) + 1);
    b.ln_part = off; off += (size_t)n_k * b.ln_stride;
    b.dq_part = off; off += (size_t)n_k * b.dq_stride;
    zero_out = mx(zero_out, r4(nd * F));
  }
  L->zero_out = off; off += zero_out;
  L->aexp_b = off;   off += aexp;
  L->scratch_total = mx(off, 4);

  size_t w = 16;
  for (int k = 0; k < 2; ++k) {
    if (!L->has[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    w = mx(w, hsg_wswgat_bwd_workspace_bytes(P.H, P.d, P.in_dim, P.d_hid, L->n_src[k], L->n_dst[k]));
    if (L->drop_attn) w = mx(w, hsg_gemm_tn_workspace_bytes(L->n_src[k], L->ldz[k], P.H * P.in_dim));
  }
  L->ws_half = (w + 255) & ~(size_t)255;
  L->ws_bytes = 2 * L->ws_half;              // second half: reduction workspace of the side stream
  return HSG_OK;
}

bool params_ok(const hsg_layer_params& P) {
  return P.W && P.Wf && P.a && P.w1 && P.b1 && P.w2 && P.b2 && P.gamma && P.beta;
}

bool grads_ok(const hsg_layer_grads& G, bool need_bf) {
  return G.dW && G.dWf && G.da && G.dw1 && G.db1 && G.dw2 && G.db2 && G.dgamma && G.dbeta && (!need_bf || G.dbf);
}

int common_checks(const hsg_loop_args* a, const Layout& L) {
  if (!a->csc_super || !a->csc_word || !a->T || !a->state) return HSG_ERR_ARG;
  if ((a->n_word > 0 && !a->word_feature) || (a->n_super > 0 && !a->super_feature)) return HSG_ERR_ARG;
  for (int k = 0; k < 2; ++k)
    if (L.has[k] && !params_ok(layer(a, k))) return HSG_ERR_ARG;
  if (a->state_floats < L.state_total) return HSG_ERR_WORKSPACE;
  if (!aligned16(a->state) || !aligned16(a->word_feature) || !aligned16(a->super_feature)) return HSG_ERR_ALIGN;
  if (a->csc_super->n_dst != a->n_super || a->csc_super->n_src != a->n_word || a->csc_word->n_dst != a->n_word ||
      a->csc_word->n_src != a->n_super)
    return HSG_ERR_SHAPE;
  return HSG_OK;
}

// dropout stream ids of application i
inline unsigned int stream_attn(int i) { return 2u * (unsigned int)i; }
inline unsigned int stream_ffn(int i) { return 2u * (unsigned int)i + 1u; }

// forward of application i
int app_fwd(const hsg_loop_args* a, const Layout& L, int i, const float* neighbor, const float* origin,
            cudaStream_t s) {
  const int k = L.kind(i);
  const hsg_layer_params& P = layer(a, k);
  const AppOff& o = L.app[k];
  float* st = a->state;
  float* blk = st + L.app_start(i);
  const int F = P.H * P.d, n_src = L.n_src[k], n_dst = L.n_dst[k], ldz = L.ldz[k];
  const hsg_csc* csc = k == 0 ? a->csc_super : a->csc_word;
  float *zp = blk + o.zp, *sh = blk + o.sh, *x = blk + o.x, *stat = blk + o.stat, *hdn = blk + o.hdn,
        *r = blk + o.r, *ln = blk + o.ln, *out = blk + o.out;
  // zp = neighbor . W_aug^T   (fc of all heads + p = a_src . z, GATLayer.py:110,146); with input dropout the
  // per-head masked copies of the input multiply the head-blocked weight (hsg_dropout.cu)
  if (L.drop_attn) {
    float* aexp = st + L.aexp;
    HSG_TRY(dropout_expand(n_src, P.in_dim, P.H, neighbor, aexp, make_drop(a->attn_p, a->seed, stream_attn(i), a->seed_dev), s));
    HSG_TRY(hsg_gemm_nt(n_src, ldz, P.H * P.in_dim, aexp, P.H * P.in_dim, st + L.wblk[k], P.H * P.in_dim, zp, ldz,
                        nullptr, nullptr, 0, 0, s));
  } else {
    HSG_TRY(hsg_gemm_nt(n_src, ldz, P.in_dim, neighbor, P.in_dim, st + L.waug[k], P.in_dim, zp, ldz, nullptr, nullptr,
                        0, 0, s));
  }
  // sh, x = elu(sh) + origin, stat   (GATLayer.py:88-102,112-113; GAT.py:56-57); when the backward prep recomputes sh
  // (hsg_edge_rc.cu) the forward does not store it
  if (edge_recompute_use(csc, P.H, P.d, ldz)) sh = nullptr;
  HSG_TRY(hsg_edge_fwd(csc, P.H, P.d, zp, ldz, st + L.q[k], origin, sh, x, stat, s));
  // FFN (GATLayer.py:35-44); a small destination set (the sentence side) takes the one-launch row kernel
  if (!L.drop_ffn && ffn_rows_ok(n_dst, F, P.d_hid))
    return ffn_rows_fwd(n_dst, F, P.d_hid, x, P.w1, P.b1, P.w2, P.b2, P.gamma, P.beta, hdn, r, out, ln, s);
  HSG_TRY(hsg_gemm_nt(n_dst, P.d_hid, F, x, F, P.w1, F, hdn, P.d_hid, P.b1, nullptr, 0, HSG_EPI_BIAS | HSG_EPI_RELU, s));
  if (L.drop_ffn) {
    HSG_TRY(hsg_gemm_nt(n_dst, F, P.d_hid, hdn, P.d_hid, P.w2, P.d_hid, r, F, P.b2, nullptr, 0, HSG_EPI_BIAS, s));
    return layernorm_fwd_dropres(n_dst, F, r, x, make_drop(a->ffn_p, a->seed, stream_ffn(i), a->seed_dev), P.gamma, P.beta, out, ln,
                                 s);
  }
  HSG_TRY(hsg_gemm_nt(n_dst, F, P.d_hid, hdn, P.d_hid, P.w2, P.d_hid, r, F, P.b2, x, F, HSG_EPI_BIAS | HSG_EPI_ADD, s));
  return hsg_layernorm_fwd(n_dst, F, r, P.gamma, P.beta, out, ln, s);
}

// backward of application i.  dnb == nullptr skips the d_neighbor product; dnb_add (may be nullptr) is summed into it.
// sd != nullptr: weight-gradient products go to the side stream (workspace ws2); side_pending[k] tells whether an
// earlier application of the same kind still has side work in flight on this kind's scratch buffers.
int app_bwd(const hsg_loop_args* a, const Layout& L, int i, const float* neighbor, const float* dout, float* dx,
            float* dnb, const float* dnb_add, float* sc, int acc_aug, const hsg_layer_grads& G, int acc_ffn, void* ws,
            size_t ws_bytes, void* ws2, SideRes* sd, bool* side_pending, int slot, int* ln_blocks, int* dq_blocks,
            cudaStream_t s, DeferredTN* defer = nullptr) {
  const int k = L.kind(i);
  const hsg_layer_params& P = layer(a, k);
  const AppOff& o = L.app[k];
  const KindBuf& b = L.kb[k];
  float* st = a->state;
  float* blk = st + L.app_start(i);
  const int F = P.H * P.d, n_src = L.n_src[k], n_dst = L.n_dst[k], ldz = L.ldz[k];
  const hsg_csc* csc_t = k == 0 ? a->csc_word : a->csc_super;     // transposed structure = the other direction's CSC
  const float *zp = blk + o.zp, *sh = blk + o.sh, *x = blk + o.x, *hdn = blk + o.hdn, *r = blk + o.r, *ln = blk + o.ln;
  float* stat = blk + o.stat;
  float *dr = sc + b.dr, *dhp = sc + b.dhp, *g = sc + b.g, *dzp = sc + b.dzp;
  float *dW_aug = sc + L.dWaug[k], *dq = sc + L.dq[k];
  cudaStream_t s2 = sd ? sd->stream : s;            // stream of the weight-gradient products
  void* wsw = sd ? ws2 : ws;
  if (sd && side_pending[k]) {                      // this kind's dr / dhp / dzp are about to be overwritten
    if (cudaStreamWaitEvent(s, sd->done[k], 0) != cudaSuccess) return HSG_ERR_CUDA;
    side_pending[k] = false;
  }
  const bool rows_kernel = !L.drop_ffn && ffn_rows_ok(n_dst, F, P.d_hid);
  const float* drm = dr;
  // (dgamma, dbeta) and dq leave the kernels as per-block partials in this application's slot; the fixed-order reduces
  // run once per layer after its last application (hsg_update_loop_bwd), not here on the critical chain
  void* ln_ws = sc + b.ln_part + (size_t)slot * b.ln_stride;
  const size_t ln_ws_bytes = b.ln_stride * sizeof(float);
  void* dq_ws = sc + b.dq_part + (size_t)slot * b.dq_stride;
  const size_t dq_ws_bytes = b.dq_stride * sizeof(float);
  if (rows_kernel) {
    // small destination set: LayerNorm backward, dhp and dx in ONE launch (+ the dgamma / dbeta reduce)
    HSG_TRY(ffn_rows_bwd(n_dst, F, P.d_hid, dout, r, ln, P.gamma, hdn, P.w1, P.w2, dr, dhp, dx, G.dgamma, G.dbeta, ln_ws,
                         ln_ws_bytes, acc_ffn, s, ln_blocks));
  } else {
    // LayerNorm
    HSG_TRY(layernorm_bwd_ex(n_dst, F, dout, r, ln, P.gamma, dr, G.dgamma, G.dbeta, ln_ws, ln_ws_bytes, acc_ffn, s,
                             ln_blocks));
    // FFN dropout: the residual path keeps dr, the W2 path sees dr * mask / (1-p)
    if (L.drop_ffn) {
      HSG_TRY(dropout_mul((size_t)n_dst * F, dr, sc + b.drm, make_drop(a->ffn_p, a->seed, stream_ffn(i), a->seed_dev), s));
      drm = sc + b.drm;
    }
    // FFN: dhp = (drm . W2) * relu', dW2 = drm^T hdn, dW1 = dhp^T x, dx = dhp . W1 + dr
    HSG_TRY(hsg_gemm_nn(n_dst, P.d_hid, F, drm, F, P.w2, P.d_hid, dhp, P.d_hid, hdn, P.d_hid, HSG_EPI_RELU_MASK, s));
  }
  if (sd) {
    if (cudaEventRecord(sd->fork, s) != cudaSuccess || cudaStreamWaitEvent(s2, sd->fork, 0) != cudaSuccess)
      return HSG_ERR_CUDA;
  }
  const int budget = sd ? side_ctas() : 0;
  HSG_TRY(gemm_tn_ex(n_dst, F, P.d_hid, drm, F, hdn, P.d_hid, G.dw2, P.d_hid, G.db2, wsw, ws_bytes, acc_ffn, s2, budget));
  // end of the chain (no d_neighbor wanted) over a small destination set: the side stream would run the two FFN weight
  // gradients one after the other while the caller's stream idles behind its last projection product; dW1 follows
  // that product on the caller's stream instead (same kernel, same operands: the placement changes no bit)
  const bool dw1_on_main = sd && !dnb && rows_kernel && !L.drop_attn && tail_on_main() && tail_dw1_main();
  if (defer && sd && !rows_kernel) {                // launched by the caller at the tail of the step (side stream)
    defer->set = true;
    defer->M = n_dst; defer->N1 = P.d_hid; defer->N2 = F;
    defer->A = dhp; defer->lda = P.d_hid; defer->B = x; defer->ldb = F;
    defer->C = G.dw1; defer->ldc = F; defer->colsum = G.db1; defer->acc = acc_ffn;
  } else if (!dw1_on_main) {
    HSG_TRY(gemm_tn_ex(n_dst, P.d_hid, F, dhp, P.d_hid, x, F, G.dw1, F, G.db1, wsw, ws_bytes, acc_ffn, s2, budget));
  }
  if (!rows_kernel)
    HSG_TRY(hsg_gemm_nn(n_dst, F, P.d_hid, dhp, P.d_hid, P.w1, F, dx, F, dr, F, HSG_EPI_ADD, s));
  // edge backward (d origin = dx, GAT.py:57)
  const hsg_csc* csc_f = k == 0 ? a->csc_super : a->csc_word;
  if (edge_recompute_use(csc_f, P.H, P.d, ldz))
    HSG_TRY(hsg_edge_bwd_prep_rc(csc_f, P.H, P.d, zp, ldz, st + L.q[k], dx, g, stat, s));
  else
    HSG_TRY(hsg_edge_bwd_prep(n_dst, P.H, P.d, dx, nullptr, sh, g, stat, s));
  HSG_TRY(edge_bwd_ex(csc_t, P.H, P.d, zp, ldz, st + L.q[k], g, stat, dzp, dq, dq_ws, dq_ws_bytes, acc_aug, s,
                      dq_blocks));
  // projection backward
  int rc;
  if (L.drop_attn) {
    const int KW = P.H * P.in_dim;
    const DropCfg dc = make_drop(a->attn_p, a->seed, stream_attn(i), a->seed_dev);
    if (dnb) {
      HSG_TRY(hsg_gemm_nn(n_src, KW, ldz, dzp, ldz, st + L.wblk[k], KW, sc + b.dA, KW, nullptr, 0, 0, s));
      HSG_TRY(dropout_reduce(n_src, P.in_dim, P.H, sc + b.dA, dnb_add, dnb, dc, s));
    }
    float* aexp = sc + L.aexp_b;                    // shared by both kinds: stays on the main stream
    HSG_TRY(dropout_expand(n_src, P.in_dim, P.H, neighbor, aexp, dc, s));
    HSG_TRY(gemm_tn_ex(n_src, ldz, KW, dzp, ldz, aexp, KW, sc + b.dWblk, KW, nullptr, ws, ws_bytes, 0, s));
    rc = wblk_gather(P.H, P.d, P.in_dim, ldz, sc + b.dWblk, dW_aug, acc_aug, s);
  } else {
    // the chain ends here when no d_neighbor is wanted (application 0 over a frozen embedding): the caller's stream
    // has nothing left to run, so this product stays on it (own workspace half) instead of queueing behind the FFN
    // weight gradients of the side stream
    const bool tn_on_main = sd && !dnb && tail_on_main();
    if (sd && !tn_on_main) {
      if (cudaEventRecord(sd->dzp, s) != cudaSuccess || cudaStreamWaitEvent(s2, sd->dzp, 0) != cudaSuccess)
        return HSG_ERR_CUDA;
    }
    if (tn_on_main && cudaEventRecord(sd->dzp, s) != cudaSuccess) return HSG_ERR_CUDA;   // edge backward done (dq partials)
    HSG_TRY(gemm_tn_ex(n_src, ldz, P.in_dim, dzp, ldz, neighbor, P.in_dim, dW_aug, P.in_dim, nullptr,
                       tn_on_main ? ws : wsw, ws_bytes, acc_aug, tn_on_main ? s : s2, tn_on_main ? tail_tn_ctas() : budget));
    if (dw1_on_main)
      HSG_TRY(gemm_tn_ex(n_dst, P.d_hid, F, dhp, P.d_hid, x, F, G.dw1, F, G.db1, ws, ws_bytes, acc_ffn, s, 0));
    rc = HSG_OK;
    if (dnb)
      rc = hsg_gemm_nn(n_src, P.in_dim, ldz, dzp, ldz, st + L.waug[k], P.in_dim, dnb, P.in_dim, dnb_add, P.in_dim,
                       dnb_add ? HSG_EPI_ADD : 0, s);
  }
  if (rc) return rc;
  if (sd) {
    if (cudaEventRecord(sd->done[k], s2) != cudaSuccess) return HSG_ERR_CUDA;
    side_pending[k] = true;
  }
  return HSG_OK;
}

// Everything of layer `k` that waits for its LAST application in backward order: the fixed-order reduces of the
// (dgamma, dbeta) and dq partials of its n_k applications (one launch each), then the attention-prep backward
// (dW_aug, dq) -> fc / feat_fc / attn_fc / TF-IDF table.
// parts: 1 = the two reduces (need the layer's edge / LayerNorm backward kernels only), 2 = attention-prep backward
// (needs dq and the layer's complete dW_aug)
int finish_kind(const hsg_loop_args* a, const Layout& L, int k, int n_k, int ln_blocks, int dq_blocks, float* sc,
                const hsg_layer_grads& G, float* dT, int acc_params, int acc_T, cudaStream_t s, int parts = 3) {
  const hsg_layer_params& P = layer(a, k);
  const KindBuf& b = L.kb[k];
  const int F = P.H * P.d;
  if (parts & 1) {
    HSG_TRY(ln_partials_reduce(n_k, ln_blocks, b.ln_stride, F, sc + b.ln_part, G.dgamma, G.dbeta, acc_params, s));
    HSG_TRY(edge_dq_reduce(n_k, dq_blocks, b.dq_stride, HSG_N_BINS * P.H, sc + b.dq_part, sc + L.dq[k], 0, s));
  }
  if (!(parts & 2)) return HSG_OK;
  return attn_prep_bwd_ex(P.H, P.d, P.in_dim, P.feat_dim, L.ldz[k], P.W, P.Wf, P.bf, P.a, a->T, sc + L.dWaug[k],
                          sc + L.dq[k], G.dW, G.dWf, G.dbf, G.da, dT, acc_params, acc_T, s);
}

}  // namespace

extern "C" {

int hsg_set_bwd_overlap(int on) {
  g_overlap.store(on ? 1 : 0);
  return HSG_OK;
}

int hsg_set_side_ctas(int n) {
  if (n < 0) return HSG_ERR_ARG;
  g_side_ctas.store(n);
  return HSG_OK;
}

int hsg_update_loop_plan(const hsg_loop_args* a, hsg_loop_plan* plan) {
  if (!plan) return HSG_ERR_ARG;
  Layout L;
  HSG_TRY(make_layout(a, &L));
  plan->state_floats = L.state_total;
  plan->scratch_floats = L.scratch_total;
  plan->ws_bytes = L.ws_bytes;
  const int ls = L.last_of_kind(0), lw = L.last_of_kind(1);
  plan->super_state_off = ls >= 0 ? L.app_start(ls) + L.app[0].out : (size_t)-1;
  plan->word_state_off = lw >= 0 ? L.app_start(lw) + L.app[1].out : (size_t)-1;
  plan->hdn_off[0] = L.app_start(0) + L.app[L.kind(0)].hdn;
  plan->hdn_off[1] = L.n_apps > 1 ? L.app_start(1) + L.app[L.kind(1)].hdn : (size_t)-1;
  plan->pair_stride = L.app[0].total + L.app[1].total;
  return HSG_OK;
}

int hsg_update_loop_fwd(const hsg_loop_args* a, void* stream) {
  Layout L;
  HSG_TRY(make_layout(a, &L));
  HSG_TRY(common_checks(a, L));
  cudaStream_t s = (cudaStream_t)stream;
  float* st = a->state;
  // attention prep: once per layer (parameters only)
  // The prep of the kind application 0 needs runs on the caller's stream; the other kind's (first needed by
  // application 1) runs on the side stream next to application 0 and is joined before application 1.
  SideRes* sd = (overlap_enabled() && L.n_apps > 1) ? side_res() : nullptr;
  const int k_first = L.kind(0);
  for (int kk = 0; kk < 2; ++kk) {
    const int k = kk == 0 ? k_first : (k_first ^ 1);
    if (!L.has[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    cudaStream_t sp = s;
    if (kk == 1 && sd) {
      if (cudaEventRecord(sd->fork, s) != cudaSuccess || cudaStreamWaitEvent(sd->stream, sd->fork, 0) != cudaSuccess)
        return HSG_ERR_CUDA;
      sp = sd->stream;
    }
    HSG_TRY(hsg_attn_prep_fwd(P.H, P.d, P.in_dim, P.feat_dim, L.ldz[k], P.W, P.Wf, P.bf, P.a, a->T, st + L.waug[k],
                              st + L.q[k], (void*)sp));
    if (L.drop_attn) HSG_TRY(wblk_build(P.H, P.d, P.in_dim, L.ldz[k], st + L.waug[k], st + L.wblk[k], sp));
    if (kk == 1 && sd && cudaEventRecord(sd->done[k], sp) != cudaSuccess) return HSG_ERR_CUDA;
  }
  const float* word = a->word_feature;
  const float* sup = a->super_feature;
  if (a->input_ready && cudaStreamWaitEvent(s, (cudaEvent_t)a->input_ready, 0) != cudaSuccess) return HSG_ERR_CUDA;
  for (int i = 0; i < L.n_apps; ++i) {
    const int k = L.kind(i);
    if (i == 1 && sd && cudaStreamWaitEvent(s, sd->done[k], 0) != cudaSuccess) return HSG_ERR_CUDA;
    HSG_TRY(app_fwd(a, L, i, k == 0 ? word : sup, k == 0 ? sup : word, s));
    float* out = st + L.app_start(i) + L.app[k].out;
    if (k == 0) sup = out; else word = out;
  }
  return HSG_OK;
}

int hsg_update_loop_bwd(const hsg_loop_args* a, const hsg_loop_bwd_args* b, void* stream) {
  Layout L;
  HSG_TRY(make_layout(a, &L));
  HSG_TRY(common_checks(a, L));
  if (!b || !b->dT || !b->scratch || !b->ws) return HSG_ERR_ARG;
  for (int k = 0; k < 2; ++k)
    if (L.has[k] && !grads_ok(k == 0 ? b->w2s : b->s2w, layer(a, k).bf != nullptr)) return HSG_ERR_ARG;
  if (!b->d_super_state && !b->d_word_state) return HSG_ERR_ARG;
  if (b->scratch_floats < L.scratch_total || b->ws_bytes < L.ws_bytes) return HSG_ERR_WORKSPACE;
  if (!aligned16(b->scratch) || (b->d_super_feature && !aligned16(b->d_super_feature)) ||
      (b->d_word_feature && !aligned16(b->d_word_feature)) || (b->d_super_state && !aligned16(b->d_super_state)) ||
      (b->d_word_state && !aligned16(b->d_word_state)))
    return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  float* st = a->state;
  float* sc = b->scratch;
  const int acc = b->accumulate ? 1 : 0;

  // pending gradient w.r.t. the current state of each node type, indexed by the kind that PRODUCES that state
  // (0: supernode state, 1: word state); NULL = zero
  const float* gst[2] = {b->d_super_state, b->d_word_state};
  int done[2] = {0, 0};
  SideRes* sd = overlap_enabled() ? side_res() : nullptr;
  bool side_pending[2] = {false, false};
  bool prep_done[2] = {false, false};
  bool fin_pending = false;
  DeferredTN deferred;
  int ln_blocks[2] = {0, 0}, dq_blocks[2] = {0, 0};
  void* ws2 = reinterpret_cast<char*>(b->ws) + L.ws_half;
  for (int i = L.n_apps - 1; i >= 0; --i) {
    const int k = L.kind(i);
    const hsg_layer_params& P = layer(a, k);
    // neighbor = the state the previous application (kind k^1: kinds alternate) produced, or the chain input
    const float* neighbor = i > 0 ? st + L.app_start(i - 1) + L.app[k ^ 1].out
                                  : (k == 0 ? a->word_feature : a->super_feature);
    const float* dout = gst[k];
    if (!dout) {                                       // this application's result has no consumer: zero cotangent
      const size_t n = (size_t)L.n_dst[k] * P.H * P.d;
      if (cudaMemsetAsync(sc + L.zero_out, 0, n * sizeof(float), s) != cudaSuccess) return HSG_ERR_CUDA;
      dout = sc + L.zero_out;
    }
    // dx  = gradient of the origin (the kind-k state BEFORE this application)       -> replaces gst[k]
    // dnb = gradient of the neighbor through this application + its pending gradient -> replaces gst[k^1]
    float* dx = sc + L.kb[k].dx;
    float* dnb = sc + L.kb[k].gstate;
    if (i == 0) {                                      // chain inputs: deliver into the caller's buffers
      float* ext_origin = k == 0 ? b->d_super_feature : b->d_word_feature;
      float* ext_neighbor = k == 0 ? b->d_word_feature : b->d_super_feature;
      if (ext_origin) dx = ext_origin;
      dnb = ext_neighbor;                              // NULL: not wanted (e.g. frozen embedding) - product skipped
    }
    const hsg_layer_grads& G = k == 0 ? b->w2s : b->s2w;
    int lnb = 0, dqb = 0;
    // HSG_TAIL_DEFER_DW1=1: the big dW1 product of application 1 (the last of its kind) waits for application 0's edge
    // backward and runs next to the tail's small kernels instead of next to application 0's chain
    const bool defer_here = i == 1 && sd && tail_on_main() && tail_defer_dw1() && !L.drop_attn && !L.drop_ffn &&
                            (L.kind(0) == 0 ? b->d_word_feature : b->d_super_feature) == nullptr;
    HSG_TRY(app_bwd(a, L, i, neighbor, dout, dx, dnb, gst[k ^ 1], sc, done[k] > 0, G, acc || done[k] > 0, b->ws,
                    L.ws_half, ws2, sd, side_pending, done[k], &lnb, &dqb, s, defer_here ? &deferred : nullptr));
    if (i == 0 && deferred.set) {
      if (cudaStreamWaitEvent(sd->stream, sd->dzp, 0) != cudaSuccess) return HSG_ERR_CUDA;
      HSG_TRY(gemm_tn_ex(deferred.M, deferred.N1, deferred.N2, deferred.A, deferred.lda, deferred.B, deferred.ldb,
                         deferred.C, deferred.ldc, deferred.colsum, ws2, L.ws_half, deferred.acc, sd->stream, 0));
      if (cudaEventRecord(sd->done[k], sd->stream) != cudaSuccess) return HSG_ERR_CUDA;
      side_pending[k] = true;
    }
    if (done[k] > 0 && (lnb != ln_blocks[k] || dqb != dq_blocks[k])) return HSG_ERR_SHAPE;   // same launch per kind
    ln_blocks[k] = lnb;
    dq_blocks[k] = dqb;
    ++done[k];
    gst[k] = dx;
    gst[k ^ 1] = dnb;
    if (i == 1) {
      // application 1 is the LAST one of its kind in this backward order: its layer's (dW_aug, dq) are complete.  The
      // attention-prep backward of that layer runs now - on the side stream, next to application 0 - instead of at
      // the end of the step; the other layer's follows application 0 and ADDS its share of dT (two terms: the sum
      // does not depend on the order).
      cudaStream_t sp = s;
      const bool own = sd && tail_on_main();         // its own branch: does not delay application 0's weight gradients
      if (sd) {
        sp = own ? sd->stream2 : sd->stream;
        if (cudaEventRecord(sd->fork, s) != cudaSuccess || cudaStreamWaitEvent(sp, sd->fork, 0) != cudaSuccess)
          return HSG_ERR_CUDA;
        if (own && side_pending[k] && cudaStreamWaitEvent(sp, sd->done[k], 0) != cudaSuccess) return HSG_ERR_CUDA;
      }
      HSG_TRY(finish_kind(a, L, k, done[k], ln_blocks[k], dq_blocks[k], sc, G, b->dT, acc, acc, sp));
      prep_done[k] = true;
      if (sd) {
        if (own) {
          if (cudaEventRecord(sd->fin, sp) != cudaSuccess) return HSG_ERR_CUDA;
          fin_pending = true;
        } else {
          if (cudaEventRecord(sd->done[k], sp) != cudaSuccess) return HSG_ERR_CUDA;
          side_pending[k] = true;
        }
      }
    }
  }
  // the reduces of the layer application 0 belongs to wait for its edge backward only (event dzp, recorded when the last
  // projection product stayed on this stream): they run on the second branch next to that product
  int early_reduce = -1;
  if (sd && tail_on_main() && !L.drop_attn && L.n_apps > 0) {
    const int k0 = L.kind(0);
    const bool dnb0 = (k0 == 0 ? b->d_word_feature : b->d_super_feature) != nullptr;
    if (!dnb0 && !prep_done[k0]) {
      const hsg_layer_grads& G = k0 == 0 ? b->w2s : b->s2w;
      if (cudaStreamWaitEvent(sd->stream2, sd->dzp, 0) != cudaSuccess) return HSG_ERR_CUDA;
      HSG_TRY(finish_kind(a, L, k0, done[k0], ln_blocks[k0], dq_blocks[k0], sc, G, b->dT, acc, acc, sd->stream2, 1));
      if (cudaEventRecord(sd->fin, sd->stream2) != cudaSuccess) return HSG_ERR_CUDA;
      fin_pending = true;
      early_reduce = k0;
    }
  }
  if (sd) {                                             // join: every weight gradient is complete from here on
    for (int k = 0; k < 2; ++k)
      if (side_pending[k] && cudaStreamWaitEvent(s, sd->done[k], 0) != cudaSuccess) return HSG_ERR_CUDA;
    if (fin_pending && cudaStreamWaitEvent(s, sd->fin, 0) != cudaSuccess) return HSG_ERR_CUDA;
  }
  // attention-prep backward: (dW_aug, dq) summed over the applications -> fc / feat_fc / attn_fc / TF-IDF table
  int t_written = (prep_done[0] || prep_done[1]) ? 1 : 0;
  for (int k = 0; k < 2; ++k) {
    if (!L.has[k] || prep_done[k]) continue;
    const hsg_layer_params& P = layer(a, k);
    const hsg_layer_grads& G = k == 0 ? b->w2s : b->s2w;
    HSG_TRY(finish_kind(a, L, k, done[k], ln_blocks[k], dq_blocks[k], sc, G, b->dT, acc, acc || t_written, s,
                        k == early_reduce ? 2 : 3));
    t_written = 1;
  }
  return HSG_OK;
}

}  // extern "C"
