/* hsg_b200.h - C ABI of the B200-native WSWGAT message-passing path.
 *
 * Drop-in boundary for the reference's (yellow-binary-tree/HeterSumGraph, pure
 * Python + DGL 0.4) WSWGAT path.  The reference has no FFI of its own; what a
 * replacement has to stand behind is the nn.Module surface
 *     WSWGAT.forward(g, w, s)              module/GAT.py:45-59
 *     MultiHeadLayer.forward(g, h)         module/GATStackLayer.py:55-63
 *     WSGATLayer / SWGATLayer.forward      module/GATLayer.py:104-116, 142-152
 *     PositionwiseFeedForward.forward      module/GATLayer.py:35-44
 *     ExampleSet.CreateGraph / graph_collate_fn (dgl.batch)
 *                                          module/dataloader.py:201-268, 328-406, 472-481
 * and each entry point below names the reference lines it replaces.  The
 * Python host side (hetersumgraph_b200/) binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the caller owns all
 *     memory (incl. workspaces); nothing here allocates or frees.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never synchronise, and are CUDA-graph capturable.
 *   - return 0 on success, a negative hsg_status otherwise; never throws.
 *   - float tensors are fp32 row-major; index tensors are int32; bins are uint8.
 *   - feature rows are 16-byte aligned: leading dimensions are multiples of 4.
 */
#ifndef HSG_B200_H_
#define HSG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSG_ABI_VERSION 1
#define HSG_N_BINS 10          /* TF-IDF boxes, HiGraph.py:52, dataloader.py:253 */
#define HSG_LEAKY_SLOPE 0.01f  /* F.leaky_relu default, GATLayer.py:92,131 */
#define HSG_LN_EPS 1e-5f       /* nn.LayerNorm default, GATLayer.py:32 */

typedef enum {
  HSG_OK = 0,
  HSG_ERR_ARG = -1,        /* null pointer / negative size */
  HSG_ERR_SHAPE = -2,      /* (heads, head_dim) or size combination not supported */
  HSG_ERR_ALIGN = -3,      /* pointer or leading dimension not 16-byte aligned */
  HSG_ERR_WORKSPACE = -4,  /* workspace too small */
  HSG_ERR_CUDA = -5,       /* cudaGetLastError() != cudaSuccess after launch */
  HSG_ERR_ARCH = -6,       /* device is not sm_100 */
  HSG_ERR_CAPACITY = -7    /* a per-graph limit of the builder was exceeded */
} hsg_status;

int hsg_version(void);
/* sizeof of the i-th argument structure below, in declaration order (hsg_token_batch = 0 ... hsg_doc_map = 13), 0 past
 * the end: lets a binding (ctypes / cgo / JNI) check its own struct layouts against the library's before the first call. */
size_t hsg_abi_sizeof(int i);
const char* hsg_strerror(int status);
/* 0 when the current device can run the sm_100a kernels, HSG_ERR_ARCH otherwise. */
int hsg_device_check(void);
int hsg_num_sms(void);
/* Programmatic dependent launch for every kernel of the library (default on; HSG_PDL=0 in the environment or
 * hsg_set_pdl(0) turns it off): each kernel's launch latency and prologue overlap its predecessor's tail. */
int hsg_set_pdl(int on);

/* Per-kernel CUDA-event timing (bench.py's roofline leg).  When enabled every
 * launch is bracketed by events on its stream; hsg_profile_read synchronises
 * and returns, for kernel slot i, the launch count and the summed milliseconds.
 * Not capturable: disable before CUDA-graph capture. */
int hsg_profile_enable(int on);
int hsg_profile_reset(void);
int hsg_profile_num_slots(void);
const char* hsg_profile_slot_name(int slot);
int hsg_profile_read(int slot, int* host_count, float* host_ms);
/* total number of kernel launches issued by this library since load (always counted) */
long long hsg_launch_count(void);
/* cudaMemsetAsync on the given stream (a memset node under stream capture, not a kernel launch) */
int hsg_memset(void* ptr, int value, size_t bytes, void* stream);

/* ------------------------------------------------------------------------
 * K0  device-side graph builder
 *     replaces ExampleSet.AddWordNode / CreateGraph (dataloader.py:201-268),
 *     MultiExampleSet.CreateGraph (:328-406) and dgl.batch (:480).
 * Input: the padded token ids of every sentence of every graph, graphs already
 * in batch order (stable sort by #sentences descending, dataloader.py:479), the
 * TF-IDF box of every token (-1: the word is not a TF-IDF key of that
 * sentence, dataloader.py:251) and the filter-id bitmap (dataloader.py:167-182).
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_graphs;
  int32_t n_sent;          /* total sentences S */
  int32_t sent_len;        /* L (sent_max_len) */
  int32_t hdsg;            /* 0: HSG (sent<->sent extras), 1: HDSG (doc nodes) */
  int32_t vocab_size;
  int32_t n_doc;           /* HDSG: total documents D, else 0 */
  int32_t n_doc_tok;       /* HDSG: total document tokens T, else 0 */
  int32_t max_sent_per_graph; /* max over graphs of the sentence count (sizes the per-CTA tables) */
  const int32_t* tokens;          /* [S, L] */
  const int8_t* sent_bin;         /* [S, L] */
  const int32_t* graph_sent_ptr;  /* [B+1] */
  const uint32_t* filter_bitmap;  /* [ceil(V/32)] bit set = filtered id */
  const int32_t* graph_doc_ptr;   /* [B+1]  HDSG */
  const int32_t* sent_doc;        /* [S]    HDSG: local doc index of each sentence */
  const int32_t* doc_tok_ptr;     /* [D+1]  HDSG */
  const int32_t* doc_tokens;      /* [T]    HDSG: unpadded doc token ids (Example2.enc_doc_input) */
  const int8_t* doc_bin;          /* [T]    HDSG */
} hsg_token_batch;

/* Per-graph counts and offsets, [B+1] exclusive prefix sums (entry B = total). */
typedef struct {
  int32_t* word_ptr;    /* word rows   */
  int32_t* super_ptr;   /* supernode rows (sentences then docs per graph) */
  int32_t* node_ptr;    /* DGL node ids */
  int32_t* edge_ptr;    /* DGL edge ids (all edges incl. sent<->sent / sent->doc) */
  int32_t* pair_ptr;    /* word<->supernode pairs = CSC entries per direction */
} hsg_graph_offsets;

/* CSC of one direction: in-edges of every destination row, ascending DGL edge id. */
typedef struct {
  int32_t n_dst;
  int32_t n_src;
  int32_t n_edges;
  int32_t reserved;
  const int32_t* indptr;  /* [n_dst+1] */
  const int32_t* nbr;     /* [E] source ROW (rank among unit==0 resp. unit==1 nodes) */
  const uint8_t* bin;     /* [E] tffrac */
  const int32_t* extra;   /* [n_dst] in-edges that are not word<->supernode edges (e=0, z_src=0); may be NULL */
} hsg_csc;

typedef struct {
  /* capacities of the output arrays (checked) */
  int32_t cap_word, cap_super, cap_pair, reserved;
  hsg_graph_offsets off;
  int32_t* word_wid;      /* [Nw] ndata["id"] of every word row */
  int32_t* word_nid;      /* [Nw] DGL node id of every word row  (== filter_nodes(unit==0)) */
  int32_t* super_nid;     /* [Ns] DGL node id of every supernode row (== filter_nodes(unit==1)) */
  int8_t* super_type;     /* [Ns] ndata["dtype"]: 1 sentence, 2 document */
  int32_t* super_graph;   /* [Ns] graph index */
  int32_t* super_indptr;  /* [Ns+1]  word->supernode CSC */
  int32_t* super_src;     /* [E] word row */
  uint8_t* super_bin;     /* [E] */
  int32_t* super_eid;     /* [E] DGL edge id of the w->s edge */
  int32_t* super_extra;   /* [Ns] */
  int32_t* word_indptr;   /* [Nw+1]  supernode->word CSC */
  int32_t* word_src;      /* [E] supernode row */
  uint8_t* word_bin;      /* [E] */
  int32_t* word_eid;      /* [E] DGL edge id of the s->w edge */
  int32_t* status;        /* [1] device-side error flag (0 ok, HSG_ERR_CAPACITY ...) */
} hsg_graph_out;

size_t hsg_build_workspace_bytes(const hsg_token_batch* tb);
/* Phase 1: per-graph counts and the [B+1] offsets (off.*[B] = totals).  The host
 * reads the totals (one small D2H) to size the arrays of phase 2. */
int hsg_build_count(const hsg_token_batch* tb, hsg_graph_offsets off, int32_t* status,
                    void* ws, size_t ws_bytes, void* stream);
/* Phase 2: fill node maps and both CSCs. */
int hsg_build_fill(const hsg_token_batch* tb, const hsg_graph_out* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * K2  attention prep: folds attn_fc into the projection and the TF-IDF table.
 *     W_aug[perm(c)] = W[c]                 (fc.weight of all heads, GATLayer.py:84,123; rows lane-interleaved)
 *     W_aug[fp+k]  = sum_j a_k[j] W[k d + j]   so that  p = h W_aug[fp:fp+H]^T = a_src . z
 *     q[b,k]       = a_k[2d:3d] . (Wf_k T[b] + bf_k)      (GATLayer.py:90-92,129-131; HiGraph.py:150-151)
 *     layout holes and rows fp+H .. ld_rows-1 of W_aug are zero (ld_rows = ldz of hsg_edge_layout).
 * ------------------------------------------------------------------------ */
int hsg_attn_prep_fwd(int H, int d, int in_dim, int feat_dim, int ld_rows,
                      const float* W, const float* Wf, const float* bf /* may be NULL */,
                      const float* a /* [H,3d] */, const float* T /* [10,feat] */,
                      float* W_aug /* [ld_rows, in_dim] */, float* q /* [10,H] */, void* stream);
/* Backward of the above: given dW_aug and dq produce dW, dWf, dbf, da, dT (all overwritten). */
int hsg_attn_prep_bwd(int H, int d, int in_dim, int feat_dim, int ld_rows,
                      const float* W, const float* Wf, const float* bf, const float* a, const float* T,
                      const float* dW_aug, const float* dq,
                      float* dW, float* dWf, float* dbf /* may be NULL */, float* da, float* dT, void* stream);

/* ------------------------------------------------------------------------
 * Dense tall-skinny products (K1/K6 projections, K4 FFN pieces).
 *   epilogue flags (combine with |):
 * ------------------------------------------------------------------------ */
#define HSG_EPI_BIAS 1      /* C += bias[n]            */
#define HSG_EPI_RELU 2      /* C = max(C, 0)           */
#define HSG_EPI_ADD 4       /* C += R[m,n]             */
#define HSG_EPI_RELU_MASK 8 /* C = R[m,n] > 0 ? C : 0  */
/* Arithmetic of the three products (process-wide setting):
 *   0  FFMA, exact fp32
 *   1  tcgen05.mma kind::tf32 with a 3-product hi/lo split ("3xTF32", ~2^-21 relative) - default; this is the
 *      fp32-parity mode (BASELINE.json bound 1e-5)
 *   2  tcgen05.mma kind::tf32, single product (bound 2e-2)
 *   3  bf16 projection mode: tcgen05.mma kind::f16 on bf16 operands (the fp32 tiles are rounded to bf16 by the
 *      converter warps on their way to the tensor core), fp32 accumulation in TMEM - the "bf16 projections <= 2e-2"
 *      class of BASELINE.json's north_star for fc (module/GATLayer.py:110,146) and the FFN linears (:38); activations
 *      and weights stay fp32 in HBM, so results of every other kernel are unchanged
 * Shapes the tensor-core path cannot take (K or a leading dimension not a multiple of 4) fall back to mode 0. */
int hsg_set_gemm_mode(int mode);
int hsg_get_gemm_mode(void);
/* Products of fewer than `flops` (2 M N K) run on 64x64 FFMA tiles in every mode: the 128x128 tcgen05 tile pipeline
 * has ~15 us of fill/drain latency and idles most SMs on the sentence-side shapes (M ~ 1 k).  Default 3e8; 0 sends
 * everything the TMA path can take to the tensor cores. */
int hsg_set_gemm_small_flops(double flops);
/* Profiling aid for the tensor-core pipeline: on = 1/0 arms/disarms a trace of CTA 0 (synchronous call); on < 0 reads
 * up to max_events (event id, k-block counter, SM clock) triples into host_out and returns their number. */
int hsg_gemm_trace(int on, unsigned long long* host_out, int max_events);
/* the same for the first cluster of the CTA-pair kernel (events of the peer CTA carry id + 10) */
int hsg_gemm_pair_trace(int on, unsigned long long* host_out, int max_events);
/* C[M,N] = A[M,K] . B[N,K]^T  (nn.Linear / Conv1d(k=1) forward).
 * lda may be SMALLER than K: rows of A then overlap (row m = the K floats starting at A + m*lda) - the sentence
 * encoder's convolution windows (hsg_enc_* below) are read that way without an im2col copy.
 * With HSG_EPI_ADD, R may be the same buffer as C (ldr == ldc): every element is read and then written by one thread
 * (in-place accumulation of K-chunks). */
int hsg_gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const float* bias, const float* R, int ldr, int epi, void* stream);
/* C[M,N] = A[M,K] . B[K,N]    (input gradient) */
int hsg_gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const float* R, int ldr, int epi, void* stream);
/* C[N1,N2] = sum_m A[m,N1] B[m,N2]  (weight gradient), deterministic split over m.
 * colsum (may be NULL): [N1] = sum_m A[m,:] (bias gradient), same pass. */
size_t hsg_gemm_tn_workspace_bytes(int M, int N1, int N2);
int hsg_gemm_tn(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                float* colsum, void* ws, size_t ws_bytes, void* stream);
/* same product; accumulate != 0 ADDS the result to C (and to colsum): a weight gradient accumulated straight into an
 * existing .grad buffer (what autograd's AccumulateGrad would do with one more launch per parameter) */
int hsg_gemm_tn_acc(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                    float* colsum, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * K3  fused edge kernel (forward): for every destination row v
 *        e_uv  = leaky_relu(p_u + q[bin_uv])                         GATLayer.py:88-92 / 127-131
 *        sh_v  = sum_act softmax(e)_uv z_u   over ALL in-edges       GATLayer.py:98-102 via pull :113/:149
 *                (extra[v] never-written edges contribute exp(0) to the denominator, z = 0)
 *        x_v   = elu(sh_v) + origin_v                                GAT.py:56-57   (optional)
 *     zp row u = [ perm(z_u) (fp) | p_u (H) | pad ],  leading dimension ldz (hsg_edge_layout).
 *     stat row v = [ m (H) | den (H) | s (H, written by hsg_edge_bwd_prep) ].
 * ------------------------------------------------------------------------ */
/* Lane-interleaved row layout of the gathered tensors (zp, g, dzp), see csrc/hsg_edge_layout.cuh:
 * fp = permuted width of the F = H*d feature block (>= F), ldz = round_up(fp + H, 8) = leading dimension
 * of zp / dzp / rows of W_aug.  hsg_edge_perm maps an original column (head-major, the order of
 * torch.cat(head_outs), GATStackLayer.py:59) to its position inside a gathered row. */
int hsg_edge_layout(int H, int d, int* host_fp, int* host_ldz);
int hsg_edge_perm(int H, int d, int col);
/* Row mapping of hsg_edge_fwd when a warp holds several GROUP-lane groups: -1 auto (each group walks its own
 * destination row when there are >= 16 384 rows - large shards; the groups share one row's edge list otherwise),
 * 0 never, 1 whenever the layout allows.  Same results up to summation order (tested). */
int hsg_set_edge_fwd_rowpar(int mode);
/* Forward mapping for WIDE rows of LOW degree (layouts with one lane group per warp and the staged epilogue, e.g. the
 * S2W default (6,50) on word rows): source rows and the origin row prefetched one destination ahead, straight-line
 * softmax for rows of at most two in-edges.  -1 auto (>= 16 384 destination rows of average in-degree <= 4), 0 never,
 * 1 whenever the layout allows.  Same results up to summation order (tested). */
int hsg_set_edge_fwd_lowdeg(int mode);
int hsg_edge_fwd(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q,
                 const float* origin /* [n_dst,F] or NULL */, float* sh /* [n_dst,F]; NULL: not stored (x != NULL) */,
                 float* x /* [n_dst,F] or NULL */, float* stat /* [n_dst,3H] */, void* stream);
/* K5a: g = dx * elu'(sh) (or g = dsh when dx == NULL) written lane-interleaved [n_dst, fp]; s[v,k] = g_v[k] . sh_v[k]. */
int hsg_edge_bwd_prep(int n_dst, int H, int d, const float* dx /* or NULL */, const float* dsh /* or NULL */,
                      const float* sh, float* g /* [n_dst,fp] */, float* stat, void* stream);
/* K5b: source-centric backward over the TRANSPOSED structure (csc_t: rows = forward
 * sources, nbr = forward destinations; in HSG/HDSG graphs this is the other
 * direction's CSC because every w->s edge has an s->w twin, dataloader.py:254-257):
 *     dzp_u = [ perm(sum_e alpha_e g_v) | sum_e dpre_e | 0 ],   dq[b,k] = sum_{e: bin=b} dpre_e
 * dq is reduced deterministically through ws. */
size_t hsg_edge_bwd_workspace_bytes(int H);
/* Row mapping of hsg_edge_bwd when a warp holds several GROUP-lane groups ((8,8): two): -1 auto (each group walks its
 * own row when the average degree is low - word rows; the groups share one row's edge list otherwise), 0 never,
 * 1 always.  Same results either way (tested); a tuning / test knob. */
int hsg_set_edge_rowpar(int mode);
/* CTA-per-row mapping of hsg_edge_bwd for FEW high-degree rows (<= 8 192 rows, average degree above 4 per group -
 * the supernode rows of a small batch): the eight warps of a CTA share one row's edge list and their partial rows are
 * summed in warp order.  -1 auto, 0 never, 1 always (overrides the warp-per-row mappings). */
int hsg_set_edge_blockrow(int mode);
/* Asynchronous-gather mapping of hsg_edge_bwd (layouts with one lane group per warp, e.g. the S2W default (6,50)):
 * neighbour rows are copied global -> shared with cp.async into a per-warp ring several rows ahead, across row
 * boundaries.  Opt-in only: -1 auto (= never: measured slower than the register-gather kernel, see hsg_edge.cu),
 * 0 never, 1 whenever the layout allows (and neither the row-parallel nor the CTA-per-row mapping is forced). */
int hsg_set_edge_bwd_async(int mode);
int hsg_edge_bwd(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q,
                 const float* g, const float* stat, float* dzp, float* dq /* [10,H] */,
                 void* ws, size_t ws_bytes, void* stream);

/* K5a': backward prep that RECOMPUTES sh from the saved softmax state instead of reading it (csrc/hsg_edge_rc.cu):
 * walks the layer's FORWARD csc, gathers the few source rows [z | p] of every destination (L1 / L2 resident on
 * word rows), forms sh, g = dx * elu'(sh) (lane-interleaved [n_dst, fp]) and s = g . sh (stat[:, 2H:3H]).  With it
 * hsg_edge_fwd is called with sh = NULL: the forward's store of sh and the prep's read of it - 2 x n_dst x F x 4
 * bytes that SURVEY.md 8(d)'s algorithmic byte counts do not contain - disappear.  Layouts with one lane group per
 * warp and F % 4 == 0 (the S2W default (6,50)); hsg_edge_bwd_prep_rc_ok tells.  Same g / s as hsg_edge_bwd_prep up to
 * the rounding of the recomputed sh (<= 2e-6 normalised, tested).  hsg_set_edge_recompute: what the update loop
 * picks: -1 auto (>= 65 536 destination rows with at most 2 in-edges on average), 0 never, 1 whenever the layout
 * allows; set it before the forward. */
int hsg_set_edge_recompute(int mode);
int hsg_edge_bwd_prep_rc_ok(int H, int d, int ldz);
int hsg_edge_bwd_prep_rc(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q,
                         const float* dx /* [n_dst,F] */, float* g /* [n_dst,fp] */, float* stat, void* stream);

/* ------------------------------------------------------------------------
 * K4  position-wise FFN pieces (GATLayer.py:35-44); the two products use hsg_gemm_*.
 *   hsg_layernorm_fwd: y = LN(r) * gamma + beta, stats[row] = (mean, rstd)
 *   hsg_layernorm_bwd: dr from dy; dgamma/dbeta reduced deterministically through ws.
 * ------------------------------------------------------------------------ */
int hsg_layernorm_fwd(int N, int D, const float* r, const float* gamma, const float* beta, float* y,
                      float* stats, void* stream);
size_t hsg_layernorm_bwd_workspace_bytes(int N, int D);
int hsg_layernorm_bwd(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma,
                      float* dr, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* Whole PositionwiseFeedForward.forward (module/GATLayer.py:35-44) of a SMALL node set in ONE launch each way
 * (exact fp32): y = LayerNorm(W2 relu(W1 x + b1) + b2 + x), leaving hdn [n, d_hid], r [n, F] and stats [n, 2] behind
 * exactly as hsg_gemm_nt + hsg_gemm_nt + hsg_layernorm_fwd would; the backward produces dr, dhp, dx and
 * (accumulate != 0: adds to) dgamma / dbeta; the weight-gradient products dW2 = dr^T hdn, dW1 = dhp^T x stay with
 * hsg_gemm_tn.  hsg_ffn_rows_ok: 1 when the shape qualifies (F == 64, d_hid % 64 == 0, d_hid <= 1024, n below the
 * small-product threshold of hsg_set_gemm_small_flops) - the sentence side of the 32-graph step; the update-loop
 * entry points pick it by themselves, anything else returns HSG_ERR_SHAPE here.  ws as for hsg_layernorm_bwd. */
int hsg_ffn_rows_ok(int n, int F, int d_hid);
int hsg_ffn_rows_fwd(int n, int F, int d_hid, const float* x, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* gamma, const float* beta, float* hdn, float* r, float* y,
                     float* stats, void* stream);
int hsg_ffn_rows_bwd(int n, int F, int d_hid, const float* dy, const float* r, const float* stats, const float* gamma,
                     const float* hdn, const float* w1, const float* w2, float* dr, float* dhp, float* dx,
                     float* dgamma, float* dbeta, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Coarse-grained entry points: every kernel of ONE WSWGAT application (module/GAT.py:45-59), forward or
 * backward, from one call.  W_aug / q come from hsg_attn_prep_fwd (once per layer and step: the update loop
 * of HiGraph.py:98-106 re-applies the same two weight sets); dW_aug / dq are summed over the applications by
 * the caller and go through hsg_attn_prep_bwd once.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t H, d, in_dim, d_hid, n_src, n_dst, ldz, reserved;
  const hsg_csc* csc;                  /* in-edges of the destination rows */
  const float *neighbor, *origin;      /* [n_src,in_dim], [n_dst,H*d] */
  const float *W_aug, *q;              /* [ldz,in_dim], [10,H] */
  const float *w1, *b1, *w2, *b2, *gamma, *beta;   /* FFN: [d_hid,F] [d_hid] [F,d_hid] [F] [F] [F] */
  float *zp, *sh, *x, *stat, *hdn, *r, *ln_stats;  /* saved for backward */
  float* out;                          /* [n_dst,F] */
} hsg_wswgat_fwd_args;
int hsg_wswgat_fwd(const hsg_wswgat_fwd_args* args, void* stream);

typedef struct {
  int32_t H, d, in_dim, d_hid, n_src, n_dst, ldz, reserved;
  const hsg_csc* csc_t;                /* transposed structure (rows = forward sources) */
  const float* dout;                   /* [n_dst,F] */
  const float *neighbor, *W_aug, *q, *w1, *w2, *gamma;
  const float *zp, *sh, *x, *hdn, *r, *ln_stats;
  float* stat;                         /* (m, den) from forward; s is written here */
  float *dr, *dhp, *g, *dzp;           /* scratch: [n_dst,F] [n_dst,d_hid] [n_dst,fp] [n_src,ldz] */
  float *dx;                           /* [n_dst,F]  = d origin */
  float *d_neighbor;                   /* [n_src,in_dim] */
  float *dW_aug, *dq;                  /* [ldz,in_dim], [10,H] */
  float *dw1, *db1, *dw2, *db2, *dgamma, *dbeta;
  void* ws;
  size_t ws_bytes;
} hsg_wswgat_bwd_args;
size_t hsg_wswgat_bwd_workspace_bytes(int H, int d, int in_dim, int d_hid, int n_src, int n_dst);
int hsg_wswgat_bwd(const hsg_wswgat_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------
 * Whole update loop from one call each way: HSumGraph.forward's
 *     sent = W2S(word, sent);  n_iter x { word = S2W(word, sent);  sent = W2S(word, sent) }
 * (HiGraph.py:98-106, HSumDocGraph :205-214), attention prep included.  The two weight sets are shared by
 * the 1 + 2 n_iter applications: their gradients are accumulated inside the producing kernels' last stage,
 * and the state gradients that meet at a node set (origin path of the later application + neighbor path of
 * the earlier one, GAT.py:47-57) are summed in the projection-backward epilogue - no separate add launches.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t H, d, in_dim, feat_dim, d_hid, reserved;
  const float *W, *Wf, *bf, *a;  /* heads packed: fc [H*d,in_dim], feat_fc [H*d,feat_dim], its bias [H*d] or NULL, attn_fc [H,3d] */
  const float *w1, *b1, *w2, *b2, *gamma, *beta; /* FFN + LayerNorm (GATLayer.py:25-44) */
} hsg_layer_params;

typedef struct {
  float *dW, *dWf, *dbf, *da, *dw1, *db1, *dw2, *db2, *dgamma, *dbeta;
} hsg_layer_grads;

typedef struct {
  int32_t n_apps;       /* number of WSWGAT applications, >= 1; HSG / HDSG: 1 + 2 n_iter */
  int32_t start_kind;   /* kind of application 0: 0 = W2S (HSG / HDSG), 1 = S2W; kinds alternate from there, so
                           (n_apps = 1, start_kind) is one stand-alone WSWGAT.forward (GAT.py:45-59) */
  int32_t n_word, n_super;
  const hsg_csc *csc_super, *csc_word;  /* in-edges of supernodes from words / of words from supernodes */
  hsg_layer_params w2s, s2w;            /* a layer that no application uses is ignored */
  const float* T;                       /* _TFembed.weight [10, feat_dim]  (HiGraph.py:52) */
  const float *word_feature;            /* [n_word,  word dim = s2w.H * s2w.d = w2s.in_dim]   */
  const float *super_feature;           /* [n_super, hidden   = w2s.H * w2s.d = s2w.in_dim]   */
  float* state;                         /* forward arena: everything backward needs + the two results */
  size_t state_floats;
  /* training-mode dropout (0 = off): attn_p on the layer input, drawn independently per head
   * (GATStackLayer.py:56); ffn_p on the FFN output before the residual (GATLayer.py:41-42).  Masks are a pure
   * function of (seed, application index, element), regenerated in backward. */
  float attn_p, ffn_p;
  unsigned long long seed;
  /* optional DEVICE step counter mixed into the mask key at run time (NULL: masks depend on `seed` only).  A CUDA
   * graph replays identical kernel arguments every step; with this pointer (advanced by hsg_adam_step_dev) every
   * replay still draws fresh masks, identical in forward and backward of the same step. */
  const unsigned long long* seed_dev;
  /* optional cudaEvent_t (NULL: none): word_feature / super_feature are being produced on ANOTHER stream (e.g. the
   * embedding gather, HiGraph.py:147-148) and this event marks their completion.  hsg_update_loop_fwd enqueues the
   * parameter-only attention prep first and makes `stream` wait for the event right before application 0, so the two
   * overlap; backward ignores it. */
  void* input_ready;
} hsg_loop_args;

typedef struct {
  size_t state_floats;    /* size of hsg_loop_args.state                                 */
  size_t scratch_floats;  /* size of hsg_loop_bwd_args.scratch                           */
  size_t ws_bytes;        /* size of hsg_loop_bwd_args.ws                                */
  size_t word_state_off;  /* float offset of the final word state in `state`; (size_t)-1: no application produced one
                             (it is word_feature) */
  size_t super_state_off; /* same for the supernode state ((size_t)-1: it is super_feature) */
  size_t hdn_off[2];      /* float offset of the FFN hidden activation of application 0 / 1 - test hook */
  size_t pair_stride;     /* floats between application i and i+2 (same layer type) in `state` */
} hsg_loop_plan;

typedef struct {
  const float *d_word_state, *d_super_state; /* upstream gradients; either may be NULL (= 0) */
  float* d_word_feature;                     /* [n_word, word dim] or NULL (skips that product) */
  float* d_super_feature;                    /* [n_super, hidden] or NULL */
  hsg_layer_grads w2s, s2w;
  float* dT;                                 /* [10, feat_dim] */
  int32_t accumulate;                        /* 1: parameter gradients are ADDED to the given buffers (fused
                                                accumulation into .grad), 0: overwritten */
  int32_t reserved;
  float* scratch;
  size_t scratch_floats;
  void* ws;
  size_t ws_bytes;
} hsg_loop_bwd_args;

/* hsg_update_loop_bwd forks the weight-gradient products (dW2, dW1, dW_aug of every application) onto an internal
 * second stream, joined by events before the call's last kernels, so they overlap the serial dx -> edge backward ->
 * d_neighbor chain (default on; HSG_BWD_OVERLAP=0 or hsg_set_bwd_overlap(0): everything on the caller's stream).
 * Same kernels and per-buffer order either way: bitwise identical results. */
int hsg_set_bwd_overlap(int on);
/* SMs the side-stream weight-gradient products may occupy (0 = all = default, or HSG_SIDE_CTAS).  Tuning knob: on
 * the 32-graph step fewer SMs measured slower at every setting (profiles/r02c_sweep.jsonl). */
int hsg_set_side_ctas(int n);
/* 1 (default; HSG_GEMM_PAIR=0): NT / NN tensor-core products run on CTA pairs (tcgen05 cta_group::2, 256-row tiles,
 * B tile shared by the two SMs); 0: the single-CTA kernel.  Bit-identical results either way. */
int hsg_set_gemm_pair(int on);
/* shortest reduction range (rows) one split of a tensor-core weight-gradient product may have (default 256) */
int hsg_set_tn_min_rows(int rows);
/* rows > 0 (or HSG_TN_ITEM_ROWS): weight-gradient products whose reduction fits in <= 48 splits of `rows` rows are cut
 * into tiles x splits items of that length and launched with one CTA per item instead of one persistent CTA per SM, so
 * that the SMs go back to the block scheduler (and to the higher-priority main stream) every item; 0: persistent plan. */
int hsg_set_tn_item_rows(int rows);
/* Sizes/offsets for the given dimensions (pointers inside `a` are not read). */
int hsg_update_loop_plan(const hsg_loop_args* a, hsg_loop_plan* plan);
int hsg_update_loop_fwd(const hsg_loop_args* a, void* stream);
int hsg_update_loop_bwd(const hsg_loop_args* a, const hsg_loop_bwd_args* b, void* stream);
/* Test hook: keep flags (1/0) of the dropout mask of `n` consecutive element indices for (p, seed, stream_id).
 * Stream ids used by the loop: 2*app for the attention-input mask (element ((head * n_src) + row) * in_dim + col),
 * 2*app + 1 for the FFN mask (element row * F + col). */
int hsg_dropout_mask(size_t n, float p, unsigned long long seed, unsigned int stream_id, unsigned char* out,
                     void* stream);

/* ------------------------------------------------------------------------
 * Readout, loss, extraction and optimizer on the device (the step right after the update loop):
 *   logits = wh(sentence state)                       HiGraph.py:108; HDSG: wh(cat(sentence, its document)), :216-228
 *   loss   = mean_graphs sum_sentences CE(logits, y)  train.py:114-119
 *   top-m  = per graph torch.topk(logits[:, 1], m)    Tester.py:128
 *   Adam (+ optional clip_grad_norm_)                 train.py:90,132-135
 * All reductions run in a fixed order.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_sent, n_super, hidden, two_part; /* two_part = 1 (HDSG): wh input is [sentence | document], 2*hidden wide */
  int32_t n_graphs, reserved;
  const float* state;            /* [n_super, hidden] final supernode state */
  const int32_t* sent_row;       /* [n_sent] supernode row of every sentence; NULL = identity (HSG: every supernode is a sentence) */
  const int32_t* doc_row;        /* [n_sent] supernode row of the sentence's document (two_part only) */
  const int32_t* graph_sent_ptr; /* [n_graphs+1] sentence offsets per graph (two_part only) */
  const float* wh_w;             /* [2, hidden * (1 + two_part)] */
  const float* wh_b;             /* [2] */
  const int64_t* labels;         /* [n_sent] in {0, 1} */
  float inv_graphs;              /* 1 / (global number of graphs) */
  float reserved2;
} hsg_head_args;
size_t hsg_head_workspace_bytes(int n_sent, int width);
/* logits [n_sent,2]; dlogits [n_sent,2] = d loss / d logits (saved for hsg_head_bwd); loss [1] */
int hsg_head_fwd(const hsg_head_args* a, float* logits, float* dlogits, float* loss, void* ws, size_t ws_bytes,
                 void* stream);
/* gout: device scalar d L / d loss, or NULL (= 1).  d_state [n_super, hidden] is fully written. */
int hsg_head_bwd(const hsg_head_args* a, const float* dlogits, const float* gout, float* d_state, float* d_wh_w,
                 float* d_wh_b, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* hsg_head_fwd followed by hsg_head_bwd(gout = NULL) in one launch (the training step, train.py:114-121, where the
 * loss is the root of backward); every output is bit-identical to the two calls. */
int hsg_head_fwd_bwd(const hsg_head_args* a, float* logits, float* dlogits, float* loss, float* d_state, float* d_wh_w,
                     float* d_wh_b, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* out_idx [n_graphs, m]: local sentence indices by descending class-1 logit (ties: lower index first), -1 padded */
int hsg_topm(const float* logits, const int32_t* graph_sent_ptr, int n_graphs, int m, int32_t* out_idx, void* stream);
/* torch.optim.Adam semantics (no amsgrad / weight decay) on flat fp32 arrays, `step` counts from 1.
 * max_grad_norm > 0 applies clip_grad_norm_'s coefficient min(1, max_norm / (||g|| + 1e-6)) on the fly (needs ws). */
size_t hsg_adam_workspace_bytes(void);
int hsg_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, float max_grad_norm, void* ws, size_t ws_bytes, void* stream);
/* The same update with the step number kept on the DEVICE: step_state[0] = number of completed steps (the update uses
 * step_state[0] + 1 and then advances it), step_state[1] = internal block ticket (must start at 0).  Identical
 * arguments every step, so the launch can be replayed from a CUDA graph (train.py:131-135 as one captured step).
 * zero_grad != 0 clears `grad` after use (the optimizer.zero_grad() of train.py:130 folded in).  The same counter is
 * what hsg_loop_args.seed_dev reads for the dropout masks. */
int hsg_adam_step_dev(size_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                      float beta2, float eps, unsigned long long* step_state, int zero_grad, float max_grad_norm,
                      void* ws, size_t ws_bytes, void* stream);
/* Gradient all-reduce (sum over the data-parallel ranks of one box) + Adam + zero_grad in ONE kernel over NVLink peer
 * memory (csrc/hsg_head.cu: push into every rank's receive slots, system-scope flags, reduce in rank order from local
 * memory): replaces ncclAllReduce + hsg_adam_step_dev for the small (latency-bound) gradient arena of the path.
 * Every rank owns a symmetric buffer of hsg_allreduce_adam_buffer_floats(n, world) floats, ZEROED once before the
 * first step; peer_bufs is a DEVICE array of the `world` base pointers (peer mappings, rank order, own buffer
 * included).  step_state: [4] device u64, [0] = completed steps (the same on every rank), the rest kernel-internal and
 * zero.  n must keep 16-byte alignment of the slots (n % 4 == 0).  Every rank must enqueue the call once per step;
 * bit-identical results on all ranks (fixed summation order). */
size_t hsg_allreduce_adam_buffer_floats(size_t n, int world);
int hsg_allreduce_adam_step(size_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                            float beta1, float beta2, float eps, unsigned long long* step_state,
                            float* const* peer_bufs, int rank, int world, void* stream);
/* out[i, :] = table[ids[i], :]: the frozen word-embedding lookup of set_wnfeature (HiGraph.py:147-148).  dim % 4 == 0. */
int hsg_embed_gather(int n, int dim, const int32_t* ids, const float* table, float* out, void* stream);

/* ------------------------------------------------------------------------
 * S2S layer type: SGATLayer / MultiHeadSGATLayer (module/GATLayer.py:49-78, module/GATStackLayer.py:27-44), the
 * "S2S" branch of WSWGAT (module/GAT.py:38-39,50-52).  Never instantiated by the reference's models; built for
 * completeness.  z = fc(h) (hsg_gemm_nt by the caller, head-major columns);
 *     sh_v = mult * sum_{j: xmember[j] == xgrp[v]} z_j / (deg_v exp(leaky_relu(a[d:2d] . z_v)) + extra_v)
 * (see csrc/hsg_s2s.cu for the derivation from DGL-0.4's pull over all in-edges); x = elu(sh) + origin (optional).
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_graphs, n_super, H, d, mult, reserved;
  const int32_t* super_ptr;   /* [n_graphs+1] supernode rows of every graph (contiguous) */
  const int32_t* deg_indptr;  /* [n_super+1] indptr of the word->supernode CSC (deg_v = word in-degree) */
  const int32_t* extra;       /* [n_super] number of non-word in-edges x_v */
  const int32_t* xgrp;        /* [n_super] row id of the group whose member sum v reads, -1: none */
  const int32_t* xmember;     /* [n_super] row id of the group v contributes its z to, -1: none */
} hsg_s2s_graph;
/* S [n_super, H*d]: group sums (saved for backward) */
int hsg_s2s_fwd(const hsg_s2s_graph* g, const float* z, const float* a /* [H, 2d] */, const float* origin /* or NULL */,
                float* S, float* sh, float* x /* or NULL */, void* stream);
size_t hsg_s2s_bwd_workspace_bytes(int n_graphs, int H, int d);
/* dx (gradient of x) or dsh (gradient of sh) given; dS: scratch [n_super, H*d]; dz [n_super, H*d]; da [H, 2d] */
int hsg_s2s_bwd(const hsg_s2s_graph* g, const float* z, const float* a, const float* S, const float* dx,
                const float* dsh, float* dS, float* dz, float* da, int accumulate, void* ws, size_t ws_bytes,
                void* stream);

/* ------------------------------------------------------------------------
 * HDSG document-node init: HSumDocGraph.forward / set_dnfeature (HiGraph.py:196-203, 231-244).
 *   doc_mean[j]   = mean of the init features of document j's sentences          hsg_doc_mean
 *   doc_feature   = dn_feature_proj(doc_mean)                                     hsg_gemm_nt (caller)
 *   super_feature = sentence / document rows interleaved per graph               hsg_super_assemble
 * backward: hsg_doc_init_bwd in two phases around the caller's projection-backward products.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_sent, n_doc, hidden, reserved;
  const int32_t* sent_row;        /* [n_sent] supernode row of every sentence */
  const int32_t* doc_row;         /* [n_doc]  supernode row of every document */
  const int32_t* sent_doc;        /* [n_sent] GLOBAL document index of every sentence */
  const int32_t* doc_graph;       /* [n_doc]  graph index of every document */
  const int32_t* graph_sent_ptr;  /* [n_graphs+1] */
} hsg_doc_map;
int hsg_doc_mean(const hsg_doc_map* m, const float* sent_feature, float* doc_mean, void* stream);
int hsg_super_assemble(const hsg_doc_map* m, const float* sent_feature, const float* doc_feature,
                       float* super_feature, void* stream);
/* phase 1 (d_doc_mean == NULL): d_doc_feature[j] = d_super[doc_row[j]];
 * phase 2: d_sent[i] = d_super[sent_row[i]] + d_doc_mean[doc(i)] / #sentences(doc(i)) */
int hsg_doc_init_bwd(const hsg_doc_map* m, const float* d_super, const float* d_doc_mean, float* d_doc_feature,
                     float* d_sent, void* stream);

/* ------------------------------------------------------------------------
 * Sentence encoder in front of the path: the n-gram CNN of module/Encoder.py:56-76
 * (sentEncoder.forward), called from HSumGraph._sent_cnn_feature, HiGraph.py:127-133.
 *   x[s,t,:]   = embed[tok[s,t]] + pos_table[t < len_s ? t+1 : 0]                       Encoder.py:58-69
 *   out[s, (h-2)*50 + c] = max_t relu(conv_h(x)[s,c,t]),  h = 2..7, 50 channels each    Encoder.py:71-73
 * Sentences are stored COMPACT: sentence s keeps its first n_s = min(tail_s + 7, L) rows (tail_s = index after its
 * last non-zero id; every window behind it sees the same PAD row and gives the same value), rows back to back in
 * xc [n_rows + 8, D] (8 zeroed tail rows), row_ptr [n_sent+1].  The six convolutions are the product
 *   y [n_rows, 312] = A [n_rows, 7D] . wpad [312, 7D]^T   with A = xc and lda = D (overlapping rows, no im2col),
 * wpad = the kernels zero-padded to height 7 (hsg_enc_pack_weights); height h occupies columns (h-2)*52 .. +49
 * (52-column groups start 16-byte aligned).  The caller issues it as K-chunks of hsg_gemm_nt over kernel rows
 * {0,1}, {2,3}, {4,5}, {6}, each restricted to the column groups whose kernels reach those rows and accumulated
 * in place (HSG_EPI_ADD with R == C).  D % 4 == 0, L >= 7.
 * Backward (frozen embedding, train.py:340-342): only the kernels and biases get gradients; d out / d y is one-hot
 * per (sentence, channel), so dW is a sparse accumulation of <= n_sent*300 input slabs in a fixed order.
 * ------------------------------------------------------------------------ */
/* HOST function (all pointers are host memory, no device work): per sentence the number of non-zero ids
 * (sent_len, Encoder.py:58), the compact row offsets (row_ptr [n_sent+1], row_ptr[n_sent] = n_rows) and the position
 * of the sentence inside its graph, 1-based (sent_pos, dataloader.py:241).  Replaces the reference's per-sentence host
 * loop with a device sync per sentence (Encoder.py:61-66). */
int hsg_enc_plan_host(int n_sent, int L, const int32_t* tokens, int n_graphs, const int32_t* graph_sent_ptr,
                      int32_t* sent_len, int32_t* row_ptr, int32_t* sent_pos);
int hsg_enc_gather(int n_sent, int L, int D, int n_rows, const int32_t* tokens /* [n_sent, L] */,
                   const int32_t* sent_len /* [n_sent] non-zero ids, Encoder.py:58 */, const int32_t* row_ptr,
                   const float* embed, const float* pos_table /* [L+1, D] */, float* xc, void* stream);
/* conv_w[i]: [50, 1, i+2, D] (Conv2d weight of height i+2) -> wpad [312, 7D] */
int hsg_enc_pack_weights(int D, const float* const* conv_w, float* wpad, void* stream);
/* out[s, (h-2)*50 + c] = relu(max over the valid windows of y[row_ptr[s]+t, (h-2)*52 + c] + bias);
 * arg_t [300, n_sent] = compact row of the maximum (first one), -1 where the ReLU is inactive */
int hsg_enc_pool_fwd(int n_sent, const int32_t* row_ptr, const float* y, int ldy, const float* const* conv_b,
                     float* out, int ldo, int32_t* arg_t, void* stream);
size_t hsg_enc_conv_wgrad_workspace_bytes(int n_sent, int D);
/* d_conv_w[i] [50, 1, i+2, D], d_conv_b[i] [50]: written (accumulate = 0) or added to (accumulate = 1) */
int hsg_enc_conv_wgrad(int n_sent, int D, const float* xc, const float* d_out, int ldo, const int32_t* arg_t,
                       float* const* d_conv_w, float* const* d_conv_b, int accumulate, void* ws, size_t ws_bytes,
                       void* stream);
/* out[r,:] = x[r,:] + table[idx[r],:]   (ngram_feature + sent_pos_embed(position), HiGraph.py:130-132) */
int hsg_add_rows(int n, int D, const float* x, int ldx, const int32_t* idx, const float* table, float* out, int ldo,
                 void* stream);

/* ------------------------------------------------------------------------
 * Recurrent part of the sentence-level (Bi)LSTM - torch.nn.LSTM on the packed per-graph sentence sequences,
 * HiGraph.py:118-119,135-142 (get_snode_feat :247-255).  One layer per call; gate order i, f, g, o.
 * Rows of graph b are graph_sent_ptr[b] .. graph_sent_ptr[b+1]-1 (batched sentence order), direction 1 walks them
 * backwards.  The caller computes the input products of all time steps with hsg_gemm_nt:
 *   xproj [S, ndir*4H] (direction d in columns d*4H .., NO bias: b_ih + b_hh are added here).
 * forward : out [S, ndir*H] (= cat(h_fwd, h_bwd), the layer output), and saved for backward:
 *           gates [S, ndir, 4H] (post-activation), cst [S, ndir, H] (cell states), hprev [S, ndir, H] (h_{t-1}).
 * backward: da [S, ndir*4H] = gradient of the pre-activation gates (xproj layout); the caller finishes with GEMMs:
 *           dW_ih = da_d^T x, dW_hh = da_d^T hprev_d, db_ih = db_hh = colsum(da_d), dx = sum_d da_d W_ih_d.
 * H <= 128, H % 4 == 0; ndir 1 or 2; w_hh / b_ih / b_hh: arrays of ndir device pointers.
 * ------------------------------------------------------------------------ */
int hsg_lstm_fwd(int n_graphs, int H, int ndir, const int32_t* graph_sent_ptr, const float* xproj,
                 const float* const* w_hh, const float* const* b_ih, const float* const* b_hh, float* out,
                 float* gates, float* cst, float* hprev, void* stream);
int hsg_lstm_bwd(int n_graphs, int H, int ndir, const int32_t* graph_sent_ptr, const float* d_out, const float* gates,
                 const float* cst, const float* const* w_hh, float* da, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HSG_B200_H_ */
