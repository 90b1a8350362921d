// K0: device-side heterogeneous-graph builder (integer work, bit-exact contract).
//
// Replaces the reference's per-example CPU graph construction and batching:
//   ExampleSet.AddWordNode            module/dataloader.py:201-220
//   ExampleSet.CreateGraph            module/dataloader.py:222-268   (HSG)
//   MultiExampleSet.CreateGraph       module/dataloader.py:328-406   (HDSG)
//   dgl.batch in graph_collate_fn     module/dataloader.py:480
// and the three filter_nodes/filter_edges UDF scans every head performs
// (module/GATLayer.py:105-107, 143-145): the id sets they return are emitted once
// per batch as row maps + two CSCs.
//
// One CTA per document graph.  DGL numbering that must be reproduced exactly:
//   * word node ids: distinct unfiltered token ids in first-occurrence order over
//     the padded sentences; then sentence nodes; then document nodes.
//   * per sentence i the reference walks Counter(tokens_i).keys() (first occurrence
//     inside the sentence) and adds the pair (w->s, s->w) for every word that is a
//     node and a TF-IDF key, then (HSG) 2N sent<->sent edges / (HDSG) one s->d edge:
//        HSG : base_i = sum_{j<i} (2 k_j + 2N),  HDSG: base_i = sum_{j<i} (2 k_j + 1)
//        eid(w->s, t-th) = base_i + 2t,  eid(s->w) = base_i + 2t + 1
//     HDSG then adds the (w->d, d->w) pairs per document in Counter(doc tokens) order.
//   * dgl.batch offsets node / edge ids by the cumulative counts of the earlier graphs.
#include <limits.h>

#include "hsg_common.cuh"

namespace hsg {

constexpr int BLD_THREADS = 256;
constexpr int BLD_WARPS = BLD_THREADS / 32;

struct BuildWs {
  int32_t* cnt_word;   // [B]
  int32_t* cnt_super;  // [B]
  int32_t* cnt_node;   // [B]
  int32_t* cnt_edge;   // [B]
  int32_t* cnt_pair;   // [B]
  int32_t* k_super;    // [S + D]  pairs per sentence (global sentence index), then per doc
  int32_t* pos_info;   // [S * L]  local word nid of an edge-producing token position, else -1
  int32_t* doc_info;   // [T]
  int32_t* wid_local;  // [S * L]  vocabulary id of local word node u of graph g at [s0 * L + u]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t ws_layout(const hsg_token_batch* tb, void* base, BuildWs* w) {
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off = align_up(off + n * sizeof(int32_t), 16);
    return base ? reinterpret_cast<int32_t*>(reinterpret_cast<char*>(base) + o) : nullptr;
  };
  const size_t B = tb->n_graphs, S = tb->n_sent, L = tb->sent_len, D = tb->n_doc, T = tb->n_doc_tok;
  int32_t* p;
  p = take(B); if (w) w->cnt_word = p;
  p = take(B); if (w) w->cnt_super = p;
  p = take(B); if (w) w->cnt_node = p;
  p = take(B); if (w) w->cnt_edge = p;
  p = take(B); if (w) w->cnt_pair = p;
  p = take(S + D); if (w) w->k_super = p;
  p = take(S * L); if (w) w->pos_info = p;
  p = take(T); if (w) w->doc_info = p;
  p = take(S * L); if (w) w->wid_local = p;
  return off + 16;
}

__device__ __forceinline__ bool is_filtered(const uint32_t* bm, int vocab, int wid) {
  if (wid < 0 || wid >= vocab) return true;
  return (__ldg(bm + (wid >> 5)) >> (wid & 31)) & 1u;
}

__device__ __forceinline__ uint32_t hash_slot(int wid, uint32_t mask) {
  uint32_t h = (uint32_t)wid * 2654435761u;
  h ^= h >> 15;
  return h & mask;
}

// exclusive scan of one int per thread across the block; returns the exclusive prefix and sets *total
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* warp_buf /* [BLD_WARPS + 1] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect warp_buf reuse
  if (lane == 31) warp_buf[w] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < BLD_WARPS; ++i) {
    int c = warp_buf[i];
    if (i < w) woff += c;
    tot += c;
  }
  *total = tot;
  return woff + inc - v;
}

__device__ __forceinline__ int table_find(const int32_t* keys, uint32_t mask, int wid) {
  uint32_t s = hash_slot(wid, mask);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    int k = keys[s];
    if (k == wid) return (int)s;
    if (k == -1) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}

// ---------------------------------------------------------------------------
// Phase 1: per-graph word-node numbering, edge-producing positions, counts
// dynamic smem: keys[HT] | vals[HT] | first_pos[cap_tok] (HDSG)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLD_THREADS)
build_count_kernel(hsg_token_batch tb, BuildWs ws, int ht_size, int cap_tok, int32_t* status) {
  pdl_prologue();
  extern __shared__ int32_t smem[];
  __shared__ int warp_buf[BLD_WARPS + 1];
  __shared__ int s_pairs;
  int32_t* keys = smem;
  int32_t* vals = smem + ht_size;
  int32_t* first_pos = smem + 2 * ht_size;
  const uint32_t mask = (uint32_t)ht_size - 1u;
  const int g = blockIdx.x;
  const int L = tb.sent_len;
  const int s0 = __ldg(tb.graph_sent_ptr + g), s1 = __ldg(tb.graph_sent_ptr + g + 1);
  const int n = s1 - s0;
  const int ntok = n * L;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  if (ntok > cap_tok) {   // host sizes the tables from max_sent_per_graph
    if (tid == 0) atomicMin(status, (int)HSG_ERR_CAPACITY);
    if (tid == 0) {
      ws.cnt_word[g] = 0; ws.cnt_super[g] = 0; ws.cnt_node[g] = 0; ws.cnt_edge[g] = 0; ws.cnt_pair[g] = 0;
    }
    return;
  }
  for (int i = tid; i < ht_size; i += BLD_THREADS) {
    keys[i] = -1;
    vals[i] = INT_MAX;
  }
  if (tid == 0) s_pairs = 0;
  __syncthreads();
  const int32_t* tok = tb.tokens + (size_t)s0 * L;
  // 1. min position of every distinct unfiltered id
  for (int t = tid; t < ntok; t += BLD_THREADS) {
    const int wid = __ldg(tok + t);
    if (wid == 0 || is_filtered(tb.filter_bitmap, tb.vocab_size, wid)) continue;
    uint32_t s = hash_slot(wid, mask);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
      const int prev = atomicCAS(&keys[s], -1, wid);
      if (prev == -1 || prev == wid) {
        atomicMin(&vals[s], t);
        break;
      }
      s = (s + 1) & mask;
    }
  }
  __syncthreads();
  // 2. number the first occurrences in position order (AddWordNode, dataloader.py:205-210)
  int running = 0;
  for (int c = 0; c < ntok; c += BLD_THREADS) {
    const int t = c + tid;
    int flag = 0, slot = -1, wid = 0;
    if (t < ntok) {
      wid = __ldg(tok + t);
      if (wid != 0 && !is_filtered(tb.filter_bitmap, tb.vocab_size, wid)) {
        slot = table_find(keys, mask, wid);
        flag = (slot >= 0 && vals[slot] == t) ? 1 : 0;
      }
    }
    int tot;
    const int ex = block_excl_scan(flag, &tot, warp_buf);
    if (flag) {
      const int nid = running + ex;
      vals[slot] = -1 - nid;                       // from now on the slot holds the node id
      ws.wid_local[(size_t)s0 * L + nid] = wid;
    }
    running += tot;
  }
  const int nw = running;
  __syncthreads();
  // 3. sentence pairs: first occurrence inside the sentence, word is a node, token is a TF-IDF key
  for (int i = wib; i < n; i += BLD_WARPS) {
    const int32_t* st = tok + (size_t)i * L;
    const int8_t* sb = tb.sent_bin + ((size_t)(s0 + i)) * L;
    int cnt = 0;
    for (int c = 0; c < L; c += 32) {
      const int j = c + lane;
      int info = -1;
      if (j < L) {
        const int wid = __ldg(st + j);
        if (wid != 0 && __ldg(sb + j) >= 0 && !is_filtered(tb.filter_bitmap, tb.vocab_size, wid)) {
          bool first = true;
          for (int jj = 0; jj < j; ++jj)
            if (__ldg(st + jj) == wid) {
              first = false;
              break;
            }
          if (first) {
            const int slot = table_find(keys, mask, wid);
            if (slot >= 0) info = -1 - vals[slot];
          }
        }
        ws.pos_info[((size_t)(s0 + i)) * L + j] = info;
      }
      cnt += __popc(__ballot_sync(0xffffffffu, info >= 0));
    }
    if (lane == 0) {
      ws.k_super[s0 + i] = cnt;
      atomicAdd(&s_pairs, cnt);
    }
  }
  __syncthreads();
  // 4. HDSG document pairs: Counter(doc tokens) order (dataloader.py:388-400)
  int nd = 0;
  if (tb.hdsg) {
    const int d0 = __ldg(tb.graph_doc_ptr + g), d1 = __ldg(tb.graph_doc_ptr + g + 1);
    nd = d1 - d0;
    for (int jd = 0; jd < nd; ++jd) {
      const int t0 = __ldg(tb.doc_tok_ptr + d0 + jd), t1 = __ldg(tb.doc_tok_ptr + d0 + jd + 1);
      for (int u = tid; u < nw; u += BLD_THREADS) first_pos[u] = INT_MAX;
      __syncthreads();
      for (int t = t0 + tid; t < t1; t += BLD_THREADS) {
        const int wid = __ldg(tb.doc_tokens + t);
        int info = -1;
        if (wid != 0 && __ldg(tb.doc_bin + t) >= 0 && !is_filtered(tb.filter_bitmap, tb.vocab_size, wid)) {
          const int slot = table_find(keys, mask, wid);
          if (slot >= 0) {
            info = -1 - vals[slot];
            atomicMin(&first_pos[info], t);
          }
        }
        ws.doc_info[t] = info;   // provisional: node id of every candidate position
      }
      __syncthreads();
      int cnt = 0;
      for (int t = t0 + tid; t < t1; t += BLD_THREADS) {
        const int info = ws.doc_info[t];
        if (info >= 0) {
          if (first_pos[info] == t) cnt += 1; else ws.doc_info[t] = -1;
        }
      }
      int tot;
      block_excl_scan(cnt, &tot, warp_buf);
      if (tid == 0) {
        ws.k_super[tb.n_sent + d0 + jd] = tot;
        s_pairs += tot;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int pairs = s_pairs;
    ws.cnt_word[g] = nw;
    ws.cnt_super[g] = n + nd;
    ws.cnt_node[g] = nw + n + nd;
    ws.cnt_pair[g] = pairs;
    ws.cnt_edge[g] = tb.hdsg ? (2 * pairs + n) : (2 * pairs + 2 * n * n);
  }
}

// ---------------------------------------------------------------------------
// Phase 1b: exclusive scans over graphs -> [B+1] offsets
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLD_THREADS)
build_scan_kernel(int B, BuildWs ws, hsg_graph_offsets off) {
  pdl_prologue();
  __shared__ int warp_buf[BLD_WARPS + 1];
  const int32_t* in[5] = {ws.cnt_word, ws.cnt_super, ws.cnt_node, ws.cnt_edge, ws.cnt_pair};
  int32_t* out[5] = {off.word_ptr, off.super_ptr, off.node_ptr, off.edge_ptr, off.pair_ptr};
  for (int a = 0; a < 5; ++a) {
    int running = 0;
    for (int c = 0; c < B; c += BLD_THREADS) {
      const int i = c + threadIdx.x;
      const int v = i < B ? in[a][i] : 0;
      int tot;
      const int ex = block_excl_scan(v, &tot, warp_buf);
      if (i < B) out[a][i] = running + ex;
      running += tot;
    }
    if (threadIdx.x == 0) out[a][B] = running;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Phase 2: node maps + both CSCs.  dynamic smem: wscan[cap_tok + 1] | cursor[cap_tok] | kscan[cap_sup + 1]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLD_THREADS)
build_fill_kernel(hsg_token_batch tb, BuildWs ws, hsg_graph_out out, int cap_tok, int cap_sup) {
  pdl_prologue();
  extern __shared__ int32_t smem[];
  __shared__ int warp_buf[BLD_WARPS + 1];
  int32_t* wscan = smem;
  int32_t* cursor = smem + cap_tok + 1;
  int32_t* kscan = cursor + cap_tok;
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int L = tb.sent_len;
  const int s0 = __ldg(tb.graph_sent_ptr + g), s1 = __ldg(tb.graph_sent_ptr + g + 1);
  const int n = s1 - s0;
  const int w0 = out.off.word_ptr[g], nw = out.off.word_ptr[g + 1] - w0;
  const int sr0 = out.off.super_ptr[g], nsup = out.off.super_ptr[g + 1] - sr0;
  const int nid0 = out.off.node_ptr[g];
  const int e0 = out.off.edge_ptr[g];
  const int p0 = out.off.pair_ptr[g];
  const int nd = nsup - n;
  int d0 = 0;
  if (tb.hdsg) d0 = __ldg(tb.graph_doc_ptr + g);
  if (nw > cap_tok || nsup > cap_sup) return;   // already flagged by phase 1
  if (w0 + nw > out.cap_word || sr0 + nsup > out.cap_super || p0 + (out.off.pair_ptr[g + 1] - p0) > out.cap_pair) {
    if (tid == 0) atomicMin(out.status, (int)HSG_ERR_CAPACITY);
    return;
  }
  // node maps
  for (int u = tid; u < nw; u += BLD_THREADS) {
    out.word_wid[w0 + u] = ws.wid_local[(size_t)s0 * L + u];
    out.word_nid[w0 + u] = nid0 + u;
    wscan[u] = 0;
    cursor[u] = 0;
  }
  if (tid == 0) wscan[nw] = 0;
  // pairs-per-supernode scan (serial: nsup <= ~105)
  if (tid == 0) {
    int run = 0;
    for (int r = 0; r < nsup; ++r) {
      kscan[r] = run;
      run += (r < n) ? ws.k_super[s0 + r] : ws.k_super[tb.n_sent + d0 + (r - n)];
    }
    kscan[nsup] = run;
  }
  __syncthreads();
  for (int r = tid; r < nsup; r += BLD_THREADS) {
    out.super_nid[sr0 + r] = nid0 + nw + r;
    out.super_type[sr0 + r] = r < n ? 1 : 2;
    out.super_graph[sr0 + r] = g;
    out.super_indptr[sr0 + r] = p0 + kscan[r];
    int ex;
    if (!tb.hdsg) {
      ex = 2 * n;                                   // dataloader.py:262-263
    } else if (r < n) {
      ex = 0;
    } else {
      ex = 0;                                       // one s->d edge per sentence of this document (:383-385)
      for (int i = 0; i < n; ++i) ex += (__ldg(tb.sent_doc + s0 + i) == r - n) ? 1 : 0;
    }
    out.super_extra[sr0 + r] = ex;
  }
  if (tid == 0) out.super_indptr[sr0 + nsup] = p0 + kscan[nsup];
  // word in-degrees
  for (int t = tid; t < n * L; t += BLD_THREADS) {
    const int info = ws.pos_info[(size_t)s0 * L + t];
    if (info >= 0) atomicAdd(&wscan[info], 1);
  }
  for (int jd = 0; jd < nd; ++jd) {
    const int t0 = __ldg(tb.doc_tok_ptr + d0 + jd), t1 = __ldg(tb.doc_tok_ptr + d0 + jd + 1);
    for (int t = t0 + tid; t < t1; t += BLD_THREADS) {
      const int info = ws.doc_info[t];
      if (info >= 0) atomicAdd(&wscan[info], 1);
    }
  }
  __syncthreads();
  {
    int running = 0;
    for (int c = 0; c < nw; c += BLD_THREADS) {
      const int u = c + tid;
      const int v = u < nw ? wscan[u] : 0;
      int tot;
      const int ex = block_excl_scan(v, &tot, warp_buf);
      if (u < nw) {
        wscan[u] = running + ex;
        out.word_indptr[w0 + u] = p0 + running + ex;
      }
      running += tot;
    }
    if (tid == 0) out.word_indptr[w0 + nw] = p0 + running;
  }
  __syncthreads();
  // walk the supernodes in DGL insertion order so that every word's in-edge list is ascending in edge id
  for (int r = 0; r < nsup; ++r) {
    const bool is_sent = r < n;
    int t0, t1;
    const int32_t* info_arr;
    const int8_t* bin_arr;
    int base;
    if (is_sent) {
      t0 = 0; t1 = L;
      info_arr = ws.pos_info + ((size_t)(s0 + r)) * L;
      bin_arr = tb.sent_bin + ((size_t)(s0 + r)) * L;
      base = tb.hdsg ? (2 * kscan[r] + r) : (2 * kscan[r] + 2 * n * r);
    } else {
      const int jd = r - n;
      t0 = __ldg(tb.doc_tok_ptr + d0 + jd);
      t1 = __ldg(tb.doc_tok_ptr + d0 + jd + 1);
      info_arr = ws.doc_info;
      bin_arr = tb.doc_bin;
      base = 2 * kscan[n] + n + 2 * (kscan[r] - kscan[n]);
    }
    int running = 0;
    for (int c = t0; c < t1; c += BLD_THREADS) {
      const int t = c + tid;
      int info = -1;
      if (t < t1) info = info_arr[t];
      int tot;
      const int ex = block_excl_scan(info >= 0 ? 1 : 0, &tot, warp_buf);
      if (info >= 0) {
        const int rank = running + ex;
        const int b = bin_arr[t];
        const int p = p0 + kscan[r] + rank;
        out.super_src[p] = w0 + info;
        out.super_bin[p] = (uint8_t)b;
        out.super_eid[p] = e0 + base + 2 * rank;
        const int qp = p0 + wscan[info] + cursor[info];
        cursor[info] += 1;                           // a word occurs at most once per supernode
        out.word_src[qp] = sr0 + r;
        out.word_bin[qp] = (uint8_t)b;
        out.word_eid[qp] = e0 + base + 2 * rank + 1;
      }
      running += tot;
    }
    __syncthreads();
  }
}

static int pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace hsg

using namespace hsg;

static int check_tb(const hsg_token_batch* tb) {
  if (!tb || tb->n_graphs < 0 || tb->n_sent < 0 || tb->sent_len <= 0) return HSG_ERR_ARG;
  if (tb->n_graphs > 0 && (!tb->tokens || !tb->sent_bin || !tb->graph_sent_ptr || !tb->filter_bitmap)) return HSG_ERR_ARG;
  if (tb->hdsg && tb->n_graphs > 0 && (!tb->graph_doc_ptr || !tb->sent_doc || !tb->doc_tok_ptr)) return HSG_ERR_ARG;
  if (tb->hdsg && tb->n_doc_tok > 0 && (!tb->doc_tokens || !tb->doc_bin)) return HSG_ERR_ARG;
  if (tb->max_sent_per_graph <= 0 && tb->n_sent > 0) return HSG_ERR_ARG;
  return HSG_OK;
}

extern "C" {

size_t hsg_build_workspace_bytes(const hsg_token_batch* tb) {
  if (!tb) return 0;
  return ws_layout(tb, nullptr, nullptr);
}

int hsg_build_count(const hsg_token_batch* tb, hsg_graph_offsets off, int32_t* status, void* ws, size_t ws_bytes,
                    void* stream) {
  int rc = check_tb(tb);
  if (rc) return rc;
  if (!off.word_ptr || !off.super_ptr || !off.node_ptr || !off.edge_ptr || !off.pair_ptr || !status || !ws)
    return HSG_ERR_ARG;
  if (ws_bytes < hsg_build_workspace_bytes(tb)) return HSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  BuildWs w;
  ws_layout(tb, ws, &w);
  const int cap_tok = tb->max_sent_per_graph * tb->sent_len;
  const int ht = pow2_ge((cap_tok > 0 ? cap_tok : 1) * 8 / 5 + 8);   // load factor <= 0.625
  const size_t smem = ((size_t)2 * ht + (tb->hdsg ? cap_tok : 0)) * sizeof(int32_t);
  if (smem > 220 * 1024) return HSG_ERR_CAPACITY;
  if (tb->n_graphs > 0) {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      if (cudaFuncSetAttribute(build_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return HSG_ERR_CUDA;
      configured = smem;
    }
    LaunchScope ls(SLOT_BUILD_COUNT, s);
    launch_k(build_count_kernel, dim3(tb->n_graphs), dim3(BLD_THREADS), smem, s, *tb, w, ht, cap_tok, status);
    rc = check_launch();
    if (rc) return rc;
  }
  LaunchScope ls(SLOT_BUILD_SCAN, s);
  launch_k(build_scan_kernel, dim3(1), dim3(BLD_THREADS), 0, s, tb->n_graphs, w, off);
  return check_launch();
}

int hsg_build_fill(const hsg_token_batch* tb, const hsg_graph_out* out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_tb(tb);
  if (rc) return rc;
  if (!out || !ws || !out->status) return HSG_ERR_ARG;
  if (ws_bytes < hsg_build_workspace_bytes(tb)) return HSG_ERR_WORKSPACE;
  if (tb->n_graphs == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  BuildWs w;
  ws_layout(tb, ws, &w);
  const int cap_tok = tb->max_sent_per_graph * tb->sent_len;
  const int cap_sup = tb->max_sent_per_graph + (tb->hdsg ? tb->max_sent_per_graph : 0);
  const size_t smem = ((size_t)2 * cap_tok + 1 + cap_sup + 1) * sizeof(int32_t);
  if (smem > 220 * 1024) return HSG_ERR_CAPACITY;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(build_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return HSG_ERR_CUDA;
    configured = smem;
  }
  // The builder runs next to the step's tensor-core products (one ~180 KB CTA per SM).  An SM keeps the shared-memory
  // carve-out of the kernel it is running: with the maximum carve-out requested here a GEMM CTA can join a builder CTA
  // on the same SM instead of waiting for it to finish.
  static bool carve = false;
  if (!carve) {
    cudaFuncSetAttribute(build_fill_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(build_count_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(build_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carve = true;
  }
  LaunchScope ls(SLOT_BUILD_FILL, s);
  launch_k(build_fill_kernel, dim3(tb->n_graphs), dim3(BLD_THREADS), smem, s, *tb, w, *out, cap_tok, cap_sup);
  return check_launch();
}

}  // extern "C"
