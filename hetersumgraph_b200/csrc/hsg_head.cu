// Readout, loss and extraction on the device - the step right after the update loop:
//   logits = wh(sentence state)                         HiGraph.py:108   (HDSG: wh(cat(sentence, its document)), :216-228)
//   loss   = mean_graphs sum_sentences CE(logits, y)    train.py:114-119 (CrossEntropyLoss(reduction='none'),
//                                                                          dgl.sum_nodes, .mean())
//   top-m  = per graph torch.topk(logits[:, 1], m)      Tester.py:128
// plus the fused Adam update over the flat parameter arena (train.py:90,135; optional clip_grad_norm_, :132-133).
// Everything is reduced in a fixed order (bitwise reproducible); the reference does the same work with ~20 tiny
// launches of stock kernels and a Python loop over dgl.unbatch(G).
#include "hsg_common.cuh"

namespace hsg {

constexpr int HEAD_ROWS_PER_BLOCK = 8;

// warp per sentence: logits, per-sentence CE, d loss / d logits (already scaled by 1/n_graphs)
__global__ void __launch_bounds__(256)
head_fwd_kernel(hsg_head_args a, float* __restrict__ logits, float* __restrict__ dlogits,
                float* __restrict__ row_loss) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= a.n_sent) return;
  const int srow = a.sent_row ? a.sent_row[i] : i;
  const int width = a.hidden * (a.two_part ? 2 : 1);
  float l0 = 0.f, l1 = 0.f;
  for (int j = lane; j < a.hidden; j += 32) {
    const float s = a.state[(size_t)srow * a.hidden + j];
    l0 = fmaf(s, a.wh_w[j], l0);
    l1 = fmaf(s, a.wh_w[width + j], l1);
  }
  if (a.two_part) {
    const int drow = a.doc_row[i];
    for (int j = lane; j < a.hidden; j += 32) {
      const float s = a.state[(size_t)drow * a.hidden + j];
      l0 = fmaf(s, a.wh_w[a.hidden + j], l0);
      l1 = fmaf(s, a.wh_w[width + a.hidden + j], l1);
    }
  }
  l0 = warp_sum(l0) + a.wh_b[0];
  l1 = warp_sum(l1) + a.wh_b[1];
  if (lane == 0) {
    const int y = (int)a.labels[i];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float lse = mx + logf(e0 + e1);
    const float inv = 1.f / (e0 + e1);
    logits[2 * i] = l0;
    logits[2 * i + 1] = l1;
    row_loss[i] = lse - (y ? l1 : l0);
    dlogits[2 * i] = a.inv_graphs * (e0 * inv - (y == 0 ? 1.f : 0.f));
    dlogits[2 * i + 1] = a.inv_graphs * (e1 * inv - (y == 1 ? 1.f : 0.f));
  }
}

// loss = scale * sum(row_loss) in a fixed order (one block)
__global__ void __launch_bounds__(1024) head_loss_reduce_kernel(int n, const float* __restrict__ row_loss, float scale,
                                                                float* __restrict__ loss) {
  pdl_prologue();
  __shared__ float red[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += row_loss[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * scale;
}

// block = `width` threads (<= 256), HEAD_ROWS_PER_BLOCK sentences: thread j owns column j of the wh input.
// Writes the sentence rows of d_state and the per-block partials of d wh.
__global__ void __launch_bounds__(256)
head_bwd_kernel(hsg_head_args a, const float* __restrict__ dlogits, const float* __restrict__ gout,
                float* __restrict__ d_state, float* __restrict__ part /* [blocks][2*width + 2] */) {
  pdl_prologue();
  const int width = a.hidden * (a.two_part ? 2 : 1);
  const int j = threadIdx.x;
  const float g = gout ? gout[0] : 1.f;
  const int i0 = blockIdx.x * HEAD_ROWS_PER_BLOCK, i1 = min(a.n_sent, i0 + HEAD_ROWS_PER_BLOCK);
  float w0 = 0.f, w1 = 0.f, b0 = 0.f, b1 = 0.f;
  const float wc0 = j < width ? a.wh_w[j] : 0.f, wc1 = j < width ? a.wh_w[width + j] : 0.f;
  for (int i = i0; i < i1; ++i) {
    const float d0 = dlogits[2 * i] * g, d1 = dlogits[2 * i + 1] * g;
    if (j < width) {
      const int srow = a.sent_row ? a.sent_row[i] : i;
      const int row = j < a.hidden ? srow : a.doc_row[i];
      const int jj = j < a.hidden ? j : j - a.hidden;
      const float f = a.state[(size_t)row * a.hidden + jj];
      w0 = fmaf(d0, f, w0);
      w1 = fmaf(d1, f, w1);
      if (j < a.hidden) d_state[(size_t)srow * a.hidden + j] = d0 * wc0 + d1 * wc1;   // sentence rows are unique
    }
    b0 += d0;
    b1 += d1;
  }
  float* p = part + (size_t)blockIdx.x * (2 * width + 2);
  if (j < width) {
    p[j] = w0;
    p[width + j] = w1;
  }
  if (j == 0) {
    p[2 * width] = b0;
    p[2 * width + 1] = b1;
  }
}

// HDSG: document rows collect the second half of wh's input gradient from their sentences.  One block per graph,
// thread j owns column j and walks the graph's sentences in order (deterministic; rows were zero-filled).
__global__ void __launch_bounds__(256)
head_bwd_doc_kernel(hsg_head_args a, const float* __restrict__ dlogits, const float* __restrict__ gout,
                    float* __restrict__ d_state) {
  pdl_prologue();
  const int gph = blockIdx.x, j = threadIdx.x;
  if (j >= a.hidden) return;
  const int width = 2 * a.hidden;
  const float g = gout ? gout[0] : 1.f;
  const float wc0 = a.wh_w[a.hidden + j], wc1 = a.wh_w[width + a.hidden + j];
  for (int i = a.graph_sent_ptr[gph]; i < a.graph_sent_ptr[gph + 1]; ++i) {
    float* o = d_state + (size_t)a.doc_row[i] * a.hidden + j;
    *o += (dlogits[2 * i] * wc0 + dlogits[2 * i + 1] * wc1) * g;
  }
}

// one warp per output: lanes take the block partials strided, then a fixed-order shuffle tree
__global__ void __launch_bounds__(256)
head_bwd_reduce_kernel(int nblocks, int n_out, const float* __restrict__ part, float* __restrict__ d_w,
                       float* __restrict__ d_b, int accumulate) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_out) return;
  float s = 0.f;
  for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * n_out + i];
  s = warp_sum(s);
  if (lane == 0) {
    float* o = i < n_out - 2 ? d_w + i : d_b + (i - (n_out - 2));
    *o = accumulate ? *o + s : s;
  }
}

// ---- forward + backward in ONE launch (training step: d L / d loss = 1) ------------------------------------------
// Same block shape as the two kernels above (8 sentences per block: warp w does head_fwd_kernel's work for sentence
// i0 + w, then thread j does head_bwd_kernel's column j over the block's sentences from the shared-memory copy of
// dlogits); head_reduce_both_kernel then runs head_loss_reduce_kernel's and head_bwd_reduce_kernel's reductions with
// their exact summation trees.  Every output is bit-identical to hsg_head_fwd + hsg_head_bwd(gout = NULL); four launches
// of the serial chain between the update loop's forward and backward become two.  (Both reductions inside the first
// kernel, by the block that takes the last ticket, was measured: 130 outputs x 127 partials on one block's 8 warps is a
// chain of dependent L2 reads, +20 us per step.)
__global__ void __launch_bounds__(256)
head_fwd_bwd_kernel(hsg_head_args a, float* __restrict__ logits, float* __restrict__ dlogits,
                    float* __restrict__ row_loss, float* __restrict__ d_state, float* __restrict__ part) {
  pdl_prologue();
  __shared__ float s_d[HEAD_ROWS_PER_BLOCK][2];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int width = a.hidden * (a.two_part ? 2 : 1);
  const int i0 = blockIdx.x * HEAD_ROWS_PER_BLOCK, i1 = min(a.n_sent, i0 + HEAD_ROWS_PER_BLOCK);
  {
    const int i = i0 + wrp;
    if (i < a.n_sent) {
      const int srow = a.sent_row ? a.sent_row[i] : i;
      float l0 = 0.f, l1 = 0.f;
      for (int j = lane; j < a.hidden; j += 32) {
        const float sv = a.state[(size_t)srow * a.hidden + j];
        l0 = fmaf(sv, a.wh_w[j], l0);
        l1 = fmaf(sv, a.wh_w[width + j], l1);
      }
      if (a.two_part) {
        const int drow = a.doc_row[i];
        for (int j = lane; j < a.hidden; j += 32) {
          const float sv = a.state[(size_t)drow * a.hidden + j];
          l0 = fmaf(sv, a.wh_w[a.hidden + j], l0);
          l1 = fmaf(sv, a.wh_w[width + a.hidden + j], l1);
        }
      }
      l0 = warp_sum(l0) + a.wh_b[0];
      l1 = warp_sum(l1) + a.wh_b[1];
      if (lane == 0) {
        const int y = (int)a.labels[i];
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        const float lse = mx + logf(e0 + e1);
        const float inv = 1.f / (e0 + e1);
        logits[2 * i] = l0;
        logits[2 * i + 1] = l1;
        row_loss[i] = lse - (y ? l1 : l0);
        const float d0 = a.inv_graphs * (e0 * inv - (y == 0 ? 1.f : 0.f));
        const float d1 = a.inv_graphs * (e1 * inv - (y == 1 ? 1.f : 0.f));
        dlogits[2 * i] = d0;
        dlogits[2 * i + 1] = d1;
        s_d[wrp][0] = d0;
        s_d[wrp][1] = d1;
      }
    }
  }
  __syncthreads();
  {
    const int j = threadIdx.x;
    float w0 = 0.f, w1 = 0.f, b0 = 0.f, b1 = 0.f;
    const float wc0 = j < width ? a.wh_w[j] : 0.f, wc1 = j < width ? a.wh_w[width + j] : 0.f;
    for (int i = i0; i < i1; ++i) {
      const float d0 = s_d[i - i0][0], d1 = s_d[i - i0][1];
      if (j < width) {
        const int srow = a.sent_row ? a.sent_row[i] : i;
        const int row = j < a.hidden ? srow : a.doc_row[i];
        const int jj = j < a.hidden ? j : j - a.hidden;
        const float f = a.state[(size_t)row * a.hidden + jj];
        w0 = fmaf(d0, f, w0);
        w1 = fmaf(d1, f, w1);
        if (j < a.hidden) d_state[(size_t)srow * a.hidden + j] = d0 * wc0 + d1 * wc1;
      }
      b0 += d0;
      b1 += d1;
    }
    float* p = part + (size_t)blockIdx.x * (2 * width + 2);
    if (j < width) {
      p[j] = w0;
      p[width + j] = w1;
    }
    if (j == 0) {
      p[2 * width] = b0;
      p[2 * width + 1] = b1;
    }
  }
}

// both fixed-order reductions of the head in one launch: block 0 = head_loss_reduce_kernel's tree over its 1024 virtual
// threads (4 per thread), blocks 1.. = head_bwd_reduce_kernel (one warp per output)
__global__ void __launch_bounds__(256)
head_reduce_both_kernel(int n, const float* __restrict__ row_loss, float scale, float* __restrict__ loss, int nblocks,
                        int n_out, const float* __restrict__ part, float* __restrict__ d_w, float* __restrict__ d_b,
                        int accumulate) {
  pdl_prologue();
  if (blockIdx.x == 0) {
    __shared__ float red[1024];
    for (int v = threadIdx.x; v < 1024; v += 256) {
      float sm = 0.f;
      for (int i = v; i < n; i += 1024) sm += row_loss[i];
      red[v] = sm;
    }
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      for (int v = threadIdx.x; v < o; v += 256) red[v] += red[v + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = red[0] * scale;
    return;
  }
  const int lane = threadIdx.x & 31;
  const int i = ((blockIdx.x - 1) * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_out) return;
  float sm = 0.f;
  for (int b = lane; b < nblocks; b += 32) sm += part[(size_t)b * n_out + i];
  sm = warp_sum(sm);
  if (lane == 0) {
    float* o = i < n_out - 2 ? d_w + i : d_b + (i - (n_out - 2));
    *o = accumulate ? *o + sm : sm;
  }
}

// one warp per graph: out[g, rank] = local sentence index with the rank-th largest class-1 logit (ties: lower index
// first), -1 padded.  n is at most a few hundred, so rank counting is cheaper than a sort.
__global__ void __launch_bounds__(256)
topm_kernel(const float* __restrict__ logits, const int* __restrict__ graph_sent_ptr, int n_graphs, int m,
            int* __restrict__ out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int gph = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gph >= n_graphs) return;
  const int s0 = graph_sent_ptr[gph], n = graph_sent_ptr[gph + 1] - s0;
  for (int r = lane; r < m; r += 32) out[(size_t)gph * m + r] = -1;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const float pi = logits[2 * (s0 + i) + 1];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const float pk = logits[2 * (s0 + k) + 1];
      rank += (pk > pi || (pk == pi && k < i)) ? 1 : 0;
    }
    if (rank < m) out[(size_t)gph * m + rank] = i;
  }
}

// ---- optimizer -----------------------------------------------------------------------------------------------
constexpr int SUMSQ_BLOCKS = 296;

__global__ void __launch_bounds__(256) sumsq_part_kernel(size_t n, const float* __restrict__ g, float* __restrict__ part) {
  pdl_prologue();
  __shared__ float red[256];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s = fmaf(g[i], g[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(512) sumsq_final_kernel(int nparts, const float* __restrict__ part, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[512];
  red[threadIdx.x] = threadIdx.x < nparts ? part[threadIdx.x] : 0.f;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// torch.optim.Adam (no amsgrad, no weight decay) on flat arrays; optional clip_grad_norm_ coefficient from sumsq.
__global__ void __launch_bounds__(256)
adam_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            float beta1, float beta2, float step_size, float inv_sqrt_bc2, float eps, const float* __restrict__ sumsq,
            float max_norm) {
  pdl_prologue();
  float coef = 1.f;
  if (sumsq) coef = fminf(1.f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// Same update with the step number read from DEVICE memory (state[0] = completed steps, state[1] = block ticket):
// the kernel arguments are identical every step, so the launch can live in a replayed CUDA graph.  The last block to
// finish advances the counter (every block has read it by then).  zero_grad: the gradient is cleared once consumed, so
// the next step's kernels accumulate into a zeroed arena without a separate fill launch.
__global__ void __launch_bounds__(256)
adam_dev_kernel(size_t n, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float lr, float beta1, float beta2, float eps, const float* __restrict__ sumsq, float max_norm,
                unsigned long long* __restrict__ state, int zero_grad) {
  pdl_prologue();
  const unsigned long long t = state[0] + 1ull;
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  float coef = 1.f;
  if (sumsq) coef = fminf(1.f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    if (zero_grad) g[i] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(&state[1], 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1ull) {
      state[1] = 0ull;
      state[0] = t;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient all-reduce + Adam in ONE kernel over NVLink peer memory (data-parallel ranks of one box; replaces
// ncclAllReduce + adam_dev_kernel, train.py:131-135 under DistributedDataParallel-style training).  The 433 k-float
// arena is latency-, not bandwidth-bound: NCCL needs ~75 us for it inside the step graph, this kernel ~15-25 us.
//   every rank owns a symmetric buffer  [ flags: world x u64 | recv: 2 parities x world slots x n floats ];
//   phase 1  PUSH: the rank's gradient arena is stored into slot `rank` of EVERY rank's recv area (parity = step & 1;
//            plain 128-bit stores through the peer mappings) and the arena is zeroed in the same sweep - nobody else
//            ever reads it, so no exit barrier is needed before the next step accumulates into it;
//   phase 2  the last block to finish publishes flag[rank] = step on every rank (fence.sys + st.release.sys);
//   phase 3  every block waits until all `world` flags of its own rank show this step (ld.acquire.sys);
//   phase 4  REDUCE + UPDATE from LOCAL memory: g = sum over slots in rank order (a fixed order: every rank computes
//            bit-identical sums, so the replicas stay bit-identical), then the Adam update of adam_dev_kernel.
// Step s + 2 reuses the parity of step s; a rank can only push for s + 2 after the barrier of s + 1, which every peer
// passes after finishing its phase 4 of step s (program order) - the slots are free by then.
// All blocks of the grid must be co-resident (phase 3 spins): the launcher sizes the grid to one block per SM.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int PEER_FLAG_FLOATS = 64;                       // 256 bytes of flags in front of the recv area

__global__ void __launch_bounds__(256)
allreduce_adam_kernel(size_t n, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                      float* __restrict__ v, float lr, float beta1, float beta2, float eps,
                      unsigned long long* __restrict__ state, float* const* __restrict__ peer, int rank, int world) {
  pdl_prologue();
  __shared__ int last_flag;
  const unsigned long long t = state[0] + 1ull;           // this step (the same number on every rank)
  const size_t par_off = (size_t)PEER_FLAG_FLOATS + (size_t)(t & 1ull) * world * n;
  const size_t tid = blockIdx.x * (size_t)256 + threadIdx.x, nthr = (size_t)gridDim.x * 256;
  const size_t n4 = n / 4;
  // ---- phase 1: push + zero ----
  {
    float4* g4 = reinterpret_cast<float4*>(g);
    for (size_t i = tid; i < n4; i += nthr) {
      const float4 x = g4[i];
      for (int r = 0; r < world; ++r)
        reinterpret_cast<float4*>(peer[r] + par_off + (size_t)rank * n)[i] = x;
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (size_t i = 4 * n4 + tid; i < n; i += nthr) {
      const float x = g[i];
      for (int r = 0; r < world; ++r) peer[r][par_off + (size_t)rank * n + i] = x;
      g[i] = 0.f;
    }
  }
  // ---- phase 2: the last block publishes this rank's flag on every rank ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long ticket = atomicAdd(&state[1], 1ull);
    last_flag = ticket == (unsigned long long)gridDim.x - 1ull;
  }
  __syncthreads();
  if (last_flag) {
    if (threadIdx.x == 0) state[1] = 0ull;
    __threadfence_system();
    if ((int)threadIdx.x < world)
      st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer[threadIdx.x]) + rank, t);
  }
  // ---- phase 3: wait for every rank's push ----
  if ((int)threadIdx.x < world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(peer[rank]) + threadIdx.x;
    unsigned long long spins = 0;
    while (ld_acquire_sys_u64(f) < t)
      if (++spins > (1ull << 28)) __trap();               // a rank that never arrives: fail loudly, do not hang
  }
  __syncthreads();
  // ---- phase 4: reduce the local slots in rank order + Adam ----
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float* recv = peer[rank] + par_off;
  auto upd = [&](float gi, float& pi, float& mi_, float& vi_) {
    const float mi = beta1 * mi_ + (1.f - beta1) * gi;
    const float vi = beta2 * vi_ + (1.f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  };
  {
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (size_t i = tid; i < n4; i += nthr) {
      float4 s = __ldcg(reinterpret_cast<const float4*>(recv) + i);
      for (int r = 1; r < world; ++r) {
        const float4 x = __ldcg(reinterpret_cast<const float4*>(recv + (size_t)r * n) + i);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      upd(s.x, pp.x, mm.x, vv.x);
      upd(s.y, pp.y, mm.y, vv.y);
      upd(s.z, pp.z, mm.z, vv.z);
      upd(s.w, pp.w, mm.w, vv.w);
      p4[i] = pp;
      m4[i] = mm;
      v4[i] = vv;
    }
    for (size_t i = 4 * n4 + tid; i < n; i += nthr) {
      float s = __ldcg(recv + i);
      for (int r = 1; r < world; ++r) s += __ldcg(recv + (size_t)r * n + i);
      upd(s, p[i], m[i], v[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(&state[2], 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1ull) {
      state[2] = 0ull;
      state[0] = t;
      __threadfence();
    }
  }
}

// out[i, :] = table[ids[i], :]   (the word-embedding lookup of set_wnfeature, HiGraph.py:147-148), float4 rows
__global__ void __launch_bounds__(256)
embed_gather_kernel(int n, int dim4, const int32_t* __restrict__ ids, const float4* __restrict__ table,
                    float4* __restrict__ out) {
  pdl_prologue();
  const int rows_per_block = 256 / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * rows_per_block + warp; r < n; r += gridDim.x * rows_per_block) {
    const float4* src = table + (size_t)__ldg(ids + r) * dim4;
    float4* dst = out + (size_t)r * dim4;
    for (int c = lane; c < dim4; c += 32) dst[c] = __ldg(src + c);
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_embed_gather(int n, int dim, const int32_t* ids, const float* table, float* out, void* stream) {
  if (n < 0 || dim <= 0 || (dim & 3) || (n > 0 && (!ids || !table || !out))) return HSG_ERR_ARG;
  if (!aligned16(table) || !aligned16(out)) return HSG_ERR_ALIGN;
  if (n == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = ceil_div(n, 8);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  LaunchScope ls(SLOT_HEAD, s);
  launch_k(embed_gather_kernel, dim3(blocks), dim3(256), 0, s, n, dim / 4, ids, reinterpret_cast<const float4*>(table),
           reinterpret_cast<float4*>(out));
  return check_launch();
}

size_t hsg_head_workspace_bytes(int n_sent, int width) {
  const int blocks = ceil_div(n_sent > 0 ? n_sent : 1, HEAD_ROWS_PER_BLOCK);
  return ((size_t)(n_sent > 0 ? n_sent : 1) + (size_t)blocks * (2 * (size_t)width + 2)) * sizeof(float) + 16;
}

static bool head_args_ok(const hsg_head_args* a) {
  if (!a || a->n_sent < 0 || a->n_super < 0 || a->hidden <= 0 || !a->state || !a->wh_w || !a->wh_b || !a->labels)
    return false;
  if (a->two_part && (!a->doc_row || !a->graph_sent_ptr || a->n_graphs < 0)) return false;
  return a->hidden * (a->two_part ? 2 : 1) <= 256;
}

int hsg_head_fwd(const hsg_head_args* a, float* logits, float* dlogits, float* loss, void* ws, size_t ws_bytes,
                 void* stream) {
  if (!head_args_ok(a) || !logits || !dlogits || !loss || !ws) return HSG_ERR_ARG;
  const int width = a->hidden * (a->two_part ? 2 : 1);
  if (ws_bytes < hsg_head_workspace_bytes(a->n_sent, width)) return HSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* row_loss = reinterpret_cast<float*>(ws);
  LaunchScope ls(SLOT_HEAD, s);
  if (a->n_sent > 0) launch_k(head_fwd_kernel, dim3(ceil_div(a->n_sent, 8)), dim3(256), 0, s, *a, logits, dlogits, row_loss);
  launch_k(head_loss_reduce_kernel, dim3(1), dim3(1024), 0, s, a->n_sent, row_loss, a->inv_graphs, loss);
  return check_launch();
}

int hsg_head_bwd(const hsg_head_args* a, const float* dlogits, const float* gout, float* d_state, float* d_wh_w,
                 float* d_wh_b, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (!head_args_ok(a) || !dlogits || !d_state || !d_wh_w || !d_wh_b || !ws) return HSG_ERR_ARG;
  const int width = a->hidden * (a->two_part ? 2 : 1);
  if (ws_bytes < hsg_head_workspace_bytes(a->n_sent, width)) return HSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws) + (a->n_sent > 0 ? a->n_sent : 1);
  const int blocks = ceil_div(a->n_sent, HEAD_ROWS_PER_BLOCK);
  LaunchScope ls(SLOT_HEAD, s);
  // rows that are not sentences (documents) and sentences without ... every row starts at zero
  if (a->sent_row || a->two_part) {
    if (cudaMemsetAsync(d_state, 0, (size_t)a->n_super * a->hidden * sizeof(float), s) != cudaSuccess) return HSG_ERR_CUDA;
  }
  if (blocks > 0) {
    launch_k(head_bwd_kernel, dim3(blocks), dim3(256), 0, s, *a, dlogits, gout, d_state, part);
    if (a->two_part && a->n_graphs > 0) launch_k(head_bwd_doc_kernel, dim3(a->n_graphs), dim3(256), 0, s, *a, dlogits, gout, d_state);
  }
  const int n_out = 2 * width + 2;
  launch_k(head_bwd_reduce_kernel, dim3(ceil_div(n_out, 8)), dim3(256), 0, s, blocks, n_out, part, d_wh_w, d_wh_b, accumulate);
  return check_launch();
}

int hsg_head_fwd_bwd(const hsg_head_args* a, float* logits, float* dlogits, float* loss, float* d_state, float* d_wh_w,
                     float* d_wh_b, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (!head_args_ok(a) || !logits || !dlogits || !loss || !d_state || !d_wh_w || !d_wh_b || !ws) return HSG_ERR_ARG;
  if (a->n_sent <= 0) {                                  // nothing to fuse: the two calls handle the empty batch
    const int rc = hsg_head_fwd(a, logits, dlogits, loss, ws, ws_bytes, stream);
    if (rc != HSG_OK) return rc;
    return hsg_head_bwd(a, dlogits, nullptr, d_state, d_wh_w, d_wh_b, accumulate, ws, ws_bytes, stream);
  }
  const int width = a->hidden * (a->two_part ? 2 : 1);
  if (ws_bytes < hsg_head_workspace_bytes(a->n_sent, width)) return HSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* row_loss = reinterpret_cast<float*>(ws);
  float* part = row_loss + a->n_sent;
  const int blocks = ceil_div(a->n_sent, HEAD_ROWS_PER_BLOCK);
  LaunchScope ls(SLOT_HEAD, s);
  if (a->sent_row || a->two_part) {
    if (cudaMemsetAsync(d_state, 0, (size_t)a->n_super * a->hidden * sizeof(float), s) != cudaSuccess) return HSG_ERR_CUDA;
  }
  launch_k(head_fwd_bwd_kernel, dim3(blocks), dim3(256), 0, s, *a, logits, dlogits, row_loss, d_state, part);
  if (a->two_part && a->n_graphs > 0)
    launch_k(head_bwd_doc_kernel, dim3(a->n_graphs), dim3(256), 0, s, *a, (const float*)dlogits, (const float*)nullptr, d_state);
  const int n_out = 2 * width + 2;
  launch_k(head_reduce_both_kernel, dim3(1 + ceil_div(n_out, 8)), dim3(256), 0, s, a->n_sent, (const float*)row_loss,
           a->inv_graphs, loss, blocks, n_out, (const float*)part, d_wh_w, d_wh_b, accumulate);
  return check_launch();
}

int hsg_topm(const float* logits, const int32_t* graph_sent_ptr, int n_graphs, int m, int32_t* out_idx, void* stream) {
  if (!logits || !graph_sent_ptr || n_graphs < 0 || m <= 0 || !out_idx) return HSG_ERR_ARG;
  if (n_graphs == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_HEAD, s);
  launch_k(topm_kernel, dim3(ceil_div(n_graphs, 8)), dim3(256), 0, s, logits, graph_sent_ptr, n_graphs, m, out_idx);
  return check_launch();
}

size_t hsg_adam_workspace_bytes(void) { return (size_t)(SUMSQ_BLOCKS + 4) * sizeof(float); }

int hsg_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, float max_grad_norm, void* ws, size_t ws_bytes, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return HSG_ERR_ARG;
  if (n == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const float* sumsq = nullptr;
  LaunchScope ls(SLOT_ADAM, s);
  if (max_grad_norm > 0.f) {
    if (!ws || ws_bytes < hsg_adam_workspace_bytes()) return HSG_ERR_WORKSPACE;
    float* part = reinterpret_cast<float*>(ws);
    launch_k(sumsq_part_kernel, dim3(SUMSQ_BLOCKS), dim3(256), 0, s, n, grad, part);
    launch_k(sumsq_final_kernel, dim3(1), dim3(512), 0, s, SUMSQ_BLOCKS, part, part + SUMSQ_BLOCKS);
    sumsq = part + SUMSQ_BLOCKS;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  size_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  launch_k(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, s, n, param, grad, exp_avg, exp_avg_sq, beta1, beta2, step_size,
                                              inv_sqrt_bc2, eps, sumsq, max_grad_norm);
  return check_launch();
}

int hsg_adam_step_dev(size_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                      float beta2, float eps, unsigned long long* step_state, int zero_grad, float max_grad_norm,
                      void* ws, size_t ws_bytes, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_state) return HSG_ERR_ARG;
  if (n == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const float* sumsq = nullptr;
  LaunchScope ls(SLOT_ADAM, s);
  if (max_grad_norm > 0.f) {
    if (!ws || ws_bytes < hsg_adam_workspace_bytes()) return HSG_ERR_WORKSPACE;
    float* part = reinterpret_cast<float*>(ws);
    launch_k(sumsq_part_kernel, dim3(SUMSQ_BLOCKS), dim3(256), 0, s, n, (const float*)grad, part);
    launch_k(sumsq_final_kernel, dim3(1), dim3(512), 0, s, SUMSQ_BLOCKS, (const float*)part, part + SUMSQ_BLOCKS);
    sumsq = part + SUMSQ_BLOCKS;
  }
  size_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  launch_k(adam_dev_kernel, dim3((unsigned)blocks), dim3(256), 0, s, n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2,
           eps, sumsq, max_grad_norm, step_state, zero_grad);
  return check_launch();
}

size_t hsg_allreduce_adam_buffer_floats(size_t n, int world) {
  return (size_t)PEER_FLAG_FLOATS + 2 * (size_t)(world > 0 ? world : 1) * n;
}

int hsg_allreduce_adam_step(size_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                            float beta1, float beta2, float eps, unsigned long long* step_state,
                            float* const* peer_bufs, int rank, int world, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_state || !peer_bufs) return HSG_ERR_ARG;
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return HSG_ERR_ARG;
  if (n % 4 != 0) return HSG_ERR_SHAPE;                    // the receive slots are n floats apart: 128-bit accesses
  if (!aligned16(param) || !aligned16(grad) || !aligned16(exp_avg) || !aligned16(exp_avg_sq)) return HSG_ERR_ALIGN;
  if (n == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_ADAM, s);
  size_t blocks = (n / 4 + 255) / 256;
  const size_t cap = (size_t)num_sms();                    // phase 3 spins: every block must be resident
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_k(allreduce_adam_kernel, dim3((unsigned)blocks), dim3(256), 0, s, n, param, grad, exp_avg, exp_avg_sq, lr, beta1,
           beta2, eps, step_state, peer_bufs, rank, world);
  return check_launch();
}

}  // extern "C"

// ---- HDSG document-node init ------------------------------------------------------------
// HSumDocGraph.forward / set_dnfeature (HiGraph.py:196-203, 231-244): a document node starts from the MEAN of its
// sentences' init features, projected by dn_feature_proj; the supernode init tensor interleaves sentence and
// document rows per graph.  The reference does this with a Python loop over graph.predecessors(dnode).
namespace hsg {

// one block per document, thread per column: doc_mean[j] = mean_{i in graph(j): sent_doc[i] == j} sent_feature[i]
__global__ void __launch_bounds__(256)
doc_mean_kernel(hsg_doc_map m, const float* __restrict__ sent_feature, float* __restrict__ doc_mean) {
  pdl_prologue();
  const int j = blockIdx.x, g = m.doc_graph[j];
  const int i0 = m.graph_sent_ptr[g], i1 = m.graph_sent_ptr[g + 1];
  for (int c = threadIdx.x; c < m.hidden; c += blockDim.x) {
    float s = 0.f;
    int cnt = 0;
    for (int i = i0; i < i1; ++i)                                   // sentence order: fixed summation order
      if (m.sent_doc[i] == j) {
        s += sent_feature[(size_t)i * m.hidden + c];
        ++cnt;
      }
    doc_mean[(size_t)j * m.hidden + c] = cnt > 0 ? s / (float)cnt : 0.f;
  }
}

// rows 0..n_sent-1: super[sent_row[i]] = sent_feature[i]; rows n_sent..: super[doc_row[j]] = doc_feature[j]
__global__ void __launch_bounds__(256)
super_assemble_kernel(hsg_doc_map m, const float* __restrict__ sent_feature, const float* __restrict__ doc_feature,
                      float* __restrict__ super_feature) {
  pdl_prologue();
  const int r = blockIdx.x;
  const float* src;
  int dst;
  if (r < m.n_sent) {
    src = sent_feature + (size_t)r * m.hidden;
    dst = m.sent_row[r];
  } else {
    src = doc_feature + (size_t)(r - m.n_sent) * m.hidden;
    dst = m.doc_row[r - m.n_sent];
  }
  for (int c = threadIdx.x; c < m.hidden; c += blockDim.x) super_feature[(size_t)dst * m.hidden + c] = src[c];
}

// d_doc_feature[j] = d_super[doc_row[j]]
__global__ void __launch_bounds__(256)
doc_gather_kernel(hsg_doc_map m, const float* __restrict__ d_super, float* __restrict__ d_doc_feature) {
  pdl_prologue();
  const int j = blockIdx.x;
  for (int c = threadIdx.x; c < m.hidden; c += blockDim.x)
    d_doc_feature[(size_t)j * m.hidden + c] = d_super[(size_t)m.doc_row[j] * m.hidden + c];
}

// d_sent[i] = d_super[sent_row[i]] + d_doc_mean[doc(i)] / #sentences(doc(i))
__global__ void __launch_bounds__(256)
doc_init_bwd_kernel(hsg_doc_map m, const float* __restrict__ d_super, const float* __restrict__ d_doc_mean,
                    float* __restrict__ d_sent) {
  pdl_prologue();
  __shared__ int cnt_s;
  const int i = blockIdx.x, j = m.sent_doc[i];
  if (threadIdx.x == 0) {
    const int g = m.doc_graph[j];
    int cnt = 0;
    for (int t = m.graph_sent_ptr[g]; t < m.graph_sent_ptr[g + 1]; ++t) cnt += m.sent_doc[t] == j ? 1 : 0;
    cnt_s = cnt;
  }
  __syncthreads();
  const float inv = 1.f / (float)cnt_s;
  for (int c = threadIdx.x; c < m.hidden; c += blockDim.x)
    d_sent[(size_t)i * m.hidden + c] =
        d_super[(size_t)m.sent_row[i] * m.hidden + c] + d_doc_mean[(size_t)j * m.hidden + c] * inv;
}

static bool doc_map_ok(const hsg_doc_map* m) {
  return m && m->n_sent >= 0 && m->n_doc >= 0 && m->hidden > 0 && m->sent_row && m->doc_row && m->sent_doc &&
         m->doc_graph && m->graph_sent_ptr;
}

}  // namespace hsg

extern "C" {

int hsg_doc_mean(const hsg_doc_map* m, const float* sent_feature, float* doc_mean, void* stream) {
  if (!doc_map_ok(m) || !sent_feature || !doc_mean) return HSG_ERR_ARG;
  if (m->n_doc == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_HEAD, s);
  launch_k(doc_mean_kernel, dim3(m->n_doc), dim3(256), 0, s, *m, sent_feature, doc_mean);
  return check_launch();
}

int hsg_super_assemble(const hsg_doc_map* m, const float* sent_feature, const float* doc_feature,
                       float* super_feature, void* stream) {
  if (!doc_map_ok(m) || !sent_feature || !doc_feature || !super_feature) return HSG_ERR_ARG;
  if (m->n_sent + m->n_doc == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_HEAD, s);
  launch_k(super_assemble_kernel, dim3(m->n_sent + m->n_doc), dim3(256), 0, s, *m, sent_feature, doc_feature,
           super_feature);
  return check_launch();
}

int hsg_doc_init_bwd(const hsg_doc_map* m, const float* d_super, const float* d_doc_mean /* or NULL: gather phase */,
                     float* d_doc_feature /* gather phase */, float* d_sent /* scatter phase */, void* stream) {
  if (!doc_map_ok(m) || !d_super) return HSG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_HEAD, s);
  if (d_doc_mean == nullptr) {
    if (!d_doc_feature) return HSG_ERR_ARG;
    if (m->n_doc > 0) launch_k(doc_gather_kernel, dim3(m->n_doc), dim3(256), 0, s, *m, d_super, d_doc_feature);
  } else {
    if (!d_sent) return HSG_ERR_ARG;
    if (m->n_sent > 0) launch_k(doc_init_bwd_kernel, dim3(m->n_sent), dim3(256), 0, s, *m, d_super, d_doc_mean, d_sent);
  }
  return check_launch();
}

}  // extern "C"
