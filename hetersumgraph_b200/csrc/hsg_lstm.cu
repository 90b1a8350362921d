// Recurrent part of the sentence-level BiLSTM (HiGraph.py:118-119, 135-142: nn.LSTM over the packed per-graph
// sentence sequences), forward and backward-through-time, one layer per call.
//
// The reference pads / packs the per-graph sentence lists and lets cuDNN walk the time steps with one small launch
// sequence per step (measured here: ~9 ms per training step at 32 graphs, 12x the whole WSWGAT path).  Every graph is
// an independent sequence whose rows are already contiguous in the batched sentence order, so:
//   * the input products  x . W_ih^T  of ALL time steps are one GEMM per direction (caller, hsg_gemm_nt);
//   * the recurrence runs in ONE persistent kernel per layer: CTA (graph, direction), 2H threads, each owning TWO gate
//     rows (t and t + 2H).  W_hh stays on chip for the whole sequence, HALF of it in registers: columns k < 64 in
//     shared memory (row pitch KS + 4 floats: the float4 reads of 8 consecutive rows cover all 32 banks), columns
//     64..127 in 2 x 64 registers per thread.  The step time is the shared-memory read of W_hh (ncu: mio / short-
//     scoreboard stalls on the LDS.128 stream), so moving half of it into the register file halves it
//     (first version: 96 of 128 columns in shared memory, one row per thread: 2.0 us per step);
//     h in shared memory, c in a register.  No global traffic on the critical path except the prefetched
//     x-projection values;
//   * backward: the same CTA walks its sequence in reverse, recomputes nothing (post-activation gates, c and
//     h_{t-1} were saved by the forward), keeps dh / dc on chip and writes the pre-activation gradients da [S, 4H]
//     per direction; the weight / input gradients are GEMMs over all time steps at once (caller: gemm_tn, gemm_nn).
// Gate order i, f, g, o (torch.nn.LSTM); accurate expf / tanhf (fp32 parity with the CPU reference <= 1e-5).
#include "hsg_common.cuh"

namespace hsg {

constexpr int LSTM_MAX_H = 128;
constexpr int LSTM_KS_MAX = 64;      // W_hh columns kept in shared memory
constexpr int LSTM_KR = 64;          // W_hh columns kept in registers (k = KS .. KS + 63), per gate row
constexpr int LSTM_NJ_MAX = 16;      // gate-row chunks of the shared-memory part of the transposed product
constexpr int LSTM_MAX_WARPS = 8;    // 2 * LSTM_MAX_H / 32

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// W_hh[:, col0 : col0 + ncols] -> shared memory rows of `pitch` floats: coalesced float4 copies, 8 independent loads in
// flight per thread (a scalar load -> store loop costs ~100 serial L2 latencies per thread, ncu: 17 % of the samples)
__device__ __forceinline__ void load_whh_cols(float* W_s, const float* __restrict__ whh, int H, int col0, int ncols,
                                              int pitch) {
  const int nk4 = ncols >> 2, tot4 = 4 * H * nk4;
  for (int base = threadIdx.x; base < tot4; base += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i4 = base + q * blockDim.x;
      if (i4 < tot4) {
        const int jj = i4 / nk4, k4 = i4 - jj * nk4;
        v[q] = __ldg(reinterpret_cast<const float4*>(whh + (size_t)jj * H + col0) + k4);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i4 = base + q * blockDim.x;
      if (i4 < tot4) {
        const int jj = i4 / nk4, k4 = i4 - jj * nk4;
        *reinterpret_cast<float4*>(W_s + (size_t)jj * pitch + 4 * k4) = v[q];
      }
    }
  }
}

// On-chip copy of W_hh for this CTA: columns [KS, H) into the registers of the threads that own gate rows j0 / j1
// (zeros past H), columns [0, KS) into shared memory.  The register part is staged through the shared-memory region
// first, so that every global read is a coalesced float4 stream (per-thread row reads of 256 B each thrash the small
// L1 next to a 147 KB shared-memory carve-out: ncu showed ~20 % of the kernel in this prologue).  Ends with a barrier.
__device__ __forceinline__ void load_whh_onchip(float* W_s, float (&wr0)[LSTM_KR], float (&wr1)[LSTM_KR],
                                                const float* __restrict__ whh, int H, int KS, int pitch, int j0, int j1,
                                                bool on) {
#pragma unroll
  for (int i = 0; i < LSTM_KR; ++i) {
    wr0[i] = 0.f;
    wr1[i] = 0.f;
  }
  if (H > KS) {
    const int nr = H - KS;                                   // <= LSTM_KR <= pitch
    load_whh_cols(W_s, whh, H, KS, nr, pitch);
    __syncthreads();
    if (on) {
#pragma unroll
      for (int i = 0; i < LSTM_KR; i += 4) {
        if (i < nr) {
          const float4 a = *reinterpret_cast<const float4*>(W_s + (size_t)j0 * pitch + i);
          const float4 b = *reinterpret_cast<const float4*>(W_s + (size_t)j1 * pitch + i);
          wr0[i] = a.x; wr0[i + 1] = a.y; wr0[i + 2] = a.z; wr0[i + 3] = a.w;
          wr1[i] = b.x; wr1[i + 1] = b.y; wr1[i + 2] = b.z; wr1[i + 3] = b.w;
        }
      }
    }
    __syncthreads();
  }
  load_whh_cols(W_s, whh, H, 0, KS, pitch);
  __syncthreads();
}

struct LstmDirPtrs {
  const float* w_hh[2];   // [4H, H]
  const float* b_ih[2];   // [4H]
  const float* b_hh[2];   // [4H]
};

// xproj [S, ndir*4H] (no bias), out [S, ndir*H], gates [S, ndir, 4H], cst [S, ndir, H], hprev [S, ndir, H]
// HT: compile-time hidden size (128: every loop bound is a constant, the shared-memory and register halves of the product
// are interleaved k-block by k-block and fully unrolled so that the LDS.128 stream runs ahead of the FMAs - ncu showed
// 26 % short-scoreboard stalls on the first FMA after each shared-memory read), 0: runtime H.
template <int HT>
__global__ void __launch_bounds__(2 * LSTM_MAX_H, 1)
lstm_fwd_kernel(int Hrt, int ndir, const int32_t* __restrict__ gptr, const float* __restrict__ xproj, LstmDirPtrs p,
                float* __restrict__ out, float* __restrict__ gates, float* __restrict__ cst,
                float* __restrict__ hprev) {
  pdl_prologue();
  extern __shared__ float4 lstm_smem4[];
  float* smem = reinterpret_cast<float*>(lstm_smem4);
  const int H = HT ? HT : Hrt;
  const int G4 = 4 * H, H2 = 2 * H;
  const int KS = H < LSTM_KS_MAX ? H : LSTM_KS_MAX;
  const int pitch = KS + 4;
  float* W_s = smem;                               // [4H][pitch]
  float* h_s = W_s + (size_t)G4 * pitch;           // [LSTM_KS_MAX + LSTM_KR] zero padded
  float* a_s = h_s + LSTM_KS_MAX + LSTM_KR;        // [4H]
  const int b = blockIdx.x, dir = blockIdx.y;
  const int t = threadIdx.x;
  const bool on = t < H2;
  const int j0 = t, j1 = t + H2;                   // the two gate rows of this thread
  const float* whh = p.w_hh[dir];
  for (int i = threadIdx.x; i < LSTM_KS_MAX + LSTM_KR; i += blockDim.x) h_s[i] = 0.f;
  float wr0[LSTM_KR], wr1[LSTM_KR];
  load_whh_onchip(W_s, wr0, wr1, whh, H, KS, pitch, j0, j1, on);
  const float bias0 = on ? __ldg(p.b_ih[dir] + j0) + __ldg(p.b_hh[dir] + j0) : 0.f;
  const float bias1 = on ? __ldg(p.b_ih[dir] + j1) + __ldg(p.b_hh[dir] + j1) : 0.f;
  const bool tanh1 = on && t < H;                  // row t + 2H is a g-gate row (tanh) for t < H, an o-gate row otherwise
  const int r0 = gptr[b], T = gptr[b + 1] - r0;
  const int ldx = ndir * G4;
  float c = 0.f;
  __syncthreads();
  float xa0 = 0.f, xa1 = 0.f;
  if (on && T > 0) {
    const float* xr = xproj + (size_t)(r0 + (dir ? T - 1 : 0)) * ldx + dir * G4;
    xa0 = __ldg(xr + j0);
    xa1 = __ldg(xr + j1);
  }
  for (int step = 0; step < T; ++step) {
    const int row = r0 + (dir ? T - 1 - step : step);
    float xn0 = 0.f, xn1 = 0.f;
    if (on && step + 1 < T) {
      const float* xr = xproj + (size_t)(r0 + (dir ? T - 2 - step : step + 1)) * ldx + dir * G4;
      xn0 = __ldg(xr + j0);
      xn1 = __ldg(xr + j1);
    }
    if (on) {
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;      // two partial sums per row
      const float* w0 = W_s + (size_t)j0 * pitch;
      const float* w1 = W_s + (size_t)j1 * pitch;
      if constexpr (HT == 2 * LSTM_KS_MAX) {
#pragma unroll
        for (int k = 0; k < LSTM_KS_MAX; k += 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(h_s + k);
          const float4 g4 = *reinterpret_cast<const float4*>(h_s + LSTM_KS_MAX + k);
          const float4 u4 = *reinterpret_cast<const float4*>(w0 + k);
          const float4 v4 = *reinterpret_cast<const float4*>(w1 + k);
          a0 = fmaf(wr0[k], g4.x, a0);
          a1 = fmaf(wr0[k + 1], g4.y, a1);
          b0 = fmaf(wr1[k], g4.x, b0);
          b1 = fmaf(wr1[k + 1], g4.y, b1);
          a0 = fmaf(wr0[k + 2], g4.z, a0);
          a1 = fmaf(wr0[k + 3], g4.w, a1);
          b0 = fmaf(wr1[k + 2], g4.z, b0);
          b1 = fmaf(wr1[k + 3], g4.w, b1);
          a0 = fmaf(u4.x, h4.x, a0);
          a1 = fmaf(u4.y, h4.y, a1);
          b0 = fmaf(v4.x, h4.x, b0);
          b1 = fmaf(v4.y, h4.y, b1);
          a0 = fmaf(u4.z, h4.z, a0);
          a1 = fmaf(u4.w, h4.w, a1);
          b0 = fmaf(v4.z, h4.z, b0);
          b1 = fmaf(v4.w, h4.w, b1);
        }
      } else {
#pragma unroll 4
      for (int k = 0; k < KS; k += 4) {
        const float4 h4 = *reinterpret_cast<const float4*>(h_s + k);
        const float4 u4 = *reinterpret_cast<const float4*>(w0 + k);
        const float4 v4 = *reinterpret_cast<const float4*>(w1 + k);
        a0 = fmaf(u4.x, h4.x, a0);
        a1 = fmaf(u4.y, h4.y, a1);
        a0 = fmaf(u4.z, h4.z, a0);
        a1 = fmaf(u4.w, h4.w, a1);
        b0 = fmaf(v4.x, h4.x, b0);
        b1 = fmaf(v4.y, h4.y, b1);
        b0 = fmaf(v4.z, h4.z, b0);
        b1 = fmaf(v4.w, h4.w, b1);
      }
      if (H > KS) {
#pragma unroll
        for (int i = 0; i < LSTM_KR; i += 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(h_s + KS + i);
          a0 = fmaf(wr0[i], h4.x, a0);
          a1 = fmaf(wr0[i + 1], h4.y, a1);
          a0 = fmaf(wr0[i + 2], h4.z, a0);
          a1 = fmaf(wr0[i + 3], h4.w, a1);
          b0 = fmaf(wr1[i], h4.x, b0);
          b1 = fmaf(wr1[i + 1], h4.y, b1);
          b0 = fmaf(wr1[i + 2], h4.z, b0);
          b1 = fmaf(wr1[i + 3], h4.w, b1);
        }
      }
      }
      const float act0 = sigmoid_acc((a0 + a1) + xa0 + bias0);            // rows < 2H: i and f gates
      const float pre1 = (b0 + b1) + xa1 + bias1;
      const float act1 = tanh1 ? tanhf(pre1) : sigmoid_acc(pre1);
      a_s[j0] = act0;
      a_s[j1] = act1;
      float* gr = gates + ((size_t)row * ndir + dir) * G4;
      gr[j0] = act0;
      gr[j1] = act1;
    }
    __syncthreads();
    if (t < H) {
      const float ig = a_s[t], fg = a_s[H + t], gg = a_s[2 * H + t], og = a_s[3 * H + t];
      const size_t o = ((size_t)row * ndir + dir) * H + t;
      hprev[o] = h_s[t];
      c = fmaf(fg, c, ig * gg);
      const float hv = og * tanhf(c);
      cst[o] = c;
      out[o] = hv;
      h_s[t] = hv;
    }
    xa0 = xn0;
    xa1 = xn1;
    __syncthreads();
  }
}

// d_out [S, ndir*H]; da [S, ndir*4H] (pre-activation gate gradients, the x-projection layout)
//
// The transposed product dh_{t-1}[k] = sum_j da[j] W_hh[j][k] uses the SAME on-chip copy of W_hh as the forward:
//   * columns k < KS (shared memory): thread (kg, jc) owns 4 consecutive columns and a chunk of gate rows - float4
//     reads along k (conflict-free), the row chunks reduced through shared memory;
//   * columns k >= KS (registers of the thread that owns gate rows t, t + 2H): every thread forms
//     da[t] W[t][k] + da[t+2H] W[t+2H][k] for its 64 columns, and the column sums are reduced across the warp with two
//     31-shuffle butterflies (32 columns each), then across the warps through shared memory.
__global__ void __launch_bounds__(2 * LSTM_MAX_H, 1)
lstm_bwd_kernel(int H, int ndir, const int32_t* __restrict__ gptr, const float* __restrict__ d_out,
                const float* __restrict__ gates, const float* __restrict__ cst, LstmDirPtrs p,
                float* __restrict__ da) {
  pdl_prologue();
  extern __shared__ float4 lstm_smem4[];
  float* smem = reinterpret_cast<float*>(lstm_smem4);
  const int G4 = 4 * H, H2 = 2 * H;
  const int KS = H < LSTM_KS_MAX ? H : LSTM_KS_MAX;
  const int pitch = KS + 4;
  float* W_s = smem;                                    // [4H][pitch]
  float* da_s = W_s + (size_t)G4 * pitch;               // [4 * LSTM_MAX_H]
  float* part_s = da_s + 4 * LSTM_MAX_H;                // [LSTM_NJ_MAX][LSTM_KS_MAX]  partial sums, columns k < KS
  float* regp_s = part_s + LSTM_NJ_MAX * LSTM_KS_MAX;   // [LSTM_MAX_WARPS][LSTM_KR]   partial sums, columns k >= KS
  const int b = blockIdx.x, dir = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool on = tid < H2;
  const float* whh = p.w_hh[dir];
  for (int i = threadIdx.x; i < LSTM_NJ_MAX * LSTM_KS_MAX + LSTM_MAX_WARPS * LSTM_KR; i += blockDim.x) part_s[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * LSTM_MAX_H; i += blockDim.x) da_s[i] = 0.f;
  const bool has_reg = H > KS;
  float wr0[LSTM_KR], wr1[LSTM_KR];
  load_whh_onchip(W_s, wr0, wr1, whh, H, KS, pitch, tid, tid + H2, on);
  const int nkg = KS >> 2;                                    // float4 column groups in shared memory
  const int nj = (H2 >= LSTM_NJ_MAX * nkg) ? LSTM_NJ_MAX : LSTM_NJ_MAX / 2;   // gate-row chunks (2H >= 8 nkg always)
  const int kg = tid % nkg, jc = tid / nkg;                   // this thread's column group and gate-row chunk
  const bool smem_on = jc < nj;
  const int jper = G4 / nj;                                   // gate rows per chunk
  const int r0 = gptr[b], T = gptr[b + 1] - r0;
  const int u = tid;                                          // pointwise: hidden unit (u < H)
  float dc_carry = 0.f;
  __syncthreads();
  float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cc = 0.f, cp = 0.f, go = 0.f;
  auto load_step = [&](int step) {
    const int row = r0 + (dir ? T - 1 - step : step);
    const size_t gb = ((size_t)row * ndir + dir) * G4;
    ig = __ldg(gates + gb + u);
    fg = __ldg(gates + gb + H + u);
    gg = __ldg(gates + gb + 2 * H + u);
    og = __ldg(gates + gb + 3 * H + u);
    cc = __ldg(cst + ((size_t)row * ndir + dir) * H + u);
    const int prow = dir ? row + 1 : row - 1;                // row of forward step - 1
    cp = step > 0 ? __ldg(cst + ((size_t)prow * ndir + dir) * H + u) : 0.f;
    go = __ldg(d_out + (size_t)row * ndir * H + dir * H + u);
  };
  if (u < H && T > 0) load_step(T - 1);
  for (int step = T - 1; step >= 0; --step) {
    const int row = r0 + (dir ? T - 1 - step : step);
    if (u < H) {
      float rec = 0.f;                                        // dh from step + 1, fixed summation order
      if (u < KS) {
        for (int q = 0; q < nj; ++q) rec += part_s[q * LSTM_KS_MAX + u];
      } else {
#pragma unroll
        for (int q = 0; q < LSTM_MAX_WARPS; ++q) rec += regp_s[q * LSTM_KR + (u - KS)];
      }
      const float dh = go + rec;
      const float tc = tanhf(cc);
      const float d_o = dh * tc;
      const float dc = fmaf(dh * og, 1.f - tc * tc, dc_carry);
      const float d_i = dc * gg, d_g = dc * ig, d_f = dc * cp;
      dc_carry = dc * fg;
      const float ai = d_i * ig * (1.f - ig), af = d_f * fg * (1.f - fg), ag = d_g * (1.f - gg * gg),
                  ao = d_o * og * (1.f - og);
      da_s[u] = ai;
      da_s[H + u] = af;
      da_s[2 * H + u] = ag;
      da_s[3 * H + u] = ao;
      float* dst = da + (size_t)row * ndir * G4 + dir * G4;
      dst[u] = ai;
      dst[H + u] = af;
      dst[2 * H + u] = ag;
      dst[3 * H + u] = ao;
      if (step > 0) load_step(step - 1);                      // prefetch: lands during the product below
    }
    __syncthreads();
    if (step > 0) {                                           // the product feeds the next (earlier) step only
      if (has_reg) {                                          // columns KS .. KS+63 from registers, butterfly reduce
        const float d0 = on ? da_s[tid] : 0.f, d1 = on ? da_s[tid + H2] : 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(d0, wr0[half * 32 + i], d1 * wr1[half * 32 + i]);
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float keep = up ? v[i + off] : v[i];
              const float send = up ? v[i] : v[i + off];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          regp_s[warp * LSTM_KR + half * 32 + lane] = v[0];   // lane l holds column KS + 32 half + l
        }
      }
      if (smem_on) {                                          // columns k < KS from shared memory
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int jb = jc * jper;
        const float* wp = W_s + (size_t)jb * pitch + 4 * kg;
#pragma unroll 4
        for (int jj = 0; jj < jper; ++jj) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + (size_t)jj * pitch);
          const float dj = da_s[jb + jj];
          acc.x = fmaf(dj, w4.x, acc.x);
          acc.y = fmaf(dj, w4.y, acc.y);
          acc.z = fmaf(dj, w4.z, acc.z);
          acc.w = fmaf(dj, w4.w, acc.w);
        }
        *reinterpret_cast<float4*>(part_s + jc * LSTM_KS_MAX + 4 * kg) = acc;
      }
    }
    __syncthreads();
  }
}

static size_t lstm_smem_bytes(int H) {
  const int KS = H < LSTM_KS_MAX ? H : LSTM_KS_MAX;
  const size_t w = (size_t)4 * H * (KS + 4);
  const size_t fwd = w + LSTM_KS_MAX + LSTM_KR + 4 * H;
  const size_t bwd = w + 4 * LSTM_MAX_H + LSTM_NJ_MAX * LSTM_KS_MAX + LSTM_MAX_WARPS * LSTM_KR;
  return (fwd > bwd ? fwd : bwd) * sizeof(float);
}

static int lstm_check(int n_graphs, int H, int ndir) {
  if (n_graphs < 0 || H <= 0 || (ndir != 1 && ndir != 2)) return HSG_ERR_ARG;
  if (H > LSTM_MAX_H || H % 4 != 0) return HSG_ERR_SHAPE;
  return HSG_OK;
}

// opt in to the large dynamic shared-memory carve-out once per kernel and size (not on every launch)
template <int TAG, typename K>                      // TAG: one static per KERNEL (the three share one function type)
static int lstm_attr(K kernel, int H) {
  static int done_bytes = 0;
  const int need = (int)lstm_smem_bytes(H);
  if (need <= done_bytes) return HSG_OK;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need) != cudaSuccess) return HSG_ERR_CUDA;
  done_bytes = need;
  return HSG_OK;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_lstm_fwd(int n_graphs, int H, int ndir, const int32_t* graph_sent_ptr, const float* xproj,
                 const float* const* w_hh, const float* const* b_ih, const float* const* b_hh, float* out,
                 float* gates, float* cst, float* hprev, void* stream) {
  int rc = lstm_check(n_graphs, H, ndir);
  if (rc != HSG_OK) return rc;
  if (n_graphs == 0) return HSG_OK;
  if (!graph_sent_ptr || !xproj || !w_hh || !b_ih || !b_hh || !out || !gates || !cst || !hprev) return HSG_ERR_ARG;
  LstmDirPtrs p;
  for (int d = 0; d < 2; ++d) {
    const int s = d < ndir ? d : 0;
    if (!w_hh[s] || !b_ih[s] || !b_hh[s]) return HSG_ERR_ARG;
    if (!aligned16(w_hh[s])) return HSG_ERR_ALIGN;
    p.w_hh[d] = w_hh[s];
    p.b_ih[d] = b_ih[s];
    p.b_hh[d] = b_hh[s];
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = ((2 * H + 31) / 32) * 32;
  if (H == 2 * LSTM_KS_MAX) {
    if ((rc = lstm_attr<0>(lstm_fwd_kernel<2 * LSTM_KS_MAX>, H)) != HSG_OK) return rc;
    LaunchScope ls(SLOT_LSTM, s);
    launch_k(lstm_fwd_kernel<2 * LSTM_KS_MAX>, dim3(n_graphs, ndir), dim3(threads), lstm_smem_bytes(H), s, H, ndir,
             graph_sent_ptr, xproj, p, out, gates, cst, hprev);
  } else {
    if ((rc = lstm_attr<1>(lstm_fwd_kernel<0>, H)) != HSG_OK) return rc;
    LaunchScope ls(SLOT_LSTM, s);
    launch_k(lstm_fwd_kernel<0>, dim3(n_graphs, ndir), dim3(threads), lstm_smem_bytes(H), s, H, ndir, graph_sent_ptr,
             xproj, p, out, gates, cst, hprev);
  }
  return check_launch();
}

int hsg_lstm_bwd(int n_graphs, int H, int ndir, const int32_t* graph_sent_ptr, const float* d_out, const float* gates,
                 const float* cst, const float* const* w_hh, float* da, void* stream) {
  int rc = lstm_check(n_graphs, H, ndir);
  if (rc != HSG_OK) return rc;
  if (n_graphs == 0) return HSG_OK;
  if (!graph_sent_ptr || !d_out || !gates || !cst || !w_hh || !da) return HSG_ERR_ARG;
  LstmDirPtrs p;
  for (int d = 0; d < 2; ++d) {
    const int s = d < ndir ? d : 0;
    if (!w_hh[s]) return HSG_ERR_ARG;
    if (!aligned16(w_hh[s])) return HSG_ERR_ALIGN;
    p.w_hh[d] = w_hh[s];
    p.b_ih[d] = nullptr;
    p.b_hh[d] = nullptr;
  }
  if ((rc = lstm_attr<2>(lstm_bwd_kernel, H)) != HSG_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_LSTM, s);
  const int threads = ((2 * H + 31) / 32) * 32;
  launch_k(lstm_bwd_kernel, dim3(n_graphs, ndir), dim3(threads), lstm_smem_bytes(H), s, H, ndir, graph_sent_ptr, d_out,
           gates, cst, p, da);
  return check_launch();
}

}  // extern "C"
