// Recurrent part of the sentence-level BiLSTM (HiGraph.py:118-119, 135-142: nn.LSTM over the packed per-graph
// sentence sequences), forward and backward-through-time, one layer per call.
//
// The reference pads / packs the per-graph sentence lists and lets cuDNN walk the time steps with one small launch
// sequence per step (measured here: ~9 ms per training step at 32 graphs, 12x the whole WSWGAT path).  Every graph is
// an independent sequence whose rows are already contiguous in the batched sentence order, so:
//   * the input products  x . W_ih^T  of ALL time steps are one GEMM per direction (caller, hsg_gemm_nt);
//   * the recurrence runs in ONE persistent kernel per layer: CTA (graph, direction), one thread per gate row
//     (4H <= 512 threads).  W_hh stays on chip for the whole sequence: columns k < 96 in shared memory (row pitch
//     KS + 4 floats: the float4 reads of 8 consecutive rows cover all 32 banks), columns 96..127 in 32 registers
//     per thread; h in shared memory, c in a register.  No global traffic on the critical path except the
//     prefetched x-projection row;
//   * backward: the same CTA walks its sequence in reverse, recomputes nothing (post-activation gates, c and
//     h_{t-1} were saved by the forward), keeps dh / dc on chip and writes the pre-activation gradients da [S, 4H]
//     per direction; the weight / input gradients are GEMMs over all time steps at once (caller: gemm_tn, gemm_nn).
// Gate order i, f, g, o (torch.nn.LSTM); accurate expf / tanhf (fp32 parity with the CPU reference <= 1e-5).
#include "hsg_common.cuh"

namespace hsg {

constexpr int LSTM_MAX_H = 128;
constexpr int LSTM_KS_MAX = 96;      // W_hh columns kept in shared memory
constexpr int LSTM_KR = 32;          // W_hh columns kept in registers (k = KS .. KS + 31)

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// W_hh[:, 0:KS] -> shared memory rows of `pitch` floats: float4 copies, 8 independent loads in flight per thread (a
// scalar load -> store loop costs ~100 serial L2 latencies per thread, ncu: 17 % of the kernel's samples)
__device__ __forceinline__ void load_whh_smem(float* W_s, const float* __restrict__ whh, int H, int KS, int pitch) {
  const int nk4 = KS >> 2, tot4 = 4 * H * nk4;
  for (int base = threadIdx.x; base < tot4; base += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i4 = base + q * blockDim.x;
      if (i4 < tot4) {
        const int jj = i4 / nk4, k4 = i4 - jj * nk4;
        v[q] = __ldg(reinterpret_cast<const float4*>(whh + (size_t)jj * H) + k4);
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

struct LstmDirPtrs {
  const float* w_hh[2];   // [4H, H]
  const float* b_ih[2];   // [4H]
  const float* b_hh[2];   // [4H]
};

// xproj [S, ndir*4H] (no bias), out [S, ndir*H], gates [S, ndir, 4H], cst [S, ndir, H], hprev [S, ndir, H]
__global__ void __launch_bounds__(512, 1)
lstm_fwd_kernel(int H, int ndir, const int32_t* __restrict__ gptr, const float* __restrict__ xproj, LstmDirPtrs p,
                float* __restrict__ out, float* __restrict__ gates, float* __restrict__ cst,
                float* __restrict__ hprev) {
  pdl_prologue();
  extern __shared__ float4 lstm_smem4[];
  float* smem = reinterpret_cast<float*>(lstm_smem4);
  const int G4 = 4 * H;
  const int KS = H < LSTM_KS_MAX ? H : LSTM_KS_MAX;
  const int pitch = KS + 4;
  float* W_s = smem;                               // [4H][pitch]
  float* h_s = W_s + (size_t)G4 * pitch;           // [LSTM_MAX_H + LSTM_KR] zero padded
  float* a_s = h_s + LSTM_MAX_H + LSTM_KR;         // [4H]
  const int b = blockIdx.x, dir = blockIdx.y;
  const int j = threadIdx.x;
  const bool on = j < G4;
  const float* whh = p.w_hh[dir];
  load_whh_smem(W_s, whh, H, KS, pitch);
  for (int i = threadIdx.x; i < LSTM_MAX_H + LSTM_KR; i += blockDim.x) h_s[i] = 0.f;
  float wreg[LSTM_KR];
#pragma unroll
  for (int i = 0; i < LSTM_KR; ++i) wreg[i] = (on && KS + i < H) ? __ldg(whh + (size_t)j * H + KS + i) : 0.f;
  const float bias = on ? __ldg(p.b_ih[dir] + j) + __ldg(p.b_hh[dir] + j) : 0.f;
  const bool is_g = on && (j / H) == 2;
  const int r0 = gptr[b], T = gptr[b + 1] - r0;
  const int ldx = ndir * G4;
  float c = 0.f;
  __syncthreads();
  float xa = (on && T > 0) ? __ldg(xproj + (size_t)(r0 + (dir ? T - 1 : 0)) * ldx + dir * G4 + j) : 0.f;
  for (int step = 0; step < T; ++step) {
    const int row = r0 + (dir ? T - 1 - step : step);
    float xa_next = 0.f;
    if (on && step + 1 < T) xa_next = __ldg(xproj + (size_t)(r0 + (dir ? T - 2 - step : step + 1)) * ldx + dir * G4 + j);
    float acc0 = 0.f, acc1 = 0.f;
    if (on) {
      const float* wr = W_s + j * pitch;
      for (int k = 0; k < KS; k += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
        const float4 h4 = *reinterpret_cast<const float4*>(h_s + k);
        acc0 = fmaf(w4.x, h4.x, acc0);
        acc1 = fmaf(w4.y, h4.y, acc1);
        acc0 = fmaf(w4.z, h4.z, acc0);
        acc1 = fmaf(w4.w, h4.w, acc1);
      }
      if (H > KS) {
#pragma unroll
        for (int i = 0; i < LSTM_KR; i += 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(h_s + KS + i);
          acc0 = fmaf(wreg[i], h4.x, acc0);
          acc1 = fmaf(wreg[i + 1], h4.y, acc1);
          acc0 = fmaf(wreg[i + 2], h4.z, acc0);
          acc1 = fmaf(wreg[i + 3], h4.w, acc1);
        }
      }
      const float a = (acc0 + acc1) + xa + bias;
      const float act = is_g ? tanhf(a) : sigmoid_acc(a);
      a_s[j] = act;
      gates[((size_t)row * ndir + dir) * G4 + j] = act;
    }
    __syncthreads();
    if (j < H) {
      const float ig = a_s[j], fg = a_s[H + j], gg = a_s[2 * H + j], og = a_s[3 * H + j];
      const size_t o = ((size_t)row * ndir + dir) * H + j;
      hprev[o] = h_s[j];
      c = fmaf(fg, c, ig * gg);
      const float hv = og * tanhf(c);
      cst[o] = c;
      out[o] = hv;
      h_s[j] = hv;
    }
    xa = xa_next;
    __syncthreads();
  }
}

// d_out [S, ndir*H]; da [S, ndir*4H] (pre-activation gate gradients, the x-projection layout)
//
// The transposed product dh_{t-1}[k] = sum_j da[j] W_hh[j][k] uses the SAME on-chip copy of W_hh as the forward:
//   * columns k < KS (shared memory): thread (kg, jc) owns 4 consecutive columns and a chunk of H/4 gate rows -
//     float4 reads along k (conflict-free), 16 row chunks reduced through shared memory;
//   * columns k >= KS (registers of the thread that owns gate row j): every thread forms da[j] * W[j][KS..KS+31]
//     and the 32 column sums are reduced across the warp with a 31-shuffle butterfly, then across the 16 warps.
constexpr int LSTM_NJ = 16;          // gate-row chunks of the shared-memory part

__global__ void __launch_bounds__(512, 1)
lstm_bwd_kernel(int H, int ndir, const int32_t* __restrict__ gptr, const float* __restrict__ d_out,
                const float* __restrict__ gates, const float* __restrict__ cst, LstmDirPtrs p,
                float* __restrict__ da) {
  pdl_prologue();
  extern __shared__ float4 lstm_smem4[];
  float* smem = reinterpret_cast<float*>(lstm_smem4);
  const int G4 = 4 * H;
  const int KS = H < LSTM_KS_MAX ? H : LSTM_KS_MAX;
  const int pitch = KS + 4;
  float* W_s = smem;                               // [4H][pitch]
  float* da_s = W_s + (size_t)G4 * pitch;          // [4 * LSTM_MAX_H]
  float* part_s = da_s + 4 * LSTM_MAX_H;           // [LSTM_NJ][LSTM_KS_MAX]  partial sums, columns k < KS
  float* regp_s = part_s + LSTM_NJ * LSTM_KS_MAX;  // [16 warps][LSTM_KR]     partial sums, columns k >= KS
  const int b = blockIdx.x, dir = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool on = tid < G4;
  const float* whh = p.w_hh[dir];
  load_whh_smem(W_s, whh, H, KS, pitch);
  for (int i = threadIdx.x; i < LSTM_NJ * LSTM_KS_MAX + 16 * LSTM_KR; i += blockDim.x) part_s[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * LSTM_MAX_H; i += blockDim.x) da_s[i] = 0.f;
  const bool has_reg = H > KS;
  float wreg[LSTM_KR];
#pragma unroll
  for (int i = 0; i < LSTM_KR; ++i) wreg[i] = (on && KS + i < H) ? __ldg(whh + (size_t)tid * H + KS + i) : 0.f;
  const int nkg = KS >> 2;                                    // float4 column groups in shared memory
  const int kg = tid % nkg, jc = tid / nkg;                   // this thread's column group and gate-row chunk
  const bool smem_on = jc < LSTM_NJ;
  const int jper = G4 / LSTM_NJ;                              // gate rows per chunk (= H / 4)
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
#pragma unroll
        for (int q = 0; q < LSTM_NJ; ++q) rec += part_s[q * LSTM_KS_MAX + u];
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) rec += regp_s[q * LSTM_KR + (u - KS)];
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
      if (has_reg) {                                          // columns KS .. KS+31 from registers, butterfly reduce
        const float dj = on ? da_s[tid] : 0.f;
        float v[LSTM_KR];
#pragma unroll
        for (int i = 0; i < LSTM_KR; ++i) v[i] = dj * wreg[i];
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
        regp_s[warp * LSTM_KR + lane] = v[0];                 // lane l holds column KS + l (warps past 4H hold zeros)
      }
      if (smem_on) {                                          // columns k < KS from shared memory
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int j0 = jc * jper;
        const float* wp = W_s + (size_t)j0 * pitch + 4 * kg;
#pragma unroll 4
        for (int jj = 0; jj < jper; ++jj) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + (size_t)jj * pitch);
          const float dj = da_s[j0 + jj];
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
  const size_t fwd = w + LSTM_MAX_H + LSTM_KR + 4 * H;
  const size_t bwd = w + 4 * LSTM_MAX_H + LSTM_NJ * LSTM_KS_MAX + 16 * LSTM_KR;
  return (fwd > bwd ? fwd : bwd) * sizeof(float);
}

static int lstm_check(int n_graphs, int H, int ndir) {
  if (n_graphs < 0 || H <= 0 || (ndir != 1 && ndir != 2)) return HSG_ERR_ARG;
  if (H > LSTM_MAX_H || H % 4 != 0) return HSG_ERR_SHAPE;
  return HSG_OK;
}

template <typename K>
static int lstm_attr(K kernel, int H) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem_bytes(H)) == cudaSuccess
             ? HSG_OK
             : HSG_ERR_CUDA;
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
  if ((rc = lstm_attr(lstm_fwd_kernel, H)) != HSG_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_LSTM, s);
  const int threads = ((4 * H + 31) / 32) * 32;
  launch_k(lstm_fwd_kernel, dim3(n_graphs, ndir), dim3(threads), lstm_smem_bytes(H), s, H, ndir, graph_sent_ptr, xproj, p,
           out, gates, cst, hprev);
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
  if ((rc = lstm_attr(lstm_bwd_kernel, H)) != HSG_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_LSTM, s);
  const int threads = ((4 * H + 31) / 32) * 32;
  launch_k(lstm_bwd_kernel, dim3(n_graphs, ndir), dim3(threads), lstm_smem_bytes(H), s, H, ndir, graph_sent_ptr, d_out,
           gates, cst, p, da);
  return check_launch();
}

}  // extern "C"
