// K4 row-wise pieces of the position-wise FFN (module/GATLayer.py:35-44):
//   y = LayerNorm(r) * gamma + beta  with r = W2 relu(W1 x + b1) + b2 + x  (r comes from hsg_gemm_nt's epilogue)
// One warp per row, the row lives in registers (D <= 512, D % 4 == 0), 128-bit accesses.
// dgamma / dbeta are reduced in a fixed order (per-warp registers -> per-block -> second stage).
#include "hsg_common.cuh"
#include "hsg_internal.cuh"

namespace hsg {

constexpr int LN_WARPS = 8;
constexpr int LN_THREADS = LN_WARPS * 32;
constexpr int LN_MAX_BLOCKS = 148 * 4;

// DROPRES: r holds the FFN output y = W2 relu(.) + b2; the kernel forms r = dropout(y) + resid (GATLayer.py:41-42),
// writes it back (backward needs r) and normalises it.
template <int NV4, bool DROPRES>  // float4 per lane; D <= 128 * NV4
__global__ void __launch_bounds__(LN_THREADS)
layernorm_fwd_kernel(int N, int D, const float* r, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ stats,
                     float* r_out, const float* __restrict__ resid, DropCfg dc) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int D4 = D >> 2;
  float4 gm[NV4], bt[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < D4 ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = c < D4 ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.f / (float)D;
  for (int row = warp; row < N; row += nwarps) {
    const float4* rr = reinterpret_cast<const float4*>(r + (size_t)row * D);
    float4 v[NV4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < D4 ? (DROPRES ? rr[c] : __ldg(rr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (DROPRES && c < D4) {
        const float4 x4 = __ldg(reinterpret_cast<const float4*>(resid + (size_t)row * D) + c);
        const unsigned long long e = (unsigned long long)row * (unsigned long long)D + 4ull * c;
        v[i].x = (drop_keep(dc, e + 0) ? v[i].x * dc.scale : 0.f) + x4.x;
        v[i].y = (drop_keep(dc, e + 1) ? v[i].y * dc.scale : 0.f) + x4.y;
        v[i].z = (drop_keep(dc, e + 2) ? v[i].z * dc.scale : 0.f) + x4.z;
        v[i].w = (drop_keep(dc, e + 3) ? v[i].w * dc.scale : 0.f) + x4.w;
        reinterpret_cast<float4*>(r_out + (size_t)row * D)[c] = v[i];
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * invD;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < D4) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
        ss += (a * a + b * b) + (cc * cc + dd * dd);
      }
    }
    const float var = warp_sum(ss) * invD;
    const float rstd = rsqrtf(var + HSG_LN_EPS);
    float4* yy = reinterpret_cast<float4*>(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < D4) {
        float4 o;
        o.x = (v[i].x - mean) * rstd * gm[i].x + bt[i].x;
        o.y = (v[i].y - mean) * rstd * gm[i].y + bt[i].y;
        o.z = (v[i].z - mean) * rstd * gm[i].z + bt[i].z;
        o.w = (v[i].w - mean) * rstd * gm[i].w + bt[i].w;
        yy[c] = o;
      }
    }
    if (lane == 0) {
      stats[2 * (size_t)row] = mean;
      stats[2 * (size_t)row + 1] = rstd;
    }
  }
}

template <int NV4>
__global__ void __launch_bounds__(LN_THREADS)
layernorm_bwd_kernel(int N, int D, const float* __restrict__ dy, const float* __restrict__ r,
                     const float* __restrict__ stats, const float* __restrict__ gamma, float* __restrict__ dr,
                     float* __restrict__ part /* [gridDim.x][2][D] */) {
  pdl_prologue();
  extern __shared__ float red[];  // [LN_WARPS][2][D]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int D4 = D >> 2;
  float4 gm[NV4], dg[NV4], db[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < D4 ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.f / (float)D;
  // software pipeline over rows: the NEXT grid-stride row of this warp is fetched into registers while the current one
  // goes through its two warp reductions (measured at 756 k x 300: 3.5 TB/s = 0.53 of the HBM peak without it - a
  // warp had loads in flight for only part of its load -> reduce -> store cycle)
  float4 prv[NV4], pdv[NV4];
  float pmean = 0.f, prstd = 0.f;
  auto prefetch_row = [&](int rw) {
    if (rw < N) {
      const float4* rr = reinterpret_cast<const float4*>(r + (size_t)rw * D);
      const float4* dd = reinterpret_cast<const float4*>(dy + (size_t)rw * D);
      pmean = __ldg(stats + 2 * (size_t)rw);
      prstd = __ldg(stats + 2 * (size_t)rw + 1);
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int c = lane + 32 * i;
        if (c < D4) {
          prv[i] = __ldg(rr + c);
          pdv[i] = __ldg(dd + c);
        }
      }
    }
  };
  prefetch_row(warp);
  for (int row = warp; row < N; row += nwarps) {
    const float mean = pmean, rstd = prstd;
    float4 crv[NV4], cdv[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      crv[i] = prv[i];
      cdv[i] = pdv[i];
    }
    prefetch_row(row + nwarps);
    float4 xh[NV4], gy[NV4];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < D4) {
        const float4 rv = crv[i], dv = cdv[i];
        xh[i] = make_float4((rv.x - mean) * rstd, (rv.y - mean) * rstd, (rv.z - mean) * rstd, (rv.w - mean) * rstd);
        gy[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
        dg[i].x = fmaf(dv.x, xh[i].x, dg[i].x);
        dg[i].y = fmaf(dv.y, xh[i].y, dg[i].y);
        dg[i].z = fmaf(dv.z, xh[i].z, dg[i].z);
        dg[i].w = fmaf(dv.w, xh[i].w, dg[i].w);
        db[i].x += dv.x;
        db[i].y += dv.y;
        db[i].z += dv.z;
        db[i].w += dv.w;
        c1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        c2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      } else {
        xh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        gy[i] = xh[i];
      }
    }
    c1 = warp_sum(c1) * invD;
    c2 = warp_sum(c2) * invD;
    float4* oo = reinterpret_cast<float4*>(dr + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < D4) {
        float4 o;
        o.x = rstd * (gy[i].x - c1 - xh[i].x * c2);
        o.y = rstd * (gy[i].y - c1 - xh[i].y * c2);
        o.z = rstd * (gy[i].z - c1 - xh[i].z * c2);
        o.w = rstd * (gy[i].w - c1 - xh[i].w * c2);
        oo[c] = o;
      }
    }
  }
  // block reduction of dgamma / dbeta in a fixed order
  float* my = red + (size_t)wib * 2 * D;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + 32 * i;
    if (c < D4) {
      reinterpret_cast<float4*>(my)[c] = dg[i];
      reinterpret_cast<float4*>(my + D)[c] = db[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[(size_t)w * 2 * D + i];
    part[(size_t)blockIdx.x * 2 * D + i] = s;
  }
}

// dgamma / dbeta = column sums of the per-block partials of nseg launches (segment j at part + j * seg_stride);
// 32 columns x 32 row groups per CTA, fixed order (row group r sums rows r, r+32, ... of the concatenated segments;
// the 32 group sums are then added in order)
__global__ void __launch_bounds__(1024) layernorm_bwd_reduce_kernel(int nseg, int nblocks, size_t seg_stride, int D,
                                                                    const float* __restrict__ part,
                                                                    float* __restrict__ dgamma,
                                                                    float* __restrict__ dbeta, int accumulate) {
  pdl_prologue();
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < 2 * D)
    for (int b = ty; b < nseg * nblocks; b += 32)
      s += part[(size_t)(b / nblocks) * seg_stride + (size_t)(b % nblocks) * 2 * D + i];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < 2 * D) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += red[r][tx];
    float* o = i < D ? dgamma + i : dbeta + (i - D);
    *o = accumulate ? *o + t : t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Whole position-wise FFN of a SMALL node set in one launch each way (the sentence side of the 32-graph step:
// ~1 000 supernode rows, F = 64, d_hid = 512).  There the three launches forward (GEMM, GEMM, LayerNorm) and three
// backward (LayerNorm, GEMM, GEMM) are 4-10 us each for 66 MFLOP of exact-fp32 work, back to back on the critical
// path of the step.  Here a CTA keeps its 16 rows, their hidden activations and the LayerNorm state in shared memory
// and streams BOTH weight matrices (L2-resident, 128 KB each) through one 16 KB shared-memory chunk buffer: every
// chunk is fetched with four coalesced 128-bit loads per thread that stay in flight while the previous chunk is
// multiplied (register double buffering), so the kernel runs at the CTA's L2 fetch rate instead of one dependent L2
// round trip per weight row (the first version: 39 us per launch).  Exact fp32 (FFMA), fixed summation order, no atomics.
// Outputs are the same tensors the three-kernel path leaves behind (hdn, r, stats / dr, dhp, dx), so the
// weight-gradient products and everything downstream are unchanged.
// Shape: F == 64, d_hid % 256 == 0, d_hid <= 1024; the caller falls back to the GEMM path otherwise.
// ---------------------------------------------------------------------------------------------------------
constexpr int FS_F = 64;
constexpr int FS_ROWS = 16;         // rows per CTA
constexpr int FS_THREADS = 256;
constexpr int FS_CH = 64;           // chunk: 64 x 64 floats = 16 KB
constexpr int FS_WLD = 68;          // chunk row pitch in shared memory: 128-bit reads of 8 consecutive rows hit 32 banks
constexpr int FS_GRP = 4;           // chunks per pipeline step (64 KB): one step = a 16 x 256 slab of work
constexpr int FS_SLOTS = 2 * FS_GRP;   // two groups: one is multiplied while the next one (64 KB) is in flight
constexpr int FS_SLOT_FLOATS = FS_CH * FS_WLD;
constexpr int FS_MAX_DH = 1024;
constexpr int FS_MAX_BLOCKS = LN_MAX_BLOCKS;

__host__ __device__ inline size_t fs_smem_floats(int Dh) {
  return (size_t)2 * FS_ROWS * FS_F + (size_t)FS_ROWS * Dh + (size_t)FS_SLOTS * FS_SLOT_FLOATS;
}

__device__ __forceinline__ void fs_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void fs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fs_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Weight stream of one pass: 2 * nch chunks, the first nch from one matrix, the rest from the other.  A "row chunk" is 64
// full rows of the [Dh, 64] matrix W1, a "column chunk" is columns 64 c .. 64 c + 63 of all 64 rows of the [64, Dh]
// matrix W2.  Group gi = chunks 4 gi .. 4 gi + 3, copied into ring slots 4 (gi & 1) .. + 3 by cp.async (every thread
// 16 pieces of 16 bytes, 16 threads per 256-byte row segment) and committed as ONE async group.
__device__ __forceinline__ void fs_issue_group(float* ring, const float* __restrict__ w1, const float* __restrict__ w2,
                                               bool w1_first, int gi, int nch, int Dh, int tid) {
  if (gi * FS_GRP < 2 * nch) {
#pragma unroll
    for (int t = 0; t < FS_GRP; ++t) {
      const int i = gi * FS_GRP + t;
      const bool from_w1 = (i < nch) == w1_first;
      const int c = i < nch ? i : i - nch;
      const float* base = from_w1 ? w1 + (size_t)c * FS_CH * FS_F : w2 + (size_t)c * FS_CH;
      const int ld = from_w1 ? FS_F : Dh;
      float* slot = ring + ((gi & 1) * FS_GRP + t) * FS_SLOT_FLOATS;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = tid + FS_THREADS * k, rr = f >> 4, q = f & 15;
        fs_cp16(slot + rr * FS_WLD + 4 * q, base + (size_t)rr * ld + 4 * q);
      }
    }
  }
  fs_commit();                                           // always: keeps the group count in step with the step index
}

#define FS_FMA4x4(ACC, A, B)                                                              \
  do {                                                                                    \
    _Pragma("unroll") for (int u_ = 0; u_ < U; ++u_) {                                    \
      _Pragma("unroll") for (int t_ = 0; t_ < 4; ++t_) {                                  \
        ACC[u_][t_] = fmaf(A[u_].x, B[t_].x, ACC[u_][t_]);                                \
        ACC[u_][t_] = fmaf(A[u_].y, B[t_].y, ACC[u_][t_]);                                \
        ACC[u_][t_] = fmaf(A[u_].z, B[t_].z, ACC[u_][t_]);                                \
        ACC[u_][t_] = fmaf(A[u_].w, B[t_].w, ACC[u_][t_]);                                \
      }                                                                                   \
    }                                                                                     \
  } while (0)

// U = rows per thread, R = 4 U rows per CTA: 16 rows (U = 4) by default, 8 rows (U = 2) for small node sets (fs_rows:
// more, shorter CTAs; the weights are re-streamed from L2 by twice as many CTAs).  Every output keeps its summation
// order, so results do not depend on U (except the per-CTA dgamma / dbeta partials of the backward).
template <int U>
__global__ void __launch_bounds__(FS_THREADS)
ffn_rows_fwd_kernel(int n, int Dh, const float* __restrict__ x, const float* __restrict__ w1,
                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ hdn,
                    float* __restrict__ r, float* __restrict__ y, float* __restrict__ stats) {
  pdl_prologue();
  extern __shared__ __align__(16) float fs_smem[];
  float* xs = fs_smem;                                   // [16][64]
  float* rs = xs + FS_ROWS * FS_F;                       // [16][64]
  float* hs = rs + FS_ROWS * FS_F;                       // [16][Dh]
  float* ring = hs + (size_t)FS_ROWS * Dh;               // [8][64][68]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int R = 4 * U;                               // rows of this CTA
  const int rq = tid >> 6;                               // rows U rq .. U rq + U - 1
  const int row0 = blockIdx.x * R;
  const int nch = Dh / FS_CH, ngrp = nch / FS_GRP;       // groups per matrix
  fs_issue_group(ring, w1, w2, true, 0, nch, Dh, tid);
  fs_issue_group(ring, w1, w2, true, 1, nch, Dh, tid);
  {
    const int rr = tid >> 4, c4 = tid & 15;              // 16 rows x 16 float4
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < R && row0 + rr < n) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)(row0 + rr) * FS_F) + c4);
    *reinterpret_cast<float4*>(xs + rr * FS_F + 4 * c4) = v;
  }
  // second product: thread = (4 output columns cq + 16 t, k quarter kq of every chunk, 4 rows)
  const int cq = tid & 15, kq = (tid >> 4) & 3;
  float acc2[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc2[u][t] = 0.f;
  for (int gi = 0; gi < 2 * ngrp; ++gi) {
    fs_wait1();                                          // this thread's pieces of group gi have landed ...
    __syncthreads();                                     // ... and everybody's (first pass: xs is complete too)
    const float* grp = ring + (gi & 1) * FS_GRP * FS_SLOT_FLOATS;
    if (gi < ngrp) {
      // ---- hdn = relu(x W1^T + b1): hidden units 256 gi + 64 t + jq, t = 0..3 (slot t holds W1 rows 256 gi + 64 t ..) ----
      const int jq = tid & 63;
      float acc[U][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float bj = __ldg(b1 + gi * FS_GRP * FS_CH + FS_CH * t + jq);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u][t] = bj;
      }
#pragma unroll 2
      for (int c4 = 0; c4 < FS_F / 4; ++c4) {
        float4 xv[U], wv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) wv[t] = *reinterpret_cast<const float4*>(grp + t * FS_SLOT_FLOATS + jq * FS_WLD + 4 * c4);
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = *reinterpret_cast<const float4*>(xs + (U * rq + u) * FS_F + 4 * c4);   // broadcast
        FS_FMA4x4(acc, xv, wv);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = U * rq + u;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = gi * FS_GRP * FS_CH + FS_CH * t + jq;
          const float h = fmaxf(acc[u][t], 0.f);
          hs[(size_t)rr * Dh + j] = h;
          if (row0 + rr < n) hdn[(size_t)(row0 + rr) * Dh + j] = h;
        }
      }
    } else {
      // ---- r += hdn W2^T over k = 256 (gi - ngrp) + 64 s + 16 kq .. + 15, s = 0..3 (slot s: W2[:, that 64-k slice]) ----
      const int kbase = (gi - ngrp) * FS_GRP * FS_CH;
#pragma unroll
      for (int sl = 0; sl < FS_GRP; ++sl) {
        const float* ws = grp + sl * FS_SLOT_FLOATS;
#pragma unroll 2
        for (int k4 = 0; k4 < 4; ++k4) {
          const int kl = 16 * kq + 4 * k4;
          float4 hv[U], wv[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) wv[t] = *reinterpret_cast<const float4*>(ws + (cq + 16 * t) * FS_WLD + kl);
#pragma unroll
          for (int u = 0; u < U; ++u)
            hv[u] = *reinterpret_cast<const float4*>(hs + (size_t)(U * rq + u) * Dh + kbase + FS_CH * sl + kl);
          FS_FMA4x4(acc2, hv, wv);
        }
      }
    }
    __syncthreads();                                     // group gi consumed by all (and hs complete after the last W1 group)
    fs_issue_group(ring, w1, w2, true, gi + 2, nch, Dh, tid);
  }
  // reduce the four k quarters in order (staging: the ring is free now), add bias and the residual
  {
    float* ps = ring;                                    // [4 kq][16 rows][64 c]
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int t = 0; t < 4; ++t) ps[(kq * R + U * rq + u) * FS_F + cq + 16 * t] = acc2[u][t];
    __syncthreads();
    const int rr = tid >> 4, c4 = tid & 15;
    if (rr < R) {
      float4 o = __ldg(reinterpret_cast<const float4*>(b2) + c4);
      const float4 xv = *reinterpret_cast<const float4*>(xs + rr * FS_F + 4 * c4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 pv = *reinterpret_cast<const float4*>(ps + (q * R + rr) * FS_F + 4 * c4);
        o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
      }
      o.x += xv.x; o.y += xv.y; o.z += xv.z; o.w += xv.w;
      *reinterpret_cast<float4*>(rs + rr * FS_F + 4 * c4) = o;
      if (row0 + rr < n) *reinterpret_cast<float4*>(r + (size_t)(row0 + rr) * FS_F + 4 * c4) = o;
    }
  }
  __syncthreads();
  // ---- LayerNorm: warp w normalises rows w and w + 8 (same formulas as layernorm_fwd_kernel) ----
#pragma unroll
  for (int h2 = 0; h2 < R / 8; ++h2) {
    const int rr = warp + 8 * h2, row = row0 + rr;
    const float v0 = rs[rr * FS_F + lane], v1 = rs[rr * FS_F + lane + 32];
    const float mean = warp_sum(v0 + v1) * (1.f / (float)FS_F);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / (float)FS_F) + HSG_LN_EPS);
    if (row < n) {
      y[(size_t)row * FS_F + lane] = d0 * rstd * __ldg(gamma + lane) + __ldg(beta + lane);
      y[(size_t)row * FS_F + lane + 32] = d1 * rstd * __ldg(gamma + lane + 32) + __ldg(beta + lane + 32);
      if (lane == 0) {
        stats[2 * (size_t)row] = mean;
        stats[2 * (size_t)row + 1] = rstd;
      }
    }
  }
}

template <int U>
__global__ void __launch_bounds__(FS_THREADS)
ffn_rows_bwd_kernel(int n, int Dh, const float* __restrict__ dy, const float* __restrict__ r,
                    const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ hdn,
                    const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ dr,
                    float* __restrict__ dhp, float* __restrict__ dx, float* __restrict__ part /* [grid][2][64] */) {
  pdl_prologue();
  extern __shared__ __align__(16) float fs_smem[];
  float* drs = fs_smem;                                  // [16][64]
  float* pg = drs + FS_ROWS * FS_F;                      // [16][64] dgamma terms
  float* dhs = pg + FS_ROWS * FS_F;                      // [16][Dh]; its head doubles as the dbeta staging area
  float* ring = dhs + (size_t)FS_ROWS * Dh;              // [8][64][68]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int R = 4 * U;
  const int rq = tid >> 6;
  const int row0 = blockIdx.x * R;
  const int nch = Dh / FS_CH, ngrp = nch / FS_GRP;
  // weight stream of this pass: W2 column chunks, then W1 row chunks; the first two groups travel during the LayerNorm part
  fs_issue_group(ring, w1, w2, false, 0, nch, Dh, tid);
  fs_issue_group(ring, w1, w2, false, 1, nch, Dh, tid);
  float* pb = dhs;                                       // [16][64] dbeta terms (dhs is written only after the barriers below)
  // ---- LayerNorm backward: warp w owns rows w and w + 8 ----
#pragma unroll
  for (int h2 = 0; h2 < R / 8; ++h2) {
    const int rr = warp + 8 * h2, row = row0 + rr;
    float mean = 0.f, rstd = 0.f, dv[2] = {0.f, 0.f}, rv[2] = {0.f, 0.f};
    if (row < n) {
      mean = __ldg(stats + 2 * (size_t)row);
      rstd = __ldg(stats + 2 * (size_t)row + 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dv[i] = __ldg(dy + (size_t)row * FS_F + lane + 32 * i);
        rv[i] = __ldg(r + (size_t)row * FS_F + lane + 32 * i);
      }
    }
    float xh[2], gy[2], c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      xh[i] = (rv[i] - mean) * rstd;
      gy[i] = dv[i] * __ldg(gamma + lane + 32 * i);
      c1 += gy[i];
      c2 += gy[i] * xh[i];
    }
    c1 = warp_sum(c1) * (1.f / (float)FS_F);
    c2 = warp_sum(c2) * (1.f / (float)FS_F);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      const float o = rstd * (gy[i] - c1 - xh[i] * c2);
      drs[rr * FS_F + c] = o;
      if (row < n) dr[(size_t)row * FS_F + c] = o;
      pg[rr * FS_F + c] = dv[i] * xh[i];
      pb[rr * FS_F + c] = dv[i];
    }
  }
  __syncthreads();
  // dgamma / dbeta partial of this CTA: rows summed in order
  if (tid < 2 * FS_F) {
    const float* src = tid < FS_F ? pg + tid : pb + (tid - FS_F);
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < R; ++w) sacc += src[w * FS_F];
    part[(size_t)blockIdx.x * 2 * FS_F + tid] = sacc;
  }
  // (the barrier at the top of the first loop pass separates these reads of pb from the writes to dhs)
  const int cq = tid & 15, kq = (tid >> 4) & 3;
  float accx[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t) accx[u][t] = 0.f;
  for (int gi = 0; gi < 2 * ngrp; ++gi) {
    fs_wait1();
    __syncthreads();
    const float* grp = ring + (gi & 1) * FS_GRP * FS_SLOT_FLOATS;
    if (gi < ngrp) {
      // ---- dhp = (dr W2) * relu'(hdn): thread = hidden units j = 256 gi + 4 jq .. + 3 (slot jq / 16, columns 4 (jq % 16) ..) ----
      const int jq = tid & 63;
      const float* ws = grp + (jq >> 4) * FS_SLOT_FLOATS + 4 * (jq & 15);
      float acc[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[u][t] = 0.f;
#pragma unroll 2
      for (int c4 = 0; c4 < FS_F / 4; ++c4) {
        float4 dv[U], wv[4];                              // wv[e] = W2[4 c4 + e][j .. j + 3]
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = *reinterpret_cast<const float4*>(ws + (4 * c4 + e) * FS_WLD);
#pragma unroll
        for (int u = 0; u < U; ++u) dv[u] = *reinterpret_cast<const float4*>(drs + (U * rq + u) * FS_F + 4 * c4);   // broadcast
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[u][0] = fmaf(dv[u].x, wv[0].x, acc[u][0]); acc[u][1] = fmaf(dv[u].x, wv[0].y, acc[u][1]);
          acc[u][2] = fmaf(dv[u].x, wv[0].z, acc[u][2]); acc[u][3] = fmaf(dv[u].x, wv[0].w, acc[u][3]);
          acc[u][0] = fmaf(dv[u].y, wv[1].x, acc[u][0]); acc[u][1] = fmaf(dv[u].y, wv[1].y, acc[u][1]);
          acc[u][2] = fmaf(dv[u].y, wv[1].z, acc[u][2]); acc[u][3] = fmaf(dv[u].y, wv[1].w, acc[u][3]);
          acc[u][0] = fmaf(dv[u].z, wv[2].x, acc[u][0]); acc[u][1] = fmaf(dv[u].z, wv[2].y, acc[u][1]);
          acc[u][2] = fmaf(dv[u].z, wv[2].z, acc[u][2]); acc[u][3] = fmaf(dv[u].z, wv[2].w, acc[u][3]);
          acc[u][0] = fmaf(dv[u].w, wv[3].x, acc[u][0]); acc[u][1] = fmaf(dv[u].w, wv[3].y, acc[u][1]);
          acc[u][2] = fmaf(dv[u].w, wv[3].z, acc[u][2]); acc[u][3] = fmaf(dv[u].w, wv[3].w, acc[u][3]);
        }
      }
      const int j = gi * FS_GRP * FS_CH + 4 * jq;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = U * rq + u;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + rr < n) {
          const float4 hm = __ldg(reinterpret_cast<const float4*>(hdn + (size_t)(row0 + rr) * Dh + j));
          o = make_float4(hm.x > 0.f ? acc[u][0] : 0.f, hm.y > 0.f ? acc[u][1] : 0.f, hm.z > 0.f ? acc[u][2] : 0.f,
                          hm.w > 0.f ? acc[u][3] : 0.f);
          *reinterpret_cast<float4*>(dhp + (size_t)(row0 + rr) * Dh + j) = o;
        }
        *reinterpret_cast<float4*>(dhs + (size_t)rr * Dh + j) = o;
      }
    } else {
      // ---- dx += dhp W1 over k = 256 (gi - ngrp) + 64 s + 16 kq .. + 15 (slot s: W1 rows of that 64-k slice); thread =
      // output columns 4 cq .. 4 cq + 3 ----
      const int kbase = (gi - ngrp) * FS_GRP * FS_CH;
#pragma unroll
      for (int sl = 0; sl < FS_GRP; ++sl) {
        const float* ws = grp + sl * FS_SLOT_FLOATS + 4 * cq;
#pragma unroll 2
        for (int k4 = 0; k4 < 4; ++k4) {
          const int kl = 16 * kq + 4 * k4;
          float4 dv[U], wv[4];                            // wv[e] = W1[k + e][4 cq .. + 3]
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[e] = *reinterpret_cast<const float4*>(ws + (kl + e) * FS_WLD);
#pragma unroll
          for (int u = 0; u < U; ++u)
            dv[u] = *reinterpret_cast<const float4*>(dhs + (size_t)(U * rq + u) * Dh + kbase + FS_CH * sl + kl);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            accx[u][0] = fmaf(dv[u].x, wv[0].x, accx[u][0]); accx[u][1] = fmaf(dv[u].x, wv[0].y, accx[u][1]);
            accx[u][2] = fmaf(dv[u].x, wv[0].z, accx[u][2]); accx[u][3] = fmaf(dv[u].x, wv[0].w, accx[u][3]);
            accx[u][0] = fmaf(dv[u].y, wv[1].x, accx[u][0]); accx[u][1] = fmaf(dv[u].y, wv[1].y, accx[u][1]);
            accx[u][2] = fmaf(dv[u].y, wv[1].z, accx[u][2]); accx[u][3] = fmaf(dv[u].y, wv[1].w, accx[u][3]);
            accx[u][0] = fmaf(dv[u].z, wv[2].x, accx[u][0]); accx[u][1] = fmaf(dv[u].z, wv[2].y, accx[u][1]);
            accx[u][2] = fmaf(dv[u].z, wv[2].z, accx[u][2]); accx[u][3] = fmaf(dv[u].z, wv[2].w, accx[u][3]);
            accx[u][0] = fmaf(dv[u].w, wv[3].x, accx[u][0]); accx[u][1] = fmaf(dv[u].w, wv[3].y, accx[u][1]);
            accx[u][2] = fmaf(dv[u].w, wv[3].z, accx[u][2]); accx[u][3] = fmaf(dv[u].w, wv[3].w, accx[u][3]);
          }
        }
      }
    }
    __syncthreads();
    fs_issue_group(ring, w1, w2, false, gi + 2, nch, Dh, tid);
  }
  // reduce the four k quarters in order, add dr
  {
    float* ps = ring;                                    // [4 kq][16 rows][64 c]
#pragma unroll
    for (int u = 0; u < U; ++u)
      *reinterpret_cast<float4*>(ps + (kq * R + U * rq + u) * FS_F + 4 * cq) =
          make_float4(accx[u][0], accx[u][1], accx[u][2], accx[u][3]);
    __syncthreads();
    const int rr = tid >> 4, c4 = tid & 15;
    if (rr < R) {
      float4 o = *reinterpret_cast<const float4*>(drs + rr * FS_F + 4 * c4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 pv = *reinterpret_cast<const float4*>(ps + (q * R + rr) * FS_F + 4 * c4);
        o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
      }
      if (row0 + rr < n) *reinterpret_cast<float4*>(dx + (size_t)(row0 + rr) * FS_F + 4 * c4) = o;
    }
  }
}

static int ln_grid(int N) {
  int blocks = ceil_div(N, LN_WARPS);
  if (blocks > LN_MAX_BLOCKS) blocks = LN_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  return blocks;
}

}  // namespace hsg

using namespace hsg;

namespace hsg {
int layernorm_fwd_dropres(int N, int D, float* r, const float* resid, DropCfg dc, const float* gamma, const float* beta,
                          float* y, float* stats, cudaStream_t s) {
  if (N < 0 || D <= 0 || !r || !resid || !gamma || !beta || !y || !stats) return HSG_ERR_ARG;
  if (D % 4 != 0 || D > 512) return HSG_ERR_SHAPE;
  if (!aligned16(r) || !aligned16(y) || !aligned16(resid) || !aligned16(gamma) || !aligned16(beta)) return HSG_ERR_ALIGN;
  if (N == 0) return HSG_OK;
  LaunchScope ls(SLOT_LN_FWD, s);
  const int grid = ln_grid(N);
  switch (ceil_div(D, 128)) {
    case 1: launch_k(layernorm_fwd_kernel<1, true>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, r, resid, dc); break;
    case 2: launch_k(layernorm_fwd_kernel<2, true>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, r, resid, dc); break;
    case 3: launch_k(layernorm_fwd_kernel<3, true>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, r, resid, dc); break;
    default: launch_k(layernorm_fwd_kernel<4, true>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, r, resid, dc); break;
  }
  return check_launch();
}

int ln_partials_reduce(int nseg, int nblocks, size_t seg_stride, int D, const float* part, float* dgamma, float* dbeta,
                       int accumulate, cudaStream_t s) {
  LaunchScope ls(SLOT_LN_BWD_REDUCE, s);
  launch_k(layernorm_bwd_reduce_kernel, dim3(ceil_div(2 * D, 32)), dim3(1024), 0, s, nseg, nblocks, seg_stride, D, part,
           dgamma, dbeta, accumulate);
  return check_launch();
}

int layernorm_bwd_ex(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma, float* dr,
                     float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s,
                     int* defer_blocks) {
  if (N < 0 || D <= 0 || !dy || !r || !stats || !gamma || !dr || (!defer_blocks && (!dgamma || !dbeta)) || !ws)
    return HSG_ERR_ARG;
  if (D % 4 != 0 || D > 512) return HSG_ERR_SHAPE;
  if (ws_bytes < hsg_layernorm_bwd_workspace_bytes(N, D)) return HSG_ERR_WORKSPACE;
  if (!aligned16(dy) || !aligned16(r) || !aligned16(dr) || !aligned16(gamma)) return HSG_ERR_ALIGN;
  const int grid = ln_grid(N);
  const int nv4 = ceil_div(D, 128);
  const size_t smem = (size_t)LN_WARPS * 2 * D * sizeof(float);
  float* part = reinterpret_cast<float*>(ws);
  {
    LaunchScope ls(SLOT_LN_BWD, s);
    switch (nv4) {
      case 1: launch_k(layernorm_bwd_kernel<1>, dim3(grid), dim3(LN_THREADS), smem, s, N, D, dy, r, stats, gamma, dr, part); break;
      case 2: launch_k(layernorm_bwd_kernel<2>, dim3(grid), dim3(LN_THREADS), smem, s, N, D, dy, r, stats, gamma, dr, part); break;
      case 3: launch_k(layernorm_bwd_kernel<3>, dim3(grid), dim3(LN_THREADS), smem, s, N, D, dy, r, stats, gamma, dr, part); break;
      default: launch_k(layernorm_bwd_kernel<4>, dim3(grid), dim3(LN_THREADS), smem, s, N, D, dy, r, stats, gamma, dr, part); break;
    }
    int rc = check_launch();
    if (rc) return rc;
  }
  if (defer_blocks) {                                   // the caller reduces the partials of several launches at once
    *defer_blocks = grid;
    return HSG_OK;
  }
  return ln_partials_reduce(1, grid, 0, D, part, dgamma, dbeta, accumulate, s);
}


// rows per CTA (HSG_FFN_ROWS_PER_CTA = 8 / 16 forces one).  Measured at n = 1 009 (gpurun r02x): alone, 8 rows per CTA
// take 14.6 / 15.9 us forward / backward against 17.4 / 17.4 us - a CTA is bound by the latency of its shared-memory
// reads at two warps per scheduler (ncu: 37 % issue, short-scoreboard stalls), not by its FMA count - but inside the
// 32-graph step the 127 CTAs share SMs with the side-stream products and the step does not move (0.604 against 0.602
// ms); 8 rows are used while they leave half of the SMs to the other streams.
static int fs_rows(int n) {
  static const int forced = [] {
    const char* e = getenv("HSG_FFN_ROWS_PER_CTA");
    const int v = e ? atoi(e) : 0;
    return (v == 8 || v == 16) ? v : 0;
  }();
  if (forced) return forced;
  return ceil_div(n, 8) <= 74 ? 8 : FS_ROWS;
}

// fused small-node-set FFN (see ffn_rows_fwd_kernel)
bool ffn_rows_ok(int n, int F, int d_hid) {
  if (n <= 0 || F != FS_F || d_hid <= 0 || d_hid > FS_MAX_DH || d_hid % (FS_GRP * FS_CH) != 0) return false;
  if (ceil_div(n, fs_rows(n)) > FS_MAX_BLOCKS) return false;        // dgamma / dbeta partials fit the LayerNorm workspace
  return gemm_is_small(n, d_hid, F);                                // same latency-vs-throughput threshold as the GEMMs
}

static bool fs_attr(const void* fn, size_t bytes) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

int ffn_rows_fwd(int n, int F, int d_hid, const float* x, const float* w1, const float* b1, const float* w2,
                 const float* b2, const float* gamma, const float* beta, float* hdn, float* r, float* y, float* stats,
                 cudaStream_t s) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !gamma || !beta || !hdn || !r || !y || !stats) return HSG_ERR_ARG;
  if (!ffn_rows_ok(n, F, d_hid)) return HSG_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(w1) || !aligned16(w2)) return HSG_ERR_ALIGN;
  static bool attr_done = false;
  if (!attr_done) {
    if (!fs_attr((const void*)ffn_rows_fwd_kernel<4>, fs_smem_floats(FS_MAX_DH) * sizeof(float)) ||
        !fs_attr((const void*)ffn_rows_fwd_kernel<2>, fs_smem_floats(FS_MAX_DH) * sizeof(float)))
      return HSG_ERR_CUDA;
    attr_done = true;
  }
  LaunchScope ls(SLOT_FFN_ROWS, s);
  const int rows = fs_rows(n);
  if (rows == 8)
    launch_k(ffn_rows_fwd_kernel<2>, dim3(ceil_div(n, 8)), dim3(FS_THREADS), fs_smem_floats(d_hid) * sizeof(float), s, n,
             d_hid, x, w1, b1, w2, b2, gamma, beta, hdn, r, y, stats);
  else
    launch_k(ffn_rows_fwd_kernel<4>, dim3(ceil_div(n, FS_ROWS)), dim3(FS_THREADS), fs_smem_floats(d_hid) * sizeof(float), s,
             n, d_hid, x, w1, b1, w2, b2, gamma, beta, hdn, r, y, stats);
  return check_launch();
}

int ffn_rows_bwd(int n, int F, int d_hid, const float* dy, const float* r, const float* stats, const float* gamma,
                 const float* hdn, const float* w1, const float* w2, float* dr, float* dhp, float* dx, float* dgamma,
                 float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s, int* defer_blocks) {
  if (!dy || !r || !stats || !gamma || !hdn || !w1 || !w2 || !dr || !dhp || !dx ||
      (!defer_blocks && (!dgamma || !dbeta)) || !ws)
    return HSG_ERR_ARG;
  if (!ffn_rows_ok(n, F, d_hid)) return HSG_ERR_SHAPE;
  if (ws_bytes < hsg_layernorm_bwd_workspace_bytes(n, F)) return HSG_ERR_WORKSPACE;
  if (!aligned16(w1) || !aligned16(w2)) return HSG_ERR_ALIGN;
  static bool attr_done = false;
  if (!attr_done) {
    if (!fs_attr((const void*)ffn_rows_bwd_kernel<4>, fs_smem_floats(FS_MAX_DH) * sizeof(float)) ||
        !fs_attr((const void*)ffn_rows_bwd_kernel<2>, fs_smem_floats(FS_MAX_DH) * sizeof(float)))
      return HSG_ERR_CUDA;
    attr_done = true;
  }
  float* part = reinterpret_cast<float*>(ws);
  const int rows = fs_rows(n);
  const int nblocks = ceil_div(n, rows);
  {
    LaunchScope ls(SLOT_FFN_ROWS, s);
    if (rows == 8)
      launch_k(ffn_rows_bwd_kernel<2>, dim3(nblocks), dim3(FS_THREADS), fs_smem_floats(d_hid) * sizeof(float), s, n, d_hid,
               dy, r, stats, gamma, hdn, w1, w2, dr, dhp, dx, part);
    else
      launch_k(ffn_rows_bwd_kernel<4>, dim3(nblocks), dim3(FS_THREADS), fs_smem_floats(d_hid) * sizeof(float), s, n, d_hid,
               dy, r, stats, gamma, hdn, w1, w2, dr, dhp, dx, part);
    int rc = check_launch();
    if (rc) return rc;
  }
  if (defer_blocks) {
    *defer_blocks = nblocks;
    return HSG_OK;
  }
  return ln_partials_reduce(1, nblocks, 0, F, part, dgamma, dbeta, accumulate, s);
}

}  // namespace hsg

extern "C" {

int hsg_layernorm_fwd(int N, int D, const float* r, const float* gamma, const float* beta, float* y, float* stats,
                      void* stream) {
  if (N < 0 || D <= 0 || !r || !gamma || !beta || !y || !stats) return HSG_ERR_ARG;
  if (D % 4 != 0 || D > 512) return HSG_ERR_SHAPE;
  if (!aligned16(r) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta)) return HSG_ERR_ALIGN;
  if (N == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_LN_FWD, s);
  const int grid = ln_grid(N);
  const int nv4 = ceil_div(D, 128);
  switch (nv4) {
    case 1: launch_k(layernorm_fwd_kernel<1, false>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, nullptr, nullptr, DropCfg{0, 0, 1.f, nullptr}); break;
    case 2: launch_k(layernorm_fwd_kernel<2, false>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, nullptr, nullptr, DropCfg{0, 0, 1.f, nullptr}); break;
    case 3: launch_k(layernorm_fwd_kernel<3, false>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, nullptr, nullptr, DropCfg{0, 0, 1.f, nullptr}); break;
    default: launch_k(layernorm_fwd_kernel<4, false>, dim3(grid), dim3(LN_THREADS), 0, s, N, D, r, gamma, beta, y, stats, nullptr, nullptr, DropCfg{0, 0, 1.f, nullptr}); break;
  }
  return check_launch();
}

int hsg_ffn_rows_ok(int n, int F, int d_hid) { return ffn_rows_ok(n, F, d_hid) ? 1 : 0; }

int hsg_ffn_rows_fwd(int n, int F, int d_hid, const float* x, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* gamma, const float* beta, float* hdn, float* r, float* y,
                     float* stats, void* stream) {
  return ffn_rows_fwd(n, F, d_hid, x, w1, b1, w2, b2, gamma, beta, hdn, r, y, stats, (cudaStream_t)stream);
}

int hsg_ffn_rows_bwd(int n, int F, int d_hid, const float* dy, const float* r, const float* stats, const float* gamma,
                     const float* hdn, const float* w1, const float* w2, float* dr, float* dhp, float* dx,
                     float* dgamma, float* dbeta, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  return ffn_rows_bwd(n, F, d_hid, dy, r, stats, gamma, hdn, w1, w2, dr, dhp, dx, dgamma, dbeta, ws, ws_bytes,
                      accumulate, (cudaStream_t)stream);
}

size_t hsg_layernorm_bwd_workspace_bytes(int N, int D) {
  (void)N;
  return (size_t)LN_MAX_BLOCKS * 2 * (D > 0 ? D : 1) * sizeof(float) + 16;
}

int hsg_layernorm_bwd(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma,
                      float* dr, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  return layernorm_bwd_ex(N, D, dy, r, stats, gamma, dr, dgamma, dbeta, ws, ws_bytes, 0, (cudaStream_t)stream);
}

}  // extern "C"
