// Compile-time lane mapping of the edge kernels (hsg_edge.cu, hsg_edge_seg.cu) and the small device helpers they
// share.  See hsg_edge_layout.cuh for the lane-interleaved row layout of the gathered tensors.
#pragma once
#include <math_constants.h>

#include "hsg_common.cuh"
#include "hsg_edge_layout.cuh"

namespace hsg {

template <int H_, int D_>
struct EdgeCfg {
  static constexpr int H = H_, D = D_, F = H_ * D_;
  static constexpr int VEC = (D_ % 4 == 0) ? 4 : ((D_ % 2 == 0) ? 2 : 1);
  static constexpr int NV = D_ / VEC;                       // vectors per head
  static constexpr int LPH = edge_lph(H_, NV);              // lanes per head (hsg_edge_layout.cuh)
  static constexpr int VPL = (NV + LPH - 1) / LPH;          // vectors per lane
  static constexpr int GROUP = H_ * LPH;                    // lanes per edge row
  static constexpr int EPS = 32 / GROUP;                    // edge rows per warp step
  static constexpr int NE = VPL * VEC;                      // elements per lane
  static constexpr int FP = VPL * GROUP * VEC;              // permuted row width
  // a lane's elements are original columns k*D + VEC*(l + LPH*i) + t: with LPH == 1 (lane owns its whole head) or
  // VPL == 1 every lane vector is a contiguous piece of the row; otherwise row I/O is staged through shared memory
  static constexpr bool STAGED = VPL > 1 && LPH > 1;
  static_assert(H_ <= 32 && LPH >= 1 && EPS >= 1, "bad edge config");
  static_assert(!STAGED || (EPS == 1 && F % 4 == 0), "staged epilogue assumes one row per warp step");
};

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float* out) {
  if (VEC == 4) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if (VEC == 2) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    out[0] = v.x; out[1] = v.y;
  } else {
    out[0] = __ldg(p);
  }
}

template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float* v) {
  if (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    p[0] = v[0];
  }
}

// sum over the LPH lanes that own one head (lanes [base, base+LPH) of the warp)
template <int LPH>
__device__ __forceinline__ float head_sum(float v, int lane, int l) {
  if ((LPH & (LPH - 1)) == 0) {
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  } else {
    const int base = lane - l;
    float s = v;
#pragma unroll
    for (int o = 1; o < LPH; ++o) {
      int src = base + ((l + o) % LPH);
      s += __shfl_sync(0xffffffffu, v, src & 31);
    }
    return s;
  }
}

constexpr int EDGE_WARPS = 8;
constexpr int EDGE_THREADS = EDGE_WARPS * 32;

// ---------------------------------------------------------------------------
// forward.  U = edge rows gathered back-to-back per group before any is consumed (memory-level parallelism):
// large for high-degree destinations (supernodes), small for low-degree ones (words) where it only costs registers.
// ---------------------------------------------------------------------------
// exp(x) as ONE multiply + MUFU.EX2 (flush-to-zero).  __expf is the same approximation plus a range fix-up for
// results below 2^-126 that none of the edge kernels' arguments needs - three more instructions per call, which the
// per-ELEMENT calls of the wide S2W rows (issue-bound kernels) pay ten times per lane and row.
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

__device__ __forceinline__ float elu1(float o) {            // F.elu, alpha = 1 (GAT.py:56), branch-free
  const float e = exp_fast(fminf(o, 0.f)) - 1.f;
  return o > 0.f ? o : e;
}

}  // namespace hsg
