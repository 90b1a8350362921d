// K3 / K5: fused WSWGAT edge kernels (forward and backward) over a CSC.
//
// Replaces, for ALL heads at once, what the reference executes per head through
// DGL's UDF runtime (module/GATLayer.py:88-102 / 127-140 driven by apply_edges
// :112/:148 and the degree-bucketed pull :113/:149):
//     e_uv  = leaky_relu(a_src . z_u + a_dst . 0 + a_feat . feat_fc(T[bin_uv]))  = leaky_relu(p_u + q[bin])
//     alpha = softmax over ALL in-edges of v (extra[v] never-written edges: e = 0, z = 0)
//     sh_v  = sum alpha z_u ;   x_v = elu(sh_v) + origin_v          (GAT.py:56-57)
//
// Mapping (compile-time per (H, D), see hsg_edge_layout.cuh): one warp walks the in-edge list of
// one destination row.  A row is spread over GROUP = H*LPH lanes so that EVERY LANE OWNS ELEMENTS
// OF EXACTLY ONE HEAD: the attention logit, the online-softmax state (m, den) and the weight alpha
// are per-lane scalars, the forward edge loop has no cross-lane reduction, and the per-edge dot
// product of the backward pass is an LPH-lane reduction.  EPS = 32/GROUP edge rows per warp step.
// Memory behaviour:
//   * gathered tensors (zp, g) are stored lane-interleaved, so each gather instruction of a warp
//     reads one contiguous slab (128-bit per lane, 64-bit when D % 4 != 0);
//   * the neighbour ids / bins of up to 32 edges are fetched with ONE coalesced load and
//     broadcast by shuffle; U edge rows per group are gathered back-to-back before any of them
//     is consumed (memory-level parallelism), and the next row's indptr is prefetched;
//   * destination-side streams (origin, sh, x, dx) use 128-bit coalesced accesses, staged through
//     shared memory when the lane layout is not contiguous (VPL > 1).
// Everything is deterministic: no floating-point atomics anywhere.
#include <math_constants.h>

#include <atomic>
#include <cstdlib>

#include "hsg_common.cuh"
#include "hsg_internal.cuh"
#include "hsg_edge_layout.cuh"
#include "hsg_edge_cfg.cuh"

namespace hsg {


template <int H, int D, int U>
__global__ void __launch_bounds__(EDGE_THREADS, (EdgeCfg<H, D>::NE <= 8) ? 4 : 3)
edge_fwd_kernel(int n_dst, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                const uint8_t* __restrict__ bin, const int32_t* __restrict__ extra, const float* __restrict__ zp,
                int ldz, const float* __restrict__ q, const float* __restrict__ origin, float* __restrict__ sh,
                float* __restrict__ x, float* __restrict__ stat) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  constexpr int OV = C::STAGED ? (C::F / 4 + 31) / 32 : 1;   // origin float4 per lane (staged epilogue)
  __shared__ float q_s[HSG_N_BINS * H];
  __shared__ __align__(16) float stage[C::STAGED ? EDGE_WARPS * C::F : 4];
  for (int i = threadIdx.x; i < HSG_N_BINS * H; i += blockDim.x) q_s[i] = q[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;       // head owned by this lane
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;

  // software pipeline over rows: indptr two rows ahead, first neighbour chunk one row ahead
  int v = warp;
  int beg = 0, end = 0, begn = 0, endn = 0, u0 = 0, b0 = 0;
  if (v < n_dst) {
    beg = __ldg(indptr + v);
    end = __ldg(indptr + v + 1);
    if (beg + lane < end) {
      u0 = __ldg(nbr + beg + lane);
      b0 = __ldg(bin + beg + lane);
    }
  }
  if (v + nwarps < n_dst) {
    begn = __ldg(indptr + v + nwarps);
    endn = __ldg(indptr + v + nwarps + 1);
  }
  while (v < n_dst) {
    const int v2 = v + 2 * nwarps;
    int beg2 = 0, end2 = 0, u0n = 0, b0n = 0;
    if (v2 < n_dst) {
      beg2 = __ldg(indptr + v2);
      end2 = __ldg(indptr + v2 + 1);
    }
    if (begn + lane < endn) {                           // next row's first chunk (one coalesced fetch)
      u0n = __ldg(nbr + begn + lane);
      b0n = __ldg(bin + begn + lane);
    }
    const float xcnt = extra ? (float)__ldg(extra + v) : 0.f;
    // origin row of this destination: issued now, consumed in the epilogue
    float4 og[OV];
    float ogd[C::STAGED ? 1 : C::NE];
    if (x != nullptr) {
      if (!C::STAGED) {
#pragma unroll
        for (int i = 0; i < C::NE; ++i) ogd[i] = 0.f;
        if (grp == 0) {
#pragma unroll
          for (int i = 0; i < C::VPL; ++i)
            if (l + C::LPH * i < C::NV)
              ld_vec<C::VEC>(origin + (size_t)v * C::F + k * D + C::VEC * (l + C::LPH * i), ogd + i * C::VEC);
        }
      } else {
#pragma unroll
        for (int i = 0; i < OV; ++i) {
          const int c4 = lane + 32 * i;
          og[i] = c4 < C::F / 4 ? __ldg(reinterpret_cast<const float4*>(origin + (size_t)v * C::F) + c4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    float m = -CUDART_INF_F, den = 0.f;
    float acc[C::NE];
#pragma unroll
    for (int i = 0; i < C::NE; ++i) acc[i] = 0.f;

    for (int c0 = beg; c0 < end; c0 += 32) {
      const int cnt = min(32, end - c0);
      int my_u = u0, my_b = b0;
      if (c0 != beg) {
        my_u = 0;
        my_b = 0;
        if (lane < cnt) {
          my_u = __ldg(nbr + c0 + lane);
          my_b = __ldg(bin + c0 + lane);
        }
      }
      for (int j0 = 0; j0 < cnt; j0 += C::EPS * U) {     // warp-uniform trip count
        float zv[U][C::NE], pe[U];
        int bb[U];
        bool ok[U];
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {                // issue all gathers first
          const int j = j0 + uu * C::EPS + grp;
          const int u = __shfl_sync(0xffffffffu, my_u, j & 31);
          bb[uu] = __shfl_sync(0xffffffffu, my_b, j & 31);
          ok[uu] = lane_on && j < cnt;
          pe[uu] = 0.f;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) zv[uu][i] = 0.f;
          if (ok[uu]) {
            const float* row = zp + (size_t)u * ldz;
            pe[uu] = __ldg(row + C::FP + k);
#pragma unroll
            for (int i = 0; i < C::VPL; ++i)
              if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(row + (i * C::GROUP + gl) * C::VEC, zv[uu] + i * C::VEC);
          }
        }
        // consume them: online softmax with ONE rescale per U rows (maximum first), branch-free
        float lgv[U];
        float mx = m;
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          lgv[uu] = ok[uu] ? leaky(pe[uu] + q_s[bb[uu] * H + k]) : -CUDART_INF_F;
          mx = fmaxf(mx, lgv[uu]);
        }
        const float sc = (m == mx) ? 1.f : exp_fast(m - mx);   // m == mx also covers -inf == -inf (no row seen yet)
        den *= sc;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) acc[i] *= sc;
        m = mx;
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          const float w = ok[uu] ? exp_fast(lgv[uu] - mx) : 0.f;
          den += w;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(w, zv[uu][i], acc[i]);
        }
      }
    }
    // merge the EPS partial states into group 0 (fixed order)
#pragma unroll
    for (int g2 = 1; g2 < C::EPS; ++g2) {
      const int src = (gl + g2 * C::GROUP) & 31;
      const float m2 = __shfl_sync(0xffffffffu, m, src);
      const float d2 = __shfl_sync(0xffffffffu, den, src);
      const float mn = fmaxf(m, m2);
      const float s1 = (m == -CUDART_INF_F) ? 0.f : exp_fast(m - mn);
      const float s2 = (m2 == -CUDART_INF_F) ? 0.f : exp_fast(m2 - mn);
      // only group 0 accumulates: the other groups must keep their own partial state for later reads
      if (grp == 0) den = den * s1 + d2 * s2;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) {
        const float a2 = __shfl_sync(0xffffffffu, acc[i], src);
        if (grp == 0) acc[i] = acc[i] * s1 + a2 * s2;
      }
      if (grp == 0) m = mn;
    }
    float mf = 0.f, inv = 0.f;
    if (m == -CUDART_INF_F) {  // no word<->supernode in-edge: DGL's zero fill (or softmax over z = 0 extras)
      den = xcnt > 0.f ? xcnt : 1.f;
    } else {
      mf = xcnt > 0.f ? fmaxf(m, 0.f) : m;
      const float sc = exp_fast(m - mf);
      den = den * sc + xcnt * exp_fast(-mf);
      inv = sc / den;
    }
    if (grp == 0 && l == 0) {
      stat[(size_t)v * 3 * H + k] = mf;
      stat[(size_t)v * 3 * H + H + k] = den;
    }
    if (!C::STAGED) {
      if (grp == 0) {                                   // every lane vector is a contiguous piece of the row
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            const size_t off = (size_t)v * C::F + k * D + C::VEC * (l + C::LPH * i);
            float o[C::VEC];
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) o[t] = acc[i * C::VEC + t] * inv;
            if (sh != nullptr) st_vec<C::VEC>(sh + off, o);
            if (x != nullptr) {
              float xo[C::VEC];
#pragma unroll
              for (int t = 0; t < C::VEC; ++t) xo[t] = ogd[i * C::VEC + t] + elu1(o[t]);
              st_vec<C::VEC>(x + off, xo);
            }
          }
        }
      }
    } else {
      float* st_row = stage + wib * C::F;
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            float o[C::VEC];
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) o[t] = acc[i * C::VEC + t] * inv;
            st_vec<C::VEC>(st_row + k * D + C::VEC * (l + C::LPH * i), o);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < OV; ++i) {                     // coalesced 128-bit row stores
        const int c4 = lane + 32 * i;
        if (c4 < C::F / 4) {
          const float4 o = *reinterpret_cast<const float4*>(st_row + 4 * c4);
          const size_t off = (size_t)v * C::F + 4 * c4;
          if (sh != nullptr) *reinterpret_cast<float4*>(sh + off) = o;
          if (x != nullptr)
            *reinterpret_cast<float4*>(x + off) =
                make_float4(og[i].x + elu1(o.x), og[i].y + elu1(o.y), og[i].z + elu1(o.z), og[i].w + elu1(o.w));
        }
      }
      __syncwarp();
    }
    v += nwarps;
    beg = begn;
    end = endn;
    begn = beg2;
    endn = end2;
    u0 = u0n;
    b0 = b0n;
  }
}

// ---------------------------------------------------------------------------
// forward over WIDE rows of LOW degree (the word rows of the S2W layer: 300 floats, 1-3 in-edges; layouts with one
// lane group per warp and the staged epilogue).  ncu on the 2 048-graph shard showed edge_fwd_kernel<6,50,1> neither
// bandwidth- nor occupancy-bound there: half of the samples wait on the ONE source-row gather a row issues and
// consumes in the same iteration, and a row costs 390 instructions (chunk loops, id shuffles, the online-softmax
// rescale).  Here the software pipeline is one stage deeper - row pointers three rows ahead, the ids of the first two
// in-edges two rows ahead (broadcast loads, no shuffles), their source rows [z | p] and the destination's origin row
// ONE ROW AHEAD in registers - and rows of at most two in-edges take a straight-line path (both logits, one maximum,
// no rescale); further edges of a row are folded in by the online update one at a time.  Same results as
// edge_fwd_kernel up to the summation order inside a row.
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS, 2)
edge_fwd_lowdeg_kernel(int n_dst, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                       const uint8_t* __restrict__ bin, const int32_t* __restrict__ extra,
                       const float* __restrict__ zp, int ldz, const float* __restrict__ q,
                       const float* __restrict__ origin, float* __restrict__ sh, float* __restrict__ x,
                       float* __restrict__ stat) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  static_assert(C::EPS == 1 && C::STAGED, "one lane group per warp, staged row I/O");
  constexpr int OV = (C::F / 4 + 31) / 32;                 // float4 per lane of a raw row
  constexpr int NE = C::NE;
  __shared__ float q_s[HSG_N_BINS * H];
  __shared__ __align__(16) float stage[EDGE_WARPS * C::F];
  for (int i = threadIdx.x; i < HSG_N_BINS * H; i += blockDim.x) q_s[i] = q[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;       // head owned by this lane
  const int l = gl % C::LPH;
  const bool lane_on = lane < C::GROUP;
  float* st_row = stage + wib * C::F;

  auto load_ip = [&](int vr, int& b_, int& e_) {
    b_ = 0;
    e_ = 0;
    if (vr < n_dst) {
      b_ = __ldg(indptr + vr);
      e_ = __ldg(indptr + vr + 1);
    }
  };
  auto load_ids = [&](int b_, int e_, int* us, int* bs) {   // first two in-edges: every lane reads the same words
    us[0] = us[1] = 0;
    bs[0] = bs[1] = 0;
    if (b_ < e_) {
      us[0] = __ldg(nbr + b_);
      bs[0] = __ldg(bin + b_);
    }
    if (b_ + 1 < e_) {
      us[1] = __ldg(nbr + b_ + 1);
      bs[1] = __ldg(bin + b_ + 1);
    }
  };
  auto gather = [&](int u, float* zz, float& pp, bool on) { // [z | p] of source row u in the lane layout
    pp = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) zz[i] = 0.f;
    if (lane_on && on) {
      const float* row = zp + (size_t)u * ldz;
      pp = __ldg(row + C::FP + k);
#pragma unroll
      for (int i = 0; i < C::VPL; ++i)
        if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(row + (i * C::GROUP + gl) * C::VEC, zz + i * C::VEC);
    }
  };
  auto load_origin = [&](int vr, float4* og) {
#pragma unroll
    for (int i = 0; i < OV; ++i) {
      const int c4 = lane + 32 * i;
      og[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x != nullptr && vr < n_dst && c4 < C::F / 4)
        og[i] = __ldg(reinterpret_cast<const float4*>(origin + (size_t)vr * C::F) + c4);
    }
  };

  int v = warp;
  int beg, end, begn, endn, beg2, end2;
  int us[2], bs[2], usn[2], bsn[2];
  float z0[NE], z1[NE], p0, p1;
  float4 og[OV];
  load_ip(v, beg, end);
  load_ip(v + nwarps, begn, endn);
  load_ip(v + 2 * nwarps, beg2, end2);
  load_ids(beg, end, us, bs);
  load_ids(begn, endn, usn, bsn);
  load_origin(v, og);
  gather(us[0], z0, p0, end > beg);
  gather(us[1], z1, p1, end > beg + 1);
  while (v < n_dst) {
    int beg3, end3, us2[2], bs2[2];
    load_ip(v + 3 * nwarps, beg3, end3);
    load_ids(beg2, end2, us2, bs2);
    const float xcnt = extra ? (float)__ldg(extra + v) : 0.f;
    const int deg = end - beg;
    // ---- softmax over the in-edges: the first two from the prefetched rows, straight line ----
    float m = -CUDART_INF_F, den = 0.f;
    float acc[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = 0.f;
    if (deg > 0) {
      const float lg0 = leaky(p0 + q_s[bs[0] * H + k]);
      const float lg1 = deg > 1 ? leaky(p1 + q_s[bs[1] * H + k]) : -CUDART_INF_F;
      m = fmaxf(lg0, lg1);
      const float w0 = exp_fast(lg0 - m), w1 = exp_fast(lg1 - m);      // exp(-inf) = 0 for a missing second edge
      den = w0 + w1;
#pragma unroll
      for (int i = 0; i < NE; ++i) acc[i] = fmaf(w1, z1[i], w0 * z0[i]);
      for (int e = beg + 2; e < end; ++e) {                // further in-edges (rare on word rows): online update
        float zt[NE], pt;
        gather(__ldg(nbr + e), zt, pt, true);
        const float lg = leaky(pt + q_s[(int)__ldg(bin + e) * H + k]);
        const float mx = fmaxf(m, lg);
        const float sc = exp_fast(m - mx), w = exp_fast(lg - mx);
        den = den * sc + w;
#pragma unroll
        for (int i = 0; i < NE; ++i) acc[i] = fmaf(w, zt[i], acc[i] * sc);
        m = mx;
      }
    }
    // the next row's source rows and origin start their way now (its ids were fetched one iteration ago)
    float4 ogc[OV];
#pragma unroll
    for (int i = 0; i < OV; ++i) ogc[i] = og[i];
    gather(usn[0], z0, p0, endn > begn);
    gather(usn[1], z1, p1, endn > begn + 1);
    load_origin(v + nwarps, og);

    float mf = 0.f, inv = 0.f;
    if (m == -CUDART_INF_F) {  // no word<->supernode in-edge: DGL's zero fill (or softmax over z = 0 extras)
      den = xcnt > 0.f ? xcnt : 1.f;
    } else {
      mf = xcnt > 0.f ? fmaxf(m, 0.f) : m;
      const float sc = exp_fast(m - mf);
      den = den * sc + xcnt * exp_fast(-mf);
      inv = sc / den;
    }
    if (lane_on) {
      if (l == 0) {
        stat[(size_t)v * 3 * H + k] = mf;
        stat[(size_t)v * 3 * H + H + k] = den;
      }
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i < C::NV) {
          float o[C::VEC];
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) o[t] = acc[i * C::VEC + t] * inv;
          st_vec<C::VEC>(st_row + k * D + C::VEC * (l + C::LPH * i), o);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < OV; ++i) {                         // coalesced 128-bit row stores
      const int c4 = lane + 32 * i;
      if (c4 < C::F / 4) {
        const float4 o = *reinterpret_cast<const float4*>(st_row + 4 * c4);
        const size_t off = (size_t)v * C::F + 4 * c4;
        if (sh != nullptr) *reinterpret_cast<float4*>(sh + off) = o;
        if (x != nullptr)
          *reinterpret_cast<float4*>(x + off) = make_float4(ogc[i].x + elu1(o.x), ogc[i].y + elu1(o.y),
                                                            ogc[i].z + elu1(o.z), ogc[i].w + elu1(o.w));
      }
    }
    __syncwarp();
    v += nwarps;
    beg = begn;
    end = endn;
    begn = beg2;
    endn = end2;
    beg2 = beg3;
    end2 = end3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      us[i] = usn[i];
      bs[i] = bsn[i];
      usn[i] = us2[i];
      bsn[i] = bs2[i];
    }
  }
}

// ---------------------------------------------------------------------------
// forward, ROW-PARALLEL: every GROUP-lane group of a warp walks ITS OWN destination row (EPS rows per warp in flight,
// U gathered rows each).  No merge of partial softmax states, no index shuffles: on the W2S default (8,8) (one lane
// per head, four rows per warp) this executes ~2.5x fewer instructions per edge than the shared-row mapping, which
// ncu showed issue-bound on large shards.  Used when there are enough rows to fill the machine (hsg_edge_fwd picks).
// Same arithmetic per row as edge_fwd_kernel with one group (edges in CSC order, one rescale per U rows).
// ---------------------------------------------------------------------------
#ifndef HSG_ROWPAR_4CTA_MAX
#define HSG_ROWPAR_4CTA_MAX 16
#endif
template <int H, int D, int U>
__global__ void __launch_bounds__(EDGE_THREADS, (U * EdgeCfg<H, D>::NE <= HSG_ROWPAR_4CTA_MAX) ? 4 : 3)
edge_fwd_rowpar_kernel(int n_dst, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                       const uint8_t* __restrict__ bin, const int32_t* __restrict__ extra,
                       const float* __restrict__ zp, int ldz, const float* __restrict__ q,
                       const float* __restrict__ origin, float* __restrict__ sh, float* __restrict__ x,
                       float* __restrict__ stat) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  static_assert(!C::STAGED && C::EPS > 1, "row-parallel forward needs direct row I/O and several groups per warp");
  __shared__ float q_s[HSG_N_BINS * H];
  for (int i = threadIdx.x; i < HSG_N_BINS * H; i += blockDim.x) q_s[i] = q[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;
  const int nsteps = ceil_div(n_dst, C::EPS);

  int st = warp;
  int beg = 0, end = 0;
  {
    const int v = st * C::EPS + grp;
    if (lane_on && st < nsteps && v < n_dst) {
      beg = __ldg(indptr + v);
      end = __ldg(indptr + v + 1);
    }
  }
  while (st < nsteps) {
    const int v = st * C::EPS + grp;
    const bool row_on = lane_on && v < n_dst;
    int begn = 0, endn = 0;                              // next step's edge range: issued now, used at the bottom
    {
      const int vn = (st + nwarps) * C::EPS + grp;
      if (lane_on && st + nwarps < nsteps && vn < n_dst) {
        begn = __ldg(indptr + vn);
        endn = __ldg(indptr + vn + 1);
      }
    }
    const int deg = row_on ? end - beg : 0;
    int maxdeg = deg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    // ids / bins of the first U edges of this group's row
    int un[U], bn[U];
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      un[uu] = 0;
      bn[uu] = 0;
      if (uu < deg) {
        un[uu] = __ldg(nbr + beg + uu);
        bn[uu] = __ldg(bin + beg + uu);
      }
    }
    const float xcnt = (extra && row_on) ? (float)__ldg(extra + v) : 0.f;
    float ogd[C::NE];
#pragma unroll
    for (int i = 0; i < C::NE; ++i) ogd[i] = 0.f;
    if (x != nullptr && row_on) {
#pragma unroll
      for (int i = 0; i < C::VPL; ++i)
        if (l + C::LPH * i < C::NV)
          ld_vec<C::VEC>(origin + (size_t)v * C::F + k * D + C::VEC * (l + C::LPH * i), ogd + i * C::VEC);
    }
    float m = -CUDART_INF_F, den = 0.f;
    float acc[C::NE];
#pragma unroll
    for (int i = 0; i < C::NE; ++i) acc[i] = 0.f;

    for (int j0 = 0; j0 < maxdeg; j0 += U) {             // warp-uniform trip count
      float zv[U][C::NE], pe[U];
      int bb[U];
      bool ok[U];
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {                  // gathers of this round
        ok[uu] = j0 + uu < deg;
        bb[uu] = bn[uu];
        pe[uu] = 0.f;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) zv[uu][i] = 0.f;
        if (ok[uu]) {
          const float* row = zp + (size_t)un[uu] * ldz;
          pe[uu] = __ldg(row + C::FP + k);
#pragma unroll
          for (int i = 0; i < C::VPL; ++i)
            if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(row + (i * C::GROUP + gl) * C::VEC, zv[uu] + i * C::VEC);
        }
      }
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {                  // ids of the next round, in flight with the gathers
        const int j = j0 + U + uu;
        un[uu] = 0;
        bn[uu] = 0;
        if (j < deg) {
          un[uu] = __ldg(nbr + beg + j);
          bn[uu] = __ldg(bin + beg + j);
        }
      }
      float lgv[U];
      float mx = m;
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        lgv[uu] = ok[uu] ? leaky(pe[uu] + q_s[bb[uu] * H + k]) : -CUDART_INF_F;
        mx = fmaxf(mx, lgv[uu]);
      }
      const float sc = (m == mx) ? 1.f : exp_fast(m - mx);
      den *= sc;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) acc[i] *= sc;
      m = mx;
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        const float w = ok[uu] ? exp_fast(lgv[uu] - mx) : 0.f;
        den += w;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(w, zv[uu][i], acc[i]);
      }
    }
    float mf = 0.f, inv = 0.f;
    if (m == -CUDART_INF_F) {  // no word<->supernode in-edge: DGL's zero fill (or softmax over z = 0 extras)
      den = xcnt > 0.f ? xcnt : 1.f;
    } else {
      mf = xcnt > 0.f ? fmaxf(m, 0.f) : m;
      const float sc = exp_fast(m - mf);
      den = den * sc + xcnt * exp_fast(-mf);
      inv = sc / den;
    }
    if (row_on) {
      if (l == 0) {
        stat[(size_t)v * 3 * H + k] = mf;
        stat[(size_t)v * 3 * H + H + k] = den;
      }
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i < C::NV) {
          const size_t off = (size_t)v * C::F + k * D + C::VEC * (l + C::LPH * i);
          float o[C::VEC];
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) o[t] = acc[i * C::VEC + t] * inv;
          if (sh != nullptr) st_vec<C::VEC>(sh + off, o);
          if (x != nullptr) {
            float xo[C::VEC];
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) xo[t] = ogd[i * C::VEC + t] + elu1(o[t]);
            st_vec<C::VEC>(x + off, xo);
          }
        }
      }
    }
    st += nwarps;
    beg = begn;
    end = endn;
  }
}

// ---------------------------------------------------------------------------
// backward prep: g = dx * elu'(sh)  (or g = dsh) written lane-interleaved [n_dst, FP],
//                s[v,k] = g_v[k] . sh_v[k]
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS)
edge_bwd_prep_kernel(int n_dst, const float* __restrict__ dx, const float* __restrict__ dsh,
                     const float* __restrict__ sh, float* __restrict__ g, float* __restrict__ stat) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  __shared__ __align__(16) float stage[C::STAGED ? 2 * EDGE_WARPS * C::F : 4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const int nsteps = ceil_div(n_dst, C::EPS);
  // staged layouts: the row of the NEXT grid-stride step is fetched into registers while the current one goes through
  // shared memory (ncu on the 2 048-graph shard: DRAM 55 % of peak, 56 % of the warps active - a warp had loads in
  // flight for only part of its per-row load -> stage -> reduce -> store cycle)
  constexpr int NV4 = C::STAGED ? (C::F / 4 + 31) / 32 : 1;
  float4 ps[NV4], pg[NV4];
  const float* gsrc = dx != nullptr ? dx : dsh;
  auto prefetch_row = [&](int vr) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c4 = lane + 32 * i;
      if (vr < n_dst && c4 < C::F / 4) {
        const size_t off = (size_t)vr * C::F + 4 * c4;
        ps[i] = __ldg(reinterpret_cast<const float4*>(sh + off));
        pg[i] = __ldg(reinterpret_cast<const float4*>(gsrc + off));
      }
    }
  };
  if (C::STAGED) prefetch_row(warp);
  for (int st = warp; st < nsteps; st += nwarps) {
    const int v = st * C::EPS + grp;
    const bool on = grp < C::EPS && v < n_dst;
    float part = 0.f;
    float gv[C::NE];
#pragma unroll
    for (int i = 0; i < C::NE; ++i) gv[i] = 0.f;
    if (!C::STAGED) {
      if (on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            const size_t off = (size_t)v * C::F + k * D + C::VEC * (l + C::LPH * i);
            float s_[C::VEC];
            float* gi = gv + i * C::VEC;
            ld_vec<C::VEC>(sh + off, s_);
            if (dx != nullptr) {
              ld_vec<C::VEC>(dx + off, gi);
#pragma unroll
              for (int t = 0; t < C::VEC; ++t) gi[t] *= (s_[t] > 0.f ? 1.f : exp_fast(s_[t]));
            } else {
              ld_vec<C::VEC>(dsh + off, gi);
            }
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) part = fmaf(gi[t], s_[t], part);
          }
        }
      }
    } else {
      float* g_row = stage + (2 * wib) * C::F;
      float* s_row = stage + (2 * wib + 1) * C::F;
      const int vr = st;                                 // EPS == 1: the whole warp streams row `st`, 128-bit
      if (vr < n_dst) {
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          const int c4 = lane + 32 * i;
          if (c4 < C::F / 4) {
            const float4 s4 = ps[i];
            float4 g4 = pg[i];
            if (dx != nullptr) {
              g4.x *= (s4.x > 0.f ? 1.f : exp_fast(s4.x));
              g4.y *= (s4.y > 0.f ? 1.f : exp_fast(s4.y));
              g4.z *= (s4.z > 0.f ? 1.f : exp_fast(s4.z));
              g4.w *= (s4.w > 0.f ? 1.f : exp_fast(s4.w));
            }
            *reinterpret_cast<float4*>(g_row + 4 * c4) = g4;
            *reinterpret_cast<float4*>(s_row + 4 * c4) = s4;
          }
        }
      }
      prefetch_row(st + nwarps);                       // lands while this row is reduced and stored
      __syncwarp();
      if (on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            const int col = k * D + C::VEC * (l + C::LPH * i);
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) {
              gv[i * C::VEC + t] = g_row[col + t];
              part = fmaf(g_row[col + t], s_row[col + t], part);
            }
          }
        }
      }
      __syncwarp();
    }
    const float s = head_sum<C::LPH>(part, lane, l);
    if (on) {
      float* out = g + (size_t)v * C::FP;
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) st_vec<C::VEC>(out + (i * C::GROUP + gl) * C::VEC, gv + i * C::VEC);
      if (l == 0) stat[(size_t)v * 3 * H + 2 * H + k] = s;
    }
  }
}

// ---------------------------------------------------------------------------
// backward, source-centric over the transposed structure
// ---------------------------------------------------------------------------
template <int H, int D, int U>
__global__ void __launch_bounds__(EDGE_THREADS, (EdgeCfg<H, D>::NE <= 4) ? 4 : ((U * EdgeCfg<H, D>::NE > 24) ? 2 : 3))
edge_bwd_kernel(int n_src, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz, const float* __restrict__ q,
                const float* __restrict__ g, const float* __restrict__ stat, float* __restrict__ dzp,
                float* __restrict__ dq_part) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  constexpr int NQ = HSG_N_BINS * H;
  __shared__ float q_s[NQ];
  __shared__ float dq_s[EDGE_WARPS][C::EPS][NQ];
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) q_s[i] = q[i];
  for (int i = threadIdx.x; i < EDGE_WARPS * C::EPS * NQ; i += blockDim.x) (&dq_s[0][0][0])[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;
  float* my_dq = &dq_s[wib][lane_on ? grp : 0][0];

  // software pipeline over rows: indptr two rows ahead, first neighbour chunk one row ahead
  int u = warp;
  int beg = 0, end = 0, begn = 0, endn = 0, v0 = 0, b0 = 0;
  if (u < n_src) {
    beg = __ldg(indptr + u);
    end = __ldg(indptr + u + 1);
    if (beg + lane < end) {
      v0 = __ldg(nbr + beg + lane);
      b0 = __ldg(bin + beg + lane);
    }
  }
  if (u + nwarps < n_src) {
    begn = __ldg(indptr + u + nwarps);
    endn = __ldg(indptr + u + nwarps + 1);
  }
  while (u < n_src) {
    const int u2 = u + 2 * nwarps;
    int beg2 = 0, end2 = 0, v0n = 0, b0n = 0;
    if (u2 < n_src) {
      beg2 = __ldg(indptr + u2);
      end2 = __ldg(indptr + u2 + 1);
    }
    if (begn + lane < endn) {
      v0n = __ldg(nbr + begn + lane);
      b0n = __ldg(bin + begn + lane);
    }
    const float* zrow = zp + (size_t)u * ldz;
    float zv[C::NE], acc[C::NE];
    float pu = 0.f, acc_dp = 0.f;
#pragma unroll
    for (int i = 0; i < C::NE; ++i) {
      zv[i] = 0.f;
      acc[i] = 0.f;
    }
    if (lane_on && end > beg) {
      pu = __ldg(zrow + C::FP + k);
#pragma unroll
      for (int i = 0; i < C::VPL; ++i)
        if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(zrow + (i * C::GROUP + gl) * C::VEC, zv + i * C::VEC);
    }
    for (int c0 = beg; c0 < end; c0 += 32) {
      const int cnt = min(32, end - c0);
      int my_v = v0, my_b = b0;
      if (c0 != beg) {
        my_v = 0;
        my_b = 0;
        if (lane < cnt) {
          my_v = __ldg(nbr + c0 + lane);
          my_b = __ldg(bin + c0 + lane);
        }
      }
      for (int j0 = 0; j0 < cnt; j0 += C::EPS * U) {   // warp-uniform (shuffles inside)
        float gv[U][C::NE], mk[U], dk[U], sk[U];
        int bb[U];
        bool ok[U];
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          const int j = j0 + uu * C::EPS + grp;
          const int v = __shfl_sync(0xffffffffu, my_v, j & 31);
          bb[uu] = __shfl_sync(0xffffffffu, my_b, j & 31);
          ok[uu] = lane_on && j < cnt;
          mk[uu] = 0.f;
          dk[uu] = 1.f;
          sk[uu] = 0.f;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) gv[uu][i] = 0.f;
          if (ok[uu]) {
            const float* grow = g + (size_t)v * C::FP;
            const float* st = stat + (size_t)v * 3 * H;
            mk[uu] = __ldg(st + k);
            dk[uu] = __ldg(st + H + k);
            sk[uu] = __ldg(st + 2 * H + k);
#pragma unroll
            for (int i = 0; i < C::VPL; ++i)
              if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(grow + (i * C::GROUP + gl) * C::VEC, gv[uu] + i * C::VEC);
          }
        }
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) part = fmaf(gv[uu][i], zv[i], part);
          const float t = head_sum<C::LPH>(part, lane, l);
          if (ok[uu]) {
            const float pre = pu + q_s[bb[uu] * H + k];
            const float lg = pre > 0.f ? pre : HSG_LEAKY_SLOPE * pre;
            const float alpha = __fdividef(exp_fast(lg - mk[uu]), dk[uu]);
            const float de = alpha * (t - sk[uu]);
            const float dpre = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
#pragma unroll
            for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(alpha, gv[uu][i], acc[i]);
            acc_dp += dpre;
            if (l == 0) my_dq[bb[uu] * H + k] += dpre;
          }
        }
      }
    }
    // merge groups (fixed order) and write [dz | dp | 0]
#pragma unroll
    for (int g2 = 1; g2 < C::EPS; ++g2) {
      const int src = (gl + g2 * C::GROUP) & 31;
      const float dp2 = __shfl_sync(0xffffffffu, acc_dp, src);
      if (grp == 0) acc_dp += dp2;     // only group 0 accumulates (others keep their partials intact)
#pragma unroll
      for (int i = 0; i < C::NE; ++i) {
        const float a2 = __shfl_sync(0xffffffffu, acc[i], src);
        if (grp == 0) acc[i] += a2;
      }
    }
    float* drow = dzp + (size_t)u * ldz;
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i >= C::NV) {
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) acc[i * C::VEC + t] = 0.f;   // layout holes must be finite zeros
        }
        st_vec<C::VEC>(drow + (i * C::GROUP + gl) * C::VEC, acc + i * C::VEC);
      }
      if (l == 0) drow[C::FP + k] = acc_dp;
    }
    for (int c = C::FP + H + lane; c < ldz; c += 32) drow[c] = 0.f;
    u += nwarps;
    beg = begn;
    end = endn;
    begn = beg2;
    endn = end2;
    v0 = v0n;
    b0 = b0n;
  }
  __syncthreads();
  // per-block partial of dq, fixed summation order over (warp, group)
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < EDGE_WARPS; ++w)
#pragma unroll
      for (int g2 = 0; g2 < C::EPS; ++g2) s += dq_s[w][g2][i];
    dq_part[(size_t)blockIdx.x * NQ + i] = s;
  }
}

// ---------------------------------------------------------------------------
// backward over LOW-DEGREE rows (word rows: 1-3 in-edges): every GROUP-lane group of a warp walks ITS OWN row, so a
// warp has EPS rows in flight, instead of the EPS groups sharing one row's (mostly single-entry) edge list.  Same
// arithmetic and summation order per row as edge_bwd_kernel with one group; ncu on the 2 048-graph shard showed
// the shared-row mapping issue-bound on these rows (sm throughput 67 %, DRAM 21 %).
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS, 4)
edge_bwd_rowpar_kernel(int n_src, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                       const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                       const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ stat,
                       float* __restrict__ dzp, float* __restrict__ dq_part) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  constexpr int NQ = HSG_N_BINS * H;
  constexpr int U = 2;
  __shared__ float q_s[NQ];
  __shared__ float dq_s[EDGE_WARPS][C::EPS][NQ];
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) q_s[i] = q[i];
  for (int i = threadIdx.x; i < EDGE_WARPS * C::EPS * NQ; i += blockDim.x) (&dq_s[0][0][0])[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;
  float* my_dq = &dq_s[wib][lane_on ? grp : 0][0];
  const int nsteps = ceil_div(n_src, C::EPS);

  int st = warp;
  int beg = 0, end = 0, v0 = 0, b0 = 0;
  {
    const int u = st * C::EPS + grp;
    if (lane_on && st < nsteps && u < n_src) {
      beg = __ldg(indptr + u);
      end = __ldg(indptr + u + 1);
      if (beg < end) {
        v0 = __ldg(nbr + beg);
        b0 = __ldg(bin + beg);
      }
    }
  }
  while (st < nsteps) {
    const int u = st * C::EPS + grp;
    const bool row_on = lane_on && u < n_src;
    // next step's edge range: issued now, consumed at the bottom of the loop
    int begn = 0, endn = 0;
    {
      const int un = (st + nwarps) * C::EPS + grp;
      if (lane_on && st + nwarps < nsteps && un < n_src) {
        begn = __ldg(indptr + un);
        endn = __ldg(indptr + un + 1);
      }
    }
    const int deg = row_on ? end - beg : 0;
    int maxdeg = deg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    float zv[C::NE], acc[C::NE];
    float pu = 0.f, acc_dp = 0.f;
#pragma unroll
    for (int i = 0; i < C::NE; ++i) {
      zv[i] = 0.f;
      acc[i] = 0.f;
    }
    if (deg > 0) {
      const float* zrow = zp + (size_t)u * ldz;
      pu = __ldg(zrow + C::FP + k);
#pragma unroll
      for (int i = 0; i < C::VPL; ++i)
        if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(zrow + (i * C::GROUP + gl) * C::VEC, zv + i * C::VEC);
    }
    for (int j0 = 0; j0 < maxdeg; j0 += U) {           // warp-uniform trip count (shuffles inside)
      float gv[U][C::NE], mk[U], dk[U], sk[U];
      int bb[U];
      bool ok[U];
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        const int j = j0 + uu;
        ok[uu] = j < deg;
        int v = 0;
        bb[uu] = 0;
        if (ok[uu]) {
          v = j == 0 ? v0 : __ldg(nbr + beg + j);
          bb[uu] = j == 0 ? b0 : (int)__ldg(bin + beg + j);
        }
        mk[uu] = 0.f;
        dk[uu] = 1.f;
        sk[uu] = 0.f;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) gv[uu][i] = 0.f;
        if (ok[uu]) {
          const float* grow = g + (size_t)v * C::FP;
          const float* stp = stat + (size_t)v * 3 * H;
          mk[uu] = __ldg(stp + k);
          dk[uu] = __ldg(stp + H + k);
          sk[uu] = __ldg(stp + 2 * H + k);
#pragma unroll
          for (int i = 0; i < C::VPL; ++i)
            if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(grow + (i * C::GROUP + gl) * C::VEC, gv[uu] + i * C::VEC);
        }
      }
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) part = fmaf(gv[uu][i], zv[i], part);
        const float t = head_sum<C::LPH>(part, lane, l);
        if (ok[uu]) {
          const float pre = pu + q_s[bb[uu] * H + k];
          const float lg = pre > 0.f ? pre : HSG_LEAKY_SLOPE * pre;
          const float alpha = __fdividef(exp_fast(lg - mk[uu]), dk[uu]);
          const float de = alpha * (t - sk[uu]);
          const float dpre = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(alpha, gv[uu][i], acc[i]);
          acc_dp += dpre;
          if (l == 0) my_dq[bb[uu] * H + k] += dpre;
        }
      }
    }
    if (row_on) {
      float* drow = dzp + (size_t)u * ldz;
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i >= C::NV) {
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) acc[i * C::VEC + t] = 0.f;   // layout holes must be finite zeros
        }
        st_vec<C::VEC>(drow + (i * C::GROUP + gl) * C::VEC, acc + i * C::VEC);
      }
      if (l == 0) drow[C::FP + k] = acc_dp;
      for (int c = C::FP + H + gl; c < ldz; c += C::GROUP) drow[c] = 0.f;
    }
    st += nwarps;
    beg = begn;
    end = endn;
    v0 = 0;
    b0 = 0;
    if (beg < end) {                                   // first edge of the next row
      v0 = __ldg(nbr + beg);
      b0 = __ldg(bin + beg);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < EDGE_WARPS; ++w)
#pragma unroll
      for (int g2 = 0; g2 < C::EPS; ++g2) s += dq_s[w][g2][i];
    dq_part[(size_t)blockIdx.x * NQ + i] = s;
  }
}

// ---------------------------------------------------------------------------
// backward over FEW, HIGH-degree rows (batch 32: 1 009 supernode rows of ~17, at most ~60, word neighbours): with a
// warp per row only n_rows/8 CTAs exist and the kernel lasts as long as its longest row.  Here a whole CTA walks one
// row: warp w takes the w-th contiguous slice of the edge list, the eight partial [dz | dp] rows meet in shared
// memory and are summed in warp order (fixed, deterministic).
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS, 3)
edge_bwd_blockrow_kernel(int n_src, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                         const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                         const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ stat,
                         float* __restrict__ dzp, float* __restrict__ dq_part) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  constexpr int NQ = HSG_N_BINS * H;
  constexpr int U = 2;
  constexpr int PW = C::FP + H;                      // floats of one partial row
  __shared__ float q_s[NQ];
  __shared__ float dq_s[EDGE_WARPS][C::EPS][NQ];
  __shared__ __align__(16) float part_s[EDGE_WARPS][PW + 2];
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) q_s[i] = q[i];
  for (int i = threadIdx.x; i < EDGE_WARPS * C::EPS * NQ; i += blockDim.x) (&dq_s[0][0][0])[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;
  float* my_dq = &dq_s[wib][lane_on ? grp : 0][0];

  for (int u = blockIdx.x; u < n_src; u += gridDim.x) {
    const int beg = __ldg(indptr + u), end = __ldg(indptr + u + 1);
    const int per = (end - beg + EDGE_WARPS - 1) / EDGE_WARPS;
    const int my_beg = min(end, beg + wib * per), my_end = min(end, my_beg + per);
    const float* zrow = zp + (size_t)u * ldz;
    float zv[C::NE], acc[C::NE];
    float pu = 0.f, acc_dp = 0.f;
#pragma unroll
    for (int i = 0; i < C::NE; ++i) {
      zv[i] = 0.f;
      acc[i] = 0.f;
    }
    if (lane_on && my_end > my_beg) {
      pu = __ldg(zrow + C::FP + k);
#pragma unroll
      for (int i = 0; i < C::VPL; ++i)
        if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(zrow + (i * C::GROUP + gl) * C::VEC, zv + i * C::VEC);
    }
    for (int c0 = my_beg; c0 < my_end; c0 += 32) {
      const int cnt = min(32, my_end - c0);
      int my_v = 0, my_b = 0;
      if (lane < cnt) {
        my_v = __ldg(nbr + c0 + lane);
        my_b = __ldg(bin + c0 + lane);
      }
      for (int j0 = 0; j0 < cnt; j0 += C::EPS * U) {   // warp-uniform (shuffles inside)
        float gv[U][C::NE], mk[U], dk[U], sk[U];
        int bb[U];
        bool ok[U];
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          const int j = j0 + uu * C::EPS + grp;
          const int v = __shfl_sync(0xffffffffu, my_v, j & 31);
          bb[uu] = __shfl_sync(0xffffffffu, my_b, j & 31);
          ok[uu] = lane_on && j < cnt;
          mk[uu] = 0.f;
          dk[uu] = 1.f;
          sk[uu] = 0.f;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) gv[uu][i] = 0.f;
          if (ok[uu]) {
            const float* grow = g + (size_t)v * C::FP;
            const float* stp = stat + (size_t)v * 3 * H;
            mk[uu] = __ldg(stp + k);
            dk[uu] = __ldg(stp + H + k);
            sk[uu] = __ldg(stp + 2 * H + k);
#pragma unroll
            for (int i = 0; i < C::VPL; ++i)
              if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(grow + (i * C::GROUP + gl) * C::VEC, gv[uu] + i * C::VEC);
          }
        }
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) part = fmaf(gv[uu][i], zv[i], part);
          const float t = head_sum<C::LPH>(part, lane, l);
          if (ok[uu]) {
            const float pre = pu + q_s[bb[uu] * H + k];
            const float lg = pre > 0.f ? pre : HSG_LEAKY_SLOPE * pre;
            const float alpha = __fdividef(exp_fast(lg - mk[uu]), dk[uu]);
            const float de = alpha * (t - sk[uu]);
            const float dpre = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
#pragma unroll
            for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(alpha, gv[uu][i], acc[i]);
            acc_dp += dpre;
            if (l == 0) my_dq[bb[uu] * H + k] += dpre;
          }
        }
      }
    }
    // merge the EPS groups of this warp (fixed order), park the warp's partial row in shared memory
#pragma unroll
    for (int g2 = 1; g2 < C::EPS; ++g2) {
      const int src = (gl + g2 * C::GROUP) & 31;
      const float dp2 = __shfl_sync(0xffffffffu, acc_dp, src);
      if (grp == 0) acc_dp += dp2;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) {
        const float a2 = __shfl_sync(0xffffffffu, acc[i], src);
        if (grp == 0) acc[i] += a2;
      }
    }
    if (grp == 0) {
      float* pr = &part_s[wib][0];
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i >= C::NV) {
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) acc[i * C::VEC + t] = 0.f;   // layout holes must be finite zeros
        }
#pragma unroll
        for (int t = 0; t < C::VEC; ++t) pr[(i * C::GROUP + gl) * C::VEC + t] = acc[i * C::VEC + t];
      }
      if (l == 0) pr[C::FP + k] = acc_dp;
    }
    __syncthreads();
    float* drow = dzp + (size_t)u * ldz;
    for (int t = threadIdx.x; t < ldz; t += blockDim.x) {
      float sres = 0.f;
      if (t < PW) {
#pragma unroll
        for (int w = 0; w < EDGE_WARPS; ++w) sres += part_s[w][t];   // warp order: fixed
      }
      drow[t] = sres;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) {
    float sres = 0.f;
#pragma unroll
    for (int w = 0; w < EDGE_WARPS; ++w)
#pragma unroll
      for (int g2 = 0; g2 < C::EPS; ++g2) sres += dq_s[w][g2][i];
    dq_part[(size_t)blockIdx.x * NQ + i] = sres;
  }
}

// ---------------------------------------------------------------------------
// backward over HIGH-degree rows with ASYNCHRONOUS gathers (the supernode rows of the S2W layer on a large shard:
// ~17 word neighbours of 1.2 KB each).  ncu showed edge_bwd_kernel latency-bound there (long-scoreboard stalls on the
// first use of a gathered row, 24 warps/SM x 2 rows in flight because the gathered rows live in registers).  Here the
// neighbour rows [g_v | (m, den, s)_v] are copied global -> shared with cp.async into a per-warp ring of R slots,
// R - 1 rows ahead of the row being consumed, across row boundaries: a warp owns a CHUNK of consecutive rows, whose
// edge lists are one contiguous range of the CSC, and streams through that range.  Same arithmetic and summation
// order per row as edge_bwd_kernel (one group per warp).  Layouts with one GROUP per warp (EPS == 1) only.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int ASYNC_ROWS_PER_CHUNK = 8;

template <int H, int D>
struct AsyncCfg {
  using C = EdgeCfg<H, D>;
  static constexpr int SLOT = (C::FP + 3 * H + 1 + 3) & ~3;          // floats: [g row | m den s | bin], 16-byte multiple
  // ring depth: as deep as ~7.5 KB per warp allows (3 CTAs of 8 warps per SM), at least 3, at most 8
  static constexpr int R_FIT = (7680 / 4) / SLOT;
#ifdef HSG_ASYNC_R
  static constexpr int R = HSG_ASYNC_R;
#else
  static constexpr int R = R_FIT < 3 ? 3 : (R_FIT > 8 ? 8 : R_FIT);
#endif
  static constexpr size_t SMEM = (size_t)EDGE_WARPS * R * SLOT * sizeof(float);
};

template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS, 3)
edge_bwd_async_kernel(int n_src, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                      const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                      const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ stat,
                      float* __restrict__ dzp, float* __restrict__ dq_part, int RC) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  using A = AsyncCfg<H, D>;
  static_assert(C::EPS == 1 && C::FP % 4 == 0, "one group per warp, 16-byte row copies");
  constexpr int NQ = HSG_N_BINS * H;
  constexpr int R = A::R, SLOT = A::SLOT;
  extern __shared__ float4 ring_raw[];
  __shared__ float q_s[NQ];
  __shared__ float dq_s[EDGE_WARPS][NQ];
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) q_s[i] = q[i];
  for (int i = threadIdx.x; i < EDGE_WARPS * NQ; i += blockDim.x) (&dq_s[0][0])[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const bool lane_on = lane < C::GROUP;
  float* my_dq = &dq_s[wib][0];
  float* ring = reinterpret_cast<float*>(ring_raw) + (size_t)wib * R * SLOT;
  const int nchunks = ceil_div(n_src, RC);

  for (int chunk = warp; chunk < nchunks; chunk += nwarps) {
    const int r0 = chunk * RC, nrows = min(RC, n_src - r0);
    const int my_ip = lane <= nrows ? __ldg(indptr + r0 + lane) : 0;        // row pointers of the chunk
    const int e_beg = __shfl_sync(0xffffffffu, my_ip, 0), e_end = __shfl_sync(0xffffffffu, my_ip, nrows);
    // ---- producer state: neighbour ids / bins in blocks of 32 edges, one block ahead ----
    int pblk = e_beg;
    int p_ids = 0, p_bins = 0, p_ids_n = 0, p_bins_n = 0;
    if (pblk + lane < e_end) {
      p_ids = __ldg(nbr + pblk + lane);
      p_bins = __ldg(bin + pblk + lane);
    }
    if (pblk + 32 + lane < e_end) {
      p_ids_n = __ldg(nbr + pblk + 32 + lane);
      p_bins_n = __ldg(bin + pblk + 32 + lane);
    }
    int pe = e_beg;
    auto issue = [&]() {                                   // copies of edge pe into its ring slot (warp-uniform)
      if (pe < e_end) {
        if (pe - pblk >= 32) {
          pblk += 32;
          p_ids = p_ids_n;
          p_bins = p_bins_n;
          p_ids_n = 0;
          p_bins_n = 0;
          if (pblk + 32 + lane < e_end) {
            p_ids_n = __ldg(nbr + pblk + 32 + lane);
            p_bins_n = __ldg(bin + pblk + 32 + lane);
          }
        }
        const int v = __shfl_sync(0xffffffffu, p_ids, pe - pblk);
        const int b = __shfl_sync(0xffffffffu, p_bins, pe - pblk);
        float* slot = ring + ((pe - e_beg) % R) * SLOT;
        const float* grow = g + (size_t)v * C::FP;
        for (int c4 = lane; c4 < C::FP / 4; c4 += 32) cp_async_16(slot + 4 * c4, grow + 4 * c4);
        const float* srow = stat + (size_t)v * 3 * H;
        for (int c = lane; c < 3 * H; c += 32) cp_async_4(slot + C::FP + c, srow + c);
        if (lane == 0) slot[C::FP + 3 * H] = __int_as_float(b);
      }
      cp_async_commit();                                   // one group per edge slot, empty or not
      ++pe;
    };
#pragma unroll 1
    for (int i = 0; i < R - 1; ++i) issue();

    // ---- consumer state ----
    int row = 0;                                            // row index inside the chunk
    int row_end = __shfl_sync(0xffffffffu, my_ip, 1);
    float zv[C::NE], zn[C::NE], acc[C::NE];
    float pu = 0.f, pn = 0.f, acc_dp = 0.f;
#pragma unroll
    for (int i = 0; i < C::NE; ++i) {
      zv[i] = 0.f;
      zn[i] = 0.f;
      acc[i] = 0.f;
    }
    auto load_z = [&](int r, float* zdst, float& pdst) {    // [z | p] of source row r0 + r (zeros past the chunk)
#pragma unroll
      for (int i = 0; i < C::NE; ++i) zdst[i] = 0.f;
      pdst = 0.f;
      if (lane_on && r < nrows) {
        const float* zrow = zp + (size_t)(r0 + r) * ldz;
        pdst = __ldg(zrow + C::FP + k);
#pragma unroll
        for (int i = 0; i < C::VPL; ++i)
          if (l + C::LPH * i < C::NV) ld_vec<C::VEC>(zrow + (i * C::GROUP + gl) * C::VEC, zdst + i * C::VEC);
      }
    };
    auto finalize = [&]() {                                 // write [dz | dp | 0] of the current row, reset
      float* drow = dzp + (size_t)(r0 + row) * ldz;
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i >= C::NV) {
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) acc[i * C::VEC + t] = 0.f;   // layout holes must be finite zeros
          }
          st_vec<C::VEC>(drow + (i * C::GROUP + gl) * C::VEC, acc + i * C::VEC);
        }
        if (l == 0) drow[C::FP + k] = acc_dp;
      }
      for (int c = C::FP + H + lane; c < ldz; c += 32) drow[c] = 0.f;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) acc[i] = 0.f;
      acc_dp = 0.f;
      ++row;
      row_end = __shfl_sync(0xffffffffu, my_ip, min(row + 1, 31));
#pragma unroll
      for (int i = 0; i < C::NE; ++i) zv[i] = zn[i];
      pu = pn;
      load_z(row + 1, zn, pn);                             // one row ahead
    };
    load_z(0, zv, pu);
    load_z(1, zn, pn);

    for (int ce = e_beg; ce < e_end; ++ce) {
      issue();
      cp_async_wait<R - 1>();
      __syncwarp();
      while (ce >= row_end) finalize();                    // rows that ended (or are empty) before this edge
      const float* slot = ring + ((ce - e_beg) % R) * SLOT;
      float gv[C::NE];
#pragma unroll
      for (int i = 0; i < C::NE; ++i) gv[i] = 0.f;
      float mk = 0.f, dk = 1.f, sk = 0.f;
      int bb = 0;
      if (lane_on) {
        mk = slot[C::FP + k];
        dk = slot[C::FP + H + k];
        sk = slot[C::FP + 2 * H + k];
        bb = __float_as_int(slot[C::FP + 3 * H]);
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            const float* sp = slot + (i * C::GROUP + gl) * C::VEC;
            if (C::VEC == 4) {
              const float4 t4 = *reinterpret_cast<const float4*>(sp);
              gv[i * C::VEC] = t4.x; gv[i * C::VEC + 1] = t4.y; gv[i * C::VEC + 2] = t4.z; gv[i * C::VEC + 3] = t4.w;
            } else if (C::VEC == 2) {
              const float2 t2 = *reinterpret_cast<const float2*>(sp);
              gv[i * C::VEC] = t2.x; gv[i * C::VEC + 1] = t2.y;
            } else {
              gv[i * C::VEC] = sp[0];
            }
          }
        }
      }
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) part = fmaf(gv[i], zv[i], part);
      const float t = head_sum<C::LPH>(part, lane, l);
      if (lane_on) {
        const float pre = pu + q_s[bb * H + k];
        const float lg = pre > 0.f ? pre : HSG_LEAKY_SLOPE * pre;
        const float alpha = __fdividef(exp_fast(lg - mk), dk);
        const float de = alpha * (t - sk);
        const float dpre = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(alpha, gv[i], acc[i]);
        acc_dp += dpre;
        if (l == 0) my_dq[bb * H + k] += dpre;
      }
      __syncwarp();                                        // the slot is free for the copy issued next iteration
    }
    while (row < nrows) finalize();                        // last row with edges and trailing empty rows
    cp_async_wait<0>();
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) {
    float sres = 0.f;
#pragma unroll
    for (int w = 0; w < EDGE_WARPS; ++w) sres += dq_s[w][i];
    dq_part[(size_t)blockIdx.x * NQ + i] = sres;
  }
}

// dq[i] = sum over the per-block partials of nseg launches (segment j at dq_part + j * seg_stride), fixed order: one
// CTA per output, strided partial sums + smem tree
__global__ void __launch_bounds__(128) edge_bwd_dq_kernel(int nseg, int nblocks, size_t seg_stride, int nq,
                                                          const float* __restrict__ dq_part, float* __restrict__ dq,
                                                          int accumulate) {
  pdl_prologue();
  __shared__ float red[128];
  const int i = blockIdx.x;
  float s = 0.f;
  for (int j = 0; j < nseg; ++j) {
    const float* part = dq_part + (size_t)j * seg_stride;
    for (int b = threadIdx.x; b < nblocks; b += 128) s += part[(size_t)b * nq + i];
  }
  red[threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dq[i] = accumulate ? dq[i] + red[0] : red[0];
}

int edge_dq_reduce(int nseg, int nblocks, size_t seg_stride, int nq, const float* dq_part, float* dq, int accumulate,
                   cudaStream_t s) {
  LaunchScope ls(SLOT_EDGE_BWD_DQ, s);
  launch_k(edge_bwd_dq_kernel, dim3(nq), dim3(128), 0, s, nseg, nblocks, seg_stride, nq, dq_part, dq, accumulate);
  return check_launch();
}

#ifndef HSG_BWD_UHI_WIDE
#define HSG_BWD_UHI_WIDE 2
#endif
constexpr int EDGE_MAX_BLOCKS = 148 * 32;  // upper bound of the grid (sizes the dq partial workspace)
constexpr int EDGE_DEFAULT_BLOCKS = 148 * 8;

// Grid caps (blocks of 8 warps; a warp walks rows in a grid-stride loop with a software pipeline over rows).  Measured
// on the 2 048-graph shard (L2 flushed): the streaming-heavy kernels (bwd_prep; forward over wide rows) gain from
// 32 blocks per SM in flight-order (S2W prep 0.60 -> 0.67 of the HBM peak, S2W fwd 0.705 -> 0.735), the gather-heavy
// backward kernels prefer 8 per SM.  HSG_EDGE_MAX_BLOCKS overrides all of them (tuning).
static int edge_block_cap(int dflt) {
  static int env = -1;
  if (env < 0) {
    env = 0;
    const char* e = getenv("HSG_EDGE_MAX_BLOCKS");
    if (e) {
      const int v = atoi(e);
      if (v >= 1 && v <= EDGE_MAX_BLOCKS) env = v;
    }
  }
  return env > 0 ? env : dflt;
}

static int edge_grid(int n_rows_steps, int cap = EDGE_DEFAULT_BLOCKS) {
  int blocks = ceil_div(n_rows_steps, EDGE_WARPS);
  cap = edge_block_cap(cap);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return blocks;
}

// (H, D) instantiations: defaults (8,8) W2S and (6,50) S2W (HiGraph.py:57-76), plus
// other hidden/embedding sizes and the small shapes used by the test fixtures.
#define HSG_EDGE_CONFIGS(X) \
  X(8, 8) X(6, 50) X(8, 16) X(6, 16) X(8, 32) X(6, 32) X(4, 4) X(6, 8) X(4, 16) X(1, 64) X(16, 4) X(2, 32) X(4, 32) X(12, 25)

static std::atomic<int> g_fwd_rowpar{-1};   // -1 auto (many rows), 0 never, 1 whenever the layout allows
static std::atomic<int> g_fwd_lowdeg{-1};   // -1 auto (many low-degree wide rows), 0 never, 1 whenever the layout allows
constexpr int FWD_LOWDEG_MIN_ROWS = 16384;
constexpr int FWD_ROWPAR_MIN_ROWS = 16384;

template <int H, int D>
static bool launch_fwd_rowpar(int blocks, bool deep, const hsg_csc* c, const float* zp, int ldz, const float* q,
                              const float* origin, float* sh, float* x, float* stat, cudaStream_t s) {
  if constexpr (EdgeCfg<H, D>::EPS > 1 && !EdgeCfg<H, D>::STAGED) {
    constexpr int UR = EdgeCfg<H, D>::NE <= 8 ? 4 : 2;
    if (deep)
      launch_k(edge_fwd_rowpar_kernel<H, D, UR>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr,
               c->bin, c->extra, zp, ldz, q, origin, sh, x, stat);
    else
      launch_k(edge_fwd_rowpar_kernel<H, D, 2>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr,
               c->bin, c->extra, zp, ldz, q, origin, sh, x, stat);
    return true;
  }
  return false;
}

// unroll depth: deep for high-degree rows (supernodes), shallow for low-degree rows (words)
template <int H, int D>
static int launch_fwd(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* origin, float* sh,
                      float* x, float* stat, cudaStream_t s) {
  using C = EdgeCfg<H, D>;
  constexpr int UHI = C::NE <= 4 ? 4 : 2, ULO = C::NE <= 4 ? 2 : 1;
  const bool deep = (double)c->n_edges > 4.0 * C::EPS * (double)c->n_dst;
  const bool deep_row = (double)c->n_edges > 4.0 * (double)c->n_dst;
  const int rp = g_fwd_rowpar.load(std::memory_order_relaxed);
  const bool rowpar = C::EPS > 1 && !C::STAGED && (rp == 1 || (rp < 0 && c->n_dst >= FWD_ROWPAR_MIN_ROWS));
  LaunchScope ls(SLOT_EDGE_FWD, s);
  if (rowpar && launch_fwd_rowpar<H, D>(edge_grid(ceil_div(c->n_dst, C::EPS)), deep_row, c, zp, ldz, q, origin, sh, x,
                                        stat, s))
    return check_launch();
  if constexpr (C::EPS == 1 && C::STAGED) {
    // wide rows of low degree in numbers (the word rows of a shard): the pipelined straight-line kernel
    const int ld = g_fwd_lowdeg.load(std::memory_order_relaxed);
    if (ld == 1 || (ld < 0 && !deep && c->n_dst >= FWD_LOWDEG_MIN_ROWS)) {
      launch_k(edge_fwd_lowdeg_kernel<H, D>, dim3(edge_grid(c->n_dst, EDGE_MAX_BLOCKS)), dim3(EDGE_THREADS), 0, s,
               c->n_dst, c->indptr, c->nbr, c->bin, c->extra, zp, ldz, q, origin, sh, x, stat);
      return check_launch();
    }
  }
  if (deep)
    launch_k(edge_fwd_kernel<H, D, UHI>, dim3(edge_grid(c->n_dst, C::STAGED ? EDGE_MAX_BLOCKS : EDGE_DEFAULT_BLOCKS)), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin,
                                                                           c->extra, zp, ldz, q, origin, sh, x, stat);
  else
    launch_k(edge_fwd_kernel<H, D, ULO>, dim3(edge_grid(c->n_dst, C::STAGED ? EDGE_MAX_BLOCKS : EDGE_DEFAULT_BLOCKS)), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin,
                                                                           c->extra, zp, ldz, q, origin, sh, x, stat);
  return check_launch();
}

template <int H, int D>
static int launch_prep(int n_dst, const float* dx, const float* dsh, const float* sh, float* g, float* stat,
                       cudaStream_t s) {
  LaunchScope ls(SLOT_EDGE_BWD_PREP, s);
  launch_k(edge_bwd_prep_kernel<H, D>, dim3(edge_grid(ceil_div(n_dst, EdgeCfg<H, D>::EPS), EDGE_MAX_BLOCKS)), dim3(EDGE_THREADS), 0, s, n_dst, dx, dsh,
                                                                                                   sh, g, stat);
  return check_launch();
}

static std::atomic<int> g_rowpar{-1};   // -1 auto (low average degree), 0 never, 1 whenever the layout allows
static std::atomic<int> g_blockrow{-1}; // -1 auto (few high-degree rows), 0 never, 1 always
constexpr int BLOCKROW_MAX_ROWS = 8192;

template <int H, int D>
static void launch_rowpar(int blocks, const hsg_csc* c, const float* zp, int ldz, const float* q, const float* g,
                          const float* stat, float* dzp, float* ws, cudaStream_t s) {
  if constexpr (EdgeCfg<H, D>::EPS > 1 && !EdgeCfg<H, D>::STAGED) {
    launch_k(edge_bwd_rowpar_kernel<H, D>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin,
             zp, ldz, q, g, stat, dzp, ws);
  }
}

// -1 auto, 0 never, 1 whenever the layout allows.  Auto NEVER picks this mapping: measured on the 2 048-graph shard
// (S2W backward, 35 k rows x ~30 neighbours of 1.2 KB) it takes 414 us against 360 us for the register-gather kernel -
// ncu (profiles/r01k_full_edge_bwd_async.txt): issue slots 69 % busy, DRAM 38 %: the ring bookkeeping and the shared-
// memory round trip cost more issue slots than the removed scoreboard stalls gave back.  Kept as a parity-tested opt-in.
static std::atomic<int> g_bwd_async{-1};

template <int H, int D>
static bool launch_bwd_async(int* blocks_out, const hsg_csc* c, const float* zp, int ldz, const float* q,
                             const float* g, const float* stat, float* dzp, float* ws, cudaStream_t s) {
  if constexpr (EdgeCfg<H, D>::EPS == 1 && EdgeCfg<H, D>::FP % 4 == 0) {
    using A = AsyncCfg<H, D>;
    static bool attr_done = false;
    if (!attr_done) {
      if (cudaFuncSetAttribute(edge_bwd_async_kernel<H, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)A::SMEM) != cudaSuccess)
        return false;
      attr_done = true;
    }
    static int rc = 0;
    if (!rc) {
      const char* e = getenv("HSG_ASYNC_RC");
      rc = e ? atoi(e) : ASYNC_ROWS_PER_CHUNK;
      if (rc < 1 || rc > 30) rc = ASYNC_ROWS_PER_CHUNK;
    }
    const int blocks = edge_grid(ceil_div(c->n_dst, rc), 148 * 3);
    launch_k(edge_bwd_async_kernel<H, D>, dim3(blocks), dim3(EDGE_THREADS), A::SMEM, s, c->n_dst, c->indptr, c->nbr,
             c->bin, zp, ldz, q, g, stat, dzp, ws, rc);
    *blocks_out = blocks;
    return true;
  }
  return false;
}

template <int H, int D>
static int launch_bwd(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* g, const float* stat,
                      float* dzp, float* dq, float* ws, int accumulate_dq, cudaStream_t s, int* defer_blocks) {
  using C = EdgeCfg<H, D>;
  constexpr int UHI = C::NE <= 4 ? 4 : (HSG_BWD_UHI_WIDE), ULO = C::NE <= 4 ? 2 : 1;
  const bool deep = (double)c->n_edges > 4.0 * C::EPS * (double)c->n_dst;
  const int rp = g_rowpar.load(std::memory_order_relaxed);
  const bool rowpar = C::EPS > 1 && !C::STAGED && (rp == 1 || (rp < 0 && !deep));
  // few high-degree rows (a warp per row would leave most SMs idle and last as long as the longest row): CTA per row
  const int br = g_blockrow.load(std::memory_order_relaxed);
  const bool blockrow = !rowpar && (br == 1 || (br < 0 && deep && c->n_dst <= BLOCKROW_MAX_ROWS));
  int blocks = blockrow ? min(c->n_dst, edge_block_cap(EDGE_DEFAULT_BLOCKS))
                        : (rowpar ? edge_grid(ceil_div(c->n_dst, C::EPS)) : edge_grid(c->n_dst));
  const int as = g_bwd_async.load(std::memory_order_relaxed);
  const bool async = !rowpar && !blockrow && as == 1;
  {
    LaunchScope ls(SLOT_EDGE_BWD, s);
    if (async && launch_bwd_async<H, D>(&blocks, c, zp, ldz, q, g, stat, dzp, ws, s)) {
    } else if (blockrow)
      launch_k(edge_bwd_blockrow_kernel<H, D>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin,
               zp, ldz, q, g, stat, dzp, ws);
    else if (rowpar)
      launch_rowpar<H, D>(blocks, c, zp, ldz, q, g, stat, dzp, ws, s);
    else if (deep)
      launch_k(edge_bwd_kernel<H, D, UHI>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin, zp, ldz, q, g,
                                                                 stat, dzp, ws);
    else
      launch_k(edge_bwd_kernel<H, D, ULO>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin, zp, ldz, q, g,
                                                                 stat, dzp, ws);
    int rc = check_launch();
    if (rc) return rc;
  }
  if (defer_blocks) {                                   // the caller reduces the partials of several launches at once
    *defer_blocks = blocks;
    return HSG_OK;
  }
  return edge_dq_reduce(1, blocks, 0, HSG_N_BINS * H, ws, dq, accumulate_dq, s);
}

static bool layout_ok(int H, int d, int ldz) {
  if (H <= 0 || d <= 0 || H > 32) return false;
  const EdgeLayout L = make_edge_layout(H, d);
  return ldz % 4 == 0 && ldz >= L.fp + H;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_edge_layout(int H, int d, int* fp, int* ldz) {
  if (H <= 0 || d <= 0 || H > 32 || !fp || !ldz) return HSG_ERR_ARG;
  bool found = false;
#define X(HH, DD) \
  if (H == HH && d == DD) found = true;
  HSG_EDGE_CONFIGS(X)
#undef X
  if (!found) return HSG_ERR_SHAPE;
  const EdgeLayout L = make_edge_layout(H, d);
  *fp = L.fp;
  *ldz = L.ldz;
  return HSG_OK;
}

int hsg_edge_perm(int H, int d, int col) {
  if (H <= 0 || d <= 0 || H > 32 || col < 0 || col >= H * d) return HSG_ERR_ARG;
  const EdgeLayout L = make_edge_layout(H, d);
  return edge_perm(L, col);
}

int hsg_edge_fwd(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q, const float* origin,
                 float* sh, float* x, float* stat, void* stream) {
  if (!csc || csc->n_dst < 0) return HSG_ERR_ARG;
  if (csc->n_dst == 0) return HSG_OK;                 // empty destination set (row pointers may be NULL)
  if (!zp || !q || !stat || (!sh && !x)) return HSG_ERR_ARG;   // sh may be NULL (segment backward recomputes it)
  if (x != nullptr && origin == nullptr) return HSG_ERR_ARG;
  if (!csc->indptr || (csc->n_edges > 0 && (!csc->nbr || !csc->bin))) return HSG_ERR_ARG;
  if (!layout_ok(H, d, ldz) || !aligned16(zp) || (sh && !aligned16(sh)) || (x && (!aligned16(x) || !aligned16(origin))))
    return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
#define X(HH, DD) \
  if (H == HH && d == DD) return launch_fwd<HH, DD>(csc, zp, ldz, q, origin, sh, x, stat, s);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

int hsg_edge_bwd_prep(int n_dst, int H, int d, const float* dx, const float* dsh, const float* sh, float* g,
                      float* stat, void* stream) {
  if (n_dst < 0 || (!dx && !dsh) || !sh || !g || !stat) return HSG_ERR_ARG;
  if (n_dst == 0) return HSG_OK;
  if (!aligned16(sh) || !aligned16(g) || (dx && !aligned16(dx)) || (dsh && !aligned16(dsh))) return HSG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
#define X(HH, DD) \
  if (H == HH && d == DD) return launch_prep<HH, DD>(n_dst, dx, dsh, sh, g, stat, s);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

int hsg_set_edge_rowpar(int mode) {
  g_rowpar.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_set_edge_fwd_lowdeg(int mode) {
  g_fwd_lowdeg.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_set_edge_fwd_rowpar(int mode) {
  g_fwd_rowpar.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_set_edge_bwd_async(int mode) {
  g_bwd_async.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_set_edge_blockrow(int mode) {
  g_blockrow.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

size_t hsg_edge_bwd_workspace_bytes(int H) { return (size_t)EDGE_MAX_BLOCKS * HSG_N_BINS * H * sizeof(float) + 16; }

int hsg_edge_bwd(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q, const float* g,
                 const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, void* stream) {
  return edge_bwd_ex(csc_t, H, d, zp, ldz, q, g, stat, dzp, dq, ws, ws_bytes, 0, (cudaStream_t)stream, nullptr);
}

}  // extern "C"

namespace hsg {
int edge_bwd_ex(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q, const float* g,
                const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, int accumulate_dq, cudaStream_t s,
                int* defer_blocks) {
  if (!csc_t || !zp || !q || !g || !stat || !dzp || (!dq && !defer_blocks) || !ws || csc_t->n_dst < 0)
    return HSG_ERR_ARG;
  if (ws_bytes < hsg_edge_bwd_workspace_bytes(H)) return HSG_ERR_WORKSPACE;
  if (!layout_ok(H, d, ldz) || !aligned16(zp) || !aligned16(g) || !aligned16(dzp)) return HSG_ERR_ALIGN;
  if (csc_t->n_dst > 0 && (!csc_t->indptr || (csc_t->n_edges > 0 && (!csc_t->nbr || !csc_t->bin)))) return HSG_ERR_ARG;
#define X(HH, DD) \
  if (H == HH && d == DD) \
    return launch_bwd<HH, DD>(csc_t, zp, ldz, q, g, stat, dzp, dq, (float*)ws, accumulate_dq, s, defer_blocks);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

}  // namespace hsg
