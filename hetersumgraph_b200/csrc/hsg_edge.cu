// K3 / K5: fused WSWGAT edge kernels (forward and backward) over a CSC.
//
// Replaces, for ALL heads at once, what the reference executes per head through
// DGL's UDF runtime (module/GATLayer.py:88-102 / 127-140 driven by apply_edges
// :112/:148 and the degree-bucketed pull :113/:149):
//     e_uv  = leaky_relu(a_src . z_u + a_dst . 0 + a_feat . feat_fc(T[bin_uv]))  = leaky_relu(p_u + q[bin])
//     alpha = softmax over ALL in-edges of v (extra[v] never-written edges: e = 0, z = 0)
//     sh_v  = sum alpha z_u ;   x_v = elu(sh_v) + origin_v          (GAT.py:56-57)
//
// Mapping (compile-time per (H, D)): one warp walks the in-edge list of one
// destination row.  A row of F = H*D floats is spread over GROUP = H*LPH lanes so
// that EVERY LANE OWNS ELEMENTS OF EXACTLY ONE HEAD: the attention logit, the
// online-softmax state (m, den) and the weight alpha are per-lane scalars, there
// is no cross-lane traffic inside the edge loop of the forward pass, and the
// per-edge dot product of the backward pass is an LPH-lane reduction.  When a row
// needs fewer than 32 lanes, EPS = 32/GROUP edges are processed per warp step.
// Source rows are gathered with 128-bit (64-bit when D % 4 != 0) loads.
// Everything is deterministic: no floating-point atomics anywhere.
#include <math_constants.h>

#include "hsg_common.cuh"

namespace hsg {

template <int H_, int D_>
struct EdgeCfg {
  static constexpr int H = H_, D = D_, F = H_ * D_;
  static constexpr int VEC = (D_ % 4 == 0) ? 4 : ((D_ % 2 == 0) ? 2 : 1);
  static constexpr int NV = D_ / VEC;                       // vectors per head
  static constexpr int LPH_MAX = 32 / H_;
  static constexpr int LPH = NV < LPH_MAX ? NV : LPH_MAX;   // lanes per head
  static constexpr int VPL = (NV + LPH - 1) / LPH;          // vectors per lane
  static constexpr int GROUP = H_ * LPH;                    // lanes per edge row
  static constexpr int EPS = 32 / GROUP;                    // edge rows per warp step
  static constexpr int NE = VPL * VEC;                      // elements per lane
  static_assert(H_ <= 32 && LPH >= 1 && EPS >= 1, "bad edge config");
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
// forward
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS)
edge_fwd_kernel(int n_dst, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                const uint8_t* __restrict__ bin, const int32_t* __restrict__ extra, const float* __restrict__ zp,
                int ldz, const float* __restrict__ q, const float* __restrict__ origin, float* __restrict__ sh,
                float* __restrict__ x, float* __restrict__ stat) {
  using C = EdgeCfg<H, D>;
  __shared__ float q_s[HSG_N_BINS * H];
  for (int i = threadIdx.x; i < HSG_N_BINS * H; i += blockDim.x) q_s[i] = q[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;       // head owned by this lane
  const int l = gl % C::LPH;
  const bool lane_on = grp < C::EPS;

  for (int v = warp; v < n_dst; v += nwarps) {
    const int beg = __ldg(indptr + v), end = __ldg(indptr + v + 1);
    const float xcnt = extra ? (float)__ldg(extra + v) : 0.f;
    float m = -CUDART_INF_F, den = 0.f;
    float acc[C::NE];
#pragma unroll
    for (int i = 0; i < C::NE; ++i) acc[i] = 0.f;

    if (lane_on) {
      int e = beg + grp;
      int u = 0, b = 0;
      if (e < end) {
        u = __ldg(nbr + e);
        b = __ldg(bin + e);
      }
      while (e < end) {
        const int en = e + C::EPS;
        int un = 0, bn = 0;
        if (en < end) {
          un = __ldg(nbr + en);
          bn = __ldg(bin + en);
        }
        const float* row = zp + (size_t)u * ldz;
        const float pe = __ldg(row + C::F + k);
        float zv[C::NE];
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          const int nv = l + C::LPH * i;
          if (nv < C::NV) {
            ld_vec<C::VEC>(row + k * D + C::VEC * nv, zv + i * C::VEC);
          } else {
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) zv[i * C::VEC + t] = 0.f;
          }
        }
        const float lg = leaky(pe + q_s[b * H + k]);
        if (lg > m) {
          const float sc = __expf(m - lg);
          den *= sc;
#pragma unroll
          for (int i = 0; i < C::NE; ++i) acc[i] *= sc;
          m = lg;
        }
        const float w = __expf(lg - m);
        den += w;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(w, zv[i], acc[i]);
        e = en;
        u = un;
        b = bn;
      }
    }
    // merge the EPS partial states into group 0 (fixed order)
#pragma unroll
    for (int g2 = 1; g2 < C::EPS; ++g2) {
      const int src = (gl + g2 * C::GROUP) & 31;
      const float m2 = __shfl_sync(0xffffffffu, m, src);
      const float d2 = __shfl_sync(0xffffffffu, den, src);
      const float mn = fmaxf(m, m2);
      const float s1 = (m == -CUDART_INF_F) ? 0.f : __expf(m - mn);
      const float s2 = (m2 == -CUDART_INF_F) ? 0.f : __expf(m2 - mn);
      // only group 0 accumulates: the other groups must keep their own partial state for later reads
      if (grp == 0) den = den * s1 + d2 * s2;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) {
        const float a2 = __shfl_sync(0xffffffffu, acc[i], src);
        if (grp == 0) acc[i] = acc[i] * s1 + a2 * s2;
      }
      if (grp == 0) m = mn;
    }
    if (grp == 0) {
      float mf, inv;
      if (m == -CUDART_INF_F) {  // no word<->supernode in-edge: DGL's zero fill (or softmax over z = 0 extras)
        mf = 0.f;
        den = xcnt > 0.f ? xcnt : 1.f;
        inv = 0.f;
      } else {
        mf = xcnt > 0.f ? fmaxf(m, 0.f) : m;
        const float sc = __expf(m - mf);
        den = den * sc + xcnt * __expf(-mf);
        inv = sc / den;
      }
      float* sh_row = sh + (size_t)v * C::F;
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        const int nv = l + C::LPH * i;
        if (nv < C::NV) {
          const int col = k * D + C::VEC * nv;
          float o[C::VEC];
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) o[t] = acc[i * C::VEC + t] * inv;
          st_vec<C::VEC>(sh_row + col, o);
          if (x != nullptr) {
            float og[C::VEC];
            ld_vec<C::VEC>(origin + (size_t)v * C::F + col, og);
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) og[t] += (o[t] > 0.f ? o[t] : expm1f(o[t]));
            st_vec<C::VEC>(x + (size_t)v * C::F + col, og);
          }
        }
      }
      if (l == 0) {
        stat[(size_t)v * 3 * H + k] = mf;
        stat[(size_t)v * 3 * H + H + k] = den;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward prep: g = dx * elu'(sh)  (or g = dsh), s[v,k] = g_v[k] . sh_v[k]
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS)
edge_bwd_prep_kernel(int n_dst, const float* __restrict__ dx, const float* __restrict__ dsh,
                     const float* __restrict__ sh, float* __restrict__ g, float* __restrict__ stat) {
  using C = EdgeCfg<H, D>;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int grp = lane / C::GROUP;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;
  const int l = gl % C::LPH;
  const int nsteps = ceil_div(n_dst, C::EPS);
  for (int st = warp; st < nsteps; st += nwarps) {
    const int v = st * C::EPS + grp;
    const bool on = grp < C::EPS && v < n_dst;
    float part = 0.f;
    if (on) {
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        const int nv = l + C::LPH * i;
        if (nv < C::NV) {
          const size_t off = (size_t)v * C::F + k * D + C::VEC * nv;
          float s_[C::VEC], d_[C::VEC];
          ld_vec<C::VEC>(sh + off, s_);
          if (dx != nullptr) {
            ld_vec<C::VEC>(dx + off, d_);
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) d_[t] *= (s_[t] > 0.f ? 1.f : __expf(s_[t]));
          } else {
            ld_vec<C::VEC>(dsh + off, d_);
          }
#pragma unroll
          for (int t = 0; t < C::VEC; ++t) part = fmaf(d_[t], s_[t], part);
          st_vec<C::VEC>(g + off, d_);
        }
      }
    }
    const float s = head_sum<C::LPH>(part, lane, l);
    if (on && l == 0) stat[(size_t)v * 3 * H + 2 * H + k] = s;
  }
}

// ---------------------------------------------------------------------------
// backward, source-centric over the transposed structure
// ---------------------------------------------------------------------------
template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS)
edge_bwd_kernel(int n_src, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz, const float* __restrict__ q,
                const float* __restrict__ g, const float* __restrict__ stat, float* __restrict__ dzp,
                float* __restrict__ dq_part) {
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
  float* my_dq = lane_on ? &dq_s[wib][grp][0] : nullptr;

  for (int u = warp; u < n_src; u += nwarps) {
    const int beg = __ldg(indptr + u), end = __ldg(indptr + u + 1);
    const float* zrow = zp + (size_t)u * ldz;
    float zv[C::NE], acc[C::NE];
    float pu = 0.f, acc_dp = 0.f;
#pragma unroll
    for (int i = 0; i < C::NE; ++i) {
      zv[i] = 0.f;
      acc[i] = 0.f;
    }
    if (lane_on && end > beg) {
      pu = __ldg(zrow + C::F + k);
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        const int nv = l + C::LPH * i;
        if (nv < C::NV) ld_vec<C::VEC>(zrow + k * D + C::VEC * nv, zv + i * C::VEC);
      }
    }
    for (int e0 = beg; e0 < end; e0 += C::EPS) {   // warp-uniform trip count (shuffles inside)
      const int e = e0 + grp;
      const bool on = lane_on && e < end;
      float gv[C::NE];
#pragma unroll
      for (int i = 0; i < C::NE; ++i) gv[i] = 0.f;
      float mk = 0.f, dk = 1.f, sk = 0.f;
      int b = 0;
      if (on) {
        const int v = __ldg(nbr + e);
        b = __ldg(bin + e);
        const float* grow = g + (size_t)v * C::F;
        const float* st = stat + (size_t)v * 3 * H;
        mk = __ldg(st + k);
        dk = __ldg(st + H + k);
        sk = __ldg(st + 2 * H + k);
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          const int nv = l + C::LPH * i;
          if (nv < C::NV) ld_vec<C::VEC>(grow + k * D + C::VEC * nv, gv + i * C::VEC);
        }
      }
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < C::NE; ++i) part = fmaf(gv[i], zv[i], part);
      const float t = head_sum<C::LPH>(part, lane, l);
      if (on) {
        const float pre = pu + q_s[b * H + k];
        const float lg = pre > 0.f ? pre : HSG_LEAKY_SLOPE * pre;
        const float alpha = __expf(lg - mk) / dk;
        const float de = alpha * (t - sk);
        const float dpre = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
#pragma unroll
        for (int i = 0; i < C::NE; ++i) acc[i] = fmaf(alpha, gv[i], acc[i]);
        acc_dp += dpre;
        if (l == 0) my_dq[b * H + k] += dpre;
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
        const int nv = l + C::LPH * i;
        if (nv < C::NV) st_vec<C::VEC>(drow + k * D + C::VEC * nv, acc + i * C::VEC);
      }
      if (l == 0) drow[C::F + k] = acc_dp;
    }
    for (int c = C::F + H + lane; c < ldz; c += 32) drow[c] = 0.f;
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

__global__ void edge_bwd_dq_kernel(int nblocks, int nq, const float* __restrict__ dq_part, float* __restrict__ dq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += dq_part[(size_t)b * nq + i];
  dq[i] = s;
}

constexpr int EDGE_MAX_BLOCKS = 148 * 8;

static int edge_grid(int n_rows_steps) {
  int blocks = ceil_div(n_rows_steps, EDGE_WARPS);
  if (blocks > EDGE_MAX_BLOCKS) blocks = EDGE_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  return blocks;
}

// (H, D) instantiations: defaults (8,8) W2S and (6,50) S2W (HiGraph.py:57-76), plus
// other hidden/embedding sizes and the small shapes used by the test fixtures.
#define HSG_EDGE_CONFIGS(X) X(8, 8) X(6, 50) X(8, 16) X(6, 16) X(8, 32) X(6, 32) X(4, 4) X(6, 8) X(4, 16) X(1, 64)

template <int H, int D>
static int launch_fwd(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* origin, float* sh,
                      float* x, float* stat, cudaStream_t s) {
  LaunchScope ls(SLOT_EDGE_FWD, s);
  edge_fwd_kernel<H, D><<<edge_grid(c->n_dst), EDGE_THREADS, 0, s>>>(c->n_dst, c->indptr, c->nbr, c->bin, c->extra,
                                                                    zp, ldz, q, origin, sh, x, stat);
  return check_launch();
}

template <int H, int D>
static int launch_prep(int n_dst, const float* dx, const float* dsh, const float* sh, float* g, float* stat,
                       cudaStream_t s) {
  LaunchScope ls(SLOT_EDGE_BWD_PREP, s);
  edge_bwd_prep_kernel<H, D><<<edge_grid(ceil_div(n_dst, EdgeCfg<H, D>::EPS)), EDGE_THREADS, 0, s>>>(n_dst, dx, dsh,
                                                                                                   sh, g, stat);
  return check_launch();
}

template <int H, int D>
static int launch_bwd(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* g, const float* stat,
                      float* dzp, float* dq, float* ws, cudaStream_t s) {
  const int blocks = edge_grid(c->n_dst);
  {
    LaunchScope ls(SLOT_EDGE_BWD, s);
    edge_bwd_kernel<H, D><<<blocks, EDGE_THREADS, 0, s>>>(c->n_dst, c->indptr, c->nbr, c->bin, zp, ldz, q, g, stat,
                                                          dzp, ws);
    int rc = check_launch();
    if (rc) return rc;
  }
  LaunchScope ls(SLOT_EDGE_BWD_DQ, s);
  const int nq = HSG_N_BINS * H;
  edge_bwd_dq_kernel<<<ceil_div(nq, 128), 128, 0, s>>>(blocks, nq, ws, dq);
  return check_launch();
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_edge_fwd(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q, const float* origin,
                 float* sh, float* x, float* stat, void* stream) {
  if (!csc || !zp || !q || !sh || !stat || csc->n_dst < 0) return HSG_ERR_ARG;
  if (x != nullptr && origin == nullptr) return HSG_ERR_ARG;
  if (csc->n_dst == 0) return HSG_OK;
  if (!csc->indptr || (csc->n_edges > 0 && (!csc->nbr || !csc->bin))) return HSG_ERR_ARG;
  if (ldz % 4 != 0 || ldz < H * d + H || !aligned16(zp) || !aligned16(sh) || (x && (!aligned16(x) || !aligned16(origin))))
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

size_t hsg_edge_bwd_workspace_bytes(int H) { return (size_t)EDGE_MAX_BLOCKS * HSG_N_BINS * H * sizeof(float) + 16; }

int hsg_edge_bwd(const hsg_csc* csc_t, int H, int d, const float* zp, int ldz, const float* q, const float* g,
                 const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, void* stream) {
  if (!csc_t || !zp || !q || !g || !stat || !dzp || !dq || !ws || csc_t->n_dst < 0) return HSG_ERR_ARG;
  if (ws_bytes < hsg_edge_bwd_workspace_bytes(H)) return HSG_ERR_WORKSPACE;
  if (ldz % 4 != 0 || ldz < H * d + H || !aligned16(zp) || !aligned16(g) || !aligned16(dzp)) return HSG_ERR_ALIGN;
  if (csc_t->n_dst > 0 && (!csc_t->indptr || (csc_t->n_edges > 0 && (!csc_t->nbr || !csc_t->bin)))) return HSG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
#define X(HH, DD) \
  if (H == HH && d == DD) return launch_bwd<HH, DD>(csc_t, zp, ldz, q, g, stat, dzp, dq, (float*)ws, s);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

}  // extern "C"
