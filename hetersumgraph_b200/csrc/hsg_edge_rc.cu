// K5a': backward prep of the fused WSWGAT edge stage that RECOMPUTES sh instead of reading it.
//
// hsg_edge_bwd_prep streams dx AND sh (the pre-ELU aggregate the forward saved) to form g = dx * elu'(sh) and
// s = g . sh: on the S2W layer of a data-parallel shard (731 k word rows x 300 floats) the forward's store of sh and
// this read of it are 2 x 877 MB of DRAM traffic that SURVEY.md 8(d)'s B_fwd / B_bwd do not contain.  A word row has
// 1-3 in-edges and its sources are a handful of sentence rows that stay in L1 / L2, so sh_v = sum alpha_e z_u is
// cheaper to recompute from the saved softmax state (m, den) than to move: this kernel walks the FORWARD CSC exactly
// like edge_fwd_kernel (one warp per destination row, lane layout of hsg_edge_layout.cuh, index pipeline two rows
// deep), reads dx through a shared-memory transposition with the next row already in flight, gathers the first two source
// rows of the NEXT destination while the current one is worked on (the gather latency, not DRAM, bounded the
// first version), and writes g in the lane-interleaved layout + s.  hsg_edge_fwd is then called
// with sh = NULL.  Same g / s as hsg_edge_bwd_prep up to the rounding of the recomputed sh (alpha through
// ex2.approx; <= 2e-6 normalised, tested).  Layouts with one lane group per warp (EPS == 1) and F % 4 == 0.
#include <atomic>
#include <cstdlib>

#include "hsg_common.cuh"
#include "hsg_internal.cuh"
#include "hsg_edge_cfg.cuh"

namespace hsg {
namespace rc {

template <int H, int D>
__global__ void __launch_bounds__(EDGE_THREADS, 2)
edge_bwd_prep_rc_kernel(int n_dst, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                        const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                        const float* __restrict__ q, const float* __restrict__ dx, float* __restrict__ g,
                        float* __restrict__ stat) {
  pdl_prologue();
  using C = EdgeCfg<H, D>;
  static_assert(C::EPS == 1 && C::F % 4 == 0, "one lane group per warp, 16-byte rows");
  constexpr int F = C::F, FP = C::FP, NQ = HSG_N_BINS * H, NE = C::NE, VEC = C::VEC;
  constexpr int NV4 = (F / 4 + 31) / 32;                   // float4 per lane of a raw row
  constexpr int U = 2;                                     // gathered source rows in flight per warp
  __shared__ float q_s[NQ];
  __shared__ __align__(16) float stage[EDGE_WARPS * F];
  for (int i = threadIdx.x; i < NQ; i += EDGE_THREADS) q_s[i] = q[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;          // head owned by this lane
  const int l = gl % C::LPH;
  const bool lane_on = lane < C::GROUP;
  float* st_row = stage + wib * F;

  auto load_dx = [&](int vr, float4* d4) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c4 = lane + 32 * i;
      d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vr < n_dst && c4 < F / 4) d4[i] = __ldg(reinterpret_cast<const float4*>(dx + (size_t)vr * F) + c4);
    }
  };
  auto load_state = [&](int vr, float& mm, float& dd) {
    mm = 0.f;
    dd = 1.f;
    if (vr < n_dst && lane_on) {
      const float* st = stat + (size_t)vr * 3 * H;
      mm = st[k];
      dd = st[H + k];
    }
  };

  // gathers of the first U in-edges of a row (ids in the lanes of `ids` / `bins`, `deg` edges): issued one row ahead
  auto gather_first = [&](int ids, int bins, int deg, float (*zz)[NE], float* pp, int* bbv) {
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      const int u = __shfl_sync(0xffffffffu, ids, uu);
      bbv[uu] = __shfl_sync(0xffffffffu, bins, uu);
      pp[uu] = 0.f;
#pragma unroll
      for (int i = 0; i < NE; ++i) zz[uu][i] = 0.f;
      if (lane_on && uu < deg) {
        const float* row = zp + (size_t)u * ldz;
        pp[uu] = __ldg(row + FP + k);
#pragma unroll
        for (int i = 0; i < C::VPL; ++i)
          if (l + C::LPH * i < C::NV) ld_vec<VEC>(row + (i * C::GROUP + gl) * VEC, zz[uu] + i * VEC);
      }
    }
  };

  // software pipeline over this warp's rows v, v + nwarps, ...: row pointers THREE rows ahead, the first chunk of
  // neighbour ids two rows ahead, the first U source rows + dx + softmax state one row ahead
  int v = warp;
  int beg = 0, end = 0, begn = 0, endn = 0, beg2 = 0, end2 = 0, u0 = 0, b0 = 0, u0n = 0, b0n = 0;
  float4 dx0[NV4], dx1[NV4];
  float m_k, den_k, m_n, den_n;
  float zf[U][NE], pf[U];                                  // first U source rows of the current row
  int bf[U];
  auto load_ip = [&](int vr, int& b_, int& e_) {
    b_ = 0;
    e_ = 0;
    if (vr < n_dst) {
      b_ = __ldg(indptr + vr);
      e_ = __ldg(indptr + vr + 1);
    }
  };
  auto load_ids = [&](int b_, int e_, int& ids, int& bins) {
    ids = 0;
    bins = 0;
    if (b_ + lane < e_) {
      ids = __ldg(nbr + b_ + lane);
      bins = __ldg(bin + b_ + lane);
    }
  };
  load_ip(v, beg, end);
  load_ip(v + nwarps, begn, endn);
  load_ip(v + 2 * nwarps, beg2, end2);
  load_ids(beg, end, u0, b0);
  load_ids(begn, endn, u0n, b0n);
  load_dx(v, dx0);
  load_state(v, m_k, den_k);
  gather_first(u0, b0, end - beg, zf, pf, bf);
  while (v < n_dst) {
    int beg3, end3, u02, b02;
    load_ip(v + 3 * nwarps, beg3, end3);
    load_ids(beg2, end2, u02, b02);                         // ids of the row two ahead (its pointers are one iteration old)
    load_dx(v + nwarps, dx1);                               // next row's dx and softmax state
    load_state(v + nwarps, m_n, den_n);
    // this row's dx through shared memory into the lane layout
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < F / 4) *reinterpret_cast<float4*>(st_row + 4 * c4) = dx0[i];
    }
    const float m_c = m_k, rden_c = __fdividef(1.f, den_k);
    __syncwarp();
    float gv[NE], shv[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      gv[i] = 0.f;
      shv[i] = 0.f;
    }
    if (lane_on) {
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i < C::NV) {
          const float* p = st_row + k * D + VEC * (l + C::LPH * i);
#pragma unroll
          for (int t = 0; t < VEC; ++t) gv[i * VEC + t] = p[t];
        }
      }
    }
    __syncwarp();
    const int deg = end - beg;
    // the first U in-edges were gathered while the previous row was being worked on
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      if (lane_on && uu < deg) {
        const float a = exp_fast(leaky(pf[uu] + q_s[bf[uu] * H + k]) - m_c) * rden_c;
#pragma unroll
        for (int i = 0; i < NE; ++i) shv[i] = fmaf(a, zf[uu][i], shv[i]);
      }
    }
    if (deg > U) {
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
        for (int j0 = (c0 == beg ? U : 0); j0 < cnt; j0 += U) {   // warp-uniform trip count
          float zv[U][NE], pe[U];
          int bb[U];
          bool ok[U];
#pragma unroll
          for (int uu = 0; uu < U; ++uu) {                  // issue the gathers first
            const int j = j0 + uu;
            const int u = __shfl_sync(0xffffffffu, my_u, j & 31);
            bb[uu] = __shfl_sync(0xffffffffu, my_b, j & 31);
            ok[uu] = lane_on && j < cnt;
            pe[uu] = 0.f;
#pragma unroll
            for (int i = 0; i < NE; ++i) zv[uu][i] = 0.f;
            if (ok[uu]) {
              const float* row = zp + (size_t)u * ldz;
              pe[uu] = __ldg(row + FP + k);
#pragma unroll
              for (int i = 0; i < C::VPL; ++i)
                if (l + C::LPH * i < C::NV) ld_vec<VEC>(row + (i * C::GROUP + gl) * VEC, zv[uu] + i * VEC);
            }
          }
#pragma unroll
          for (int uu = 0; uu < U; ++uu) {
            if (ok[uu]) {
              const float a = exp_fast(leaky(pe[uu] + q_s[bb[uu] * H + k]) - m_c) * rden_c;
#pragma unroll
              for (int i = 0; i < NE; ++i) shv[i] = fmaf(a, zv[uu][i], shv[i]);
            }
          }
        }
      }
    }
    // the next row's first source rows start their way (its ids were fetched at the top of this iteration)
    gather_first(u0n, b0n, endn - begn, zf, pf, bf);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      gv[i] *= (shv[i] > 0.f ? 1.f : exp_fast(shv[i]));
      part = fmaf(gv[i], shv[i], part);
    }
    const float s = head_sum<C::LPH>(part, lane, l);
    if (lane_on) {
      float* out = g + (size_t)v * FP;
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        if (l + C::LPH * i >= C::NV) {
#pragma unroll
          for (int t = 0; t < VEC; ++t) gv[i * VEC + t] = 0.f;          // layout holes must be finite zeros
        }
        st_vec<VEC>(out + (i * C::GROUP + gl) * VEC, gv + i * VEC);
      }
      if (l == 0) stat[(size_t)v * 3 * H + 2 * H + k] = s;
    }
    v += nwarps;
    beg = begn;
    end = endn;
    begn = beg2;
    endn = end2;
    beg2 = beg3;
    end2 = end3;
    u0 = u0n;
    b0 = b0n;
    u0n = u02;
    b0n = b02;
#pragma unroll
    for (int i = 0; i < NV4; ++i) dx0[i] = dx1[i];
    m_k = m_n;
    den_k = den_n;
  }
}

static std::atomic<int> g_mode{-1};    // -1 auto (many destination rows), 0 never, 1 whenever the layout allows
constexpr int AUTO_MIN_ROWS = 65536;

#define HSG_EDGE_CONFIGS(X) \
  X(8, 8) X(6, 50) X(8, 16) X(6, 16) X(8, 32) X(6, 32) X(4, 4) X(6, 8) X(4, 16) X(1, 64) X(16, 4) X(2, 32) X(4, 32) X(12, 25)

template <int H, int D>
static constexpr bool shape_ok() {
  return EdgeCfg<H, D>::EPS == 1 && (H * D) % 4 == 0;
}

template <int H, int D>
static int launch(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* dx, float* g, float* stat,
                  cudaStream_t s) {
  if constexpr (shape_ok<H, D>()) {
    int blocks = ceil_div(c->n_dst, EDGE_WARPS);
    const int cap = num_sms() * 32;
    if (blocks > cap) blocks = cap;
    LaunchScope ls(SLOT_EDGE_BWD_PREP, s);
    launch_k(edge_bwd_prep_rc_kernel<H, D>, dim3(blocks), dim3(EDGE_THREADS), 0, s, c->n_dst, c->indptr, c->nbr, c->bin,
             zp, ldz, q, dx, g, stat);
    return check_launch();
  } else {
    return HSG_ERR_SHAPE;
  }
}

static bool applicable(int H, int d, int ldz) {
  if (ldz % 4 != 0) return false;
#define X(HH, DD) \
  if (H == HH && d == DD) return shape_ok<HH, DD>();
  HSG_EDGE_CONFIGS(X)
#undef X
  return false;
}

}  // namespace rc

bool edge_recompute_use(const hsg_csc* csc_fwd, int H, int d, int ldz) {
  const int mode = rc::g_mode.load(std::memory_order_relaxed);
  if (mode == 0 || !csc_fwd || !rc::applicable(H, d, ldz)) return false;
  // auto: many destination rows of LOW in-degree (word rows: ~1.5 sentences each).  Measured on the B200: at 1.45
  // in-edges per row forward + prep take 0.94 ms instead of 0.99 ms on the 2 048-graph shard; at 4 per row (stress
  // graph) recomputing costs more than re-reading sh (prep 0.33 ms against 0.16 ms)
  return mode == 1 || (csc_fwd->n_dst >= rc::AUTO_MIN_ROWS && (long long)csc_fwd->n_edges <= 2ll * csc_fwd->n_dst);
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_set_edge_recompute(int mode) {
  rc::g_mode.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_edge_bwd_prep_rc_ok(int H, int d, int ldz) { return rc::applicable(H, d, ldz) ? 1 : 0; }

int hsg_edge_bwd_prep_rc(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q, const float* dx,
                         float* g, float* stat, void* stream) {
  if (!csc || csc->n_dst < 0) return HSG_ERR_ARG;
  if (csc->n_dst == 0) return HSG_OK;
  if (!zp || !q || !dx || !g || !stat) return HSG_ERR_ARG;
  if (!csc->indptr || (csc->n_edges > 0 && (!csc->nbr || !csc->bin))) return HSG_ERR_ARG;
  if (!aligned16(zp) || !aligned16(dx) || !aligned16(g)) return HSG_ERR_ALIGN;
  if (!rc::applicable(H, d, ldz)) return HSG_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
#define X(HH, DD) \
  if (H == HH && d == DD) return rc::launch<HH, DD>(csc, zp, ldz, q, dx, g, stat, s);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

}  // extern "C"
