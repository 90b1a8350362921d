// K5-seg: SEGMENT-RESIDENT backward of the fused WSWGAT edge stage for layers whose SOURCE side is small per graph
// (S2W, module/GATLayer.py:127-152: a document's <= ~55 sentence / document nodes feed its few hundred words).
//
// A batch is a disjoint union of graphs (dgl.batch, dataloader.py:480): rows and edges of graph g never mix with
// those of another graph.  The general backward (hsg_edge.cu) ignores that: edge_bwd_prep streams dx and sh and
// writes g = dx * elu'(sh) (three passes over the [n_dst, F] side), then edge_bwd gathers the g rows again per
// source row - 4 x n_dst x F x 4 bytes of DRAM traffic plus the forward's store of sh, against the ONE compulsory
// read of dx in SURVEY.md 8(d)'s B_bwd.  Here a CTA owns whole graphs:
//   * the graph's source rows [z | p] (<= 55 x 1.2 KB) come into shared memory with ONE bulk (TMA) copy and its
//     accumulators [dz | dp] live next to them; they leave with coalesced stores when the graph is finished;
//   * the destination side streams through a ring of 16-row tiles of dx filled by cp.async.bulk + mbarrier several
//     tiles ahead (rows of consecutive destinations are contiguous: a tile is ONE 19 KB copy, no tensor map);
//   * phase A, one warp per destination row (same lane mapping as hsg_edge.cu): sh_v is RECOMPUTED from the saved
//     softmax state (m, den) and the source rows in shared memory (the forward does not store sh at all),
//     g = dx * elu'(sh) overwrites the row in place, s = g . sh, and per in-edge alpha, t = g_v . z_u,
//     dpre = leaky'(pre) alpha (t - s) go to a small record list;
//   * phase B, one thread per column of the [dz | dp] row: walks the tile's records IN ORDER and adds
//     alpha_e g_v[c] into the accumulator of source u_e - a fixed summation order, no atomics, bitwise
//     reproducible whatever the grid.  dq is a per-thread private column of a shared table, reduced over CTAs by
//     edge_bwd_dq_kernel in block order.
// DRAM traffic: dx once, [z | p] once, [dz | dp] once, the CSC once - the survey's B_bwd.
#include <atomic>
#include <cstdlib>

#include "hsg_common.cuh"
#include "hsg_internal.cuh"
#include "hsg_edge_cfg.cuh"

namespace hsg {
namespace seg {

constexpr int T = 16;                  // destination rows per tile = warps per CTA
constexpr int WARPS = 16;
constexpr int THREADS = WARPS * 32;
constexpr int CAP = 96;                // edge records per round (a tile with more edges takes several rounds)
constexpr int EXT = 64;                // segments per CTA (extent table)
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Layout {
  int q, dq, ext, rec_u, rec_vb, rec_a, rec_d, zs, dzs, tiles, total;
};

__host__ __device__ inline int up16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline Layout make_layout(int H, int F, int ldz, int cap_src, int nstage) {
  Layout L;
  int o = 128;                                            // mbarriers
  L.q = o;      o += up16(HSG_N_BINS * H * 4);
  L.dq = o;     o += up16(HSG_N_BINS * H * 4);
  L.ext = o;    o += EXT * 4 * 4;
  L.rec_u = o;  o += 2 * CAP * 4;
  L.rec_vb = o; o += 2 * CAP * 4;
  L.rec_a = o;  o += up16(2 * CAP * H * 4);
  L.rec_d = o;  o += up16(2 * CAP * H * 4);
  o = (o + 127) & ~127;
  L.zs = o;     o += cap_src * ldz * 4;
  L.dzs = o;    o += cap_src * ldz * 4;
  o = (o + 127) & ~127;
  L.tiles = o;  o += nstage * T * F * 4;
  L.total = o;
  return L;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1ull << 24)) __trap();        // never hang the GPU: fail loudly instead
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Cursor {
  int gi, t;
};

template <int H, int D>
__global__ void __launch_bounds__(THREADS, 1)
edge_bwd_seg_kernel(int n_seg, const int32_t* __restrict__ seg_dst, const int32_t* __restrict__ seg_src, int cap_src,
                    int nstage, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                    const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                    const float* __restrict__ q, const float* __restrict__ dx, const float* __restrict__ stat,
                    float* __restrict__ dzp, float* __restrict__ dq_part) {
  using C = EdgeCfg<H, D>;
  static_assert(C::EPS == 1 && C::F % 4 == 0, "one lane group per warp, 16-byte rows");
  constexpr int F = C::F, FP = C::FP, NQ = HSG_N_BINS * H, NE = C::NE;
  extern __shared__ __align__(128) unsigned char smem[];
  const Layout L = make_layout(H, F, ldz, cap_src, nstage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);     // [0, nstage): tile landed; [MAX_STAGES]: source rows landed
  float* q_s = reinterpret_cast<float*>(smem + L.q);
  float* dq_s = reinterpret_cast<float*>(smem + L.dq);
  int* ext = reinterpret_cast<int*>(smem + L.ext);        // per segment of this CTA: w0, nw, s0, ns
  int* rec_u = reinterpret_cast<int*>(smem + L.rec_u);
  int* rec_vb = reinterpret_cast<int*>(smem + L.rec_vb);
  float* rec_a = reinterpret_cast<float*>(smem + L.rec_a);
  float* rec_d = reinterpret_cast<float*>(smem + L.rec_d);
  float* zs = reinterpret_cast<float*>(smem + L.zs);
  float* dzs = reinterpret_cast<float*>(smem + L.dzs);
  float* tiles = reinterpret_cast<float*>(smem + L.tiles);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gl = lane % C::GROUP;
  const int k = gl / C::LPH;          // head owned by this lane (phase A)
  const int l = gl % C::LPH;
  const bool lane_on = lane < C::GROUP;
  const uint32_t bar_z = smem_u32(bars + MAX_STAGES);

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(smem_u32(bars + i), 1);
    mbar_init(bar_z, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < NQ; i += THREADS) dq_s[i] = 0.f;
  for (int i = tid; i < cap_src * ldz; i += THREADS) dzs[i] = 0.f;
  pdl_prologue();
  for (int i = tid; i < NQ; i += THREADS) q_s[i] = q[i];
  const int n_mine = (int)blockIdx.x < n_seg ? (n_seg - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  for (int i = tid; i < n_mine; i += THREADS) {
    const int g = blockIdx.x + i * gridDim.x;
    const int w0 = __ldg(seg_dst + g), s0 = __ldg(seg_src + g);
    ext[4 * i] = w0;
    ext[4 * i + 1] = __ldg(seg_dst + g + 1) - w0;
    ext[4 * i + 2] = s0;
    ext[4 * i + 3] = __ldg(seg_src + g + 1) - s0;
  }
  __syncthreads();
  // rows of segment i that take part: none when it has no sources, no destinations, or more sources than fit
  auto eff_nw = [&](int i) {
    const int ns = ext[4 * i + 3];
    return (ns > 0 && ns <= cap_src) ? ext[4 * i + 1] : 0;
  };
  // source rows no tile will reach: zero gradient, or NaN when the caller's bound seg_max_src was wrong
  for (int i = 0; i < n_mine; ++i) {
    const int nw = ext[4 * i + 1], s0 = ext[4 * i + 2], ns = ext[4 * i + 3];
    if (ns > 0 && eff_nw(i) <= 0) {
      const float fill = (ns > cap_src && nw > 0) ? CUDART_NAN_F : 0.f;
      float* out = dzp + (size_t)s0 * ldz;
      for (int j = tid; j < ns * ldz; j += THREADS) out[j] = fill;
    }
  }
  auto c_norm = [&](Cursor& c) {
    while (c.gi < n_mine && c.t * T >= eff_nw(c.gi)) {
      ++c.gi;
      c.t = 0;
    }
  };
  auto c_next = [&](Cursor& c) {
    ++c.t;
    c_norm(c);
  };

  // ---- producer (thread 0): dx tiles, nstage ahead of the consumer; source rows one segment ahead ----
  Cursor cp{0, 0};
  int tp = 0;
  c_norm(cp);
  auto issue_tile = [&]() {
    if (cp.gi >= n_mine) return;
    const int w0 = ext[4 * cp.gi], nw = ext[4 * cp.gi + 1];
    const int rows = min(T, nw - cp.t * T);
    const uint32_t bytes = (uint32_t)rows * F * 4u;
    const int slot = tp % nstage;
    const uint32_t bar = smem_u32(bars + slot);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(tiles + (size_t)slot * T * F), dx + (size_t)(w0 + cp.t * T) * F, bytes, bar);
    ++tp;
    c_next(cp);
  };
  auto issue_sources = [&](int gi) {
    const int s0 = ext[4 * gi + 2], ns = ext[4 * gi + 3];
    const uint32_t bytes = (uint32_t)ns * ldz * 4u;
    mbar_expect_tx(bar_z, bytes);
    bulk_g2s(smem_u32(zs), zp + (size_t)s0 * ldz, bytes, bar_z);
  };

  // ---- index pipeline of this warp's row: row pointers two tiles ahead, first neighbours + softmax state one ----
  // quad (one value per lane): lane 0/1 = indptr[v], indptr[v+1] of this warp's row, lane 2/3 = first / last edge
  // of the tile
  auto load_quad = [&](const Cursor& c) {
    int val = 0;
    if (c.gi < n_mine && lane < 4) {
      const int w0 = ext[4 * c.gi], nw = ext[4 * c.gi + 1];
      const int v0 = w0 + c.t * T, rows = min(T, nw - c.t * T);
      if (lane >= 2)
        val = __ldg(indptr + v0 + (lane == 2 ? 0 : rows));
      else if (w < rows)
        val = __ldg(indptr + v0 + w + lane);
    }
    return val;
  };
  struct Pre {
    int u, b;
    float m, den;
  };
  auto load_edges = [&](int quad, const Cursor& c) {
    Pre p{0, 0, 0.f, 1.f};
    const int ip0 = __shfl_sync(0xffffffffu, quad, 0), ip1 = __shfl_sync(0xffffffffu, quad, 1);
    if (c.gi < n_mine) {
      const int w0 = ext[4 * c.gi], nw = ext[4 * c.gi + 1];
      const int rows = min(T, nw - c.t * T);
      if (w < rows) {
        if (ip0 + lane < ip1) {
          p.u = __ldg(nbr + ip0 + lane);
          p.b = __ldg(bin + ip0 + lane);
        }
        if (lane_on) {
          const float* st = stat + (size_t)(w0 + c.t * T + w) * 3 * H;
          p.m = __ldg(st + k);
          p.den = __ldg(st + H + k);
        }
      }
    }
    return p;
  };

  Cursor cc{0, 0};
  c_norm(cc);
  Cursor c1 = cc;
  if (c1.gi < n_mine) c_next(c1);
  Cursor c2 = c1;
  if (c2.gi < n_mine) c_next(c2);
  if (tid == 0 && cc.gi < n_mine) {
    fence_async_smem();
    issue_sources(cc.gi);
    for (int i = 0; i < nstage; ++i) issue_tile();
  }
  int quad_c = load_quad(cc), quad_1 = load_quad(c1);
  Pre pre_c = load_edges(quad_c, cc);

  // phase-B identity of this thread: column tid of a [dz | dp] row
  int pk = 0, pcol = -1;
  if (tid < FP) {
    const int slab = C::GROUP * C::VEC;
    const int pi = tid / slab, rem = tid % slab;
    const int pgl = rem / C::VEC, pt = rem % C::VEC;
    pk = pgl / C::LPH;
    const int nv = pgl % C::LPH + C::LPH * pi;
    if (nv < C::NV) pcol = pk * D + nv * C::VEC + pt;
  } else if (tid < FP + H) {
    pk = tid - FP;
  }

  int tc = 0, buf = 0;
  uint32_t zpar = 0;
  while (cc.gi < n_mine) {
    const int quad_2 = load_quad(c2);
    const Pre pre_1 = load_edges(quad_1, c1);
    const int nw = ext[4 * cc.gi + 1], s0 = ext[4 * cc.gi + 2], ns = ext[4 * cc.gi + 3];
    const int rows = min(T, nw - cc.t * T);
    const bool last_tile = (cc.t + 1) * T >= nw;
    if (cc.t == 0) {                                       // this segment's source rows have landed
      mbar_wait(bar_z, zpar);
      zpar ^= 1u;
    }
    const int slot = tc % nstage;
    mbar_wait(smem_u32(bars + slot), (uint32_t)(tc / nstage) & 1u);
    float* tile = tiles + (size_t)slot * T * F;

    const int ip0 = __shfl_sync(0xffffffffu, quad_c, 0), ip1 = __shfl_sync(0xffffffffu, quad_c, 1);
    const int te0 = __shfl_sync(0xffffffffu, quad_c, 2), te1 = __shfl_sync(0xffffffffu, quad_c, 3);
    const int n_te = te1 - te0;
    const int rounds = n_te > CAP ? ceil_div(n_te, CAP) : 1;
    const bool row_on = w < rows;
    float* trow = tile + w * F;
    float gv[NE];
    float s_head = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) gv[i] = 0.f;

    // ---- phase A1: sh_v recomputed, g = dx * elu'(sh) in place, s = g . sh ----
    if (row_on) {
      float shv[NE];
#pragma unroll
      for (int i = 0; i < NE; ++i) shv[i] = 0.f;
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            const float* p = trow + k * D + C::VEC * (l + C::LPH * i);
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) gv[i * C::VEC + t] = p[t];
          }
        }
      }
      for (int c0 = ip0; c0 < ip1; c0 += 32) {
        const int cnt = min(32, ip1 - c0);
        int my_u = pre_c.u, my_b = pre_c.b;
        if (c0 != ip0) {
          my_u = 0;
          my_b = 0;
          if (lane < cnt) {
            my_u = __ldg(nbr + c0 + lane);
            my_b = __ldg(bin + c0 + lane);
          }
        }
        for (int j = 0; j < cnt; ++j) {
          const int u = min(max(__shfl_sync(0xffffffffu, my_u, j) - s0, 0), ns - 1);
          const int b = __shfl_sync(0xffffffffu, my_b, j);
          if (lane_on) {
            const float* zrow = zs + u * ldz;
            const float pre = zrow[FP + k] + q_s[b * H + k];
            const float a = __fdividef(__expf(leaky(pre) - pre_c.m), pre_c.den);
#pragma unroll
            for (int i = 0; i < C::VPL; ++i) {
              if (l + C::LPH * i < C::NV) {
                const float* p = zrow + (i * C::GROUP + gl) * C::VEC;
#pragma unroll
                for (int t = 0; t < C::VEC; ++t) shv[i * C::VEC + t] = fmaf(a, p[t], shv[i * C::VEC + t]);
              }
            }
          }
        }
      }
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        gv[i] *= (shv[i] > 0.f ? 1.f : __expf(shv[i]));
        part = fmaf(gv[i], shv[i], part);
      }
      s_head = head_sum<C::LPH>(part, lane, l);
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            float* p = trow + k * D + C::VEC * (l + C::LPH * i);
#pragma unroll
            for (int t = 0; t < C::VEC; ++t) p[t] = gv[i * C::VEC + t];
          }
        }
      }
    }

    for (int r = 0; r < rounds; ++r) {
      if (r > 0) __syncthreads();                          // the previous round's records have been consumed
      const int win0 = te0 + r * CAP;
      int* ru = rec_u + buf * CAP;
      int* rvb = rec_vb + buf * CAP;
      float* ra = rec_a + buf * CAP * H;
      float* rd = rec_d + buf * CAP * H;
      // ---- phase A2: per in-edge alpha, dpre of this round's window ----
      if (row_on) {
        for (int c0 = ip0; c0 < ip1; c0 += 32) {
          const int cnt = min(32, ip1 - c0);
          if (c0 + cnt <= win0 || c0 >= win0 + CAP) continue;
          int my_u = pre_c.u, my_b = pre_c.b;
          if (c0 != ip0) {
            my_u = 0;
            my_b = 0;
            if (lane < cnt) {
              my_u = __ldg(nbr + c0 + lane);
              my_b = __ldg(bin + c0 + lane);
            }
          }
          for (int j = 0; j < cnt; ++j) {
            const int ei = c0 + j - win0;
            if (ei < 0 || ei >= CAP) continue;             // warp-uniform
            const int u = min(max(__shfl_sync(0xffffffffu, my_u, j) - s0, 0), ns - 1);
            const int b = __shfl_sync(0xffffffffu, my_b, j);
            const float* zrow = zs + u * ldz;
            float part = 0.f, pre = 0.f;
            if (lane_on) {
              pre = zrow[FP + k] + q_s[b * H + k];
#pragma unroll
              for (int i = 0; i < C::VPL; ++i) {
                if (l + C::LPH * i < C::NV) {
                  const float* p = zrow + (i * C::GROUP + gl) * C::VEC;
#pragma unroll
                  for (int t = 0; t < C::VEC; ++t) part = fmaf(gv[i * C::VEC + t], p[t], part);
                }
              }
            }
            const float tdot = head_sum<C::LPH>(part, lane, l);
            if (lane_on && l == 0) {
              const float a = __fdividef(__expf(leaky(pre) - pre_c.m), pre_c.den);
              const float de = a * (tdot - s_head);
              ra[ei * H + k] = a;
              rd[ei * H + k] = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
            }
            if (lane == 0) {
              ru[ei] = u;
              rvb[ei] = w | (b << 8);
            }
          }
        }
      }
      __syncthreads();
      if (tid == 0) {
        // every thread has left phase B of the previous tile: its slot takes the tile nstage ahead; after the last
        // phase A of a segment nobody reads its source rows any more: the next segment's rows start their way
        fence_async_smem();
        if (r == 0 && tc > 0) issue_tile();
        if (r == rounds - 1 && last_tile && c1.gi < n_mine) issue_sources(c1.gi);
      }
      // ---- phase B: column threads walk the records in order ----
      const int n_rec = min(CAP, n_te - r * CAP);
      if (tid < FP) {
        if (pcol >= 0) {
          int cur = -1;
          float acc = 0.f;
          for (int e = 0; e < n_rec; ++e) {
            const int u = ru[e];
            const float val = ra[e * H + pk] * tile[(rvb[e] & 0xff) * F + pcol];
            if (u != cur) {
              if (cur >= 0) dzs[cur * ldz + tid] += acc;
              cur = u;
              acc = val;
            } else {
              acc += val;
            }
          }
          if (cur >= 0) dzs[cur * ldz + tid] += acc;
        }
      } else if (tid < FP + H) {
        int cur = -1;
        float acc = 0.f;
        for (int e = 0; e < n_rec; ++e) {
          const int u = ru[e];
          const float val = rd[e * H + pk];
          dq_s[(rvb[e] >> 8) * H + pk] += val;
          if (u != cur) {
            if (cur >= 0) dzs[cur * ldz + tid] += acc;
            cur = u;
            acc = val;
          } else {
            acc += val;
          }
        }
        if (cur >= 0) dzs[cur * ldz + tid] += acc;
      }
    }

    if (last_tile) {                                       // [dz | dp | 0] of this segment's source rows
      __syncthreads();
      float4* acc4 = reinterpret_cast<float4*>(dzs);
      float4* out4 = reinterpret_cast<float4*>(dzp + (size_t)s0 * ldz);
      const int n4 = ns * ldz / 4;
      for (int j = tid; j < n4; j += THREADS) {
        out4[j] = acc4[j];
        acc4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cc = c1;
    c1 = c2;
    if (c2.gi < n_mine) c_next(c2);
    quad_c = quad_1;
    quad_1 = quad_2;
    pre_c = pre_1;
    ++tc;
    buf ^= 1;
  }
  __syncthreads();
  for (int i = tid; i < NQ; i += THREADS) dq_part[(size_t)blockIdx.x * NQ + i] = dq_s[i];
}

static std::atomic<int> g_mode{-1};    // -1 auto (many segments), 0 never, 1 whenever the layout allows
constexpr int AUTO_MIN_SEGMENTS = 512;

template <int H, int D>
static int pick_stages(int ldz, int cap_src) {
  for (int n = MAX_STAGES; n >= 2; --n)
    if (make_layout(H, H * D, ldz, cap_src, n).total <= SMEM_LIMIT) return n;
  return 0;
}

template <int H, int D>
static bool shape_ok(int ldz, int cap_src) {
  if constexpr (EdgeCfg<H, D>::EPS == 1 && (H * D) % 4 == 0)
    return ldz % 4 == 0 && pick_stages<H, D>(ldz, cap_src) >= 2;
  else
    return false;
}

template <int H, int D>
static int launch(const hsg_csc* c, const float* zp, int ldz, const float* q, const float* dx, const float* stat,
                  float* dzp, float* dq, float* ws, int accumulate_dq, cudaStream_t s) {
  if constexpr (EdgeCfg<H, D>::EPS == 1 && (H * D) % 4 == 0) {
    const int cap_src = c->seg_max_src;
    const int nstage = pick_stages<H, D>(ldz, cap_src);
    if (nstage < 2) return HSG_ERR_SHAPE;
    const int smem_bytes = make_layout(H, H * D, ldz, cap_src, nstage).total;
    static int attr_bytes = 0;
    if (smem_bytes > attr_bytes) {
      if (cudaFuncSetAttribute(edge_bwd_seg_kernel<H, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) !=
          cudaSuccess)
        return HSG_ERR_CUDA;
      attr_bytes = SMEM_LIMIT;
    }
    int grid = min(c->n_seg, num_sms());
    if (ceil_div(c->n_seg, grid) > EXT) grid = ceil_div(c->n_seg, EXT);
    {
      LaunchScope ls(SLOT_EDGE_BWD, s);
      launch_k(edge_bwd_seg_kernel<H, D>, dim3(grid), dim3(THREADS), (size_t)smem_bytes, s, c->n_seg, c->seg_dst_ptr,
               c->seg_src_ptr, cap_src, nstage, c->indptr, c->nbr, c->bin, zp, ldz, q, dx, stat, dzp, ws);
      const int rc = check_launch();
      if (rc) return rc;
    }
    return edge_dq_reduce(grid, HSG_N_BINS * H, ws, dq, accumulate_dq, s);
  } else {
    return HSG_ERR_SHAPE;
  }
}

#define HSG_EDGE_CONFIGS(X) \
  X(8, 8) X(6, 50) X(8, 16) X(6, 16) X(8, 32) X(6, 32) X(4, 4) X(6, 8) X(4, 16) X(1, 64) X(16, 4) X(2, 32) X(4, 32) X(12, 25)

static bool applicable(const hsg_csc* c, int H, int d, int ldz) {
  if (!c || c->n_seg <= 0 || !c->seg_dst_ptr || !c->seg_src_ptr || c->seg_max_src <= 0) return false;
  if (ceil_div(c->n_seg, EXT) > 148 * 32) return false;       // dq partial workspace rows (EDGE_MAX_BLOCKS)
#define X(HH, DD) \
  if (H == HH && d == DD) return shape_ok<HH, DD>(ldz, c->seg_max_src);
  HSG_EDGE_CONFIGS(X)
#undef X
  return false;
}

}  // namespace seg

bool edge_bwd_seg_use(const hsg_csc* csc_fwd, int H, int d, int ldz) {
  const int mode = seg::g_mode.load(std::memory_order_relaxed);
  if (mode == 0 || !seg::applicable(csc_fwd, H, d, ldz)) return false;
  return mode == 1 || csc_fwd->n_seg >= seg::AUTO_MIN_SEGMENTS;
}

int edge_bwd_seg_ex(const hsg_csc* c, int H, int d, const float* zp, int ldz, const float* q, const float* dx,
                    const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, int accumulate_dq,
                    cudaStream_t s) {
  if (!c || !zp || !q || !dx || !stat || !dzp || !dq || !ws || c->n_dst < 0) return HSG_ERR_ARG;
  if (ws_bytes < hsg_edge_bwd_workspace_bytes(H)) return HSG_ERR_WORKSPACE;
  if (!aligned16(zp) || !aligned16(dx) || !aligned16(dzp)) return HSG_ERR_ALIGN;
  if (!seg::applicable(c, H, d, ldz)) return HSG_ERR_SHAPE;
  if (!c->indptr || (c->n_edges > 0 && (!c->nbr || !c->bin))) return HSG_ERR_ARG;
#define X(HH, DD) \
  if (H == HH && d == DD) return seg::launch<HH, DD>(c, zp, ldz, q, dx, stat, dzp, dq, (float*)ws, accumulate_dq, s);
  HSG_EDGE_CONFIGS(X)
#undef X
  return HSG_ERR_SHAPE;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_set_edge_seg(int mode) {
  seg::g_mode.store(mode < 0 ? -1 : (mode ? 1 : 0));
  return HSG_OK;
}

int hsg_edge_bwd_seg_ok(const hsg_csc* csc, int H, int d, int ldz) { return seg::applicable(csc, H, d, ldz) ? 1 : 0; }

int hsg_edge_bwd_seg(const hsg_csc* csc, int H, int d, const float* zp, int ldz, const float* q, const float* dx,
                     const float* stat, float* dzp, float* dq, void* ws, size_t ws_bytes, void* stream) {
  return edge_bwd_seg_ex(csc, H, d, zp, ldz, q, dx, stat, dzp, dq, ws, ws_bytes, 0, (cudaStream_t)stream);
}

}  // extern "C"
