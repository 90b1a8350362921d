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

constexpr int T = 16;                  // destination rows per tile = phase-A warps
constexpr int A_WARPS = 16;            // one destination row of the tile each
constexpr int CAP = 96;                // edge records per round (a tile with more edges takes several rounds)
constexpr int NR = 3;                  // record buffers (phase A runs at most NR rounds ahead of phase B)
constexpr int EXT = 64;                // segments per CTA (extent table)
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;

// mbarrier slots
constexpr int BAR_FULL = 0;                       // [MAX_STAGES] tile landed                (TMA -> A)
constexpr int BAR_EMPTY = MAX_STAGES;             // [MAX_STAGES] tile consumed              (B -> producer)
constexpr int BAR_READY = 2 * MAX_STAGES;         // [NR] records + g of a round written     (A -> B)
constexpr int BAR_CONSUMED = BAR_READY + NR;      // [NR] records of a round consumed        (B -> A)
constexpr int BAR_ZFULL = BAR_CONSUMED + NR;      // source rows of a segment landed         (TMA -> A)
constexpr int BAR_ZFREE = BAR_ZFULL + 1;          // every A warp has left the segment       (A -> producer)
constexpr int BAR_COUNT = BAR_ZFREE + 1;

struct Layout {
  int q, dq, ext, hdr, rec_u, rec_vb, rec_a, rec_d, zs, dzs, tiles, total;
};

__host__ __device__ inline int up16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline Layout make_layout(int H, int F, int ldz, int cap_src, int nstage) {
  Layout L;
  int o = 256;                                            // mbarriers
  L.q = o;      o += up16(HSG_N_BINS * H * 4);
  L.dq = o;     o += up16(3 * HSG_N_BINS * H * 4);
  L.ext = o;    o += EXT * 4 * 4;
  L.hdr = o;    o += NR * 16;
  L.rec_u = o;  o += NR * CAP * 4;
  L.rec_vb = o; o += NR * CAP * 4;
  L.rec_a = o;  o += up16(NR * CAP * H * 4);
  L.rec_d = o;  o += up16(NR * CAP * H * 4);
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned long long spins = 0;
  while (!mbar_test(bar, parity))
    if (++spins > (1ull << 24)) __trap();                 // never hang the GPU: fail loudly instead
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// shared-memory vector access of a lane's VEC consecutive floats (addresses are VEC*4-byte aligned by construction)
template <int VEC>
__device__ __forceinline__ void lds_vec(const float* p, float* out) {
  if (VEC == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if (VEC == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    out[0] = v.x; out[1] = v.y;
  } else {
    out[0] = p[0];
  }
}
template <int VEC>
__device__ __forceinline__ void sts_vec(float* p, const float* v) {
  if (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    p[0] = v[0];
  }
}

// position in the flat tile sequence of a CTA: segment gi (index into this CTA's extent table), tile t inside it;
// v0 = first destination row of the tile, nwc = destination rows of the segment (cached from the table)
struct Cursor {
  int gi, t, v0, nwc;
};

template <int H, int D>
struct SegCfg {
  static constexpr int B_WARPS = 3;                                // phase B: source row u belongs to warp u % 3
  static constexpr int WARPS = A_WARPS + B_WARPS + 1;              // + the producer warp: 20 warps -> 96 registers
  static constexpr int THREADS = WARPS * 32;                       //   (warps are allocated in fours)
};

// exp(x) for x <= ~0 as ONE multiply + MUFU.EX2 (flush-to-zero; __expf adds range fix-ups this path never needs)
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

template <int H, int D>
__global__ void __launch_bounds__(SegCfg<H, D>::THREADS, 1)
edge_bwd_seg_kernel(int n_seg, const int32_t* __restrict__ seg_dst, const int32_t* __restrict__ seg_src, int cap_src,
                    int nstage, const int32_t* __restrict__ indptr, const int32_t* __restrict__ nbr,
                    const uint8_t* __restrict__ bin, const float* __restrict__ zp, int ldz,
                    const float* __restrict__ q, const float* __restrict__ dx, const float* __restrict__ stat,
                    float* __restrict__ dzp, float* __restrict__ dq_part) {
  using C = EdgeCfg<H, D>;
  using S = SegCfg<H, D>;
  static_assert(C::EPS == 1 && C::F % 4 == 0, "one lane group per warp, 16-byte rows");
  constexpr int F = C::F, FP = C::FP, NQ = HSG_N_BINS * H, NE = C::NE, VEC = C::VEC;
  constexpr int THREADS = S::THREADS, B_WARPS = S::B_WARPS;
  extern __shared__ __align__(128) unsigned char smem[];
  const Layout L = make_layout(H, F, ldz, cap_src, nstage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float* q_s = reinterpret_cast<float*>(smem + L.q);
  float* dq_s = reinterpret_cast<float*>(smem + L.dq);
  int* ext = reinterpret_cast<int*>(smem + L.ext);        // per segment of this CTA: w0, nw (0: skipped), s0, ns
  int* hdr = reinterpret_cast<int*>(smem + L.hdr);        // per record buffer: edges of the tile
  int* rec_u = reinterpret_cast<int*>(smem + L.rec_u);
  int* rec_vb = reinterpret_cast<int*>(smem + L.rec_vb);
  float* rec_a = reinterpret_cast<float*>(smem + L.rec_a);
  float* rec_d = reinterpret_cast<float*>(smem + L.rec_d);
  float* zs = reinterpret_cast<float*>(smem + L.zs);
  float* dzs = reinterpret_cast<float*>(smem + L.dzs);
  float* tiles = reinterpret_cast<float*>(smem + L.tiles);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  auto bar = [&](int i) { return smem_u32(bars + i); };

  if (tid == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(bar(BAR_FULL + i), 1);
      mbar_init(bar(BAR_EMPTY + i), B_WARPS);
    }
    for (int i = 0; i < NR; ++i) {
      mbar_init(bar(BAR_READY + i), A_WARPS);
      mbar_init(bar(BAR_CONSUMED + i), B_WARPS);
    }
    mbar_init(bar(BAR_ZFULL), 1);
    mbar_init(bar(BAR_ZFREE), A_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < B_WARPS * NQ; i += THREADS) dq_s[i] = 0.f;
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
  // A segment takes part when it has sources that fit and destinations.  Source rows no tile will reach get a
  // zero gradient - or NaN when the caller's bound seg_max_src was wrong (fail loudly, never out of bounds).
  for (int i = 0; i < n_mine; ++i) {
    const int nw = ext[4 * i + 1], s0 = ext[4 * i + 2], ns = ext[4 * i + 3];
    if (ns > 0 && (ns > cap_src || nw <= 0)) {
      const float fill = (ns > cap_src && nw > 0) ? CUDART_NAN_F : 0.f;
      float* out = dzp + (size_t)s0 * ldz;
      for (int j = tid; j < ns * ldz; j += THREADS) out[j] = fill;
    }
  }
  __syncthreads();
  for (int i = tid; i < n_mine; i += THREADS) {
    const int ns = ext[4 * i + 3];
    if (!(ns > 0 && ns <= cap_src) || ext[4 * i + 1] < 0) ext[4 * i + 1] = 0;
  }
  __syncthreads();
  // every role walks the same flat sequence of tiles
  auto c_seek = [&](Cursor& c) {                           // first segment at or after c.gi that takes part
    while (c.gi < n_mine && ext[4 * c.gi + 1] <= 0) ++c.gi;
    c.t = 0;
    c.v0 = c.gi < n_mine ? ext[4 * c.gi] : 0;
    c.nwc = c.gi < n_mine ? ext[4 * c.gi + 1] : 0;
  };
  auto c_next = [&](Cursor& c) {
    ++c.t;
    c.v0 += T;
    if (c.t * T >= c.nwc) {
      ++c.gi;
      c_seek(c);
    }
  };
  auto c_rows = [&](const Cursor& c) { return min(T, c.nwc - c.t * T); };

  if (w == A_WARPS + B_WARPS) {
    // =========================== producer: one thread feeds the TMA engine ===========================
    if (lane == 0) {
      Cursor cp{0, 0, 0, 0};
      c_seek(cp);
      int zgi = cp.gi;                                     // next segment whose source rows are to be loaded
      int slot = 0, zc = 0;
      uint32_t use_par = 0;                                // parity of the slot's previous use
      bool first_lap = true;
      while (cp.gi < n_mine || zgi < n_mine) {
        bool progress = false;
        if (zgi < n_mine && (zc == 0 || mbar_test(bar(BAR_ZFREE), (uint32_t)(zc - 1) & 1u))) {
          const int s0 = ext[4 * zgi + 2], ns = ext[4 * zgi + 3];
          const uint32_t bytes = (uint32_t)ns * ldz * 4u;
          fence_async_smem();
          mbar_expect_tx(bar(BAR_ZFULL), bytes);
          bulk_g2s(smem_u32(zs), zp + (size_t)s0 * ldz, bytes, bar(BAR_ZFULL));
          ++zc;
          ++zgi;
          while (zgi < n_mine && ext[4 * zgi + 1] <= 0) ++zgi;
          progress = true;
        }
        if (cp.gi < n_mine && (first_lap || mbar_test(bar(BAR_EMPTY + slot), use_par))) {
          const uint32_t bytes = (uint32_t)c_rows(cp) * F * 4u;
          fence_async_smem();
          mbar_expect_tx(bar(BAR_FULL + slot), bytes);
          bulk_g2s(smem_u32(tiles + (size_t)slot * T * F), dx + (size_t)cp.v0 * F, bytes, bar(BAR_FULL + slot));
          c_next(cp);
          if (++slot == nstage) {
            slot = 0;
            if (!first_lap) use_par ^= 1u;
            first_lap = false;
          }
          progress = true;
        }
        if (!progress) __nanosleep(32);
      }
    }
  } else if (w < A_WARPS) {
    // =========================== phase A: warp w owns row w of every tile ===========================
    const int gl = lane % C::GROUP;
    const int k = gl / C::LPH;          // head owned by this lane
    const int l = gl % C::LPH;
    const bool lane_on = lane < C::GROUP;
    // index pipeline: row pointers two tiles ahead, first neighbours + softmax state one tile ahead.
    // quad (one value per lane): lanes 0/1 = indptr[v], indptr[v+1] of this warp's row; 2/3 = the tile's edge range
    auto load_quad = [&](const Cursor& c) {
      int val = 0;
      if (c.gi < n_mine && lane < 4) {
        const int rows = c_rows(c);
        if (lane >= 2)
          val = __ldg(indptr + c.v0 + (lane == 2 ? 0 : rows));
        else if (w < rows)
          val = __ldg(indptr + c.v0 + w + lane);
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
      if (c.gi < n_mine && w < c_rows(c)) {
        if (ip0 + lane < ip1) {
          p.u = __ldg(nbr + ip0 + lane);
          p.b = __ldg(bin + ip0 + lane);
        }
        if (lane_on) {
          const float* st = stat + (size_t)(c.v0 + w) * 3 * H;
          p.m = __ldg(st + k);
          p.den = __ldg(st + H + k);
        }
      }
      return p;
    };
    // lanes of this lane's head, for the head reductions (LPH need not be a power of two)
    int hsrc[C::LPH > 1 ? C::LPH - 1 : 1];
#pragma unroll
    for (int o = 1; o < C::LPH; ++o) hsrc[o - 1] = ((lane - l) + (l + o) % C::LPH) & 31;
    auto hsum = [&](float v) {
      float s = v;
#pragma unroll
      for (int o = 1; o < C::LPH; ++o) s += __shfl_sync(0xffffffffu, v, hsrc[o - 1]);
      return s;
    };

    Cursor cc{0, 0, 0, 0};
    c_seek(cc);
    Cursor c1 = cc;
    if (c1.gi < n_mine) c_next(c1);
    Cursor c2 = c1;
    if (c2.gi < n_mine) c_next(c2);
    int quad_c = load_quad(cc), quad_1 = load_quad(c1);
    Pre pre_c = load_edges(quad_c, cc);
    int slot = 0, rb = 0;
    uint32_t full_par = 0, zs_par = 0, cons_par = 0;
    bool rec_first_lap = true;
    const int lofs = k * D + VEC * l;                      // this lane's first column inside a raw row
    const int pofs = gl * VEC;                             // ... inside a lane-interleaved row
    while (cc.gi < n_mine) {
      const int quad_2 = load_quad(c2);
      const Pre pre_1 = load_edges(quad_1, c1);
      const int s0 = ext[4 * cc.gi + 2], ns = ext[4 * cc.gi + 3];
      const int rows = c_rows(cc);
      const bool last_tile = (cc.t + 1) * T >= cc.nwc;
      if (cc.t == 0) {                                     // this segment's source rows have landed
        mbar_wait(bar(BAR_ZFULL), zs_par);
        zs_par ^= 1u;
      }
      mbar_wait(bar(BAR_FULL + slot), full_par);
      float* trow = tiles + (size_t)slot * T * F + w * F;
      const int ip0 = __shfl_sync(0xffffffffu, quad_c, 0), ip1 = __shfl_sync(0xffffffffu, quad_c, 1);
      const int te0 = __shfl_sync(0xffffffffu, quad_c, 2), te1 = __shfl_sync(0xffffffffu, quad_c, 3);
      const int n_te = te1 - te0;
      const int rounds = n_te > CAP ? ceil_div(n_te, CAP) : 1;
      const bool row_on = w < rows;
      const float m_k = pre_c.m;
      const float rden_k = __fdividef(1.f, pre_c.den);
      float gv[NE], z0[NE];
      float s_head = 0.f, a0 = 0.f, pre0 = 0.f;
      int u0 = 0, b0 = 0;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        gv[i] = 0.f;
        z0[i] = 0.f;
      }
      // alpha of an edge from source row zrow with TF-IDF bin b (softmax state of this row: m_k, 1 / den_k)
      auto edge_alpha = [&](const float* zrow, int b, float& pre) {
        pre = zrow[FP + k] + q_s[b * H + k];
        return exp_fast(leaky(pre) - m_k) * rden_k;
      };
      // ---- phase A1: sh_v recomputed, g = dx * elu'(sh) in place, s = g . sh ----
      if (row_on) {
        float shv[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) shv[i] = 0.f;
        if (lane_on) {
#pragma unroll
          for (int i = 0; i < C::VPL; ++i)
            if (l + C::LPH * i < C::NV) lds_vec<VEC>(trow + lofs + VEC * C::LPH * i, gv + i * VEC);
        }
        if (ip1 > ip0) {                                   // first in-edge: its source row stays in registers
          u0 = min(max(__shfl_sync(0xffffffffu, pre_c.u, 0) - s0, 0), ns - 1);
          b0 = __shfl_sync(0xffffffffu, pre_c.b, 0);
          if (lane_on) {
            const float* zrow = zs + u0 * ldz;
            a0 = edge_alpha(zrow, b0, pre0);
#pragma unroll
            for (int i = 0; i < C::VPL; ++i)
              if (l + C::LPH * i < C::NV) lds_vec<VEC>(zrow + pofs + C::GROUP * VEC * i, z0 + i * VEC);
#pragma unroll
            for (int i = 0; i < NE; ++i) shv[i] = a0 * z0[i];
          }
        }
        if (ip1 > ip0 + 1) {
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
            for (int j = (c0 == ip0 ? 1 : 0); j < cnt; ++j) {
              const int u = min(max(__shfl_sync(0xffffffffu, my_u, j) - s0, 0), ns - 1);
              const int b = __shfl_sync(0xffffffffu, my_b, j);
              if (lane_on) {
                const float* zrow = zs + u * ldz;
                float pre;
                const float a = edge_alpha(zrow, b, pre);
#pragma unroll
                for (int i = 0; i < C::VPL; ++i) {
                  if (l + C::LPH * i < C::NV) {
                    float zt[VEC];
                    lds_vec<VEC>(zrow + pofs + C::GROUP * VEC * i, zt);
#pragma unroll
                    for (int t = 0; t < VEC; ++t) shv[i * VEC + t] = fmaf(a, zt[t], shv[i * VEC + t]);
                  }
                }
              }
            }
          }
        }
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          gv[i] *= (shv[i] > 0.f ? 1.f : exp_fast(shv[i]));
          part = fmaf(gv[i], shv[i], part);
        }
        s_head = hsum(part);
        if (lane_on) {
#pragma unroll
          for (int i = 0; i < C::VPL; ++i)
            if (l + C::LPH * i < C::NV) sts_vec<VEC>(trow + lofs + VEC * C::LPH * i, gv + i * VEC);
        }
      }

      for (int r = 0; r < rounds; ++r) {
        if (!rec_first_lap) mbar_wait(bar(BAR_CONSUMED + rb), cons_par);   // record buffer free again
        const int win0 = te0 + r * CAP;
        int* ru = rec_u + rb * CAP;
        int* rvb = rec_vb + rb * CAP;
        float* ra = rec_a + rb * CAP * H;
        float* rd = rec_d + rb * CAP * H;
        // ---- phase A2: per in-edge alpha, dpre = leaky'(pre) alpha (g_v . z_u - s) of this round's window ----
        if (row_on) {
          auto emit = [&](int e, int u, int b, float a, float pre, float part) {
            const float tdot = hsum(part);
            const int ei = e - win0;
            if (ei >= 0 && ei < CAP) {                     // warp-uniform
              if (lane_on && l == 0) {
                const float de = a * (tdot - s_head);
                ra[ei * H + k] = a;
                rd[ei * H + k] = pre > 0.f ? de : HSG_LEAKY_SLOPE * de;
              }
              if (lane == 0) {
                ru[ei] = u;
                rvb[ei] = w | (b << 8);
              }
            }
          };
          if (ip1 > ip0) {
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < NE; ++i) part = fmaf(gv[i], z0[i], part);
            emit(ip0, u0, b0, a0, pre0, part);
          }
          if (ip1 > ip0 + 1) {
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
              for (int j = (c0 == ip0 ? 1 : 0); j < cnt; ++j) {
                const int u = min(max(__shfl_sync(0xffffffffu, my_u, j) - s0, 0), ns - 1);
                const int b = __shfl_sync(0xffffffffu, my_b, j);
                const float* zrow = zs + u * ldz;
                float part = 0.f, pre = 0.f, a = 0.f;
                if (lane_on) {
                  a = edge_alpha(zrow, b, pre);
#pragma unroll
                  for (int i = 0; i < C::VPL; ++i) {
                    if (l + C::LPH * i < C::NV) {
                      float zt[VEC];
                      lds_vec<VEC>(zrow + pofs + C::GROUP * VEC * i, zt);
#pragma unroll
                      for (int t = 0; t < VEC; ++t) part = fmaf(gv[i * VEC + t], zt[t], part);
                    }
                  }
                }
                emit(c0 + j, u, b, a, pre, part);
              }
            }
          }
        }
        if (w == 0 && lane == 0) hdr[rb * 4] = n_te;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_READY + rb));
        if (++rb == NR) {
          rb = 0;
          if (!rec_first_lap) cons_par ^= 1u;
          rec_first_lap = false;
        }
      }
      if (last_tile) {                                     // this warp no longer reads the segment's source rows
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_ZFREE));
      }
      cc = c1;
      c1 = c2;
      if (c2.gi < n_mine) c_next(c2);
      quad_c = quad_1;
      quad_1 = quad_2;
      pre_c = pre_1;
      if (++slot == nstage) {
        slot = 0;
        full_par ^= 1u;
      }
    }
  } else {
    // ====== phase B: warp b accumulates the source rows u = b (mod B_WARPS), lanes as in phase A ======
    const int b_id = w - A_WARPS;
    const int gl = lane % C::GROUP;
    const int k = gl / C::LPH;
    const int l = gl % C::LPH;
    const bool lane_on = lane < C::GROUP;
    const int lofs = k * D + VEC * l;
    const int pofs = gl * VEC;
    float* dq_mine = dq_s + b_id * NQ;
    Cursor cb{0, 0, 0, 0};
    c_seek(cb);
    int slot = 0, rb = 0, cur = -1;
    uint32_t ready_par = 0;
    float acc[NE], accd = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = 0.f;
    auto flush = [&]() {                                   // pending sums -> the accumulator row of source `cur`
      if (cur >= 0 && lane_on) {
        float* drow = dzs + cur * ldz;
#pragma unroll
        for (int i = 0; i < C::VPL; ++i) {
          if (l + C::LPH * i < C::NV) {
            float t[VEC];
            lds_vec<VEC>(drow + pofs + C::GROUP * VEC * i, t);
#pragma unroll
            for (int jj = 0; jj < VEC; ++jj) t[jj] += acc[i * VEC + jj];
            sts_vec<VEC>(drow + pofs + C::GROUP * VEC * i, t);
          }
        }
        if (l == 0) drow[FP + k] += accd;
      }
#pragma unroll
      for (int i = 0; i < NE; ++i) acc[i] = 0.f;
      accd = 0.f;
    };
    while (cb.gi < n_mine) {
      const int s0 = ext[4 * cb.gi + 2], ns = ext[4 * cb.gi + 3];
      const bool last_tile = (cb.t + 1) * T >= cb.nwc;
      const float* tile = tiles + (size_t)slot * T * F;
      int r = 0, rounds = 1;
      do {
        mbar_wait(bar(BAR_READY + rb), ready_par);
        const int n_te = hdr[rb * 4];
        rounds = n_te > CAP ? ceil_div(n_te, CAP) : 1;
        const int n_rec = max(0, min(CAP, n_te - r * CAP));
        const int* ru = rec_u + rb * CAP;
        const int* rvb = rec_vb + rb * CAP;
        const float* ra = rec_a + rb * CAP * H;
        const float* rd = rec_d + rb * CAP * H;
        // operands of record e into registers (issued one record ahead of their use)
        float na = 0.f, nd = 0.f, ng[NE];
        int nvb = 0;
#pragma unroll
        for (int i = 0; i < NE; ++i) ng[i] = 0.f;
        auto fetch = [&](int e) {
          nvb = rvb[e];
          if (lane_on) {
            na = ra[e * H + k];
            nd = rd[e * H + k];
            const float* grow = tile + (nvb & 0xff) * F + lofs;
#pragma unroll
            for (int i = 0; i < C::VPL; ++i)
              if (l + C::LPH * i < C::NV) lds_vec<VEC>(grow + VEC * C::LPH * i, ng + i * VEC);
          }
        };
        for (int e0 = 0; e0 < n_rec; e0 += 32) {
          const int my_u = e0 + lane < n_rec ? ru[e0 + lane] : -1;
          unsigned mask = __ballot_sync(0xffffffffu, my_u >= 0 && (my_u % B_WARPS) == b_id);
          if (mask) fetch(e0 + __ffs(mask) - 1);
          while (mask) {                                   // this warp's records, in order: a fixed summation order
            const int jj = __ffs(mask) - 1;
            mask &= mask - 1;
            const int u = __shfl_sync(0xffffffffu, my_u, jj);
            const float a = na, dv = nd;
            const int vb = nvb;
            float gt[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) gt[i] = ng[i];
            if (mask) fetch(e0 + __ffs(mask) - 1);
            if (u != cur) {
              flush();
              cur = u;
            }
#pragma unroll
            for (int i = 0; i < NE; ++i) acc[i] = fmaf(a, gt[i], acc[i]);
            if (lane_on && l == 0) {
              accd += dv;
              dq_mine[(vb >> 8) * H + k] += dv;
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(BAR_CONSUMED + rb));
          if (r == rounds - 1) mbar_arrive(bar(BAR_EMPTY + slot));
        }
        if (++rb == NR) {
          rb = 0;
          ready_par ^= 1u;
        }
        ++r;
      } while (r < rounds);
      if (last_tile) {                                     // [dz | dp | 0] of this warp's source rows leave
        flush();
        cur = -1;
        __syncwarp();
        for (int u = b_id; u < ns; u += B_WARPS) {
          float4* acc4 = reinterpret_cast<float4*>(dzs + u * ldz);
          float4* out4 = reinterpret_cast<float4*>(dzp + (size_t)(s0 + u) * ldz);
          for (int jj = lane; jj < ldz / 4; jj += 32) {
            out4[jj] = acc4[jj];
            acc4[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        __syncwarp();
      }
      c_next(cb);
      if (++slot == nstage) slot = 0;
    }
  }
  __syncthreads();
  for (int i = tid; i < NQ; i += THREADS) {
    float s = 0.f;
#pragma unroll
    for (int b = 0; b < B_WARPS; ++b) s += dq_s[b * NQ + i];
    dq_part[(size_t)blockIdx.x * NQ + i] = s;
  }
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
      launch_k(edge_bwd_seg_kernel<H, D>, dim3(grid), dim3(SegCfg<H, D>::THREADS), (size_t)smem_bytes, s, c->n_seg, c->seg_dst_ptr,
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
