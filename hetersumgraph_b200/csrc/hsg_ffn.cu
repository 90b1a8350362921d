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

// dgamma / dbeta = column sums of the per-block partials; 32 columns x 32 row groups per CTA, fixed order
// (row group r sums blocks r, r+32, ... ; the 32 group sums are then added in order)
__global__ void __launch_bounds__(1024) layernorm_bwd_reduce_kernel(int nblocks, int D, const float* __restrict__ part,
                                                                    float* __restrict__ dgamma,
                                                                    float* __restrict__ dbeta, int accumulate) {
  pdl_prologue();
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < 2 * D)
    for (int b = ty; b < nblocks; b += 32) s += part[(size_t)b * 2 * D + i];
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

int layernorm_bwd_ex(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma, float* dr,
                     float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s) {
  if (N < 0 || D <= 0 || !dy || !r || !stats || !gamma || !dr || !dgamma || !dbeta || !ws) return HSG_ERR_ARG;
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
  LaunchScope ls(SLOT_LN_BWD_REDUCE, s);
  launch_k(layernorm_bwd_reduce_kernel, dim3(ceil_div(2 * D, 32)), dim3(1024), 0, s, grid, D, part, dgamma, dbeta, accumulate);
  return check_launch();
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

size_t hsg_layernorm_bwd_workspace_bytes(int N, int D) {
  (void)N;
  return (size_t)LN_MAX_BLOCKS * 2 * (D > 0 ? D : 1) * sizeof(float) + 16;
}

int hsg_layernorm_bwd(int N, int D, const float* dy, const float* r, const float* stats, const float* gamma,
                      float* dr, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  return layernorm_bwd_ex(N, D, dy, r, stats, gamma, dr, dgamma, dbeta, ws, ws_bytes, 0, (cudaStream_t)stream);
}

}  // extern "C"
