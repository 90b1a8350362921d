// S2S layer type: SGATLayer / MultiHeadSGATLayer (module/GATLayer.py:49-78, module/GATStackLayer.py:27-44) and the
// "S2S" branch of WSWGAT (module/GAT.py:38-39,50-52).  The reference never instantiates it (HiGraph.py:57-76); it is
// built for completeness of the module surface.
//
// Exact semantics on the reference's graphs (DGL-0.4 zero fills, see oracle/closed_form.py:s2s_multi_head_cf, pinned
// to the reference's own classes in tests/golden/s2s_*.npz):
//     z = fc(h) on the supernodes;   e_v = leaky_relu(a[d:2d] . z_v)   (logit of every word in-edge of v: z_src = 0)
//     pull over ALL in-edges of v:   deg_v word edges (message 0) + x_v extra edges (logit 0, message z_src)
//     sh_v = sum_{extra j->v} z_j / (deg_v exp(e_v) + x_v)
// The extra in-edges are implicit: supernode j contributes its z to the group xmember[j], supernode v reads
// mult * (sum of the group xgrp[v]) - HSG: every sentence belongs to and reads its graph's group, each ordered pair
// twice (dataloader.py:262-263, mult = 2); HDSG: sentences belong to their document's group, documents read it
// (dataloader.py:385, mult = 1).  One CTA per graph (a graph's supernodes are contiguous rows), one thread per
// feature column, sequential walks over the graph's <= ~100 supernodes: deterministic, no atomics.
#include "hsg_common.cuh"

namespace hsg {

constexpr int S2S_MAX_F = 512;

__device__ __forceinline__ float s2s_head_sum(float v, float* tmp, int c, int d) {
  // sum over the d columns of this thread's head; all threads call it (block-wide barriers inside)
  tmp[c] = v;
  __syncthreads();
  const int k0 = (c / d) * d;
  float s = 0.f;
  for (int j = 0; j < d; ++j) s += tmp[k0 + j];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(S2S_MAX_F)
s2s_fwd_kernel(hsg_s2s_graph gph, const float* __restrict__ z, const float* __restrict__ a,
               const float* __restrict__ origin, float* __restrict__ S, float* __restrict__ sh,
               float* __restrict__ x) {
  pdl_prologue();
  __shared__ float tmp[S2S_MAX_F];
  const int F = gph.H * gph.d, d = gph.d;
  const int c = threadIdx.x;                       // blockDim.x == F
  const int g = blockIdx.x;
  const int r0 = gph.super_ptr[g], r1 = gph.super_ptr[g + 1];
  const float ad = a[(c / d) * 2 * d + d + (c % d)];
  // group sums (S is indexed by supernode row; a group's row and all its members lie in this graph)
  for (int r = r0; r < r1; ++r) S[(size_t)r * F + c] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int m = gph.xmember[r];
    if (m >= 0) S[(size_t)m * F + c] += z[(size_t)r * F + c];
  }
  for (int r = r0; r < r1; ++r) {
    const float t = s2s_head_sum(ad * z[(size_t)r * F + c], tmp, c, d);
    const float e = t > 0.f ? t : HSG_LEAKY_SLOPE * t;
    const int grp = gph.xgrp[r];
    const float xv = (float)gph.extra[r];
    const float deg = (float)(gph.deg_indptr[r + 1] - gph.deg_indptr[r]);
    float o = 0.f;
    if (grp >= 0 && xv > 0.f) o = (float)gph.mult * S[(size_t)grp * F + c] / (deg * __expf(e) + xv);
    sh[(size_t)r * F + c] = o;
    if (x != nullptr) x[(size_t)r * F + c] = (o > 0.f ? o : __expf(o) - 1.f) + origin[(size_t)r * F + c];
  }
}

__global__ void __launch_bounds__(S2S_MAX_F)
s2s_bwd_kernel(hsg_s2s_graph gph, const float* __restrict__ z, const float* __restrict__ a,
               const float* __restrict__ S, const float* __restrict__ dx, const float* __restrict__ dsh,
               float* __restrict__ dS, float* __restrict__ dz, float* __restrict__ da_part) {
  pdl_prologue();
  __shared__ float tmp[S2S_MAX_F];
  const int F = gph.H * gph.d, d = gph.d;
  const int c = threadIdx.x;
  const int g = blockIdx.x;
  const int r0 = gph.super_ptr[g], r1 = gph.super_ptr[g + 1];
  const float ad = a[(c / d) * 2 * d + d + (c % d)];
  float da_c = 0.f;
  for (int r = r0; r < r1; ++r) dS[(size_t)r * F + c] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float zc = z[(size_t)r * F + c];
    const float t = s2s_head_sum(ad * zc, tmp, c, d);
    const float e = t > 0.f ? t : HSG_LEAKY_SLOPE * t;
    const int grp = gph.xgrp[r];
    const float xv = (float)gph.extra[r];
    const float deg = (float)(gph.deg_indptr[r + 1] - gph.deg_indptr[r]);
    float dzs = 0.f;
    if (grp >= 0 && xv > 0.f) {
      const float ex = __expf(e);
      const float den = deg * ex + xv;
      const float cv = (float)gph.mult / den;
      const float Sg = S[(size_t)grp * F + c];
      const float o = cv * Sg;
      float gs;
      if (dx != nullptr) gs = dx[(size_t)r * F + c] * (o > 0.f ? 1.f : __expf(o));
      else gs = dsh[(size_t)r * F + c];
      const float dc = s2s_head_sum(gs * Sg, tmp, c, d);          // d sh / d c_v summed over the head
      const float dden = -cv * dc / den;
      const float de = dden * deg * ex;
      const float dt = t > 0.f ? de : HSG_LEAKY_SLOPE * de;
      dzs = dt * ad;
      da_c = fmaf(dt, zc, da_c);
      dS[(size_t)grp * F + c] += cv * gs;                          // sequential over r in this thread
    } else {
      (void)s2s_head_sum(0.f, tmp, c, d);                          // keep the barriers uniform
    }
    dz[(size_t)r * F + c] = dzs;
  }
  for (int r = r0; r < r1; ++r) {
    const int m = gph.xmember[r];
    if (m >= 0) dz[(size_t)r * F + c] += dS[(size_t)m * F + c];
  }
  da_part[(size_t)g * F + c] = da_c;
}

// da[k, j] = 0 (a_src multiplies DGL's zero-filled word z), da[k, d + j] = sum over graphs (fixed order)
__global__ void __launch_bounds__(256)
s2s_da_reduce_kernel(int n_graphs, int H, int d, const float* __restrict__ da_part, float* __restrict__ da,
                     int accumulate) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H * d) return;
  float s = 0.f;
  for (int g = 0; g < n_graphs; ++g) s += da_part[(size_t)g * H * d + c];
  const int k = c / d, j = c % d;
  float* o0 = da + k * 2 * d + j;
  float* o1 = da + k * 2 * d + d + j;
  *o0 = accumulate ? *o0 : 0.f;
  *o1 = accumulate ? *o1 + s : s;
}

static bool s2s_ok(const hsg_s2s_graph* g) {
  return g && g->n_graphs >= 0 && g->n_super >= 0 && g->H > 0 && g->d > 0 && g->H * g->d <= S2S_MAX_F && g->mult > 0 &&
         g->super_ptr && g->deg_indptr && g->extra && g->xgrp && g->xmember;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_s2s_fwd(const hsg_s2s_graph* g, const float* z, const float* a, const float* origin, float* S, float* sh,
                float* x, void* stream) {
  if (!s2s_ok(g) || !z || !a || !S || !sh || (x && !origin)) return HSG_ERR_ARG;
  if (g->n_graphs == 0 || g->n_super == 0) return HSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_S2S, s);
  launch_k(s2s_fwd_kernel, dim3(g->n_graphs), dim3(g->H * g->d), 0, s, *g, z, a, origin, S, sh, x);
  return check_launch();
}

size_t hsg_s2s_bwd_workspace_bytes(int n_graphs, int H, int d) {
  return (size_t)(n_graphs > 0 ? n_graphs : 1) * H * d * sizeof(float) + 16;
}

int hsg_s2s_bwd(const hsg_s2s_graph* g, const float* z, const float* a, const float* S, const float* dx,
                const float* dsh, float* dS, float* dz, float* da, int accumulate, void* ws, size_t ws_bytes,
                void* stream) {
  if (!s2s_ok(g) || !z || !a || !S || (!dx && !dsh) || !dS || !dz || !da || !ws) return HSG_ERR_ARG;
  if (ws_bytes < hsg_s2s_bwd_workspace_bytes(g->n_graphs, g->H, g->d)) return HSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(SLOT_S2S, s);
  float* part = reinterpret_cast<float*>(ws);
  if (g->n_graphs > 0 && g->n_super > 0)
    launch_k(s2s_bwd_kernel, dim3(g->n_graphs), dim3(g->H * g->d), 0, s, *g, z, a, S, dx, dsh, dS, dz, part);
  launch_k(s2s_da_reduce_kernel, dim3(ceil_div(g->H * g->d, 256)), dim3(256), 0, s,
           (g->n_super > 0 ? g->n_graphs : 0), g->H, g->d, part, da, accumulate);
  return check_launch();
}

}  // extern "C"
