// Lane-interleaved ("permuted") feature layout shared by the edge kernels, the attention-prep
// kernels and the host side.
//
// A feature row of F = H*D floats is spread over GROUP = H*LPH lanes so that every lane owns
// elements of exactly one head.  Tensors that are GATHERED per edge (zp = [z | p], g, dzp) are
// stored in the order the lanes read them:
//     original column c = k*D + VEC*(l + LPH*i) + t      (head k, lane-in-head l, vector i, element t)
//     permuted position  = (i*GROUP + k*LPH + l)*VEC + t
// so each of the VPL gather instructions of a warp reads one contiguous GROUP*VEC*4-byte slab.
// The permutation costs nothing: it is a row permutation of the augmented projection weight.
// Holes (when NV is not a multiple of LPH) are zero columns.  FP = VPL*GROUP*VEC >= F.
#pragma once
#include <cuda_runtime.h>

namespace hsg {

struct EdgeLayout {
  int H, D, vec, nv, lph, vpl, group, eps, fp, ldz;
};

// lanes per head.  Short heads (at most two vectors, e.g. the W2S default D = 8) are owned by ONE lane: the logit,
// softmax weight and the backward dot product of a head then need no cross-lane traffic and no redundant math, and
// 32/H edge rows are in flight per warp instruction.  Longer heads are spread over up to 32/H lanes.
__host__ __device__ constexpr int edge_lph(int H, int nv) {
  return nv <= 2 ? 1 : (nv < 32 / H ? nv : 32 / H);
}

__host__ __device__ inline EdgeLayout make_edge_layout(int H, int D) {
  EdgeLayout L;
  L.H = H;
  L.D = D;
  L.vec = (D % 4 == 0) ? 4 : ((D % 2 == 0) ? 2 : 1);
  L.nv = D / L.vec;
  L.lph = edge_lph(H, L.nv);
  L.vpl = (L.nv + L.lph - 1) / L.lph;
  L.group = H * L.lph;
  L.eps = 32 / L.group;
  L.fp = L.vpl * L.group * L.vec;
  L.ldz = (L.fp + H + 7) / 8 * 8;
  return L;
}

// original column -> permuted position
__host__ __device__ inline int edge_perm(const EdgeLayout& L, int c) {
  const int k = c / L.D, r = c % L.D;
  const int nv = r / L.vec, t = r % L.vec;
  const int l = nv % L.lph, i = nv / L.lph;
  return (i * L.group + k * L.lph + l) * L.vec + t;
}

// permuted position -> original column, or -1 for a hole / out of range
__host__ __device__ inline int edge_unperm(const EdgeLayout& L, int pos) {
  if (pos < 0 || pos >= L.fp) return -1;
  const int slab = L.group * L.vec;
  const int i = pos / slab, rem = pos % slab;
  const int gl = rem / L.vec, t = rem % L.vec;
  const int k = gl / L.lph, l = gl % L.lph;
  const int nv = l + L.lph * i;
  if (nv >= L.nv) return -1;
  return k * L.D + nv * L.vec + t;
}

}  // namespace hsg
