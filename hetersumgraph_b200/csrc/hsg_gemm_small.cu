// Small dense products of the WSWGAT path: the sentence-side shapes (M ~ 1 k rows: FFN on supernodes, the S2W
// projection and their gradients) where a 128x128 tensor-core tile pipeline would occupy 8-32 SMs with a long
// serial k-loop.  Exact fp32 FFMA, 64x64 output tiles, and the reduction dimension split over a THREAD-BLOCK
// CLUSTER: the 1-8 CTAs of a cluster each accumulate a K-slice of the same output tile, park the partial tile in
// their shared memory and reduce it through distributed shared memory (cluster.map_shared_rank) in a fixed rank
// order - deterministic, no workspace, no second launch, and the epilogue (bias / ReLU / residual / ReLU-mask /
// gradient accumulation) runs once on the reduced tile.
//
//   MODE 0  NT : C[M,N]   = A[M,K] . B[N,K]^T      A, B k-contiguous
//   MODE 1  NN : C[M,N]   = A[M,K] . B[K,N]        A k-contiguous, B n-contiguous
//   MODE 2  TN : C[N1,N2] = A[R,N1]^T . B[R,N2]    both mn-contiguous, reduction over the R node rows;
//                colsum[N1] = column sums of A (bias gradient) in the same pass
#include <cooperative_groups.h>

#include "hsg_common.cuh"
#include "hsg_internal.cuh"

namespace cg = cooperative_groups;

namespace hsg {

namespace small {

constexpr int T = 64;        // output tile T x T
constexpr int KB = 16;       // reduction elements per stage
constexpr int PAD = 4;
constexpr int THREADS = 256;

struct Params {
  int M, N, K;               // output M x N, reduction length K
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  const float* bias;
  const float* R; int ldr;
  int epi;
  int accumulate;            // C += (TN weight gradients summed over applications)
  float* colsum;             // TN only
  int k_per_rank;            // multiple of KB
};

// stage loader: `rows` indexes the output dimension (m or n), k the reduction dimension.
//   K_CONTIG: element (row, k) at P[row*ld + k]   - one float4 along k per thread, transposed into smem
//   else    : element (row, k) at P[k*ld + row]   - one float4 along row per thread
template <bool K_CONTIG, bool VEC>
__device__ __forceinline__ float4 load_stage(const float* __restrict__ P, int ld, int row0, int n_rows, int k0,
                                             int k_end, int tid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (K_CONTIG) {
    const int r = row0 + (tid >> 2), k = k0 + (tid & 3) * 4;
    if (r < n_rows) {
      const float* p = P + (size_t)r * ld + k;
      if (VEC) {
        if (k < k_end) v = __ldg(reinterpret_cast<const float4*>(p));   // k_end, k multiples of 4
      } else {
        if (k + 0 < k_end) v.x = __ldg(p + 0);
        if (k + 1 < k_end) v.y = __ldg(p + 1);
        if (k + 2 < k_end) v.z = __ldg(p + 2);
        if (k + 3 < k_end) v.w = __ldg(p + 3);
      }
    }
  } else {
    const int k = k0 + (tid >> 4), r = row0 + (tid & 15) * 4;
    if (k < k_end) {
      const float* p = P + (size_t)k * ld + r;
      if (VEC) {
        if (r < n_rows) v = __ldg(reinterpret_cast<const float4*>(p));   // n_rows multiple of 4
      } else {
        if (r + 0 < n_rows) v.x = __ldg(p + 0);
        if (r + 1 < n_rows) v.y = __ldg(p + 1);
        if (r + 2 < n_rows) v.z = __ldg(p + 2);
        if (r + 3 < n_rows) v.w = __ldg(p + 3);
      }
    }
  }
  return v;
}

template <bool K_CONTIG>
__device__ __forceinline__ void store_stage(float (*S)[T + PAD], float4 v, int tid) {
  if (K_CONTIG) {
    const int r = tid >> 2, kq = (tid & 3) * 4;
    S[kq + 0][r] = v.x;
    S[kq + 1][r] = v.y;
    S[kq + 2][r] = v.z;
    S[kq + 3][r] = v.w;
  } else {
    *reinterpret_cast<float4*>(&S[tid >> 4][(tid & 15) * 4]) = v;
  }
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(THREADS) gemm_small_kernel(Params p) {
  pdl_prologue();
  constexpr bool A_KC = MODE != 2, B_KC = MODE == 0;
  __shared__ __align__(16) float As[2][KB][T + PAD];
  __shared__ __align__(16) float Bs[2][KB][T + PAD];
  __shared__ __align__(16) float Ps[T][T + PAD];    // partial tile for the cluster reduction
  __shared__ float Pc[T];                           // partial column sums (TN)
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();       // == blockIdx.z (cluster spans z only)
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * T, n0 = blockIdx.y * T;
  const int k_beg = rank * p.k_per_rank;
  const int k_end = min(p.K, k_beg + p.k_per_rank);
  const bool do_col = MODE == 2 && p.colsum != nullptr && blockIdx.y == 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;

  const int nst = k_end > k_beg ? (k_end - k_beg + KB - 1) / KB : 0;
  if (nst > 0) {
    const float4 ra = load_stage<A_KC, VEC>(p.A, p.lda, m0, p.M, k_beg, k_end, tid);
    const float4 rb = load_stage<B_KC, VEC>(p.B, p.ldb, n0, p.N, k_beg, k_end, tid);
    store_stage<A_KC>(As[0], ra, tid);
    store_stage<B_KC>(Bs[0], rb, tid);
  }
  __syncthreads();
  for (int st = 0; st < nst; ++st) {
    const int buf = st & 1;
    float4 ra, rb;
    if (st + 1 < nst) {
      ra = load_stage<A_KC, VEC>(p.A, p.lda, m0, p.M, k_beg + (st + 1) * KB, k_end, tid);
      rb = load_stage<B_KC, VEC>(p.B, p.ldb, n0, p.N, k_beg + (st + 1) * KB, k_end, tid);
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_col && tid < T) {
#pragma unroll
      for (int k = 0; k < KB; ++k) csum += As[buf][k][tid];
    }
    if (st + 1 < nst) {
      store_stage<A_KC>(As[buf ^ 1], ra, tid);
      store_stage<B_KC>(Bs[buf ^ 1], rb, tid);
      __syncthreads();
    }
  }

  // ---- reduction over the cluster + epilogue ------------------------------------------------------
  auto epilogue_store = [&](int m, int n, float4 v) {       // 4 consecutive columns n..n+3 of row m
    if (m >= p.M) return;
    float o[4] = {v.x, v.y, v.z, v.w};
    const bool vec = VEC && (n + 3 < p.N) && ((p.ldc & 3) == 0) && (p.R == nullptr || (p.ldr & 3) == 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n + j;
      if (nn < p.N) {
        float c = o[j];
        if (p.epi & HSG_EPI_BIAS) c += __ldg(p.bias + nn);
        if (p.epi & HSG_EPI_RELU) c = fmaxf(c, 0.f);
        if (p.epi & HSG_EPI_ADD) c += p.R[(size_t)m * p.ldr + nn];
        if (p.epi & HSG_EPI_RELU_MASK) c = p.R[(size_t)m * p.ldr + nn] > 0.f ? c : 0.f;
        if (p.accumulate) c += p.C[(size_t)m * p.ldc + nn];
        o[j] = c;
      }
    }
    if (vec) {
      *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.N) p.C[(size_t)m * p.ldc + n + j] = o[j];
    }
  };

  if (cs == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      epilogue_store(m0 + ty * 4 + i, n0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    if (do_col && tid < T && m0 + tid < p.M) p.colsum[m0 + tid] = p.accumulate ? p.colsum[m0 + tid] + csum : csum;
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&Ps[ty * 4 + i][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  if (tid < T) Pc[tid] = csum;
  cluster.sync();
  // rank r owns rows [r*T/cs, (r+1)*T/cs) of the tile: T/cs rows x 16 float4 columns
  const int rows_per = T / cs;
  for (int e = tid; e < rows_per * (T / 4); e += THREADS) {
    const int rl = rank * rows_per + e / (T / 4), c4 = (e % (T / 4)) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < cs; ++r) {                                  // fixed rank order: deterministic
      const float* remote = cluster.map_shared_rank(&Ps[0][0], r);
      const float4 v = *reinterpret_cast<const float4*>(remote + rl * (T + PAD) + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    epilogue_store(m0 + rl, n0 + c4, s);
  }
  if (do_col && rank == 0 && tid < T && m0 + tid < p.M) {
    float s = 0.f;
    for (int r = 0; r < cs; ++r) s += cluster.map_shared_rank(&Pc[0], r)[tid];
    p.colsum[m0 + tid] = p.accumulate ? p.colsum[m0 + tid] + s : s;
  }
  cluster.sync();   // nobody leaves while a peer may still read its shared memory
}

static int pick_cluster(int tiles, int K) {
  int cs = 1;
  while (cs < 8 && tiles * cs < 148 && K / (cs * 2) >= 32) cs *= 2;
  return cs;
}

template <int MODE>
static int launch(Params p, bool vec, cudaStream_t s) {
  const int mt = ceil_div(p.M, T), nt = ceil_div(p.N, T);
  const int cs = pick_cluster(mt * nt, p.K);
  p.k_per_rank = ceil_div(ceil_div(p.K, cs), KB) * KB;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(mt, nt, cs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = cs;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = vec ? cudaLaunchKernelEx(&cfg, gemm_small_kernel<MODE, true>, p)
                      : cudaLaunchKernelEx(&cfg, gemm_small_kernel<MODE, false>, p);
  if (e != cudaSuccess) return HSG_ERR_CUDA;
  return check_launch();
}

}  // namespace small

int gemm_small_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* bias, const float* R, int ldr, int epi, cudaStream_t s) {
  small::Params p{M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi, 0, nullptr, 0};
  const bool vec = (K % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B) && aligned16(C) &&
                   (R == nullptr || aligned16(R));
  return small::launch<0>(p, vec, s);
}

int gemm_small_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* R, int ldr, int epi, cudaStream_t s) {
  small::Params p{M, N, K, A, lda, B, ldb, C, ldc, nullptr, R, ldr, epi, 0, nullptr, 0};
  const bool vec = (K % 4 == 0) && (N % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B) &&
                   aligned16(C) && (R == nullptr || aligned16(R));
  return small::launch<1>(p, vec, s);
}

int gemm_small_tn(int Rows, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  float* colsum, int accumulate, cudaStream_t s) {
  small::Params p{N1, N2, Rows, A, lda, B, ldb, C, ldc, nullptr, nullptr, 0, 0, accumulate, colsum, 0};
  const bool vec = (N1 % 4 == 0) && (N2 % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) &&
                   aligned16(B) && aligned16(C);
  return small::launch<2>(p, vec, s);
}

}  // namespace hsg
