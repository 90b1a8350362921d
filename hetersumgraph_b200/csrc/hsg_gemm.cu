// Tall-skinny dense products of the WSWGAT path (projections K1/K6 and the FFN
// products of K4) in exact fp32 (FFMA) - the parity mode that meets the 1e-5
// bound of BASELINE.json.  M (nodes) is tall, N and K are 64..512.
//
//   gemm_nt : C = A . B^T   nn.Linear / Conv1d(k=1) forward   (GATLayer.py:38,110,146)
//   gemm_nn : C = A . B     input gradients
//   gemm_tn : C = A^T . B   weight gradients, split over the node dimension with a
//                           fixed-order second stage (bitwise reproducible)
#include <atomic>
#include <cstdlib>

#include "hsg_common.cuh"
#include "hsg_internal.cuh"

namespace hsg {

// tcgen05 path (hsg_gemm_tc.cu)
namespace tc {
int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, const float* R, int ldr, int epi, int precise, cudaStream_t s);
int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* R,
            int ldr, int epi, int precise, cudaStream_t s);
int gemm_tn(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* part, float* part_col,
            int splits, int rows_per_split, int precise, cudaStream_t s, bool cta_per_item = false);
int trace_ctl(int on, unsigned long long* host_out, int max_events);
}  // namespace tc

// tcgen05 cta_group::2 path (hsg_gemm_tc2.cu): 256-row tiles over a CTA pair, B tile shared between the two SMs
namespace tc2 {
int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, const float* R, int ldr, int epi, int precise, cudaStream_t s);
int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* R,
            int ldr, int epi, int precise, cudaStream_t s);
int trace_ctl(int on, unsigned long long* host_out, int max_events);
}  // namespace tc2

// 1: NT / NN products of at least g_pair_min_rows rows run on CTA pairs (default; HSG_GEMM_PAIR=0 turns it off)
static std::atomic<int> g_pair{-1};
static std::atomic<int> g_pair_min_rows{512};

static bool pair_enabled(int M) {
  int v = g_pair.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HSG_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
    g_pair.store(v);
  }
  return v != 0 && M >= g_pair_min_rows.load(std::memory_order_relaxed);
}

// 0: FFMA exact fp32, 1: tcgen05 3xTF32 (fp32-parity, default), 2: tcgen05 single-pass TF32
static std::atomic<int> g_gemm_mode{1};
// arithmetic selector of the tensor-core kernels: 1 = 3xTF32 (fp32 parity), 0 = one TF32 pass, 2 = bf16 operands
// (kind::f16, fp32 accumulation)
static inline int tc_arith(int mode) { return mode == 1 ? 1 : (mode == 3 ? 2 : 0); }


constexpr int BN = 64, BK = 16;
constexpr int TN = 4;
constexpr int GEMM_THREADS = 256;
constexpr int APAD = 4;
// Products below this many flops run on the FFMA tiles even in the tensor-core modes: a 128x128 tcgen05 tile
// pipeline has ~15-20 us of fill/drain latency and leaves most SMs idle on the sentence-side shapes (M ~ 1 k),
// where 64x64 FFMA tiles finish in a few microseconds (and are exact fp32).
static std::atomic<double> g_small_flops{3.0e8};

// ---------------------------------------------------------------------------
// C[M,N] = A[M,K] * op(B);  B_NT: B is [N,K] (K contiguous), else B is [K,N].
// VEC: all of K%4, lda%4, ldb%4 (and N%4 for !B_NT) hold and pointers are 16B aligned.
// ---------------------------------------------------------------------------
// BM = 128 (8x4 micro-tile) for tall products, BM = 64 (4x4) for small ones (more CTAs).
template <bool B_NT, bool VEC, int BM>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
            float* __restrict__ C, int ldc, const float* __restrict__ bias, const float* __restrict__ R, int ldr,
            int epi) {
  pdl_prologue();
  constexpr int TM = BM / 16, AH = BM / 64;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + APAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // staging registers
  float4 ra[AH], rb;
  const int a_row = tid / 4, a_kq = (tid % 4) * 4;  // A tile: rows a_row and a_row+64, k = a_kq..a_kq+3
  const int bnt_row = tid / 4, bnt_kq = (tid % 4) * 4;  // B_NT tile: BN rows x BK
  const int bnn_k = tid / 16, bnn_n = (tid % 16) * 4;   // B_NN tile: BK rows x BN

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < AH; ++h) {
      int m = m0 + a_row + 64 * h;
      int k = k0 + a_kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const float* p = A + (size_t)m * lda + k;
        if (VEC) {
          if (k < K) v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k + 0 < K) v.x = p[0];
          if (k + 1 < K) v.y = p[1];
          if (k + 2 < K) v.z = p[2];
          if (k + 3 < K) v.w = p[3];
        }
      }
      ra[h] = v;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (B_NT) {
      int n = n0 + bnt_row, k = k0 + bnt_kq;
      if (n < N) {
        const float* p = B + (size_t)n * ldb + k;
        if (VEC) {
          if (k < K) v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k + 0 < K) v.x = p[0];
          if (k + 1 < K) v.y = p[1];
          if (k + 2 < K) v.z = p[2];
          if (k + 3 < K) v.w = p[3];
        }
      }
    } else {
      int k = k0 + bnn_k, n = n0 + bnn_n;
      if (k < K) {
        const float* p = B + (size_t)k * ldb + n;
        if (VEC) {
          if (n < N) v = *reinterpret_cast<const float4*>(p);
        } else {
          if (n + 0 < N) v.x = p[0];
          if (n + 1 < N) v.y = p[1];
          if (n + 2 < N) v.z = p[2];
          if (n + 3 < N) v.w = p[3];
        }
      }
    }
    rb = v;
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < AH; ++h) {
      int r = a_row + 64 * h;
      As[buf][a_kq + 0][r] = ra[h].x;
      As[buf][a_kq + 1][r] = ra[h].y;
      As[buf][a_kq + 2][r] = ra[h].z;
      As[buf][a_kq + 3][r] = ra[h].w;
    }
    if (B_NT) {
      Bs[buf][bnt_kq + 0][bnt_row] = rb.x;
      Bs[buf][bnt_kq + 1][bnt_row] = rb.y;
      Bs[buf][bnt_kq + 2][bnt_row] = rb.z;
      Bs[buf][bnt_kq + 3][bnt_row] = rb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][bnn_k][bnn_n]) = rb;
    }
  };

  const int nk = ceil_div(K, BK);
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM];
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4 * h]);
        a[4 * h] = av.x; a[4 * h + 1] = av.y; a[4 * h + 2] = av.z; a[4 * h + 3] = av.w;
      }
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
      float b[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const int n = n0 + tx * TN;
  const bool vec_out = VEC && (ldc % 4 == 0) && (n + 3 < N) && (R == nullptr || ldr % 4 == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float c = acc[i][j];
      int nn = n + j;
      if (nn < N) {
        if (epi & HSG_EPI_BIAS) c += bias[nn];
        if (epi & HSG_EPI_RELU) c = fmaxf(c, 0.f);
        if (epi & HSG_EPI_ADD) c += R[(size_t)m * ldr + nn];
        if (epi & HSG_EPI_RELU_MASK) c = R[(size_t)m * ldr + nn] > 0.f ? c : 0.f;
      }
      v[j] = c;
    }
    if (vec_out) {
      *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n + j < N) C[(size_t)m * ldc + n + j] = v[j];
    }
  }
}

// ---------------------------------------------------------------------------
// Weight gradient: part[z][n1][n2] = sum_{m in chunk z} A[m,n1] B[m,n2]
// ---------------------------------------------------------------------------
constexpr int TB = 64;      // output tile TB x TB
constexpr int TBK = 16;     // rows of the reduction dimension per stage

template <bool VEC>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(int M, int N1, int N2, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
               float* __restrict__ part, float* __restrict__ part_col, int rows_per_split) {
  pdl_prologue();
  __shared__ __align__(16) float As[2][TBK][TB + 4];
  __shared__ __align__(16) float Bs[2][TBK][TB + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int n1_0 = blockIdx.x * TB, n2_0 = blockIdx.y * TB;
  const int z = blockIdx.z;
  const int m_beg = z * rows_per_split;
  const int m_end = min(M, m_beg + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;  // column sum of A for thread tid < TB (only blockIdx.y == 0)
  const bool do_col = (part_col != nullptr) && (blockIdx.y == 0);

  const int lr = tid / 16, lc = (tid % 16) * 4;
  float4 ra, rb;
  auto load = [&](int mrow) {
    int m = mrow + lr;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (m < m_end) {
      const float* pa = A + (size_t)m * lda + n1_0 + lc;
      const float* pb = B + (size_t)m * ldb + n2_0 + lc;
      if (VEC) {
        if (n1_0 + lc < N1) ra = *reinterpret_cast<const float4*>(pa);
        if (n2_0 + lc < N2) rb = *reinterpret_cast<const float4*>(pb);
      } else {
        if (n1_0 + lc + 0 < N1) ra.x = pa[0];
        if (n1_0 + lc + 1 < N1) ra.y = pa[1];
        if (n1_0 + lc + 2 < N1) ra.z = pa[2];
        if (n1_0 + lc + 3 < N1) ra.w = pa[3];
        if (n2_0 + lc + 0 < N2) rb.x = pb[0];
        if (n2_0 + lc + 1 < N2) rb.y = pb[1];
        if (n2_0 + lc + 2 < N2) rb.z = pb[2];
        if (n2_0 + lc + 3 < N2) rb.w = pb[3];
      }
    }
  };
  auto store = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = rb;
  };
  const int nst = ceil_div(max(m_end - m_beg, 0), TBK);
  if (nst > 0) {
    load(m_beg);
    store(0);
  }
  __syncthreads();
  for (int st = 0; st < nst; ++st) {
    const int buf = st & 1;
    if (st + 1 < nst) load(m_beg + (st + 1) * TBK);
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_col && tid < TB) {
#pragma unroll
      for (int k = 0; k < TBK; ++k) csum += As[buf][k][tid];
    }
    if (st + 1 < nst) {
      store(buf ^ 1);
      __syncthreads();
    }
  }
  float* out = part + (size_t)z * N1 * N2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n1 = n1_0 + ty * 4 + i;
    if (n1 >= N1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n2 = n2_0 + tx * 4 + j;
      if (n2 < N2) out[(size_t)n1 * N2 + n2] = acc[i][j];
    }
  }
  if (do_col && tid < TB && n1_0 + tid < N1) part_col[(size_t)z * N1 + n1_0 + tid] = csum;
}

// Column sums of A [M, N1] over row chunks of `rows` rows: part_col[chunk][c] = sum of A[row][c] in row order per row
// lane (8 lanes: rows l, l + 8, ...), the 8 lane sums added in order.  Used instead of the ones column of the
// tensor-core product when that column would open a new column of output tiles (N2 a multiple of 128: the FFN's
// dW2 = dr^T hdn with db2, 300 x 512 + 1 -> 15 tiles instead of 12, 57 us instead of 41 us per launch at 11 817 rows).
constexpr int CS_LANES = 8;
__global__ void __launch_bounds__(128 * CS_LANES)
colsum_part_kernel(int M, int N1, const float* __restrict__ A, int lda, int rows, float* __restrict__ part_col) {
  pdl_prologue();
  __shared__ float red[CS_LANES][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int c = blockIdx.y * 128 + tx;
  const int r0 = blockIdx.x * rows, r1 = min(M, r0 + rows);
  float sm = 0.f;
  if (c < N1) {
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += CS_LANES) sm += __ldg(A + (size_t)r * lda + c);
  }
  red[ty][tx] = sm;
  __syncthreads();
  if (ty == 0 && c < N1) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < CS_LANES; ++l) t += red[l][tx];
    part_col[(size_t)blockIdx.x * N1 + c] = t;
  }
}
constexpr int CS_MAX_PARTS = 592;
static void colsum_plan(int M, int* parts, int* rows) {
  int r = 256;
  if (ceil_div(M, r) > CS_MAX_PARTS) r = ceil_div(ceil_div(M, CS_MAX_PARTS), 8) * 8;
  *rows = r;
  *parts = ceil_div(M, r);
}

// part_col: [ncol][N1] column-sum partials (ncol == nsplit when the product kernel wrote them, its own count when
// colsum_part_kernel did)
// one thread per output, the splits summed in order: few splits (the FFN weight gradients: 9-12)
__global__ void gemm_tn_reduce_kernel(int nsplit, int ncol, int N1, int N2, const float* __restrict__ part,
                                      const float* __restrict__ part_col, float* __restrict__ C, int ldc,
                                      float* __restrict__ colsum, int accumulate) {
  pdl_prologue();
  const int total = N1 * N2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + N1; i += gridDim.x * blockDim.x) {
    if (i < total) {
      float s = 0.f;
      for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * total + i];
      float* o = C + (size_t)(i / N2) * ldc + (i % N2);
      *o = accumulate ? *o + s : s;
    } else if (colsum != nullptr) {
      int c = i - total;
      float s = 0.f;
      for (int z = 0; z < ncol; ++z) s += part_col[(size_t)z * N1 + c];
      colsum[c] = accumulate ? colsum[c] + s : s;
    }
  }
}

// MANY splits (the projection weight gradients: 47 splits of a 72 x 300 output, whose reduce sits on the critical tail of
// the step).  Block = 32 consecutive outputs x 8 split lanes: warp w sums the partials z = w, w + 8, ... of its 32 outputs in that
// order (coalesced 128-byte rows, up to ceil(nsplit / 8) loads in flight per thread instead of a chain of nsplit), the
// eight lane sums are added in order.  Fixed order: bitwise reproducible.  (One thread per output summing all splits
// took 7 us for 47 splits of the 72 x 300 projection gradient - on the critical tail of the step.)
__global__ void __launch_bounds__(256)
gemm_tn_reduce_split_kernel(int nsplit, int ncol, int N1, int N2, const float* __restrict__ part,
                      const float* __restrict__ part_col, float* __restrict__ C, int ldc, float* __restrict__ colsum,
                      int accumulate) {
  pdl_prologue();
  __shared__ float red[8][32];
  const int total = N1 * N2, all = total + (colsum != nullptr ? N1 : 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = blockIdx.x * 32; base < all; base += gridDim.x * 32) {
    const int i = base + lane;
    float s = 0.f;
    if (i < total) {
#pragma unroll 4
      for (int z = w; z < nsplit; z += 8) s += part[(size_t)z * total + i];
    } else if (i < all) {
      const int c = i - total;
#pragma unroll 4
      for (int z = w; z < ncol; z += 8) s += part_col[(size_t)z * N1 + c];
    }
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < all) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red[r][lane];
      if (i < total) {
        float* o = C + (size_t)(i / N2) * ldc + (i % N2);
        *o = accumulate ? *o + t : t;
      } else {
        const int c = i - total;
        colsum[c] = accumulate ? colsum[c] + t : t;
      }
    }
    __syncthreads();
  }
}

// split of the node dimension for the tcgen05 weight-gradient kernel: 128 x 128 output tiles
// cta_budget: how many SMs the product may occupy (0 = all); g_tn_min_rows: shortest reduction range of one split.
// Both are tuning knobs (hsg_set_side_ctas / hsg_set_tn_min_rows).  Measured on the 32-graph step (profiles/
// r02c_sweep.jsonl): the step gets monotonically SLOWER when the side-stream weight-gradient products are confined to
// fewer SMs (0.674 ms with all SMs, 0.746 ms with 72) or cut into fewer, longer splits - the step is bound by the sum of
// the kernels' work, not by the small-kernel chain those products delay - so the defaults are "all SMs" and 256 rows.
static std::atomic<int> g_tn_min_rows{256};

// Short work items (hsg_set_tn_item_rows / HSG_TN_ITEM_ROWS, 0 = off): when the whole reduction fits in <= 48 splits of
// `item_rows` rows, the product is cut into tiles x splits items of that length and launched with ONE CTA PER ITEM instead
// of one persistent CTA per SM.  The weight-gradient products run on the low-priority side stream next to the serial
// dx -> edge-backward chain; a persistent CTA keeps its SM (and all of its shared memory) for the whole product, so the
// short kernels of that chain waited 40-50 us for an SM (CUPTI timeline r02t); with short items an SM is handed back
// every few microseconds and the block scheduler gives it to the higher-priority stream first.
static std::atomic<int> g_tn_item_rows{-1};
static int tn_item_rows() {
  int v = g_tn_item_rows.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HSG_TN_ITEM_ROWS");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
    v = (v + 31) & ~31;
    g_tn_item_rows.store(v);
  }
  return v;
}
static bool tn_short_items(int M) {
  const int r = tn_item_rows();
  return r > 0 && ceil_div(M, r) <= 48;
}

// HSG_TN_COLSUM_APART=0: keep the ones column inside the product even when it opens a new column of tiles
static bool colsum_apart() {
  static const int v = [] {
    const char* e = getenv("HSG_TN_COLSUM_APART");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v != 0;
}

static void tn_plan_tc(int M, int N1, int N2, bool colsum, int* splits, int* rows, int cta_budget = 0) {
  if (tn_short_items(M)) {
    *rows = tn_item_rows();
    *splits = ceil_div(M, *rows);
    return;
  }
  const int tiles = ceil_div(N1, 128) * ceil_div(N2 + (colsum ? 1 : 0), 128);
  const int sms = (cta_budget > 0 && cta_budget < 148) ? cta_budget : 148;
  int want = sms / tiles;   // floor: tiles * splits <= budget, ONE wave of the persistent kernel (ceil made 150 > 148)
  if (want < 1) want = 1;
  int r = ceil_div(ceil_div(M, want), 32) * 32;
  const int min_rows = g_tn_min_rows.load(std::memory_order_relaxed);
  if (r < min_rows) r = min_rows;
  *rows = r;
  *splits = ceil_div(M, r);
}

static int tn_splits(int M, int N1, int N2) {
  int tiles = ceil_div(N1, TB) * ceil_div(N2, TB);
  int want = ceil_div(2 * 148, tiles);
  int maxs = ceil_div(M, 256);
  int s = want < maxs ? want : maxs;
  if (s < 1) s = 1;
  if (s > 128) s = 128;
  return s;
}

static bool is_small(int M, int N, int K) {
  return 2.0 * (double)M * (double)N * (double)K < g_small_flops.load(std::memory_order_relaxed);
}

bool gemm_is_small(int M, int N, int K) { return is_small(M, N, K); }


template <bool B_NT>
static int launch_ffma(bool small, bool vec, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* bias, const float* R, int ldr, int epi, cudaStream_t s) {
  if (small)
    return B_NT ? gemm_small_nt(M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi, s)
                : gemm_small_nn(M, N, K, A, lda, B, ldb, C, ldc, R, ldr, epi, s);
  dim3 grid(ceil_div(M, 128), ceil_div(N, BN));
  if (vec)
    launch_k(gemm_kernel<B_NT, true, 128>, dim3(grid), dim3(GEMM_THREADS), 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi);
  else
    launch_k(gemm_kernel<B_NT, false, 128>, dim3(grid), dim3(GEMM_THREADS), 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi);
  return check_launch();
}

int gemm_tn_ex(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
               float* colsum, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s, int cta_budget) {
  if (M < 0 || N1 <= 0 || N2 <= 0 || !C || !ws || (M > 0 && (!A || !B))) return HSG_ERR_ARG;
  if (ws_bytes < hsg_gemm_tn_workspace_bytes(M, N1, N2)) return HSG_ERR_WORKSPACE;
  if (M > 0 && is_small(M, N1, N2)) {   // cluster split over the node rows: no partials, no reduce launch
    LaunchScope ls(SLOT_GEMM_TN, s);
    return gemm_small_tn(M, N1, N2, A, lda, B, ldb, C, ldc, colsum, accumulate, s);
  }
  int nsplit = M > 0 ? tn_splits(M, N1, N2) : 0;
  float* part = reinterpret_cast<float*>(ws);
  float* part_col = part + (size_t)nsplit * N1 * N2;
  const int mode = g_gemm_mode.load(std::memory_order_relaxed);
  const bool vec_tc = (N1 % 4 == 0) && (N2 % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B);
  int ncol = -1;                                      // column-sum partials: -1 = one per split, written by the product
  if (M > 0 && mode != 0 && vec_tc) {
    int rows = 0;
    // the ones column would be the only column of a fifth (ninth, ...) column of tiles: sum the columns of A apart
    const bool col_apart = colsum != nullptr && (N2 % 128 == 0) && colsum_apart();
    tn_plan_tc(M, N1, N2, colsum != nullptr && !col_apart, &nsplit, &rows, cta_budget);
    int crows = 0, cparts = 0;
    if (col_apart) colsum_plan(M, &cparts, &crows);
    if (((size_t)nsplit * (size_t)N1 * N2 + (size_t)(col_apart ? cparts : nsplit) * N1) * sizeof(float) > ws_bytes)
      return HSG_ERR_WORKSPACE;
    part_col = part + (size_t)nsplit * N1 * N2;
    LaunchScope ls(SLOT_GEMM_TN, s);
    if (col_apart) {
      launch_k(colsum_part_kernel, dim3(cparts, ceil_div(N1, 128)), dim3(128 * CS_LANES), 0, s, M, N1, A, lda, crows, part_col);
      ncol = cparts;
    }
    int rc = tc::gemm_tn(M, N1, N2, A, lda, B, ldb, part, (colsum && !col_apart) ? part_col : nullptr, nsplit, rows,
                         tc_arith(mode), s, tn_short_items(M));
    if (rc) return rc;
  } else if (M > 0) {
    int rows = ceil_div(M, nsplit);
    rows = ceil_div(rows, TBK) * TBK;
    nsplit = ceil_div(M, rows);
    part_col = part + (size_t)nsplit * N1 * N2;
    dim3 grid(ceil_div(N1, TB), ceil_div(N2, TB), nsplit);
    bool vec = (N1 % 4 == 0) && (N2 % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B);
    LaunchScope ls(SLOT_GEMM_TN, s);
    if (vec)
      launch_k(gemm_tn_kernel<true>, dim3(grid), dim3(256), 0, s, M, N1, N2, A, lda, B, ldb, part, colsum ? part_col : nullptr, rows);
    else
      launch_k(gemm_tn_kernel<false>, dim3(grid), dim3(256), 0, s, M, N1, N2, A, lda, B, ldb, part, colsum ? part_col : nullptr, rows);
    int rc = check_launch();
    if (rc) return rc;
  }
  {
    LaunchScope ls(SLOT_GEMM_TN_REDUCE, s);
    int total = N1 * N2 + N1;
    if (nsplit >= 24) {                                // measured: slower than the plain kernel at 12 splits (0.605 / 0.5925 ms)
      int blocks = ceil_div(total, 32);
      if (blocks > 2368) blocks = 2368;
      launch_k(gemm_tn_reduce_split_kernel, dim3(blocks), dim3(256), 0, s, nsplit, ncol < 0 ? nsplit : ncol, N1, N2, part,
               part_col, C, ldc, colsum, accumulate);
    } else {
      int blocks = ceil_div(total, 256);
      if (blocks > 1184) blocks = 1184;
      launch_k(gemm_tn_reduce_kernel, dim3(blocks), dim3(256), 0, s, nsplit, ncol < 0 ? nsplit : ncol, N1, N2, part, part_col,
               C, ldc, colsum, accumulate);
    }
  }
  return check_launch();
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const float* bias, const float* R, int ldr, int epi, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return HSG_ERR_ARG;
  if (M == 0) return HSG_OK;                          // empty node set: nothing to do (pointers may be NULL)
  if (!A || !B || !C) return HSG_ERR_ARG;
  if ((epi & HSG_EPI_BIAS) && !bias) return HSG_ERR_ARG;
  if ((epi & (HSG_EPI_ADD | HSG_EPI_RELU_MASK)) && !R) return HSG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (K % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B) && aligned16(C) &&
             (R == nullptr || aligned16(R));
  LaunchScope ls(SLOT_GEMM_NT, s);
  const int mode = g_gemm_mode.load(std::memory_order_relaxed);
  const bool small = is_small(M, N, K);
  if (mode != 0 && !small && vec && (ldc % 4 == 0)) {
    if (pair_enabled(M)) return tc2::gemm_nt(M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi, tc_arith(mode), s);
    return tc::gemm_nt(M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi, tc_arith(mode), s);
  }
  return launch_ffma<true>(small, vec, M, N, K, A, lda, B, ldb, C, ldc, bias, R, ldr, epi, s);
}

int hsg_gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const float* R, int ldr, int epi, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return HSG_ERR_ARG;
  if (epi & (HSG_EPI_BIAS | HSG_EPI_RELU)) return HSG_ERR_ARG;
  if (M == 0) return HSG_OK;                          // empty node set: nothing to do (pointers may be NULL)
  if (!A || !B || !C) return HSG_ERR_ARG;
  if ((epi & (HSG_EPI_ADD | HSG_EPI_RELU_MASK)) && !R) return HSG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (K % 4 == 0) && (N % 4 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && aligned16(A) && aligned16(B) &&
             aligned16(C) && (R == nullptr || aligned16(R));
  LaunchScope ls(SLOT_GEMM_NN, s);
  const int mode = g_gemm_mode.load(std::memory_order_relaxed);
  const bool small = is_small(M, N, K);
  if (mode != 0 && !small && vec && (ldc % 4 == 0)) {
    if (pair_enabled(M)) return tc2::gemm_nn(M, N, K, A, lda, B, ldb, C, ldc, R, ldr, epi, tc_arith(mode), s);
    return tc::gemm_nn(M, N, K, A, lda, B, ldb, C, ldc, R, ldr, epi, tc_arith(mode), s);
  }
  return launch_ffma<false>(small, vec, M, N, K, A, lda, B, ldb, C, ldc, nullptr, R, ldr, epi, s);
}

int hsg_set_gemm_pair(int on) {
  g_pair.store(on ? 1 : 0);
  return HSG_OK;
}

int hsg_set_tn_item_rows(int rows) {
  if (rows < 0) return HSG_ERR_ARG;
  g_tn_item_rows.store((rows + 31) & ~31);
  return HSG_OK;
}

int hsg_set_tn_min_rows(int rows) {
  if (rows < 32) return HSG_ERR_ARG;
  g_tn_min_rows.store((rows + 31) & ~31);
  return HSG_OK;
}

int hsg_set_gemm_mode(int mode) {
  if (mode < 0 || mode > 3) return HSG_ERR_ARG;
  g_gemm_mode.store(mode);
  return HSG_OK;
}

int hsg_get_gemm_mode(void) { return g_gemm_mode.load(); }

int hsg_set_gemm_small_flops(double flops) {
  if (!(flops >= 0.0)) return HSG_ERR_ARG;
  g_small_flops.store(flops);
  return HSG_OK;
}

int hsg_gemm_trace(int on, unsigned long long* host_out, int max_events) { return tc::trace_ctl(on, host_out, max_events); }
int hsg_gemm_pair_trace(int on, unsigned long long* host_out, int max_events) { return tc2::trace_ctl(on, host_out, max_events); }

size_t hsg_gemm_tn_workspace_bytes(int M, int N1, int N2) {
  if (M <= 0 || N1 <= 0 || N2 <= 0) return 16;
  size_t s = (size_t)tn_splits(M, N1, N2);
  // both plans of the tensor-core path: WITHOUT the column-sum column the product has fewer column tiles and therefore
  // MORE splits (e.g. M = 4 099, 300 x 512: 12 splits against 9) - sizing for the column-sum plan alone let the
  // partials of the other one run past the workspace (round-1 bug, found by the round-2 test order)
  for (int cs = 0; cs < 2; ++cs) {
    int s_tc = 0, rows_tc = 0;
    tn_plan_tc(M, N1, N2, cs != 0, &s_tc, &rows_tc);
    if ((size_t)s_tc > s) s = (size_t)s_tc;
  }
  int cparts = 0, crows = 0;
  colsum_plan(M, &cparts, &crows);                   // column sums taken apart (colsum_part_kernel)
  if ((size_t)cparts < s) cparts = (int)s;
  return (s * (size_t)N1 * N2 + (size_t)cparts * N1) * sizeof(float) + 16;
}

int hsg_gemm_tn_acc(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                    float* colsum, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  return gemm_tn_ex(M, N1, N2, A, lda, B, ldb, C, ldc, colsum, ws, ws_bytes, accumulate ? 1 : 0, (cudaStream_t)stream, 0);
}

int hsg_gemm_tn(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                float* colsum, void* ws, size_t ws_bytes, void* stream) {
  return gemm_tn_ex(M, N1, N2, A, lda, B, ldb, C, ldc, colsum, ws, ws_bytes, 0, (cudaStream_t)stream, 0);
}

}  // extern "C"
