// Tensor-core (tcgen05 / TMEM) version of the tall-skinny products of the WSWGAT path.
//
//   D[Md, Nd] = sum_k A(md, k) * B(nd, k)        accumulators in TMEM (fp32), 128 x BN tile per CTA
//
// One templated kernel serves the three products of hsg_gemm.cu:
//   NT  (A_MN=0, B_MN=0)  C = A . B^T : A [M,K] and B [N,K] are K-major
//   NN  (A_MN=0, B_MN=1)  C = A . B   : B [K,N] is MN-major
//   TN  (A_MN=1, B_MN=1)  C = A^T . B : both operands MN-major, K = node dimension, split over gridDim.z
//
// Operands are fp32 in HBM.  A CTA stages 32-deep k-blocks through registers into shared memory in
// the canonical UMMA layouts (K-major: SWIZZLE_128B, 8-row x 128 B atoms; MN-major: SWIZZLE_128B_BASE32B,
// 4-k x 128 B atoms - the only MN-major layout the hardware accepts for 32-bit operands),
// splitting every value on the way into hi = x rounded to the nearest TF32 value and lo = x - hi.  One elected thread issues
//      D += A_hi B_hi ;  D += A_hi B_lo ;  D += A_lo B_hi          (precise mode, "3xTF32")
// with tcgen05.mma.kind::tf32 - error ~2^-21 relative, inside BASELINE.json's fp32 bound of 1e-5 -
// or only the first product in fast mode (TF32, bound 2e-2).  Two smem stages; tcgen05.commit
// -> mbarrier releases a stage while the next k-block is being staged.  The epilogue reads the
// accumulators with tcgen05.ld (each warp its own 32 TMEM lanes) and applies bias / ReLU /
// residual / ReLU-mask before the global store.  Weight-gradient column sums come for free from
// an extra all-ones B column.
#include "hsg_common.cuh"

namespace hsg {
namespace tc {

constexpr int TM = 128;          // MMA M (TMEM lanes)
constexpr int BK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int BN_MAX = 256;      // MMA N per CTA (TMEM columns)
constexpr int THREADS = 256;
constexpr int A_BYTES = TM * 128;          // one hi or lo A tile
constexpr int B_BYTES = BN_MAX * 128;      // one hi or lo B tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 1024 + 64;  // + alignment slack + barriers

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
    if (!done && ++spins > (1ull << 26)) __trap();   // never hang the GPU: fail loudly instead
  }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (sm_100 version bit 46).  layout: 2 = SWIZZLE_128B (K-major operands),
// 1 = SWIZZLE_128B_BASE32B (the only layout the hardware takes for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// hi = x rounded to nearest TF32 (10 explicit mantissa bits), lo = x - hi (exact in fp32, <= 12 significant
// bits, so the tensor core's own TF32 truncation of lo costs at most one bit: ~2^-23 relative overall).
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void split_store(float4 v, char* hi, char* lo, uint32_t off, bool want_lo) {
  float4 h;
  if (want_lo) {
    h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    *reinterpret_cast<float4*>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  } else {
    h = v;   // single-pass mode: the tensor core truncates to TF32 itself
  }
  *reinterpret_cast<float4*>(hi + off) = h;
}

struct Operand {
  const float* p;
  int ld;
  int ext_mn;   // valid extent along the MMA M / N dimension
  int ext_k;    // valid extent along K
};

// ---- staging: global fp32 -> (hi, lo) SWIZZLE_128B tiles -------------------------------------
// K-major tile: `rows` rows (M or N index) x 32 k.  smem: row r -> (r/8)*1024 + (r%8)*128, 16B chunk c at c ^ (r%8)
constexpr int A_CHUNKS = TM * 8 / THREADS;        // float4 per thread for a 128-row tile  (4)
constexpr int B_CHUNKS = BN_MAX * 8 / THREADS;    // float4 per thread for a 256-row tile  (8)

template <int NCH>
__device__ __forceinline__ void load_kmajor(const Operand& op, int mn0, int k0, int rows, float4* regs) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int id = threadIdx.x + i * THREADS;
    const int r = id >> 3, c = id & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int gm = mn0 + r, gk = k0 + 4 * c;
    if (r < rows && gm < op.ext_mn && gk < op.ext_k)
      v = __ldg(reinterpret_cast<const float4*>(op.p + (size_t)gm * op.ld + gk));
    regs[i] = v;
  }
}

template <int NCH>
__device__ __forceinline__ void store_kmajor(char* hi, char* lo, int rows, const float4* regs, bool want_lo) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int id = threadIdx.x + i * THREADS;
    const int r = id >> 3, c = id & 7;
    if (r < rows) {
      const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
      split_store(regs[i], hi, lo, off, want_lo);
    }
  }
}

// MN-major tile: 32 k x `width` (M or N index, multiple of 32), SWIZZLE_128B_BASE32B:
// atoms of 4 k-rows x 128 B (32 fp32 along MN); k -> (k/4)*sbo + (k%4)*128, mn -> (mn/32)*512,
// inside a row the 32-byte unit j = (mn%32)/8 sits at j ^ (k%4); sbo = (width/32)*512
template <int NCH>
__device__ __forceinline__ void load_mnmajor(const Operand& op, int mn0, int k0, int width, int ones_col,
                                             float4* regs) {
  const int w4 = width >> 2;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int id = threadIdx.x + i * THREADS;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id < 32 * w4) {
      const int k = id / w4, n4 = id - k * w4;
      const int gk = k0 + k, gn = mn0 + 4 * n4;
      if (gk < op.ext_k) {
        if (gn < op.ext_mn) v = __ldg(reinterpret_cast<const float4*>(op.p + (size_t)gk * op.ld + gn));
        if (ones_col >= gn && ones_col < gn + 4) {       // all-ones column -> column sums of the other operand
          if (ones_col == gn) v.x = 1.f;
          else if (ones_col == gn + 1) v.y = 1.f;
          else if (ones_col == gn + 2) v.z = 1.f;
          else v.w = 1.f;
        }
      }
    }
    regs[i] = v;
  }
}

template <int NCH>
__device__ __forceinline__ void store_mnmajor(char* hi, char* lo, int width, const float4* regs, bool want_lo) {
  const int w4 = width >> 2;
  const uint32_t sbo = (uint32_t)(width >> 5) * 512u;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int id = threadIdx.x + i * THREADS;
    if (id < 32 * w4) {
      const int k = id / w4, n4 = id - k * w4;
      const uint32_t c16 = (uint32_t)(n4 & 7), kr = (uint32_t)(k & 3);
      const uint32_t off = (uint32_t)(k >> 2) * sbo + (uint32_t)(n4 >> 3) * 512u + kr * 128u +
                           (((c16 >> 1) ^ kr) << 5) + ((c16 & 1) << 4);
      split_store(regs[i], hi, lo, off, want_lo);
    }
  }
}

struct Epilogue {
  float* D;            // output (or split-K partial base)
  int ldd;
  size_t split_stride; // elements between split-K partials (0 when not split)
  const float* bias;
  const float* R;
  int ldr;
  int epi;
  float* colsum_part;  // [splits][ext_m] column sums (TN) or nullptr
  int ones_col;        // D column that holds the column sums, -1 if none
};

// grid: (ceil(Md/128), ceil(Nd_total/bn), splits).  k range of split z: [z*k_per_split, min(K, (z+1)*k_per_split))
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(Operand A, Operand B, int Md, int Nd, int K, int bn, int k_per_split, int precise, Epilogue ep) {
  extern __shared__ char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES);   // [0,1]: stage free, [2]: unused
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * bn;
  const int n_valid = min(bn, Nd - n0);                 // valid D columns of this tile (incl. a ones column)
  const int n_mma = (n_valid + 15) & ~15;               // MMA N (multiple of 16, <= 256)
  const int n_stage = B_MN ? ((n_mma + 31) & ~31) : n_mma;   // staged B extent
  const int k_beg = blockIdx.z * k_per_split;
  const int k_end = min(K, k_beg + k_per_split);
  const int nkb = (k_end - k_beg + BK - 1) / BK;

  // precise mode keeps the small correction products in a second accumulator (columns 256..511): the
  // tensor core's fp32 accumulation truncates, so the error grows with the number of updates of one
  // accumulator; splitting leaves K/8 updates on the main one and adds the two in the epilogue.
  const uint32_t tmem_cols = precise ? 2u * BN_MAX : (uint32_t)BN_MAX;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc(A_MN, B_MN, n_mma);
  const bool want_lo = precise != 0;

  Operand Ak = A, Bk = B;
  Ak.ext_k = min(A.ext_k, k_end);
  Bk.ext_k = min(B.ext_k, k_end);

  float4 ra[A_CHUNKS], rb[B_CHUNKS];
  auto load_block = [&](int kb) {
    const int k0 = k_beg + kb * BK;
    if (A_MN) load_mnmajor<A_CHUNKS>(Ak, m0, k0, TM, -1, ra);
    else load_kmajor<A_CHUNKS>(Ak, m0, k0, TM, ra);
    if (B_MN) load_mnmajor<B_CHUNKS>(Bk, n0, k0, n_stage, ep.ones_col, rb);
    else load_kmajor<B_CHUNKS>(Bk, n0, k0, n_stage, rb);
  };
  auto store_block = [&](int s) {
    char* st = smem + s * STAGE_BYTES;
    if (A_MN) store_mnmajor<A_CHUNKS>(st, st + A_BYTES, TM, ra, want_lo);
    else store_kmajor<A_CHUNKS>(st, st + A_BYTES, TM, ra, want_lo);
    char* sb = st + 2 * A_BYTES;
    if (B_MN) store_mnmajor<B_CHUNKS>(sb, sb + B_BYTES, n_stage, rb, want_lo);
    else store_kmajor<B_CHUNKS>(sb, sb + B_BYTES, n_stage, rb, want_lo);
  };

  if (nkb > 0) load_block(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= 2) mbar_wait(smem_u32(&bars[s]), ((kb >> 1) - 1) & 1);   // MMAs of k-block kb-2 have drained stage s
    store_block(s);
    if (kb + 1 < nkb) load_block(kb + 1);                                // prefetch next k-block into registers
    fence_async_smem();                                                  // generic-proxy writes -> async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
      const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
      // K-major: 8-row groups 1024 B apart, a k-step of 8 fp32 = 32 B inside the swizzled row.
      // MN-major: MN atoms 512 B apart (LBO), 4-k groups sbo apart, a k-step of 8 = two 4-k groups.
      const uint32_t a_sbo = A_MN ? (TM / 32) * 512u : 1024u, a_lbo = A_MN ? 512u : 16u, a_lay = A_MN ? 1u : 2u;
      const uint32_t b_sbo = B_MN ? (uint32_t)(n_stage >> 5) * 512u : 1024u, b_lbo = B_MN ? 512u : 16u,
                     b_lay = B_MN ? 1u : 2u;
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint32_t a_off = A_MN ? ks * 2u * a_sbo : ks * 32u;
        const uint32_t b_off = B_MN ? ks * 2u * b_sbo : ks * 32u;
        const uint64_t dah = make_desc(a_hi + a_off, a_lbo, a_sbo, a_lay),
                       dbh = make_desc(b_hi + b_off, b_lbo, b_sbo, b_lay);
        tc_mma_tf32(tmem_d, dah, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        if (want_lo) {
          const uint64_t dal = make_desc(a_lo + a_off, a_lbo, a_sbo, a_lay),
                         dbl = make_desc(b_lo + b_off, b_lbo, b_sbo, b_lay);
          tc_mma_tf32(tmem_d + BN_MAX, dah, dbl, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          tc_mma_tf32(tmem_d + BN_MAX, dal, dbh, idesc, 1u);
        }
      }
      tc_commit(smem_u32(&bars[s]));
    }
  }
  // wait for the last commit (it tracks all earlier MMAs of this thread)
  if (nkb > 0) {
    const int last = nkb - 1;
    mbar_wait(smem_u32(&bars[last & 1]), (last >> 1) & 1);
  }
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global ----
  const int lg = warp & 3;                     // TMEM lane group of this warp
  const int row = m0 + lg * 32 + lane;
  float* Dz = ep.D + (size_t)blockIdx.z * ep.split_stride;
  for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += 64) {
    float v[32];
    if (nkb > 0) {
      tc_ld32(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
      if (want_lo) {
        float w[32];
        tc_ld32(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(BN_MAX + c0), w);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += w[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (row < Md) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = n0 + c0 + j;
        float o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float c = v[j + t];
          const int cc = col + t;
          if (cc < n0 + n_valid && cc != ep.ones_col) {
            if (ep.epi & HSG_EPI_BIAS) c += __ldg(ep.bias + cc);
            if (ep.epi & HSG_EPI_RELU) c = fmaxf(c, 0.f);
            if (ep.epi & HSG_EPI_ADD) c += __ldg(ep.R + (size_t)row * ep.ldr + cc);
            if (ep.epi & HSG_EPI_RELU_MASK) c = __ldg(ep.R + (size_t)row * ep.ldr + cc) > 0.f ? c : 0.f;
          }
          o[t] = c;
        }
        // real output columns of THIS tile (never touch the neighbouring tile's columns)
        const int n_out = min(ep.ones_col >= 0 ? ep.ones_col : Nd, n0 + n_valid);
        if (col + 3 < n_out && (ep.ldd & 3) == 0) {
          *reinterpret_cast<float4*>(Dz + (size_t)row * ep.ldd + col) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (col + t < n_out) Dz[(size_t)row * ep.ldd + col + t] = o[t];
        }
        if (ep.colsum_part != nullptr && ep.ones_col >= col && ep.ones_col < col + 4)
          ep.colsum_part[(size_t)blockIdx.z * Md + row] = o[ep.ones_col - col];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
}

static bool g_attr_done[3] = {false, false, false};

template <bool A_MN, bool B_MN>
static int launch(int which, dim3 grid, Operand A, Operand B, int Md, int Nd, int K, int bn, int k_per_split,
                  int precise, Epilogue ep, cudaStream_t s) {
  if (!g_attr_done[which]) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess)
      return HSG_ERR_CUDA;
    g_attr_done[which] = true;
  }
  gemm_tc_kernel<A_MN, B_MN><<<grid, THREADS, SMEM_BYTES, s>>>(A, B, Md, Nd, K, bn, k_per_split, precise, ep);
  return check_launch();
}

static int pick_bn(int n_total) {
  // D column tile: a multiple of 16 up to 256, tiles as even as possible
  const int tiles = ceil_div(n_total, BN_MAX);
  int bn = ceil_div(ceil_div(n_total, tiles), 16) * 16;
  return bn > BN_MAX ? BN_MAX : bn;
}

int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, const float* R, int ldr, int epi, int precise, cudaStream_t s) {
  Operand a{A, lda, M, K}, b{B, ldb, N, K};
  const int bn = pick_bn(N);
  Epilogue ep{C, ldc, 0, bias, R, ldr, epi, nullptr, -1};
  dim3 grid(ceil_div(M, TM), ceil_div(N, bn), 1);
  return launch<false, false>(0, grid, a, b, M, N, K, bn, ceil_div(K, BK) * BK, precise, ep, s);
}

int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* R,
            int ldr, int epi, int precise, cudaStream_t s) {
  Operand a{A, lda, M, K}, b{B, ldb, N, K};
  const int bn = pick_bn(N);
  Epilogue ep{C, ldc, 0, nullptr, R, ldr, epi, nullptr, -1};
  dim3 grid(ceil_div(M, TM), ceil_div(N, bn), 1);
  return launch<false, true>(1, grid, a, b, M, N, K, bn, ceil_div(K, BK) * BK, precise, ep, s);
}

// part[z][N1][N2] (+ part_col[z][N1]) for z < splits; rows_per_split multiple of 32
int gemm_tn(int M, int N1, int N2, const float* A, int lda, const float* B, int ldb, float* part, float* part_col,
            int splits, int rows_per_split, int precise, cudaStream_t s) {
  Operand a{A, lda, N1, M}, b{B, ldb, N2, M};
  const int n_total = N2 + (part_col ? 1 : 0);
  const int bn = pick_bn(n_total);
  Epilogue ep{part, N2, (size_t)N1 * N2, nullptr, nullptr, 0, 0, part_col, part_col ? N2 : -1};
  dim3 grid(ceil_div(N1, TM), ceil_div(n_total, bn), splits);
  return launch<true, true>(2, grid, a, b, N1, n_total, M, bn, rows_per_split, precise, ep, s);
}

}  // namespace tc
}  // namespace hsg
