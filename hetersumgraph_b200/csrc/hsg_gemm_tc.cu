// Tensor-core (tcgen05 / TMEM / TMA) version of the tall-skinny products of the WSWGAT path.
//
//   D[Md, Nd] = sum_k A(md, k) * B(nd, k)        accumulators in TMEM (fp32), 128 x BN tile per CTA
//
// One templated kernel serves the three products of hsg_gemm.cu:
//   NT  (A_MN=0, B_MN=0)  C = A . B^T : A [M,K] and B [N,K] are K-major
//   NN  (A_MN=0, B_MN=1)  C = A . B   : B [K,N] is MN-major
//   TN  (A_MN=1, B_MN=1)  C = A^T . B : both operands MN-major, K = node dimension, split over gridDim.z
//
// Operands are fp32 in HBM.  Warp-specialised pipeline, STAGES shared-memory stages:
//   * one thread issues TMA (cp.async.bulk.tensor) loads of raw 32-deep k-blocks straight into the canonical
//     UMMA layouts - K-major operands with SWIZZLE_128B (8-row x 128 B atoms), MN-major operands with
//     SWIZZLE_128B_ATOM_32B (4-k x 128 B atoms: the only MN-major layout the tensor core takes for 32-bit
//     data); tile edges are zero-filled by the TMA unit;
//   * six converter warps sweep the landed tile linearly and split every value into hi = x truncated to the
//     TF32 grid (what the tensor core reads from the raw tile: nothing to store) and lo = x - hi (second tile);
//   * one thread issues   D += A_hi B_hi ;  D' += A_hi B_lo ;  D' += A_lo B_hi     (precise mode, "3xTF32")
//     with tcgen05.mma.kind::tf32 - error ~2^-22 relative, inside BASELINE.json's fp32 bound of 1e-5 - or
//     only the first product on the raw tile in fast mode (TF32, the tensor core truncates itself);
//   * tcgen05.commit -> mbarrier hands the stage back to the TMA thread.
// The epilogue reads the accumulators with tcgen05.ld (each warp its own 32 TMEM lanes) and applies bias /
// ReLU / residual / ReLU-mask before the global store.  Weight-gradient column sums come for free from an
// extra all-ones B column.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "hsg_common.cuh"

namespace hsg {
namespace tc {

constexpr int TM = 128;          // MMA M (TMEM lanes)
constexpr int BK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int BN_MAX = 128;      // MMA N per CTA (TMEM columns per accumulator)
constexpr int THREADS = 384;     // warp 0: TMA producer, warp 1: MMA issuer, warps 2-7: converters, warps 8-11: epilogue
constexpr int NCONV = 192;
constexpr int NEPI_WARPS = 4;
constexpr int TILE_BYTES = TM * 128;             // one hi or lo tile (A: 128 rows, B: up to 128 rows/cols) = 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi | A_lo | B_hi | B_lo = 64 KB
constexpr int STAGES = 3;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + NEPI_WARPS * 32 * 33 * 4;  // + alignment slack + barriers

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

// bf16 x bf16 -> fp32 (kind::f16), K = 16 per instruction
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// bf16 mode: fp32 accumulate, bf16 x bf16, both operands K-major (the converters transpose MN-major raw tiles)
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);          // .x = a (low half), round to nearest even
  return *reinterpret_cast<const uint32_t*>(&v);
}

// ---- bf16 mode (hsg_set_gemm_mode(3)): operand conversion ---------------------------------------------------------
// Both raw tile kinds are rewritten by the converter warps into ONE bf16 layout, K-major SWIZZLE_128B with only the
// first 64 bytes (32 bf16) of every 128-byte row in use:   element (r, k) at  r*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2.
// The tensor core then reads it with the same descriptors as the fp32 K-major tiles (8-row groups 1024 B apart, a
// k-step of 16 bf16 = 32 B inside the swizzled row), kind::f16, two k-steps per 32-deep k-block.
// raw K-major tile (box {32 k, rows}, SWIZZLE_128B): one 16-byte chunk per work item, linear sweep
__device__ __forceinline__ void bf16_from_kmajor(const char* raw, char* out, int rows, int ct, int nthreads) {
  for (int id = ct; id < rows * 8; id += nthreads) {
    const int r = id >> 3, c = (id & 7) ^ (r & 7);               // logical chunk c = fp32 elements 4c .. 4c+3 of row r
    const float4 v = *reinterpret_cast<const float4*>(raw + id * 16);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8) = o;
  }
}
// raw MN-major tile (boxes {32 mn, 32 k}, SWIZZLE_128B_ATOM_32B, atoms 4096 B apart): work item = (column mn, group of
// 8 k); lanes walk consecutive mn, so the eight loads of a warp each read one 128-byte row (conflict-free) and the
// 16-byte stores of a quarter warp fall into eight different chunk positions
__device__ __forceinline__ void bf16_from_mnmajor(const char* raw, char* out, int mn_ext, int ct, int nthreads) {
  for (int id = ct; id < mn_ext * 4; id += nthreads) {
    const int g = id / mn_ext, m = id - g * mn_ext;
    const char* col = raw + (m >> 5) * 4096 + (m & 7) * 4;
    const int j = (m & 31) >> 3;
    float f[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = 8 * g + t;
      f[t] = *reinterpret_cast<const float*>(col + k * 128 + ((j ^ (k & 3)) << 5));
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + m * 128 + ((g ^ (m & 7)) << 4)) = o;
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// hi = x rounded to nearest TF32 (10 explicit mantissa bits), lo = x - hi (exact in fp32, <= 12 significant
// bits, so the tensor core's own TF32 truncation of lo costs at most one bit: ~2^-23 relative overall).
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// hi as the tensor core itself sees a raw fp32 operand: kind::tf32 ignores the low 13 mantissa bits (truncation).  With
// HSG_TC_HI_INPLACE == 0 the converters leave the landed tile untouched (it IS the hi operand) and only write
// lo = x - trunc(x): one shared-memory store per chunk instead of two.  |lo| < 2^-10 |x| instead of 2^-11 with
// round-to-nearest; the dropped lo*lo term and the hardware's truncation of lo stay at ~2^-20 relative.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
#ifndef HSG_TC_HI_INPLACE
#define HSG_TC_HI_INPLACE 0
#endif

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

// Shared-memory tile layouts produced by the TMA boxes (all offsets relative to a 1024-byte aligned tile):
//   K-major  operand: one box {32 k, rows}: row r at r*128, 16-byte chunk c at c ^ (r%8)          (SWIZZLE_128B)
//   MN-major operand: one box {32 mn, 32 k} per 32-wide MN atom, atoms 4096 B apart; k row at k*128, 32-byte
//                     unit j at j ^ (k%4)                                                   (SWIZZLE_128B_ATOM_32B)
//
// Persistent kernel: grid = min(#tiles, #SMs); CTA b processes tiles b, b+grid, ... (N fastest, so the CTAs that
// run together share the A tile through L2).  Two TMEM accumulator sets: the epilogue warps drain set a while the
// MMA warp already accumulates the next tile into set a^1.
// optional pipeline trace of CTA 0 (debug / profiling aid): (event id, k-block counter, clock) triples
__device__ unsigned long long g_trace[3 * 2048];
__device__ unsigned int g_trace_n;
__device__ int g_trace_on;
__device__ __forceinline__ void trace(int ev, uint32_t it) {   // fire-and-forget stores: slot = ev * 256 + it
  if (g_trace_on && blockIdx.x == 0 && it < 256) {
    const unsigned int i = (unsigned int)ev * 256u + it;
    g_trace[3 * i] = (unsigned long long)ev;
    g_trace[3 * i + 1] = it;
    g_trace[3 * i + 2] = clock64();
  }
}

struct TileInfo {
  int m0, n0, n_valid, n_mma, k_beg, nkb, z, chunk;
};

// Work item w = (n tile, m tile, split z) as before; with k_chunk > 0 a split's K range is cut into chunks of k_chunk
// and tile index t = chunk * n_items + w: launched with exactly n_items CTAs, CTA w walks ALL chunks of its own item
// one after the other and its epilogue ADDS chunk > 0 onto what it stored for chunk 0.  The tensor core's fp32
// accumulation truncates, so the error of one accumulation chain grows linearly with its length (measured on the
// weight-gradient products: 2.8e-6 at 1 k rows per chain, 2.2e-5 at 8 k, 1.5e-4 at 63 k - a 2 048-graph shard);
// chunks keep every chain at <= k_chunk / 8 updates and the chunks are summed in fp32 with round-to-nearest.
__device__ __forceinline__ TileInfo tile_info(int t, int m_tiles, int n_tiles, int Nd, int bn, int K, int k_per_split,
                                              int k_chunk, int n_items) {
  TileInfo ti;
  ti.chunk = k_chunk > 0 ? t / n_items : 0;
  const int w = k_chunk > 0 ? t - ti.chunk * n_items : t;
  const int nt = w % n_tiles;
  const int r = w / n_tiles;
  const int mt = r % m_tiles;
  ti.z = r / m_tiles;
  ti.m0 = mt * TM;
  ti.n0 = nt * bn;
  ti.n_valid = min(bn, Nd - ti.n0);
  ti.n_mma = (ti.n_valid + 15) & ~15;
  const int split_end = min(K, ti.z * k_per_split + k_per_split);
  ti.k_beg = ti.z * k_per_split + ti.chunk * k_chunk;
  const int k_end = k_chunk > 0 ? min(split_end, ti.k_beg + k_chunk) : split_end;
  ti.nkb = k_end > ti.k_beg ? (k_end - ti.k_beg + BK - 1) / BK : 0;
  return ti;
}

template <bool A_MN, bool B_MN>
// Register cap (HSG_TC_MAXREG, 0 = none, the default): at 168 registers x 384 threads a CTA of this kernel takes the
// whole register file of its SM, so while a weight-gradient product runs on the side stream no CTA of the caller's
// chain can be placed next to it.  Measured with a cap of 112 (gpurun r02t): edge_bwd_blockrow then does run next to the
// dW2 product (62 -> 17.5 us in the step), but the cluster kernel and ffn_rows_bwd behind it (shared memory) still wait
// for whole CTAs to retire, and the capped kernel spills (40 B) and is 6 % slower: step 0.6045 against 0.589 ms.
#ifndef HSG_TC_MAXREG
#define HSG_TC_MAXREG 0
#endif
#if HSG_TC_MAXREG > 0
__global__ void __maxnreg__(HSG_TC_MAXREG)
#else
__global__ void __launch_bounds__(THREADS, 1)
#endif
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Md, int Nd,
               int K, int bn, int nb_box, int k_per_split, int m_tiles, int n_tiles, int total_tiles, int precise,
               int k_chunk, int n_items, Epilogue ep) {
  extern __shared__ char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // 4 warps x 32 x 33 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool want_lo = precise == 1;                     // 3xTF32
  const bool bf16 = precise == 2;                        // bf16 operands, one kind::f16 product
  const bool use_conv = want_lo || bf16 || ep.ones_col >= 0;   // converter warps touch the stage before the MMAs
  // accumulator set a: main columns [a*acc_cols, +128), precise mode adds correction columns [.. +128, +256):
  // the tensor core's fp32 accumulation truncates, so the small hi*lo / lo*hi products get their own accumulator
  // (K/8 updates on the main one instead of 3K/8) and are added in the epilogue.
  const uint32_t acc_cols = want_lo ? 2u * BN_MAX : (uint32_t)BN_MAX;
  const uint32_t tmem_cols = 2u * acc_cols;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_ready = smem_u32(&bars[STAGES]),
                 bar_empty = smem_u32(&bars[2 * STAGES]), bar_tfull = smem_u32(&bars[3 * STAGES]),
                 bar_tempty = smem_u32(&bars[3 * STAGES + 2]);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);                  // TMA: one arrive.expect_tx + the transaction bytes
      mbar_init(bar_ready + 8 * i, NCONV);             // converters done with the stage
      mbar_init(bar_empty + 8 * i, 1);                 // tcgen05.commit: MMAs drained the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);                 // tcgen05.commit: accumulator set complete
      mbar_init(bar_tempty + 8 * i, NEPI_WARPS);       // epilogue warps drained the accumulator set
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: everything above (TMEM allocation, barrier init, tensor-map prefetch) overlapped the previous kernel's tail;
  // from here on global memory written by earlier kernels is read, so wait for them.
  pdl_prologue();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t stage_tx = (uint32_t)(TM * 128 + nb_box * 128);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = tile_info(t, m_tiles, n_tiles, Nd, bn, K, k_per_split, k_chunk, n_items);
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + 8 * slot, ph ^ 1);                          // MMAs drained the slot
          trace(1, it);
          const uint32_t st = smem_u32(smem + slot * STAGE_BYTES);
          const int k0 = ti.k_beg + kb * BK;
          const uint32_t full = bar_full + 8 * slot;
          mbar_expect_tx(full, stage_tx);
          if (A_MN) {
#pragma unroll
            for (int a = 0; a < TM / 32; ++a) tma_load_2d(st + a * 4096, &tmA, ti.m0 + 32 * a, k0, full);
          } else {
            tma_load_2d(st, &tmA, k0, ti.m0, full);
          }
          if (B_MN) {
            for (int a = 0; a < nb_box / 32; ++a)
              tma_load_2d(st + 2 * TILE_BYTES + a * 4096, &tmB, ti.n0 + 32 * a, k0, full);
          } else {
            tma_load_2d(st + 2 * TILE_BYTES, &tmB, k0, ti.n0, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // K-major: 8-row groups 1024 B apart, a k-step of 8 fp32 = 32 B inside the swizzled row.
      // MN-major: MN atoms 4096 B apart (LBO), 4-k groups 512 B apart (SBO), a k-step of 8 = two groups = 1024 B.
      const uint32_t a_sbo = A_MN ? 512u : 1024u, a_lbo = A_MN ? 4096u : 16u, a_lay = A_MN ? 1u : 2u;
      const uint32_t b_sbo = B_MN ? 512u : 1024u, b_lbo = B_MN ? 4096u : 16u, b_lay = B_MN ? 1u : 2u;
      uint32_t it = 0, tl = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
        const TileInfo ti = tile_info(t, m_tiles, n_tiles, Nd, bn, K, k_per_split, k_chunk, n_items);
        const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);                           // epilogue drained this accumulator set
        tc_fence_after();
        const uint32_t d_main = tmem_d + acc * acc_cols, d_corr = d_main + BN_MAX;
        const uint32_t idesc = make_idesc(A_MN, B_MN, ti.n_mma);
        // precise mode: A_hi.B_hi and A_hi.B_lo in ONE instruction of N = 128 + n_mma - the B_lo tile directly
        // follows the B_hi tile in shared memory and the correction accumulator directly follows the main one in
        // TMEM, so [B_hi | B_lo] is one operand and [main | corr] one accumulator: A_hi is read from shared memory
        // once instead of twice (the pipeline is shared-memory-bandwidth bound).  Columns n_mma..127 are scratch.
        const uint32_t idesc_wide = make_idesc(A_MN, B_MN, BN_MAX + ti.n_mma);
        const uint32_t idesc_bf = make_idesc_bf16(ti.n_mma);
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait((use_conv ? bar_ready : bar_full) + 8 * slot, ph);
          trace(2, it);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + slot * STAGE_BYTES), a_lo = a_hi + TILE_BYTES;
          const uint32_t b_hi = a_hi + 2 * TILE_BYTES, b_lo = b_hi + TILE_BYTES;
          if (bf16) {                                      // the converted tiles live where the lo planes would be
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks)
              tc_mma_bf16(d_main, make_desc(a_lo + ks * 32u, 16u, 1024u, 2u), make_desc(b_lo + ks * 32u, 16u, 1024u, 2u),
                          idesc_bf, (kb > 0 || ks > 0) ? 1u : 0u);
            tc_commit(bar_empty + 8 * slot);
            trace(3, it);
            continue;
          }
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t a_off = A_MN ? ks * 1024u : ks * 32u;
            const uint32_t b_off = B_MN ? ks * 1024u : ks * 32u;
            const uint64_t dah = make_desc(a_hi + a_off, a_lbo, a_sbo, a_lay),
                           dbh = make_desc(b_hi + b_off, b_lbo, b_sbo, b_lay);
            if (want_lo) {
              const uint64_t dal = make_desc(a_lo + a_off, a_lbo, a_sbo, a_lay);
              tc_mma_tf32(d_main, dah, dbh, idesc_wide, (kb > 0 || ks > 0) ? 1u : 0u);   // [main | corr] (+)= A_hi [B_hi | B_lo]
              tc_mma_tf32(d_corr, dal, dbh, idesc, 1u);                                   // corr += A_lo B_hi
            } else {
              tc_mma_tf32(d_main, dah, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            }
          }
          tc_commit(bar_empty + 8 * slot);                                  // frees the slot when the MMAs retire
          trace(3, it);
        }
        tc_commit(bar_tfull + 8 * acc);                                     // accumulators of this tile complete
      }
    }
  } else if (warp < 8) {
    // ===== converters (warps 2-7): linear hi/lo sweep over the landed stage =====
    if (use_conv) {
      const int ct = tid - 64;
      const int b_chunks = nb_box * 8;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = tile_info(t, m_tiles, n_tiles, Nd, bn, K, k_per_split, k_chunk, n_items);
        const bool ones_here = B_MN && ep.ones_col >= ti.n0 && ep.ones_col < ti.n0 + nb_box;
        const int k_end = min(min(K, ti.z * k_per_split + k_per_split), k_chunk > 0 ? ti.k_beg + k_chunk : K);
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * slot, ph);                                // this k-block has landed
          if (ct == 0) trace(4, it);
          char* st = smem + slot * STAGE_BYTES;
          if (ones_here) {                                                   // all-ones column (TMA zero-filled it)
            if (ct < 32 && ti.k_beg + kb * BK + ct < k_end) {
              const int nl = ep.ones_col - ti.n0, k = ct;
              const uint32_t off = (uint32_t)(nl >> 5) * 4096u + (uint32_t)k * 128u +
                                   (uint32_t)((((nl & 31) >> 3) ^ (k & 3)) << 5) + (uint32_t)(nl & 7) * 4u;
              *reinterpret_cast<float*>(st + 2 * TILE_BYTES + off) = 1.f;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NCONV) : "memory");
          }
          if (bf16) {
            if (A_MN) bf16_from_mnmajor(st, st + TILE_BYTES, TM, ct, NCONV);
            else      bf16_from_kmajor(st, st + TILE_BYTES, TM, ct, NCONV);
            if (B_MN) bf16_from_mnmajor(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, nb_box, ct, NCONV);
            else      bf16_from_kmajor(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, nb_box, ct, NCONV);
          }
          if (want_lo) {
            // (128 + nb_box) * 8 <= 2048 chunks of 16 B over 192 threads: up to 11 per thread, loads issued first
            constexpr int CPT = (TM * 8 + BN_MAX * 8 + NCONV - 1) / NCONV;
            const int total = TM * 8 + b_chunks;
            float4 v[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
              const int id = ct + i * NCONV;
              const char* hi = id < TM * 8 ? st + id * 16 : st + 2 * TILE_BYTES + (id - TM * 8) * 16;
              v[i] = id < total ? *reinterpret_cast<const float4*>(hi) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
              const int id = ct + i * NCONV;
              if (id < total) {
                char* hi = id < TM * 8 ? st + id * 16 : st + 2 * TILE_BYTES + (id - TM * 8) * 16;
#if HSG_TC_HI_INPLACE
                const float4 h = make_float4(tf32_rn(v[i].x), tf32_rn(v[i].y), tf32_rn(v[i].z), tf32_rn(v[i].w));
                *reinterpret_cast<float4*>(hi) = h;
#else
                const float4 h = make_float4(tf32_trunc(v[i].x), tf32_trunc(v[i].y), tf32_trunc(v[i].z),
                                             tf32_trunc(v[i].w));
#endif
                *reinterpret_cast<float4*>(hi + TILE_BYTES) =
                    make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
              }
            }
          }
          fence_async_smem();                                                // generic-proxy writes -> async proxy
          if (ct == 0) trace(5, it);
          mbar_arrive(bar_ready + 8 * slot);
        }
      }
    }
  } else {
    // ===== epilogue (warps 8-11): TMEM -> registers -> global, overlapped with the next tile's main loop =====
    const int lg = warp & 3;                     // TMEM lane group of this warp (warp 8 -> lanes 0..31, ...)
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const TileInfo ti = tile_info(t, m_tiles, n_tiles, Nd, bn, K, k_per_split, k_chunk, n_items);
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, aph);
      if (tid == 256) trace(6, tl);
      tc_fence_after();
      const uint32_t d_main = tmem_d + acc * acc_cols + ((uint32_t)(lg * 32) << 16);
      float* Dz = ep.D + (size_t)ti.z * ep.split_stride;
      const int n_lim = ti.n0 + ti.n_valid;
      const int n_out = min(ep.ones_col >= 0 ? ep.ones_col : Nd, n_lim);   // real output columns of THIS tile
      // Each thread holds one accumulator ROW (32 columns per TMEM load).  Writing rows from registers would make
      // every store instruction touch 32 different lines with 16 B each (partial sectors: ~6x slower than the whole
      // main loop, measured).  So each 32x32 chunk is transposed through a padded shared-memory tile and leaves the
      // SM as full 128-byte lines: 8 lanes x 16 B per row, 4 rows per store instruction.
      float* tile = epi_stage + (warp - 8) * (32 * 33);
      const int r_sub = lane >> 3, c_sub = (lane & 7) * 4;
      // chunk > 0 of a K-chunked item: add onto what THIS thread stored for the previous chunk (plain loads: the data
      // was written by this kernel, the read-only path must not be used for it)
      const bool acc_chunk = ti.chunk > 0;
      const float* r_base = acc_chunk ? Dz : ep.R;
      const int r_ld = acc_chunk ? ep.ldd : ep.ldr;
      const bool has_r = acc_chunk || (ep.epi & (HSG_EPI_ADD | HSG_EPI_RELU_MASK)) != 0;
      const bool ld_vec = (ep.ldd & 3) == 0, lr_vec = (r_ld & 3) == 0;
      for (int c0 = 0; c0 < ti.n_mma; c0 += 32) {
        const int col = ti.n0 + c0 + c_sub;
        // residual / ReLU-mask operand of the whole 32x32 chunk first: eight independent 16-byte loads per lane in
        // flight while the accumulators come out of TMEM (one exposed L2/HBM latency per chunk instead of eight)
        float4 rv4[8];
        if (has_r) {
#pragma unroll
          for (int p8 = 0; p8 < 8; ++p8) {
            const int row = ti.m0 + lg * 32 + p8 * 4 + r_sub;
            rv4[p8] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < Md) {
              const float* rp = r_base + (size_t)row * r_ld + col;
              if (acc_chunk) {
                if (lr_vec && col + 3 < n_out) {
                  rv4[p8] = *reinterpret_cast<const float4*>(rp);
                } else {
                  if (col + 0 < n_out) rv4[p8].x = rp[0];
                  if (col + 1 < n_out) rv4[p8].y = rp[1];
                  if (col + 2 < n_out) rv4[p8].z = rp[2];
                  if (col + 3 < n_out) rv4[p8].w = rp[3];
                }
              } else if (lr_vec && col + 3 < n_out) {
                rv4[p8] = __ldg(reinterpret_cast<const float4*>(rp));
              } else {
                if (col + 0 < n_out) rv4[p8].x = __ldg(rp + 0);
                if (col + 1 < n_out) rv4[p8].y = __ldg(rp + 1);
                if (col + 2 < n_out) rv4[p8].z = __ldg(rp + 2);
                if (col + 3 < n_out) rv4[p8].w = __ldg(rp + 3);
              }
            }
          }
        }
        // bias of this lane's four columns: loaded once per chunk (not once per row group)
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (ep.epi & HSG_EPI_BIAS) {
#pragma unroll
          for (int tt = 0; tt < 4; ++tt)
            if (col + tt < n_out) bz[tt] = __ldg(ep.bias + col + tt);
        }
        float v[32];
        if (ti.nkb > 0) {
          tc_ld32(d_main + (uint32_t)c0, v);
          if (want_lo) {
            float w[32];
            tc_ld32(d_main + (uint32_t)(BN_MAX + c0), w);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = v[i];           // bank (lane + i) % 32: conflict-free
        __syncwarp();
        const bool do_relu = (ep.epi & HSG_EPI_RELU) != 0, do_add = acc_chunk || (ep.epi & HSG_EPI_ADD) != 0,
                   do_mask = (ep.epi & HSG_EPI_RELU_MASK) != 0;
        const bool vec_col = col + 3 < n_out && ld_vec;
        const int ones_d = ep.colsum_part != nullptr ? ep.ones_col - col : -1;   // 0..3 when this lane holds the sums
        // everything below stays in registers: no dynamically indexed arrays (they would live in local memory)
#pragma unroll
        for (int p8 = 0; p8 < 8; ++p8) {
          const int rl = p8 * 4 + r_sub;
          const int row = ti.m0 + lg * 32 + rl;
          float o0 = tile[rl * 33 + c_sub + 0], o1 = tile[rl * 33 + c_sub + 1], o2 = tile[rl * 33 + c_sub + 2],
                o3 = tile[rl * 33 + c_sub + 3];
          if (ones_d >= 0 && ones_d < 4 && row < Md)
          {
            float* cp = ep.colsum_part + (size_t)ti.z * Md + row;
            const float cv = ones_d == 0 ? o0 : (ones_d == 1 ? o1 : (ones_d == 2 ? o2 : o3));
            *cp = acc_chunk ? *cp + cv : cv;
          }
          o0 += bz[0]; o1 += bz[1]; o2 += bz[2]; o3 += bz[3];
          if (do_relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f); }
          if (do_add) { o0 += rv4[p8].x; o1 += rv4[p8].y; o2 += rv4[p8].z; o3 += rv4[p8].w; }
          if (do_mask) {
            o0 = rv4[p8].x > 0.f ? o0 : 0.f; o1 = rv4[p8].y > 0.f ? o1 : 0.f;
            o2 = rv4[p8].z > 0.f ? o2 : 0.f; o3 = rv4[p8].w > 0.f ? o3 : 0.f;
          }
          if (row < Md) {
            float* dp = Dz + (size_t)row * ep.ldd + col;
            if (vec_col) {
              *reinterpret_cast<float4*>(dp) = make_float4(o0, o1, o2, o3);
            } else {
              if (col + 0 < n_out) dp[0] = o0;
              if (col + 1 < n_out) dp[1] = o1;
              if (col + 2 < n_out) dp[2] = o2;
              if (col + 3 < n_out) dp[3] = o3;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (tid == 256) trace(7, tl);
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);                     // hand the accumulator set back
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
}

// ---- host side: tensor maps + launch -----------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

typedef CUresult (*ReplaceAddrFn)(CUtensorMap*, void*);

static ReplaceAddrFn replace_fn() {
  static ReplaceAddrFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapReplaceAddress", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<ReplaceAddrFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [outer, inner] with row pitch ld (elements); box = {32 inner, box_outer}; elements outside
// [inner, outer] are zero-filled by the TMA unit.  Encoded maps are cached by geometry (the same few shapes recur
// every step); only the base address is patched per call.
struct MapKey {
  int inner, outer, ld, box_outer, mn;
  bool operator==(const MapKey& o) const {
    return inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer && mn == o.mn;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.inner * 1000003u;
    h = (h ^ (size_t)k.outer) * 1000003u;
    h = (h ^ (size_t)k.ld) * 1000003u;
    h = (h ^ (size_t)k.box_outer) * 1000003u;
    return h ^ (size_t)k.mn;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

static bool make_map(CUtensorMap* m, const float* ptr, int inner, int outer, int ld, int box_outer, bool mn_major) {
  static const bool no_cache = getenv("HSG_TMAP_NOCACHE") != nullptr;     // debugging aid: encode every map afresh
  ReplaceAddrFn rep = no_cache ? nullptr : replace_fn();
  const MapKey key{inner, outer, ld, box_outer, mn_major ? 1 : 0};
  if (rep) {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *m = it->second;
      return rep(m, const_cast<float*>(ptr)) == CUDA_SUCCESS;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  const bool ok =
      fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  if (ok && rep) {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() < 4096) g_maps.emplace(key, *m);
  }
  return ok;
}

// shared with the CTA-pair kernel (hsg_gemm_tc2.cu)
bool make_tensor_map(CUtensorMap* m, const float* ptr, int inner, int outer, int ld, int box_outer, bool mn_major) {
  return make_map(m, ptr, inner, outer, ld, box_outer, mn_major);
}

struct Operand {
  const float* p;
  int ld;
  int ext_mn;   // valid extent along the MMA M / N dimension
  int ext_k;    // valid extent along K
};

static bool g_attr_done[3] = {false, false, false};

// longest accumulation chain of one TMEM accumulator, in K elements (see tile_info): weight-gradient splits longer
// than this are walked in chunks by their CTA
constexpr int K_CHUNK = 1024;

template <bool A_MN, bool B_MN>
static int launch(int which, dim3 grid, Operand A, Operand B, int Md, int Nd, int K, int bn, int k_per_split,
                  int precise, Epilogue ep, cudaStream_t s, int k_chunk = 0, bool cta_per_item = false) {
  if (!g_attr_done[which]) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess)
      return HSG_ERR_CUDA;
    g_attr_done[which] = true;
  }
  const int nb_box = B_MN ? ((bn + 31) & ~31) : ((bn + 15) & ~15);      // staged B rows / columns per stage
  CUtensorMap tmA, tmB;
  const bool okA = A_MN ? make_map(&tmA, A.p, A.ext_mn, A.ext_k, A.ld, 32, true)
                        : make_map(&tmA, A.p, A.ext_k, A.ext_mn, A.ld, TM, false);
  const bool okB = B_MN ? make_map(&tmB, B.p, B.ext_mn, B.ext_k, B.ld, 32, true)
                        : make_map(&tmB, B.p, B.ext_k, B.ext_mn, B.ld, nb_box, false);
  if (!okA || !okB) return HSG_ERR_CUDA;
  const int m_tiles = (int)grid.x, n_tiles = (int)grid.y, n_items = m_tiles * n_tiles * (int)grid.z;
  int total = n_items, ctas = n_items < num_sms() ? n_items : num_sms();
  if (cta_per_item) {
    // short items on a low-priority stream: one CTA each, so that SMs return to the block scheduler every item
    ctas = n_items;
    if (k_per_split > k_chunk) k_chunk = 0;        // (items are planned <= K_CHUNK rows; never chunk across CTAs)
  }
  if (k_chunk > 0) {
    // chunked items: CTA w must own every chunk of item w (its epilogue adds them up in place), i.e. exactly one CTA
    // per item; the plan keeps n_items within one wave, otherwise run unchunked
    if (n_items <= num_sms() && k_per_split > k_chunk) {
      total = n_items * ((k_per_split + k_chunk - 1) / k_chunk);
      ctas = n_items;
    } else {
      k_chunk = 0;
    }
  }
  launch_k(gemm_tc_kernel<A_MN, B_MN>, dim3(ctas), dim3(THREADS), SMEM_BYTES, s, tmA, tmB, Md, Nd, K, bn, nb_box, k_per_split, m_tiles,
                                                               n_tiles, total, precise, k_chunk, n_items, ep);
  return check_launch();
}

static int pick_bn(int n_total) {
  // D column tile: a multiple of 16 up to BN_MAX, tiles as even as possible
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
            int splits, int rows_per_split, int precise, cudaStream_t s, bool cta_per_item) {
  Operand a{A, lda, N1, M}, b{B, ldb, N2, M};
  const int n_total = N2 + (part_col ? 1 : 0);
  const int bn = pick_bn(n_total);
  Epilogue ep{part, N2, (size_t)N1 * N2, nullptr, nullptr, 0, 0, part_col, part_col ? N2 : -1};
  dim3 grid(ceil_div(N1, TM), ceil_div(n_total, bn), splits);
  return launch<true, true>(2, grid, a, b, N1, n_total, M, bn, rows_per_split, precise, ep, s, precise ? K_CHUNK : 0,
                            cta_per_item);
}

int trace_ctl(int on, unsigned long long* host_out, int max_events) {
  if (on >= 0) {
    static unsigned long long zeros[3 * 2048];
    cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
    cudaMemcpyToSymbol(g_trace_on, &on, sizeof(on));
    return 0;
  }
  unsigned int n = 2048;
  if ((int)n > max_events) n = max_events;
  if (host_out && n) cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * 3 * sizeof(unsigned long long));
  return (int)n;
}

}  // namespace tc
}  // namespace hsg
