// CTA-pair (tcgen05 cta_group::2) version of the tall-skinny fp32-parity products  C = A . B^T  (NT) and  C = A . B
// (NN) of the WSWGAT path: FFN-1 / FFN-2 / the projections and their input-gradient products (GATLayer.py:38,110,146).
//
// Why a second kernel next to hsg_gemm_tc.cu: the single-CTA 3xTF32 pipeline is bound by shared-memory bandwidth
// (per 32-deep k-block of a 128x128 tile: TMA write 32 KB, hi/lo sweep 64 KB, operand reads 80 KB = 176 KB against
// 768 tensor-core cycles) and holds only three 64 KB stages.  Here two CTAs of a cluster share ONE 256 x BN tile:
// each CTA stages its own 128 rows of A but only HALF of the B tile (BN/2 rows / columns), the other half is read by
// the pair's MMA straight from the peer's shared memory.  Per CTA and k-block: TMA write 24 KB, hi/lo sweep 48 KB,
// operand reads 72 KB = 144 KB, 24 KB from L2 instead of 32, and a stage is 48 KB (four stages fit; three are used - see HSG_TC2_STAGES below).
//
//   * both CTAs: one thread issues TMA loads of the CTA's own A rows and B half into its own shared memory;
//     converter warps split the landed tiles into hi (= the raw tile: kind::tf32 truncates) and lo = x - hi;
//     one converter thread per CTA then arrives on the LEADER's `ready` barrier (cluster-scope release);
//   * leader CTA only: one thread issues   main += A_hi B_hi ; corr += A_hi B_lo ; corr += A_lo B_hi   with
//     tcgen05.mma.cta_group::2.kind::tf32 (M = 256: 128 TMEM lanes in each CTA) and hands stages / accumulator sets
//     back with tcgen05.commit ... multicast::cluster to the barriers of BOTH CTAs;
//   * both CTAs: four epilogue warps drain the CTA's own 128 accumulator rows (same epilogue as hsg_gemm_tc.cu) and
//     arrive on the leader's `tempty` barrier.
// Arithmetic is identical to the single-CTA kernel product by product (same hi/lo split, same k order, separate
// correction accumulator), so results are bit-identical to it; tests/test_gpu_parity.py runs both.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "hsg_common.cuh"

namespace hsg {
namespace tc {
bool make_tensor_map(CUtensorMap* m, const float* ptr, int inner, int outer, int ld, int box_outer, bool mn_major);
}
namespace tc2 {

constexpr int TM = 128;
constexpr int BK = 32;
constexpr int BN_MAX = 128;
constexpr int THREADS = 384;     // warp 0: TMA, warp 1: MMA (leader), warps 2-7: converters, warps 8-11: epilogue
constexpr int NCONV = 192;
constexpr int NEPI_WARPS = 4;
constexpr int A_TILE = TM * 128;                 // 16 KB
constexpr int B_HALF = (BN_MAX / 2) * 128;       // 8 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_HALF;   // A_raw | A_lo | B_raw | B_lo = 48 KB
// Three stages and a 112-register cap (no spills; HSG_TC2_STAGES=4 -DHSG_TC2_MAXREG=0 restores the first version): the
// kernel's own time is the same as with four stages and 151 registers (32.7-33.8 us on the four FFN shapes at 11 817
// rows; +-6 % per shape at 731 k rows, equal in sum), but a 178 KB / 43 k-register CTA leaves room on its SM for a CTA of
// the graph builder (build_fill: 40 KB, 10 k registers), which otherwise holds 32 SMs for ~130 us next to the word-side
// FFN products: FFN-2 of the forward 46 -> 40 us inside the step, e2e 53.3 -> 54.2 k graphs/s (gpurun r02u).
#ifndef HSG_TC2_STAGES
#define HSG_TC2_STAGES 3
#endif
#ifndef HSG_TC2_MAXREG
#define HSG_TC2_MAXREG 112
#endif
constexpr int STAGES = HSG_TC2_STAGES;
constexpr int EPI_BUF = 32 * 128;                                // one 32 x 32 fp32 chunk, SWIZZLE_128B rows of 128 B
constexpr int EPI_BYTES = NEPI_WARPS * 2 * EPI_BUF;              // two chunks in flight per epilogue warp
constexpr int BARS_BYTES = 256;                                  // barriers + TMEM slot
constexpr int BIAS_BYTES = 2 * BN_MAX * 4;                       // the tile's bias slice, double-buffered by tile parity
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + EPI_BYTES + BARS_BYTES + BIAS_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

// wait on a barrier of THIS CTA (arrivals may come from the peer CTA; default .acquire.cta semantics like CUTLASS's
// ClusterBarrier - a cluster-scope acquire / release pair measured ~2 000 cycles per k-block on the arriving thread)
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
    if (!done && ++spins > (1ull << 24)) __trap();   // never hang the GPU: fail loudly instead
  }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_rank(uint32_t local_bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar), "r"(rank)
      : "memory");
}

// all MMAs issued so far by this thread retire -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// bf16 x bf16 -> fp32 (kind::f16), K = 16 per instruction, M = 256 over the pair
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// fp32 accumulate, tf32 x tf32, A K-major, B K- or MN-major, M = 256 over the pair, N = n
__device__ __forceinline__ uint32_t make_idesc(bool b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

// bf16 mode (see hsg_gemm_tc.cu): both operands K-major bf16 in the first 64 bytes of SWIZZLE_128B rows
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void bf16_from_kmajor(const char* raw, char* out, int rows, int ct, int nthreads) {
  for (int id = ct; id < rows * 8; id += nthreads) {
    const int r = id >> 3, c = (id & 7) ^ (r & 7);
    const float4 v = *reinterpret_cast<const float4*>(raw + id * 16);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8) = o;
  }
}
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

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// shared memory (SWIZZLE_128B box {32 fp32, 32 rows}) -> global, clipped at the tensor bounds by the TMA unit
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// optional pipeline trace of the first cluster (debug / profiling aid): slot = (ev + 10 * rank) * 128 + it
__device__ unsigned long long g_trace2[3 * 4096];
__device__ int g_trace2_on;
__device__ __forceinline__ void trace(int ev, uint32_t it, uint32_t rank) {
  if (g_trace2_on && blockIdx.x < 2 && it < 128) {
    const unsigned int i = ((unsigned int)ev + 10u * rank) * 128u + it;
    g_trace2[3 * i] = (unsigned long long)(ev + 10 * rank);
    g_trace2[3 * i + 1] = it;
    g_trace2[3 * i + 2] = clock64();
  }
}

// per-CTA span of the traced launch: (globaltimer at entry, globaltimer at exit, SM id) - shows whether all 74 CTA
// pairs of the persistent grid are co-resident or some start only after others have finished
__device__ unsigned long long g_span2[3 * 160];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Epilogue {
  float* D;
  int ldd;
  const float* bias;
  const float* R;
  int ldr;
  int epi;
};

struct TileInfo {
  int m0, n0, n_valid, n_mma, nkb;
};

// tile t = (n tile fastest, 256-row m tile); this CTA owns rows m0 = 256 * mt + 128 * rank
__device__ __forceinline__ TileInfo tile_info(int t, int n_tiles, int Nd, int bn, int K, int rank) {
  TileInfo ti;
  const int nt = t % n_tiles, mt = t / n_tiles;
  ti.m0 = mt * 2 * TM + rank * TM;
  ti.n0 = nt * bn;
  ti.n_valid = min(bn, Nd - ti.n0);
  ti.n_mma = (ti.n_valid + 15) & ~15;          // cta_group::2: N in steps of 16, each CTA stages n_mma / 2
  ti.nkb = (K + BK - 1) / BK;
  return ti;
}

template <bool B_MN>
#if HSG_TC2_MAXREG > 0
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(HSG_TC2_MAXREG)
#else
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
#endif
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, int Md, int Nd, int K,
                int bn, int nb_half_box, int n_tiles, int total_tiles, int precise, Epilogue ep) {
  extern __shared__ char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* epi_stage = smem + STAGES * STAGE_BYTES;                                    // 1 KB aligned: 8 swizzled 4 KB buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  float* bias_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + EPI_BYTES + BARS_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (g_trace2_on && tid == 0 && blockIdx.x < 160) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_span2[3 * blockIdx.x] = globaltimer_ns();
    g_span2[3 * blockIdx.x + 2] = smid;
  }
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool want_lo = precise == 1;               // 3xTF32
  const bool bf16 = precise == 2;                  // bf16 operands, one kind::f16 product
  const uint32_t acc_cols = want_lo ? 2u * BN_MAX : (uint32_t)BN_MAX;
  const uint32_t tmem_cols = 2u * acc_cols;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_ready = smem_u32(&bars[STAGES]),
                 bar_empty = smem_u32(&bars[2 * STAGES]), bar_tfull = smem_u32(&bars[3 * STAGES]),
                 bar_tempty = smem_u32(&bars[3 * STAGES + 2]);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  if (tid == 32) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);                  // local TMA: one arrive.expect_tx + the transaction bytes
      mbar_init(bar_ready + 8 * i, 2);                 // (leader's copy is used) one arrival per CTA: converters done
      mbar_init(bar_empty + 8 * i, 1);                 // multicast tcgen05.commit: the pair's MMAs drained the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);                 // multicast tcgen05.commit: accumulator set complete
      mbar_init(bar_tempty + 8 * i, 2 * NEPI_WARPS);   // (leader's copy) epilogue warps of both CTAs drained the set
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of BOTH CTAs are initialised before anyone arrives remotely; TMEM allocated
  tc_fence_after();
  pdl_prologue();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t stage_tx = (uint32_t)(TM * 128 + nb_half_box * 128);

  if (warp == 0) {
    // ===== TMA producer (each CTA: its own A rows and its own half of B) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < total_tiles; t += n_pairs) {
        const TileInfo ti = tile_info(t, n_tiles, Nd, bn, K, (int)rank);
        const int nb0 = ti.n0 + (int)rank * (ti.n_mma >> 1);                // first B row / column of this CTA's half
        // (An L2 tensor prefetch of the next tile's A rows was tried here: no measurable gain - the loop is bound by the
        // stage round trip, not by DRAM latency alone - so it is not issued.)
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + 8 * slot, ph ^ 1);
          trace(1, it, rank);
          const uint32_t st = smem_u32(smem + slot * STAGE_BYTES);
          const int k0 = kb * BK;
          const uint32_t full = bar_full + 8 * slot;
          mbar_expect_tx(full, stage_tx);
          tma_load_2d(st, &tmA, k0, ti.m0, full);
          if (B_MN) {
            for (int a = 0; a < nb_half_box / 32; ++a)
              tma_load_2d(st + 2 * A_TILE + a * 4096, &tmB, nb0 + 32 * a, k0, full);
          } else {
            tma_load_2d(st + 2 * A_TILE, &tmB, k0, nb0, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the LEADER CTA drives the tensor cores of both SMs =====
    if (leader && lane == 0) {
      const uint32_t b_sbo = B_MN ? 512u : 1024u, b_lbo = B_MN ? 4096u : 16u, b_lay = B_MN ? 1u : 2u;
      uint32_t it = 0, tl = 0;
      for (int t = pair; t < total_tiles; t += n_pairs, ++tl) {
        const TileInfo ti = tile_info(t, n_tiles, Nd, bn, K, 0);
        const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_d + acc * acc_cols, d_corr = d_main + BN_MAX;
        const uint32_t idesc = make_idesc(B_MN, ti.n_mma);
        const uint32_t idesc_bf = make_idesc_bf16(ti.n_mma);
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_ready + 8 * slot, ph);
          trace(2, it, rank);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + slot * STAGE_BYTES), a_lo = a_hi + A_TILE;
          const uint32_t b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_HALF;
          if (bf16) {                                      // the converted tiles live where the lo planes would be
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks)
              tc_mma_bf16_pair(d_main, make_desc(a_lo + ks * 32u, 16u, 1024u, 2u),
                               make_desc(b_lo + ks * 32u, 16u, 1024u, 2u), idesc_bf, (kb > 0 || ks > 0) ? 1u : 0u);
            tc_commit_pair(bar_empty + 8 * slot);
            trace(3, it, rank);
            continue;
          }
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t a_off = ks * 32u;
            const uint32_t b_off = B_MN ? ks * 1024u : ks * 32u;
            const uint64_t dah = make_desc(a_hi + a_off, 16u, 1024u, 2u),
                           dbh = make_desc(b_hi + b_off, b_lbo, b_sbo, b_lay);
            const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
            tc_mma_tf32_pair(d_main, dah, dbh, idesc, first);
            if (want_lo) {
              const uint64_t dal = make_desc(a_lo + a_off, 16u, 1024u, 2u),
                             dbl = make_desc(b_lo + b_off, b_lbo, b_sbo, b_lay);
              tc_mma_tf32_pair(d_corr, dah, dbl, idesc, first);              // corr (+)= A_hi B_lo
              tc_mma_tf32_pair(d_corr, dal, dbh, idesc, 1u);                 // corr  += A_lo B_hi
            }
          }
          tc_commit_pair(bar_empty + 8 * slot);                              // frees the slot in both CTAs
          trace(3, it, rank);
        }
        tc_commit_pair(bar_tfull + 8 * acc);                                 // accumulators complete in both CTAs
      }
    }
  } else if (warp < 8) {
    // ===== converters (warps 2-7) of each CTA: lo = x - trunc_tf32(x) for the CTA's own A tile and B half =====
    const int ct = tid - 64;
    const int b_chunks = nb_half_box * 8;
    uint32_t it = 0;
    for (int t = pair; t < total_tiles; t += n_pairs) {
      const TileInfo ti = tile_info(t, n_tiles, Nd, bn, K, (int)rank);
      for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
        const uint32_t slot = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(bar_full + 8 * slot, ph);
        if (ct == 0) trace(4, it, rank);
        char* st = smem + slot * STAGE_BYTES;
        if (bf16) {
          bf16_from_kmajor(st, st + A_TILE, TM, ct, NCONV);
          if (B_MN) bf16_from_mnmajor(st + 2 * A_TILE, st + 2 * A_TILE + B_HALF, nb_half_box, ct, NCONV);
          else      bf16_from_kmajor(st + 2 * A_TILE, st + 2 * A_TILE + B_HALF, nb_half_box, ct, NCONV);
        }
        if (want_lo) {
          // (128 + nb_half_box) * 8 <= 1536 chunks of 16 B over 192 threads: up to 8 per thread, loads issued first
          constexpr int CPT = (TM * 8 + (BN_MAX / 2) * 8 + NCONV - 1) / NCONV;
          const int total = TM * 8 + b_chunks;
          float4 v[CPT];
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const int id = ct + i * NCONV;
            const char* hi = id < TM * 8 ? st + id * 16 : st + 2 * A_TILE + (id - TM * 8) * 16;
            v[i] = id < total ? *reinterpret_cast<const float4*>(hi) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const int id = ct + i * NCONV;
            if (id < total) {
              char* lo = id < TM * 8 ? st + A_TILE + id * 16 : st + 2 * A_TILE + B_HALF + (id - TM * 8) * 16;
              *reinterpret_cast<float4*>(lo) =
                  make_float4(v[i].x - tf32_trunc(v[i].x), v[i].y - tf32_trunc(v[i].y), v[i].z - tf32_trunc(v[i].z),
                              v[i].w - tf32_trunc(v[i].w));
            }
          }
        }
        fence_async_smem();                                                  // generic-proxy writes -> async proxy
        asm volatile("bar.sync 1, %0;" ::"r"(NCONV) : "memory");             // every converter of this CTA is done
        if (ct == 0) {
          trace(5, it, rank);
          mbar_arrive_rank(bar_ready + 8 * slot, 0);                         // one arrival per CTA on the leader
          trace(6, it, rank);
        }
      }
    }
  } else {
    // ===== epilogue (warps 8-11) of each CTA: its own 128 accumulator rows =====
    // Each lane holds one accumulator ROW (32 columns per TMEM load).  The row goes (bias / ReLU applied) into a
    // SWIZZLE_128B staging chunk as eight conflict-free 16-byte stores; a residual / ReLU-mask operand is then applied
    // in a second, row-group sweep whose global loads are coalesced (8 lanes x 16 B per row); one lane hands the chunk
    // to the TMA unit, which writes full lines and clips at the matrix edge.  Two staging chunks per warp: the store
    // of chunk c is in flight while chunk c + 1 is produced.
    const int lg = warp & 3;
    char* my_stage = epi_stage + (warp - 8) * 2 * EPI_BUF;
    const int r_sub = lane >> 3, j_sub = lane & 7;
    const bool has_r = (ep.epi & (HSG_EPI_ADD | HSG_EPI_RELU_MASK)) != 0;
    const bool do_relu = (ep.epi & HSG_EPI_RELU) != 0, do_add = (ep.epi & HSG_EPI_ADD) != 0,
               do_mask = (ep.epi & HSG_EPI_RELU_MASK) != 0, do_bias = (ep.epi & HSG_EPI_BIAS) != 0;
    const bool lr_vec = (ep.ldr & 3) == 0;
    uint32_t tl = 0, chunk_no = 0;
    for (int t = pair; t < total_tiles; t += n_pairs, ++tl) {
      const TileInfo ti = tile_info(t, n_tiles, Nd, bn, K, (int)rank);
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      const int n_out = min(Nd, ti.n0 + ti.n_valid);
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (do_bias) {                                       // the tile's bias slice (4 columns per lane), fetched early
        const int c = ti.n0 + 4 * lane;
        if (c + 0 < n_out) b4.x = __ldg(ep.bias + c + 0);
        if (c + 1 < n_out) b4.y = __ldg(ep.bias + c + 1);
        if (c + 2 < n_out) b4.z = __ldg(ep.bias + c + 2);
        if (c + 3 < n_out) b4.w = __ldg(ep.bias + c + 3);
      }
      mbar_wait(bar_tfull + 8 * acc, aph);
      tc_fence_after();
      if (warp == 8 && lane == 0) trace(7, tl, rank);
      // Bias slice in shared memory, one copy per accumulator set: every epilogue warp writes the same values.  A warp
      // reaches this point for tile tl only after ALL epilogue warps released set `acc` for tile tl - 2, so nobody
      // still reads the copy being overwritten.
      float* my_bias = bias_stage + acc * BN_MAX;
      if (do_bias) {
        *reinterpret_cast<float4*>(my_bias + 4 * lane) = b4;
        __syncwarp();
      }
      const uint32_t d_main = tmem_d + acc * acc_cols + ((uint32_t)(lg * 32) << 16);
      const int row0 = ti.m0 + lg * 32;
      for (int c0 = 0; c0 < ti.n_mma; c0 += 32, ++chunk_no) {
        const int col0 = ti.n0 + c0;
        // residual / ReLU-mask operand of the whole chunk first (coalesced: 8 lanes x 16 B per row, 4 rows per load)
        float4 rv4[8];
        if (has_r) {
          const int col = col0 + 4 * j_sub;
#pragma unroll
          for (int p8 = 0; p8 < 8; ++p8) {
            const int row = row0 + p8 * 4 + r_sub;
            rv4[p8] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < Md) {
              const float* rp = ep.R + (size_t)row * ep.ldr + col;
              if (lr_vec && col + 3 < n_out) {
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
        float v[32];
        tc_ld32(d_main + (uint32_t)c0, v);
        if (want_lo) {
          float w[32];
          tc_ld32(d_main + (uint32_t)(BN_MAX + c0), w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
        char* buf = my_stage + (chunk_no & 1) * EPI_BUF;
        // the TMA store that last read this buffer (two chunks ago) must have finished reading it
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        {
          char* rowp = buf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (do_bias) {
              const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c0 + 4 * j);    // broadcast read
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            if (do_relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = o;
          }
        }
        if (has_r) {
          __syncwarp();
#pragma unroll
          for (int p8 = 0; p8 < 8; ++p8) {
            const int rl = p8 * 4 + r_sub;
            float4* sp = reinterpret_cast<float4*>(buf + rl * 128 + ((j_sub ^ (rl & 7)) << 4));
            float4 o = *sp;
            if (do_add) { o.x += rv4[p8].x; o.y += rv4[p8].y; o.z += rv4[p8].z; o.w += rv4[p8].w; }
            if (do_mask) {
              o.x = rv4[p8].x > 0.f ? o.x : 0.f; o.y = rv4[p8].y > 0.f ? o.y : 0.f;
              o.z = rv4[p8].z > 0.f ? o.z : 0.f; o.w = rv4[p8].w > 0.f ? o.w : 0.f;
            }
            *sp = o;
          }
        }
        fence_async_smem();                                 // generic-proxy writes -> async proxy (TMA store)
        __syncwarp();
        if (lane == 0 && row0 < Md && col0 < n_out) {
          tma_store_2d(&tmD, smem_u32(buf), col0, row0);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_rank(bar_tempty + 8 * acc, 0);              // hand the set back to the leader
      if (warp == 8 && lane == 0) trace(8, tl, rank);
    }
    if (lane == 0) bulk_wait_all();                         // every store of this warp has reached global memory
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
  }
  if (g_trace2_on && tid == 0 && blockIdx.x < 160) g_span2[3 * blockIdx.x + 1] = globaltimer_ns();
}

// ---- host side ---------------------------------------------------------------------------------------------------
static bool g_attr_done[2] = {false, false};

template <bool B_MN>
static int launch(int which, const float* A, int lda, const float* B, int ldb, int M, int N, int K, int bn, int precise,
                  Epilogue ep, cudaStream_t s) {
  if (!g_attr_done[which]) {
    if (cudaFuncSetAttribute(gemm_tc2_kernel<B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess)
      return HSG_ERR_CUDA;
    g_attr_done[which] = true;
  }
  // rows (NT) / columns (NN) of B one CTA stages: half of the padded tile width
  const int nb_half_box = B_MN ? ((bn / 2 + 31) & ~31) : ((bn / 2 + 7) & ~7);
  CUtensorMap tmA, tmB;
  const bool okA = tc::make_tensor_map(&tmA, A, K, M, lda, TM, false);
  const bool okB = B_MN ? tc::make_tensor_map(&tmB, B, N, K, ldb, 32, true)
                        : tc::make_tensor_map(&tmB, B, K, N, ldb, nb_half_box, false);
  CUtensorMap tmD;
  if (!okA || !okB || !tc::make_tensor_map(&tmD, ep.D, N, M, ep.ldd, 32, false)) return HSG_ERR_CUDA;
  const int m_tiles = ceil_div(M, 2 * TM), n_tiles = ceil_div(N, bn), total = m_tiles * n_tiles;
  int pairs = num_sms() / 2;
  if (total < pairs) pairs = total;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<B_MN>, tmA, tmB, tmD, M, N, K, bn, nb_half_box, n_tiles, total, precise, ep);
  return check_launch();
}

// D column tile: a multiple of 32 - the epilogue stores whole 32-column chunks and relies on the TMA unit clipping at the
// matrix edge, so only the LAST column tile may be ragged
static int pick_bn(int n_total) {
  const int tiles = ceil_div(n_total, BN_MAX);
  int bn = ceil_div(ceil_div(n_total, tiles), 32) * 32;
  return bn > BN_MAX ? BN_MAX : bn;
}

int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, const float* R, int ldr, int epi, int precise, cudaStream_t s) {
  Epilogue ep{C, ldc, bias, R, ldr, epi};
  return launch<false>(0, A, lda, B, ldb, M, N, K, pick_bn(N), precise, ep, s);
}

int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* R,
            int ldr, int epi, int precise, cudaStream_t s) {
  Epilogue ep{C, ldc, nullptr, R, ldr, epi};
  return launch<true>(1, A, lda, B, ldb, M, N, K, pick_bn(N), precise, ep, s);
}

// on == -2: copy the per-CTA spans (3 x 160 values) instead of the event trace; on == -3: returns the number of CTA
// pairs of gemm_tc2_kernel<false> the device can hold at once (cudaOccupancyMaxActiveClusters)
int trace_ctl(int on, unsigned long long* host_out, int max_events) {
  if (on == -2) {
    int n = 3 * 160 < max_events ? 3 * 160 : max_events;
    if (host_out && n > 0) cudaMemcpyFromSymbol(host_out, g_span2, (size_t)n * sizeof(unsigned long long));
    return n;
  }
  if (on == -3) {
    cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<false>, &cfg) != cudaSuccess) return -1;
    return n;
  }
  if (on >= 0) {
    static unsigned long long zeros[3 * 4096];
    cudaMemcpyToSymbol(g_trace2, zeros, sizeof(zeros));
    cudaMemcpyToSymbol(g_trace2_on, &on, sizeof(on));
    return 0;
  }
  unsigned int n = 4096;
  if ((int)n > max_events) n = max_events;
  if (host_out && n) cudaMemcpyFromSymbol(host_out, g_trace2, (size_t)n * 3 * sizeof(unsigned long long));
  return (int)n;
}

}  // namespace tc2
}  // namespace hsg
