// Shared helpers for the sm_100a WSWGAT kernels: launch bookkeeping, per-kernel
// CUDA-event profiling, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsg_b200.h"

namespace hsg {

enum Slot {
  SLOT_BUILD_COUNT = 0,
  SLOT_BUILD_SCAN,
  SLOT_BUILD_FILL,
  SLOT_PREP_FWD,
  SLOT_PREP_BWD,
  SLOT_GEMM_NT,
  SLOT_GEMM_NN,
  SLOT_GEMM_TN,
  SLOT_GEMM_TN_REDUCE,
  SLOT_EDGE_FWD,
  SLOT_EDGE_BWD_PREP,
  SLOT_EDGE_BWD,
  SLOT_EDGE_BWD_DQ,
  SLOT_LN_FWD,
  SLOT_LN_BWD,
  SLOT_LN_BWD_REDUCE,
  SLOT_HEAD,
  SLOT_ADAM,
  SLOT_DROPOUT,
  SLOT_S2S,
  SLOT_ENCODER,
  SLOT_LSTM,
  SLOT_FFN_ROWS,
  SLOT_COUNT
};

// implemented in hsg_abi.cu
void prof_begin(int slot, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
int num_sms();

struct LaunchScope {
  int slot;
  cudaStream_t s;
  LaunchScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_) { prof_begin(slot, s); }
  ~LaunchScope() { prof_end(slot, s); }
};

// Programmatic dependent launch (PDL): every kernel of the library is launched with the programmatic-stream-
// serialization attribute and starts with pdl_prologue(): it lets ITS dependents be scheduled at once
// (griddepcontrol.launch_dependents) and then waits until everything before it in the stream has completed and
// flushed (griddepcontrol.wait) before touching global memory.  The ~45 short kernels of a step are a dependent
// chain; this overlaps each kernel's launch latency and prologue (barrier init, TMEM allocation, tensor-map
// prefetch, index loads of parameters) with its predecessor's execution.  Without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}

void count_launch();   // hsg_abi.cu: one tick per kernel launch of the library (hsg_launch_count)
bool pdl_enabled();   // hsg_abi.cu (HSG_PDL=0 in the environment or hsg_set_pdl(0) turns it off)

template <typename... P, typename... A>
inline void launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? HSG_OK : HSG_ERR_CUDA; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : HSG_LEAKY_SLOPE * x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hsg
