// C-ABI bookkeeping: version, error strings, device check, per-kernel profiling.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "hsg_common.cuh"

namespace hsg {

static const char* kSlotNames[SLOT_COUNT] = {
    "build_count", "build_scan", "build_fill", "attn_prep_fwd", "attn_prep_bwd", "gemm_nt",
    "gemm_nn", "gemm_tn", "gemm_tn_reduce", "edge_fwd", "edge_bwd_prep", "edge_bwd",
    "edge_bwd_dq", "layernorm_fwd", "layernorm_bwd", "layernorm_bwd_reduce", "head", "adam", "dropout", "s2s", "encoder", "lstm", "ffn_rows"};

struct EventPair {
  cudaEvent_t a, b;
  int slot;
};

static std::mutex g_mu;
static bool g_prof = false;
static std::vector<EventPair> g_pending;
static std::vector<cudaEvent_t> g_free;
static int g_count[SLOT_COUNT];
static float g_ms[SLOT_COUNT];
static std::atomic<long long> g_launches{0};
static int g_sms = 0;

static cudaEvent_t get_event() {
  if (!g_free.empty()) {
    cudaEvent_t e = g_free.back();
    g_free.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void prof_begin(int slot, cudaStream_t s) {
  if (!g_prof) return;
  std::lock_guard<std::mutex> lk(g_mu);
  EventPair p;
  p.a = get_event();
  p.b = get_event();
  p.slot = slot;
  cudaEventRecord(p.a, s);
  g_pending.push_back(p);
}

void prof_end(int slot, cudaStream_t s) {
  if (!g_prof) return;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = (int)g_pending.size() - 1; i >= 0; --i) {
    if (g_pending[i].slot == slot) {
      cudaEventRecord(g_pending[i].b, s);
      break;
    }
  }
}

static void drain() {
  for (auto& p : g_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_count[p.slot] += 1;
      g_ms[p.slot] += ms;
    }
    g_free.push_back(p.a);
    g_free.push_back(p.b);
  }
  g_pending.clear();
}

static std::atomic<int> g_pdl{-1};

bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HSG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
    g_pdl.store(v);
  }
  return v != 0;
}

int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_version(void) { return HSG_ABI_VERSION; }

// sizeof of the i-th argument structure of the header, in declaration order (binding self-check: a foreign-function
// binding compares its own layout with these before the first call); 0 past the end
size_t hsg_abi_sizeof(int i) {
  switch (i) {
    case 0: return sizeof(hsg_token_batch);
    case 1: return sizeof(hsg_graph_offsets);
    case 2: return sizeof(hsg_csc);
    case 3: return sizeof(hsg_graph_out);
    case 4: return sizeof(hsg_wswgat_fwd_args);
    case 5: return sizeof(hsg_wswgat_bwd_args);
    case 6: return sizeof(hsg_layer_params);
    case 7: return sizeof(hsg_layer_grads);
    case 8: return sizeof(hsg_loop_args);
    case 9: return sizeof(hsg_loop_plan);
    case 10: return sizeof(hsg_loop_bwd_args);
    case 11: return sizeof(hsg_head_args);
    case 12: return sizeof(hsg_s2s_graph);
    case 13: return sizeof(hsg_doc_map);
    default: return 0;
  }
}

const char* hsg_strerror(int status) {
  switch (status) {
    case HSG_OK: return "ok";
    case HSG_ERR_ARG: return "invalid argument (null pointer or negative size)";
    case HSG_ERR_SHAPE: return "unsupported shape ((heads, head_dim) not instantiated or size out of range)";
    case HSG_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case HSG_ERR_WORKSPACE: return "workspace too small";
    case HSG_ERR_CUDA: return "CUDA launch failed";
    case HSG_ERR_ARCH: return "device is not sm_100 (B200)";
    case HSG_ERR_CAPACITY: return "graph builder capacity exceeded";
    default: return "unknown hsg status";
  }
}

int hsg_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return HSG_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return HSG_ERR_CUDA;
  return major == 10 ? HSG_OK : HSG_ERR_ARCH;
}

int hsg_num_sms(void) { return num_sms(); }

int hsg_set_pdl(int on) {
  g_pdl.store(on ? 1 : 0);
  return HSG_OK;
}

int hsg_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!on) drain();
  g_prof = on != 0;
  return HSG_OK;
}

int hsg_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  drain();
  for (int i = 0; i < SLOT_COUNT; ++i) {
    g_count[i] = 0;
    g_ms[i] = 0.f;
  }
  return HSG_OK;
}

int hsg_profile_num_slots(void) { return SLOT_COUNT; }

const char* hsg_profile_slot_name(int slot) { return (slot >= 0 && slot < SLOT_COUNT) ? kSlotNames[slot] : ""; }

int hsg_profile_read(int slot, int* host_count, float* host_ms) {
  if (slot < 0 || slot >= SLOT_COUNT || !host_count || !host_ms) return HSG_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  drain();
  *host_count = g_count[slot];
  *host_ms = g_ms[slot];
  return HSG_OK;
}

long long hsg_launch_count(void) { return g_launches.load(); }

int hsg_memset(void* ptr, int value, size_t bytes, void* stream) {
  if (!ptr && bytes) return HSG_ERR_ARG;
  if (bytes == 0) return HSG_OK;
  return cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream) == cudaSuccess ? HSG_OK : HSG_ERR_CUDA;
}

}  // extern "C"
