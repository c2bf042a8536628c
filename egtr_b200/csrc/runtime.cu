// Library-wide state: per-thread error string, launch counter, device properties.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace egtr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// EGTR_B200_PDL: 0 = programmatic dependent launch off, 1 = every launch, 2 = launches of at least one CTA per SM
// Split-K scratch is per "slot": forwards that may run concurrently (two CUDA graphs on two streams) use different slots.
static thread_local int g_scratch_slot = 0;
int scratch_slot() { return g_scratch_slot; }

// Upper bound on split-K factors of the tensor-core GEMMs.  Default 64 (a lone forward is latency-bound); the engine sets 1 for
// forwards that run several in flight (serving): idle SMs are filled by other images and the partial-sum round trip is pure cost.
static std::atomic<int> g_splitk_max{-1};
int splitk_max() {
  int v = g_splitk_max.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("EGTR_GEMM_SPLITK_MAX");
    v = e ? atoi(e) : 64;
    if (v < 1) v = 1;
    g_splitk_max.store(v);
  }
  return v;
}

// Persistent GEMM grids are capped at num_sms / grid_div: with several forwards in flight, GEMMs of different images then run
// side by side on disjoint SMs (better wave quantisation, prologues and tails overlap) instead of time-slicing the whole GPU.
static std::atomic<int> g_grid_div{-1};
int grid_div() {
  int v = g_grid_div.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("EGTR_GEMM_GRID_DIV");
    v = e ? atoi(e) : 1;
    if (v < 1) v = 1;
    g_grid_div.store(v);
  }
  return v;
}

int pdl_mode() {
  static const int mode = [] { const char* e = getenv("EGTR_B200_PDL"); return e ? atoi(e) : 2; }();
  return mode;
}

}  // namespace egtr

extern "C" int egtr_set_scratch_slot(int slot) {
  EGTR_CHECK(slot >= 0 && slot < 32, EGTR_ERR_ARG, "egtr_set_scratch_slot: slot %d outside 0..31", slot);
  egtr::g_scratch_slot = slot;
  return EGTR_OK;
}
extern "C" int egtr_set_grid_div(int div) {
  EGTR_CHECK(div >= 1 && div <= 16, EGTR_ERR_ARG, "egtr_set_grid_div: %d outside 1..16", div);
  egtr::g_grid_div.store(div);
  return EGTR_OK;
}
extern "C" int egtr_set_splitk_max(int max_splits) {
  EGTR_CHECK(max_splits >= 1 && max_splits <= 64, EGTR_ERR_ARG, "egtr_set_splitk_max: %d outside 1..64", max_splits);
  egtr::g_splitk_max.store(max_splits);
  return EGTR_OK;
}
extern "C" const char* egtr_last_error(void) { return egtr::g_err; }
extern "C" int egtr_abi_version(void) { return 1; }
extern "C" long long egtr_launch_count(void) { return egtr::g_launches.load(); }
extern "C" void egtr_launch_count_reset(void) { egtr::g_launches.store(0); }
