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

// The library keeps a few per-process device objects (error flags, split-K scratch, SM count) that belong to the device that was
// current at first use: one process drives ONE GPU (the path's multi-GPU model).  A call issued with another
// device current would hand kernels pointers of the wrong GPU — refuse it loudly instead.
int bound_device_ok() {
  static std::atomic<int> bound{-1};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int expected = -1;
  if (bound.compare_exchange_strong(expected, dev)) return 1;
  return expected == dev;
}

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

// Balanced persistent grids: a grid capped at `slots` CTAs (or pairs) runs ceil(work / slots) rounds of tiles whatever its size
// between ceil(work / rounds) and slots, so the smallest such grid finishes at the same time and leaves the other SMs to the
// kernels of the other forwards in flight (87 pair tiles on 37 slots: 3 rounds either way, 29 pairs instead of 37).
// EGTR_GEMM_BALANCE=0 restores min(work, slots) (dev A/B).
static std::atomic<int> g_grid_balance{-1};
int balanced_grid(long long work, int slots) {
  int on = g_grid_balance.load(std::memory_order_relaxed);
  if (on < 0) {
    const char* e = getenv("EGTR_GEMM_BALANCE");
    on = e ? (atoi(e) != 0) : 1;
    g_grid_balance.store(on);
  }
  if (slots < 1) slots = 1;
  if (work <= slots) return (int)work;
  if (!on || grid_div() == 1) return slots;  // a lone forward has nobody to leave SMs to (and full grids keep their PDL attribute)
  const long long rounds = (work + slots - 1) / slots;
  return (int)((work + rounds - 1) / rounds);
}

static std::atomic<int> g_pdl_mode{-1};
int pdl_mode() {
  int v = g_pdl_mode.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("EGTR_B200_PDL");
    v = e ? atoi(e) : 2;
    g_pdl_mode.store(v);
  }
  return v;
}
// dev/diagnostic switches (tools/diag_race.py): bit 0 MSDA gathers bypass L1 (ld.global.cg), bit 1 the TMA-store epilogue waits
// for full completion of its bulk stores (not only for the reads of its staging tiles) before the CTA exits, bit 2 the GEMM
// kernels do not trigger their dependents early
static std::atomic<int> g_debug_flags{-1};
int debug_flags() {
  int v = g_debug_flags.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("EGTR_B200_DEBUG_FLAGS");
    v = e ? atoi(e) : 0;
    g_debug_flags.store(v);
  }
  return v;
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
extern "C" int egtr_set_pdl_mode(int mode) {
  EGTR_CHECK(mode >= 0 && mode <= 2, EGTR_ERR_ARG, "egtr_set_pdl_mode: %d outside 0..2", mode);
  egtr::g_pdl_mode.store(mode);
  return EGTR_OK;
}
extern "C" int egtr_set_debug_flags(int flags) {
  egtr::g_debug_flags.store(flags & 0xff);
  return EGTR_OK;
}
extern "C" int egtr_set_grid_balance(int on) {
  egtr::g_grid_balance.store(on != 0);
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
