// Shared device/host helpers for the EGTR B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egtr_b200.h"

namespace egtr {

// ----------------------------------------------------------------------------- host errors
void set_last_error(const char* fmt, ...);

#define EGTR_CHECK(cond, code, ...)                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      ::egtr::set_last_error(__VA_ARGS__);                           \
      return (code);                                                 \
    }                                                                \
  } while (0)

#define EGTR_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::egtr::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                         \
      return EGTR_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int num_sms();
int bound_device_ok();  // runtime.cu: 1 when the current device is the one this process first used the library on
#define EGTR_ONE_DEVICE() \
  EGTR_CHECK(::egtr::bound_device_ok(), EGTR_ERR_UNSUPPORTED, "libegtr_b200 drives one GPU per process: the current CUDA device is not the one first used")
int splitk_max();  // runtime.cu: egtr_set_splitk_max / EGTR_GEMM_SPLITK_MAX
int grid_div();    // runtime.cu: egtr_set_grid_div / EGTR_GEMM_GRID_DIV
int balanced_grid(long long work, int slots);  // runtime.cu: smallest grid <= slots that needs no more rounds of tiles
int pdl_mode();  // 0 off, 1 every launch, 2 only grids of at least one CTA per SM
int debug_flags();  // runtime.cu: diagnostic switches (egtr_set_debug_flags)

// Every kernel of the library is launched with programmatic dependent launch (PDL): the next kernel's CTAs may become
// resident (and run their prologue) while the previous grid drains, and block in `pdl_entry()` / `pdl_wait()` until that grid
// has completed and its writes are visible.  Rule: a kernel executes griddepcontrol.wait before its first global access.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  const int mode = pdl_mode();
  if (mode == 1 || (mode == 2 && (long long)grid.x * grid.y * grid.z >= num_sms())) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_cluster_pdl(kernel, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}
#endif

// ----------------------------------------------------------------------------- A-operand source
// Where the rows of a GEMM's left operand come from.  One description serves the tcgen05 kernel's
// software producer and the SIMT kernel: rows are 64-float (256 B) contiguous runs in global memory.
//   mode 0  plain   : row m = a + m*lda                 (+ a2 + m*lda if a2 != nullptr, e.g. x + pos)
//   mode 1  conv    : implicit im2col over an NHWC tensor; k = (ky*KW + kx)*C + c, zero padding.
using ASrc = egtr_asrc_t;

struct RowInfo {  // per tile row, decoded once per tile
  long long base;  // plain: m*lda ; conv: b*H*W (pixel index of the image origin)
  int iy0, ix0;    // conv: top-left input coordinate of the receptive field (may be negative)
  int valid;       // 0 -> row beyond M: reads as zeros
};

__device__ __forceinline__ RowInfo decode_row(const ASrc& s, long long m, long long M) {
  RowInfo r;
  r.valid = m < M;
  if (s.mode == 0) {
    r.base = m * (long long)s.lda;
    r.iy0 = r.ix0 = 0;
  } else if (s.mode == 2) {
    long long ohw = (long long)s.OH * s.OW;
    long long b = m / ohw;
    int rem = (int)(m - b * ohw);
    int oy = rem / s.OW, ox = rem - oy * s.OW;
    r.base = b;  // image index; elements are gathered one by one from the NCHW planes
    r.iy0 = oy * s.stride - s.pad;
    r.ix0 = ox * s.stride - s.pad;
  } else {  // modes 1 and 3: NHWC pixel addressing (mode 3: zero-padded NHWC4 image, pad == 0 here)
    long long ohw = (long long)s.OH * s.OW;
    long long b = m / ohw;
    int rem = (int)(m - b * ohw);
    int oy = rem / s.OW, ox = rem - oy * s.OW;
    r.base = b * (long long)s.H * s.W;
    r.iy0 = oy * s.stride - s.pad;
    r.ix0 = ox * s.stride - s.pad;
  }
  return r;
}

// Offset (in floats) of the 16B-aligned run starting at column k0 of this row, or -1 for zeros.
// For conv mode k0..k0+63 must not straddle a filter tap (C % 64 == 0 or the run is < C).
__device__ __forceinline__ long long row_offset(const ASrc& s, const RowInfo& r, int k0) {
  if (!r.valid) return -1;
  if (s.mode == 0) return r.base + k0;
  int tap = k0 / s.C;
  int c0 = k0 - tap * s.C;
  int ky = tap / s.KW, kx = tap - ky * s.KW;
  int iy = r.iy0 + ky, ix = r.ix0 + kx;
  if ((unsigned)iy >= (unsigned)s.H || (unsigned)ix >= (unsigned)s.W) return -1;
  return (r.base + (long long)iy * s.W + ix) * s.C + c0;
}

// mode 3 (zero-padded NHWC4 image, the stem): float4 number `q` of the row = filter tap q (4 channels, the 4th
// zero); offset in floats, or -1 for the K-padding taps.  The image's own zero border replaces bounds checks.
__device__ __forceinline__ long long tap_offset_nhwc4(const ASrc& s, const RowInfo& r, int q) {
  if (!r.valid || q >= s.KH * s.KW) return -1;
  const int ky = q / s.KW, kx = q - ky * s.KW;
  return (r.base + (long long)(r.iy0 + ky) * s.W + (r.ix0 + kx)) * 4;
}

// mode 2 (NCHW stem gather): four consecutive k of one row, k = (ky*KW + kx)*C + c, zero beyond KH*KW*C.
__device__ __forceinline__ float4 gather4_nchw(const ASrc& s, long long img, int iy0, int ix0, int k) {
  float v[4];
  const int kmax = s.KH * s.KW * s.C;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = k + j;
    float x = 0.f;
    if (kk < kmax) {
      const int tap = kk / s.C, c = kk - tap * s.C;
      const int ky = tap / s.KW, kx = tap - ky * s.KW;
      const int iy = iy0 + ky, ix = ix0 + kx;
      if ((unsigned)iy < (unsigned)s.H && (unsigned)ix < (unsigned)s.W)
        x = __ldg(s.a + ((img * s.C + c) * s.H + iy) * (long long)s.W + ix);
    }
    v[j] = x;
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

// ----------------------------------------------------------------------------- epilogue
// out[orow(m)*ldo + n] = act(acc + bias[n] + res[orow(m)*ldr + n])
//   orow(m) = (m / rows_per_b) * bstride + off + m % rows_per_b   (level slices of [B,S,C] buffers)
using Epilogue = egtr_epilogue_t;

// Relation pair tiles: 128 consecutive rows = 8 subjects x 16 objects of one image.  Inside a tile the rows are
// ordered so that the 16 rows one producer thread of the GEMM owns (r = 32*p + 2*i + rsub, i = 0..15) form a
// 4 subjects x 4 objects block: every U/V vector it loads is then used four times.
//   r = 32*p + 2*(4*a + c) + rsub   ->   subject = 4*(p & 1) + a,   object = 8*(p >> 1) + 4*rsub + c
__device__ __forceinline__ void pair_local(int r, int& s_loc, int& o_loc) {
  const int p = r >> 5, i = (r & 31) >> 1, rsub = r & 1;
  s_loc = 4 * (p & 1) + (i >> 2);
  o_loc = 8 * (p >> 1) + 4 * rsub + (i & 3);
}
__device__ __forceinline__ bool pair_decode(long long m, int n_q, int& b, int& i, int& j) {
  const int ti_n = (n_q + 7) >> 3, tj_n = (n_q + 15) >> 4;
  const long long per_img = (long long)ti_n * tj_n * 128;
  b = (int)(m / per_img);
  const int rem = (int)(m - b * per_img);
  const int tile = rem >> 7;
  const int ti = tile / tj_n, tj = tile - ti * tj_n;
  int s_loc, o_loc;
  pair_local(rem & 127, s_loc, o_loc);
  i = ti * 8 + s_loc;
  j = tj * 16 + o_loc;
  return i < n_q && j < n_q;
}

// Output row of GEMM row m, or -1 when the row has no output (padding rows of pair tiles).
__device__ __forceinline__ long long out_row(const Epilogue& e, long long m) {
  if (e.pair_n > 0) {
    int b, i, j;
    if (!pair_decode(m, e.pair_n, b, i, j)) return -1;
    return ((long long)b * e.pair_n + i) * e.pair_n + j;
  }
  if (e.rows_per_b <= 0) return m;
  long long b = m / e.rows_per_b;
  return b * e.bstride + e.off + (m - b * e.rows_per_b);
}

// ----------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// fp32 -> (hi, lo) bf16 pair with hi + lo == x to ~2^-17 relative.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace egtr
