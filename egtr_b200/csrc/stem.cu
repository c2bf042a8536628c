// ResNet stem: 7x7 / stride 2 / pad 3 convolution + folded FrozenBN + ReLU (timm resnet50 `conv1` / `bn1` / `act1`, reached through
// model/deformable_detr.py:772-787) as a TMA-fed tcgen05 kernel.
//
// Round 1/2 ran it on gemm_sbf16_kernel's software producer (eight warps gather 49 taps per output pixel from a zero-padded NHWC4
// fp32 image, split them to bf16 hi/lo and store the swizzled operand tile): 96 us alone, 6 % of the throughput step with the
// max-pool behind it.  Here the image is stored ONCE as two zero-bordered NHWC4 bf16 planes (hi, lo: x = hi + lo to 2^-17), and a
// filter ROW of the window of output pixel (oy, ox) — 7 pixels x 4 channels = 28 contiguous elements starting at padded pixel 2*ox —
// is a contiguous run of memory.  A 3-D tensor map whose "pixel" dimension advances by 2 pixels (16 bytes) while its inner box covers
// KROW elements (overlapping rows) turns the 128 windows of a 128-pixel output tile, for one filter row ky, into ONE TMA box per
// plane: no gather code, no conversion, no im2col buffer.  K = 7 filter rows x 32 elements (28 real + 4 that meet zero weights).
//   persistent CTAs, 192 threads: warps 0-3 epilogue (TMEM lane quadrants), warp 4 TMA, warp 5 MMA
//   weights (64 x 7 x KROW, hi/lo planes) resident in shared memory; operand ring of (hi, lo) boxes per filter row
//   per tile 7 x 2 k-steps x 3 products (lo*hi + hi*lo + hi*hi) = 42 MMAs of 128 x 64 x 16 into one of two TMEM accumulators
//   epilogue: bias + ReLU -> fp32 NHWC rows through a per-warp transposition tile (full 128-byte lines per store)
#include <cuda.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace egtr {
void count_launch();
int tmap_tiled(const void* ptr, int dtype, int rank, const unsigned long long* dims, const unsigned long long* strides, const unsigned* box,
               int swizzle64, CUtensorMap* out);  // gemm_p32.cu: the library's tensor-map cache
namespace {

// Image layouts (EGTR_STEM_LAYOUT):
//   1  two planes (hi, lo) of NHWC4 bf16, 64-byte operand rows (SWIZZLE_64B): two boxes per filter row, 6 MMAs (lo*hi, hi*lo, hi*hi)
//   2  ONE plane, hi and lo interleaved per pixel (h0 h1 h2 0 l0 l1 l2 0 = 16 bytes): the 8-pixel window is one 128-byte operand row,
//      ONE box per filter row (half the TMA requests of layout 1).
//      The split moves into two zero-padded weight sets over the same 64-element row: B13 = w_hi at the hi AND the lo positions
//      (A . B13 = hi*w_hi + lo*w_hi), B2 = w_lo at the hi positions (A . B2 = hi*w_lo): 8 MMAs per filter row, half of whose
//      multiplies meet zeros.  Measured 64.4 us against layout 1's 57.4 us (six 16 KB stages in flight instead of nine next to the
//      112 KB of resident weight sets): the kernel is bound by bytes in flight, not by the request rate.  Kept as an A/B variant.
#ifndef EGTR_STEM_LAYOUT
#define EGTR_STEM_LAYOUT 1
#endif
constexpr int LAYOUT = EGTR_STEM_LAYOUT;
constexpr int KROW = LAYOUT == 2 ? 64 : 32;      // elements of an operand row
constexpr int ROW_BYTES = KROW * 2;
constexpr int PIX_BYTES = LAYOUT == 2 ? 16 : 8;  // bytes per pixel of a plane
constexpr int NPLANES = LAYOUT == 2 ? 1 : 2;
constexpr int KH = 7, NOUT = 64, TILE = 128;
constexpr int A_BYTES = TILE * ROW_BYTES;        // one plane's box of one filter row
constexpr int W_BYTES = NOUT * ROW_BYTES;        // one weight tile (hi / lo plane, or set B13 / B2) of one filter row
constexpr int STAGES = LAYOUT == 2 ? 6 : 9;
constexpr int STAGE_BYTES = NPLANES * A_BYTES;
constexpr int W_RES_BYTES = KH * 2 * W_BYTES;
constexpr int STG_BYTES = 4 * 4096;
constexpr int SMEM_BYTES = W_RES_BYTES + STAGES * STAGE_BYTES + STG_BYTES + 256 + 1024;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 128;  // two accumulators of 64 columns

struct StemArgs {
  const float* bias;
  float* out;       // [B, OH, OW, 64] fp32
  int B, OH, OW, tiles_x;
  int* err;
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// K-major operand tile descriptor: 128-byte rows / SWIZZLE_128B (8-row groups 1024 B apart) or 64-byte rows / SWIZZLE_64B (512 B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}

__global__ void __launch_bounds__(THREADS, 1)
stem_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_w,
            const StemArgs a) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* wres = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // [KH][hi | lo][64 rows]
  uint8_t* ring = wres + W_RES_BYTES;
  uint8_t* stg_all = ring + STAGES * STAGE_BYTES;
  uint64_t* full_bar = (uint64_t*)(stg_all + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_holder = (uint32_t*)(w_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = a.OH * a.tiles_x, total = a.B * tiles_per_img;
  if (warp == 4 && lane == 0) {
    ptx::prefetch_tensormap(&map_hi);
    ptx::prefetch_tensormap(&map_lo);
    ptx::prefetch_tensormap(&map_w);
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) ptx::tmem_alloc<TMEM_COLS>(tmem_holder);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_holder;

  if (warp == 4) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_full, W_RES_BYTES);
      for (int ky = 0; ky < KH; ++ky) {  // weight planes [2][64][KH * KROW]: rows 0-63 hi, 64-127 lo
        tma_load_2d(ptx::smem_u32(wres + ky * 2 * W_BYTES), &map_w, w_full, ky * KROW, 0);
        tma_load_2d(ptx::smem_u32(wres + ky * 2 * W_BYTES + W_BYTES), &map_w, w_full, ky * KROW, NOUT);
      }
      int stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / tiles_per_img, r = t - b * tiles_per_img;
        const int oy = r / a.tiles_x, ox0 = (r - oy * a.tiles_x) * TILE;
        for (int ky = 0; ky < KH; ++ky) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, a.err, 401);
          const uint32_t st = ptx::smem_u32(ring + stage * STAGE_BYTES);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_4d(st, &map_hi, &full_bar[stage], 0, ox0, 2 * oy + ky, b);
          if (NPLANES == 2) tma_load_4d(st + A_BYTES, &map_lo, &full_bar[stage], 0, ox0, 2 * oy + ky, b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(TILE, NOUT);
    ptx::mbar_wait(w_full, 0, a.err, 402);
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      ptx::mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1, a.err, 403);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem + acc * NOUT;
      for (int ky = 0; ky < KH; ++ky) {
        ptx::mbar_wait(&full_bar[stage], phase, a.err, 404);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = ptx::smem_u32(ring + stage * STAGE_BYTES), a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = ptx::smem_u32(wres + ky * 2 * W_BYTES), b_lo = b_hi + W_BYTES;
          if (LAYOUT == 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // the interleaved row's 64 elements (8 pixels x (4 hi | 4 lo)): four 16-wide k-steps
              const uint64_t da = umma_desc(a_hi + ks * 32);
              ptx::umma_bf16(d_tmem, da, umma_desc(b_lo + ks * 32), idesc, (ky != 0) || (ks != 0));  // set B2: hi * w_lo (small term first)
              ptx::umma_bf16(d_tmem, da, umma_desc(b_hi + ks * 32), idesc, 1);                        // set B13: hi * w_hi + lo * w_hi
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {  // the filter row's 32 elements (28 real): two 16-wide k-steps
              const uint64_t dah = umma_desc(a_hi + ks * 32), dal = umma_desc(a_lo + ks * 32);
              const uint64_t dbh = umma_desc(b_hi + ks * 32), dbl = umma_desc(b_lo + ks * 32);
              ptx::umma_bf16(d_tmem, dal, dbh, idesc, (ky != 0) || (ks != 0));  // small terms first
              ptx::umma_bf16(d_tmem, dah, dbl, idesc, 1);
              ptx::umma_bf16(d_tmem, dah, dbh, idesc, 1);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (ky == KH - 1) ptx::umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const uint32_t stg = ptx::smem_u32(stg_all + warp * 4096);
    const uint32_t mine = stg + lane * 128;
    const int sw = lane & 7, c16 = lane & 7;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int b = t / tiles_per_img, r = t - b * tiles_per_img;
      const int oy = r / a.tiles_x, ox0 = (r - oy * a.tiles_x) * TILE + warp * 32;  // this warp's first output pixel
      const int acc = it & 1;
      ptx::mbar_wait(&acc_full[acc], (it >> 1) & 1, a.err, 405);
      ptx::tc_fence_after();
      const int nrows = max(0, min(32, a.OW - ox0));
      float* orow0 = a.out + (((long long)b * a.OH + oy) * a.OW + ox0) * NOUT;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t rr[32];
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + acc * NOUT + ch * 32, rr);
        ptx::tmem_ld_wait();
        if (ch == 1) {  // accumulator drained into registers: the MMA warp may start the tile after next
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
        }
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b4 = __ldg((const float4*)(a.bias + ch * 32) + q4);
          rr[4 * q4] = __float_as_uint(fmaxf(__uint_as_float(rr[4 * q4]) + b4.x, 0.f));
          rr[4 * q4 + 1] = __float_as_uint(fmaxf(__uint_as_float(rr[4 * q4 + 1]) + b4.y, 0.f));
          rr[4 * q4 + 2] = __float_as_uint(fmaxf(__uint_as_float(rr[4 * q4 + 2]) + b4.z, 0.f));
          rr[4 * q4 + 3] = __float_as_uint(fmaxf(__uint_as_float(rr[4 * q4 + 3]) + b4.w, 0.f));
        }
        // lane = pixel holds 32 channels: transpose through the warp's tile so that a store instruction covers 4 pixels x 128 bytes
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + ((c ^ sw) << 4)), "r"(rr[4 * c]), "r"(rr[4 * c + 1]), "r"(rr[4 * c + 2]),
                       "r"(rr[4 * c + 3]) : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + row * 128 + ((c16 ^ (row & 7)) << 4)) : "memory");
          if (row < nrows) *(uint4*)((uint8_t*)(orow0 + (long long)row * NOUT + ch * 32) + c16 * 16) = v;
        }
        __syncwarp();
      }
    }
  }
  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// NCHW fp32 image -> zero-bordered NHWC4 bf16, Hp = H + 6, Wp = W + 6 rounded up to even: two planes [2][B][Hp][Wp][4] (hi, then lo:
// layout 1) or one plane [B][Hp][Wp][8] with hi and lo interleaved per pixel (layout 2)
__global__ void pad_split_kernel(const float* __restrict__ img, int B, int H, int W, int Hp, int Wp, uint2* __restrict__ out) {
  pdl_entry();
  const long long n = (long long)B * Hp * Wp;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int xp = (int)(i % Wp);
  const long long t = i / Wp;
  const int yp = (int)(t % Hp), b = (int)(t / Hp);
  const int x = xp - 3, y = yp - 3;
  float v[3] = {0.f, 0.f, 0.f};
  if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
    const long long plane = (long long)H * W, o = (long long)b * 3 * plane + (long long)y * W + x;
    v[0] = __ldg(img + o); v[1] = __ldg(img + o + plane); v[2] = __ldg(img + o + 2 * plane);
  }
  __nv_bfloat16 h[3], l[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) split_bf16(v[c], h[c], l[c]);
  auto pk = [](__nv_bfloat16 a0, __nv_bfloat16 a1) { return (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(a1) << 16); };
  const uint2 hi = make_uint2(pk(h[0], h[1]), pk(h[2], __float2bfloat16_rn(0.f))), lo = make_uint2(pk(l[0], l[1]), pk(l[2], __float2bfloat16_rn(0.f)));
  if (LAYOUT == 2) {
    ((uint4*)out)[i] = make_uint4(hi.x, hi.y, lo.x, lo.y);
  } else {
    out[i] = hi;
    out[n + i] = lo;
  }
}

int* stem_error_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" long long egtr_stem_planes_bytes(int B, int H, int W) {
  const long long Hp = H + 6, Wp = (W + 6 + 1) & ~1;
  return 2 * B * Hp * Wp * 8 + 4096;  // + slack: the last windows of the last row read past it (into zero weights)
}
extern "C" int egtr_stem_krow(void) { return KROW; }
extern "C" int egtr_stem_layout(void) { return LAYOUT; }

extern "C" int egtr_stem_pad_split_bf16(const float* img, int B, int H, int W, void* planes, egtr_stream_t s) {
  EGTR_CHECK(img && planes && B > 0 && H > 0 && W > 0, EGTR_ERR_ARG, "egtr_stem_pad_split_bf16: bad arguments");
  EGTR_CHECK(((uintptr_t)planes & 15) == 0, EGTR_ERR_ARG, "egtr_stem_pad_split_bf16: planes must be 16-byte aligned");
  const int Hp = H + 6, Wp = (W + 6 + 1) & ~1;
  const long long total = (long long)B * Hp * Wp;
  launch_pdl(pad_split_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)0, (cudaStream_t)s, img, B, H, W, Hp, Wp, (uint2*)planes);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_stem_conv7x7s2_bf16x3(const void* planes, int B, int H, int W, const void* w_planes, const float* bias, float* out, egtr_stream_t s) {
  EGTR_ONE_DEVICE();
  EGTR_CHECK(planes && w_planes && bias && out && B > 0 && H > 0 && W > 0, EGTR_ERR_ARG, "egtr_stem_conv7x7s2_bf16x3: bad arguments");
  EGTR_CHECK(((uintptr_t)planes & 127) == 0 && ((uintptr_t)w_planes & 127) == 0 && ((uintptr_t)out & 15) == 0, EGTR_ERR_ARG,
             "egtr_stem_conv7x7s2_bf16x3: alignment");
  const int Hp = H + 6, Wp = (W + 6 + 1) & ~1;
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const int swz64 = ROW_BYTES == 128 ? 0 : 1;
  // {element in window, output pixel (stride 2 pixels: the windows overlap), padded row, image}
  const unsigned long long row_pitch = (unsigned long long)Wp * PIX_BYTES, img_pitch = row_pitch * Hp;
  const unsigned long long dims[4] = {(unsigned long long)KROW, (unsigned long long)OW, (unsigned long long)Hp, (unsigned long long)B};
  const unsigned long long strides[3] = {2ull * PIX_BYTES, row_pitch, img_pitch};
  const unsigned box[4] = {(unsigned)KROW, TILE, 1, 1};
  CUtensorMap m_hi, m_lo, m_w;
  int rc;
  if ((rc = tmap_tiled(planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, swz64, &m_hi)) != EGTR_OK) return rc;
  m_lo = m_hi;
  if (NPLANES == 2 && (rc = tmap_tiled((const uint8_t*)planes + img_pitch * B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, swz64, &m_lo)) != EGTR_OK)
    return rc;
  // weights [2][64][KH * KROW] bf16: rows 0-63 = hi plane (layout 2: set B13), rows 64-127 = lo plane (set B2)
  const unsigned long long wdims[2] = {(unsigned long long)KH * KROW, 2 * NOUT}, wstrides[1] = {(unsigned long long)KH * KROW * 2};
  const unsigned wbox[2] = {(unsigned)KROW, NOUT};
  if ((rc = tmap_tiled(w_planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wdims, wstrides, wbox, swz64, &m_w)) != EGTR_OK) return rc;
  StemArgs a = {};
  a.bias = bias; a.out = out; a.B = B; a.OH = OH; a.OW = OW; a.tiles_x = cdiv(OW, TILE);
  a.err = stem_error_flag();
  static bool attr = false;
  if (!attr) {
    EGTR_CUDA(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  const long long total = (long long)B * OH * a.tiles_x;
  const int grid = (int)(total < num_sms() / grid_div() ? total : num_sms() / grid_div());
  EGTR_CUDA(launch_pdl(stem_kernel, dim3(grid > 0 ? grid : 1), dim3(THREADS), (size_t)SMEM_BYTES, (cudaStream_t)s, m_hi, m_lo, m_w, a));
  count_launch();
  return EGTR_OK;
}
