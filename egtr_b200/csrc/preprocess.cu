// Input staging on the device: what HuggingFace DetrFeatureExtractor does on the host for the reference
// (data/visual_genome.py:64-66, train_egtr.py:176-186) — PIL bilinear resize, /255, ImageNet normalisation, zero padding to
// the batch's largest size, pixel_mask — as two integer resampling passes that are bit-identical to Pillow's
// ImagingResampleHorizontal_8bpc / Vertical_8bpc (22-bit fixed-point taps, uint8 intermediate) and an fp32 epilogue with
// IEEE divisions.  The uint8 image is what crosses PCIe (0.9 MB for 480x640 instead of 21 MB of fp32 pixels + int64 mask).
#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// one thread = one output pixel (all C <= 4 channels): out[y, xo, c] = clip8((2^21 + sum_t src[y, xmin + t, c] * k[xo][t]) >> 22)
__global__ void __launch_bounds__(256)
resample_h_u8_kernel(const uint8_t* __restrict__ src, int H, int W, int C, int OW, const int* __restrict__ bounds,
                     const int* __restrict__ kk, int ksize, uint8_t* __restrict__ dst) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)H * OW) return;
  const int y = (int)(i / OW), xo = (int)(i - (long long)y * OW);
  const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
  const int* k = kk + (long long)xo * ksize;
  int acc[4] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  const uint8_t* p = src + ((long long)y * W + xmin) * C;
  for (int t = 0; t < n; ++t) {
    const int kv = __ldg(k + t);
    for (int c = 0; c < C; ++c) acc[c] += (int)p[t * C + c] * kv;
  }
  for (int c = 0; c < C; ++c) dst[i * C + c] = (uint8_t)clip8(acc[c] >> kPrecisionBits);
}

// vertical pass + (x / 255 - mean) / std in fp32 (IEEE divisions, as numpy computes it) into a plane of the padded NCHW batch
// tensor; also sets the pixel_mask of the valid region
__global__ void __launch_bounds__(256)
resample_v_norm_kernel(const uint8_t* __restrict__ src, int H, int W, int C, int OH, const int* __restrict__ bounds,
                       const int* __restrict__ kk, int ksize, float3 mean, float3 stdv, float* __restrict__ dst, long long plane_stride,
                       int row_stride, long long* __restrict__ mask, int mask_row_stride) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)OH * W) return;
  const int yo = (int)(i / W), x = (int)(i - (long long)yo * W);
  const int ymin = bounds[2 * yo], n = bounds[2 * yo + 1];
  const int* k = kk + (long long)yo * ksize;
  int acc[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int t = 0; t < n; ++t) {
    const int kv = __ldg(k + t);
    const uint8_t* p = src + ((long long)(ymin + t) * W + x) * C;
    for (int c = 0; c < 3; ++c) acc[c] += (int)p[c] * kv;
  }
  const float m[3] = {mean.x, mean.y, mean.z}, s[3] = {stdv.x, stdv.y, stdv.z};
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn((float)clip8(acc[c] >> kPrecisionBits), 255.0f);
    dst[c * plane_stride + (long long)yo * row_stride + x] = __fdiv_rn(__fsub_rn(v, m[c]), s[c]);
  }
  if (mask) mask[(long long)yo * mask_row_stride + x] = 1;
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_resample_h_u8(const uint8_t* src, int H, int W, int C, int OW, const int* bounds, const int* kk, int ksize,
                                  uint8_t* dst, egtr_stream_t s) {
  EGTR_CHECK(src && bounds && kk && dst && H > 0 && W > 0 && OW > 0 && C >= 1 && C <= 4 && ksize >= 1, EGTR_ERR_ARG,
             "egtr_resample_h_u8: bad arguments (H=%d W=%d C=%d OW=%d ksize=%d)", H, W, C, OW, ksize);
  const long long total = (long long)H * OW;
  launch_pdl(resample_h_u8_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)0, (cudaStream_t)s, src, H, W, C, OW, bounds, kk, ksize, dst);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_resample_v_normalize_f32(const uint8_t* src, int H, int W, int C, int OH, const int* bounds, const int* kk,
                                             int ksize, const float* mean3, const float* std3, float* dst, long long plane_stride,
                                             int row_stride, int64_t* mask, int mask_row_stride, egtr_stream_t s) {
  EGTR_CHECK(src && bounds && kk && dst && mean3 && std3 && H > 0 && W > 0 && OH > 0 && C >= 3 && C <= 4 && ksize >= 1 && row_stride >= W,
             EGTR_ERR_ARG, "egtr_resample_v_normalize_f32: bad arguments (H=%d W=%d C=%d OH=%d)", H, W, C, OH);
  const long long total = (long long)OH * W;
  launch_pdl(resample_v_norm_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)0, (cudaStream_t)s, src, H, W, C, OH, bounds, kk, ksize,
             make_float3(mean3[0], mean3[1], mean3[2]), make_float3(std3[0], std3[1], std3[2]), dst, plane_stride, row_stride,
             (long long*)mask, mask_row_stride);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
