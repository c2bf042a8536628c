// Row-wise and image-geometry kernels of the EGTR path (everything that is not a GEMM, the
// deformable gather or the relation head).  All fp32, all HBM-bound: coalesced float4 traffic,
// one pass over the data wherever the reduction structure allows it.
#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

// ------------------------------------------------------------------ LayerNorm(x + res), C == 256
__global__ void __launch_bounds__(256)
add_layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                        const float* __restrict__ beta, int rows, float* __restrict__ out) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xp = (const float4*)(x + (long long)row * 256);
  float4 a = xp[lane], b = xp[lane + 32];
  if (res) {
    const float4* rp = (const float4*)(res + (long long)row * 256);
    const float4 c = rp[lane], d = rp[lane + 32];
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
  const float mean = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.f / 256.f);
  a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
  b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
  const float var = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w) * (1.f / 256.f);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
  const float4 g0 = ((const float4*)gamma)[lane], g1 = ((const float4*)gamma)[lane + 32];
  const float4 b0 = ((const float4*)beta)[lane], b1 = ((const float4*)beta)[lane + 32];
  float4 o0, o1;
  o0.x = a.x * rstd * g0.x + b0.x; o0.y = a.y * rstd * g0.y + b0.y; o0.z = a.z * rstd * g0.z + b0.z; o0.w = a.w * rstd * g0.w + b0.w;
  o1.x = b.x * rstd * g1.x + b1.x; o1.y = b.y * rstd * g1.y + b1.y; o1.z = b.z * rstd * g1.z + b1.z; o1.w = b.w * rstd * g1.w + b1.w;
  float4* op = (float4*)(out + (long long)row * 256);
  op[lane] = o0;
  op[lane + 32] = o1;
}

// ------------------------------------------------------------------ P32 rows (include/egtr_b200.h)
// four consecutive channels c..c+3 (c % 4 == 0) of a row: hi at byte (c/32)*128 + (c%32)*2, lo 64 bytes further
__device__ __forceinline__ void p32_store4(uint8_t* row, int c, const float4& v) {
  __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
  uint2 ph, pl;
  ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
  uint8_t* g = row + (c >> 5) * 128 + (c & 31) * 2;
  *(uint2*)g = ph;
  *(uint2*)(g + 64) = pl;
}
__device__ __forceinline__ float4 p32_load4(const uint8_t* row, int c) {
  const uint8_t* g = row + (c >> 5) * 128 + (c & 31) * 2;
  const uint2 ph = *(const uint2*)g, pl = *(const uint2*)(g + 64);
  float4 v;
  v.x = __uint_as_float(ph.x << 16) + __uint_as_float(pl.x << 16);
  v.y = __uint_as_float(ph.x & 0xffff0000u) + __uint_as_float(pl.x & 0xffff0000u);
  v.z = __uint_as_float(ph.y << 16) + __uint_as_float(pl.y << 16);
  v.w = __uint_as_float(ph.y & 0xffff0000u) + __uint_as_float(pl.y & 0xffff0000u);
  return v;
}

__global__ void __launch_bounds__(256)
rows_to_p32_kernel(const float* __restrict__ x, const float* __restrict__ addend, long long rows, int C4, int ldx, uint8_t* __restrict__ out) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long row = i / C4;
  if (row >= rows) return;
  const int c = (int)(i - row * C4) * 4;
  float4 v = *(const float4*)(x + row * ldx + c);
  if (addend) {
    const float4 a = *(const float4*)(addend + row * ldx + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  p32_store4(out + row * (long long)C4 * 16, c, v);
}

__global__ void __launch_bounds__(256)
p32_to_rows_kernel(const uint8_t* __restrict__ p32, long long rows, int C4, float* __restrict__ out, int ldo) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long row = i / C4;
  if (row >= rows) return;
  const int c = (int)(i - row * C4) * 4;
  *(float4*)(out + row * ldo + c) = p32_load4(p32 + row * (long long)C4 * 16, c);
}

// LayerNorm(x + res) with P32 / fp32 outputs and the optional "+ addend" second P32 output (next layer's x + pos)
__global__ void __launch_bounds__(256)
add_layernorm256_p32_kernel(const float* __restrict__ x, const uint8_t* __restrict__ res, int res_fmt, const float* __restrict__ gamma,
                            const float* __restrict__ beta, int rows, uint8_t* __restrict__ out_p32, float* __restrict__ out_f32,
                            const float* __restrict__ addend, uint8_t* __restrict__ out_plus) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4* xp = (const float4*)(x + (long long)row * 256);
  float4 a = xp[lane], b = xp[lane + 32];
  if (res) {
    float4 c, d;
    if (res_fmt) {
      c = p32_load4(res + (long long)row * 1024, c0);
      d = p32_load4(res + (long long)row * 1024, c1);
    } else {
      const float4* rp = (const float4*)(res + (long long)row * 1024);
      c = rp[lane]; d = rp[lane + 32];
    }
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
  const float mean = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.f / 256.f);
  a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
  b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
  const float var = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w) * (1.f / 256.f);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
  const float4 g0 = ((const float4*)gamma)[lane], g1 = ((const float4*)gamma)[lane + 32];
  const float4 b0 = ((const float4*)beta)[lane], b1 = ((const float4*)beta)[lane + 32];
  float4 o0, o1;
  o0.x = a.x * rstd * g0.x + b0.x; o0.y = a.y * rstd * g0.y + b0.y; o0.z = a.z * rstd * g0.z + b0.z; o0.w = a.w * rstd * g0.w + b0.w;
  o1.x = b.x * rstd * g1.x + b1.x; o1.y = b.y * rstd * g1.y + b1.y; o1.z = b.z * rstd * g1.z + b1.z; o1.w = b.w * rstd * g1.w + b1.w;
  if (out_f32) {
    float4* op = (float4*)(out_f32 + (long long)row * 256);
    op[lane] = o0;
    op[lane + 32] = o1;
  }
  if (out_p32) {
    p32_store4(out_p32 + (long long)row * 1024, c0, o0);
    p32_store4(out_p32 + (long long)row * 1024, c1, o1);
  }
  if (out_plus) {
    const float4* ap = (const float4*)(addend + (long long)row * 256);
    const float4 p0 = ap[lane], p1 = ap[lane + 32];
    o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
    o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
    p32_store4(out_plus + (long long)row * 1024, c0, o0);
    p32_store4(out_plus + (long long)row * 1024, c1, o1);
  }
}

// LayerNorm(sum_s partial[s] + bias + res): consumes the raw split-K sums of egtr_gemm_f32_splitk; optional second (strided) copy
__global__ void __launch_bounds__(256)
sum_layernorm256_kernel(const float* __restrict__ partial, int splits, long long split_stride, const float* __restrict__ bias,
                        const float* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ beta, int rows,
                        float* __restrict__ out, float* __restrict__ out2, int rows_per_b2, long long bstride2) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 a = ((const float4*)bias)[lane], b = ((const float4*)bias)[lane + 32];
  for (int s = 0; s < splits; ++s) {
    const float4* xp = (const float4*)(partial + s * split_stride + (long long)row * 256);
    const float4 c = xp[lane], d = xp[lane + 32];
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
  if (res) {
    const float4* rp = (const float4*)(res + (long long)row * 256);
    const float4 c = rp[lane], d = rp[lane + 32];
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
  const float mean = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.f / 256.f);
  a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
  b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
  const float var = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w) * (1.f / 256.f);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
  const float4 g0 = ((const float4*)gamma)[lane], g1 = ((const float4*)gamma)[lane + 32];
  const float4 b0 = ((const float4*)beta)[lane], b1 = ((const float4*)beta)[lane + 32];
  float4 o0, o1;
  o0.x = a.x * rstd * g0.x + b0.x; o0.y = a.y * rstd * g0.y + b0.y; o0.z = a.z * rstd * g0.z + b0.z; o0.w = a.w * rstd * g0.w + b0.w;
  o1.x = b.x * rstd * g1.x + b1.x; o1.y = b.y * rstd * g1.y + b1.y; o1.z = b.z * rstd * g1.z + b1.z; o1.w = b.w * rstd * g1.w + b1.w;
  float4* op = (float4*)(out + (long long)row * 256);
  op[lane] = o0;
  op[lane + 32] = o1;
  if (out2) {
    const int bb = row / rows_per_b2;
    float4* op2 = (float4*)(out2 + bb * bstride2 + (long long)(row - bb * rows_per_b2) * 256);
    op2[lane] = o0;
    op2[lane + 32] = o1;
  }
}

// ------------------------------------------------------------------ zero masked rows
__global__ void mask_rows_kernel(float* __restrict__ x, int ld, int C4, const uint8_t* __restrict__ keep, long long rows) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long row = i / C4;
  if (row >= rows) return;
  if (!keep[row]) ((float4*)(x + row * ld))[i - row * C4] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ NCHW (C=3) -> zero-padded NHWC4
// one float4 per padded pixel; the border and the 4th channel are zero so the stem's gather needs no bounds checks
__global__ void pad_nhwc4_kernel(const float* __restrict__ img, int B, int H, int W, int pad, float4* __restrict__ out) {
  pdl_entry();
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * Hp * Wp) return;
  const int xp = (int)(i % Wp);
  const long long t = i / Wp;
  const int yp = (int)(t % Hp), b = (int)(t / Hp);
  const int x = xp - pad, y = yp - pad;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
    const long long plane = (long long)H * W, o = (long long)b * 3 * plane + (long long)y * W + x;
    v.x = __ldg(img + o); v.y = __ldg(img + o + plane); v.z = __ldg(img + o + 2 * plane);
  }
  out[i] = v;
}

// ------------------------------------------------------------------ 3x3/2 pad 1 max-pool, NHWC
__global__ void maxpool_kernel(const float* __restrict__ x, int B, int H, int W, int C4, int OH, int OW, float* __restrict__ out, int out_fmt) {
  pdl_entry();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C4;
  if (i >= total) return;
  const int c = (int)(i % C4);
  long long t = i / C4;
  const int ox = (int)(t % OW); t /= OW;
  const int oy = (int)(t % OH);
  const int b = (int)(t / OH);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float4 v = __ldg((const float4*)x + (((long long)b * H + iy) * W + ix) * C4 + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  if (out_fmt == 0) ((float4*)out)[i] = m;
  else p32_store4((uint8_t*)out + (i / C4) * (long long)C4 * 16, c * 4, m);
}

// ------------------------------------------------------------------ GroupNorm (C == 256, 32 groups of 8)
// pass 1: per (b, row chunk) partial sums per group in double; pass 2: normalise in place.
constexpr int GN_ROWS = 64;  // rows per CTA
__global__ void __launch_bounds__(256)
groupnorm_partial_kernel(const float* __restrict__ x, int rows_per_b, int bstride, int off, double* __restrict__ part) {
  pdl_entry();
  // thread = (row lane r8 in 0..7, float4 column c in 0..63 -> group c/2)
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int c = threadIdx.x & 63, r8 = threadIdx.x >> 6;  // 4 row lanes x 64 float4 columns
  const int row0 = chunk * GN_ROWS;
  float s = 0.f, ss = 0.f;
  for (int r = row0 + r8; r < min(row0 + GN_ROWS, rows_per_b); r += 4) {
    const float4 v = ((const float4*)(x + ((long long)b * bstride + off + r) * 256))[c];
    s += v.x + v.y + v.z + v.w;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ double sh[256][2];
  sh[threadIdx.x][0] = (double)s;
  sh[threadIdx.x][1] = (double)ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    double a = 0.0, q = 0.0;
    for (int rr = 0; rr < 4; ++rr)
      for (int cc = 0; cc < 2; ++cc) {
        a += sh[rr * 64 + g * 2 + cc][0];
        q += sh[rr * 64 + g * 2 + cc][1];
      }
    double* p = part + (((long long)b * gridDim.x + chunk) * 32 + g) * 2;
    p[0] = a;
    p[1] = q;
  }
}
// one warp per (image, group): fold the chunk partials into mean / rstd (float2 at the head of the scratch area)
__global__ void __launch_bounds__(32)
groupnorm_finalize_kernel(const double* __restrict__ part, int nchunks, int rows_per_b, float2* __restrict__ stats) {
  pdl_entry();
  const int g = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  double a = 0.0, q = 0.0;
  for (int k = lane; k < nchunks; k += 32) {
    const double* p = part + (((long long)b * nchunks + k) * 32 + g) * 2;
    a += p[0];
    q += p[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double n = (double)rows_per_b * 8.0;
    const double mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[b * 32 + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
  }
}
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(float* __restrict__ x, int rows_per_b, int bstride, int off, const float2* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, uint8_t* __restrict__ out_p32,
                       const float* __restrict__ addend, uint8_t* __restrict__ out_plus_p32) {
  pdl_entry();
  const int b = blockIdx.y, chunk = blockIdx.x;
  __shared__ float mean_s[32], rstd_s[32];
  if (threadIdx.x < 32) {
    const float2 st = stats[b * 32 + threadIdx.x];
    mean_s[threadIdx.x] = st.x;
    rstd_s[threadIdx.x] = st.y;
  }
  __syncthreads();
  const int c = threadIdx.x & 63, r8 = threadIdx.x >> 6;
  const float mean = mean_s[c >> 1], rstd = rstd_s[c >> 1];
  const float4 g = ((const float4*)gamma)[c], be = ((const float4*)beta)[c];
  const int row0 = chunk * GN_ROWS;
  for (int r = row0 + r8; r < min(row0 + GN_ROWS, rows_per_b); r += 4) {
    float4* p = (float4*)(x + ((long long)b * bstride + off + r) * 256) + c;
    float4 v = *p;
    v.x = (v.x - mean) * rstd * g.x + be.x; v.y = (v.y - mean) * rstd * g.y + be.y;
    v.z = (v.z - mean) * rstd * g.z + be.z; v.w = (v.w - mean) * rstd * g.w + be.w;
    *p = v;
    const long long row = (long long)b * bstride + off + r;
    if (out_p32) p32_store4(out_p32 + row * 1024, c * 4, v);  // the encoder's operands, written here instead of by a conversion pass
    if (out_plus_p32) {
      const float4 a = *((const float4*)(addend + row * 256) + c);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      p32_store4(out_plus_p32 + row * 1024, c * 4, v);
    }
  }
}

// ------------------------------------------------------------------ masks, cumulative sums, valid ratios
struct GeoLevels {
  int L;
  int h[8], w[8], start[8];
};
// nearest-neighbour level masks: one thread per token (legacy 'nearest': src = floor(dst * in/out), fp32 scale)
__global__ void __launch_bounds__(256)
level_mask_gather_kernel(const int64_t* __restrict__ pixel_mask, int H, int W, GeoLevels lv, int S, uint8_t* __restrict__ mask_flat) {
  pdl_entry();
  const int tok = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (tok >= S) return;
  int l = 0;
  while (l + 1 < lv.L && tok >= lv.start[l + 1]) ++l;
  const int h = lv.h[l], w = lv.w[l];
  const int i = tok - lv.start[l];
  const int y = i / w, x = i - y * w;
  const float sy = (float)H / (float)h, sx = (float)W / (float)w;
  const int yy = min((int)floorf((float)y * sy), H - 1), xx = min((int)floorf((float)x * sx), W - 1);
  mask_flat[(long long)b * S + tok] = pixel_mask[((long long)b * H + yy) * W + xx] != 0;
}
// one CTA per (level, image): column/row inclusive scans of the level mask (staged in smem) and valid ratios
__global__ void __launch_bounds__(256)
level_scans_kernel(GeoLevels lv, int S, const uint8_t* __restrict__ mask_flat, float* __restrict__ ycum, float* __restrict__ xcum,
                   float* __restrict__ valid_ratios) {
  pdl_entry();
  const int l = blockIdx.x, b = blockIdx.y;
  const int h = lv.h[l], w = lv.w[l];
  const uint8_t* mk = mask_flat + (long long)b * S + lv.start[l];
  float* yc = ycum + (long long)b * S + lv.start[l];
  float* xc = xcum + (long long)b * S + lv.start[l];
  const uint8_t* msrc = mk;
  // the loads of a scan are independent of the running sum: unrolling lets 8 of them be in flight per thread
  for (int x = threadIdx.x; x < w; x += blockDim.x) {  // y_embed = cumsum over rows (deformable_detr.py:853)
    float c = 0.f;
#pragma unroll 8
    for (int y = 0; y < h; ++y) { c += (float)__ldg(msrc + y * w + x); yc[y * w + x] = c; }
  }
  for (int y = threadIdx.x; y < h; y += blockDim.x) {  // x_embed = cumsum over columns (854)
    float c = 0.f;
#pragma unroll 8
    for (int x = 0; x < w; ++x) { c += (float)__ldg(msrc + y * w + x); xc[y * w + x] = c; }
  }
  if (threadIdx.x == 0) {  // get_valid_ratio (2064-2073): first column / first row
    int vh = 0, vw = 0;
    for (int y = 0; y < h; ++y) vh += msrc[y * w];
    for (int x = 0; x < w; ++x) vw += msrc[x];
    valid_ratios[((long long)b * lv.L + l) * 2 + 0] = (float)vw / (float)w;
    valid_ratios[((long long)b * lv.L + l) * 2 + 1] = (float)vh / (float)h;
  }
}
// sine embedding (850-876; normalize=True, scale=2*pi, T=10000) + level_embed (2262); C == 256
__global__ void __launch_bounds__(256)
pos_embed_kernel(const float* __restrict__ ycum, const float* __restrict__ xcum, GeoLevels lv, int S, const float* __restrict__ level_embed,
                 const float* __restrict__ dim_t_tab, float* __restrict__ pos) {
  pdl_entry();
  // 2 tokens per CTA; thread = one (sin, cos) channel pair: channels 2j, 2j+1 share dim_t (860-865), so one sincosf serves both
  const int tok = blockIdx.x * 2 + (threadIdx.x >> 7), b = blockIdx.y, j = threadIdx.x & 127;
  if (tok >= S) return;
  int l = 0;
  while (l + 1 < lv.L && tok >= lv.start[l + 1]) ++l;
  const int w = lv.w[l], h = lv.h[l];
  const int pix = tok - lv.start[l];
  const int y = pix / w, x = pix - y * w;
  const long long base = (long long)b * S + lv.start[l];
  const bool is_y = j < 64;
  const int i = (is_y ? j : j - 64) * 2;  // even channel index inside the y (or x) half
  float e, last;
  if (is_y) { e = ycum[base + pix]; last = ycum[base + (h - 1) * w + x]; }
  else      { e = xcum[base + pix]; last = xcum[base + y * w + (w - 1)]; }
  const float v = (e - 0.5f) / (last + 1e-6f) * 6.283185307179586f;
  // dim_t = 10000^(2*(i//2)/128) comes from a host-made table so that fully padded columns, whose normalised
  // coordinate is -0.5/1e-6 (the reference divides by last+eps, 857-858), see bit-identical arguments.
  const float a = v / dim_t_tab[i];
  float sn, cs;
  sincosf(a, &sn, &cs);
  const int c = (is_y ? 0 : 128) + i;
  const float2 le = *(const float2*)(level_embed + l * 256 + c);
  *(float2*)(pos + ((long long)b * S + tok) * 256 + c) = make_float2(sn + le.x, cs + le.y);
}

// ------------------------------------------------------------------ decoder self-attention core
// grid (heads, B, ceil(N/8)); 8 warps, one query per warp; K,V of the head staged in smem.
__global__ void __launch_bounds__(256)
mha_core_kernel(const float* __restrict__ qkv, int ld, int N, int C, float* __restrict__ out) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                   // [N][33]
  float* vs = sm + (((size_t)N * 33 + 3) & ~(size_t)3);  // [N][32], 16-byte aligned for the float4 staging stores
  const int hd = blockIdx.x, b = blockIdx.y;
  const float* base = qkv + (long long)b * N * ld + hd * 32;
  for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {  // float4 loads, all independent: one round trip for the whole head
    const int j = i >> 3, d = (i & 7) * 4;
    const float4 kv = __ldg((const float4*)(base + (long long)j * ld + C + d));
    const float4 vv = __ldg((const float4*)(base + (long long)j * ld + 2 * C + d));
    ks[j * 33 + d] = kv.x; ks[j * 33 + d + 1] = kv.y; ks[j * 33 + d + 2] = kv.z; ks[j * 33 + d + 3] = kv.w;
    *(float4*)(vs + j * 32 + d) = vv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.z * 8 + warp;
  if (qi >= N) return;
  const float qd = base[(long long)qi * ld + lane];  // lane d holds q[d] (already scaled)
  float sc[10];  // N <= 320
  float mx = -INFINITY;
  const int nj = (N + 31) >> 5;
  for (int t = 0; t < nj; ++t) {
    const int j = t * 32 + lane;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float q = __shfl_sync(0xffffffffu, qd, d);
      if (j < N) s = fmaf(q, ks[j * 33 + d], s);
    }
    sc[t] = (j < N) ? s : -INFINITY;
    mx = fmaxf(mx, sc[t]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = 0; t < nj; ++t) {
    sc[t] = (t * 32 + lane < N) ? expf(sc[t] - mx) : 0.f;
    sum += sc[t];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float acc = 0.f;
  for (int t = 0; t < nj; ++t) {
    const float p = sc[t] * inv;
    const int jmax = min(32, N - t * 32);
    for (int jj = 0; jj < jmax; ++jj) acc = fmaf(__shfl_sync(0xffffffffu, p, jj), vs[(t * 32 + jj) * 32 + lane], acc);
  }
  out[((long long)b * N + qi) * C + hd * 32 + lane] = acc;
}

// ------------------------------------------------------------------ tiny-N linear (N <= 8), one warp per row
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ bvec, int rows, int K,
                    int N, int act, const float* __restrict__ ref, int ld_ref, int ref_rows, float* __restrict__ y, int ldy) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float xv = x[(long long)row * ldx + k];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < N) acc[n] = fmaf(xv, __ldg(w + (long long)n * K + k), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = warp_sum(acc[n]);
  if (lane < N) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n == lane) v = acc[n];
    if (bvec) v += bvec[lane];
    if (act == 2 && lane < 2) {  // bbox head: += inverse_sigmoid(reference point) (egtr.py:291-298, deformable_detr.py:658-662)
      float r = fminf(fmaxf(ref[(long long)(row % ref_rows) * ld_ref + lane], 0.f), 1.f);
      v += logf(fmaxf(r, 1e-5f) / fmaxf(1.f - r, 1e-5f));
    }
    if (act >= 1) v = sigmoidf_(v);
    y[(long long)row * ldy + lane] = v;
  }
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_add_layernorm_f32(const float* x, const float* res, const float* gamma, const float* beta, int rows,
                                      int C, float* out, egtr_stream_t s) {
  EGTR_CHECK(x && gamma && beta && out && rows > 0, EGTR_ERR_ARG, "egtr_add_layernorm_f32: bad arguments");
  EGTR_CHECK(C == 256, EGTR_ERR_UNSUPPORTED, "egtr_add_layernorm_f32: built for d_model 256 (got %d)", C);
  launch_pdl(add_layernorm256_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)(0), (cudaStream_t)s, x, res, gamma, beta, rows, out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_add_layernorm_p32(const float* x, const void* res, int res_fmt, const float* gamma, const float* beta, int rows,
                                      int C, void* out_p32, float* out_f32, const float* addend, void* out_plus_p32, egtr_stream_t s) {
  EGTR_CHECK(x && gamma && beta && rows > 0 && (out_p32 || out_f32 || out_plus_p32), EGTR_ERR_ARG, "egtr_add_layernorm_p32: bad arguments");
  EGTR_CHECK(!out_plus_p32 || addend, EGTR_ERR_ARG, "egtr_add_layernorm_p32: out_plus_p32 needs an addend");
  EGTR_CHECK(C == 256, EGTR_ERR_UNSUPPORTED, "egtr_add_layernorm_p32: built for d_model 256 (got %d)", C);
  launch_pdl(add_layernorm256_p32_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)(0), (cudaStream_t)s, x, (const uint8_t*)res, res_fmt, gamma, beta, rows,
                                                                         (uint8_t*)out_p32, out_f32, addend, (uint8_t*)out_plus_p32);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_sum_layernorm_f32(const float* partial, int splits, long long split_stride, const float* bias, const float* res,
                                      const float* gamma, const float* beta, int rows, int C, float* out, float* out2, int rows_per_b2,
                                      long long bstride2, egtr_stream_t s) {
  EGTR_CHECK(partial && bias && gamma && beta && out && rows > 0 && splits >= 1, EGTR_ERR_ARG, "egtr_sum_layernorm_f32: bad arguments");
  EGTR_CHECK(C == 256, EGTR_ERR_UNSUPPORTED, "egtr_sum_layernorm_f32: built for d_model 256 (got %d)", C);
  EGTR_CHECK(!out2 || rows_per_b2 > 0, EGTR_ERR_ARG, "egtr_sum_layernorm_f32: out2 needs rows_per_b2");
  launch_pdl(sum_layernorm256_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)0, (cudaStream_t)s, partial, splits, split_stride, bias, res,
             gamma, beta, rows, out, out2, rows_per_b2 > 0 ? rows_per_b2 : rows, bstride2);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_rows_to_p32(const float* x, const float* addend, int rows, int C, int ldx, void* out, egtr_stream_t s) {
  EGTR_CHECK(x && out && rows > 0 && C > 0 && C % 32 == 0 && ldx % 4 == 0 && ldx >= C, EGTR_ERR_ARG, "egtr_rows_to_p32: bad arguments (C=%d ldx=%d)", C, ldx);
  const long long total = (long long)rows * (C / 4);
  launch_pdl(rows_to_p32_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, x, addend, rows, C / 4, ldx, (uint8_t*)out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_p32_to_rows(const void* p32, int rows, int C, float* out, int ldo, egtr_stream_t s) {
  EGTR_CHECK(p32 && out && rows > 0 && C > 0 && C % 32 == 0 && ldo % 4 == 0 && ldo >= C, EGTR_ERR_ARG, "egtr_p32_to_rows: bad arguments (C=%d ldo=%d)", C, ldo);
  const long long total = (long long)rows * (C / 4);
  launch_pdl(p32_to_rows_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, (const uint8_t*)p32, rows, C / 4, out, ldo);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_mask_rows_f32(float* x, int ld, int C, const uint8_t* keep, int rows, egtr_stream_t s) {
  EGTR_CHECK(x && keep && rows > 0 && C % 4 == 0 && ld % 4 == 0, EGTR_ERR_ARG, "egtr_mask_rows_f32: bad arguments");
  const long long total = (long long)rows * (C / 4);
  launch_pdl(mask_rows_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, x, ld, C / 4, keep, rows);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_pad_nchw3_to_nhwc4_f32(const float* img, int B, int H, int W, int pad, float* out, egtr_stream_t s) {
  EGTR_CHECK(img && out && B > 0 && H > 0 && W > 0 && pad >= 0, EGTR_ERR_ARG, "egtr_pad_nchw3_to_nhwc4_f32: bad arguments");
  const long long total = (long long)B * (H + 2 * pad) * (W + 2 * pad);
  launch_pdl(pad_nhwc4_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, img, B, H, W, pad, (float4*)out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_maxpool3x3s2_nhwc_f32(const float* x, int B, int H, int W, int C, float* out, egtr_stream_t s) {
  return egtr_maxpool3x3s2_nhwc_ex(x, B, H, W, C, out, EGTR_FMT_F32, s);
}

extern "C" int egtr_maxpool3x3s2_nhwc_ex(const float* x, int B, int H, int W, int C, void* out, int out_fmt, egtr_stream_t s) {
  EGTR_CHECK(x && out && B > 0 && H > 0 && W > 0 && C % 4 == 0 && (out_fmt == EGTR_FMT_F32 || C % 32 == 0), EGTR_ERR_ARG,
             "egtr_maxpool3x3s2_nhwc: bad arguments");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)B * OH * OW * (C / 4);
  launch_pdl(maxpool_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, x, B, H, W, C / 4, OH, OW, (float*)out, out_fmt);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_groupnorm_f32(float* x, int B, int rows_per_b, int bstride, int off, int C, int groups,
                                  const float* gamma, const float* beta, double* scratch, egtr_stream_t s) {
  return egtr_groupnorm_ex(x, B, rows_per_b, bstride, off, C, groups, gamma, beta, scratch, nullptr, nullptr, nullptr, s);
}

extern "C" int egtr_groupnorm_ex(float* x, int B, int rows_per_b, int bstride, int off, int C, int groups, const float* gamma,
                                 const float* beta, double* scratch, void* out_p32, const float* addend, void* out_plus_p32,
                                 egtr_stream_t s) {
  EGTR_CHECK(x && gamma && beta && scratch && B > 0 && rows_per_b > 0 && (!out_plus_p32 || addend), EGTR_ERR_ARG,
             "egtr_groupnorm: bad arguments");
  EGTR_CHECK(C == 256 && groups == 32, EGTR_ERR_UNSUPPORTED, "egtr_groupnorm_f32: built for GroupNorm(32, 256)");
  dim3 grid(cdiv(rows_per_b, GN_ROWS), B);
  double* part = scratch + B * 32;  // first B*32 doubles hold the float2 (mean, rstd) table
  float2* stats = (float2*)scratch;
  launch_pdl(groupnorm_partial_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, x, rows_per_b, bstride, off, part);
  launch_pdl(groupnorm_finalize_kernel, dim3(dim3(32, B)), dim3(32), (size_t)(0), (cudaStream_t)s, part, grid.x, rows_per_b, stats);
  launch_pdl(groupnorm_apply_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, x, rows_per_b, bstride, off, stats, gamma, beta,
             (uint8_t*)out_p32, addend, (uint8_t*)out_plus_p32);
  count_launch();
  count_launch();
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" long long egtr_groupnorm_scratch_doubles(int B, int rows_per_b) {
  return (long long)B * 32 + (long long)B * cdiv(rows_per_b, GN_ROWS) * 32 * 2;
}

extern "C" int egtr_levels_geometry_f32(const int64_t* pixel_mask, int B, int H, int W, const int* shapes_hw, int L,
                                        const float* level_embed, const float* dim_t, int C, uint8_t* mask_flat, float* pos_flat,
                                        float* valid_ratios, float* scratch, egtr_stream_t s) {
  EGTR_CHECK(pixel_mask && shapes_hw && level_embed && dim_t && mask_flat && pos_flat && valid_ratios && scratch, EGTR_ERR_ARG,
             "egtr_levels_geometry_f32: null pointer");
  EGTR_CHECK(C == 256 && L >= 1 && L <= 8 && B > 0 && B <= 65535, EGTR_ERR_UNSUPPORTED, "egtr_levels_geometry_f32: C=%d L=%d", C, L);
  GeoLevels lv;
  lv.L = L;
  int S = 0;
  for (int l = 0; l < L; ++l) {
    lv.h[l] = shapes_hw[2 * l];
    lv.w[l] = shapes_hw[2 * l + 1];
    lv.start[l] = S;
    S += lv.h[l] * lv.w[l];
  }
  float* ycum = scratch;
  float* xcum = scratch + (long long)B * S;
  launch_pdl(level_mask_gather_kernel, dim3(dim3(cdiv(S, 256), B)), dim3(256), (size_t)(0), (cudaStream_t)s, pixel_mask, H, W, lv, S, mask_flat);
  launch_pdl(level_scans_kernel, dim3(dim3(L, B)), dim3(256), (size_t)(0), (cudaStream_t)s, lv, S, mask_flat, ycum, xcum, valid_ratios);
  count_launch();
  launch_pdl(pos_embed_kernel, dim3(dim3(cdiv(S, 2), B)), dim3(256), (size_t)(0), (cudaStream_t)s, ycum, xcum, lv, S, level_embed, dim_t, pos_flat);
  count_launch();
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_mha_core_f32(const float* qkv, int ld, int B, int N, int heads, int D, float* out, egtr_stream_t s) {
  EGTR_CHECK(qkv && out && B > 0 && N > 0, EGTR_ERR_ARG, "egtr_mha_core_f32: bad arguments");
  EGTR_CHECK(D == 32 && N <= 320 && ld >= 3 * heads * D, EGTR_ERR_UNSUPPORTED, "egtr_mha_core_f32: head_dim 32, N <= 320 (D=%d N=%d)", D, N);
  const size_t smem = ((size_t)N * (33 + 32) + 4) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    EGTR_CUDA(cudaFuncSetAttribute(mha_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (320 * 65 + 4) * 4));
    attr = true;
  }
  launch_pdl(mha_core_kernel, dim3(dim3(heads, B, cdiv(N, 8))), dim3(256), (size_t)(smem), (cudaStream_t)s, qkv, ld, N, heads * D, out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_small_linear_f32(const float* x, int ldx, const float* w, const float* b, int rows, int K, int N,
                                     int act, const float* ref, int ld_ref, int ref_rows, float* y, int ldy, egtr_stream_t s) {
  EGTR_CHECK(x && w && y && rows > 0 && K > 0 && N >= 1 && N <= 8, EGTR_ERR_ARG, "egtr_small_linear_f32: bad arguments (N=%d)", N);
  EGTR_CHECK(act != 2 || ref != nullptr, EGTR_ERR_ARG, "egtr_small_linear_f32: act 2 needs reference points");
  launch_pdl(small_linear_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)(0), (cudaStream_t)s, x, ldx, w, b, rows, K, N, act, ref, ld_ref, ref_rows > 0 ? ref_rows : rows, y, ldy);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
