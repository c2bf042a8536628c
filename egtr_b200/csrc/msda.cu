// Multi-scale deformable attention forward (K1 of SURVEY.md §2.3), rewritten for B200.
//
// Reference semantics: model/custom_kernel/cuda/ms_deform_im2col_cuda.cuh:237-299 (+ bilinear 33-84):
//   out[b,q,m,:] = sum_{l,p} w[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p]),
//   pixel coords x = loc_x*W - 0.5, a sample counts only if -1 < y < H and -1 < x < W, corners outside
//   the map read as zero.
//
// The kernel is bound by the SM's L1/L2 gather path (each (q,m) pulls 16 samples x 4 corners x 128 B),
// not by compulsory HBM traffic, so the design goals are (1) every gather is a full 128-byte line read
// by 8 lanes x float4, (2) sample arithmetic is done once per sample, not once per channel lane, and
// (3) queries that share cache lines run in the same CTA: a CTA owns ONE head and a 8x4 pixel patch of
// encoder queries (neighbouring pixels sample neighbouring tokens), or 32 consecutive decoder queries.
//   phase 1: 256 threads = 16 (query) x 16 (sample) lanes: read the raw offsets/logits (fused form) or
//            the precomputed locations/weights (drop-in form), softmax over the 16 samples with
//            16-lane shuffles, turn each sample into 4 corner offsets (32-bit elements) + 4 combined weights in smem;
//   phase 2: 32 groups of 8 lanes: stream the 16 samples of one query, 4 LDG.128 each (unconditional in the fused form:
//            absent corners carry weight 0 and point at token 0).
#include <cuda_fp16.h>

#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

constexpr int MAX_L = 8;
constexpr int SLOT_WORDS = 8;                      // 4 corner offsets (elements, token * row stride) + 4 weights
constexpr int Q_STRIDE = 16 * SLOT_WORDS + 8;      // words per query, padded against bank conflicts

struct Levels {
  int L;
  int H[MAX_L], W[MAX_L], start[MAX_L];
  int patch_start[MAX_L + 1];  // encoder patch enumeration (8x4 pixel patches per level)
  int patches_x[MAX_L];
};

struct MsdaArgs {
  const float* value; int ld_value;       // [B,S,...] row stride in floats; head m at column m*32
  const float* offaw; int ld_offaw;       // fused: per query [M*L*P*2 offsets | M*L*P logits]
  const float* ref_points;                // fused decoder: [Lq,2], scaled by valid_ratios[b,l] in-kernel
  const float* valid_ratios;              // fused encoder: [B,L,2]
  const float* loc;                       // drop-in: [B,Lq,M,L,P,2]
  const float* attw;                      // drop-in: [B,Lq,M,L,P]
  const int64_t* dev_shapes;              // drop-in: [L,2] int64 on device
  const int64_t* dev_start;               // drop-in: [L] int64 on device
  float* out;                             // [B,Lq,M*32]
  int out_fmt;                            // 1: P32 rows (one head = one 32-channel group: hi 64 B | lo 64 B)
  int B, S, M, Lq;
  int enc_patches;                        // 1: queries are the S tokens, enumerated in 8x4 patches
  // H16 form: value as fp16 pair records [heads][records][2 slots][32] (include/egtr_b200.h, EGTR_FMT_H16PAIR)
  const uint8_t* value_h16;
  long long h16_records;                  // records per head = B*S + 1
  int head0;                              // first head of this launch inside the record tensor (decoder: layer * M)
};

// QPB queries per CTA, 8 threads each: 32 for the encoder's 8x4 pixel patches, 8 for the decoder's few hundred queries
// (25 x 8 CTAs instead of 7 x 8, and half as many dependent gather batches per thread).
#ifndef EGTR_MSDA_MINB   // dev A/B knobs: resident CTAs per SM and samples in flight per thread of the 32-query form
#define EGTR_MSDA_MINB 6
#endif
#ifndef EGTR_MSDA_UNROLL
#define EGTR_MSDA_UNROLL 4
#endif
constexpr int kMsdaUnroll = EGTR_MSDA_UNROLL;
#ifndef EGTR_MSDA_FMA2   // H16 form: accumulate channel pairs with packed fma.rn.f32x2 (same IEEE result per lane)
#define EGTR_MSDA_FMA2 1
#endif
__device__ __forceinline__ unsigned long long pack_f32x2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long y, unsigned long long z) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  return r;
}
#ifndef EGTR_MSDA_ENC_QPB  // encoder patch: 32 = 8x4 pixels (256 threads), 64 = 8x8 pixels (512 threads)
#define EGTR_MSDA_ENC_QPB 32
#endif
template <bool FUSED, int QPB, bool BYPASS_L1 = false, bool H16 = false>
__global__ void __launch_bounds__(QPB * 8, QPB == 32 ? EGTR_MSDA_MINB : (QPB == 64 ? EGTR_MSDA_MINB / 2 : 8))
msda_kernel(const MsdaArgs a, const Levels lv_in) {
  pdl_entry();
  __shared__ float slots[QPB * Q_STRIDE];
  __shared__ int q_of[QPB];
  __shared__ float2 q_ref[QPB];  // encoder form: pixel centre / (valid_ratio * size) of the query's own level
  __shared__ int lvH[MAX_L], lvW[MAX_L], lvS[MAX_L];

  const int tid = threadIdx.x;
  const int m = blockIdx.y;
  const int b = blockIdx.z;
  const int L = lv_in.L;

  if (tid < L) {
    if (FUSED) {
      lvH[tid] = lv_in.H[tid]; lvW[tid] = lv_in.W[tid]; lvS[tid] = lv_in.start[tid];
    } else {  // the reference op takes its shape tensors on the device (ms_deform_attn.h:20-39)
      lvH[tid] = (int)a.dev_shapes[2 * tid]; lvW[tid] = (int)a.dev_shapes[2 * tid + 1]; lvS[tid] = (int)a.dev_start[tid];
    }
  }
  // ---- which queries does this CTA own?
  if (tid < QPB) {
    int q = -1;
    if (a.enc_patches) {
      int l = 0;
      while (l + 1 < L && (int)blockIdx.x >= lv_in.patch_start[l + 1]) ++l;
      const int pid = blockIdx.x - lv_in.patch_start[l];
      const int py = pid / lv_in.patches_x[l], px = pid - py * lv_in.patches_x[l];
      const int y = py * (QPB / 8) + (tid >> 3), x = px * 8 + (tid & 7);
      if (y < lv_in.H[l] && x < lv_in.W[l]) {
        q = lv_in.start[l] + y * lv_in.W[l] + x;
        // deformable_detr.py:1642-1644: linspace(0.5, n-0.5, n)[i] / (valid_ratio * n); the per-level
        // valid-ratio factor of line 1647 is applied per sample below
        const float* vr = a.valid_ratios + (long long)b * L * 2;
        q_ref[tid] = make_float2(((float)x + 0.5f) / (vr[l * 2 + 0] * (float)lv_in.W[l]),
                                 ((float)y + 0.5f) / (vr[l * 2 + 1] * (float)lv_in.H[l]));
      }
    } else {
      q = blockIdx.x * QPB + tid;
      if (q >= a.Lq) q = -1;
    }
    q_of[tid] = q;
  }
  __syncthreads();

  // ---- phase 1: one thread per (query, sample); two passes cover the 32 queries
  const int s = tid & 15;          // sample index = l*P + p   (L*P == 16)
  // the raw offsets / logits of BOTH passes are requested before any of the softmax shuffles: one exposed load latency per CTA
  float2 off_in[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  float logit_in[2] = {0.f, 0.f};
  if (FUSED) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int q = q_of[(tid >> 4) + pass * (QPB / 2)];
      if (q >= 0) {
        const float* row = a.offaw + ((long long)b * a.Lq + q) * a.ld_offaw;
        if (BYPASS_L1) {
          off_in[pass] = __ldcg((const float2*)(row + (m * 16 + s) * 2));
          logit_in[pass] = __ldcg(row + a.M * 32 + m * 16 + s);
        } else {
          off_in[pass] = *(const float2*)(row + (m * 16 + s) * 2);
          logit_in[pass] = row[a.M * 32 + m * 16 + s];
        }
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int qi = (tid >> 4) + pass * (QPB / 2);
    const int q = q_of[qi];
    float* slot = &slots[qi * Q_STRIDE + s * SLOT_WORDS];
    // corners outside the map (and samples outside the image) keep weight 0 and point at token 0: phase 2 can then load
    // unconditionally — no predicates and no register zeroing per gather (values are finite, so 0 * v contributes nothing)
    // (fused form only: its value tensor is this library's own finite GEMM output; the drop-in op keeps predicated loads)
    constexpr int kNone = FUSED ? 0 : -1;
    int idx[4] = {kNone, kNone, kNone, kNone};
    float cw[4] = {0.f, 0.f, 0.f, 0.f};
    // L*P == 16 with P == 4 is asserted on the host.
    const int l = s >> 2;
    const float2 off = off_in[pass];
    float wgt = 0.f;
    if (FUSED) {
      // softmax over the 16 samples of this (query, head): 16-lane butterflies, executed by every
      // lane (absent queries feed zeros) so the full-mask shuffles stay convergent.
      const float logit = logit_in[pass];
      float mx = logit;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      // softmax over 16 logits: ex2.approx + fast divide are within 2 ulp — far inside the 1e-3 parity bar and cheap
      // enough that phase 1 stops being a third of the kernel's instruction count
      const float e = __expf(logit - mx);
      float sum = e;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      wgt = __fdividef(e, sum);
    }
    if (q >= 0) {
      const int H = lvH[l], W = lvW[l];
      float lx, ly;
      if (FUSED) {
        float rx, ry;
        if (a.enc_patches) {
          // deformable_detr.py:1647: reference_points[:, :, None] * valid_ratios[:, None]
          const float* vr = a.valid_ratios + (long long)b * L * 2;
          const float2 r0 = q_ref[qi];
          rx = r0.x * vr[l * 2 + 0];
          ry = r0.y * vr[l * 2 + 1];
        } else {
          // decoder: reference_points[:, :, None] * valid_ratios[:, None] (deformable_detr.py:1865-1867);
          // the sigmoid'ed points are shared by the batch ([Lq,2]) because there is no box refinement.
          const float2 r = *(const float2*)(a.ref_points + (long long)q * 2);
          const float* vr = a.valid_ratios + (long long)b * L * 2;
          rx = r.x * vr[l * 2 + 0]; ry = r.y * vr[l * 2 + 1];
        }
        lx = rx + __fdividef(off.x, (float)W);  // deformable_detr.py:1066-1073 (offset / (W_l, H_l))
        ly = ry + __fdividef(off.y, (float)H);
      } else {
        const long long base = (((long long)b * a.Lq + q) * a.M + m) * 16 + s;
        const float2 lc = *(const float2*)(a.loc + base * 2);
        lx = lc.x; ly = lc.y;
        wgt = a.attw[base];
      }
      const float him = ly * (float)H - 0.5f, wim = lx * (float)W - 0.5f;  // cuh:285-286
      if (him > -1.f && wim > -1.f && him < (float)H && wim < (float)W) {   // cuh:288
        const int hl = (int)floorf(him), wl = (int)floorf(wim);
        const float lh = him - (float)hl, lw = wim - (float)wl;
        const float hh = 1.f - lh, hw = 1.f - lw;
        // corners are stored as 32-bit ELEMENT offsets into this image's value rows (token * row stride; the host entry points
        // guarantee S * ld_value < 2^31), so phase 2 forms each gather address with one IMAD.WIDE instead of a 64-bit multiply
        const bool y0 = hl >= 0, y1 = hl + 1 <= H - 1, x0 = wl >= 0, x1 = wl + 1 <= W - 1;
        if (H16) {
          // pair records: tokens (t, t+1) of one map row are record t + 1 (wl = -1 -> the record whose slot 1 is the row's
          // first token); idx[0] / idx[1] = records of the top / bottom row, cw = (top-left, top-right, bottom-left, bottom-right)
          const int rec = b * a.S + lvS[l] + hl * W + wl + 1;
          if (y0) idx[0] = rec;
          if (y1) idx[1] = rec + W;
          if (y0 && x0) cw[0] = hh * hw * wgt;
          if (y0 && x1) cw[1] = hh * lw * wgt;
          if (y1 && x0) cw[2] = lh * hw * wgt;
          if (y1 && x1) cw[3] = lh * lw * wgt;
        } else {
          const int ld = a.ld_value;
          const int base_off = (lvS[l] + hl * W + wl) * ld;
          if (y0 && x0) { idx[0] = base_off;               cw[0] = hh * hw * wgt; }
          if (y0 && x1) { idx[1] = base_off + ld;          cw[1] = hh * lw * wgt; }
          if (y1 && x0) { idx[2] = base_off + W * ld;      cw[2] = lh * hw * wgt; }
          if (y1 && x1) { idx[3] = base_off + W * ld + ld; cw[3] = lh * lw * wgt; }
        }
      }
    }
    *(int4*)slot = make_int4(idx[0], idx[1], idx[2], idx[3]);
    *(float4*)(slot + 4) = make_float4(cw[0], cw[1], cw[2], cw[3]);
  }
  __syncthreads();

  // ---- phase 2: 8 lanes x float4 per query, 16 samples x 4 corners
  const int g = tid >> 3, c4 = (tid & 7) * 4;
  const int q = q_of[g];
  if (q < 0) return;
  if constexpr (H16) {
    // 8 lanes per query: lanes 0-3 read slot 0 (the left corner), lanes 4-7 slot 1 (the right corner) of the SAME 128-byte
    // record — one L1 wavefront per row of the bilinear footprint; lane j & 3 owns channels 8j .. 8j+7
    const int half = (tid >> 2) & 1, j8 = (tid & 3) * 8;
    const uint8_t* hb = a.value_h16 + ((long long)(a.head0 + m) * a.h16_records) * 128 + half * 64 + j8 * 2;
    unsigned long long vb;
    asm("mov.b64 %0, %1;" : "=l"(vb) : "l"(hb));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* myslots = &slots[g * Q_STRIDE];
#if EGTR_MSDA_FMA2
    // channel pairs stay packed as f32x2 register pairs: one FFMA2 per corner row per two channels (the kernel is issue-bound)
    unsigned long long acc2[4] = {0ull, 0ull, 0ull, 0ull};
#endif
#pragma unroll (QPB >= 32 ? kMsdaUnroll : 8)
    for (int ss = 0; ss < 16; ++ss) {
      const int2 id = *(const int2*)(myslots + ss * SLOT_WORDS);
      const float4 w = *(const float4*)(myslots + ss * SLOT_WORDS + 4);
      const float wt = half ? w.y : w.x, wb = half ? w.w : w.z;
      const uint4 t4 = __ldg((const uint4*)(vb + (unsigned long long)(uint32_t)id.x * 128ull));
      const uint4 b4 = __ldg((const uint4*)(vb + (unsigned long long)(uint32_t)id.y * 128ull));
      const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#if EGTR_MSDA_FMA2
      const unsigned long long wt2 = pack_f32x2(wt, wt), wb2 = pack_f32x2(wb, wb);
#endif
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&tw[k]));
        const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&bw[k]));
#if EGTR_MSDA_FMA2
        acc2[k] = fma2(wt2, pack_f32x2(tf.x, tf.y), acc2[k]);
        acc2[k] = fma2(wb2, pack_f32x2(bf.x, bf.y), acc2[k]);
#else
        acc[2 * k] = fmaf(wt, tf.x, acc[2 * k]); acc[2 * k + 1] = fmaf(wt, tf.y, acc[2 * k + 1]);
        acc[2 * k] = fmaf(wb, bf.x, acc[2 * k]); acc[2 * k + 1] = fmaf(wb, bf.y, acc[2 * k + 1]);
#endif
      }
    }
#if EGTR_MSDA_FMA2
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack_f32x2(acc2[k], acc[2 * k], acc[2 * k + 1]);
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffu << (threadIdx.x & 24), acc[k], 4);  // left + right corners (lanes j, j + 4 of this query's 8; other queries of the warp may have exited)
    float* orow = a.out + ((long long)b * a.Lq + q) * (a.M * 32) + m * 32;
    if (a.out_fmt == 0) {
      *(float4*)(orow + j8 + half * 4) = half ? make_float4(acc[4], acc[5], acc[6], acc[7]) : make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {  // P32 row: lanes 0-3 store the hi halves (16 B each), lanes 4-7 the lo halves 64 bytes further
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
        const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&h2);
        if (half) {
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(acc[2 * k] - __uint_as_float(hbits << 16), acc[2 * k + 1] - __uint_as_float(hbits & 0xffff0000u));
          o[k] = *reinterpret_cast<const uint32_t*>(&l2);
        } else {
          o[k] = hbits;
        }
      }
      *(uint4*)((uint8_t*)orow + half * 64 + j8 * 2) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  const float* vbase = a.value + (long long)b * a.S * a.ld_value + m * 32 + c4;
  // keep the per-thread base opaque in one register pair: address = IMAD.WIDE.U32(offset, 4, base), one instruction per gather
  // (left to itself the compiler re-associates base = uniform pointer + 64-bit element offset: four instructions per address)
  unsigned long long vb;
  asm("mov.b64 %0, %1;" : "=l"(vb) : "l"(vbase));
  auto ldv = [](const float4* ptr) { return BYPASS_L1 ? __ldcg(ptr) : __ldg(ptr); };
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* myslots = &slots[g * Q_STRIDE];
#pragma unroll (QPB >= 32 ? kMsdaUnroll : 8)
  for (int ss = 0; ss < 16; ++ss) {
    const int4 id = *(const int4*)(myslots + ss * SLOT_WORDS);
    const float4 w = *(const float4*)(myslots + ss * SLOT_WORDS + 4);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
    if (FUSED || id.x >= 0) v0 = ldv((const float4*)(vb + (unsigned long long)(uint32_t)id.x * 4ull));
    if (FUSED || id.y >= 0) v1 = ldv((const float4*)(vb + (unsigned long long)(uint32_t)id.y * 4ull));
    if (FUSED || id.z >= 0) v2 = ldv((const float4*)(vb + (unsigned long long)(uint32_t)id.z * 4ull));
    if (FUSED || id.w >= 0) v3 = ldv((const float4*)(vb + (unsigned long long)(uint32_t)id.w * 4ull));
    acc.x = fmaf(w.x, v0.x, acc.x); acc.y = fmaf(w.x, v0.y, acc.y); acc.z = fmaf(w.x, v0.z, acc.z); acc.w = fmaf(w.x, v0.w, acc.w);
    acc.x = fmaf(w.y, v1.x, acc.x); acc.y = fmaf(w.y, v1.y, acc.y); acc.z = fmaf(w.y, v1.z, acc.z); acc.w = fmaf(w.y, v1.w, acc.w);
    acc.x = fmaf(w.z, v2.x, acc.x); acc.y = fmaf(w.z, v2.y, acc.y); acc.z = fmaf(w.z, v2.z, acc.z); acc.w = fmaf(w.z, v2.w, acc.w);
    acc.x = fmaf(w.w, v3.x, acc.x); acc.y = fmaf(w.w, v3.y, acc.y); acc.z = fmaf(w.w, v3.z, acc.z); acc.w = fmaf(w.w, v3.w, acc.w);
  }
  float* orow = a.out + ((long long)b * a.Lq + q) * (a.M * 32) + m * 32;
  if (a.out_fmt == 0) {
    *(float4*)(orow + c4) = acc;
  } else {
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    split_bf16(acc.x, h0, l0); split_bf16(acc.y, h1, l1); split_bf16(acc.z, h2, l2); split_bf16(acc.w, h3, l3);
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
    uint8_t* g = (uint8_t*)orow + c4 * 2;
    *(uint2*)g = ph;
    *(uint2*)(g + 64) = pl;
  }
}

int fill_levels(const int* shapes_hw, int L, Levels* lv, int* S_out, int patch_h = 4) {
  lv->L = L;
  int start = 0, pstart = 0;
  for (int l = 0; l < L; ++l) {
    lv->H[l] = shapes_hw[2 * l];
    lv->W[l] = shapes_hw[2 * l + 1];
    lv->start[l] = start;
    start += lv->H[l] * lv->W[l];
    lv->patch_start[l] = pstart;
    lv->patches_x[l] = (lv->W[l] + 7) / 8;
    pstart += lv->patches_x[l] * ((lv->H[l] + patch_h - 1) / patch_h);
  }
  lv->patch_start[L] = pstart;
  *S_out = start;
  return pstart;
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_msda_fwd_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                 const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D, int L,
                                 int Lq, int P, float* out, egtr_stream_t s) {
  EGTR_CHECK(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out, EGTR_ERR_ARG,
             "egtr_msda_fwd_f32: null pointer");
  EGTR_CHECK(B > 0 && S > 0 && M > 0 && Lq > 0, EGTR_ERR_ARG, "egtr_msda_fwd_f32: empty shape");
  EGTR_CHECK(D == 32 && L * P == 16 && P == 4 && L <= MAX_L, EGTR_ERR_UNSUPPORTED,
             "egtr_msda_fwd_f32: built for head_dim 32 and L*P = 4*4 (got D=%d L=%d P=%d)", D, L, P);
  EGTR_CHECK((long long)B * S * M * D < (1LL << 31), EGTR_ERR_ARG, "egtr_msda_fwd_f32: B*S*M*D must be < 2^31");
  EGTR_CHECK(B <= 65535 && M <= 65535, EGTR_ERR_ARG, "egtr_msda_fwd_f32: grid limits");
  MsdaArgs a = {};
  a.value = value; a.ld_value = M * D;
  a.loc = sampling_loc; a.attw = attn_weight;
  a.dev_shapes = spatial_shapes; a.dev_start = level_start_index;
  a.out = out; a.B = B; a.S = S; a.M = M; a.Lq = Lq; a.enc_patches = 0;
  Levels lv = {};
  lv.L = L;
  dim3 grid(cdiv(Lq, 32), M, B);
  launch_pdl(msda_kernel<false, 32>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, a, lv);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_msda_fused_fwd_f32(const float* value, int ld_value, const int* shapes_hw, const float* offaw,
                                       int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                                       int B, int S, int M, int D, int L, int Lq, int P, float* out, egtr_stream_t s) {
  return egtr_msda_fused_fwd_ex(value, ld_value, shapes_hw, offaw, ld_offaw, ref_points, valid_ratios, enc_ref, B, S, M, D, L, Lq,
                                P, out, EGTR_FMT_F32, s);
}

extern "C" int egtr_msda_fused_fwd_ex(const float* value, int ld_value, const int* shapes_hw, const float* offaw,
                                      int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                                      int B, int S, int M, int D, int L, int Lq, int P, void* out_v, int out_fmt, egtr_stream_t s) {
  float* out = (float*)out_v;
  EGTR_CHECK(value && shapes_hw && offaw && out, EGTR_ERR_ARG, "egtr_msda_fused_fwd_f32: null pointer");
  EGTR_CHECK(valid_ratios != nullptr && (enc_ref || ref_points != nullptr), EGTR_ERR_ARG,
             "egtr_msda_fused_fwd_f32: reference points missing");
  EGTR_CHECK(D == 32 && L * P == 16 && P == 4 && L <= MAX_L, EGTR_ERR_UNSUPPORTED,
             "egtr_msda_fused_fwd_f32: built for head_dim 32 and L*P = 4*4 (got D=%d L=%d P=%d)", D, L, P);
  EGTR_CHECK(ld_value % 4 == 0 && ld_value >= M * D && ld_offaw >= M * L * P * 3 && ld_offaw % 2 == 0, EGTR_ERR_ARG,
             "egtr_msda_fused_fwd_f32: bad leading dimensions");
  EGTR_CHECK(B <= 65535 && M <= 65535, EGTR_ERR_ARG, "egtr_msda_fused_fwd_f32: grid limits");
  EGTR_CHECK((long long)S * ld_value < (1LL << 31), EGTR_ERR_ARG, "egtr_msda_fused_fwd_f32: S*ld_value must be < 2^31");
  Levels lv = {};
  int S_chk = 0;
  constexpr int kEncQ = EGTR_MSDA_ENC_QPB;
  const int patches = fill_levels(shapes_hw, L, &lv, &S_chk, kEncQ / 8);
  EGTR_CHECK(S_chk == S, EGTR_ERR_ARG, "egtr_msda_fused_fwd_f32: sum(H*W)=%d != S=%d", S_chk, S);
  EGTR_CHECK(!enc_ref || Lq == S, EGTR_ERR_ARG, "egtr_msda_fused_fwd_f32: encoder form needs Lq == S");
  MsdaArgs a = {};
  a.value = value; a.ld_value = ld_value;
  a.offaw = offaw; a.ld_offaw = ld_offaw;
  a.ref_points = ref_points; a.valid_ratios = valid_ratios;
  a.out = out; a.out_fmt = out_fmt; a.B = B; a.S = S; a.M = M; a.Lq = Lq; a.enc_patches = enc_ref ? 1 : 0;
  const bool cg = debug_flags() & 1;  // diagnostic (egtr_set_debug_flags bit 0): gathers and offset reads bypass L1
  if (!enc_ref && (long long)Lq * B <= 4096) {  // decoder-sized query sets: small CTAs for parallelism and latency
    dim3 grid(cdiv(Lq, 8), M, B);
    launch_pdl(msda_kernel<true, 8>, dim3(grid), dim3(64), (size_t)(0), (cudaStream_t)s, a, lv);
  } else if (enc_ref) {
    dim3 grid(patches, M, B);
    if (cg) launch_pdl(msda_kernel<true, kEncQ, true>, dim3(grid), dim3(kEncQ * 8), (size_t)(0), (cudaStream_t)s, a, lv);
    else launch_pdl(msda_kernel<true, kEncQ>, dim3(grid), dim3(kEncQ * 8), (size_t)(0), (cudaStream_t)s, a, lv);
  } else {
    dim3 grid(cdiv(Lq, 32), M, B);
    launch_pdl(msda_kernel<true, 32>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, a, lv);
  }
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_msda_fused_fwd_h16(const void* value_h16, long long records, int head0, int heads_total, const int* shapes_hw,
                                       const float* offaw, int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                                       int B, int S, int M, int D, int L, int Lq, int P, void* out_v, int out_fmt, egtr_stream_t s) {
  float* out = (float*)out_v;
  EGTR_CHECK(value_h16 && shapes_hw && offaw && out, EGTR_ERR_ARG, "egtr_msda_fused_fwd_h16: null pointer");
  EGTR_CHECK(valid_ratios != nullptr && (enc_ref || ref_points != nullptr), EGTR_ERR_ARG, "egtr_msda_fused_fwd_h16: reference points missing");
  EGTR_CHECK(D == 32 && L * P == 16 && P == 4 && L <= MAX_L, EGTR_ERR_UNSUPPORTED,
             "egtr_msda_fused_fwd_h16: built for head_dim 32 and L*P = 4*4 (got D=%d L=%d P=%d)", D, L, P);
  EGTR_CHECK(records == (long long)B * S + 1 && records < (1LL << 25) && head0 >= 0 && head0 + M <= heads_total, EGTR_ERR_ARG,
             "egtr_msda_fused_fwd_h16: records must be B*S + 1 (< 2^25), heads [%d, %d) of %d", head0, head0 + M, heads_total);
  EGTR_CHECK(((uintptr_t)value_h16 & 127) == 0 && ld_offaw >= M * L * P * 3 && ld_offaw % 2 == 0, EGTR_ERR_ARG,
             "egtr_msda_fused_fwd_h16: alignment / leading dimension");
  EGTR_CHECK(B <= 65535 && M <= 65535, EGTR_ERR_ARG, "egtr_msda_fused_fwd_h16: grid limits");
  Levels lv = {};
  int S_chk = 0;
  constexpr int kEncQ = EGTR_MSDA_ENC_QPB;
  const int patches = fill_levels(shapes_hw, L, &lv, &S_chk, kEncQ / 8);
  EGTR_CHECK(S_chk == S, EGTR_ERR_ARG, "egtr_msda_fused_fwd_h16: sum(H*W)=%d != S=%d", S_chk, S);
  EGTR_CHECK(!enc_ref || Lq == S, EGTR_ERR_ARG, "egtr_msda_fused_fwd_h16: encoder form needs Lq == S");
  MsdaArgs a = {};
  a.value_h16 = (const uint8_t*)value_h16; a.h16_records = records; a.head0 = head0;
  a.offaw = offaw; a.ld_offaw = ld_offaw;
  a.ref_points = ref_points; a.valid_ratios = valid_ratios;
  a.out = out; a.out_fmt = out_fmt; a.B = B; a.S = S; a.M = M; a.Lq = Lq; a.enc_patches = enc_ref ? 1 : 0;
  if (!enc_ref && (long long)Lq * B <= 4096) {
    dim3 grid(cdiv(Lq, 8), M, B);
    launch_pdl(msda_kernel<true, 8, false, true>, dim3(grid), dim3(64), (size_t)(0), (cudaStream_t)s, a, lv);
  } else if (enc_ref) {
    dim3 grid(patches, M, B);
    launch_pdl(msda_kernel<true, kEncQ, false, true>, dim3(grid), dim3(kEncQ * 8), (size_t)(0), (cudaStream_t)s, a, lv);
  } else {
    dim3 grid(cdiv(Lq, 32), M, B);
    launch_pdl(msda_kernel<true, 32, false, true>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, a, lv);
  }
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
