// The whole Deformable-DETR decoder stack for small query sets as ONE kernel: a thread-block cluster of eight CTAs per image walks
// the layers (reference: model/deformable_detr.py:1390-1489 layer, 1149-1262 self-attention with Q/K capture, 1774-1968 stack),
// every 200-row GEMM on tcgen05, phases separated by hardware cluster barriers instead of kernel launches.
//
// Round 1/2 ran each decoder layer as ten launches of latency-bound CUDA-core kernels (36 skinny GEMMs, 0.76 ms of a 3.4 ms
// forward, 10 % of the throughput configuration's step).  Here CTA r of the cluster owns HEAD r and column slice r:
//   QKV    q_r|k_r|v_r = h . W_r^T + (query_pos . W^T + b)          N = 96 columns, K = 256      (captured as fp32 [N,768] rows)
//   MHA    S = q_r k_r^T (two 128-row tiles, N = 256 keys, K = 32) -> softmax in registers from TMEM -> P (bf16 hi/lo) in
//          shared memory -> O = P v_r (K = keys)                      everything of head r stays on this SM
//   OPROJ  o[:, 32r:32r+32] = attn . Wo^T + b          LN1   t1 = LayerNorm(h + o)                (rows dealt over CTAs / warps)
//   OFFAW  sampling offsets | attention logits of t1 (+ query_pos term)        N = 64 columns on six CTAs
//   MSDA   head r of the multi-scale deformable gather over the fp16 pair records (msda.cu's decoder form)
//   OUTPROJ + LN2, FC1 (ReLU, 128 hidden columns per CTA, P32 rows), FC2 + LN3 (-> next layer's h, stacked intermediates)
// The `(x + query_pos) . W` projections are split as `x . W + query_pos . W`: the second term depends on weights only and is
// precomputed per layer in fp64 at load time (engine.py), so every GEMM streams ONE activation operand.
// Activations cross CTAs through L2 (a few hundred KB): generic stores -> fence.proxy.async -> barrier.cluster (release /
// acquire) -> TMA loads or ld.global.cg in the consumers.  bf16x3 split products as everywhere (fp32-equivalent).
#include <cuda.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace egtr {

void count_launch();
int tmap_p32_rows(const void* ptr, int channels, int rows_per_b, int nb, int box_rows, CUtensorMap* out);  // gemm_p32.cu
int tmap_weight_planes(const void* planes, int K, long long rows, int box_rows, CUtensorMap* out);        // gemm_p32.cu

namespace {

constexpr int CL = 8;                 // cluster size = heads = column slices
constexpr int THREADS = 256;          // warps 0-3: TMEM lane quadrants (epilogue / softmax), 4: TMA, 5: MMA, 6-7: SIMT phases only
constexpr int TMA_WARP = 4, MMA_WARP = 5;
constexpr int BK = 64;
constexpr int GROUP_BYTES = 128 * 128;  // one 32-channel P32 group of 128 rows
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 2 * GROUP_BYTES + 2 * 128 * BK * 2;  // activation k-block (32 KB) + weight hi/lo tiles of up to 128 rows
constexpr int RING_BYTES = STAGES * STAGE_BYTES;                 // 192 KB
// MHA phase (aliases the ring): Q (2 x 16 KB) | K (32 KB) | V^T (8 key groups x 4 KB) | P (8 key groups x 16 KB)
constexpr int MHA_Q = 0, MHA_K = 2 * GROUP_BYTES, MHA_VT = MHA_K + 2 * GROUP_BYTES, MHA_P = MHA_VT + 8 * 4096;
constexpr int MHA_BYTES = MHA_P + 8 * GROUP_BYTES;               // 224 KB
constexpr int CTRL_BYTES = 1024;
constexpr int SMEM_BYTES = (MHA_BYTES > RING_BYTES ? MHA_BYTES : RING_BYTES) + CTRL_BYTES + 1024 /*align*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int TMEM_COLS = 512;
constexpr int VEC_FLOATS = 9 * 256 + 1024;  // per-layer vector block: bo | bout | b2 | ln1 g,b | ln2 g,b | ln3 g,b | b1

enum Phase { PH_INIT = 0, PH_QKV, PH_MHA, PH_OPROJ, PH_LN1, PH_OFFAW, PH_MSDA, PH_OUTPROJ, PH_LN2, PH_FC1, PH_FC2, PH_LN3, PH_END };

struct DecMaps {
  CUtensorMap a_h, a_attn, a_t1, a_attn2, a_t2, a_f, a_qk, a_vt;
  CUtensorMap w_qkv, w_o, w_offaw, w_out, w_fc1, w_fc2;
};

struct DecArgs {
  int B, N, L, S, Lv, MT;
  int layer0, layer1, phase0, phase1, mha_mode;
  int plane_rows[6];  // rows of one bf16 plane per weight tensor (qkv, o, offaw, out, fc1, fc2)
  const float* vec;
  const float* qkv_pos;
  const float* off_pos;
  const float* tgt;
  const float* ref;
  const float* valid_ratios;
  float *hf, *t1f, *t2f, *o, *offaw;
  uint8_t *hp, *t1p, *t2p, *attn_p, *attn2_p, *f_p, *qk_p, *vt_p;
  float* qkv;    // [L][B*N][768]
  float* inter;  // [B][L][N][256]
  const uint8_t* value_h16;
  long long records;
  int lvH[4], lvW[4], lvS[4];
  int* err;
};

struct Pipe {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
};

struct Ctx {
  uint8_t* ring;
  uint64_t *full, *empty, *acc_full, *mha_bar;
  uint32_t tmem;
  int warp, lane, r, b;
  Pipe pt, pm;     // operand ring position of the TMA thread / the MMA warp (identical sequences)
  uint32_t gp;     // parity of acc_full: flips with every GEMM phase
  uint32_t mp;     // parity of the MHA barriers: flips with every MHA phase
  uint32_t pp;     // parity of the P-tile barrier: completes once per m-tile
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t smem_dst, const void* tmap, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 32 fp32 values -> one P32 group: words 0-15 = bf16 hi pairs, 16-31 = bf16 lo pairs
__device__ __forceinline__ void split_group(const float (&v)[32], uint32_t (&o)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t h = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    o[j] = h;
    o[16 + j] = pack_bf16x2(v[2 * j] - __uint_as_float(h << 16), v[2 * j + 1] - __uint_as_float(h & 0xffff0000u));
  }
}
__device__ __forceinline__ void store_group_global(uint8_t* dst, const uint32_t (&o)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) *(uint4*)(dst + c * 16) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
}

// End of a phase: this CTA's global writes become visible to the other CTAs of the cluster — to their generic loads (release /
// acquire of the cluster barrier) and to their TMA loads (the asynchronous proxy: fence.proxy.async on both sides).
__device__ __forceinline__ void phase_end() {
  __threadfence();
  fence_proxy_async_all();
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  fence_proxy_async_all();
}

// ------------------------------------------------------------------------------------------------ GEMM phase
// D[rows of image b, this CTA's ncols columns] = A[rows, K] . W[w_row0 .. w_row0 + ncols, K]^T; `epi(row, chunk, v)` consumes 32
// columns of one row (row < N).  ncols in {32, 64, 96, 128} (0: this CTA idles), K = 64 * kblocks.
template <class Epi>
__device__ __forceinline__ void gemm_phase(Ctx& c, const DecArgs& a, const CUtensorMap* ta, const CUtensorMap* tw, int w_row0, int plane_rows,
                                           int ncols, int kblocks, Epi epi) {
  if (ncols > 0) {
    if (c.warp == TMA_WARP) {
      if (c.lane == 0) {
        for (int mt = 0; mt < a.MT; ++mt)
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&c.empty[c.pt.stage], c.pt.phase ^ 1, a.err, 301);
            const uint32_t st = ptx::smem_u32(c.ring + c.pt.stage * STAGE_BYTES);
            uint64_t* bar = &c.full[c.pt.stage];
            ptx::mbar_arrive_expect_tx(bar, 2 * GROUP_BYTES + 2 * ncols * 128);
            tma_load_4d(st, ta, bar, kb * 128, mt * 128, 0, c.b);
            tma_load_4d(st + GROUP_BYTES, ta, bar, kb * 128 + 64, mt * 128, 0, c.b);
            tma_load_2d_s(st + 2 * GROUP_BYTES, tw, bar, kb * BK, w_row0);
            tma_load_2d_s(st + 2 * GROUP_BYTES + ncols * 128, tw, bar, kb * BK, plane_rows + w_row0);
            c.pt.advance();
          }
      }
    } else if (c.warp == MMA_WARP) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, ncols);
      for (int mt = 0; mt < a.MT; ++mt) {
        const uint32_t d_tmem = c.tmem + mt * 128;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&c.full[c.pm.stage], c.pm.phase, a.err, 302);
          ptx::tc_fence_after();
          if (c.lane == 0) {
            const uint32_t a0 = ptx::smem_u32(c.ring + c.pm.stage * STAGE_BYTES);
            const uint32_t b_hi = a0 + 2 * GROUP_BYTES, b_lo = b_hi + ncols * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t at = a0 + (ks >> 1) * GROUP_BYTES + (ks & 1) * 32;
              const uint64_t dah = ptx::umma_desc_sw128(at), dal = ptx::umma_desc_sw128(at + 64);
              const uint64_t dbh = ptx::umma_desc_sw128(b_hi + ks * 32), dbl = ptx::umma_desc_sw128(b_lo + ks * 32);
              ptx::umma_bf16(d_tmem, dal, dbh, idesc, (kb != 0) || (ks != 0));  // small terms first
              ptx::umma_bf16(d_tmem, dah, dbl, idesc, 1);
              ptx::umma_bf16(d_tmem, dah, dbh, idesc, 1);
            }
            ptx::umma_commit(&c.empty[c.pm.stage]);
            if (kb == kblocks - 1) ptx::umma_commit(&c.acc_full[mt]);
          }
          __syncwarp();
          c.pm.advance();
        }
      }
    } else if (c.warp < 4) {
      for (int mt = 0; mt < a.MT; ++mt) {
        ptx::mbar_wait(&c.acc_full[mt], c.gp, a.err, 303);
        ptx::tc_fence_after();
        const int row = mt * 128 + c.warp * 32 + c.lane;
        const uint32_t t_addr = c.tmem + ((uint32_t)(c.warp * 32) << 16) + mt * 128;
        for (int ch = 0; ch < (ncols >> 5); ++ch) {
          uint32_t rr[32];
          ptx::tmem_ld_32x32(t_addr + ch * 32, rr);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          if (row < a.N) epi(row, ch, v);
        }
      }
    }
  }
  if (ncols > 0) c.gp ^= 1;  // an idling CTA's accumulator barriers did not complete a phase
  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();  // accumulators drained, operand ring idle: the next phase may reuse TMEM and shared memory
  ptx::tc_fence_after();
}

__device__ __forceinline__ void add_vec32(float (&v)[32], const float* p) {  // read-only data (weights-derived): the nc path is fine
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = __ldg((const float4*)p + j);
    v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
  }
}
__device__ __forceinline__ void store_row32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) *((float4*)dst + j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// ------------------------------------------------------------------------------------------------ row-wise phases
// rows of image b are dealt over (CTA, warp): row = 64 i + 8 r + warp; a lane owns channels 8*lane .. 8*lane + 7 (= 16 bytes of a
// P32 group's hi half and 16 of its lo half)
template <class F>
__device__ __forceinline__ void for_rows(const Ctx& c, int N, F f) {
  for (int row = 8 * c.r + c.warp; row < N; row += 64) f(row);
}
__device__ __forceinline__ void store_row_p32(uint8_t* base, long long grow, int lane, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    l[j] = pack_bf16x2(x[2 * j] - __uint_as_float(h[j] << 16), x[2 * j + 1] - __uint_as_float(h[j] & 0xffff0000u));
  }
  uint8_t* g = base + grow * 1024 + (lane >> 2) * 128 + (lane & 3) * 16;
  *(uint4*)g = make_uint4(h[0], h[1], h[2], h[3]);
  *(uint4*)(g + 64) = make_uint4(l[0], l[1], l[2], l[3]);
}

// out = LayerNorm(o + res) * gamma + beta over 256 channels (deformable_detr.py:1417, 1447, 1477); also the P32 copy the next
// GEMM streams and, for the layer output, the stacked intermediate state
__device__ __forceinline__ void ln_phase(const Ctx& c, const DecArgs& a, const float* res, const float* gamma, const float* beta, float* outf,
                                         uint8_t* outp, float* out2) {
  for_rows(c, a.N, [&](int row) {
    const long long grow = (long long)c.b * a.N + row;
    const float4* op = (const float4*)(a.o + grow * 256) + 2 * c.lane;
    const float4* rp = (const float4*)(res + grow * 256) + 2 * c.lane;
    const float4 o0 = __ldcg(op), o1 = __ldcg(op + 1), r0 = __ldcg(rp), r1 = __ldcg(rp + 1);
    float x[8] = {o0.x + r0.x, o0.y + r0.y, o0.z + r0.z, o0.w + r0.w, o1.x + r1.x, o1.y + r1.y, o1.z + r1.z, o1.w + r1.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    const float mean = warp_sum(s) * (1.f / 256.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] -= mean; q = fmaf(x[j], x[j], q); }
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / 256.f) + 1e-5f);
    const float4 g0 = __ldg((const float4*)gamma + 2 * c.lane), g1 = __ldg((const float4*)gamma + 2 * c.lane + 1);
    const float4 b0 = __ldg((const float4*)beta + 2 * c.lane), b1 = __ldg((const float4*)beta + 2 * c.lane + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = x[j] * rstd * gg[j] + bb[j];
    float4* of = (float4*)(outf + grow * 256) + 2 * c.lane;
    of[0] = make_float4(x[0], x[1], x[2], x[3]);
    of[1] = make_float4(x[4], x[5], x[6], x[7]);
    if (out2) {
      float4* o2 = (float4*)(out2 + (long long)row * 256) + 2 * c.lane;
      o2[0] = make_float4(x[0], x[1], x[2], x[3]);
      o2[1] = make_float4(x[4], x[5], x[6], x[7]);
    }
    store_row_p32(outp, grow, c.lane, x);
  });
}

// ------------------------------------------------------------------------------------------------ MSDA phase (head r)
// msda.cu's fused decoder form over fp16 pair records, 32 queries at a time: phase 1 = one thread per (query, sample): softmax over
// the 16 logits, sampling location, record indices of the top / bottom row + 4 corner weights; phase 2 = 8 lanes per query.
constexpr int SLOT_WORDS = 8, Q_STRIDE = 16 * SLOT_WORDS + 8;
__device__ __forceinline__ void msda_phase(const Ctx& c, const DecArgs& a, int layer, float* slots /* smem, 32 * Q_STRIDE floats */) {
  const int tid = threadIdx.x, m = c.r, b = c.b;
  const int s = tid & 15, l = s >> 2;
  const float* vr = a.valid_ratios + (long long)b * a.Lv * 2;
  const int H = a.lvH[l], W = a.lvW[l], S0 = a.lvS[l];
  for (int q0 = 0; q0 < a.N; q0 += 32) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int qi = (tid >> 4) + pass * 16;
      const int q = q0 + qi;
      float2 off = make_float2(0.f, 0.f);
      float logit = 0.f;
      if (q < a.N) {
        const float* row = a.offaw + ((long long)b * a.N + q) * 384;
        off = __ldcg((const float2*)(row + (m * 16 + s) * 2));
        logit = __ldcg(row + 256 + m * 16 + s);
      }
      float mx = logit;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = __expf(logit - mx);
      float sum = e;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float wgt = __fdividef(e, sum);
      int idx[2] = {0, 0};
      float cw[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < a.N) {
        const float2 rp = __ldg((const float2*)(a.ref + (long long)q * 2));
        const float lx = rp.x * vr[l * 2 + 0] + __fdividef(off.x, (float)W);
        const float ly = rp.y * vr[l * 2 + 1] + __fdividef(off.y, (float)H);
        const float him = ly * (float)H - 0.5f, wim = lx * (float)W - 0.5f;
        if (him > -1.f && wim > -1.f && him < (float)H && wim < (float)W) {
          const int hl = (int)floorf(him), wl = (int)floorf(wim);
          const float lh = him - (float)hl, lw = wim - (float)wl;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const bool y0 = hl >= 0, y1 = hl + 1 <= H - 1, x0 = wl >= 0, x1 = wl + 1 <= W - 1;
          const int rec = b * a.S + S0 + hl * W + wl + 1;
          if (y0) idx[0] = rec;
          if (y1) idx[1] = rec + W;
          if (y0 && x0) cw[0] = hh * hw * wgt;
          if (y0 && x1) cw[1] = hh * lw * wgt;
          if (y1 && x0) cw[2] = lh * hw * wgt;
          if (y1 && x1) cw[3] = lh * lw * wgt;
        }
      }
      float* slot = &slots[qi * Q_STRIDE + s * SLOT_WORDS];
      *(int2*)slot = make_int2(idx[0], idx[1]);
      *(float4*)(slot + 4) = make_float4(cw[0], cw[1], cw[2], cw[3]);
    }
    __syncthreads();
    const int g = tid >> 3, q = q0 + g;
    if (q < a.N) {
      const int half = (tid >> 2) & 1, j8 = (tid & 3) * 8;
      const uint8_t* hb = a.value_h16 + ((long long)(layer * CL + m) * a.records) * 128 + half * 64 + j8 * 2;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* myslots = &slots[g * Q_STRIDE];
#pragma unroll 8
      for (int ss = 0; ss < 16; ++ss) {
        const int2 id = *(const int2*)(myslots + ss * SLOT_WORDS);
        const float4 w = *(const float4*)(myslots + ss * SLOT_WORDS + 4);
        const float wt = half ? w.y : w.x, wb = half ? w.w : w.z;
        const uint4 t4 = __ldg((const uint4*)(hb + (unsigned long long)(uint32_t)id.x * 128ull));
        const uint4 b4 = __ldg((const uint4*)(hb + (unsigned long long)(uint32_t)id.y * 128ull));
        const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&tw[k]));
          const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&bw[k]));
          acc[2 * k] = fmaf(wt, tf.x, acc[2 * k]); acc[2 * k + 1] = fmaf(wt, tf.y, acc[2 * k + 1]);
          acc[2 * k] = fmaf(wb, bf.x, acc[2 * k]); acc[2 * k + 1] = fmaf(wb, bf.y, acc[2 * k + 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffu << (threadIdx.x & 24), acc[k], 4);  // left + right corners
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t hbits = pack_bf16x2(acc[2 * k], acc[2 * k + 1]);
        o[k] = half ? pack_bf16x2(acc[2 * k] - __uint_as_float(hbits << 16), acc[2 * k + 1] - __uint_as_float(hbits & 0xffff0000u)) : hbits;
      }
      *(uint4*)(a.attn2_p + ((long long)b * a.N + q) * 1024 + m * 128 + half * 64 + j8 * 2) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ MHA phase (head r)
// mha_bar: [0] operand tiles landed, [1..2] S tile of m-tile 0/1 complete, [3] P tile written (128 arrivals), [4..5] O complete
__device__ __forceinline__ void mha_phase(Ctx& c, const DecArgs& a, const DecMaps& maps) {
  uint8_t* sm = c.ring;
  const int kg = (a.N + 31) >> 5;  // key groups of 32
  if (c.warp == TMA_WARP) {
    if (c.lane == 0) {
      uint64_t* bar = &c.mha_bar[0];
      ptx::mbar_arrive_expect_tx(bar, (a.MT + 2) * GROUP_BYTES + kg * 4096);
      for (int mt = 0; mt < a.MT; ++mt) tma_load_4d(ptx::smem_u32(sm + MHA_Q + mt * GROUP_BYTES), &maps.a_qk, bar, c.r * 64, mt * 128, 0, c.b);
      for (int t = 0; t < 2; ++t) tma_load_4d(ptx::smem_u32(sm + MHA_K + t * GROUP_BYTES), &maps.a_qk, bar, (8 + c.r) * 64, t * 128, 0, c.b);
      for (int g = 0; g < kg; ++g) tma_load_2d_s(ptx::smem_u32(sm + MHA_VT + g * 4096), &maps.a_vt, bar, g * 64, c.b * 256 + c.r * 32);
    }
  } else if (c.warp == MMA_WARP) {
    ptx::mbar_wait(&c.mha_bar[0], c.mp, a.err, 311);
    ptx::tc_fence_after();
    if (c.lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 256);
      const uint32_t kt = ptx::smem_u32(sm + MHA_K);
      for (int mt = 0; mt < a.MT; ++mt) {
        const uint32_t qt = ptx::smem_u32(sm + MHA_Q + mt * GROUP_BYTES);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t dah = ptx::umma_desc_sw128(qt + ks * 32), dal = ptx::umma_desc_sw128(qt + ks * 32 + 64);
          const uint64_t dbh = ptx::umma_desc_sw128(kt + ks * 32), dbl = ptx::umma_desc_sw128(kt + ks * 32 + 64);
          ptx::umma_bf16(c.tmem + mt * 256, dal, dbh, idesc, ks != 0);
          ptx::umma_bf16(c.tmem + mt * 256, dah, dbl, idesc, 1);
          ptx::umma_bf16(c.tmem + mt * 256, dah, dbh, idesc, 1);
        }
        ptx::umma_commit(&c.mha_bar[1 + mt]);
      }
    }
    __syncwarp();
    for (int mt = 0; mt < a.MT; ++mt) {
      ptx::mbar_wait(&c.mha_bar[3], c.pp ^ (uint32_t)(mt & 1), a.err, 312);
      ptx::tc_fence_after();
      if (c.lane == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 32);
        for (int g = 0; g < kg; ++g) {
          const uint32_t pt = ptx::smem_u32(sm + MHA_P + g * GROUP_BYTES), vt = ptx::smem_u32(sm + MHA_VT + g * 4096);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t dah = ptx::umma_desc_sw128(pt + ks * 32), dal = ptx::umma_desc_sw128(pt + ks * 32 + 64);
            const uint64_t dbh = ptx::umma_desc_sw128(vt + ks * 32), dbl = ptx::umma_desc_sw128(vt + ks * 32 + 64);
            ptx::umma_bf16(c.tmem + mt * 256, dal, dbh, idesc, (g != 0) || (ks != 0));
            ptx::umma_bf16(c.tmem + mt * 256, dah, dbl, idesc, 1);
            ptx::umma_bf16(c.tmem + mt * 256, dah, dbh, idesc, 1);
          }
        }
        ptx::umma_commit(&c.mha_bar[4 + mt]);
      }
      __syncwarp();
    }
  } else if (c.warp < 4) {
    const int sw = c.lane & 7;
    for (int mt = 0; mt < a.MT; ++mt) {
      ptx::mbar_wait(&c.mha_bar[1 + mt], c.mp, a.err, 313);
      ptx::tc_fence_after();
      const int row = mt * 128 + c.warp * 32 + c.lane;
      const uint32_t t_addr = c.tmem + ((uint32_t)(c.warp * 32) << 16) + mt * 256;
      uint32_t rr[32];
      float mx = -INFINITY;
      for (int g = 0; g < kg; ++g) {  // pass 1: row maximum over the real keys
        ptx::tmem_ld_32x32(t_addr + g * 32, rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (g * 32 + j < a.N) mx = fmaxf(mx, __uint_as_float(rr[j]));
      }
      float sum = 0.f;
      const uint32_t prow = ptx::smem_u32(sm + MHA_P) + (c.warp * 32 + c.lane) * 128;
      for (int g = 0; g < kg; ++g) {  // pass 2: unnormalised probabilities as bf16 hi/lo, the PV product's A operand
        ptx::tmem_ld_32x32(t_addr + g * 32, rr);
        ptx::tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = (g * 32 + j < a.N) ? __expf(__uint_as_float(rr[j]) - mx) : 0.f;
          sum += p[j];
        }
        uint32_t o[32];
        split_group(p, o);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + g * GROUP_BYTES + ((ch ^ sw) << 4)), "r"(o[4 * ch]),
                       "r"(o[4 * ch + 1]), "r"(o[4 * ch + 2]), "r"(o[4 * ch + 3]) : "memory");
      }
      ptx::fence_proxy_async_smem();  // generic-proxy stores of P -> the tensor core's reads
      ptx::tc_fence_before();         // ... and this thread's TMEM reads of S before the PV product overwrites its first columns
      ptx::mbar_arrive(&c.mha_bar[3]);
      ptx::mbar_wait(&c.mha_bar[4 + mt], c.mp, a.err, 314);
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32(t_addr, rr);
      ptx::tmem_ld_wait();
      if (row < a.N) {
        const float inv = 1.f / sum;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) * inv;
        uint32_t o[32];
        split_group(v, o);
        store_group_global(a.attn_p + ((long long)c.b * a.N + row) * 1024 + c.r * 128, o);
      }
      // the next m-tile's softmax rewrites the P tile: its PV product has completed (mha_bar[4 + mt])
    }
  }
  c.mp ^= 1;
  c.pp ^= (uint32_t)(a.MT & 1);
  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();
  ptx::tc_fence_after();
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
decoder_kernel(const __grid_constant__ DecMaps maps, const DecArgs a) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ctrl = ring + (MHA_BYTES > RING_BYTES ? MHA_BYTES : RING_BYTES);
  Ctx c;
  c.ring = ring;
  c.full = (uint64_t*)ctrl;        // [STAGES]
  c.empty = c.full + 4;            // [STAGES]
  c.acc_full = c.empty + 4;        // [2]
  c.mha_bar = c.acc_full + 2;      // [6]
  uint32_t* tmem_holder = (uint32_t*)(c.mha_bar + 6);
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.r = blockIdx.x;  // == %cluster_ctarank: the cluster spans x
  c.b = blockIdx.y;
  c.pt = {0, 0};
  c.pm = {0, 0};
  c.gp = 0;
  c.mp = 0;
  c.pp = 0;
  if (c.warp == TMA_WARP && c.lane == 0) {
    const CUtensorMap* mp = &maps.a_h;
    for (int i = 0; i < (int)(sizeof(DecMaps) / sizeof(CUtensorMap)); ++i) ptx::prefetch_tensormap(mp + i);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&c.full[i], 1);
      ptx::mbar_init(&c.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) ptx::mbar_init(&c.acc_full[i], 1);
    for (int i = 0; i < 6; ++i) ptx::mbar_init(&c.mha_bar[i], i == 3 ? 128 : 1);
    ptx::fence_barrier_init();
  }
  if (c.warp == MMA_WARP) ptx::tmem_alloc<TMEM_COLS>(tmem_holder);
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  pdl_wait();
  c.tmem = *tmem_holder;

  const int N = a.N, r = c.r, b = c.b;
  for (int l = a.layer0; l < a.layer1; ++l) {
    const int p_lo = (l == a.layer0) ? a.phase0 : (int)PH_QKV, p_hi = (l == a.layer1 - 1) ? a.phase1 : (int)PH_END;
    const float* vec = a.vec + (long long)l * VEC_FLOATS;
    for (int ph = p_lo; ph < p_hi; ++ph) {
      if (ph == PH_INIT) {
        if (l == 0)
          for_rows(c, N, [&](int row) {  // h = the learned query embeddings (deformable_detr.py:2290-2292)
            const long long grow = (long long)b * N + row;
            const float4 x0 = __ldg((const float4*)(a.tgt + (long long)row * 256) + 2 * c.lane), x1 = __ldg((const float4*)(a.tgt + (long long)row * 256) + 2 * c.lane + 1);
            float4* of = (float4*)(a.hf + grow * 256) + 2 * c.lane;
            of[0] = x0;
            of[1] = x1;
            const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            store_row_p32(a.hp, grow, c.lane, x);
          });
      } else if (ph == PH_QKV) {
        // weight rows of layer l: head-major, per head q(32) | k(32) | v(32); columns land in the standard q | k | v row layout
        float* qkv_l = a.qkv + (long long)l * a.B * N * 768;
        const float* pos_l = a.qkv_pos + (long long)l * N * 768;
        gemm_phase(c, a, &maps.a_h, &maps.w_qkv, l * 768 + r * 96, a.plane_rows[0], 96, 4, [&](int row, int ch, float (&v)[32]) {
          const int gcol = ch * 256 + r * 32;
          add_vec32(v, pos_l + (long long)row * 768 + gcol);
          const long long grow = (long long)b * N + row;
          store_row32(qkv_l + grow * 768 + gcol, v);
          if (a.mha_mode) {
            if (ch < 2) {  // q_r, k_r as P32 groups r and 8 + r of the [B*N, 512] operand rows
              uint32_t o[32];
              split_group(v, o);
              store_group_global(a.qk_p + grow * 2048 + (ch * 8 + r) * 128, o);
            } else {       // v_r transposed: row (b, 32 r + d) of the [B*256, 256-key] P32 matrix, key = this query
              uint8_t* vt = a.vt_p + ((long long)b * 256 + r * 32) * 1024 + (row >> 5) * 128 + (row & 31) * 2;
#pragma unroll
              for (int d = 0; d < 32; ++d) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v[d]);
                const __nv_bfloat16 lo = __float2bfloat16_rn(v[d] - __bfloat162float(h));
                *(__nv_bfloat16*)(vt + d * 1024) = h;
                *(__nv_bfloat16*)(vt + d * 1024 + 64) = lo;
              }
            }
          }
        });
      } else if (ph == PH_MHA) {
        if (a.mha_mode) mha_phase(c, a, maps);
      } else if (ph == PH_OPROJ || ph == PH_OUTPROJ || ph == PH_FC2) {
        const CUtensorMap* ta = ph == PH_OPROJ ? &maps.a_attn : (ph == PH_OUTPROJ ? &maps.a_attn2 : &maps.a_f);
        const CUtensorMap* tw = ph == PH_OPROJ ? &maps.w_o : (ph == PH_OUTPROJ ? &maps.w_out : &maps.w_fc2);
        const int pr = a.plane_rows[ph == PH_OPROJ ? 1 : (ph == PH_OUTPROJ ? 3 : 5)];
        const float* bias = vec + (ph == PH_OPROJ ? 0 : (ph == PH_OUTPROJ ? 256 : 512)) + r * 32;
        gemm_phase(c, a, ta, tw, l * 256 + r * 32, pr, 32, ph == PH_FC2 ? 16 : 4, [&](int row, int ch, float (&v)[32]) {
          add_vec32(v, bias);
          store_row32(a.o + ((long long)b * N + row) * 256 + r * 32, v);
        });
      } else if (ph == PH_LN1) {
        ln_phase(c, a, a.hf, vec + 768, vec + 1024, a.t1f, a.t1p, nullptr);
      } else if (ph == PH_OFFAW) {
        const float* pos_l = a.off_pos + (long long)l * N * 384;
        gemm_phase(c, a, &maps.a_t1, &maps.w_offaw, l * 384 + r * 64, a.plane_rows[2], r < 6 ? 64 : 0, 4, [&](int row, int ch, float (&v)[32]) {
          const int gcol = r * 64 + ch * 32;
          add_vec32(v, pos_l + (long long)row * 384 + gcol);
          store_row32(a.offaw + ((long long)b * N + row) * 384 + gcol, v);
        });
      } else if (ph == PH_MSDA) {
        msda_phase(c, a, l, (float*)ring);
      } else if (ph == PH_LN2) {
        ln_phase(c, a, a.t1f, vec + 1280, vec + 1536, a.t2f, a.t2p, nullptr);
      } else if (ph == PH_FC1) {
        const float* bias = vec + 2304 + r * 128;
        gemm_phase(c, a, &maps.a_t2, &maps.w_fc1, l * 1024 + r * 128, a.plane_rows[4], 128, 4, [&](int row, int ch, float (&v)[32]) {
          add_vec32(v, bias + ch * 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          uint32_t o[32];
          split_group(v, o);
          store_group_global(a.f_p + ((long long)b * N + row) * 4096 + (r * 4 + ch) * 128, o);
        });
      } else if (ph == PH_LN3) {
        ln_phase(c, a, a.t2f, vec + 1792, vec + 2048, a.hf, a.hp, a.inter + ((long long)b * a.L + l) * N * 256);
      }
      phase_end();
    }
  }

  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();
  if (c.warp == MMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(c.tmem);
  }
}

// Fault code of a barrier wait that timed out (ptx::mbar_wait traps after ~2 s): host-mapped memory, so that it can still be read
// after the trap has killed the context (egtr_decoder_fault).
int* g_fault_host = nullptr;
int* device_error_flag_dec() {
  static int* dev = nullptr;
  if (!dev) {
    if (cudaHostAlloc((void**)&g_fault_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return nullptr;
    *g_fault_host = 0;
    if (cudaHostGetDevicePointer((void**)&dev, g_fault_host, 0) != cudaSuccess) return nullptr;
  }
  return dev;
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_decoder_fault() { return g_fault_host ? *(volatile int*)g_fault_host : 0; }

extern "C" long long egtr_decoder_scratch_bytes(int B, int N) {
  const long long R = (long long)B * N;
  // hf t1f t2f o (fp32 rows) + hp t1p t2p attn_p attn2_p (P32 rows) = 9 KB per row; offaw 1.5 KB; f_p 4 KB; qk_p 2 KB; + V^T 256 KB per image
  return R * (9 * 1024 + 1536 + 4096 + 2048) + (long long)B * 256 * 1024 + 1024;
}

extern "C" int egtr_decoder_fused_f32(const egtr_decoder_weights_t* w, void* scratch, const void* value_h16, long long records,
                                      const int* shapes_hw, int n_levels, const float* valid_ratios, int B, int S, float* qkv_out,
                                      float* inter, int layer0, int layer1, int phase0, int phase1, int mha_mode, egtr_stream_t s) {
  EGTR_ONE_DEVICE();
  EGTR_CHECK(w && scratch && value_h16 && shapes_hw && valid_ratios && qkv_out && inter, EGTR_ERR_ARG, "egtr_decoder_fused_f32: null pointer");
  const int N = w->n_queries, L = w->layers;
  EGTR_CHECK(N > 0 && N <= 256 && L > 0 && B > 0 && B <= 65535, EGTR_ERR_UNSUPPORTED, "egtr_decoder_fused_f32: built for up to 256 queries (N=%d)", N);
  EGTR_CHECK(n_levels == 4, EGTR_ERR_UNSUPPORTED, "egtr_decoder_fused_f32: 4 feature levels x 4 points (got %d levels)", n_levels);
  EGTR_CHECK(0 <= layer0 && layer0 < layer1 && layer1 <= L && 0 <= phase0 && phase0 < PH_END && 0 < phase1 && phase1 <= PH_END, EGTR_ERR_ARG,
             "egtr_decoder_fused_f32: layers [%d, %d) phases [%d, %d)", layer0, layer1, phase0, phase1);
  EGTR_CHECK(records == (long long)B * S + 1 && records < (1LL << 25), EGTR_ERR_ARG, "egtr_decoder_fused_f32: records must be B*S + 1");
  EGTR_CHECK(((uintptr_t)scratch & 1023) == 0 && ((uintptr_t)value_h16 & 127) == 0, EGTR_ERR_ARG, "egtr_decoder_fused_f32: scratch must be 1 KB aligned");
  DecArgs a = {};
  a.B = B; a.N = N; a.L = L; a.S = S; a.Lv = n_levels; a.MT = cdiv(N, 128);
  a.layer0 = layer0; a.layer1 = layer1; a.phase0 = phase0; a.phase1 = phase1; a.mha_mode = mha_mode;
  a.vec = w->vec; a.qkv_pos = w->qkv_pos; a.off_pos = w->off_pos; a.tgt = w->tgt; a.ref = w->ref_points; a.valid_ratios = valid_ratios;
  int start = 0;
  for (int l = 0; l < n_levels; ++l) {
    a.lvH[l] = shapes_hw[2 * l]; a.lvW[l] = shapes_hw[2 * l + 1]; a.lvS[l] = start;
    start += a.lvH[l] * a.lvW[l];
  }
  EGTR_CHECK(start == S, EGTR_ERR_ARG, "egtr_decoder_fused_f32: sum(H*W)=%d != S=%d", start, S);
  const long long R = (long long)B * N;
  uint8_t* p = (uint8_t*)scratch;
  auto take = [&](long long bytes) { uint8_t* q = p; p += bytes; return q; };
  a.hf = (float*)take(R * 1024); a.t1f = (float*)take(R * 1024); a.t2f = (float*)take(R * 1024); a.o = (float*)take(R * 1024);
  a.hp = take(R * 1024); a.t1p = take(R * 1024); a.t2p = take(R * 1024); a.attn_p = take(R * 1024); a.attn2_p = take(R * 1024);
  a.offaw = (float*)take(R * 1536);
  a.f_p = take(R * 4096);
  a.qk_p = take(R * 2048);
  p = (uint8_t*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
  a.vt_p = take((long long)B * 256 * 1024);
  a.qkv = qkv_out; a.inter = inter;
  a.value_h16 = (const uint8_t*)value_h16; a.records = records;
  a.err = device_error_flag_dec();
  const int nl[6] = {768, 256, 384, 256, 1024, 256};
  for (int i = 0; i < 6; ++i) a.plane_rows[i] = L * nl[i];
  DecMaps m;
  int rc;
  if ((rc = tmap_p32_rows(a.hp, 256, N, B, 128, &m.a_h)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.attn_p, 256, N, B, 128, &m.a_attn)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.t1p, 256, N, B, 128, &m.a_t1)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.attn2_p, 256, N, B, 128, &m.a_attn2)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.t2p, 256, N, B, 128, &m.a_t2)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.f_p, 1024, N, B, 128, &m.a_f)) != EGTR_OK) return rc;
  if ((rc = tmap_p32_rows(a.qk_p, 512, N, B, 128, &m.a_qk)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(a.vt_p, 512, (long long)B * 256, 32, &m.a_vt)) != EGTR_OK) return rc;  // [B*256 rows][256 keys] P32 = 512 bf16 per row
  if ((rc = tmap_weight_planes(w->w_qkv, 256, 2ll * L * 768, 96, &m.w_qkv)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(w->w_o, 256, 2ll * L * 256, 32, &m.w_o)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(w->w_offaw, 256, 2ll * L * 384, 64, &m.w_offaw)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(w->w_out, 256, 2ll * L * 256, 32, &m.w_out)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(w->w_fc1, 256, 2ll * L * 1024, 128, &m.w_fc1)) != EGTR_OK) return rc;
  if ((rc = tmap_weight_planes(w->w_fc2, 1024, 2ll * L * 256, 32, &m.w_fc2)) != EGTR_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    EGTR_CUDA(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  EGTR_CUDA(launch_pdl(decoder_kernel, dim3(CL, B), dim3(THREADS), (size_t)SMEM_BYTES, (cudaStream_t)s, m, a));
  count_launch();
  return EGTR_OK;
}
