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
//   OFFAW  head r's 32 sampling offsets | 16 attention logits of t1 (+ query_pos term)        N = 48 columns
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
constexpr int MAX_STAGES = 6;
// operand ring of a GEMM phase: a stage = activation k-block (32 KB) + weight hi/lo tiles of this CTA's ncols rows (256 B per row),
// as many stages as fit the region (5 at 32 / 48 columns, 4 at 96, 3 at 128): these phases are latency-bound, loads in flight count
// MHA phase (aliases the ring): Q (2 x 16 KB) | K (32 KB) | V^T (8 key groups x 4 KB) | P (8 key groups x 16 KB)
constexpr int MHA_Q = 0, MHA_K = 2 * GROUP_BYTES, MHA_VT = MHA_K + 2 * GROUP_BYTES, MHA_P = MHA_VT + 8 * 4096;
constexpr int MHA_BYTES = MHA_P + 8 * GROUP_BYTES;               // 224 KB
constexpr int CTRL_BYTES = 1024;
constexpr int RING_BYTES = MHA_BYTES;
constexpr int SMEM_BYTES = RING_BYTES + CTRL_BYTES + 1024 /*align*/;
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
  int dbg;  // egtr_set_debug_flags (dev experiments)
  unsigned long long* prof;  // dev: globaltimer at kernel start [0] and after every phase [1 + layer*12 + phase] (CTA 0)
};

struct Pipe {  // ring position; the stage count changes from phase to phase, so every stage's barrier keeps its own parity bit
  int stage;
  uint32_t bits;
  __device__ __forceinline__ uint32_t parity() const { return (bits >> stage) & 1u; }
  __device__ __forceinline__ void advance(int nstages) {
    bits ^= 1u << stage;
    if (++stage == nstages) stage = 0;
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
// `local`: the next phase reads only what THIS CTA wrote (head r's q / k / v, head r's offsets): a CTA barrier replaces the cluster's.
__device__ __forceinline__ void phase_end(bool local, int dbg) {
  if (!(dbg & 16)) __threadfence();
  if (dbg & 32) asm volatile("fence.proxy.async.global;" ::: "memory");
  else fence_proxy_async_all();
  ptx::tc_fence_before();
  __syncwarp();
  if (local) __syncthreads();
  else ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (!(dbg & 64)) {
    if (dbg & 32) asm volatile("fence.proxy.async.global;" ::: "memory");
    else fence_proxy_async_all();
  }
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The kernel's code is executed ONCE per layer, phase after phase: straight-line code that outgrows the SM's instruction cache
// runs at L2 fetch latency (the first version, 10.7 K instructions of inlined and unrolled phases, spent ~130 of its 145 us per
// layer on instruction fetch).  Hence: one out-of-line routine per phase kind, runtime epilogue kinds instead of template
// instantiations, rolled loops with explicit software prefetch — the layer loop's whole footprint stays cache-resident.
__device__ __noinline__ void mbar_wait_ni(uint64_t* bar, uint32_t parity, int* err, int code) { ptx::mbar_wait(bar, parity, err, code); }

// ------------------------------------------------------------------------------------------------ GEMM phase
// D[rows of image b, this CTA's ncols columns] = A[rows, K] . W[w_row0 .. w_row0 + ncols, K]^T, K = 64 * kblocks, ncols in
// {32, 48, 96, 128}.  Epilogue kinds (32-column chunks of one row):
enum GemmKind { GK_QKV = 0, GK_O, GK_OFFAW, GK_FC1 };
struct GemmJob {
  const CUtensorMap* ta;
  const CUtensorMap* tw;
  int w_row0, plane_rows, ncols, kblocks, kind, layer;
  const float* bias;  // GK_O: 32 bias values of this CTA's columns; GK_FC1: 128
};

__device__ __forceinline__ void gemm_phase(Ctx& c, const DecArgs& a, const GemmJob& j) {
  const int ncols = j.ncols, kblocks = j.kblocks;
  const int stage_bytes = 2 * GROUP_BYTES + ncols * 256;
  const int nstages = min(MAX_STAGES, RING_BYTES / stage_bytes);
  c.pt.stage = c.pm.stage = 0;
  if (c.warp == TMA_WARP) {
    if (c.lane == 0) {
#pragma unroll 1
      for (int i = 0; i < a.MT * kblocks; ++i) {
        const int mt = i / kblocks, kb = i - mt * kblocks;
        mbar_wait_ni(&c.empty[c.pt.stage], c.pt.parity() ^ 1, a.err, 301);
        const uint32_t st = ptx::smem_u32(c.ring + c.pt.stage * stage_bytes);
        uint64_t* bar = &c.full[c.pt.stage];
        ptx::mbar_arrive_expect_tx(bar, 2 * GROUP_BYTES + 2 * ncols * 128);
        tma_load_4d(st, j.ta, bar, kb * 128, mt * 128, 0, c.b);
        tma_load_4d(st + GROUP_BYTES, j.ta, bar, kb * 128 + 64, mt * 128, 0, c.b);
        tma_load_2d_s(st + 2 * GROUP_BYTES, j.tw, bar, kb * BK, j.w_row0);
        tma_load_2d_s(st + 2 * GROUP_BYTES + ncols * 128, j.tw, bar, kb * BK, j.plane_rows + j.w_row0);
        c.pt.advance(nstages);
      }
    }
  } else if (c.warp == MMA_WARP) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, ncols);
#pragma unroll 1
    for (int i = 0; i < a.MT * kblocks; ++i) {
      const int mt = i / kblocks, kb = i - mt * kblocks;
      mbar_wait_ni(&c.full[c.pm.stage], c.pm.parity(), a.err, 302);
      ptx::tc_fence_after();
      if (c.lane == 0) {
        const uint32_t d_tmem = c.tmem + mt * 128;
        const uint32_t a0 = ptx::smem_u32(c.ring + c.pm.stage * stage_bytes);
        const uint32_t b_hi = a0 + 2 * GROUP_BYTES, b_lo = b_hi + ncols * 128;
#pragma unroll 1
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
      c.pm.advance(nstages);
    }
  } else if (c.warp < 4) {
    const int N = a.N, r = c.r, l = j.layer, kind = j.kind;
#pragma unroll 1
    for (int mt = 0; mt < a.MT; ++mt) {
      mbar_wait_ni(&c.acc_full[mt], c.gp, a.err, 303);
      ptx::tc_fence_after();
      const int row = mt * 128 + c.warp * 32 + c.lane;
      const long long grow = (long long)c.b * N + row;
      const uint32_t t_addr = c.tmem + ((uint32_t)(c.warp * 32) << 16) + mt * 128;
#pragma unroll 1
      for (int ch = 0; ch < ((ncols + 31) >> 5); ++ch) {  // a 16-column tail chunk reads 16 junk columns: ignored below
        uint32_t rr[32];
        ptx::tmem_ld_32x32(t_addr + ch * 32, rr);
        ptx::tmem_ld_wait();
        if (row >= N) continue;
        // where this chunk goes: additive term (row bias or bias), fp32 destination, P32 destination, V^T destination
        const float* addp;
        float* dstf = nullptr;
        uint8_t* dstp = nullptr;
        uint8_t* dstvt = nullptr;
        int nvalid = 32;
        float floor_v = -INFINITY;
        if (kind == GK_QKV) {          // head-major weight rows q_r | k_r | v_r -> the standard q | k | v row layout
          const int gcol = ch * 256 + r * 32;
          addp = a.qkv_pos + ((long long)l * N + row) * 768 + gcol;
          dstf = a.qkv + ((long long)l * a.B * N + grow) * 768 + gcol;
          if (a.mha_mode) {
            if (ch < 2) dstp = a.qk_p + grow * 2048 + (ch * 8 + r) * 128;  // P32 groups r (q) and 8 + r (k) of [B*N, 512]
            else dstvt = a.vt_p + ((long long)c.b * 256 + r * 32) * 1024 + (row >> 5) * 128 + (row & 31) * 2;
          }
        } else if (kind == GK_O) {
          addp = j.bias;
          dstf = a.o + grow * 256 + r * 32;
        } else if (kind == GK_OFFAW) {  // head r: 32 sampling offsets | 16 attention logits -> [256 offsets | 128 logits] rows
          const int gcol = ch == 0 ? r * 32 : 256 + r * 16;
          nvalid = ch == 0 ? 32 : 16;
          addp = a.off_pos + ((long long)l * N + row) * 384 + gcol;
          dstf = a.offaw + grow * 384 + gcol;
        } else {                        // GK_FC1: ReLU, P32 rows
          addp = j.bias + ch * 32;
          floor_v = 0.f;
          dstp = a.f_p + grow * 4096 + (r * 4 + ch) * 128;
        }
        float v[32];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * q4 < nvalid) b4 = __ldg((const float4*)addp + q4);
          v[4 * q4] = fmaxf(__uint_as_float(rr[4 * q4]) + b4.x, floor_v);
          v[4 * q4 + 1] = fmaxf(__uint_as_float(rr[4 * q4 + 1]) + b4.y, floor_v);
          v[4 * q4 + 2] = fmaxf(__uint_as_float(rr[4 * q4 + 2]) + b4.z, floor_v);
          v[4 * q4 + 3] = fmaxf(__uint_as_float(rr[4 * q4 + 3]) + b4.w, floor_v);
        }
        if (dstf != nullptr) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            if (4 * q4 < nvalid) *((float4*)dstf + q4) = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
        }
        if (dstp != nullptr) {
          uint32_t o[32];
          split_group(v, o);
          store_group_global(dstp, o);
        }
        if (dstvt != nullptr) {  // v_r transposed: row (b, 32 r + d) of the [B*256, 256-key] P32 matrix, key = this query
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v[d]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v[d] - __bfloat162float(h));
            *(__nv_bfloat16*)(dstvt + d * 1024) = h;
            *(__nv_bfloat16*)(dstvt + d * 1024 + 64) = lo;
          }
        }
      }
    }
  }
  c.gp ^= 1;
  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();  // accumulators drained, operand ring idle: the next phase may reuse TMEM and shared memory
  ptx::tc_fence_after();
}

// ------------------------------------------------------------------------------------------------ row-wise phases
// rows of image b are dealt over (CTA, warp): row = 64 i + 8 r + warp; a lane owns channels 8*lane .. 8*lane + 7 (= 16 bytes of a
// P32 group's hi half and 16 of its lo half)
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
// GEMM streams and, for the layer output, the stacked intermediate state.  `o == nullptr`: out = res (the layer-0 input).
__device__ __forceinline__ void ln_phase(const Ctx& c, const DecArgs& a, const float* o, const float* res, long long res_bstride, const float* gamma,
                                      const float* beta, float* outf, uint8_t* outp, float* out2) {
  float4 g0, g1, b0, b1;
  if (o != nullptr) {
    g0 = __ldg((const float4*)gamma + 2 * c.lane); g1 = __ldg((const float4*)gamma + 2 * c.lane + 1);
    b0 = __ldg((const float4*)beta + 2 * c.lane); b1 = __ldg((const float4*)beta + 2 * c.lane + 1);
  }
  // software prefetch: the next row's loads are in flight while this row is reduced (one exposed L2 latency per phase)
  float4 no0, no1, nr0, nr1;
  auto load = [&](int row) {
    const long long grow = (long long)c.b * a.N + row;
    const float4* rp = (const float4*)(res + (long long)c.b * res_bstride + (long long)row * 256) + 2 * c.lane;
    nr0 = __ldcg(rp); nr1 = __ldcg(rp + 1);
    no0 = no1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o != nullptr) {
      const float4* op = (const float4*)(o + grow * 256) + 2 * c.lane;
      no0 = __ldcg(op); no1 = __ldcg(op + 1);
    }
  };
  int row = 8 * c.r + c.warp;
  if (row < a.N) load(row);
#pragma unroll 1
  for (; row < a.N; row += 64) {
    const long long grow = (long long)c.b * a.N + row;
    float x[8] = {no0.x + nr0.x, no0.y + nr0.y, no0.z + nr0.z, no0.w + nr0.w, no1.x + nr1.x, no1.y + nr1.y, no1.z + nr1.z, no1.w + nr1.w};
    if (row + 64 < a.N) load(row + 64);
    if (o != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += x[k];
      const float mean = warp_sum(s) * (1.f / 256.f);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { x[k] -= mean; q = fmaf(x[k], x[k], q); }
      const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / 256.f) + 1e-5f);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = x[k] * rstd * gg[k] + bb[k];
    }
    float4* of = (float4*)(outf + grow * 256) + 2 * c.lane;
    of[0] = make_float4(x[0], x[1], x[2], x[3]);
    of[1] = make_float4(x[4], x[5], x[6], x[7]);
    if (out2) {
      float4* o2 = (float4*)(out2 + (long long)row * 256) + 2 * c.lane;
      o2[0] = make_float4(x[0], x[1], x[2], x[3]);
      o2[1] = make_float4(x[4], x[5], x[6], x[7]);
    }
    store_row_p32(outp, grow, c.lane, x);
  }
}

// ------------------------------------------------------------------------------------------------ MSDA phase (head r)
// msda.cu's fused decoder form over fp16 pair records.  Phase 1, all queries: one thread per (query, sample): softmax over the 16
// logits, sampling location, record indices of the top / bottom row + 4 corner weights -> shared memory; phase 2: 8 lanes per query.
constexpr int SLOT_WORDS = 8, Q_STRIDE = 16 * SLOT_WORDS + 8;
__device__ __forceinline__ void msda_phase(const Ctx& c, const DecArgs& a, int layer, float* slots /* smem, N * Q_STRIDE floats */) {
  const int tid = threadIdx.x, m = c.r, b = c.b;
  const int s = tid & 15, l = s >> 2;
  const float* vr = a.valid_ratios + (long long)b * a.Lv * 2;
  const int H = a.lvH[l], W = a.lvW[l], S0 = a.lvS[l];
  const float vrx = vr[l * 2 + 0], vry = vr[l * 2 + 1];
  float2 off_n = make_float2(0.f, 0.f);
  float lg_n = 0.f;
  auto load = [&](int q) {
    off_n = make_float2(0.f, 0.f);
    lg_n = 0.f;
    if (q < a.N) {
      const float* row = a.offaw + ((long long)b * a.N + q) * 384;
      off_n = __ldcg((const float2*)(row + (m * 16 + s) * 2));
      lg_n = __ldcg(row + 256 + m * 16 + s);
    }
  };
  load(tid >> 4);
#pragma unroll 1
  for (int q = tid >> 4; q < ((a.N + 15) & ~15); q += 16) {  // warp-uniform trip count: the 16-lane shuffles stay convergent
    const float2 off = off_n;
    const float lg = lg_n;
    load(q + 16);
    float mx = lg;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = __expf(lg - mx);
    float sum = e;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float wgt = __fdividef(e, sum);
    if (q < a.N) {
      int idx0 = 0, idx1 = 0;
      float cw[4] = {0.f, 0.f, 0.f, 0.f};
      const float2 rp = __ldg((const float2*)(a.ref + (long long)q * 2));
      const float lx = rp.x * vrx + __fdividef(off.x, (float)W);
      const float ly = rp.y * vry + __fdividef(off.y, (float)H);
      const float him = ly * (float)H - 0.5f, wim = lx * (float)W - 0.5f;
      if (him > -1.f && wim > -1.f && him < (float)H && wim < (float)W) {
        const int hl = (int)floorf(him), wl = (int)floorf(wim);
        const float lh = him - (float)hl, lw = wim - (float)wl;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const bool y0 = hl >= 0, y1 = hl + 1 <= H - 1, x0 = wl >= 0, x1 = wl + 1 <= W - 1;
        const int rec = b * a.S + S0 + hl * W + wl + 1;
        if (y0) idx0 = rec;
        if (y1) idx1 = rec + W;
        if (y0 && x0) cw[0] = hh * hw * wgt;
        if (y0 && x1) cw[1] = hh * lw * wgt;
        if (y1 && x0) cw[2] = lh * hw * wgt;
        if (y1 && x1) cw[3] = lh * lw * wgt;
      }
      float* slot = &slots[q * Q_STRIDE + s * SLOT_WORDS];
      *(int2*)slot = make_int2(idx0, idx1);
      *(float4*)(slot + 4) = make_float4(cw[0], cw[1], cw[2], cw[3]);
    }
  }
  __syncthreads();
  // ---- phase 2: 8 lanes per query (lanes 0-3 the left corners, 4-7 the right ones of the same 128-byte pair record)
  const int half = (tid >> 2) & 1, j8 = (tid & 3) * 8;
  const uint8_t* hb = a.value_h16 + ((long long)(layer * CL + m) * a.records) * 128 + half * 64 + j8 * 2;
#pragma unroll 1
  for (int q = tid >> 3; q < a.N; q += THREADS / 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* myslots = &slots[q * Q_STRIDE];
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
#pragma unroll 1
      for (int g = 0; g < kg; ++g) tma_load_2d_s(ptx::smem_u32(sm + MHA_VT + g * 4096), &maps.a_vt, bar, g * 64, c.b * 256 + c.r * 32);
    }
  } else if (c.warp == MMA_WARP) {
    mbar_wait_ni(&c.mha_bar[0], c.mp, a.err, 311);
    ptx::tc_fence_after();
    if (c.lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 256);
      const uint32_t kt = ptx::smem_u32(sm + MHA_K);
#pragma unroll 1
      for (int i = 0; i < 2 * a.MT; ++i) {
        const int mt = i >> 1, ks = i & 1;
        const uint32_t qt = ptx::smem_u32(sm + MHA_Q + mt * GROUP_BYTES);
        const uint64_t dah = ptx::umma_desc_sw128(qt + ks * 32), dal = ptx::umma_desc_sw128(qt + ks * 32 + 64);
        const uint64_t dbh = ptx::umma_desc_sw128(kt + ks * 32), dbl = ptx::umma_desc_sw128(kt + ks * 32 + 64);
        ptx::umma_bf16(c.tmem + mt * 256, dal, dbh, idesc, ks != 0);
        ptx::umma_bf16(c.tmem + mt * 256, dah, dbl, idesc, 1);
        ptx::umma_bf16(c.tmem + mt * 256, dah, dbh, idesc, 1);
        if (ks == 1) ptx::umma_commit(&c.mha_bar[1 + mt]);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int mt = 0; mt < a.MT; ++mt) {
      mbar_wait_ni(&c.mha_bar[3], c.pp ^ (uint32_t)(mt & 1), a.err, 312);
      ptx::tc_fence_after();
      if (c.lane == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 32);
#pragma unroll 1
        for (int i = 0; i < 2 * kg; ++i) {
          const int g = i >> 1, ks = i & 1;
          const uint32_t pt = ptx::smem_u32(sm + MHA_P + g * GROUP_BYTES), vt = ptx::smem_u32(sm + MHA_VT + g * 4096);
          const uint64_t dah = ptx::umma_desc_sw128(pt + ks * 32), dal = ptx::umma_desc_sw128(pt + ks * 32 + 64);
          const uint64_t dbh = ptx::umma_desc_sw128(vt + ks * 32), dbl = ptx::umma_desc_sw128(vt + ks * 32 + 64);
          ptx::umma_bf16(c.tmem + mt * 256, dal, dbh, idesc, i != 0);
          ptx::umma_bf16(c.tmem + mt * 256, dah, dbl, idesc, 1);
          ptx::umma_bf16(c.tmem + mt * 256, dah, dbh, idesc, 1);
        }
        ptx::umma_commit(&c.mha_bar[4 + mt]);
      }
      __syncwarp();
    }
  } else if (c.warp < 4) {
    const int sw = c.lane & 7;
#pragma unroll 1
    for (int mt = 0; mt < a.MT; ++mt) {
      mbar_wait_ni(&c.mha_bar[1 + mt], c.mp, a.err, 313);
      ptx::tc_fence_after();
      const int row = mt * 128 + c.warp * 32 + c.lane;
      const uint32_t t_addr = c.tmem + ((uint32_t)(c.warp * 32) << 16) + mt * 256;
      uint32_t rr[32];
      float mx = -INFINITY;
#pragma unroll 1
      for (int g = 0; g < kg; ++g) {  // pass 1: row maximum over the real keys
        ptx::tmem_ld_32x32(t_addr + g * 32, rr);
        ptx::tmem_ld_wait();
        const int nk = a.N - g * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, j < nk ? __uint_as_float(rr[j]) : -INFINITY);
      }
      float sum = 0.f;
      const uint32_t prow = ptx::smem_u32(sm + MHA_P) + (c.warp * 32 + c.lane) * 128;
#pragma unroll 1
      for (int g = 0; g < kg; ++g) {  // pass 2: unnormalised probabilities as bf16 hi/lo, the PV product's A operand
        ptx::tmem_ld_32x32(t_addr + g * 32, rr);
        ptx::tmem_ld_wait();
        const int nk = a.N - g * 32;
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = j < nk ? __expf(__uint_as_float(rr[j]) - mx) : 0.f;
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
      mbar_wait_ni(&c.mha_bar[4 + mt], c.mp, a.err, 314);
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
decoder_kernel(const __grid_constant__ DecMaps maps, const __grid_constant__ DecArgs a) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ctrl = ring + RING_BYTES;
  Ctx c;
  c.ring = ring;
  c.full = (uint64_t*)ctrl;        // [MAX_STAGES]
  c.empty = c.full + 8;            // [MAX_STAGES]
  c.acc_full = c.empty + 8;        // [2]
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
    for (int i = 0; i < MAX_STAGES; ++i) {
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
  if (a.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    a.prof[0] = gtime();
    a.prof[1 + a.L * PH_END] = clock64();
  }
  if (a.dbg & (128 | 256)) {  // dev: the bare cost of 64 phase ends (cluster / CTA-local), stamped into the slots after the clocks
    for (int i = 0; i < 64; ++i) phase_end((a.dbg & 256) != 0, a.dbg);
    if (a.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.prof[3 + a.L * PH_END] = gtime();
  }

  const int N = a.N, r = c.r, b = c.b;
#pragma unroll 1
  for (int l = a.layer0; l < a.layer1; ++l) {
    const int p_lo = (l == a.layer0) ? a.phase0 : (int)PH_QKV, p_hi = (l == a.layer1 - 1) ? a.phase1 : (int)PH_END;
    const float* vec = a.vec + (long long)l * VEC_FLOATS;
#pragma unroll 1
    for (int ph = p_lo; ph < p_hi; ++ph) {
      GemmJob j;
      j.layer = l;
      j.kblocks = 4;
      j.bias = nullptr;
      bool gemm = true;
      if (ph == PH_QKV) {
        j.ta = &maps.a_h; j.tw = &maps.w_qkv; j.w_row0 = l * 768 + r * 96; j.plane_rows = a.plane_rows[0]; j.ncols = 96; j.kind = GK_QKV;
      } else if (ph == PH_OPROJ) {
        j.ta = &maps.a_attn; j.tw = &maps.w_o; j.w_row0 = l * 256 + r * 32; j.plane_rows = a.plane_rows[1]; j.ncols = 32; j.kind = GK_O;
        j.bias = vec + r * 32;
      } else if (ph == PH_OFFAW) {
        j.ta = &maps.a_t1; j.tw = &maps.w_offaw; j.w_row0 = l * 384 + r * 48; j.plane_rows = a.plane_rows[2]; j.ncols = 48; j.kind = GK_OFFAW;
      } else if (ph == PH_OUTPROJ) {
        j.ta = &maps.a_attn2; j.tw = &maps.w_out; j.w_row0 = l * 256 + r * 32; j.plane_rows = a.plane_rows[3]; j.ncols = 32; j.kind = GK_O;
        j.bias = vec + 256 + r * 32;
      } else if (ph == PH_FC1) {
        j.ta = &maps.a_t2; j.tw = &maps.w_fc1; j.w_row0 = l * 1024 + r * 128; j.plane_rows = a.plane_rows[4]; j.ncols = 128; j.kind = GK_FC1;
        j.bias = vec + 2304 + r * 128;
      } else if (ph == PH_FC2) {
        j.ta = &maps.a_f; j.tw = &maps.w_fc2; j.w_row0 = l * 256 + r * 32; j.plane_rows = a.plane_rows[5]; j.ncols = 32; j.kind = GK_O;
        j.kblocks = 16;
        j.bias = vec + 512 + r * 32;
      } else {
        gemm = false;
      }
      if (gemm) {
        gemm_phase(c, a, j);
      } else if (ph == PH_INIT) {  // h = the learned query embeddings (deformable_detr.py:2290-2292)
        if (l == 0) ln_phase(c, a, nullptr, a.tgt, 0, nullptr, nullptr, a.hf, a.hp, nullptr);
      } else if (ph == PH_MHA) {
        if (a.mha_mode) mha_phase(c, a, maps);
      } else if (ph == PH_MSDA) {
        msda_phase(c, a, l, (float*)ring);
      } else {  // the three LayerNorms: residual, affine pair, outputs
        const long long nb = (long long)N * 256;
        const float* res = ph == PH_LN1 ? a.hf : (ph == PH_LN2 ? a.t1f : a.t2f);
        const float* gam = vec + (ph == PH_LN1 ? 768 : (ph == PH_LN2 ? 1280 : 1792));
        float* outf = ph == PH_LN1 ? a.t1f : (ph == PH_LN2 ? a.t2f : a.hf);
        uint8_t* outp = ph == PH_LN1 ? a.t1p : (ph == PH_LN2 ? a.t2p : a.hp);
        ln_phase(c, a, a.o, res, nb, gam, gam + 256, outf, outp, ph == PH_LN3 ? a.inter + ((long long)b * a.L + l) * N * 256 : nullptr);
      }
      phase_end((ph == PH_QKV && a.mha_mode) || ph == PH_OFFAW, a.dbg);
      if (a.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.prof[1 + l * PH_END + ph] = gtime();
    }
  }

  if (a.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.prof[2 + a.L * PH_END] = clock64();
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

unsigned long long* g_dec_prof = nullptr;
extern "C" int egtr_decoder_debug_profile(unsigned long long* dev_buf) {  // dev: [1 + layers*12] globaltimer stamps, NULL = off
  g_dec_prof = dev_buf;
  return EGTR_OK;
}

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
  a.prof = g_dec_prof;
  a.dbg = debug_flags();
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
  if ((rc = tmap_weight_planes(w->w_offaw, 256, 2ll * L * 384, 48, &m.w_offaw)) != EGTR_OK) return rc;
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
