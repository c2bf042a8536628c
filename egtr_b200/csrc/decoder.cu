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
  int zsplit;  // 1: cluster of 8 (CTA r = head / column slice r); 2: cluster of 16, CTA rank = 8 z + r owns m-tile z
  int plane_rows[6];  // rows of one bf16 plane per weight tensor (qkv, o, offaw, out, fc1, fc2)
  const float* vec;
  const float* qkv_pos;
  const float* off_pos;
  const float* tgt;
  const float* ref;
  const float* valid_ratios;
  float *hf, *t1f, *t2f, *o, *offaw, *part;
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
  int mt_lo, nmt;          // this CTA's m-tiles [mt_lo, mt_lo + nmt): all of them, or one when a 16-CTA cluster splits the rows
  int row_lo, row_hi;      // ... = its rows for the row-wise phases
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

// ------------------------------------------------------------------------------------------------ warp row transposition
// The TMEM read hands every lane one ROW of a 32-column chunk (128 contiguous bytes).  Written (or read) straight from there, each
// vector instruction touches 32 different cache lines and the load/store unit serialises them (the first version spent ~7 of the
// QKV phase's 18 us there).  A 4 KB per-warp staging tile (chunk index XOR row, conflict-free both ways) turns that into
// instructions that cover 4 rows x 128 bytes.
constexpr int STG_BYTES = 4 * 4096;  // four epilogue warps; the top of the ring region (a GEMM phase's stages stay below it)
__device__ __forceinline__ void warp_store_rows(uint32_t stg, int lane, const uint32_t (&o)[32], uint8_t* base, long long stride, int nrows, int nbytes) {
  const uint32_t mine = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + ((ch ^ sw) << 4)), "r"(o[4 * ch]), "r"(o[4 * ch + 1]), "r"(o[4 * ch + 2]),
                 "r"(o[4 * ch + 3]) : "memory");
  __syncwarp();
  const int c16 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + row * 128 + ((c16 ^ (row & 7)) << 4)) : "memory");
    if (row < nrows && c16 * 16 < nbytes) *(uint4*)(base + row * stride + c16 * 16) = v;
  }
  __syncwarp();
}
// v[32] += the lane's row of a [rows x nbytes] fp32 block read the same way (weights-derived data: the nc path is fine)
__device__ __forceinline__ void warp_add_rows(uint32_t stg, int lane, float (&v)[32], const uint8_t* base, long long stride, int nrows, int nbytes) {
  const int c16 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (row < nrows && c16 * 16 < nbytes) x = __ldg((const uint4*)(base + row * stride + c16 * 16));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + row * 128 + ((c16 ^ (row & 7)) << 4)), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
  }
  __syncwarp();
  const uint32_t mine = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 x;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(mine + ((ch ^ sw) << 4)) : "memory");
    v[4 * ch] += __uint_as_float(x.x); v[4 * ch + 1] += __uint_as_float(x.y); v[4 * ch + 2] += __uint_as_float(x.z); v[4 * ch + 3] += __uint_as_float(x.w);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ GEMM phase
// D[this CTA's rows of image b, its ncols columns] = A[rows, a_col0 + K] . W[w_row0 .. w_row0 + ncols, w_col0 + K]^T, K = 64 * kblocks,
// ncols in {32, 48, 96, 128, 256}.  Epilogue kinds (32-column chunks of 32 rows per warp):
//   GK_QKV   + row bias -> fp32 q|k|v rows (captured), q / k also as P32 operand rows, v transposed (the PV product's B operand)
//   GK_O     + bias -> fp32 `o` columns (out_proj / output_proj; the LayerNorm phase follows)
//   GK_OFFAW + row bias -> fp32 offsets | logits
//   GK_FC1   + bias, ReLU -> P32 rows of this CTA's 128 hidden columns
//   GK_PART  raw partial sums of fc2 over this CTA's 128 hidden columns (split-K: the A operand is what THIS CTA wrote in FC1 —
//            no cluster barrier, 1/4 of the operand bytes, 24 MMAs of N = 256 instead of 192 of N = 32); LN3 adds the eight planes
enum GemmKind { GK_QKV = 0, GK_O, GK_OFFAW, GK_FC1, GK_PART };
struct GemmJob {
  const CUtensorMap* ta;
  const CUtensorMap* tw;
  int a_col0, w_row0, w_col0, plane_rows, ncols, kblocks, kind, layer;
  const float* bias;  // GK_O: 32 bias values of this CTA's columns; GK_FC1: 128
};

__device__ __forceinline__ void gemm_phase(Ctx& c, const DecArgs& a, const GemmJob& j) {
  const int ncols = j.ncols, kblocks = j.kblocks;
  const int stage_bytes = 2 * GROUP_BYTES + ncols * 256;
  const int nstages = min(MAX_STAGES, (RING_BYTES - STG_BYTES) / stage_bytes);
  const int acc_stride = ncols > 128 ? 256 : 128;  // TMEM columns between the accumulators of two m-tiles
  c.pt.stage = c.pm.stage = 0;
  if (c.warp == TMA_WARP) {
    if (c.lane == 0) {
#pragma unroll 1
      for (int i = 0; i < c.nmt * kblocks; ++i) {
        const int lm = i / kblocks, kb = i - lm * kblocks, mt = c.mt_lo + lm;
        mbar_wait_ni(&c.empty[c.pt.stage], c.pt.parity() ^ 1, a.err, 301);
        const uint32_t st = ptx::smem_u32(c.ring + c.pt.stage * stage_bytes);
        uint64_t* bar = &c.full[c.pt.stage];
        ptx::mbar_arrive_expect_tx(bar, 2 * GROUP_BYTES + 2 * ncols * 128);
        tma_load_4d(st, j.ta, bar, 2 * j.a_col0 + kb * 128, mt * 128, 0, c.b);
        tma_load_4d(st + GROUP_BYTES, j.ta, bar, 2 * j.a_col0 + kb * 128 + 64, mt * 128, 0, c.b);
        tma_load_2d_s(st + 2 * GROUP_BYTES, j.tw, bar, j.w_col0 + kb * BK, j.w_row0);
        tma_load_2d_s(st + 2 * GROUP_BYTES + ncols * 128, j.tw, bar, j.w_col0 + kb * BK, j.plane_rows + j.w_row0);
        c.pt.advance(nstages);
      }
    }
  } else if (c.warp == MMA_WARP) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, ncols);
#pragma unroll 1
    for (int i = 0; i < c.nmt * kblocks; ++i) {
      const int lm = i / kblocks, kb = i - lm * kblocks;
      mbar_wait_ni(&c.full[c.pm.stage], c.pm.parity(), a.err, 302);
      ptx::tc_fence_after();
      if (c.lane == 0) {
        const uint32_t d_tmem = c.tmem + lm * acc_stride;
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
        if (kb == kblocks - 1) ptx::umma_commit(&c.acc_full[lm]);
      }
      __syncwarp();
      c.pm.advance(nstages);
    }
  } else if (c.warp < 4) {
    const int N = a.N, r = c.r, l = j.layer, kind = j.kind;
    const uint32_t stg = ptx::smem_u32(c.ring + RING_BYTES - STG_BYTES + c.warp * 4096);
#pragma unroll 1
    for (int lm = 0; lm < c.nmt; ++lm) {
      mbar_wait_ni(&c.acc_full[lm], c.gp, a.err, 303);
      ptx::tc_fence_after();
      const int row0 = (c.mt_lo + lm) * 128 + c.warp * 32;  // this warp's first row; lane = row0 + lane
      const long long grow0 = (long long)c.b * N + row0;
      const int nrows = max(0, min(32, N - row0));
      const uint32_t t_addr = c.tmem + ((uint32_t)(c.warp * 32) << 16) + lm * acc_stride;
#pragma unroll 1
      for (int ch = 0; ch < ((ncols + 31) >> 5); ++ch) {  // a 16-column tail chunk reads 16 junk columns: ignored below
        uint32_t rr[32];
        ptx::tmem_ld_32x32(t_addr + ch * 32, rr);
        ptx::tmem_ld_wait();
        if (nrows == 0) continue;  // warp-uniform
        // where this chunk goes: additive term (row bias block or bias vector), fp32 / P32 / V^T destinations of the warp's rows
        const uint8_t* addrows = nullptr;
        const float* addvec = nullptr;
        uint8_t *dstf = nullptr, *dstp = nullptr, *dstvt = nullptr;
        long long add_stride = 0, f_stride = 0, p_stride = 0;
        int nbytes = 128;
        float floor_v = -INFINITY;
        if (kind == GK_QKV) {          // head-major weight rows q_r | k_r | v_r -> the standard q | k | v row layout
          const int gcol = ch * 256 + r * 32;
          addrows = (const uint8_t*)(a.qkv_pos + ((long long)l * N + row0) * 768 + gcol); add_stride = 3072;
          dstf = (uint8_t*)(a.qkv + ((long long)l * a.B * N + grow0) * 768 + gcol); f_stride = 3072;
          if (a.mha_mode) {
            if (ch < 2) { dstp = a.qk_p + grow0 * 2048 + (ch * 8 + r) * 128; p_stride = 2048; }  // P32 groups r (q), 8 + r (k) of [B*N, 512]
            else dstvt = a.vt_p + ((long long)c.b * 256 + r * 32) * 1024 + ((row0 + c.lane) >> 5) * 128 + ((row0 + c.lane) & 31) * 2;
          }
        } else if (kind == GK_O) {
          addvec = j.bias;
          dstf = (uint8_t*)(a.o + grow0 * 256 + r * 32); f_stride = 1024;
        } else if (kind == GK_OFFAW) {  // head r: 32 sampling offsets | 16 attention logits -> [256 offsets | 128 logits] rows
          const int gcol = ch == 0 ? r * 32 : 256 + r * 16;
          nbytes = ch == 0 ? 128 : 64;
          addrows = (const uint8_t*)(a.off_pos + ((long long)l * N + row0) * 384 + gcol); add_stride = 1536;
          dstf = (uint8_t*)(a.offaw + grow0 * 384 + gcol); f_stride = 1536;
        } else if (kind == GK_FC1) {    // ReLU, P32 rows
          addvec = j.bias + ch * 32;
          floor_v = 0.f;
          dstp = a.f_p + grow0 * 4096 + (r * 4 + ch) * 128; p_stride = 4096;
        } else {                        // GK_PART: plane r of the fc2 partial sums
          dstf = (uint8_t*)(a.part + ((long long)r * a.B * N + grow0) * 256 + ch * 32); f_stride = 1024;
        }
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(rr[k]);
        if (addrows != nullptr) warp_add_rows(stg, c.lane, v, addrows, add_stride, nrows, nbytes);
        if (addvec != nullptr) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = __ldg((const float4*)addvec + q4);
            v[4 * q4] += b4.x; v[4 * q4 + 1] += b4.y; v[4 * q4 + 2] += b4.z; v[4 * q4 + 3] += b4.w;
          }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], floor_v);
        if (dstf != nullptr) {
#pragma unroll
          for (int k = 0; k < 32; ++k) rr[k] = __float_as_uint(v[k]);
          warp_store_rows(stg, c.lane, rr, dstf, f_stride, nrows, nbytes);
        }
        if (dstp != nullptr) {
          split_group(v, rr);
          warp_store_rows(stg, c.lane, rr, dstp, p_stride, nrows, 128);
        }
        if (dstvt != nullptr && c.lane < nrows) {  // v_r transposed: row (b, 32 r + d) of the [B*256, 256-key] P32 matrix, key = this query
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
// `nparts` > 1: `o` holds that many partial-sum planes [nparts][B*N][256] (fc2's split-K) and `obias` their bias.
__device__ __forceinline__ void ln_phase(const Ctx& c, const DecArgs& a, const float* o, int nparts, const float* obias, const float* res,
                                      long long res_bstride, const float* gamma, const float* beta, float* outf, uint8_t* outp, float* out2) {
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
      if (obias != nullptr) { no0 = __ldg((const float4*)obias + 2 * c.lane); no1 = __ldg((const float4*)obias + 2 * c.lane + 1); }
#pragma unroll 1
      for (int p = 0; p < nparts; ++p) {  // fixed order: bit-identical from run to run
        const float4* op = (const float4*)(o + ((long long)p * a.B * a.N + grow) * 256) + 2 * c.lane;
        const float4 x0 = __ldcg(op), x1 = __ldcg(op + 1);
        no0.x += x0.x; no0.y += x0.y; no0.z += x0.z; no0.w += x0.w;
        no1.x += x1.x; no1.y += x1.y; no1.z += x1.z; no1.w += x1.w;
      }
    }
  };
  int row = c.row_lo + 8 * c.r + c.warp;
  if (row < c.row_hi) load(row);
#pragma unroll 1
  for (; row < c.row_hi; row += 64) {
    const long long grow = (long long)c.b * a.N + row;
    float x[8] = {no0.x + nr0.x, no0.y + nr0.y, no0.z + nr0.z, no0.w + nr0.w, no1.x + nr1.x, no1.y + nr1.y, no1.z + nr1.z, no1.w + nr1.w};
    if (row + 64 < c.row_hi) load(row + 64);
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
    if (q < c.row_hi) {
      const float* row = a.offaw + ((long long)b * a.N + q) * 384;
      off_n = __ldcg((const float2*)(row + (m * 16 + s) * 2));
      lg_n = __ldcg(row + 256 + m * 16 + s);
    }
  };
  load(c.row_lo + (tid >> 4));
#pragma unroll 1
  for (int q = c.row_lo + (tid >> 4); q < c.row_lo + ((c.row_hi - c.row_lo + 15) & ~15); q += 16) {  // warp-uniform trip count: the 16-lane shuffles stay convergent
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
    if (q < c.row_hi) {
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
      float* slot = &slots[(q - c.row_lo) * Q_STRIDE + s * SLOT_WORDS];
      *(int2*)slot = make_int2(idx0, idx1);
      *(float4*)(slot + 4) = make_float4(cw[0], cw[1], cw[2], cw[3]);
    }
  }
  __syncthreads();
  // ---- phase 2: 8 lanes per query (lanes 0-3 the left corners, 4-7 the right ones of the same 128-byte pair record)
  const int half = (tid >> 2) & 1, j8 = (tid & 3) * 8;
  const uint8_t* hb = a.value_h16 + ((long long)(layer * CL + m) * a.records) * 128 + half * 64 + j8 * 2;
#pragma unroll 1
  for (int q = c.row_lo + (tid >> 3); q < c.row_hi; q += THREADS / 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* myslots = &slots[(q - c.row_lo) * Q_STRIDE];
#pragma unroll 16
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
      ptx::mbar_arrive_expect_tx(bar, (c.nmt + 2) * GROUP_BYTES + kg * 4096);
      for (int lm = 0; lm < c.nmt; ++lm) tma_load_4d(ptx::smem_u32(sm + MHA_Q + lm * GROUP_BYTES), &maps.a_qk, bar, c.r * 64, (c.mt_lo + lm) * 128, 0, c.b);
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
      for (int i = 0; i < 2 * c.nmt; ++i) {
        const int mt = i >> 1, ks = i & 1;  // local m-tile
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
    for (int mt = 0; mt < c.nmt; ++mt) {
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
    for (int mt = 0; mt < c.nmt; ++mt) {
      mbar_wait_ni(&c.mha_bar[1 + mt], c.mp, a.err, 313);
      ptx::tc_fence_after();
      const int row = (c.mt_lo + mt) * 128 + c.warp * 32 + c.lane;
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
      {
        const int row0 = row - c.lane, nrows = max(0, min(32, a.N - row0));
        const float inv = 1.f / sum;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) * inv;
        split_group(v, rr);
        // (the P tile's last group may overlap the staging tile: the PV product that read it has completed)
        if (nrows > 0)
          warp_store_rows(ptx::smem_u32(sm + RING_BYTES - STG_BYTES + c.warp * 4096), c.lane, rr, a.attn_p + ((long long)c.b * a.N + row0) * 1024 + c.r * 128,
                          1024, nrows, 128);
      }
      // the next m-tile's softmax rewrites the P tile: its PV product has completed (mha_bar[4 + mt])
    }
  }
  c.mp ^= 1;
  c.pp ^= (uint32_t)(c.nmt & 1);
  ptx::tc_fence_before();
  __syncwarp();
  __syncthreads();
  ptx::tc_fence_after();
}

__global__ void __launch_bounds__(THREADS, 1)
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
  c.r = blockIdx.x & (CL - 1);  // blockIdx.x == %cluster_ctarank: the cluster spans x
  c.b = blockIdx.y;
  c.mt_lo = a.zsplit == 2 ? (int)(blockIdx.x >> 3) : 0;
  c.nmt = a.zsplit == 2 ? 1 : a.MT;
  c.row_lo = c.mt_lo * 128;
  c.row_hi = min(a.N, (c.mt_lo + c.nmt) * 128);
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
      j.a_col0 = j.w_col0 = 0;
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
        // split-K: this CTA's 128 hidden columns (written by itself in FC1) x all 256 output columns -> partial-sum plane r
        j.ta = &maps.a_f; j.tw = &maps.w_fc2; j.w_row0 = l * 256; j.plane_rows = a.plane_rows[5]; j.ncols = 256; j.kind = GK_PART;
        j.kblocks = 2;
        j.a_col0 = j.w_col0 = r * 128;
      } else {
        gemm = false;
      }
      if (gemm) {
        gemm_phase(c, a, j);
      } else if (ph == PH_INIT) {  // h = the learned query embeddings (deformable_detr.py:2290-2292)
        if (l == 0) ln_phase(c, a, nullptr, 0, nullptr, a.tgt, 0, nullptr, nullptr, a.hf, a.hp, nullptr);
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
        if (ph == PH_LN3) ln_phase(c, a, a.part, CL, vec + 512, res, nb, gam, gam + 256, outf, outp, a.inter + ((long long)b * a.L + l) * N * 256);
        else ln_phase(c, a, a.o, 1, nullptr, res, nb, gam, gam + 256, outf, outp, nullptr);
      }
      phase_end((ph == PH_QKV && a.mha_mode && a.zsplit == 1) || ph == PH_OFFAW || ph == PH_FC1, a.dbg);
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
  // hf t1f t2f o (fp32 rows) + hp t1p t2p attn_p attn2_p (P32 rows) = 9 KB per row; offaw 1.5 KB; f_p 4 KB; qk_p 2 KB; fc2 partial sums 8 KB; + V^T 256 KB per image
  return R * (9 * 1024 + 1536 + 4096 + 2048 + 8 * 1024) + (long long)B * 256 * 1024 + 1024;
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
  a.part = (float*)take(R * 1024 * CL);
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
  if ((rc = tmap_weight_planes(w->w_fc2, 1024, 2ll * L * 256, 256, &m.w_fc2)) != EGTR_OK) return rc;
  // A lone forward spreads the stack over a 16-CTA cluster (rank = 8 z + r: m-tile z of head / column slice r) — twice the SMs'
  // worth of operand ingest, load/store and SIMT throughput for the row-local phases; forwards in flight keep the 8-CTA cluster
  // (16 co-scheduled SMs of one GPC are hard to come by next to other images' persistent GEMMs).  EGTR_DECODER_CLUSTER=8|16 forces.
  static int max16 = -1;
  const char* fe = getenv("EGTR_DECODER_CLUSTER");
  const int forced = fe ? atoi(fe) : 0;
  if (max16 < 0) {
    EGTR_CUDA(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    max16 = 0;
    if (cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16, 1);
      cfg.blockDim = dim3(THREADS);
      cfg.dynamicSmemBytes = SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, decoder_kernel, &cfg) == cudaSuccess) max16 = n;
    }
    (void)cudaGetLastError();
  }
  a.zsplit = (a.MT == 2 && max16 >= 1 && (forced == 16 || (forced == 0 && grid_div() == 1))) ? 2 : 1;
  EGTR_CUDA(launch_cluster_pdl(decoder_kernel, dim3(CL * a.zsplit, B), dim3(THREADS), (size_t)SMEM_BYTES, (cudaStream_t)s, CL * a.zsplit, m, a));
  count_launch();
  return EGTR_OK;
}
