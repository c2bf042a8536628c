// Relation head (model/egtr.py:322-418, 507-516) as ONE fused pair kernel on tcgen05 CTA pairs.
//
// Inputs are the per-query layer-1 partials U_l(i), V_l(j) (relation.cu's algebra: the reference's N x N x 7 x 512
// `relation_source` is never formed).  For every tile of 16 subjects x 16 objects a cluster of two CTAs (8 subjects each)
//   1. stages the tile's U / V slices (32 channels x 7 layers x 24 queries, 21 KB) in shared memory by TMA — each vector
//      is read from L2 once per tile instead of once per thread;
//   2. producer warps form  h1 = relu(b1 + sum_l gate_l(i,j) (U_l(i) + V_l(j)))  for 128 pairs x 32 channels, split it into
//      bf16 hi/lo and store it as a swizzled K-major operand tile (never in HBM);
//   3. layer 2 runs as cta_group::2 tcgen05.mma (M = 256 pairs, N = 256, bf16x3 products, fp32 accumulators in TMEM), the
//      weights streamed by TMA, half of the tile per CTA;
//   4. relation items: epilogue warps read the accumulator, add bias, ReLU, re-split and re-stage the hidden rows in shared
//      memory as the A operand of the layer-3 MMA (N = P padded to 64, issued as two N = 32 halves whose accumulators alias
//      the already drained columns [0,32) and [128,160) of the same TMEM buffer), then finish
//      sigmoid(acc + b3 + triplet_dist[c_i, c_j] - tau log rel_dist) straight into pred_rel;
//      connectivity items: the 256 -> 1 last layer is a dot product in the epilogue, sigmoid, pred_connectivity.
// Nothing of the pair dimension except the two outputs ever reaches HBM.
//
// Work items are (pair tile, phase) with phase 0 = relation MLP, 1 = connectivity MLP (the two MLPs read disjoint
// channel halves of h1), dealt round-robin to the persistent CTA pairs.
//
// Roles per CTA (576 threads): warps 0-7 epilogue (TMEM quadrant = warp & 3, column half = warp >> 2), warps 8-15
// h1 producers, warp 16 TMA (weights, U/V, layer-3 weights; a non-blocking three-cursor loop), warp 17 MMA issue (leader CTA).
#include <cuda.h>

#include <mutex>
#include <string.h>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace egtr {

void count_launch();

namespace {

constexpr int RH_THREADS = 576;
constexpr int PROD_WARP0 = 8, TMA_WARP = 16, MMA_WARP = 17;
constexpr int MAX_SW = 4, MAX_SH = 3, MAX_SU = 3;  // barrier arrays are sized for the larger variant
constexpr int TILE_BYTES = 128 * 128;
constexpr int U_ROWS = 8, V_ROWS = 16, MAX_LR = 7;
constexpr int UV_BYTES = (U_ROWS + V_ROWS) * MAX_LR * 128;
constexpr int GROUPS = 8;  // 32-channel groups per 256-wide MLP input

// BIG == false: P <= 64.  Two 256-column layer-2 accumulators in TMEM (the epilogue of item k overlaps the MMAs of item k + 1); the
//   layer-3 accumulators (two N = 32 halves) alias the already drained columns [0,32) and [128,160) of the item's own buffer.
// BIG == true: 64 < P <= 256 (stress config E: 200 predicates).  One layer-2 accumulator (columns 0..255) and one layer-3
//   accumulator of P3 = 64 * ceil(P / 64) columns (256..); layer 3 is one N = P3 MMA per k-step, its weight stage P3 / 2 rows per CTA.
template <bool BIG>
struct RhCfg {
  static constexpr int SW = BIG ? 3 : 4;   // layer-2 weight stages (16 KB each: this CTA's 128 weight rows x one 32-channel group)
  static constexpr int SH = 3;             // h1 operand stages (16 KB each: 128 pairs x one group, hi 64 B | lo 64 B per row)
  static constexpr int SU = BIG ? 2 : 3;   // U/V staging buffers
  static constexpr int W3_BYTES = BIG ? 128 * 128 : 32 * 128;  // this CTA's layer-3 weight rows x one group
  static constexpr int OFF_W2 = 0;
  static constexpr int OFF_H1 = OFF_W2 + SW * TILE_BYTES;
  static constexpr int OFF_H2 = OFF_H1 + SH * TILE_BYTES;
  static constexpr int OFF_W3 = OFF_H2 + 2 * TILE_BYTES;
  static constexpr int OFF_UV = OFF_W3 + 2 * W3_BYTES;
  static constexpr int OFF_GATE = OFF_UV + SU * UV_BYTES;    // [128 rows][8] gate values of the current item
  static constexpr int OFF_XCH = OFF_GATE + 128 * 8 * 4;     // [2][8 warps][32] connectivity partial dots
  static constexpr int OFF_B3 = OFF_XCH + 2 * 8 * 32 * 4;    // [P3] b3 - tau log rel_dist
  static constexpr int OFF_BAR = OFF_B3 + (BIG ? 256 : 64) * 4;
  static constexpr int SMEM = OFF_BAR + 512 + 1024 /*align slack*/;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct RelArgs {
  const float* U;        // [B*N, Lr, ldu]  subject-side partials: rel 0..255 | conn 256..511 | gate logit at 512
  const float* V;        // same, object side (gate bias folded in)
  const float* b1;       // [512]
  const float* b2;       // [512] layer-2 bias, rel | conn
  const float* b3;       // [>= P] layer-3 bias of the relation MLP
  const float* rel_dist; // [P] or NULL (no logit adjustment)
  const float* w3c;      // [256] last layer of the connectivity MLP
  const int* cls;        // [B*N] argmax classes, or NULL (no frequency bias)
  const float* triplet;  // [K1, K1, P]
  float* pred_rel;       // [B, N, N, P]
  float* pred_conn;      // [B, N, N]
  float b3c, tau;
  int B, N, P, Lr, ldu, k1;
  int tiles_i, tiles_j;  // pair tiles per image (16 x 16)
  int P3;                // BIG: layer-3 columns, 64 * ceil(P / 64)
};

struct Item {
  int phase, b, i0, j0;
};

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {  // low 16 bits = a
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// packed f32x2 arithmetic (Blackwell FADD2 / FFMA2) on 64-bit register pairs: low word = first element
__device__ __forceinline__ unsigned long long pack_f32x2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long x, unsigned long long y) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long y, unsigned long long z) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  return r;
}

template <bool BIG>
__global__ void __launch_bounds__(RH_THREADS, 1)
relhead_kernel(const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_w3,
               const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_v, const RelArgs a,
               int* __restrict__ err) {
  pdl_launch_dependents();
  using Cfg = RhCfg<BIG>;
  constexpr int SW = Cfg::SW, SH = Cfg::SH, SU = Cfg::SU, W3_BYTES = Cfg::W3_BYTES;
  constexpr int OFF_W2 = Cfg::OFF_W2, OFF_H1 = Cfg::OFF_H1, OFF_H2 = Cfg::OFF_H2, OFF_W3 = Cfg::OFF_W3, OFF_UV = Cfg::OFF_UV;
  constexpr int OFF_GATE = Cfg::OFF_GATE, OFF_XCH = Cfg::OFF_XCH, OFF_B3 = Cfg::OFF_B3, OFF_BAR = Cfg::OFF_BAR;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_s = ptx::smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // [SW]  leader: both CTAs' weight boxes of the stage have landed
  uint64_t* w_empty = w_full + MAX_SW;     // [SW]  the MMAs that read the stage have completed (commit, both CTAs)
  uint64_t* h1_full = w_empty + MAX_SW;    // [SH]  leader: 16 producer warps of the pair have stored the stage
  uint64_t* h1_empty = h1_full + MAX_SH;   // [SH]
  uint64_t* uv_full = h1_empty + MAX_SH;   // [SU]  local
  uint64_t* uv_empty = uv_full + MAX_SU;   // [SU]  local: 8 producer warps
  uint64_t* l2_full = uv_empty + MAX_SU;   // [2]   layer-2 accumulator of buffer b complete (commit, both CTAs)
  uint64_t* acc_free = l2_full + 2;    // [2]   leader: 16 epilogue warps have read everything they need from buffer b
  uint64_t* h2_full = acc_free + 2;    // [2]   leader: restaged hidden group (8 warps of the pair) + layer-3 weights (tx)
  uint64_t* h2_empty = h2_full + 2;    // [2]
  uint64_t* l3_full = h2_empty + 2;    // [1]
  uint32_t* tmem_holder = (uint32_t*)(l3_full + 1);
  float* gate_tab = (float*)(smem + OFF_GATE);
  float* xch = (float*)(smem + OFF_XCH);
  float* b3adj = (float*)(smem + OFF_B3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int slots = gridDim.x >> 1, me = blockIdx.x >> 1;
  const int tiles_img = a.tiles_i * a.tiles_j;
  const int T = a.B * tiles_img;  // pair tiles; items = 2 * T, the relation items first
  const int items = 2 * T;
  auto decode = [&](int w) {
    Item it;
    it.phase = w >= T ? 1 : 0;
    const int t = w - it.phase * T;
    it.b = t / tiles_img;
    const int r = t - it.b * tiles_img;
    const int ti = r / a.tiles_j;
    it.i0 = ti * 16 + rank * 8;
    it.j0 = (r - ti * a.tiles_j) * 16;
    return it;
  };
  const int n_items = me < items ? (items - me + slots - 1) / slots : 0;   // items of this CTA pair
  const int n_rel = me < T ? (T - me + slots - 1) / slots : 0;             // ... of which relation items (they come first)

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmap_w2);
    ptx::prefetch_tensormap(&tmap_w3);
    ptx::prefetch_tensormap(&tmap_u);
    ptx::prefetch_tensormap(&tmap_v);
    for (int i = 0; i < SW; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < SH; ++i) { ptx::mbar_init(&h1_full[i], 16); ptx::mbar_init(&h1_empty[i], 1); }
    for (int i = 0; i < SU; ++i) { ptx::mbar_init(&uv_full[i], 1); ptx::mbar_init(&uv_empty[i], 8); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&l2_full[i], 1);
      ptx::mbar_init(&acc_free[i], 16);
      ptx::mbar_init(&h2_full[i], 9);
      ptx::mbar_init(&h2_empty[i], 1);
    }
    ptx::mbar_init(l3_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == MMA_WARP) ptx::tmem_alloc_2cta<512>(tmem_holder);
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------ TMA: three independent cursors, never blocking on one
    if (lane == 0) {
      const int tot = n_items * GROUPS, tot3 = n_rel * GROUPS;
      int cw = 0, cu = 0, c3 = 0;
      const uint32_t uv_tx = (uint32_t)(a.Lr * (U_ROWS + V_ROWS) * 128);
      long long t0 = clock64();
      while (cw < tot || cu < tot || c3 < tot3) {
        bool progress = false;
        if (cw < tot) {  // layer-2 weights of (item, group): this CTA's 128 rows
          const int st = cw % SW;
          if (ptx::mbar_try_wait(&w_empty[st], ((cw / SW) & 1) ^ 1)) {
            const int phase = (me + (cw >> 3) * slots) >= T ? 1 : 0;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&w_full[st], 2 * TILE_BYTES);
            ptx::tma_load_2d_2cta(smem_s + OFF_W2 + st * TILE_BYTES, &tmap_w2, &w_full[st], (cw & 7) * 64, phase * 256 + rank * 128);
            ++cw;
            progress = true;
          }
        }
        if (cu < tot) {  // U (8 subjects) and V (16 objects) slices of (item, group)
          const int bf = cu % SU;
          if (ptx::mbar_try_wait(&uv_empty[bf], ((cu / SU) & 1) ^ 1)) {
            const Item it = decode(me + (cu >> 3) * slots);
            const int c0 = it.phase * 256 + (cu & 7) * 32;
            const uint32_t dst = smem_s + OFF_UV + bf * UV_BYTES;
            ptx::mbar_arrive_expect_tx(&uv_full[bf], uv_tx);
            tma_load_3d(dst, &tmap_u, &uv_full[bf], c0, 0, it.b * a.N + it.i0);
            tma_load_3d(dst + U_ROWS * a.Lr * 128, &tmap_v, &uv_full[bf], c0, 0, it.b * a.N + it.j0);
            ++cu;
            progress = true;
          }
        }
        if (c3 < tot3) {  // layer-3 weights of (relation item, chunk, column set): this CTA's 32 rows of group set*4 + chunk
          const int set = c3 & 1, use = c3 >> 1;
          if (ptx::mbar_try_wait(&h2_empty[set], (use & 1) ^ 1)) {
            const int group = set * 4 + (use & 3);
            const int rows3 = BIG ? a.P3 / 2 : 32;  // layer-3 weight rows this CTA stages (the map's box height)
            if (rank == 0) ptx::mbar_arrive_expect_tx(&h2_full[set], 2 * rows3 * 128);
            ptx::tma_load_2d_2cta(smem_s + OFF_W3 + set * W3_BYTES, &tmap_w3, &h2_full[set], group * 64, rank * rows3);
            ++c3;
            progress = true;
          }
        }
        if (progress) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) { if (err) atomicExch(err, 301); __threadfence_system(); __trap(); }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issue (leader CTA)
    if (rank == 0) {
      constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(256, 256);
      const uint32_t idesc3 = BIG ? ptx::umma_idesc_bf16(256, a.P3) : ptx::umma_idesc_bf16(256, 32);
      int n = 0, u3[2] = {0, 0}, n_l3 = 0;
      int prev_rel = 0, prev_buf = 0;
      // layer 3 of a relation item, one 32-channel group of the hidden layer per step, in the order the two column sets restage them
      auto l3_step = [&](int idx) {
        const int set = idx & 1;
        if (idx == 0) {
          if (BIG) {  // the layer-3 accumulator is its own buffer: the previous relation item's result has been read
            ptx::mbar_wait(&acc_free[1], (n_l3 & 1) ^ 1, err, 317);
            ptx::mbar_wait(&h2_full[0], u3[0] & 1, err, 311);
          } else {    // columns [0,32) and [128,160) of the item's buffer are overwritten: both first chunks must be drained
            ptx::mbar_wait(&h2_full[0], u3[0] & 1, err, 311);
            ptx::mbar_wait(&h2_full[1], u3[1] & 1, err, 312);
          }
        } else if (idx >= 2 || BIG) {
          ptx::mbar_wait(&h2_full[set], u3[set] & 1, err, 313);
        }
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_s + OFF_H2 + set * TILE_BYTES, b0 = smem_s + OFF_W3 + set * W3_BYTES;
          if (BIG) {
            const uint32_t d = tmem_base + 256;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t dah = ptx::umma_desc_sw128(a0 + j * 32), dal = ptx::umma_desc_sw128(a0 + j * 32 + 64);
              const uint64_t dbh = ptx::umma_desc_sw128(b0 + j * 32), dbl = ptx::umma_desc_sw128(b0 + j * 32 + 64);
              ptx::umma_bf16_2cta(d, dal, dbh, idesc3, (idx != 0) || (j != 0));
              ptx::umma_bf16_2cta(d, dah, dbl, idesc3, 1);
              ptx::umma_bf16_2cta(d, dah, dbh, idesc3, 1);
            }
          } else {
            const uint32_t d0 = tmem_base + prev_buf * 256;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t dah = ptx::umma_desc_sw128(a0 + j * 32), dal = ptx::umma_desc_sw128(a0 + j * 32 + 64);
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const uint32_t bb = b0 + hf * 2048 + j * 32;
                const uint64_t dbh = ptx::umma_desc_sw128(bb), dbl = ptx::umma_desc_sw128(bb + 64);
                const uint32_t d = d0 + hf * 128;
                ptx::umma_bf16_2cta(d, dal, dbh, idesc3, (idx != 0) || (j != 0));
                ptx::umma_bf16_2cta(d, dah, dbl, idesc3, 1);
                ptx::umma_bf16_2cta(d, dah, dbh, idesc3, 1);
              }
            }
          }
          ptx::umma_commit_2cta(&h2_empty[set]);
          if (idx == 7) ptx::umma_commit_2cta(l3_full);
        }
        __syncwarp();
        ++u3[set];
        if (idx == 7) ++n_l3;
      };
      for (int k = 0; k < n_items; ++k) {
        const int w = me + k * slots;
        const int buf = BIG ? 0 : (k & 1);
        ptx::mbar_wait(&acc_free[buf], (BIG ? (k & 1) : ((k >> 1) & 1)) ^ 1, err, 314);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int g = 0; g < GROUPS; ++g, ++n) {
          const int st = n % SW, sh = n % SH;
          ptx::mbar_wait(&w_full[st], (n / SW) & 1, err, 315);
          ptx::mbar_wait(&h1_full[sh], (n / SH) & 1, err, 316);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t a0 = smem_s + OFF_H1 + sh * TILE_BYTES, b0 = smem_s + OFF_W2 + st * TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t dah = ptx::umma_desc_sw128(a0 + j * 32), dal = ptx::umma_desc_sw128(a0 + j * 32 + 64);
              const uint64_t dbh = ptx::umma_desc_sw128(b0 + j * 32), dbl = ptx::umma_desc_sw128(b0 + j * 32 + 64);
              ptx::umma_bf16_2cta(d_tmem, dal, dbh, idesc2, (g != 0) || (j != 0));  // small terms first
              ptx::umma_bf16_2cta(d_tmem, dah, dbl, idesc2, 1);
              ptx::umma_bf16_2cta(d_tmem, dah, dbh, idesc2, 1);
            }
            ptx::umma_commit_2cta(&w_empty[st]);
            ptx::umma_commit_2cta(&h1_empty[sh]);
            if (g == GROUPS - 1) ptx::umma_commit_2cta(&l2_full[buf]);
          }
          __syncwarp();
          if (!BIG && prev_rel && g >= 1) l3_step(g - 1);  // layer 3 of the previous item rides behind this item's layer 2
        }
        if (BIG) {
          // one accumulator: the next item's layer 2 cannot start before this item's hidden layer has left TMEM, and the restage
          // of chunk c waits for the layer-3 MMAs of chunk c - 1 — so layer 3 of THIS item runs here, paced by the epilogue
          if (w < T)
            for (int s = 0; s < 8; ++s) l3_step(s);
        } else {
          if (prev_rel) l3_step(7);
          prev_rel = w < T;
          prev_buf = buf;
        }
      }
      if (!BIG && prev_rel)
        for (int s = 0; s < 8; ++s) l3_step(s);
    }
  } else if (warp >= PROD_WARP0) {
    // ------------------------------------------------------------------ h1 producers (8 warps)
    const int tid_p = threadIdx.x - PROD_WARP0 * 32;
    const int pw = tid_p >> 5;
    // a thread owns a 2 subjects x 2 objects block (tile rows r0 .. r0+3, row = r0 + 2*ds + do) and 4 of the group's 32 channels
    const int blkid = pw * 4 + (lane >> 3), c4 = lane & 7;
    const int sp = blkid >> 3, op = blkid & 7, r0 = blkid * 4;
    // gate table duty: this thread computes 4 of the 128 x 8 gate values of an item (row gr, layers gl0 .. gl0+3)
    const int gr = tid_p >> 1, gl0 = (tid_p & 1) * 4;
    const int g_s = ((gr >> 5) << 1) | ((gr >> 1) & 1), g_o = (((gr >> 2) & 7) << 1) | (gr & 1);
    const long long qstride = (long long)a.Lr * a.ldu;
    auto gate_logits = [&](const Item& it, float (&gl)[4]) {
      const long long qi = (long long)it.b * a.N + min(it.i0 + g_s, a.N - 1), qj = (long long)it.b * a.N + min(it.j0 + g_o, a.N - 1);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int l = gl0 + kk;
        gl[kk] = l < a.Lr ? __ldg(a.U + qi * qstride + l * a.ldu + 512) + __ldg(a.V + qj * qstride + l * a.ldu + 512) : 0.f;
      }
    };
    float glog[4] = {0.f, 0.f, 0.f, 0.f};
    if (n_items > 0) gate_logits(decode(me), glog);
    const uint32_t u_off = (uint32_t)(sp * 2 * a.Lr * 128 + c4 * 16);
    const uint32_t v_off = (uint32_t)(U_ROWS * a.Lr * 128 + op * 2 * a.Lr * 128 + c4 * 16);
    const uint32_t row_stride = (uint32_t)(a.Lr * 128);  // bytes between consecutive queries of a staged slice
    int n = 0;
    for (int k = 0; k < n_items; ++k) {
      const Item it = decode(me + k * slots);
      ptx::named_bar_sync(1, 256);  // every producer has taken its gates of the previous item out of the table
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) gate_tab[gr * 8 + gl0 + kk] = (gl0 + kk < a.Lr) ? fast_sigmoid(glog[kk]) : 0.f;
      ptx::named_bar_sync(1, 256);
      float gt[4][8];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float4 g0 = *(const float4*)(gate_tab + (r0 + p) * 8), g1 = *(const float4*)(gate_tab + (r0 + p) * 8 + 4);
        gt[p][0] = g0.x; gt[p][1] = g0.y; gt[p][2] = g0.z; gt[p][3] = g0.w;
        gt[p][4] = g1.x; gt[p][5] = g1.y; gt[p][6] = g1.z; gt[p][7] = g1.w;
      }
      if (k + 1 < n_items) gate_logits(decode(me + (k + 1) * slots), glog);  // in flight behind this item's arithmetic
      for (int g = 0; g < GROUPS; ++g, ++n) {
        const int bf = n % SU, sh = n % SH;
        const float4 bias = __ldg((const float4*)(a.b1 + it.phase * 256 + g * 32 + c4 * 4));
        // accumulators and operands stay packed as f32x2 register pairs: one FADD2 + one FFMA2 per two channels
        unsigned long long acc[4][2];
#pragma unroll
        for (int p = 0; p < 4; ++p) { acc[p][0] = pack_f32x2(bias.x, bias.y); acc[p][1] = pack_f32x2(bias.z, bias.w); }
        ptx::mbar_wait(&uv_full[bf], (n / SU) & 1, err, 321);
        const uint32_t ub = smem_s + OFF_UV + bf * UV_BYTES + u_off, vb = smem_s + OFF_UV + bf * UV_BYTES + v_off;
        // software-pipelined over the MAX_LR layer slots: the loads of slot l + 1 are in flight behind the arithmetic of slot l
        // (slots >= Lr re-read the last layer with gate 0 — no data-dependent branch, so the loads can be hoisted)
        unsigned long long U[2][2][2], V[2][2][2];  // [buffer][query of the 2 x 2 block][channel pair]
        auto ld_layer = [&](int l, int bsel) {
          const uint32_t lo = (uint32_t)min(l, a.Lr - 1) * 128u;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(U[bsel][d][0]), "=l"(U[bsel][d][1]) : "r"(ub + d * row_stride + lo));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(V[bsel][d][0]), "=l"(V[bsel][d][1]) : "r"(vb + d * row_stride + lo));
          }
        };
        ld_layer(0, 0);
#pragma unroll
        for (int l = 0; l < MAX_LR; ++l) {
          if (l + 1 < MAX_LR) ld_layer(l + 1, (l + 1) & 1);
          const int bsel = l & 1;
#pragma unroll
          for (int ds = 0; ds < 2; ++ds)
#pragma unroll
            for (int dd = 0; dd < 2; ++dd) {
              const int p = ds * 2 + dd;
              const unsigned long long gg = pack_f32x2(gt[p][l], gt[p][l]);
              acc[p][0] = fma2(gg, add2(U[bsel][ds][0], V[bsel][dd][0]), acc[p][0]);
              acc[p][1] = fma2(gg, add2(U[bsel][ds][1], V[bsel][dd][1]), acc[p][1]);
            }
        }
        __syncwarp();  // every lane has read the staged slices
        if (lane == 0) ptx::mbar_arrive(&uv_empty[bf]);
        uint32_t hi[4][2], lo[4][2];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float a0, a1;
            unpack_f32x2(acc[p][h], a0, a1);
            const float x0 = fmaxf(a0, 0.f), x1 = fmaxf(a1, 0.f);
            const uint32_t hh = pack2_bf16(x0, x1);
            hi[p][h] = hh;
            lo[p][h] = pack2_bf16(x0 - __uint_as_float(hh << 16), x1 - __uint_as_float(hh & 0xffff0000u));
          }
        ptx::mbar_wait(&h1_empty[sh], ((n / SH) & 1) ^ 1, err, 322);
        const uint32_t tile = smem_s + OFF_H1 + sh * TILE_BYTES;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int r = r0 + p;
          const uint32_t rowb = tile + r * 128 + (c4 & 1) * 8;
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + ((((c4 >> 1)) ^ (r & 7)) << 4)), "r"(hi[p][0]), "r"(hi[p][1]) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + (((4 + (c4 >> 1)) ^ (r & 7)) << 4)), "r"(lo[p][0]), "r"(lo[p][1]) : "memory");
        }
        ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's asynchronous proxy
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(&h1_full[sh]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3, set = warp >> 2;
    if (threadIdx.x < (BIG ? 256 : 64)) {
      float v = 0.f;
      if ((int)threadIdx.x < a.P) v = a.b3[threadIdx.x] - (a.rel_dist ? a.tau * logf(a.rel_dist[threadIdx.x]) : 0.f);  // egtr.py:509-512
      b3adj[threadIdx.x] = v;
    }
    ptx::named_bar_sync(6, 256);
    const int row = q * 32 + lane;
    const int s_loc = 2 * q + ((lane >> 1) & 1), o_loc = 2 * (lane >> 2) + (lane & 1);
    const uint32_t my_row_s = smem_s + OFF_H2 + set * TILE_BYTES + row * 128;
    const int sw = row & 7;
    int u3 = 0, n_r = 0, n_c = 0;
    for (int k = 0; k < n_items; ++k) {
      const Item it = decode(me + k * slots);
      const int buf = BIG ? 0 : (k & 1);
      const int i = it.i0 + s_loc, j = it.j0 + o_loc;
      const bool valid = i < a.N && j < a.N;
      const long long pair = ((long long)it.b * a.N + i) * a.N + j;
      ptx::mbar_wait(&l2_full[buf], BIG ? (k & 1) : ((k >> 1) & 1), err, 331);
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
      uint32_t r[32];
      if (it.phase == 0) {
        // ---- relation MLP: hidden = relu(acc + b2) re-split and re-staged as the A operand of layer 3
#pragma unroll 1
        for (int c = 0; c < 4; ++c, ++u3) {
          const int group = set * 4 + c;
          ptx::tmem_ld_32x32(t_acc + group * 32, r);
          ptx::tmem_ld_wait();
          uint32_t o[32];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 b4 = __ldg((const float4*)(a.b2 + group * 32) + jj);
            const float x0 = fmaxf(__uint_as_float(r[4 * jj]) + b4.x, 0.f), x1 = fmaxf(__uint_as_float(r[4 * jj + 1]) + b4.y, 0.f);
            const float x2 = fmaxf(__uint_as_float(r[4 * jj + 2]) + b4.z, 0.f), x3 = fmaxf(__uint_as_float(r[4 * jj + 3]) + b4.w, 0.f);
            const uint32_t h0 = pack2_bf16(x0, x1), h1 = pack2_bf16(x2, x3);
            o[2 * jj] = h0;
            o[2 * jj + 1] = h1;
            o[16 + 2 * jj] = pack2_bf16(x0 - __uint_as_float(h0 << 16), x1 - __uint_as_float(h0 & 0xffff0000u));
            o[16 + 2 * jj + 1] = pack2_bf16(x2 - __uint_as_float(h1 << 16), x3 - __uint_as_float(h1 & 0xffff0000u));
          }
          ptx::mbar_wait(&h2_empty[set], (u3 & 1) ^ 1, err, 332);
#pragma unroll
          for (int cc = 0; cc < 8; ++cc)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_s + ((cc ^ sw) << 4)), "r"(o[4 * cc]), "r"(o[4 * cc + 1]),
                         "r"(o[4 * cc + 2]), "r"(o[4 * cc + 3]) : "memory");
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_before();  // the accumulator columns of this chunk are read: layer 3 may overwrite them
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(&h2_full[set]);
        }
        if (BIG) {  // the hidden layer has left the layer-2 accumulator: the next item's MMAs may overwrite it
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(&acc_free[0]);
        }
        ptx::mbar_wait(l3_full, n_r & 1, err, 333);
        ++n_r;
        ptx::tc_fence_after();
        const float* trip = nullptr;
        if (valid && a.cls != nullptr) trip = a.triplet + ((long long)a.cls[(long long)it.b * a.N + i] * a.k1 + a.cls[(long long)it.b * a.N + j]) * a.P;  // egtr.py:405-413
        float* out = a.pred_rel + pair * a.P;
        auto finish32 = [&](int p0) {  // r = logits of predicates p0 .. p0 + 31 of this lane's pair
          if (!valid) return;
          if ((a.P & 1) == 0) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              const int p = p0 + jj;
              if (p < a.P) {
                float x0 = __uint_as_float(r[jj]) + b3adj[p], x1 = __uint_as_float(r[jj + 1]) + b3adj[p + 1];
                if (trip) { const float2 t2 = __ldg((const float2*)(trip + p)); x0 += t2.x; x1 += t2.y; }
                *(float2*)(out + p) = make_float2(fast_sigmoid(x0), fast_sigmoid(x1));
              }
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int p = p0 + jj;
              if (p < a.P) {
                float x0 = __uint_as_float(r[jj]) + b3adj[p];
                if (trip) x0 += __ldg(trip + p);
                out[p] = fast_sigmoid(x0);
              }
            }
          }
        };
        if (BIG) {
          // this warp's half of the P3 layer-3 columns, 32 at a time; the accumulator is released after the last read
          const int half3 = a.P3 >> 1, nch3 = half3 >> 5;
          for (int c = 0; c < nch3; ++c) {
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256 + set * half3 + c * 32, r);
            ptx::tmem_ld_wait();
            if (c == nch3 - 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_leader(&acc_free[1]);
            }
            finish32(set * half3 + c * 32);
          }
        } else {
          ptx::tmem_ld_32x32(t_acc + set * 128, r);  // predicates set*32 .. set*32+31
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(&acc_free[buf]);
          finish32(set * 32);
        }
      } else {
        // ---- connectivity MLP: last layer (256 -> 1) as a dot product over this warp's 128 columns
        float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int group = set * 4 + c;
          ptx::tmem_ld_32x32(t_acc + group * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 b4 = __ldg((const float4*)(a.b2 + 256 + group * 32) + jj), w4 = __ldg((const float4*)(a.w3c + group * 32) + jj);
            dot = fmaf(fmaxf(__uint_as_float(r[4 * jj]) + b4.x, 0.f), w4.x, dot);
            dot = fmaf(fmaxf(__uint_as_float(r[4 * jj + 1]) + b4.y, 0.f), w4.y, dot);
            dot = fmaf(fmaxf(__uint_as_float(r[4 * jj + 2]) + b4.z, 0.f), w4.z, dot);
            dot = fmaf(fmaxf(__uint_as_float(r[4 * jj + 3]) + b4.w, 0.f), w4.w, dot);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(&acc_free[buf]);
        float* mine = xch + ((n_c & 1) * 8 + warp) * 32;
        mine[lane] = dot;
        ptx::named_bar_sync(2 + q, 64);  // the two warps that share these 32 rows
        if (set == 0 && valid) a.pred_conn[pair] = fast_sigmoid(dot + xch[((n_c & 1) * 8 + warp + 4) * 32 + lane] + a.b3c);  // egtr.py:416, 516
        ++n_c;
      }
    }
  }

  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync_all();  // the peer may still signal this CTA's barriers / read its operand tiles
  if (warp == MMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<512>(tmem_base);
  }
}

// fp32 [N,K] rows -> "P32 group" weight rows: row n = K/32 groups of (32 bf16 hi | 32 bf16 lo); output row r takes source
// row perm[r] (perm == NULL: r); source rows >= N are zero.
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int K, int rows_out, const int* __restrict__ perm,
                                   __nv_bfloat16* __restrict__ out) {
  pdl_entry();
  const long long total = (long long)rows_out * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i - (long long)r * K);
    const int src = perm ? perm[r] : r;
    __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
    if (src >= 0 && src < N) split_bf16(w[(long long)src * K + k], h, l);
    __nv_bfloat16* g = out + (long long)r * 2 * K + (k >> 5) * 64 + (k & 31);
    g[0] = h;
    g[32] = l;
  }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn rh_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  });
  return fn;
}

struct RhMapKey {
  const void* ptr;
  unsigned long long d0, d1, d2, s1, s2;
  unsigned b0, b1, b2;
  int dtype, rank, swizzle;
};
struct RhMapHash {
  size_t operator()(const RhMapKey& d) const {
    const unsigned char* b = (const unsigned char*)&d;
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(RhMapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
};
struct RhMapEq {
  bool operator()(const RhMapKey& x, const RhMapKey& y) const { return memcmp(&x, &y, sizeof(RhMapKey)) == 0; }
};

int rh_map(const void* ptr, int dtype, int rank, unsigned long long d0, unsigned long long d1, unsigned long long d2,
           unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1, unsigned b2, int swizzle, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<RhMapKey, CUtensorMap, RhMapHash, RhMapEq> cache;
  RhMapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.d0 = d0; key.d1 = d1; key.d2 = d2; key.s1 = s1; key.s2 = s2; key.b0 = b0; key.b1 = b1; key.b2 = b2;
  key.dtype = dtype; key.rank = rank; key.swizzle = swizzle;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return EGTR_OK; }
  EncodeTiledFn enc = rh_encode_fn();
  EGTR_CHECK(enc != nullptr, EGTR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, (CUtensorMapDataType)dtype, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGTR_CHECK(r == CUDA_SUCCESS, EGTR_ERR_CUDA, "cuTensorMapEncodeTiled (relation head) failed with CUresult %d", (int)r);
  if (cache.size() >= (1u << 15)) cache.clear();  // bounded: maps are passed to kernels by value, re-encoding is cheap
  cache.emplace(key, m);
  *out = m;
  return EGTR_OK;
}

int* rh_error_flag() {
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

extern "C" int egtr_pack_weight_p32g(const float* w, int N, int K, int rows_out, const int* perm, void* out, egtr_stream_t s) {
  EGTR_CHECK(w && out && N > 0 && K > 0 && K % 32 == 0 && rows_out > 0, EGTR_ERR_ARG,
             "egtr_pack_weight_p32g: bad arguments (N=%d K=%d rows_out=%d)", N, K, rows_out);
  const long long total = (long long)rows_out * K;
  int grid = cdiv(total, 256);
  if (grid > 4 * 148) grid = 4 * 148;
  launch_pdl(pack_weight_kernel, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)s, w, N, K, rows_out, perm, (__nv_bfloat16*)out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_relation_pairs_fused_f32(const float* U, const float* V, int ldu, int layers, const egtr_relhead_weights_t* w,
                                             const int* cls, const float* triplet_dist, int k1, const float* rel_dist, float tau,
                                             int B, int N, int P, float* pred_rel, float* pred_conn, egtr_stream_t s) {
  EGTR_ONE_DEVICE();
  EGTR_CHECK(U && V && w && pred_rel && pred_conn && w->b1 && w->w2g && w->b2 && w->w3g && w->b3 && w->w3c, EGTR_ERR_ARG,
             "egtr_relation_pairs_fused_f32: null pointer");
  EGTR_CHECK(B > 0 && N > 0 && P > 0 && layers >= 1, EGTR_ERR_ARG, "egtr_relation_pairs_fused_f32: empty shape");
  EGTR_CHECK(P <= 256 && layers <= MAX_LR && ldu >= 513 && ldu % 4 == 0, EGTR_ERR_UNSUPPORTED,
             "egtr_relation_pairs_fused_f32: built for P <= 256 predicates and <= 7 layers (P=%d layers=%d ldu=%d)", P, layers, ldu);
  EGTR_CHECK(cls == nullptr || triplet_dist != nullptr, EGTR_ERR_ARG, "egtr_relation_pairs_fused_f32: triplet_dist missing");
  EGTR_CHECK(((uintptr_t)U & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)w->w2g & 127) == 0 && ((uintptr_t)w->w3g & 127) == 0 &&
                 ((uintptr_t)w->b1 & 15) == 0 && ((uintptr_t)w->b2 & 15) == 0 && ((uintptr_t)w->w3c & 15) == 0 &&
                 ((uintptr_t)pred_rel & 7) == 0 && (triplet_dist == nullptr || ((uintptr_t)triplet_dist & 7) == 0),
             EGTR_ERR_ARG, "egtr_relation_pairs_fused_f32: alignment");
  EGTR_CHECK((long long)B * N * N * P < (1LL << 40) && (long long)B * N < (1LL << 31), EGTR_ERR_ARG, "egtr_relation_pairs_fused_f32: size");
  RelArgs a = {};
  a.U = U; a.V = V; a.b1 = w->b1; a.b2 = w->b2; a.b3 = w->b3; a.rel_dist = rel_dist; a.w3c = w->w3c; a.cls = cls; a.triplet = triplet_dist;
  a.pred_rel = pred_rel; a.pred_conn = pred_conn; a.b3c = w->b3c; a.tau = tau;
  a.B = B; a.N = N; a.P = P; a.Lr = layers; a.ldu = ldu; a.k1 = k1;
  a.tiles_i = cdiv(N, 16); a.tiles_j = cdiv(N, 16);
  const bool big = P > 64;
  a.P3 = big ? 64 * cdiv(P, 64) : 64;  // rows of w3g
  CUtensorMap tw2, tw3, tu, tv;
  int rc = rh_map(w->w2g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 512, 512, 1, 1024, 0, 64, 128, 1, 1, &tw2);
  if (rc != EGTR_OK) return rc;
  rc = rh_map(w->w3g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 512, (unsigned long long)a.P3, 1, 1024, 0, 64, big ? (unsigned)a.P3 / 2 : 32u, 1, 1, &tw3);
  if (rc != EGTR_OK) return rc;
  const unsigned long long rows = (unsigned long long)B * N;
  rc = rh_map(U, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (unsigned long long)ldu, (unsigned long long)layers, rows, 4ull * ldu, 4ull * ldu * layers,
              32, (unsigned)layers, U_ROWS, 0, &tu);
  if (rc != EGTR_OK) return rc;
  rc = rh_map(V, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (unsigned long long)ldu, (unsigned long long)layers, rows, 4ull * ldu, 4ull * ldu * layers,
              32, (unsigned)layers, V_ROWS, 0, &tv);
  if (rc != EGTR_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    EGTR_CUDA(cudaFuncSetAttribute(relhead_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RhCfg<false>::SMEM));
    EGTR_CUDA(cudaFuncSetAttribute(relhead_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RhCfg<true>::SMEM));
    attr_set = true;
  }
  const long long items = 2ll * B * a.tiles_i * a.tiles_j;
  int slots = num_sms() / 2 / grid_div();
  if (slots < 1) slots = 1;
  const int grid = balanced_grid(items, slots) * 2;
  if (big) {
    EGTR_CUDA(launch_cluster_pdl(relhead_kernel<true>, dim3(grid), dim3(RH_THREADS), (size_t)RhCfg<true>::SMEM, (cudaStream_t)s, 2, tw2, tw3, tu,
                                 tv, a, rh_error_flag()));
  } else {
    EGTR_CUDA(launch_cluster_pdl(relhead_kernel<false>, dim3(grid), dim3(RH_THREADS), (size_t)RhCfg<false>::SMEM, (cudaStream_t)s, 2, tw2, tw3, tu,
                                 tv, a, rh_error_flag()));
  }
  count_launch();
  return EGTR_OK;
}

// The relation head behind one entry point: captured decoder self-attention queries / keys of every layer and
// the last hidden state in, pred_rel / pred_connectivity out.  Three launches: the 14 per-query projections (proj_q / proj_k /
// final_sub/obj_proj composed with layer 1 of both MLPs and the gate, egtr.py:336-397 -> U, V), the class argmax for the
// frequency bias, and the fused pair kernel above.
extern "C" int egtr_relation_head_fwd_f32(const float* const* q_ptrs, const float* const* k_ptrs, int ld_qk, const float* h_last, int ld_h,
                                          const float* logits, int K, const egtr_relhead_weights_t* w, const float* triplet_dist,
                                          const float* rel_dist, float tau, int use_freq_bias, int logit_adjustment, int B, int N, int P,
                                          float* U_scratch, float* V_scratch, int* cls_scratch, float* pred_rel, float* pred_conn,
                                          egtr_stream_t s) {
  EGTR_CHECK(q_ptrs && k_ptrs && h_last && w && U_scratch && V_scratch && pred_rel && pred_conn, EGTR_ERR_ARG,
             "egtr_relation_head_fwd_f32: null pointer");
  const int Lr = w->layers;
  EGTR_CHECK(Lr >= 1 && Lr <= MAX_LR && w->uv_planes && w->uv_bias, EGTR_ERR_ARG, "egtr_relation_head_fwd_f32: weights (layers=%d)", Lr);
  EGTR_CHECK(!use_freq_bias || (logits && triplet_dist && cls_scratch), EGTR_ERR_ARG, "egtr_relation_head_fwd_f32: frequency bias inputs missing");
  EGTR_CHECK(!logit_adjustment || rel_dist, EGTR_ERR_ARG, "egtr_relation_head_fwd_f32: rel_dist missing");
  const int G = 2 * Lr, ldu = 516;
  const float* ap[16];
  float* op[16];
  int nb[16], lda[16];
  for (int l = 0; l < Lr; ++l) {
    const bool last = l == Lr - 1;
    ap[l] = last ? h_last : q_ptrs[l];
    ap[Lr + l] = last ? h_last : k_ptrs[l];
    lda[l] = lda[Lr + l] = last ? ld_h : ld_qk;
    op[l] = U_scratch + l * ldu;
    op[Lr + l] = V_scratch + l * ldu;
    EGTR_CHECK(ap[l] && ap[Lr + l], EGTR_ERR_ARG, "egtr_relation_head_fwd_f32: layer %d pointer", l);
  }
  for (int g = 0; g < G; ++g) nb[g] = g * w->uv_npad;
  egtr_epilogue_t ep = {};
  ep.bias = w->uv_bias;
  ep.out = op[0];
  ep.ldo = ep.ldr = Lr * ldu;
  int rc = egtr_gemm_sbf16_grouped(ap, nullptr, op, nb, G, lda, w->uv_planes, G * w->uv_npad, B * N, 513, w->uv_npad, 256, &ep, s);
  if (rc != EGTR_OK) return rc;
  if (use_freq_bias) {
    rc = egtr_argmax_rows_f32(logits, K, B * N, cls_scratch, s);
    if (rc != EGTR_OK) return rc;
  }
  return egtr_relation_pairs_fused_f32(U_scratch, V_scratch, ldu, Lr, w, use_freq_bias ? cls_scratch : nullptr, triplet_dist, K + 1,
                                       logit_adjustment ? rel_dist : nullptr, tau, B, N, P, pred_rel, pred_conn, s);
}
