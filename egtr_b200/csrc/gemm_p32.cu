// D[M,N] = A[M,K] * W[N,K]^T (+bias, +residual, ReLU, row mask) on tcgen05 with BOTH operands fed by TMA.
//
// Activations between GEMM-class kernels live in HBM in the "P32" row format: an fp32 [rows, C] matrix
// stored with the same 4*C-byte row pitch, every group of 32 channels = 128 bytes = 32 bf16 hi values
// followed by 32 bf16 lo values (x ~= hi + lo to 2^-17, exactly the split the bf16x3 product scheme of
// gemm_tc.cu performs in registers).  Writing the split once in the producing kernel's epilogue means the
// consuming GEMM needs no operand-producer warps at all: one 128 x 128-byte SW128 TMA box per channel
// group lands hi in 16-byte chunks 0-3 and lo in chunks 4-7 of each smem row, and the UMMA descriptors of
// the hi / lo operands are the same tile at byte offsets 0 / 64 (+32 for the second 16-wide k-step).
//
// Convolutions use the same kernel: the activation map is 4-D {channels, W, H, B} over the NHWC P32 tensor and an M tile is
// a BW x BH patch of output pixels (BW * BH = 128), so filter tap (ky, kx) of a 64-channel slab is ONE tiled TMA box at
// pixel offset (w0*stride + kx - pad, h0*stride + ky - pad): the zero padding is TMA's out-of-bounds fill, the stride is
// the map's element stride, and there is no im2col buffer and no gather code.  Plain rows are the BW = 128, BH = 1 case.
//
// Roles (persistent, one CTA per SM, 320 threads):
//   warps 0-7  epilogue : two warps per TMEM lane quadrant, each owns half of the tile's columns;
//                         tcgen05.ld 32x32 -> bias / residual / ReLU / mask -> fp32 or P32 packing ->
//                         st.shared into a 4 KB SW128 staging box -> TMA store (clips the M tail);
//                         residual boxes are TMA-loaded into the same staging box beforehand.
//   warp  8    TMA      : per k-block 2 activation boxes + weight hi/lo boxes into the stage ring.
//   warp  9    MMA      : one lane issues 12 tcgen05.mma per k-block; owns the TMEM allocation
//                         (two accumulator stages so the epilogue of tile t overlaps the MMAs of t+1).
#include <cuda.h>

#include <mutex>
#include <string.h>
#include <unordered_map>

#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace egtr {

void count_launch();
int scratch_slot();

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int MAX_EPI_WARPS = 8;
constexpr int A_GROUP_BYTES = BLOCK_M * 128;  // one 32-channel group of 128 rows: hi 64 B | lo 64 B per row
constexpr int STG_BYTES = 32 * 128;           // one epilogue box: 32 rows x 32 channels

// CTAS == 2: a cluster of two CTAs computes a 256-row tile with one cta_group::2 MMA — each CTA stages its own 128 activation
// rows and only HALF of the weight tile, so a k-block costs 64 KB instead of 96 KB of L2->SM traffic per SM at BLOCK_N = 256
// and a third pipeline stage fits.
// (A weight-stationary variant — this CTA's half of an n-tile's weights resident across its m-tiles, activations alone streamed —
// was built and measured in round 1: 8-15 % slower than streaming both operands, because the resident region leaves 64 KB instead
// of 192 KB of loads in flight on a latency-bound operand path.  Removed in round 2.)
template <int BLOCK_N, int CTAS>
struct PCfg {
  static constexpr int EPI_WARPS = 8;
  static constexpr int TMA_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
  static constexpr int NUM_THREADS = (EPI_WARPS + 2) * 32;
  static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);
  static constexpr int B_ROWS = BLOCK_N / CTAS;  // weight rows staged by one CTA
  static constexpr int B_TILE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = 2 * A_GROUP_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGE_BUDGET = 192 * 1024;
  static constexpr int STAGES = STAGE_BUDGET / STAGE_BYTES > 4 ? 4 : STAGE_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int CTRL_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + CTRL_BYTES + 1024 /*align slack*/;
  static constexpr int CHUNKS = COLS_PER_WARP / 32;  // 32-column chunks per epilogue warp
  static_assert(STAGES >= 2, "operand ring needs at least two stages");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct PArgs {
  const float* bias;
  const uint8_t* row_keep;
  int M, N, K;
  int nb;                 // batches (images); tiles never straddle a batch
  int tiles_w, tiles_h;   // tiles per batch along w (rows mode: 128-row tiles) and h (rows mode: 1)
  int bw_log2;            // tile = BW x BH output pixels, BW = 1 << bw_log2, BH = 128 >> bw_log2 (rows mode: 128 x 1)
  int lim_w, lim_h;       // output extent per batch (rows mode: rows_per_b, 1)
  int stride, pad, kw, slabs;  // conv geometry; k-block kb = filter tap kb / slabs, 64-channel slab kb % slabs
  int keep_bstride, keep_off;
  int relu, out_fmt, res_fmt, has_res;
  int splits, kb_per_split, plane_rows;
  int ncols;              // columns that exist in the output map (N, or Npad for split-K partial sums)
  // LayerNorm epilogue (N == BLOCK_N == 256): out = LN(acc + bias + res) * gamma + beta; out2 = out + addend (optional)
  const float* ln_gamma;
  const float* ln_beta;
  const float* ln_addend;
  int ln, ln_out2, ln_ld;
  int dbg;                // egtr_set_debug_flags
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(ptx::smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int c, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tmap), "r"(smem_src), "r"(c), "r"(x), "r"(y), "r"(z) : "memory");
}

#ifdef EGTR_P32_PROF
// dev-only phase timestamps of CTA 0 (globaltimer, ns): [0] entry, [1] after setup sync + pdl wait, [2] first stage full (MMA warp),
// [3] first accumulator full (epilogue warp 0), [4] epilogue warp 0 done with its last store, [5] exit
__device__ unsigned long long g_p32_prof[8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define P32_STAMP(i) do { if (blockIdx.x == 0) g_p32_prof[i] = gtime(); } while (0)
#else
#define P32_STAMP(i)
#endif

struct TileCoord {
  int b, w0, h0, n0, sp;
};

template <int CTAS>
__device__ __forceinline__ void release_tmem_stage(uint64_t* bar, int lane) {  // whole warp, after its last tcgen05.wait::ld
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (CTAS == 2) ptx::mbar_arrive_leader(bar);
    else ptx::mbar_arrive(bar);
  }
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {  // low 16 bits = first element
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int BLOCK_N, int CTAS>
__global__ void __launch_bounds__((PCfg<BLOCK_N, CTAS>::NUM_THREADS), 1)
gemm_p32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                const __grid_constant__ CUtensorMap tmap_out2, const PArgs p, int* __restrict__ err) {
  if (!(p.dbg & 4)) pdl_launch_dependents();  // the next kernel may take SMs as this grid's CTAs retire
  if (threadIdx.x == 0) P32_STAMP(0);
  using C = PCfg<BLOCK_N, CTAS>;
  constexpr int EPI_WARPS = C::EPI_WARPS, TMA_WARP = C::TMA_WARP, MMA_WARP = C::MMA_WARP;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // operand stage ring
  uint8_t* stg_all = smem + C::STAGES * C::STAGE_BYTES;         // [EPI_WARPS][4096], 1024-aligned
  uint8_t* ctrl = stg_all + EPI_WARPS * STG_BYTES;
  uint64_t* full_bar = (uint64_t*)ctrl;          // [STAGES]
  uint64_t* empty_bar = full_bar + 4;            // [STAGES]
  uint64_t* tmem_full = empty_bar + 4;           // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* res_bar = tmem_empty + 2;            // [EPI_WARPS]
  uint32_t* tmem_holder = (uint32_t*)(res_bar + MAX_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? (int)ptx::cluster_ctarank() : 0;  // position in the CTA pair (0 = leader)
  const int mtiles_per_b = p.tiles_w * p.tiles_h;
  const int m_tiles = mtiles_per_b * p.nb;
  const int n_tiles = (p.ncols + BLOCK_N - 1) / BLOCK_N;
  // work item = (pair of consecutive m-tiles | one m-tile, n-tile, K split); CTA `rank` of a pair takes m-tile 2j + rank.  A
  // phantom odd tile decodes to batch index nb: its loads are out of bounds (zero fill) and it stores nothing.
  const int m_tiles2 = (m_tiles + CTAS - 1) / CTAS;
  const int total_all = m_tiles2 * n_tiles * p.splits;
  // items dealt round-robin, n fastest (neighbouring CTAs share activation rows in L2)
  const int slots = gridDim.x / CTAS, me = blockIdx.x / CTAS;
  const int w_first = me, total = total_all, w_step = slots;
  const int k_blocks_all = p.K / BLOCK_K;
  auto decode = [&](int w) {
    TileCoord t;
    const int tg = w / p.splits;
    t.sp = w - tg * p.splits;
    const int mt2 = tg / n_tiles;
    t.n0 = (tg - mt2 * n_tiles) * BLOCK_N;
    const int mt = mt2 * CTAS + rank;
    t.b = mt / mtiles_per_b;
    const int r = mt - t.b * mtiles_per_b;
    const int th = r / p.tiles_w;
    t.w0 = (r - th * p.tiles_w) << p.bw_log2;
    t.h0 = th * (BLOCK_M >> p.bw_log2);
    return t;
  };

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
    ptx::prefetch_tensormap(&tmap_out);
    if (p.has_res) ptx::prefetch_tensormap(&tmap_res);
    for (int i = 0; i < C::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);   // the TMA thread's expect_tx arrive
      ptx::mbar_init(&empty_bar[i], 1);  // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], EPI_WARPS * CTAS);  // lane 0 of every epilogue warp of the pair
    }
    for (int i = 0; i < EPI_WARPS; ++i) ptx::mbar_init(&res_bar[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    if (CTAS == 2) ptx::tmem_alloc_2cta<C::TMEM_COLS>(tmem_holder);
    else ptx::tmem_alloc<C::TMEM_COLS>(tmem_holder);
  }
  ptx::tc_fence_before();
  __syncwarp();
  if (CTAS == 2) ptx::cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();  // barriers, TMEM and tensor-map prefetch above overlapped the previous kernel's tail
  if (threadIdx.x == 0) P32_STAMP(1);
  const uint32_t tmem_base = *tmem_holder;

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------ TMA: activation groups + weight tiles
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int w = w_first; w < total; w += w_step) {
        const TileCoord t = decode(w);
        const int n0 = t.n0;
        const int kb_lo = t.sp * p.kb_per_split, kb_hi = min(k_blocks_all, kb_lo + p.kb_per_split);
        const int ax = t.w0 * p.stride - p.pad, ay = t.h0 * p.stride - p.pad;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          const int tap = kb / p.slabs, slab = kb - tap * p.slabs;
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 201);
          const uint32_t st = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          if (CTAS == 2) {
            // both CTAs' boxes are counted on the leader's barrier, which alone expects the pair's bytes
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            const int nr = n0 + rank * C::B_ROWS;
            ptx::tma_load_4d_2cta(st, &tmap_a, &full_bar[stage], slab * 128, ax + kx, ay + ky, t.b);
            ptx::tma_load_4d_2cta(st + A_GROUP_BYTES, &tmap_a, &full_bar[stage], slab * 128 + 64, ax + kx, ay + ky, t.b);
            ptx::tma_load_2d_2cta(st + 2 * A_GROUP_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, nr);
            ptx::tma_load_2d_2cta(st + 2 * A_GROUP_BYTES + C::B_TILE_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, p.plane_rows + nr);
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_4d(st, &tmap_a, &full_bar[stage], slab * 128, ax + kx, ay + ky, t.b);
            tma_load_4d(st + A_GROUP_BYTES, &tmap_a, &full_bar[stage], slab * 128 + 64, ax + kx, ay + ky, t.b);
            ptx::tma_load_2d(smem + stage * C::STAGE_BYTES + 2 * A_GROUP_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, n0);
            ptx::tma_load_2d(smem + stage * C::STAGE_BYTES + 2 * A_GROUP_BYTES + C::B_TILE_BYTES, &tmap_w, &full_bar[stage],
                             kb * BLOCK_K, p.plane_rows + n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (the pair's leader only)
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M * CTAS, BLOCK_N);
    int stage = 0, phase = 0, it = 0;
    for (int w = w_first; w < total; w += w_step, ++it) {
      const int sp = w % p.splits;
      const int kb_lo = sp * p.kb_per_split, kb_hi = min(k_blocks_all, kb_lo + p.kb_per_split);
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, err, 202);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase, err, 203);
        if (lane == 0 && it == 0 && kb == kb_lo) P32_STAMP(2);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_hi = a0 + 2 * A_GROUP_BYTES;
          const uint32_t b_lo = b_hi + C::B_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 16-wide k-steps: channel group ks>>1, half (ks&1) of its 32 channels
            const uint32_t at = a0 + (ks >> 1) * A_GROUP_BYTES + (ks & 1) * 32;
            const uint64_t dah = ptx::umma_desc_sw128(at), dal = ptx::umma_desc_sw128(at + 64);
            const uint64_t dbh = ptx::umma_desc_sw128(b_hi + ks * 32), dbl = ptx::umma_desc_sw128(b_lo + ks * 32);
            if (CTAS == 2) {
              ptx::umma_bf16_2cta(d_tmem, dal, dbh, idesc, (kb != kb_lo) || (ks != 0));
              ptx::umma_bf16_2cta(d_tmem, dah, dbl, idesc, 1);
              ptx::umma_bf16_2cta(d_tmem, dah, dbh, idesc, 1);
            } else {
              ptx::umma_bf16(d_tmem, dal, dbh, idesc, (kb != kb_lo) || (ks != 0));  // small terms first
              ptx::umma_bf16(d_tmem, dah, dbl, idesc, 1);
              ptx::umma_bf16(d_tmem, dah, dbh, idesc, 1);
            }
          }
          if (CTAS == 2) {
            ptx::umma_commit_2cta(&empty_bar[stage]);  // frees the stage in both CTAs
            if (kb == kb_hi - 1) ptx::umma_commit_2cta(&tmem_full[acc]);
          } else {
            ptx::umma_commit(&empty_bar[stage]);
            if (kb == kb_hi - 1) ptx::umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;    // TMEM lane quadrant = rows q*32 .. q*32+31 of the tile
    const int hf = warp >> 2;  // column half
    uint8_t* stg = stg_all + warp * STG_BYTES;
    const uint32_t stg_s = ptx::smem_u32(stg);
    const uint32_t my_row_s = stg_s + lane * 128;
    const int sw = lane & 7;
    uint64_t* rbar = &res_bar[warp];
    uint32_t res_phase = 0;
    int it = 0;
    for (int w = w_first; w < total; w += w_step, ++it) {
      const TileCoord t = decode(w);
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      // this warp's 32 tile rows = a (min(BW,32) x 32/min(BW,32)) box of output pixels starting at (ow, oh)
      const int ow = t.w0 + ((q * 32) & ((1 << p.bw_log2) - 1));
      const int oh = t.h0 + ((q * 32) >> p.bw_log2);
      const int zc = t.sp * p.nb + t.b;    // batch coordinate (split-K partial sums: one batch block per split)
      const int col_base = t.n0 + hf * C::COLS_PER_WARP;
      int nch = (p.ncols - col_base + 31) / 32;
      nch = nch < 0 ? 0 : (nch > C::CHUNKS ? C::CHUNKS : nch);
      if (ow >= p.lim_w || oh >= p.lim_h || t.b >= p.nb) nch = 0;  // box entirely in the tail (or a phantom tile): nothing to store
      bool keep = true;
      if (p.row_keep != nullptr && nch > 0 && ow + lane < p.lim_w)  // rows mode only
        keep = p.row_keep[(long long)t.b * p.keep_bstride + p.keep_off + ow + lane] != 0;
      if (p.has_res && nch > 0 && lane == 0) {  // first residual box of the tile, in flight while the MMAs run
        ptx::mbar_arrive_expect_tx(rbar, STG_BYTES);
        tma_load_4d(stg_s, &tmap_res, rbar, col_base, ow, oh, zc);
      }
      ptx::mbar_wait(&tmem_full[acc], acc_phase, err, 205);
      if (warp == 0 && lane == 0 && it == 0) P32_STAMP(3);
      ptx::tc_fence_after();
      if (nch == 0) {
        release_tmem_stage<CTAS>(&tmem_empty[acc], lane);
        continue;
      }
      uint32_t r[32];
      if (C::CHUNKS == 4 && p.ln) {
        // ---- Linear + residual + LayerNorm in one epilogue.  A row's 256 columns live in two warps (column halves): pass A
        // adds bias and residual, writes the sums back into the TMEM accumulator and reduces sum / sum of squares; the two
        // warps swap their partial sums through their staging tiles; pass B re-reads TMEM, normalises and stores P32 rows.
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + hf * C::COLS_PER_WARP;
        float s1 = 0.f, s2 = 0.f;
        ptx::tmem_ld_32x32(t_addr, r);
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int n = col_base + ci * 32;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(r[j]));
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (ci + 1 < 4) ptx::tmem_ld_32x32(t_addr + (ci + 1) * 32, r);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg((const float4*)(p.bias + n) + j);
              v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          }
          if (p.has_res) {
            ptx::mbar_wait(rbar, res_phase, err, 206);
            res_phase ^= 1;
            uint32_t x[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                           : "r"(my_row_s + ((c ^ sw) << 4)) : "memory");
            // The next residual box lands in the same staging tile.  Generic-proxy reads followed by an asynchronous-proxy write
            // of the same shared memory need a proxy fence (the plain path gets one from its output staging; this path requested
            // the next box right after issuing its ld.shared, which __syncwarp alone does not order: WARPSYNC does not wait for
            // the loads' data, and with a lightly loaded memory system the box can land inside a congested LDS latency).
            ptx::fence_proxy_async_smem();
            __syncwarp();  // every lane has read this residual box: the next one may land in the staging tile
            if (lane == 0 && ci + 1 < 4) {
              ptx::mbar_arrive_expect_tx(rbar, STG_BYTES);
              tma_load_4d(stg_s, &tmap_res, rbar, n + 32, ow, oh, zc);
            }
            if (p.res_fmt == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(x[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                v[2 * j] += __uint_as_float(x[j] << 16) + __uint_as_float(x[16 + j] << 16);
                v[2 * j + 1] += __uint_as_float(x[j] & 0xffff0000u) + __uint_as_float(x[16 + j] & 0xffff0000u);
              }
            }
          }
          uint32_t vb[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); vb[j] = __float_as_uint(v[j]); }
          ptx::tmem_st_32x32(t_addr + ci * 32, vb);
        }
        ptx::tmem_st_wait();
        {  // swap (sum, sum of squares) with the warp that owns the other 128 columns of the same 32 rows
          float2* mine = (float2*)stg;
          const float2* theirs = (const float2*)(stg_all + (warp ^ 4) * STG_BYTES);
          mine[lane] = make_float2(s1, s2);
          ptx::named_bar_sync(1 + q, 64);
          float2 o2 = theirs[lane];
          // the partner's staging tile is next written by ITS asynchronous-proxy traffic (residual boxes): order this generic-proxy
          // read before it explicitly (bar.sync orders the accesses of the participating threads, the proxy fence the proxies)
          asm volatile("" : "+f"(o2.x), "+f"(o2.y));  // the loaded values exist before the fence and the barrier are issued
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1 + q, 64);  // both warps have read before either overwrites its staging tile with output
          s1 += o2.x; s2 += o2.y;
        }
        const float mean = s1 * (1.f / 256.f);
        const float rstd = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mean * mean, 0.f) + 1e-5f);
        const long long grow = (long long)t.b * p.keep_bstride + p.keep_off + ow + lane;  // rows mode: this lane's output row
        const bool row_ok = ow + lane < p.lim_w;
        ptx::tmem_ld_32x32(t_addr, r);
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int n = col_base + ci * 32;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(r[j]));
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (ci + 1 < 4) ptx::tmem_ld_32x32(t_addr + (ci + 1) * 32, r);
          else release_tmem_stage<CTAS>(&tmem_empty[acc], lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg((const float4*)(p.ln_gamma + n) + j), b4 = __ldg((const float4*)(p.ln_beta + n) + j);
            v[4 * j] = (v[4 * j] - mean) * rstd * g4.x + b4.x; v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
            v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z; v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
          }
#pragma unroll 1
          for (int pass = 0; pass < (p.ln_out2 ? 2 : 1); ++pass) {
            if (pass == 1) {  // second output: + addend (this lane's own row, 128 contiguous bytes)
              const float4* ap = (const float4*)(p.ln_addend + grow * p.ln_ld + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 a4 = row_ok ? __ldg(ap + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * j] += a4.x; v[4 * j + 1] += a4.y; v[4 * j + 2] += a4.z; v[4 * j + 3] += a4.w;
              }
            }
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t h = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              const float l0 = v[2 * j] - __uint_as_float(h << 16), l1 = v[2 * j + 1] - __uint_as_float(h & 0xffff0000u);
              o[j] = h;
              o[16 + j] = pack_bf16x2(l0, l1);
            }
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_s + ((c ^ sw) << 4)), "r"(o[4 * c]), "r"(o[4 * c + 1]),
                           "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(pass == 0 ? &tmap_out : &tmap_out2, stg_s, n, ow, oh, zc);
              bulk_commit();
            }
          }
        }
        if (lane == 0) bulk_wait_read0();  // the next tile's first residual box lands in the staging tile
        __syncwarp();
        continue;
      }
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + hf * C::COLS_PER_WARP, r);
#pragma unroll 1
      for (int ci = 0; ci < nch; ++ci) {
        const int n = col_base + ci * 32;
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(r[j]));  // uses of r must not be scheduled above the wait
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ci + 1 < nch) {  // next chunk's accumulator read overlaps this chunk's arithmetic and store
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + hf * C::COLS_PER_WARP + (ci + 1) * 32, r);
        } else {  // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp right away
          release_tmem_stage<CTAS>(&tmem_empty[acc], lane);
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg((const float4*)(p.bias + n) + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        if (p.has_res) {
          ptx::mbar_wait(rbar, res_phase, err, 206);
          res_phase ^= 1;
          uint32_t x[32];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                         : "r"(my_row_s + ((c ^ sw) << 4)) : "memory");
          if (p.res_fmt == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(x[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[2 * j] += __uint_as_float(x[j] << 16) + __uint_as_float(x[16 + j] << 16);
              v[2 * j + 1] += __uint_as_float(x[j] & 0xffff0000u) + __uint_as_float(x[16 + j] & 0xffff0000u);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (!keep) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (p.out_fmt == 2) {
          // H16 pair records (include/egtr_b200.h): this chunk = 32 tokens x one head; the 64-byte fp16 row of token t goes to
          // slot 0 of record t + 1 and to slot 1 of record t — two TMA stores of the same staging box (SWIZZLE_64B rows)
          uint32_t hh[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            hh[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          const uint32_t row_s = stg_s + lane * 64;
          const int sw2 = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_s + ((c ^ sw2) << 4)), "r"(hh[4 * c]), "r"(hh[4 * c + 1]),
                         "r"(hh[4 * c + 2]), "r"(hh[4 * c + 3]) : "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int head = zc * (p.ncols >> 5) + (n >> 5);
            tma_store_4d(&tmap_out, stg_s, 0, 0, ow + 1, head);
            tma_store_4d(&tmap_out, stg_s, 0, 1, ow, head);
            bulk_commit();
          }
          continue;
        }
        uint32_t o[32];
        if (p.out_fmt == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t h = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            const float l0 = v[2 * j] - __uint_as_float(h << 16), l1 = v[2 * j + 1] - __uint_as_float(h & 0xffff0000u);
            o[j] = h;
            o[16 + j] = pack_bf16x2(l0, l1);
          }
        }
        if (!p.has_res) {
          // the previous store's read of the staging box was left in flight behind this chunk's arithmetic
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_s + ((c ^ sw) << 4)), "r"(o[4 * c]), "r"(o[4 * c + 1]),
                       "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_out, stg_s, n, ow, oh, zc);
          bulk_commit();
          if (p.has_res) {
            bulk_wait_read0();  // the staging box doubles as the residual landing zone: it must be free before the next load
            if (ci + 1 < nch) {
              ptx::mbar_arrive_expect_tx(rbar, STG_BYTES);
              tma_load_4d(stg_s, &tmap_res, rbar, n + 32, ow, oh, zc);
            }
          }
        }
        if (p.has_res) __syncwarp();
      }
    }
    if (lane == 0) {  // all stores of this warp have been read out of shared memory (global visibility: kernel end)
      if (p.dbg & 2) bulk_wait0();
      else bulk_wait_read0();
    }
    if (warp == 0 && lane == 0) P32_STAMP(4);
  }

  ptx::tc_fence_before();
  __syncwarp();  // single-lane roles reconverge before the aligned barrier
  if (CTAS == 2) ptx::cluster_sync_all();  // the peer may still signal this CTA's barriers / read its operand tiles
  else __syncthreads();
  if (warp == MMA_WARP) {
    ptx::tc_fence_after();
    if (CTAS == 2) ptx::tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
    else ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
    if (lane == 0) P32_STAMP(5);
  }
}

// --------------------------------------------------------------------------- split-K reduction
// out = epilogue(sum over splits of partial[s][m][n]); one warp per (row, 32-channel group), lane = channel.
struct ReduceArgs {
  const float* partial;
  int splits, M, N, Npad;
  const float* bias;
  const void* res;
  void* out;
  int ldo, ldr, relu, out_fmt, res_fmt;
  int rows_per_b, bstride, off;  // output row of GEMM row m (0: identity), as in egtr_epilogue_t
  const uint8_t* row_keep;
};

__device__ __forceinline__ float p32_load(const void* base, long long row, int ld, int col) {
  const __nv_bfloat16* g = (const __nv_bfloat16*)((const uint8_t*)base + (row * ld + (col & ~31)) * 4);
  return __bfloat162float(g[col & 31]) + __bfloat162float(g[32 + (col & 31)]);
}

__global__ void __launch_bounds__(256)
p32_reduce_kernel(const ReduceArgs a) {
  pdl_entry();
  // one thread = four consecutive channels of one row: float4 partial-sum loads, four splits in flight at a time
  const int n4 = a.N >> 2;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)a.M * n4) return;
  const long long m = i / n4;
  const int n = (int)(i - m * n4) * 4;
  const float* src = a.partial + m * a.Npad + n;
  const long long sstride = (long long)a.M * a.Npad;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  for (; s + 4 <= a.splits; s += 4) {
    const float4 v0 = *(const float4*)(src + (s + 0) * sstride), v1 = *(const float4*)(src + (s + 1) * sstride);
    const float4 v2 = *(const float4*)(src + (s + 2) * sstride), v3 = *(const float4*)(src + (s + 3) * sstride);
    acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
    acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
  }
  for (; s < a.splits; ++s) {
    const float4 v = *(const float4*)(src + s * sstride);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  long long orow = m;
  if (a.rows_per_b > 0) {
    const long long b = m / a.rows_per_b;
    orow = b * a.bstride + a.off + (m - b * a.rows_per_b);
  }
  if (a.bias) {
    const float4 b4 = __ldg((const float4*)(a.bias + n));
    acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
  }
  if (a.res) {
    if (a.res_fmt) {
      acc.x += p32_load(a.res, orow, a.ldr, n); acc.y += p32_load(a.res, orow, a.ldr, n + 1);
      acc.z += p32_load(a.res, orow, a.ldr, n + 2); acc.w += p32_load(a.res, orow, a.ldr, n + 3);
    } else {
      const float4 r4 = *(const float4*)((const float*)a.res + orow * a.ldr + n);
      acc.x += r4.x; acc.y += r4.y; acc.z += r4.z; acc.w += r4.w;
    }
  }
  if (a.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
  if (a.row_keep && !a.row_keep[orow]) acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.out_fmt == 0) {
    *(float4*)((float*)a.out + orow * a.ldo + n) = acc;
  } else {
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    split_bf16(acc.x, h0, l0); split_bf16(acc.y, h1, l1); split_bf16(acc.z, h2, l2); split_bf16(acc.w, h3, l3);
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
    uint8_t* g = (uint8_t*)a.out + (orow * a.ldo + (n & ~31)) * 4 + (n & 31) * 2;
    *(uint2*)g = ph;
    *(uint2*)(g + 64) = pl;
  }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  });
  return fn;
}

struct MapDesc {  // everything that determines a tensor map (POD, zero-initialised, compared bytewise)
  const void* ptr;
  unsigned long long dim[4], stride[3];
  unsigned box[4], estr[4];
  int dtype, rank;
  int swizzle64;  // 0: SWIZZLE_128B (every operand / fp32 / P32 map), 1: SWIZZLE_64B (the 64-byte rows of the H16 pair-record output)
};
struct MapDescHash {
  size_t operator()(const MapDesc& d) const {
    const unsigned char* b = (const unsigned char*)&d;
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapDesc); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
};
struct MapDescEq {
  bool operator()(const MapDesc& a, const MapDesc& b) const { return memcmp(&a, &b, sizeof(MapDesc)) == 0; }
};

// Tensor maps are immutable per (pointer, geometry): encode once, reuse for every launch (workspaces are persistent).
int cached_map(const MapDesc& d, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapDesc, CUtensorMap, MapDescHash, MapDescEq> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(d);
  if (it != cache.end()) {
    *out = it->second;
    return EGTR_OK;
  }
  EncodeTiledFn enc = encode_fn();
  EGTR_CHECK(enc != nullptr, EGTR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4] = {d.dim[0], d.dim[1], d.dim[2], d.dim[3]};
  cuuint64_t gstride[3] = {d.stride[0], d.stride[1], d.stride[2]};
  cuuint32_t box[4] = {d.box[0], d.box[1], d.box[2], d.box[3]};
  cuuint32_t estr[4] = {d.estr[0], d.estr[1], d.estr[2], d.estr[3]};
  CUtensorMap m;
  static const int l2promo = [] { const char* e = getenv("EGTR_TMA_L2PROMO"); return e ? atoi(e) : 3; }();  // dev: 0 none, 1 64B, 2 128B, 3 256B
  const CUtensorMapL2promotion promo = l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(&m, (CUtensorMapDataType)d.dtype, (cuuint32_t)d.rank, const_cast<void*>(d.ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, d.swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGTR_CHECK(r == CUDA_SUCCESS, EGTR_ERR_CUDA,
             "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u,%u)",
             (int)r, d.rank, d.dim[0], d.dim[1], d.dim[2], d.dim[3], d.stride[0], d.stride[1], d.stride[2], d.box[0], d.box[1],
             d.box[2], d.box[3]);
  if (cache.size() >= (1u << 15)) cache.clear();  // bounded: maps are passed to kernels by value, re-encoding is cheap
  cache.emplace(d, m);
  *out = m;
  return EGTR_OK;
}

// 4-D map {d0, d1, d2, d3} (d0 innermost, contiguous), byte strides s1..s3 of dims 1..3, box {b0, b1, b2, 1},
// element (traversal) strides {1, e, e, 1}.
MapDesc desc4(const void* ptr, int dtype, unsigned long long d0, unsigned long long d1, unsigned long long d2, unsigned long long d3,
              unsigned long long s1, unsigned long long s2, unsigned long long s3, unsigned b0, unsigned b1, unsigned b2, unsigned e = 1) {
  MapDesc d;
  memset(&d, 0, sizeof(d));
  d.ptr = ptr; d.dtype = dtype; d.rank = 4;
  d.dim[0] = d0; d.dim[1] = d1; d.dim[2] = d2; d.dim[3] = d3;
  d.stride[0] = s1; d.stride[1] = s2; d.stride[2] = s3;
  d.box[0] = b0; d.box[1] = b1; d.box[2] = b2; d.box[3] = 1;
  d.estr[0] = 1; d.estr[1] = e; d.estr[2] = e; d.estr[3] = 1;
  return d;
}
MapDesc desc2(const void* ptr, int dtype, unsigned long long d0, unsigned long long d1, unsigned long long s1, unsigned b0, unsigned b1) {
  MapDesc d;
  memset(&d, 0, sizeof(d));
  d.ptr = ptr; d.dtype = dtype; d.rank = 2;
  d.dim[0] = d0; d.dim[1] = d1;
  d.stride[0] = s1;
  d.box[0] = b0; d.box[1] = b1;
  d.estr[0] = d.estr[1] = 1;
  return d;
}

int* device_error_flag_p32() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

// Grow-only per scratch slot.  A replaced buffer is NOT freed: CUDA graphs captured earlier hold its address.  Growth is geometric
// (x1.5 over the request), so everything ever left behind by a slot sums to less than twice its final capacity — bounded by the
// largest shape the process sees, not by the number of shapes.
float* partial_buffer_p32(size_t floats) {
  static float* buf[32] = {};
  static size_t cap[32] = {};
  const int slot = scratch_slot();
  if (floats > cap[slot]) {
    float* nb = nullptr;
    const size_t want = floats + floats / 2;
    if (cudaMalloc(&nb, want * sizeof(float)) != cudaSuccess) return nullptr;
    buf[slot] = nb;
    cap[slot] = want;
  }
  return buf[slot];
}

// Output pixel tile of a convolution: BW x BH = 128 with BW a power of two in [8, 128]; fewest tiles wins, wider wins ties.
int pick_bw_log2(int OW, int OH) {
  int best = 7;
  long long best_tiles = -1;
  for (int l = 7; l >= 3; --l) {
    const int bw = 1 << l, bh = BLOCK_M >> l;
    const long long tiles = (long long)cdiv(OW, bw) * cdiv(OH, bh);
    if (best_tiles < 0 || tiles < best_tiles) { best_tiles = tiles; best = l; }
  }
  return best;
}

template <int BLOCK_N, int CTAS>
int launch_p32(const ASrc& a, const void* planes, int plane_rows, int M, int N, int Npad, int K, const Epilogue& ep,
               cudaStream_t st) {
  using C = PCfg<BLOCK_N, CTAS>;
  const bool conv = a.mode == 1;
  PArgs p = {};
  p.M = M; p.N = N; p.K = K;
  p.plane_rows = plane_rows;
  int rows_per_b, out_w, out_h;  // per batch: rows, output extent
  if (conv) {
    p.nb = M / (a.OH * a.OW);
    rows_per_b = a.OH * a.OW; out_w = a.OW; out_h = a.OH;
    p.bw_log2 = pick_bw_log2(a.OW, a.OH);
    p.stride = a.stride; p.pad = a.pad; p.kw = a.KW; p.slabs = a.C / 64;
  } else {
    rows_per_b = ep.rows_per_b > 0 ? ep.rows_per_b : M;
    p.nb = M / rows_per_b;
    out_w = rows_per_b; out_h = 1;
    p.bw_log2 = 7;
    p.stride = 1; p.pad = 0; p.kw = 1; p.slabs = K / 64;
  }
  const int BW = 1 << p.bw_log2, BH = BLOCK_M >> p.bw_log2;
  p.tiles_w = cdiv(out_w, BW); p.tiles_h = cdiv(out_h, BH);
  p.lim_w = out_w; p.lim_h = out_h;
  const int m_tiles = p.tiles_w * p.tiles_h * p.nb;
  const int n_tiles = cdiv(N, BLOCK_N);
  const int k_blocks = K / BLOCK_K;
  // split-K spreads few-tile / long-K problems over the SMs for latency; capped by egtr_set_splitk_max (default 1 = off)
  // ... except that a handful of CTAs never walk more than 32 k-blocks alone (the extra level's 3x3/2 conv on C5: 273 rows, K = 18432)
  int splitk_cap = splitk_max();
  if (k_blocks / 32 > splitk_cap) splitk_cap = k_blocks / 32;
  int splits = 1;
  if (splitk_cap > 1 && m_tiles * n_tiles * 2 <= num_sms() && k_blocks >= 8) {
    splits = num_sms() / (m_tiles * n_tiles);
    if (splits > k_blocks / 4) splits = k_blocks / 4;
    if (splits > splitk_cap) splits = splitk_cap;
    if (splits < 1) splits = 1;
  }
  if (ep.ln_gamma != nullptr || ep.out_fmt == EGTR_FMT_H16PAIR) splits = 1;  // (the LayerNorm epilogue needs the complete row sums in one tile)
  const int kbps = cdiv(k_blocks, splits);
  splits = cdiv(k_blocks, kbps);
  p.splits = splits; p.kb_per_split = kbps;
  p.dbg = debug_flags();

  CUtensorMap ta, tw, to, tr, to2;
  int rc;
  if (conv) {
    const unsigned long long pix = 4ull * a.C;  // bytes per pixel
    rc = cached_map(desc4(a.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2ull * a.C, a.W, a.H, p.nb, pix, pix * a.W, pix * a.W * a.H, 64,
                          BW * a.stride, BH * a.stride, a.stride), &ta);
  } else {
    const unsigned long long pitch = 4ull * a.lda;
    rc = cached_map(desc4(a.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2ull * K, rows_per_b, 1, p.nb, pitch, pitch * rows_per_b,
                          pitch * rows_per_b, 64, BLOCK_M, 1), &ta);
  }
  if (rc != EGTR_OK) return rc;
  rc = cached_map(desc2(planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, K, 2ull * plane_rows, 2ull * K, BLOCK_K, C::B_ROWS), &tw);
  if (rc != EGTR_OK) return rc;
  const unsigned box_w = BW < 32 ? BW : 32, box_h = 32 / box_w;  // one epilogue warp's 32 tile rows
  float* partial = nullptr;
  if (splits > 1) {
    partial = partial_buffer_p32((size_t)splits * M * Npad);
    EGTR_CHECK(partial != nullptr, EGTR_ERR_CUDA, "split-K scratch allocation failed");
    const unsigned long long pitch = 4ull * Npad;
    rc = cached_map(desc4(partial, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, Npad, out_w, out_h, (unsigned long long)p.nb * splits, pitch,
                          pitch * out_w, pitch * rows_per_b, 32, box_w, box_h), &to);
    if (rc != EGTR_OK) return rc;
    tr = to;
    to2 = to;
    p.ncols = Npad;
  } else {
    const unsigned long long bstride = ep.rows_per_b > 0 ? (unsigned long long)ep.bstride : (unsigned long long)rows_per_b;
    const uint8_t* obase = (const uint8_t*)ep.out + 4ll * ep.off * ep.ldo;
    const unsigned long long opitch = 4ull * ep.ldo;
    if (ep.out_fmt == EGTR_FMT_H16PAIR) {
      // [nb * N/32 heads][rows_per_b + 1 records][2 slots][32 fp16]: box = 32 records of one slot
      MapDesc d = desc4(ep.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 32, 2, (unsigned long long)rows_per_b + 1,
                        (unsigned long long)p.nb * (N / 32), 64, 128, 128ull * (rows_per_b + 1), 32, 1, 32);
      d.swizzle64 = 1;
      rc = cached_map(d, &to);
    } else {
      rc = cached_map(desc4(obase, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, N, out_w, out_h, p.nb, opitch, opitch * out_w, opitch * bstride, 32,
                            box_w, box_h), &to);
    }
    if (rc != EGTR_OK) return rc;
    tr = to;
    if (ep.res != nullptr) {
      const uint8_t* rbase = (const uint8_t*)ep.res + 4ll * ep.off * ep.ldr;
      const unsigned long long rpitch = 4ull * ep.ldr;
      rc = cached_map(desc4(rbase, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, N, out_w, out_h, p.nb, rpitch, rpitch * out_w, rpitch * bstride, 32,
                            box_w, box_h), &tr);
      if (rc != EGTR_OK) return rc;
      p.has_res = 1;
      p.res_fmt = ep.res_fmt;
    }
    to2 = to;
    if (ep.ln_gamma != nullptr) {
      p.ln = 1;
      p.ln_gamma = ep.ln_gamma;
      p.ln_beta = ep.ln_beta;
      p.ln_ld = ep.ldo;
      if (ep.ln_out2 != nullptr) {
        const uint8_t* o2base = (const uint8_t*)ep.ln_out2 + 4ll * ep.off * ep.ldo;
        rc = cached_map(desc4(o2base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, N, out_w, out_h, p.nb, opitch, opitch * out_w, opitch * bstride, 32,
                              box_w, box_h), &to2);
        if (rc != EGTR_OK) return rc;
        p.ln_out2 = 1;
        p.ln_addend = ep.ln_addend;
      }
    }
    p.bias = ep.bias;
    p.row_keep = conv ? nullptr : ep.row_keep;
    p.keep_bstride = (int)bstride;
    p.keep_off = ep.off;
    p.relu = ep.relu;
    p.out_fmt = ep.out_fmt;
    p.ncols = N;
  }
  static bool attr_set = false;  // one flag per instantiation
  if (!attr_set) {
    EGTR_CUDA(cudaFuncSetAttribute(gemm_p32_kernel<BLOCK_N, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int work = cdiv(m_tiles, CTAS) * cdiv(p.ncols, BLOCK_N) * splits;  // per CTA (CTAS == 1) or per CTA pair
  // throughput mode (egtr_set_grid_div): the persistent grid takes 1/div of the GPU, so that GEMMs of the other forwards in
  // flight run beside it on disjoint SMs instead of time-slicing the whole machine
  const int grid = balanced_grid(work, num_sms() / CTAS / grid_div()) * CTAS;
  EGTR_CUDA(launch_cluster_pdl(gemm_p32_kernel<BLOCK_N, CTAS>, dim3(grid), dim3(C::NUM_THREADS), (size_t)(C::SMEM_BYTES), st, CTAS, ta, tw, to, tr, to2, p,
                       device_error_flag_p32()));
  if (splits > 1) {
    ReduceArgs r = {};
    r.partial = partial; r.splits = splits; r.M = M; r.N = N; r.Npad = Npad;
    r.bias = ep.bias; r.res = ep.res; r.out = ep.out; r.ldo = ep.ldo; r.ldr = ep.ldr; r.relu = ep.relu;
    r.out_fmt = ep.out_fmt; r.res_fmt = ep.res_fmt; r.row_keep = conv ? nullptr : ep.row_keep;
    r.rows_per_b = ep.rows_per_b; r.bstride = ep.bstride; r.off = ep.off;
    const long long threads = (long long)M * (N / 4);
    launch_pdl(p32_reduce_kernel, dim3(cdiv(threads, 256)), dim3(256), (size_t)(0), st, r);
    count_launch();
    EGTR_CUDA(cudaGetLastError());
  }
  return EGTR_OK;
}

}  // namespace

// Tensor maps for the other TMA-fed kernels of the library (decoder.cu), from the same cache.
// P32 rows [nb][rows_per_b][channels] (fp32 pitch): box = 64 bf16 (one group's hi | lo = 128 bytes) x box_rows rows; rows beyond
// rows_per_b of an image read as zeros.
int tmap_p32_rows(const void* ptr, int channels, int rows_per_b, int nb, int box_rows, CUtensorMap* out) {
  const unsigned long long pitch = 4ull * channels;
  return cached_map(desc4(ptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2ull * channels, rows_per_b, 1, nb, pitch, pitch * rows_per_b,
                          pitch * rows_per_b, 64, box_rows, 1), out);
}
// bf16 matrix [rows][K] (K-major: weight planes, transposed P32 operands): box = 64 elements x box_rows rows.
int tmap_weight_planes(const void* planes, int K, long long rows, int box_rows, CUtensorMap* out) {
  return cached_map(desc2(planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, K, rows, 2ull * K, BLOCK_K, box_rows), out);
}

// Any tiled map (rank <= 4, unit element strides) through the same mutex-protected cache: stem.cu's overlapping-row image map.
int tmap_tiled(const void* ptr, int dtype, int rank, const unsigned long long* dims, const unsigned long long* strides, const unsigned* box,
               int swizzle64, CUtensorMap* out) {
  MapDesc d;
  memset(&d, 0, sizeof(d));
  d.ptr = ptr; d.dtype = dtype; d.rank = rank; d.swizzle64 = swizzle64;
  for (int i = 0; i < rank; ++i) { d.dim[i] = dims[i]; d.box[i] = box[i]; d.estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) d.stride[i] = strides[i];
  return cached_map(d, out);
}

// Entry used by egtr_gemm_sbf16 when the operand source is P32 (a.fmt == 1): mode 0 rows or mode 1 NHWC convolution.
int gemm_p32_dispatch(const ASrc& a, const void* planes, int plane_rows, int M, int N, int Npad, int K, const Epilogue& ep,
                      cudaStream_t st) {
  EGTR_ONE_DEVICE();
  EGTR_CHECK(a.a2 == nullptr && (a.mode == 0 || a.mode == 1), EGTR_ERR_UNSUPPORTED,
             "P32 operand: plain rows or NHWC convolution only (fold addends into the producer)");
  EGTR_CHECK(N % 32 == 0 && K % 64 == 0 && ep.ldo % 4 == 0 && ep.ldo >= N, EGTR_ERR_ARG,
             "egtr_gemm_sbf16 (P32): need N %% 32 == 0, K %% 64 == 0, 16-byte row pitches (N=%d K=%d ldo=%d)", N, K, ep.ldo);
  if (a.mode == 0) {
    EGTR_CHECK(a.lda % 4 == 0 && a.lda >= K, EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32 rows): lda=%d", a.lda);
    EGTR_CHECK(ep.rows_per_b <= 0 || M % ep.rows_per_b == 0, EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32 rows): M %% rows_per_b != 0");
  } else {
    EGTR_CHECK(a.C % 64 == 0 && K == a.KH * a.KW * a.C && a.stride >= 1 && a.stride <= 2 && a.OH > 0 && a.OW > 0 &&
                   M % (a.OH * a.OW) == 0, EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32 conv): C=%d K=%d stride=%d M=%d", a.C, K, a.stride, M);
    EGTR_CHECK(ep.rows_per_b <= 0 || ep.rows_per_b == a.OH * a.OW, EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32 conv): rows_per_b must be OH*OW");
  }
  EGTR_CHECK(((uintptr_t)a.a & 127) == 0 && ((uintptr_t)ep.out & 127) == 0 && (!ep.res || ((uintptr_t)ep.res & 127) == 0) && ((uintptr_t)planes & 127) == 0,
             EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32): 128-byte aligned buffers required");
  EGTR_CHECK(!ep.res || (ep.ldr % 4 == 0 && ep.ldr >= N), EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32): ldr=%d", ep.ldr);
  EGTR_CHECK(ep.pair_n == 0 && !ep.fin && !ep.dot_w, EGTR_ERR_UNSUPPORTED, "egtr_gemm_sbf16 (P32): relation epilogues are not built here");
  EGTR_CHECK(ep.out_fmt != EGTR_FMT_H16PAIR || (a.mode == 0 && !ep.res && !ep.ln_gamma && ep.rows_per_b <= 0 && !ep.relu), EGTR_ERR_UNSUPPORTED,
             "egtr_gemm_sbf16 (P32): H16 pair-record output takes plain rows (one batch = M rows), bias and row_keep only");
  static const int forced_bn = [] { const char* e = getenv("EGTR_GEMM_BLOCK_N"); return e ? atoi(e) : 0; }();  // dev experiments only
  int bn = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : (N >= 192 ? 128 : 64));
  if (bn == 256) {
    // wave quantisation: a persistent grid of 74 CTA pairs runs ceil(items / 74) rounds of (bn + fixed) cost each; 87 pair
    // items of 256 columns (the encoder's 22 223 tokens, N = 256) take two rounds, 174 items of 128 columns take three halves
    const long long pairs_m = cdiv(cdiv(M, BLOCK_M), 2), slots = num_sms() / 2 / grid_div() > 0 ? num_sms() / 2 / grid_div() : 1;
    const long long c256 = cdiv(pairs_m * (N / 256), slots) * (256 + 32), c128 = cdiv(pairs_m * (N / 128), slots) * (128 + 32);
    if (c128 < c256) bn = 128;
  }
  if (forced_bn == 64 || forced_bn == 128 || forced_bn == 256) bn = forced_bn;
  if (ep.ln_gamma != nullptr) {
    EGTR_CHECK(N == 256 && a.mode == 0 && ep.ln_beta != nullptr && ep.out_fmt == EGTR_FMT_P32 && ep.row_keep == nullptr &&
                   (!ep.ln_out2 || ep.ln_addend),
               EGTR_ERR_ARG, "egtr_gemm_sbf16 (P32): the LayerNorm epilogue needs N == 256, plain rows, P32 output");
    bn = 256;  // the whole row in one tile
  }
  // CTA pairs (cta_group::2) for every 128/256-wide tile shape with at least two m-tiles: measured through the whole forward,
  // pairs everywhere beat both 1-CTA tiles and a size threshold (less L2->SM weight traffic, one more pipeline stage)
  static const int forced_ctas = [] { const char* e = getenv("EGTR_GEMM_CTAS"); return e ? atoi(e) : 0; }();  // dev experiments only
  const long long tiles = (long long)cdiv(M, BLOCK_M) * cdiv(N, bn);
  bool pair = bn >= 128 && tiles >= 2;
  if (forced_ctas == 1) pair = false;
  if (forced_ctas == 2 && bn >= 128) pair = true;
  if (pair) {
    if (bn == 256) return launch_p32<256, 2>(a, planes, plane_rows, M, N, Npad, K, ep, st);
    return launch_p32<128, 2>(a, planes, plane_rows, M, N, Npad, K, ep, st);
  }
  if (bn == 256) return launch_p32<256, 1>(a, planes, plane_rows, M, N, Npad, K, ep, st);
  if (bn == 128) return launch_p32<128, 1>(a, planes, plane_rows, M, N, Npad, K, ep, st);
  return launch_p32<64, 1>(a, planes, plane_rows, M, N, Npad, K, ep, st);
}

}  // namespace egtr

#ifdef EGTR_P32_PROF
extern "C" int egtr_debug_p32_prof(unsigned long long* host_out) {
  EGTR_CUDA(cudaMemcpyFromSymbol(host_out, egtr::g_p32_prof, sizeof(unsigned long long) * 8));
  return EGTR_OK;
}
#endif
