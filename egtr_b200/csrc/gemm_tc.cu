// D[M,N] = A[M,K] * W[N,K]^T (+bias, +residual, ReLU) on Blackwell tensor cores.
//
// Precision: the reference path is fp32 end to end and parity is 1e-3 on the final outputs.  A single
// bf16 or tf32 product per term measurably fails that (tools/precision_probe.py: 9e-1 / 3e-3 on
// pred_rel); three bf16 products per term  a*w ~= ah*wh + ah*wl + al*wh  (a = ah + al, w = wh + wl)
// reproduce fp32 to ~5e-5.  So every k-step issues three tcgen05.mma (kind::f16, BF16 inputs, FP32
// accumulate in TMEM) — which costs nothing extra against a TF32 kernel because at 128x256 tiles both
// are bound by operand delivery from L2, not by the tensor pipe.
//
// Structure: persistent, warp-specialised, one CTA per SM.
//   warps 0-3  epilogue : tcgen05.ld accumulator rows -> bias/residual/ReLU -> global fp32
//   warp  4    TMA      : weight tiles (hi and lo planes, pre-split once at load) -> 128B-swizzled smem
//   warp  5    MMA      : one lane issues tcgen05.mma; owns the TMEM allocation
//   warps 6-13 A-producer: two groups of 4 warps take alternate k-blocks, so two k-blocks of global loads are
//                          always in flight per CTA; each: LDG fp32 activation rows (plain rows, x+pos, or an implicit-im2col gather
//                          for convolutions), split to bf16 hi/lo in registers, st.shared into the same
//                          swizzled K-major layout TMA would have produced — activations stay fp32 in
//                          HBM and there is never an im2col buffer.
// Pipelines: smem stage full/empty mbarriers (producers+TMA -> MMA), two TMEM accumulator stages
// with full/empty mbarriers (MMA -> epilogue) so the epilogue of tile t overlaps the MMAs of tile t+1.
#include <cuda.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace egtr {

void count_launch();
int scratch_slot();
int gemm_p32_dispatch(const ASrc& a, const void* planes, int plane_rows, int M, int N, int Npad, int K, const Epilogue& ep,
                      cudaStream_t st);  // gemm_p32.cu: TMA-fed operand rows

namespace {

#ifdef EGTR_GEMM_PROF
// dev-only cycle accounting per role (tools/gemm_bench.py --prof): [cta][slot]
__device__ unsigned long long g_prof[148][16];
#define PROF_T0(var) long long var = clock64()
#define PROF_ADD(slot, var) do { if (lane == 0) g_prof_local[slot] += clock64() - (var); } while (0)
#else
#define PROF_T0(var)
#define PROF_ADD(slot, var)
#endif

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 448;  // 4 epilogue + TMA + MMA + 2 x 4 producer warps
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // one bf16 plane
constexpr int STG_LD = 36;                           // floats per row of an epilogue transpose tile (32 + 4 pad)

template <int BLOCK_N>
struct Cfg {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 2 : (BLOCK_N == 128 ? 3 : 4);
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 4608 /*barriers, row info*/ +
                                    4 * 32 * 36 * 4 /*epilogue transpose tiles*/ + 2 * 128 * 8 * 4 /*pair-stage gate tables*/;
};

// Grouped launch: `groups` independent GEMMs of identical (M, N, K) in one grid — different left operands,
// different weight rows (all groups' planes stacked in one [plane_rows, K] matrix) and different outputs.
// Used for the seven relation-head projections per side and for the decoder's q|k / v pair.
constexpr int MAX_GROUPS = 16;
struct GroupTab {
  const float* a[MAX_GROUPS];
  const float* a2[MAX_GROUPS];
  float* out[MAX_GROUPS];
  int n_base[MAX_GROUPS];  // first weight/bias row of the group inside the stacked planes
  int lda[MAX_GROUPS];     // row stride of a / a2 (plain mode)
};

struct RowSlot {  // RowInfo packed for shared memory
  long long base;
  int iy0, ix0;
};

// MODE selects the one operand-producer path that is compiled in (0 plain rows, 5 plain rows + addend, 1 implicit
// im2col, 3 stem taps, 4 relation pair gating); REL = true adds the relation-head epilogues (dot / finish).  One
// path per instantiation keeps each kernel's code — and its instruction-cache footprint — small.
template <int BLOCK_N, int MODE, bool REL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_sbf16_kernel(const __grid_constant__ CUtensorMap tmap_w, const ASrc src, const Epilogue ep, int M, int N,
                  int Npad, int K, int splits, int kb_per_split, float* __restrict__ partial, int groups, int plane_rows,
                  const GroupTab tab, int* __restrict__ err) {
  pdl_launch_dependents();  // the next kernel may take SMs as this grid's CTAs retire
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ctrl = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = (uint64_t*)ctrl;           // [STAGES]
  uint64_t* empty_bar = full_bar + C::STAGES;     // [STAGES]
  uint64_t* tmem_full = empty_bar + C::STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_holder = (uint32_t*)(tmem_empty + 2);
  RowSlot* rows_all = (RowSlot*)(ctrl + 256);     // [2 groups][128] x 16 B
  float* stage_out = (float*)(ctrl + 4608);       // [4 warps][32][STG_LD]
  float* gates_all = (float*)(ctrl + 4608 + 4 * 32 * STG_LD * 4);  // [2 groups][128 rows][8] relation gate values

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef EGTR_GEMM_PROF
  long long g_prof_local[4] = {0, 0, 0, 0};
  const long long prof_start = clock64();
#endif
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = Npad / BLOCK_N;
  // work item = (output tile, K split): few-tile GEMMs with a long K (3x3 convs on C5, decoder FFN) are
  // spread over the SMs along K; their raw partial sums go to `partial` and a reduce kernel applies the epilogue
  const int tiles_per_group = m_tiles * n_tiles;
  const int total_tiles = groups * tiles_per_group * splits;
  const int k_blocks_all = K / BLOCK_K;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_w);
    for (int i = 0; i < C::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 128 + 1);  // 128 producer threads + the TMA thread's expect_tx arrive
      ptx::mbar_init(&empty_bar[i], 1);       // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 5) ptx::tmem_alloc<C::TMEM_COLS>(tmem_holder);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();  // barriers, TMEM and tensor-map prefetch above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA: weight tiles
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
        const int tg = w / splits, sp = w - tg * splits;
        const int g = tg / tiles_per_group, t = tg - g * tiles_per_group;
        const int n0 = tab.n_base[g] + (t % n_tiles) * BLOCK_N;
        const int kb_lo = sp * kb_per_split, kb_hi = min(k_blocks_all, kb_lo + kb_per_split);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          PROF_T0(t_a);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 101);
          PROF_ADD(0, t_a);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::B_TILE_BYTES);
          ptx::tma_load_2d(st + 2 * A_TILE_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, n0);
          ptx::tma_load_2d(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, plane_rows + n0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, BLOCK_N);
    int stage = 0, phase = 0, it = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x, ++it) {
      const int sp = w % splits;
      const int kb_lo = sp * kb_per_split, kb_hi = min(k_blocks_all, kb_lo + kb_per_split);
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      PROF_T0(t_b);
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, err, 102);
      PROF_ADD(0, t_b);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        PROF_T0(t_c);
        ptx::mbar_wait(&full_bar[stage], phase, err, 103);
        PROF_ADD(1, t_c);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;
          const uint32_t b_lo = b_hi + C::B_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
            const uint32_t koff = ks * UMMA_K * 2;  // bytes along K inside the swizzle row
            const uint64_t dah = ptx::umma_desc_sw128(a_hi + koff), dal = ptx::umma_desc_sw128(a_lo + koff);
            const uint64_t dbh = ptx::umma_desc_sw128(b_hi + koff), dbl = ptx::umma_desc_sw128(b_lo + koff);
            ptx::umma_bf16(d_tmem, dal, dbh, idesc, (kb != kb_lo) || (ks != 0));  // small terms first
            ptx::umma_bf16(d_tmem, dah, dbl, idesc, 1);
            ptx::umma_bf16(d_tmem, dah, dbh, idesc, 1);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (kb == kb_hi - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 6) {
    // ------------------------------------------------------------------ A producers (2 groups x 4 warps)
    const int grp = (warp - 6) >> 2;  // group g fills the k-blocks whose running number is == g (mod 2)
    const int p = (warp - 6) & 3;
    const int ptid = p * 32 + lane;
    const int kc = lane & 15;   // which float4 of the 64-float run
    const int rsub = lane >> 4; // 0/1: two rows per warp-wide load
    RowSlot* rows = rows_all + grp * 128;
    const uint32_t rows_s = ptx::smem_u32(rows);
    auto ld_rowslot = [](uint32_t addr) {  // explicit 16-byte shared load of one RowSlot
      uint32_t a, b, c, d;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
      RowSlot r;
      r.base = (long long)(((unsigned long long)b << 32) | a);
      r.iy0 = (int)c;
      r.ix0 = (int)d;
      return r;
    };
    int seq = 0;  // k-blocks issued so far by this CTA (all work items)
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
      const int tg = w / splits, sp = w - tg * splits;
      const int g = tg / tiles_per_group, t = tg - g * tiles_per_group;
      const float* __restrict__ a_ptr = tab.a[g];
      const float* __restrict__ a2_ptr = tab.a2[g];
      const int kb_lo = sp * kb_per_split, kb_hi = min(k_blocks_all, kb_lo + kb_per_split);
      const long long m0 = (long long)(t / n_tiles) * BLOCK_M;
      ptx::named_bar_sync(1 + grp, 128);  // previous tile's readers are done with `rows`
      {
        RowSlot rs;
        if (MODE == 4) {
          // relation pair tile: this thread's row is the pair (subject i, object j) of image b
          int pb, pi, pj;
          const bool ok = pair_decode(m0 + ptid, src.H, pb, pi, pj) && (m0 + ptid) < M;
          rs.base = ok ? pb : -1;
          rs.iy0 = min(pi, src.H - 1);
          rs.ix0 = min(pj, src.H - 1);
          // gate_l(i,j) = sigmoid(a_l(i) + b_l(j) + bias): separable logit, columns 2C of U and V (egtr.py:400)
          const float* ug = src.a + ((long long)pb * src.H + rs.iy0) * src.W * src.lda + 2 * src.C;
          const float* vg = src.a2 + ((long long)pb * src.H + rs.ix0) * src.W * src.lda + 2 * src.C;
          float* gt = gates_all + (grp * 128 + ptid) * 8;
#pragma unroll
          for (int l = 0; l < 8; ++l) gt[l] = (l < src.W) ? sigmoidf_(__ldg(ug + l * src.lda) + __ldg(vg + l * src.lda)) : 0.f;
        } else {
          RowInfo ri = decode_row(src, m0 + ptid, M);
          if (MODE == 0 || MODE == 5) ri.base = (m0 + ptid) * (long long)tab.lda[g];
          rs.base = ri.valid ? ri.base : -1;
          rs.iy0 = ri.iy0;
          rs.ix0 = ri.ix0;
        }
        rows[ptid] = rs;
      }
      ptx::named_bar_sync(1 + grp, 128);
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++seq) {
        if ((seq & 1) != grp) continue;
        const int stage = seq % C::STAGES, phase = (seq / C::STAGES) & 1;
        const int k0 = kb * BLOCK_K;
        int ky = 0, kx = 0, c0 = k0;
        if (MODE == 1) {
          const int tap = k0 / src.C;
          c0 = k0 - tap * src.C;
          ky = tap / src.KW;
          kx = tap - ky * src.KW;
        }
        const uint32_t a_hi_s = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t a_lo_s = a_hi_s + A_TILE_BYTES;
        // All global loads of a batch are issued back to back (predicated, no branches) before anything
        // consumes them, so the 16 (or 2 x 8 with the x+pos addend) 16-byte loads of a thread overlap
        // in the memory system; the stage's empty barrier is only waited for once they are in flight.
        auto store_row = [&](int i, const float4& x) {  // split one float4 to bf16 hi/lo and store it (interleaves ALU and LSU work)
          const int r = p * 32 + 2 * i + rsub;
          __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
          split_bf16(x.x, h0, l0);
          split_bf16(x.y, h1, l1);
          split_bf16(x.z, h2, l2);
          split_bf16(x.w, h3, l3);
          const uint32_t o = ptx::sw128_offset(r, kc >> 1) + (kc & 1) * 8;
          uint2 ph, pl;
          ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
          pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
          // explicit shared-space stores (the realigned dynamic-smem pointer is generic to the compiler)
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi_s + o), "r"(ph.x), "r"(ph.y));
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_lo_s + o), "r"(pl.x), "r"(pl.y));
        };
        auto row_off = [&](int i) -> long long {
          const RowSlot rs = ld_rowslot(rows_s + (p * 32 + 2 * i + rsub) * 16);
          long long off = rs.base + k0;  // plain rows
          if (MODE == 1) {
            const int iy = rs.iy0 + ky, ix = rs.ix0 + kx;
            const bool in = (unsigned)iy < (unsigned)src.H && (unsigned)ix < (unsigned)src.W;
            off = in ? (rs.base + (long long)iy * src.W + ix) * src.C + c0 : -1;
          }
          return rs.base >= 0 ? off : -1;
        };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MODE == 4) {
          // h1 = relu(b1 + sum_l gate_l(s,o) * (U_l(s) + V_l(o))) for this thread's 16 pairs x 4 channels.  The rows
          // of a thread are a 4 x 4 block of (subject, object) (pair_local), processed as two passes of 2 subjects
          // x 4 objects: 6 L1/L2-resident float4 loads feed 8 pairs per layer; two layers are in flight at a time.
          const int cbase = (t % n_tiles) * src.C + k0 + kc * 4;  // channel inside [rel 0..C) | conn C..2C)
          const float4 b1v = __ldg((const float4*)(src.aux + cbase));
          int img, ti0, tj0;
          pair_decode(m0, src.H, img, ti0, tj0);  // (image, first subject, first object) of the tile: row 0 is (s 0, o 0)
          const int nq1 = src.H - 1;
          const long long qstride = (long long)src.W * src.lda;
          const float* ubase = src.a + (long long)img * src.H * qstride + cbase;
          const float* vbase = src.a2 + (long long)img * src.H * qstride + cbase;
          const uint32_t gts = ptx::smem_u32(gates_all + grp * 128 * 8);
          const int s_first = ti0 + 4 * (p & 1), o_first = tj0 + 8 * (p >> 1) + 4 * rsub;
          int vo[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) vo[c] = (int)(min(o_first + c, nq1) * qstride);
          bool waited = false;
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {  // subjects a = 2*pass, 2*pass + 1
            const int uo0 = (int)(min(s_first + 2 * pass, nq1) * qstride), uo1 = (int)(min(s_first + 2 * pass + 1, nq1) * qstride);
            float4 acc[8];  // [a_local (2)][c (4)]
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = b1v;
            auto fma_layer = [&](int l, const float4 (&u)[2], const float4 (&v)[4]) {
#pragma unroll
              for (int al = 0; al < 2; ++al)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int r = p * 32 + 2 * (4 * (2 * pass + al) + c) + rsub;
                  float gv;
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gv) : "r"(gts + (r * 8 + l) * 4));
                  float4& a4 = acc[al * 4 + c];
                  a4.x = fmaf(gv, u[al].x + v[c].x, a4.x); a4.y = fmaf(gv, u[al].y + v[c].y, a4.y);
                  a4.z = fmaf(gv, u[al].z + v[c].z, a4.z); a4.w = fmaf(gv, u[al].w + v[c].w, a4.w);
                }
            };
            auto load_layer = [&](int l, float4 (&u)[2], float4 (&v)[4]) {
              u[0] = __ldg((const float4*)(ubase + uo0 + l * src.lda));
              u[1] = __ldg((const float4*)(ubase + uo1 + l * src.lda));
#pragma unroll
              for (int c = 0; c < 4; ++c) v[c] = __ldg((const float4*)(vbase + vo[c] + l * src.lda));
            };
            float4 ua[2], va[4], ub[2], vb[4];
            load_layer(0, ua, va);
#pragma unroll 1
            for (int l = 0; l < src.W; l += 2) {
              if (l + 1 < src.W) load_layer(l + 1, ub, vb);
              fma_layer(l, ua, va);
              if (l + 2 < src.W) load_layer(l + 2, ua, va);
              if (l + 1 < src.W) fma_layer(l + 1, ub, vb);
            }
            if (!waited) { ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 104); waited = true; }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              acc[q].x = fmaxf(acc[q].x, 0.f); acc[q].y = fmaxf(acc[q].y, 0.f);
              acc[q].z = fmaxf(acc[q].z, 0.f); acc[q].w = fmaxf(acc[q].w, 0.f);
              store_row(4 * (2 * pass + (q >> 2)) + (q & 3), acc[q]);
            }
          }
        } else if constexpr (MODE == 3) {
          // stem on the zero-padded NHWC4 image: this thread's float4 of the run is filter tap k0/4 + kc
          const int tap = (k0 >> 2) + kc;
          const int tky = tap / src.KW, tkx = tap - tky * src.KW;
          const bool tap_ok = tap < src.KH * src.KW;
          float4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const RowSlot rs = ld_rowslot(rows_s + (p * 32 + 2 * i + rsub) * 16);
            const long long off = (rs.base + (long long)(rs.iy0 + tky) * src.W + (rs.ix0 + tkx)) * 4;
            v[i] = (tap_ok && rs.base >= 0) ? __ldg((const float4*)(a_ptr + off)) : zero4;
          }
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 104);
#pragma unroll
          for (int i = 0; i < 16; ++i) store_row(i, v[i]);
        } else if constexpr (MODE != 5) {
          long long off[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) off[i] = row_off(i);
          float4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = off[i] >= 0 ? __ldg((const float4*)(a_ptr + off[i]) + kc) : zero4;
          PROF_T0(t_d);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 104);
          PROF_ADD(0, t_d);
          PROF_T0(t_e);
#pragma unroll
          for (int i = 0; i < 16; ++i) store_row(i, v[i]);
          PROF_ADD(1, t_e);
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            long long off[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) off[i] = row_off(half * 8 + i);
            float4 v[8], w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = off[i] >= 0 ? __ldg((const float4*)(a_ptr + off[i]) + kc) : zero4;
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = (off[i] >= 0 && a2_ptr != nullptr) ? __ldg((const float4*)(a2_ptr + off[i]) + kc) : zero4;
            if (half == 0) ptx::mbar_wait(&empty_bar[stage], phase ^ 1, err, 104);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i].x += w[i].x; v[i].y += w[i].y; v[i].z += w[i].z; v[i].w += w[i].w;
              store_row(half * 8 + i, v[i]);
            }
          }
        }
        ptx::fence_proxy_async_smem();  // make the st.shared visible to the tensor core's async proxy
        ptx::mbar_arrive(&full_bar[stage]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3 = TMEM lane quadrants)
    int it = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x, ++it) {
      const int tg = w / splits, sp = w - tg * splits;
      const int g = tg / tiles_per_group, t = tg - g * tiles_per_group;
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      const int n0 = (t % n_tiles) * BLOCK_N;
      if (ep.res != nullptr && splits == 1) {
        // pull this tile's residual rows into L2 while the MMAs are still running: the epilogue's own loads
        // (8 x 16 B per lane per chunk) are too few bytes in flight to stream them from HBM at speed
        const long long mr = (long long)(t / n_tiles) * BLOCK_M + warp * 32 + lane;
        const long long pr = mr < M ? out_row(ep, mr) : -1;
        if (pr >= 0) {
          const float* rp = ep.res + pr * ep.ldr + n0;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 32; ++j)
            if (n0 + j * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + j * 32));
        }
      }
      PROF_T0(t_f);
      ptx::mbar_wait(&tmem_full[acc], acc_phase, err, 105);
      PROF_ADD(0, t_f);
      ptx::tc_fence_after();
      PROF_T0(t_g);
      // Accumulator rows live one per thread in TMEM (lane = row).  Each 32x32 chunk is transposed through a
      // padded smem tile so that global traffic is row-contiguous: 8 lanes x float4 cover 128 B of one output
      // row (4 rows per warp instruction) for the stores and for the residual loads alike.
      float* stg = stage_out + warp * (32 * STG_LD);
      const int rsub = lane >> 3, cc = (lane & 7) * 4;
      const long long m_base = (long long)(t / n_tiles) * BLOCK_M + warp * 32;
      long long orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long mr = m_base + i * 4 + rsub;
        orow[i] = mr < M ? (splits > 1 ? (long long)sp * M + mr : out_row(ep, mr)) : -1;
      }
      // value.masked_fill(~mask, 0) of deformable_detr.py:1050-1052 folded into the store: masked rows write zeros
      uint32_t keep_bits = 0xffu;
      if (ep.row_keep != nullptr && splits == 1) {
        keep_bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (orow[i] >= 0 && ep.row_keep[orow[i]]) keep_bits |= 1u << i;
      }
      float* __restrict__ obase = splits > 1 ? partial : tab.out[g];
      const int ldo = splits > 1 ? Npad : ep.ldo;
      const float* __restrict__ rbase = splits > 1 ? nullptr : ep.res;
      const float* __restrict__ bias = (splits > 1 || !ep.bias) ? nullptr : ep.bias + tab.n_base[g];
      const int relu = splits > 1 ? 0 : ep.relu;
      const int ncols = splits > 1 ? Npad : N;
      // relation-head epilogues: the connectivity MLP's 256->1 last layer as a row dot product (dot tile), and the
      // frequency-bias / logit-adjustment / sigmoid finish of pred_rel (fin)
      const bool dot_tile = REL && (ep.dot_w != nullptr) && (splits == 1) && (n0 >= ep.dot_col0);
      const bool fin = REL && ep.fin && (splits == 1);
      float dpart[REL ? 8 : 1];
#pragma unroll
      for (int i = 0; i < (REL ? 8 : 1); ++i) dpart[i] = 0.f;
      const float* trip[REL ? 8 : 1];
      if (fin) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          trip[i] = nullptr;
          if (orow[i] >= 0 && ep.triplet != nullptr) {
            const long long pr = orow[i];               // pair index (b*N + s)*N + o
            const long long bs = pr / ep.fin_n;         // b*N + s
            const int o = (int)(pr - bs * ep.fin_n);
            const long long bb = bs / ep.fin_n;
            trip[i] = ep.triplet + ((long long)ep.cls[bs] * ep.k1 + ep.cls[bb * ep.fin_n + o]) * N;
          }
        }
      }
      const uint32_t stg_w = ptx::smem_u32(stg) + lane * (STG_LD * 4);               // this lane's row (TMEM order)
      const uint32_t stg_r = ptx::smem_u32(stg) + (rsub * STG_LD + cc) * 4;           // this lane's float4 column (row order)
      const bool aligned = ((ldo & 3) == 0) && (!rbase || (ep.ldr & 3) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        const int n = n0 + c0 + cc;  // first of this lane's 4 columns
        if (n0 + c0 >= ncols) break;  // whole chunk beyond N (padded weight rows)
        const bool chunk_full = aligned && (n0 + c0 + 32 <= ncols);  // warp-uniform fast path
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) {
          if (chunk_full) b4 = __ldg((const float4*)(bias + n));
          else { if (n < ncols) b4.x = __ldg(bias + n); if (n + 1 < ncols) b4.y = __ldg(bias + n + 1);
                 if (n + 2 < ncols) b4.z = __ldg(bias + n + 2); if (n + 3 < ncols) b4.w = __ldg(bias + n + 3); }
        }
        uint32_t r[32];
        PROF_T0(t_h);
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * BLOCK_N + c0, r);
        ptx::tmem_ld_wait();
        PROF_ADD(2, t_h);
        PROF_T0(t_i);
        __syncwarp();  // previous chunk's readers are done with the staging tile
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_w + 16 * j), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                       "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
        __syncwarp();
        float4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)  // all eight row reads are issued together (no control flow in between)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[i].x), "=f"(o[i].y), "=f"(o[i].z), "=f"(o[i].w)
                       : "r"(stg_r + i * 4 * STG_LD * 4) : "memory");
        if (chunk_full) {
          float4 rs[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rs[i] = (rbase && orow[i] >= 0) ? *(const float4*)(rbase + orow[i] * ep.ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v = o[i];
            v.x += b4.x + rs[i].x; v.y += b4.y + rs[i].y; v.z += b4.z + rs[i].z; v.w += b4.w + rs[i].w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (!((keep_bits >> i) & 1u)) v = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (REL) {
            if (dot_tile) {
              const float4 w4 = __ldg((const float4*)(ep.dot_w + (n - ep.dot_col0)));
              dpart[i] = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, dpart[i]))));
              continue;
            }
            if (fin && orow[i] >= 0) {
              if (trip[i]) { v.x += __ldg(trip[i] + n); v.y += __ldg(trip[i] + n + 1); v.z += __ldg(trip[i] + n + 2); v.w += __ldg(trip[i] + n + 3); }
              if (ep.adj) { v.x -= __ldg(ep.adj + n); v.y -= __ldg(ep.adj + n + 1); v.z -= __ldg(ep.adj + n + 2); v.w -= __ldg(ep.adj + n + 3); }
              v.x = sigmoidf_(v.x); v.y = sigmoidf_(v.y); v.z = sigmoidf_(v.z); v.w = sigmoidf_(v.w);
            }
            }
            if (orow[i] >= 0) {
              if (ep.out_fmt == EGTR_FMT_P32 && splits == 1) {  // P32 rows: hi 8 bytes at (n % 32) * 2 of the group, lo 64 bytes further
                __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
                uint2 ph, pl;
                ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
                pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
                uint8_t* g = (uint8_t*)(obase + orow[i] * ldo + (n & ~31)) + (n & 31) * 2;
                *(uint2*)g = ph;
                *(uint2*)(g + 64) = pl;
              } else {
                *(float4*)(obase + orow[i] * ldo + n) = v;
              }
            }
          }
        } else {  // N tail or unaligned leading dimension: per-element, still fully unrolled (registers only)
          auto put = [&](long long row, int col, float v, float bv, bool keep, const float* tp) {
            if (col < ncols) {
              v += bv;
              if (rbase) v += rbase[row * ep.ldr + col];
              if (relu) v = fmaxf(v, 0.f);
              if (fin) {
                if (tp) v += __ldg(tp + col);
                if (ep.adj) v -= __ldg(ep.adj + col);
                v = sigmoidf_(v);
              }
              obase[row * ldo + col] = keep ? v : 0.f;
            }
          };
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (orow[i] >= 0) {
              const bool keep = (keep_bits >> i) & 1u;
              const float* tp = fin ? trip[REL ? i : 0] : nullptr;
              put(orow[i], n, o[i].x, b4.x, keep, tp);
              put(orow[i], n + 1, o[i].y, b4.y, keep, tp);
              put(orow[i], n + 2, o[i].z, b4.z, keep, tp);
              put(orow[i], n + 3, o[i].w, b4.w, keep, tp);
            }
          }
        }
        PROF_ADD(3, t_i);
      }
      if (dot_tile) {
#pragma unroll
        for (int i = 0; i < (REL ? 8 : 1); ++i) {  // the 8 lanes that share a row hold its partial sums
          float d = dpart[i];
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (cc == 0 && orow[i] >= 0) ep.dot_out[orow[i]] = sigmoidf_(d + ep.dot_b);
        }
      }
      PROF_ADD(1, t_g);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }
#ifdef EGTR_GEMM_PROF
  if (lane == 0 && blockIdx.x < 148) {
    // slots: role*4 + {0: first wait, 1: second counter, 2: role total}; roles: 0 TMA(warp4) 1 MMA(warp5) 2 producer(warp6) 3 epilogue(warp0)
    const int role = warp == 4 ? 0 : warp == 5 ? 1 : warp == 6 ? 2 : warp == 0 ? 3 : -1;
    if (role >= 0) {
      g_prof[blockIdx.x][role * 4 + 0] = g_prof_local[0];
      g_prof[blockIdx.x][role * 4 + 1] = g_prof_local[1];
      g_prof[blockIdx.x][role * 4 + 2] = clock64() - prof_start;
      if (role == 3) { g_prof[blockIdx.x][14] = g_prof_local[2]; g_prof[blockIdx.x][15] = g_prof_local[3]; }
    }
  }
#endif

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int npad, k, block_n;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && npad == o.npad && k == o.k && block_n == o.block_n; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.npad * 1315423911u ^ (size_t)k.k * 2654435761u ^ (size_t)k.block_n;
  }
};

// Weight tensor maps are immutable per (pointer, shape): encode once, reuse for every launch.
int weight_tensor_map(const void* planes, int Npad, int K, int block_n, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  MapKey key{planes, Npad, K, block_n};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EGTR_OK;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  EGTR_CHECK(enc != nullptr, EGTR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)2 * Npad};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(planes), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGTR_CHECK(r == CUDA_SUCCESS, EGTR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  if (cache.size() >= (1u << 15)) cache.clear();  // bounded: maps are passed to kernels by value, re-encoding is cheap
  cache.emplace(key, m);
  *out = m;
  return EGTR_OK;
}

int* device_error_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

// out = epilogue(sum over splits of partial[s][m][n])
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, int Npad, const Epilogue ep) {
  pdl_entry();
  const int n4 = (N + 3) / 4;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * n4) return;
  const long long m = i / n4;
  const int n = (int)(i - m * n4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *(const float4*)(partial + ((long long)s * M + m) * Npad + n);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const long long orow = out_row(ep, m);
  if (orow < 0) return;
  const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n + j < N) {
      float v = o[j];
      if (ep.bias) v += __ldg(ep.bias + n + j);
      if (ep.res) v += ep.res[orow * ep.ldr + n + j];
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.row_keep && !ep.row_keep[orow]) v = 0.f;
      ep.out[orow * ep.ldo + n + j] = v;
    }
  }
}

// grow-only scratch for split-K partial sums (allocated outside graph capture, during warm-up)
float* partial_buffer(size_t floats) {  // grow-only per scratch slot; older buffers stay alive for captured graphs
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

template <int BLOCK_N, int MODE, bool REL>
int launch(const ASrc& a, const void* planes, int M, int N, int Npad, int K, const Epilogue& ep, cudaStream_t st,
           int groups = 1, int plane_rows = 0, const GroupTab* gtab = nullptr) {
  using C = Cfg<BLOCK_N>;
  if (plane_rows == 0) plane_rows = Npad;
  GroupTab tab = {};
  if (gtab) tab = *gtab;
  else { tab.a[0] = a.a; tab.a2[0] = a.a2; tab.out[0] = ep.out; tab.n_base[0] = 0; tab.lda[0] = a.lda; }
  CUtensorMap tmap;
  int rc = weight_tensor_map(planes, plane_rows, K, BLOCK_N, &tmap);
  if (rc != EGTR_OK) return rc;
  static bool attr_set = false;  // one flag per instantiation
  if (!attr_set) {
    EGTR_CUDA(cudaFuncSetAttribute(gemm_sbf16_kernel<BLOCK_N, MODE, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = groups * cdiv(M, BLOCK_M) * (Npad / BLOCK_N);
  const int k_blocks = K / BLOCK_K;
  int splits = 1;
  if (groups == 1 && splitk_max() > 1 && tiles * 2 <= num_sms() && k_blocks >= 8) {
    splits = num_sms() / tiles;
    if (splits > k_blocks / 4) splits = k_blocks / 4;
    if (splits > splitk_max()) splits = splitk_max();
    if (splits < 1) splits = 1;
  }
  int kbps = cdiv(k_blocks, splits);
  splits = cdiv(k_blocks, kbps);
  float* partial = nullptr;
  if (splits > 1) {
    partial = partial_buffer((size_t)splits * M * Npad);
    EGTR_CHECK(partial != nullptr, EGTR_ERR_CUDA, "split-K scratch allocation failed");
  }
  const int work = tiles * splits;
  const int grid = balanced_grid(work, num_sms() / grid_div());  // throughput mode: a share of the GPU per GEMM
  launch_pdl(gemm_sbf16_kernel<BLOCK_N, MODE, REL>, dim3(grid), dim3(NUM_THREADS), (size_t)(C::SMEM_BYTES), st, tmap, a, ep, M, N, Npad, K, splits, kbps, partial, groups, plane_rows, tab, device_error_flag());
  EGTR_CUDA(cudaGetLastError());
  if (splits > 1) {
    const long long n = (long long)M * ((N + 3) / 4);
    launch_pdl(splitk_reduce_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), st, partial, splits, M, N, Npad, ep);
    count_launch();
    EGTR_CUDA(cudaGetLastError());
  }
  return EGTR_OK;
}

// fp32 [N,K] -> bf16 hi/lo planes [2][Npad][K]
__global__ void split_weight_kernel(const float* __restrict__ w, int N, int K, int Npad, __nv_bfloat16* __restrict__ planes) {
  pdl_entry();
  const long long total = (long long)Npad * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K);
    __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
    if (n < N) split_bf16(w[i], h, l);
    planes[i] = h;
    planes[total + i] = l;
  }
}


// Picks the instantiation: tile width from the padded N, producer path from the operand source, relation epilogues
// only where the epilogue descriptor asks for them.
int dispatch(const ASrc& a, const void* planes, int M, int N, int Npad, int K, const Epilogue& ep, cudaStream_t st, int groups,
             int plane_rows, const GroupTab* gtab) {
  EGTR_ONE_DEVICE();
  static const int forced_bn = [] { const char* e = getenv("EGTR_GEMM_BLOCK_N"); return e ? atoi(e) : 0; }();  // dev experiments only
  int bn = (Npad % 256 == 0) ? 256 : (Npad % 128 == 0 ? 128 : 64);
  if (forced_bn == 64 || (forced_bn == 128 && Npad % 128 == 0)) bn = forced_bn;
  const bool rel = ep.fin || ep.dot_w || ep.pair_n > 0;
  const int mode = (a.mode == 0 && a.a2) ? 5 : a.mode;
#define EGTR_GO(BN, MD, RL) return launch<BN, MD, RL>(a, planes, M, N, Npad, K, ep, st, groups, plane_rows, gtab)
#define EGTR_BN(MD, RL) do { if (bn == 256) EGTR_GO(256, MD, RL); if (bn == 128) EGTR_GO(128, MD, RL); EGTR_GO(64, MD, RL); } while (0)
  if (mode == 4) { if (bn == 256) EGTR_GO(256, 4, true); set_last_error("pair stage needs Npad %% 256 == 0"); return EGTR_ERR_ARG; }
  if (rel) { if (mode == 0) EGTR_BN(0, true); set_last_error("relation epilogues are built for plain rows"); return EGTR_ERR_UNSUPPORTED; }
  if (mode == 0) EGTR_BN(0, false);
  if (mode == 5) EGTR_BN(5, false);
  if (mode == 1) EGTR_BN(1, false);
  if (mode == 3) EGTR_BN(3, false);
#undef EGTR_BN
#undef EGTR_GO
  set_last_error("egtr_gemm_sbf16: operand mode %d has no tensor-core producer (use egtr_gemm_f32)", a.mode);
  return EGTR_ERR_UNSUPPORTED;
}

}  // namespace

}  // namespace egtr

using namespace egtr;

extern "C" int egtr_split_weight_bf16(const float* w, int N, int K, int Npad, void* planes, egtr_stream_t s) {
  EGTR_CHECK(w && planes && N > 0 && K > 0 && Npad >= N && Npad % 64 == 0, EGTR_ERR_ARG,
             "egtr_split_weight_bf16: bad arguments (N=%d K=%d Npad=%d)", N, K, Npad);
  const long long total = (long long)Npad * K;
  int grid = cdiv(total, 256);
  if (grid > 4 * 148) grid = 4 * 148;
  launch_pdl(split_weight_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, w, N, K, Npad, (__nv_bfloat16*)planes);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_gemm_sbf16(const egtr_asrc_t* a, const void* w_planes, int M, int N, int Npad, int K,
                               const egtr_epilogue_t* ep, egtr_stream_t s) {
  EGTR_CHECK(a && w_planes && ep && a->a && ep->out, EGTR_ERR_ARG, "egtr_gemm_sbf16: null argument");
  EGTR_CHECK(M > 0 && N > 0 && K > 0 && K % 64 == 0 && Npad % 64 == 0 && Npad >= N, EGTR_ERR_ARG,
             "egtr_gemm_sbf16: need K %% 64 == 0 and Npad %% 64 == 0 (M=%d N=%d Npad=%d K=%d)", M, N, Npad, K);
  EGTR_CHECK(a->mode == 0 || (a->mode == 1 && a->C % 64 == 0 && K == a->KH * a->KW * a->C) ||
                 (a->mode == 2 && K >= a->KH * a->KW * a->C) || (a->mode == 3 && a->C == 4 && a->pad == 0 && K >= a->KH * a->KW * 4) ||
                 (a->mode == 4 && a->a2 && a->aux && K == a->C && a->C % 64 == 0 && a->W <= 8 && ep->pair_n == a->H && Npad % 256 == 0),
             EGTR_ERR_ARG, "egtr_gemm_sbf16: conv source needs C %% 64 == 0 and K == KH*KW*C (mode=%d C=%d K=%d)", a->mode, a->C, K);
  EGTR_CHECK(a->mode != 0 || (a->lda % 4 == 0 && a->lda >= K), EGTR_ERR_ARG, "egtr_gemm_sbf16: lda=%d", a->lda);
  EGTR_CHECK(((uintptr_t)a->a & 15) == 0 && ((uintptr_t)w_planes & 127) == 0, EGTR_ERR_ARG, "egtr_gemm_sbf16: alignment");
  count_launch();
  cudaStream_t st = (cudaStream_t)s;
  if (a->fmt == EGTR_FMT_P32) return gemm_p32_dispatch(*a, w_planes, Npad, M, N, Npad, K, *ep, st);
  EGTR_CHECK(ep->res_fmt == EGTR_FMT_F32 && (ep->out_fmt == EGTR_FMT_F32 || (N % 32 == 0 && ep->ldo % 4 == 0 && !ep->fin)), EGTR_ERR_UNSUPPORTED,
             "egtr_gemm_sbf16: fp32-operand kernel: no P32 residuals; P32 output needs N %% 32 == 0 and no split-K");
  return dispatch(*a, w_planes, M, N, Npad, K, *ep, st, 1, 0, nullptr);
}

extern "C" int egtr_gemm_sbf16_grouped(const float* const* a_ptrs, const float* const* a2_ptrs, float* const* out_ptrs,
                                       const int* n_base, int groups, const int* lda, const void* w_planes, int plane_rows, int M,
                                       int N, int Npad, int K, const egtr_epilogue_t* ep, egtr_stream_t s) {
  EGTR_CHECK(a_ptrs && out_ptrs && n_base && w_planes && ep, EGTR_ERR_ARG, "egtr_gemm_sbf16_grouped: null argument");
  EGTR_CHECK(groups >= 1 && groups <= MAX_GROUPS, EGTR_ERR_ARG, "egtr_gemm_sbf16_grouped: 1..%d groups (got %d)", MAX_GROUPS, groups);
  EGTR_CHECK(M > 0 && N > 0 && K > 0 && K % 64 == 0 && Npad % 64 == 0 && Npad >= N && plane_rows % 64 == 0, EGTR_ERR_ARG,
             "egtr_gemm_sbf16_grouped: bad shape (M=%d N=%d Npad=%d K=%d)", M, N, Npad, K);
  EGTR_CHECK(lda && ep->res == nullptr && ep->rows_per_b == 0, EGTR_ERR_ARG,
             "egtr_gemm_sbf16_grouped: plain rows, no residual / row remap");
  GroupTab tab = {};
  for (int g = 0; g < groups; ++g) {
    EGTR_CHECK(a_ptrs[g] && out_ptrs[g] && n_base[g] % 64 == 0 && n_base[g] + Npad <= plane_rows && lda[g] % 4 == 0 &&
                   lda[g] >= K, EGTR_ERR_ARG, "egtr_gemm_sbf16_grouped: group %d", g);
    tab.a[g] = a_ptrs[g];
    tab.a2[g] = a2_ptrs ? a2_ptrs[g] : nullptr;
    tab.out[g] = out_ptrs[g];
    tab.n_base[g] = n_base[g];
    tab.lda[g] = lda[g];
  }
  ASrc a = {};
  a.a = a_ptrs[0];
  a.mode = 0;
  a.lda = lda[0];
  count_launch();
  cudaStream_t st = (cudaStream_t)s;
  for (int g = 0; g < groups; ++g)
    if (tab.a2[g]) a.a2 = tab.a2[g];  // any addend -> the addend-capable producer (groups without one skip it at run time)
  return dispatch(a, w_planes, M, N, Npad, K, *ep, st, groups, plane_rows, &tab);
}

#ifdef EGTR_GEMM_PROF
extern "C" int egtr_debug_gemm_prof(unsigned long long* host_out) {
  EGTR_CUDA(cudaMemcpyFromSymbol(host_out, egtr::g_prof, sizeof(unsigned long long) * 148 * 16));
  return EGTR_OK;
}
#endif
