// fp32 CUDA-core GEMM with the same operand-source / epilogue contract as the tcgen05 kernel.
// Used where tensor-core tiles do not fit (K % 64 != 0, a few hundred rows with N < 64) and as the
// independent cross-check of the tensor-core path in tests.  64x64 tile, BK = 16, 4x4 per thread.
#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const ASrc src, const float* __restrict__ w, int M, int N, int K, const Epilogue ep) {
  pdl_entry();
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const long long m0 = (long long)blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int lrow = tid >> 2, lq = tid & 3;  // loader mapping: one float4 per thread per tile
  const RowInfo ri = decode_row(src, m0 + lrow, M);
  const int wn = n0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    const long long off = src.mode >= 2 ? -1 : row_offset(src, ri, k0);
    if (src.mode == 2) {
      if (ri.valid) av = gather4_nchw(src, ri.base, ri.iy0, ri.ix0, k0 + lq * 4);
    } else if (src.mode == 3) {
      const long long o3 = tap_offset_nhwc4(src, ri, (k0 >> 2) + lq);
      if (o3 >= 0) av = __ldg((const float4*)(src.a + o3));
    } else if (off >= 0) {
      av = __ldg((const float4*)(src.a + off) + lq);
      if (src.a2) {
        const float4 p = __ldg((const float4*)(src.a2 + off) + lq);
        av.x += p.x; av.y += p.y; av.z += p.z; av.w += p.w;
      }
    }
    if (wn < N) wv = __ldg((const float4*)(w + (long long)wn * K + k0) + lq);
    __syncthreads();
    As[lq * 4 + 0][lrow] = av.x; As[lq * 4 + 1][lrow] = av.y; As[lq * 4 + 2][lrow] = av.z; As[lq * 4 + 3][lrow] = av.w;
    Ws[lq * 4 + 0][lrow] = wv.x; Ws[lq * 4 + 1][lrow] = wv.y; Ws[lq * 4 + 2][lrow] = wv.z; Ws[lq * 4 + 3][lrow] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *(const float4*)&As[k][ty * 4];
      const float4 b = *(const float4*)&Ws[k][tx * 4];
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const long long orow = out_row(ep, m);
    if (orow < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float o = acc[i][j];
      if (ep.bias) o += __ldg(ep.bias + n);
      if (ep.res) o += ep.res[orow * ep.ldr + n];
      if (ep.relu) o = fmaxf(o, 0.f);
      if (ep.row_keep && !ep.row_keep[orow]) o = 0.f;
      ep.out[orow * ep.ldo + n] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny GEMM: a few hundred rows (the decoder's N_q queries, the relation head's per-query tensors).
// These launches are latency-bound, not throughput-bound, so the goal is many small CTAs and no setup
// cost: 16x64 output tile per 128-thread CTA, K in chunks of 32 with the next chunk's global loads in
// flight while the current one is multiplied.  fp32 FMA throughout (exact w.r.t. the reference).
// blockIdx.z = group (stacked weights / per-group operands), as in the tensor-core grouped launch.
struct SkinnyGroups {
  const float* a[16];
  const float* a2[16];
  float* out[16];
  int n_base[16];
  int lda[16];
};

constexpr int SK_TM = 16, SK_TN = 64, SK_KC = 256, SK_LD = SK_KC + 4;  // +4 floats: conflict-free float4 rows
constexpr int SK_STAGE_FLOATS = (2 * SK_TM + SK_TN) * SK_LD;              // A, A2 and W tiles of one K chunk
constexpr int SK_STAGE_BYTES = SK_STAGE_FLOATS * 4;                       // 99.8 KB: K <= 256 needs one stage -> two CTAs per SM
constexpr int SK_SMEM_BYTES = 2 * SK_STAGE_BYTES;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA (256 threads) = 16 x 64 outputs, 4 per thread.  A whole 256-wide K chunk of both operands is fetched with
// cp.async in ONE burst (a single memory round trip for every K <= 256 shape of the decoder); longer K double-buffers
// chunks.  The launch asks for one stage of smem when one chunk suffices, so two CTAs stay resident per SM.
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const SkinnyGroups gt, const float* __restrict__ w, int M, int N, int K, const Epilogue ep, int splits, int k_per,
                   float* __restrict__ partial) {
  pdl_entry();
  extern __shared__ __align__(16) float sk_smem[];
  // blockIdx.z = group (grouped launch) or K split (split-K launch: raw partial sums, the consumer applies the epilogue)
  const int g = splits > 1 ? 0 : blockIdx.z;
  const int sp = splits > 1 ? blockIdx.z : 0;
  const int k_lo = sp * k_per, k_hi = min(K, k_lo + k_per);
  const float* __restrict__ A = gt.a[g];
  const float* __restrict__ A2 = gt.a2[g];
  const int lda = gt.lda[g];
  const float* __restrict__ Wg = w + (long long)gt.n_base[g] * K;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SK_TM, n0 = blockIdx.x * SK_TN;
  const int r = tid >> 4, c16 = tid & 15;  // compute mapping: row r, columns c16 + 16*j (j < 4)

  auto issue = [&](int k0, int buf) {
    float* As = sk_smem + buf * SK_STAGE_FLOATS;
    float* A2s = As + SK_TM * SK_LD;
    float* Ws = A2s + SK_TM * SK_LD;
    const int kc = min(SK_KC, k_hi - k0) >> 2;  // float4 per row in this chunk
    for (int i = tid; i < SK_TM * (SK_KC / 4); i += 256) {
      const int row = i / (SK_KC / 4), q = i - row * (SK_KC / 4);
      const bool ok = (m0 + row) < M && q < kc;
      float* d = As + row * SK_LD + q * 4;
      if (ok) cp_async16(d, A + (long long)(m0 + row) * lda + k0 + q * 4);
      else *(float4*)d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A2) {
        float* d2 = A2s + row * SK_LD + q * 4;
        if (ok) cp_async16(d2, A2 + (long long)(m0 + row) * lda + k0 + q * 4);
        else *(float4*)d2 = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int i = tid; i < SK_TN * (SK_KC / 4); i += 256) {
      const int row = i / (SK_KC / 4), q = i - row * (SK_KC / 4);
      float* d = Ws + row * SK_LD + q * 4;
      if ((n0 + row) < N && q < kc) cp_async16(d, Wg + (long long)(n0 + row) * K + k0 + q * 4);
      else *(float4*)d = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = 0.f;
  const int chunks = (k_hi - k_lo + SK_KC - 1) / SK_KC;
  issue(k_lo, 0);
  for (int c = 0; c < chunks; ++c) {
    if (c + 1 < chunks) { issue(k_lo + (c + 1) * SK_KC, (c + 1) & 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const float* As = sk_smem + (c & 1) * SK_STAGE_FLOATS + r * SK_LD;
    const float* A2s = As + SK_TM * SK_LD;
    const float* Ws = sk_smem + (c & 1) * SK_STAGE_FLOATS + 2 * SK_TM * SK_LD + c16 * SK_LD;
    const int klen = min(SK_KC, k_hi - (k_lo + c * SK_KC));
#pragma unroll 4
    for (int k = 0; k < klen; k += 4) {
      float4 a = *(const float4*)(As + k);
      if (A2) { const float4 p = *(const float4*)(A2s + k); a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wv = *(const float4*)(Ws + j * 16 * SK_LD + k);
        acc[j] = fmaf(a.x, wv.x, acc[j]); acc[j] = fmaf(a.y, wv.y, acc[j]);
        acc[j] = fmaf(a.z, wv.z, acc[j]); acc[j] = fmaf(a.w, wv.w, acc[j]);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the chunk after next overwrites it
  }
  const int m = m0 + r;
  if (m >= M) return;
  if (splits > 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + c16 + 16 * j;
      if (n < N) partial[((long long)sp * M + m) * N + n] = acc[j];
    }
    return;
  }
  const long long orow = out_row(ep, m);
  if (orow < 0) return;
  float* __restrict__ out = gt.out[g];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + c16 + 16 * j;
    if (n < N) {
      float o = acc[j];
      if (ep.bias) o += __ldg(ep.bias + gt.n_base[g] + n);
      if (ep.res) o += ep.res[orow * ep.ldr + n];
      if (ep.relu) o = fmaxf(o, 0.f);
      if (ep.row_keep && !ep.row_keep[orow]) o = 0.f;
      out[orow * ep.ldo + n] = o;
    }
  }
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_gemm_f32_grouped(const float* const* a_ptrs, const float* const* a2_ptrs, float* const* out_ptrs,
                                     const int* n_base, int groups, const int* lda, const float* w, int M, int N, int K,
                                     const egtr_epilogue_t* ep, egtr_stream_t s) {
  EGTR_CHECK(a_ptrs && out_ptrs && n_base && lda && w && ep, EGTR_ERR_ARG, "egtr_gemm_f32_grouped: null argument");
  EGTR_CHECK(groups >= 1 && groups <= 16 && M > 0 && M <= 16 * 65535 && N > 0 && K > 0 && K % 4 == 0, EGTR_ERR_ARG,
             "egtr_gemm_f32_grouped: bad shape (groups=%d M=%d N=%d K=%d; K %% 4 == 0)", groups, M, N, K);
  EGTR_CHECK(groups == 1 || (ep->res == nullptr && ep->rows_per_b == 0), EGTR_ERR_ARG, "egtr_gemm_f32_grouped: no residual / remap with groups");
  SkinnyGroups gt = {};
  for (int g = 0; g < groups; ++g) {
    EGTR_CHECK(a_ptrs[g] && out_ptrs[g] && lda[g] % 4 == 0 && lda[g] >= K, EGTR_ERR_ARG, "egtr_gemm_f32_grouped: group %d", g);
    gt.a[g] = a_ptrs[g];
    gt.a2[g] = a2_ptrs ? a2_ptrs[g] : nullptr;
    gt.out[g] = out_ptrs[g];
    gt.n_base[g] = n_base[g];
    gt.lda[g] = lda[g];
  }
  static bool attr = false;
  if (!attr) {
    EGTR_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
    attr = true;
  }
  dim3 grid(cdiv(N, SK_TN), cdiv(M, SK_TM), groups);
  launch_pdl(gemm_skinny_kernel, dim3(grid), dim3(256), (size_t)(K > SK_KC ? SK_SMEM_BYTES : SK_STAGE_BYTES), (cudaStream_t)s, gt, w, M, N, K,
             *ep, 1, K, (float*)nullptr);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_gemm_f32_splitk(const float* a, const float* a2, int lda, const float* w, int M, int N, int K, int splits,
                                    float* partial, egtr_stream_t s) {
  EGTR_CHECK(a && w && partial && M > 0 && M <= 16 * 65535 && N > 0 && K > 0 && K % 4 == 0 && lda % 4 == 0 && lda >= K, EGTR_ERR_ARG,
             "egtr_gemm_f32_splitk: bad shape (M=%d N=%d K=%d lda=%d)", M, N, K, lda);
  EGTR_CHECK(splits >= 2 && splits <= 64 && K % splits == 0 && (K / splits) % 4 == 0, EGTR_ERR_ARG,
             "egtr_gemm_f32_splitk: splits=%d must divide K=%d into multiples of 4", splits, K);
  SkinnyGroups gt = {};
  gt.a[0] = a; gt.a2[0] = a2; gt.lda[0] = lda;
  static bool attr = false;
  if (!attr) {
    EGTR_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
    attr = true;
  }
  Epilogue ep = {};
  dim3 grid(cdiv(N, SK_TN), cdiv(M, SK_TM), splits);
  launch_pdl(gemm_skinny_kernel, dim3(grid), dim3(256), (size_t)(K / splits > SK_KC ? SK_SMEM_BYTES : SK_STAGE_BYTES), (cudaStream_t)s, gt, w,
             M, N, K, ep, splits, K / splits, partial);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_gemm_f32(const egtr_asrc_t* a, const float* w, int M, int N, int K, const egtr_epilogue_t* ep,
                             egtr_stream_t s) {
  EGTR_CHECK(a && w && ep && a->a && ep->out, EGTR_ERR_ARG, "egtr_gemm_f32: null argument");
  EGTR_CHECK(M > 0 && N > 0 && K > 0 && K % 16 == 0, EGTR_ERR_ARG, "egtr_gemm_f32: need K %% 16 == 0 (M=%d N=%d K=%d)", M, N, K);
  EGTR_CHECK(a->mode == 0 || (a->mode == 1 && a->C % 16 == 0 && K == a->KH * a->KW * a->C) ||
                 (a->mode == 2 && K >= a->KH * a->KW * a->C) || (a->mode == 3 && a->C == 4 && a->pad == 0 && K >= a->KH * a->KW * 4), EGTR_ERR_ARG,
             "egtr_gemm_f32: conv source needs C %% 16 == 0 and K == KH*KW*C");
  EGTR_CHECK(a->mode != 0 || a->lda % 4 == 0, EGTR_ERR_ARG, "egtr_gemm_f32: lda %% 4 != 0");
  dim3 grid(cdiv(N, TN), cdiv(M, TM));
  EGTR_CHECK(grid.y <= 65535, EGTR_ERR_ARG, "egtr_gemm_f32: M too large for this path (%d)", M);
  launch_pdl(gemm_f32_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)s, *a, w, M, N, K, *ep);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
