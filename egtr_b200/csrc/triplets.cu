// Triplet extraction on the device — the post-processing that follows the model in the reference's evaluation loop
// (`/root/reference/train_egtr.py:43-173`, `lib/pytorch_misc.py:27-34`), SURVEY.md §8f row 1.
//
//   obj_scores, pred_classes = max(softmax(logits)[:, :num_labels])                    (56-58)
//   sub_ob = outer(obj_scores, obj_scores), zero diagonal                              (59-62)
//   rel    = clamp(pred_rel, 0, 1) * clamp(pred_connectivity, 0, 1)                    (65-68)
//   "multiple" : score[s,o,p] = rel[s,o,p] * sub_ob[s,o]      -> top-k (s,o,p), rel_scores[k]        (84-94)
//   "single"   : score[s,o]   = max_p rel[s,o,p] * sub_ob[s,o] -> top-k (s,o),   rel_scores[k,P]      (120-128)
//
// The reference sorts ALL N^2 (or N^2 P) scores on the CPU (numpy argsort) to keep 100.  Here: scores are written
// once, a two-pass radix select on the float bit patterns (scores are >= 0, so the IEEE bits order like integers)
// finds the k-th largest value exactly, the <= k survivors are compacted and ranked (score descending, flat index
// ascending on ties — numpy's quicksort leaves tie order unspecified).  Everything is HBM-bound: ~3 passes over 4 N^2 P bytes.
#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

// ---------------------------------------------------------------- object scores / classes
__global__ void __launch_bounds__(256)
obj_scores_kernel(const float* __restrict__ logits, int K, int num_labels, int rows, float* __restrict__ score, int* __restrict__ cls) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* x = logits + (long long)row * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, x[k]);
  mx = warp_max(mx);
  float sum = 0.f, best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float e = expf(x[k] - mx);
    sum += e;
    if (k < num_labels && e > best) { best = e; bi = k; }
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { score[row] = best / sum; cls[row] = bi; }
}

// ---------------------------------------------------------------- scores
// multiple: one thread per (pair, p); single: one warp-lane group per pair
__global__ void __launch_bounds__(256)
triplet_scores_kernel(const float* __restrict__ rel, const float* __restrict__ conn, const float* __restrict__ obj, int N, int P,
                      int single, long long total, float* __restrict__ out) {
  pdl_entry();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long pair = single ? t : t / P;
  const long long bi = pair / N;  // b*N + s
  const int o = (int)(pair - bi * N);
  const long long b = bi / N;
  const int s = (int)(bi - b * N);
  const float so = (s == o) ? 0.f : obj[bi] * obj[b * N + o];
  const float c = conn ? fminf(fmaxf(conn[pair], 0.f), 1.f) : 1.f;
  float v;
  if (single) {
    v = 0.f;
    const float* r = rel + pair * P;
    for (int p = 0; p < P; ++p) v = fmaxf(v, fminf(fmaxf(r[p], 0.f), 1.f) * c);
  } else {
    v = fminf(fmaxf(rel[t], 0.f), 1.f) * c;
  }
  out[t] = v * so;
}

// ---------------------------------------------------------------- radix select (two 16-bit passes)
__global__ void __launch_bounds__(256)
hist_hi_kernel(const float* __restrict__ sc, long long n, unsigned* __restrict__ hist) {
  pdl_entry();  // hist [B][65536]
  const int b = blockIdx.y;
  const float* x = sc + (long long)b * n;
  unsigned* h = hist + (long long)b * 65536;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&h[__float_as_uint(x[i]) >> 16], 1u);
}
// one CTA per image: walk the histogram from the top, find the bin holding the k-th largest
__global__ void __launch_bounds__(1024)
pick_bin_kernel(const unsigned* __restrict__ hist, int k, unsigned* __restrict__ sel) {
  pdl_entry();  // sel [B][4]: bin, need_in_bin, (lo passes reuse)
  const int b = blockIdx.x;
  const unsigned* h = hist + (long long)b * 65536;
  __shared__ unsigned part[1024];
  // thread t owns bins [64 t, 64 t + 64) counted from the TOP (bin index 65535 - ...)
  unsigned s = 0;
  for (int j = 0; j < 64; ++j) s += h[65535 - (threadIdx.x * 64 + j)];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    int t = 0;
    while (t < 1024 && acc + part[t] < (unsigned)k) acc += part[t++];
    unsigned bin = 0, need = 0;
    if (t < 1024) {
      for (int j = 0; j < 64; ++j) {
        const unsigned c = h[65535 - (t * 64 + j)];
        if (acc + c >= (unsigned)k) { bin = 65535 - (t * 64 + j); need = k - acc; break; }
        acc += c;
      }
    }
    sel[b * 4 + 0] = bin;   // all elements in higher bins are selected
    sel[b * 4 + 1] = need;  // plus the `need` largest of this bin
  }
}
__global__ void __launch_bounds__(256)
hist_lo_kernel(const float* __restrict__ sc, long long n, const unsigned* __restrict__ sel, unsigned* __restrict__ hist) {
  pdl_entry();
  const int b = blockIdx.y;
  const float* x = sc + (long long)b * n;
  const unsigned bin = sel[b * 4];
  unsigned* h = hist + (long long)b * 65536;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned u = __float_as_uint(x[i]);
    if ((u >> 16) == bin) atomicAdd(&h[u & 0xffffu], 1u);
  }
}
__global__ void __launch_bounds__(1024)
pick_lo_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ sel) {
  pdl_entry();  // -> sel[2] = exact threshold bits, sel[3] = ties to take
  const int b = blockIdx.x;
  const unsigned* h = hist + (long long)b * 65536;
  const unsigned need = sel[b * 4 + 1];
  __shared__ unsigned part[1024];
  unsigned s = 0;
  for (int j = 0; j < 64; ++j) s += h[65535 - (threadIdx.x * 64 + j)];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0, lo = 0, ties = 0;
    int t = 0;
    while (t < 1024 && acc + part[t] < need) acc += part[t++];
    if (t < 1024) {
      for (int j = 0; j < 64; ++j) {
        const unsigned c = h[65535 - (t * 64 + j)];
        if (acc + c >= need) { lo = 65535 - (t * 64 + j); ties = need - acc; break; }
        acc += c;
      }
    }
    sel[b * 4 + 2] = (sel[b * 4] << 16) | lo;
    sel[b * 4 + 3] = ties;
  }
}
// compaction: everything above the threshold, plus `ties` elements equal to it
__global__ void __launch_bounds__(256)
collect_kernel(const float* __restrict__ sc, long long n, const unsigned* __restrict__ sel, int k, unsigned* __restrict__ counters,
               float* __restrict__ cand_score, long long* __restrict__ cand_idx) {
  pdl_entry();
  const int b = blockIdx.y;
  const float* x = sc + (long long)b * n;
  const unsigned thr = sel[b * 4 + 2], ties = sel[b * 4 + 3];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned u = __float_as_uint(x[i]);
    bool take = u > thr;
    if (u == thr) take = atomicAdd(&counters[b * 2 + 1], 1u) < ties;
    if (take) {
      const unsigned slot = atomicAdd(&counters[b * 2], 1u);
      if (slot < (unsigned)k) { cand_score[(long long)b * k + slot] = x[i]; cand_idx[(long long)b * k + slot] = i; }
    }
  }
}
// rank the k survivors (score descending, flat index ascending) and emit indices / relation scores
__global__ void __launch_bounds__(128)
emit_kernel(const float* __restrict__ cand_score, const long long* __restrict__ cand_idx, const unsigned* __restrict__ counters, int k,
            int N, int P, int single, const float* __restrict__ rel, const float* __restrict__ conn, int* __restrict__ inds,
            float* __restrict__ rel_scores) {
  pdl_entry();
  const int b = blockIdx.x;
  extern __shared__ unsigned char sm_raw[];
  float* sc = (float*)sm_raw;
  long long* ix = (long long*)(sm_raw + ((k * 4 + 7) / 8) * 8);
  const int cnt = min((int)counters[b * 2], k);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    sc[i] = i < cnt ? cand_score[(long long)b * k + i] : -1.f;
    ix[i] = i < cnt ? cand_idx[(long long)b * k + i] : (long long)0x7fffffffffffffffLL;
  }
  __syncthreads();
  const int W = single ? 2 : 3;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int rank = 0;
    for (int j = 0; j < k; ++j) rank += (sc[j] > sc[i]) || (sc[j] == sc[i] && ix[j] < ix[i]);
    int* dst = inds + ((long long)b * k + rank) * W;
    if (i >= cnt) {  // fewer than k candidates can only happen when n < k: pad with -1
      for (int w = 0; w < W; ++w) dst[w] = -1;
      continue;
    }
    long long f = ix[i];
    int p = 0;
    if (!single) { p = (int)(f % P); f /= P; }
    const int o = (int)(f % N), s = (int)(f / N);
    dst[0] = s; dst[1] = o;
    if (!single) dst[2] = p;
    const long long pair = ((long long)b * N + s) * N + o;
    const float c = conn ? fminf(fmaxf(conn[pair], 0.f), 1.f) : 1.f;
    if (single) {
      for (int q = 0; q < P; ++q) rel_scores[((long long)b * k + rank) * P + q] = fminf(fmaxf(rel[pair * P + q], 0.f), 1.f) * c;
    } else {
      rel_scores[(long long)b * k + rank] = fminf(fmaxf(rel[pair * P + p], 0.f), 1.f) * c;
    }
  }
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" long long egtr_triplets_scratch_bytes(int B, int N, int P, int single, int k) {
  const long long n = (long long)N * N * (single ? 1 : P);
  return (long long)B * n * 4 + (long long)B * 65536 * 4 * 2 + (long long)B * 16 + (long long)B * 8 + (long long)B * k * 12 + 256;
}

extern "C" int egtr_triplets_f32(const float* logits, const float* pred_rel, const float* pred_conn, int B, int N, int K,
                                 int num_labels, int P, int single, int k, void* scratch, float* obj_scores, int* pred_classes,
                                 int* rel_inds, float* rel_scores, egtr_stream_t s) {
  EGTR_CHECK(logits && pred_rel && scratch && obj_scores && pred_classes && rel_inds && rel_scores, EGTR_ERR_ARG, "egtr_triplets_f32: null pointer");
  EGTR_CHECK(B > 0 && N > 0 && K > 0 && P > 0 && k > 0 && k <= 1024 && num_labels > 0 && num_labels <= K && B <= 65535, EGTR_ERR_ARG,
             "egtr_triplets_f32: bad sizes (B=%d N=%d K=%d P=%d k=%d)", B, N, K, P, k);
  cudaStream_t st = (cudaStream_t)s;
  const long long n = (long long)N * N * (single ? 1 : P);
  EGTR_CHECK(n >= k, EGTR_ERR_ARG, "egtr_triplets_f32: fewer candidates (%lld) than k (%d)", n, k);
  unsigned char* base = (unsigned char*)scratch;
  float* scores = (float*)base;                       base += (long long)B * n * 4;
  unsigned* hist_hi = (unsigned*)base;                base += (long long)B * 65536 * 4;
  unsigned* hist_lo = (unsigned*)base;                base += (long long)B * 65536 * 4;
  unsigned* sel = (unsigned*)base;                    base += (long long)B * 16;
  unsigned* counters = (unsigned*)base;               base += (long long)B * 8;
  base = (unsigned char*)(((uintptr_t)base + 7) & ~(uintptr_t)7);
  long long* cand_idx = (long long*)base;             base += (long long)B * k * 8;
  float* cand_score = (float*)base;
  EGTR_CUDA(cudaMemsetAsync(hist_hi, 0, (size_t)B * 65536 * 4 * 2 + (size_t)B * 24, st));
  launch_pdl(obj_scores_kernel, dim3(cdiv((long long)B * N, 8)), dim3(256), (size_t)(0), st, logits, K, num_labels, B * N, obj_scores, pred_classes);
  const long long total = (long long)B * n;
  launch_pdl(triplet_scores_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), st, pred_rel, pred_conn, obj_scores, N, P, single, total, scores);
  int gx = cdiv(n, 256 * 8);
  if (gx > 592) gx = 592;
  launch_pdl(hist_hi_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, hist_hi);
  launch_pdl(pick_bin_kernel, dim3(B), dim3(1024), (size_t)(0), st, hist_hi, k, sel);
  launch_pdl(hist_lo_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, sel, hist_lo);
  launch_pdl(pick_lo_kernel, dim3(B), dim3(1024), (size_t)(0), st, hist_lo, sel);
  launch_pdl(collect_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, sel, k, counters, cand_score, cand_idx);
  launch_pdl(emit_kernel, dim3(B), dim3(128), (size_t)(((k * 4 + 7) / 8) * 8 + k * 8), st, cand_score, cand_idx, counters, k, N, P, single, pred_rel, pred_conn, rel_inds, rel_scores);
  for (int i = 0; i < 8; ++i) count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
