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

// ---------------------------------------------------------------- radix select (three passes: 11 + 11 + 10 bits)
// Scores are >= 0, so their IEEE bit patterns order like unsigned integers.  Pass p histograms the next digit of the elements
// that match the digits chosen so far, in SHARED memory (2048 bins per CTA, warp-aggregated adds: scores cluster in a few
// exponents, so most lanes of a warp hit the same bin), flushes the non-empty bins to the image's global histogram, and the
// NEXT kernel picks the digit that holds the k-th largest element (every CTA repeats the 2048-bin suffix scan: cheaper than a
// launch).  sel[b] = {digits chosen so far, elements still needed inside that prefix}.
constexpr int BINS = 2048;
struct Pick { unsigned prefix, need, count; };

// suffix scan over hist[BINS] from the top: the bin where the running count first reaches `need`; 256 threads
__device__ __forceinline__ Pick pick_digit(const unsigned* __restrict__ hist, unsigned need, unsigned* sm /* [256 + 4] */) {
  const int t = threadIdx.x;
  unsigned loc[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { loc[j] = hist[BINS - 1 - (t * 8 + j)]; s += loc[j]; }  // thread t owns the 8 bins below BINS - 8t
  sm[t] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {  // inclusive scan over threads (counts from the top)
    const unsigned v = t >= o ? sm[t - o] : 0;
    __syncthreads();
    sm[t] += v;
    __syncthreads();
  }
  const unsigned incl = sm[t], excl = incl - s;
  if (excl < need && incl >= need) {  // exactly one thread (need >= 1 and the total is >= need)
    unsigned acc = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (acc + loc[j] >= need) { sm[256] = BINS - 1 - (t * 8 + j); sm[257] = need - acc; sm[258] = loc[j]; break; }
      acc += loc[j];
    }
  }
  __syncthreads();
  Pick p = {sm[256], sm[257], sm[258]};
  __syncthreads();
  return p;
}

template <int PASS>
__global__ void __launch_bounds__(256)
select_pass_kernel(const float* __restrict__ sc, long long n, int k, unsigned* __restrict__ hist /* [B][3][BINS] */, unsigned* __restrict__ sel /* [B][8] */) {
  pdl_entry();
  __shared__ unsigned h[BINS];
  __shared__ unsigned sm[260];
  const int b = blockIdx.y;
  const float* x = sc + (long long)b * n;
  unsigned* H = hist + (long long)b * 3 * BINS;
  unsigned prefix = 0;
  if (PASS > 0) {  // digit(s) chosen by the previous pass(es)
    const unsigned need = PASS == 1 ? (unsigned)k : sel[b * 8 + 1];
    const Pick p = pick_digit(H + (PASS - 1) * BINS, need, sm);
    prefix = PASS == 1 ? p.prefix : ((sel[b * 8 + 0] << 11) | p.prefix);
    if (blockIdx.x == 0 && threadIdx.x == 0) { sel[b * 8 + (PASS == 1 ? 0 : 2)] = prefix; sel[b * 8 + (PASS == 1 ? 1 : 3)] = p.need; }
    // (PASS 1 writes sel[0..1] = first digit / need; PASS 2 writes sel[2..3] = first two digits / need: distinct words, so CTAs of
    //  this launch that still read sel[0..1] are not disturbed)
  }
  for (int i = threadIdx.x; i < BINS; i += 256) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * 1024;
  for (long long i0 = blockIdx.x * 1024ll; i0 < n; i0 += stride) {
    // four coalesced loads in flight per thread before the (serialising) warp votes
    unsigned u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = i0 + e * 256 + threadIdx.x;
      u[e] = i < n ? __float_as_uint(x[i]) : 0xffffffffu;  // 0xffffffff: a NaN pattern no score has
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unsigned key = 0xffffffffu;  // sentinel: not counted
      if (u[e] != 0xffffffffu) {
        if (PASS == 0) key = u[e] >> 21;
        else if (PASS == 1) { if ((u[e] >> 21) == prefix) key = (u[e] >> 10) & 2047u; }
        else { if ((u[e] >> 10) == prefix) key = u[e] & 1023u; }
      }
      if (PASS == 0) {  // every element counts and scores cluster in a few bins: one add per distinct bin of the warp
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&h[key], (unsigned)__popc(peers));
      } else if (key != 0xffffffffu) {  // later passes see only the elements inside the chosen digit: rare, plain adds
        atomicAdd(&h[key], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BINS; i += 256)
    if (h[i]) atomicAdd(&H[PASS * BINS + i], h[i]);
}

// Elements equal to the threshold, counted per CTA chunk (contiguous index ranges): the compaction below takes the `ties`
// SMALLEST flat indices among them, so the selected set — not only its order — is the same on every run and equals a stable
// descending argsort (atomics used to pick ties by arrival order).
__global__ void __launch_bounds__(256)
tie_count_kernel(const float* __restrict__ sc, long long n, int k, const unsigned* __restrict__ hist, unsigned* __restrict__ sel,
                 unsigned* __restrict__ tie_counts /* [B][gridDim.x] */) {
  pdl_entry();
  __shared__ unsigned sm[260];
  const int b = blockIdx.y;
  const Pick p = pick_digit(hist + ((long long)b * 3 + 2) * BINS, sel[b * 8 + 3], sm);
  const unsigned thr = (sel[b * 8 + 2] << 10) | p.prefix;
  if (blockIdx.x == 0 && threadIdx.x == 0) { sel[b * 8 + 4] = thr; sel[b * 8 + 5] = p.need; sel[b * 8 + 6] = p.count; }
  const float* x = sc + (long long)b * n;
  const long long chunk = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  unsigned c = 0;
  for (long long i0 = lo; i0 < hi; i0 += 1024) {
    unsigned u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = i0 + e * 256 + threadIdx.x;
      u[e] = i < hi ? __float_as_uint(x[i]) : ~thr;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) c += u[e] == thr;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    tie_counts[(long long)b * gridDim.x + blockIdx.x] = t;
  }
}

// compaction: everything above the threshold, plus the `ties` lowest-index elements equal to it
__global__ void __launch_bounds__(256)
collect_kernel(const float* __restrict__ sc, long long n, const unsigned* __restrict__ sel, const unsigned* __restrict__ tie_counts, int k,
               unsigned* __restrict__ counters, float* __restrict__ cand_score, long long* __restrict__ cand_idx) {
  pdl_entry();
  __shared__ unsigned sm[16];
  const int b = blockIdx.y;
  const float* x = sc + (long long)b * n;
  const unsigned thr = sel[b * 8 + 4], ties = sel[b * 8 + 5];
  // rank of this chunk's first tie among all ties of the image (ascending index)
  unsigned before = 0;
  for (int c = threadIdx.x; c < (int)blockIdx.x; c += 256) before += tie_counts[(long long)b * gridDim.x + c];
  before = __reduce_add_sync(0xffffffffu, before);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = before;
  __syncthreads();
  unsigned base = 0;
  for (int w = 0; w < 8; ++w) base += sm[w];
  __syncthreads();
  const long long chunk = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  for (long long i0 = lo; i0 < hi; i0 += 256) {
    const long long i = i0 + threadIdx.x;
    const unsigned u = i < hi ? __float_as_uint(x[i]) : 0u;
    const bool eq = i < hi && u == thr;
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) sm[8 + warp] = __popc(bal);
    __syncthreads();
    unsigned wbase = 0, tile = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) wbase += sm[8 + w]; tile += sm[8 + w]; }
    const bool take = (i < hi && u > thr) || (eq && base + wbase + __popc(bal & ((1u << lane) - 1u)) < ties);
    if (take) {
      const unsigned slot = atomicAdd(&counters[b * 2], 1u);
      if (slot < (unsigned)k) { cand_score[(long long)b * k + slot] = x[i]; cand_idx[(long long)b * k + slot] = i; }
    }
    base += tile;
    __syncthreads();
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

constexpr int SELECT_GRID = 592;  // CTAs per image of the select / tie / collect passes (4 x 148)

extern "C" long long egtr_triplets_scratch_bytes(int B, int N, int P, int single, int k) {
  const long long n = (long long)N * N * (single ? 1 : P);
  return (long long)B * n * 4 + (long long)B * 3 * BINS * 4 + (long long)B * 32 + (long long)B * 8 + (long long)B * SELECT_GRID * 4 +
         (long long)B * k * 12 + 256;
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
  unsigned* hist = (unsigned*)base;                   base += (long long)B * 3 * BINS * 4;
  unsigned* sel = (unsigned*)base;                    base += (long long)B * 32;
  unsigned* counters = (unsigned*)base;               base += (long long)B * 8;
  unsigned* tie_counts = (unsigned*)base;             base += (long long)B * SELECT_GRID * 4;
  base = (unsigned char*)(((uintptr_t)base + 7) & ~(uintptr_t)7);
  long long* cand_idx = (long long*)base;             base += (long long)B * k * 8;
  float* cand_score = (float*)base;
  EGTR_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * 3 * BINS * 4 + (size_t)B * 40, st));  // histograms, sel, counters
  launch_pdl(obj_scores_kernel, dim3(cdiv((long long)B * N, 8)), dim3(256), (size_t)(0), st, logits, K, num_labels, B * N, obj_scores, pred_classes);
  const long long total = (long long)B * n;
  launch_pdl(triplet_scores_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), st, pred_rel, pred_conn, obj_scores, N, P, single, total, scores);
  int gx = cdiv(n, 256 * 8);
  if (gx > SELECT_GRID) gx = SELECT_GRID;
  launch_pdl(select_pass_kernel<0>, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, k, hist, sel);
  launch_pdl(select_pass_kernel<1>, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, k, hist, sel);
  launch_pdl(select_pass_kernel<2>, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, k, hist, sel);
  launch_pdl(tie_count_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, k, (const unsigned*)hist, sel, tie_counts);
  launch_pdl(collect_kernel, dim3(dim3(gx, B)), dim3(256), (size_t)(0), st, scores, n, (const unsigned*)sel, (const unsigned*)tie_counts, k, counters,
             cand_score, cand_idx);
  launch_pdl(emit_kernel, dim3(B), dim3(128), (size_t)(((k * 4 + 7) / 8) * 8 + k * 8), st, cand_score, cand_idx, counters, k, N, P, single, pred_rel, pred_conn, rel_inds, rel_scores);
  for (int i = 0; i < 8; ++i) count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
