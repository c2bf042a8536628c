// Relation head, pair stage (model/egtr.py:366-418, 507-516).
//
// Algebra used everywhere in this file (SURVEY.md §7 "hard parts"): with s_l(i) = subject-side and
// o_l(j) = object-side 256-vectors of "layer" l (6 projected Q/K pairs + the final hidden state),
//   gate_l(i,j)   = sigmoid(wg_s . s_l(i) + wg_o . o_l(j) + bg)                      (separable logit)
//   W1 . sum_l gate_l (s_l(i) (+) o_l(j)) = sum_l gate_l(i,j) * (U_l(i) + V_l(j)),   U_l = W1[:, :256] s_l, V_l = W1[:, 256:] o_l
// so the reference's N x N x 7 x 512 `relation_source` (573 MB / image at N = 200) is never formed:
// the per-query tensors U, V (N x 7 x 513 each, with the gate logit in the last column) are two small
// GEMMs, and the pair stage only combines them.
#include "common.cuh"

namespace egtr {
void count_launch();
namespace {

// H1[b,i,j,c] = relu(b1[c] + sum_l sigmoid(U[b,i,l,512] + V[b,j,l,512]) * (U[b,i,l,c] + V[b,j,l,c]))
// grid (ceil(N/32), N, B); 256 threads: 32 object rows j x 8 channel lanes; thread loops channels.
__global__ void __launch_bounds__(256)
pair_hidden_kernel(const float* __restrict__ U, const float* __restrict__ V, int ldu, const float* __restrict__ b1, int N, int Lr,
                   float* __restrict__ H1) {
  pdl_entry();
  __shared__ float us[7 * 516];
  __shared__ float gate[32][8];
  const int b = blockIdx.z, i = blockIdx.y, j0 = blockIdx.x * 32;
  const float* Ui = U + ((long long)b * N + i) * Lr * ldu;
  for (int t = threadIdx.x; t < Lr * 513; t += 256) {
    const int l = t / 513, c = t - l * 513;
    us[l * 516 + c] = Ui[l * ldu + c];
  }
  __syncthreads();
  const int jl = threadIdx.x >> 3, cl = threadIdx.x & 7;
  const int j = j0 + jl;
  const float* Vj = V + ((long long)b * N + min(j, N - 1)) * Lr * ldu;
  if (cl < Lr) gate[jl][cl] = sigmoidf_(us[cl * 516 + 512] + Vj[cl * ldu + 512]);
  __syncthreads();
  if (j >= N) return;
  float g[7];
#pragma unroll
  for (int l = 0; l < 7; ++l) g[l] = l < Lr ? gate[jl][l] : 0.f;
  float* out = H1 + (((long long)b * N + i) * N + j) * 512;
  for (int c = cl * 4; c < 512; c += 32) {
    float4 acc = *(const float4*)(b1 + c);
#pragma unroll
    for (int l = 0; l < 7; ++l) {
      if (l < Lr) {
        const float4 u = *(const float4*)(us + l * 516 + c);
        const float4 v = __ldg((const float4*)(Vj + l * ldu + c));
        acc.x = fmaf(g[l], u.x + v.x, acc.x); acc.y = fmaf(g[l], u.y + v.y, acc.y);
        acc.z = fmaf(g[l], u.z + v.z, acc.z); acc.w = fmaf(g[l], u.w + v.w, acc.w);
      }
    }
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    *(float4*)(out + c) = acc;
  }
}

// argmax over class logits, first maximum wins (torch.argmax)
__global__ void argmax_kernel(const float* __restrict__ logits, int K, int rows, int* __restrict__ out) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float v = logits[(long long)row * K + k];
    if (v > best) { best = v; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

// pred_rel / pred_conn from the MLP outputs: frequency bias, logit adjustment, sigmoid.
__global__ void __launch_bounds__(256)
finish_kernel(const float* __restrict__ rel_logits, int ld_rel, const float* __restrict__ conn_logits, int ld_conn,
              const int* __restrict__ cls, int K1, const float* __restrict__ triplet, const float* __restrict__ rel_dist, float tau,
              int use_freq, int logit_adj, int N, int P, long long pairs, float* __restrict__ pred_rel, float* __restrict__ pred_conn) {
  pdl_entry();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= pairs * P) return;
  const long long pair = t / P;
  const int p = (int)(t - pair * P);
  float v = rel_logits[pair * ld_rel + p];
  if (use_freq) {
    const long long bi = pair / N;  // b*N + i
    const int j = (int)(pair - bi * N);
    const long long b = bi / N;
    const int ci = cls[bi], cj = cls[b * N + j];
    v += __ldg(triplet + ((long long)ci * K1 + cj) * P + p);  // egtr.py:405-413
  }
  if (logit_adj) v -= tau * logf(rel_dist[p]);  // egtr.py:509-512
  pred_rel[t] = sigmoidf_(v);
  if (p == 0 && conn_logits != nullptr) pred_conn[pair] = sigmoidf_(conn_logits[pair * ld_conn]);
}

}  // namespace
}  // namespace egtr

using namespace egtr;

extern "C" int egtr_relation_pair_hidden_f32(const float* U, const float* V, int ldu, const float* b1, int B, int N, int Lr,
                                             float* H1, egtr_stream_t s) {
  EGTR_CHECK(U && V && b1 && H1 && B > 0 && N > 0, EGTR_ERR_ARG, "egtr_relation_pair_hidden_f32: bad arguments");
  EGTR_CHECK(Lr >= 1 && Lr <= 7 && ldu >= 513 && ldu <= 516 && ldu % 4 == 0 && N <= 65535 && B <= 65535, EGTR_ERR_UNSUPPORTED,
             "egtr_relation_pair_hidden_f32: Lr=%d ldu=%d", Lr, ldu);
  launch_pdl(pair_hidden_kernel, dim3(dim3(cdiv(N, 32), N, B)), dim3(256), (size_t)(0), (cudaStream_t)s, U, V, ldu, b1, N, Lr, H1);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_relation_finish_f32(const float* rel_logits, int ld_rel, const float* conn_logits, int ld_conn,
                                        const float* logits, int K, const float* triplet_dist, const float* rel_dist,
                                        float tau, int use_freq_bias, int logit_adjustment, int B, int N, int P,
                                        int* cls_scratch, float* pred_rel, float* pred_conn, egtr_stream_t s) {
  EGTR_CHECK(rel_logits && pred_rel && cls_scratch && (!conn_logits || pred_conn), EGTR_ERR_ARG,
             "egtr_relation_finish_f32: null pointer");
  EGTR_CHECK(!use_freq_bias || triplet_dist, EGTR_ERR_ARG, "egtr_relation_finish_f32: triplet_dist missing");
  EGTR_CHECK(!logit_adjustment || rel_dist, EGTR_ERR_ARG, "egtr_relation_finish_f32: rel_dist missing");
  if (logits != nullptr) {  // NULL: cls_scratch already holds the argmax classes (egtr_argmax_rows_f32)
    launch_pdl(argmax_kernel, dim3(cdiv((long long)B * N, 8)), dim3(256), (size_t)(0), (cudaStream_t)s, logits, K, B * N, cls_scratch);
    count_launch();
  }
  const long long pairs = (long long)B * N * N;
  launch_pdl(finish_kernel, dim3(cdiv(pairs * P, 256)), dim3(256), (size_t)(0), (cudaStream_t)s, rel_logits, ld_rel, conn_logits, ld_conn, cls_scratch, K + 1,
                                                                  triplet_dist, rel_dist, tau, use_freq_bias, logit_adjustment, N, P,
                                                                  pairs, pred_rel, pred_conn);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}

extern "C" int egtr_argmax_rows_f32(const float* x, int cols, int rows, int* out, egtr_stream_t s) {
  EGTR_CHECK(x && out && cols > 0 && rows > 0, EGTR_ERR_ARG, "egtr_argmax_rows_f32: bad arguments");
  launch_pdl(argmax_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)(0), (cudaStream_t)s, x, cols, rows, out);
  count_launch();
  EGTR_CUDA(cudaGetLastError());
  return EGTR_OK;
}
