/*
 * egtr_b200 — C ABI of the B200-native EGTR inference hot path (libegtr_b200.so).
 *
 * Every entry point takes raw DEVICE pointers, plain sizes and a CUDA stream (passed as void*,
 * i.e. a cudaStream_t), enqueues work on that stream without synchronising, and returns a status
 * code (0 = ok; `egtr_last_error()` describes the last failure on this thread).  Inputs are
 * borrowed, outputs are caller-allocated.  No torch/ATen types cross this boundary.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   egtr_msda_fwd_f32          <- ms_deform_attn_forward, model/custom_kernel/vision.cpp:13,
 *                                 ms_deform_attn.h:20-39, cuda/ms_deform_attn_cuda.cu:23-83
 *   egtr_msda_fused_fwd_f32    <- softmax + sampling-location arithmetic + the same op,
 *                                 model/deformable_detr.py:1056-1095
 *   egtr_gemm_* / egtr_conv_*  <- nn.Linear / nn.Conv2d call sites of the path (cuBLAS / cuDNN in
 *                                 the reference), model/deformable_detr.py:778,1049-1102,1166-1168,
 *                                 1255,1333-1338,1995-2011; model/egtr.py:292-293,339,355,380-416
 *   egtr_relation_*            <- the inline relation head, model/egtr.py:322-418,507-516
 *   the remaining egtr_* ops   <- ATen library kernels the reference reaches implicitly
 *                                 (layer_norm, group_norm, max_pool2d, softmax/bmm attention,
 *                                 sine position embedding, mask interpolation), SURVEY.md §2.3
 */
#ifndef EGTR_B200_H_
#define EGTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  EGTR_OK = 0,
  EGTR_ERR_ARG = 1,         /* bad shape / null pointer / unsupported size */
  EGTR_ERR_CUDA = 2,        /* a CUDA runtime or driver call failed */
  EGTR_ERR_UNSUPPORTED = 3  /* valid in the reference, not implemented here (documented) */
};

typedef void* egtr_stream_t; /* cudaStream_t */

const char* egtr_last_error(void);
int egtr_abi_version(void);
/* Launches issued by this library since the last reset (bench.py's gpu_launches). */
long long egtr_launch_count(void);
void egtr_launch_count_reset(void);

/* Internal split-K scratch is kept per slot (0..31, thread-local selection, default 0): forwards that may execute
 * concurrently on different streams (e.g. two captured CUDA graphs) must be enqueued under different slots. */
int egtr_set_scratch_slot(int slot);
/* Upper bound (1..64, process-wide) on the split-K factor of the tensor-core GEMMs.  Default 64: minimises the latency of a
 * single forward; 1 (off) is the throughput configuration, where several forwards in flight fill the SMs. */
int egtr_set_splitk_max(int max_splits);
/* Persistent GEMM grids use num_sms / div SMs (1..16, process-wide, default 1).  With several forwards in flight div > 1 runs
 * their GEMMs side by side on disjoint SMs instead of time-slicing the whole GPU (+10 % at 800x1333) — EXPERIMENTAL: round 1
 * measured nondeterministic deviations in 10-28 % of full-size forwards with div > 1; keep 1 in production. */
int egtr_set_grid_div(int div);
/* With div > 1: size each persistent grid as the smallest one that needs no more rounds of tiles than num_sms / div CTAs would
 * (default on; 0 = always take num_sms / div).  The SMs it leaves go to the kernels of the other forwards in flight. */
int egtr_set_grid_balance(int on);
/* Programmatic dependent launch: 0 off, 1 every launch, 2 (default) grids of at least one CTA per SM. */
int egtr_set_pdl_mode(int mode);
/* Diagnostic switches (tools/diag_race.py), default 0: bit 0 MSDA gathers bypass L1, bit 1 GEMM epilogues wait for full
 * completion of their bulk stores before exit, bit 2 GEMM kernels do not trigger dependent launches early. */
int egtr_set_debug_flags(int flags);

/* ---------------------------------------------------------------- GEMM-class operators ---- */
/* Left-operand source: rows of 64-float runs.  mode 0: row m = a + m*lda (+ a2 + m*lda when a2
 * is non-null, the "x + pos" of deformable_detr.py:1040,1163).  mode 1: implicit im2col over an
 * NHWC tensor [B,H,W,C] for a KHxKW/stride/pad convolution, k = (ky*KW + kx)*C + c.  mode 2: the
 * same gather over an NCHW image with few channels (the 7x7/2 stem on pixel_values [B,3,H,W]):
 * k = (ky*KW + kx)*C + c for k < KH*KW*C, zero for the padding columns up to K.  mode 3: the stem on
 * the zero-padded NHWC4 copy of the image made by egtr_pad_nchw3_to_nhwc4_f32 (C = 4, pad = 0, H/W the
 * padded sizes): k = (ky*KW + kx)*4 + c, every tap one aligned float4 and no bounds checks.  mode 4: the
 * relation head's pair stage (model/egtr.py:366-401 + layer 1 of both MLPs, factorised — DESIGN.md §4.3):
 * rows enumerate subject x object pairs in tiles of 8 x 16, a = U, a2 = V ([B*H, W*lda] per-query layer-1
 * partials with the gate logit in column 2*C; H = N_q, W = 7 "layers", C = 256, OH/OW = tiles per image side),
 * and the operand row is  relu(aux + sum_l sigmoid(U_l[i,2C] + V_l[j,2C]) * (U_l[i,c] + V_l[j,c]))  for the
 * c-range of the n-tile (block-diagonal layer 2: columns >= C of the output read channels C..2C-1). */
typedef struct {
  const float* a;
  const float* a2;
  int mode;
  int lda;
  int H, W, C, OH, OW, KH, KW, stride, pad;
  const float* aux; /* mode 4: b1 [2*C] (layer-1 bias of the relation and connectivity MLPs) */
  int fmt;          /* 0: fp32.  1: P32 (see below) — mode 0 rows and mode 1 NHWC convolutions (a 128-pixel M tile is a
                     * BW x BH patch of output pixels, each filter tap one tiled TMA box, zero padding = TMA's out-of-bounds
                     * fill) are streamed by TMA with no operand-producer warps; a2 must be NULL (fold addends into the
                     * kernel that wrote the tensor) */
} egtr_asrc_t;

/* P32 row format ("split-bf16 planes at fp32 pitch"): an fp32 [rows, C] matrix (C % 32 == 0) stored with the same
 * 4*C-byte row pitch; every group of 32 channels occupies 128 bytes = 32 bf16 `hi` values then 32 bf16 `lo` values with
 * hi = bf16_rn(x), lo = bf16_rn(x - hi), i.e. x == hi + lo to 2^-17 relative.  It is what the bf16x3 tensor-core product
 * consumes, written once by the producing kernel instead of being re-split by every consumer. */
enum { EGTR_FMT_F32 = 0, EGTR_FMT_P32 = 1, EGTR_FMT_H16PAIR = 2 };
/* H16 pair records — the MSDeformAttn `value` tensor laid out for the bilinear gather (output of the value_proj GEMM, input of
 * egtr_msda_fused_fwd_h16): fp16, head-major, [heads = N/32][rows + 1 records][2 slots][32 channels]; record r holds token r-1 in
 * slot 0 and token r in slot 1 (record 0 slot 0 and record `rows` slot 1 are padding and must hold finite values — allocate the
 * buffer zeroed).  The two x-neighbours (t, t+1) of a bilinear sample are then ONE 128-byte line (record t+1) instead of two
 * lines of an fp32 [rows, heads*32] matrix: half the L1 wavefronts per sample, the same bytes. */

/* out[orow(m)*ldo + n] = keep(act(acc + bias[n] + res[orow(m)*ldr + n]));
 * orow(m) = (m / rows_per_b)*bstride + off + m % rows_per_b when rows_per_b > 0, else m;
 * keep(x) = row_keep[orow(m)] ? x : 0 when row_keep is non-null (value.masked_fill, deformable_detr.py:1050-1052). */
typedef struct {
  const float* bias;
  const float* res;
  float* out;
  int ldo, ldr;
  int relu;
  int rows_per_b, bstride, off;
  const uint8_t* row_keep;
  /* relation-head extensions (all zero / NULL elsewhere) */
  int pair_n;           /* > 0: rows are padded 8x16 pair tiles of N_q = pair_n queries; output row = pair index */
  const float* dot_w;   /* non-NULL: output columns >= dot_col0 are not stored; the row's */
  float* dot_out;       /*   sigmoid(sum_n act(..)[n] * dot_w[n - dot_col0] + dot_b) goes to dot_out[pair] */
  float dot_b;          /*   (the connectivity MLP's last layer, model/egtr.py:416, 516) */
  int dot_col0;
  int fin;              /* 1: out = sigmoid(acc + bias + triplet[cls[s], cls[o], n] - adj[n]) (model/egtr.py:405-413, 509-515) */
  int fin_n;            /*    N_q of the finishing GEMM whose rows are pair indices (b*N + i)*N + j */
  const int* cls;       /*    argmax class per query [B*N] */
  const float* triplet; /*    [K1, K1, P] or NULL */
  const float* adj;     /*    tau * log(rel_dist) [P] or NULL */
  int k1;
  int out_fmt;          /* EGTR_FMT_F32 / EGTR_FMT_P32: storage of `out` (P32 needs N % 32 == 0) */
  int res_fmt;          /* storage of `res` */
  /* LayerNorm epilogue (P32-operand kernel, N == 256 == one tile row): out = LayerNorm(acc + bias + res) * ln_gamma + ln_beta
   * (eps 1e-5) — the Linear + residual + LayerNorm tail of an encoder sub-layer (deformable_detr.py:1325-1343) without the
   * fp32 round trip; ln_out2 (optional, P32 rows, same pitch) receives out + ln_addend (the next layer's x + pos). */
  const float* ln_gamma;
  const float* ln_beta;
  const float* ln_addend; /* fp32 rows [*, ldo] at the output row mapping */
  void* ln_out2;
} egtr_epilogue_t;

/* fp32 weight [N,K] -> split-bf16 planes [2][Npad][K] (hi, lo; rows >= N zero). Npad % 64 == 0. */
int egtr_split_weight_bf16(const float* w, int N, int K, int Npad, void* planes, egtr_stream_t s);

/* D = A * W^T on tcgen05 tensor cores (bf16x3 split products, fp32 accumulate in TMEM); W planes
 * from egtr_split_weight_bf16, streamed by TMA.  K % 64 == 0, Npad % 64 == 0. */
int egtr_gemm_sbf16(const egtr_asrc_t* a, const void* w_planes, int M, int N, int Npad, int K,
                    const egtr_epilogue_t* ep, egtr_stream_t s);

/* `groups` (<= 16) independent GEMMs of identical (M, N, K) in ONE launch: group g reads rows a_ptrs[g]
 * (+ a2_ptrs[g] when non-null; row stride lda[g]), multiplies by weight rows [n_base[g], n_base[g]+Npad) of the
 * stacked planes [2][plane_rows][K], adds ep->bias[n_base[g] + n] and writes out_ptrs[g] (+ ep->ldo, relu).
 * The pointer arrays are HOST arrays.  Replaces the 14 proj_q/proj_k/final_* Linear calls of the relation
 * head (model/egtr.py:336-397) and the decoder's q|k and v projections (model/deformable_detr.py:1166-1168). */
int egtr_gemm_sbf16_grouped(const float* const* a_ptrs, const float* const* a2_ptrs, float* const* out_ptrs,
                            const int* n_base, int groups, const int* lda, const void* w_planes, int plane_rows, int M,
                            int N, int Npad, int K, const egtr_epilogue_t* ep, egtr_stream_t s);

/* Same contract on fp32 CUDA cores with W fp32 [N,K]: the small/odd-shape path (K % 16 == 0). */
int egtr_gemm_f32(const egtr_asrc_t* a, const float* w, int M, int N, int K,
                  const egtr_epilogue_t* ep, egtr_stream_t s);

/* Few-hundred-row GEMMs (decoder queries, per-query relation tensors): latency-bound, so many small fp32
 * CUDA-core CTAs instead of 128-row tensor-core tiles.  Same grouped contract as above with fp32 weights
 * w [rows,K] (group g uses rows n_base[g] .. n_base[g]+N); groups == 1 additionally allows ep->res and the
 * row remap.  K % 32 == 0. */
int egtr_gemm_f32_grouped(const float* const* a_ptrs, const float* const* a2_ptrs, float* const* out_ptrs,
                          const int* n_base, int groups, const int* lda, const float* w, int M, int N, int K,
                          const egtr_epilogue_t* ep, egtr_stream_t s);

/* Split-K form for long-K few-row products (decoder out_proj / fc2): raw partial sums partial[s][M][N] over K slices
 * s = 0..splits-1 of (a + a2) * w^T; the consumer (egtr_sum_layernorm_f32) adds bias / residual.  K %% splits == 0. */
int egtr_gemm_f32_splitk(const float* a, const float* a2, int lda, const float* w, int M, int N, int K, int splits,
                         float* partial, egtr_stream_t s);

/* x (+ addend) [rows, C] fp32 (row stride ldx floats) -> P32 rows [rows, C] (pitch 4*C bytes).  C % 32 == 0. */
int egtr_rows_to_p32(const float* x, const float* addend, int rows, int C, int ldx, void* out, egtr_stream_t s);
/* P32 rows -> fp32 (debug / tests / handing a P32 tensor back to the caller). */
int egtr_p32_to_rows(const void* p32, int rows, int C, float* out, int ldo, egtr_stream_t s);

/* ---------------------------------------------------------------- MSDeformAttn ----------- */
/* Drop-in for ms_deform_attn_forward: value [B,S,M,D], spatial_shapes [L,2] int64 (device),
 * level_start_index [L] int64 (device), sampling_loc [B,Lq,M,L,P,2], attn_weight [B,Lq,M,L,P]
 * -> out [B,Lq,M*D].  All contiguous fp32.  D must be 32; L <= 8. */
int egtr_msda_fwd_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D,
                      int L, int Lq, int P, float* out, egtr_stream_t s);

/* Fused form: `offaw` [B*Lq, ld_offaw] holds the raw sampling offsets (M*L*P*2 floats) followed by
 * the raw attention logits (M*L*P floats) of one query row; shapes_hw is a HOST array of L (H,W)
 * pairs.  Computes softmax over L*P, loc = ref + off/(W,H) and the gather in one kernel.  value row
 * stride = ld_value floats.  Reference points: enc_ref == 0 (decoder): ref_points [Lq,2] (sigmoid
 * outputs, shared by the batch) times valid_ratios [B,L,2] (deformable_detr.py:1865-1867);
 * enc_ref != 0 (encoder): ref_points is NULL and the pixel-centre points are generated in-kernel
 * from valid_ratios (deformable_detr.py:1616-1648). */
int egtr_msda_fused_fwd_f32(const float* value, int ld_value, const int* shapes_hw, const float* offaw,
                            int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                            int B, int S, int M, int D, int L, int Lq, int P, float* out, egtr_stream_t s);

/* Same with the output written as P32 rows when out_fmt == EGTR_FMT_P32 (one head = one 32-channel group). */
int egtr_msda_fused_fwd_ex(const float* value, int ld_value, const int* shapes_hw, const float* offaw,
                           int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                           int B, int S, int M, int D, int L, int Lq, int P, void* out, int out_fmt, egtr_stream_t s);

/* The fused form on a `value` tensor stored as H16 pair records (EGTR_FMT_H16PAIR, written by the value_proj GEMM's epilogue):
 * value_h16 [heads_total][records = B*S + 1][2][32] fp16; this launch reads heads head0 .. head0 + M - 1 (decoder: the six
 * layers' value projections are one GEMM, layer i = heads 8i .. 8i+7).  Every bilinear sample costs two 128-byte L1 wavefronts
 * (top pair, bottom pair) instead of four; accumulation stays fp32. */
int egtr_msda_fused_fwd_h16(const void* value_h16, long long records, int head0, int heads_total, const int* shapes_hw,
                            const float* offaw, int ld_offaw, const float* ref_points, const float* valid_ratios, int enc_ref,
                            int B, int S, int M, int D, int L, int Lq, int P, void* out, int out_fmt, egtr_stream_t s);

/* ---------------------------------------------------------------- row-wise / image ops --- */
/* out = LayerNorm(x + res) over the last dim C (== 256), eps 1e-5; res may be NULL. */
int egtr_add_layernorm_f32(const float* x, const float* res, const float* gamma, const float* beta,
                           int rows, int C, float* out, egtr_stream_t s);
/* y = LayerNorm(x + res), C == 256; x fp32, res fp32 or P32 (res_fmt) or NULL.  Any of the three outputs may be NULL:
 * out_p32 (y as P32 rows), out_f32 (y as fp32), out_plus_p32 (y + addend as P32 rows: the `x + pos` operand of the next
 * layer's sampling_offsets / attention_weights projections, deformable_detr.py:1040). */
int egtr_add_layernorm_p32(const float* x, const void* res, int res_fmt, const float* gamma, const float* beta, int rows,
                           int C, void* out_p32, float* out_f32, const float* addend, void* out_plus_p32, egtr_stream_t s);
/* out = LayerNorm(sum_s partial[s*split_stride + row*C ..] + bias + res), C == 256: the Linear + residual + LayerNorm tail
 * of a decoder sub-layer (deformable_detr.py:1414-1417, 1441-1443, 1474-1477) on split-K sums.  out2 (optional) receives
 * a second copy at out2[(row / rows_per_b2) * bstride2 + (row % rows_per_b2) * C] (the stacked intermediate states). */
int egtr_sum_layernorm_f32(const float* partial, int splits, long long split_stride, const float* bias, const float* res,
                           const float* gamma, const float* beta, int rows, int C, float* out, float* out2, int rows_per_b2,
                           long long bstride2, egtr_stream_t s);
/* zero rows of x[rows, C] where keep[row] == 0 (value.masked_fill, deformable_detr.py:1050-1052). */
int egtr_mask_rows_f32(float* x, int ld, int C, const uint8_t* keep, int rows, egtr_stream_t s);
/* pixel_values [B,3,H,W] -> zero-bordered NHWC4 [B,H+2*pad,W+2*pad,4] (4th channel zero). */
int egtr_pad_nchw3_to_nhwc4_f32(const float* img, int B, int H, int W, int pad, float* out, egtr_stream_t s);
/* NHWC 3x3/2 pad 1 max-pool. */
int egtr_maxpool3x3s2_nhwc_f32(const float* x, int B, int H, int W, int C, float* out, egtr_stream_t s);
/* Same; out_fmt == EGTR_FMT_P32 writes the pooled map as P32 rows (C % 32 == 0). */
int egtr_maxpool3x3s2_nhwc_ex(const float* x, int B, int H, int W, int C, void* out, int out_fmt, egtr_stream_t s);
/* GroupNorm(32 groups) in place over x[B, rows_per_b (at row offset `off`, batch stride `bstride`
 * rows), C], eps 1e-5 (deformable_detr.py:1996). */
int egtr_groupnorm_f32(float* x, int B, int rows_per_b, int bstride, int off, int C, int groups,
                       const float* gamma, const float* beta, double* scratch, egtr_stream_t s);
/* Same, additionally writing the normalised rows as P32 (out_p32) and rows + addend as P32 (out_plus_p32; the encoder's
 * x + pos operand) at the same [B, bstride rows, C] mapping; either may be NULL. */
int egtr_groupnorm_ex(float* x, int B, int rows_per_b, int bstride, int off, int C, int groups, const float* gamma,
                      const float* beta, double* scratch, void* out_p32, const float* addend, void* out_plus_p32,
                      egtr_stream_t s);
/* doubles of scratch egtr_groupnorm_f32 needs for (B, rows_per_b). */
long long egtr_groupnorm_scratch_doubles(int B, int rows_per_b);
/* pixel_mask [B,H,W] int64 -> per-level nearest-neighbour masks (uint8 [B,S]), sine position
 * embedding + level_embed ([B,S,C]), valid ratios [B,L,2] (deformable_detr.py:783-785, 850-876,
 * 2064-2073, 2262).  shapes_hw: HOST array of L (h,w).  dim_t: device table of C/2 floats,
 * 10000^(2*(i//2)/(C/2)) (deformable_detr.py:860-865). */
int egtr_levels_geometry_f32(const int64_t* pixel_mask, int B, int H, int W, const int* shapes_hw, int L,
                             const float* level_embed, const float* dim_t, int C, uint8_t* mask_flat, float* pos_flat,
                             float* valid_ratios, float* scratch /* 2*B*S floats */, egtr_stream_t s);
/* Decoder self-attention core: qkv [B*N, ld] with q (already scaled) at col 0, k at col C, v at
 * col 2C; softmax(q k^T) v per head -> out [B*N, C] (deformable_detr.py:1190-1253). */
int egtr_mha_core_f32(const float* qkv, int ld, int B, int N, int heads, int D, float* out, egtr_stream_t s);
/* y[r, n] = act(x[r, :K] . w[n, :K] + b[n]) for tiny N (<= 8), one warp per row.  act 0: none,
 * 1: sigmoid, 2: bbox head — add inverse_sigmoid(ref[r, n]) (eps 1e-5) to n < 2, then sigmoid
 * (model/egtr.py:291-303, model/deformable_detr.py:658-662). */
int egtr_small_linear_f32(const float* x, int ldx, const float* w, const float* b, int rows, int K, int N,
                          int act, const float* ref, int ld_ref, int ref_rows /* ref row = r % ref_rows */,
                          float* y, int ldy, egtr_stream_t s);

/* ---------------------------------------------------------------- relation head ---------- */
/* Pair stage of the relation head (model/egtr.py:366-416, 507-516) given the per-query tensors
 *   U [B,N,7,ldu]: subject-side layer-1 partials (rel 0..255 | conn 256..511) then gate scalar at 512
 *   V [B,N,7,ldu]: object-side partials, same layout (gate bias folded into V's scalar)
 * writes H1 [B*N*N, 512] = relu(b1 + sum_l sigmoid(a_l(i)+b_l(j)) * (U_l(i) + V_l(j))).
 * (bring-up path; the fused kernel below never materialises H1) */
int egtr_relation_pair_hidden_f32(const float* U, const float* V, int ldu, const float* b1, int B, int N, int Lr,
                                  float* H1, egtr_stream_t s);
/* pred_rel = sigmoid(rel_logits + triplet_dist[c_i, c_j, :] - tau*log(rel_dist)), c = argmax(logits);
 * pred_conn = sigmoid(conn_logits).  rel_logits [B*N*N, ld_rel], conn_logits [B*N*N, ld_conn].  conn_logits may be NULL
 * (connectivity already finished elsewhere); logits may be NULL when cls_scratch already holds the argmax classes. */
int egtr_relation_finish_f32(const float* rel_logits, int ld_rel, const float* conn_logits, int ld_conn,
                             const float* logits, int K, const float* triplet_dist, const float* rel_dist,
                             float tau, int use_freq_bias, int logit_adjustment, int B, int N, int P,
                             int* cls_scratch /* B*N ints */, float* pred_rel, float* pred_conn, egtr_stream_t s);

/* ---- The relation head behind one entry point (model/egtr.py:322-418, 507-516) ----
 * Weights prepared once at load.  The per-query stage composes proj_q[l] / proj_k[l] / final_sub_proj / final_obj_proj with
 * layer 1 of both MLPs and the gate (two stacked Linears with nothing in between, composed in fp64 by the host):
 * group l < layers gives U_l (subject side), group layers + l gives V_l (object side, gate bias folded in); each group has 513
 * output rows: relation MLP 0..255 | connectivity MLP 256..511 | gate logit 512. */
typedef struct {
  int layers;            /* decoder_layers + 1, <= 7 */
  const void* uv_planes; /* egtr_split_weight_bf16 planes of the stacked [2*layers*uv_npad, 256] weight */
  const float* uv_bias;  /* [2*layers*uv_npad] */
  int uv_npad;           /* rows per group in the stack (>= 513, % 64 == 0) */
  const float* b1;       /* [512] layer-1 bias: rel_predictor.layers.0.bias | connectivity_layer.layers.0.bias */
  const void* w2g;       /* egtr_pack_weight_p32g of [512,256]: rel_predictor.layers.1.weight | connectivity_layer.layers.1.weight */
  const float* b2;       /* [512] */
  const void* w3g;       /* egtr_pack_weight_p32g of rel_predictor.layers.2.weight [P,256] into p3 rows (see below) */
  const float* b3;       /* [P] */
  const float* w3c;      /* [256] connectivity_layer.layers.2.weight */
  float b3c;             /* connectivity_layer.layers.2.bias */
} egtr_relhead_weights_t;

/* fp32 [N,K] rows -> bf16 "P32 group" rows for the fused kernel's TMA boxes: output row r = K/32 groups of (32 hi | 32 lo),
 * taken from source row perm[r] (DEVICE int array; NULL: r); rows whose source is outside [0,N) are zero.  K % 32 == 0.
 * Layer 3 (w3g): P <= 64 -> rows_out = 64 and perm[r] = 32*((r%32)/16) + 16*(r/32) + r%16 (each CTA of a pair feeds 16 rows to
 * each of the two N = 32 MMAs); P > 64 -> rows_out = 64*ceil(P/64), no permutation. */
int egtr_pack_weight_p32g(const float* w, int N, int K, int rows_out, const int* perm, void* out, egtr_stream_t s);

/* Pair stage as ONE kernel (relhead.cu): U, V [B*N, layers, ldu] per-query partials (ldu % 4 == 0, >= 513) ->
 * pred_rel [B,N,N,P] = sigmoid(MLP_rel(h1) + triplet_dist[cls_i, cls_j] - tau*log(rel_dist)),
 * pred_conn [B,N,N] = sigmoid(MLP_conn(h1)), h1 = relu(b1 + sum_l sigmoid(g_l(i,j)) (U_l(i) + V_l(j))).
 * The gated pair tensor, both hidden layers and the logits stay in shared memory / TMEM.  cls == NULL: no frequency bias;
 * rel_dist == NULL: no logit adjustment.  k1 = row length of triplet_dist's class axes (num_labels + 1). */
int egtr_relation_pairs_fused_f32(const float* U, const float* V, int ldu, int layers, const egtr_relhead_weights_t* w,
                                  const int* cls, const float* triplet_dist, int k1, const float* rel_dist, float tau,
                                  int B, int N, int P, float* pred_rel, float* pred_conn, egtr_stream_t s);

/* The whole head: q_ptrs[l] / k_ptrs[l] (HOST arrays of layers-1 device pointers) are the captured decoder self-attention
 * queries (scaled, as captured) / keys of layer l as rows [B*N, 256] with row stride ld_qk; h_last [B*N, 256] (stride ld_h) the
 * decoder output; logits [B*N, K] the class logits (argmax -> frequency-bias classes).  Scratch: U, V [B*N*layers*516] floats,
 * cls [B*N] ints.  Three launches: the 2*layers per-query projections, the class argmax, the fused pair kernel. */
int egtr_relation_head_fwd_f32(const float* const* q_ptrs, const float* const* k_ptrs, int ld_qk, const float* h_last, int ld_h,
                               const float* logits, int K, const egtr_relhead_weights_t* w, const float* triplet_dist,
                               const float* rel_dist, float tau, int use_freq_bias, int logit_adjustment, int B, int N, int P,
                               float* U_scratch, float* V_scratch, int* cls_scratch, float* pred_rel, float* pred_conn,
                               egtr_stream_t s);

/* ---------------------------------------------------------------- ResNet stem on TMA (stem.cu) */
/* conv1 7x7 / stride 2 / pad 3 + folded FrozenBN + ReLU (timm resnet50 stem, model/deformable_detr.py:772-787) without a gather:
 * the image is stored as zero-bordered NHWC4 bf16 (hi, lo) in which a filter row of an output pixel's window is a contiguous run,
 * so 128 windows are one TMA box over a tensor map with overlapping rows.  planes: egtr_stem_planes_bytes(B, H, W) bytes,
 * ZERO-INITIALISED once by the caller (the slack behind the last row is read against zero weights), 128-byte aligned, filled by
 * egtr_stem_pad_split_bf16.  Weights, bf16 [2][64][7 * krow], krow = egtr_stem_krow(), by egtr_stem_layout():
 *   1 (two image planes, krow 32; the default): row o of plane 0 / 1 = hi / lo of w[o][ky][4 * kx + c] (egtr_split_weight_bf16 of that matrix);
 *   2 (one plane, hi and lo interleaved per pixel, krow 64; build-time A/B variant, -DEGTR_STEM_LAYOUT=2): element ky*64 + 8*kx + j of row o of set 0 =
 *     hi(w[o][c][ky][kx]) for j = c and for j = 4 + c, of set 1 = lo(w[o][c][ky][kx]) for j = c; zeros elsewhere (c < 3, kx < 7).
 * out: fp32 NHWC [B, OH, OW, 64]. */
long long egtr_stem_planes_bytes(int B, int H, int W);
int egtr_stem_krow(void);
int egtr_stem_layout(void);
int egtr_stem_pad_split_bf16(const float* img_nchw, int B, int H, int W, void* planes, egtr_stream_t s);
int egtr_stem_conv7x7s2_bf16x3(const void* planes, int B, int H, int W, const void* w_planes, const float* bias, float* out, egtr_stream_t s);

/* ---------------------------------------------------------------- fused decoder stack (decoder.cu) */
/* Replaces, for small query sets (N <= 256), the launch sequence of DeformableDetrDecoder.forward
 * (model/deformable_detr.py:1774-1968; layer 1390-1489; self-attention with Q/K capture 1149-1262) by ONE kernel: a cluster of
 * eight CTAs per image, every GEMM on tcgen05.  Weights are stacked over the layers as bf16 hi/lo planes [2][L*rows][K]
 * (egtr_split_weight_bf16 with Npad = L*rows): */
typedef struct {
  int layers;              /* L */
  int n_queries;           /* N */
  const void* w_qkv;       /* rows per layer 768, head-major: head r = q_r (32 rows, pre-scaled by head_dim^-0.5) | k_r | v_r; K = 256 */
  const void* w_o;         /* 256 rows: self_attn.out_proj */
  const void* w_offaw;     /* 384 rows, head-major: head r = its 32 encoder_attn.sampling_offsets rows | its 16 attention_weights rows */
  const void* w_out;       /* 256 rows: encoder_attn.output_proj */
  const void* w_fc1;       /* 1024 rows */
  const void* w_fc2;       /* 256 rows, K = 1024 */
  const float* vec;        /* [L][3328]: out_proj.bias | output_proj.bias | fc2.bias | ln1 w,b | ln2 w,b | ln3 w,b (256 each) | fc1.bias (1024) */
  const float* qkv_pos;    /* [L][N][768] = query_pos . Wq|Wk^T + bias (v columns: bias only), columns q | k | v as captured */
  const float* off_pos;    /* [L][N][384] = query_pos . Woffaw^T + bias */
  const float* tgt;        /* [N][256] learned query embeddings (layer-0 input) */
  const float* ref_points; /* [N][2] sigmoid'ed reference points */
} egtr_decoder_weights_t;

long long egtr_decoder_scratch_bytes(int B, int N);
/* Diagnostic: code of the barrier wait that timed out inside the decoder kernel (0: none); readable after the launch failure. */
int egtr_decoder_fault(void);
/* Diagnostic: device buffer of 1 + layers*12 uint64 that receives %globaltimer at kernel start and after every phase (NULL: off). */
int egtr_decoder_debug_profile(unsigned long long* dev_buf);
/* value_h16: the six cross-attention value tensors as fp16 pair records [layers*8 heads][B*S + 1][2][32] (EGTR_FMT_H16PAIR),
 * head index = layer*8 + head.  Outputs: qkv_out [L][B*N][768] (q scaled | k | v of every layer's self-attention: the captured
 * states of model/egtr.py:322-345), inter [B][L][N][256] (hidden state after every layer).  scratch: 1 KB aligned,
 * egtr_decoder_scratch_bytes(B, N) bytes, ZERO-INITIALISED once by the caller.  Layers [layer0, layer1); phase0 / phase1 bound
 * the phases of the first / last of them (0 .. 12: init, qkv, mha, out_proj, ln1, offsets, msda, output_proj, ln2, fc1, fc2,
 * ln3) — the full stack is (0, L, 0, 12).  mha_mode 1: self-attention core inside the kernel; 0: skipped (bring-up: the caller
 * runs egtr_mha_core between two launches and provides attn as P32 rows at scratch offset 7*B*N KB). */
int egtr_decoder_fused_f32(const egtr_decoder_weights_t* w, void* scratch, const void* value_h16, long long records,
                           const int* shapes_hw, int n_levels, const float* valid_ratios, int B, int S, float* qkv_out,
                           float* inter, int layer0, int layer1, int phase0, int phase1, int mha_mode, egtr_stream_t s);

/* ---------------------------------------------------------------- triplet extraction */
/* Device-side equivalent of the model-output post-processing in train_egtr.py:56-94 (multiple predicates per pair,
 * single == 0) and 56-69 + 120-128 (one entry per pair, single == 1): object scores/classes from softmax(logits)
 * [:, :num_labels], sub_ob = outer(obj, obj) without diagonal, rel = clamp(pred_rel) * clamp(pred_conn) (pred_conn
 * may be NULL), top-k of rel * sub_ob (resp. max_p rel * sub_ob) by a radix select instead of a full argsort.
 * Outputs: obj_scores [B,N], pred_classes [B,N] int32, rel_inds [B,k,3] (s,o,p) or [B,k,2] (s,o) int32 sorted by
 * descending score, rel_scores [B,k] or [B,k,P].  scratch: egtr_triplets_scratch_bytes() bytes of device memory. */
long long egtr_triplets_scratch_bytes(int B, int N, int P, int single, int k);
int egtr_triplets_f32(const float* logits, const float* pred_rel, const float* pred_conn, int B, int N, int K,
                      int num_labels, int P, int single, int k, void* scratch, float* obj_scores, int* pred_classes,
                      int* rel_inds, float* rel_scores, egtr_stream_t s);

/* ---------------------------------------------------------------- input staging */
/* PIL-exact bilinear resampling of uint8 images on the device (Pillow src/libImaging/Resample.c, 8-bit path): `bounds`
 * [out][2] = (first input index, taps) and `kk` [out][ksize] = taps in 22-bit fixed point, both DEVICE arrays built by the host
 * (egtr_b200/preprocess.py, same arithmetic as Pillow's precompute_coeffs / normalize_coeffs_8bpc).
 * Horizontal pass: src [H,W,C] -> dst [H,OW,C]. */
int egtr_resample_h_u8(const uint8_t* src, int H, int W, int C, int OW, const int* bounds, const int* kk, int ksize,
                       uint8_t* dst, egtr_stream_t s);
/* Vertical pass + DetrFeatureExtractor's rescale / normalise ((x / 255 - mean) / std, fp32, IEEE divisions; mean3 / std3 are
 * HOST arrays) written into a [3, *, row_stride] region of the zero-padded NCHW batch tensor (plane_stride floats between
 * channels); also sets pixel_mask (int64, 1 = real pixel) over the [OH, W] region when mask is non-NULL. */
int egtr_resample_v_normalize_f32(const uint8_t* src, int H, int W, int C, int OH, const int* bounds, const int* kk,
                                  int ksize, const float* mean3, const float* std3, float* dst, long long plane_stride,
                                  int row_stride, int64_t* mask, int mask_row_stride, egtr_stream_t s);

/* out[r] = argmax over x[r, :cols], first maximum wins (torch.argmax; model/egtr.py:406). */
int egtr_argmax_rows_f32(const float* x, int cols, int rows, int* out, egtr_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* EGTR_B200_H_ */
