/*
 * flexam_b200 — C ABI of the B200-native (sm_100a) FlexAM denoising-step operators.
 *
 * One shared library (flexam_b200/csrc/libflexam_b200.so) exports exactly these entry points. They are
 * what a binding of the reference's hot path (FlexAM/models/wan_transformer3d_FlexAM.py, cited per
 * function as file:line relative to the reference root) calls instead of torch/cuBLAS/cuDNN/flash-attn.
 *
 * Conventions
 *   - plain pointers to DEVICE memory + sizes; no torch types, no allocation, no synchronisation and no
 *     ownership transfer inside the library. The caller owns every buffer and passes the CUDA stream
 *     (a cudaStream_t cast to void*; NULL = legacy default stream).
 *   - "bf16" buffers are uint16 storage of bfloat16; "f32" are float.
 *   - every function returns FX_OK (0) or a negative FX_ERR_* code; fx_last_error() gives the text.
 *     Launch errors are reported; asynchronous execution errors surface at the caller's next sync.
 *   - row-major everywhere; `ld*` arguments are leading dimensions in ELEMENTS.
 */
#ifndef FLEXAM_B200_H_
#define FLEXAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FX_OK 0
#define FX_ERR_ARG (-1)   /* bad pointer / size / alignment / unsupported shape */
#define FX_ERR_CUDA (-2)  /* CUDA runtime or driver error at launch time */
#define FX_ERR_ARCH (-3)  /* device is not sm_100 */

#define FX_ABI_VERSION 1

int fx_abi_version(void);
const char* fx_last_error(void);
/* FX_OK when `device` is a compute-capability 10.x part this library was built for. */
int fx_check_device(int device);

/* ---------------------------------------------------------------------------------------------------------
 * Dense projection on tcgen05/TMEM (replaces every nn.Linear on the path: self_attn.{q,k,v,o} :242-261,
 * cross_attn.{q,k,v,o} :363-370, ffn :414-416,467, head.head :506, text_embedding :959-964, and the
 * patch/ref/CNN convolutions once their input is gathered into rows :885,896,874-880).
 *
 *   acc[m,n] = sum_k A[m,k] * W[n,k]          A: bf16 [M,K] (lda), W: bf16 [N,K] (ldw) = nn.Linear.weight
 *   y        = bf16_round(acc + bias[n])      bias: bf16 [N] or NULL  (one rounding, like the autocast Linear)
 *
 * epilogue:
 *   FX_EPI_BF16        out bf16 [M,N] (ldo)  = y
 *   FX_EPI_GELU_BF16   out bf16              = bf16(gelu_tanh(y))                     (ffn.1 :415)
 *   FX_EPI_F32         out f32               = float(y)                               (patch embed -> fp32 stream)
 *   FX_EPI_RESID_F32   out f32 (in/out)     += y * gate[m,n]                          (x + y*e[2], x + y*e[5] :456,468;
 *                      gate[m,n] = gate_mod[n] + gate_e[row_idx[m]*gate_e_stride + n]   x + cross_attn :461 with no gate)
 *                      gate_mod/gate_e NULL -> that term is 0; both NULL -> gate = 1; row_idx NULL -> row 0.
 * Requirements: K % 8 == 0, N % 8 == 0, lda/ldw % 8 == 0, 16-byte aligned bases.
 */
#define FX_EPI_BF16 0
#define FX_EPI_GELU_BF16 1
#define FX_EPI_F32 2
#define FX_EPI_RESID_F32 3
#define FX_EPI_F32_EXACT 4 /* out f32 = acc + bias, no bf16 rounding (fp32 verification mode, see below) */
int fx_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out, int64_t ldo,
                 int M, int N, int K, int epilogue, const float* gate_mod, const float* gate_e,
                 int64_t gate_e_stride, const int32_t* row_idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LayerNorm (no affine, eps) + adaLN modulation + density shift, fp32 in -> bf16 out
 * (WanAttentionBlock :444-453, :464-465; Head :493-507):
 *   out[m,:] = LN(x[m,:]) * (1 + scale_mod[:] + scale_e[u,:]) + shift_mod[:] + shift_e[u,:] + (dens_mod[:] + dens[b,:])
 *   u = row_idx[m] (NULL -> 0), b = m / rows_per_batch.  scale_e/shift_e rows are e_stride floats apart,
 *   dens rows dens_stride floats apart (dens NULL -> no density term; dens_mod may be NULL).
 * fx_ln_affine: out = LN(x) * gamma + beta  (norm3 :405-407,461), gamma/beta bf16 [D].
 * D % 256 == 0, D <= 8192.
 */
int fx_ln_modulate(const float* x, void* out, int M, int D, float eps, const float* shift_mod,
                   const float* scale_mod, const float* shift_e, const float* scale_e, int64_t e_stride,
                   const int32_t* row_idx, const float* dens_mod, const float* dens, int64_t dens_stride,
                   int rows_per_batch, void* stream);
int fx_ln_affine(const float* x, void* out, int M, int D, float eps, const void* gamma, const void* beta,
                 void* stream);
/* The block-level LayerNorm sites with the modulation combined once per (distinct timestep u, sample b) instead of
 * per token (same arithmetic as fx_ln_modulate, :444-453,:464-465):
 *   fx_modulation_tables: tab f32 [2][U*B][2][D]; site s (0 = self-attention LN, 1 = ffn LN), row u*B + b:
 *     tab[s][u*B+b][0][:] = 1 + (mod[3s+1] + e0[u][3s+1]);  tab[s][u*B+b][1][:] = (mod[3s] + e0[u][3s]) + (dmod[s] + de0[b][s])
 *     mod f32 [6,D] (blocks.i.modulation), dmod f32 [2,D], e0 f32 [U,6,D], de0 f32 [B,2,D].
 *   fx_ln_scale_shift: out[m,:] = bf16(LN(x[m,:]) * scale[r,:] + shift[r,:]), r = row_idx[m] (NULL -> 0); scale/shift
 *     rows are row_stride floats apart (for the table above: scale = tab[s], shift = tab[s] + D, row_stride = 2D,
 *     row_idx[m] = u(m)*B + b(m)). */
int fx_modulation_tables(const float* mod, const float* dmod, const float* e0, const float* de0, int U, int B, int D,
                         float* tab, void* stream);
int fx_ln_scale_shift(const float* x, void* out, int M, int D, float eps, const float* scale, const float* shift,
                      int64_t row_stride, const int32_t* row_idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Full-width RMSNorm (+ optional 3-axis RoPE), in place on bf16 rows (WanRMSNorm :173-189 applied to the
 * whole D-wide row :242-243,363-364; rope_apply :135-164).
 *   r      = bf16(rsqrt(mean_D(x^2) + eps));  x = bf16(bf16(x * r) * w[:])
 *   rope (freqs != NULL): per head of 128, adjacent pairs (2j,2j+1) rotated by freqs[pos(j)][j] where
 *   pos = frame for j<22, row for 22<=j<43, col for j>=43 of token t = tok_offset + (m % rows_per_batch);
 *   frame/row/col from t over the grid (gf,gh,gw); tokens >= gf*gh*gw are left unrotated.
 *   freqs: f32 [1024][64][2] (cos,sin).  x: bf16 [M, D] with row stride ldx.  D % 256 == 0, head_dim 128.
 *   weight2 != NULL: the same launch also norms (and rotates) the next D columns [D, 2D) of every row with
 *   weight2 — q and k of the packed q|k|v projection output in one pass.
 */
int fx_rmsnorm_rope(void* x, int64_t ldx, int M, int D, float eps, const void* weight, const void* weight2,
                    const float* freqs, int gf, int gh, int gw, int tok_offset, int rows_per_batch, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Non-causal attention forward on tcgen05/TMEM, head_dim 128 (attention()/flash_attention,
 * FlexAM/models/attention_utils.py:174-233; call sites :251-256 and :367):
 *   o[b,i,h,:] = softmax_j(q[b,i,h,:].k[b,j,h,:] * scale) v[b,j,h,:],  j < Lk
 * q/k/v/o: bf16, element (b,i,h,d) at base + b*stride_b + i*stride_l + h*128 + d  (strides in elements,
 * multiples of 8). This matches [B,L,H,128] views of the packed projection outputs.
 */
int fx_fmha_fwd(const void* q, int64_t q_stride_b, int64_t q_stride_l, const void* k, int64_t k_stride_b,
                int64_t k_stride_l, const void* v, int64_t v_stride_b, int64_t v_stride_l, void* o,
                int64_t o_stride_b, int64_t o_stride_l, int B, int H, int Lq, int Lk, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Front end (patch_embedding Conv3d k=s=(1,2,2) :624-625,885; ref_conv Conv2d k=s=2 :675-678,896):
 * gathers 2x2 patches into GEMM rows, K order (c, q, r) = weight.flatten(1).
 *   src k: bf16, channel-major [C_k, F, H, W] (chan_last_k = 0) or channel-last [F, H, W, C_k] (1), k < nsrc <= 4;
 *   rows[(f*H/2 + h)*W/2 + w][(c*2 + q)*2 + r] = cat_k(src_k)[c][f][2h+q][2w+r];  rows: bf16 [F*H/2*W/2, ldr]
 */
int fx_patchify(const void* const* src, const int* channels, const int* chan_last, int nsrc, int F, int H, int W,
                void* rows, int64_t ldr, void* stream);
/* unpatchify :1126-1149 (+ ref strip :1106-1109): out[c][f][2h+q][2w+r] = head[tok(f,h,w)][(q*2+r)*C + c];
 * head: bf16 [F*H/2*W/2, ldh] (already offset past the ref tokens), out: bf16 [C, F, H, W]. */
int fx_unpatchify(const void* head, int64_t ldh, void* out, int C, int F, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * fp32 embedding MLPs (sinusoidal_embedding_1d :31-41; time_embedding/time_projection :630-632,928-944;
 * density_embedding/projection :634-636,951-955). Computed on the de-duplicated timesteps only.
 *   fx_sinusoid: out f32 [n, dim] = [cos(t*f_j) | sin(t*f_j)], f_j = 10000^(-j/(dim/2)), evaluated in fp64.
 *   fx_linear_f32: out f32 [M,N] = act_in(in f32 [M,K]) @ W[N,K]^T (bf16 weights up-cast) + bias; fp32 accumulate.
 *                  act_in: 0 none, 1 SiLU.  K % 8 == 0.
 */
int fx_sinusoid(const float* t, float* out, int n, int dim, void* stream);
/* fx_linear_f32_tc: the same fp32 linear for MANY rows (per-token timesteps that do not de-duplicate: fg/bg edit
 *   masks, pipeline_wan2_2_fun_control_FlexAM.py:686-690, 891-898; the reference runs 1.56 TFLOP of SGEMM per sample
 *   there, :928-944) on tcgen05: act_in(in) is split exactly into `planes` (1..3) bf16 planes hi [+ mid [+ lo]] written
 *   side by side into planes_ws (bf16 [M, planes*K]) and ONE GEMM accumulates all planes against the bf16 weight.
 *   planes = 2 keeps 16 significant bits of the input (relative error <= 2^-17). K % 64 == 0, N % 8 == 0.
 * fx_dedup_f32: distinct values of t[n] in order of first appearance, at most cap (<= 64): uniq f32 [cap] (unused
 *   tail = last value), inv int32 [n] (index into uniq), count int32 [1] = number found, cap + 1 = "more than cap"
 *   (inv is then meaningless). Stream-ordered, no host synchronisation: replaces torch.unique on the step's path. */
int fx_linear_f32_tc(const float* in, int64_t ldi, const void* w, int64_t ldw, const void* bias, float* out,
                     int64_t ldo, int M, int N, int K, int act_in, int planes, void* planes_ws, void* stream);
int fx_dedup_f32(const float* t, int n, int cap, float* uniq, int32_t* inv, int32_t* count, void* stream);
int fx_linear_f32(const float* in, int64_t ldi, const void* w, int64_t ldw, const void* bias, float* out,
                  int64_t ldo, int M, int N, int K, int act_in, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * CNN control fuser pieces (cnn_conv1..5 :680-711,868-881). Activations are channel-last [P, C] with
 * P = F*H*W pixels of one sample; convs run as fx_gemm_bf16 over fx_im2col3x3 rows.
 *   fx_nchw_to_nhwc: dst bf16 [P, ldd] columns [c0, c0+C) = src bf16 [C, P]
 *   fx_im2col3x3   : rows bf16 [P, 9*C], column (c*3+kh)*3+kw = in[f, y+kh-1, x+kw-1, c] (zero padded per frame)
 *   fx_groupnorm_silu: y = SiLU(GroupNorm_G(x; gamma, beta, eps)) (+ resid), stats over all pixels of the sample;
 *                    x bf16 [P, C] (conv output), y_f32 [P, C] and y_bf16 [P, C] (either may be NULL),
 *                    resid f32 [P, C] or NULL; stats: f32 workspace [2*G].
 *   fx_groupnorm_partials / fx_groupnorm_silu_partials: the same normalisation in two deterministic stages, so the
 *                    frames of a sample may live on different ranks (the fuser is sharded by frames across a
 *                    sequence-parallel group) with bit-identical statistics on every layout: partials f64 [F, G, 2] =
 *                    (sum, sum of squares) per (frame, group) of x bf16 [F*pp, C] (pp pixels per frame); the second
 *                    call sums the partials of ALL Ft frames of the sample in frame order (gathered from the peers
 *                    when sharded), then applies GroupNorm + SiLU (+ resid) to the P local pixels.
 */
int fx_nchw_to_nhwc(const void* src, void* dst, int64_t ldd, int c0, int C, int64_t P, void* stream);
int fx_im2col3x3(const void* in, void* rows, int F, int H, int W, int C, void* stream);
int fx_groupnorm_silu(const void* x, int64_t P, int C, int G, float eps, const void* gamma, const void* beta,
                      const float* resid, float* y_f32, void* y_bf16, float* stats, void* stream);
int fx_groupnorm_partials(const void* x, int F, int64_t pp, int C, int G, double* partials, void* stream);
int fx_groupnorm_silu_partials(const void* x, int64_t P, int C, int G, float eps, const void* gamma, const void* beta,
                               const double* partials, int Ft, int64_t pp, const float* resid, float* y_f32,
                               void* y_bf16, int pad_H, int pad_W, int64_t ld_bf16, float* stats, void* stream);
/* Convolution as an implicit GEMM on the tcgen05 kernels, no im2col buffer (the control fuser's cnn_conv1..4
 * :680-711,874-880; replaces cuDNN Conv3d with kernel (1,3,3)). act: bf16 channel-last, ZERO-PADDED
 * [T + kt - 1, Hp, Wp, Cin] with Hp = H + 2, Wp = W + 2 when ks == 3 (H, W when ks == 1) and kt - 1 leading frames of
 * causal history; Cin % 64 == 0 (pad channels with zeros). w: bf16 [Cout, kt*ks*ks*Cin], K order (dt, dy, dx, cin).
 * out: DENSE [T*H*W, ldo] with the fx_gemm_bf16 epilogues FX_EPI_BF16 / _GELU_BF16 / _F32 / _F32_EXACT (+ bias).
 * Tap (dt, dy, dx) reads the activation matrix shifted by dt*Hp*Wp + (dy-1)*Wp + (dx-1) rows: one TMA coordinate change
 * per tap; positions of the padded grid that are halo are computed and dropped.
 * fx_nchw_to_nhwc_padded: fx_nchw_to_nhwc into that padded layout (interior positions only: the halo stays as the caller
 * zeroed it). fx_groupnorm_silu_partials with pad_H > 0 writes y_bf16 into the padded layout [.., pad_H+2, pad_W+2, ld_bf16].
 * stride_s = 2 (ks == 3): nn.ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) of the VAE encoder (wan_vae3_8.py:101-107): out is
 * [T*(H/2)*(W/2), ldo]; stride_t = 2 (kt == 3): Conv3d((3,1,1), stride (2,1,1)) over [last cached frame | chunk]
 * (:108-109, :150-153): out is [(T/2)*H*W, ldo]. */
int fx_conv_gemm_bf16(const void* act, const void* w, const void* bias, void* out, int64_t ldo, int T, int H, int W,
                      int Cin, int Cout, int kt, int ks, int stride_s, int stride_t, int epilogue, void* stream);
int fx_nchw_to_nhwc_padded(const void* src, void* dst, int64_t ldd, int c0, int C, int F, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sampler glue (pipeline_wan2_2_fun_control_FlexAM.py:926-934): CFG combine + Euler flow step + first-frame
 * re-pin in one pass:  v = vu + s*(vc - vu);  lat = lat + dsigma*v;  lat = (1-mask)*pinned + mask*lat.
 *   vu, vc, pinned: bf16 [n]; lat: f32 [n] in/out; mask: f32 [n] or NULL.
 */
int fx_cfg_euler_step(const void* vu, const void* vc, float guidance, float dsigma, float* lat, const float* mask,
                      const void* pinned, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Ulysses sequence-parallel exchange layout (the head-scatter all-to-all around self-attention that the
 * reference delegates to the absent FlexAM/dist usp_attn_forward, :801-815): out[b][a][:] = in[a][b][:] for
 * bf16 blocks of `inner` contiguous elements (inner % 8 == 0); `in` rows may be strided (ld_a elements between
 * consecutive a). Packs [tokens][3*P head groups][Hl*128] into per-destination send buffers and unpacks
 * [P][tokens][Hl*128] back into [tokens][H*128].
 */
int fx_swap01_bf16(const void* in, int64_t ld_a, void* out, int A, int B, int inner, void* stream);

/* Fused exchange (one kernel = compute + NVLink peer stores; replaces pack + all_to_all + unpack around attention):
 *   fx_qkv_norm_rope_scatter: fx_rmsnorm_rope on the q and k thirds of the packed [M, 3D] projection output (v is
 *     moved as is) with the result written head-scattered: head h of tensor w in {q,k,v} of row (b, t),
 *     b = m / rows_per_batch, goes to peers[h / heads_per_peer] at
 *     [((b*3 + w) * dst_rows + dst_row0 + t) * heads_per_peer + h % heads_per_peer][128].
 *     peers: HOST array of n_peers (<= 8) DEVICE pointers to every rank's exchange buffer [B][3][dst_rows][D/n_peers]
 *     (peer-mapped symmetric memory; entry `own rank` is the local buffer); dst_row0 = own rank * rows_per_batch.
 *   fx_fmha_fwd_scatter: fx_fmha_fwd whose output row i is written to o_peers[i / rows_per_peer] at
 *     b*o_stride_b + (i % rows_per_peer)*o_stride_l + h*128 (bases already offset by this rank's first head).
 * The caller orders the peer writes against their consumers with a cross-rank barrier on the stream. */
int fx_qkv_norm_rope_scatter(const void* qkv, int64_t ldx, int M, int D, float eps, const void* weight_q,
                             const void* weight_k, const float* freqs, int gf, int gh, int gw, int tok_offset,
                             int rows_per_batch, void* const* peers, int n_peers, int heads_per_peer,
                             int64_t dst_rows, int dst_row0, void* stream);
int fx_fmha_fwd_scatter(const void* q, int64_t q_stride_b, int64_t q_stride_l, const void* k, int64_t k_stride_b,
                        int64_t k_stride_l, const void* v, int64_t v_stride_b, int64_t v_stride_l,
                        void* const* o_peers, int n_peers, int rows_per_peer, int64_t o_stride_b, int64_t o_stride_l,
                        int B, int H, int Lq, int Lk, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * fp32 verification mode (flexam_b200/precise.py): the same step with fp32 activations, for the north-star check
 * "<= 1e-4 relative L2 against an fp32 run" of the reference. Contractions stay on tcgen05: an fp32 matrix is split
 * EXACTLY into three bf16 planes (hi + mid + lo), each plane runs through fx_gemm_bf16 with FX_EPI_F32_EXACT and the
 * three results are added with fx_add_f32; the bf16 gather kernels above move fp32 data losslessly plane by plane.
 *   fx_split3_f32 : planes bf16 [3][M][K] (hi, mid, lo) from in f32 [M,K] (ldi); in == hi + mid + lo exactly
 *   fx_join3_f32  : out f32 [n] = (lo + mid) + hi from planes bf16 [3][n]
 *   fx_ln_f32     : fx_ln_modulate (gamma == NULL) or fx_ln_affine (gamma/beta bf16 [D]) with fp32 output, any D
 *   fx_rmsnorm_rope_f32 : fx_rmsnorm_rope on fp32 rows without the bf16 roundings (the reference's fp32 flow :186-189)
 *   fx_gelu_f32   : x = gelu_tanh(x) in place (:415)
 *   fx_gated_residual_f32 : x[m,n] += y[m,n] * gate[m,n], gate as in FX_EPI_RESID_F32 (:456,:461,:468)
 *   fx_attention_f32 : fx_fmha_fwd semantics on fp32 tensors (strides in elements), fp32 SIMT arithmetic
 *   fx_groupnorm_silu_f32 : fx_groupnorm_silu on an fp32 conv output, fp32 result
 */
int fx_split3_f32(const float* in, int64_t ldi, int M, int K, void* planes, void* stream);
int fx_join3_f32(const void* planes, int64_t n, float* out, void* stream);
int fx_ln_f32(const float* x, float* out, int M, int D, float eps, const float* shift_mod, const float* scale_mod,
              const float* shift_e, const float* scale_e, int64_t e_stride, const int32_t* row_idx,
              const float* dens_mod, const float* dens, int64_t dens_stride, int rows_per_batch, const void* gamma,
              const void* beta, void* stream);
int fx_rmsnorm_rope_f32(float* x, int64_t ldx, int M, int D, float eps, const void* weight, const void* weight2,
                        const float* freqs, int gf, int gh, int gw, int tok_offset, int rows_per_batch, void* stream);
int fx_gelu_f32(float* x, int64_t n, void* stream);
int fx_gated_residual_f32(float* x, const float* y, int M, int N, const float* gate_mod, const float* gate_e,
                          int64_t gate_e_stride, const int32_t* row_idx, void* stream);
int fx_attention_f32(const float* q, int64_t q_stride_b, int64_t q_stride_l, const float* k, int64_t k_stride_b,
                     int64_t k_stride_l, const float* v, int64_t v_stride_b, int64_t v_stride_l, float* o,
                     int64_t o_stride_b, int64_t o_stride_l, int B, int H, int Lq, int Lk, float scale, void* stream);
int fx_groupnorm_silu_f32(const float* x, int64_t P, int C, int G, float eps, const void* gamma, const void* beta,
                          const float* resid, float* y, float* stats, void* stream);

/* TeaCache residual bookkeeping on the fp32 token stream (:1003-1051): dst += src, out = a - b. */
int fx_add_f32(float* dst, const float* src, int64_t n, void* stream);
int fx_sub_f32(float* out, const float* a, const float* b, int64_t n, void* stream);

/* Sampled fingerprint of n weight tensors (DEVICE arrays ptrs[n], nbytes[n]; every tensor 16-byte aligned): out[t] =
 * order-free 64-bit hash-sum over every `stride`-th 16-byte word of tensor t. The host keeps the values taken when it
 * packed / cached anything derived from those weights and compares later: edits that bypass torch's version counters
 * (`weight.data += delta`, the reference's merge_lora / unmerge_lora, FlexAM/utils/lora_utils.py:481-485, :595-599)
 * are dense, so any sample catches them. out: uint64 [n] (zeroed inside, stream-ordered). */
int fx_fingerprint(const void* const* ptrs, const int64_t* nbytes, int n, int stride, uint64_t* out, void* stream);

/* Developer knobs by name (the FX_<NAME> environment variables seed the same table at first use): "gemm_group_m",
 * "gemm_n_span" (rasterisation of the CTA-pair GEMM; 0 / -1 = built-in choice). Not a compute-path switch. */
int fx_tune(const char* name, int value);

/* ---------------------------------------------------------------------------------------------------------
 * umT5 text encoder (FlexAM/models/wan_text_encoder.py): WanT5EncoderModel.forward :291-304 as the
 * pipeline calls it (_get_t5_prompt_embeds: 2 x 512 token ids + the tokenizer's attention mask, bf16 weights AND
 * activations). The q|k|v / o / gated-FFN projections are fx_gemm_bf16 calls; these are the remaining operators, each
 * with the reference's bf16 rounding points:
 *   fx_embedding_bf16   out bf16 [rows, D] = table bf16 [vocab, D][ids[rows]]                        (token_embedding :296)
 *   fx_t5_layernorm     out = bf16(w * bf16(x * rsqrt(mean_fp32(x^2) + eps)))  x, out bf16 [M, D]        (T5LayerNorm :51-56)
 *   fx_t5_attention     qkv bf16 [B*L, ld] = q | k | v (each H*64 wide); head_dim 64, L <= 512, NO 1/sqrt(d) scaling:
 *                       s_ij = bf16(bf16(q_i.k_j) + bias_rel[h][j-i+L-1]), keys with mask[b][j] == 0 get finfo(bf16).min,
 *                       p = bf16(softmax_fp32(s)), out[b*L+i][h*64 ..] = bf16(sum_j p_ij v_j)              (T5Attention :75-109)
 *                       bias_rel bf16 [H][2L-1] = pos_embedding[bucket(j-i)][h] (T5RelativeEmbedding :219-253), mask int32
 *                       [B, L] or NULL
 *   fx_add_bf16         x = bf16(x + y), n % 8 == 0                                       (bf16 residual stream :161-162)
 *   fx_gated_gelu_bf16  out[M, N] = bf16(fc1 * GELU(gate)), GELU as the reference's chain of bf16 tensor ops :38-41, :126;
 *                       row strides ld1 / ldg / ldo in elements (fc1 and gate are the two halves of ONE packed GEMM output)
 *   fx_t5_attention runs on tcgen05: S = Q K^T for all (<= 512) keys of a 128-query tile into the 512 TMEM columns, softmax
 *                       in TMEM, P (bf16 pairs, in place) as the A operand of O = P V
 */
int fx_embedding_bf16(const int64_t* ids, const void* table, void* out, int64_t rows, int D, int64_t vocab,
                      void* stream);
int fx_t5_layernorm(const void* x, const void* weight, void* out, int M, int D, float eps, void* stream);
int fx_t5_attention(const void* qkv, int64_t ld, const void* bias_rel, const int32_t* mask, void* out, int64_t ldo,
                    int B, int L, int H, void* stream);
int fx_add_bf16(void* x, const void* y, int64_t n, void* stream);
int fx_gated_gelu_bf16(const void* fc1, int64_t ld1, const void* gate, int64_t ldg, void* out, int64_t ldo, int64_t M,
                       int N, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Wan2.2 VAE decoder (decode half; FlexAM/models/wan_vae3_8.py). Activations are channel-last bf16
 * [frames, H, W, C]; every convolution (CausalConv3d :22-47, Resample's per-frame 3x3 :93-97 and (3,1,1) time convolution
 * :98-99, 1x1 shortcuts, the attention block's 1x1 projections) is fx_conv_gemm_bf16 / fx_gemm_bf16. The rest:
 *   fx_vae_norm_act        out[grid row of pixel p][:C] = act(x[p] / max(||x[p]||_2, 1e-12) * sqrt(C) * gamma): RMS_norm
 *                          :50-64 (+ SiLU when silu != 0) written into the zero-padded grid [frames, H+2*pad, W+2*pad, ldo]
 *                          the next convolution reads, starting at frame `frame0` (the frames before it are the causal
 *                          history the caller keeps); gamma == NULL: plain copy into that layout; pad = 0: dense rows
 *   fx_vae_upsample2x      nearest-exact 2x of [F, H, W, C] into the padded grid [F, 2H+2, 2W+2, C]        (:67-73, :93-95)
 *   fx_vae_time_interleave x[2t+k][p][c] = y[t][p][k*C+c]: the frame doubling after the time convolution    (:139-142)
 *   fx_vae_dupup_add       main += DupUp3D(x) (channel repeat + 2x2(x2) pixel shuffle, first chunk drops ft-1 frames),
 *                          bf16 add                                                                   (:395-417, :497-500)
 *   fx_softmax_rows_f32    p bf16 [rows, cols] = softmax(s f32 * scale) per row (the attention block's single head of
 *                          width C over the H*W tokens of a frame: S = Q K^T and P V are fx_gemm_bf16 calls)  (:260-282)
 *   fx_vae_unpatchify      video bf16 [3, Ttot, 2H, 2W] frames [frame0, frame0+T) = clamp(unpatchify(y [T,H,W,>=12]))
 *                                                                                                      (:304-318, :1043)
 */
int fx_vae_norm_act(const void* x, int64_t ldx, const void* gamma, void* out, int64_t ldo, int64_t npix, int C, int H,
                    int W, int pad, int frame0, int silu, void* stream);
int fx_vae_upsample2x(const void* x, void* out, int F, int H, int W, int C, void* stream);
int fx_vae_time_interleave(const void* y, void* x, int T, int64_t P, int C, void* stream);
int fx_vae_dupup_add(void* main_io, const void* x, int Tout, int H, int W, int Cin, int Cout, int ft, int drop,
                     void* stream);
int fx_softmax_rows_f32(const float* s, int64_t lds, void* p, int64_t ldp, int rows, int cols, float scale, void* stream);
int fx_vae_unpatchify(const void* y, int64_t ldy, void* video, int T, int H, int W, int Ttot, int frame0, void* stream);
/* Encoder half (:788-819, Encoder3d :505-618, Down_ResidualBlock :420-457):
 *   fx_vae_patchify     rows[(f*h+y)*w+x][c*4+r*2+q] = video bf16 [3, Ttot, 2h, 2w][c][frame0+f][2y+q][2x+r]     (:285-301)
 *   fx_vae_avgdown_add  main += AvgDown3D(x): (ft, fs, fs) blocks folded into channels, time front-padded to a multiple of
 *                       ft, mean over groups of Cin*ft*fs*fs/Cout channels; x bf16 [T,H,W,Cin], main bf16
 *                       [ceil(T/ft), H/fs, W/fs, Cout]                                                     (:340-372, :457)
 * The stride-2 convolutions of Resample downsample2d / downsample3d are fx_conv_gemm_bf16 with stride_s / stride_t = 2. */
int fx_vae_patchify(const void* video, void* rows, int64_t ldr, int T, int h, int w, int Ttot, int frame0, void* stream);
int fx_vae_avgdown_add(void* main_io, const void* x, int T, int H, int W, int Cin, int Cout, int ft, int fs, void* stream);
/* Decode across the GPUs of one box (flexam_b200/dist.py SlabExchange; the reference decodes the whole clip on every rank,
 * pipeline_wan2_2_fun_control_FlexAM.py:955-958): each rank owns a band of image rows of every padded grid
 * [frames, Hp, Wp, C]. fx_vae_halo_push copies this rank's first / last interior row of frames [frame0, frame0+T) into the
 * bottom halo row (Hp-1) of `up_grid` / the top halo row (0) of `dn_grid` — the same grid on the neighbouring ranks, peer
 * memory; NULL = no neighbour (the image border keeps its zero padding). The caller orders it against the neighbours'
 * reads (a stream barrier over the symmetric buffers). */
int fx_vae_halo_push(const void* grid, void* up_grid, void* dn_grid, int frame0, int T, int Hp, int Wp, int C, void* stream);

/* Small utility kernels used by the host glue. */
int fx_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int fx_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXAM_B200_H_ */
