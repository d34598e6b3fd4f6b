/* ltx_b200.h — C ABI of libltx_b200.so: hand-written sm_100a kernels for the LTX-Video / Wan2.1
 * denoising hot path.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * Conventions (SURVEY.md §8b "What a C-ABI replacement must export"):
 *   - every pointer is a DEVICE pointer unless named host_*; bf16 tensors are `const void*`
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - functions never allocate, never synchronise, and are thread-safe on distinct streams
 *   - return 0 on success, a negative LTXB200_ERR_* code otherwise (never throw)
 *   - leading dimensions / strides are in ELEMENTS
 *
 * The reference has no FFI: every entry below replaces a PyTorch call site, cited as file:line of
 * /root/reference (soasme/LTX-Video-GPUPoor).
 */
#ifndef LTX_B200_H_
#define LTX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTXB200_OK 0
#define LTXB200_ERR_BAD_SHAPE (-1)
#define LTXB200_ERR_BAD_ALIGN (-2)
#define LTXB200_ERR_CUDA (-3)
#define LTXB200_ERR_TENSORMAP (-4)
#define LTXB200_ERR_UNSUPPORTED (-5)

#define LTXB200_ACT_NONE 0
#define LTXB200_ACT_GELU_TANH 1
#define LTXB200_ACT_SILU 2
#define LTXB200_ACT_GELU_ERF 3   /* exact GELU (torch.nn.GELU()): Wan MLPProj, wan/modules/model.py:583 */

#define LTXB200_CONV_STORE_NDHWC 0   /* out[b,t,h,w,co] bf16 */
#define LTXB200_CONV_STORE_D2S 1     /* depth-to-space 2x2x2, first frame dropped; co order (p1,p2,p3,c) */
#define LTXB200_CONV_STORE_UNPATCH 2 /* unpatchify 4x4 to [b,c,t,4h,4w]; co order (c,q,r) */

int ltxb200_abi_version(void);
const char* ltxb200_error_string(int code);
/* number of kernels launched through this library since load (per process); for bench.py's gpu_launches */
long long ltxb200_launch_count(void);

/* out[M,N] = epilogue(A[M,K] @ W[N,K]^T): tcgen05/TMEM GEMM, TMA-fed, fp32 accumulate.
 * epilogue: v = acc + bias[n]; v = act(v); v = v * gate[m / rows_per_gate, n]; v = v + residual[m, n].
 * Replaces every nn.Linear on the path: ltx_video/models/transformers/attention.py:543-558,1040-1059,
 * 1147,339-340 (FFN), transformer3d.py:418,448,503; wan/modules/model.py:168-171,390-391; fused gate /
 * residual replace attention.py:282-288,345-351.  out may alias residual.  K % 8 == 0, N % 8 == 0. */
int ltxb200_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                      void* out, int64_t ldc, int out_f32, const void* bias, int act,
                      const void* residual, int64_t ldr, const void* gate, int64_t gate_ld,
                      int rows_per_gate, void* stream);

/* 3x3x3 convolution as TMA-tiled implicit GEMM on tensor cores; x is NDHWC bf16 [B,T,H,W,Cin],
 * w is [Cout, 27*Cin] bf16 with k = ((kt*3+kh)*3+kw)*Cin + ci.  Spatial zero padding, temporal replicate
 * padding (causal: 2 leading frames; non-causal: 1 each side).  Replaces CausalConv3d.forward
 * (ltx_video/models/autoencoders/causal_conv3d.py:44-59), DepthToSpaceUpsample.forward pixel-shuffle
 * + frame drop (causal_video_autoencoder.py:1051-1065), the residual add (:1256) and the final
 * unpatchify (:1282-1299) via store_mode.  Cin % 64 == 0, Cout % 8 == 0. */
int ltxb200_conv3d_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W,
                        int Cin, int Cout, int causal, int store_mode, int out_f32, const void* residual,
                        void* stream);

/* Non-causal softmax attention, layout [B, L, H, d] (d = 64 or 128), fp32 softmax, bf16 in/out.
 * q/k/v may be strided views (token stride ld*, batch stride bs*, head stride = d).
 * key_bias: optional fp32 [B, Lk] additive mask bias (e.g. 0 / -10000).  scale <= 0 means d^-0.5.
 * Replaces pay_attention / sdpa_wrapper (utils/attention.py:99-116,161-398). */
int ltxb200_attention_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                           const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso,
                           int B, int H, int Lq, int Lk, int d, float scale, const float* key_bias,
                           void* stream);

/* y = norm(x) [* weight + bias] ; then y = y*(1+scale[g]) + shift[g], g = row / rows_per_group.
 * layer_norm = 0: RMSNorm (attention.py:233-251,314-320); 1: LayerNorm (transformer3d.py:494-502,
 * wan/modules/model.py:437-441,461,467-472).  D % 256 == 0, D <= 8192.  scale/shift/weight/bias may be NULL. */
int ltxb200_norm_mod_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int M, int D, const void* scale,
                          const void* shift, int64_t mod_ld, int rows_per_group, const void* weight,
                          const void* bias, float eps, int layer_norm, void* stream);

/* In-place q/k RMSNorm (affine, over the whole inner dim) + LTX interleaved RoPE with [tokens, D] bf16
 * cos/sin tables (attention.py:1040-1055, 960-975).  k or cos/sin may be NULL. */
int ltxb200_qk_norm_rope_bf16(void* q, int64_t ldq, int Mq, void* k, int64_t ldk, int Mk, int D,
                              const void* wq, const void* wk, const void* cos_table, const void* sin_table,
                              int tokens_per_batch, float eps, void* stream);

/* Wan variant: in-place q/k WanRMSNorm (wan/modules/model.py:99-111, two bf16 roundings) + per-head 3-axis RoPE
 * from fp32 [tokens, head_dim] tables, fp32 math (wan/modules/posemb_layers.py:222-276).
 * token = token_offset + row % tokens_per_batch: token_offset is the rank's first global token under Ulysses
 * sequence parallelism (wan/distributed/xdit_context_parallel.py:52-57). */
int ltxb200_qk_norm_rope_wan_bf16(void* q, int64_t ldq, int Mq, void* k, int64_t ldk, int Mk, int D, const void* wq,
                                  const void* wk, const float* cos_table, const float* sin_table, int head_dim,
                                  int tokens_per_batch, int token_offset, float eps, void* stream);

/* out = sum_j coefs[j] * xs[j] (fp32, terms <= 6, n % 4 == 0; xs and coefs are HOST arrays of device pointers /
 * scalars).  The UniPC predictor/corrector, x0 conversion and the CFG combine are such combinations
 * (wan/utils/fm_solvers_unipc.py:321,458-484,590-626; wan/text2video.py:562).  out may alias an input. */
int ltxb200_lincomb_f32(float* out, int64_t n, int terms, const float* const* xs, const float* coefs, void* stream);

/* RectifiedFlowScheduler.step with PER-TOKEN timesteps (ltx_video/schedulers/rf.py:361-375): token r (timestep tok_timesteps[r])
 * looks up the next entry of the descending `schedule` strictly below its timestep - 1e-6 (0 if none) and takes
 *   out = x - (t - lower) * v                              (noise == NULL), or
 *   out = (1 - next) * (x - t * v) + next * noise, next = t - (t - lower)   (stochastic sampling, rf.py:370-373; add_noise :382-392).
 * x, v, noise, out: [tokens, channels] fp32 (channels % 4 == 0); fp32 arithmetic with the reference expression's rounding
 * points (no contraction): equal to the PyTorch result bit for bit.  out may alias x.  (ABI version 4.) */
int ltxb200_rf_step_tokens_f32(float* out, const float* x, const float* v, const float* noise, const float* tok_timesteps,
                               int64_t tokens, int channels, const float* schedule, int num_steps, void* stream);

/* ada[l,g,j,:] = table[l,j,:] + temb[g, j*D:(j+1)*D]  (attention.py:239-241), JD = 6*D. */
int ltxb200_ada_add_bf16(const void* table, const void* temb, void* out, int L, int G, int JD, void* stream);

/* y = act(x) elementwise, n % 8 == 0.  mode: LTXB200_ACT_* */
int ltxb200_act_bf16(const void* x, void* y, int64_t n, int mode, void* stream);

/* STG AttentionValues blend: a[b] = a[b]*mask[b] + v[b]*(1-mask[b]) (attention.py:1134-1139). */
int ltxb200_stg_blend_bf16(void* a, const void* v, int64_t ldv, const float* mask, int B, int64_t rows, int D,
                           void* stream);

/* out = bf16(a*x + b*y), n % 8 == 0, out may alias x or y.  TeaCache residual reuse / capture
 * (wan/modules/model.py:1051-1054 `x += previous_residual[i]`, :1090-1099 `torch.sub(x, ori)`). */
int ltxb200_axpby_bf16(const void* x, const void* y, void* out, int64_t n, float a, float b, void* stream);

/* TeaCache step distance (wan/modules/model.py:1039 `(e - prev).abs().mean() / prev.abs().mean()`):
 * out2[0] = sum |bf16(a - b)|, out2[1] = sum |b| (fp32, device).  a, b: bf16 [n]. */
int ltxb200_rel_l1_bf16(const void* a, const void* b, int64_t n, float* out2, void* stream);

/* ---- LTX multi-scale flow (pipeline_ltx_video.py:1709-1903, latent_upsampler.py) ---- */
#define LTXB200_GROUPNORM_CHUNKS 64
/* y = [SiLU]( GroupNorm(32 groups, eps)(x) * gamma + beta [+ residual] ) on NDHWC bf16 x [B, voxels, C], C in {256,...,2048}
 * (latent_upsampler.py:30-39,73-75: conv -> norm -> SiLU, and SiLU(norm2(..) + residual)).  scratch: B*64*32*2 floats. */
int ltxb200_groupnorm_silu_bf16(const void* x, void* y, int B, int64_t voxels, int C, const void* gamma, const void* beta,
                                const void* residual, float eps, int apply_silu, float* scratch, void* stream);

/* adain_filter_latent (pipeline_ltx_video.py:1709-1737): rows = batch*channels; x [rows, n], ref [rows, m], out [rows, n] fp32. */
int ltxb200_adain_f32(const float* x, const float* ref, float* out, int rows, int64_t n, int64_t m, float factor, void* stream);

/* NDHWC bf16 -> NCDHW fp32, (x - mean[c]) / std[c] when std/mean are given (normalize_latents, vae_encode.py:228-237). */
int ltxb200_latent_from_ndhwc(const void* x, float* out, int B, int C, int64_t FHW, const float* stdv, const float* meanv,
                              void* stream);

/* F.interpolate(mode="bilinear", align_corners=False) per plane (pipeline_ltx_video.py:1894-1899): [planes,h,w] -> [planes,H,W]. */
int ltxb200_bilinear_resize_f32(const float* x, float* y, int64_t planes, int h, int w, int H, int W, void* stream);

/* diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): out[n, dim] bf16 from t[n] fp32 */
int ltxb200_timestep_embed(const float* t, void* out, int n, int dim, void* stream);

int ltxb200_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);

/* Guidance combine (cfg-star CFG, STG, std-rescale) + RectifiedFlow Euler step with per-token dt and
 * conditioning mask; latents fp32 updated in place (pipeline_ltx_video.py:1183-1222,1309-1342,
 * ltx_video/schedulers/rf.py:350-375).  pred: cond c at pred + c*cond_stride, n = tokens*channels.
 * scratch: >= 8*148 floats.  latents_bf16 (optional) receives the bf16 copy for the next forward. */
int ltxb200_guidance_step(const void* pred, int64_t cond_stride, int64_t n, int channels, int has_cfg,
                          int has_stg, int do_rescale, float guidance_scale, float stg_scale, float rescale,
                          float* latents, void* latents_bf16, const float* timesteps, int num_steps, float t,
                          const float* cond_mask, float* scratch, void* stream);

/* Same, with the stochastic update of rf.py:370-373: x0 = x - t*v, x <- (1 - t_next)*x0 + t_next*noise (noise: fp32 [n], N(0,1)). */
int ltxb200_guidance_step_stochastic(const void* pred, int64_t cond_stride, int64_t n, int channels, int has_cfg,
                          int has_stg, int do_rescale, float guidance_scale, float stg_scale, float rescale,
                          float* latents, void* latents_bf16, const float* timesteps, int num_steps, float t,
                          const float* cond_mask, float* scratch, const float* noise, void* stream);

/* Wan classifier-free guidance in fp32, optionally with the CFG-Zero* projection of the unconditional branch
 * (wan/text2video.py:31-42 optimized_scale, :551-562): out = a*u + g*(c - a*u), a = <c,u>/(|u|^2+1e-8) or 1.
 * scratch: >= 2*148 floats (only read when use_alpha). */
int ltxb200_cfg_combine_f32(const float* cond, const float* uncond, float* out, int64_t n, float guide_scale,
                            int use_alpha, float* scratch, void* stream);

/* PixelNorm over channels (+ optional SiLU) on NDHWC bf16 (pixel_norm.py:12; causal_video_autoencoder.py:1212,1240).
 * C in {64,128,256,512,1024}. */
int ltxb200_pixelnorm_silu_bf16(const void* x, void* y, int64_t voxels, int C, float eps, int apply_silu,
                                void* stream);

/* PixelNorm -> x * (1 + scale[c]) + shift[c] -> optional SiLU: the timestep-conditioned ResnetBlock3D / decoder tail
 * (causal_video_autoencoder.py:773-797,1212-1240); scale, shift: bf16 [C] (one video per call). */
int ltxb200_pixelnorm_mod_silu_bf16(const void* x, void* y, int64_t voxels, int C, float eps, const void* scale,
                                    const void* shift, int apply_silu, void* stream);

/* latents NCDHW (fp32 if is_f32 else bf16) -> x*std[c]+mean[c] -> NDHWC bf16 (vae_encode.py:239-247). */
int ltxb200_latent_to_ndhwc(const void* z, int is_f32, void* out, int B, int C, int64_t FHW, const float* stdv,
                            const float* meanv, void* stream);

/* ---- `mixed` precision (transformer3d.py:343,439-442; pipeline_ltx_video.py:1061,1152-1177; ltxv.py:186): the residual stream and the
 * AdaLN modulation stay fp32, only the Linear inputs are bf16 (the reference wraps the forward in torch.autocast). ---- */
/* out[M,N] (fp32, may alias residual) = act(A @ W^T + bias) * gate[m / rows_per_gate] + residual, gate / residual fp32 */
int ltxb200_gemm_bf16_f32res(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, float* out, int64_t ldc,
                             const void* bias, int act, const float* residual, int64_t ldr, const float* gate, int64_t gate_ld,
                             int rows_per_gate, void* stream);
/* y (bf16) = norm(x fp32) * (1 + scale[g]) + shift[g], scale / shift fp32 rows (NULL = plain norm), one rounding at the end */
int ltxb200_norm_mod_f32in(const float* x, int64_t ldx, void* y, int64_t ldy, int M, int D, const float* scale, const float* shift,
                           int64_t mod_ld, int rows_per_group, float eps, int layer_norm, void* stream);
/* ada32[l, g, :] = fp32(table[l, :]) + fp32(temb[g, :])  (table [L, JD], temb [G, JD] bf16 -> [L, G, JD] fp32) */
int ltxb200_ada_add_f32(const void* table, const void* temb, float* out, int L, int G, int JD, void* stream);

/* ltxb200_conv3d_bf16 (NDHWC store) with ResnetBlock3D's next PixelNorm + SiLU (causal_video_autoencoder.py:1212-1240, pixel_norm.py:12)
 * fused into the epilogue, for Cout <= 256 (one N tile holds the whole channel vector): out2[b,t,h,w,:] =
 * silu(bf16(y / sqrt(mean_c(y^2) + eps))) with y the row as stored (bf16, after bias and residual).
 * norm_mode 1: `out` (the raw row: the next block's residual) AND out2;  2: out2 only (`out` may be NULL). */
int ltxb200_conv3d_norm_bf16(const void* x, const void* w, const void* bias, void* out, void* out2, int B, int T, int H, int W,
                             int Cin, int Cout, int causal, const void* residual, int norm_mode, float eps, void* stream);
/* 3x3x3 CAUSAL convolution (two replicated leading frames, zero spatial padding) with output strides stride_t / stride_hw in
 * {1, 2}: x [B,T,H,W,Cin] -> out [B,(T-1)/stride_t+1,(H-1)/stride_hw+1,(W-1)/stride_hw+1,Cout].  The "compress_all/time/space"
 * blocks of the LTX VAE Encoder (causal_video_autoencoder.py:404-447 via make_conv_nd(stride=...)); the TMA descriptor strides
 * over the input, so no im2col or space-to-depth copy is made. */
int ltxb200_conv3d_strided_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W, int Cin,
                                int Cout, int stride_t, int stride_hw, void* stream);

/* ---- Wan2.1 VAE decode (wan/modules/vae.py:386-493): the pieces the LTX decoder kernels do not already cover ---- */
/* convolution with taps_t x taps_hw x taps_hw taps (each 1 or 3) on NDHWC bf16, w = [Cout, taps*Cin] tap-major, causal in time
 * (taps t-2,t-1,t) with ZERO temporal padding when causal_zero_pad == 1 (Wan CausalConv3d, vae.py:17-37; Conv2d 3x3 of
 * Resample with taps_t = 1, :80-88; time_conv with taps_hw = 1, :86-88), replicate padding when 0; causal_zero_pad == 2 is
 * the centred, zero-padded nn.Conv3d / nn.Conv2d(kernel 3, padding 1) of the LatentUpsampler (latent_upsampler.py:25-27,
 * 73,90-93,107); optional residual. */
int ltxb200_conv_taps_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W, int Cin,
                           int Cout, int taps_t, int taps_hw, int causal_zero_pad, const void* residual, void* stream);
/* ---- Wan2.1 VAE encode (wan/modules/vae.py:275-383, 536-575): the strided pieces of Resample('downsample2d'|'downsample3d') ----
 * zero-padded causal convolution as above with output strides stride_t / stride_hw in {1, 2}:
 *   out [B, (T-1)/stride_t+1, (H-1-off_hw)/stride_hw+1, (W-1-off_hw)/stride_hw+1, Cout];
 * off_hw = 1 moves the spatial taps to (h, h+1, h+2): nn.ZeroPad2d((0,1,0,1)) + Conv2d(dim, dim, 3, stride 2) (:90-93, taps_t = 1);
 * taps_t = 3, taps_hw = 1, stride_t = 2 is `time_conv` = CausalConv3d(dim, dim, (3,1,1), stride (2,1,1)) applied to
 * [cached last frame, chunk] (:150-165): output frame m >= 1 reads input frames (2m-2, 2m-1, 2m); frame 0 is replaced by the
 * caller (the first chunk bypasses time_conv). */
int ltxb200_conv_taps_strided_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W, int Cin,
                                   int Cout, int taps_t, int taps_hw, int stride_t, int stride_hw, int off_hw, void* stream);
/* RMS_norm (vae.py:41-58: F.normalize over channels * sqrt(c_real) * gamma) + optional SiLU on [voxels, C] bf16 */
int ltxb200_l2norm_silu_bf16(const void* x, void* y, int64_t voxels, int C, int c_real, const void* gamma, int apply_silu,
                             void* stream);
/* nearest(-exact) x2 upsample of NHWC frames (vae.py:61-67,80-82): [frames, H, W, C] -> [frames, 2H, 2W, C] */
int ltxb200_upsample2x_nhwc_bf16(const void* x, void* y, int64_t frames, int H, int W, int C, void* stream);
/* P = softmax(scale * S) row-wise, S fp32 [rows, ld_s] -> P bf16 [rows, ld_p] (AttentionBlock, vae.py:257-263: one head of
 * width C = 384, computed as two tcgen05 GEMMs around this kernel) */
int ltxb200_softmax_rows_f32_bf16(const float* s, int64_t ld_s, void* p, int64_t ld_p, int rows, int cols, float scale,
                                  void* stream);

/* ltxb200_attention_bf16 with per-batch key lengths instead of an additive bias: batch element b attends to keys [0, key_lens[b]) only
 * (device int32 [B], 1 <= key_lens[b] <= Lk).  For a right-padded prompt this is the reference's (1 - mask) * -10000 bias
 * (transformer3d.py:411-415; utils/attention.py:179-180 forces sdpa for masks) without the bias pass and without the padded key blocks:
 * exp(-10000 + s - m) is exactly 0 in fp32, so the result is the same. */
int ltxb200_attention_klens_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk, const void* v,
                                 int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso, int B, int H, int Lq, int Lk, int d,
                                 float scale, const int* key_lens, void* stream);

/* as ltxb200_attention_bf16, but out += attention(q, k, v) (bf16 read-modify-write in the epilogue): the image-token
 * branch of WanI2VCrossAttention, `x += img_x` (wan/modules/model.py:329-337), without a separate add pass. */
int ltxb200_attention_acc_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                               const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso, int B, int H,
                               int Lq, int Lk, int d, float scale, const float* key_bias, void* stream);

/* ---- peer-memory exchange for Ulysses sequence parallelism (one process per GPU, NVLink P2P) ----
 * Replaces xFuserLongContextAttention's head<->sequence all-to-alls as called at
 * wan/distributed/xdit_context_parallel.py:179-184: the producing kernels store straight into the destination rank's
 * buffer; see csrc/comm.cuh for the flag protocol.  comm_alloc/open/close/free are the only entry points of the library
 * that allocate or synchronise (setup time, not the data path). */
int ltxb200_comm_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64B);   /* cudaMalloc (zeroed) + cudaIpcGetMemHandle */
int ltxb200_comm_open(const void* ipc_handle_64B, void** dev_ptr);            /* map a peer's buffer (enables P2P lazily) */
int ltxb200_comm_close(void* dev_ptr);
int ltxb200_comm_free(void* dev_ptr);
/* stream-ordered wait until flags[0..P) (this rank's flag array) have all reached `epoch`.  The wait is bounded
 * (LTXB200_COMM_TIMEOUT_MS, default 20 s): a peer that never publishes costs the timeout, not a hung GPU. */
int ltxb200_comm_wait(const void* flags, int P, unsigned int epoch, void* stream);
/* the same with an error path: on timeout (timeout_ms, 0 = the default above) the kernel stores (source rank + 1) | epoch << 8
 * into *status_dev (device word; sticky — later waits given the same word return immediately) and *status_host (a word of
 * pinned, device-accessible host memory the host polls without synchronising), then returns.  Either may be NULL. */
int ltxb200_comm_wait_status(const void* flags, int P, unsigned int epoch, void* status_dev, void* status_host,
                             unsigned int timeout_ms, void* stream);
/* all-gather by peer stores (the head output gather, xdit_context_parallel.py:142 `get_sp_group().all_gather(x, dim=1)`):
 * segment b (of nseg, seg_bytes each, 16-byte multiple) of the contiguous local `src` is stored at byte offset
 * (b*P + rank)*seg_bytes of dst_ptrs[r] on every rank r, then flag[rank] = epoch is published on every peer. */
int ltxb200_peer_allgather(const void* src, int64_t seg_bytes, int nseg, int P, int rank, void* const* dst_ptrs,
                           void* const* flag_ptrs, unsigned int epoch, void* counter, void* stream);
/* Wan q/k RMSNorm + RoPE (as ltxb200_qk_norm_rope_wan_bf16) on the local fused QKV rows [B*n_loc, 3*D], v passed through,
 * each head group stored into recv_ptrs[g] laid out [N, B, 3, H/P, head_dim] in global token order; then flag[rank] =
 * epoch is published on every peer.  recv_ptrs / flag_ptrs: HOST arrays of P device pointers (peer mappings). */
int ltxb200_qk_norm_rope_wan_scatter_bf16(const void* qkv, int64_t ld, int M, int D, const void* wq, const void* wk,
                                          const float* cos_table, const float* sin_table, int head_dim,
                                          int tokens_per_batch, int token_offset, float eps, int B, int P, int rank,
                                          void* const* recv_ptrs, void* const* flag_ptrs, unsigned int epoch,
                                          void* counter, void* stream);
/* the same over rows [row0, row0 + rows) of the M local rows only: one exchange may be issued as several launches (token
 * chunks, so that the peer stores of chunk i overlap the QKV projection of chunk i+1 on another stream).  `signal_ctas` =
 * the sum of ltxb200_scatter_signal_ctas(rows_c) * nsel / 3 over all launches of the exchange; the last CTA of the last launch
 * to finish publishes the flag.  nsel = 3: q, k and v;  nsel = 2: q and k only. */
int ltxb200_qk_norm_rope_wan_scatter_rows_bf16(const void* qkv, int64_t ld, int M, int row0, int rows, int D, const void* wq,
                                               const void* wk, const float* cos_table, const float* sin_table, int head_dim,
                                               int tokens_per_batch, int token_offset, float eps, int B, int P, int rank,
                                               void* const* recv_ptrs, void* const* flag_ptrs, unsigned int epoch,
                                               void* counter, unsigned int signal_ctas, int nsel, void* stream);
/* CTAs one launch over `rows` rows adds to the arrival counter with nsel = 3 (q, k, v); with nsel = 2 (V already sent by
 * ltxb200_gemm_qkv_vscatter_bf16) it is two thirds of that */
unsigned int ltxb200_scatter_signal_ctas(int rows);
/* The fused QKV projection out[M, 3D] = A @ W^T + bias of the LOCAL token shard whose V third (columns 2D..3D, no normalisation needed) is
 * stored by the GEMM epilogue straight into recv_ptrs[g] ([N, B, 3, H/P, head_dim], global token order) of the rank that owns the heads,
 * while q and k (columns 0..2D) stay local for the scatter kernel above (nsel = 2), which also publishes the exchange's flag:
 * a third of the head<->sequence all-to-all (xdit_context_parallel.py:179-184) leaves the GPU during the GEMM.  row0 = first local row of A. */
int ltxb200_gemm_qkv_vscatter_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, int D, void* out, int64_t ldc,
                                   const void* bias, int head_dim, int tokens_per_batch, int token_offset, int row0, int B, int P,
                                   int rank, void* const* recv_ptrs, void* stream);
/* ltxb200_attention_bf16 whose epilogue stores query token t's row to out_ptrs[t / tokens_per_peer] at row
 * b*tokens_per_peer + t % tokens_per_peer, head head_offset + h of a [B*tokens_per_peer, ldo] matrix (the Ulysses
 * return exchange, xdit_context_parallel.py:186-190), then publishes the epoch flag on every peer. */
int ltxb200_attention_scatter_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                                   const void* v, int64_t ldv, int64_t bsv, int64_t ldo, int B, int H, int Lq, int Lk,
                                   int d, float scale, const float* key_bias, int P, int rank, void* const* out_ptrs,
                                   void* const* flag_ptrs, unsigned int epoch, void* counter, int tokens_per_peer,
                                   int head_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTX_B200_H_ */
