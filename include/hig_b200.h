/* hig_b200 — C ABI of the B200-native denoising hot path of line/Human-Interaction-Generation.
 *
 * The reference has no FFI: its hot path is PyTorch eager code.  Each entry point below names the reference
 * Python it replaces (paths relative to the reference's codes/ directory).  All pointers are DEVICE pointers
 * owned by the caller (PyTorch), nothing is allocated inside, every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*), and every call returns 0 or a negative HIG_ERR_* code
 * (message via hig_last_error()).  There is no CPU fallback: without a CUDA device the calls fail.
 *
 * dtype arguments: HIG_BF16 (0) = __nv_bfloat16 storage, HIG_F32 (1) = float storage ("fp32 mode").
 */
#ifndef HIG_B200_H
#define HIG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HIG_OK 0
#define HIG_ERR_INVALID (-1)
#define HIG_ERR_CUDA (-2)
#define HIG_ERR_UNSUPPORTED (-3)
#define HIG_ERR_NO_DRIVER (-4)

#define HIG_BF16 0
#define HIG_F32 1
#define HIG_F16 2 /* residual-stream storage of the product path (11-bit mantissa; saturating stores) */

#define HIG_ACT_NONE 0
#define HIG_ACT_GELU 1 /* exact erf GELU, nn.GELU() */
#define HIG_ACT_SILU 2
#define HIG_ACT_QUICKGELU 3 /* x sigmoid(1.702 x): CLIP text transformer MLP (hig_act_fwd only) */

#define HIG_ATTN_SELF 0    /* LinearTemporalSelfAttention              models/interaction_transformer.py:112-130 */
#define HIG_ATTN_INTER 1   /* LinearTemporalInteractionCrossAttention  models/interaction_transformer.py:181-207 */
#define HIG_ATTN_KV_ONLY 2 /* LinearTemporalCrossAttention, K/V half   models/interaction_transformer.py:155-161 */
#define HIG_ATTN_Q_ONLY 3  /* LinearTemporalCrossAttention, Q half     models/interaction_transformer.py:153,156,162 */

/* library info */
int hig_version(void);
const char* hig_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long hig_launch_count(void);

/* Debug aid (no reference counterpart): when buf != NULL the resident-W projection kernel writes 32 clock64 / globaltimer
 * slots per CTA pair into buf, each launch taking the next block of 74 * 32 entries; tracing stops by itself after
 * max_launches launches (buf holds max_launches * 74 * 32 entries) — tile-by-tile timeline read by tools/gemm_trace.py;
 * NULL or max_launches <= 0 disables. */
int hig_debug_trace(unsigned long long* buf, int max_launches);

/* Grid sizing (no reference counterpart): the persistent kernels (tcgen05 GEMMs, attention apply) launch one CTA (pair) per SM.
 * n > 0 makes them size their grids for n SMs instead of all of them — data-parallel training leaves a few SMs to the NCCL
 * all-reduce kernels (tools/train.py:78-82 is where the reference's DDP reducer would run) so that a gradient segment's
 * exchange overlaps the next segment's backward; n <= 0 restores the device's SM count. */
int hig_set_sm_limit(int n);

/* Debug aid (no reference counterpart): while `counter` (a device unsigned long long) is non-NULL, every fp16 residual-stream
 * store of hig_gemm_stream (HIG_GS_RES_H) whose fp32 value exceeds +-65504 — i.e. that the saturating conversion clamps —
 * adds 1 to *counter.  The reference's stream is fp32 (models/interaction_transformer.py:129,164,203,263): a non-zero count
 * means this checkpoint does not fit the fp16 stream and `precision="fp32"` must be used.  NULL disables (default). */
int hig_debug_saturation(unsigned long long* counter);

/* L2 residency hint (no reference counterpart): pins [ptr, ptr+bytes) — the fp32 residual stream — in the 126 MB L2
 * through an access-policy window on `stream`; ptr == NULL clears it. */
int hig_l2_persist(const void* ptr, unsigned long long bytes, float hit_ratio, void* stream);

/* out = act(A[M,K] . W[N,K]^T + bias + residual), bf16 operands, fp32 accumulate on tcgen05/TMEM.
 * Replaces every nn.Linear on the path (models/interaction_transformer.py:74-97,105-128,137-163,172-205,
 * 254-263,471-478,508-509) with the bias add, GELU (FFN, :262), SiLU (time_embed, :476) and the residual
 * `x + proj_out(...)` (:129,164,203,263) fused in the epilogue.
 * lda/ldw/K multiples of 8; residual is fp32 [*, ldr], row = m % res_row_mod when res_row_mod > 0;
 * either/both of out_f32 / out_bf16 may be given. */
int hig_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                  const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, void* out_bf16,
                  int ldo_bf16, int act, void* stream);

/* hig_gemm_bf16 with typed residual / main output: each HIG_F32 or HIG_F16.  The product path keeps the residual
 * stream `x = x + proj_out(...)` (:129,164,203,263) in fp16 — half the epilogue traffic of fp32 at 8x the
 * precision of bf16 (measured: fp16 storage costs 1e-3 relative error per denoiser step, bf16 8e-3). */
int hig_gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                     const void* residual, int res_dtype, int ldr, int res_row_mod, void* out, int out_dtype, int ldo,
                     void* out_bf16, int ldo_bf16, int act, void* stream);

/* The token-sized projections of the sampling path with a TMA-staged epilogue (csrc/gemm_stream.cu): same CTA-pair
 * tcgen05 main loop, but the epilogue packs each accumulator row into a 128B-swizzled shared-memory slab and one lane
 * issues a TMA store (residual tiles arrive by TMA load).  A [M,K], W [N,K] both HIG_BF16 or both HIG_F16 (op_dtype),
 * N % 64 == 0, bias fp32 [N].
 *   HIG_GS_BF16      out bf16 [M,ldo] = A.W^T + bias                       (FFN linear2 :263, text-CA query :153)
 *   HIG_GS_BF16_GELU out bf16 = GELU(A.W^T + bias)                         (FFN linear1 :262)
 *   HIG_GS_RES_H     out fp16 [M,ldo] += A.W^T + bias, in place: the residual `x + proj_out(...)` (:97,129,164,203,
 *                    263); stats_out (nullable, N == 512) fp32 [M,8] receives four (sum, sum of squares) partials of
 *                    every updated row
 *   HIG_GS_LN_BF16   out bf16 = rstd_m (A.W^T - mu_m wsum) + bias: nn.LayerNorm(ln_width) folded into the projection
 *                    that consumes it (:119-121,153,190-194).  A is the raw stream, W = gamma o W_lin, wsum[n] =
 *                    sum_k W[n,k], bias[n] = b[n] + sum_k beta[k] W_lin[n,k]; (mu, rstd) of row m from stats_in[m]
 *                    (the partials HIG_GS_RES_H or hig_row_stats left), eps = 1e-5. */
#define HIG_GS_BF16 0
#define HIG_GS_BF16_GELU 1
#define HIG_GS_RES_H 2
#define HIG_GS_LN_BF16 3
#define HIG_GS_F16 4 /* out fp16 = A.W^T + bias (K <= 512, N % 256 == 0): the output heads out / out2 (:613-616) */
/* HIG_GS_LN_BF16 whose first 512 output columns (the query block of a Q or Q|K|V projection) are replaced by the softmax
 * over each head's 64 features — F.softmax(query, dim=-1) of :120,156,195 — computed lane-locally in the epilogue
 * (K <= 512, N % 256 == 0).  hig_attn_apply_stylize then takes q with flag bit 1 set. */
#define HIG_GS_LN_QSM 5
int hig_gemm_stream(int kind, const void* A, int lda, const void* W, int ldw, int op_dtype, int M, int N, int K,
                    const float* bias, const float* wsum, const float* stats_in, float* stats_out, int ln_width,
                    void* out, int ldo, void* stream);

/* stats[row] = {sum, sum of squares, 0 x 6} of fp16 rows of width 512: the LayerNorm statistics of the motion-embedding
 * output (models/interaction_transformer.py:593-602 feeding :119) in the layout HIG_GS_LN_BF16 reads. */
int hig_row_stats(const void* x, int x_dtype, int rows, int width, float* stats, void* stream);

/* the same contract in fp32 storage and fp32 FFMA arithmetic ("fp32 mode", parity <= 1e-5) */
int hig_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias,
                 const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, int act, void* stream);

/* out = [SiLU]( LayerNorm(x; gamma, beta, eps=1e-5) * (1 + scale[s]) + shift[s] ), s = row / rows_per_seq.
 * Replaces StylizationBlock.forward's norm/FiLM/SiLU (models/interaction_transformer.py:86-97) and the
 * pre-attention LayerNorms (:119,153,155,190,194).  scale_shift (nullable) points at fp32
 * [S, ss_stride] with scale at [0,width) and shift at [width, 2*width) of each row.  width = 512 or 256. */
int hig_ln_film_silu(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int apply_silu, void* out,
                     int out_dtype, void* stream);

/* fused efficient attention, one CTA per (sequence, head), head dim 64; see HIG_ATTN_* for the reference lines.
 * q [S*T, ldq], k/v [S*T, ldkv] (pointers at head 0), y [S*T, ldy]; a_in/a_out [S,H,64,64] (dtype storage);
 * length int32 [S] (nullable = no mask); INTER reads K,V of sequence (s + pair_shift) % S and masks with
 * length[s] (the reference's query-side quirk, :194). */
int hig_eff_attn(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                 void* a_out, void* y, int ldy, const int* length, int S, int T, int H, int pair_shift, int mask_v,
                 int dtype, void* stream);

/* Query half of the efficient attention fused with the block's StylizationBlock front end, bf16 storage:
 *   out[t,:] = SiLU( LN_512( concat_h softmax_feat(Q[t,h]) . A[s,h] ) * (1 + scale_s) + shift_s )
 * (einsum + reshape at models/interaction_transformer.py:128,162,201, then StylizationBlock.forward :86-97 up to its
 * SiLU).  q [S*T, ldq] at head 0, a_in [S,8,64,64] from hig_eff_attn KV_ONLY (which honours `length` and
 * `pair_shift`: the K/V half of the self / inter-person attention), scale_shift as in hig_ln_film_silu, out [S*T,512].
 * apply_silu is a flag word: bit 0 = SiLU after the affine; bit 1 = q already holds softmax_feat(Q) (the projection
 * that produced it ran with HIG_GS_LN_QSM), so the kernel feeds it to the contraction as it is. */
int hig_attn_apply_stylize(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                           const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                           void* stream);

/* out[s] = [cos(t_s f) | sin(t_s f)] — timestep_embedding (models/interaction_transformer.py:26-43);
 * freqs = fp32 [half] table computed once on the host exactly as the reference does. */
int hig_timestep_embed(const long long* t, const float* freqs, int S, int half, void* out, int out_dtype,
                       void* stream);

/* x fp32 [S,T,C] -> operand [S*T, ld_out] laid out so one GEMM with [joint_embed.weight | joint_embed2.weight]
 * reproduces embed_motion (models/interaction_transformer.py:593-602). */
int hig_pack_motion(const float* x, int S, int T, int C, int ld_out, void* out, int out_dtype, void* stream);

/* one reverse-diffusion update, in place on x: p_sample + p_mean_variance (EPSILON / FIXED_SMALL,
 * clip_denoised=False) — models/gaussian_diffusion.py:606-666,443-537.  coef = fp32 [5][n_steps]
 * {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
 * exp(0.5*posterior_log_variance_clipped)}.  noise nullable -> Philox4x32-10(seed, t, index) + Box-Muller
 * (replaces th.randn_like, :657); seed_dev nullable: when given, the key is read from device memory instead of `seed`,
 * so a CUDA graph holding this launch can be replayed for a new sample after an 8-byte update.
 * eps [S*T, ld_eps] is HIG_F32 or HIG_F16 (eps_dtype; the product path's output heads write fp16).
 * packed nullable: also writes the next step's pack_motion operand.  t_next nullable: t_next[s] = t[s]-1
 * (may alias t; issued as a trailing launch). */
int hig_ddpm_step(float* x, const void* eps, int ld_eps, int eps_dtype, const float* noise, const long long* t, const float* coef,
                  int n_steps, int S, int T, int C, unsigned long long seed, const unsigned long long* seed_dev, void* packed,
                  int ld_packed, int packed_dtype, long long* t_next, void* stream);

/* out[s,:] = SiLU(table[t[s],:] + xf_proj[s,:]) — the input of every StylizationBlock's emb linear (SiLU folded in,
 * models/interaction_transformer.py:74-77) when emb = time_embed(timestep_embedding(t)) + xf_proj (:591) is read from a
 * table of the time MLP over the whole schedule (table fp32 [n_steps,E], built once per set of weights with hig_gemm_bf16)
 * instead of being recomputed every sampling step.  t int64 [S] (clamped to the table), xf_proj fp32 [S,E],
 * out HIG_BF16 or HIG_F32 [S,E]. */
int hig_time_table_silu(const float* table, int n_steps, const long long* t, const float* xf_proj, int S, int E, void* out,
                        int out_dtype, void* stream);

/* out[r,:] = fp16(table[r % period,:]), table fp32 [period,width], out fp16 [rows,width], width % 8 == 0: seeds the fp16
 * residual stream with sequence_embedding + the joint_embed biases (models/interaction_transformer.py:593-602), after
 * which embed_motion is the in-place projection hig_gemm_stream(HIG_GS_RES_H) on the packed motion operand. */
int hig_tile_rows(const float* table, int period, int width, long long rows, void* out_f16, void* stream);

/* Sampled motion -> 3-D joints in one launch: tools/visualization.py:149-155 (x[1:]*std+mean, x[0,:4]*init_std+init_mean)
 * followed by utils/motion_process.py recover_from_ric2 (:418-456; recover_root_rot_pos :362-381, qrot/qinv
 * utils/quaternion.py:16-20,54-73).  x fp32 [S,T,C] contiguous, persons stacked on dim 0 (each sequence is processed on
 * its own, as the reference does); init_row = 0 (the sampler's layout: init state in row 0) or T-1 (recover_from_ric2's
 * layout: init state moved to the end); mean/std [C] and init_mean/init_std [4] nullable pairs (NULL = already
 * de-normalised); length int32 [S] nullable: rows >= length[s] are padding, their joints are written as zeros.
 * joints fp32 [S, T-1, joints_num, 3].  The two prefix sums accumulate in fp64 like torch's CPU cumsum. */
int hig_recover_joints(const float* x, int S, int T, int C, int init_row, const float* mean, const float* std_,
                       const float* init_mean, const float* init_std, const int* length, int joints_num, float* joints,
                       void* stream);

/* x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) noise — GaussianDiffusion.q_sample (models/gaussian_diffusion.py:399-417) */
int hig_q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                 const float* sqrt_1mac, int S, int TC, float* out, void* stream);

/* hig_attn_apply_stylize that also writes y_out (bf16 [S*T, 512], 16-byte aligned) = concat_h softmax_feat(Q_h) . A[s,h], the
 * attention output before the StylizationBlock (models/interaction_transformer.py:128,162,201): the training forward keeps it
 * for the LayerNorm backward, so the unfused attention kernel + a separate LayerNorm pass are not needed there either. */
int hig_attn_apply_stylize_y(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                             const float* scale_shift, int ss_stride, int apply_silu, void* out, void* y_out, int S, int T,
                             int H, void* stream);

/* K/V half of the efficient attention, bf16: a_out[s,h] = softmax_time(K_masked)^T (V mask) (64 x 64 per head;
 * models/interaction_transformer.py:121-127 / :194-200), K/V of sequence (s + pair_shift) % S, rows >= length[s] masked.
 * transposed != 0 writes A^T ([l][d]), the K-major B operand hig_attn_apply_stylize_tc consumes. */
int hig_attn_kv(const void* k, const void* v, int ldkv, void* a_out, const int* length, int S, int T, int H,
                int pair_shift, int transposed, void* stream);

/* hig_attn_apply_stylize on tcgen05 / TMEM (the product path's kernel): q holds softmax_feat(Q) already (written by a
 * HIG_GS_LN_QSM projection), a_t = A^T [S, 8, 64, 64] from hig_attn_kv(transposed = 1); one 128-row tile of a sequence per
 * CTA step, 8 heads x tcgen05.mma (M=128, N=64, K=64) into the SM's 512 TMEM columns, one row per TMEM lane so the
 * LayerNorm of models/interaction_transformer.py:93 is lane-local; FiLM + SiLU (:94-97) in registers, TMA store.
 * apply_silu: bit 0 only.  T <= 256. */
int hig_attn_apply_stylize_tc(const void* q, int ldq, const void* a_t, const float* gamma, const float* beta,
                              const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training path.  The reference has no backward code: torch.autograd differentiates the forward above
 * (loss.backward() in DDPMMulTrainer.update, trainers/mul_ddpm_trainer.py:249-256).  These are the hand-written
 * derivatives of the forward kernels; hig_b200/autograd.py schedules them.
 * ------------------------------------------------------------------------------------------------------------ */

/* out_f32[M,N] += A[M,K] . W[N,K]^T with the K range split over CTAs and combined by fp32 atomics (the caller
 * initialises out_f32).  Weight gradients dW = dY^T . X: M, N are weight dims, K = tokens.  k_splits <= 0: auto. */
int hig_gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32,
                         int ldo_f32, int k_splits, void* stream);

/* C[M,N] (+)= opA(A) . opB(W)^T on the CTA-pair tcgen05 kernel with either operand consumed MN-major (as it lies in
 * memory, no transposed copy): trans_a: A stored [K, M]; trans_b: W stored [K, N].  What torch.autograd computes for
 * nn.Linear (every nn.Linear call site of models/interaction_transformer.py, see hig_gemm_bf16):
 *   grad_weight = dY^T . X   -> trans_a = trans_b = 1, A = dY [tok, out], W = X [tok, in], K = tok, split_k = -1
 *   grad_input  = dY . W     -> trans_b = 1, A = dY [tok, out], W = weight [out, in] as stored, K = out
 * bias (fp32 [N]) / residual (fp32 [M, ldr], may alias out_f32 for accumulation) / out_bf16 as in hig_gemm_bf16;
 * split_k != 0 accumulates with fp32 atomics into out_f32 (caller-initialised), raw products only. */
int hig_gemm_bf16_t(int trans_a, int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const float* bias, const float* residual, int ldr, float* out_f32, int ldo_f32, void* out_bf16,
                    int ldo_bf16, int split_k, void* stream);

/* An nn.Linear next to an activation, fused for the training path (the FFN of models/interaction_transformer.py:261-264,
 * linear1 -> GELU -> linear2, and what torch.autograd does on the way back):
 *   forward  (trans_b = 0, W [N,K]):  pre = A W^T + bias;  out_pre_bf16 (nullable) = pre;  out_bf16 = act(pre as stored)
 *   backward (trans_b = 1, W [K,N] = the next layer's weight as stored):  out_bf16 = (A W + bias) * act'(gate_bf16[M,N]),
 *            gate = the saved pre-activation (nullable: no gating), gate_act 1 GELU / 2 SiLU
 * act: 0 none, 1 GELU (erf form, as hig_act_fwd), 2 SiLU.  bias fp32 [N] is required (pass zeros).  N % 32 == 0. */
int hig_gemm_bf16_fused(int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                        int act, void* out_bf16, int ldo_bf16, void* out_pre_bf16, int ldo_pre, const void* gate_bf16,
                        int ld_gate, int gate_act, void* stream);

/* in [M,N] -> outT [N,M] (nullable), copy [M,N] in the output dtype (nullable), colsum[n] += sum_m in[m,n] (nullable,
 * fp32 atomics: bias gradients).  rows_zero_mod > 0 treats rows with m % rows_zero_mod == 0 as zero (frame 0 of a
 * sequence belongs to the out2 head, models/interaction_transformer.py:613-616). */
int hig_transpose(const void* in, int in_dtype, int M, int N, int ld_in, void* outT, int ld_t, void* copy, int ld_c,
                  int out_dtype, float* colsum, int rows_zero_mod, void* stream);

/* out[n] += sum_m in[m,n] (fp32 atomics; the caller zeroes out) */
int hig_colsum(const void* in, int dtype, int M, int N, int ld, float* out, void* stream);

/* out = act(x) and dx = dy * act'(x) for HIG_ACT_* (exact erf GELU of FFN :257/:262, SiLU of time_embed :476) */
int hig_act_fwd(const void* x, int x_dtype, long long n, int act, void* out, int out_dtype, void* stream);
int hig_act_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, long long n, int act, void* dx, int dx_dtype,
                void* stream);

/* backward of hig_ln_film_silu: dx (overwritten, or += when dx_accumulate), d_ss[s] += (dscale | dshift) (nullable),
 * d_gb[s] += per-sequence partials of (dgamma | dbeta) (nullable; column-sum over s gives the parameter gradients) */
int hig_ln_film_silu_bwd(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                         const float* beta, const float* scale_shift, int ss_stride, int apply_silu, const void* dout,
                         int dout_dtype, void* dx, int dx_dtype, int dx_accumulate, float* d_ss, int dss_stride,
                         float* d_gb, int dgb_stride, void* stream);

/* backward of hig_eff_attn (same modes).  SELF/INTER: q,k,v,dy -> dq,dk,dv (INTER writes dk,dv to the partner's
 * rows).  KV_ONLY: k,v,dA(in, fp32 [S,H,64,64]) -> dk,dv.  Q_ONLY: q,a_in,dy -> dq, dA(out). */
int hig_eff_attn_bwd(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                     const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                     const int* length, int S, int T, int H, int pair_shift, int dtype, void* stream);

/* The same, plus the bias gradients of the projections that produced Q / K / V (torch.autograd's sum over the token rows of
 * the nn.Linear outputs' gradients, models/interaction_transformer.py:108-110): q_sum / k_sum / v_sum [H*64] fp32 (each
 * nullable) += column sums of dq / dk / dv as stored.  The bf16 tensor-core kernel takes them from the tiles it is about to
 * store; the other kernels (and HIG_DETERMINISTIC=1) run hig_colsum afterwards. */
int hig_eff_attn_bwd_sums(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                          const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                          const int* length, int S, int T, int H, int pair_shift, int dtype, float* q_sum, float* k_sum,
                          float* v_sum, void* stream);

/* Softmax multi-head attention of the text transformers in MotionInteractionTransformer.encode_text
 * (models/interaction_transformer.py:533-559: CLIP's causal 8-head text transformer, and the 4-head nn.TransformerEncoder of
 * :446-455): out[b,i,h,:] = softmax_j(q_i . k_j / 8 (+ causal mask)) v_j, head dim 64, rows [B*N, ld] with q/k/v pointing at
 * head 0 of their blocks, N <= 128 tokens, storage dtype bf16 or fp32 (fp32 arithmetic). */
int hig_mha_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int B, int N, int H,
                      int causal, int dtype, void* stream);

/* The reference's training loss and its gradient — DDPMMulTrainer.backward_G, trainers/mul_ddpm_trainer.py:223-247:
 * per frame the mean over features of (pred - tgt)^2 (frame 0: its first 4 features only; frames >= 1: all C), weighted by
 * the length mask, normalised by the mask sum.  pit != 0 (unlabelled mode, S = 4B sequences stacked (m1,c1) (m1,c2) (m2,c2)
 * (m2,c1)): the two persons of an assignment are summed and the cheaper assignment of each pair is kept (:235-242).
 * rows [S] and w [S] are fp32 scratch (per-sequence loss sums / gradient weights), loss a device scalar,
 * d_pred (nullable) = d loss / d pred, fp32 [S,T,C]. */
int hig_masked_mse(const float* pred, const float* tgt, const int* length, int S, int T, int C, int pit, float* rows,
                   float* w, float* loss, float* d_pred, void* stream);

/* *out (double, caller-zeroed) += sum_i x[i]^2 — the global gradient norm of clip_grad_norm_ (mul_ddpm_trainer.py:253) */
int hig_sumsq(const float* x, long long n, double* out, void* stream);

/* own[i] = (own[i] + sum_{j < count} staged[j * stride + i]) * scale, fp32, n elements: the reduction step of the data-parallel
 * gradient mean (the reference's DDP all-reduce, tools/train.py:78-82) when the ranks exchange their gradient slices through
 * NVLink peer memory with the copy engines (hig_b200.ddp.PeerGradExchange) instead of an SM-resident NCCL kernel. */
int hig_mean_slices(float* own, const float* staged, long long n, long long stride, int count, float scale, void* stream);

/* clip-by-global-norm + torch.optim.Adam step (mul_ddpm_trainer.py:253-255, :291) over flat fp32 buffers in one pass;
 * p_bf16 (nullable) receives the bf16 mirror of the updated parameters (the tcgen05 GEMM operands).  gnorm2 (nullable):
 * device double holding the squared global gradient norm; the gradient is scaled by min(1, max_norm / (norm + 1e-6)). */
int hig_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                  float beta2, float eps, int step, const double* gnorm2, float max_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIG_B200_H */
