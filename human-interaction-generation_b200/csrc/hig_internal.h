// Internal (C++) interfaces between the kernel translation units and the C ABI in capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/hig_b200.h"  // HIG_OK / HIG_ERR_* / HIG_BF16 / HIG_F32

#include <utility>

namespace hig {

// HIG_PDL=0 disables programmatic dependent launch (default on)
bool pdl_enabled();
// HIG_DETERMINISTIC=1: cross-CTA floating-point atomics are replaced by single-contributor reductions (bit-reproducible gradients)
bool deterministic();

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute: consecutive
// kernels of the denoiser step overlap prologue and tail (also inside CUDA-graph capture: programmatic edges)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// records a message retrievable through hig_last_error() and returns `code`
int set_error(int code, const std::string& msg);
// every kernel launch made by this library bumps a process-wide counter (hig_launch_count)
void count_launch();

int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
              const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, void* out_bf16,
              int ldo_bf16, int act, cudaStream_t stream);

int gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                 const void* residual, int res_dtype, int ldr, int res_row_mod, void* out, int out_dtype, int ldo,
                 void* out_bf16, int ldo_bf16, int act, cudaStream_t stream);

// kind: 0 bias->bf16, 1 bias->GELU->bf16, 2 fp16 stream in place (+ row statistics), 3 LayerNorm-folded -> bf16
int gemm_stream(int kind, const void* A, int lda, const void* W, int ldw, int op_dtype, int M, int N, int K,
                const float* bias, const float* wsum, const float* stats_in, float* stats_out, int ln_width,
                void* out, int ldo, cudaStream_t stream);

void set_gemm_trace(unsigned long long* buf, int max_launches);
void set_saturation_counter(unsigned long long* counter);
void set_sm_limit(int n);

int row_stats(const void* x, int x_dtype, int rows, int width, float* stats, cudaStream_t stream);

int gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias,
             const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, int act,
             cudaStream_t stream);

int ln_film_silu(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                 const float* beta, const float* scale_shift, int ss_stride, int apply_silu, void* out, int out_dtype,
                 cudaStream_t stream);

int eff_attn(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in, void* a_out,
             void* y, int ldy, const int* length, int S, int T, int H, int pair_shift, int mask_v, int dtype,
             cudaStream_t stream);

int timestep_embed(const long long* t, const float* freqs, int S, int half, void* out, int out_dtype,
                   cudaStream_t stream);

int time_table_silu(const float* table, int n_steps, const long long* t, const float* xf_proj, int S, int E, void* out,
                    int out_dtype, cudaStream_t stream);

int tile_rows(const float* table, int period, int width, long long rows, void* out_f16, cudaStream_t stream);

int pack_motion(const float* x, int S, int T, int C, int ld_out, void* out, int out_dtype, cudaStream_t stream);

int ddpm_step(float* x, const void* eps, int ld_eps, int eps_dtype, const float* noise, const long long* t, const float* coef,
              int n_steps, int S, int T, int C, unsigned long long seed, const unsigned long long* seed_dev, void* packed,
              int ld_packed, int packed_dtype, long long* t_next, cudaStream_t stream);

int recover_joints(const float* x, int S, int T, int C, int init_row, const float* mean, const float* stdv,
                   const float* init_mean, const float* init_std, const int* length, int joints_num, float* joints,
                   cudaStream_t stream);

int q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac, const float* sqrt_1mac,
             int S, int TC, float* out, cudaStream_t stream);

// transposed != 0: a_out[s, h] is written as A^T ([l][d]) — the K-major B operand of attn_apply_stylize_tc
int attn_kv(const void* k, const void* v, int ldkv, void* a_out, const int* length, int S, int T, int H,
            int pair_shift, int transposed, cudaStream_t stream);

// y_out (nullable): also write the attention output itself (bf16 [S*T, 512]) — saved by the training forward
int attn_apply_stylize(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                       const float* scale_shift, int ss_stride, int apply_silu, void* out, void* y_out, int S, int T, int H,
                       cudaStream_t stream);

// tcgen05 / TMEM variant (attn_apply_tc.cu): q holds softmax_feat(Q) already, a_t = A^T [S, 8, 64 (l), 64 (d)]
int attn_apply_stylize_tc(const void* q, int ldq, const void* a_t, const float* gamma, const float* beta,
                          const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                          cudaStream_t stream);

// ---- training path (bwd_ops.cu, eff_attn_bwd.cu, gemm_tcgen05.cu) ----
int gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32, int ldo_f32,
                     int k_splits, cudaStream_t stream);

int gemm_bf16_t(int trans_a, int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                const float* bias, const float* residual, int ldr, float* out_f32, int ldo_f32, void* out_bf16,
                int ldo_bf16, int split_k, cudaStream_t stream);

int gemm_bf16_fused(int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                    int act, void* out_bf16, int ldo_bf16, void* out_pre_bf16, int ldo_pre, const void* gate_bf16,
                    int ld_gate, int gate_act, cudaStream_t stream);

int transpose(const void* in, int in_dtype, int M, int N, int ld_in, void* outT, int ld_t, void* copy, int ld_c,
              int out_dtype, float* colsum, int rows_zero_mod, cudaStream_t stream);

int colsum(const void* in, int dtype, int M, int N, int ld, float* out, cudaStream_t stream);

int act_fwd(const void* x, int x_dtype, long long n, int act, void* out, int out_dtype, cudaStream_t stream);
int act_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, long long n, int act, void* dx, int dx_dtype,
            cudaStream_t stream);

int ln_film_silu_bwd(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int apply_silu, const void* dout,
                     int dout_dtype, void* dx, int dx_dtype, int dx_accumulate, float* d_ss, int dss_stride,
                     float* d_gb, int dgb_stride, cudaStream_t stream);

int eff_attn_bwd(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                 const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                 const int* length, int S, int T, int H, int pair_shift, int dtype, float* q_sum, float* k_sum,
                 float* v_sum, cudaStream_t stream);

// ---- text conditioning path (text_ops.cu) ----
int mha_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int B, int N, int H, int causal,
                  int dtype, cudaStream_t stream);

// ---- training step around the denoiser (train_ops.cu) ----
int masked_mse(const float* pred, const float* tgt, const int* length, int S, int T, int C, int pit, float* rows, float* w,
               float* loss, float* d_pred, cudaStream_t stream);
int sumsq(const float* x, long long n, double* out, cudaStream_t stream);
int mean_slices(float* own, const float* staged, long long n, long long stride, int count, float scale, cudaStream_t stream);
int adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
              float eps, int step, const double* gnorm2, float max_norm, cudaStream_t stream);

}  // namespace hig
