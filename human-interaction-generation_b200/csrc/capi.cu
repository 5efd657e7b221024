// extern "C" boundary of libhig_b200.so — see include/hig_b200.h for the contract of every entry point.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include "hig_internal.h"

namespace hig {
static std::mutex g_err_mu;
static std::string g_err;
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const std::string& msg) {
  std::lock_guard<std::mutex> g(g_err_mu);
  g_err = msg;
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// HIG_DETERMINISTIC=1 (read at every call, so a test can toggle it): reductions that are normally combined with floating-point
// atomics across CTAs (column sums, the gradient norm) run with ONE contributor per address, in a fixed order
bool deterministic() {
  const char* e = getenv("HIG_DETERMINISTIC");
  return e != nullptr && e[0] != '\0' && e[0] != '0';
}

bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("HIG_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace hig

extern "C" {

int hig_version(void) { return 100; }

const char* hig_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> g(hig::g_err_mu);
  copy = hig::g_err;
  return copy.c_str();
}

unsigned long long hig_launch_count(void) { return hig::g_launches.load(std::memory_order_relaxed); }

int hig_l2_persist(const void* ptr, unsigned long long bytes, float hit_ratio, void* stream) {
  // Pin a hot buffer (the fp32 residual stream) in the 126 MB L2 with an access-policy window on `stream`:
  // every kernel launched (or captured) on that stream afterwards treats [ptr, ptr+bytes) as persisting.
  int dev = 0, max_persist = 0, max_window = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (ptr && bytes > 0 && max_persist > 0) {
    size_t want = bytes < (size_t)max_persist ? (size_t)bytes : (size_t)max_persist;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cap);
    if (cap == cudaStreamCaptureStatusNone) {  // device limits cannot change while a capture is open
      cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      if (e != cudaSuccess) return hig::set_error(HIG_ERR_CUDA, std::string("cudaDeviceSetLimit: ") + cudaGetErrorString(e));
    }
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
    attr.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? (size_t)bytes : (size_t)max_window;
    attr.accessPolicyWindow.hitRatio = hit_ratio;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;  // disable
  }
  cudaError_t e = cudaStreamSetAttribute(static_cast<cudaStream_t>(stream), cudaStreamAttributeAccessPolicyWindow, &attr);
  if (e != cudaSuccess) return hig::set_error(HIG_ERR_CUDA, std::string("cudaStreamSetAttribute: ") + cudaGetErrorString(e));
  return HIG_OK;
}

int hig_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                  const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, void* out_bf16,
                  int ldo_bf16, int act, void* stream) {
  return hig::gemm_bf16(A, lda, W, ldw, M, N, K, bias, residual, ldr, res_row_mod, out_f32, ldo_f32, out_bf16,
                        ldo_bf16, act, static_cast<cudaStream_t>(stream));
}

int hig_gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                     const void* residual, int res_dtype, int ldr, int res_row_mod, void* out, int out_dtype, int ldo,
                     void* out_bf16, int ldo_bf16, int act, void* stream) {
  return hig::gemm_bf16_ex(A, lda, W, ldw, M, N, K, bias, residual, res_dtype, ldr, res_row_mod, out, out_dtype, ldo,
                           out_bf16, ldo_bf16, act, static_cast<cudaStream_t>(stream));
}

int hig_gemm_stream(int kind, const void* A, int lda, const void* W, int ldw, int op_dtype, int M, int N, int K,
                    const float* bias, const float* wsum, const float* stats_in, float* stats_out, int ln_width,
                    void* out, int ldo, void* stream) {
  return hig::gemm_stream(kind, A, lda, W, ldw, op_dtype, M, N, K, bias, wsum, stats_in, stats_out, ln_width, out, ldo,
                          static_cast<cudaStream_t>(stream));
}

int hig_row_stats(const void* x, int x_dtype, int rows, int width, float* stats, void* stream) {
  return hig::row_stats(x, x_dtype, rows, width, stats, static_cast<cudaStream_t>(stream));
}

int hig_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias,
                 const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, int act, void* stream) {
  return hig::gemm_f32(A, lda, W, ldw, M, N, K, bias, residual, ldr, res_row_mod, out_f32, ldo_f32, act,
                       static_cast<cudaStream_t>(stream));
}

int hig_ln_film_silu(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int apply_silu, void* out,
                     int out_dtype, void* stream) {
  return hig::ln_film_silu(x, x_dtype, rows, width, rows_per_seq, gamma, beta, scale_shift, ss_stride, apply_silu, out,
                           out_dtype, static_cast<cudaStream_t>(stream));
}

int hig_eff_attn(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                 void* a_out, void* y, int ldy, const int* length, int S, int T, int H, int pair_shift, int mask_v,
                 int dtype, void* stream) {
  return hig::eff_attn(mode, q, ldq, k, v, ldkv, a_in, a_out, y, ldy, length, S, T, H, pair_shift, mask_v, dtype,
                       static_cast<cudaStream_t>(stream));
}

int hig_timestep_embed(const long long* t, const float* freqs, int S, int half, void* out, int out_dtype,
                       void* stream) {
  return hig::timestep_embed(t, freqs, S, half, out, out_dtype, static_cast<cudaStream_t>(stream));
}

int hig_pack_motion(const float* x, int S, int T, int C, int ld_out, void* out, int out_dtype, void* stream) {
  return hig::pack_motion(x, S, T, C, ld_out, out, out_dtype, static_cast<cudaStream_t>(stream));
}

int hig_ddpm_step(float* x, const void* eps, int ld_eps, int eps_dtype, const float* noise, const long long* t, const float* coef,
                  int n_steps, int S, int T, int C, unsigned long long seed, const unsigned long long* seed_dev, void* packed,
                  int ld_packed, int packed_dtype, long long* t_next, void* stream) {
  return hig::ddpm_step(x, eps, ld_eps, eps_dtype, noise, t, coef, n_steps, S, T, C, seed, seed_dev, packed, ld_packed, packed_dtype,
                        t_next, static_cast<cudaStream_t>(stream));
}

int hig_time_table_silu(const float* table, int n_steps, const long long* t, const float* xf_proj, int S, int E, void* out,
                        int out_dtype, void* stream) {
  return hig::time_table_silu(table, n_steps, t, xf_proj, S, E, out, out_dtype, static_cast<cudaStream_t>(stream));
}

int hig_tile_rows(const float* table, int period, int width, long long rows, void* out_f16, void* stream) {
  return hig::tile_rows(table, period, width, rows, out_f16, static_cast<cudaStream_t>(stream));
}

int hig_debug_trace(unsigned long long* buf, int max_launches) {
  hig::set_gemm_trace(buf, max_launches);
  return HIG_OK;
}

int hig_set_sm_limit(int n) {
  hig::set_sm_limit(n);
  return HIG_OK;
}

int hig_debug_saturation(unsigned long long* counter) {
  hig::set_saturation_counter(counter);
  return HIG_OK;
}

int hig_recover_joints(const float* x, int S, int T, int C, int init_row, const float* mean, const float* std_,
                       const float* init_mean, const float* init_std, const int* length, int joints_num, float* joints,
                       void* stream) {
  return hig::recover_joints(x, S, T, C, init_row, mean, std_, init_mean, init_std, length, joints_num, joints,
                             static_cast<cudaStream_t>(stream));
}

int hig_q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                 const float* sqrt_1mac, int S, int TC, float* out, void* stream) {
  return hig::q_sample(x0, noise, t, sqrt_ac, sqrt_1mac, S, TC, out, static_cast<cudaStream_t>(stream));
}

int hig_attn_apply_stylize(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                           const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                           void* stream) {
  return hig::attn_apply_stylize(q, ldq, a_in, gamma, beta, scale_shift, ss_stride, apply_silu, out, nullptr, S, T, H,
                                 static_cast<cudaStream_t>(stream));
}

int hig_attn_apply_stylize_y(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                             const float* scale_shift, int ss_stride, int apply_silu, void* out, void* y_out, int S, int T,
                             int H, void* stream) {
  return hig::attn_apply_stylize(q, ldq, a_in, gamma, beta, scale_shift, ss_stride, apply_silu, out, y_out, S, T, H,
                                 static_cast<cudaStream_t>(stream));
}

int hig_attn_kv(const void* k, const void* v, int ldkv, void* a_out, const int* length, int S, int T, int H,
                int pair_shift, int transposed, void* stream) {
  return hig::attn_kv(k, v, ldkv, a_out, length, S, T, H, pair_shift, transposed, static_cast<cudaStream_t>(stream));
}

int hig_attn_apply_stylize_tc(const void* q, int ldq, const void* a_t, const float* gamma, const float* beta,
                              const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                              void* stream) {
  return hig::attn_apply_stylize_tc(q, ldq, a_t, gamma, beta, scale_shift, ss_stride, apply_silu, out, S, T, H,
                                    static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- training path
int hig_gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32,
                         int ldo_f32, int k_splits, void* stream) {
  return hig::gemm_bf16_splitk(A, lda, W, ldw, M, N, K, out_f32, ldo_f32, k_splits, static_cast<cudaStream_t>(stream));
}

int hig_gemm_bf16_t(int trans_a, int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const float* bias, const float* residual, int ldr, float* out_f32, int ldo_f32, void* out_bf16,
                    int ldo_bf16, int split_k, void* stream) {
  return hig::gemm_bf16_t(trans_a, trans_b, A, lda, W, ldw, M, N, K, bias, residual, ldr, out_f32, ldo_f32, out_bf16,
                          ldo_bf16, split_k, static_cast<cudaStream_t>(stream));
}

int hig_gemm_bf16_fused(int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                        int act, void* out_bf16, int ldo_bf16, void* out_pre_bf16, int ldo_pre, const void* gate_bf16,
                        int ld_gate, int gate_act, void* stream) {
  return hig::gemm_bf16_fused(trans_b, A, lda, W, ldw, M, N, K, bias, act, out_bf16, ldo_bf16, out_pre_bf16, ldo_pre,
                              gate_bf16, ld_gate, gate_act, static_cast<cudaStream_t>(stream));
}

int hig_transpose(const void* in, int in_dtype, int M, int N, int ld_in, void* outT, int ld_t, void* copy, int ld_c,
                  int out_dtype, float* colsum, int rows_zero_mod, void* stream) {
  return hig::transpose(in, in_dtype, M, N, ld_in, outT, ld_t, copy, ld_c, out_dtype, colsum, rows_zero_mod,
                        static_cast<cudaStream_t>(stream));
}

int hig_colsum(const void* in, int dtype, int M, int N, int ld, float* out, void* stream) {
  return hig::colsum(in, dtype, M, N, ld, out, static_cast<cudaStream_t>(stream));
}

int hig_act_fwd(const void* x, int x_dtype, long long n, int act, void* out, int out_dtype, void* stream) {
  return hig::act_fwd(x, x_dtype, n, act, out, out_dtype, static_cast<cudaStream_t>(stream));
}

int hig_act_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, long long n, int act, void* dx, int dx_dtype,
                void* stream) {
  return hig::act_bwd(x, x_dtype, dy, dy_dtype, n, act, dx, dx_dtype, static_cast<cudaStream_t>(stream));
}

int hig_ln_film_silu_bwd(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                         const float* beta, const float* scale_shift, int ss_stride, int apply_silu, const void* dout,
                         int dout_dtype, void* dx, int dx_dtype, int dx_accumulate, float* d_ss, int dss_stride,
                         float* d_gb, int dgb_stride, void* stream) {
  return hig::ln_film_silu_bwd(x, x_dtype, rows, width, rows_per_seq, gamma, beta, scale_shift, ss_stride, apply_silu,
                               dout, dout_dtype, dx, dx_dtype, dx_accumulate, d_ss, dss_stride, d_gb, dgb_stride,
                               static_cast<cudaStream_t>(stream));
}

int hig_eff_attn_bwd(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                     const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                     const int* length, int S, int T, int H, int pair_shift, int dtype, void* stream) {
  return hig::eff_attn_bwd(mode, q, ldq, k, v, ldkv, a_in, dy, lddy, dq, lddq, dk, dv, lddkv, dA, length, S, T, H,
                           pair_shift, dtype, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int hig_eff_attn_bwd_sums(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                          const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                          const int* length, int S, int T, int H, int pair_shift, int dtype, float* q_sum, float* k_sum,
                          float* v_sum, void* stream) {
  return hig::eff_attn_bwd(mode, q, ldq, k, v, ldkv, a_in, dy, lddy, dq, lddq, dk, dv, lddkv, dA, length, S, T, H,
                           pair_shift, dtype, q_sum, k_sum, v_sum, static_cast<cudaStream_t>(stream));
}

int hig_mha_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int B, int N, int H,
                      int causal, int dtype, void* stream) {
  return hig::mha_attention(q, k, v, ld, out, ldo, B, N, H, causal, dtype, static_cast<cudaStream_t>(stream));
}

int hig_masked_mse(const float* pred, const float* tgt, const int* length, int S, int T, int C, int pit, float* rows,
                   float* w, float* loss, float* d_pred, void* stream) {
  return hig::masked_mse(pred, tgt, length, S, T, C, pit, rows, w, loss, d_pred, static_cast<cudaStream_t>(stream));
}

int hig_sumsq(const float* x, long long n, double* out, void* stream) {
  return hig::sumsq(x, n, out, static_cast<cudaStream_t>(stream));
}

int hig_mean_slices(float* own, const float* staged, long long n, long long stride, int count, float scale, void* stream) {
  return hig::mean_slices(own, staged, n, stride, count, scale, static_cast<cudaStream_t>(stream));
}

int hig_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                  float beta2, float eps, int step, const double* gnorm2, float max_norm, void* stream) {
  return hig::adam_flat(p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, step, gnorm2, max_norm,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
