// Small HBM-bound kernels around the denoiser: timestep embedding, motion packing, the fused DDPM
// posterior update (one kernel per sampling step) and q_sample for training.
//
// Reference: timestep_embedding                  codes/models/interaction_transformer.py:26-43
//            embed_motion (two_embed layout)     :593-602
//            p_sample / p_mean_variance / _predict_xstart_from_eps / q_posterior_mean_variance
//                                                codes/models/gaussian_diffusion.py:606-666, 443-537, 539-544, 419-441
//            q_sample                            :399-417
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

// ------------------------------------------------------------------------------------------------
// timestep embedding: out[s, :] = [cos(t_s * f_i) | sin(t_s * f_i)],  f = host-computed fp32 table
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void timestep_embed_kernel(const long long* __restrict__ t, const float* __restrict__ freqs, int S, int half,
                                      TOut* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * half) return;
  const int s = idx / half, i = idx - s * half;
  const float arg = __fmul_rn(static_cast<float>(t[s]), freqs[i]);
  out[(size_t)s * 2 * half + i] = static_cast<TOut>(cosf(arg));
  out[(size_t)s * 2 * half + half + i] = static_cast<TOut>(sinf(arg));
}

int timestep_embed(const long long* t, const float* freqs, int S, int half, void* out, int out_dtype,
                   cudaStream_t stream) {
  if (!t || !freqs || !out || S <= 0 || half <= 0) return set_error(HIG_ERR_INVALID, "timestep_embed: bad arguments");
  const int n = S * half, blocks = (n + 255) / 256;
  if (out_dtype == HIG_BF16)
    timestep_embed_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(t, freqs, S, half, (__nv_bfloat16*)out);
  else
    timestep_embed_kernel<float><<<blocks, 256, 0, stream>>>(t, freqs, S, half, (float*)out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("timestep_embed launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// tile_rows: out[r, :] = fp16(table[r % period, :]) — seeds the fp16 residual stream with the positional rows
// (sequence_embedding + joint_embed bias, :593-602) so that the motion embedding becomes an in-place
// `stream += A.W^T` projection that also leaves the LayerNorm row statistics behind.
// ------------------------------------------------------------------------------------------------
__global__ void tile_rows_kernel(const float* __restrict__ table, int period, int width, long long rows, __half* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-column group
  const int gpr = width >> 3;
  pdl_wait();      // the stream is still read by the previous step's last kernels
  pdl_trigger();
  if (idx >= rows * gpr) return;
  const long long r = idx / gpr;
  const int c = (int)(idx - r * gpr) * 8;
  const float* src = table + (size_t)(r % period) * width + c;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
  __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
  *reinterpret_cast<uint4*>(out + (size_t)r * width + c) = *reinterpret_cast<uint4*>(h);
}

int tile_rows(const float* table, int period, int width, long long rows, void* out_f16, cudaStream_t stream) {
  if (!table || !out_f16 || period <= 0 || width <= 0 || (width % 8) || rows <= 0)
    return set_error(HIG_ERR_INVALID, "tile_rows: bad arguments (width must be a multiple of 8)");
  const long long n = rows * (width / 8);
  cudaError_t e = launch_pdl(tile_rows_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, table, period, width, rows,
                             (__half*)out_f16);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("tile_rows launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// time_table_silu: out[s, :] = SiLU(table[t_s, :] + xf_proj[s, :]).  table[n, :] = time_embed(timestep_embedding(n))
// (:474-478, :591) for every timestep of the schedule, computed once per set of weights with the same GEMM kernels, so
// the sampling loop replaces the sinusoid kernel and the two M = S GEMMs of the time MLP (8 CTAs each, latency-bound)
// by this gather.  Same association order as the GEMM epilogue it replaces: (acc + bias) + residual, then SiLU.
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void time_table_silu_kernel(const float* __restrict__ table, int n_steps, const long long* __restrict__ t,
                                       const float* __restrict__ xf_proj, int S, int E, TOut* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();      // t is decremented by the previous step's posterior kernel
  pdl_trigger();
  if (idx >= S * (E / 4)) return;
  const int s = idx / (E / 4), c = (idx - s * (E / 4)) * 4;
  long long ts = t[s];
  ts = ts < 0 ? 0 : (ts >= n_steps ? n_steps - 1 : ts);
  const float4 a = *reinterpret_cast<const float4*>(table + (size_t)ts * E + c);
  const float4 b = *reinterpret_cast<const float4*>(xf_proj + (size_t)s * E + c);
  const float v[4] = {silu_f(a.x + b.x), silu_f(a.y + b.y), silu_f(a.z + b.z), silu_f(a.w + b.w)};
#pragma unroll
  for (int j = 0; j < 4; ++j) out[(size_t)s * E + c + j] = static_cast<TOut>(v[j]);
}

int time_table_silu(const float* table, int n_steps, const long long* t, const float* xf_proj, int S, int E, void* out,
                    int out_dtype, cudaStream_t stream) {
  if (!table || !t || !xf_proj || !out || S <= 0 || E <= 0 || (E % 4) || n_steps <= 0)
    return set_error(HIG_ERR_INVALID, "time_table_silu: bad arguments (E must be a multiple of 4)");
  const int n = S * (E / 4), blocks = (n + 255) / 256;
  cudaError_t e;
  if (out_dtype == HIG_BF16)
    e = launch_pdl(time_table_silu_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, table, n_steps, t, xf_proj, S, E,
                   (__nv_bfloat16*)out);
  else
    e = launch_pdl(time_table_silu_kernel<float>, dim3(blocks), dim3(256), 0, stream, table, n_steps, t, xf_proj, S, E,
                   (float*)out);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("time_table_silu launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// pack_motion: x fp32 [S,T,C] -> GEMM operand [S*T, ld] such that ONE projection with the augmented weight
// [joint_embed.weight | joint_embed2.weight | 0] reproduces embed_motion:
//   row t>=1 : cols [0,C) = x[s,t,:],     cols [C,ld) = 0
//   row t==0 : cols [0,C) = 0,            cols [C,C+4) = x[s,0,:4], rest 0
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void pack_motion_kernel(const float* __restrict__ x, int rows, int T, int C, int ld, TOut* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * ld) return;
  const int row = (int)(idx / ld), c = (int)(idx - (long long)row * ld);
  const int t = row % T;
  float v = 0.f;
  if (t > 0) {
    if (c < C) v = x[(size_t)row * C + c];
  } else {
    if (c >= C && c < C + 4) v = x[(size_t)row * C + (c - C)];
  }
  out[idx] = static_cast<TOut>(v);
}

int pack_motion(const float* x, int S, int T, int C, int ld_out, void* out, int out_dtype, cudaStream_t stream) {
  if (!x || !out || S <= 0 || T <= 0 || C <= 0 || ld_out < C + 4)
    return set_error(HIG_ERR_INVALID, "pack_motion: bad arguments");
  const long long n = (long long)S * T * ld_out;
  const int blocks = (int)((n + 255) / 256);
  if (out_dtype == HIG_BF16)
    pack_motion_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(x, S * T, T, C, ld_out, (__nv_bfloat16*)out);
  else
    pack_motion_kernel<float><<<blocks, 256, 0, stream>>>(x, S * T, T, C, ld_out, (float*)out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("pack_motion launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (production noise; parity runs inject a noise tensor instead)
// ------------------------------------------------------------------------------------------------
HIG_DEVICE uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
HIG_DEVICE float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;           // [0,1)
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// ------------------------------------------------------------------------------------------------
// ddpm_step: x <- c1_t (r_t x - m_t eps) + c2_t x + 1[t>0] sigma_t z          (fp32, reference op order, no FMA
// contraction) and, in the same pass, the packed GEMM operand of the next denoiser call.
// coef = [5][n_steps] fp32: r = sqrt(1/abar), m = sqrt(1/abar - 1), c1, c2, sigma = exp(0.5*logvar_clipped)
// ------------------------------------------------------------------------------------------------
template <typename TPack, typename TEps>
__global__ void __launch_bounds__(256)
ddpm_step_kernel(float* __restrict__ x, const TEps* __restrict__ eps, int ld_eps, const float* __restrict__ noise,
                 const long long* __restrict__ t, const float* __restrict__ coef, int n_steps, int S, int T, int C,
                 unsigned long long seed, const unsigned long long* __restrict__ seed_dev, TPack* __restrict__ packed,
                 int ld_packed) {
  const long long per_seq = (long long)T * C;
  const long long total = (long long)S * per_seq;
  const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  pdl_wait();      // eps is the output of the previous kernel
  pdl_trigger();
  if (base >= total) return;
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (noise) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j < total) z[j] = noise[base + j];
  } else {
    // counter = (element-group index, timestep of the first element's sequence), key = seed; a seed kept in device
    // memory lets one captured graph serve every sampling call (the host rewrites 8 bytes instead of re-capturing)
    if (seed_dev != nullptr) seed = *seed_dev;
    const int s0 = (int)((unsigned)base / (unsigned)per_seq);
    const long long ts = t[s0];
    const uint4 ctr = make_uint4((uint32_t)(base >> 2), (uint32_t)((base >> 2) >> 32), (uint32_t)ts, 0x48494742u);
    const uint4 rnd = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 n0 = box_muller(rnd.x, rnd.y), n1 = box_muller(rnd.z, rnd.w);
    z[0] = n0.x; z[1] = n0.y; z[2] = n1.x; z[3] = n1.y;
  }
  // decompose the first element once (32-bit arithmetic: the host guarantees total < 2^31), then step
  int s = (int)((unsigned)base / (unsigned)per_seq);
  int rem = (int)((unsigned)base - (unsigned)s * (unsigned)per_seq);
  int tt = rem / C, c = rem - tt * C;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long idx = base + j;
    if (idx >= total) break;
    long long ts = t[s];
    ts = ts < 0 ? 0 : (ts >= n_steps ? n_steps - 1 : ts);
    const float r = coef[ts], m = coef[n_steps + ts], c1 = coef[2 * n_steps + ts], c2 = coef[3 * n_steps + ts];
    const float sigma = ts > 0 ? coef[4 * n_steps + ts] : 0.f;
    const float xv = x[idx];
    const float ev = static_cast<float>(eps[((size_t)s * T + tt) * ld_eps + c]);
    const float x0 = __fsub_rn(__fmul_rn(r, xv), __fmul_rn(m, ev));
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
    const float xn = __fadd_rn(mean, __fmul_rn(sigma, z[j]));
    x[idx] = xn;
    if (packed) {
      TPack* prow = packed + ((size_t)s * T + tt) * ld_packed;
      if (tt > 0) prow[c] = static_cast<TPack>(xn);
      else if (c < 4) prow[C + c] = static_cast<TPack>(xn);
    }
    if (++c == C) {
      c = 0;
      if (++tt == T) { tt = 0; ++s; }
    }
  }
}

__global__ void advance_t_kernel(const long long* __restrict__ t, long long* __restrict__ t_next, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) t_next[i] = t[i] - 1;
}

int ddpm_step(float* x, const void* eps, int ld_eps, int eps_dtype, const float* noise, const long long* t, const float* coef,
              int n_steps, int S, int T, int C, unsigned long long seed, const unsigned long long* seed_dev, void* packed,
              int ld_packed, int packed_dtype, long long* t_next, cudaStream_t stream) {
  if (!x || !eps || !t || !coef || S <= 0 || T <= 0 || C <= 0 || n_steps <= 0)
    return set_error(HIG_ERR_INVALID, "ddpm_step: bad arguments");
  if (eps_dtype != HIG_F32 && eps_dtype != HIG_F16) return set_error(HIG_ERR_INVALID, "ddpm_step: eps is fp32 or fp16");
  if (packed && ld_packed < C + 4) return set_error(HIG_ERR_INVALID, "ddpm_step: ld_packed < C+4");
  const long long total = (long long)S * T * C;
  if (total >= (1LL << 31)) return set_error(HIG_ERR_UNSUPPORTED, "ddpm_step: more than 2^31 elements");
  const int blocks = (int)(((total + 3) / 4 + 255) / 256);
  const bool pk16 = packed && packed_dtype == HIG_BF16;
  if (eps_dtype == HIG_F16) {
    if (!pk16) return set_error(HIG_ERR_UNSUPPORTED, "ddpm_step: fp16 eps goes with the bf16 packed operand (product path)");
    launch_pdl(ddpm_step_kernel<__nv_bfloat16, __half>, dim3(blocks), dim3(256), 0, stream, x, (const __half*)eps, ld_eps,
               noise, t, coef, n_steps, S, T, C, seed, seed_dev, (__nv_bfloat16*)packed, ld_packed);
  } else if (pk16) {
    launch_pdl(ddpm_step_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(256), 0, stream, x, (const float*)eps, ld_eps, noise,
               t, coef, n_steps, S, T, C, seed, seed_dev, (__nv_bfloat16*)packed, ld_packed);
  } else {
    ddpm_step_kernel<float, float><<<blocks, 256, 0, stream>>>(x, (const float*)eps, ld_eps, noise, t, coef, n_steps, S, T, C,
                                                                seed, seed_dev, (float*)packed, ld_packed);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("ddpm_step launch: ") + cudaGetErrorString(e));
  count_launch();
  if (t_next) {
    advance_t_kernel<<<(S + 255) / 256, 256, 0, stream>>>(t, t_next, S);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("advance_t launch: ") + cudaGetErrorString(e));
    count_launch();
  }
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// q_sample: x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) noise    (per-sequence t)
// ------------------------------------------------------------------------------------------------
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const long long* __restrict__ t, const float* __restrict__ sqrt_ac,
                                const float* __restrict__ sqrt_1mac, long long total, int TC, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long ts = t[idx / TC];
  out[idx] = __fadd_rn(__fmul_rn(sqrt_ac[ts], x0[idx]), __fmul_rn(sqrt_1mac[ts], noise[idx]));
}

int q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac, const float* sqrt_1mac,
             int S, int TC, float* out, cudaStream_t stream) {
  if (!x0 || !noise || !t || !sqrt_ac || !sqrt_1mac || !out || S <= 0 || TC <= 0)
    return set_error(HIG_ERR_INVALID, "q_sample: bad arguments");
  const long long total = (long long)S * TC;
  q_sample_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(x0, noise, t, sqrt_ac, sqrt_1mac, total, TC, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("q_sample launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
