// hig_ln_film_silu: fused LayerNorm (+ FiLM scale/shift from the timestep/text embedding) (+ SiLU).
//
// Replaces, in one coalesced pass, the chain  nn.LayerNorm -> h*(1+scale)+shift -> nn.SiLU  of
// StylizationBlock.forward (codes/models/interaction_transformer.py:86-97) and the plain pre-attention
// LayerNorms (:119,153,155,190,194).  The reference computes self.norm(x) three times per attention block;
// here it is computed once and written in the GEMM operand type.
//
// A warp processes one row at a time (row width 512 = latent_dim, or 256 = text_latent_dim); every lane owns 8
// contiguous elements per 256-wide chunk, so loads are 2x float4 (fp32 in) or 1x uint4 (bf16 in), stores 16 B.
// Statistics are two-pass in registers (mean, then centred variance) in fp32, eps = 1e-5, biased variance —
// the same arithmetic as ATen's native_layer_norm.
#include <cuda_fp16.h>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

template <typename T> struct Io;
template <> struct Io<float> {
  static HIG_DEVICE void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static HIG_DEVICE void store8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Io<__half> {
  static HIG_DEVICE void load8(const __half* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <> struct Io<__nv_bfloat16> {
  static HIG_DEVICE void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
  }
  static HIG_DEVICE void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// Each warp owns a run of `rows_per_warp` consecutive rows inside ONE sequence, so gamma, beta and that
// sequence's (scale, shift) are loaded once per warp into registers and the row loop only streams x: parameter
// traffic through L1 drops from 8 KB per row to 8 KB per run.  The arithmetic keeps the reference's order
// ((n * gamma + beta) * (1 + scale) + shift).  Two rows are in flight per iteration to cover HBM latency.
constexpr int LN_WARPS = 4;

template <int WIDTH, typename TIn, typename TOut>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_film_silu_kernel(const TIn* __restrict__ x, int rows, int rows_per_seq, int rows_per_warp, int runs_per_seq,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ scale_shift, int ss_stride, int apply_silu, TOut* __restrict__ out) {
  constexpr int CHUNKS = WIDTH / 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int run = blockIdx.x * LN_WARPS + warp;
  const int seq = run / runs_per_seq, piece = run - seq * runs_per_seq;
  int r0 = seq * rows_per_seq + piece * rows_per_warp;
  int r1 = min(r0 + rows_per_warp, (seq + 1) * rows_per_seq);
  r1 = min(r1, rows);
  if (r0 >= r1) return;

  float A[CHUNKS][8], B[CHUNKS][8];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    Io<float>::load8(gamma + c * 256 + lane * 8, A[c]);
    Io<float>::load8(beta + c * 256 + lane * 8, B[c]);
  }
  // bf16 output (product path): fold FiLM into the affine once per run, y = n * A' + B' with A' = gamma (1 + scale),
  // B' = beta (1 + scale) + shift, and use single-MUFU sigmoid (tanh.approx) — the kernel is issue-bound, not
  // HBM-bound, at ~25 instructions per element.  fp32 output (fp32 mode): reference evaluation order, precise expf.
  constexpr bool kFast = sizeof(TOut) == 2;
  // gamma / beta are parameters (loaded while the previous kernel drains); scale_shift and x are kernel outputs
  pdl_wait();
  pdl_trigger();
  float SC[CHUNKS][8], SH[CHUNKS][8];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int col = c * 256 + lane * 8;
    if (scale_shift) {
      const float* ssp = scale_shift + (size_t)seq * ss_stride;
      Io<float>::load8(ssp + col, SC[c]);
      Io<float>::load8(ssp + WIDTH + col, SH[c]);
      if (kFast) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float m1 = 1.0f + SC[c][j];
          A[c][j] *= m1;
          B[c][j] = fmaf(B[c][j], m1, SH[c][j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { SC[c][j] = 0.f; SH[c][j] = 0.f; }
    }
  }

  auto finish = [&](float (&v)[CHUNKS][8], int row) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    const float mean = warp_sum(s) * (1.0f / WIDTH);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[c][j] -= mean;
        ss = fmaf(v[c][j], v[c][j], ss);
      }
    const float var = warp_sum(ss) * (1.0f / WIDTH) + 1e-5f;
    const float rstd = kFast ? rsqrtf(var) : 1.0f / sqrtf(var);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t;
        if (kFast) {
          t = fmaf(v[c][j] * rstd, A[c][j], B[c][j]);
          if (apply_silu) {
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * t));
            t = t * fmaf(th, 0.5f, 0.5f);
          }
        } else {
          t = v[c][j] * rstd * A[c][j] + B[c][j];
          if (scale_shift) t = t * (1.0f + SC[c][j]) + SH[c][j];
          if (apply_silu) t = t / (1.0f + expf(-t));
        }
        o[j] = t;
      }
      Io<TOut>::store8(out + (size_t)row * WIDTH + c * 256 + lane * 8, o);
    }
  };

  int r = r0;
  for (; r + 1 < r1; r += 2) {
    float v0[CHUNKS][8], v1[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      Io<TIn>::load8(x + (size_t)r * WIDTH + c * 256 + lane * 8, v0[c]);
      Io<TIn>::load8(x + (size_t)(r + 1) * WIDTH + c * 256 + lane * 8, v1[c]);
    }
    finish(v0, r);
    finish(v1, r + 1);
  }
  if (r < r1) {
    float v0[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) Io<TIn>::load8(x + (size_t)r * WIDTH + c * 256 + lane * 8, v0[c]);
    finish(v0, r);
  }
}

template <int WIDTH, typename TIn, typename TOut>
static void launch_ln(const void* x, int rows, int rows_per_seq, const float* gamma, const float* beta,
                      const float* scale_shift, int ss_stride, int apply_silu, void* out, cudaStream_t stream) {
  // aim for ~24 runs (warps) per SM; a run never crosses a sequence boundary
  const int n_seq = (rows + rows_per_seq - 1) / rows_per_seq;
  int rows_per_warp = (int)(((long long)rows + 148 * 24 - 1) / (148 * 24));
  if (rows_per_warp < 1) rows_per_warp = 1;
  if (rows_per_warp > rows_per_seq) rows_per_warp = rows_per_seq;
  const int runs_per_seq = (rows_per_seq + rows_per_warp - 1) / rows_per_warp;
  const long long runs = (long long)n_seq * runs_per_seq;
  const int blocks = (int)((runs + LN_WARPS - 1) / LN_WARPS);
  launch_pdl(ln_film_silu_kernel<WIDTH, TIn, TOut>, dim3(blocks), dim3(LN_WARPS * 32), 0, stream,
             reinterpret_cast<const TIn*>(x), rows, rows_per_seq, rows_per_warp, runs_per_seq, gamma, beta, scale_shift,
             ss_stride, apply_silu, reinterpret_cast<TOut*>(out));
}

int ln_film_silu(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                 const float* beta, const float* scale_shift, int ss_stride, int apply_silu, void* out, int out_dtype,
                 cudaStream_t stream) {
  if (!x || !out || !gamma || !beta || rows <= 0) return set_error(HIG_ERR_INVALID, "ln_film_silu: bad arguments");
  if (width != 512 && width != 256) return set_error(HIG_ERR_UNSUPPORTED, "ln_film_silu: width must be 256 or 512");
  if (rows_per_seq <= 0) rows_per_seq = 1;
  if (!scale_shift) rows_per_seq = rows;  // no per-sequence parameters: runs may span sequences
  if (scale_shift && (ss_stride % 4)) return set_error(HIG_ERR_INVALID, "ln_film_silu: ss_stride % 4 != 0");
  using bf = __nv_bfloat16;
#define HIG_LN_CASE(W, TI, TO) \
  launch_ln<W, TI, TO>(x, rows, rows_per_seq, gamma, beta, scale_shift, ss_stride, apply_silu, out, stream)
  if (width == 512) {
    if (x_dtype == HIG_F32 && out_dtype == HIG_BF16) HIG_LN_CASE(512, float, bf);
    else if (x_dtype == HIG_F16 && out_dtype == HIG_BF16) HIG_LN_CASE(512, __half, bf);
    else if (x_dtype == HIG_BF16 && out_dtype == HIG_BF16) HIG_LN_CASE(512, bf, bf);
    else if (x_dtype == HIG_F32 && out_dtype == HIG_F32) HIG_LN_CASE(512, float, float);
    else return set_error(HIG_ERR_UNSUPPORTED, "ln_film_silu: dtype combination");
  } else {
    if (x_dtype == HIG_F32 && out_dtype == HIG_BF16) HIG_LN_CASE(256, float, bf);
    else if (x_dtype == HIG_BF16 && out_dtype == HIG_BF16) HIG_LN_CASE(256, bf, bf);
    else if (x_dtype == HIG_F32 && out_dtype == HIG_F32) HIG_LN_CASE(256, float, float);
    else return set_error(HIG_ERR_UNSUPPORTED, "ln_film_silu: dtype combination");
  }
#undef HIG_LN_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("ln_film_silu launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// Row statistics of the fp16 residual stream for the LayerNorm-folded projections of gemm_stream.cu: stats[row] =
// {sum, sum of squares, 0, 0, 0, 0, 0, 0} (the four-partial layout the out-projection epilogue writes; here the whole
// row goes into partial 0).  Only the motion-embedding output needs it — every later LayerNorm input gets its
// statistics from the epilogue that produced it.
__global__ void __launch_bounds__(256) row_stats_kernel(const __half* __restrict__ x, int rows, float* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    float a[8], b[8];
    Io<__half>::load8(x + (size_t)row * 512 + lane * 8, a);
    Io<__half>::load8(x + (size_t)row * 512 + 256 + lane * 8, b);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += a[i] + b[i]; s2 = fmaf(a[i], a[i], fmaf(b[i], b[i], s2)); }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane < 2) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane == 0) { o.x = s1; o.y = s2; }
      reinterpret_cast<float4*>(stats + (size_t)row * 8)[lane] = o;
    }
  }
}

int row_stats(const void* x, int x_dtype, int rows, int width, float* stats, cudaStream_t stream) {
  if (!x || !stats || rows <= 0) return set_error(HIG_ERR_INVALID, "row_stats: bad arguments");
  if (x_dtype != HIG_F16 || width != 512) return set_error(HIG_ERR_UNSUPPORTED, "row_stats: fp16 rows of width 512");
  const int blocks = min((rows + 7) / 8, 148 * 8);
  cudaError_t e = launch_pdl(row_stats_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const __half*>(x), rows, stats);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("row_stats launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
