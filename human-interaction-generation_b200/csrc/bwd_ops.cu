// Training-path helper kernels (HBM-bound): operand transposition for the dgrad/wgrad GEMMs, column sums (bias and
// positional-table gradients), elementwise activation forward/backward, and the backward of the fused
// LayerNorm + FiLM + SiLU kernel.
//
// Reference: the reference has no backward code of its own — torch.autograd differentiates
//   StylizationBlock.forward            codes/models/interaction_transformer.py:86-97
//   nn.LayerNorm call sites             :119,153,155,190,194
//   FFN.forward (exact erf GELU)        :261-264
//   time_embed (SiLU)                   :474-478
// These kernels are the hand-written derivatives of the forward kernels in ln_film.cu / gemm_*.cu.
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

template <typename T> HIG_DEVICE float ld_as_f(const T* p);
template <> HIG_DEVICE float ld_as_f<float>(const float* p) { return *p; }
template <> HIG_DEVICE float ld_as_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> HIG_DEVICE void st_from_f(T* p, float v);
template <> HIG_DEVICE void st_from_f<float>(float* p, float v) { *p = v; }
template <> HIG_DEVICE void st_from_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// 8 contiguous elements <-> fp32 registers with 16-byte (bf16) / 2 x 16-byte (fp32) accesses
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static HIG_DEVICE void load(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static HIG_DEVICE void load_rw(const float* p, float (&v)[8]) {  // coherent load (buffer also written by this kernel)
    const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static HIG_DEVICE void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static HIG_DEVICE void unpack(const uint4 u, float (&v)[8]) {
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
  }
  static HIG_DEVICE void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(__ldg(reinterpret_cast<const uint4*>(p)), v); }
  static HIG_DEVICE void load_rw(const __nv_bfloat16* p, float (&v)[8]) { unpack(*reinterpret_cast<const uint4*>(p), v); }
  static HIG_DEVICE void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// ------------------------------------------------------------------------------------------------
// transpose: in [M,N] (ld_in) -> outT [N,M] (ld_t), optional straight copy in the output type (ld_c) and optional
// column sums (fp32, atomically accumulated: the caller zeroes them).  64x64 tiles through padded shared memory,
// coalesced on both sides.  rows_keep_mod > 0: rows with (m % rows_keep_mod) == 0 are treated as zero (frame 0 of
// every sequence belongs to the out2 head, :613-616).
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
transpose_kernel(const TIn* __restrict__ in, int M, int N, int ld_in, TOut* __restrict__ outT, int ld_t,
                 TOut* __restrict__ copy, int ld_c, float* __restrict__ colsum, int rows_zero_mod) {
  __shared__ float tile[64][65];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  // load: thread -> (row = tid / 16 + 16 i, 4 consecutive columns)
  {
    const int c4 = (tid & 15) * 4, r = tid >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = r + 16 * i, m = m0 + rl;
      const bool zero_row = rows_zero_mod > 0 && (m % rows_zero_mod) == 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + c4 + j;
        float v = 0.f;
        if (m < M && n < N && !zero_row) v = ld_as_f(in + (size_t)m * ld_in + n);
        tile[rl][c4 + j] = v;
        if (copy && m < M && n < N) st_from_f(copy + (size_t)m * ld_c + n, v);
      }
    }
  }
  __syncthreads();
  // store: thread -> (output row = input column nl = tid / 8 + 32 p, 8 consecutive input rows)
  const int seg = (tid & 7) * 8;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int nl = (tid >> 3) + 32 * p, n = n0 + nl;
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = tile[seg + j][nl];
      part += v;
      const int m = m0 + seg + j;
      if (outT && n < N && m < M) st_from_f(outT + (size_t)n * ld_t + m, v);
    }
    if (colsum) {
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if ((tid & 7) == 0 && n < N) atomicAdd(colsum + n, part);
    }
  }
}

// cast (+ column sums) without a transposed output: the fp32 residual-stream gradient -> bf16 GEMM operand, and its column
// sums = the bias gradient of the block's output projection.  A thread owns 8 adjacent columns (16 / 32-byte accesses),
// threads that share a column group take interleaved rows and meet in shared memory; one atomicAdd per column and CTA.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
cast_colsum_kernel(const TIn* __restrict__ in, int M, int N, int ld_in, TOut* __restrict__ copy, int ld_c,
                   float* __restrict__ colsum, int rows_per_chunk) {
  __shared__ float part[256 * 8];
  const int groups = N >> 3;
  int gpc = 256;
  while (gpc > groups) gpc >>= 1;
  const int rpt = 256 / gpc;
  const int gi = threadIdx.x % gpc, ri = threadIdx.x / gpc;
  const int grp = blockIdx.x * gpc + gi;
  const int m0 = blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (grp < groups) {
#pragma unroll 4
    for (int m = m0 + ri; m < m1; m += rpt) {
      float v[8];
      Vec8<TIn>::load(in + (size_t)m * ld_in + grp * 8, v);
      if (copy) Vec8<TOut>::store(copy + (size_t)m * ld_c + grp * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < gpc * 8; c += 256) {
    const int g2 = c >> 3, j = c & 7;
    if (blockIdx.x * gpc + g2 >= groups) continue;
    float sum = 0.f;
    for (int k = 0; k < rpt; ++k) sum += part[(k * gpc + g2) * 8 + j];
    atomicAdd(colsum + (blockIdx.x * gpc + g2) * 8 + j, sum);
  }
}

int transpose(const void* in, int in_dtype, int M, int N, int ld_in, void* outT, int ld_t, void* copy, int ld_c,
              int out_dtype, float* colsum, int rows_zero_mod, cudaStream_t stream) {
  if (!in || M <= 0 || N <= 0 || (!outT && !copy && !colsum)) return set_error(HIG_ERR_INVALID, "transpose: bad arguments");
  if (!outT && rows_zero_mod == 0 && (N % 8) == 0 && (ld_in % 8) == 0 && (!copy || (ld_c % 8) == 0) &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(copy)) & 15) == 0 &&
      ((in_dtype == HIG_F32 && out_dtype == HIG_BF16) || (in_dtype == HIG_BF16 && out_dtype == HIG_BF16))) {
    const int groups = N / 8;
    int gpc = 256;
    while (gpc > groups) gpc >>= 1;
    const int xb = (groups + gpc - 1) / gpc, rpt = 256 / gpc;
    int chunks = (148 * 4 + xb - 1) / xb;
    const int max_chunks = (M + rpt * 4 - 1) / (rpt * 4);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1 || (colsum != nullptr && deterministic())) chunks = 1;
    const int rpc = (M + chunks - 1) / chunks;
    dim3 grid(xb, (M + rpc - 1) / rpc);
    using bf = __nv_bfloat16;
    if (in_dtype == HIG_F32)
      cast_colsum_kernel<float, bf><<<grid, 256, 0, stream>>>((const float*)in, M, N, ld_in, (bf*)copy, ld_c, colsum, rpc);
    else
      cast_colsum_kernel<bf, bf><<<grid, 256, 0, stream>>>((const bf*)in, M, N, ld_in, (bf*)copy, ld_c, colsum, rpc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("cast_colsum launch: ") + cudaGetErrorString(e));
    count_launch();
    return HIG_OK;
  }
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  using bf = __nv_bfloat16;
#define HIG_TR(TI, TO) \
  transpose_kernel<TI, TO><<<grid, 256, 0, stream>>>((const TI*)in, M, N, ld_in, (TO*)outT, ld_t, (TO*)copy, ld_c, colsum, rows_zero_mod)
  if (in_dtype == HIG_F32 && out_dtype == HIG_F32) HIG_TR(float, float);
  else if (in_dtype == HIG_F32 && out_dtype == HIG_BF16) HIG_TR(float, bf);
  else if (in_dtype == HIG_BF16 && out_dtype == HIG_BF16) HIG_TR(bf, bf);
  else if (in_dtype == HIG_BF16 && out_dtype == HIG_F32) HIG_TR(bf, float);
  else return set_error(HIG_ERR_INVALID, "transpose: bad dtype");
#undef HIG_TR
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("transpose launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// colsum: out[n] (+)= sum_m in[m, n].  Threads own columns (coalesced), blockIdx.y owns a chunk of rows; the
// partial sums are combined with fp32 atomics, so the caller zeroes `out` (accumulate semantics).
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_kernel(const TIn* __restrict__ in, int M, int N, int ld, int rows_per_chunk, float* __restrict__ out) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += ld_as_f(in + (size_t)m * ld + n);
  atomicAdd(out + n, acc);
}

// N % 8 == 0, ld % 8 == 0, 16-byte aligned base: a thread owns 8 adjacent columns (16 / 32-byte loads), the threads of a CTA
// that share a column group take interleaved rows and meet in shared memory; one atomicAdd per column and CTA.
template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const TIn* __restrict__ in, int M, int N, int ld, int rows_per_chunk, float* __restrict__ out) {
  __shared__ float part[256 * 8];
  const int groups = N >> 3;                       // column groups of 8
  int gpc = 256;                                   // groups per CTA: the largest power of two <= min(groups, 256)
  while (gpc > groups) gpc >>= 1;
  const int rpt = 256 / gpc;                       // row lanes per group inside the CTA
  const int gi = threadIdx.x % gpc, ri = threadIdx.x / gpc;
  const int grp = blockIdx.x * gpc + gi;
  const int m0 = blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (grp < groups && ri < rpt) {
#pragma unroll 4
    for (int m = m0 + ri; m < m1; m += rpt) {     // unrolled: four independent 16 / 32-byte loads in flight per thread
      float v[8];
      Vec8<TIn>::load(in + (size_t)m * ld + grp * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  // thread t < gpc * 8 owns column (t / 8 group, t % 8)
  const int t = threadIdx.x;
  for (int c = t; c < gpc * 8; c += 256) {
    const int g2 = c >> 3, j = c & 7;
    if (blockIdx.x * gpc + g2 >= groups) continue;
    float sum = 0.f;
    for (int k = 0; k < rpt; ++k) sum += part[(k * gpc + g2) * 8 + j];
    atomicAdd(out + (blockIdx.x * gpc + g2) * 8 + j, sum);
  }
}

int colsum(const void* in, int dtype, int M, int N, int ld, float* out, cudaStream_t stream) {
  if (!in || !out || M <= 0 || N <= 0) return set_error(HIG_ERR_INVALID, "colsum: bad arguments");
  if (dtype != HIG_BF16 && dtype != HIG_F32) return set_error(HIG_ERR_INVALID, "colsum: bad dtype");
  const bool vec = (N % 8) == 0 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (vec) {
    const int groups = N / 8;
    int gpc = 256;
    while (gpc > groups) gpc >>= 1;
    {
      const int xb = (groups + gpc - 1) / gpc;
      const int rpt = 256 / gpc;
      int chunks = (148 * 4 + xb - 1) / xb;
      const int max_chunks = (M + rpt * 4 - 1) / (rpt * 4);    // at least ~4 rows per thread
      if (chunks > max_chunks) chunks = max_chunks;
      if (chunks < 1 || deterministic()) chunks = 1;          // one CTA per column group: a single atomicAdd per column
      const int rpc = (M + chunks - 1) / chunks;
      dim3 grid(xb, (M + rpc - 1) / rpc);
      if (dtype == HIG_BF16) colsum_vec_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, M, N, ld, rpc, out);
      else colsum_vec_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, M, N, ld, rpc, out);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("colsum launch: ") + cudaGetErrorString(e));
      count_launch();
      return HIG_OK;
    }
  }
  const int xb = (N + 255) / 256;
  int chunks = (148 * 8 + xb - 1) / xb;
  if (chunks > M) chunks = M;
  if (chunks < 1 || deterministic()) chunks = 1;
  const int rpc = (M + chunks - 1) / chunks;
  dim3 grid(xb, (M + rpc - 1) / rpc);
  if (dtype == HIG_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, M, N, ld, rpc, out);
  else colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, M, N, ld, rpc, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("colsum launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// elementwise activation forward / backward (exact erf GELU :257, SiLU :476/:77); also a plain cast when act == 0.
//   fwd: out = act(x)            bwd: dx = dy * act'(x)   (x = the saved pre-activation)
// ------------------------------------------------------------------------------------------------
// 8 elements per thread and iteration (16 / 32-byte accesses); `n8` full groups, the < 8-element tail goes scalar
template <typename TIn, typename TOut>
__global__ void act_fwd_kernel(const TIn* __restrict__ x, long long n, int act, TOut* __restrict__ out, int vec) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n8 = n >> 3;
    for (long long g = i; g < n8; g += stride) {
      float v[8];
      Vec8<TIn>::load(x + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_fwd_f(v[j], act);
      Vec8<TOut>::store(out + g * 8, v);
    }
    for (long long k = (n8 << 3) + i; k < n; k += stride) st_from_f(out + k, act_fwd_f(ld_as_f(x + k), act));
    return;
  }
  for (; i < n; i += stride) st_from_f(out + i, act_fwd_f(ld_as_f(x + i), act));
}
template <typename TX, typename TG, typename TOut>
__global__ void act_bwd_kernel(const TX* __restrict__ x, const TG* __restrict__ dy, long long n, int act,
                               TOut* __restrict__ dx, int vec) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n8 = n >> 3;
    for (long long g = i; g < n8; g += stride) {
      float v[8], d[8];
      Vec8<TX>::load(x + g * 8, v);
      Vec8<TG>::load(dy + g * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] *= act_grad_f(v[j], act);
      Vec8<TOut>::store(dx + g * 8, d);
    }
    for (long long k = (n8 << 3) + i; k < n; k += stride) st_from_f(dx + k, ld_as_f(dy + k) * act_grad_f(ld_as_f(x + k), act));
    return;
  }
  for (; i < n; i += stride) st_from_f(dx + i, ld_as_f(dy + i) * act_grad_f(ld_as_f(x + i), act));
}

static int ew_blocks(long long n) {
  long long b = (n / 8 + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

int act_fwd(const void* x, int x_dtype, long long n, int act, void* out, int out_dtype, cudaStream_t stream) {
  if (!x || !out || n <= 0 || act < 0 || act > 3) return set_error(HIG_ERR_INVALID, "act_fwd: bad arguments");
  using bf = __nv_bfloat16;
  const int b = ew_blocks(n);
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (x_dtype == HIG_F32 && out_dtype == HIG_F32) act_fwd_kernel<float, float><<<b, 256, 0, stream>>>((const float*)x, n, act, (float*)out, vec);
  else if (x_dtype == HIG_F32 && out_dtype == HIG_BF16) act_fwd_kernel<float, bf><<<b, 256, 0, stream>>>((const float*)x, n, act, (bf*)out, vec);
  else if (x_dtype == HIG_BF16 && out_dtype == HIG_BF16) act_fwd_kernel<bf, bf><<<b, 256, 0, stream>>>((const bf*)x, n, act, (bf*)out, vec);
  else if (x_dtype == HIG_BF16 && out_dtype == HIG_F32) act_fwd_kernel<bf, float><<<b, 256, 0, stream>>>((const bf*)x, n, act, (float*)out, vec);
  else return set_error(HIG_ERR_INVALID, "act_fwd: bad dtype");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("act_fwd launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

int act_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, long long n, int act, void* dx, int dx_dtype,
            cudaStream_t stream) {
  if (!x || !dy || !dx || n <= 0 || act < 0 || act > 2) return set_error(HIG_ERR_INVALID, "act_bwd: bad arguments");
  using bf = __nv_bfloat16;
  const int b = ew_blocks(n);
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
#define HIG_AB(TX, TG, TO) act_bwd_kernel<TX, TG, TO><<<b, 256, 0, stream>>>((const TX*)x, (const TG*)dy, n, act, (TO*)dx, vec)
  if (x_dtype == HIG_BF16 && dy_dtype == HIG_BF16 && dx_dtype == HIG_BF16) HIG_AB(bf, bf, bf);
  else if (x_dtype == HIG_F32 && dy_dtype == HIG_F32 && dx_dtype == HIG_F32) HIG_AB(float, float, float);
  else if (x_dtype == HIG_BF16 && dy_dtype == HIG_F32 && dx_dtype == HIG_BF16) HIG_AB(bf, float, bf);
  else if (x_dtype == HIG_BF16 && dy_dtype == HIG_F32 && dx_dtype == HIG_F32) HIG_AB(bf, float, float);
  else if (x_dtype == HIG_F32 && dy_dtype == HIG_BF16 && dx_dtype == HIG_F32) HIG_AB(float, bf, float);
  else if (x_dtype == HIG_F32 && dy_dtype == HIG_BF16 && dx_dtype == HIG_BF16) HIG_AB(float, bf, bf);
  else return set_error(HIG_ERR_INVALID, "act_bwd: unsupported dtype combination");
#undef HIG_AB
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("act_bwd launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of ln_film_silu:   u = n_hat * gamma + beta,  t = u * (1 + scale_s) + shift_s,  out = [SiLU](t)
//   dt      = dout * SiLU'(t)
//   dscale  = sum_rows(dt * u)      dshift = sum_rows(dt)            (per sequence)
//   du      = dt * (1 + scale)
//   dgamma  = sum_rows(du * n_hat)  dbeta  = sum_rows(du)            (per-sequence partials; the caller column-sums)
//   dx      = rstd * (dn - mean(dn) - n_hat * mean(dn * n_hat)),  dn = du * gamma
// One CTA per (sequence, row slice); a warp walks rows, a lane owns 8 contiguous columns per 256-wide chunk exactly
// as the forward kernel does; the per-lane column partials are combined through shared-memory atomics and flushed
// with global fp32 atomics (partial buffers are zeroed by the caller).
// ------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 8;

// raw 8-element vector of the storage type: issued one row ahead of its use (software prefetch)
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  HIG_DEVICE void load(const float* p) { a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1); }
  HIG_DEVICE void get(float (&v)[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 u;
  HIG_DEVICE void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
  HIG_DEVICE void get(float (&v)[8]) const { Vec8<__nv_bfloat16>::unpack(u, v); }
};

// Parameters (gamma, beta, 1 + scale, shift) sit in shared memory (not 64 registers), the per-column partial sums of a
// warp are combined warp by warp through shared memory (no shared-memory atomics), transcendental math is the fast
// intrinsic kind, and the next row's loads are in flight while a row is processed: 2 CTAs per SM instead of 1.
// (Round 1: 95-105 us per launch at [23296, 512] against ~11 us of HBM traffic — 22 % of the training step.)
template <int WIDTH, typename TX, typename TG, typename TDX>
__global__ void __launch_bounds__(LNB_WARPS * 32, 2)
ln_film_silu_bwd_kernel(const TX* __restrict__ x, int rows, int rows_per_seq, int slices, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ scale_shift, int ss_stride,
                        int apply_silu, const TG* __restrict__ dout, TDX* __restrict__ dx, int dx_accumulate,
                        float* __restrict__ d_ss, int dss_stride, float* __restrict__ d_gb, int dgb_stride) {
  constexpr int CH = WIDTH / 256;
  // per-column vectors in a lane-major layout: column c*256 + lane*8 + j lives at ((c*2 + j/4)*32 + lane)*4 + j%4, so a
  // warp's 16-byte accesses are contiguous (the natural layout put the lanes 32 B apart: 2-way bank conflicts, 1.7 M per launch)
  __shared__ __align__(16) float par[4][WIDTH];   // gamma, beta, 1 + scale, shift
  __shared__ __align__(16) float red[2][WIDTH];   // X = sum dt n_hat, Y = sum dt
  auto perm = [](int col) { const int c = col >> 8, l = (col & 255) >> 3, j = col & 7; return ((c * 2 + (j >> 2)) * 32 + l) * 4 + (j & 3); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x / slices, slice = blockIdx.x - seq * slices;
  for (int i = threadIdx.x; i < WIDTH; i += LNB_WARPS * 32) {
    const int p = perm(i);
    par[0][p] = gamma[i];
    par[1][p] = beta[i];
    par[2][p] = scale_shift ? 1.0f + scale_shift[(size_t)seq * ss_stride + i] : 1.0f;
    par[3][p] = scale_shift ? scale_shift[(size_t)seq * ss_stride + WIDTH + i] : 0.f;
    red[0][p] = red[1][p] = 0.f;
  }
  __syncthreads();

  // Two accumulators per column instead of four: with  X = sum_rows dt * n_hat  and  Y = sum_rows dt  (dt = dout * SiLU'),
  //   dshift = Y,  dscale = gamma X + beta Y,  dbeta = (1 + scale) Y,  dgamma = (1 + scale) X
  // because gamma, beta and (1 + scale) are constant over the rows a CTA (one sequence) walks.
  float aX[CH][8], aY[CH][8];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) aX[c][j] = aY[c][j] = 0.f;

  const bool need_u = apply_silu != 0 || d_ss != nullptr;   // u = n_hat gamma + beta feeds SiLU' and d(scale)
  const int r_begin = seq * rows_per_seq;
  const int r_end = min(rows, r_begin + rows_per_seq);
  const int r_step = slices * LNB_WARPS;
  int r = r_begin + slice * LNB_WARPS + warp;
  Raw8<TX> nx[CH];
  Raw8<TG> ng[CH];
  if (r < r_end) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const size_t idx = (size_t)r * WIDTH + c * 256 + lane * 8;
      nx[c].load(x + idx);
      ng[c].load(dout + idx);
    }
  }
  for (; r < r_end; r += r_step) {
    float v[CH][8], g[CH][8];
    float s = 0.f, ssq = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      nx[c].get(v[c]);
      ng[c].get(g[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[c][j];
        ssq = fmaf(v[c][j], v[c][j], ssq);
      }
    }
    if (r + r_step < r_end) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const size_t idx = (size_t)(r + r_step) * WIDTH + c * 256 + lane * 8;
        nx[c].load(x + idx);
        ng[c].load(dout + idx);
      }
    }
    // the gradient this row accumulates into (dx_accumulate): requested now, consumed after the second reduction round
    float prev[CH][8];
    if (dx_accumulate) {
#pragma unroll
      for (int c = 0; c < CH; ++c) Vec8<TDX>::load_rw(dx + (size_t)r * WIDTH + c * 256 + lane * 8, prev[c]);
    }
    // one reduction round for both moments (two independent shuffle chains); var = E[x^2] - mean^2 in fp32, as the forward's
    // LayerNorm-folded projections take it
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    }
    const float mean = s * (1.0f / WIDTH);
    const float rstd = rsqrtf(fmaxf(fmaf(ssq, 1.0f / WIDTH, -mean * mean), 0.f) + 1e-5f);
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] -= mean;
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int p0 = (c * 64 + lane) * 4, p1 = p0 + 128;   // lane-major positions of columns j = 0..3 / 4..7
      float G[8], Bt[8], S1[8], SH[8];
      *reinterpret_cast<float4*>(G) = *reinterpret_cast<const float4*>(&par[0][p0]);
      *reinterpret_cast<float4*>(G + 4) = *reinterpret_cast<const float4*>(&par[0][p1]);
      *reinterpret_cast<float4*>(S1) = *reinterpret_cast<const float4*>(&par[2][p0]);
      *reinterpret_cast<float4*>(S1 + 4) = *reinterpret_cast<const float4*>(&par[2][p1]);
      if (need_u) {
        *reinterpret_cast<float4*>(Bt) = *reinterpret_cast<const float4*>(&par[1][p0]);
        *reinterpret_cast<float4*>(Bt + 4) = *reinterpret_cast<const float4*>(&par[1][p1]);
        *reinterpret_cast<float4*>(SH) = *reinterpret_cast<const float4*>(&par[3][p0]);
        *reinterpret_cast<float4*>(SH + 4) = *reinterpret_cast<const float4*>(&par[3][p1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float nh = v[c][j] * rstd;
        float dt = g[c][j];
        float u = 0.f;
        if (need_u) {
          u = fmaf(nh, G[j], Bt[j]);
          if (apply_silu) {
            const float t = fmaf(u, S1[j], SH[j]);
            const float sg = __fdividef(1.0f, 1.0f + __expf(-t));
            dt *= sg * fmaf(t, 1.0f - sg, 1.0f);
          }
        }
        aX[c][j] = fmaf(dt, nh, aX[c][j]);
        aY[c][j] += dt;
        const float dn = dt * S1[j] * G[j];
        m1 += dn;
        m2 = fmaf(dn, nh, m2);
        v[c][j] = nh;
        g[c][j] = dn;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.0f / WIDTH);
    m2 *= (1.0f / WIDTH);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const size_t idx = (size_t)r * WIDTH + c * 256 + lane * 8;
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = rstd * (g[c][j] - m1 - v[c][j] * m2);
        if (dx_accumulate) d[j] += prev[c][j];
      }
      Vec8<TDX>::store(dx + idx, d);
    }
  }
  // column partials: warp after warp adds its registers into the shared accumulators (16-byte accesses, no atomics)
  for (int w = 0; w < LNB_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int p0 = (c * 64 + lane) * 4;
        auto add8 = [&](float* dst, const float (&a)[8]) {
          float4 p = *reinterpret_cast<float4*>(dst), q = *reinterpret_cast<float4*>(dst + 128);
          p.x += a[0]; p.y += a[1]; p.z += a[2]; p.w += a[3]; q.x += a[4]; q.y += a[5]; q.z += a[6]; q.w += a[7];
          *reinterpret_cast<float4*>(dst) = p;
          *reinterpret_cast<float4*>(dst + 128) = q;
        };
        add8(&red[0][p0], aX[c]);
        add8(&red[1][p0], aY[c]);
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < WIDTH; i += LNB_WARPS * 32) {
    const int p = perm(i);
    const float X = red[0][p], Y = red[1][p];
    if (d_ss) {
      atomicAdd(d_ss + (size_t)seq * dss_stride + i, fmaf(par[0][p], X, par[1][p] * Y));
      atomicAdd(d_ss + (size_t)seq * dss_stride + WIDTH + i, Y);
    }
    if (d_gb) {
      atomicAdd(d_gb + (size_t)seq * dgb_stride + i, par[2][p] * X);
      atomicAdd(d_gb + (size_t)seq * dgb_stride + WIDTH + i, par[2][p] * Y);
    }
  }
}

int ln_film_silu_bwd(const void* x, int x_dtype, int rows, int width, int rows_per_seq, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int apply_silu, const void* dout,
                     int dout_dtype, void* dx, int dx_dtype, int dx_accumulate, float* d_ss, int dss_stride,
                     float* d_gb, int dgb_stride, cudaStream_t stream) {
  if (!x || !dout || !dx || !gamma || !beta || rows <= 0 || rows_per_seq <= 0)
    return set_error(HIG_ERR_INVALID, "ln_film_silu_bwd: bad arguments");
  if (width != 512 && width != 256) return set_error(HIG_ERR_UNSUPPORTED, "ln_film_silu_bwd: width must be 256 or 512");
  if (scale_shift && ((ss_stride % 4) || (reinterpret_cast<uintptr_t>(scale_shift) & 15)))
    return set_error(HIG_ERR_INVALID, "ln_film_silu_bwd: scale_shift must be 16-byte aligned with ss_stride % 4 == 0");
  const int n_seq = (rows + rows_per_seq - 1) / rows_per_seq;
  // one wave of two resident CTAs per SM: split a sequence's rows over `slices` CTAs only when there are few sequences
  // (256 sequences x 2 slices = 512 CTAs ran as 1.73 waves: 58 us; 256 x 1 fits one wave)
  int slices = (148 * 2) / n_seq;
  const int max_slices = (rows_per_seq + LNB_WARPS - 1) / LNB_WARPS;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1 || deterministic()) slices = 1;      // one CTA per sequence: every (scale | shift) gradient has one contributor
  const int blocks = n_seq * slices;
  using bf = __nv_bfloat16;
#define HIG_LNB(W, TX, TG, TD)                                                                                        \
  ln_film_silu_bwd_kernel<W, TX, TG, TD><<<blocks, LNB_WARPS * 32, 0, stream>>>(                                      \
      (const TX*)x, rows, rows_per_seq, slices, gamma, beta, scale_shift, ss_stride, apply_silu, (const TG*)dout,     \
      (TD*)dx, dx_accumulate, d_ss, dss_stride, d_gb, dgb_stride)
#define HIG_LNB_W(W)                                                                                                  \
  if (x_dtype == HIG_F32 && dout_dtype == HIG_BF16 && dx_dtype == HIG_F32) HIG_LNB(W, float, bf, float);             \
  else if (x_dtype == HIG_BF16 && dout_dtype == HIG_BF16 && dx_dtype == HIG_BF16) HIG_LNB(W, bf, bf, bf);            \
  else if (x_dtype == HIG_F32 && dout_dtype == HIG_F32 && dx_dtype == HIG_F32) HIG_LNB(W, float, float, float);      \
  else if (x_dtype == HIG_F32 && dout_dtype == HIG_BF16 && dx_dtype == HIG_BF16) HIG_LNB(W, float, bf, bf);          \
  else return set_error(HIG_ERR_UNSUPPORTED, "ln_film_silu_bwd: dtype combination")
  if (width == 512) { HIG_LNB_W(512); } else { HIG_LNB_W(256); }
#undef HIG_LNB_W
#undef HIG_LNB
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("ln_film_silu_bwd launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
