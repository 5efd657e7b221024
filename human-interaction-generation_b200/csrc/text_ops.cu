// Text-conditioning path kernels: softmax multi-head attention for the two text transformers of
// MotionInteractionTransformer.encode_text (codes/models/interaction_transformer.py:533-559) — the CLIP text transformer
// (12 layers, 8 heads, causal mask; nn.MultiheadAttention inside CLIP's ResidualAttentionBlock) and the 4-layer
// nn.TransformerEncoder (:446-455, 4 heads, no mask) — at 77 tokens per caption.  The projections / MLPs run on the library's
// GEMM kernels and the LayerNorms on hig_ln_film_silu; this file adds the one operation those do not cover.
//
// out[b, i, h, :] = sum_j softmax_j( q[b,i,h,:] . k[b,j,h,:] / sqrt(64) [+ causal mask] ) v[b,j,h,:]
// One CTA per (caption, head): K and V of the head in shared memory (fp32, padded rows), a warp per query row: lanes own
// keys for the scores (77 tokens -> 3 keys per lane) and output features for the weighted sum.  The whole text path is
// < 0.5 % of a sampling call and runs once per batch, so this kernel is written for clarity, not for a roofline.
#include <string>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

template <typename T> HIG_DEVICE float tx_ld(const T* p);
template <> HIG_DEVICE float tx_ld<float>(const float* p) { return *p; }
template <> HIG_DEVICE float tx_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> HIG_DEVICE void tx_st(T* p, float v);
template <> HIG_DEVICE void tx_st<float>(float* p, float v) { *p = v; }
template <> HIG_DEVICE void tx_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

constexpr int MHA_HD = 64, MHA_LD = 65, MHA_WARPS = 8, MHA_MAXN = 128;

template <typename T>
__global__ void __launch_bounds__(MHA_WARPS * 32)
mha_attention_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ld, T* __restrict__ out,
                     int ldo, int N, int causal) {
  extern __shared__ float mha_smem[];
  float* sK = mha_smem;                 // [N][65]
  float* sV = sK + N * MHA_LD;          // [N][65]
  float* sQ = sV + N * MHA_LD;          // [warps][64]
  float* sP = sQ + MHA_WARPS * MHA_HD;  // [warps][MHA_MAXN]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t base = (size_t)b * N * ld + h * MHA_HD;
  for (int i = tid; i < N * MHA_HD; i += MHA_WARPS * 32) {
    const int r = i >> 6, c = i & 63;
    sK[r * MHA_LD + c] = tx_ld(k + base + (size_t)r * ld + c);
    sV[r * MHA_LD + c] = tx_ld(v + base + (size_t)r * ld + c);
  }
  __syncthreads();
  float* myq = sQ + warp * MHA_HD;
  float* myp = sP + warp * MHA_MAXN;
  for (int i = warp; i < N; i += MHA_WARPS) {
    myq[lane] = tx_ld(q + base + (size_t)i * ld + lane) * 0.125f;
    myq[lane + 32] = tx_ld(q + base + (size_t)i * ld + lane + 32) * 0.125f;
    __syncwarp();
    const int nk = causal ? i + 1 : N;
    float sc[MHA_MAXN / 32];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MHA_MAXN / 32; ++jj) {
      const int j = lane + 32 * jj;
      float d = -INFINITY;
      if (j < nk) {
        d = 0.f;
#pragma unroll 16
        for (int c = 0; c < MHA_HD; ++c) d = fmaf(myq[c], sK[j * MHA_LD + c], d);
      }
      sc[jj] = d;
      m = fmaxf(m, d);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < MHA_MAXN / 32; ++jj) {
      const int j = lane + 32 * jj;
      const float e = j < nk ? expf(sc[jj] - m) : 0.f;
      if (j < N) myp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nk; ++j) {
      const float p = myp[j];
      o0 = fmaf(p, sV[j * MHA_LD + lane], o0);
      o1 = fmaf(p, sV[j * MHA_LD + lane + 32], o1);
    }
    T* og = out + ((size_t)b * N + i) * ldo + h * MHA_HD;
    tx_st(og + lane, o0 * inv);
    tx_st(og + lane + 32, o1 * inv);
    __syncwarp();
  }
}

int mha_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int B, int N, int H, int causal,
                  int dtype, cudaStream_t stream) {
  if (!q || !k || !v || !out || B <= 0 || N <= 0 || H <= 0) return set_error(HIG_ERR_INVALID, "mha_attention: bad arguments");
  if (N > MHA_MAXN) return set_error(HIG_ERR_UNSUPPORTED, "mha_attention: at most 128 tokens (the text context is 77)");
  const size_t smem = ((size_t)2 * N * MHA_LD + MHA_WARPS * (MHA_HD + MHA_MAXN)) * sizeof(float);
  dim3 grid(H, B);
  cudaError_t e;
  if (dtype == HIG_BF16) {
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(mha_attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("mha_attention attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    using bf = __nv_bfloat16;
    mha_attention_kernel<bf><<<grid, MHA_WARPS * 32, smem, stream>>>((const bf*)q, (const bf*)k, (const bf*)v, ld, (bf*)out, ldo, N, causal);
  } else if (dtype == HIG_F32) {
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(mha_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("mha_attention attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    mha_attention_kernel<float><<<grid, MHA_WARPS * 32, smem, stream>>>((const float*)q, (const float*)k, (const float*)v, ld,
                                                                       (float*)out, ldo, N, causal);
  } else {
    return set_error(HIG_ERR_INVALID, "mha_attention: bad dtype");
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("mha_attention launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
