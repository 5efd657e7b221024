// hig_gemm_f32: the fp32-mode projection  C[M,N] = act( A[M,K] · W[N,K]^T + bias + residual ).
//
// This is the "fp32 mode" of the north star (per-step parity <= 1e-5 against the reference's fp32 PyTorch):
// plain FFMA accumulation in fp32, no tensor cores (tcgen05 kind::tf32 keeps 10 mantissa bits and cannot hold
// 1e-5).  It exists to pin the algorithm of every other kernel bit-tightly; the product path is gemm_tcgen05.cu.
// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread, smem-staged and transposed so inner loads are
// conflict-free float4 reads.
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, int M, int N, int K,
                const float* __restrict__ bias, const float* __restrict__ residual, int ldr, int res_row_mod,
                float* __restrict__ out, int ldo, int act) {
  __shared__ float sA[SG_BK][SG_BM + 4];
  __shared__ float sW[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // each thread loads 4 elements of A and 4 of W: row = tid/4 (0..63), kk = (tid%4)*4 .. +3
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + kk + u;
        const int m = m0 + r, n = n0 + r;
        sA[kk + u][r] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
        sW[kk + u][r] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int rr = res_row_mod > 0 ? (m % res_row_mod) : m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (residual) v += residual[(size_t)rr * ldr + n];
      if (act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      else if (act == 2) v = v / (1.0f + expf(-v));
      out[(size_t)m * ldo + n] = v;
    }
  }
}

int gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias,
             const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, int act,
             cudaStream_t stream) {
  if (!A || !W || !out_f32 || M <= 0 || N <= 0 || K <= 0) return set_error(HIG_ERR_INVALID, "gemm_f32: bad arguments");
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(A, lda, W, ldw, M, N, K, bias, residual, ldr, res_row_mod, out_f32,
                                            ldo_f32, act);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("gemm_f32 launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
