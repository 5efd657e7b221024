// Training-step kernels around the denoiser's forward / backward (all HBM-bound, fp32):
//
//   hig_masked_mse        the reference's training loss and its gradient in three launches — DDPMMulTrainer.backward_G,
//                         codes/trainers/mul_ddpm_trainer.py:223-247: frame 0 scores its first 4 features, frames >= 1
//                         all C, mean over features, weighted by src_mask, normalised by the mask sum; the unlabelled
//                         (PIT) variant sums the two persons of an assignment and keeps the cheaper of the two caption
//                         assignments per pair (:235-242).  Replaces ~10 eager kernels + torch.autograd's backward of them.
//   hig_sumsq             sum of squares of a flat fp32 buffer (global gradient norm of clip_grad_norm_, :253).
//   hig_adam_flat         clip-by-global-norm + Adam (torch.optim.Adam semantics, :291 / :254-255) over flat fp32
//                         parameter / gradient / moment buffers in ONE pass that also refreshes the bf16 operand mirror the
//                         tcgen05 GEMMs read — the optimizer step, the gradient clipping multiply and the per-iteration
//                         fp32 -> bf16 weight casts are a single sweep over the 107 M parameters.
#include <string>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

// ---------------------------------------------------------------------------------------------- loss
// rows[s] = sum_t mask[s,t] * mean_c (pred - tgt)^2     one CTA per sequence, a warp per frame
__global__ void __launch_bounds__(256)
mse_rows_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int* __restrict__ length, int S, int T,
                int C, float* __restrict__ rows) {
  __shared__ float part[8];
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int len = length ? length[s] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  float acc = 0.f;
  for (int t = warp; t < len; t += 8) {
    const size_t off = ((size_t)s * T + t) * C;
    const int nc = t == 0 ? 4 : C;
    float e = 0.f;
    for (int c = lane; c < nc; c += 32) {
      const float d = pred[off + c] - tgt[off + c];
      e = fmaf(d, d, e);
    }
    acc += warp_sum(e) / (float)nc;
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += part[i];
    rows[s] = r;
  }
}

// loss + per-sequence gradient weight.  labelled: loss = sum_s rows[s] / M, w[s] = 1 / M, M = sum_s len_s.
// PIT (S = 4B: (m1,c1) (m1,c2) (m2,c2) (m2,c1)): p[j] = rows[j] + rows[j + 2B] for j < 2B (the two persons of assignment j),
// loss = sum_{i<B} min(p[i], p[i+B]) / (M / 2); sequences of the losing assignment get w = 0.
__global__ void __launch_bounds__(256)
mse_finalize_kernel(const float* __restrict__ rows, const int* __restrict__ length, int S, int T, int pit,
                    float* __restrict__ w, float* __restrict__ loss) {
  __shared__ float red[256];
  float m = 0.f;
  for (int s = threadIdx.x; s < S; s += 256) {
    int len = length ? length[s] : T;
    m += (float)(len < 0 ? 0 : (len > T ? T : len));
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float M = red[0];
  __syncthreads();
  float acc = 0.f;
  if (!pit) {
    const float inv = M > 0.f ? 1.0f / M : 0.f;
    for (int s = threadIdx.x; s < S; s += 256) {
      w[s] = inv;
      acc += rows[s];
    }
    acc *= inv;
  } else {
    const int B = S / 4;
    const float inv = M > 0.f ? 2.0f / M : 0.f;
    for (int i = threadIdx.x; i < B; i += 256) {
      const float p0 = rows[i] + rows[i + 2 * B], p1 = rows[i + B] + rows[i + 3 * B];
      const bool first = p0 <= p1;          // torch.min keeps the first index on a tie
      acc += (first ? p0 : p1) * inv;
      w[i] = w[i + 2 * B] = first ? inv : 0.f;
      w[i + B] = w[i + 3 * B] = first ? 0.f : inv;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0];
}

// d_pred[s,t,c] = 2 (pred - tgt) mask[s,t] w[s] / C_t   (0 for the features frame 0 does not score)
__global__ void __launch_bounds__(256)
mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int* __restrict__ length,
                const float* __restrict__ w, int S, int T, int C, float* __restrict__ d_pred) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)S * T) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(row / T), t = (int)(row - (long long)s * T);
  int len = length ? length[s] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  const int nc = t == 0 ? 4 : C;
  const float k = t < len ? 2.0f * w[s] / (float)nc : 0.f;
  const size_t off = (size_t)row * C;
  for (int c = lane; c < C; c += 32) d_pred[off + c] = c < nc ? k * (pred[off + c] - tgt[off + c]) : 0.f;
}

int masked_mse(const float* pred, const float* tgt, const int* length, int S, int T, int C, int pit, float* rows, float* w,
               float* loss, float* d_pred, cudaStream_t stream) {
  if (!pred || !tgt || !rows || !w || !loss || S <= 0 || T <= 0 || C < 4)
    return set_error(HIG_ERR_INVALID, "masked_mse: bad arguments");
  if (pit && (S % 4)) return set_error(HIG_ERR_INVALID, "masked_mse: the PIT batch stacks 4 x B sequences");
  mse_rows_kernel<<<S, 256, 0, stream>>>(pred, tgt, length, S, T, C, rows);
  mse_finalize_kernel<<<1, 256, 0, stream>>>(rows, length, S, T, pit, w, loss);
  int launches = 2;
  if (d_pred) {
    const long long nrows = (long long)S * T;
    mse_grad_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, stream>>>(pred, tgt, length, w, S, T, C, d_pred);
    ++launches;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("masked_mse launch: ") + cudaGetErrorString(e));
  for (int i = 0; i < launches; ++i) count_launch();
  return HIG_OK;
}

// ---------------------------------------------------------------------------------------------- gradient norm
__global__ void __launch_bounds__(512)
sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  __shared__ float part[16];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < 16 ? part[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) atomicAdd(out, (double)r);
  }
}

// out (a double the caller zeroed) += sum x^2;  x must be 16-byte aligned
int sumsq(const float* x, long long n, double* out, cudaStream_t stream) {
  if (!x || !out || n <= 0) return set_error(HIG_ERR_INVALID, "sumsq: bad arguments");
  if (reinterpret_cast<uintptr_t>(x) & 15) return set_error(HIG_ERR_INVALID, "sumsq: x must be 16-byte aligned");
  long long blocks = (n / 4 + 511) / 512;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1 || deterministic()) blocks = 1;      // HIG_DETERMINISTIC: one block, fixed summation order
  sumsq_kernel<<<(unsigned)blocks, 512, 0, stream>>>(x, n, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("sumsq launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ---------------------------------------------------------------------------------------------- slice mean
// own[i] = (own[i] + sum_j staged[j * stride + i]) * scale — the reduction step of the peer-memory gradient exchange
// (ddp.PeerGradExchange): `count` contributions pulled from the other ranks by the copy engines.  A few blocks only: it runs
// beside the backward kernels and must not take the machine from them.
__global__ void __launch_bounds__(256)
mean_slices_kernel(float* __restrict__ own, const float* __restrict__ staged, long long n, long long stride, int count,
                   float scale, int vec) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += nth) {
      float4 a = reinterpret_cast<float4*>(own)[i];
      for (int j = 0; j < count; ++j) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(staged + (long long)j * stride) + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
      reinterpret_cast<float4*>(own)[i] = a;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nth) {
      float a = own[i];
      for (int j = 0; j < count; ++j) a += staged[(long long)j * stride + i];
      own[i] = a * scale;
    }
    return;
  }
  for (long long i = tid; i < n; i += nth) {
    float a = own[i];
    for (int j = 0; j < count; ++j) a += staged[(long long)j * stride + i];
    own[i] = a * scale;
  }
}

int mean_slices(float* own, const float* staged, long long n, long long stride, int count, float scale, cudaStream_t stream) {
  if (!own || !staged || n <= 0 || count < 0) return set_error(HIG_ERR_INVALID, "mean_slices: bad arguments");
  const int vec = ((reinterpret_cast<uintptr_t>(own) | reinterpret_cast<uintptr_t>(staged)) & 15) == 0 && (stride % 4) == 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 32) blocks = 32;
  if (blocks < 1) blocks = 1;
  mean_slices_kernel<<<(unsigned)blocks, 256, 0, stream>>>(own, staged, n, stride, count, scale, vec);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("mean_slices launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// ---------------------------------------------------------------------------------------------- Adam
// torch.optim.Adam (no weight decay, no amsgrad):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)          bc_i = 1 - b_i^step
// g is first scaled by clip = min(1, max_norm / (sqrt(gnorm2) + 1e-6)) (torch.nn.utils.clip_grad_norm_) when gnorm2 != null.
__global__ void __launch_bounds__(512)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __nv_bfloat16* __restrict__ pb, long long n, float step_size, float b1, float b2, float inv_sqrt_bc2,
                 float eps, const double* __restrict__ gnorm2, float max_norm) {
  float clip = 1.0f;
  if (gnorm2 != nullptr) {
    const float nrm = (float)sqrt(*gnorm2);
    clip = fminf(1.0f, max_norm / (nrm + 1e-6f));
  }
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp);
    const float* G = reinterpret_cast<const float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm);
    float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = G[j] * clip;
      M[j] = fmaf(b1, M[j], (1.0f - b1) * gj);
      V[j] = fmaf(b2, V[j], (1.0f - b2) * gj * gj);
      P[j] -= step_size * M[j] / (sqrtf(V[j]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb != nullptr)
      reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16x2(P[0], P[1]), pack_bf16x2(P[2], P[3]));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float gj = g[i] * clip;
    const float mj = fmaf(b1, m[i], (1.0f - b1) * gj), vj = fmaf(b2, v[i], (1.0f - b2) * gj * gj);
    const float pj = p[i] - step_size * mj / (sqrtf(vj) * inv_sqrt_bc2 + eps);
    p[i] = pj; m[i] = mj; v[i] = vj;
    if (pb != nullptr) pb[i] = __float2bfloat16(pj);
  }
}

int adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
              float eps, int step, const double* gnorm2, float max_norm, cudaStream_t stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) return set_error(HIG_ERR_INVALID, "adam_flat: bad arguments");
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(m) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(p_bf16) & 7))
    return set_error(HIG_ERR_INVALID, "adam_flat: buffers must be 16-byte aligned (bf16 mirror: 8)");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = (n / 4 + 511) / 512;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_flat_kernel<<<(unsigned)blocks, 512, 0, stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n,
                                                        (float)(lr / bc1), beta1, beta2, (float)(1.0 / sqrt(bc2)), eps,
                                                        gnorm2, max_norm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("adam_flat launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
