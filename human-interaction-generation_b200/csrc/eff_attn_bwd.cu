// hig_eff_attn_bwd: backward of the fused efficient attention (eff_attn.cu), one CTA per (sequence, head).
//
// Forward (reference: LinearTemporalSelfAttention / CrossAttention / InteractionCrossAttention.forward,
// codes/models/interaction_transformer.py:112-130, 145-165, 181-207; differentiated there by torch.autograd):
//   Qs = softmax_feat(Q)   Ks = softmax_time(K, masked)   A = Ks^T V   Y = Qs A
// Backward, given dY (nothing but Q, K, V is saved by the forward — Qs, Ks, A are recomputed here in shared memory):
//   dA  = Qs^T dY                      dQs = dY A^T        dQ = Qs * (dQs - rowsum(dQs * Qs))
//   dV  = Ks dA                        dKs = V dA^T        dK = Ks * (dKs - colsum(dKs * Ks))
// Masked key rows have Ks == 0 exactly, so dK = dV = 0 there (the -1e6 additive mask of the reference passes no
// gradient either: its softmax output is exactly 0 in fp32).
//
// modes: 0 SELF     q,k,v of sequence s                      -> dq, dk, dv rows of s
//        1 INTER    q of s; k,v of the partner (s+shift)%S   -> dq rows of s; dk, dv rows of the partner
//        2 KV_ONLY  k,v, dA (fp32 [S,H,64,64], input)        -> dk, dv                (text K/V side)
//        3 Q_ONLY   q, a_in, dY                              -> dq, dA (fp32, output)  (text query side)
// All arithmetic is fp32 on CUDA cores; storage type (bf16 / fp32) is a template parameter.  Shared memory holds
// Ks and V for the whole sequence, A and dA, and a 32-row chunk of (Qs, dY) at a time.
#include <cstdlib>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

int eff_attn_bwd_tc(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                    const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                    const int* length, int S, int T, int H, int pair_shift, float* q_sum, float* k_sum, float* v_sum,
                    cudaStream_t stream);

constexpr int BW_HD = 64;
constexpr int BW_LD = 65;       // padded fp32 row
constexpr int BW_THREADS = 256;
constexpr int BW_CHUNK = 32;    // query rows per chunk (8 warps x 4 rows)

template <typename T> HIG_DEVICE float bw_ld(const T* p);
template <> HIG_DEVICE float bw_ld<float>(const float* p) { return *p; }
template <> HIG_DEVICE float bw_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> HIG_DEVICE void bw_st(T* p, float v);
template <> HIG_DEVICE void bw_st<float>(float* p, float v) { *p = v; }
template <> HIG_DEVICE void bw_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(BW_THREADS)
eff_attn_bwd_kernel(int mode, const T* __restrict__ q, int ldq, const T* __restrict__ k, const T* __restrict__ v,
                    int ldkv, const T* __restrict__ a_in, const T* __restrict__ dy, int lddy, T* __restrict__ dq,
                    int lddq, T* __restrict__ dk, T* __restrict__ dv, int lddkv, float* __restrict__ dA_g,
                    const int* __restrict__ length, int S, int T_, int pair_shift) {
  extern __shared__ __align__(16) uint8_t bw_smem[];
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  const int Tkv = do_kv ? T_ : 0;
  float* sK = reinterpret_cast<float*>(bw_smem);  // [Tkv][65]  Ks (normalised), later dK scratch
  float* sV = sK + Tkv * BW_LD;                    // [Tkv][65]  V, later dKs
  float* sA = sV + Tkv * BW_LD;                    // [64][65]
  float* sdA = sA + BW_HD * BW_LD;                 // [64][65]
  float* sQ = sdA + BW_HD * BW_LD;                 // [32][65]  Qs chunk
  float* sdY = sQ + BW_CHUNK * BW_LD;              // [32][65]  dY chunk
  float* sred = sdY + BW_CHUNK * BW_LD;            // [4][64]

  const int h = blockIdx.x, s = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s_kv = (mode == 1) ? (s + pair_shift) % S : s;
  int len = T_;
  if (length && (mode == 0 || mode == 1)) {
    len = length[s];
    len = len < 0 ? 0 : (len > T_ ? T_ : len);
  }
  const int c = tid & 63, part = tid >> 6;  // column / 4-way row partition used by the column reductions

  // ---------------- P1: Ks, V, A ----------------
  if (do_kv) {
    const T* kg = k + (size_t)s_kv * T_ * ldkv + h * BW_HD;
    const T* vg = v + (size_t)s_kv * T_ * ldkv + h * BW_HD;
    for (int i = tid; i < T_ * BW_HD; i += BW_THREADS) {
      const int r = i >> 6, cc = i & 63;
      sK[r * BW_LD + cc] = bw_ld(kg + (size_t)r * ldkv + cc);
      sV[r * BW_LD + cc] = (r < len) ? bw_ld(vg + (size_t)r * ldkv + cc) : 0.f;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int t = part; t < len; t += 4) m = fmaxf(m, sK[t * BW_LD + c]);
    sred[part * 64 + c] = m;
    __syncthreads();
    m = fmaxf(fmaxf(sred[c], sred[64 + c]), fmaxf(sred[128 + c], sred[192 + c]));
    __syncthreads();
    float sum = 0.f;
    for (int t = part; t < len; t += 4) {
      const float e = expf(sK[t * BW_LD + c] - m);
      sK[t * BW_LD + c] = e;
      sum += e;
    }
    sred[part * 64 + c] = sum;
    __syncthreads();
    const float tot = sred[c] + sred[64 + c] + sred[128 + c] + sred[192 + c];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    for (int t = part; t < T_; t += 4) sK[t * BW_LD + c] = (t < len) ? sK[t * BW_LD + c] * inv : 0.f;
    __syncthreads();
  }
  const int d4 = tid >> 2, l0 = (tid & 3) * 16;  // thread -> A / dA element block [d4][l0 .. l0+16)
  if (do_q) {
    if (do_kv) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      for (int t = 0; t < len; ++t) {
        const float kk = sK[t * BW_LD + d4];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(kk, sV[t * BW_LD + l0 + j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) sA[d4 * BW_LD + l0 + j] = acc[j];
    } else {
      const T* ag = a_in + ((size_t)s * H + h) * BW_HD * BW_HD;
      for (int i = tid; i < BW_HD * BW_HD; i += BW_THREADS) sA[(i >> 6) * BW_LD + (i & 63)] = bw_ld(ag + i);
    }
    __syncthreads();

    // ---------------- P2: per 32-row chunk: Qs, dQ; dA accumulates in registers ----------------
    float dacc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dacc[j] = 0.f;
    for (int t0 = 0; t0 < T_; t0 += BW_CHUNK) {
#pragma unroll
      for (int rr = 0; rr < BW_CHUNK / 8; ++rr) {
        const int rl = warp * (BW_CHUNK / 8) + rr, t = t0 + rl;
        if (t < T_) {
          const T* qg = q + ((size_t)s * T_ + t) * ldq + h * BW_HD;
          const T* gy = dy + ((size_t)s * T_ + t) * lddy + h * BW_HD;
          const float a = bw_ld(qg + lane), b = bw_ld(qg + lane + 32);
          const float mx = warp_max(fmaxf(a, b));
          const float ea = expf(a - mx), eb = expf(b - mx);
          const float inv = 1.0f / warp_sum(ea + eb);
          const float qa = ea * inv, qb = eb * inv;
          sQ[rl * BW_LD + lane] = qa;
          sQ[rl * BW_LD + lane + 32] = qb;
          sdY[rl * BW_LD + lane] = bw_ld(gy + lane);
          sdY[rl * BW_LD + lane + 32] = bw_ld(gy + lane + 32);
          __syncwarp();
          float da = 0.f, db = 0.f;  // dQs[d] = sum_l dY[t,l] A[d,l]
#pragma unroll 8
          for (int l = 0; l < BW_HD; ++l) {
            const float g = sdY[rl * BW_LD + l];
            da = fmaf(g, sA[lane * BW_LD + l], da);
            db = fmaf(g, sA[(lane + 32) * BW_LD + l], db);
          }
          const float dot = warp_sum(da * qa + db * qb);
          T* dqg = dq + ((size_t)s * T_ + t) * lddq + h * BW_HD;
          bw_st(dqg + lane, qa * (da - dot));
          bw_st(dqg + lane + 32, qb * (db - dot));
        } else {
          sQ[rl * BW_LD + lane] = 0.f;
          sQ[rl * BW_LD + lane + 32] = 0.f;
          sdY[rl * BW_LD + lane] = 0.f;
          sdY[rl * BW_LD + lane + 32] = 0.f;
        }
      }
      __syncthreads();
#pragma unroll 4
      for (int rl = 0; rl < BW_CHUNK; ++rl) {
        const float qs = sQ[rl * BW_LD + d4];
#pragma unroll
        for (int j = 0; j < 16; ++j) dacc[j] = fmaf(qs, sdY[rl * BW_LD + l0 + j], dacc[j]);
      }
      __syncthreads();
    }
    if (mode == 3) {
      float* og = dA_g + ((size_t)s * H + h) * BW_HD * BW_HD;
#pragma unroll
      for (int j = 0; j < 16; ++j) og[d4 * BW_HD + l0 + j] = dacc[j];
      return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) sdA[d4 * BW_LD + l0 + j] = dacc[j];
  } else {
    const float* ig = dA_g + ((size_t)s * H + h) * BW_HD * BW_HD;
    for (int i = tid; i < BW_HD * BW_HD; i += BW_THREADS) sdA[(i >> 6) * BW_LD + (i & 63)] = ig[i];
  }
  __syncthreads();

  // ---------------- P3: dV = Ks dA (to global), dKs = V dA^T (in place over V), dK ----------------
  T* dkg = dk + (size_t)s_kv * T_ * lddkv + h * BW_HD;
  T* dvg = dv + (size_t)s_kv * T_ * lddkv + h * BW_HD;
  for (int t = warp; t < T_; t += BW_THREADS / 32) {
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    if (t < len) {
#pragma unroll 8
      for (int j = 0; j < BW_HD; ++j) {
        const float ks = sK[t * BW_LD + j];     // dV[t,l] = sum_d Ks[t,d] dA[d,l]   (l = lane, lane+32)
        v0 = fmaf(ks, sdA[j * BW_LD + lane], v0);
        v1 = fmaf(ks, sdA[j * BW_LD + lane + 32], v1);
        const float vv = sV[t * BW_LD + j];     // dKs[t,d] = sum_l V[t,l] dA[d,l]   (d = lane, lane+32)
        k0 = fmaf(vv, sdA[lane * BW_LD + j], k0);
        k1 = fmaf(vv, sdA[(lane + 32) * BW_LD + j], k1);
      }
    }
    bw_st(dvg + (size_t)t * lddkv + lane, v0);
    bw_st(dvg + (size_t)t * lddkv + lane + 32, v1);
    __syncwarp();
    sV[t * BW_LD + lane] = k0;
    sV[t * BW_LD + lane + 32] = k1;
  }
  __syncthreads();
  float cs = 0.f;
  for (int t = part; t < len; t += 4) cs = fmaf(sV[t * BW_LD + c], sK[t * BW_LD + c], cs);
  sred[part * 64 + c] = cs;
  __syncthreads();
  cs = sred[c] + sred[64 + c] + sred[128 + c] + sred[192 + c];
  for (int t = part; t < T_; t += 4) {
    const float ks = sK[t * BW_LD + c];
    bw_st(dkg + (size_t)t * lddkv + c, ks * (sV[t * BW_LD + c] - cs));
  }
}

// bias-gradient column sums of dQ / dK / dV as separate passes (the kernels that do not fuse them)
static int attn_bwd_sums(bool do_q, bool do_kv, const void* dq, int lddq, const void* dk, const void* dv, int lddkv, int rows,
                         int width, int dtype, float* q_sum, float* k_sum, float* v_sum, cudaStream_t stream) {
  int rc = HIG_OK;
  if (do_q && q_sum) rc = colsum(dq, dtype, rows, width, lddq, q_sum, stream);
  if (rc == HIG_OK && do_kv && k_sum) rc = colsum(dk, dtype, rows, width, lddkv, k_sum, stream);
  if (rc == HIG_OK && do_kv && v_sum) rc = colsum(dv, dtype, rows, width, lddkv, v_sum, stream);
  return rc;
}

int eff_attn_bwd(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                 const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                 const int* length, int S, int T, int H, int pair_shift, int dtype, float* q_sum, float* k_sum,
                 float* v_sum, cudaStream_t stream) {
  if (mode < 0 || mode > 3) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: bad mode");
  if (S <= 0 || T <= 0 || H <= 0) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: empty shape");
  if (T > 256) return set_error(HIG_ERR_UNSUPPORTED, "eff_attn_bwd: T > 256 not supported");
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  if (do_kv && (!k || !v || !dk || !dv)) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: K/V/dK/dV required");
  if (do_q && (!q || !dy || !dq)) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: Q/dY/dQ required");
  if ((mode == 2 || mode == 3) && !dA) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: dA required");
  if (mode == 3 && !a_in) return set_error(HIG_ERR_INVALID, "eff_attn_bwd: a_in required");
  const size_t smem = ((size_t)(do_kv ? 2 * T : 0) * BW_LD + 2 * BW_HD * BW_LD + 2 * BW_CHUNK * BW_LD + 4 * 64) * sizeof(float);
  dim3 grid(H, S);
  cudaError_t e;
  if (dtype == HIG_BF16) {
    // tensor-core kernel (eff_attn_bwd_tc.cu); HIG_ATTN_BWD_TC=0 keeps the CUDA-core kernel below for A/B runs
    static const bool use_tc = []() { const char* ev = getenv("HIG_ATTN_BWD_TC"); return !(ev && ev[0] == '0'); }();
    if (use_tc) {
      // (bit-reproducible mode: the fused column sums are atomics in CTA order — take them with hig::colsum below instead)
      const bool fuse = !deterministic();
      const int rc = eff_attn_bwd_tc(mode, q, ldq, k, v, ldkv, a_in, dy, lddy, dq, lddq, dk, dv, lddkv, dA, length, S, T,
                                     H, pair_shift, fuse ? q_sum : nullptr, fuse ? k_sum : nullptr, fuse ? v_sum : nullptr,
                                     stream);
      if (rc == HIG_OK && !fuse) return attn_bwd_sums(do_q, do_kv, dq, lddq, dk, dv, lddkv, S * T, H * BW_HD, dtype, q_sum, k_sum, v_sum, stream);
      if (rc != HIG_ERR_UNSUPPORTED) return rc;
    }
  }
  if (dtype == HIG_BF16) {
    using bf = __nv_bfloat16;
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(eff_attn_bwd_kernel<bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn_bwd attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    eff_attn_bwd_kernel<bf><<<grid, BW_THREADS, smem, stream>>>(
        mode, (const bf*)q, ldq, (const bf*)k, (const bf*)v, ldkv, (const bf*)a_in, (const bf*)dy, lddy, (bf*)dq, lddq,
        (bf*)dk, (bf*)dv, lddkv, dA, length, S, T, pair_shift);
  } else if (dtype == HIG_F32) {
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(eff_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn_bwd attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    eff_attn_bwd_kernel<float><<<grid, BW_THREADS, smem, stream>>>(
        mode, (const float*)q, ldq, (const float*)k, (const float*)v, ldkv, (const float*)a_in, (const float*)dy, lddy,
        (float*)dq, lddq, (float*)dk, (float*)dv, lddkv, dA, length, S, T, pair_shift);
  } else {
    return set_error(HIG_ERR_INVALID, "eff_attn_bwd: bad dtype");
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn_bwd launch: ") + cudaGetErrorString(e));
  count_launch();
  return attn_bwd_sums(do_q, do_kv, dq, lddq, dk, dv, lddkv, S * T, H * BW_HD, dtype, q_sum, k_sum, v_sum, stream);
}

}  // namespace hig
