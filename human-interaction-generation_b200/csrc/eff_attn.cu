// hig_eff_attn: fused "efficient attention" for one (sequence, head) per CTA, no HBM round trip for the
// intermediates.
//
//   Qs = softmax(Q, over the 64 head features)            [T,64]
//   Ks = softmax(K + (1-mask)*(-1e6), over time)           [Tkv,64]   (masked rows come out exactly 0 in fp32)
//   A  = Ks^T · (V * mask)                                 [64,64]
//   Y  = Qs · A                                            [T,64]
//
// Reference semantics: LinearTemporalSelfAttention.forward            interaction_transformer.py:112-130
//                      LinearTemporalCrossAttention.forward           :145-165   (K,V from text, no mask)
//                      LinearTemporalInteractionCrossAttention.forward :181-207  (K,V from the partner, masked with
//                          the *query-side* sequence's length, V not multiplied by the mask — equal because Ks==0)
//
// modes: 0 SELF   Q,K,V of sequence s, length[s]
//        1 INTER  Q of s; K,V of (s + pair_shift) % S; mask from length[s]
//        2 KV_ONLY  K,V -> A written to a_out [S,H,64,64]   (text cross-attention precompute, step invariant)
//        3 Q_ONLY   Q, a_in -> Y                            (text cross-attention apply)
//
// bf16 path: tiles are staged with cp.async into padded shared memory (row stride 144 B => conflict-free
// ldmatrix), the column (time) softmax runs on CUDA cores in fp32, both contractions run on mma.sync
// m16n8k16 bf16 with fp32 accumulate, the feature softmax of Q is done on the MMA A-fragments in registers.
// fp32 path ("fp32 mode"): the same algorithm with FFMA contractions and expf, used to pin the algorithm to 1e-5.
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int HD = 64;           // head dim (latent_dim / num_heads = 512 / 8)
constexpr int ATT_STRIDE = 72;   // bf16 elements per smem row (144 B)
constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;

HIG_DEVICE void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
HIG_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> HIG_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

HIG_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage a [rows,64] bf16 tile (global row stride ld) into padded smem with 16-byte cp.async
HIG_DEVICE void stage_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int ld, int rows) {
  for (int i = threadIdx.x; i < rows * 8; i += ATT_THREADS) {
    const int r = i >> 3, c = (i & 7) * 8;
    cp_async16(s + r * ATT_STRIDE + c, g + (size_t)r * ld + c);
  }
}
HIG_DEVICE void zero_rows(__nv_bfloat16* s, int r0, int r1) {
  for (int i = threadIdx.x + r0 * 8; i < r1 * 8; i += ATT_THREADS) {
    const int r = i >> 3, c = (i & 7) * 8;
    *reinterpret_cast<uint4*>(s + r * ATT_STRIDE + c) = make_uint4(0, 0, 0, 0);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
eff_attn_bf16_kernel(int mode, const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, int ldkv, const __nv_bfloat16* __restrict__ a_in,
                     __nv_bfloat16* __restrict__ a_out, __nv_bfloat16* __restrict__ y, int ldy,
                     const int* __restrict__ length, int S, int T, int pair_shift, int mask_v) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int TP = (T + 15) & ~15;
  // Q_ONLY (text apply) only stages Q and A: a third of the shared memory, so five CTAs fit per SM instead of two
  // KV_ONLY stages K and V only (three CTAs per SM)
  const int kv_rows = (mode == 3) ? 0 : TP;
  const int q_rows = (mode == 2) ? 0 : TP;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + kv_rows * ATT_STRIDE;
  __nv_bfloat16* sQ = sV + kv_rows * ATT_STRIDE;
  __nv_bfloat16* sA = sQ + q_rows * ATT_STRIDE;
  float* sred = reinterpret_cast<float*>(sA + HD * ATT_STRIDE);  // [ATT_WARPS][64] partials, then [64] inverse sums
  float* sinv = sred + ATT_WARPS * 64;

  const int h = blockIdx.x, s = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  // INTER, and KV_ONLY with a pair shift (the K/V half of the inter-person attention): partner's K,V, own length
  const int s_kv = (mode == 1 || mode == 2) ? (s + pair_shift) % S : s;
  int len = T;
  if (length && mode != 3) {
    len = length[s];
    len = len < 0 ? 0 : (len > T ? T : len);
  }

  // ---------------- async staging ----------------
  if (do_kv) {
    stage_tile(sK, k + (size_t)s_kv * T * ldkv + h * HD, ldkv, T);
    stage_tile(sV, v + (size_t)s_kv * T * ldkv + h * HD, ldkv, T);
  } else {
    stage_tile(sA, a_in + ((size_t)s * H + h) * HD * HD, HD, HD);  // A tile [64,64] from global (bf16, dense)
  }
  cp_async_commit();
  if (do_q) stage_tile(sQ, q + (size_t)s * T * ldq + h * HD, ldq, T);
  cp_async_commit();
  if (do_kv) { zero_rows(sK, T, TP); zero_rows(sV, T, TP); }
  if (do_q) zero_rows(sQ, T, TP);
  cp_async_wait<1>();
  __syncthreads();

  if (do_kv) {
    // ---------------- phase 1: time softmax of K over rows [0,len); lane owns the column pair (2*lane, 2*lane+1),
    // warp w owns rows w, w+8, ...  Unnormalised exp() is written back in place; 1/sum is folded into A. ----------
    uint32_t* sK32 = reinterpret_cast<uint32_t*>(sK);
    constexpr int ROW32 = ATT_STRIDE / 2;
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int t = warp; t < len; t += ATT_WARPS) {
      const float2 kv2 = unpack_bf16x2(sK32[t * ROW32 + lane]);
      m0 = fmaxf(m0, kv2.x);
      m1 = fmaxf(m1, kv2.y);
    }
    sred[warp * 64 + 2 * lane] = m0;
    sred[warp * 64 + 2 * lane + 1] = m1;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) {
      m0 = fmaxf(m0, sred[w * 64 + 2 * lane]);
      m1 = fmaxf(m1, sred[w * 64 + 2 * lane + 1]);
    }
    __syncthreads();
    float s0 = 0.f, s1 = 0.f;
    for (int t = warp; t < T; t += ATT_WARPS) {
      uint32_t packed = 0u;
      if (t < len) {
        const float2 kv2 = unpack_bf16x2(sK32[t * ROW32 + lane]);
        // sum what the MMA will actually see (the bf16-rounded weights) so the normalisation is exact
        const __nv_bfloat162 e2 = __floats2bfloat162_rn(__expf(kv2.x - m0), __expf(kv2.y - m1));
        const float2 ef = __bfloat1622float2(e2);
        s0 += ef.x;
        s1 += ef.y;
        packed = *reinterpret_cast<const uint32_t*>(&e2);
      }
      sK32[t * ROW32 + lane] = packed;
    }
    if (mask_v) {
      uint32_t* sV32 = reinterpret_cast<uint32_t*>(sV);
      for (int t = len + warp; t < T; t += ATT_WARPS) sV32[t * ROW32 + lane] = 0u;
    }
    sred[warp * 64 + 2 * lane] = s0;
    sred[warp * 64 + 2 * lane + 1] = s1;
    __syncthreads();
    if (tid < 64) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < ATT_WARPS; ++w) tot += sred[w * 64 + tid];
      sinv[tid] = tot > 0.f ? 1.0f / tot : 0.f;
    }
    __syncthreads();

    // ---------------- phase 2: A[d,l] = (sum_t e[t,d] V[t,l]) / sum_t e[t,d]
    // warp w owns d in [16*(w&3), +16) and l in [32*(w>>2), +32) ----------------
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int dw = (warp & 3) * 16, lw = (warp >> 2) * 32;
    const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
    for (int kt = 0; kt < TP; kt += 16) {
      uint32_t a[4];
      {
        const int row = kt + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int col = dw + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(sK_u + (row * ATT_STRIDE + col) * 2, a[0], a[1], a[2], a[3]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kt + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = lw + np * 16 + ((lane >> 4) & 1) * 8;
        ldsm_x4_t(sV_u + (row * ATT_STRIDE + col) * 2, b0, b1, b2, b3);
        mma_bf16_16816(acc[2 * np], a, b0, b1);
        mma_bf16_16816(acc[2 * np + 1], a, b2, b3);
      }
    }
    const int g = lane >> 2, tg = lane & 3;
    const float i0 = sinv[dw + g], i1 = sinv[dw + g + 8];
    __nv_bfloat16* dst = (mode == 2) ? a_out + ((size_t)s * H + h) * HD * HD : sA;
    const int dstride = (mode == 2) ? HD : ATT_STRIDE;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = lw + nt * 8 + 2 * tg;
      *reinterpret_cast<uint32_t*>(dst + (dw + g) * dstride + col) = pack_bf16x2(acc[nt][0] * i0, acc[nt][1] * i0);
      *reinterpret_cast<uint32_t*>(dst + (dw + g + 8) * dstride + col) = pack_bf16x2(acc[nt][2] * i1, acc[nt][3] * i1);
    }
    if (mode == 2) return;
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---------------- phase 3: Y = softmax_feat(Q) · A, one 16-row tile per warp iteration ----------------
  const uint32_t sQ_u = smem_u32(sQ), sA_u = smem_u32(sA);
  uint32_t bfrag[4][8][2];  // [k-step][n-tile][2]
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int col = np * 16 + ((lane >> 4) & 1) * 8;
      ldsm_x4_t(sA_u + (row * ATT_STRIDE + col) * 2, bfrag[kk][2 * np][0], bfrag[kk][2 * np][1],
                bfrag[kk][2 * np + 1][0], bfrag[kk][2 * np + 1][1]);
    }
  const int g = lane >> 2, tg = lane & 3;
  for (int mt = warp; mt * 16 < T; mt += ATT_WARPS) {
    uint32_t af[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int col = kk * 16 + ((lane >> 4) & 1) * 8;
      ldsm_x4(sQ_u + (row * ATT_STRIDE + col) * 2, af[kk][0], af[kk][1], af[kk][2], af[kk][3]);
    }
    // feature softmax: row g uses regs {0,2} of each k-step, row g+8 uses regs {1,3}
    float x0[16], x1[16];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float2 f;
      f = unpack_bf16x2(af[kk][0]); x0[4 * kk + 0] = f.x; x0[4 * kk + 1] = f.y;
      f = unpack_bf16x2(af[kk][2]); x0[4 * kk + 2] = f.x; x0[4 * kk + 3] = f.y;
      f = unpack_bf16x2(af[kk][1]); x1[4 * kk + 0] = f.x; x1[4 * kk + 1] = f.y;
      f = unpack_bf16x2(af[kk][3]); x1[4 * kk + 2] = f.x; x1[4 * kk + 3] = f.y;
    }
    float m0 = x0[0], m1 = x1[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) { m0 = fmaxf(m0, x0[j]); m1 = fmaxf(m1, x1[j]); }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x0[j] = __expf(x0[j] - m0); s0 += x0[j];
      x1[j] = __expf(x1[j] - m1); s1 += x1[j];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      af[kk][0] = pack_bf16x2(x0[4 * kk + 0] * i0, x0[4 * kk + 1] * i0);
      af[kk][2] = pack_bf16x2(x0[4 * kk + 2] * i0, x0[4 * kk + 3] * i0);
      af[kk][1] = pack_bf16x2(x1[4 * kk + 0] * i1, x1[4 * kk + 1] * i1);
      af[kk][3] = pack_bf16x2(x1[4 * kk + 2] * i1, x1[4 * kk + 3] * i1);
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_bf16_16816(acc[nt], af[kk], bfrag[kk][nt][0], bfrag[kk][nt][1]);

    // stage the 16x64 bf16 result through this warp's own (already consumed) Q rows, then 16-byte row stores
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * tg;
      *reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g) * ATT_STRIDE + col) = pack_bf16x2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g + 8) * ATT_STRIDE + col) = pack_bf16x2(acc[nt][2], acc[nt][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int r = mt * 16 + (idx >> 3), cc = (idx & 7) * 8;
      if (r < T) {
        const uint4 val = *reinterpret_cast<const uint4*>(sQ + r * ATT_STRIDE + cc);
        *reinterpret_cast<uint4*>(y + ((size_t)s * T + r) * ldy + h * HD + cc) = val;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 mode
// ------------------------------------------------------------------------------------------------
constexpr int F32_STRIDE = 65;
constexpr int F32_THREADS = 256;

__global__ void __launch_bounds__(F32_THREADS)
eff_attn_f32_kernel(int mode, const float* __restrict__ q, int ldq, const float* __restrict__ k,
                    const float* __restrict__ v, int ldkv, const float* __restrict__ a_in, float* __restrict__ a_out,
                    float* __restrict__ y, int ldy, const int* __restrict__ length, int S, int T, int pair_shift,
                    int mask_v) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  float* sK = reinterpret_cast<float*>(att_smem);   // [T][65]
  float* sV = sK + T * F32_STRIDE;                   // [T][65]
  float* sA = sV + T * F32_STRIDE;                   // [64][65]
  float* sred = sA + HD * F32_STRIDE;                // [4][64] partial max / sum
  float* sq = sred + 4 * 64;                         // [8 warps][64] softmaxed q row

  const int h = blockIdx.x, s = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  const int s_kv = (mode == 1 || mode == 2) ? (s + pair_shift) % S : s;
  int len = T;
  if (length && mode != 3) {
    len = length[s];
    len = len < 0 ? 0 : (len > T ? T : len);
  }

  if (do_kv) {
    const float* kg = k + (size_t)s_kv * T * ldkv + h * HD;
    const float* vg = v + (size_t)s_kv * T * ldkv + h * HD;
    for (int i = tid; i < T * HD; i += F32_THREADS) {
      const int r = i >> 6, c = i & 63;
      sK[r * F32_STRIDE + c] = kg[(size_t)r * ldkv + c];
      float vv = vg[(size_t)r * ldkv + c];
      if (mask_v && r >= len) vv = 0.f;
      sV[r * F32_STRIDE + c] = vv;
    }
    __syncthreads();
    const int c = tid & 63, part = tid >> 6;  // 4 row partitions
    float m = -INFINITY;
    for (int t = part; t < len; t += 4) m = fmaxf(m, sK[t * F32_STRIDE + c]);
    sred[part * 64 + c] = m;
    __syncthreads();
    m = fmaxf(fmaxf(sred[c], sred[64 + c]), fmaxf(sred[128 + c], sred[192 + c]));
    __syncthreads();
    float sum = 0.f;
    for (int t = part; t < len; t += 4) {
      const float e = expf(sK[t * F32_STRIDE + c] - m);
      sK[t * F32_STRIDE + c] = e;
      sum += e;
    }
    sred[part * 64 + c] = sum;
    __syncthreads();
    const float inv = 1.0f / (sred[c] + sred[64 + c] + sred[128 + c] + sred[192 + c]);
    for (int t = part; t < T; t += 4) sK[t * F32_STRIDE + c] = (t < len) ? sK[t * F32_STRIDE + c] * inv : 0.f;
    __syncthreads();
    // A[d][l]: thread -> d = tid/4, l in [16*(tid%4), +16)
    const int d = tid >> 2, l0 = (tid & 3) * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float kk = sK[t * F32_STRIDE + d];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(kk, sV[t * F32_STRIDE + l0 + j], acc[j]);
    }
    if (mode == 2) {
      float* ao = a_out + ((size_t)s * H + h) * HD * HD;
#pragma unroll
      for (int j = 0; j < 16; ++j) ao[d * HD + l0 + j] = acc[j];
      return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) sA[d * F32_STRIDE + l0 + j] = acc[j];
  } else {
    const float* ag = a_in + ((size_t)s * H + h) * HD * HD;
    for (int i = tid; i < HD * HD; i += F32_THREADS) sA[(i >> 6) * F32_STRIDE + (i & 63)] = ag[i];
  }
  __syncthreads();
  if (!do_q) return;
  // one warp per query row
  float* myq = sq + warp * 64;
  for (int t = warp; t < T; t += F32_THREADS / 32) {
    const float* qg = q + ((size_t)s * T + t) * ldq + h * HD;
    const float a = qg[lane], b = qg[lane + 32];
    const float m = warp_max(fmaxf(a, b));
    const float ea = expf(a - m), eb = expf(b - m);
    const float inv = 1.0f / warp_sum(ea + eb);
    myq[lane] = ea * inv;
    myq[lane + 32] = eb * inv;
    __syncwarp();
    float y0 = 0.f, y1 = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float qs = myq[d];
      y0 = fmaf(qs, sA[d * F32_STRIDE + lane], y0);
      y1 = fmaf(qs, sA[d * F32_STRIDE + lane + 32], y1);
    }
    float* yg = y + ((size_t)s * T + t) * ldy + h * HD;
    yg[lane] = y0;
    yg[lane + 32] = y1;
    __syncwarp();
  }
}

int eff_attn(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in, void* a_out,
             void* y, int ldy, const int* length, int S, int T, int H, int pair_shift, int mask_v, int dtype,
             cudaStream_t stream) {
  if (mode < 0 || mode > 3) return set_error(HIG_ERR_INVALID, "eff_attn: bad mode");
  if (S <= 0 || T <= 0 || H <= 0) return set_error(HIG_ERR_INVALID, "eff_attn: empty shape");
  if (T > 256) return set_error(HIG_ERR_UNSUPPORTED, "eff_attn: T > 256 not supported (reference max is 196 frames)");
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  if (do_kv && (!k || !v)) return set_error(HIG_ERR_INVALID, "eff_attn: K/V required");
  if (do_q && (!q || !y)) return set_error(HIG_ERR_INVALID, "eff_attn: Q/Y required");
  if (mode == 2 && !a_out) return set_error(HIG_ERR_INVALID, "eff_attn: a_out required");
  if (mode == 3 && !a_in) return set_error(HIG_ERR_INVALID, "eff_attn: a_in required");
  dim3 grid(H, S);
  cudaError_t e;
  // bf16 K/V half: dedicated kernel (4 CTAs per SM).  The mask makes V*mask and unmasked V equivalent (Ks == 0).
  if (dtype == HIG_BF16 && mode == 2 && (ldkv % 8) == 0 && !(reinterpret_cast<uintptr_t>(k) & 15) &&
      !(reinterpret_cast<uintptr_t>(v) & 15))
    return attn_kv(k, v, ldkv, a_out, length, S, T, H, pair_shift, 0, stream);
  if (dtype == HIG_BF16) {
    if ((do_q && ((ldq % 8) || (ldy % 8))) || (do_kv && (ldkv % 8)))
      return set_error(HIG_ERR_INVALID, "eff_attn: bf16 leading dimensions must be multiples of 8");
    const int TP = (T + 15) & ~15;
    const size_t smem = (size_t)((mode == 3 ? 1 : (mode == 2 ? 2 : 3)) * TP + HD) * ATT_STRIDE * 2 + (ATT_WARPS + 1) * 64 * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(eff_attn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    using bf = __nv_bfloat16;
    eff_attn_bf16_kernel<<<grid, ATT_THREADS, smem, stream>>>(
        mode, (const bf*)q, ldq, (const bf*)k, (const bf*)v, ldkv, (const bf*)a_in, (bf*)a_out, (bf*)y, ldy, length, S,
        T, pair_shift, mask_v);
  } else if (dtype == HIG_F32) {
    const size_t smem = (size_t)(2 * T + HD) * F32_STRIDE * 4 + (4 * 64 + 8 * 64) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
      e = cudaFuncSetAttribute(eff_attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn attr: ") + cudaGetErrorString(e));
      configured = smem;
    }
    eff_attn_f32_kernel<<<grid, F32_THREADS, smem, stream>>>(
        mode, (const float*)q, ldq, (const float*)k, (const float*)v, ldkv, (const float*)a_in, (float*)a_out,
        (float*)y, ldy, length, S, T, pair_shift, mask_v);
  } else {
    return set_error(HIG_ERR_INVALID, "eff_attn: bad dtype");
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
