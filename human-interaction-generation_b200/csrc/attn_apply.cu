// hig_attn_apply_stylize: the query half of the efficient attention fused with the StylizationBlock's
// LayerNorm + FiLM + SiLU, for all 8 heads of a row at once:
//
//   Y[t, h*64 : (h+1)*64] = softmax_feat(Q[t, h]) . A[s, h]          (A = softmax_time(K)^T V, 64 x 64 per head,
//                                                                      produced by hig_eff_attn KV_ONLY)
//   out[t, :] = SiLU( LayerNorm_512(Y[t, :]) * (1 + scale_s) + shift_s )
//
// Reference: the einsum 'bnhd,bhdl->bnhl' + reshape of LinearTemporal{Self,Cross,InteractionCross}Attention.forward
// (codes/models/interaction_transformer.py:128,162,201) followed by StylizationBlock.forward's norm / FiLM / SiLU
// (:86-97).  Y never reaches HBM: a CTA owns full 512-wide rows (warp w = head w), so the LayerNorm statistics are
// exchanged between the 8 warps through shared memory.  Traffic: read Q + A (L2-resident), write out.
//
// CTA = 8 warps, one 16-row tile per iteration; Q tiles are prefetched with cp.async into a per-warp double buffer
// (XOR-swizzled 128-byte rows, conflict-free ldmatrix); A for the 8 heads sits in shared memory (64 KB, swizzled).
#include <cstdlib>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int AP_THREADS = 256;
constexpr int AP_HD = 64;
constexpr int AP_D = 512;

HIG_DEVICE void ap_cp_async16(uint32_t smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
HIG_DEVICE void ap_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> HIG_DEVICE void ap_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
HIG_DEVICE void ap_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void ap_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void ap_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `c` of row `r` in a [rows][128 B] tile whose chunks are XOR-swizzled with r & 7
HIG_DEVICE uint32_t ap_swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

constexpr int AP_SMEM_A = 8 * AP_HD * 128;          // 64 KB: A[h][d][l]
constexpr int AP_SMEM_Q = 8 * 2 * 16 * 128;         // 32 KB: per warp, 2 buffers of 16 rows x 128 B
constexpr int AP_SMEM_GB = 2 * AP_D * 4;            // gamma', beta'
constexpr int AP_SMEM_RED = 2 * 16 * 8 * 8;         // two buffers of row partials (sum, sum of squares) [16 rows][8 warps]
constexpr int AP_SMEM = AP_SMEM_A + AP_SMEM_Q + AP_SMEM_GB + AP_SMEM_RED;

__global__ void __launch_bounds__(AP_THREADS, 2)
attn_apply_stylize_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ a_in,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ scale_shift, int ss_stride, int apply_silu,
                          __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ y_out, int S, int T) {
  // y_out (nullable, training forward): the attention output itself, bf16 [S*T, 512] — the LayerNorm backward needs it
  extern __shared__ __align__(128) uint8_t ap_smem[];
  const uint32_t sA = smem_u32(ap_smem);
  const uint32_t sQ = sA + AP_SMEM_A;
  float* sG = reinterpret_cast<float*>(ap_smem + AP_SMEM_A + AP_SMEM_Q);
  float* sB = sG + AP_D;
  float2* sR = reinterpret_cast<float2*>(sB + AP_D);   // [2][16][8]

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int n_tiles = (T + 15) >> 4;
  // balanced static schedule: the S * n_tiles row tiles are cut into gridDim.x contiguous ranges (sizes differ by at
  // most one tile); a range touches at most two sequences when gridDim.x >= S, so A is (re)loaded at most twice
  const long long G = (long long)S * n_tiles;
  const int g_lo = (int)(G * blockIdx.x / gridDim.x), g_hi = (int)(G * (blockIdx.x + 1) / gridDim.x);
  if (g_lo >= g_hi) return;

  const uint32_t sQw = sQ + w * (2 * 16 * 128);
  // this warp's Q tile of global tile index gi: 16 rows x 8 chunks = 128 x 16 B, 4 per lane
  auto prefetch = [&](int gi, int buf) {
    const int ps = gi / n_tiles, ptile = gi - ps * n_tiles;
    const __nv_bfloat16* qg = q + (size_t)ps * T * ldq + w * AP_HD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int r = idx >> 3, c = idx & 7;
      const int t = ptile * 16 + r;
      ap_cp_async16(sQw + buf * 2048 + ap_swz(r, c), qg + (size_t)min(t, T - 1) * ldq + c * 8, t < T);
    }
  };
  pdl_wait();      // q and a_in are outputs of the previous kernels
  pdl_trigger();
  prefetch(g_lo, 0);
  ap_commit();

  const int g = lane >> 2, tg = lane & 3;
  // flags: bit 0 = SiLU after the FiLM affine; bit 1 = q already holds softmax_feat(Q) (written by the Q / Q|K|V
  // projection's epilogue, HIG_GS_LN_QSM): the fragments ldmatrix returns are the MMA operands as they are
  const bool q_ready = (apply_silu & 2) != 0;
  apply_silu &= 1;
  const float hs = apply_silu ? 0.5f : 1.0f;
  const uint32_t sAw = sA + w * (AP_HD * 128);
  int buf = 0, cur_s = -1;
  for (int gi = g_lo; gi < g_hi; ++gi, buf ^= 1) {
    const int s = gi / n_tiles, tile = gi - s * n_tiles;
    if (s != cur_s) {
      // ---- new sequence: A of its 8 heads (8 * 64 rows * 8 chunks of 16 B) and the folded FiLM affine
      //      out = n_hat * G + B,  G = gamma (1 + scale),  B = beta (1 + scale) + shift
      __syncthreads();   // every warp is done with the previous sequence's A / sG / sB
      const __nv_bfloat16* ag = a_in + (size_t)s * 8 * AP_HD * AP_HD;
      for (int i = tid; i < 8 * AP_HD * 8; i += AP_THREADS) {
        const int row = i >> 3, c = i & 7;  // row = h*64 + d
        ap_cp_async16(sA + ap_swz(row, c), ag + (size_t)row * AP_HD + c * 8, true);
      }
      ap_commit();
      for (int i = tid; i < AP_D; i += AP_THREADS) {
        float gg = gamma[i], bb = beta[i];
        if (scale_shift) {
          const float m1 = 1.0f + scale_shift[(size_t)s * ss_stride + i];
          gg *= m1;
          bb = fmaf(bb, m1, scale_shift[(size_t)s * ss_stride + AP_D + i]);
        }
        sG[i] = gg * hs;
        sB[i] = bb * hs;
      }
      ap_wait<0>();
      __syncthreads();
      cur_s = s;
    }
    if (gi + 1 < g_hi) prefetch(gi + 1, buf ^ 1);
    ap_commit();
    ap_wait<1>();
    __syncwarp();
    const uint32_t sQt = sQw + buf * 2048;
    uint32_t af[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = kk * 2 + ((lane >> 4) & 1);
      ap_ldsm_x4(sQt + ap_swz(row, c), af[kk][0], af[kk][1], af[kk][2], af[kk][3]);
    }
    // feature softmax on the fragments: row g <- regs {0,2}, row g+8 <- regs {1,3} of every k-step.  Column pairs stay
    // packed (FFMA2 / FADD2 / FMUL2); exp(x - m) = ex2(x log2e - m log2e) is one packed FMA + one MUFU per element.
    if (!q_ready) {
      uint64_t x0[8], x1[8];   // pair j of row g / g+8: columns 16 kk + {0,1} (j = 2kk) and 16 kk + 8 + {0,1} (j = 2kk+1) + 2tg
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        x0[2 * kk + 0] = f2_pack_u(af[kk][0] << 16, af[kk][0] & 0xffff0000u);
        x0[2 * kk + 1] = f2_pack_u(af[kk][2] << 16, af[kk][2] & 0xffff0000u);
        x1[2 * kk + 0] = f2_pack_u(af[kk][1] << 16, af[kk][1] & 0xffff0000u);
        x1[2 * kk + 1] = f2_pack_u(af[kk][3] << 16, af[kk][3] & 0xffff0000u);
      }
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a, b;
        f2_unpack(x0[j], a, b); m0 = fmaxf(m0, fmaxf(a, b));
        f2_unpack(x1[j], a, b); m1 = fmaxf(m1, fmaxf(a, b));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      {
        const float kL2E = 1.4426950408889634f;
        const uint64_t l2e = f2_pack(kL2E, kL2E);
        const uint64_t nm0 = f2_pack(-m0 * kL2E, -m0 * kL2E), nm1 = f2_pack(-m1 * kL2E, -m1 * kL2E);
        uint64_t sum0 = 0ull, sum1 = 0ull;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a, b;
          f2_unpack(f2_fma(x0[j], l2e, nm0), a, b);
          x0[j] = f2_pack(ex2_ftz(a), ex2_ftz(b));
          sum0 = f2_add(sum0, x0[j]);
          f2_unpack(f2_fma(x1[j], l2e, nm1), a, b);
          x1[j] = f2_pack(ex2_ftz(a), ex2_ftz(b));
          sum1 = f2_add(sum1, x1[j]);
        }
        float s0, s1, t0, t1;
        f2_unpack(sum0, s0, t0);
        f2_unpack(sum1, s1, t1);
        s0 += t0;
        s1 += t1;
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float i0 = 1.0f / s0, i1 = 1.0f / s1;
        const uint64_t i02 = f2_pack(i0, i0), i12 = f2_pack(i1, i1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float a, b;
          f2_unpack(f2_mul(x0[2 * kk + 0], i02), a, b); af[kk][0] = pack_bf16x2(a, b);
          f2_unpack(f2_mul(x0[2 * kk + 1], i02), a, b); af[kk][2] = pack_bf16x2(a, b);
          f2_unpack(f2_mul(x1[2 * kk + 0], i12), a, b); af[kk][1] = pack_bf16x2(a, b);
          f2_unpack(f2_mul(x1[2 * kk + 1], i12), a, b); af[kk][3] = pack_bf16x2(a, b);
        }
      }
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = np * 2 + ((lane >> 4) & 1);
        ap_ldsm_x4_t(sAw + ap_swz(row, c), b0, b1, b2, b3);
        ap_mma(acc[2 * np], af[kk], b0, b1);
        ap_mma(acc[2 * np + 1], af[kk], b2, b3);
      }
    }
    if (y_out != nullptr) {
      // attention output before the LayerNorm, staged through this warp's consumed Q buffer like the final result
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQt + ap_swz(g, nt) + 4 * tg), "r"(pack_bf16x2(acc[nt][0], acc[nt][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQt + ap_swz(g + 8, nt) + 4 * tg), "r"(pack_bf16x2(acc[nt][2], acc[nt][3])) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = i * 32 + lane;
        const int r = idx >> 3, c = idx & 7;
        const int t = tile * 16 + r;
        if (t < T) {
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(sQt + ap_swz(r, c)) : "memory");
          *reinterpret_cast<uint4*>(y_out + ((size_t)s * T + t) * AP_D + w * AP_HD + c * 8) = val;
        }
      }
      __syncwarp();
    }
    // ---- LayerNorm over the 512 columns of a row = 8 warps x 64 columns: (sum, sum of squares) partials of every warp
    //      meet in shared memory behind ONE barrier per tile (the partial buffer alternates between tiles)
    uint64_t y0[8], y1[8];
    uint64_t ps0 = 0ull, pq0 = 0ull, ps1 = 0ull, pq1 = 0ull;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      y0[nt] = f2_pack(acc[nt][0], acc[nt][1]);
      y1[nt] = f2_pack(acc[nt][2], acc[nt][3]);
      ps0 = f2_add(ps0, y0[nt]); pq0 = f2_fma(y0[nt], y0[nt], pq0);
      ps1 = f2_add(ps1, y1[nt]); pq1 = f2_fma(y1[nt], y1[nt], pq1);
    }
    float r0, r1, q0, q1;
    {
      float a, b;
      f2_unpack(ps0, a, b); r0 = a + b;
      f2_unpack(pq0, a, b); q0 = a + b;
      f2_unpack(ps1, a, b); r1 = a + b;
      f2_unpack(pq1, a, b); q1 = a + b;
    }
    r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    float2* sRb = sR + (buf ? 16 * 8 : 0);     // [16 rows][8 warps] (sum, sum of squares)
    if (tg == 0) { sRb[g * 8 + w] = make_float2(r0, q0); sRb[(g + 8) * 8 + w] = make_float2(r1, q1); }
    __syncthreads();
    float mean0, mean1, rstd0, rstd1;
    {
      float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 u = *reinterpret_cast<const float4*>(sRb + g * 8 + 2 * i);
        const float4 v = *reinterpret_cast<const float4*>(sRb + (g + 8) * 8 + 2 * i);
        sa += u.x + u.z; qa += u.y + u.w;
        sb += v.x + v.z; qb += v.y + v.w;
      }
      mean0 = sa * (1.0f / AP_D);
      mean1 = sb * (1.0f / AP_D);
      rstd0 = rsqrtf(fmaxf(fmaf(qa, 1.0f / AP_D, -mean0 * mean0), 0.f) + 1e-5f);
      rstd1 = rsqrtf(fmaxf(fmaf(qb, 1.0f / AP_D, -mean1 * mean1), 0.f) + 1e-5f);
    }
    // ---- affine + SiLU, staged through this warp's consumed Q buffer, then 16-byte row stores.  sG / sB already carry
    //      the factor 1/2 when SiLU follows:  h = x/2,  SiLU(x) = h + h tanh(h)
    __syncwarp();
    {
      const uint64_t rs0 = f2_pack(rstd0, rstd0), rs1 = f2_pack(rstd1, rstd1);
      const uint64_t c0 = f2_pack(-mean0 * rstd0, -mean0 * rstd0), c1 = f2_pack(-mean1 * rstd1, -mean1 * rstd1);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = w * AP_HD + nt * 8 + 2 * tg;
        const uint64_t G2 = *reinterpret_cast<const uint64_t*>(sG + col), B2 = *reinterpret_cast<const uint64_t*>(sB + col);
        uint64_t h0 = f2_fma(y0[nt], f2_mul(G2, rs0), f2_fma(c0, G2, B2));
        uint64_t h1 = f2_fma(y1[nt], f2_mul(G2, rs1), f2_fma(c1, G2, B2));
        float a, b, c, d;
        if (apply_silu) {
          f2_unpack(h0, a, b);
          f2_unpack(h1, c, d);
          h0 = f2_fma(h0, f2_pack(tanh_approx_f(a), tanh_approx_f(b)), h0);
          h1 = f2_fma(h1, f2_pack(tanh_approx_f(c), tanh_approx_f(d)), h1);
        }
        f2_unpack(h0, a, b);
        f2_unpack(h1, c, d);
        // element (row, nt*8 + 2tg) -> chunk nt, byte offset 4*tg inside the chunk
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQt + ap_swz(g, nt) + 4 * tg), "r"(pack_bf16x2(a, b)) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQt + ap_swz(g + 8, nt) + 4 * tg), "r"(pack_bf16x2(c, d)) : "memory");
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int r = idx >> 3, c = idx & 7;
      const int t = tile * 16 + r;
      if (t < T) {
        uint4 val;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                     : "r"(sQt + ap_swz(r, c)) : "memory");
        *reinterpret_cast<uint4*>(out + ((size_t)s * T + t) * AP_D + w * AP_HD + c * 8) = val;
      }
    }
    __syncwarp();
  }
  ap_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// attn_kv_kernel: the K/V half,  A[s,h] = softmax_time(K_masked)^T . (V * mask)   -> bf16 [S,H,64,64]
// (LinearTemporal*Attention.forward, :121-127 / :155-161 / :194-200).  One CTA per (sequence, head); K and V
// [T,64] staged with cp.async into unpadded XOR-swizzled shared memory (50 KB at T=196 -> 4 CTAs per SM), the time
// softmax runs column-wise in fp32 (lane = column pair, warp strides rows), the contraction on mma.sync.  K/V of
// sequence (s + pair_shift) % S, rows >= length[s] masked (the inter-person block's query-side mask quirk, :194).
// ------------------------------------------------------------------------------------------------
constexpr int KV_THREADS = 256;
constexpr int KV_WARPS = 8;

// TRANSPOSED: the output is A^T[s,h] ([l][d]) — the K-major B operand attn_apply_tc_kernel's tcgen05.mma reads.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(KV_THREADS, 4)
attn_kv_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int ldkv,
               __nv_bfloat16* __restrict__ a_out, const int* __restrict__ length, int S, int T, int pair_shift) {
  extern __shared__ __align__(128) uint8_t kv_smem[];
  const int TP = (T + 15) & ~15;
  const uint32_t sK = smem_u32(kv_smem);
  const uint32_t sV = sK + TP * 128;
  float* sred = reinterpret_cast<float*>(kv_smem + 2 * TP * 128);  // [KV_WARPS][64]
  float* sinv = sred + KV_WARPS * 64;                               // [64]
  const int h = blockIdx.x, s = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s_kv = (s + pair_shift) % S;
  int len = T;
  if (length) {
    len = length[s];
    len = len < 0 ? 0 : (len > T ? T : len);
  }
  const __nv_bfloat16* kg = k + (size_t)s_kv * T * ldkv + h * AP_HD;
  const __nv_bfloat16* vg = v + (size_t)s_kv * T * ldkv + h * AP_HD;
  pdl_wait();      // K, V are outputs of the previous kernel (length[] is older)
  pdl_trigger();
  // rows >= len are zero-filled: Ks is forced to 0 there and V*mask == 0 (unmasked V rows meet Ks == 0: same A)
  for (int i = tid; i < TP * 8; i += KV_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < len;
    const size_t off = (size_t)min(r, T - 1) * ldkv + c * 8;
    ap_cp_async16(sK + ap_swz(r, c), kg + off, ok);
    ap_cp_async16(sV + ap_swz(r, c), vg + off, ok);
  }
  ap_commit();
  ap_wait<0>();
  __syncthreads();

  // ---- time softmax of K: lane owns the 16-byte chunk (lane & 7) = 8 columns of rows (lane >> 3) + 4 warp + 32 i.
  //      Maxima stay packed bf16x2 (HMNMX2), exp(k - m) = ex2(k log2e - m log2e) is one packed FMA + one MUFU.
  const int rr = lane >> 3, cc = lane & 7;
  auto kchunk = [&](int t) { return sK + t * 128 + ((cc ^ (t & 7)) << 4); };
  auto hmax2 = [](uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  };
  uint32_t mx[4] = {0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u};   // (-inf, -inf)
  for (int t = warp * 4 + rr; t < len; t += 4 * KV_WARPS) {
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(kchunk(t)));
    mx[0] = hmax2(mx[0], u.x); mx[1] = hmax2(mx[1], u.y); mx[2] = hmax2(mx[2], u.z); mx[3] = hmax2(mx[3], u.w);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mx[i] = hmax2(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 8));
    mx[i] = hmax2(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 16));
  }
  uint32_t* sredu = reinterpret_cast<uint32_t*>(sred);   // [KV_WARPS][32 words]
  if (rr == 0) *reinterpret_cast<uint4*>(sredu + warp * 32 + cc * 4) = make_uint4(mx[0], mx[1], mx[2], mx[3]);
  __syncthreads();
#pragma unroll
  for (int w = 0; w < KV_WARPS; ++w) {
    const uint4 o = *reinterpret_cast<const uint4*>(sredu + w * 32 + cc * 4);
    mx[0] = hmax2(mx[0], o.x); mx[1] = hmax2(mx[1], o.y); mx[2] = hmax2(mx[2], o.z); mx[3] = hmax2(mx[3], o.w);
  }
  __syncthreads();   // sred is reused for the sums
  const float kL2E = 1.4426950408889634f;
  const uint64_t l2e = f2_pack(kL2E, kL2E);
  uint64_t nm[4], sm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 m = unpack_bf16x2(mx[i]);
    nm[i] = f2_pack(-m.x * kL2E, -m.y * kL2E);
    sm[i] = 0ull;
  }
  for (int t = warp * 4 + rr; t < len; t += 4 * KV_WARPS) {
    uint32_t u[4];
    const uint32_t addr = kchunk(t);
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(addr));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a, b;
      f2_unpack(f2_fma(f2_pack_u(u[i] << 16, u[i] & 0xffff0000u), l2e, nm[i]), a, b);
      // sum what the MMA will see (the bf16-rounded weights) so the normalisation is exact
      u[i] = pack_bf16x2(ex2_ftz(a), ex2_ftz(b));
      sm[i] = f2_add(sm[i], f2_pack_u(u[i] << 16, u[i] & 0xffff0000u));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
  }
  {
    float ssum[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) f2_unpack(sm[i], ssum[2 * i], ssum[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 8);
      ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 16);
    }
    if (rr == 0) {
      *reinterpret_cast<float4*>(sred + warp * 64 + cc * 8) = make_float4(ssum[0], ssum[1], ssum[2], ssum[3]);
      *reinterpret_cast<float4*>(sred + warp * 64 + cc * 8 + 4) = make_float4(ssum[4], ssum[5], ssum[6], ssum[7]);
    }
  }
  __syncthreads();
  if (tid < 64) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < KV_WARPS; ++w) tot += sred[w * 64 + tid];
    sinv[tid] = tot > 0.f ? 1.0f / tot : 0.f;
  }
  __syncthreads();

  // ---- A[d,l] = sum_t e[t,d] V[t,l] / sum_t e[t,d]: warp -> d in [16*(w&3), +16), l in [32*(w>>2), +32)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int dw = (warp & 3) * 16, lw = (warp >> 2) * 32;
  const int kend = (len + 15) & ~15;
  for (int kt = 0; kt < kend; kt += 16) {
    // TRANSPOSED: A^T[l,d] = sum_t V[t,l] e[t,d] — the same loop with the roles of the two tiles swapped
    const uint32_t sRowOp = TRANSPOSED ? sV : sK, sColOp = TRANSPOSED ? sK : sV;
    uint32_t a[4];
    {
      const int row = kt + (lane & 7) + ((lane >> 4) & 1) * 8;
      const int c = (dw >> 3) + ((lane >> 3) & 1);
      ap_ldsm_x4_t(sRowOp + ap_swz(row, c), a[0], a[1], a[2], a[3]);
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      const int row = kt + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = (lw >> 3) + np * 2 + ((lane >> 4) & 1);
      ap_ldsm_x4_t(sColOp + ap_swz(row, c), b0, b1, b2, b3);
      ap_mma(acc[2 * np], a, b0, b1);
      ap_mma(acc[2 * np + 1], a, b2, b3);
    }
  }
  const int g = lane >> 2, tg = lane & 3;
  const float i0 = sinv[dw + g], i1 = sinv[dw + g + 8];
  __nv_bfloat16* dst = a_out + ((size_t)s * H + h) * AP_HD * AP_HD;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = lw + nt * 8 + 2 * tg;
    if (TRANSPOSED) {   // rows are l, columns are d: the softmax normaliser belongs to the column
      const float c0 = sinv[col], c1 = sinv[col + 1];
      *reinterpret_cast<uint32_t*>(dst + (dw + g) * AP_HD + col) = pack_bf16x2(acc[nt][0] * c0, acc[nt][1] * c1);
      *reinterpret_cast<uint32_t*>(dst + (dw + g + 8) * AP_HD + col) = pack_bf16x2(acc[nt][2] * c0, acc[nt][3] * c1);
    } else {
      *reinterpret_cast<uint32_t*>(dst + (dw + g) * AP_HD + col) = pack_bf16x2(acc[nt][0] * i0, acc[nt][1] * i0);
      *reinterpret_cast<uint32_t*>(dst + (dw + g + 8) * AP_HD + col) = pack_bf16x2(acc[nt][2] * i1, acc[nt][3] * i1);
    }
  }
}

int attn_kv(const void* k, const void* v, int ldkv, void* a_out, const int* length, int S, int T, int H,
            int pair_shift, int transposed, cudaStream_t stream) {
  if (!k || !v || !a_out || S <= 0 || T <= 0 || H <= 0) return set_error(HIG_ERR_INVALID, "attn_kv: bad arguments");
  if (T > 256) return set_error(HIG_ERR_UNSUPPORTED, "attn_kv: T > 256 not supported");
  if (ldkv % 8) return set_error(HIG_ERR_INVALID, "attn_kv: ldkv must be a multiple of 8");
  const int TP = (T + 15) & ~15;
  const size_t smem = (size_t)2 * TP * 128 + (KV_WARPS + 1) * 64 * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_kv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_kv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_kv attr: ") + cudaGetErrorString(e));
    configured = smem;
  }
  cudaError_t e = launch_pdl(transposed ? attn_kv_kernel<true> : attn_kv_kernel<false>, dim3(H, S), dim3(KV_THREADS), smem,
                             stream, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv, (__nv_bfloat16*)a_out, length, S,
                             T, pair_shift);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_kv launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

int attn_apply_stylize(const void* q, int ldq, const void* a_in, const float* gamma, const float* beta,
                       const float* scale_shift, int ss_stride, int apply_silu, void* out, void* y_out, int S, int T, int H,
                       cudaStream_t stream) {
  if (!q || !a_in || !gamma || !beta || !out || S <= 0 || T <= 0)
    return set_error(HIG_ERR_INVALID, "attn_apply_stylize: bad arguments");
  if (H != 8) return set_error(HIG_ERR_UNSUPPORTED, "attn_apply_stylize: built for 8 heads x 64 (latent_dim 512)");
  if (ldq % 8) return set_error(HIG_ERR_INVALID, "attn_apply_stylize: ldq must be a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(a_in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(y_out) & 15))
    return set_error(HIG_ERR_INVALID, "attn_apply_stylize: pointers must be 16-byte aligned");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_apply_stylize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_apply_stylize attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  const int n_tiles = (T + 15) / 16;
  // two resident CTAs per SM, every CTA gets the same number of row tiles (+-1)
  static const int n_sm = []() { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
  long long ctas = 2LL * n_sm;
  if (ctas > (long long)S * n_tiles) ctas = (long long)S * n_tiles;
  cudaError_t e = launch_pdl(attn_apply_stylize_kernel, dim3((int)ctas), dim3(AP_THREADS), AP_SMEM, stream,
                             (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)a_in, gamma, beta, scale_shift, ss_stride,
                             apply_silu, (__nv_bfloat16*)out, (__nv_bfloat16*)y_out, S, T);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_apply_stylize launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
