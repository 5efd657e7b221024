// hig_eff_attn_bwd, bf16 storage: backward of the efficient attention on tensor cores (mma.sync m16n8k16, fp32
// accumulate), one CTA per (sequence, head), everything between the loads of Q / K / V / dY and the stores of
// dQ / dK / dV stays in shared memory and registers.
//
// Forward (LinearTemporalSelfAttention / CrossAttention / InteractionCrossAttention.forward,
// codes/models/interaction_transformer.py:112-130, 145-165, 181-207; differentiated in the reference by torch.autograd):
//   Qs = softmax_feat(Q)   Ks = softmax_time(K, masked)   A = Ks^T V   Y = Qs A
// Backward, given dY (Qs, Ks, A are recomputed here):
//   dA  = Qs^T dY                      dQs = dY A^T        dQ = Qs * (dQs - rowsum(dQs * Qs))
//   dV  = Ks dA                        dKs = V dA^T        dK = Ks * (dKs - colsum(dKs * Ks))
// Same modes as the CUDA-core kernel in eff_attn_bwd.cu (which remains the fp32-mode implementation):
//   0 SELF, 1 INTER (K, V, dK, dV of the partner sequence, mask of the query side), 2 KV_ONLY (dA in), 3 Q_ONLY (dA out).
//
// Shared memory: four [TP, 64] bf16 tiles (K -> Ks -> dK, V -> dV, Q -> Qs -> dQ, dY) in 128-byte rows whose 16-byte
// chunks are XOR-swizzled with row & 7 (conflict-free ldmatrix, plain and transposed), A and dA as [64, 64] bf16.
// The round-1 kernel did these contractions as per-thread fp32 dot products out of padded fp32 shared memory:
// 512 us per launch at the C4 shape (S = 256, T = 91) against ~26 us of HBM traffic.
#include <cstdlib>
#include <string>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

namespace tcb {

constexpr int TC_THREADS = 256;
constexpr int TC_WARPS = 8;
constexpr int TC_HD = 64;

HIG_DEVICE void tcb_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
HIG_DEVICE void tcb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
HIG_DEVICE void tcb_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
HIG_DEVICE void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HIG_DEVICE void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk c of row r in a [rows][128 B] tile, chunks XOR-swizzled with r & 7
HIG_DEVICE uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
// byte offset of the bf16 pair at (row r, even column col)
HIG_DEVICE uint32_t swz_el(int r, int col) { return swz(r, col >> 3) + (uint32_t)((col & 7) * 2); }
HIG_DEVICE uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
HIG_DEVICE void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// A operand (16 rows m0.. x 16 k of chunk pair kc) from a tile stored [m][k]
HIG_DEVICE void frag_a(uint32_t tile, int m0, int kc, int lane, uint32_t (&a)[4]) {
  const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = kc * 2 + ((lane >> 4) & 1);
  ldsm4(tile + swz(row, c), a[0], a[1], a[2], a[3]);
}
// A operand (16 rows m0.. x 16 k starting at k0) from a tile stored [k][m]  (A^T in memory)
HIG_DEVICE void frag_a_t(uint32_t tile, int m0, int k0, int lane, uint32_t (&a)[4]) {
  const int row = k0 + (lane & 7) + ((lane >> 4) & 1) * 8;
  const int c = (m0 >> 3) + ((lane >> 3) & 1);
  ldsm4t(tile + swz(row, c), a[0], a[1], a[2], a[3]);
}
// B operands of two adjacent 8-column n tiles (n0.., n0+8..) x 16 k starting at k0, from a tile stored [k][n]
HIG_DEVICE void frag_b_kn(uint32_t tile, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = (n0 >> 3) + ((lane >> 4) & 1);
  ldsm4t(tile + swz(row, c), b[0], b[1], b[2], b[3]);
}
// the same from a tile stored [n][k]  (k contiguous), k chunk pair kc
HIG_DEVICE void frag_b_nk(uint32_t tile, int n0, int kc, int lane, uint32_t (&b)[4]) {
  const int row = n0 + (lane & 7) + ((lane >> 4) & 1) * 8;
  const int c = kc * 2 + ((lane >> 3) & 1);
  ldsm4(tile + swz(row, c), b[0], b[1], b[2], b[3]);
}

}  // namespace tcb
using namespace tcb;

// MINB: resident CTAs per SM the register allocation aims for (2: 127 registers, no spills; 3: 80 registers, ~350 B spilled)
template <int MINB>
__global__ void __launch_bounds__(TC_THREADS, MINB)
eff_attn_bwd_tc_kernel(int mode, const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k,
                       const __nv_bfloat16* __restrict__ v, int ldkv, const __nv_bfloat16* __restrict__ a_in,
                       const __nv_bfloat16* __restrict__ dy, int lddy, __nv_bfloat16* __restrict__ dq, int lddq,
                       __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int lddkv,
                       float* __restrict__ dA_g, const int* __restrict__ length, int S, int T, int pair_shift,
                       float* __restrict__ q_sum, float* __restrict__ k_sum, float* __restrict__ v_sum) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  const int TP = (T + 15) & ~15;
  const int ntiles = TP >> 4;
  const uint32_t base = smem_u32(tc_smem);
  const uint32_t sK = base;
  const uint32_t sV = sK + (do_kv ? TP * 128 : 0);
  const uint32_t sQ = sV + (do_kv ? TP * 128 : 0);
  const uint32_t sY = sQ + (do_q ? TP * 128 : 0);
  const uint32_t sA = sY + (do_q ? TP * 128 : 0);
  const uint32_t sdA = sA + TC_HD * 128;
  float* fred = reinterpret_cast<float*>(tc_smem + (sdA - base) + TC_HD * 128);   // [8][64] partials, then reused
  float* finv = fred + TC_WARPS * 64;                                               // [64]
  float* fcs = finv + 64;                                                           // [64]

  const int h = blockIdx.x, s = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int s_kv = (mode == 1) ? (s + pair_shift) % S : s;
  int len = T;
  if (length != nullptr && (mode == 0 || mode == 1)) {
    len = length[s];
    len = len < 0 ? 0 : (len > T ? T : len);
  }

  // ---------------------------------------------------------------- loads
  if (do_kv) {
    const __nv_bfloat16* kg = k + (size_t)s_kv * T * ldkv + h * TC_HD;
    const __nv_bfloat16* vg = v + (size_t)s_kv * T * ldkv + h * TC_HD;
    for (int i = tid; i < TP * 8; i += TC_THREADS) {
      const int r = i >> 3, c = i & 7;
      const bool ok = r < len;   // rows >= len: Ks == 0 exactly and V is masked (or meets Ks == 0)
      const size_t off = (size_t)min(r, T - 1) * ldkv + c * 8;
      tcb_cp16(sK + swz(r, c), kg + off, ok);
      tcb_cp16(sV + swz(r, c), vg + off, ok);
    }
  }
  if (do_q) {
    const __nv_bfloat16* qg = q + (size_t)s * T * ldq + h * TC_HD;
    const __nv_bfloat16* yg = dy + (size_t)s * T * lddy + h * TC_HD;
    for (int i = tid; i < TP * 8; i += TC_THREADS) {
      const int r = i >> 3, c = i & 7;
      const bool ok = r < T;
      tcb_cp16(sQ + swz(r, c), qg + (size_t)min(r, T - 1) * ldq + c * 8, ok);
      tcb_cp16(sY + swz(r, c), yg + (size_t)min(r, T - 1) * lddy + c * 8, ok);
    }
    if (!do_kv) {   // Q_ONLY: A comes from the text K/V side
      const __nv_bfloat16* ag = a_in + ((size_t)s * H + h) * TC_HD * TC_HD;
      for (int i = tid; i < TC_HD * 8; i += TC_THREADS) tcb_cp16(sA + swz(i >> 3, i & 7), ag + (size_t)i * 8, true);
    }
  } else {
    // KV_ONLY: dA is an input (fp32) -> bf16 operand tile
    const float* ig = dA_g + ((size_t)s * H + h) * TC_HD * TC_HD;
    for (int i = tid; i < TC_HD * TC_HD / 2; i += TC_THREADS) {
      const int d = i >> 5, l = (i & 31) * 2;
      const float2 f = *reinterpret_cast<const float2*>(ig + d * TC_HD + l);
      sts32(sdA + swz_el(d, l), pack_bf16x2(f.x, f.y));
    }
  }
  tcb_commit();
  tcb_wait_all();
  __syncthreads();

  // ---------------------------------------------------------------- softmaxes
  if (do_kv) {
    // time softmax of K (columns): lane owns chunk cc (8 columns) of rows rr + 4 warp + 32 i
    const int rr = lane >> 3, cc = lane & 7;
    float mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
    for (int t = warp * 4 + rr; t < len; t += 4 * TC_WARPS) {
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(sK + swz(t, cc)));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        mx[2 * j] = fmaxf(mx[2 * j], f.x);
        mx[2 * j + 1] = fmaxf(mx[2 * j + 1], f.y);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 8));
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 16));
    }
    if (rr == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) fred[warp * 64 + cc * 8 + j] = mx[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m = fred[cc * 8 + j];
#pragma unroll
      for (int w = 1; w < TC_WARPS; ++w) m = fmaxf(m, fred[w * 64 + cc * 8 + j]);
      mx[j] = m;
    }
    __syncthreads();
    float sm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j] = 0.f;
    for (int t = warp * 4 + rr; t < len; t += 4 * TC_WARPS) {
      const uint32_t addr = sK + swz(t, cc);
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        const float e0 = __expf(f.x - mx[2 * j]), e1 = __expf(f.y - mx[2 * j + 1]);
        sm[2 * j] += e0;
        sm[2 * j + 1] += e1;
        // keep fp32 exponentials for the normalisation pass: stash as bf16 now, renormalise below from the same values
        w[j] = pack_bf16x2(e0, e1);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[j] += __shfl_xor_sync(0xffffffffu, sm[j], 8);
      sm[j] += __shfl_xor_sync(0xffffffffu, sm[j], 16);
    }
    if (rr == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) fred[warp * 64 + cc * 8 + j] = sm[j];
    }
    __syncthreads();
    if (tid < 64) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < TC_WARPS; ++w) tot += fred[w * 64 + tid];
      finv[tid] = tot > 0.f ? 1.0f / tot : 0.f;
      fcs[tid] = 0.f;
    }
    __syncthreads();
    float iv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) iv[j] = finv[cc * 8 + j];
    for (int t = warp * 4 + rr; t < len; t += 4 * TC_WARPS) {
      const uint32_t addr = sK + swz(t, cc);
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        w[j] = pack_bf16x2(f.x * iv[2 * j], f.y * iv[2 * j + 1]);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
  }
  if (do_q) {
    // feature softmax of Q (rows): four lanes per row (16 columns = two 16-byte chunks each), a warp takes 8 rows at a
    // time; the row maximum / sum are two quad shuffles instead of two full-warp reductions
    const int qr = lane >> 2, qc = lane & 3;
    // (warp-uniform trip count: the quad shuffles below are full-mask; rows t >= T are padding rows of the tile — they exist
    //  in shared memory, hold zeros and are simply not written back)
    for (int tb = warp * 8; tb < T; tb += TC_WARPS * 8) {
      const int t = tb + qr;
      float f[16];
      uint32_t w[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4 * c]), "=r"(w[4 * c + 1]), "=r"(w[4 * c + 2]), "=r"(w[4 * c + 3])
                     : "r"(sQ + swz(t, 2 * qc + c)));
      }
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 p = unpack_bf16x2(w[j]);
        f[2 * j] = p.x;
        f[2 * j + 1] = p.y;
        m = fmaxf(m, fmaxf(p.x, p.y));
      }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[j] = __expf(f[j] - m);
        sum += f[j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(f[2 * j] * inv, f[2 * j + 1] * inv);
      if (t < T) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sQ + swz(t, 2 * qc + c)), "r"(w[4 * c]), "r"(w[4 * c + 1]),
                       "r"(w[4 * c + 2]), "r"(w[4 * c + 3]) : "memory");
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- A = Ks^T V,  dA = Qs^T dY   (64 x 64 each)
  // warp -> rows d in [dw, dw+16), columns l in [lw, lw+32).  The column sums the K-side softmax backward needs,
  //   cs[d] = sum_t dKs[t,d] Ks[t,d] = sum_t Ks[t,d] sum_l V[t,l] dA[d,l] = sum_l dA[d,l] A[d,l],
  // are row dot products of the two 64 x 64 matrices this phase holds in registers: no per-tile reduction later.
  {
    const int dw = (warp & 3) * 16, lw = (warp >> 2) * 32;
    float accA[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) accA[i][j] = 0.f;
    if (do_kv) {
      const int kend = (len + 15) & ~15;
      for (int kt = 0; kt < kend; kt += 16) {
        uint32_t a[4], b[4];
        frag_a_t(sK, dw, kt, lane, a);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          frag_b_kn(sV, kt, lw + np * 16, lane, b);
          mma16816(accA[2 * np], a, b[0], b[1]);
          mma16816(accA[2 * np + 1], a, b[2], b[3]);
        }
      }
      if (do_q) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = lw + nt * 8 + 2 * tg;
          sts32(sA + swz_el(dw + g, col), pack_bf16x2(accA[nt][0], accA[nt][1]));
          sts32(sA + swz_el(dw + g + 8, col), pack_bf16x2(accA[nt][2], accA[nt][3]));
        }
      }
    }
    uint32_t dAb[4][2];   // bf16 pairs of dA at this thread's fragment positions (rows dw+g / dw+g+8)
    if (do_q) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int kt = 0; kt < TP; kt += 16) {
        uint32_t a[4], b[4];
        frag_a_t(sQ, dw, kt, lane, a);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          frag_b_kn(sY, kt, lw + np * 16, lane, b);
          mma16816(acc[2 * np], a, b[0], b[1]);
          mma16816(acc[2 * np + 1], a, b[2], b[3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        dAb[nt][0] = pack_bf16x2(acc[nt][0], acc[nt][1]);
        dAb[nt][1] = pack_bf16x2(acc[nt][2], acc[nt][3]);
      }
      if (mode == 3) {
        float* og = dA_g + ((size_t)s * H + h) * TC_HD * TC_HD;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = lw + nt * 8 + 2 * tg;
          *reinterpret_cast<float2*>(og + (dw + g) * TC_HD + col) = make_float2(acc[nt][0], acc[nt][1]);
          *reinterpret_cast<float2*>(og + (dw + g + 8) * TC_HD + col) = make_float2(acc[nt][2], acc[nt][3]);
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = lw + nt * 8 + 2 * tg;
          sts32(sdA + swz_el(dw + g, col), dAb[nt][0]);
          sts32(sdA + swz_el(dw + g + 8, col), dAb[nt][1]);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = lw + nt * 8 + 2 * tg;
        dAb[nt][0] = lds32(sdA + swz_el(dw + g, col));
        dAb[nt][1] = lds32(sdA + swz_el(dw + g + 8, col));
      }
    }
    if (do_kv) {
      float c0 = 0.f, c1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 d0 = unpack_bf16x2(dAb[nt][0]), d1 = unpack_bf16x2(dAb[nt][1]);
        c0 = fmaf(d0.x, accA[nt][0], fmaf(d0.y, accA[nt][1], c0));
        c1 = fmaf(d1.x, accA[nt][2], fmaf(d1.y, accA[nt][3], c1));
      }
      c0 += __shfl_xor_sync(0xffffffffu, c0, 1); c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
      c1 += __shfl_xor_sync(0xffffffffu, c1, 1); c1 += __shfl_xor_sync(0xffffffffu, c1, 2);
      if (tg == 0) {          // the two column halves (warp >> 2) of a row meet in shared memory
        atomicAdd(&fcs[dw + g], c0);
        atomicAdd(&fcs[dw + g + 8], c1);
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- per 16-row tile of t (a warp owns its tiles' rows)
  for (int tt = warp; tt < ntiles; tt += TC_WARPS) {
    const int t0 = tt * 16;
    if (do_q) {
      // dQs[t, d] = sum_l dY[t, l] A[d, l]
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t a[4], b[4];
        frag_a(sY, t0, kc, lane, a);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          frag_b_nk(sA, np * 16, kc, lane, b);
          mma16816(acc[2 * np], a, b[0], b[1]);
          mma16816(acc[2 * np + 1], a, b[2], b[3]);
        }
      }
      // The bf16-rounded Qs of a row sum to 1 only to ~4e-3; dividing the row dot product by that sum keeps
      // sum_d dQ[t, d] == 0 (and dQ == 0 when A has identical rows) exactly as with unrounded softmax weights.
      float2 qs0[8], qs1[8];
      float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * tg;
        qs0[nt] = unpack_bf16x2(lds32(sQ + swz_el(t0 + g, col)));
        qs1[nt] = unpack_bf16x2(lds32(sQ + swz_el(t0 + g + 8, col)));
        d0 = fmaf(acc[nt][0], qs0[nt].x, fmaf(acc[nt][1], qs0[nt].y, d0));
        d1 = fmaf(acc[nt][2], qs1[nt].x, fmaf(acc[nt][3], qs1[nt].y, d1));
        n0 += qs0[nt].x + qs0[nt].y;
        n1 += qs1[nt].x + qs1[nt].y;
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      n0 += __shfl_xor_sync(0xffffffffu, n0, 1); n0 += __shfl_xor_sync(0xffffffffu, n0, 2);
      n1 += __shfl_xor_sync(0xffffffffu, n1, 1); n1 += __shfl_xor_sync(0xffffffffu, n1, 2);
      d0 = n0 > 0.f ? d0 / n0 : 0.f;     // padding rows (t >= T) carry Qs == 0
      d1 = n1 > 0.f ? d1 / n1 : 0.f;
      __syncwarp();   // every lane has read its Qs values before the tile's rows are overwritten with dQ
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * tg;
        sts32(sQ + swz_el(t0 + g, col), pack_bf16x2(qs0[nt].x * (acc[nt][0] - d0), qs0[nt].y * (acc[nt][1] - d0)));
        sts32(sQ + swz_el(t0 + g + 8, col), pack_bf16x2(qs1[nt].x * (acc[nt][2] - d1), qs1[nt].y * (acc[nt][3] - d1)));
      }
    }
    if (do_kv) {
      // dKs[t, d] = sum_l V[t, l] dA[d, l]       dV[t, l] = sum_d Ks[t, d] dA[d, l]
      float dks[8][4], dvv[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dks[i][j] = 0.f; dvv[i][j] = 0.f; }
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t av[4], ak[4], b[4];
        frag_a(sV, t0, kc, lane, av);
        frag_a(sK, t0, kc, lane, ak);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          frag_b_nk(sdA, np * 16, kc, lane, b);
          mma16816(dks[2 * np], av, b[0], b[1]);
          mma16816(dks[2 * np + 1], av, b[2], b[3]);
          frag_b_kn(sdA, kc * 16, np * 16, lane, b);
          mma16816(dvv[2 * np], ak, b[0], b[1]);
          mma16816(dvv[2 * np + 1], ak, b[2], b[3]);
        }
      }
      __syncwarp();   // all K / V fragments of this tile are in registers before its rows become dK / dV
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * tg;
        sts32(sV + swz_el(t0 + g, col), pack_bf16x2(dvv[nt][0], dvv[nt][1]));
        sts32(sV + swz_el(t0 + g + 8, col), pack_bf16x2(dvv[nt][2], dvv[nt][3]));
        const float c0 = fcs[col], c1 = fcs[col + 1];
        const uint32_t a0 = sK + swz_el(t0 + g, col), a1 = sK + swz_el(t0 + g + 8, col);
        const float2 k0 = unpack_bf16x2(lds32(a0)), k1 = unpack_bf16x2(lds32(a1));
        sts32(a0, pack_bf16x2(k0.x * (dks[nt][0] - c0), k0.y * (dks[nt][1] - c1)));
        sts32(a1, pack_bf16x2(k1.x * (dks[nt][2] - c0), k1.y * (dks[nt][3] - c1)));
      }
    }
  }
  // column sums of the gradient tiles (the projections' bias gradients), optional: a thread stores the same 16-byte chunk
  // (8 columns) of every row it handles, so the sums ride on the store loop; partials meet in fred[3][64]
  const bool sums = (q_sum != nullptr) || (k_sum != nullptr) || (v_sum != nullptr);
  if (sums && tid < 192) fred[tid] = 0.f;
  __syncthreads();

  auto add8 = [](float (&acc)[8], const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  };
  auto flush8 = [&](float (&acc)[8], float* dst) {   // lanes with equal (lane & 7) own the same chunk: fold rows, then 8 lanes publish
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dst + lane * 8 + j, acc[j]);
    }
  };

  // ---------------------------------------------------------------- coalesced stores (16 bytes per lane)
  if (do_q) {
    __nv_bfloat16* og = dq + (size_t)s * T * lddq + h * TC_HD;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < T * 8; i += TC_THREADS) {
      const int r = i >> 3, c = i & 7;
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(sQ + swz(r, c)));
      *reinterpret_cast<uint4*>(og + (size_t)r * lddq + c * 8) = u;
      if (q_sum != nullptr) add8(acc, u);
    }
    if (q_sum != nullptr) flush8(acc, fred);
  }
  if (do_kv) {
    __nv_bfloat16* okg = dk + (size_t)s_kv * T * lddkv + h * TC_HD;
    __nv_bfloat16* ovg = dv + (size_t)s_kv * T * lddkv + h * TC_HD;
    float acck[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < T * 8; i += TC_THREADS) {
      const int r = i >> 3, c = i & 7;
      uint4 u, w;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(sK + swz(r, c)));
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(sV + swz(r, c)));
      *reinterpret_cast<uint4*>(okg + (size_t)r * lddkv + c * 8) = u;
      *reinterpret_cast<uint4*>(ovg + (size_t)r * lddkv + c * 8) = w;
      if (k_sum != nullptr) add8(acck, u);
      if (v_sum != nullptr) add8(accv, w);
    }
    if (k_sum != nullptr) flush8(acck, fred + 64);
    if (v_sum != nullptr) flush8(accv, fred + 128);
  }
  if (sums) {
    __syncthreads();
    if (tid < 192) {
      float* dst = tid < 64 ? q_sum : (tid < 128 ? k_sum : v_sum);
      if (dst != nullptr) atomicAdd(dst + h * TC_HD + (tid & 63), fred[tid]);
    }
  }
}

// returns HIG_ERR_UNSUPPORTED when the shape / alignment does not fit this kernel (the caller falls back to the
// CUDA-core kernel of eff_attn_bwd.cu)
int eff_attn_bwd_tc(int mode, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* a_in,
                    const void* dy, int lddy, void* dq, int lddq, void* dk, void* dv, int lddkv, float* dA,
                    const int* length, int S, int T, int H, int pair_shift, float* q_sum, float* k_sum, float* v_sum,
                    cudaStream_t stream) {
  const bool do_kv = (mode != 3), do_q = (mode != 2);
  auto mis = [](const void* p, int ld) { return p != nullptr && ((reinterpret_cast<uintptr_t>(p) & 15) || (ld % 8)); };
  if (T > 256 || mis(q, ldq) || mis(k, ldkv) || mis(v, ldkv) || mis(dy, lddy) || mis(dq, lddq) || mis(dk, lddkv) ||
      mis(dv, lddkv) || mis(a_in, 8) || (dA && (reinterpret_cast<uintptr_t>(dA) & 7)))
    return HIG_ERR_UNSUPPORTED;
  const int TP = (T + 15) & ~15;
  const size_t smem = (size_t)((do_kv ? 2 : 0) + (do_q ? 2 : 0)) * TP * 128 + 2 * TC_HD * 128 + (TC_WARPS * 64 + 128) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(eff_attn_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(eff_attn_bwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn_bwd_tc attr: ") + cudaGetErrorString(e));
    configured = smem;
  }
  using bf = __nv_bfloat16;
  // three CTAs per SM when their shared memory fits (T <= 96 at 4 tiles): HIG_ATTN_BWD_OCC=2 keeps the 127-register build
  static const int occ = []() { const char* ev = getenv("HIG_ATTN_BWD_OCC"); return ev ? atoi(ev) : 3; }();
  auto kern = (occ >= 3 && smem * 3 <= 227 * 1024) ? eff_attn_bwd_tc_kernel<3> : eff_attn_bwd_tc_kernel<2>;
  kern<<<dim3(H, S), TC_THREADS, smem, stream>>>(
      mode, (const bf*)q, ldq, (const bf*)k, (const bf*)v, ldkv, (const bf*)a_in, (const bf*)dy, lddy, (bf*)dq, lddq,
      (bf*)dk, (bf*)dv, lddkv, dA, length, S, T, pair_shift, do_q ? q_sum : nullptr, do_kv ? k_sum : nullptr,
      do_kv ? v_sum : nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("eff_attn_bwd_tc launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
