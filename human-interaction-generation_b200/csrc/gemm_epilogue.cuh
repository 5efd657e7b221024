// Shared epilogue of the tcgen05 GEMM kernels: TMEM -> registers -> swizzled smem slab -> coalesced global I/O.
#pragma once
#include <cuda_fp16.h>
#include "hig_common.cuh"

namespace hig {

struct GemmEpilogue {
  const float* bias;      // [N] or nullptr
  const float* residual;  // fp32 [rows, ldr] or nullptr
  int ldr;
  int res_row_mod;        // >0: residual row = m % res_row_mod (positional tables)
  float* out_f32;         // nullable
  int ldo_f32;
  __nv_bfloat16* out_bf16;  // nullable
  int ldo_bf16;
  int act;                // 0 none, 1 GELU(erf), 2 SiLU
  int atomic;             // 1: split-K partial -> atomicAdd into out_f32 (generic kind only; no bias/act)
  // fp16 residual stream (product path): residual16 / out16 replace residual / out_f32 when set
  const __half* residual16;
  int ldr16;
  __half* out16;
  int ldo16;
  // training fusions (generic kind, vector path only; null = off):
  //   out_pre: bf16 copy of the PRE-activation value (the activation is then applied to that stored, rounded value — what a
  //            separate activation kernel reading it back would see);  gate: the result is multiplied by act'(gate[row, col])
  //            (backward of an activation fused into the GEMM that produces its incoming gradient)
  __nv_bfloat16* out_pre = nullptr;
  int ldo_pre = 0;
  const __nv_bfloat16* gate = nullptr;
  int ld_gate = 0;
  int gate_act = 0;
};

// two fp32 -> packed fp16x2 with saturation to +-65504 (no inf in the fp16 residual stream)
HIG_DEVICE uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
HIG_DEVICE float2 unpack_f16x2(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

HIG_DEVICE void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
HIG_DEVICE float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
HIG_DEVICE float ld_shared_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Rare path (N or a leading dimension not 16-byte friendly, e.g. the 263-wide output heads): element-wise with
// bounds checks, still coalesced across the warp.  Kept out of line so the hot path stays small in the I-cache.
static __device__ __noinline__ void epilogue_chunk_scalar(uint32_t slab, GemmEpilogue ep, int row0, int M, int col0,
                                                          int N, int lane) {
  const int sub = lane >> 3, cg = lane & 7;
  const int gcol = col0 + cg * 4;
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + sub;
    const int row = row0 + rl;
    if (row >= M) continue;
    const int rr = ep.res_row_mod > 0 ? (row % ep.res_row_mod) : row;
    for (int j = 0; j < 4; ++j) {
      const int col = gcol + j;
      if (col < N) {
        float x = ld_shared_f1(slab + rl * 128 + ((cg ^ (rl & 7)) << 4) + j * 4);
        if (ep.bias) x += __ldg(ep.bias + col);
        if (ep.residual) x += __ldg(ep.residual + (size_t)rr * ep.ldr + col);
        if (ep.residual16) x += __half2float(ep.residual16[(size_t)rr * ep.ldr16 + col]);
        if (ep.act == 1) x = gelu_fast_f(x);
        else if (ep.act == 2) x = silu_f(x);
        if (ep.out_f32) {
          if (ep.atomic) atomicAdd(ep.out_f32 + (size_t)row * ep.ldo_f32 + col, x);
          else ep.out_f32[(size_t)row * ep.ldo_f32 + col] = x;
        }
        if (ep.out16) {
          const uint32_t h2 = pack_f16x2_sat(x, 0.f);
          ep.out16[(size_t)row * ep.ldo16 + col] = *reinterpret_cast<const __half*>(&h2);
        }
        if (ep.out_bf16) ep.out_bf16[(size_t)row * ep.ldo_bf16 + col] = __float2bfloat16(x);
      }
    }
  }
}

// Compile-time epilogue variants: the epilogue warps are instruction-latency bound (two warps per scheduler), so
// the common cases are specialised to strip every runtime flag test and keep address arithmetic out of the loop.
enum EpiKind : int {
  EPI_GENERIC = 0,      // anything (runtime flags)
  EPI_BF16 = 1,         // bias -> bf16                       (QKV / Q / FFN linear2 / text KV)
  EPI_BF16_GELU = 2,    // bias -> GELU -> bf16               (FFN linear1)
  EPI_RES_F32 = 3,      // bias + fp32 residual -> fp32       (block output projection, in place on the stream)
  EPI_RES_F32_BF16 = 4, // ... and a bf16 copy of the stream  (feeds the FFN / output heads)
  EPI_RES_H = 5,        // bias + fp16 residual -> fp16       (product path: the residual stream is stored in fp16)
  EPI_RES_H_BF16 = 6,   // ... and a bf16 copy of the stream
  EPI_FUSED = 7         // generic + the training fusions (out_pre / gate / erf-form GELU); kept apart so the generic kind
                        // — every split-K weight gradient runs it — does not carry their registers and branches
};

// Epilogue of one 32-row x 32-column fp32 chunk owned by one warp.
// tcgen05.ld hands every lane one ROW (32 consecutive columns): storing that straight to global memory makes each
// warp-wide store touch 32 different cache lines.  The chunk is therefore transposed through a per-warp 4 KB
// shared-memory slab (32 rows x 128 B, 16-byte chunks XOR-swizzled with row & 7: conflict-free both for the
// row-owner writes and for the coalesced reads), after which lane l owns columns 4*(l%8)..+3 of rows l/8 + 4 i
// (i = 0..7): bias / residual / activation are applied and stored from that layout, so every global access of a
// warp covers 4 rows x 128 contiguous bytes.  Everything that does not depend on the accumulator (bias and
// residual loads, addresses) is issued before the shared-memory round trip: the epilogue warps are latency-bound.
struct EpiLane {
  uint32_t slab_st;    // shared address of this lane's row in the slab (row-owner writes)
  uint32_t slab_ld[2]; // shared address of (row sub, chunk cg) for even / odd i; row sub + 4 i = + i * 512 B
  int sw, cg, row_first;
  uint32_t row_ok;     // bit i: row_first + 4 i < M
  float* o32;          // element (row_first, cg*4) of each tensor; null when unused
  __nv_bfloat16* o16;
  const float* res;
  __half* oh;
  const __half* resh;
  __nv_bfloat16* pre;
  const __nv_bfloat16* gate;
};

HIG_DEVICE void epi_setup(EpiLane& L, uint32_t slab, const GemmEpilogue& ep, int row0, int M, int lane) {
  const int sub = lane >> 3;
  L.cg = lane & 7;
  L.sw = lane & 7;
  L.slab_st = slab + lane * 128;
  L.slab_ld[0] = slab + sub * 128 + ((L.cg ^ sub) << 4);
  L.slab_ld[1] = slab + sub * 128 + ((L.cg ^ sub ^ 4) << 4);
  L.row_first = row0 + sub;
  L.row_ok = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) L.row_ok |= ((L.row_first + 4 * i < M) ? 1u : 0u) << i;
  const size_t r = (size_t)L.row_first;
  L.o32 = ep.out_f32 ? ep.out_f32 + r * ep.ldo_f32 + L.cg * 4 : nullptr;
  L.o16 = ep.out_bf16 ? ep.out_bf16 + r * ep.ldo_bf16 + L.cg * 4 : nullptr;
  L.res = ep.residual ? ep.residual + r * ep.ldr + L.cg * 4 : nullptr;
  L.oh = ep.out16 ? ep.out16 + r * ep.ldo16 + L.cg * 4 : nullptr;
  L.resh = ep.residual16 ? ep.residual16 + r * ep.ldr16 + L.cg * 4 : nullptr;
  L.pre = ep.out_pre ? ep.out_pre + r * ep.ldo_pre + L.cg * 4 : nullptr;
  L.gate = ep.gate ? ep.gate + r * ep.ld_gate + L.cg * 4 : nullptr;
}

template <int KIND>
HIG_DEVICE void epilogue_chunk(const uint32_t (&r)[32], const EpiLane& L, uint32_t slab, const GemmEpilogue& ep,
                               int row0, int M, int col0, int N, int vec_ok, int lane) {
  constexpr bool GEN = (KIND == EPI_GENERIC || KIND == EPI_FUSED);   // runtime-flag kinds
  const int gcol = col0 + L.cg * 4;
  // warp-uniform: the whole chunk takes the vector path or the scalar one
  const bool vec = (!GEN) || (vec_ok && (col0 + 32 <= N));
  const bool has_res = GEN ? (ep.residual != nullptr) : (KIND >= EPI_RES_F32);
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 rs[8];
  uint2 gt[8];
  if (vec) {
    if (!GEN || ep.bias) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol));
    if (KIND == EPI_FUSED && L.gate) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        gt[i] = make_uint2(0u, 0u);
        if ((L.row_ok >> i) & 1u) gt[i] = __ldg(reinterpret_cast<const uint2*>(L.gate + (size_t)(4 * i) * ep.ld_gate + col0));
      }
    }
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((L.row_ok >> i) & 1u) {
          const float* rp = L.res + (size_t)(4 * i) * ep.ldr + col0;
          if (GEN && ep.res_row_mod > 0)
            rp = ep.residual + (size_t)((L.row_first + 4 * i) % ep.res_row_mod) * ep.ldr + gcol;
          rs[i] = __ldg(reinterpret_cast<const float4*>(rp));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_f4(L.slab_st + ((j ^ L.sw) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  if (vec) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = ld_shared_f4(L.slab_ld[i & 1] + i * 512);
      v[i].x += bias4.x; v[i].y += bias4.y; v[i].z += bias4.z; v[i].w += bias4.w;
    }
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i].x += rs[i].x; v[i].y += rs[i].y; v[i].z += rs[i].z; v[i].w += rs[i].w; }
    }
    if (GEN && L.resh) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((L.row_ok >> i) & 1u) {
          const uint2 h = __ldg(reinterpret_cast<const uint2*>(L.resh + (size_t)(4 * i) * ep.ldr16 + col0));
          const float2 a = unpack_f16x2(h.x), b = unpack_f16x2(h.y);
          v[i].x += a.x; v[i].y += a.y; v[i].z += b.x; v[i].w += b.y;
        }
      }
    }
    if (KIND == EPI_FUSED && L.pre) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint2 h = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        if ((L.row_ok >> i) & 1u) *reinterpret_cast<uint2*>(L.pre + (size_t)(4 * i) * ep.ldo_pre + col0) = h;
        const float2 a = unpack_bf16x2(h.x), b = unpack_bf16x2(h.y);
        v[i] = make_float4(a.x, a.y, b.x, b.y);
      }
    }
    const int act = GEN ? ep.act : (KIND == EPI_BF16_GELU ? 1 : 0);
    if (KIND == EPI_FUSED && act == 3) {        // erf-form GELU (Abramowitz & Stegun), as hig_act_fwd computes it
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = gelu_as_f(v[i].x); v[i].y = gelu_as_f(v[i].y); v[i].z = gelu_as_f(v[i].z); v[i].w = gelu_as_f(v[i].w);
      }
    }
    if (KIND == EPI_FUSED && L.gate) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a = unpack_bf16x2(gt[i].x), b = unpack_bf16x2(gt[i].y);
        v[i].x *= act_grad_f(a.x, ep.gate_act); v[i].y *= act_grad_f(a.y, ep.gate_act);
        v[i].z *= act_grad_f(b.x, ep.gate_act); v[i].w *= act_grad_f(b.y, ep.gate_act);
      }
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = gelu_fast_f(v[i].x); v[i].y = gelu_fast_f(v[i].y); v[i].z = gelu_fast_f(v[i].z); v[i].w = gelu_fast_f(v[i].w);
      }
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = silu_f(v[i].x); v[i].y = silu_f(v[i].y); v[i].z = silu_f(v[i].z); v[i].w = silu_f(v[i].w);
      }
    }
    const bool w32 = GEN ? (L.o32 != nullptr) : (KIND >= EPI_RES_F32);
    const bool w16 = GEN ? (L.o16 != nullptr) : (KIND != EPI_RES_F32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if ((L.row_ok >> i) & 1u) {
        if (GEN && L.oh)
          *reinterpret_cast<uint2*>(L.oh + (size_t)(4 * i) * ep.ldo16 + col0) =
              make_uint2(pack_f16x2_sat(v[i].x, v[i].y), pack_f16x2_sat(v[i].z, v[i].w));
        if (w32) {
          float* o = L.o32 + (size_t)(4 * i) * ep.ldo_f32 + col0;
          if (GEN && ep.atomic) {
            // one 16-byte vector reduction instead of four scalar atomics (split-K weight gradients)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[i].x), "f"(v[i].y), "f"(v[i].z),
                         "f"(v[i].w) : "memory");
          } else {
            *reinterpret_cast<float4*>(o) = v[i];
          }
        }
        if (w16)
          *reinterpret_cast<uint2*>(L.o16 + (size_t)(4 * i) * ep.ldo_bf16 + col0) =
              make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
      }
    }
  } else {
    epilogue_chunk_scalar(slab, ep, row0, M, col0, N, lane);
  }
  __syncwarp();  // the slab is rewritten by the next chunk
}

// ---- split form used by the specialised kinds: global loads one chunk ahead of the accumulator ----------------
struct EpiPre {
  float4 bias4;
  float4 rs[8];   // fp32 residual kinds; the fp16 kinds keep the raw 8-byte loads in rs[i].x / .y
};

template <int KIND>
HIG_DEVICE void epi_prefetch(EpiPre& P, const EpiLane& L, const GemmEpilogue& ep, int col0) {
  P.bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + L.cg * 4));
  if (KIND == EPI_RES_F32 || KIND == EPI_RES_F32_BF16) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      P.rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((L.row_ok >> i) & 1u)
        P.rs[i] = __ldg(reinterpret_cast<const float4*>(L.res + (size_t)(4 * i) * ep.ldr + col0));
    }
  }
  if (KIND == EPI_RES_H || KIND == EPI_RES_H_BF16) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint2 h = make_uint2(0u, 0u);
      if ((L.row_ok >> i) & 1u) h = __ldg(reinterpret_cast<const uint2*>(L.resh + (size_t)(4 * i) * ep.ldr16 + col0));
      P.rs[i].x = __uint_as_float(h.x);
      P.rs[i].y = __uint_as_float(h.y);
    }
  }
}

template <int KIND>
HIG_DEVICE void epi_finish(const uint32_t (&r)[32], const EpiPre& P, const EpiLane& L, const GemmEpilogue& ep,
                           int col0) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_f4(L.slab_st + ((j ^ L.sw) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  __syncwarp();
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = ld_shared_f4(L.slab_ld[i & 1] + i * 512);
    v[i].x += P.bias4.x; v[i].y += P.bias4.y; v[i].z += P.bias4.z; v[i].w += P.bias4.w;
    if (KIND == EPI_RES_F32 || KIND == EPI_RES_F32_BF16) {
      v[i].x += P.rs[i].x; v[i].y += P.rs[i].y; v[i].z += P.rs[i].z; v[i].w += P.rs[i].w;
    }
    if (KIND == EPI_RES_H || KIND == EPI_RES_H_BF16) {
      const float2 a = unpack_f16x2(__float_as_uint(P.rs[i].x)), b = unpack_f16x2(__float_as_uint(P.rs[i].y));
      v[i].x += a.x; v[i].y += a.y; v[i].z += b.x; v[i].w += b.y;
    }
    if (KIND == EPI_BF16_GELU) {
      v[i].x = gelu_fast_f(v[i].x); v[i].y = gelu_fast_f(v[i].y); v[i].z = gelu_fast_f(v[i].z); v[i].w = gelu_fast_f(v[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if ((L.row_ok >> i) & 1u) {
      if (KIND == EPI_RES_F32 || KIND == EPI_RES_F32_BF16)
        *reinterpret_cast<float4*>(L.o32 + (size_t)(4 * i) * ep.ldo_f32 + col0) = v[i];
      if (KIND == EPI_RES_H || KIND == EPI_RES_H_BF16)
        *reinterpret_cast<uint2*>(L.oh + (size_t)(4 * i) * ep.ldo16 + col0) =
            make_uint2(pack_f16x2_sat(v[i].x, v[i].y), pack_f16x2_sat(v[i].z, v[i].w));
      if (KIND != EPI_RES_F32 && KIND != EPI_RES_H)
        *reinterpret_cast<uint2*>(L.o16 + (size_t)(4 * i) * ep.ldo_bf16 + col0) =
            make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
    }
  }
  __syncwarp();  // the slab is rewritten by the next chunk
}

// classify a runtime epilogue into the specialised kinds (host side)
inline int classify_epilogue(const GemmEpilogue& ep, int vec_ok, int N) {
  if (ep.out_pre || ep.gate || ep.act == 3) return EPI_FUSED;
  if (!vec_ok || (N % 32) != 0 || !ep.bias || ep.res_row_mod > 0 || ep.atomic) return EPI_GENERIC;
  if (ep.residual16 || ep.out16) {
    if (ep.residual16 && ep.out16 && !ep.residual && !ep.out_f32 && ep.act == 0)
      return ep.out_bf16 ? EPI_RES_H_BF16 : EPI_RES_H;
    return EPI_GENERIC;
  }
  if (!ep.residual && !ep.out_f32 && ep.out_bf16) {
    if (ep.act == 0) return EPI_BF16;
    if (ep.act == 1) return EPI_BF16_GELU;
  }
  if (ep.residual && ep.out_f32 && ep.act == 0) return ep.out_bf16 ? EPI_RES_F32_BF16 : EPI_RES_F32;
  return EPI_GENERIC;
}

}  // namespace hig
