// Shared epilogue of the tcgen05 GEMM kernels: TMEM -> registers -> swizzled smem slab -> coalesced global I/O.
#pragma once
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
};

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
static __device__ __noinline__ void epilogue_half_scalar(uint32_t slab, GemmEpilogue ep, int row0, int M, int col0, int N,
                                                  int lane) {
  const int sub = lane >> 2, cg = lane & 3;
  const int gcol = col0 + cg * 4;
  for (int i = 0; i < 4; ++i) {
    const int rl = i * 8 + sub;
    const int row = row0 + rl;
    if (row >= M) continue;
    const int rr = ep.res_row_mod > 0 ? (row % ep.res_row_mod) : row;
    for (int j = 0; j < 4; ++j) {
      const int col = gcol + j;
      if (col < N) {
        float x = ld_shared_f1(slab + (rl * 16 + ((cg ^ ((rl >> 1) & 3)) << 2) + j) * 4);
        if (ep.bias) x += __ldg(ep.bias + col);
        if (ep.residual) x += __ldg(ep.residual + (size_t)rr * ep.ldr + col);
        if (ep.act == 1) x = gelu_fast_f(x);
        else if (ep.act == 2) x = silu_f(x);
        if (ep.out_f32) ep.out_f32[(size_t)row * ep.ldo_f32 + col] = x;
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
  EPI_RES_F32_BF16 = 4  // ... and a bf16 copy of the stream  (feeds the FFN / output heads)
};

// Per-warp, per-tile constants hoisted out of the chunk loop.  Lane l works on rows sub + 8 i (i = 0..3) and the
// 16-byte column group cg of every half-chunk; its four rows are 8 rows apart, so one base pointer per tensor plus a
// constant stride reaches all of them.  (rl >> 1) & 3 is the same for the four rows, hence one swizzled slab address.
struct EpiLane {
  uint32_t slab_st;    // shared address of this lane's row in the slab (row-owner writes)
  uint32_t slab_ld0;   // shared address of (row sub, chunk cg) in the coalesced layout; row i*8+sub = + i*512 B
  int sw, cg, row_first;
  uint32_t row_ok;     // bit i: row_first + 8 i < M
  float* o32;          // element (row_first, cg*4) of each tensor; null when unused
  __nv_bfloat16* o16;
  const float* res;
};

HIG_DEVICE void epi_setup(EpiLane& L, uint32_t slab, const GemmEpilogue& ep, int row0, int M, int lane) {
  const int sub = lane >> 2;
  L.cg = lane & 3;
  L.sw = (lane >> 1) & 3;
  L.slab_st = slab + lane * 64;
  L.slab_ld0 = slab + (sub * 16 + ((L.cg ^ ((sub >> 1) & 3)) << 2)) * 4;
  L.row_first = row0 + sub;
  L.row_ok = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) L.row_ok |= ((L.row_first + 8 * i < M) ? 1u : 0u) << i;
  const size_t r = (size_t)L.row_first;
  L.o32 = ep.out_f32 ? ep.out_f32 + r * ep.ldo_f32 + L.cg * 4 : nullptr;
  L.o16 = ep.out_bf16 ? ep.out_bf16 + r * ep.ldo_bf16 + L.cg * 4 : nullptr;
  L.res = ep.residual ? ep.residual + r * ep.ldr + L.cg * 4 : nullptr;
}

// One 32-row x 16-column half-chunk: row-owner registers -> swizzled slab -> coalesced layout -> global.
// col0 is the global column of the half-chunk's first column.
template <int KIND>
HIG_DEVICE void epilogue_half(const uint32_t* r16, const EpiLane& L, uint32_t slab, const GemmEpilogue& ep, int row0,
                              int M, int col0, int N, int vec_ok, int lane) {
  const int gcol = col0 + L.cg * 4;
  const bool vec = (KIND != EPI_GENERIC) || (vec_ok && (gcol + 4 <= N));
  const bool has_res = (KIND == EPI_GENERIC) ? (ep.residual != nullptr) : (KIND >= EPI_RES_F32);
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 rs[4];
  if (vec) {
    if (KIND != EPI_GENERIC || ep.bias) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol));
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((L.row_ok >> i) & 1u) {
          const float* rp = L.res + (size_t)(8 * i) * ep.ldr + col0;
          if (KIND == EPI_GENERIC && ep.res_row_mod > 0)
            rp = ep.residual + (size_t)((L.row_first + 8 * i) % ep.res_row_mod) * ep.ldr + gcol;
          rs[i] = __ldg(reinterpret_cast<const float4*>(rp));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_shared_f4(L.slab_st + ((j ^ L.sw) << 4), __uint_as_float(r16[4 * j]), __uint_as_float(r16[4 * j + 1]),
                 __uint_as_float(r16[4 * j + 2]), __uint_as_float(r16[4 * j + 3]));
  __syncwarp();
  if (vec) {
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = ld_shared_f4(L.slab_ld0 + i * 512);
      v[i].x += bias4.x; v[i].y += bias4.y; v[i].z += bias4.z; v[i].w += bias4.w;
    }
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i].x += rs[i].x; v[i].y += rs[i].y; v[i].z += rs[i].z; v[i].w += rs[i].w; }
    }
    const int act = (KIND == EPI_GENERIC) ? ep.act : (KIND == EPI_BF16_GELU ? 1 : 0);
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i].x = gelu_fast_f(v[i].x); v[i].y = gelu_fast_f(v[i].y); v[i].z = gelu_fast_f(v[i].z); v[i].w = gelu_fast_f(v[i].w);
      }
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i].x = silu_f(v[i].x); v[i].y = silu_f(v[i].y); v[i].z = silu_f(v[i].z); v[i].w = silu_f(v[i].w);
      }
    }
    const bool w32 = (KIND == EPI_GENERIC) ? (L.o32 != nullptr) : (KIND >= EPI_RES_F32);
    const bool w16 = (KIND == EPI_GENERIC) ? (L.o16 != nullptr) : (KIND != EPI_RES_F32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((L.row_ok >> i) & 1u) {
        if (w32) *reinterpret_cast<float4*>(L.o32 + (size_t)(8 * i) * ep.ldo_f32 + col0) = v[i];
        if (w16)
          *reinterpret_cast<uint2*>(L.o16 + (size_t)(8 * i) * ep.ldo_bf16 + col0) =
              make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
      }
    }
  } else {
    epilogue_half_scalar(slab, ep, row0, M, col0, N, lane);
  }
  __syncwarp();  // the slab is rewritten by the next half-chunk
}

template <int KIND>
HIG_DEVICE void epilogue_chunk(const uint32_t (&r)[32], const EpiLane& L, uint32_t slab, const GemmEpilogue& ep,
                               int row0, int M, int col0, int N, int vec_ok, int lane) {
  epilogue_half<KIND>(&r[0], L, slab, ep, row0, M, col0, N, vec_ok, lane);
  if (col0 + 16 < N) epilogue_half<KIND>(&r[16], L, slab, ep, row0, M, col0 + 16, N, vec_ok, lane);
}

// classify a runtime epilogue into the specialised kinds (host side)
inline int classify_epilogue(const GemmEpilogue& ep, int vec_ok, int N) {
  if (!vec_ok || (N % 16) != 0 || !ep.bias || ep.res_row_mod > 0) return EPI_GENERIC;
  if (!ep.residual && !ep.out_f32 && ep.out_bf16) {
    if (ep.act == 0) return EPI_BF16;
    if (ep.act == 1) return EPI_BF16_GELU;
  }
  if (ep.residual && ep.out_f32 && ep.act == 0) return ep.out_bf16 ? EPI_RES_F32_BF16 : EPI_RES_F32;
  return EPI_GENERIC;
}

}  // namespace hig
