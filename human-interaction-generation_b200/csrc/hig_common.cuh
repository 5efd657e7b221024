// Shared device helpers for the hig_b200 kernels (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05 / TMEM plus small math utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef HIG_DEVICE
#define HIG_DEVICE __device__ __forceinline__
#endif

namespace hig {

// ----------------------------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------------------------
HIG_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

HIG_DEVICE uint32_t lane_id() { return threadIdx.x & 31u; }

HIG_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

HIG_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
HIG_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

HIG_DEVICE float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// erf via Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7): exact-GELU semantics at a third of erff's instruction
// count; used by the bf16 GEMM epilogue (the fp32-mode kernels call erff).
HIG_DEVICE float erf_as_f(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(e, x);
}
HIG_DEVICE float gelu_as_f(float v) { return 0.5f * v * (1.0f + erf_as_f(v * 0.70710678118654752440f)); }
HIG_DEVICE float tanh_approx_f(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// GELU for the bf16 epilogue: 0.5 v (1 + tanh(sqrt(2/pi) (v + 0.044715 v^3))) on the single-MUFU tanh.approx.
// |difference to the exact erf GELU| <= 5e-4 absolute (well inside one bf16 ulp of the stored result); the
// fp32-mode kernels keep erff.  The GEMM epilogue is issue-bound, and erff costs ~4x the instructions.
HIG_DEVICE float gelu_fast_f(float v) {
  const float u = v * fmaf(0.044715f * v, v, 1.0f);
  return 0.5f * v * (1.0f + tanh_approx_f(0.7978845608028654f * u));
}
// exact (erf) GELU, as torch.nn.GELU() default
HIG_DEVICE float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// activations of the training kernels and their derivatives (act: 1 GELU erf-form, 2 SiLU, 3 QuickGELU)
HIG_DEVICE float act_fwd_f(float v, int act) {
  if (act == 1) return gelu_as_f(v);                     // erf to 1.5e-7 absolute (Abramowitz & Stegun 7.1.26): exact-GELU semantics
  if (act == 2) return __fdividef(v, 1.0f + __expf(-v));
  if (act == 3) return v / (1.0f + expf(-1.702f * v));   // QuickGELU of CLIP's text transformer MLP: x sigmoid(1.702 x)
  return v;
}
HIG_DEVICE float act_grad_f(float v, int act) {
  if (act == 1) {
    // the kernel was instruction-bound on erff + expf (84 % issue-active, 40 us for 143 MB): A&S erf + fast exp
    const float cdf = 0.5f * (1.0f + erf_as_f(v * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
    return fmaf(v, pdf, cdf);
  }
  if (act == 2) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-v));
    return s * fmaf(v, 1.0f - s, 1.0f);
  }
  return 1.0f;
}

HIG_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
HIG_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// ----------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) for the instruction-issue-bound epilogues and attention tails
// ----------------------------------------------------------------------------------------------
HIG_DEVICE uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
HIG_DEVICE uint64_t f2_pack_u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
HIG_DEVICE void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
HIG_DEVICE uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
HIG_DEVICE uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
HIG_DEVICE uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x, flush-to-zero: one MUFU, none of the denormal range scaling __expf / exp2f wrap around it
HIG_DEVICE float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// (prologue: barrier init, TMEM allocation, tensor-map prefetch, parameter staging) while its predecessor drains;
// pdl_wait() blocks until the predecessor grid has completed and its writes are visible.  Both are no-ops for
// kernels launched without the attribute.
// ----------------------------------------------------------------------------------------------
HIG_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
HIG_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
HIG_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
HIG_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
HIG_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
HIG_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
HIG_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
HIG_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
HIG_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D tiled load, completes on an mbarrier
// ----------------------------------------------------------------------------------------------
HIG_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
HIG_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
HIG_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
HIG_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
HIG_DEVICE void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
HIG_DEVICE void tmem_dealloc(uint32_t taddr) {  // whole warp, same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16: bf16/fp16 inputs, fp32 accumulate)
HIG_DEVICE void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
HIG_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
HIG_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
HIG_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile stored as [rows][64] (128 B per row) with the
// 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are 1024 B apart (SBO), the
// leading-dimension offset is unused for swizzled K-major layouts, descriptor version 1 (Blackwell).
HIG_DEVICE uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                        // LBO (ignored for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // SBO = 1024 B, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                        // version = 1
  d |= static_cast<uint64_t>(2) << 61;                        // layout = SWIZZLE_128B
  return d;
}

// MN-major operand tile, 128B swizzle: [k rows][64 MN elements = 128 B] boxes as TMA writes them; 8-row k groups are
// 1024 B apart (SBO), consecutive 64-element MN blocks `mn_block_bytes` apart (LBO).  Canonical layout (16-byte units):
// ((8,n),(8,k)) : ((1,LBO),(8,SBO)).
HIG_DEVICE uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t mn_block_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((mn_block_bytes >> 4) & 0x3FFFu) << 16;   // LBO
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                         // SBO
  d |= static_cast<uint64_t>(1) << 46;                                  // version = 1
  d |= static_cast<uint64_t>(2) << 61;                                  // layout = SWIZZLE_128B
  return d;
}

// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, shape M x N.
HIG_DEVICE constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)            // D format = F32
         | (1u << 7)          // A format = BF16
         | (1u << 10)         // B format = BF16
         | ((N >> 3) << 17)   // N / 8
         | ((M >> 4) << 24);  // M / 16
}


// ----------------------------------------------------------------------------------------------
// thread-block clusters / CTA pairs (cta_group::2)
// ----------------------------------------------------------------------------------------------
HIG_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
HIG_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of this cluster
HIG_DEVICE void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// 2-CTA TMA load: lands in this CTA's smem, completes tx bytes on the LEADER CTA's mbarrier (peer bit cleared)
HIG_DEVICE void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
      "r"(c1)
      : "memory");
}
template <uint32_t kCols>
HIG_DEVICE void tmem_alloc_2cta(uint32_t* smem_result) {  // one warp in EACH CTA of the pair, same warp index
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
HIG_DEVICE void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA's smem] * B[N x 16: N/2 rows from each CTA's smem]
HIG_DEVICE void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once) on the mbarrier at this smem offset in every CTA of `mask` when the issued MMAs have completed
HIG_DEVICE void umma_commit_2cta_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

}  // namespace hig
