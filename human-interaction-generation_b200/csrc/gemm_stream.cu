// hig_gemm_stream: the CTA-pair tcgen05 GEMM of gemm2_tcgen05.cu with a TMA-staged epilogue, for the four
// projections that make up 95 % of a denoiser step.
//
// Why a second epilogue: with K = 512 a 256 x 256 tile holds only 32 MMAs (4096 tensor-pipe clocks), and the
// register-transposing epilogue of gemm_epilogue.cuh needs ~8000 clocks per tile (measured: QKV projection 36.0 us
// with it, 27.6 us with the epilogue math disabled) — the GEMMs were epilogue-bound.  Here every epilogue lane keeps
// the accumulator row tcgen05.ld hands it, writes its packed 2-byte outputs into a 128B-swizzled staging slab
// (conflict-free 16-byte stores) and one elected lane issues a TMA store of the 32 x 64 box; residual tiles arrive
// the same way (TMA load into the slab, updated in place).  No transposition, no per-lane global addressing, and the
// accumulator buffer is handed back to the MMA warp as soon as the last tcgen05.ld has landed.
//
// Kinds (reference lines: models/interaction_transformer.py):
//   ST_BF16       out = acc + bias                                  -> bf16     (FFN linear2 :263, text-CA query :153)
//   ST_BF16_GELU  out = GELU(acc + bias)                            -> bf16     (FFN linear1 :262)
//   ST_RES_H      x   = x + acc + bias, fp16 residual stream in place (+ per-row sum / sum-of-squares partials)
//                                                                               (StylizationBlock out_layers + residual :97,129,164,203,263)
//   ST_LN_BF16    out = rstd_m (acc - mu_m wsum_n) + bias_n         -> bf16     (pre-attention LayerNorm folded into the
//                 Q/K/V projections :119-121,153,190-194: A is the raw fp16 stream, W = gamma o W_qkv in fp16,
//                 wsum_n = sum_k W_nk, bias_n = b_n + sum_k beta_k Wqkv_nk, (mu, rstd) from the row partials the
//                 previous ST_RES_H epilogue left behind — the three LayerNorm kernels per layer disappear)
//   ST_F16        out = acc + bias                                  -> fp16     (output heads out / out2 :613-616 reading the
//                 fp16 stream directly; resident-W kernel only)
//   ST_LN_QSM     ST_LN_BF16 whose first 512 output columns — the query block of a Q / Q|K|V projection — are replaced by
//                 softmax over each head's 64 features (:120,156,195 `F.softmax(query, dim=-1)`): one accumulator row
//                 chunk pair is exactly one head of one row, so the softmax is lane-local in the epilogue and the
//                 attention-apply kernel reads ready-made MMA operands (resident-W kernel only)
// Operands are bf16 x bf16 or fp16 x fp16 (kind::f16 instruction descriptor formats), fp32 accumulate.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <string>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int GS_BM = 128;       // rows of A per CTA (256 per pair)
constexpr int GS_BN = 256;       // tile columns (128 rows of W per CTA)
constexpr int GS_BK = 64;
constexpr int GS_STAGES = 5;
constexpr int GS_EPI_WARPS = 8;
constexpr int GS_THREADS = 64 + 32 * GS_EPI_WARPS;
constexpr int GS_A_BYTES = GS_BM * GS_BK * 2;          // 16 KB
constexpr int GS_B_BYTES = (GS_BN / 2) * GS_BK * 2;    // 16 KB
constexpr int GS_SLAB = 32 * 128;                      // 32 rows x 64 two-byte columns
constexpr int GS_EPI_BYTES = GS_EPI_WARPS * 2 * GS_SLAB;  // two slabs per epilogue warp: 64 KB
constexpr int GS_BAR_BYTES = (2 * GS_STAGES + 4 + 2 * GS_EPI_WARPS) * 8 + 16;
constexpr int GS_SMEM = GS_STAGES * (GS_A_BYTES + GS_B_BYTES) + GS_EPI_BYTES + GS_BAR_BYTES + 1024;
static_assert(GS_SMEM <= 232448, "shared memory budget");

enum StreamKind : int { ST_BF16 = 0, ST_BF16_GELU = 1, ST_RES_H = 2, ST_LN_BF16 = 3, ST_F16 = 4, ST_LN_QSM = 5 };
constexpr int QSM_COLS = 512;   // ST_LN_QSM: output columns [0, 512) are the query block (8 heads x 64 features)

struct StreamEpi {
  const float* bias;       // [N]
  const float* wsum;       // [N]            ST_LN_BF16
  const float* stats_in;   // [M, 8] fp32    ST_LN_BF16: 4 x (sum, sum of squares) partials per row
  float* stats_out;        // [M, 8] or null ST_RES_H
  float inv_width;         // 1 / (LayerNorm width)
  float ln_eps;
  unsigned long long* sat_count;   // debug (hig_debug_saturation): counts fp16 stream stores that hit +-65504; normally null
  int gelu_erf;            // ST_BF16_GELU: 1 = erf-form GELU (|err| <= 1.5e-7 vs torch's exact GELU), 0 = tanh form (<= 5e-4)
};

HIG_DEVICE void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
HIG_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
HIG_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
HIG_DEVICE void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

HIG_DEVICE void st_shared_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
HIG_DEVICE uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
HIG_DEVICE uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
HIG_DEVICE float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// instruction descriptor, kind::f16: D = F32, A/B format 0 = F16, 1 = BF16, both K-major
HIG_DEVICE uint32_t umma_idesc_f16kind(uint32_t M, uint32_t N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// One 32-column chunk of this lane's accumulator row -> 64 bytes of the staging slab (16-byte chunks cb..cb+3 of the
// row, XOR-swizzled with row & 7 exactly as CU_TENSOR_MAP_SWIZZLE_128B expects).
template <int KIND>
HIG_DEVICE void stream_chunk(const uint32_t (&r)[32], uint32_t slab_row, int cb, int sw, const StreamEpi& ep, int col0,
                             float rstd, float nmr, float& s1, float& s2) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {  // 8 columns -> one 16-byte chunk
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * g));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * g + 4));
    float v[8];
    v[0] = __uint_as_float(r[8 * g + 0]); v[1] = __uint_as_float(r[8 * g + 1]);
    v[2] = __uint_as_float(r[8 * g + 2]); v[3] = __uint_as_float(r[8 * g + 3]);
    v[4] = __uint_as_float(r[8 * g + 4]); v[5] = __uint_as_float(r[8 * g + 5]);
    v[6] = __uint_as_float(r[8 * g + 6]); v[7] = __uint_as_float(r[8 * g + 7]);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const uint32_t addr = slab_row + (((cb + g) ^ sw) << 4);
    if (KIND == ST_LN_BF16) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(ep.wsum + col0 + 8 * g));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(ep.wsum + col0 + 8 * g + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(rstd, v[i], fmaf(nmr, ww[i], bb[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += bb[i];
    }
    if (KIND == ST_BF16_GELU) {
      if (ep.gelu_erf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_as_f(v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_fast_f(v[i]);
      }
    }
    if (KIND == ST_RES_H) {
      const uint4 h = ld_shared_u4(addr);
      const float2 a = unpack_h2(h.x), b = unpack_h2(h.y), c = unpack_h2(h.z), d = unpack_h2(h.w);
      v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
      st_shared_u4(addr, pack_h2_sat(v[0], v[1]), pack_h2_sat(v[2], v[3]), pack_h2_sat(v[4], v[5]), pack_h2_sat(v[6], v[7]));
      if (ep.sat_count != nullptr) {
        int n = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) n += fabsf(v[i]) > 65504.f;
        if (n) atomicAdd(ep.sat_count, (unsigned long long)n);
      }
    } else {
      st_shared_u4(addr, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GS_THREADS, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, int M, int N, int K, StreamEpi ep, int f16_ops) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + GS_STAGES * GS_A_BYTES;
  uint8_t* sEpi = sB + GS_STAGES * GS_B_BYTES;   // 1024-byte aligned: 5 x 32 KB above the aligned base
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + GS_EPI_BYTES);
  uint64_t* empty_bar = full_bar + GS_STAGES;
  uint64_t* tfull_bar = empty_bar + GS_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;            // [epilogue warp][slab]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * GS_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int m_tiles = (M + 2 * GS_BM - 1) / (2 * GS_BM);
  const int n_tiles = (N + GS_BN - 1) / GS_BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + GS_BK - 1) / GS_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < GS_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, 2 * GS_EPI_WARPS);  // every epilogue warp of BOTH CTAs arrives on the leader's
    }
    for (int s = 0; s < 2 * GS_EPI_WARPS; ++s) mbar_init(res_bar + s, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      // W does not depend on the previous kernel: the W halves of the first ring slots are requested BEFORE
      // griddepcontrol.wait, so they travel while the predecessor's last CTAs drain.  A (and everything the epilogue
      // touches) waits.
      const int pre = pair < num_tiles ? min(GS_STAGES, k_blocks) : 0;
      for (int s = 0; s < pre; ++s) {
        if (rank == 0) mbar_arrive_expect_tx(full_bar + s, 2 * (GS_A_BYTES + GS_B_BYTES));
        tma_load_2d_2cta(sB + s * GS_B_BYTES, &tmB, full_bar + s, s * GS_BK,
                         (pair % n_tiles) * GS_BN + (int)rank * (GS_BN / 2));
      }
      pdl_wait();
      pdl_trigger();
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_blk = tile % n_tiles;
        const int m_blk = tile / n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % GS_STAGES;
          const uint32_t phase = (it / GS_STAGES) & 1u;
          if (it >= pre) {
            mbar_wait(empty_bar + stage, phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * (GS_A_BYTES + GS_B_BYTES));
            tma_load_2d_2cta(sB + stage * GS_B_BYTES, &tmB, full_bar + stage, kb * GS_BK,
                             n_blk * GS_BN + (int)rank * (GS_BN / 2));
          }
          tma_load_2d_2cta(sA + stage * GS_A_BYTES, &tmA, full_bar + stage, kb * GS_BK,
                           m_blk * 2 * GS_BM + (int)rank * GS_BM);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && lane == 0) {   // touches no global memory: ordered behind the producer through the barriers
      const uint32_t idesc = umma_idesc_f16kind(2 * GS_BM, GS_BN, f16_ops ? 0u : 1u);
      uint32_t it = 0, lt = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
        const uint32_t as = lt & 1u;
        const uint32_t aphase = (lt >> 1) & 1u;
        mbar_wait(tempty_bar + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GS_BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % GS_STAGES;
          const uint32_t phase = (it / GS_STAGES) & 1u;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * GS_A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * GS_B_BYTES));
#pragma unroll
          for (int k = 0; k < GS_BK / 16; ++k)
            umma_f16_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(empty_bar + stage, 0b11);  // frees this smem slot in both CTAs
        }
        umma_commit_2cta_mc(tfull_bar + as, 0b11);  // accumulator halves ready in both CTAs
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..9, both CTAs) =================
    pdl_wait();   // statistics / residual we read and the output we overwrite belong to the previous kernels
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int ch = ew >> 2;          // column half of the tile
    uint8_t* slab0 = sEpi + ew * 2 * GS_SLAB;
    uint8_t* slab1 = slab0 + GS_SLAB;
    uint64_t* rbar = res_bar + 2 * ew;
    const int sw = lane & 7;
    const uint32_t row_s0 = smem_u32(slab0) + lane * 128;
    const uint32_t row_s1 = smem_u32(slab1) + lane * 128;
    uint32_t lt = 0;
    // Staging-slab discipline (per warp): slab 0 carries columns 0..63 of the warp's half, slab 1 columns 64..127, and
    // every half commits exactly one bulk group (possibly empty), so "at most one group pending" always means "the
    // store issued from the OTHER slab half a tile ago has been read out" — no wait ever targets a store just issued.
    if (KIND == ST_RES_H) {
      if (lane == 0 && pair < num_tiles) {
        const int n_blk = pair % n_tiles, m_blk = pair / n_tiles;
        const int row0 = m_blk * 2 * GS_BM + (int)rank * GS_BM + q * 32, gc0 = n_blk * GS_BN + ch * (GS_BN / 2);
        if (gc0 < N) { mbar_arrive_expect_tx(rbar, GS_SLAB); tma_load_2d(slab0, &tmC, rbar, gc0, row0); }
      }
    }
    uint32_t rph0 = 0, rph1 = 0;   // residual-barrier phases advance only on tiles that used the slab
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int n_blk = tile % n_tiles;
      const int m_blk = tile / n_tiles;
      const uint32_t as = lt & 1u;
      const uint32_t aphase = (lt >> 1) & 1u;
      const int row0 = m_blk * 2 * GS_BM + (int)rank * GS_BM + q * 32;
      const int gc0 = n_blk * GS_BN + ch * (GS_BN / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GS_BN + ch * (GS_BN / 2);
      const bool have0 = gc0 < N, have1 = gc0 + 64 < N;   // N % 64 == 0 (host-checked)
      float rstd = 0.f, nmr = 0.f, s1 = 0.f, s2 = 0.f;
      if (KIND == ST_LN_BF16) {
        const int row = min(row0 + lane, M - 1);
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(ep.stats_in + (size_t)row * 8));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(ep.stats_in + (size_t)row * 8 + 4));
        const float sum = (p0.x + p0.z) + (p1.x + p1.z);
        const float ssq = (p0.y + p0.w) + (p1.y + p1.w);
        const float mu = sum * ep.inv_width;
        const float var = fmaxf(fmaf(ssq, ep.inv_width, -mu * mu), 0.f);
        rstd = rsqrtf(var + ep.ln_eps);
        nmr = -mu * rstd;
      }

      mbar_wait(tfull_bar + as, aphase);
      tc_fence_after();
      uint32_t ra[32], rb[32];
      // ---- columns 0..63 of this warp's half -> slab 0
      tmem_ld_32x32(taddr, ra);
      tmem_ld_32x32(taddr + 32, rb);
      tmem_ld_wait();
      if (KIND == ST_RES_H) {
        if (have0) { mbar_wait(rbar, rph0); rph0 ^= 1u; }     // residual box landed (fetched half a tile ago)
      } else {
        if (lane == 0) bulk_wait_read<1>();                    // slab 0's previous store has been read out
        __syncwarp();
      }
      if (have0) {
        stream_chunk<KIND>(ra, row_s0, 0, sw, ep, gc0, rstd, nmr, s1, s2);
        stream_chunk<KIND>(rb, row_s0, 4, sw, ep, gc0 + 32, rstd, nmr, s1, s2);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) {
        if (have0) tma_store_2d(&tmC, slab0, gc0, row0);
        bulk_commit();
        if (KIND == ST_RES_H) {   // slab 1's previous store is the older pending group: fetch this tile's second residual box
          bulk_wait_read<1>();
          if (have1) { mbar_arrive_expect_tx(rbar + 1, GS_SLAB); tma_load_2d(slab1, &tmC, rbar + 1, gc0 + 64, row0); }
        }
      }
      // ---- columns 64..127 -> slab 1; the accumulator buffer goes back to the MMA warp once they are in registers
      tmem_ld_32x32(taddr + 64, ra);
      tmem_ld_32x32(taddr + 96, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
      if (KIND == ST_RES_H) {
        if (have1) { mbar_wait(rbar + 1, rph1); rph1 ^= 1u; }
      } else {
        if (lane == 0) bulk_wait_read<1>();                    // slab 1's previous store has been read out
        __syncwarp();
      }
      if (have1) {
        stream_chunk<KIND>(ra, row_s1, 0, sw, ep, gc0 + 64, rstd, nmr, s1, s2);
        stream_chunk<KIND>(rb, row_s1, 4, sw, ep, gc0 + 96, rstd, nmr, s1, s2);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) {
        if (have1) tma_store_2d(&tmC, slab1, gc0 + 64, row0);
        bulk_commit();
        if (KIND == ST_RES_H) {   // slab 0's store is now the older group: prefetch the NEXT tile's first residual box
          const int nt = tile + num_pairs;
          if (nt < num_tiles) {
            bulk_wait_read<1>();
            const int nn = nt % n_tiles, nm = nt / n_tiles;
            const int nrow0 = nm * 2 * GS_BM + (int)rank * GS_BM + q * 32, ngc0 = nn * GS_BN + ch * (GS_BN / 2);
            if (ngc0 < N) { mbar_arrive_expect_tx(rbar, GS_SLAB); tma_load_2d(slab0, &tmC, rbar, ngc0, nrow0); }
          }
        }
      }
      if (KIND == ST_RES_H) {
        // deterministic row statistics: partial (n_blk, ch) of row (sum, sum of squares) — summed by the consumer
        if (ep.stats_out != nullptr && row0 + lane < M && have0)
          *reinterpret_cast<float2*>(ep.stats_out + (size_t)(row0 + lane) * 8 + (n_blk * 2 + ch) * 2) = make_float2(s1, s2);
      }
    }
    if (lane == 0) bulk_wait<0>();   // all stores of this warp have completed before the CTA may retire
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

// ======================================================================================================================
// Resident-W variant (gemm_wres_kernel): the same CTA-pair MMA, but the pair keeps its 256 rows of W (all of K <= 512,
// 128 KB per CTA) in shared memory for a whole run of row tiles and streams only A.
//
// Why: the L2 delivers ~6.3 KB/clk chip-wide (B300_MICROARCH.md, "LTS throughput cap"), i.e. ~43 B/clk per SM, while a
// 256 x 256 tile with both operands streamed needs (256 + 256) x 16 x 2 B per 128-clk MMA = 64 B/clk per SM before the
// epilogue moves a byte — the streamed kernel above tops out at ~63 % tensor-pipe active on the QKV projection
// (profiles/r01c_ncu_full_summary.txt).  With W resident the operand stream is 32 B/clk per SM.  Tiles are ordered
// column-block-major and cut into one contiguous range per pair, so a pair reloads W at most once or twice per launch.
//
// Shared memory: W 128 KB + 4 A stages x 16 KB + 8 epilogue warps x 2 sub-slabs x 2 KB (32 rows x 32 columns, 64-byte
// swizzle) = 224 KB.  The fp16 residual of ST_RES_H does not fit a landing slab any more: each lane reads its own row
// with 32-byte loads (one full sector per lane per instruction), prefetched half a tile ahead in registers.
// ======================================================================================================================
constexpr int WR_STAGES = 4;
constexpr int WR_KB = 8;                                  // K <= 512
constexpr int WR_W_BYTES = WR_KB * GS_B_BYTES;            // 128 KB
constexpr int WR_SUB = 32 * 64;                           // 2 KB sub-slab
constexpr int WR_EPI_BYTES = GS_EPI_WARPS * 2 * WR_SUB;   // 32 KB
constexpr int WR_BAR_BYTES = (2 * WR_STAGES + 4 + 2) * 8 + 16;
constexpr int WR_SMEM = WR_W_BYTES + WR_STAGES * GS_A_BYTES + WR_EPI_BYTES + WR_BAR_BYTES + 1024;
static_assert(WR_SMEM <= 232448, "shared memory budget");

// tanh-form GELU on a pair (same formula as gelu_fast_f): v (0.5 + 0.5 tanh(v (c0 + c1 v^2)))
HIG_DEVICE uint64_t f2_gelu(uint64_t v) {
  const uint64_t c0 = f2_pack(0.7978845608028654f, 0.7978845608028654f);
  const uint64_t c1 = f2_pack(0.7978845608028654f * 0.044715f, 0.7978845608028654f * 0.044715f);
  const uint64_t hf = f2_pack(0.5f, 0.5f);
  const uint64_t arg = f2_mul(f2_fma(c1, f2_mul(v, v), c0), v);
  float a0, a1;
  f2_unpack(arg, a0, a1);
  const uint64_t th = f2_pack(tanh_approx_f(a0), tanh_approx_f(a1));
  const uint64_t hv = f2_mul(v, hf);
  return f2_fma(hv, th, hv);
}

// 64 fp16 columns of this lane's row of the residual stream -> 32 registers (4 x 32-byte loads, L1 no-allocate)
HIG_DEVICE void wr_load_res(const __half* p, bool ok, uint32_t (&r)[32]) {
  if (ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[8 * i + 0]), "=r"(r[8 * i + 1]), "=r"(r[8 * i + 2]), "=r"(r[8 * i + 3]), "=r"(r[8 * i + 4]),
                     "=r"(r[8 * i + 5]), "=r"(r[8 * i + 6]), "=r"(r[8 * i + 7])
                   : "l"(p + 16 * i)
                   : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0u;
  }
}

// One 32-column chunk of this lane's accumulator row -> 64 bytes of a sub-slab (16-byte chunks 0..3 of the row,
// XOR-swizzled with (row >> 1) & 3 as CU_TENSOR_MAP_SWIZZLE_64B expects).  res: 16 registers = the 32 fp16 residuals.
// The 64 accumulator columns col0 .. col0+63 of this lane's row (one head of the query block) -> in place, as fp32 bit
// patterns: softmax over the 64 features of LayerNorm-folded projection values.  Same arithmetic as the attention
// kernels' feature softmax: exp(x - m) = ex2(x log2e - m log2e), sum of the unrounded exponentials, one reciprocal.
HIG_DEVICE void wres_softmax64(uint32_t (&ra)[32], uint32_t (&rb)[32], const StreamEpi& ep, int col0, float rstd, float nmr) {
  const float kL2E = 1.4426950408889634f;
  float m = -INFINITY;
  auto affine = [&](uint32_t (&r)[32], int c0) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + c0 + 4 * g));
      const float4 w = __ldg(reinterpret_cast<const float4*>(ep.wsum + c0 + 4 * g));
      const float v0 = fmaf(rstd, __uint_as_float(r[4 * g + 0]), fmaf(nmr, w.x, b.x));
      const float v1 = fmaf(rstd, __uint_as_float(r[4 * g + 1]), fmaf(nmr, w.y, b.y));
      const float v2 = fmaf(rstd, __uint_as_float(r[4 * g + 2]), fmaf(nmr, w.z, b.z));
      const float v3 = fmaf(rstd, __uint_as_float(r[4 * g + 3]), fmaf(nmr, w.w, b.w));
      m = fmaxf(m, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));
      r[4 * g + 0] = __float_as_uint(v0); r[4 * g + 1] = __float_as_uint(v1);
      r[4 * g + 2] = __float_as_uint(v2); r[4 * g + 3] = __float_as_uint(v3);
    }
  };
  affine(ra, col0);
  affine(rb, col0 + 32);
  const float nml = -m * kL2E;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  auto expo = [&](uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float e0 = ex2_ftz(fmaf(__uint_as_float(r[i + 0]), kL2E, nml)), e1 = ex2_ftz(fmaf(__uint_as_float(r[i + 1]), kL2E, nml));
      const float e2 = ex2_ftz(fmaf(__uint_as_float(r[i + 2]), kL2E, nml)), e3 = ex2_ftz(fmaf(__uint_as_float(r[i + 3]), kL2E, nml));
      s0 += e0; s1 += e1; s2 += e2; s3 += e3;
      r[i + 0] = __float_as_uint(e0); r[i + 1] = __float_as_uint(e1); r[i + 2] = __float_as_uint(e2); r[i + 3] = __float_as_uint(e3);
    }
  };
  expo(ra);
  expo(rb);
  const float inv = 1.0f / ((s0 + s1) + (s2 + s3));
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    ra[i] = __float_as_uint(__uint_as_float(ra[i]) * inv);
    rb[i] = __float_as_uint(__uint_as_float(rb[i]) * inv);
  }
}

// final_vals: r already holds the finished fp32 outputs (query-block softmax) — pack and store only
template <int KIND, int ROFF>
HIG_DEVICE void wres_chunk(const uint32_t (&r)[32], const uint32_t (&res)[32], uint32_t sub_row, int sw2, const StreamEpi& ep,
                           int col0, uint64_t rstd2, uint64_t nmr2, uint64_t& s1, uint64_t& s2, bool final_vals = false) {
  constexpr bool IS_LN = KIND == ST_LN_BF16 || KIND == ST_LN_QSM;
  if (KIND == ST_LN_QSM && final_vals) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      st_shared_u4(sub_row + ((g ^ sw2) << 4),
                   pack_bf16x2(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])),
                   pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                   pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])),
                   pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
    return;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {  // 8 columns -> one 16-byte chunk
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * g));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * g + 4));
    uint64_t v[4], bb[4];
    bb[0] = f2_pack(b0.x, b0.y); bb[1] = f2_pack(b0.z, b0.w); bb[2] = f2_pack(b1.x, b1.y); bb[3] = f2_pack(b1.z, b1.w);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = f2_pack_u(r[8 * g + 2 * i], r[8 * g + 2 * i + 1]);
    if (IS_LN) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(ep.wsum + col0 + 8 * g));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(ep.wsum + col0 + 8 * g + 4));
      const uint64_t ww[4] = {f2_pack(w0.x, w0.y), f2_pack(w0.z, w0.w), f2_pack(w1.x, w1.y), f2_pack(w1.z, w1.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = f2_fma(rstd2, v[i], f2_fma(nmr2, ww[i], bb[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = f2_add(v[i], bb[i]);
    }
    if (KIND == ST_BF16_GELU) {
      if (ep.gelu_erf) {      // warp-uniform: exact-GELU semantics (FFN.forward :257/:262) at ~2x the epilogue instructions
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a, b;
          f2_unpack(v[i], a, b);
          v[i] = f2_pack(gelu_as_f(a), gelu_as_f(b));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = f2_gelu(v[i]);
      }
    }
    const uint32_t addr = sub_row + ((g ^ sw2) << 4);
    float lo[4], hi[4];
    if (KIND == ST_RES_H) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = unpack_h2(res[ROFF + 4 * g + i]);
        v[i] = f2_add(v[i], f2_pack(h.x, h.y));
        s1 = f2_add(s1, v[i]);
        s2 = f2_fma(v[i], v[i], s2);
        f2_unpack(v[i], lo[i], hi[i]);
      }
      st_shared_u4(addr, pack_h2_sat(lo[0], hi[0]), pack_h2_sat(lo[1], hi[1]), pack_h2_sat(lo[2], hi[2]), pack_h2_sat(lo[3], hi[3]));
      if (ep.sat_count != nullptr) {     // debug only (hig_debug_saturation): a saturating store would otherwise be silent
        int n = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) n += (fabsf(lo[i]) > 65504.f) + (fabsf(hi[i]) > 65504.f);
        if (n) atomicAdd(ep.sat_count, (unsigned long long)n);
      }
    } else if (KIND == ST_F16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) f2_unpack(v[i], lo[i], hi[i]);
      st_shared_u4(addr, pack_h2_sat(lo[0], hi[0]), pack_h2_sat(lo[1], hi[1]), pack_h2_sat(lo[2], hi[2]), pack_h2_sat(lo[3], hi[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) f2_unpack(v[i], lo[i], hi[i]);
      st_shared_u4(addr, pack_bf16x2(lo[0], hi[0]), pack_bf16x2(lo[1], hi[1]), pack_bf16x2(lo[2], hi[2]), pack_bf16x2(lo[3], hi[3]));
    }
  }
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GS_THREADS, 1)
gemm_wres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __half* __restrict__ resid, int ldr, int M, int N, int K,
                 StreamEpi ep, int f16_ops, unsigned long long* __restrict__ trace) {
  // trace (debug, normally null): 32 clock64 / globaltimer slots per pair written by the leader CTA (tools/gemm_trace.py)
  unsigned long long* tr = nullptr;
  if (trace != nullptr && cluster_ctarank() == 0) {
    tr = trace + (size_t)(blockIdx.x >> 1) * 32;
    if (threadIdx.x == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      tr[0] = clock64();
      tr[29] = gt;
    }
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = sW + WR_W_BYTES;
  uint8_t* sEpi = sA + WR_STAGES * GS_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + WR_EPI_BYTES);
  uint64_t* empty_bar = full_bar + WR_STAGES;
  uint64_t* tfull_bar = empty_bar + WR_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wfull_bar = tempty_bar + 2;
  uint64_t* wfree_bar = wfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfree_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int m_tiles = (M + 2 * GS_BM - 1) / (2 * GS_BM);
  const int n_tiles = N / GS_BN;                       // N % 256 == 0 (host-checked)
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + GS_BK - 1) / GS_BK;        // <= WR_KB (host-checked)
  // column-block-major order, one contiguous range of tiles per pair (sizes differ by at most one).  With the query softmax
  // in the epilogue the query-block tiles (first in this order) are epilogue-bound and cost ~5/4 of a plain tile
  // (measured 3.9 vs 3.2 us), so the ranges are cut by weight: pairs on query tiles take ~6.9 tiles, the others ~8.6.
  int lo, hi;
  if (KIND == ST_LN_QSM && n_tiles > QSM_COLS / GS_BN) {
    const long long tq = (long long)(QSM_COLS / GS_BN) * m_tiles;
    const long long wtot = 5 * tq + 4 * ((long long)num_tiles - tq);
    auto cut = [&](long long p) {
      const long long w = wtot * p / num_pairs;
      return (int)(w <= 5 * tq ? w / 5 : tq + (w - 5 * tq) / 4);
    };
    lo = cut(pair);
    hi = cut(pair + 1);
  } else {
    lo = (int)((long long)num_tiles * pair / num_pairs);
    hi = (int)((long long)num_tiles * (pair + 1) / num_pairs);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < WR_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, 2 * GS_EPI_WARPS);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(wfree_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr != nullptr && threadIdx.x == 0) tr[1] = clock64();

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      auto load_w = [&](int n_blk) {
        // one barrier for the whole block: waiting k-block by k-block was measured and is slower — the first tile then runs
        // while 7/8 of W still competes with A for the ~44 B/clk an SM gets from L2 (profiles/r01d_gemm_trace.txt)
        if (rank == 0) mbar_arrive_expect_tx(wfull_bar, 2 * k_blocks * GS_B_BYTES);
        for (int kb = 0; kb < k_blocks; ++kb)
          tma_load_2d_2cta(sW + kb * GS_B_BYTES, &tmB, wfull_bar, kb * GS_BK, n_blk * GS_BN + (int)rank * (GS_BN / 2));
      };
      int cur_n = -1;
      uint32_t run = 0;
      // W does not depend on the previous kernel: the whole resident block is requested BEFORE griddepcontrol.wait
      if (lo < hi) { cur_n = lo / m_tiles; load_w(cur_n); run = 1; }
      pdl_wait();
      pdl_trigger();
      if (tr != nullptr) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        tr[2] = clock64();
        tr[31] = gt;
      }
      uint32_t it = 0;
      for (int tile = lo; tile < hi; ++tile) {
        const int n_blk = tile / m_tiles;
        const int m_blk = tile - n_blk * m_tiles;
        if (n_blk != cur_n) {            // the MMAs that read the old block have completed (commit from the MMA warp)
          mbar_wait(wfree_bar, (run - 1u) & 1u);
          load_w(n_blk);
          cur_n = n_blk;
          ++run;
        }
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % WR_STAGES;
          const uint32_t phase = (it / WR_STAGES) & 1u;
          mbar_wait(empty_bar + stage, phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * GS_A_BYTES);
          tma_load_2d_2cta(sA + stage * GS_A_BYTES, &tmA, full_bar + stage, kb * GS_BK, m_blk * 2 * GS_BM + (int)rank * GS_BM);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16kind(2 * GS_BM, GS_BN, f16_ops ? 0u : 1u);
      uint32_t it = 0, lt = 0, runs = 0;
      int cur_n = -1;
      for (int tile = lo; tile < hi; ++tile, ++lt) {
        const int n_blk = tile / m_tiles;
        if (n_blk != cur_n) {
          mbar_wait(wfull_bar, runs & 1u);
          if (tr != nullptr && runs == 0) tr[3] = clock64();
          ++runs;
          cur_n = n_blk;
        }
        const uint32_t as = lt & 1u;
        const uint32_t aphase = (lt >> 1) & 1u;
        mbar_wait(tempty_bar + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GS_BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % WR_STAGES;
          const uint32_t phase = (it / WR_STAGES) & 1u;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * GS_A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + kb * GS_B_BYTES));
#pragma unroll
          for (int k = 0; k < GS_BK / 16; ++k)
            umma_f16_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(empty_bar + stage, 0b11);
        }
        umma_commit_2cta_mc(tfull_bar + as, 0b11);
        if (tr != nullptr && lt < 8) tr[4 + lt] = clock64();
        if (tile + 1 < hi && (tile + 1) / m_tiles != cur_n) umma_commit_2cta_mc(wfree_bar, 0b11);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..9, both CTAs) =================
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int ch = ew >> 2;          // column half of the tile
    uint8_t* sub0 = sEpi + ew * 2 * WR_SUB;
    const int sw2 = (lane >> 1) & 3;
    const uint32_t row_s0 = smem_u32(sub0) + lane * 64;
    const uint32_t row_s1 = row_s0 + WR_SUB;
    pdl_wait();   // statistics / residual we read and the output we overwrite belong to the previous kernels
    uint32_t res0[32], res1[32];
    if (KIND == ST_RES_H) {
      if (lo < hi) {
        const int n_blk = lo / m_tiles, m_blk = lo - n_blk * m_tiles;
        const int row = m_blk * 2 * GS_BM + (int)rank * GS_BM + q * 32 + lane;
        wr_load_res(resid + (size_t)min(row, M - 1) * ldr + n_blk * GS_BN + ch * (GS_BN / 2), row < M, res0);
      }
    }
    uint32_t lt = 0;
    for (int tile = lo; tile < hi; ++tile, ++lt) {
      const int n_blk = tile / m_tiles;
      const int m_blk = tile - n_blk * m_tiles;
      const uint32_t as = lt & 1u;
      const uint32_t aphase = (lt >> 1) & 1u;
      const int row0 = m_blk * 2 * GS_BM + (int)rank * GS_BM + q * 32;
      const int gc0 = n_blk * GS_BN + ch * (GS_BN / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GS_BN + ch * (GS_BN / 2);
      float rstd = 0.f, nmr = 0.f;
      if (KIND == ST_LN_BF16 || KIND == ST_LN_QSM) {
        const int row = min(row0 + lane, M - 1);
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(ep.stats_in + (size_t)row * 8));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(ep.stats_in + (size_t)row * 8 + 4));
        const float sum = (p0.x + p0.z) + (p1.x + p1.z);
        const float ssq = (p0.y + p0.w) + (p1.y + p1.w);
        const float mu = sum * ep.inv_width;
        const float var = fmaxf(fmaf(ssq, ep.inv_width, -mu * mu), 0.f);
        rstd = rsqrtf(var + ep.ln_eps);
        nmr = -mu * rstd;
      }
      const uint64_t rstd2 = f2_pack(rstd, rstd), nmr2 = f2_pack(nmr, nmr);
      uint64_t s1 = 0ull, s2 = 0ull;
      if (KIND == ST_RES_H) {   // second half of this tile's residual: in flight while the MMAs finish
        const int row = row0 + lane;
        wr_load_res(resid + (size_t)min(row, M - 1) * ldr + gc0 + 64, row < M, res1);
      }

      mbar_wait(tfull_bar + as, aphase);
      tc_fence_after();
      if (tr != nullptr && warp == 2 && lane == 0 && lt < 8) tr[12 + lt] = clock64();
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(taddr, ra);
      tmem_ld_32x32(taddr + 32, rb);
      tmem_ld_wait();
      // query block of a Q / Q|K|V projection: (ra, rb) is one head of this lane's row -> feature softmax in place
      const bool qsm = KIND == ST_LN_QSM && gc0 < QSM_COLS;
      if (qsm) wres_softmax64(ra, rb, ep, gc0, rstd, nmr);
      // ---- columns 0..31 -> sub-slab 0, 32..63 -> sub-slab 1 (each store group is waited for two chunks later)
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      wres_chunk<KIND, 0>(ra, res0, row_s0, sw2, ep, gc0, rstd2, nmr2, s1, s2, qsm);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmC, reinterpret_cast<const void*>(sub0), gc0, row0); bulk_commit(); bulk_wait_read<1>(); }
      __syncwarp();
      wres_chunk<KIND, 16>(rb, res0, row_s1, sw2, ep, gc0 + 32, rstd2, nmr2, s1, s2, qsm);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmC, reinterpret_cast<const void*>(sub0 + WR_SUB), gc0 + 32, row0); bulk_commit(); }
      if (KIND == ST_RES_H) {   // first half of the NEXT tile's residual
        if (tile + 1 < hi) {
          const int nn = (tile + 1) / m_tiles, nm = (tile + 1) - nn * m_tiles;
          const int row = nm * 2 * GS_BM + (int)rank * GS_BM + q * 32 + lane;
          wr_load_res(resid + (size_t)min(row, M - 1) * ldr + nn * GS_BN + ch * (GS_BN / 2), row < M, res0);
        }
      }
      // ---- columns 64..127; the accumulator buffer goes back to the MMA warp once they are in registers
      tmem_ld_32x32(taddr + 64, ra);
      tmem_ld_32x32(taddr + 96, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive_cluster(tempty_bar + as, 0); bulk_wait_read<1>(); }
      __syncwarp();
      if (qsm) wres_softmax64(ra, rb, ep, gc0 + 64, rstd, nmr);     // gc0 is a multiple of 128: same block as the first head
      wres_chunk<KIND, 0>(ra, res1, row_s0, sw2, ep, gc0 + 64, rstd2, nmr2, s1, s2, qsm);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmC, reinterpret_cast<const void*>(sub0), gc0 + 64, row0); bulk_commit(); bulk_wait_read<1>(); }
      __syncwarp();
      wres_chunk<KIND, 16>(rb, res1, row_s1, sw2, ep, gc0 + 96, rstd2, nmr2, s1, s2, qsm);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmC, reinterpret_cast<const void*>(sub0 + WR_SUB), gc0 + 96, row0); bulk_commit(); }
      if (KIND == ST_RES_H) {
        // deterministic row statistics: partial (n_blk, ch) of row (sum, sum of squares) — summed by the consumer
        if (ep.stats_out != nullptr && row0 + lane < M) {
          float a0, a1, b0, b1;
          f2_unpack(s1, a0, a1);
          f2_unpack(s2, b0, b1);
          *reinterpret_cast<float2*>(ep.stats_out + (size_t)(row0 + lane) * 8 + (n_blk * 2 + ch) * 2) = make_float2(a0 + a1, b0 + b1);
        }
      }
      if (tr != nullptr && warp == 2 && lane == 0 && lt < 8) tr[20 + lt] = clock64();
    }
    if (lane == 0) bulk_wait<0>();   // all stores of this warp have completed before the CTA may retire
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
  if (tr != nullptr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tr[28] = clock64();
    tr[30] = gt;
  }
}

// debug: device counter of saturating fp16 stream stores (hig_debug_saturation); null = off
static unsigned long long* g_sat_count = nullptr;
void set_saturation_counter(unsigned long long* p) { g_sat_count = p; }
static bool gelu_erf_enabled() {
  const char* e = getenv("HIG_GELU");
  return e && (e[0] == 'e' || e[0] == 'E');
}

// debug: set through hig_debug_trace(buf, max_launches); every traced launch takes the next block of 74 * 32 slots
static unsigned long long* g_trace = nullptr;
static int g_trace_left = 0;
void set_gemm_trace(unsigned long long* buf, int max_launches) {
  g_trace = max_launches > 0 ? buf : nullptr;
  g_trace_left = g_trace ? max_launches : 0;
}

// ---------------------------------------------------------------------------------------------- host side
int get_tmap_2b(const void* ptr, int rows, int cols, int ld, int box_rows, int is_f16, CUtensorMap* out);  // gemm_tcgen05.cu
int device_num_sms();

template <int KIND>
static int launch_stream(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int M, int N, int K,
                         const StreamEpi& ep, int f16_ops, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm_stream_kernel<KIND>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("cudaFuncSetAttribute(stream gemm): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (M + 2 * GS_BM - 1) / (2 * GS_BM);
  const int n_tiles = (N + GS_BN - 1) / GS_BN;
  int pairs = m_tiles * n_tiles;
  const int max_pairs = device_num_sms() / 2;
  if (pairs > max_pairs) pairs = max_pairs;
  if (const char* pe = getenv("HIG_GS_PAIRS")) { const int v = atoi(pe); if (v > 0 && v < pairs) pairs = v; }  // experiment knob
  cudaError_t e = launch_pdl(kern, dim3(2 * pairs), dim3(GS_THREADS), GS_SMEM, stream, tmA, tmB, tmC, M, N, K, ep, f16_ops);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("stream gemm launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

template <int KIND>
static int launch_wres(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const void* resid, int ldr,
                       int M, int N, int K, const StreamEpi& ep, int f16_ops, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm_wres_kernel<KIND>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("cudaFuncSetAttribute(resident-W gemm): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (M + 2 * GS_BM - 1) / (2 * GS_BM);
  int pairs = m_tiles * (N / GS_BN);
  const int max_pairs = device_num_sms() / 2;
  if (pairs > max_pairs) pairs = max_pairs;
  if (const char* pe = getenv("HIG_GS_PAIRS")) { const int v = atoi(pe); if (v > 0 && v < pairs) pairs = v; }  // experiment knob
  cudaError_t e = launch_pdl(kern, dim3(2 * pairs), dim3(GS_THREADS), WR_SMEM, stream, tmA, tmB, tmC,
                             static_cast<const __half*>(resid), ldr, M, N, K, ep, f16_ops, g_trace);
  if (g_trace != nullptr) {   // next traced launch takes the next block; tracing switches itself off when the buffer is full
    g_trace += 74 * 32;
    if (--g_trace_left <= 0) g_trace = nullptr;
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("resident-W gemm launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// HIG_WRES=0 routes everything through the streamed kernel (A/B timing); default: resident-W wherever it applies
static bool wres_enabled() {
  const char* e = getenv("HIG_WRES");
  return !(e && e[0] == '0');
}

int gemm_stream(int kind, const void* A, int lda, const void* W, int ldw, int op_dtype, int M, int N, int K,
                const float* bias, const float* wsum, const float* stats_in, float* stats_out, int ln_width,
                void* out, int ldo, cudaStream_t stream) {
  if (!A || !W || !bias || !out || M <= 0 || N <= 0 || K <= 0) return set_error(HIG_ERR_INVALID, "gemm_stream: null operand or empty shape");
  if (kind < ST_BF16 || kind > ST_LN_QSM) return set_error(HIG_ERR_INVALID, "gemm_stream: bad kind");
  if (op_dtype != HIG_BF16 && op_dtype != HIG_F16) return set_error(HIG_ERR_INVALID, "gemm_stream: operands are bf16 or fp16");
  if ((lda % 8) || (ldw % 8) || (ldo % 8) || (K % 8)) return set_error(HIG_ERR_INVALID, "gemm_stream: leading dims / K must be multiples of 8");
  if (N % 64) return set_error(HIG_ERR_UNSUPPORTED, "gemm_stream: N must be a multiple of 64");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(bias) & 15))
    return set_error(HIG_ERR_INVALID, "gemm_stream: operands must be 16-byte aligned");
  if (kind == ST_LN_BF16 || kind == ST_LN_QSM) {
    if (kind == ST_LN_QSM && N < QSM_COLS) return set_error(HIG_ERR_INVALID, "gemm_stream: the query block is the first 512 output columns");
    if (!wsum || !stats_in || ln_width <= 0) return set_error(HIG_ERR_INVALID, "gemm_stream: LN kind needs wsum, stats_in, ln_width");
    if ((reinterpret_cast<uintptr_t>(wsum) & 15) || (reinterpret_cast<uintptr_t>(stats_in) & 15))
      return set_error(HIG_ERR_INVALID, "gemm_stream: wsum / stats_in must be 16-byte aligned");
  }
  if (stats_out) {
    if (kind != ST_RES_H || N != 2 * GS_BN) return set_error(HIG_ERR_UNSUPPORTED, "gemm_stream: row statistics come from the N = 512 residual kind");
    if (reinterpret_cast<uintptr_t>(stats_out) & 7) return set_error(HIG_ERR_INVALID, "gemm_stream: stats_out must be 8-byte aligned");
  }
  StreamEpi ep;
  ep.bias = bias; ep.wsum = wsum; ep.stats_in = stats_in; ep.stats_out = stats_out;
  ep.inv_width = ln_width > 0 ? 1.0f / (float)ln_width : 0.f;
  ep.ln_eps = 1e-5f;
  ep.sat_count = g_sat_count;
  ep.gelu_erf = gelu_erf_enabled() ? 1 : 0;      // read per call: HIG_GELU=erf / tanh (default) can be A/B-toggled
  const int f16 = op_dtype == HIG_F16;
  const int tm_f16 = f16;
  CUtensorMap tmA, tmB, tmC;
  int rc = get_tmap_2b(A, M, K, lda, 128, tm_f16, &tmA);
  if (rc) return rc;
  rc = get_tmap_2b(W, N, K, ldw, 128, tm_f16, &tmB);
  if (rc) return rc;
  // resident-W kernel: whole K in shared memory, full 256-column blocks; the residual is read with 32-byte loads
  const bool wres = wres_enabled() && K <= WR_KB * GS_BK && (N % GS_BN) == 0 &&
                    (kind != ST_RES_H || (reinterpret_cast<uintptr_t>(out) & 31) == 0);
  if (wres) {
    rc = get_tmap_2b(out, M, N, ldo, 32, ((kind == ST_RES_H || kind == ST_F16) ? 1 : 0) | 2, &tmC);   // 32 x 32 boxes, 64-byte swizzle
    if (rc) return rc;
    switch (kind) {
      case ST_F16: return launch_wres<ST_F16>(tmA, tmB, tmC, nullptr, 0, M, N, K, ep, f16, stream);
      case ST_LN_QSM: return launch_wres<ST_LN_QSM>(tmA, tmB, tmC, nullptr, 0, M, N, K, ep, f16, stream);
      case ST_BF16: return launch_wres<ST_BF16>(tmA, tmB, tmC, nullptr, 0, M, N, K, ep, f16, stream);
      case ST_BF16_GELU: return launch_wres<ST_BF16_GELU>(tmA, tmB, tmC, nullptr, 0, M, N, K, ep, f16, stream);
      case ST_RES_H: return launch_wres<ST_RES_H>(tmA, tmB, tmC, out, ldo, M, N, K, ep, f16, stream);
      default: return launch_wres<ST_LN_BF16>(tmA, tmB, tmC, nullptr, 0, M, N, K, ep, f16, stream);
    }
  }
  if (kind == ST_F16 || kind == ST_LN_QSM)
    return set_error(HIG_ERR_UNSUPPORTED, "gemm_stream: this kind exists on the resident-W kernel only (K <= 512, N % 256 == 0)");
  rc = get_tmap_2b(out, M, N, ldo, 32, kind == ST_RES_H, &tmC);
  if (rc) return rc;
  switch (kind) {
    case ST_BF16: return launch_stream<ST_BF16>(tmA, tmB, tmC, M, N, K, ep, f16, stream);
    case ST_BF16_GELU: return launch_stream<ST_BF16_GELU>(tmA, tmB, tmC, M, N, K, ep, f16, stream);
    case ST_RES_H: return launch_stream<ST_RES_H>(tmA, tmB, tmC, M, N, K, ep, f16, stream);
    default: return launch_stream<ST_LN_BF16>(tmA, tmB, tmC, M, N, K, ep, f16, stream);
  }
}

}  // namespace hig
