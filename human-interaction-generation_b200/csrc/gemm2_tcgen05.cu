// hig_gemm_bf16, CTA-pair variant:  C[M,N] = act( A[M,K] · W[N,K]^T + bias + residual ) with tcgen05 cta_group::2.
//
// Two CTAs of a cluster (one TPC) cooperate on a 256 x 256 output tile: each CTA TMA-loads ITS 128 rows of A and ITS
// 128 rows of W per k-block (32 KB per stage instead of 48 KB), the leader CTA issues one
// tcgen05.mma.cta_group::2 (M=256, N=256, K=16) that reads both CTAs' shared memory, and each CTA ends up with its
// 128 x 256 fp32 accumulator half in its own TMEM.  Operand traffic per flop drops by a third against the
// single-CTA 128 x 256 kernel — at the K=512 shapes of the denoiser the single-CTA kernel is bound by L2->SMEM
// bandwidth (~7.8 TB/s measured), not by the tensor pipe, so this is the lever that matters.
//
// Pipeline (per CTA): warp 0 = TMA producer, warp 1 = TMEM allocator (+ MMA issuer in the leader),
// warps 2..9 = epilogue.  Barriers: full[s] lives in the leader (both CTAs' TMA bytes complete on it),
// empty[s] / tmem_full[a] are signalled in BOTH CTAs by multicast tcgen05.commit, tmem_empty[a] lives in the leader
// and collects the 16 epilogue warps of the pair.
#include <cuda.h>
#include <string>
#include "hig_common.cuh"
#include "gemm_epilogue.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int G2_BM = 128;       // rows of A per CTA (256 per pair)
constexpr int G2_BN = 256;       // tile columns (128 rows of W per CTA)
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;          // 16 KB
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;    // 16 KB
constexpr int G2_EPI_BYTES = G2_EPI_WARPS * 32 * 32 * 4;
constexpr int G2_BAR_BYTES = (2 * G2_STAGES + 4) * 8 + 16;
constexpr int G2_SMEM = G2_STAGES * (G2_A_BYTES + G2_B_BYTES) + G2_EPI_BYTES + G2_BAR_BYTES + 1024;

// TA / TB = 1: the operand is given "MN-major" — A as [K, M] row-major (A^T in memory), W as [K, N] row-major — and is
// consumed as it lies: TMA boxes of 64 k-rows x 64 MN-columns (128 B, 128B swizzle), two per CTA per stage, described to
// the tensor core with MN-major shared-memory descriptors (instruction-descriptor bits 15 / 16).  This is what lets the
// training path compute  dW = dY^T X  (both operands token-major) and  dX = dY W  (W as stored) without ever
// materialising a transposed copy.
template <int KIND, int TA, int TB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                      int K, GemmEpilogue ep, int vec_ok, int k_splits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + G2_STAGES * G2_A_BYTES;
  float* sEpi = reinterpret_cast<float*>(sB + G2_STAGES * G2_B_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sEpi) + G2_EPI_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int m_tiles = (M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int n_tiles = (N + G2_BN - 1) / G2_BN;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * k_splits;  // split-K slices accumulate atomically in the epilogue
  const int k_blocks_all = (K + G2_BK - 1) / G2_BK;
  const int kb_per = (k_blocks_all + k_splits - 1) / k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, 2 * G2_EPI_WARPS);  // every epilogue warp of BOTH CTAs arrives on the leader's
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // its outputs (our operands / residual) are visible after the wait.  Dependents are released only after the wait,
  // so a kernel's pre-wait code may rely on everything but its immediate predecessor's outputs.
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mn = tile % mn_tiles, ks = tile / mn_tiles;
        const int n_blk = mn % n_tiles;
        const int m_blk = mn / n_tiles;
        const int kb0 = ks * kb_per, kb1 = min(k_blocks_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1 && ep.act != 101; ++kb, ++it) {
          const uint32_t stage = it % G2_STAGES;
          const uint32_t phase = (it / G2_STAGES) & 1u;
          mbar_wait(empty_bar + stage, phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * (G2_A_BYTES + G2_B_BYTES));
          if (TA) {   // [K, M] source: two boxes of 64 k-rows x 64 m-columns
            tma_load_2d_2cta(sA + stage * G2_A_BYTES, &tmA, full_bar + stage, m_blk * 2 * G2_BM + (int)rank * G2_BM, kb * G2_BK);
            tma_load_2d_2cta(sA + stage * G2_A_BYTES + G2_A_BYTES / 2, &tmA, full_bar + stage,
                             m_blk * 2 * G2_BM + (int)rank * G2_BM + 64, kb * G2_BK);
          } else {
            tma_load_2d_2cta(sA + stage * G2_A_BYTES, &tmA, full_bar + stage, kb * G2_BK,
                             m_blk * 2 * G2_BM + (int)rank * G2_BM);
          }
          if (TB) {
            tma_load_2d_2cta(sB + stage * G2_B_BYTES, &tmB, full_bar + stage, n_blk * G2_BN + (int)rank * (G2_BN / 2), kb * G2_BK);
            tma_load_2d_2cta(sB + stage * G2_B_BYTES + G2_B_BYTES / 2, &tmB, full_bar + stage,
                             n_blk * G2_BN + (int)rank * (G2_BN / 2) + 64, kb * G2_BK);
          } else {
            tma_load_2d_2cta(sB + stage * G2_B_BYTES, &tmB, full_bar + stage, kb * G2_BK,
                             n_blk * G2_BN + (int)rank * (G2_BN / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_BM, G2_BN) | (TA ? (1u << 15) : 0u) | (TB ? (1u << 16) : 0u);
      uint32_t it = 0, lt = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
        const uint32_t as = lt & 1u;
        const uint32_t aphase = (lt >> 1) & 1u;
        mbar_wait(tempty_bar + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * G2_BN;
        const int ks = tile / mn_tiles;
        const int kb0 = ks * kb_per, kb1 = min(k_blocks_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1 && ep.act != 101; ++kb, ++it) {
          const uint32_t stage = it % G2_STAGES;
          const uint32_t phase = (it / G2_STAGES) & 1u;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          // K-major: a k-step of 16 elements is 32 B inside the 128-byte swizzle row (+2 in 16-byte units);
          // MN-major: a k-step is two 8-row groups of 1024 B (+128), the two 64-wide MN blocks lie 8 KB apart (LBO)
          const uint64_t da = TA ? umma_desc_mn_sw128(smem_u32(sA + stage * G2_A_BYTES), G2_A_BYTES / 2)
                                 : umma_desc_k_sw128(smem_u32(sA + stage * G2_A_BYTES));
          const uint64_t db = TB ? umma_desc_mn_sw128(smem_u32(sB + stage * G2_B_BYTES), G2_B_BYTES / 2)
                                 : umma_desc_k_sw128(smem_u32(sB + stage * G2_B_BYTES));
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_f16_2cta(tmem_d, da + (TA ? 128 : 2) * k, db + (TB ? 128 : 2) * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(empty_bar + stage, 0b11);  // frees this smem slot in both CTAs
        }
        umma_commit_2cta_mc(tfull_bar + as, 0b11);  // accumulator halves ready in both CTAs
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..9, both CTAs) =================
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    constexpr int COLS_PER_WARP = G2_BN / 2;
    const uint32_t slab = smem_u32(sEpi) + (warp - 2) * 32 * 32 * 4;
    uint32_t lt = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int mn = tile % mn_tiles;
      const int n_blk = mn % n_tiles;
      const int m_blk = mn / n_tiles;
      const uint32_t as = lt & 1u;
      const uint32_t aphase = (lt >> 1) & 1u;
      const int row0 = m_blk * 2 * G2_BM + (int)rank * G2_BM + q * 32;
      const int cbase = ch * COLS_PER_WARP;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G2_BN + cbase;
      const int gc0 = n_blk * G2_BN + cbase;
      EpiLane L;
      epi_setup(L, slab, ep, row0, M, lane);
      if (KIND != EPI_GENERIC && KIND != EPI_FUSED) {
        // specialised kinds (N % 32 == 0, 16-byte friendly): bias / residual of chunk k+1 are fetched while chunk k
        // is transposed and stored; chunk 0's are in flight before the accumulator is even ready.
        EpiPre P[2];
        epi_prefetch<KIND>(P[0], L, ep, gc0);
        mbar_wait(tfull_bar + as, aphase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < COLS_PER_WARP / 32; ++k) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + k * 32, r);
          if (k + 1 < COLS_PER_WARP / 32 && gc0 + (k + 1) * 32 < N) epi_prefetch<KIND>(P[(k + 1) & 1], L, ep, gc0 + (k + 1) * 32);
          tmem_ld_wait();
          if (gc0 + k * 32 < N) epi_finish<KIND>(r, P[k & 1], L, ep, gc0 + k * 32);
        }
      } else {
        mbar_wait(tfull_bar + as, aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          if (gc0 + c < N && ep.act != 100) epilogue_chunk<KIND>(r, L, slab, ep, row0, M, gc0 + c, N, vec_ok, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

template <int KIND, int TA = 0, int TB = 0>
static int launch_2cta_kind(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& ep,
                            int vec_ok, int num_sms, int k_splits, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm_bf16_2cta_kernel<KIND, TA, TB>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("cudaFuncSetAttribute(2cta): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int n_tiles = (N + G2_BN - 1) / G2_BN;
  int pairs = m_tiles * n_tiles * k_splits;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  cudaError_t e = launch_pdl(kern, dim3(2 * pairs), dim3(G2_THREADS), G2_SMEM, stream, tmA, tmB, M, N, K, ep, vec_ok, k_splits);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("gemm 2cta launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

int launch_gemm_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& ep,
                     int vec_ok, int num_sms, int k_splits, cudaStream_t stream) {
  const int kind = (ep.act >= 100) ? EPI_GENERIC : classify_epilogue(ep, vec_ok, N);
  switch (kind) {
    case EPI_BF16: return launch_2cta_kind<EPI_BF16>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_BF16_GELU: return launch_2cta_kind<EPI_BF16_GELU>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_RES_F32: return launch_2cta_kind<EPI_RES_F32>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_RES_F32_BF16: return launch_2cta_kind<EPI_RES_F32_BF16>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_RES_H: return launch_2cta_kind<EPI_RES_H>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_RES_H_BF16: return launch_2cta_kind<EPI_RES_H_BF16>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    case EPI_FUSED: return launch_2cta_kind<EPI_FUSED>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
    default: return launch_2cta_kind<EPI_GENERIC>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
  }
}

// transposed-operand variants (training dgrad / wgrad): bf16 / fp32-residual / generic (split-K atomics) epilogues only
int launch_gemm_2cta_t(int ta, int tb, const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                       const GemmEpilogue& ep, int vec_ok, int num_sms, int k_splits, cudaStream_t stream) {
  int kind = (ep.act >= 100) ? EPI_GENERIC : classify_epilogue(ep, vec_ok, N);
  if (kind == EPI_FUSED) {
    if (ta || !tb) return set_error(HIG_ERR_UNSUPPORTED, "gemm_fused: transposed form is trans_b only");
    return launch_2cta_kind<EPI_FUSED, 0, 1>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);
  }
  if (kind != EPI_BF16 && kind != EPI_RES_F32) kind = EPI_GENERIC;
#define HIG_G2T(KD)                                                                                                   \
  do {                                                                                                                \
    if (ta && tb) return launch_2cta_kind<KD, 1, 1>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);         \
    if (ta) return launch_2cta_kind<KD, 1, 0>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);               \
    return launch_2cta_kind<KD, 0, 1>(tmA, tmB, M, N, K, ep, vec_ok, num_sms, k_splits, stream);                       \
  } while (0)
  if (kind == EPI_BF16) HIG_G2T(EPI_BF16);
  if (kind == EPI_RES_F32) HIG_G2T(EPI_RES_F32);
  HIG_G2T(EPI_GENERIC);
#undef HIG_G2T
}

}  // namespace hig
