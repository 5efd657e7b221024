// hig_attn_apply_stylize on tcgen05 / TMEM (product path): the query half of the efficient attention fused with the
// StylizationBlock front end, one 128-row tile of one sequence per step of a persistent CTA:
//
//   Y[t, 64h : 64h+64] = Qs[t, 64h : 64h+64] . A[s, h]            8 heads -> 8 x (4 x tcgen05.mma M=128 N=64 K=16)
//   out[t, :]          = SiLU( LayerNorm_512(Y[t, :]) (1 + scale_s) + shift_s )
//
// (einsum 'bnhd,bhdl->bnhl' + reshape of LinearTemporal{Self,Cross,InteractionCross}Attention.forward,
// codes/models/interaction_transformer.py:128,162,201, then StylizationBlock.forward's norm / FiLM / SiLU :86-97.)
//
// Why tcgen05 here although the contraction is < 2 % of the step's FLOPs: the 128 x 512 fp32 tile lands in the SM's
// whole TMEM (128 lanes x 512 columns) with ONE ROW PER LANE, so the LayerNorm statistics of a row are a lane-local
// sum over tcgen05.ld chunks — no ldmatrix / mma.sync / quad-shuffle / shared-memory-statistics dependency chain, which
// is what held the mma.sync kernel of attn_apply.cu at 26.6 us against a 7.9 us HBM floor (0.30 of peak, 44 % issue
// active at 16 warps per SM; profiles/r01e_ncu_full_summary.txt).  Y never exists outside TMEM.
//
// Operands: Qs = softmax_feat(Q) as the Q / Q|K|V projection's epilogue leaves it (bf16, HIG_GS_LN_QSM), streamed
// head by head (128 rows x 64 columns = 16 KB, K-major, 128B swizzle) together with that head's A^T ([l][d], 8 KB: the
// K-major B operand, written in that layout by attn_kv_kernel / the text precompute) through a 6-stage TMA ring.
// 3-D tensor maps [S][T][cols] clip rows >= T on load (zero fill) and on store.
// Roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..17 epilogue (TMEM lane quarter = warp % 4, column
// quarter = (warp - 2) / 4): pass 1 row statistics with the next tcgen05.ld always in flight, one 128-thread named barrier
// per lane quarter to combine the four column quarters' partials, pass 2 normalise + FiLM + SiLU -> bf16 -> 64B-swizzled
// 32 x 32 sub-slabs -> TMA store.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

namespace atc {
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + 32 * EPI_WARPS;   // warp 0 producer, warp 1 MMA, warps 2..17 epilogue
constexpr int STAGES = 6;
constexpr int Q_BYTES = 128 * 64 * 2;          // one head block of the tile (128 rows x 64 columns)
constexpr int AH_BYTES = 64 * 64 * 2;          // A^T of one head
constexpr int STAGE_BYTES = Q_BYTES + AH_BYTES;   // a ring stage carries BOTH operands of one head's MMAs
constexpr int SUB = 32 * 64;                   // staging sub-slab: 32 rows x 32 bf16 columns (64-byte rows, 64B swizzle)
constexpr int EPI_BYTES = EPI_WARPS * 2 * SUB; // two sub-slabs per epilogue warp
constexpr int GB_BYTES = 2 * 2 * 512 * 4;      // (G, B) x two tile parities
constexpr int RED_BYTES = EPI_WARPS * 32 * 8;
constexpr int BAR_BYTES = (2 * STAGES + 2) * 8 + 16;
constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + GB_BYTES + RED_BYTES + BAR_BYTES + 1024;
static_assert(SMEM <= 232448, "shared memory budget");

HIG_DEVICE void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
HIG_DEVICE void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
HIG_DEVICE void bulk_commit_g() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> HIG_DEVICE void bulk_wait_read_g() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> HIG_DEVICE void bulk_wait_g() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
HIG_DEVICE void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
HIG_DEVICE float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
HIG_DEVICE void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// sum and sum of squares of one 32-column accumulator chunk, packed pairs
HIG_DEVICE void acc_stats(const uint32_t (&r)[32], uint64_t& s1, uint64_t& s2) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t v = f2_pack_u(r[2 * i], r[2 * i + 1]);
    s1 = f2_add(s1, v);
    s2 = f2_fma(v, v, s2);
  }
}
// one 32-column chunk of this lane's row: normalise, FiLM affine, SiLU, bf16 -> the 64-byte row of a sub-slab
// (16-byte chunks XOR-swizzled with (row >> 1) & 3, as CU_TENSOR_MAP_SWIZZLE_64B expects)
HIG_DEVICE void finish_chunk(const uint32_t (&r)[32], uint32_t gaddr, uint32_t baddr, uint64_t rstd2, uint64_t nmr2, bool silu,
                             uint32_t sub_row, int sw2) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {         // 8 columns -> one 16-byte chunk
    const float4 G0 = lds_f4(gaddr + g8 * 32), G1 = lds_f4(gaddr + g8 * 32 + 16);
    const float4 B0 = lds_f4(baddr + g8 * 32), B1 = lds_f4(baddr + g8 * 32 + 16);
    const uint64_t GG[4] = {f2_pack(G0.x, G0.y), f2_pack(G0.z, G0.w), f2_pack(G1.x, G1.y), f2_pack(G1.z, G1.w)};
    const uint64_t BB[4] = {f2_pack(B0.x, B0.y), f2_pack(B0.z, B0.w), f2_pack(B1.x, B1.y), f2_pack(B1.z, B1.w)};
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint64_t v = f2_fma(f2_pack_u(r[8 * g8 + 2 * i], r[8 * g8 + 2 * i + 1]), rstd2, nmr2);
      v = f2_fma(v, GG[i], BB[i]);
      float a, b;
      f2_unpack(v, a, b);
      if (silu) {
        v = f2_fma(v, f2_pack(tanh_approx_f(a), tanh_approx_f(b)), v);
        f2_unpack(v, a, b);
      }
      pk[i] = pack_bf16x2(a, b);
    }
    sts_u4(sub_row + ((g8 ^ sw2) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
}
}  // namespace atc
using namespace atc;

__global__ void __launch_bounds__(atc::THREADS, 1)
attn_apply_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmO, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ scale_shift, int ss_stride, int apply_silu,
                     int S, int T) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // ring: stage s = [Q block 16 KB | A^T head 8 KB]
  uint8_t* sEpi = sQ + STAGES * STAGE_BYTES;
  float* sGB = reinterpret_cast<float*>(sEpi + EPI_BYTES);            // [parity][G | B][512]
  float2* sRed = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(sGB) + GB_BYTES);   // [16 warps][32 lanes]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sRed) + RED_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = T > 128 ? 2 : 1;            // tiles per sequence: rows [0, 128) and [128, T)
  const int num_tiles = S * nt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile i: the first S tiles are rows [0, 128) of sequence i, the next S rows [128, T) of sequence i - S
  auto tile_seq = [&](int i) { return i < S ? i : i - S; };
  auto tile_row0 = [&](int i) { return i < S ? 0 : 128; };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      pdl_wait();        // Qs and A^T are the outputs of the two previous kernels
      pdl_trigger();
      // the ring runs ahead of the MMA warp across tile boundaries: while the epilogue drains tile i, the operands of the
      // first STAGES heads of tile i+1 are already landing (the first cut kept A^T in a single buffer gated on the previous
      // tile's MMAs and had 5 of 8 query blocks in flight: the next tile's MMAs then waited on L2 latency)
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int s = tile_seq(tile), r0 = tile_row0(tile);
        for (int h = 0; h < 8; ++h, ++it) {
          const uint32_t stage = it % STAGES, phase = (it / STAGES) & 1u;
          mbar_wait(empty_bar + stage, phase ^ 1u);
          mbar_arrive_expect_tx(full_bar + stage, STAGE_BYTES);
          tma_load_3d(sQ + stage * STAGE_BYTES, &tmQ, full_bar + stage, h * 64, r0, s);
          tma_load_2d(sQ + stage * STAGE_BYTES + Q_BYTES, &tmA, full_bar + stage, 0, (s * 8 + h) * 64);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        mbar_wait(tempty_bar, (lt & 1u) ^ 1u);     // the epilogue has drained the previous tile's accumulator
        tc_fence_after();
        for (int h = 0; h < 8; ++h, ++it) {
          const uint32_t stage = it % STAGES, phase = (it / STAGES) & 1u;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + stage * STAGE_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sQ + stage * STAGE_BYTES + Q_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + h * 64, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
          umma_commit(empty_bar + stage);
        }
        umma_commit(tfull_bar);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps 2..17 =================
    // TMEM lane quarter = warp % 4 (hardware rule); the four warps of a quarter take 128 columns (two heads) each.
    // tcgen05.ld of the next 32-column chunk is always in flight while the current one is processed.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int cq = ew >> 2;            // column quarter
    const int et = threadIdx.x - 64;   // 0..511
    const uint32_t sub0 = smem_u32(sEpi) + ew * 2 * SUB;
    const uint32_t row_s[2] = {sub0 + lane * 64, sub0 + SUB + lane * 64};
    const int sw2 = (lane >> 1) & 3;
    const bool silu = (apply_silu & 1) != 0;
    const float hs = silu ? 0.5f : 1.0f;     // SiLU(x) = h + h tanh(h), h = x / 2: the affine carries the 1/2
    // folded FiLM affine of a tile's sequence: out = n_hat * G + B,  G = gamma (1 + scale),  B = beta (1 + scale) + shift.
    // Written one tile AHEAD (double-buffered by tile parity) so that its global loads never sit on a tile's critical path.
    auto stage_affine = [&](int tile, uint32_t par) {
      const int s = tile_seq(tile);
      float* sG = sGB + par * 1024;
      float* sB = sG + 512;
      const int c = et;
      float gg = __ldg(gamma + c), bb = __ldg(beta + c);
      if (scale_shift != nullptr) {
        const float m1 = 1.0f + __ldg(scale_shift + (size_t)s * ss_stride + c);
        gg *= m1;
        bb = fmaf(bb, m1, __ldg(scale_shift + (size_t)s * ss_stride + 512 + c));
      }
      sG[c] = gg * hs;
      sB[c] = bb * hs;
    };
    pdl_wait();          // scale_shift comes from this step's FiLM GEMM
    if ((int)blockIdx.x < num_tiles) stage_affine(blockIdx.x, 0);
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int s = tile_seq(tile), r0 = tile_row0(tile);
      const uint32_t par = lt & 1u;
      float* sG = sGB + par * 1024;
      float* sB = sG + 512;
      named_bar(5, 32 * EPI_WARPS);
      const int qrow0 = r0 + q * 32;
      const bool live = qrow0 < T;       // warp-uniform: this quarter holds at least one valid row
      mbar_wait(tfull_bar, par);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cq * 128;
      uint32_t ra[32], rb[32];
      // ---- pass 1: row statistics over this warp's 128 columns
      uint64_t s1 = 0ull, s2 = 0ull;
      if (live) {
        tmem_ld_32x32(taddr, ra);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32, rb);
        acc_stats(ra, s1, s2);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 64, ra);
        acc_stats(rb, s1, s2);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 96, rb);
        acc_stats(ra, s1, s2);
        tmem_ld_wait();
        tmem_ld_32x32(taddr, ra);          // first chunk of pass 2: in flight across the statistics exchange
        acc_stats(rb, s1, s2);
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(s1, a0, a1);
        f2_unpack(s2, b0, b1);
        sRed[ew * 32 + lane] = make_float2(a0 + a1, b0 + b1);
      }
      named_bar(1 + q, 128);             // the four column quarters of this lane quarter
      float rstd, nmr;
      {
        float sum = 0.f, ssq = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 p = sRed[(j * 4 + (ew & 3)) * 32 + lane];
          sum += p.x;
          ssq += p.y;
        }
        const float mean = sum * (1.0f / 512.0f);
        const float var = fmaxf(fmaf(ssq, 1.0f / 512.0f, -mean * mean), 0.f);
        rstd = rsqrtf(var + 1e-5f);
        nmr = -mean * rstd;
      }
      const uint64_t rstd2 = f2_pack(rstd, rstd), nmr2 = f2_pack(nmr, nmr);
      // ---- pass 2: normalise + FiLM + SiLU -> bf16 -> sub-slab -> TMA store (32 columns per store)
      if (live) {
        const uint32_t gB = smem_u32(sG) + cq * 512, bB = smem_u32(sB) + cq * 512;
        auto store = [&](int c32) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, sub0 + (c32 & 1) * SUB, cq * 128 + c32 * 32, qrow0, s);
            bulk_commit_g();
            bulk_wait_read_g<1>();       // the other sub-slab's previous store has been read out: it is the next target
          }
          __syncwarp();
        };
        if (lane == 0) bulk_wait_read_g<0>();
        __syncwarp();
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32, rb);
        finish_chunk(ra, gB, bB, rstd2, nmr2, silu, row_s[0], sw2);
        store(0);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 64, ra);
        finish_chunk(rb, gB + 128, bB + 128, rstd2, nmr2, silu, row_s[1], sw2);
        store(1);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 96, rb);
        finish_chunk(ra, gB + 256, bB + 256, rstd2, nmr2, silu, row_s[0], sw2);
        store(2);
        tmem_ld_wait();
        // every tcgen05.ld of this tile has completed: hand TMEM back to the MMA warp before the last chunk's math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        finish_chunk(rb, gB + 384, bB + 384, rstd2, nmr2, silu, row_s[1], sw2);
        store(3);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
      }
      // next tile's affine into the other parity buffer: every warp passed this tile's named barrier, i.e. finished the
      // previous tile, the last reader of that buffer
      if (tile + (int)gridDim.x < num_tiles) stage_affine(tile + gridDim.x, par ^ 1u);
    }
    if (lane == 0) bulk_wait_g<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 [S][T][cols] with row pitch ld (elements) and sequence pitch T * ld; box = [1][box_rows][64 cols], 128B swizzle
// (narrow: box = [1][box_rows][32 cols], 64B swizzle — the epilogue's staging sub-slabs)
static int get_tmap_3d(const void* ptr, int S, int T, int cols, int ld, int box_rows, int narrow, CUtensorMap* out) {
  struct Key {
    const void* p; int S, T, cols, ld, br;
    bool operator==(const Key& o) const { return p == o.p && S == o.S && T == o.T && cols == o.cols && ld == o.ld && br == o.br; }
  };
  struct KH {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.p);
      h ^= (size_t)k.S * 0x9E3779B97F4A7C15ull + (h << 6);
      h ^= (size_t)k.T * 0xC2B2AE3D27D4EB4Full + (h >> 3);
      h ^= (size_t)(k.cols * 31 + k.ld) * 0x165667B19E3779F9ull + (h << 9);
      h ^= (size_t)k.br * 0x27D4EB2F165667C5ull;
      return h;
    }
  };
  static std::unordered_map<Key, CUtensorMap, KH> cache;
  static std::mutex mu;
  static PFN_encodeTiled3 enc = []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_encodeTiled3>(p);
    return (PFN_encodeTiled3) nullptr;
  }();
  Key key{ptr, S, T, cols, ld, box_rows * 2 + narrow};
  {
    std::lock_guard<std::mutex> g(mu);
    auto itr = cache.find(key);
    if (itr != cache.end()) { *out = itr->second; return HIG_OK; }
  }
  if (!enc) return set_error(HIG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled not available");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)S};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[3] = {narrow ? 32u : 64u, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HIG_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed: " + std::to_string((int)r));
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 1024) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return HIG_OK;
}

int get_tmap_2b(const void* ptr, int rows, int cols, int ld, int box_rows, int flags, CUtensorMap* out);  // gemm_tcgen05.cu
int device_num_sms();

// q: Qs (bf16, already softmaxed over each head's 64 features), a_t: A^T [S, 8, 64 (l), 64 (d)] bf16
int attn_apply_stylize_tc(const void* q, int ldq, const void* a_t, const float* gamma, const float* beta,
                          const float* scale_shift, int ss_stride, int apply_silu, void* out, int S, int T, int H,
                          cudaStream_t stream) {
  if (!q || !a_t || !gamma || !beta || !out || S <= 0 || T <= 0)
    return set_error(HIG_ERR_INVALID, "attn_apply_stylize_tc: bad arguments");
  if (H != 8) return set_error(HIG_ERR_UNSUPPORTED, "attn_apply_stylize_tc: built for 8 heads x 64 (latent_dim 512)");
  if (T > 256) return set_error(HIG_ERR_UNSUPPORTED, "attn_apply_stylize_tc: T > 256 not supported");
  if (ldq % 8) return set_error(HIG_ERR_INVALID, "attn_apply_stylize_tc: ldq must be a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(a_t) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(HIG_ERR_INVALID, "attn_apply_stylize_tc: pointers must be 16-byte aligned");
  CUtensorMap tmQ, tmA, tmO;
  int rc = get_tmap_3d(q, S, T, 512, ldq, 128, 0, &tmQ);
  if (rc) return rc;
  rc = get_tmap_2b(a_t, S * 8 * 64, 64, 64, 64, 0, &tmA);
  if (rc) return rc;
  rc = get_tmap_3d(out, S, T, 512, 512, 32, 1, &tmO);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::SMEM);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_apply_stylize_tc attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  const int tiles = S * (T > 128 ? 2 : 1);
  int ctas = device_num_sms();
  if (ctas > tiles) ctas = tiles;
  cudaError_t e = launch_pdl(attn_apply_tc_kernel, dim3(ctas), dim3(atc::THREADS), atc::SMEM, stream, tmQ, tmA, tmO,
                             gamma, beta, scale_shift, ss_stride, apply_silu, S, T);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("attn_apply_stylize_tc launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
