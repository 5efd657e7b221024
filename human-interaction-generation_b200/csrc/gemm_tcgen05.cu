// hig_gemm_bf16: C[M,N] = act( A[M,K] · W[N,K]^T + bias + residual )       (bf16 in, fp32 accumulate)
//
// Every dense projection of the denoiser runs through this kernel: QKV / Q / out-proj / FFN / stylization
// emb linears / joint embed / output heads  (reference: nn.Linear call sites in
// codes/models/interaction_transformer.py:74-97,105-128,137-163,172-205,254-263,471-478,508-509).
//
// Design (B200, sm_100a):
//   * persistent grid, one CTA per SM, static round-robin tile scheduler (n fastest so the CTAs that share an
//     A row-panel run together and hit L2);
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx),
//     warp 1 = tcgen05.mma issuer (single elected thread) + TMEM allocator,
//     warps 2..5 = epilogue (tcgen05.ld 32x32b -> registers -> bias/residual/activation -> global);
//   * 128 x BN x 64 tiles, STAGES-deep smem ring, accumulator double-buffered in TMEM (2 x BN columns) so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <string>
#include "hig_common.cuh"
#include "gemm_epilogue.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int EPI_BYTES = GEMM_EPI_WARPS * 32 * 32 * 4;  // one 32 x 16-word transpose slab per epilogue warp
  static constexpr int TOTAL = STAGES * (A_BYTES + B_BYTES) + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         int M, int N, int K, GemmEpilogue ep, int vec_ok, int k_splits) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  float* sEpi = reinterpret_cast<float*>(sB + STAGES * L::B_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * L::B_BYTES + L::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * k_splits;  // split-K: tile = (k slice, m block, n block), slices accumulate atomically
  const int k_blocks_all = (K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (k_blocks_all + k_splits - 1) / k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, GEMM_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<2 * BN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // its outputs (our operands / residual) are visible after the wait.  Dependents are released only after the wait,
  // so a kernel's pre-wait code may rely on everything but its immediate predecessor's outputs.
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile % mn_tiles, ks = tile / mn_tiles;
        const int n_blk = mn % n_tiles;
        const int m_blk = mn / n_tiles;
        const int kb0 = ks * kb_per, kb1 = min(k_blocks_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1u;
          mbar_wait(empty_bar + stage, phase ^ 1u);
          mbar_arrive_expect_tx(full_bar + stage, L::A_BYTES + L::B_BYTES);
          tma_load_2d(sA + stage * L::A_BYTES, &tmA, full_bar + stage, kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB + stage * L::B_BYTES, &tmB, full_bar + stage, kb * GEMM_BK, n_blk * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      uint32_t it = 0;
      uint32_t lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t as = lt & 1u;
        const uint32_t aphase = (lt >> 1) & 1u;
        mbar_wait(tempty_bar + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int ks = tile / mn_tiles;
        const int kb0 = ks * kb_per, kb1 = min(k_blocks_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1u;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * L::A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * L::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in 16-byte units
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);  // frees the smem slot when these MMAs retire
        }
        umma_commit(tfull_bar + as);  // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..9) =================
    // warp%4 selects the TMEM lane quarter (hardware rule); the two warps of a quarter split the tile's columns.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    constexpr int COLS_PER_WARP = BN / 2;
    const uint32_t slab = smem_u32(sEpi) + (warp - 2) * 32 * 32 * 4;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int mn = tile % mn_tiles;
      const int n_blk = mn % n_tiles;
      const int m_blk = mn / n_tiles;
      const uint32_t as = lt & 1u;
      const uint32_t aphase = (lt >> 1) & 1u;
      mbar_wait(tfull_bar + as, aphase);
      tc_fence_after();
      const int row0 = m_blk * GEMM_BM + q * 32;
      const int cbase = ch * COLS_PER_WARP;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + cbase;
      const int gc0 = n_blk * BN + cbase;
      EpiLane L;
      epi_setup(L, slab, ep, row0, M, lane);
      // two chunks per iteration: the TMEM load of the second overlaps the stores of the first
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 64) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(taddr + c, ra);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + c + 32, rb);
        if (gc0 + c < N) epilogue_chunk<EPI_GENERIC>(ra, L, slab, ep, row0, M, gc0 + c, N, vec_ok, lane);
        tmem_ld_wait();
        if (gc0 + c + 32 < N) epilogue_chunk<EPI_GENERIC>(rb, L, slab, ep, row0, M, gc0 + c + 32, N, vec_ok, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr; int rows, cols, ld, box_rows, f16;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && f16 == o.f16;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6);
    h ^= (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (h >> 3);
    h ^= (size_t)k.ld * 0x165667B19E3779F9ull + (h << 9);
    h ^= (size_t)(k.box_rows * 4 + k.f16) * 0x27D4EB2F165667C5ull;
    return h;
  }
};

// 2-byte row-major [rows, cols] (bf16, or fp16 when is_f16) with leading dimension ld (elements);
// box = [box_rows, 64 cols], 128B swizzle (flags bit 1: [box_rows, 32 cols], 64B swizzle); flags bit 0: fp16.
// Used for operands (box_rows 128/256) and for the TMA-staged epilogues of gemm_stream.cu (box_rows 32).
int get_tmap_2b(const void* ptr, int rows, int cols, int ld, int box_rows, int flags, CUtensorMap* out) {
  const int is_f16 = flags & 1, narrow = (flags >> 1) & 1;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, ld, box_rows, flags};
  {
    std::lock_guard<std::mutex> g(mu);
    auto itr = cache.find(key);
    if (itr != cache.end()) { *out = itr->second; return HIG_OK; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(HIG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)(narrow ? 32 : GEMM_BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HIG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return HIG_OK;
}

static int get_tmap(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  return get_tmap_2b(ptr, rows, cols, ld, box_rows, 0, out);
}

int launch_gemm_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& ep,
                     int vec_ok, int num_sms, int k_splits, cudaStream_t stream);

int device_num_sms();
static int num_sms() { return device_num_sms(); }
// SMs the persistent kernels (GEMMs, attention apply) size their grids for.  hig_set_sm_limit(n) caps it: the data-parallel
// training path leaves a few SMs to NCCL so that the gradient all-reduce of a segment runs BESIDE the next segment's backward
// kernels instead of waiting behind persistent grids that own every SM.
static int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }
int device_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  if (g_sm_limit > 0 && g_sm_limit < n) return g_sm_limit & ~1;   // CTA pairs: even
  return n;
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& ep,
                       int vec_ok, int k_splits, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * k_splits;
  if (grid > num_sms()) grid = num_sms();
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, stream, tmA, tmB, M, N, K, ep, vec_ok, k_splits);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

// out_pre / gate extras of the training fusions: vector path only (the scalar edge path does not implement them)
static int apply_fused(GemmEpilogue& ep, const GemmEpilogue& f, int vec_ok, int N) {
  if (!f.out_pre && !f.gate) return HIG_OK;
  auto bad = [](const void* p, int ld) { return p && ((reinterpret_cast<uintptr_t>(p) & 7) || (ld % 4)); };
  if (!vec_ok || (N % 32) != 0 || bad(f.out_pre, f.ldo_pre) || bad(f.gate, f.ld_gate) || ep.atomic)
    return set_error(HIG_ERR_UNSUPPORTED, "gemm_fused: needs N % 32 == 0, 8-byte aligned rows and no split-K");
  if (f.gate && (f.gate_act < 1 || f.gate_act > 2)) return set_error(HIG_ERR_INVALID, "gemm_fused: gate_act must be 1 (GELU) or 2 (SiLU)");
  ep.out_pre = f.out_pre; ep.ldo_pre = f.ldo_pre; ep.gate = f.gate; ep.ld_gate = f.ld_gate; ep.gate_act = f.gate_act;
  return HIG_OK;
}

// split_k != 0: the K range is cut into slices that run as independent tiles and are combined with fp32 atomics into
// out_f32 (which the caller has initialised: zeros, or a gradient to accumulate into).  Used by the weight-gradient
// GEMMs, whose output is a handful of tiles while K = tokens is long.  split_k > 0 forces that many slices.
static int gemm_bf16_impl(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                          const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, void* out_bf16,
                          int ldo_bf16, int act, int split_k, cudaStream_t stream, const void* residual16 = nullptr,
                          int ldr16 = 0, void* out16 = nullptr, int ldo16 = 0, const GemmEpilogue* fused = nullptr) {
  if (!A || !W || M <= 0 || N <= 0 || K <= 0) return set_error(HIG_ERR_INVALID, "gemm: null operand or empty shape");
  if (!out_f32 && !out_bf16 && !out16) return set_error(HIG_ERR_INVALID, "gemm: no output");
  if ((lda % 8) || (ldw % 8)) return set_error(HIG_ERR_INVALID, "gemm: lda/ldw must be multiples of 8 (TMA 16B rule)");
  if (split_k && (bias || residual || out_bf16 || act != 0 || !out_f32))
    return set_error(HIG_ERR_INVALID, "gemm: split-K accumulates raw products into out_f32 only");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(HIG_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  if ((act < 0 || act > 3) && act != 100 && act != 101) return set_error(HIG_ERR_INVALID, "gemm: bad activation");

  GemmEpilogue ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = ldr; ep.res_row_mod = res_row_mod;
  ep.out_f32 = out_f32; ep.ldo_f32 = ldo_f32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); ep.ldo_bf16 = ldo_bf16;
  ep.act = act;
  ep.atomic = split_k ? 1 : 0;
  ep.residual16 = reinterpret_cast<const __half*>(residual16); ep.ldr16 = ldr16;
  ep.out16 = reinterpret_cast<__half*>(out16); ep.ldo16 = ldo16;

  int vec_ok = 1;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) vec_ok = 0;
  if (residual && ((reinterpret_cast<uintptr_t>(residual) & 15) || (ldr % 4))) vec_ok = 0;
  if (out_f32 && ((reinterpret_cast<uintptr_t>(out_f32) & 15) || (ldo_f32 % 4))) vec_ok = 0;
  if (out_bf16 && ((reinterpret_cast<uintptr_t>(out_bf16) & 7) || (ldo_bf16 % 4))) vec_ok = 0;
  if (fused) {
    const int rc_f = apply_fused(ep, *fused, vec_ok, N);
    if (rc_f) return rc_f;
  }
  if (residual16 && ((reinterpret_cast<uintptr_t>(residual16) & 7) || (ldr16 % 4))) vec_ok = 0;
  if (out16 && ((reinterpret_cast<uintptr_t>(out16) & 7) || (ldo16 % 4))) vec_ok = 0;

  CUtensorMap tmA, tmB;
  // CTA-pair kernel (256 x 256 tiles, cta_group::2) for the large projections; HIG_GEMM_2CTA=0 disables it
  static const bool allow_2cta = []() { const char* e = getenv("HIG_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  auto pick_splits = [&](int mn_tiles, int workers) {
    if (!split_k) return 1;
    // floor, not ceil: tiles x slices must fit ONE wave of CTA pairs (ceil gave 76 / 84 / 80 units on 74 pairs for the
    // 512x512 / 1536x512 / 1024x512 weight gradients: a second, nearly empty wave doubled their time)
    int want = split_k > 0 ? split_k : workers / mn_tiles;
    if (want > k_blocks) want = k_blocks;
    if (want < 1) want = 1;
    const int per = (k_blocks + want - 1) / want;
    return (k_blocks + per - 1) / per;  // every slice owns at least one k-block
  };
  if (allow_2cta && M >= 512 && N >= 256) {
    int rc2 = get_tmap(A, M, K, lda, 128, &tmA);
    if (rc2) return rc2;
    rc2 = get_tmap(W, N, K, ldw, 128, &tmB);
    if (rc2) return rc2;
    const int mn = ((M + 255) / 256) * ((N + 255) / 256);
    return launch_gemm_2cta(tmA, tmB, M, N, K, ep, vec_ok, num_sms(), pick_splits(mn, num_sms() / 2), stream);
  }
  if (ep.out_pre || ep.gate || ep.act == 3)
    return set_error(HIG_ERR_UNSUPPORTED, "gemm_fused: implemented on the CTA-pair kernel only (M >= 512, N >= 256)");
  const bool big_n = N > 128;
  int rc = get_tmap(A, M, K, lda, GEMM_BM, &tmA);
  if (rc) return rc;
  rc = get_tmap(W, N, K, ldw, big_n ? 256 : 128, &tmB);
  if (rc) return rc;
  const int mn1 = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + (big_n ? 255 : 127)) / (big_n ? 256 : 128));
  const int ks1 = pick_splits(mn1, num_sms());
  if (big_n) return launch_gemm<256, 4>(tmA, tmB, M, N, K, ep, vec_ok, ks1, stream);
  return launch_gemm<128, 6>(tmA, tmB, M, N, K, ep, vec_ok, ks1, stream);
}

int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
              const float* residual, int ldr, int res_row_mod, float* out_f32, int ldo_f32, void* out_bf16,
              int ldo_bf16, int act, cudaStream_t stream) {
  return gemm_bf16_impl(A, lda, W, ldw, M, N, K, bias, residual, ldr, res_row_mod, out_f32, ldo_f32, out_bf16, ldo_bf16,
                        act, 0, stream);
}

// fp16 residual-stream variant: residual / out may each be fp32 (dtype HIG_F32) or fp16 (HIG_F16)
int gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                 const void* residual, int res_dtype, int ldr, int res_row_mod, void* out, int out_dtype, int ldo,
                 void* out_bf16, int ldo_bf16, int act, cudaStream_t stream) {
  const bool r16 = residual && res_dtype == HIG_F16, o16 = out && out_dtype == HIG_F16;
  if (residual && res_dtype != HIG_F16 && res_dtype != HIG_F32) return set_error(HIG_ERR_INVALID, "gemm: residual dtype");
  if (out && out_dtype != HIG_F16 && out_dtype != HIG_F32) return set_error(HIG_ERR_INVALID, "gemm: out dtype");
  if (r16 && res_row_mod > 0) return set_error(HIG_ERR_UNSUPPORTED, "gemm: row-modulo residual tables are fp32");
  return gemm_bf16_impl(A, lda, W, ldw, M, N, K, bias, r16 ? nullptr : static_cast<const float*>(residual), ldr,
                        res_row_mod, o16 ? nullptr : static_cast<float*>(out), ldo, out_bf16, ldo_bf16, act, 0, stream,
                        r16 ? residual : nullptr, ldr, o16 ? out : nullptr, ldo);
}

int launch_gemm_2cta_t(int ta, int tb, const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                       const GemmEpilogue& ep, int vec_ok, int num_sms, int k_splits, cudaStream_t stream);

// C[M,N] (+)= opA(A) . opB(W)^T with either operand given MN-major (contraction index on the rows):
//   trans_a: A is stored [K, M] (lda = its row pitch)      trans_b: W is stored [K, N] (ldw = its row pitch)
// Training path:  dW[n,k] = sum_m dY[m,n] X[m,k]  -> trans_a = trans_b = 1 (A = dY, W = X, contraction = tokens, split-K);
//                 dX[m,k] = sum_n dY[m,n] W[n,k]  -> trans_b = 1 (A = dY K-major as stored, W = the weight as stored).
// split_k != 0: fp32 atomic accumulation into out_f32 (caller initialises it); no bias / residual / bf16 output then.
static int gemm_bf16_t_impl(int trans_a, int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                            const float* bias, const float* residual, int ldr, float* out_f32, int ldo_f32, void* out_bf16,
                            int ldo_bf16, int act, int split_k, cudaStream_t stream, const GemmEpilogue* fused) {
  if (!trans_a && !trans_b)
    return gemm_bf16_impl(A, lda, W, ldw, M, N, K, bias, residual, ldr, 0, out_f32, ldo_f32, out_bf16, ldo_bf16, act,
                          split_k, stream, nullptr, 0, nullptr, 0, fused);
  if (!A || !W || M <= 0 || N <= 0 || K <= 0) return set_error(HIG_ERR_INVALID, "gemm_t: null operand or empty shape");
  if (!out_f32 && !out_bf16) return set_error(HIG_ERR_INVALID, "gemm_t: no output");
  if ((lda % 8) || (ldw % 8)) return set_error(HIG_ERR_INVALID, "gemm_t: lda/ldw must be multiples of 8 (TMA 16B rule)");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(HIG_ERR_INVALID, "gemm_t: operands must be 16-byte aligned");
  if (split_k && (bias || residual || out_bf16 || !out_f32))
    return set_error(HIG_ERR_INVALID, "gemm_t: split-K accumulates raw products into out_f32 only");
  GemmEpilogue ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = ldr; ep.res_row_mod = 0;
  ep.out_f32 = out_f32; ep.ldo_f32 = ldo_f32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); ep.ldo_bf16 = ldo_bf16;
  ep.act = act; ep.atomic = split_k ? 1 : 0;
  ep.residual16 = nullptr; ep.ldr16 = 0; ep.out16 = nullptr; ep.ldo16 = 0;
  int vec_ok = 1;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) vec_ok = 0;
  if (residual && ((reinterpret_cast<uintptr_t>(residual) & 15) || (ldr % 4))) vec_ok = 0;
  if (out_f32 && ((reinterpret_cast<uintptr_t>(out_f32) & 15) || (ldo_f32 % 4))) vec_ok = 0;
  if (out_bf16 && ((reinterpret_cast<uintptr_t>(out_bf16) & 7) || (ldo_bf16 % 4))) vec_ok = 0;
  if (fused) {
    const int rc_f = apply_fused(ep, *fused, vec_ok, N);
    if (rc_f) return rc_f;
  }
  CUtensorMap tmA, tmB;
  int rc = trans_a ? get_tmap(A, K, M, lda, 64, &tmA) : get_tmap(A, M, K, lda, 128, &tmA);
  if (rc) return rc;
  rc = trans_b ? get_tmap(W, K, N, ldw, 64, &tmB) : get_tmap(W, N, K, ldw, 128, &tmB);
  if (rc) return rc;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int mn = ((M + 255) / 256) * ((N + 255) / 256);
  int ks = 1;
  if (split_k) {
    const int workers = num_sms() / 2;
    int want = split_k > 0 ? split_k : workers / mn;      // one wave of CTA pairs (see gemm_bf16_impl)
    if (want > k_blocks) want = k_blocks;
    if (want < 1) want = 1;
    const int per = (k_blocks + want - 1) / want;
    ks = (k_blocks + per - 1) / per;
  }
  return launch_gemm_2cta_t(trans_a, trans_b, tmA, tmB, M, N, K, ep, vec_ok, num_sms(), ks, stream);
}

int gemm_bf16_t(int trans_a, int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                const float* bias, const float* residual, int ldr, float* out_f32, int ldo_f32, void* out_bf16,
                int ldo_bf16, int split_k, cudaStream_t stream) {
  return gemm_bf16_t_impl(trans_a, trans_b, A, lda, W, ldw, M, N, K, bias, residual, ldr, out_f32, ldo_f32, out_bf16,
                          ldo_bf16, 0, split_k, stream, nullptr);
}

// Training fusions around an activation (models/interaction_transformer.py:261-264, the FFN's linear1 -> GELU -> linear2):
//   forward :  pre = A W^T + bias -> out_pre (bf16);  out = act(pre as stored)                     (act 3: erf-form GELU)
//   backward:  out = (A W) * act'(gate)   with W as stored (trans_b = 1), gate = the saved pre-activation
int gemm_bf16_fused(int trans_b, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                    int act, void* out_bf16, int ldo_bf16, void* out_pre_bf16, int ldo_pre, const void* gate_bf16,
                    int ld_gate, int gate_act, cudaStream_t stream) {
  if (!out_bf16) return set_error(HIG_ERR_INVALID, "gemm_fused: out_bf16 required");
  if (act != 0 && act != 1 && act != 2) return set_error(HIG_ERR_INVALID, "gemm_fused: act must be 0, 1 (GELU) or 2 (SiLU)");
  if (!bias) return set_error(HIG_ERR_INVALID, "gemm_fused: bias required (pass zeros)");
  GemmEpilogue f;
  f.out_pre = reinterpret_cast<__nv_bfloat16*>(out_pre_bf16); f.ldo_pre = ldo_pre;
  f.gate = reinterpret_cast<const __nv_bfloat16*>(gate_bf16); f.ld_gate = ld_gate; f.gate_act = gate_act;
  return gemm_bf16_t_impl(0, trans_b, A, lda, W, ldw, M, N, K, bias, nullptr, 0, nullptr, 0, out_bf16, ldo_bf16,
                          act == 1 ? 3 : act, 0, stream, &f);
}

int gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32, int ldo_f32,
                     int k_splits, cudaStream_t stream) {
  return gemm_bf16_impl(A, lda, W, ldw, M, N, K, nullptr, nullptr, 0, 0, out_f32, ldo_f32, nullptr, 0, 0,
                        k_splits > 0 ? k_splits : -1, stream);
}

}  // namespace hig
