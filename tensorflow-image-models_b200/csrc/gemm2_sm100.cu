// CTA-pair variant of the dense contraction (see gemm_sm100.cu for the operator and its call sites):
//
//     C[M,N] = residual[M,N] + gamma[N] * act(A[M,K] @ W[N,K]^T + bias[N])
//
// Two CTAs on the two SMs of a TPC (cluster 2x1x1) compute one 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256, N = 256, K = 16):
//   * each CTA TMA-loads its own 128 rows of A and its own 128 rows (N half) of W per k-block, so the shared
//     memory traffic per SM per MMA is 2/3 of the one-CTA 128 x 256 kernel's (the operand-bandwidth limit of that
//     kernel) and every W tile is fetched from L2 once per 256 output rows instead of once per 128
//   * all TMA completions are credited to the LEADER CTA's "full" barrier; the leader's MMA warp issues the
//     pair-wide MMAs; tcgen05.commit multicasts the "slot free" and "accumulator ready" arrivals to both CTAs
//   * each CTA's TMEM holds its 128 rows x 256 columns of the fp32 accumulator (two stages: 512 columns); the
//     epilogue warps of each CTA drain their own half (same code as the one-CTA kernel) and release the
//     accumulator stage with a relaxed remote arrive on the leader's barrier
//   * bf16 output (qkv, fc1 + GELU: the instances whose tile time is set by the EPILOGUE, not the tensor pipe, when
//     K <= 768): SIXTEEN epilogue warps, four per SM sub-partition, on 32-column chunks (64-byte rows, 2 KB slabs,
//     64B-swizzled TMA stores) so that two loads / activations / stores are in flight per sub-partition scheduler
//     and each warp's registers stay under the 576-thread budget.  fp32 output (residual GEMMs): eight warps.
// Warp roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader only) + TMEM alloc, warps 2.. epilogue.
// Pairs are persistent over the tile list.
#include "gemm_epilogue.cuh"

namespace tfimm {
namespace {

constexpr int kCtaM = 128;     // rows per CTA
constexpr int kPairM = 256;    // rows per CTA pair (UMMA M)
constexpr int kBlockN = 256;   // UMMA N
constexpr int kHalfN = 128;    // W rows loaded by each CTA
constexpr int kBlockK = 64;    // 64 bf16 = one 128-byte swizzle span
constexpr int kUmmaK = 16;
constexpr int kAccStages = 2;
constexpr int kABytes = kCtaM * kBlockK * 2;
constexpr int kBBytes = kHalfN * kBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;  // per CTA
constexpr uint32_t kTmemCols = kAccStages * kBlockN;  // 512: all of this SM's tensor memory
constexpr int kCH = 32;        // output columns per epilogue chunk (both output types)

// kWideEpi (fp32 output only): sixteen epilogue warps instead of eight.  With a short contraction (K <= 1024: the proj
// GEMMs, 128 x 256 fp32 outputs + as many residual values per CTA every ~3 us) the eight-warp epilogue is the limiter
// (ViT-B proj 79.5 -> 69.6 us); at K = 3072 the MMA time hides it either way and the extra warps cost ~1 %.
template <typename OutT, bool kWideEpi = false>
struct PairCfg {
  static constexpr int kEpiGroups = (sizeof(OutT) == 2 || kWideEpi) ? 4 : 2;  // epilogue warps per TMEM lane quarter
  static constexpr int kNumEpiWarps = 4 * kEpiGroups;
  static constexpr int kNumThreads = 32 * (2 + kNumEpiWarps);
  static constexpr int kSlabBytes = 32 * kCH * (int)sizeof(OutT);                     // 2 KB (bf16) / 4 KB (fp32)
  static constexpr int kSlabTotal = kNumEpiWarps * kSlabBytes;
  static constexpr int kStages = 5;
  static constexpr int kNumBarriers = 2 * kStages + 2 * kAccStages + kNumEpiWarps;
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlabTotal + kNumBarriers * 8 + 16 + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

template <typename OutT, bool kWideEpi = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<OutT, kWideEpi>::kNumThreads, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                              const __grid_constant__ CUtensorMap tmap_b,
                              const __grid_constant__ CUtensorMap tmap_c,
                              const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  using Cfg = PairCfg<OutT, kWideEpi>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kNumEpiWarps = Cfg::kNumEpiWarps;
  constexpr int kEpiGroups = Cfg::kEpiGroups;
  constexpr int NCH = kBlockN / kCH;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (identical offsets in both CTAs: same kernel, same smem layout)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_tiles = smem_base;
  const uint32_t smem_slabs = smem_base + kStages * kStageBytes;
  const uint32_t smem_bars = smem_slabs + Cfg::kSlabTotal;
  auto full_bar = [&](int s) { return smem_bars + 8u * s; };
  auto empty_bar = [&](int s) { return smem_bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return smem_bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return smem_bars + 8u * (2 * kStages + kAccStages + s); };
  auto res_bar = [&](int w) { return smem_bars + 8u * (2 * kStages + 2 * kAccStages + w); };
  const uint32_t tmem_ptr_smem = smem_bars + 8u * Cfg::kNumBarriers;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of smem_base

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    if (p.has_res) prefetch_tmap(&tmap_r);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader's copy is the one in use: its producer's arrive.expect_tx
      mbar_init(empty_bar(s), 1);  // multicast tcgen05.commit
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(tfull_bar(s), 1);                    // multicast tcgen05.commit
      mbar_init(tempty_bar(s), 2 * kNumEpiWarps);    // leader's copy: epilogue warps of both CTAs
    }
    for (int w = 0; w < kNumEpiWarps; ++w) mbar_init(res_bar(w), 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc_pair<kTmemCols>(tmem_ptr_smem);
  // barriers of BOTH CTAs must be initialised before any remote arrive / TMA credit / multicast commit
  tcgen05_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int num_m_tiles = (p.M + kPairM - 1) / kPairM;
  const int num_n_tiles = (p.N + kBlockN - 1) / kBlockN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_k_blocks = (p.K + kBlockK - 1) / kBlockK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp_idx == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        const int m_blk = t / num_n_tiles, n_blk = t % num_n_tiles;
        const int a_row = m_blk * kPairM + (int)rank * kCtaM;
        const int b_row = n_blk * kBlockN + (int)rank * kHalfN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_tiles + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          const uint32_t full_leader = mapa_shared(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
          tma_load_2d_pair(sa, &tmap_a, full_leader, kb * kBlockK, a_row);
          tma_load_2d_pair(sb, &tmap_b, full_leader, kb * kBlockK, b_row);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------- MMA issuer (leader CTA) -------------------------------
    // ONE thread: the loop is ~20 instructions per k-block, and its issue latency is what the tensor pipe waits on when
    // the sixteen epilogue warps are busy -- no warp-wide waits, no reconvergence points.
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(kPairM, kBlockN);
      const uint64_t da0 = umma_desc_k_sw128(smem_tiles), db0 = umma_desc_k_sw128(smem_tiles + kABytes);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBlockN);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t off = (uint64_t)((uint32_t)(stage * kStageBytes) >> 4);
          const uint64_t da = da0 + off, db = db0 + off;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes per UMMA_K step inside the swizzle span -> +2 in the (addr>>4) field
            umma_bf16_ss_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                              (uint32_t)((kb | k) != 0));
          }
          umma_commit_pair(empty_bar(stage), 0x3);                          // frees the slot in both CTAs
          if (kb == num_k_blocks - 1) umma_commit_pair(tfull_bar(acc), 0x3);  // accumulator ready, both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // -------------------------------- epilogue (both CTAs) --------------------------------
    const int ew = warp_idx - 2;         // slab / residual-barrier owner
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may access
    const int grp = ew >> 2;             // which share of the column chunks this warp takes
    const uint32_t slab = smem_slabs + (uint32_t)ew * Cfg::kSlabBytes;
    uint8_t* my_row = smem_gen + (slab - smem_base) + lane * (Cfg::kSlabBytes / 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t cc = 0;  // chunks processed by this warp (residual barrier parity)
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      const int m_blk = t / num_n_tiles, n_blk = t % num_n_tiles;
      const int row0 = m_blk * kPairM + (int)rank * kCtaM + q * 32;
      const int cols_left = p.N - n_blk * kBlockN;
      const int nvalid = cols_left >= kBlockN ? NCH : (cols_left + kCH - 1) / kCH;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * kBlockN) + ((uint32_t)(q * 32) << 16);
      const uint32_t tempty_leader = mapa_shared(tempty_bar(acc), 0);
      auto release_acc = [&]() {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader);
      };
      if (grp >= nvalid) release_acc();  // nothing to read for this warp in this tile
#pragma unroll 1
      for (int c = grp; c < nvalid; c += kEpiGroups) {
        const int n0 = n_blk * kBlockN + c * kCH;
        const bool last = c + kEpiGroups >= nvalid;
        epilogue_chunk<OutT, kCH>(p, t_acc + (uint32_t)(c * kCH), n0, row0, slab, my_row, lane, res_bar(ew), cc & 1u,
                                  &tmap_c, &tmap_r, /*ct=*/nullptr, [&]() { if (last) release_acc(); });
        ++cc;
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // both CTAs are done with the pair's tensor memory and with each other's shared memory
  tcgen05_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

template <typename OutT, bool kWideEpi = false>
int launch_gemm_pair(const void* A, int lda, const void* W, int ldw, const void* residual, int ldr, void* C, int ldc,
                     const GemmParams& p, cudaStream_t stream) {
  using Cfg = PairCfg<OutT, kWideEpi>;
  const int M = p.M, N = p.N, K = p.K;
  constexpr int out_dtype = sizeof(OutT) == 2 ? kBF16 : kF32;
  constexpr int row_bytes = kCH * (int)sizeof(OutT);  // 64 (bf16) or 128 (fp32): also the TMA swizzle span
  CUtensorMap ta, tb, tc, tr;
  int st;
  if ((st = make_tmap_2d(&ta, A, kBF16, M, K, lda, kCtaM, kBlockK, "A")) != kOk) return st;
  if ((st = make_tmap_2d(&tb, W, kBF16, N, K, ldw, kHalfN, kBlockK, "W")) != kOk) return st;
  if ((st = make_tmap_2d(&tc, C, out_dtype, M, N, ldc, 32, kCH, "C", row_bytes)) != kOk) return st;
  if (residual != nullptr) {
    if ((st = make_tmap_2d(&tr, residual, out_dtype, M, N, ldr, 32, kCH, "residual", row_bytes)) != kOk) return st;
  } else {
    tr = tc;
  }
  auto kernel = gemm_bf16_tcgen05_pair_kernel<OutT, kWideEpi>;
  static unsigned long long attr_devs = 0;  // per instantiation
  if (first_use_on_device(attr_devs)) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const int tiles = ((M + kPairM - 1) / kPairM) * ((N + kBlockN - 1) / kBlockN);
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  kernel<<<2 * pairs, Cfg::kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, tr, p);
  TFIMM_LAUNCH_OK("gemm_bf16_tcgen05_pair_kernel");
  return kOk;
}

}  // namespace

int gemm_bf16_pair(const void* A, int lda, const void* W, int ldw, const void* residual, int ldr, void* C, int ldc,
                   const GemmParams& p, int out_dtype, cudaStream_t stream) {
  if (out_dtype == kBF16) return launch_gemm_pair<__nv_bfloat16>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
  return p.K <= 1024 ? launch_gemm_pair<float, true>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream)
                     : launch_gemm_pair<float, false>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
}

}  // namespace tfimm
