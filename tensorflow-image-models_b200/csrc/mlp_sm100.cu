// Fused MLP block for narrow stages (C = 96, 128, 192 or 256 channels):
//
//     out[M,C] = residual[M,C] + gamma[C] * ( act(A[M,C] @ W1[H,C]^T + b1[H]) @ W2[C,H]^T + b2[C] )
//
// Reference: MLP.call (tfimm/layers/transformers.py:208-214: fc1 -> act -> fc2) inside ConvNeXtBlock.call
// (tfimm/architectures/convnext.py:219-228: x = mlp(norm(dw(x))); x = shortcut + gamma * x) and
// SwinTransformerBlock.call (tfimm/architectures/swin.py:315-318: x = x + mlp(norm2(x))).  A is the normalised
// bf16 activation (written by the depthwise + LayerNorm kernel / the LayerNorm kernel), the residual stream is fp32.
//
// As two GEMMs the hidden tensor [M, 4C] makes a round trip through HBM: at ConvNeXt-B / Swin-B stage 0
// (M = 802,816, C = 128) that is 1.64 GB of the 2.67 GB the block moves, and both GEMMs are HBM-bound (567 us for the
// pair, tools/bench_gemm.py).  Here a CTA pair owns 256 rows and walks the hidden dimension in chunks of 128:
//
//   fc1   D1[256 x 128]  = A[256 x C] W1_j^T            tcgen05.mma cta_group::2, A / W1 chunk from shared memory
//   act   16 warps per CTA: D1 (tensor memory) -> + b1 -> activation -> bf16 -> back into the first half of the SAME
//         tensor-memory columns, laid out as the A operand of the next product (32 rows x 16 k per 8 columns)
//   fc2   D2[256 x C]   += H_j[256 x 128] W2_j^T         A operand from tensor memory, W2 chunk from shared memory
//   out   after the last chunk the same warps drain D2: + b2, * gamma, + residual (TMA-prefetched), TMA store
//
// so the hidden activations never leave the SM.  Tensor memory: two D1/H buffers of 128 columns + D2 (C columns).
// Issue order of the single MMA thread: fc1(0) fc1(1) | fc2(c) fc1(c+2) ...: the tensor pipe executes in issue order,
// so fc1(c+2) overwrites buffer c & 1 only after fc2(c) has read H from it, and the activation of chunk c+1 runs while
// fc2(c) and fc1(c+2) execute.  The TMA producer issues its loads in exactly the order the MMA thread consumes them
// (W1 two chunks ahead of W2), through separate rings for the A tile, the W1 chunks and the W2 chunks.
#include "gemm_epilogue.cuh"

namespace tfimm {
namespace {

constexpr int kMlpRows = 128;      // rows per CTA
constexpr int kMlpPairRows = 256;  // rows per CTA pair (UMMA M)
constexpr int kHC = 128;           // hidden units per chunk (UMMA N of fc1, K of fc2)
constexpr int kMlpHidWarps = 16;
constexpr int kMlpThreads = 32 * (2 + kMlpHidWarps);

template <int C>
struct MlpCfg {
  static constexpr int kKB1 = (C + 63) / 64;                 // k-blocks of fc1 (C = 96: the second one is half TMA zero fill)
  static constexpr int kABytes = kKB1 * kMlpRows * 128;      // one A tile: 32 / 48 / 64 KB
  static constexpr int kABufs = C <= 128 ? 2 : 1;
  static constexpr int kW1Bytes = kKB1 * (kHC / 2) * 128;    // this CTA's 64 rows of a W1 chunk: 16 / 24 / 32 KB
  static constexpr int kW2Bytes = 2 * (C / 2) * 128;         // this CTA's C/2 rows of a W2 chunk (2 k-blocks): 12 .. 32 KB
  static constexpr int kWStages = C <= 192 ? 3 : 2;
  static constexpr int kOutWarps = C == 128 ? 16 : (C == 96 ? 12 : 8);   // warps that also drain D2 (C / 32 chunks x 4)
  static constexpr int kOutGroups = kOutWarps / 4;
  static constexpr int kSlabTotal = kOutWarps * kEpiSlabBytes;
  static constexpr int kNumBars = 2 * kABufs + 4 * kWStages + 2 + 2 + 2 + kMlpHidWarps;
  static constexpr int kSmemBytes =
      kABufs * kABytes + kWStages * (kW1Bytes + kW2Bytes) + kSlabTotal + kNumBars * 8 + 16 + 1024;
  static constexpr uint32_t kD2Col = 256;                    // D1/H buffers at columns 0 and 128, D2 from 256
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static_assert(C == 96 || C == 128 || C == 192 || C == 256, "fused MLP: 96 / 128 / 192 / 256 channels");
  static_assert((C / 32) % kOutGroups == 0, "every output warp drains the same number of 32-column chunks");
};

struct MlpParams {
  int M, C, H;
  const float* b1;
  int act;
  GemmParams out;   // bias = b2, gamma, has_res: the output epilogue is the GEMM epilogue
};

// D[tmem] (+)= A[tmem] * B[smem desc], CTA pair: each CTA's 128 rows of A sit in its own tensor memory
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                      const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_c,
                      const __grid_constant__ CUtensorMap tmap_r, const MlpParams p) {
  using Cfg = MlpCfg<C>;
  constexpr int KB1 = Cfg::kKB1;
  constexpr int S = Cfg::kWStages;
  constexpr int AB = Cfg::kABufs;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w1 = smem_a + AB * Cfg::kABytes;
  const uint32_t smem_w2 = smem_w1 + S * Cfg::kW1Bytes;
  const uint32_t smem_slabs = smem_w2 + S * Cfg::kW2Bytes;
  const uint32_t bars = smem_slabs + Cfg::kSlabTotal;
  int nb = 0;
  auto take = [&](int n) { const uint32_t a = bars + 8u * nb; nb += n; return a; };
  const uint32_t a_full = take(AB), a_empty = take(AB);
  const uint32_t w1_full = take(S), w1_empty = take(S), w2_full = take(S), w2_empty = take(S);
  const uint32_t d1_full = take(2);    // fc1 of a chunk retired (multicast commit: each CTA's copy)
  const uint32_t h_ready = take(2);    // leader's copy: all 32 activation warps of the pair have written H
  const uint32_t d2_full = take(1);    // fc2 of a tile's last chunk retired (multicast commit)
  const uint32_t d2_free = take(1);    // leader's copy: the output warps of both CTAs have read D2
  const uint32_t res_bars = take(kMlpHidWarps);
  const uint32_t tmem_ptr_smem = bars + 8u * Cfg::kNumBars;

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w1);
    prefetch_tmap(&tmap_w2);
    prefetch_tmap(&tmap_c);
    if (p.out.has_res) prefetch_tmap(&tmap_r);
    for (int i = 0; i < AB; ++i) { mbar_init(a_full + 8u * i, 1); mbar_init(a_empty + 8u * i, 1); }
    for (int i = 0; i < S; ++i) {
      mbar_init(w1_full + 8u * i, 1); mbar_init(w1_empty + 8u * i, 1);
      mbar_init(w2_full + 8u * i, 1); mbar_init(w2_empty + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(d1_full + 8u * i, 1); mbar_init(h_ready + 8u * i, 2 * kMlpHidWarps); }
    mbar_init(d2_full, 1);
    mbar_init(d2_free, 2 * Cfg::kOutWarps);
    for (int w = 0; w < kMlpHidWarps; ++w) mbar_init(res_bars + 8u * w, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc_pair<512>(tmem_ptr_smem);
  tcgen05_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int num_tiles = (p.M + kMlpPairRows - 1) / kMlpPairRows;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int my_tiles = pair_id < num_tiles ? (num_tiles - pair_id + num_pairs - 1) / num_pairs : 0;
  const int NC = p.H / kHC;                  // hidden chunks per tile
  const int total = my_tiles * NC;           // chunks this pair walks through

  if (warp_idx == 0) {
    // ------------------------------ TMA producer (both CTAs, one thread) ------------------------------
    if (lane == 0 && total > 0) {
      auto emit_a = [&](int i) {
        const int buf = i % AB;
        mbar_wait(a_empty + 8u * buf, (((uint32_t)(i / AB)) & 1u) ^ 1u);
        const uint32_t full_leader = mapa_shared(a_full + 8u * buf, 0);
        if (leader) mbar_expect_tx(a_full + 8u * buf, 2 * Cfg::kABytes);
        const int row = (pair_id + i * num_pairs) * kMlpPairRows + (int)rank * kMlpRows;
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb)
          tma_load_2d_pair(smem_a + buf * Cfg::kABytes + kb * (kMlpRows * 128), &tmap_a, full_leader, kb * 64, row);
      };
      auto emit_w1 = [&](int c) {
        const int slot = c % S, j = c % NC;
        mbar_wait(w1_empty + 8u * slot, (((uint32_t)(c / S)) & 1u) ^ 1u);
        const uint32_t full_leader = mapa_shared(w1_full + 8u * slot, 0);
        if (leader) mbar_expect_tx(w1_full + 8u * slot, 2 * Cfg::kW1Bytes);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb)
          tma_load_2d_pair(smem_w1 + slot * Cfg::kW1Bytes + kb * ((kHC / 2) * 128), &tmap_w1, full_leader, kb * 64,
                           j * kHC + (int)rank * (kHC / 2));
      };
      auto emit_w2 = [&](int c) {
        const int slot = c % S, j = c % NC;
        mbar_wait(w2_empty + 8u * slot, (((uint32_t)(c / S)) & 1u) ^ 1u);
        const uint32_t full_leader = mapa_shared(w2_full + 8u * slot, 0);
        if (leader) mbar_expect_tx(w2_full + 8u * slot, 2 * Cfg::kW2Bytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d_pair(smem_w2 + slot * Cfg::kW2Bytes + kb * ((C / 2) * 128), &tmap_w2, full_leader,
                           j * kHC + kb * 64, (int)rank * (C / 2));
      };
      emit_a(0);
      emit_w1(0);
      if (total > 1) emit_w1(1);
      for (int c = 0; c < total; ++c) {
        emit_w2(c);
        if (c + 2 < total) {
          if ((c + 2) % NC == 0) emit_a((c + 2) / NC);
          emit_w1(c + 2);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------- MMA issuer (leader CTA, one thread) -------------------------------
    if (leader && lane == 0 && total > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16_f32(kMlpPairRows, kHC);
      constexpr uint32_t idesc2 = umma_idesc_bf16_f32(kMlpPairRows, C);
      auto fc1 = [&](int c) {
        const int i = c / NC, j = c - i * NC, slot = c % S, buf = i % AB;
        if (j == 0) mbar_wait(a_full + 8u * buf, ((uint32_t)(i / AB)) & 1u);
        mbar_wait(w1_full + 8u * slot, ((uint32_t)(c / S)) & 1u);
        tcgen05_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)((c & 1) * kHC);
        const uint64_t da0 = umma_desc_k_sw128(smem_a + buf * Cfg::kABytes);
        const uint64_t db0 = umma_desc_k_sw128(smem_w1 + slot * Cfg::kW1Bytes);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_pair(d1, da0 + (uint64_t)(kb * ((kMlpRows * 128) >> 4) + 2 * k),
                              db0 + (uint64_t)(kb * (((kHC / 2) * 128) >> 4) + 2 * k), idesc1, (uint32_t)((kb | k) != 0));
        }
        umma_commit_pair(w1_empty + 8u * slot, 0x3);
        umma_commit_pair(d1_full + 8u * (c & 1), 0x3);
        if (j == NC - 1) umma_commit_pair(a_empty + 8u * buf, 0x3);   // the last product that reads this A tile
      };
      auto fc2 = [&](int c) {
        const int i = c / NC, j = c - i * NC, slot = c % S;
        mbar_wait(h_ready + 8u * (c & 1), ((uint32_t)(c >> 1)) & 1u);
        mbar_wait(w2_full + 8u * slot, ((uint32_t)(c / S)) & 1u);
        if (j == 0) mbar_wait(d2_free, (((uint32_t)i) & 1u) ^ 1u);    // the previous tile's D2 has been drained
        tcgen05_fence_after();
        const uint32_t d2 = tmem_base + Cfg::kD2Col;
        const uint32_t h0 = tmem_base + (uint32_t)((c & 1) * kHC);
        const uint64_t db0 = umma_desc_k_sw128(smem_w2 + slot * Cfg::kW2Bytes);
#pragma unroll
        for (int s = 0; s < kHC / 16; ++s)   // 16 hidden units per step: 8 packed columns written by warp group s / 2
          umma_bf16_ts_pair(d2, h0 + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)),
                            db0 + (uint64_t)((s >> 2) * (((C / 2) * 128) >> 4) + 2 * (s & 3)), idesc2,
                            (uint32_t)((j | s) != 0));
        umma_commit_pair(w2_empty + 8u * slot, 0x3);
        if (j == NC - 1) umma_commit_pair(d2_full, 0x3);
      };
      fc1(0);
      if (total > 1) fc1(1);
      for (int c = 0; c < total; ++c) {
        fc2(c);
        if (c + 2 < total) fc1(c + 2);
      }
    }
  } else {
    // ------------------------------ activation + output warps (both CTAs) ------------------------------
    const int ew = warp_idx - 2;
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may access
    const int grp = ew >> 2;             // 32-column group of the hidden chunk
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t slab = smem_slabs + (uint32_t)(ew % Cfg::kOutWarps) * kEpiSlabBytes;
    uint8_t* my_row = smem_gen + (slab - smem_base) + lane * 128;
    const uint32_t h_ready_leader0 = mapa_shared(h_ready, 0), d2_free_leader = mapa_shared(d2_free, 0);
    uint32_t cc = 0;   // output chunks this warp has stored (residual barrier parity)
    int i = 0, j = 0;  // tile, chunk of the tile
    for (int c = 0; c < total; ++c, j = (j + 1 == NC ? 0 : j + 1), i += (j == 0)) {
      mbar_wait(d1_full + 8u * (c & 1), ((uint32_t)(c >> 1)) & 1u);
      tcgen05_fence_after();
      const uint32_t t_d1 = tmem_base + lane_off + (uint32_t)((c & 1) * kHC + grp * 32);
      uint64_t v[16];
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_d1, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = pack2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
      }
      if (p.b1 != nullptr) apply_vec<32, false>(v, p.b1, j * kHC + grp * 32, p.H);
#ifdef TFIMM_MLP_DIRECT_RCP
      apply_act_pairs<16, false>(v, p.act);
#else
      apply_act_pairs<16, true>(v, p.act);   // these warps are bound by the MUFU pipe: 1.25 instead of 2 per element
#endif
      // bf16 pairs into the first 16 of this warp's own 32 columns: two 16-unit k-steps of the second product
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a, b;
          unpack2(v[8 * h + e], a, b);
          pk[e] = pack_bf16x2(a, b);
        }
        tmem_st_32x32b_x8(t_d1 + (uint32_t)(8 * h), pk);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(h_ready_leader0 + 8u * (c & 1));
      if (j == NC - 1 && ew < Cfg::kOutWarps) {
        // ---- the tile's output: D2 -> + b2, * gamma, + residual -> TMA store ----
        mbar_wait(d2_full, ((uint32_t)i) & 1u);
        tcgen05_fence_after();
        const int row0 = (pair_id + i * num_pairs) * kMlpPairRows + (int)rank * kMlpRows + q * 32;
        const uint32_t t_d2 = tmem_base + lane_off + Cfg::kD2Col;
#pragma unroll 1
        for (int ch = grp; ch < C / 32; ch += Cfg::kOutGroups) {
          const bool last = ch + Cfg::kOutGroups >= C / 32;
          epilogue_chunk<float, 32>(p.out, t_d2 + (uint32_t)(ch * 32), ch * 32, row0, slab, my_row, lane,
                                    res_bars + 8u * ew, cc & 1u, &tmap_c, &tmap_r, /*ct=*/nullptr, [&]() {
                                      if (last) {
                                        tcgen05_fence_before();
                                        __syncwarp();
                                        if (lane == 0) mbar_arrive_cluster(d2_free_leader);
                                      }
                                    });
          ++cc;
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <int C>
int launch_mlp(const void* A, int lda, const void* W1, int ldw1, const void* W2, int ldw2, const void* residual, int ldr,
               void* out, int ldc, const MlpParams& p, cudaStream_t stream) {
  using Cfg = MlpCfg<C>;
  CUtensorMap ta, tw1, tw2, tc, tr;
  int st;
  if ((st = make_tmap_2d(&ta, A, kBF16, p.M, C, lda, kMlpRows, 64, "A")) != kOk) return st;
  if ((st = make_tmap_2d(&tw1, W1, kBF16, p.H, C, ldw1, kHC / 2, 64, "W1")) != kOk) return st;
  if ((st = make_tmap_2d(&tw2, W2, kBF16, C, p.H, ldw2, C / 2, 64, "W2")) != kOk) return st;
  if ((st = make_tmap_2d(&tc, out, kF32, p.M, C, ldc, 32, 32, "out", 128)) != kOk) return st;
  if (residual != nullptr) {
    if ((st = make_tmap_2d(&tr, residual, kF32, p.M, C, ldr, 32, 32, "residual", 128)) != kOk) return st;
  } else {
    tr = tc;
  }
  auto kernel = mlp_fused_pair_kernel<C>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs))
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int tiles = (p.M + kMlpPairRows - 1) / kMlpPairRows;
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  kernel<<<2 * pairs, kMlpThreads, Cfg::kSmemBytes, stream>>>(ta, tw1, tw2, tc, tr, p);
  TFIMM_LAUNCH_OK("mlp_fused_pair_kernel");
  return kOk;
}

}  // namespace

// Returns kUnsupported for shapes outside the kernel (the caller then runs the two GEMMs).
int mlp_fused_bf16(const void* A, int lda, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                   const float* b2, const float* gamma, const void* residual, int ldr, void* out, int ldc, int M, int C,
                   int H, int act, cudaStream_t stream) {
  if ((C != 96 && C != 128 && C != 192 && C != 256) || H % kHC != 0 || H < 2 * kHC || M < 1) {
    set_last_error("mlp_fused: needs C in {96, 128, 192, 256} and hidden %% 128 == 0, >= 256 (got C=%d hidden=%d)", C, H);
    return kUnsupported;
  }
  MlpParams p{};
  p.M = M;
  p.C = C;
  p.H = H;
  p.b1 = b1;
  p.act = act;
  p.out.M = M;
  p.out.N = C;
  p.out.K = H;
  p.out.bias = b2;
  p.out.gamma = gamma;
  p.out.act = kActNone;
  p.out.has_res = residual != nullptr;
  switch (C) {
    case 96: return launch_mlp<96>(A, lda, W1, ldw1, W2, ldw2, residual, ldr, out, ldc, p, stream);
    case 128: return launch_mlp<128>(A, lda, W1, ldw1, W2, ldw2, residual, ldr, out, ldc, p, stream);
    case 192: return launch_mlp<192>(A, lda, W1, ldw1, W2, ldw2, residual, ldr, out, ldc, p, stream);
    default: return launch_mlp<256>(A, lda, W1, ldw1, W2, ldw2, residual, ldr, out, ldc, p, stream);
  }
}

}  // namespace tfimm
