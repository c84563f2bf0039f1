// Dense contraction for the tfimm forward path on sm_100a:
//
//     C[M,N] = residual[M,N] + gamma[N] * act(A[M,K] @ W[N,K]^T + bias[N])
//
// This single kernel replaces every tf.keras.layers.Dense and 1x1 Conv2D the
// reference calls on the hot path (qkv/proj: tfimm/architectures/vit.py:142-146,
// swin.py:124-128; fc1/fc2: tfimm/layers/transformers.py:192-205; heads:
// vit.py:364-368, convnext.py:356-360; 1x1 convs: efficientnet_blocks.py:412-434,
// resnet.py:220-248) plus the patchify convolutions once their input has been
// gathered (layers/transformers.py:131-139, convnext.py:259-266,319-326).
//
// Design (one persistent CTA per SM, warp-specialised, 320 threads):
//   warp 0      TMA producer: A/W tiles -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer: tcgen05.mma 128 x BLOCK_N x 16, fp32 accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of i+1
//   warps 2..9  epilogue, two warps per TMEM lane quarter taking alternate 128-byte column
//               chunks: tcgen05.ld -> bias/act/gamma (FFMA2) -> + residual -> swizzled smem slab
//               -> TMA store.  Each warp owns its 32-row slab, so there is no cross-warp barrier;
//               the residual chunk is TMA-prefetched into the same slab while the accumulator is
//               being loaded and activated, and may alias the output (in-place residual stream).
#include "gemm_epilogue.cuh"

namespace tfimm {

int gemm_bf16_pair(const void* A, int lda, const void* W, int ldw, const void* residual, int ldr, void* C, int ldc,
                   const GemmParams& p, int out_dtype, cudaStream_t stream);

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle span
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 32 * (2 + kNumEpiWarps);
constexpr int kSlabBytes = kEpiSlabBytes;
constexpr int kAccStages = 2;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 5 : 7);
  static constexpr int kSlabTotal = kNumEpiWarps * kSlabBytes;
  static constexpr int kNumBarriers = 3 * kStages + 2 * kAccStages + kNumEpiWarps;   // full, empty, ready (gated)
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kSlabTotal + kNumBarriers * 8 + 16 + 1024 /*alignment slack*/;
  static constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;  // 512 / 256 / 128
};

constexpr int kNumGateWarps = 4;   // gated instances only: one thread per A-tile row

template <int BLOCK_N, typename OutT, bool kGated = false>
__global__ void __launch_bounds__(kNumThreads + (kGated ? 32 * kNumGateWarps : 0), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c,
                         const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int CH = 128 / (int)sizeof(OutT);  // output columns per 128-byte slab row
  constexpr int NCH = BLOCK_N / CH;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_tiles = smem_base;
  const uint32_t smem_slabs = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t smem_bars = smem_slabs + Cfg::kSlabTotal;
  auto full_bar = [&](int s) { return smem_bars + 8u * s; };
  auto empty_bar = [&](int s) { return smem_bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return smem_bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return smem_bars + 8u * (2 * kStages + kAccStages + s); };
  auto res_bar = [&](int w) { return smem_bars + 8u * (2 * kStages + 2 * kAccStages + w); };
  auto ready_bar = [&](int s) { return smem_bars + 8u * (2 * kStages + 2 * kAccStages + kNumEpiWarps + s); };
  const uint32_t tmem_ptr_smem = smem_bars + 8u * Cfg::kNumBarriers;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of smem_base

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    if (p.has_res) prefetch_tmap(&tmap_r);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kNumEpiWarps);
    }
    for (int w = 0; w < kNumEpiWarps; ++w) mbar_init(res_bar(w), 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int num_m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int num_n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_k_blocks = (p.K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / num_n_tiles, n_blk = t % num_n_tiles;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_tiles + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (p.conv == 0) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBlockK, m_blk * kBlockM);
          } else {
            // implicit convolution: tap (ky, kx) and a 64-channel slice of the input patch; padding = OOB zero fill
            const int tap = kb / p.cv_cblocks, cb = kb - tap * p.cv_cblocks;
            const int ky = tap / p.cv_ks, kx = tap - ky * p.cv_ks;
            const int tx = m_blk % p.cv_tiles_x, tyb = m_blk / p.cv_tiles_x;
            const int ty = tyb % p.cv_tiles_y, tb = tyb / p.cv_tiles_y;
            tma_load_4d(sa, &tmap_a, full_bar(stage), cb * kBlockK, tx * p.cv_pw * p.cv_stride + kx - p.cv_pad,
                        ty * p.cv_ph * p.cv_stride + ky - p.cv_pad, tb * p.cv_pb);
          }
          tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kBlockK, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // ONE thread: its issue latency is what the tensor pipe waits on while the epilogue warps are busy
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, BLOCK_N);
    if (lane == 0) {
      const uint64_t da0 = umma_desc_k_sw128(smem_tiles), db0 = umma_desc_k_sw128(smem_tiles + Cfg::kABytes);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(kGated ? ready_bar(stage) : full_bar(stage), phase);   // gated: the A tile has been rescaled
          tcgen05_fence_after();
          const uint64_t off = (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
          const uint64_t da = da0 + off, db = db0 + off;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes per UMMA_K step inside the swizzle span -> +2 in the (addr>>4) field
            umma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));                            // frees the smem slot
          if (kb == num_k_blocks - 1) umma_commit(tfull_bar(acc));  // accumulator ready
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (kGated && warp_idx >= 2 + kNumEpiWarps) {
    // ------------------------- A-operand gate (squeeze-excite) -------------------------
    // Each of the four warps owns every fourth k-block (a whole 128 x 64 tile), so four stages are being rescaled at
    // any time and the latency of one stage's chain (shared-memory loads -> multiply -> stores -> proxy fence -> arrive)
    // is overlapped with the other three.  lane = one 16-byte chunk column (8 contraction indices) x 32 rows 4 apart:
    // the gate values of a lane change only when its rows cross into the next image (two 16-byte loads, the first pair
    // issued BEFORE the wait on the TMA barrier); a quarter-warp touches one whole 128-byte row (no bank conflicts
    // under the 128B swizzle).  x * gate in fp32, back as bf16 -- the rounding of the separate scale pass
    // (csrc/conv.cu, scale_channels_kernel) -- then the proxy fence that makes the generic-proxy writes visible to the
    // tensor core's async-proxy reads, and one arrive on the stage's "ready" barrier.
    const int wt = warp_idx - 2 - kNumEpiWarps;
    const int c = lane & 7, r0 = lane >> 3;
    auto load_gate = [&](int img, int k0, float4& ga, float4& gb) {
      const float* g = p.a_scale + (long)img * p.K + k0;
      ga = __ldg(reinterpret_cast<const float4*>(g));
      gb = __ldg(reinterpret_cast<const float4*>(g + 4));
    };
    int stage = 0, turn = 0;   // stage / owner of the NEXT k-block in the CTA's global order
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / num_n_tiles;
      const long row_first = (long)m_blk * kBlockM + r0;
      long im0 = row_first / p.a_rows_per_img;
      im0 = im0 < p.a_imgs ? im0 : p.a_imgs - 1;            // rows past M are zero-filled: any gate row will do
      const int img0 = (int)im0;
      const long bound0 = (im0 + 1) * p.a_rows_per_img - (long)m_blk * kBlockM;   // first tile row of the next image
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        if (turn == wt) {
          const int k0 = kb * kBlockK + c * 8;
          const bool valid = k0 < p.K;                       // K % 8 == 0: a chunk is inside or outside as a whole
          float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
          if (valid) load_gate(img0, k0, ga, gb);
          mbar_wait(full_bar(stage), phase);
          if (valid) {
            const uint32_t sa = smem_tiles + stage * Cfg::kStageBytes;
            int cur = img0;
            long bound = bound0;
#pragma unroll 1
            for (int b8 = 0; b8 < 4; ++b8) {
              uint4 u[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + 4 * (8 * b8 + i);
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(u[i].x), "=r"(u[i].y), "=r"(u[i].z), "=r"(u[i].w)
                             : "r"(sa + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4))));
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + 4 * (8 * b8 + i);
                if (r >= bound && cur < p.a_imgs - 1) {
                  do { ++cur; bound += p.a_rows_per_img; } while (r >= bound && cur < p.a_imgs - 1);
                  load_gate(cur, k0, ga, gb);
                }
                const float2 x0 = unpack_bf16x2(u[i].x), x1 = unpack_bf16x2(u[i].y), x2 = unpack_bf16x2(u[i].z),
                             x3 = unpack_bf16x2(u[i].w);
                const uint32_t o0 = pack_bf16x2(x0.x * ga.x, x0.y * ga.y), o1 = pack_bf16x2(x1.x * ga.z, x1.y * ga.w);
                const uint32_t o2 = pack_bf16x2(x2.x * gb.x, x2.y * gb.y), o3 = pack_bf16x2(x3.x * gb.z, x3.y * gb.w);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4))),
                             "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                             : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(ready_bar(stage));
        }
        turn = turn + 1 == kNumGateWarps ? 0 : turn + 1;
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // -------------------------------- epilogue --------------------------------
    const int ew = warp_idx - 2;         // 0..7: slab / residual-barrier owner
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may access
    const int grp = ew >> 2;             // which half of the column chunks this warp takes
    const uint32_t slab = smem_slabs + (uint32_t)ew * kSlabBytes;
    uint8_t* my_row = smem_gen + (slab - smem_base) + lane * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t cc = 0;  // chunks processed by this warp (residual barrier parity)
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / num_n_tiles, n_blk = t % num_n_tiles;
      const int row0 = m_blk * kBlockM + q * 32;
      ConvTile ctile{0, 0, 0};
      if (p.conv != 0) {
        // this warp's 32 rows are a (32 / pw) x pw pixel patch of ONE image of the tile's pb x ph x pw patch
        const int tx = m_blk % p.cv_tiles_x, tyb = m_blk / p.cv_tiles_x;
        const int ty = tyb % p.cv_tiles_y, tb = tyb / p.cv_tiles_y;
        const int per_img = p.cv_ph * p.cv_pw, r0 = q * 32;
        ctile.b = tb * p.cv_pb + r0 / per_img;
        ctile.y = ty * p.cv_ph + (r0 % per_img) / p.cv_pw;
        ctile.x = tx * p.cv_pw;
      }
      const ConvTile* ct = p.conv != 0 ? &ctile : nullptr;
      const int cols_left = p.N - n_blk * BLOCK_N;
      const int nvalid = cols_left >= BLOCK_N ? NCH : (cols_left + CH - 1) / CH;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * BLOCK_N) + ((uint32_t)(q * 32) << 16);
      if (grp >= nvalid) {
        // nothing to read for this warp in this tile: hand the accumulator back right away
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
#pragma unroll 1
      for (int c = grp; c < nvalid; c += 2) {
        const int n0 = n_blk * BLOCK_N + c * CH;
        const bool last = c + 2 >= nvalid;
        epilogue_chunk<OutT>(p, t_acc + (uint32_t)(c * CH), n0, row0, slab, my_row, lane, res_bar(ew), cc & 1u,
                             &tmap_c, &tmap_r, ct, [&]() {
                               if (last) {
                                 // all TMEM reads of this accumulator stage by this warp are done
                                 tcgen05_fence_before();
                                 __syncwarp();
                                 if (lane == 0) mbar_arrive(tempty_bar(acc));
                               }
                             });
        ++cc;
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------ host side -----------------------------------
template <int BLOCK_N, typename OutT, bool kGated = false>
int launch_gemm(const void* A, int lda, const void* W, int ldw, const void* residual, int ldr, void* C, int ldc,
                const GemmParams& p, cudaStream_t stream) {
  const int M = p.M, N = p.N, K = p.K;
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int out_dtype = sizeof(OutT) == 2 ? kBF16 : kF32;
  constexpr int CH = 128 / (int)sizeof(OutT);
  CUtensorMap ta, tb, tc, tr;
  int st;
  if ((st = make_tmap_2d(&ta, A, kBF16, M, K, lda, kBlockM, kBlockK, "A")) != kOk) return st;
  if ((st = make_tmap_2d(&tb, W, kBF16, N, K, ldw, BLOCK_N, kBlockK, "W")) != kOk) return st;
  if ((st = make_tmap_2d(&tc, C, out_dtype, M, N, ldc, 32, CH, "C")) != kOk) return st;
  if (residual != nullptr) {
    if ((st = make_tmap_2d(&tr, residual, out_dtype, M, N, ldr, 32, CH, "residual")) != kOk) return st;
  } else {
    tr = tc;
  }
  auto kernel = gemm_bf16_tcgen05_kernel<BLOCK_N, OutT, kGated>;
  static unsigned long long attr_devs = 0;  // per instantiation
  if (first_use_on_device(attr_devs)) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kernel<<<grid, kNumThreads + (kGated ? 32 * kNumGateWarps : 0), Cfg::kSmemBytes, stream>>>(ta, tb, tc, tr, p);
  TFIMM_LAUNCH_OK("gemm_bf16_tcgen05_kernel");
  return kOk;
}

int pick_block_n(int M, int N);

// Implicit k x k convolution on the tensor cores: same kernel, A tensor map = the NHWC input (rank 4, traversal
// stride = conv stride), C / residual tensor maps = the NHWC output (rank 4).  See GemmParams::conv.
template <int BLOCK_N, typename OutT>
int launch_conv(const void* x, const void* W, int ldw, const void* residual, void* out, int B, int H, int Wd, int C,
                int Ho, int Wo, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int out_dtype = sizeof(OutT) == 2 ? kBF16 : kF32;
  constexpr int CH = 128 / (int)sizeof(OutT);
  const int N = p.N, s = p.cv_stride;
  CUtensorMap ta, tb, tc, tr;
  int st;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wd, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)Wd * C * 2, (uint64_t)H * Wd * C * 2};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(p.cv_pw * s), (uint32_t)(p.cv_ph * s), (uint32_t)p.cv_pb};
    const uint32_t estr[4] = {1u, (uint32_t)s, (uint32_t)s, 1u};
    if ((st = make_tmap(&ta, x, kBF16, 4, dims, strides, box, "conv input", 128, estr)) != kOk) return st;
  }
  if ((st = make_tmap_2d(&tb, W, kBF16, N, p.K, ldw, BLOCK_N, kBlockK, "conv weights")) != kOk) return st;
  {
    const uint64_t esz = sizeof(OutT);
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)N * esz, (uint64_t)Wo * N * esz, (uint64_t)Ho * Wo * N * esz};
    const uint32_t box[4] = {(uint32_t)CH, (uint32_t)p.cv_pw, (uint32_t)(32 / p.cv_pw), 1u};
    if ((st = make_tmap(&tc, out, out_dtype, 4, dims, strides, box, "conv output")) != kOk) return st;
    if (residual != nullptr) {
      if ((st = make_tmap(&tr, residual, out_dtype, 4, dims, strides, box, "conv residual")) != kOk) return st;
    } else {
      tr = tc;
    }
  }
  auto kernel = gemm_bf16_tcgen05_kernel<BLOCK_N, OutT>;
  static unsigned long long attr_devs = 0;  // per instantiation
  if (first_use_on_device(attr_devs)) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const int tiles = (p.M / kBlockM) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kernel<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, tr, p);
  TFIMM_LAUNCH_OK("gemm_bf16_tcgen05_kernel (implicit convolution)");
  return kOk;
}

int pick_block_n(int M, int N) {
  const int sms = sm_count() > 0 ? sm_count() : 148;  // no device (host-side shape queries): B200
  const int mt = (M + kBlockM - 1) / kBlockM;
  int best = 256;
  double best_cost = 1e30;
  for (int bn : {256, 128, 64}) {
    const int nt = (N + bn - 1) / bn;
    const long tiles = (long)mt * nt;
    const long waves = (tiles + sms - 1) / sms;
    // time ~ waves * per-tile MMA time (prop. to bn); small preference for wide tiles
    const double cost = (double)waves * bn * (bn == 256 ? 1.0 : (bn == 128 ? 1.04 : 1.10));
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// CTA-pair kernel (gemm2_sm100.cu) when its 256 x 256 tiles fill the machine at least as well as the best
// one-CTA tiling: cost = waves x tile width, the pair kernel's MMA rate per SM being ~15% higher.
bool prefer_pair(int M, int N) {
  if (N < 256 || M < 256) return false;
  const int sms = sm_count() > 0 ? sm_count() : 148;  // no device (host-side shape queries): B200
  const long tiles2 = (long)((M + 255) / 256) * ((N + 255) / 256);
  const long waves2 = (tiles2 + sms / 2 - 1) / (sms / 2);
  const double cost2 = (double)waves2 * 256 * 0.87;
  const int bn = pick_block_n(M, N);
  const long tiles1 = (long)((M + kBlockM - 1) / kBlockM) * ((N + bn - 1) / bn);
  const double cost1 = (double)((tiles1 + sms - 1) / sms) * bn * (bn == 256 ? 1.0 : (bn == 128 ? 1.04 : 1.10));
  return cost2 <= cost1;
}

}  // namespace

int gemm_bf16_skinny(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, int ldr,
                     void* C, int ldc, int M, int N, int K, int act, cudaStream_t stream, const float* gate = nullptr,
                     int rows_per_img = 1, int imgs = 1);

int gemm_bf16_dispatch(const void* A, int lda, const void* W, int ldw, const float* bias,
                       const float* gamma, const void* residual, int ldr, void* C, int ldc, int M,
                       int N, int K, int act, int act_post, int out_dtype, int force_block_n, cudaStream_t stream) {
  TFIMM_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: M, N, K must be positive (got %d %d %d)", M, N, K);
  TFIMM_CHECK_ARG(out_dtype == kBF16 || out_dtype == kF32, "gemm: out_dtype must be bf16 or f32");
  TFIMM_CHECK_ARG(K % 8 == 0, "gemm: K must be a multiple of 8 (got %d)", K);
  TFIMM_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0, "gemm: bias must be 16-byte aligned");
  TFIMM_CHECK_ARG(gamma == nullptr || (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0, "gemm: gamma must be 16-byte aligned");
  if (force_block_n == 0 && K <= 64 && out_dtype == kBF16 && gamma == nullptr && act_post == 0) {
    // short contraction: streaming mma.sync kernel (gemm_skinny.cu); kUnsupported = shape outside its envelope
    const int st = gemm_bf16_skinny(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act, stream);
    if (st != kUnsupported) return st;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.gamma = gamma; p.act = act; p.has_res = residual != nullptr ? 1 : 0; p.act_post = act_post;
  // force_block_n: 0 = choose; 64/128/256 = one-CTA kernel with that tile width; 2 = CTA-pair 256x256 kernel
  const bool pair = force_block_n == 2 || (force_block_n == 0 && prefer_pair(M, N));
  int bn = pair ? 256 : (force_block_n > 0 ? force_block_n : pick_block_n(M, N));
  if (pair) return gemm_bf16_pair(A, lda, W, ldw, residual, ldr, C, ldc, p, out_dtype, stream);
#define TFIMM_GEMM_CASE(BN)                                                                         \
  case BN:                                                                                          \
    return out_dtype == kBF16 ? launch_gemm<BN, __nv_bfloat16>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream) \
                              : launch_gemm<BN, float>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
  switch (bn) {
    TFIMM_GEMM_CASE(256)
    TFIMM_GEMM_CASE(128)
    TFIMM_GEMM_CASE(64)
    default:
      set_last_error("gemm: unsupported block_n %d", bn);
      return kInvalidArgument;
  }
#undef TFIMM_GEMM_CASE
}

// Dense layer whose input rows are first multiplied by a per-image channel gate (squeeze-excite): the projection
// convolutions after SEModule (tfimm/layers/attention.py, efficientnet_blocks.py:241-248, 438-453).  bf16 out.
int gemm_bf16_gated_dispatch(const void* A, int lda, const float* gate, int rows_per_img, int imgs, const void* W, int ldw,
                             const float* bias, const void* residual, int ldr, void* C, int ldc, int M, int N, int K,
                             int act, cudaStream_t stream) {
  TFIMM_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm_gated: need K %% 8 == 0 (got M=%d N=%d K=%d)", M, N, K);
  TFIMM_CHECK_ARG(gate != nullptr && rows_per_img > 0 && imgs > 0 && (reinterpret_cast<uintptr_t>(gate) & 15u) == 0,
                  "gemm_gated: gate [imgs][K] fp32, 16-byte aligned");
  TFIMM_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0, "gemm_gated: bias must be 16-byte aligned");
  if (K <= 64) {   // short contraction: the streaming kernel scales its A fragments in registers
    const int st = gemm_bf16_skinny(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act, stream, gate, rows_per_img,
                                    imgs);
    if (st != kUnsupported) return st;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.act = act; p.has_res = residual != nullptr ? 1 : 0;
  p.a_scale = gate; p.a_rows_per_img = rows_per_img; p.a_imgs = imgs;
  switch (pick_block_n(M, N)) {
    case 256: return launch_gemm<256, __nv_bfloat16, true>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
    case 128: return launch_gemm<128, __nv_bfloat16, true>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
    default: return launch_gemm<64, __nv_bfloat16, true>(A, lda, W, ldw, residual, ldr, C, ldc, p, stream);
  }
}

// k x k convolution (stride 1 or 2, symmetric padding (k-1)/2... given as `pad`) + bias + activation (+ residual),
// NHWC bf16 in, NHWC bf16/fp32 out, W[N][k*k*C] in (ky, kx, c) order: implicit GEMM, no im2col matrix in HBM.
int conv_bf16_dispatch(const void* x, const void* W, int ldw, const float* bias, const void* residual, void* out,
                       int B, int H, int Wd, int C, int N, int ks, int stride, int pad, int act, int act_post,
                       int out_dtype, cudaStream_t stream) {
  TFIMM_CHECK_ARG(B > 0 && H > 0 && Wd > 0 && C > 0 && C % 64 == 0, "conv: C must be a multiple of 64 (got %d)", C);
  TFIMM_CHECK_ARG(ks >= 1 && ks <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < ks, "conv: bad geometry");
  TFIMM_CHECK_ARG(N > 0 && N % 8 == 0, "conv: N must be a multiple of 8 (got %d)", N);
  TFIMM_CHECK_ARG(out_dtype == kBF16 || out_dtype == kF32, "conv: out_dtype must be bf16 or f32");
  TFIMM_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0, "conv: bias must be 16-byte aligned");
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (Wd + 2 * pad - ks) / stride + 1;
  TFIMM_CHECK_ARG(Ho > 0 && Wo > 0, "conv: empty output");
  GemmParams p{};
  p.N = N; p.K = ks * ks * C;
  p.bias = bias; p.act = act; p.has_res = residual != nullptr ? 1 : 0; p.act_post = act_post;
  p.conv = 1; p.cv_cblocks = C / 64; p.cv_ks = ks; p.cv_stride = stride; p.cv_pad = pad;
  // 128-pixel output patch: 8 x 16 pixels of one image, or 8 x 8 pixels of two images for small feature maps
  if (Wo > 8) { p.cv_pb = 1; p.cv_ph = 8; p.cv_pw = 16; }
  else { p.cv_pb = 2; p.cv_ph = 8; p.cv_pw = 8; }
  p.cv_tiles_x = (Wo + p.cv_pw - 1) / p.cv_pw;
  p.cv_tiles_y = (Ho + p.cv_ph - 1) / p.cv_ph;
  const int tiles_b = (B + p.cv_pb - 1) / p.cv_pb;
  p.M = tiles_b * p.cv_tiles_y * p.cv_tiles_x * kBlockM;  // padded row count: every tile is a full patch
  const int bn = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
#define TFIMM_CONV_CASE(BN)                                                                                       \
  case BN:                                                                                                        \
    return out_dtype == kBF16                                                                                     \
               ? launch_conv<BN, __nv_bfloat16>(x, W, ldw, residual, out, B, H, Wd, C, Ho, Wo, p, stream)           \
               : launch_conv<BN, float>(x, W, ldw, residual, out, B, H, Wd, C, Ho, Wo, p, stream);
  switch (bn) {
    TFIMM_CONV_CASE(256)
    TFIMM_CONV_CASE(128)
    TFIMM_CONV_CASE(64)
  }
#undef TFIMM_CONV_CASE
  return kInvalidArgument;
}

}  // namespace tfimm
