// tcgen05 self-attention for short sequences (N <= 256 keys, head_dim 64): the ViT-B/16 case
// (N = 197) of tfimm/architectures/vit.py:149-165.  Persistent, warp-specialised, one CTA per SM.
//   TMA (3D maps over the packed qkv projection, zero-fill past the sequence end)
//        Q [256 x 64], K [npad x 64], V [npad x 64] of work item i+1 -> 128B-swizzled smem while item i runs
//   S = Q K^T        one tcgen05.mma chain per 128-query tile (M=128, N=npad, K=64), fp32 scores in TMEM
//   softmax          each thread owns one query row: tcgen05.ld, max, exp2, row sum in registers;
//                    P (bf16) is written back over the consumed score columns (TMEM cols [0, npad/2))
//   O = P V          tcgen05.mma with A = P from TMEM and B = V from smem as an MN-major operand
//                    (V is [key][dh] in memory: no transpose pass), O in TMEM cols [128, 192)
//   epilogue         O / rowsum -> bf16 -> swizzled smem slab -> TMA store (clipped at the sequence end)
// The (B,H,N,N) score tensor the reference materialises never leaves the SM.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kDH = 64;
constexpr int kQRows = 128;
constexpr int kMaxKeys = 256;
constexpr uint32_t kOCol = 128;
constexpr int kKVBytesMax = kMaxKeys * 128;    // 32 KB each

// Row softmax of one 128-query tile, thread = query row, for a COMPILE-TIME key count (NFULL 32-column chunks plus
// an optional 16-column tail): scores S (fp32, TMEM columns [0, npad) of this warp's lane quarter at t_row) ->
// P = exp2((S - max) * scale_log2) as packed bf16 in columns [0, npad/2); returns the fp32 row sum of the unrounded
// P.  Two passes over TMEM (max, then exp), each fully unrolled with the load of chunk c+1 in flight while chunk c
// is processed in the OTHER register buffer (no copies); only the last chunk is masked against the true key count N
// (its padding columns hold 0 = q . 0, which must neither win the max nor enter the sum).
// Instruction budget per row of 208 keys: 14 LDTM + 104 FMNMX3 + 104 FFMA2 + 208 MUFU.EX2 + 104 FADD2 + 104 F2FP +
// 13 STTM: the MUFU pipe (16 lanes / clk / SM) is the floor.
template <int NFULL, bool TAIL16>
__device__ __forceinline__ float softmax_tile_unrolled(uint32_t t_row, int N, float scale_log2) {
  constexpr int NCH = NFULL + (TAIL16 ? 1 : 0);
  uint32_t buf[2][32];
  auto issue = [&](int c, uint32_t (&rg)[32]) {
    if (!TAIL16 || c < NFULL) {
      tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), rg);
    } else {
      uint32_t r16[16];
      tmem_ld_32x32b_x16(t_row + (uint32_t)(c * 32), r16);
#pragma unroll
      for (int j = 0; j < 16; ++j) rg[j] = r16[j];
    }
  };
  // ---- pass 1: row maximum ----
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  issue(0, buf[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tmem_ld_wait();
    if (c + 1 < NCH) issue(c + 1, buf[(c + 1) & 1]);
    const uint32_t (&v)[32] = buf[c & 1];
    const int width = (TAIL16 && c == NFULL) ? 16 : 32;
    if (c + 1 < NCH) {
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        mx[(j >> 1) & 3] = fmaxf(mx[(j >> 1) & 3], fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
    } else {
#pragma unroll
      for (int j = 0; j < width; ++j)
        mx[j & 3] = fmaxf(mx[j & 3], (c * 32 + j < N) ? __uint_as_float(v[j]) : -INFINITY);
    }
  }
  const float mx_all = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  const uint64_t sl2 = splat2(scale_log2), nmoff = splat2(-mx_all * scale_log2);
  // ---- pass 2: P = exp2(S * scale_log2 - max * scale_log2), row sum, bf16 P back into TMEM ----
  uint64_t sum[4] = {splat2(0.f), splat2(0.f), splat2(0.f), splat2(0.f)};
  issue(0, buf[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tmem_ld_wait();
    if (c + 1 < NCH) issue(c + 1, buf[(c + 1) & 1]);
    const uint32_t (&v)[32] = buf[c & 1];
    const int width = (TAIL16 && c == NFULL) ? 16 : 32;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < width / 2; ++j) {
      float t0, t1;
      unpack2(fma2(pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sl2, nmoff), t0, t1);
      float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
      if (c + 1 == NCH) {
        p0 = (c * 32 + 2 * j < N) ? p0 : 0.f;
        p1 = (c * 32 + 2 * j + 1 < N) ? p1 : 0.f;
      }
      sum[j & 3] = add2(sum[j & 3], pack2(p0, p1));
      pk[j] = pack_bf16x2(p0, p1);
    }
    // P chunk c overwrites score columns [16c, 16c + width/2): consumed already (the load in flight is chunk c+1)
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = pk[j]; hi[j] = (width == 32) ? pk[8 + j] : 0u; }
    tmem_st_32x32b_x8(t_row + (uint32_t)(c * 16), lo);
    if (width == 32) tmem_st_32x32b_x8(t_row + (uint32_t)(c * 16 + 8), hi);
  }
  tmem_st_wait();
  float s0, s1;
  unpack2(add2(add2(sum[0], sum[1]), add2(sum[2], sum[3])), s0, s1);
  return s0 + s1;
}

// ---------------------------------------------------------------------------------------------------
// persistent, warp-specialised, two-stage pipeline (one CTA per SM)
// ---------------------------------------------------------------------------------------------------
//   warp 0      TMA producer: Q (256 rows), K, V of work item i+1 while item i is being processed
//   warps 1, 10 tcgen05.mma issuers, one per group: S = Q K^T of the group's query tile, later O = P V
//   warps 2..5  softmax / epilogue group A (TMEM columns   0..255)
//   warps 6..9  softmax / epilogue group B (TMEM columns 256..511)
// A work item is one (image, head).  With two query tiles per item (197 tokens: 128 + 69 rows) the groups ALTERNATE
// between the long and the short tile from item to item (group g takes tile (item index + g) & 1) and are driven by
// separate MMA warps, so the group that had the short tile moves on to the next item (already in the other stage) while
// the other one is still busy: an item costs (long + short) / 2 per group instead of `long` (with one MMA warp and a
// fixed assignment group B idled ~40 % of every item).  Per tile the TMEM region holds S (fp32, cols [0,npad)), then
// P (bf16 packed, cols [0,npad/2)) written over the consumed scores, then O (cols [128,192)).
constexpr int kP2Threads = 352;
constexpr int kP2QBytes = 256 * 128;
constexpr int kP2StageBytes = kP2QBytes + 2 * kKVBytesMax;  // 96 KB
constexpr int kP2SmemBytes = 2 * kP2StageBytes + 256 + 1024;

__global__ void __launch_bounds__(kP2Threads, 1)
vit_attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                         const __grid_constant__ CUtensorMap tmap_o, int N, int H, int total_items,
                         float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + 2 * kP2StageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (2 + s); };
  auto sfull_bar = [&](int r) { return bars + 8u * (4 + r); };
  auto pready_bar = [&](int r) { return bars + 8u * (6 + r); };
  auto ofull_bar = [&](int r) { return bars + 8u * (8 + r); };
  auto tempty_bar = [&](int r) { return bars + 8u * (10 + r); };
  const uint32_t tmem_ptr_smem = bars + 8u * 12;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * kDH;
  const int npad = (N + 15) & ~15;
  const int mtiles = N > kQRows ? 2 : 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    prefetch_tmap(&tmap_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 4 * mtiles);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(sfull_bar(r), 1);
      mbar_init(pready_bar(r), 4);
      mbar_init(ofull_bar(r), 1);
      mbar_init(tempty_bar(r), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  if (warp == 0) {
    // ------------------------------------- TMA producer -------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        const int b = item / H, h = item % H;
        const uint32_t sQ = smem_base + s * kP2StageBytes, sK = sQ + kP2QBytes, sV = sK + kKVBytesMax;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)(kP2QBytes + 2 * npad * 128));
        tma_load_3d(sQ, &tmap_q, full_bar(s), h * kDH, 0, b);
        tma_load_3d(sK, &tmap_kv, full_bar(s), D + h * kDH, 0, b);
        tma_load_3d(sV, &tmap_kv, full_bar(s), 2 * D + h * kDH, 0, b);
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------ MMA issuer of group g (one warp each) ------------------------------
    const int g = warp == 1 ? 0 : 1;
    if (g < mtiles) {
      const uint32_t idesc_s = umma_idesc_bf16_f32(kQRows, npad);
      const uint32_t idesc_o = umma_idesc_bf16_f32(kQRows, kDH, /*b_mn_major=*/true);
      const int ksteps = npad >> 4;
      const uint32_t t0 = tmem_base + (uint32_t)(g * 256);
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u, par = (uint32_t)it & 1u;
        const int r = mtiles == 2 ? ((it + g) & 1) : 0;   // query tile of this group for this item
        const uint32_t sQ = smem_base + s * kP2StageBytes, sK = sQ + kP2QBytes, sV = sK + kKVBytesMax;
        mbar_wait(full_bar(s), ph);
        mbar_wait(tempty_bar(g), par ^ 1u);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint64_t dq = umma_desc_k_sw128(sQ + (uint32_t)r * (kQRows * 128)), dk = umma_desc_k_sw128(sK);
#pragma unroll
          for (int k = 0; k < kDH / 16; ++k)
            umma_bf16_ss(t0, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
          umma_commit(sfull_bar(g));
        }
        __syncwarp();
        mbar_wait(pready_bar(g), par);
        tcgen05_fence_after();
        if (lane == 0) {
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t dv = umma_desc_mn_sw128(sV + (uint32_t)(j * 2048), (uint32_t)(npad * 128));
            umma_bf16_ts(t0 + kOCol, t0 + (uint32_t)(j * 8), dv, idesc_o, (uint32_t)(j != 0));
          }
          umma_commit(ofull_bar(g));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- softmax + epilogue groups -------------------------------
    const int g = (warp - 2) >> 2;  // group: TMEM region g, barriers g
    const int q = warp & 3;         // TMEM lane quarter
    if (g < mtiles) {
      const uint32_t t_row = tmem_base + (uint32_t)(g * 256) + ((uint32_t)(q * 32) << 16);
      const int nfull = npad >> 5;
      const bool tail16 = (npad & 16) != 0;
      const int nchunks = nfull + (tail16 ? 1 : 0);
      const uint64_t sl2 = splat2(scale_log2);
      auto issue = [&](int c, uint32_t (&rg)[32]) {
        if (c < nfull) {
          tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), rg);
        } else {
          uint32_t t16[16];
          tmem_ld_32x32b_x16(t_row + (uint32_t)(c * 32), t16);
#pragma unroll
          for (int j = 0; j < 16; ++j) rg[j] = t16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) rg[j] = 0xff800000u;
        }
      };
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t par = (uint32_t)it & 1u;
        const int b = item / H, h = item % H;
        const int r = mtiles == 2 ? ((it + g) & 1) : 0;   // this item's query tile for this group (alternates)
        const int row0 = r * kQRows + q * 32;             // first query row of this warp inside the image
        const bool warp_has_rows = row0 < N;
        mbar_wait(sfull_bar(g), par);
        tcgen05_fence_after();
        float row_sum = 1.f;
        if (warp_has_rows && npad == 208) {
          row_sum = softmax_tile_unrolled<6, true>(t_row, N, scale_log2);   // ViT-B/16, DeiT @224: 197 / 198 tokens
        } else if (warp_has_rows && npad == 64) {
          row_sum = softmax_tile_unrolled<2, false>(t_row, N, scale_log2);  // patch-32 models @224: 50 tokens
        } else if (warp_has_rows) {
          // any other key count: run-time chunk loop
          // four independent running maxima / sums: one accumulator would be a 208-deep dependent chain per row
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          {
            uint32_t rg[32], cur[32];
            issue(0, rg);
            for (int c = 0; c < nchunks; ++c) {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) cur[j] = rg[j];
              if (c + 1 < nchunks) issue(c + 1, rg);
              if (c * 32 + 32 <= N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c * 32 + j < N) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
              }
            }
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          const uint64_t nmoff = splat2(-mx * scale_log2);
          uint64_t sum4[4] = {splat2(0.f), splat2(0.f), splat2(0.f), splat2(0.f)};
          {
            uint32_t rg[32], cur[32];
            issue(0, rg);
            for (int c = 0; c < nchunks; ++c) {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) cur[j] = rg[j];
              if (c + 1 < nchunks) issue(c + 1, rg);
              uint32_t pk[16];
              const bool full = (c * 32 + 32 <= N);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float t0, t1;
                unpack2(fma2(pack2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), sl2, nmoff), t0, t1);
                float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
                if (!full) {
                  p0 = (c * 32 + 2 * j < N) ? p0 : 0.f;
                  p1 = (c * 32 + 2 * j + 1 < N) ? p1 : 0.f;
                }
                sum4[j & 3] = add2(sum4[j & 3], pack2(p0, p1));
                pk[j] = pack_bf16x2(p0, p1);
              }
              uint32_t lo[8], hi[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { lo[j] = pk[j]; hi[j] = pk[8 + j]; }
              tmem_st_32x32b_x8(t_row + (uint32_t)(c * 16), lo);
              if (c < nfull) tmem_st_32x32b_x8(t_row + (uint32_t)(c * 16 + 8), hi);
            }
          }
          tmem_st_wait();
          float s0, s1;
          unpack2(add2(add2(sum4[0], sum4[1]), add2(sum4[2], sum4[3])), s0, s1);
          row_sum = s0 + s1;
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pready_bar(g));

        mbar_wait(ofull_bar(g), par);
        tcgen05_fence_after();
        uint32_t o0[32], o1[32];
        if (warp_has_rows) {
          tmem_ld_32x32b_x32(t_row + kOCol, o0);
          tmem_ld_32x32b_x32(t_row + kOCol + 32u, o1);
          tmem_ld_wait();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(g));  // the group's MMA warp may start the next item's S in this region
        if (warp_has_rows) {
          const float inv = 1.0f / row_sum;
          const uint32_t slab = smem_base + s * kP2StageBytes + (uint32_t)row0 * 128u;  // this warp's dead Q rows
          uint8_t* my_row = smem_gen + (slab - smem_base) + lane * 128;
          const int sw = lane & 7;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t (&o)[32] = j < 4 ? o0 : o1;
            const int e = 8 * (j & 3);
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[e + 0]) * inv, __uint_as_float(o[e + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[e + 2]) * inv, __uint_as_float(o[e + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[e + 4]) * inv, __uint_as_float(o[e + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[e + 6]) * inv, __uint_as_float(o[e + 7]) * inv);
            *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = u;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmap_o, slab, h * kDH, row0, b);
            tma_store_commit();
            tma_store_wait_read<0>();
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));  // K/V/Q of this stage may be overwritten
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int attention_bf16_tc2(const void* qkv, void* out, int B, int N, int H, float scale, cudaStream_t stream) {
  const int D = H * kDH;
  const int npad = (N + 15) & ~15;
  CUtensorMap tq, tkv, to;
  {
    const uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)N, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)N * 3 * D * 2};
    const uint32_t box_q[3] = {kDH, 256, 1};
    const uint32_t box_kv[3] = {kDH, (uint32_t)npad, 1};
    int st;
    if ((st = make_tmap(&tq, qkv, kBF16, 3, dims, strides, box_q, "attention q")) != kOk) return st;
    if ((st = make_tmap(&tkv, qkv, kBF16, 3, dims, strides, box_kv, "attention kv")) != kOk) return st;
  }
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)N, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)N * D * 2};
    const uint32_t box[3] = {kDH, 32, 1};
    int st;
    if ((st = make_tmap(&to, out, kBF16, 3, dims, strides, box, "attention out")) != kOk) return st;
  }
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs)) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(vit_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2SmemBytes));
  }
  const int items = B * H;
  const int grid = items < sm_count() ? items : sm_count();
  vit_attention_tc2_kernel<<<grid, kP2Threads, kP2SmemBytes, stream>>>(tq, tkv, to, N, H, items,
                                                                       scale * 1.4426950408889634f);
  TFIMM_LAUNCH_OK("vit_attention_tc2_kernel");
  return kOk;
}

}  // namespace tfimm
