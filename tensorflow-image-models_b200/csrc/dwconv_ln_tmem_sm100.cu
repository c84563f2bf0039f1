// ConvNeXt block head on sm_100a:  ZeroPadding2D(3) -> DepthwiseConv2D(7x7, bias) -> LayerNorm over C
// (tfimm/architectures/convnext.py:189-198, 219-223), cluster-free formulation with the fp32 results parked in
// TENSOR MEMORY.
//
// The depthwise 7x7 is 49 MACs per element: FP32-FMA bound on the CUDA cores, provided taps and input halo come from
// on-chip memory and the instruction stream is mostly FMAs.  LayerNorm needs every channel of a pixel.  Here one
// persistent CTA owns ALL channels of a TH x 7 pixel tile and walks over them in slabs of 64:
//
//   producer warp   per (tile, slab): ONE 4-D TMA box (64 ch x 13 x (TH+6) fp32 halo; the zero padding of the
//                   convolution is the TMA out-of-bounds fill) + ONE 2-D box with the slab's 49 x 64 taps, into a
//                   2/3-stage shared-memory ring (full / empty mbarriers)
//   compute warp w  owns output rows 2w, 2w+1 of the tile; lane = channel pair of the slab.  Taps of the pair in
//                   registers (49 x fp32x2, re-read from the stage per slab), 2-row x 7-column register block,
//                   686 FFMA2 per 104 LDS.64; the 28 fp32 results go to the warp's private columns of TENSOR MEMORY
//                   (tcgen05.st) -- 200 KB of exact fp32 intermediates for a 14 x 7 x 512 tile that no shared-memory
//                   budget could hold next to the halos.
//                   After the last slab the SAME warp reads its pixels back (tcgen05.ld: 2 x NSLAB values per lane
//                   and pixel), merges per-lane (mean, M2) over the warp with Chan's formula (shuffles), normalises
//                   and writes bf16 rows.  No thread-block cluster, no DSMEM, no CTA-wide barrier in the tile loop:
//                   warps only meet at the stage barriers, so one warp's LayerNorm phase overlaps the others' FMA
//                   phases.
// Replaces the thread-block-cluster kernel of round 1 (per-tile cluster barrier + serial DSMEM reads: 15 % of HBM
// peak, fp16 stash); numerics are now plain fp32 up to the bf16 output rounding.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kTW = 7;          // output tile columns
constexpr int kHW = kTW + 6;    // 13 halo columns
constexpr int kCS = 64;         // channels per slab
constexpr int kTapBytes = 49 * kCS * 4;

template <int NW>
struct TmCfg {
  static constexpr int kTH = 2 * NW;                       // output rows per tile
  static constexpr int kHH = kTH + 6;                      // halo rows
  static constexpr int kHaloBytes = kHH * kHW * kCS * 4;   // fp32
  static constexpr int kStageBytes = ((kHaloBytes + kTapBytes + 127) / 128) * 128;
  static constexpr int kStages = NW == 7 ? 2 : 3;
  static constexpr int kThreads = (NW + 1) * 32;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStages * 8 + 16 + 128;
};

__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }

// columns per pixel in TMEM: 2 * NSLAB values, padded to a power of two so that one tcgen05.ld shape fetches them
__host__ __device__ constexpr int pix_cols(int nslab) {
  return 2 * nslab <= 4 ? 4 : (2 * nslab <= 8 ? 8 : (2 * nslab <= 16 ? 16 : 32));
}

template <int NW, int NSLAB>
__global__ void __launch_bounds__(TmCfg<NW>::kThreads, 1)
dwconv7_ln_tmem_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                       __nv_bfloat16* __restrict__ out, int H, int W, int C, int tiles_x, int tiles_per_img,
                       int n_tiles, float eps) {
  using Cfg = TmCfg<NW>;
  constexpr int kStages = Cfg::kStages;
  constexpr int PC = pix_cols(NSLAB);       // TMEM columns per pixel
  constexpr int kWarpCols = 14 * PC;        // per compute warp (14 pixels: 2 rows x 7 columns)
  static_assert((NW + 3) / 4 * kWarpCols <= 512, "tile does not fit tensor memory");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t tmem_ptr_smem = bars + 8u * 2 * kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NW);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  if (warp == NW) {
    // ------------------------------------------ TMA producer ------------------------------------------
    if (lane == 0) {
      uint32_t k = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, t = tile % tiles_per_img;
        const int tx0 = (t % tiles_x) * kTW, ty0 = (t / tiles_x) * Cfg::kTH;
        for (int s = 0; s < NSLAB; ++s, ++k) {
          const int stage = (int)(k % kStages);
          const uint32_t ph = (k / kStages) & 1u;
          mbar_wait(empty_bar(stage), ph ^ 1u);
          const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
          mbar_expect_tx(full_bar(stage), Cfg::kHaloBytes + kTapBytes);
          tma_load_4d(dst, &tmap_x, full_bar(stage), s * kCS, tx0 - 3, ty0 - 3, b);
          tma_load_2d(dst + Cfg::kHaloBytes, &tmap_w, full_bar(stage), s * kCS, 0);
        }
      }
    }
  } else {
    // ------------------------------------------ compute warps ------------------------------------------
    const uint32_t t_warp = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * kWarpCols);
    const int oy0 = 2 * warp;  // first output row of this warp inside the tile
    const float inv_c = 1.0f / (float)C;
    uint32_t k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, t = tile % tiles_per_img;
      const int tx0 = (t % tiles_x) * kTW, ty0 = (t / tiles_x) * Cfg::kTH;
      const bool rows_on = ty0 + oy0 < H;
      // ---- depthwise 7x7, slab by slab ----
#pragma unroll 1
      for (int s = 0; s < NSLAB; ++s, ++k) {
        const int stage = (int)(k % kStages);
        const uint32_t ph = (k / kStages) & 1u;
        const uint64_t bv = *reinterpret_cast<const uint64_t*>(bias + s * kCS + 2 * lane);
        mbar_wait(full_bar(stage), ph);
        if (rows_on) {
          const uint64_t* halo = reinterpret_cast<const uint64_t*>(smem_gen + stage * Cfg::kStageBytes);
          const uint64_t* taps = halo + Cfg::kHaloBytes / 8;
          uint64_t w[49];
#pragma unroll
          for (int i = 0; i < 49; ++i) w[i] = taps[i * 32 + lane];
          uint64_t acc0[7], acc1[7];
#pragma unroll
          for (int i = 0; i < 7; ++i) acc0[i] = acc1[i] = bv;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint64_t* row = halo + (size_t)((oy0 + r) * kHW) * 32 + lane;
#pragma unroll
            for (int ix = 0; ix < kHW; ++ix) {
              const uint64_t v = row[ix * 32];
#pragma unroll
              for (int kx = 0; kx < 7; ++kx) {
                const int ox = ix - kx;
                if (ox >= 0 && ox < 7) {
                  if (r <= 6) acc0[ox] = fma2(v, w[r * 7 + kx], acc0[ox]);
                  if (r >= 1) acc1[ox] = fma2(v, w[(r - 1) * 7 + kx], acc1[ox]);
                }
              }
            }
          }
          // this warp is done reading the stage
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          // park the 28 fp32 results: pixel p = row * 7 + col, columns [p * PC + 2 s, +2)
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            float a0, a1;
            unpack2(acc0[i], a0, a1);
            tmem_st_x2(t_warp + (uint32_t)(i * PC + 2 * s), __float_as_uint(a0), __float_as_uint(a1));
            unpack2(acc1[i], a0, a1);
            tmem_st_x2(t_warp + (uint32_t)((7 + i) * PC + 2 * s), __float_as_uint(a0), __float_as_uint(a1));
          }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
        }
      }
      if (!rows_on) continue;
      tmem_st_wait();
      // ---- LayerNorm over the C channels of each of this warp's 14 pixels ----
      float g[2 * NSLAB], be[2 * NSLAB];
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const float2 gg = __ldg(reinterpret_cast<const float2*>(gamma + s * kCS + 2 * lane));
        const float2 bb = __ldg(reinterpret_cast<const float2*>(beta + s * kCS + 2 * lane));
        g[2 * s] = gg.x; g[2 * s + 1] = gg.y;
        be[2 * s] = bb.x; be[2 * s + 1] = bb.y;
      }
#pragma unroll 2
      for (int p = 0; p < 14; ++p) {
        const int oy = ty0 + oy0 + p / 7, ox = tx0 + p % 7;
        if (oy >= H || ox >= W) continue;  // warp-uniform
        uint32_t raw[PC];
        tmem_ld_cols<PC>(t_warp + (uint32_t)(p * PC), raw);
        tmem_ld_wait();
        float v[2 * NSLAB];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 2 * NSLAB; ++j) {
          v[j] = __uint_as_float(raw[j]);
          sum += v[j];
        }
        // per-lane (mean, M2) over 2*NSLAB values, merged pairwise over the warp (equal counts: Chan et al.)
        float mean = sum * (1.0f / (2 * NSLAB));
        float m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 2 * NSLAB; ++j) m2 = fmaf(v[j] - mean, v[j] - mean, m2);
        float half_cnt = (float)NSLAB;  // count of each side / 2
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, mean, off);
          const float o2 = __shfl_xor_sync(0xffffffffu, m2, off);
          const float d = om - mean;
          m2 = m2 + o2 + d * d * half_cnt;
          mean = 0.5f * (mean + om);
          half_cnt *= 2.0f;
        }
        const float rstd = rsqrtf(m2 * inv_c + eps);
        const float nmr = -mean * rstd;
        __nv_bfloat16* orow = out + (((long)b * H + oy) * W + ox) * C + 2 * lane;
#pragma unroll
        for (int s = 0; s < NSLAB; ++s) {
          const float y0 = fmaf(fmaf(v[2 * s], rstd, nmr), g[2 * s], be[2 * s]);
          const float y1 = fmaf(fmaf(v[2 * s + 1], rstd, nmr), g[2 * s + 1], be[2 * s + 1]);
          *reinterpret_cast<uint32_t*>(orow + s * kCS) = pack_bf16x2(y0, y1);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int NW, int NSLAB>
int launch_tmem(const void* x, const float* wgt, const float* bias, const float* gamma, const float* beta, void* out,
                int B, int H, int W, int C, float eps, cudaStream_t stream) {
  using Cfg = TmCfg<NW>;
  auto kernel = dwconv7_ln_tmem_kernel<NW, NSLAB>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs))
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + Cfg::kTH - 1) / Cfg::kTH;
  const long n_tiles = (long)B * tiles_x * tiles_y;
  CUtensorMap tx, tw;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t box[4] = {(uint32_t)kCS, (uint32_t)kHW, (uint32_t)Cfg::kHH, 1u};
    const int rc = make_tmap(&tx, x, kF32, 4, dims, strides, box, "dwconv7_ln input", /*swizzle_bytes=*/0);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, 49};
    const uint64_t strides[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {(uint32_t)kCS, 49u};
    const int rc = make_tmap(&tw, wgt, kF32, 2, dims, strides, box, "dwconv7_ln taps", /*swizzle_bytes=*/0);
    if (rc != kOk) return rc;
  }
  const long grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kernel<<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tx, tw, bias, gamma, beta,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), H, W, C,
                                                                    tiles_x, tiles_x * tiles_y, (int)n_tiles, eps);
  TFIMM_LAUNCH_OK("dwconv7_ln_tmem_kernel");
  return kOk;
}

}  // namespace

// Returns kUnsupported (without setting an error) when the shape is outside this kernel's instantiations, so the
// caller can use the generic kernel: fp32 in / bf16 out, C a multiple of 64 with C/64 in {2,3,4,6,8} (14-row tiles) or
// {12,16} and, for maps of at most 8 rows, {6,8} as well (8-row tiles, one compute warp per TMEM lane quarter).
int dwconv7_ln_tmem(const void* x, int in_dtype, const float* wgt, const float* bias, const float* gamma,
                    const float* beta, void* out, int out_dtype, int B, int H, int W, int C, float eps,
                    cudaStream_t stream) {
  if (out_dtype != kBF16 || in_dtype != kF32 || C % kCS != 0) return kUnsupported;
  if ((long)B * H * W * C >= (1L << 40)) return kUnsupported;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || (reinterpret_cast<uintptr_t>(wgt) & 15u) != 0 ||
      (reinterpret_cast<uintptr_t>(bias) & 7u) != 0 || (reinterpret_cast<uintptr_t>(gamma) & 7u) != 0 ||
      (reinterpret_cast<uintptr_t>(beta) & 7u) != 0 || (reinterpret_cast<uintptr_t>(out) & 3u) != 0)
    return kUnsupported;
  const int ns = C / kCS;
#define TFIMM_TM(NW, NS) return launch_tmem<NW, NS>(x, wgt, bias, gamma, beta, out, B, H, W, C, eps, stream)
  if (H > 8) {
    switch (ns) {
      case 2: TFIMM_TM(7, 2);
      case 3: TFIMM_TM(7, 3);
      case 4: TFIMM_TM(7, 4);
      case 6: TFIMM_TM(7, 6);
      case 8: TFIMM_TM(7, 8);
      case 12: TFIMM_TM(4, 12);
      case 16: TFIMM_TM(4, 16);
      default: return kUnsupported;
    }
  }
  switch (ns) {
    case 2: TFIMM_TM(4, 2);
    case 3: TFIMM_TM(4, 3);
    case 4: TFIMM_TM(4, 4);
    case 6: TFIMM_TM(4, 6);
    case 8: TFIMM_TM(4, 8);
    case 12: TFIMM_TM(4, 12);
    case 16: TFIMM_TM(4, 16);
    default: return kUnsupported;
  }
#undef TFIMM_TM
}

}  // namespace tfimm
