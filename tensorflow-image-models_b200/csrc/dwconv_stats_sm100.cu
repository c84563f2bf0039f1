// ConvNeXt block head with the LayerNorm moved OUT of the kernel:
//     ZeroPadding2D(3) -> DepthwiseConv2D(7x7, bias)      [-> LayerNorm -> Dense(C, 4C): folded into the GEMM]
// (tfimm/architectures/convnext.py:189-198, 219-224).  The fused dwconv+LN kernel (dwconv_ln_sm100.cu) spends ~2/3 of
// its instructions and all of its cluster barriers on the LayerNorm (statistics across the channel slabs of a
// cluster, normalisation pass); here the kernel only produces what the LayerNorm-folded GEMM needs
// (tfimm_b200_gemm_bf16_ln):
//   * the RAW convolution output in bf16 (the GEMM's A operand), and
//   * per pixel and per 64-channel slab the partial (sum, sum of squares) of the fp32 results.
// CTAs are independent (no clusters): one CTA = 14 x 7 output pixels x CS channels, persistent over the tile list;
// the fp32 input halo is ONE 4-D TMA box (zero padding = out-of-bounds fill), the taps of a lane's channel pair
// live in registers as packed fp32x2, a thread computes a 2-row x 7-column block (686 FFMA2 per 104 LDS.64), writes
// its bf16 pairs straight from registers (128 B per warp store) and the slab statistics come from warp shuffles.
// The next tile's halo is fetched while the stores / reductions of the current one run; two CTAs per SM.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kTH = 14;              // output tile rows (7 row pairs)
constexpr int kTW = 7;               // output tile columns
constexpr int kHH = kTH + 6;         // 20 halo rows
constexpr int kHW = kTW + 6;         // 13 halo columns
constexpr int kWarps = kTH / 2;      // one warp per output row pair
constexpr int kThreads = kWarps * 32;

template <int CS>
__global__ void __launch_bounds__(kThreads, 2)
dwconv7_stats_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ wgt /*[49][C]*/,
                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, float2* __restrict__ stats,
                     int H, int W, int C, int tiles_x, int tiles_per_img, long n_units, int cslabs) {
  constexpr int kPairs = CS / 2;
  constexpr int kHaloBytes = kHH * kHW * CS * 4;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint64_t* halo = reinterpret_cast<const uint64_t*>(smem);  // [20*13][CS/2] fp32x2
  const uint32_t bar = smem_u32(smem + kHaloBytes);
  const uint32_t halo_addr = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    prefetch_tmap(&tmap_x);
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto decode = [&](long unit, int& cs, int& b, int& ty0, int& tx0) {
    cs = (int)(unit % cslabs);  // channel slab fastest: neighbouring CTAs touch the same pixels
    const long t = unit / cslabs;
    b = (int)(t / tiles_per_img);
    const int ti = (int)(t % tiles_per_img);
    ty0 = (ti / tiles_x) * kTH;
    tx0 = (ti % tiles_x) * kTW;
  };
  auto issue_halo = [&](long unit) {
    int cs, b, ty0, tx0;
    decode(unit, cs, b, ty0, tx0);
    mbar_expect_tx(bar, kHaloBytes);
    tma_load_4d(halo_addr, &tmap_x, bar, cs * CS, tx0 - 3, ty0 - 3, b);
  };
  long unit = blockIdx.x;
  if (tid == 0 && unit < n_units) issue_halo(unit);

  const bool pair_on = lane < kPairs;
  int cur_cs = -1;
  uint64_t w[49];
  uint64_t bv = pack2(0.f, 0.f);
  uint32_t phase = 0;
  for (; unit < n_units; unit += gridDim.x) {
    int cs, b, ty0, tx0;
    decode(unit, cs, b, ty0, tx0);
    const int c0 = cs * CS + 2 * (pair_on ? lane : 0);
    if (cs != cur_cs) {  // taps of this lane's channel pair (a CTA keeps its slab when gridDim.x % cslabs == 0)
      cur_cs = cs;
#pragma unroll
      for (int t = 0; t < 49; ++t)
        w[t] = pack2(__ldg(wgt + (size_t)t * C + c0), __ldg(wgt + (size_t)t * C + c0 + 1));
      bv = pack2(__ldg(bias + c0), __ldg(bias + c0 + 1));
    }
    mbar_wait(bar, phase);
    phase ^= 1u;

    // ---- depthwise 7x7: warp = output row pair, lane = channel pair ----
    const int oy0 = 2 * warp;
    uint64_t acc0[7], acc1[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) acc0[i] = acc1[i] = bv;
    const bool rows_on = ty0 + oy0 < H;
    if (pair_on && rows_on) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint64_t* row = halo + (size_t)((oy0 + r) * kHW) * kPairs + lane;
#pragma unroll
        for (int ix = 0; ix < kHW; ++ix) {
          const uint64_t v = row[(size_t)ix * kPairs];
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
    }
    __syncthreads();
    // halo buffer is free: fetch this CTA's next tile while the stores / reductions below run
    if (tid == 0 && unit + gridDim.x < n_units) issue_halo(unit + gridDim.x);

    // ---- bf16 rows straight from registers + slab statistics by warp shuffles ----
    if (rows_on) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int oy = ty0 + oy0 + rr;
        if (oy >= H) break;
        const long prow = ((long)b * H + oy) * W;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int ox = tx0 + i;
          if (ox >= W) break;
          float a0, a1;
          unpack2(rr == 0 ? acc0[i] : acc1[i], a0, a1);
          if (!pair_on) a0 = a1 = 0.f;
          else *reinterpret_cast<uint32_t*>(out + (prow + ox) * C + c0) = pack_bf16x2(a0, a1);
          float s = a0 + a1, q = fmaf(a0, a0, a1 * a1);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            q += __shfl_xor_sync(0xffffffffu, q, off);
          }
          if (lane == 0) stats[(prow + ox) * cslabs + cs] = make_float2(s, q);
        }
      }
    }
  }
}

template <int CS>
int launch_dwstats(const void* x, const float* wgt, const float* bias, void* out, float* stats, int B, int H, int W,
                   int C, cudaStream_t stream) {
  constexpr int kSmemBytes = kHH * kHW * CS * 4 + 16;
  auto kernel = dwconv7_stats_kernel<CS>;
  static int resident = 0;
  if (resident == 0) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    int per_sm = 0;
    TFIMM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, kSmemBytes));
    resident = (per_sm > 0 ? per_sm : 1) * sm_count();
  }
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  const uint32_t box[4] = {(uint32_t)CS, (uint32_t)kHW, (uint32_t)kHH, 1u};
  int rc = make_tmap(&tmap, x, kF32, 4, dims, strides, box, "dwconv7_stats input", /*swizzle_bytes=*/0);
  if (rc != kOk) return rc;
  const int cslabs = C / CS;
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const long n_units = (long)B * tiles_x * tiles_y * cslabs;
  // a multiple of the slab count, so that a CTA stays on one channel slab and loads its taps once
  long grid = n_units < resident ? n_units : (resident / cslabs) * (long)cslabs;
  if (grid <= 0) grid = n_units < cslabs ? n_units : cslabs;
  kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(tmap, wgt, bias, reinterpret_cast<__nv_bfloat16*>(out),
                                                          reinterpret_cast<float2*>(stats), H, W, C, tiles_x,
                                                          tiles_x * tiles_y, n_units, cslabs);
  TFIMM_LAUNCH_OK("dwconv7_stats_kernel");
  return kOk;
}

}  // namespace

// Number of per-pixel partial statistics (channel slabs) the kernel writes for C channels.
int dwconv7_stats_parts(int C) { return C % 64 == 0 ? C / 64 : (C % 32 == 0 ? C / 32 : 0); }

int dwconv7_stats(const void* x, const float* wgt, const float* bias, void* out, float* stats, int B, int H, int W,
                  int C, cudaStream_t stream) {
  TFIMM_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0, "dwconv7_stats: need C%%32==0 (C=%d)", C);
  TFIMM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7u) == 0,
                  "dwconv7_stats: unaligned pointer");
  if (C % 64 == 0) return launch_dwstats<64>(x, wgt, bias, out, stats, B, H, W, C, stream);
  return launch_dwstats<32>(x, wgt, bias, out, stats, B, H, W, C, stream);
}

}  // namespace tfimm
