// ConvNeXt block head on sm_100a:  ZeroPadding2D(3) -> DepthwiseConv2D(7x7, bias) -> LayerNorm over C
// (tfimm/architectures/convnext.py:189-198, 219-223), channel-slab / thread-block-cluster formulation.
//
// A depthwise 7x7 is 49 MACs per element: FP32-FMA bound on the CUDA cores, not HBM bound, *if* the taps
// and the input halo are read from on-chip memory.  LayerNorm then needs every channel of a pixel.  Both are
// reconciled by splitting the channels of one 14x7 pixel tile over the CTAs of a cluster:
//
//   CTA (cluster rank r)  owns channels [r*CS, (r+1)*CS)  (CS = 96 or 128) of a 14(rows) x 7(cols) output tile
//     A. stage the 20x13 input halo of its channel slab in shared memory as bf16 (zero outside the image)
//     B. each thread keeps the 49 taps of ONE channel pair in registers (packed fp32x2) and slides over its
//        share of the tile's 7-pixel row strips: 343 FFMA2 per strip, operands from smem; results (+bias) go
//        to an fp16 stash [98][CS] in shared memory (11-bit mantissa: 8x finer than the bf16 output)
//     C. LayerNorm statistics: per-pixel partial sums over the slab, exchanged between the CTAs of the
//        cluster through distributed shared memory (two rounds: mean, then centred second moment)
//     D. normalise the own slab and write bf16 rows
// ~93 KB of shared memory per CTA, so two CTAs are resident per SM and one CTA's global loads (phase A)
// overlap the other's FMA phase.  Weights are read once per CTA instead of once per pixel strip.
#include "common.cuh"

#include <cooperative_groups.h>
#include <cuda_fp16.h>

namespace cg = cooperative_groups;

namespace tfimm {
namespace {

constexpr int kTH = 14;              // output tile rows
constexpr int kTW = 7;               // output tile columns (one 7-pixel strip per row)
constexpr int kHH = kTH + 6;         // 20 halo rows
constexpr int kHW = kTW + 6;         // 13 halo columns
constexpr int kPix = kTH * kTW;      // 98
constexpr int kThreads = 256;

template <int CS>
struct DwCfg {
  static constexpr int kPairs = CS / 2;                 // channel pairs per slab
  static constexpr int kSlots = 64;                     // thread slots per row group (>= kPairs)
  static constexpr int kHaloBytes = ((kHH * kHW * CS * 2 + 15) / 16) * 16;
  static constexpr int kStashBytes = kPix * CS * 2;     // fp16
  static constexpr int kStatBytes = kPix * 4 * 4;       // part_sum, part_sq, mean, rstd
  static constexpr int kSmemBytes = kHaloBytes + kStashBytes + kStatBytes;
};

__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t u) {
  // bf16 -> fp32 is a 16-bit shift: low half -> lane 0, high half -> lane 1
  return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <typename InT, int CS>
__global__ void __launch_bounds__(kThreads, 2)
dwconv7_ln_cluster_kernel(const InT* __restrict__ x, const float* __restrict__ wgt /*[49][C]*/,
                          const float* __restrict__ bias, const float* __restrict__ gamma,
                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int H, int W, int C,
                          int tiles_x, int tiles_per_img, int cluster_size, float eps) {
  using Cfg = DwCfg<CS>;
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* halo = reinterpret_cast<uint32_t*>(smem);                          // [20*13][CS/2] bf16x2
  __half2* stash = reinterpret_cast<__half2*>(smem + Cfg::kHaloBytes);         // [98][CS/2] fp16x2
  float* part_sum = reinterpret_cast<float*>(smem + Cfg::kHaloBytes + Cfg::kStashBytes);
  float* part_sq = part_sum + kPix;
  float* s_mean = part_sq + kPix;
  float* s_rstd = s_mean + kPix;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int crank = blockIdx.x % cluster_size;
  const int tile_id = blockIdx.x / cluster_size;
  const int b = tile_id / tiles_per_img;
  const int t_in_img = tile_id % tiles_per_img;
  const int ty0 = (t_in_img / tiles_x) * kTH, tx0 = (t_in_img % tiles_x) * kTW;
  const int c_base = crank * CS;

  // ---- A. halo tile -> smem (bf16), 4 channels per lane; loads issued in batches before any is consumed ----
  {
    constexpr int kQuads = CS / 4;
    constexpr int kTotal = kHH * kHW * kQuads;
    constexpr int kU = 8;
    for (int base = 0; base < kTotal; base += kThreads * kU) {
      uint2 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int idx = base + u * kThreads + tid;
        v[u] = make_uint2(0u, 0u);
        if (idx < kTotal) {
          const int qd = idx % kQuads;
          const int pos = idx / kQuads;
          const int gy = ty0 + pos / kHW - 3, gx = tx0 + pos % kHW - 3;
          if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const InT* src = x + (((long)b * H + gy) * W + gx) * C + c_base + qd * 4;
            if constexpr (sizeof(InT) == 4) {
              const float4 f = *reinterpret_cast<const float4*>(src);
              v[u].x = pack_bf16x2(f.x, f.y);
              v[u].y = pack_bf16x2(f.z, f.w);
            } else {
              v[u] = *reinterpret_cast<const uint2*>(src);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int idx = base + u * kThreads + tid;
        if (idx < kTotal) *reinterpret_cast<uint2*>(halo + (size_t)(idx / kQuads) * (CS / 2) + (idx % kQuads) * 2) = v[u];
      }
    }
  }
  __syncthreads();

  // ---- B. depthwise 7x7: thread = (channel pair cp, row group rg); one 7-pixel strip per tile row ----
  {
    const int cp = tid % Cfg::kSlots, rg = tid / Cfg::kSlots;
    if (cp < Cfg::kPairs) {
      const int c0 = c_base + 2 * cp;
      uint64_t w[49];
#pragma unroll
      for (int t = 0; t < 49; ++t) w[t] = pack2(__ldg(wgt + (size_t)t * C + c0), __ldg(wgt + (size_t)t * C + c0 + 1));
      const uint64_t bv = pack2(__ldg(bias + c0), __ldg(bias + c0 + 1));
#pragma unroll 1
      for (int oy = rg; oy < kTH; oy += kThreads / Cfg::kSlots) {
        if (ty0 + oy >= H) break;
        uint64_t acc[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) acc[i] = bv;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t* row = halo + (size_t)((oy + ky) * kHW) * (CS / 2) + cp;
#pragma unroll
          for (int ix = 0; ix < 13; ++ix) {
            const uint64_t v = bf16x2_to_f32x2(row[(size_t)ix * (CS / 2)]);
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
              const int ox = ix - kx;
              if (ox >= 0 && ox < 7) acc[ox] = fma2(v, w[ky * 7 + kx], acc[ox]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          float a0, a1;
          unpack2(acc[i], a0, a1);
          stash[(size_t)(oy * kTW + i) * (CS / 2) + cp] = __floats2half2_rn(a0, a1);
        }
      }
    }
  }
  __syncthreads();

  // ---- C. LayerNorm statistics over all C channels of each pixel (cluster-wide) ----
  cg::cluster_group cluster = cg::this_cluster();
  const float inv_c = 1.0f / (float)C;
  auto pixel_valid = [&](int p) { return (ty0 + p / kTW) < H && (tx0 + p % kTW) < W; };
  // lane l reads channel pairs 2l, 2l+1 (one 8-byte access) of a pixel's slab row
  auto load4 = [&](int p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(stash + (size_t)p * (CS / 2) + 2 * lane);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = c2.x; v[3] = c2.y;
  };
  const bool lane_on = lane * 4 < CS;
  for (int p = warp; p < kPix; p += kThreads / 32) {
    float s = 0.f;
    if (lane_on && pixel_valid(p)) {
      float v[4];
      load4(p, v);
      s = (v[0] + v[1]) + (v[2] + v[3]);
    }
    s = warp_sum(s);
    if (lane == 0) part_sum[p] = s;
  }
  if (cluster_size > 1) cluster.sync(); else __syncthreads();
  for (int p = tid; p < kPix; p += kThreads) {
    float s = 0.f;
    for (int r = 0; r < cluster_size; ++r)
      s += (cluster_size > 1 ? cluster.map_shared_rank(part_sum, r) : part_sum)[p];
    s_mean[p] = s * inv_c;
  }
  __syncthreads();
  for (int p = warp; p < kPix; p += kThreads / 32) {
    float s = 0.f;
    if (lane_on && pixel_valid(p)) {
      const float m = s_mean[p];
      float v[4];
      load4(p, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (v[j] - m) * (v[j] - m);
    }
    s = warp_sum(s);
    if (lane == 0) part_sq[p] = s;
  }
  if (cluster_size > 1) cluster.sync(); else __syncthreads();
  for (int p = tid; p < kPix; p += kThreads) {
    float s = 0.f;
    for (int r = 0; r < cluster_size; ++r)
      s += (cluster_size > 1 ? cluster.map_shared_rank(part_sq, r) : part_sq)[p];
    s_rstd[p] = rsqrtf(s * inv_c + eps);
  }
  __syncthreads();

  // ---- D. normalise the own slab and write (8 bytes per lane, 256 B per pixel row for CS = 128) ----
  if (lane_on) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c_base + lane * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c_base + lane * 4));
    for (int p = warp; p < kPix; p += kThreads / 32) {
      if (!pixel_valid(p)) continue;
      const float m = s_mean[p], rs = s_rstd[p];
      float v[4];
      load4(p, v);
      uint2 u;
      u.x = pack_bf16x2((v[0] - m) * rs * g.x + be.x, (v[1] - m) * rs * g.y + be.y);
      u.y = pack_bf16x2((v[2] - m) * rs * g.z + be.z, (v[3] - m) * rs * g.w + be.w);
      __nv_bfloat16* orow = out + (((long)b * H + ty0 + p / kTW) * W + tx0 + p % kTW) * C + c_base + lane * 4;
      *reinterpret_cast<uint2*>(orow) = u;
    }
  }
  // peers may still be reading this CTA's partial sums
  if (cluster_size > 1) cluster.sync();
}

template <typename InT, int CS>
int launch_cluster(const void* x, const float* wgt, const float* bias, const float* gamma, const float* beta,
                   void* out, int B, int H, int W, int C, float eps, cudaStream_t stream) {
  using Cfg = DwCfg<CS>;
  auto kernel = dwconv7_ln_cluster_kernel<InT, CS>;
  const int cl = C / CS;
  static bool attr_set = false;
  if (!attr_set) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((long)B * tiles_x * tiles_y * cl));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TFIMM_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, reinterpret_cast<const InT*>(x), wgt, bias, gamma, beta,
                                   reinterpret_cast<__nv_bfloat16*>(out), H, W, C, tiles_x, tiles_x * tiles_y, cl, eps));
  return kOk;
}

}  // namespace

// Returns kUnsupported (without setting an error) when the shape does not fit the cluster formulation, so the
// caller can use the generic kernel.
int dwconv7_ln_cluster(const void* x, int in_dtype, const float* wgt, const float* bias, const float* gamma,
                       const float* beta, void* out, int out_dtype, int B, int H, int W, int C, float eps,
                       cudaStream_t stream) {
  if (out_dtype != kBF16) return kUnsupported;
  int cs = 0;
  for (int cand : {128, 96}) {
    if (C % cand == 0) {
      const int cl = C / cand;
      if (cl == 1 || cl == 2 || cl == 4 || cl == 8) { cs = cand; break; }
    }
  }
  if (cs == 0) return kUnsupported;
#define TFIMM_DWC(IN)                                                                                 \
  return cs == 128 ? launch_cluster<IN, 128>(x, wgt, bias, gamma, beta, out, B, H, W, C, eps, stream)  \
                   : launch_cluster<IN, 96>(x, wgt, bias, gamma, beta, out, B, H, W, C, eps, stream)
  if (in_dtype == kF32) { TFIMM_DWC(float); }
  if (in_dtype == kBF16) { TFIMM_DWC(__nv_bfloat16); }
#undef TFIMM_DWC
  return kUnsupported;
}

}  // namespace tfimm
