// ConvNeXt block head on sm_100a:  ZeroPadding2D(3) -> DepthwiseConv2D(7x7, bias) -> LayerNorm over C
// (tfimm/architectures/convnext.py:189-198, 219-223), channel-slab / thread-block-cluster formulation.
//
// A depthwise 7x7 is 49 MACs per element: FP32-FMA bound on the CUDA cores, not HBM bound, *if* the taps
// and the input halo are read from on-chip memory.  LayerNorm then needs every channel of a pixel.  Both are
// reconciled by splitting the channels of one 14x14 pixel tile over the CTAs of a cluster:
//
//   CTA (cluster rank r)  owns channels [r*CS, (r+1)*CS)  (CS = 96 or 128) of a 14x14 output tile
//     A. stage the 20x20 input halo of its channel slab in shared memory as bf16 (zero outside the image)
//     B. each thread keeps the 49 taps of ONE channel pair in registers (packed fp32x2) and slides over its
//        share of the tile's 7-pixel strips: 343 FFMA2 per strip, operands from smem; results (+bias) go to
//        an fp32 stash [196][CS] in shared memory
//     C. LayerNorm statistics: per-pixel partial sums over the slab, exchanged between the CTAs of the
//        cluster through distributed shared memory (two rounds: mean, then centred second moment)
//     D. normalise the own slab and write bf16 (or fp32) rows
// Weights are read once per CTA instead of once per pixel strip, the input once per tile (+halo).
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace tfimm {
namespace {

constexpr int kT = 14;             // output tile edge
constexpr int kHalo = kT + 6;      // 20
constexpr int kPix = kT * kT;      // 196
constexpr int kThreads = 256;

template <int CS>
struct DwCfg {
  static constexpr int kPairs = CS / 2;                 // channel pairs per slab
  static constexpr int kSlots = 64;                     // thread slots per row group (>= kPairs)
  static constexpr int kHaloBytes = kHalo * kHalo * CS * 2;
  static constexpr int kStashBytes = kPix * CS * 4;
  static constexpr int kStatBytes = kPix * 4 * 4;       // part_sum, part_sq, mean, rstd
  static constexpr int kSmemBytes = kHaloBytes + kStashBytes + kStatBytes;
};

__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t u) {
  // bf16 -> fp32 is a 16-bit shift: low half -> lane 0, high half -> lane 1
  return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <typename InT, typename OutT, int CS>
__global__ void __launch_bounds__(kThreads, 1)
dwconv7_ln_cluster_kernel(const InT* __restrict__ x, const float* __restrict__ wgt /*[49][C]*/,
                          const float* __restrict__ bias, const float* __restrict__ gamma,
                          const float* __restrict__ beta, OutT* __restrict__ out, int H, int W, int C,
                          int tiles_x, int tiles_per_img, int cluster_size, float eps) {
  using Cfg = DwCfg<CS>;
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* halo = reinterpret_cast<uint32_t*>(smem);                          // [400][CS/2] bf16x2
  float* stash = reinterpret_cast<float*>(smem + Cfg::kHaloBytes);             // [196][CS]
  float* part_sum = reinterpret_cast<float*>(smem + Cfg::kHaloBytes + Cfg::kStashBytes);
  float* part_sq = part_sum + kPix;
  float* s_mean = part_sq + kPix;
  float* s_rstd = s_mean + kPix;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int crank = blockIdx.x % cluster_size;
  const int tile_id = blockIdx.x / cluster_size;
  const int b = tile_id / tiles_per_img;
  const int t_in_img = tile_id % tiles_per_img;
  const int ty0 = (t_in_img / tiles_x) * kT, tx0 = (t_in_img % tiles_x) * kT;
  const int c_base = crank * CS;

  // ---- A. halo tile -> smem (bf16), 4 channels per lane ----
  // Loads are issued in batches of kU before any is consumed: with one CTA per SM the only way to cover
  // the HBM latency is memory-level parallelism inside each thread.
  {
    constexpr int kQuads = CS / 4;  // 4-channel groups per position
    constexpr int kTotal = kHalo * kHalo * kQuads;
    constexpr int kU = 10;
    for (int base = 0; base < kTotal; base += kThreads * kU) {
      uint2 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int idx = base + u * kThreads + tid;
        v[u] = make_uint2(0u, 0u);
        if (idx < kTotal) {
          const int qd = idx % kQuads;
          const int pos = idx / kQuads;
          const int hy = pos / kHalo, hx = pos % kHalo;
          const int gy = ty0 + hy - 3, gx = tx0 + hx - 3;
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
        if (idx < kTotal) {
          const int qd = idx % kQuads;
          const int pos = idx / kQuads;
          *reinterpret_cast<uint2*>(halo + (size_t)pos * (CS / 2) + qd * 2) = v[u];
        }
      }
    }
  }
  __syncthreads();

  // ---- B. depthwise 7x7: thread = (channel pair cp, row group rg) ----
  {
    const int cp = tid % Cfg::kSlots, rg = tid / Cfg::kSlots;
    if (cp < Cfg::kPairs) {
      const int c0 = c_base + 2 * cp;
      uint64_t w[49];
#pragma unroll
      for (int t = 0; t < 49; ++t) w[t] = pack2(__ldg(wgt + (size_t)t * C + c0), __ldg(wgt + (size_t)t * C + c0 + 1));
      const uint64_t bv = pack2(__ldg(bias + c0), __ldg(bias + c0 + 1));
#pragma unroll 1
      for (int s = rg; s < 2 * kT; s += kThreads / Cfg::kSlots) {
        const int oy = s >> 1, ox0 = (s & 1) * 7;
        if (ty0 + oy >= H || tx0 + ox0 >= W) continue;
        uint64_t acc[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) acc[i] = bv;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t* row = halo + (size_t)((oy + ky) * kHalo + ox0) * (CS / 2) + cp;
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
          *reinterpret_cast<float2*>(stash + (size_t)(oy * kT + ox0 + i) * CS + 2 * cp) = make_float2(a0, a1);
        }
      }
    }
  }
  __syncthreads();

  // ---- C. LayerNorm statistics over all C channels of each pixel (cluster-wide) ----
  cg::cluster_group cluster = cg::this_cluster();
  const float inv_c = 1.0f / (float)C;
  auto pixel_valid = [&](int p) { return (ty0 + p / kT) < H && (tx0 + p % kT) < W; };
  for (int p = warp; p < kPix; p += kThreads / 32) {
    float s = 0.f;
    if (pixel_valid(p))
      for (int c = lane; c < CS; c += 32) s += stash[(size_t)p * CS + c];
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
    if (pixel_valid(p)) {
      const float m = s_mean[p];
      for (int c = lane; c < CS; c += 32) {
        const float d = stash[(size_t)p * CS + c] - m;
        s += d * d;
      }
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

  // ---- D. normalise the own slab and write ----
  for (int p = warp; p < kPix; p += kThreads / 32) {
    if (!pixel_valid(p)) continue;
    const float m = s_mean[p], rs = s_rstd[p];
    OutT* orow = out + (((long)b * H + ty0 + p / kT) * W + tx0 + p % kT) * C + c_base;
    for (int c = lane * 4; c < CS; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(stash + (size_t)p * CS + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c_base + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c_base + c));
      const float o0 = (v.x - m) * rs * g.x + be.x, o1 = (v.y - m) * rs * g.y + be.y;
      const float o2 = (v.z - m) * rs * g.z + be.z, o3 = (v.w - m) * rs * g.w + be.w;
      if constexpr (sizeof(OutT) == 2) {
        uint2 u;
        u.x = pack_bf16x2(o0, o1);
        u.y = pack_bf16x2(o2, o3);
        *reinterpret_cast<uint2*>(orow + c) = u;
      } else {
        *reinterpret_cast<float4*>(orow + c) = make_float4(o0, o1, o2, o3);
      }
    }
  }
  // peers may still be reading this CTA's partial sums
  if (cluster_size > 1) cluster.sync();
}

template <typename InT, typename OutT, int CS>
int launch_cluster(const void* x, const float* wgt, const float* bias, const float* gamma, const float* beta,
                   void* out, int B, int H, int W, int C, float eps, cudaStream_t stream) {
  using Cfg = DwCfg<CS>;
  auto kernel = dwconv7_ln_cluster_kernel<InT, OutT, CS>;
  const int cl = C / CS;
  static bool attr_set = false;
  if (!attr_set) {
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles_x = (W + kT - 1) / kT, tiles_y = (H + kT - 1) / kT;
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
                                   reinterpret_cast<OutT*>(out), H, W, C, tiles_x, tiles_x * tiles_y, cl, eps));
  return kOk;
}

}  // namespace

// Returns kUnsupported (without setting an error) when the shape does not fit the cluster formulation, so the
// caller can use the generic kernel.
int dwconv7_ln_cluster(const void* x, int in_dtype, const float* wgt, const float* bias, const float* gamma,
                       const float* beta, void* out, int out_dtype, int B, int H, int W, int C, float eps,
                       cudaStream_t stream) {
  int cs = 0;
  for (int cand : {128, 96}) {
    if (C % cand == 0) {
      const int cl = C / cand;
      if (cl == 1 || cl == 2 || cl == 4 || cl == 8) { cs = cand; break; }
    }
  }
  if (cs == 0) return kUnsupported;
#define TFIMM_DWC(IN, OUT)                                                                                  \
  return cs == 128 ? launch_cluster<IN, OUT, 128>(x, wgt, bias, gamma, beta, out, B, H, W, C, eps, stream)  \
                   : launch_cluster<IN, OUT, 96>(x, wgt, bias, gamma, beta, out, B, H, W, C, eps, stream)
  if (in_dtype == kF32 && out_dtype == kBF16) { TFIMM_DWC(float, __nv_bfloat16); }
  if (in_dtype == kBF16 && out_dtype == kBF16) { TFIMM_DWC(__nv_bfloat16, __nv_bfloat16); }
#undef TFIMM_DWC
  return kUnsupported;
}

}  // namespace tfimm
