// Epilogue shared by the tcgen05 GEMM kernels (gemm_sm100.cu: one CTA per tile; gemm2_sm100.cu: CTA pair):
//     C = residual + gamma * act(acc + bias)            (or act(residual + ...) with act_post)
// One epilogue warp owns 32 accumulator rows (one TMEM lane quarter) and moves them in 128-byte column chunks:
// tcgen05.ld -> bias/act/gamma on packed fp32 pairs (FFMA2) -> + residual (TMA-prefetched into the warp's own
// smem slab while the accumulator is loaded and activated) -> 128B-swizzled slab -> TMA store.  The residual may
// alias the output (in-place fp32 residual stream).
#pragma once
#include "common.cuh"

namespace tfimm {

struct GemmParams {
  int M, N, K;
  const float* bias;   // [N] or null
  const float* gamma;  // [N] or null
  int act;
  int has_res;
  int act_post;  // 1: activation applied after the residual add (ResNet: act(x + shortcut))
  // Squeeze-excite gate folded into the A operand (gemm_sm100.cu, gated instances): row m of A is multiplied by
  // a_scale[m / a_rows_per_img][k] (fp32 [a_imgs][K]) and rounded back to bf16 in shared memory, between the TMA load
  // and the MMA -- the values the separate scale_channels pass used to write to HBM.
  const float* a_scale;
  int a_rows_per_img, a_imgs;
  // Implicit convolution (gemm_sm100.cu): the A operand is not a matrix but the NHWC input itself.  A tile's 128
  // rows are a patch of cv_pb images x cv_ph rows x cv_pw columns of OUTPUT pixels; k-block kb is tap
  // (ky, kx) = kb / (C/64) and 64 input channels, fetched as ONE 4-D TMA box whose out-of-bounds elements are
  // the zero padding.  C and the residual are NHWC too (4-D stores of pw x 32/pw pixel patches per warp).
  int conv;                  // 0: plain GEMM
  int cv_cblocks;            // C / 64
  int cv_ks, cv_stride, cv_pad;
  int cv_pb, cv_ph, cv_pw;
  int cv_tiles_x, cv_tiles_y;  // patch grid per image group (x fastest, then y, then image group)
};

struct ConvTile {
  int x, y, b;  // output-pixel coordinates of a warp's first row
};

constexpr int kEpiSlabBytes = 32 * 128;     // 32 rows x 128 B, one per epilogue warp

// v[j] (+ or *)= vec[n0 + j] on packed pairs; full chunks use 16-byte loads.
template <int CH, bool kMul>
__device__ __forceinline__ void apply_vec(uint64_t (&v)[CH / 2], const float* __restrict__ vec, int n0, int N) {
  if (n0 + CH <= N) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(vec + n0 + j));
      if (kMul) {
        v[j / 2] = mul2(v[j / 2], pack2(b4.x, b4.y));
        v[j / 2 + 1] = mul2(v[j / 2 + 1], pack2(b4.z, b4.w));
      } else {
        v[j / 2] = add2(v[j / 2], pack2(b4.x, b4.y));
        v[j / 2 + 1] = add2(v[j / 2 + 1], pack2(b4.z, b4.w));
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < CH; j += 2) {
      const float neutral = kMul ? 1.f : 0.f;
      const float b0 = (n0 + j < N) ? __ldg(vec + n0 + j) : neutral;
      const float b1 = (n0 + j + 1 < N) ? __ldg(vec + n0 + j + 1) : neutral;
      v[j / 2] = kMul ? mul2(v[j / 2], pack2(b0, b1)) : add2(v[j / 2], pack2(b0, b1));
    }
  }
}

template <int NP, bool kSharedRcp = false>
__device__ __forceinline__ void apply_act_pairs(uint64_t (&v)[NP], int act) {
  static_assert(NP % 2 == 0, "activations are evaluated on groups of four elements");
  switch (act) {
    case kActGelu:
#pragma unroll
      for (int j = 0; j < NP; j += 2) {
        gelu4<kSharedRcp>(v[j], v[j + 1]);
      }
      break;
    case kActSwish:
#pragma unroll
      for (int j = 0; j < NP; j += 2) {
        swish4<kSharedRcp>(v[j], v[j + 1]);
      }
      break;
    case kActNone:
      break;
    default:
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        float a, b;
        unpack2(v[j], a, b);
        v[j] = pack2(apply_act<false>(a, act), apply_act<false>(b, act));
      }
      break;
  }
}

// One 128-byte column chunk (CH = 128 / sizeof(OutT) columns starting at n0) of this warp's 32 rows (row0..).
//   t_addr     TMEM address of the chunk's first column in this warp's lane quarter
//   slab       this warp's 4 KB smem slab (shared::cta address), my_row = generic pointer to this lane's row
//   res_bar    this warp's residual mbarrier, res_parity = (#chunks this warp has processed) & 1
//   after_load called once the accumulator values are in registers (caller releases the TMEM stage there)
//   ct         implicit-convolution mode: where this warp's 32 rows (a pw x 32/pw pixel patch of one image) sit
//              in the NHWC output; null for a plain row-major C
//   CHW        columns per chunk: 128 bytes of output per row by default; the CTA-pair kernel's bf16 instance uses
//              32 columns (64-byte rows, 64B swizzle, 2 KB slabs) so that sixteen epilogue warps fit
//   kStoresInFlight  1 when the caller passes alternating slabs (two per warp)
template <typename OutT, int CHW = 128 / (int)sizeof(OutT), int kStoresInFlight = 0, typename AfterLoad>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, uint32_t t_addr, int n0, int row0,
                                               uint32_t slab, uint8_t* my_row, int lane, uint32_t res_bar,
                                               uint32_t res_parity, const CUtensorMap* tmap_c,
                                               const CUtensorMap* tmap_r, const ConvTile* ct,
                                               AfterLoad after_load) {
  constexpr int CH = CHW;
  constexpr int RB = CH * (int)sizeof(OutT);  // bytes per slab row: 128 or 64
  constexpr int UNITS = RB / 16;
  static_assert(RB == 128 || RB == 64, "slab rows are 128 or 64 bytes");
  // TMA swizzle: 16-byte unit j of row r lives at j ^ (r & 7) (SWIZZLE_128B) or j ^ ((r >> 1) & 3) (SWIZZLE_64B)
  const int sw = RB == 128 ? (lane & 7) : ((lane >> 1) & 3);
  // the previous store of this warp FROM THIS SLAB must have finished reading it (kStoresInFlight = 1: the caller
  // alternates between two slabs, so the store issued one chunk ago may still be in flight)
  if (lane == 0) {
    tma_store_wait_read<kStoresInFlight>();
    if (p.has_res) {
      mbar_expect_tx(res_bar, 32 * RB);
      if (ct == nullptr) tma_load_2d(slab, tmap_r, res_bar, n0, row0);
      else tma_load_4d(slab, tmap_r, res_bar, n0, ct->x, ct->y, ct->b);
    }
  }
  __syncwarp();
  uint64_t v[CH / 2];
  {
    uint32_t r[32];
#pragma unroll
    for (int h = 0; h < CH / 32; ++h) {
      tmem_ld_32x32b_x32(t_addr + (uint32_t)(h * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[h * 16 + j] = pack2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
    }
  }
  after_load();
  if (p.bias != nullptr) apply_vec<CH, false>(v, p.bias, n0, p.N);
  if (!p.act_post) apply_act_pairs(v, p.act);
  if (p.gamma != nullptr) apply_vec<CH, true>(v, p.gamma, n0, p.N);
  if (p.has_res) {
    mbar_wait(res_bar, res_parity);
#pragma unroll
    for (int j = 0; j < UNITS; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(my_row + ((j ^ sw) << 4));
      if constexpr (sizeof(OutT) == 2) {
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y);
        const float2 f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        v[4 * j + 0] = add2(v[4 * j + 0], pack2(f0.x, f0.y));
        v[4 * j + 1] = add2(v[4 * j + 1], pack2(f1.x, f1.y));
        v[4 * j + 2] = add2(v[4 * j + 2], pack2(f2.x, f2.y));
        v[4 * j + 3] = add2(v[4 * j + 3], pack2(f3.x, f3.y));
      } else {
        v[2 * j + 0] = add2(v[2 * j + 0], pack2(__uint_as_float(u.x), __uint_as_float(u.y)));
        v[2 * j + 1] = add2(v[2 * j + 1], pack2(__uint_as_float(u.z), __uint_as_float(u.w)));
      }
    }
  }
  if (p.act_post) apply_act_pairs(v, p.act);
#pragma unroll
  for (int j = 0; j < UNITS; ++j) {
    uint4 u;
    if constexpr (sizeof(OutT) == 2) {
      float a0, a1, a2, a3, a4, a5, a6, a7;
      unpack2(v[4 * j + 0], a0, a1);
      unpack2(v[4 * j + 1], a2, a3);
      unpack2(v[4 * j + 2], a4, a5);
      unpack2(v[4 * j + 3], a6, a7);
      u.x = pack_bf16x2(a0, a1);
      u.y = pack_bf16x2(a2, a3);
      u.z = pack_bf16x2(a4, a5);
      u.w = pack_bf16x2(a6, a7);
    } else {
      float a0, a1, a2, a3;
      unpack2(v[2 * j + 0], a0, a1);
      unpack2(v[2 * j + 1], a2, a3);
      u.x = __float_as_uint(a0);
      u.y = __float_as_uint(a1);
      u.z = __float_as_uint(a2);
      u.w = __float_as_uint(a3);
    }
    *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = u;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (ct == nullptr) tma_store_2d(tmap_c, slab, n0, row0);
    else tma_store_4d(tmap_c, slab, n0, ct->x, ct->y, ct->b);
    tma_store_commit();
  }
}

}  // namespace tfimm
