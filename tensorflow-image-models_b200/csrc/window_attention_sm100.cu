// Swin (shifted-)window attention on tcgen05: tokens per window N <= 64, head_dim 32.
//
// Reference: WindowAttention.call (tfimm/architectures/swin.py:159-198) wrapped by SwinTransformerBlock.call's
// tf.roll -> window_partition -> ... -> window_reverse -> tf.roll (swin.py:299-313).  The five full-tensor copies are
// row permutations: here they are a row-index table (row_map) used by the gather and by the final scatter.
//
// A work item is TWO windows of one head (2 x 64 padded rows = one 128-row UMMA tile).  Persistent CTAs, warp roles:
//   2 loader warps   cp.async gather of the q / k / v rows of the item (16-byte chunks, 64-byte rows written in the
//                    SWIZZLE_64B pattern) into a 4-stage shared-memory ring; two items of copies in flight per thread
//   1 MMA warp       S = Q K^T as ONE 128 x 128 x 32 tcgen05.mma chain (only the two diagonal 64 x 64 blocks are used:
//                    the tensor pipe is idle anyway), later O = P V (128 x 32 x 128, A = P from tensor memory, B = V as
//                    an MN-major SWIZZLE_64B operand).  Software-pipelined: S of item i+2 is issued before P V of item i
//   2 x 4 softmax warps (groups alternate items; TWO 128-column TMEM buffers each, so the scores of a group's next
//                    item are already there when it finishes the current one): thread = query row.  Its 64
//                    scores come out of tensor memory ONCE (no second pass: a window row fits the register file),
//                    + relative-position bias row (padded [H][64][64] table, 16-byte loads) and -100 where the
//                    shift-region labels differ (one 64-bit mask per row, precomputed on the host), fp32 softmax with
//                    ex2.approx, P as packed bf16 back into tensor memory (zeros in the other window's columns so
//                    that the block-diagonal structure survives the 128-key P V product), O / rowsum -> bf16 -> one
//                    64-byte row segment per thread straight to the token's row of the output.
// Round 1 ran one warp per (window, head) on mma.sync with serial gather -> compute -> scatter: 1.4 TB/s.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kWinRows = 64;                     // padded tokens per window
constexpr int kItemRows = 2 * kWinRows;          // 128: UMMA M
constexpr int kDh = 32;
constexpr int kRowBytes = kDh * 2;               // 64
constexpr int kTileBytes = kItemRows * kRowBytes;    // 8 KB per q / k / v
constexpr int kStageBytes = 3 * kTileBytes;          // 24 KB
constexpr int kStages = 6;
constexpr int kMapMax = 4096;                    // row_map entries staged in shared memory (nw_img * N)
constexpr int kLoaderWarps = 2;
constexpr int kThreads = (8 + 1 + kLoaderWarps) * 32;   // 352
// TMEM: 4 buffers of 128 columns = (group, parity of the group's item count).  A buffer holds the scores S [0,128),
// then P (packed bf16) over [0,64) and the output accumulator O over [64,96) -- both dead score columns by then.
constexpr uint32_t kBufCols = 128;
constexpr uint32_t kOCol = 64;
constexpr int kNumBars = 2 * kStages + 16;
constexpr int kSmemBytes = kStages * kStageBytes + kNumBars * 8 + 16 + kMapMax * 4 + 1024;

// 16-byte chunk c of row r in the SWIZZLE_64B pattern (Swizzle<2,4,3>: address bits [4,6) ^= bits [7,9))
__device__ __forceinline__ uint32_t sw64(int row, int chunk) {
  return (uint32_t)(row * kRowBytes + ((chunk ^ ((row >> 1) & 3)) << 4));
}
// K-major operand, rows of 64 bytes (32 bf16), 8-row groups 512 bytes apart; layout type 4 = SWIZZLE_64B
// (cute/arch/mma_sm100_desc.hpp, SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// MN-major operand (V[key][dh] as B[N = dh][K = key]): 32 contiguous MN elements (64 bytes) per K index, 8 K indices
// per swizzle atom (512 bytes, SBO); one MN block only (LBO unused).
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
window_attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                           const float* __restrict__ bias_pad /*[H][64][64]*/, const int* __restrict__ row_map,
                           const unsigned long long* __restrict__ maskbits /*[nw_img][64] or null*/, long total_windows,
                           int nw_img, int N, int H, long total_items, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  // per TMEM buffer tb = 2 * group + (k & 1), k = index of the item among the group's items
  auto sfull_bar = [&](int tb) { return bars + 8u * (2 * kStages + tb); };
  auto pready_bar = [&](int tb) { return bars + 8u * (2 * kStages + 4 + tb); };
  auto ofull_bar = [&](int tb) { return bars + 8u * (2 * kStages + 8 + tb); };
  auto tempty_bar = [&](int tb) { return bars + 8u * (2 * kStages + 12 + tb); };
  const uint32_t tmem_ptr_smem = bars + 8u * kNumBars;
  int* s_map = reinterpret_cast<int*>(smem_gen + (bars - smem_base) + 8 * kNumBars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = nw_img * N;                 // tokens per image
  const long ld = 3L * H * kDh;             // qkv row stride (elements)
  const long pairs = (total_windows + 1) >> 1;   // window pairs; item = head * pairs + pair (pair fastest)
  const bool map_in_smem = L <= kMapMax;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), kLoaderWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int tb = 0; tb < 4; ++tb) {
      mbar_init(sfull_bar(tb), 1);
      mbar_init(pready_bar(tb), 4);
      mbar_init(ofull_bar(tb), 1);
      mbar_init(tempty_bar(tb), 4);
    }
    fence_mbar_init();
  }
  if (map_in_smem)
    for (int i = threadIdx.x; i < L; i += kThreads) s_map[i] = row_map[i];
  if (warp == 8) tmem_alloc<512>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  // token -> row of its image (32-bit: one image has < 2^31 tokens)
  auto map_at = [&](int idx) { return map_in_smem ? s_map[idx] : __ldg(row_map + idx); };

  if (warp >= 9) {
    // ------------------------------------------- loaders -------------------------------------------
    // thread = rows lt and lt + 64 of the item (the same token of the two windows), all four 16-byte chunks of q, k, v
    const int lt = (warp - 9) * 32 + lane;          // 0..63 = token
    long it = 0;
    int pending_stage = -1;
    for (long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int s = (int)(it % kStages);
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      const int h = (int)(item / pairs);
      const long gwin0 = 2 * (item % pairs);
      const long img0 = gwin0 / nw_img;
      const int wi0 = (int)(gwin0 - img0 * nw_img);
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t sQ = smem_base + s * kStageBytes, sK = sQ + kTileBytes, sV = sK + kTileBytes;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int wi = (w == 0) ? wi0 : (wi0 + 1 == nw_img ? 0 : wi0 + 1);
        const long img = (w == 0) ? img0 : (wi0 + 1 == nw_img ? img0 + 1 : img0);
        const bool valid = lt < N && gwin0 + w < total_windows;
        const long src_row = valid ? img * L + map_at(wi * N + lt) : 0;
        const __nv_bfloat16* src = qkv + src_row * ld + (long)h * kDh;
        const int r = w * kWinRows + lt;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = sw64(r, c);
          cp_async_16(sQ + off, src + c * 8, valid);
          cp_async_16(sK + off, src + H * kDh + c * 8, valid);
          cp_async_16(sV + off, src + 2 * H * kDh + c * 8, valid);
        }
      }
      cp_async_commit();
      if (pending_stage >= 0) {
        cp_async_wait<1>();            // the previous item's copies of this thread have landed
        fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(pending_stage));
      }
      pending_stage = s;
    }
    if (pending_stage >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(pending_stage));
    }
  } else if (warp == 8) {
    // ------------------------------------------ MMA issuer ------------------------------------------
    // program order: S(it), then P V (it - 2).  S of an item only needs its TMEM buffer drained by the item four
    // back, so the scores of a group's NEXT item are ready while the group is still in its softmax.
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kItemRows, kItemRows);
    constexpr uint32_t idesc_o = umma_idesc_bf16_f32(kItemRows, kDh, /*b_mn_major=*/true);
    long n_items = 0;
    for (long item = blockIdx.x; item < total_items; item += gridDim.x) ++n_items;
    for (long it = 0; it < n_items + 2; ++it) {
      if (it < n_items) {
        const int s = (int)(it % kStages);
        const int tb = (int)(2 * (it & 1) + ((it >> 1) & 1));
        mbar_wait(full_bar(s), (uint32_t)(it / kStages) & 1u);
        mbar_wait(tempty_bar(tb), ((uint32_t)(it >> 2) & 1u) ^ 1u);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sQ = smem_base + s * kStageBytes, sK = sQ + kTileBytes;
          const uint64_t dq = umma_desc_k_sw64(sQ), dk = umma_desc_k_sw64(sK);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16_ss(tmem_base + (uint32_t)tb * kBufCols, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s,
                         (uint32_t)(k != 0));
          umma_commit(sfull_bar(tb));
        }
        __syncwarp();
      }
      if (it >= 2) {
        const long jt = it - 2;
        const int s = (int)(jt % kStages);
        const int tb = (int)(2 * (jt & 1) + ((jt >> 1) & 1));
        mbar_wait(pready_bar(tb), (uint32_t)(jt >> 2) & 1u);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sV = smem_base + s * kStageBytes + 2 * kTileBytes;
          const uint32_t t0 = tmem_base + (uint32_t)tb * kBufCols;
#pragma unroll
          for (int j = 0; j < kItemRows / 16; ++j)   // 16 keys per step: 8 packed P columns, 1024 bytes of V
            umma_bf16_ts(t0 + kOCol, t0 + (uint32_t)(j * 8), umma_desc_mn_sw64(sV + (uint32_t)(j * 1024)), idesc_o,
                         (uint32_t)(j != 0));
          umma_commit(ofull_bar(tb));
          umma_commit(empty_bar(s));   // every MMA that reads this stage has completed when this arrives
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------- softmax + epilogue groups -------------------------------------
    const int g = warp >> 2;              // group: items it with (it & 1) == g
    const int q = warp & 3;               // TMEM lane quarter
    const int r = q * 32 + lane;          // row inside the item
    const int win = r >> 6, tok = r & 63;
    const float l2e = 1.4426950408889634f;
    const float sl2 = scale * l2e;
    long it = 0;
    for (long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      if ((it & 1) != g) continue;
      const int tb = 2 * g + (int)((it >> 1) & 1);
      const uint32_t par = (uint32_t)(it >> 2) & 1u;
      const uint32_t t_row = tmem_base + (uint32_t)tb * kBufCols + ((uint32_t)(q * 32) << 16);
      const int h = (int)(item / pairs);
      const long gwin0 = 2 * (item % pairs);
      const long img0 = gwin0 / nw_img;
      const int wi0 = (int)(gwin0 - img0 * nw_img);
      const int wi = (win == 0) ? wi0 : (wi0 + 1 == nw_img ? 0 : wi0 + 1);
      const long img = (win == 0) ? img0 : (wi0 + 1 == nw_img ? img0 + 1 : img0);
      const bool row_ok = tok < N && gwin0 + win < total_windows;
      // bias row and mask bits of this query row: issued before the wait on the scores
      float b[52];
      const float4* brow = reinterpret_cast<const float4*>(bias_pad + ((long)h * 64 + tok) * 64);
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const float4 v = __ldg(brow + j);
        b[4 * j] = v.x; b[4 * j + 1] = v.y; b[4 * j + 2] = v.z; b[4 * j + 3] = v.w;
      }
      unsigned long long mbits = 0ull;
      if (maskbits != nullptr && row_ok) mbits = __ldg(maskbits + (long)wi * 64 + tok);
      const long out_row = row_ok ? img * L + map_at(wi * N + tok) : 0;

      mbar_wait(sfull_bar(tb), par);
      tcgen05_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(t_row + (uint32_t)(win * 64), s0);
      tmem_ld_32x32b_x32(t_row + (uint32_t)(win * 64 + 32), s1);
      tmem_ld_wait();
      // logits (log2 domain) of the N <= 52 keys of this row's own window
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 52; ++j) {
        const float sv = __uint_as_float(j < 32 ? s0[j] : s1[j - 32]);
        float v = fmaf(sv, sl2, b[j] * l2e);
        if ((mbits >> j) & 1ull) v -= 100.0f * l2e;
        v = j < N ? v : -INFINITY;
        b[j] = v;
        mx = fmaxf(mx, v);
      }
      float sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float p0 = 0.f, p1 = 0.f;
        if (2 * j < 52) p0 = ex2_approx(b[2 * j] - mx);          // exp2(-inf) = 0 for the padded keys
        if (2 * j + 1 < 52) p1 = ex2_approx(b[2 * j + 1] - mx);
        sum += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      // every thread of the group must have its scores in registers before anybody overwrites columns [0, 64)
      // (P of window 1 lands on the score columns of window 0 -- in OTHER lanes' rows, but the same TMEM columns of
      // its own lane only: no hazard across lanes; the barrier is only needed against the MMA, which is idle here)
      {
        uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v8[j] = pk[8 * c + j];
          tmem_st_32x32b_x8(t_row + (uint32_t)(win * 32 + 8 * c), v8);
          tmem_st_32x32b_x8(t_row + (uint32_t)((1 - win) * 32 + 8 * c), z);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pready_bar(tb));

      mbar_wait(ofull_bar(tb), par);
      tcgen05_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_row + kOCol, o);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(tb));   // the MMA warp may reuse this buffer (four items from now)
      if (row_ok) {
        const float inv = 1.0f / sum;
        uint4* dst = reinterpret_cast<uint4*>(out + out_row * ((long)H * kDh) + (long)h * kDh);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          dst[c] = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// bias_pad: [H][64][64] fp32 (rows / columns beyond N are ignored); maskbits: [nw_img][64] uint64, bit j of entry
// (wi, i) set when tokens i and j of window wi lie in different shift regions (null for unshifted blocks).
int window_attention_tc_bf16(const void* qkv, void* out, const float* bias_pad, const int* row_map,
                             const unsigned long long* maskbits, int B, int nw_img, int N, int H, int dh, float scale,
                             cudaStream_t stream) {
  TFIMM_CHECK_ARG(B > 0 && nw_img > 0 && N > 0 && H > 0, "window_attention: bad shape");
  TFIMM_CHECK_ARG(bias_pad != nullptr && row_map != nullptr, "window_attention: bias and row_map are required");
  if (dh != kDh || N > 52) {
    set_last_error("window_attention: the tcgen05 kernel takes head_dim 32 and <= 52 tokens per window (got dh=%d N=%d)",
                   dh, N);
    return kUnsupported;
  }
  TFIMM_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(bias_pad) & 15u) == 0,
                  "window_attention: qkv / out / bias must be 16-byte aligned");
  const long total_windows = (long)B * nw_img;
  const long pairs = (total_windows + 1) / 2;
  const long items = pairs * H;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs))
    TFIMM_CUDA_OK(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const long grid = items < sm_count() ? items : sm_count();
  window_attention_tc_kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), bias_pad, row_map, maskbits,
      total_windows, nw_img, N, H, items, scale);
  TFIMM_LAUNCH_OK("window_attention_tc_kernel");
  return kOk;
}

}  // namespace tfimm
