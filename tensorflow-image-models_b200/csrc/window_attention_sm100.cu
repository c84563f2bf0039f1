// Swin (shifted-)window attention on tcgen05: tokens per window N <= 52, head_dim 32.
//
// Reference: WindowAttention.call (tfimm/architectures/swin.py:159-198) wrapped by SwinTransformerBlock.call's
// tf.roll -> window_partition -> ... -> window_reverse -> tf.roll (swin.py:299-313).  The five full-tensor copies are
// row permutations: here they are a row-index table (row_map) used by the gather and by the final scatter.
//
// A work item is TWO windows of one head (2 x 64 padded rows = one 128-row UMMA tile).  The two windows share one
// tensor-core instruction through a block structure in the CONTRACTION dimension:
//     Q tile [128 rows][64]:  rows of window 0 = [q | 0],  rows of window 1 = [0 | q]
//     K tile [ 64 keys][64]:  row j = [k_win0[j] | k_win1[j]]            =>  S[r][j] = q_r . k_{win(r)}[j]   (128 x 64)
//     V tile [ 64 keys][64]:  row j = [v_win0[j] | v_win1[j]]            =>  (P V)[r][32 win(r) + d] = O[r][d]  (128 x 64)
// so the scores take 64 TMEM columns (not 128 with two dead off-diagonal blocks), P needs no zero blocks, and there is
// room for separate S/P and O buffers: 2 groups x (2 + 2) x 64 columns = 512.
//
// Persistent CTAs, warp roles:
//   2 loader warps   cp.async gather of the q / k / v rows of the item (16-byte chunks, SWIZZLE_128B pattern) into a
//                    6-stage shared-memory ring, four items of copies in flight per thread; the zero halves of the Q tiles are written once at kernel start
//   1 MMA warp       S = Q K^T (128 x 64 x 64) as soon as the item's stage is full and the S buffer is free, i.e. the
//                    P V product of the group's item before last has retired -- while the group is still busy with
//                    the softmax of the item in between
//   2 x 4 softmax warps (groups alternate items): thread = query row.  Its 64 scores come out of tensor memory ONCE
//                    (a window row fits the register file), + relative-position bias row (padded [H][64][64] table,
//                    16-byte loads) and -100 where the shift-region labels differ (one 64-bit mask per row, precomputed
//                    on the host), fp32 softmax with ex2.approx, P as packed bf16 back over the scores, a 128-thread
//                    named barrier, then the group's first thread issues O = P V itself (A = P from tensor memory,
//                    B = V as an MN-major operand).  The output of an item (O / rowsum -> bf16 -> one 64-byte row
//                    segment per thread straight to the token's row) is stored after the NEXT item's P V has been
//                    issued, so neither tensor-core latency is on the group's critical path.
// Round 1 ran one warp per (window, head) on mma.sync with serial gather -> compute -> scatter: 1.4 TB/s.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kWinRows = 64;                     // padded tokens per window
constexpr int kItemRows = 2 * kWinRows;          // 128: UMMA M
constexpr int kDh = 32;
constexpr int kQBytes = kItemRows * 128;         // 16 KB: 128 rows x 64 bf16 (half of each row is zero)
constexpr int kKBytes = kWinRows * 128;          //  8 KB: 64 keys x (32 + 32) bf16
constexpr int kStageBytes = kQBytes + 2 * kKBytes;   // 32 KB
constexpr int kStages = 6;
constexpr int kInflight = 4;                     // items of cp.async copies in flight per loader thread
constexpr int kMapMax = 4096;                    // row_map entries staged in shared memory (nw_img * N)
constexpr int kLoaderWarps = 2;
constexpr int kThreads = (8 + 1 + kLoaderWarps) * 32;   // 352
// TMEM columns of group g: S/P buffers at 256 g + {0, 64}, O buffers at 256 g + {128, 192}
constexpr int kNumBars = 2 * kStages + 8;
constexpr int kSmemBytes = kStages * kStageBytes + kNumBars * 8 + 16 + kMapMax * 4 + 1024;

// 16-byte chunk c of 128-byte row r in the SWIZZLE_128B pattern
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
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
  // per (group, parity of the group's item count): tb = 2 * group + (k & 1)
  auto sfull_bar = [&](int tb) { return bars + 8u * (2 * kStages + tb); };       // scores written
  auto ofull_bar = [&](int tb) { return bars + 8u * (2 * kStages + 4 + tb); };   // P V retired: O ready, S/P buffer free
  const uint32_t tmem_ptr_smem = bars + 8u * kNumBars;
  int* s_map = reinterpret_cast<int*>(smem_gen + (bars - smem_base) + 8 * kNumBars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = nw_img * N;                 // tokens per image
  const int ld = 3 * H * kDh;               // qkv row stride (elements)
  const int ldo = H * kDh;
  const long pairs = (total_windows + 1) >> 1;   // window pairs; item = head * pairs + pair (pair fastest)
  const bool map_in_smem = L <= kMapMax;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), kLoaderWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int tb = 0; tb < 4; ++tb) {
      mbar_init(sfull_bar(tb), 1);
      mbar_init(ofull_bar(tb), 1);
    }
    fence_mbar_init();
  }
  if (map_in_smem)
    for (int i = threadIdx.x; i < L; i += kThreads) s_map[i] = row_map[i];
  // Q tiles: zero everything once; the loaders only ever write the non-zero half of a row
  for (int i = threadIdx.x; i < kStages * (kQBytes / 16); i += kThreads) {
    const int s = i / (kQBytes / 16), c = i - s * (kQBytes / 16);
    *reinterpret_cast<uint4*>(smem_gen + (size_t)s * kStageBytes + (size_t)c * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  if (warp == 8) tmem_alloc<512>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  // token -> row of its image (32-bit: one image has < 2^31 tokens)
  auto map_at = [&](int idx) { return map_in_smem ? s_map[idx] : __ldg(row_map + idx); };
  // work index t -> (head, window pair).  Pairs are taken in blocks of 8 * gridDim: inside a block the head is the
  // slow index, so a CTA keeps one head (one bias table in L1) for 8 consecutive items, and the other 64-byte halves
  // of the 128-byte lines it gathers (the neighbouring head of the same tokens) are requested within a few tens of
  // microseconds, while the line is still in L2 -- with the head outermost over ALL pairs every line came from DRAM
  // twice (1.23 GB read for 0.62 GB of q / k / v at Swin-B stage 0).
  struct ItemPos { int h, wi0; long img0, gwin0; };
  const uint32_t pairs_per_blk = 8u * gridDim.x;
  auto locate = [&](long t) {
    ItemPos p;
    const uint32_t per_blk = pairs_per_blk * (uint32_t)H;
    const uint32_t blk = (uint32_t)t / per_blk, rem = (uint32_t)t - blk * per_blk;
    const uint32_t p0 = blk * pairs_per_blk;
    const uint32_t nb = min(pairs_per_blk, (uint32_t)pairs - p0);
    const uint32_t h = rem / nb;
    p.h = (int)h;
    p.gwin0 = 2L * (p0 + rem - h * nb);
    const uint32_t img0 = (uint32_t)p.gwin0 / (uint32_t)nw_img;
    p.img0 = img0;
    p.wi0 = (int)((uint32_t)p.gwin0 - img0 * (uint32_t)nw_img);
    return p;
  };

  if (warp >= 9) {
    // ------------------------------------------- loaders -------------------------------------------
    // thread = token lt of BOTH windows of the item, all four 16-byte chunks of q, k, v
    const int lt = (warp - 9) * 32 + lane;          // 0..63 = token
    long it = 0;
    for (long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int s = (int)(it % kStages);
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      const ItemPos ip = locate(item);
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t sQ = smem_base + s * kStageBytes, sK = sQ + kQBytes, sV = sK + kKBytes;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int wi = (w == 0) ? ip.wi0 : (ip.wi0 + 1 == nw_img ? 0 : ip.wi0 + 1);
        const long img = (w == 0) ? ip.img0 : (ip.wi0 + 1 == nw_img ? ip.img0 + 1 : ip.img0);
        const bool valid = lt < N && ip.gwin0 + w < total_windows;
        const long src_row = valid ? img * L + map_at(wi * N + lt) : 0;
        const __nv_bfloat16* src = qkv + src_row * ld + ip.h * kDh;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          cp_async_16(sQ + sw128(w * kWinRows + lt, 4 * w + c), src + c * 8, valid);
          cp_async_16(sK + sw128(lt, 4 * w + c), src + ldo + c * 8, valid);
          cp_async_16(sV + sw128(lt, 4 * w + c), src + 2 * ldo + c * 8, valid);
        }
      }
      cp_async_commit();
      if (it >= kInflight - 1) {
        cp_async_wait<kInflight - 1>();   // the copies of item it - (kInflight - 1) of this thread have landed
        fence_proxy_async_smem();         // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar((int)((it - (kInflight - 1)) % kStages)));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncwarp();
    for (long d = it < kInflight - 1 ? 0 : it - (kInflight - 1); d < it; ++d)
      if (lane == 0) mbar_arrive(full_bar((int)(d % kStages)));
  } else if (warp == 8) {
    // ------------------------------------------ MMA issuer: S only ------------------------------------------
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kItemRows, kWinRows);
    long it = 0;
    for (long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int s = (int)(it % kStages);
      const int g = (int)(it & 1), kpar = (int)((it >> 1) & 1), tb = 2 * g + kpar;
      mbar_wait(full_bar(s), (uint32_t)(it / kStages) & 1u);
      mbar_wait(ofull_bar(tb), ((uint32_t)(it >> 2) & 1u) ^ 1u);   // P V of the item four back (same buffer) retired
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t sQ = smem_base + s * kStageBytes, sK = sQ + kQBytes;
        const uint64_t dq = umma_desc_k_sw128(sQ), dk = umma_desc_k_sw128(sK);
        const uint32_t t_s = tmem_base + (uint32_t)(256 * g + 64 * kpar);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 64 = 4 x 16 contraction elements, 32 bytes apart inside the swizzle span
          umma_bf16_ss(t_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
        umma_commit(sfull_bar(tb));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------- softmax + epilogue groups -------------------------------------
    const int g = warp >> 2;              // group: items it with (it & 1) == g
    const int q = warp & 3;               // TMEM lane quarter
    const int r = q * 32 + lane;          // row inside the item
    const int win = r >> 6, tok = r & 63;
    const float l2e = 1.4426950408889634f;
    const float sl2 = scale * l2e;
    constexpr uint32_t idesc_o = umma_idesc_bf16_f32(kItemRows, 2 * kDh, /*b_mn_major=*/true);
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t_group = tmem_base + (uint32_t)(256 * g);
    // epilogue of the PREVIOUS item of this group, deferred until the next item's P V has been issued
    bool prev_pending = false, prev_ok = false;
    float prev_inv = 0.f;
    long prev_out_off = 0;
    int prev_tb = 0;
    uint32_t prev_par = 0;
    auto finish_prev = [&]() {
      mbar_wait(ofull_bar(prev_tb), prev_par);
      tcgen05_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_group + lane_off + (uint32_t)(128 + 64 * (prev_tb & 1) + 32 * win), o);
      tmem_ld_wait();
      tcgen05_fence_before();   // ordered before the named barrier that precedes the next P V into this buffer
      if (prev_ok) {
        uint4* dst = reinterpret_cast<uint4*>(out + prev_out_off);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * prev_inv, __uint_as_float(o[8 * c + 1]) * prev_inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * prev_inv, __uint_as_float(o[8 * c + 3]) * prev_inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * prev_inv, __uint_as_float(o[8 * c + 5]) * prev_inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * prev_inv, __uint_as_float(o[8 * c + 7]) * prev_inv);
          dst[c] = u;
        }
      }
      prev_pending = false;
    };
    long it = g;
    for (long item = blockIdx.x + (long)g * gridDim.x; item < total_items; item += 2L * gridDim.x, it += 2) {
      const int s = (int)(it % kStages);
      const int kpar = (int)((it >> 1) & 1), tb = 2 * g + kpar;
      const uint32_t par = (uint32_t)(it >> 2) & 1u;
      const uint32_t t_s = t_group + lane_off + (uint32_t)(64 * kpar);
      const ItemPos ip = locate(item);
      const int wi = (win == 0) ? ip.wi0 : (ip.wi0 + 1 == nw_img ? 0 : ip.wi0 + 1);
      const long img = (win == 0) ? ip.img0 : (ip.wi0 + 1 == nw_img ? ip.img0 + 1 : ip.img0);
      const bool row_ok = tok < N && ip.gwin0 + win < total_windows;
      // bias row and mask bits of this query row: issued before the wait on the scores
      float b[52];
      const float4* brow = reinterpret_cast<const float4*>(bias_pad + (ip.h * 64 + tok) * 64);
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const float4 v = __ldg(brow + j);
        b[4 * j] = v.x; b[4 * j + 1] = v.y; b[4 * j + 2] = v.z; b[4 * j + 3] = v.w;
      }
      unsigned long long mbits = 0ull;
      if (maskbits != nullptr && row_ok) mbits = __ldg(maskbits + wi * 64 + tok);
      const long out_off = row_ok ? (img * L + map_at(wi * N + tok)) * ldo + ip.h * kDh : 0;

      mbar_wait(sfull_bar(tb), par);
      tcgen05_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(t_s, s0);
      tmem_ld_32x32b_x32(t_s + 32u, s1);
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
      // P: 64 keys = 32 packed columns over the first half of the scores (every warp only touches its own lanes)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = pk[8 * c + j];
        tmem_st_32x32b_x8(t_s + (uint32_t)(8 * c), v8);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      // all 128 rows of P are in tensor memory (and this group's reads of the O buffer two items back are done):
      // the group's first thread issues O = P V and the two commits
      asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
      if (q == 0 && lane == 0) {
        tcgen05_fence_after();
        const uint32_t sV = smem_base + s * kStageBytes + kQBytes + kKBytes;
        const uint32_t t_p = t_group + (uint32_t)(64 * kpar), t_o = t_group + (uint32_t)(128 + 64 * kpar);
#pragma unroll
        for (int j = 0; j < kWinRows / 16; ++j)   // 16 keys per step: 8 packed P columns, 2048 bytes of V
          umma_bf16_ts(t_o, t_p + (uint32_t)(j * 8), umma_desc_mn_sw128(sV + (uint32_t)(j * 2048), (uint32_t)kKBytes), idesc_o,
                       (uint32_t)(j != 0));
        umma_commit(ofull_bar(tb));
        umma_commit(empty_bar(s));   // q / k / v of this stage are dead once P V has retired
      }
      __syncwarp();
      if (prev_pending) finish_prev();   // the previous item's O has long been accumulated: drain and store it now
      prev_pending = true;
      prev_ok = row_ok;
      prev_inv = 1.0f / sum;
      prev_out_off = out_off;
      prev_tb = tb;
      prev_par = par;
    }
    if (prev_pending) finish_prev();
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
  if (items >= (1L << 31) || total_windows >= (1L << 30)) {
    set_last_error("window_attention: more than 2^31 work items");
    return kUnsupported;
  }
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
