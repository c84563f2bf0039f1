// Dense layer with a SHORT contraction (K <= 64): a streaming kernel, not a tile pipeline.
//
//     C[M,N] = act(A[M,K] @ W[N,K]^T + bias[N])        bf16 in / bf16 out, fp32 accumulate
//
// Call sites: the 1x1 convolutions of EfficientNet's first stages and stems -- (K, N) = (27->32, 48), (48, 24),
// (24, 144), (32, 192), (56, 336) with M up to 9.2 M pixels at 380 px -- and the patch-embedding GEMMs (K = 48) of
// Swin / ConvNeXt (tfimm/architectures/efficientnet_blocks.py:412-434, layers/transformers.py:131-139).  At K <= 64
// the work is ~40 flop per byte: HBM-bound, with the activation (one MUFU.EX2 per element) as the second term.  The
// persistent tcgen05 kernel handles such shapes one 128 x 64 tile per barrier round trip (TMA -> UMMA -> tcgen05.ld
// -> TMA store, ~2.5 us per tile): B4's 9.2 M x 24 -> 144 expansion ran at 1.5 TB/s (2.06 ms of a 28 ms forward).
// Here the whole weight matrix sits in shared memory for the life of the CTA, warps stream 16-row slices of A through
// cp.async double buffers, multiply on mma.sync (HMMA -- the tensor pipe is idle either way at this intensity) and
// write 128-byte row segments; occupancy (16 warps / SM), not a pipeline, hides the latency.
#include "common.cuh"

namespace tfimm {
namespace {

constexpr int kSkM = 128;        // rows per CTA step (8 warps x 16)
constexpr int kSkWarps = 8;
constexpr int kSkNMax = 512;
constexpr int kSkRowBytes = 128; // smem pitch of A / W rows: 64 bf16, zero-padded past K, 16-byte chunks XOR-swizzled

__device__ __forceinline__ uint32_t sk_off(int row, int chunk) {
  return (uint32_t)(row * kSkRowBytes + ((chunk ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(kSkWarps * 32, 2)
gemm_bf16_skinny_kernel(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ W, int ldw,
                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ C, int ldc, int M, int N, int K,
                        int act, int n_rows_w) {
  extern __shared__ __align__(128) uint8_t smem[];
  // [W: n_rows_w x 128 B][A buffers: 2 x 128 x 128 B][per-warp output staging: 8 x 16 rows x 128 B]
  const uint32_t sW = smem_u32(smem);
  const uint32_t sA = sW + (uint32_t)n_rows_w * kSkRowBytes;
  uint8_t* stage_gen = smem + (size_t)n_rows_w * kSkRowBytes + 2 * kSkM * kSkRowBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kchunks = (K + 7) >> 3;           // 16-byte chunks per row that hold data
  const int ksteps = (K + 15) >> 4;           // mma k-steps (<= 4)
  const int num_tiles = (M + kSkM - 1) / kSkM;

  // weights: once per CTA
  for (int idx = tid; idx < n_rows_w * 8; idx += kSkWarps * 32) {
    const int r = idx >> 3, c = idx & 7;
    const bool valid = r < N && c < kchunks;
    cp_async_16(sW + sk_off(r, c), W + (long)(valid ? r : 0) * ldw + (valid ? c : 0) * 8, valid);
  }
  auto load_a = [&](int tile, int buf) {
    const long row0 = (long)tile * kSkM;
    for (int idx = tid; idx < kSkM * 8; idx += kSkWarps * 32) {
      const int r = idx >> 3, c = idx & 7;
      const bool valid = row0 + r < M && c < kchunks;
      cp_async_16(sA + (uint32_t)buf * (kSkM * kSkRowBytes) + sk_off(r, c),
                  A + (valid ? (row0 + r) * (long)lda + c * 8 : 0), valid);
    }
  };
  int tile = blockIdx.x;
  if (tile < num_tiles) load_a(tile, 0);
  cp_async_commit();

  const int g = lane >> 2, t = lane & 3;
  uint8_t* my_stage = stage_gen + (size_t)warp * 16 * kSkRowBytes;
  int it = 0;
  for (; tile < num_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tile + (int)gridDim.x < num_tiles) load_a(tile + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();      // this tile's A (and, the first time, W) has landed
    __syncthreads();
    const uint32_t a_base = sA + (uint32_t)buf * (kSkM * kSkRowBytes);
    const long row0 = (long)tile * kSkM + warp * 16;
    if (row0 < M) {
      uint32_t af[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks < ksteps) {
          const int row = warp * 16 + (lane & 15);
          ldmatrix_x4(a_base + sk_off(row, ks * 2 + (lane >> 4)), af[ks][0], af[ks][1], af[ks][2], af[ks][3]);
        }
      }
      for (int nc = 0; nc < N; nc += 64) {
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
          if (nc + nt * 8 < N) {
            const int row = nc + nt * 8 + (lane & 7);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (2 * j < ksteps) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(sW + sk_off(row, 4 * j + (lane >> 3)), b0, b1, b2, b3);
                mma_bf16_16816(acc[nt], af[2 * j], b0, b1);
                if (2 * j + 1 < ksteps) mma_bf16_16816(acc[nt], af[2 * j + 1], b2, b3);
              }
            }
          }
        }
        // epilogue on packed pairs: (row g: cols 2t, 2t+1) and (row g+8: same cols) of each 8-column tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int col = nc + nt * 8 + 2 * t;
          float b0 = 0.f, b1 = 0.f;
          if (bias != nullptr && col < N) {
            const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + col));
            b0 = bb.x;
            b1 = bb.y;
          }
          uint64_t lo = pack2(acc[nt][0] + b0, acc[nt][1] + b1), hi = pack2(acc[nt][2] + b0, acc[nt][3] + b1);
          if (act == kActSwish) swish4(lo, hi);
          else if (act == kActGelu) gelu4(lo, hi);
          else if (act != kActNone) {
            float a0, a1, a2, a3;
            unpack2(lo, a0, a1);
            unpack2(hi, a2, a3);
            lo = pack2(apply_act<false>(a0, act), apply_act<false>(a1, act));
            hi = pack2(apply_act<false>(a2, act), apply_act<false>(a3, act));
          }
          float a0, a1, a2, a3;
          unpack2(lo, a0, a1);
          unpack2(hi, a2, a3);
          // staging tile [16 rows][64 cols] bf16, 16-byte chunks XOR-swizzled by row
          *reinterpret_cast<uint32_t*>(my_stage + g * kSkRowBytes + ((nt ^ (g & 7)) << 4) + 4 * t) = pack_bf16x2(a0, a1);
          *reinterpret_cast<uint32_t*>(my_stage + (g + 8) * kSkRowBytes + ((nt ^ ((g + 8) & 7)) << 4) + 4 * t) =
              pack_bf16x2(a2, a3);
        }
        __syncwarp();
        // 16 rows x 8 chunks of 16 B: four rows per instruction, 128 contiguous bytes per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 4 + (lane >> 3), c = lane & 7;
          if (row0 + r < M && nc + c * 8 < N) {
            const uint4 v = *reinterpret_cast<const uint4*>(my_stage + r * kSkRowBytes + ((c ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(C + (row0 + r) * (long)ldc + nc + c * 8) = v;
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();   // every warp is done with A buffer `buf` before the next iteration's prefetch overwrites it
  }
  cp_async_wait<0>();
}

}  // namespace

// Returns kUnsupported (without setting an error) for shapes outside this kernel: the caller uses the tcgen05 path.
int gemm_bf16_skinny(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M, int N,
                     int K, int act, cudaStream_t stream) {
  if (K > 64 || K % 8 != 0 || N % 8 != 0 || N > kSkNMax || M < 4096) return kUnsupported;
  if (lda % 8 != 0 || ldw % 8 != 0 || ldc % 8 != 0) return kUnsupported;
  if ((reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(W) & 15u) ||
      (reinterpret_cast<uintptr_t>(C) & 15u) || (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 7u)))
    return kUnsupported;
  const int n_rows_w = (N + 7) / 8 * 8;
  const int smem = n_rows_w * kSkRowBytes + 2 * kSkM * kSkRowBytes + kSkWarps * 16 * kSkRowBytes;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs))
    TFIMM_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSkNMax * kSkRowBytes + 2 * kSkM * kSkRowBytes + kSkWarps * 16 * kSkRowBytes));
  const int tiles = (M + kSkM - 1) / kSkM;
  const int max_ctas = 2 * sm_count();
  const int grid = tiles < max_ctas ? tiles : max_ctas;
  gemm_bf16_skinny_kernel<<<grid, kSkWarps * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W), ldw, bias,
      reinterpret_cast<__nv_bfloat16*>(C), ldc, M, N, K, act, n_rows_w);
  TFIMM_LAUNCH_OK("gemm_bf16_skinny_kernel");
  return kOk;
}

}  // namespace tfimm
