// Fused multi-head self-attention for the ViT family:
//     out[b, n, h, :] = softmax(scale * q k^T) v        (per image b, head h)
// reading q/k/v straight out of the packed qkv projection (B*N, 3*H*dh) whose
// column order is [q | k | v], each head-major -- exactly the layout produced
// by the reshape/transpose in tfimm/architectures/vit.py:149-165.  The
// (B,H,N,N) score tensor the reference materialises (vit.py:160-163) never
// leaves the SM.
//
// bf16 path: one CTA per (image, head, 224-query chunk); K and V of the head are
// staged once in XOR-swizzled shared memory with cp.async, each warp owns 16-row
// query tiles and runs a flash-style online softmax over 64-key blocks with
// mma.sync m16n8k16 (fp32 accumulate, fp32 softmax statistics).
//
// fp32 path (precision="fp32" parity mode): plain SIMT, one warp per query row.
#include "common.cuh"

#include <stdlib.h>

namespace tfimm {
namespace {

constexpr int kDH = 64;

template <int NW, int TPW>
__global__ void __launch_bounds__(NW * 32, 2)
vit_attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                          int N, int H, float scale_log2) {
  constexpr int ROWS = NW * TPW * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const int npad = (N + 15) & ~15;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + ROWS * 128;
  const uint32_t sV = sK + npad * 128;

  const int b = blockIdx.z, h = blockIdx.y;
  const int q_base = blockIdx.x * ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long ld = 3L * H * kDH;
  const __nv_bfloat16* base = qkv + (long)b * N * ld + h * kDH;

  // ---- stage Q (this chunk), K, V (whole head) in swizzled smem ----
  for (int idx = tid; idx < ROWS * 8; idx += NW * 32) {
    const int r = idx >> 3, c = idx & 7;
    const int gr = q_base + r;
    const bool valid = gr < N;
    cp_async_16(sQ + r * 128 + ((c ^ (r & 7)) << 4), base + (long)(valid ? gr : 0) * ld + c * 8, valid);
  }
  for (int idx = tid; idx < npad * 8; idx += NW * 32) {
    const int r = idx >> 3, c = idx & 7;
    const bool valid = r < N;
    const __nv_bfloat16* src = base + (long)(valid ? r : 0) * ld + c * 8;
    const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
    cp_async_16(sK + off, src + H * kDH, valid);
    cp_async_16(sV + off, src + 2 * H * kDH, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int nblocks = (npad + 63) >> 6;

#pragma unroll 1
  for (int tt = 0; tt < TPW; ++tt) {
    const int tile = tt * NW + warp;      // round-robin so short sequences stay balanced
    const int q0 = tile * 16;             // row inside this CTA's chunk
    if (q_base + q0 >= N) continue;

    uint32_t qf[kDH / 16][4];
#pragma unroll
    for (int ks = 0; ks < kDH / 16; ++ks) {
      const int row = q0 + (lane & 15);
      const int chunk = ks * 2 + (lane >> 4);
      ldmatrix_x4(sQ + row * 128 + ((chunk ^ (row & 7)) << 4), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float o[kDH / 8][4];
#pragma unroll
    for (int i = 0; i < kDH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

#pragma unroll 1
    for (int kb = 0; kb < nblocks; ++kb) {
      const int key0 = kb * 64;
      const int ntv = min(8, (npad - key0) >> 3);  // valid 8-key tiles in this block (even)
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (nt < ntv) {
#pragma unroll
          for (int j = 0; j < kDH / 32; ++j) {
            const int row = key0 + nt * 8 + (lane & 7);
            const int chunk = 4 * j + (lane >> 3);
            uint32_t k0, k1, k2, k3;
            ldmatrix_x4(sK + row * 128 + ((chunk ^ (row & 7)) << 4), k0, k1, k2, k3);
            mma_bf16_16816(s[nt], qf[2 * j], k0, k1);
            mma_bf16_16816(s[nt], qf[2 * j + 1], k2, k3);
          }
        }
      }
      // scale, mask, row max
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + nt * 8 + 2 * t + (e & 1);
          const float val = (nt < ntv && key < N) ? s[nt][e] * scale_log2 : -INFINITY;
          s[nt][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
      float alpha[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        alpha[r] = exp2f(m_run[r] - m_new);
        m_run[r] = m_new;
        l_run[r] *= alpha[r];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(s[nt][e] - m_run[e >> 1]);
          s[nt][e] = pv;
          l_run[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int i = 0; i < kDH / 8; ++i) {
        o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
        o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
      }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (2 * kk < ntv) {
          uint32_t a[4];
          a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
          a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
          a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int jp = 0; jp < kDH / 16; ++jp) {
            const int row = key0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int chunk = 2 * jp + (lane >> 4);
            uint32_t v0, v1, v2, v3;
            ldmatrix_x4_trans(sV + row * 128 + ((chunk ^ (row & 7)) << 4), v0, v1, v2, v3);
            mma_bf16_16816(o[2 * jp], a, v0, v1);
            mma_bf16_16816(o[2 * jp + 1], a, v2, v3);
          }
        }
      }
    }
    // normalise and write through the (now dead) Q tile in smem for coalesced stores
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[r] = 1.0f / l;
    }
    __syncwarp();
    uint8_t* tile_gen = smem + q0 * 128;
#pragma unroll
    for (int nt = 0; nt < kDH / 8; ++nt) {
      const int r0 = g, r1 = g + 8;
      *reinterpret_cast<uint32_t*>(tile_gen + r0 * 128 + ((nt ^ (r0 & 7)) << 4) + t * 4) =
          pack_bf16x2(o[nt][0] * inv[0], o[nt][1] * inv[0]);
      *reinterpret_cast<uint32_t*>(tile_gen + r1 * 128 + ((nt ^ (r1 & 7)) << 4) + t * 4) =
          pack_bf16x2(o[nt][2] * inv[1], o[nt][3] * inv[1]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int r = idx >> 3, c = idx & 7;
      const int gr = q_base + q0 + r;
      if (gr < N) {
        const uint4 val = *reinterpret_cast<const uint4*>(tile_gen + r * 128 + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(out + ((long)b * N + gr) * ((long)H * kDH) + h * kDH + c * 8) = val;
      }
    }
  }
}

template <int NW, int TPW>
int launch_vit_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int H, float scale,
                         cudaStream_t stream) {
  constexpr int ROWS = NW * TPW * 16;
  const int npad = (N + 15) & ~15;
  const size_t smem = (size_t)(ROWS + 2 * npad) * 128;
  if (smem > 227 * 1024) {
    set_last_error("attention: sequence length %d does not fit the resident-KV kernel (%zu B smem)", N, smem);
    return kUnsupported;
  }
  auto kernel = vit_attention_bf16_kernel<NW, TPW>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs))
    TFIMM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  dim3 grid((N + ROWS - 1) / ROWS, H, B);
  kernel<<<grid, NW * 32, smem, stream>>>(qkv, out, N, H, scale * 1.4426950408889634f);
  TFIMM_LAUNCH_OK("vit_attention_bf16_kernel");
  return kOk;
}

// ---- fp32 reference-precision path: one warp per (b, h, query) ----
// Also serves Swin windows: optional additive bias[h, n, n] and mask[w % nmask, n, n]
// (tfimm/architectures/swin.py:172-194), where "b" enumerates windows.
__global__ void attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                     const float* __restrict__ bias, const float* __restrict__ mask,
                                     int nmask, long total_rows, int N, int H, int dh, float scale,
                                     float* __restrict__ probs, const int* __restrict__ row_map,
                                     int nw_img) {
  extern __shared__ float sh[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = sh + (size_t)warp * (N + dh);
  float* qs = sc + N;
  const long rid = (long)blockIdx.x * warps + warp;
  if (rid >= total_rows) return;
  const int n = (int)(rid % N);
  const long bh = rid / N;
  const int h = (int)(bh % H);
  const long b = bh / H;
  const long ld = 3L * H * dh;
  // Token (b, j) lives in row b*N + j, or -- for Swin windows -- wherever the roll + window-partition
  // permutation put it: image (b / nw_img), token row_map[(b % nw_img) * N + j] (swin.py:299-303).
  const long row_base = row_map != nullptr ? (b / nw_img) * ((long)nw_img * N) : b * N;
  const int* rmap = row_map != nullptr ? row_map + (b % nw_img) * (long)N : nullptr;
  auto grow = [&](int j) -> long { return row_base + (rmap != nullptr ? rmap[j] : j); };
  const float* base = qkv + (long)h * dh;
  for (int d = lane; d < dh; d += 32) qs[d] = base[grow(n) * ld + d] * scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    const float* kr = base + grow(j) * ld + (long)H * dh;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(qs[d], kr[d], acc);
    if (bias != nullptr) acc += bias[((long)h * N + n) * N + j];
    if (mask != nullptr) acc += mask[((b % nmask) * N + n) * (long)N + j];
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncwarp();
  if (probs != nullptr)
    for (int j = lane; j < N; j += 32) probs[((bh * N) + n) * (long)N + j] = sc[j] * inv;
  for (int d = lane; d < dh; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(sc[j], base[grow(j) * ld + 2L * H * dh + d], acc);
    out[grow(n) * ((long)H * dh) + (long)h * dh + d] = acc * inv;
  }
}

}  // namespace

int attention_bf16_tc2(const void* qkv, void* out, int B, int N, int H, float scale, cudaStream_t stream);

int attention_bf16(const void* qkv, void* out, int B, int N, int H, int dh, float scale,
                   cudaStream_t stream) {
  TFIMM_CHECK_ARG(B > 0 && N > 0 && H > 0, "attention: bad shape B=%d N=%d H=%d", B, N, H);
  if (dh != kDH) {
    set_last_error("attention: bf16 kernel supports head_dim 64 only (got %d)", dh);
    return kUnsupported;
  }
  TFIMM_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
                  "attention: pointers must be 16-byte aligned");
  // Short sequences (ViT-B/16 @224: N = 197) run on tcgen05 (attention_sm100.cu); longer ones on the resident-KV
  // mma.sync kernel below.
  if (N <= 256) return attention_bf16_tc2(qkv, out, B, N, H, scale, stream);
  auto q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  // resident K/V + one query tile per CTA must fit 227 KB: 224-row tiles up to N = 784, 128-row tiles up to N = 832
  // (vit_base_patch8_224 has N = 785); longer sequences are kUnsupported (the host falls back to the fp32 kernel)
  if (N <= 128 || N > 784) return launch_vit_attention<4, 2>(q, o, B, N, H, scale, stream);
  return launch_vit_attention<7, 2>(q, o, B, N, H, scale, stream);
}

int attention_f32(const float* qkv, float* out, const float* bias, const float* mask, int nmask, long B,
                  int N, int H, int dh, float scale, float* probs, const int* row_map, int nw_img,
                  cudaStream_t stream) {
  TFIMM_CHECK_ARG(row_map == nullptr || (nw_img > 0 && B % nw_img == 0), "attention_f32: bad window map");
  TFIMM_CHECK_ARG(B > 0 && N > 0 && H > 0 && dh > 0, "attention_f32: bad shape");
  const int warps = 4;
  const long total = B * H * N;
  const size_t smem = (size_t)warps * (N + dh) * sizeof(float);
  if (smem > 48 * 1024) {
    set_last_error("attention_f32: sequence too long for the fp32 parity kernel (N=%d)", N);
    return kUnsupported;
  }
  const unsigned grid = (unsigned)((total + warps - 1) / warps);
  attention_f32_kernel<<<grid, warps * 32, smem, stream>>>(qkv, out, bias, mask, nmask > 0 ? nmask : 1, total,
                                                          N, H, dh, scale, probs, row_map, nw_img > 0 ? nw_img : 1);
  TFIMM_LAUNCH_OK("attention_f32_kernel");
  return kOk;
}

}  // namespace tfimm
