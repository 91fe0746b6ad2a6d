// Small-sequence multi-head attention for the RoI decoders that reuse the ViT Block (MAEBoxHeadRec /
// MAEMaskHeadPointSup run `Block(dim=256, num_heads=8)` -- head_dim 32 -- on 50 / 197 tokens per RoI:
// bbox_heads/mae_bbox_head_rec.py:148-167, mask_heads/mae_mask_head_pointSup.py:172-190; the op is VT:74-83).
// At T <= 256 a whole head fits on chip and the problem is a few MFLOP per (RoI, head): no tensor-core tiling pays here.  One
// CTA per (head, batch item): K and V of the head sit in shared memory, every warp owns query rows; scores with lanes over the
// keys (fp32 FMA), softmax in registers (warp max / sum), P V with lanes over the head dimension.  Thousands of independent
// CTAs fill the machine.  fp16 in / out, fp32 arithmetic.
#include "common.cuh"

using namespace asb;

namespace {

constexpr int SM_MAXT = 256;
constexpr int SM_THREADS = 256;

// qkv [B, T, 3, heads, D] fp16 (the layout the qkv Linear writes, VT:76) -> o [B, T, heads * D] fp16
template <int D>
__global__ void __launch_bounds__(SM_THREADS)
mhsa_small_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, int T, int heads, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  __half* k_s = reinterpret_cast<__half*>(smem_raw);             // [T][D + 2]  (+2 halves: odd word stride, conflict-free rows)
  __half* v_s = k_s + (size_t)T * (D + 2);                       // [T][D]
  float* q_s = reinterpret_cast<float*>(v_s + (size_t)T * D);    // [warps][D]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = SM_THREADS / 32;
  const size_t row_stride = (size_t)3 * heads * D;
  const __half* base = qkv + (size_t)b * T * row_stride + (size_t)h * D;
  for (int i = threadIdx.x; i < T * (D / 2); i += SM_THREADS) {
    const int t = i / (D / 2), c = (i - t * (D / 2)) * 2;
    const __half2 kv = *reinterpret_cast<const __half2*>(base + t * row_stride + (size_t)heads * D + c);
    const __half2 vv = *reinterpret_cast<const __half2*>(base + t * row_stride + (size_t)2 * heads * D + c);
    *reinterpret_cast<__half2*>(k_s + (size_t)t * (D + 2) + c) = kv;
    *reinterpret_cast<__half2*>(v_s + (size_t)t * D + c) = vv;
  }
  __syncthreads();
  float* qw = q_s + warp * D;
  constexpr int KPL = SM_MAXT / 32;                              // keys per lane
  for (int r = warp; r < T; r += nw) {
    for (int c = lane; c < D; c += 32) qw[c] = __half2float(base[r * row_stride + c]);
    __syncwarp();
    float s[KPL];
    float mx = -3.0e38f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int j = lane + 32 * i;
      float acc = 0.f;
      if (j < T) {
        const __half2* kr = reinterpret_cast<const __half2*>(k_s + (size_t)j * (D + 2));
#pragma unroll
        for (int c = 0; c < D / 2; ++c) {
          const float2 kk = __half22float2(kr[c]);
          acc = fmaf(qw[2 * c], kk.x, acc);
          acc = fmaf(qw[2 * c + 1], kk.y, acc);
        }
        acc *= scale_log2;
        mx = fmaxf(mx, acc);
      }
      s[i] = acc;
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int j = lane + 32 * i;
      s[i] = j < T ? ex2_approx(s[i] - mx) : 0.f;
      sum += s[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    // P V: lanes over the head dimension (D = 32: one column per lane; D = 64: two)
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int jn = min(32, T - 32 * i);
      for (int jj = 0; jj < jn; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, s[i], jj);
        const __half* vr = v_s + (size_t)(32 * i + jj) * D;
        acc0 = fmaf(pj, __half2float(vr[lane]), acc0);
        if (D == 64) acc1 = fmaf(pj, __half2float(vr[lane + 32]), acc1);
      }
    }
    __half* orow = o + ((size_t)b * T + r) * heads * D + (size_t)h * D;
    orow[lane] = __float2half_rn(acc0 * inv);
    if (D == 64) orow[lane + 32] = __float2half_rn(acc1 * inv);
    __syncwarp();
  }
}

}  // namespace

// VT:79-83 for short sequences and head_dim 32 or 64: qkv [B, T, 3, heads, head_dim] f16 (the qkv Linear's output), T <= 256
// -> o [B, T, heads * head_dim] f16.  scale = head_dim^-0.5 (VT:67).
extern "C" int as_mhsa_small(const void* qkv, void* o, int B, int T, int heads, int head_dim, cudaStream_t stream) {
  if (T < 1 || T > SM_MAXT || (head_dim != 32 && head_dim != 64) || B < 1 || heads < 1) return AS_ERR_BAD_ARG;
  const float scale_log2 = (float)(1.4426950408889634 / sqrt((double)head_dim));
  const size_t smem = (size_t)T * (head_dim + 2) * 2 + (size_t)T * head_dim * 2 + (SM_THREADS / 32) * head_dim * 4;
  const dim3 grid(heads, B);
  if (head_dim == 32) {
    AS_CUDA(cudaFuncSetAttribute(mhsa_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mhsa_small_kernel<32><<<grid, SM_THREADS, smem, stream>>>((const __half*)qkv, (__half*)o, T, heads, scale_log2);
  } else {
    AS_CUDA(cudaFuncSetAttribute(mhsa_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mhsa_small_kernel<64><<<grid, SM_THREADS, smem, stream>>>((const __half*)qkv, (__half*)o, T, heads, scale_log2);
  }
  AS_LAUNCH_CHECK();
  return 0;
}
