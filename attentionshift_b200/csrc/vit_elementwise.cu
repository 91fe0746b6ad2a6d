// Memory-bound pieces of the ViT block: LayerNorm (fp32 in -> fp16 GEMM operand out), patch im2col, token assembly.
// Reference: models/vision_transformer.py:110,114 (norm1/norm2, eps 1e-6 at :146), PatchEmbed conv (:136) as a GEMM,
// visual_transformer_det.py:192-214 (cls / position table / point tokens).
#include "common.cuh"

using namespace asb;

namespace {

// one warp per row; row cached in registers (C <= 32*4*MAXV)
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_f16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __half* __restrict__ y, int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
  const int nv = C / 4;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    v[i] = c < nv ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      yr[c] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

// img [B,3,H,W] fp32 -> cols [B*hp*wp, 3*16*16] fp16, column = c*256 + ky*16 + kx (= conv weight flattened)
__global__ void im2col16_kernel(const float* __restrict__ img, __half* __restrict__ cols, int B, int H, int W) {
  const int hp = H / 16, wp = W / 16;
  const size_t total = (size_t)B * hp * wp * 768 / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col4 = (int)(i % 192);           // 4 consecutive kx
    const size_t patch = i / 192;
    const int c = col4 / 64, ky = (col4 % 64) / 4, kx = (col4 % 4) * 4;
    const int px = (int)(patch % wp), py = (int)((patch / wp) % hp), b = (int)(patch / ((size_t)wp * hp));
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * H + py * 16 + ky) * W + px * 16 + kx);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(cols + patch * 768 + col4 * 4) =
        make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  }
}

// x[b,0] = cls + pos[0]; x[b,1+n] = emb[b,n] + pos[1+n]; x[b,1+N+j] = ptok[j]   (VTD:203-213)
__global__ void assemble_tokens_kernel(const float* __restrict__ emb, const float* __restrict__ cls,
                                       const float* __restrict__ pos, const float* __restrict__ ptok,
                                       float* __restrict__ x, int B, int N, int Tp, int C) {
  const int T = 1 + N + Tp;
  const int c4n = C / 4;
  const size_t total = (size_t)B * T * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const int t = (int)((i / c4n) % T);
    const int b = (int)(i / ((size_t)c4n * T));
    float4 v;
    if (t == 0) {
      const float4 a = reinterpret_cast<const float4*>(cls)[c4], p = reinterpret_cast<const float4*>(pos)[c4];
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else if (t <= N) {
      const float4 a = reinterpret_cast<const float4*>(emb + ((size_t)b * N + t - 1) * C)[c4];
      const float4 p = reinterpret_cast<const float4*>(pos + (size_t)t * C)[c4];
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      v = reinterpret_cast<const float4*>(ptok + (size_t)(t - 1 - N) * C)[c4];
    }
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

}  // namespace

extern "C" int as_layernorm_f16(const float* x, const float* gamma, const float* beta, void* y_f16, int M, int C,
                                float eps, cudaStream_t stream) {
  if (C % 4 || C > 32 * 4 * 16) return AS_ERR_BAD_ARG;
  const int rows_per_cta = 8;
  const dim3 grid((M + rows_per_cta - 1) / rows_per_cta);
  const int nv = (C / 4 + 31) / 32;
  if (nv <= 6) layernorm_f16_kernel<6><<<grid, 256, 0, stream>>>(x, gamma, beta, (__half*)y_f16, M, C, eps);
  else if (nv <= 8) layernorm_f16_kernel<8><<<grid, 256, 0, stream>>>(x, gamma, beta, (__half*)y_f16, M, C, eps);
  else layernorm_f16_kernel<16><<<grid, 256, 0, stream>>>(x, gamma, beta, (__half*)y_f16, M, C, eps);
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" int as_patch_im2col_f16(const float* img, void* cols_f16, int B, int H, int W, cudaStream_t stream) {
  if (H % 16 || W % 16) return AS_ERR_BAD_ARG;
  im2col16_kernel<<<148 * 8, 256, 0, stream>>>(img, (__half*)cols_f16, B, H, W);
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" int as_assemble_tokens(const float* emb, const float* cls, const float* pos, const float* ptok, float* x,
                                  int B, int N, int Tp, int C, cudaStream_t stream) {
  if (C % 4) return AS_ERR_BAD_ARG;
  assemble_tokens_kernel<<<148 * 8, 256, 0, stream>>>(emb, cls, pos, ptok, x, B, N, Tp, C);
  AS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ batched transpose with zero padding (training path)
// src [batch, R, C] f16 (rows contiguous) -> dst [batch, C, Rp] f16, dst[b, c, r] = src[b, r, c] for r < R and 0 for R <= r < Rp.
// The backward GEMMs and the attention backward take K-major operands: dW = dY^T X needs dY^T and X^T with the token dimension
// contiguous (and padded to the GEMM's K granule), the attention backward needs Q^T, K^T, dO^T ([64, Tpad] per head).  A
// strided torch copy does this at ~0.3 TB/s; 64 x 64 tiles through shared memory keep both sides coalesced.
namespace {
__global__ void __launch_bounds__(256) transpose_pad_f16_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int R, int C,
                                                               int Rp) {
  __shared__ __half tile[64][66];
  const int b = blockIdx.z, r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const __half* s = src + (size_t)b * R * C;
  __half* d = dst + (size_t)b * C * Rp;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    __half2 v = __floats2half2_rn(0.f, 0.f);
    if (r < R) {
      if (c + 1 < C) v = *reinterpret_cast<const __half2*>(s + (size_t)r * C + c);
      else if (c < C) v = __halves2half2(s[(size_t)r * C + c], __float2half(0.f));
    }
    tile[i][2 * tx] = __low2half(v);
    tile[i][2 * tx + 1] = __high2half(v);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;
    if (c < C && r < Rp) {
      const __half2 v = __halves2half2(tile[2 * tx][i], tile[2 * tx + 1][i]);
      if (r + 1 < Rp) *reinterpret_cast<__half2*>(d + (size_t)c * Rp + r) = v;
      else d[(size_t)c * Rp + r] = __low2half(v);
    }
  }
}
}  // namespace

extern "C" int as_transpose_pad_f16(const void* src, void* dst, int batch, int R, int C, int Rp, cudaStream_t stream) {
  if (batch < 1 || R < 1 || C < 1 || Rp < R || (C & 1) || (Rp & 1)) return AS_ERR_BAD_ARG;
  const dim3 grid((Rp + 63) / 64, (C + 63) / 64, batch);
  transpose_pad_f16_kernel<<<grid, 256, 0, stream>>>((const __half*)src, (__half*)dst, R, C, Rp);
  AS_LAUNCH_CHECK();
  return 0;
}
