// Element-wise / reduction half of the ViT block's backward ("attention backward + MLP / LayerNorm fusion").
// The reference leaves these to autograd over the unfused block (VT:109-124): one kernel (and one pass over a [tokens, C]
// tensor) per cast, per bias-gradient reduction, per GELU / LayerNorm backward, per residual-gradient add.  Here:
//   as_colsum            bias gradient = column sums of dY (fp32 or fp16), optionally fused with the fp16 cast the next
//                        GEMM wants
//   as_gelu_bwd_f16      d(pre) = dH * gelu'(pre) (exact erf GELU, VT:40 nn.GELU) fused with fc1's bias gradient
//   as_layernorm_bwd     dX = LN-backward(x, gamma, dY) (+ the residual-stream gradient, VT:113-114's `x + f(norm(x))`)
//                        with d gamma / d beta accumulated on the way (statistics recomputed from x: nothing saved)
//   as_attn_bwd_prep     delta = rowsum(dO o O) per head and the head-major copy of dO the attention backward loads
// All HBM-bound: every tensor is read once, every output written once; column partials go through a small [parts, N] buffer
// and a fixed-order second stage (no atomics: the gradients are bit-reproducible).
#include "common.cuh"

#include <cuda_fp16.h>
#include <math.h>

using namespace asb;

namespace {

constexpr int CS_MAX_PARTS = 1184;            // row slices at most (workspace bound)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4h(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// exact GELU derivative: Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// MODE 0: part = colsum(x), optional cast16 = half(x);  MODE 1: g = x * gelu'(aux) -> out16, part = colsum(g)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
colsum_part_kernel(const T* __restrict__ x, const __half* __restrict__ aux, int M, int N, float* __restrict__ part,
                   __half* __restrict__ out16) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto one = [&](float4 v, float4 a, size_t o) {
    if (MODE == 1) {
      v.x *= gelu_grad(a.x); v.y *= gelu_grad(a.y); v.z *= gelu_grad(a.z); v.w *= gelu_grad(a.w);
      // the bias gradient sums what the GEMMs will see: the fp16-rounded values
      const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out16 + o) = u;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      v = make_float4(f0.x, f0.y, f1.x, f1.y);
    } else if (out16) {
      st4h(out16 + o, v);
    }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  };
  // rows blockIdx.y, + gridDim.y, ...: four rows' loads in flight per thread
  int r = blockIdx.y;
  const size_t step = (size_t)gridDim.y * N;
  for (; r + 3 * (int)gridDim.y < M; r += 4 * gridDim.y) {
    const size_t o = (size_t)r * N + c;
    float4 v[4], a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = ld4(x + o + i * step);
      if (MODE == 1) a[i] = ld4(aux + o + i * step);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) one(v[i], a[i], o + i * step);
  }
  for (; r < M; r += gridDim.y) {
    const size_t o = (size_t)r * N + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 1) a = ld4(aux + o);
    one(ld4(x + o), a, o);
  }
  *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * N + c) = acc;
}

// out[c] = sum over parts in a fixed order: thread (tx, ty) adds parts ty, ty + 8, ... of column c, the eight row groups are
// then added in order through shared memory
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ part, int n_part, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f;
  if (c < N) {
    int p = ty;
    for (; p + 8 < n_part; p += 16) { a0 += part[(size_t)p * N + c]; a1 += part[(size_t)(p + 8) * N + c]; }
    if (p < n_part) a0 += part[(size_t)p * N + c];
  }
  red[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && c < N) {
    float a = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) a += red[w][tx];
    out[c] = a;
  }
}

int colsum_parts(int M, int N, int num_sms, dim3* grid, dim3* block) {
  const int tx = (N / 4 + 31) / 32 * 32;
  const int bx = tx < 256 ? tx : 256;
  const int gx = (N / 4 + bx - 1) / bx;
  int gy = (4 * num_sms + gx - 1) / gx;
  if (gy > M) gy = M;
  if (gy > CS_MAX_PARTS) gy = CS_MAX_PARTS;
  if (gy < 1) gy = 1;
  *grid = dim3(gx, gy);
  *block = dim3(bx);
  return gy;
}

// LayerNorm backward, one warp per row, C = 128 * NC columns: lane owns columns lane*4 + 128*i .. +3
template <int NC>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const __half* __restrict__ dy,
                     const float* __restrict__ resid_grad, int M, float eps, float* __restrict__ dx, float* __restrict__ part) {
  constexpr int C = NC * 128;
  __shared__ float red[2 * C];                     // CTA partial of d gamma | d beta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wglob = blockIdx.x * 8 + warp, wtot = gridDim.x * 8;
  float4 g[NC], dg[NC], db[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    g[i] = ld4(gamma + lane * 4 + 128 * i);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = wglob; r < M; r += wtot) {
    const size_t o = (size_t)r * C + lane * 4;
    float4 xv[NC], dv[NC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xv[i] = ld4(x + o + 128 * i);
      dv[i] = ld4(dy + o + 128 * i);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    float s1 = 0.f, s2 = 0.f;                        // sum(dy * gamma), sum(dy * gamma * xhat)
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;              // xhat
      dg[i].x += dv[i].x * xv[i].x; dg[i].y += dv[i].y * xv[i].y; dg[i].z += dv[i].z * xv[i].z; dg[i].w += dv[i].w * xv[i].w;
      db[i].x += dv[i].x; db[i].y += dv[i].y; db[i].z += dv[i].z; db[i].w += dv[i].w;
      dv[i].x *= g[i].x; dv[i].y *= g[i].y; dv[i].z *= g[i].z; dv[i].w *= g[i].w;      // dy * gamma
      s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
      s2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      float4 out;
      out.x = rstd * (dv[i].x - m1 - xv[i].x * m2); out.y = rstd * (dv[i].y - m1 - xv[i].y * m2);
      out.z = rstd * (dv[i].z - m1 - xv[i].z * m2); out.w = rstd * (dv[i].w - m1 - xv[i].w * m2);
      if (resid_grad) {
        const float4 rg = ld4(resid_grad + o + 128 * i);
        out.x += rg.x; out.y += rg.y; out.z += rg.z; out.w += rg.w;
      }
      *reinterpret_cast<float4*>(dx + o + 128 * i) = out;
    }
  }
  // CTA partial of d gamma | d beta: the warps add their registers one after the other (fixed order)
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        float4* pg = reinterpret_cast<float4*>(&red[lane * 4 + 128 * i]);
        float4* pb = reinterpret_cast<float4*>(&red[C + lane * 4 + 128 * i]);
        if (w == 0) { *pg = dg[i]; *pb = db[i]; }
        else {
          float4 a = *pg, b = *pb;
          a.x += dg[i].x; a.y += dg[i].y; a.z += dg[i].z; a.w += dg[i].w;
          b.x += db[i].x; b.y += db[i].y; b.z += db[i].z; b.w += db[i].w;
          *pg = a; *pb = b;
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < 2 * C; c += 256) part[(size_t)blockIdx.x * 2 * C + c] = red[c];
}

// thread = (token, head): delta = sum_d dO * O, dO copied head-major
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __half* __restrict__ d_o, const __half* __restrict__ o, int B, int T, int heads,
                     __half* __restrict__ d_oh, float* __restrict__ delta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * T * heads) return;
  const int h = (int)(idx % heads);
  const long long bt = idx / heads;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + idx * 64);
  const uint4* po = reinterpret_cast<const uint4*>(o + idx * 64);
  uint4* dst = reinterpret_cast<uint4*>(d_oh + (((size_t)b * heads + h) * T + t) * 64);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a = pd[i], c = po[i];
    dst[i] = a;
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* ch = reinterpret_cast<const __half2*>(&c);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]), fc = __half22float2(ch[e]);
      acc = fmaf(fa.x, fc.x, acc);
      acc = fmaf(fa.y, fc.y, acc);
    }
  }
  delta[((size_t)b * heads + h) * T + t] = acc;
}

int sm_count(int* n) {
  int dev;
  AS_CUDA(cudaGetDevice(&dev));
  AS_CUDA(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

}  // namespace

extern "C" size_t as_colsum_workspace(int N) { return (size_t)CS_MAX_PARTS * N * sizeof(float); }

// out [N] f32 = column sums of x [M, N] (x_is_f16: fp16, else fp32); cast16 (fp32 input only, may be null) receives half(x).
extern "C" int as_colsum(const void* x, int x_is_f16, int M, int N, void* cast16, float* out, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
  if (M < 1 || N < 4 || N % 4 || workspace_bytes < as_colsum_workspace(N) || (x_is_f16 && cast16)) return AS_ERR_BAD_ARG;
  int sms;
  if (int r = sm_count(&sms)) return r;
  dim3 grid, block;
  const int parts = colsum_parts(M, N, sms, &grid, &block);
  float* part = (float*)workspace;
  if (x_is_f16) colsum_part_kernel<__half, 0><<<grid, block, 0, stream>>>((const __half*)x, nullptr, M, N, part, nullptr);
  else colsum_part_kernel<float, 0><<<grid, block, 0, stream>>>((const float*)x, nullptr, M, N, part, (__half*)cast16);
  colsum_finish_kernel<<<(N + 31) / 32, 256, 0, stream>>>(part, parts, N, out);
  AS_LAUNCH_CHECK();
  return 0;
}

// d_pre [M, N] fp16 = d_hid * gelu'(pre) (erf GELU); d_bias [N] f32 = column sums of d_pre.
extern "C" int as_gelu_bwd_f16(const void* d_hid, const void* pre, int M, int N, void* d_pre, float* d_bias, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream) {
  if (M < 1 || N < 4 || N % 4 || workspace_bytes < as_colsum_workspace(N)) return AS_ERR_BAD_ARG;
  int sms;
  if (int r = sm_count(&sms)) return r;
  dim3 grid, block;
  const int parts = colsum_parts(M, N, sms, &grid, &block);
  float* part = (float*)workspace;
  colsum_part_kernel<__half, 1><<<grid, block, 0, stream>>>((const __half*)d_hid, (const __half*)pre, M, N, part, (__half*)d_pre);
  colsum_finish_kernel<<<(N + 31) / 32, 256, 0, stream>>>(part, parts, N, d_bias);
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t as_layernorm_bwd_workspace(int C) { return (size_t)CS_MAX_PARTS * 2 * C * sizeof(float); }

// x [M, C] f32 (the LayerNorm input), dy [M, C] fp16 (gradient of the fp16 LayerNorm output), resid_grad [M, C] f32 or null.
// dx [M, C] f32 = LN-backward (+ resid_grad); dgb [2, C] f32 = d gamma | d beta.  C in {128, 256, ..., 1024}.
extern "C" int as_layernorm_bwd(const float* x, const float* gamma, const void* dy, const float* resid_grad, int M, int C,
                                float eps, float* dx, float* dgb, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (M < 1 || C % 128 || C < 128 || C > 1024 || workspace_bytes < as_layernorm_bwd_workspace(C)) return AS_ERR_BAD_ARG;
  int sms;
  if (int r = sm_count(&sms)) return r;
  int grid = 4 * sms;
  if (grid > (M + 7) / 8) grid = (M + 7) / 8;
  if (grid > CS_MAX_PARTS) grid = CS_MAX_PARTS;
  float* part = (float*)workspace;
  const __half* d = (const __half*)dy;
  switch (C / 128) {
#define LN_CASE(NC) case NC: layernorm_bwd_kernel<NC><<<grid, 256, 0, stream>>>(x, gamma, d, resid_grad, M, eps, dx, part); break;
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
  }
  colsum_finish_kernel<<<(2 * C + 31) / 32, 256, 0, stream>>>(part, grid, 2 * C, dgb);
  AS_LAUNCH_CHECK();
  return 0;
}

// d_o, o [B, T, heads*64] fp16 -> d_oh [B, heads, T, 64] fp16 (head-major copy of d_o), delta [B, heads, T] f32.
extern "C" int as_attn_bwd_prep(const void* d_o, const void* o, int B, int T, int heads, void* d_oh, float* delta,
                                cudaStream_t stream) {
  if (B < 1 || T < 1 || heads < 1) return AS_ERR_BAD_ARG;
  const long long n = (long long)B * T * heads;
  attn_bwd_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const __half*)d_o, (const __half*)o, B, T, heads,
                                                                         (__half*)d_oh, delta);
  AS_LAUNCH_CHECK();
  return 0;
}
