// Attention roll-out restricted to the rows the attention-shift head reads.
// Reference: attns_project_to_feature (RH:1257-1272) builds aug = (A + I) / rowsum(A + I) for the last L layers and
// the full T x T products joint[-i] = joint[-(i-1)] @ aug[-i]; its only consumers take rows [-n_point_tokens:]
// (RH:2272).  Row r of a product depends only on row r of the left factor, so we carry an [n_rows x T] slab:
//   R_0 = aug_last[-n_rows:],   R_i = R_{i-1} @ aug_{last-i}
// with  R @ aug = R' @ A + R',  R'[r,k] = R[r,k] / rs[k],  rs[k] = rowsum(A)[k] + 1   (A = head-mean attention).
#include "common.cuh"
#include <cuda_fp16.h>

using namespace asb;

namespace {

constexpr int RM = 104;      // slab rows handled per CTA (>= n_rows, 8 groups of 13)
constexpr int RN = 128;      // output columns per CTA
constexpr int RK = 32;       // k chunk

// rs[b,k] = sum of the head-mean kernel's per-tile partial row sums (fixed order) + 1
__global__ void rollout_rowsum(const float* __restrict__ part, int ntile, int total, float* __restrict__ rs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int t = 0; t < ntile; ++t) s += part[(size_t)i * ntile + t];
  rs[i] = s + 1.f;
}

// R_0[b,r,n] = (A[b, T-n_rows+r, n] + [n == T-n_rows+r]) / rs[b, T-n_rows+r]
__global__ void rollout_first(const float* __restrict__ A, int ld, const float* __restrict__ rs, int T, int n_rows,
                              float* __restrict__ out, size_t out_bstride) {
  const int b = blockIdx.z, r = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= T) return;
  const int row = T - n_rows + r;
  const float v = A[((size_t)b * T + row) * ld + n] + (n == row ? 1.f : 0.f);
  out[b * out_bstride + (size_t)r * T + n] = v / rs[(size_t)b * T + row];
}

// Rp[b,r,k] = R[b,r,k] / rs[b,k]
__global__ void rollout_scale(const float* __restrict__ R, size_t r_bstride, const float* __restrict__ rs, int T,
                              int n_rows, float* __restrict__ Rp) {
  const int b = blockIdx.z, r = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  Rp[((size_t)b * n_rows + r) * T + k] = R[b * r_bstride + (size_t)r * T + k] / rs[(size_t)b * T + k];
}

// out[b,r,n] = sum_k Rp[b,r,k] * A[b,k,n] + Rp[b,r,n];   grid (ceil(T/128), B), 256 threads (8 row groups x 32 col groups)
__global__ void __launch_bounds__(256)
rollout_gemm(const float* __restrict__ Rp, const float* __restrict__ A, int ld, int T, int n_rows,
             float* __restrict__ out, size_t out_bstride) {
  __shared__ float a_s[RK][RM];                    // [k][row]
  __shared__ __align__(16) float b_s[RK][RN];      // [k][col]
  const int b = blockIdx.y, n0 = blockIdx.x * RN;
  const int rg = threadIdx.x >> 5, cg = threadIdx.x & 31;
  const float* Rb = Rp + (size_t)b * n_rows * T;
  const float* Ab = A + (size_t)b * T * ld;
  float acc[13][4];
#pragma unroll
  for (int i = 0; i < 13; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < T; k0 += RK) {
    for (int i = threadIdx.x; i < RM * RK; i += 256) {
      const int r = i / RK, k = i - r * RK;
      a_s[k][r] = (r < n_rows && k0 + k < T) ? Rb[(size_t)r * T + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < RK * RN; i += 256) {
      const int k = i / RN, c = i - k * RN;
      b_s[k][c] = (k0 + k < T && n0 + c < T) ? Ab[(size_t)(k0 + k) * ld + n0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < RK; ++k) {
      const float4 bv = *reinterpret_cast<const float4*>(&b_s[k][cg * 4]);
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        const float a = a_s[k][rg * 13 + i];
        acc[i][0] = fmaf(a, bv.x, acc[i][0]);
        acc[i][1] = fmaf(a, bv.y, acc[i][1]);
        acc[i][2] = fmaf(a, bv.z, acc[i][2]);
        acc[i][3] = fmaf(a, bv.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const int r = rg * 13 + i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + cg * 4 + j;
      if (n < T) out[b * out_bstride + (size_t)r * T + n] = acc[i][j] + Rb[(size_t)r * T + n];
    }
  }
}

// tensor-core path helpers: the slab R' as split-fp16 A operand [B,128,ldk] (zero padded) and R' itself (the identity
// term of aug) written straight into the output slab, where the three GEMMs then accumulate.
__global__ void rollout_scale_split(const float* __restrict__ R, size_t r_bstride, int ldr, const float* __restrict__ rs, int T,
                                    int n_rows, int ldk, float scale, __half* __restrict__ hi, __half* __restrict__ lo,
                                    float* __restrict__ out, size_t out_bstride, int ldo) {
  const int b = blockIdx.z, r = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ldk) return;
  float v = 0.f;
  if (r < n_rows && k < T) v = R[b * r_bstride + (size_t)r * ldr + k] / rs[(size_t)b * T + k];
  const float sv = v * scale;
  const __half h = __float2half_rn(sv);
  hi[((size_t)b * 128 + r) * ldk + k] = h;
  lo[((size_t)b * 128 + r) * ldk + k] = __float2half_rn(sv - __half2float(h));
  if (r < n_rows && k < ldo) out[b * out_bstride + (size_t)r * ldo + k] = v;
}
__global__ void rollout_first_ld(const float* __restrict__ A, int ld, const float* __restrict__ rs, int T, int n_rows,
                                 float* __restrict__ out, size_t out_bstride, int ldo) {
  const int b = blockIdx.z, r = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= ldo) return;
  const int row = T - n_rows + r;
  float v = 0.f;
  if (n < T) v = (A[((size_t)b * T + row) * ld + n] + (n == row ? 1.f : 0.f)) / rs[(size_t)b * T + row];
  out[b * out_bstride + (size_t)r * ldo + n] = v;
}

}  // namespace

extern "C" int as_bgemm_f16_f32(const void* x_f16, const void* w_f16, float* out, const float* resid, int batch, int M,
                                int N, int K, int x_rows, int w_rows, int ldo, long long out_bstride, float alpha,
                                cudaStream_t stream);

extern "C" size_t as_rollout_tc_workspace(int B, int T, int ldt) {
  return ((size_t)B * T * 4 + 255) / 256 * 256 + (size_t)2 * B * 128 * ldt * 2;
}

// Tensor-core roll-out.  attn[l] [B,T,ld] f32 (only the LAST layer's map is read: its last n_rows rows), t_hi[l] / t_lo[l]
// [B,ldt,ldt] split-fp16 transposed maps scaled by t_scale (as written by as_attn_headmean), rowsum_part as above.
// out [B, L, n_rows, ldt] f32 (row stride ldt; columns >= T are zero).
extern "C" int as_rollout_rows_tc(const float* const* attn, const void* const* t_hi, const void* const* t_lo,
                                  const float* const* rowsum_part, int L, int B, int T, int ld, int ldt, float t_scale,
                                  int ntile, int n_rows, float* out, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  if (n_rows > 128 || n_rows > T || L < 1 || ldt % 64 || ldt < T) return AS_ERR_BAD_ARG;
  if (workspace_bytes < as_rollout_tc_workspace(B, T, ldt)) return AS_ERR_BAD_ARG;
  float* rs = (float*)workspace;
  __half* hi = (__half*)((char*)workspace + ((size_t)B * T * 4 + 255) / 256 * 256);
  __half* lo = hi + (size_t)B * 128 * ldt;
  const size_t bstride = (size_t)L * n_rows * ldt;
  const float r_scale = 4096.f;                                   // 2^12: keeps the lo halves out of the fp16 subnormals
  const float alpha = 1.f / (r_scale * t_scale);
  for (int i = 0; i < L; ++i) {
    const int l = L - 1 - i;
    rollout_rowsum<<<(B * T + 255) / 256, 256, 0, stream>>>(rowsum_part[l], ntile, B * T, rs);
    float* dst = out + (size_t)i * n_rows * ldt;
    if (i == 0) {
      rollout_first_ld<<<dim3((ldt + 255) / 256, n_rows, B), 256, 0, stream>>>(attn[l], ld, rs, T, n_rows, dst, bstride, ldt);
    } else {
      rollout_scale_split<<<dim3((ldt + 255) / 256, 128, B), 256, 0, stream>>>(out + (size_t)(i - 1) * n_rows * ldt, bstride, ldt, rs, T,
                                                                             n_rows, ldt, r_scale, hi, lo, dst, bstride, ldt);
      int r = as_bgemm_f16_f32(hi, t_hi[l], dst, dst, B, n_rows, ldt, ldt, 128, ldt, ldt, (long long)bstride, alpha, stream);
      if (r) return r;
      r = as_bgemm_f16_f32(hi, t_lo[l], dst, dst, B, n_rows, ldt, ldt, 128, ldt, ldt, (long long)bstride, alpha, stream);
      if (r) return r;
      r = as_bgemm_f16_f32(lo, t_hi[l], dst, dst, B, n_rows, ldt, ldt, 128, ldt, ldt, (long long)bstride, alpha, stream);
      if (r) return r;
    }
  }
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t as_rollout_workspace(int B, int T, int n_rows) {
  return ((size_t)B * T * 4 + 255) / 256 * 256 + (size_t)B * n_rows * T * 4;
}

// attn[l], rowsum_part[l] : host arrays of L device pointers, oldest layer first ([B,T,ld] and [B,T,ntile]).
// out [B, L, n_rows, T]: index 0 = last layer alone ... L-1 = product over all L layers (RH:1268-1271 order).
extern "C" int as_rollout_rows(const float* const* attn, const float* const* rowsum_part, int L, int B, int T, int ld,
                               int ntile, int n_rows, float* out, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
  if (n_rows > RM || n_rows > T || L < 1) return AS_ERR_BAD_ARG;
  if (workspace_bytes < as_rollout_workspace(B, T, n_rows)) return AS_ERR_BAD_ARG;
  float* rs = (float*)workspace;
  float* Rp = (float*)((char*)workspace + ((size_t)B * T * 4 + 255) / 256 * 256);
  const size_t bstride = (size_t)L * n_rows * T;
  const dim3 ge((T + 255) / 256, n_rows, B);
  for (int i = 0; i < L; ++i) {
    const int l = L - 1 - i;
    rollout_rowsum<<<(B * T + 255) / 256, 256, 0, stream>>>(rowsum_part[l], ntile, B * T, rs);
    float* dst = out + (size_t)i * n_rows * T;
    if (i == 0) {
      rollout_first<<<ge, 256, 0, stream>>>(attn[l], ld, rs, T, n_rows, dst, bstride);
    } else {
      rollout_scale<<<ge, 256, 0, stream>>>(out + (size_t)(i - 1) * n_rows * T, bstride, rs, T, n_rows, Rp);
      rollout_gemm<<<dim3((T + RN - 1) / RN, B), 256, 0, stream>>>(Rp, attn[l], ld, T, n_rows, dst, bstride);
    }
  }
  AS_LAUNCH_CHECK();
  return 0;
}
