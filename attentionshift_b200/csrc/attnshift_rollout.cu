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

// ------------------------------------------------------------------ fused split-fp16 slab GEMM on tcgen05
// out[b, r, n] += alpha * sum_k (Ahi + Alo)[b, r, k] * (Bhi + Blo)[b, n, k]   (the lo.lo term is below fp32 resolution)
// One CTA = one image x 64 output columns, all 128 slab rows, the whole K range: the T x T operand (570 MB per layer as a
// hi / lo pair) is read from HBM exactly ONCE -- three separate GEMM launches read its hi half twice and its lo half once.
constexpr int RT_BN = 64, RT_BK = 64, RT_STAGES = 4;
constexpr int RT_A = 128 * RT_BK * 2, RT_B = RT_BN * RT_BK * 2;            // 16 KB, 8 KB
constexpr int RT_STAGE = 2 * RT_A + 2 * RT_B;                             // 48 KB: A hi, A lo, B hi, B lo
constexpr int RT_SMEM = RT_STAGES * RT_STAGE + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
rollout_mma_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                   const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, int kblocks,
                   int n_rows, int ldo, size_t out_bstride, float alpha, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RT_STAGES * RT_STAGE);
  uint64_t* empty = full + RT_STAGES;
  uint64_t* acc_full = empty + RT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * RT_BN, b = blockIdx.y;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_ahi); tma_prefetch_desc(&tm_alo); tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo);
    for (int i = 0; i < RT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % RT_STAGES;
        mbar_wait(&empty[st], ((kb / RT_STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], RT_STAGE);
        uint8_t* d = smem + st * RT_STAGE;
        tma_load_3d(d, &tm_ahi, &full[st], kb * RT_BK, 0, b);
        tma_load_3d(d + RT_A, &tm_alo, &full[st], kb * RT_BK, 0, b);
        tma_load_3d(d + 2 * RT_A, &tm_bhi, &full[st], kb * RT_BK, n0, b);
        tma_load_3d(d + 2 * RT_A + RT_B, &tm_blo, &full[st], kb * RT_BK, n0, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(0, 128, RT_BN);
    for (int kb = 0; kb < kblocks; ++kb) {
      const int st = kb % RT_STAGES;
      mbar_wait(&full[st], (kb / RT_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + st * RT_STAGE), a_lo = a_hi + RT_A, b_hi = a_hi + 2 * RT_A, b_lo = b_hi + RT_B;
#pragma unroll
        for (int k = 0; k < RT_BK / 16; ++k) {
          mma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, (kb | k) != 0);
          mma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_lo + k * 32), idesc, 1);
          mma_f16_ss(tmem, umma_desc_k_sw128(a_lo + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, 1);
        }
        tc_commit(&empty[st]);
        if (kb == kblocks - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3, r = quad * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < RT_BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + c * 32, v);
      tc_wait_ld();
      if (r < n_rows) {
        float4* o4 = reinterpret_cast<float4*>(out + b * out_bstride + (size_t)r * ldo + n0 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = o4[i];                                   // the identity term R' written by rollout_scale_split
          o.x = fmaf(__uint_as_float(v[4 * i]), alpha, o.x); o.y = fmaf(__uint_as_float(v[4 * i + 1]), alpha, o.y);
          o.z = fmaf(__uint_as_float(v[4 * i + 2]), alpha, o.z); o.w = fmaf(__uint_as_float(v[4 * i + 3]), alpha, o.w);
          o4[i] = o;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

int rollout_mma(const __half* a_hi, const __half* a_lo, const void* b_hi, const void* b_lo, float* out, int B, int n_rows,
                int ldt, size_t out_bstride, float alpha, cudaStream_t stream) {
  CUtensorMap tm[4];
  uint64_t da[3] = {(uint64_t)ldt, 128, (uint64_t)B}, sa[2] = {(uint64_t)ldt * 2, (uint64_t)128 * ldt * 2};
  uint32_t ba[3] = {RT_BK, 128, 1};
  uint64_t db[3] = {(uint64_t)ldt, (uint64_t)ldt, (uint64_t)B}, sb[2] = {(uint64_t)ldt * 2, (uint64_t)ldt * ldt * 2};
  uint32_t bb[3] = {RT_BK, RT_BN, 1};
  int r = as_encode_tmap(&tm[0], a_hi, 2, 3, da, sa, ba);
  if (!r) r = as_encode_tmap(&tm[1], a_lo, 2, 3, da, sa, ba);
  if (!r) r = as_encode_tmap(&tm[2], b_hi, 2, 3, db, sb, bb);
  if (!r) r = as_encode_tmap(&tm[3], b_lo, 2, 3, db, sb, bb);
  if (r) return r;
  static bool attr = false;
  if (!attr) {
    AS_CUDA(cudaFuncSetAttribute(rollout_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM));
    attr = true;
  }
  rollout_mma_kernel<<<dim3(ldt / RT_BN, B), 192, RT_SMEM, stream>>>(tm[0], tm[1], tm[2], tm[3], ldt / RT_BK, n_rows, ldt,
                                                                     out_bstride, alpha, out);
  AS_LAUNCH_CHECK();
  return 0;
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
      // hi.hi + hi.lo + lo.hi in ONE pass over the T x T operand (three GEMM launches would read its hi half twice)
      const int r = rollout_mma(hi, lo, t_hi[l], t_lo[l], dst, B, n_rows, ldt, bstride, alpha, stream);
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
