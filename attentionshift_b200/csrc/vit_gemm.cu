// Linear layers of the ViT block on 5th-gen tensor cores:
//   Y[M,N] = X[M,K] * W[N,K]^T (+ bias) with a fused epilogue.
// Reference ops replaced: nn.Linear qkv (models/vision_transformer.py:76), proj (:84), Mlp fc1/GELU/fc2 (:40-59)
// plus the residual adds of Block.forward (:110-115).
//
// Structure (one CTA per SM, persistent over 128x256 output tiles):
//   warp 0      : TMA producer  -- X and W tiles (64 halves = 128 B rows, SWIZZLE_128B) into a 4-stage smem ring
//   warp 1      : MMA issuer    -- tcgen05.mma.kind::f16 (M128 N256 K16), fp32 accumulators in TMEM, 2 accumulator
//                                  buffers (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of i+1
//   warps 2..5  : epilogue      -- tcgen05.ld 32x32b, bias / GELU / residual / QKV head split, vector stores
#include "common.cuh"

using namespace asb;

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;

enum { EPI_F16 = 0, EPI_GELU_F16 = 1, EPI_RESID_F32 = 2, EPI_QKV = 3, EPI_F32 = 4 };

struct GemmParams {
  int M, N, K, epi;
  const float* bias;
  void* out;
  const float* resid;
  __half* q;
  __half* k;
  __half* vt;
  int T, Tpad, heads;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void store_f16x32(__half* dst, const float* f) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h0 = __floats2half2_rn(f[8 * i + 0], f[8 * i + 1]);
    __half2 h1 = __floats2half2_rn(f[8 * i + 2], f[8 * i + 3]);
    __half2 h2 = __floats2half2_rn(f[8 * i + 4], f[8 * i + 5]);
    __half2 h3 = __floats2half2_rn(f[8 * i + 6], f[8 * i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    d4[i] = u;
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  const int kblocks = p.K / BK;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
          tma_load_2d(smem_a + stage * A_BYTES, &tm_a, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * B_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(0, BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_base = smem_u32(smem_a + stage * A_BYTES);
          const uint32_t b_base = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            mma_f16_ss(tmem_base + acc * BN, umma_desc_k_sw128(a_base + k * 32), umma_desc_k_sw128(b_base + k * 32),
                       idesc, (kb | k) != 0);
          }
          tc_commit(&empty[stage]);
          if (kb == kblocks - 1) tc_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int quad = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const int C = p.heads * 64;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int m = m_blk * BM + quad * 32 + lane;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      int b_idx = 0, t_idx = 0;
      if (p.epi == EPI_QKV && m < p.M) { b_idx = m / p.T; t_idx = m - b_idx * p.T; }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, v);
        tc_wait_ld();
        const int n0 = n_blk * BN + c * 32;
        if (m < p.M && n0 < p.N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + (p.bias ? __ldg(p.bias + n0 + i) : 0.f);
          if (p.epi == EPI_F16) {
            store_f16x32(reinterpret_cast<__half*>(p.out) + (size_t)m * p.N + n0, f);
          } else if (p.epi == EPI_GELU_F16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
            store_f16x32(reinterpret_cast<__half*>(p.out) + (size_t)m * p.N + n0, f);
          } else if (p.epi == EPI_RESID_F32 || p.epi == EPI_F32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)m * p.N + n0);
            if (p.epi == EPI_RESID_F32) {
              const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)m * p.N + n0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 r = r4[i];
                o4[i] = make_float4(r.x + f[4 * i], r.y + f[4 * i + 1], r.z + f[4 * i + 2], r.w + f[4 * i + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
          } else {  // EPI_QKV: scatter to Q [B,h,T,64], K [B,h,T,64], V^T [B,h,64,Tpad]
            const int which = n0 / C;
            const int cc = n0 - which * C;
            const int h = cc >> 6, d0 = cc & 63;
            const size_t bh = (size_t)b_idx * p.heads + h;
            if (which < 2) {
              __half* dst = (which == 0 ? p.q : p.k) + (bh * p.T + t_idx) * 64 + d0;
              store_f16x32(dst, f);
            } else {
              __half* dst = p.vt + (bh * 64 + d0) * (size_t)p.Tpad + t_idx;
#pragma unroll
              for (int i = 0; i < 32; ++i) dst[(size_t)i * p.Tpad] = __float2half_rn(f[i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int launch_linear(const void* x, const void* w, const GemmParams& p, cudaStream_t stream) {
  if (p.K % BK != 0 || p.M <= 0 || p.N <= 0 || (p.N % 32) != 0) return AS_ERR_BAD_ARG;
  CUtensorMap tm_a, tm_b;
  uint64_t dims_a[2] = {(uint64_t)p.K, (uint64_t)p.M}, str_a[1] = {(uint64_t)p.K * 2};
  uint32_t box_a[2] = {BK, BM};
  uint64_t dims_b[2] = {(uint64_t)p.K, (uint64_t)p.N}, str_b[1] = {(uint64_t)p.K * 2};
  uint32_t box_b[2] = {BK, BN};
  int r = as_encode_tmap(&tm_a, x, 2, 2, dims_a, str_a, box_a);
  if (r) return r;
  r = as_encode_tmap(&tm_b, w, 2, 2, dims_b, str_b, box_b);
  if (r) return r;
  static int num_sms = 0;
  if (!num_sms) {
    int dev;
    AS_CUDA(cudaGetDevice(&dev));
    AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    AS_CUDA(cudaFuncSetAttribute(linear_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  linear_tcgen05_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tm_a, tm_b, p);
  AS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// mode: 0 = fp16 out, 1 = GELU -> fp16 out, 2 = fp32 out = resid + y, 4 = fp32 out
extern "C" int as_linear_f16(const void* x_f16, const void* w_f16, const float* bias, void* out, const float* resid,
                             int M, int N, int K, int mode, cudaStream_t stream) {
  if (mode != EPI_F16 && mode != EPI_GELU_F16 && mode != EPI_RESID_F32 && mode != EPI_F32) return AS_ERR_BAD_ARG;
  if (mode == EPI_RESID_F32 && !resid) return AS_ERR_BAD_ARG;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.epi = mode; p.bias = bias; p.out = out; p.resid = resid; p.heads = 1;
  return launch_linear(x_f16, w_f16, p, stream);
}

// x [B*T, C] fp16, w [3C, C] fp16, bias [3C] -> q,k [B,h,T,64] fp16, vt [B,h,64,Tpad] fp16 (V transposed, K-major for P*V)
extern "C" int as_qkv_proj_f16(const void* x_f16, const void* w_f16, const float* bias, void* q, void* k, void* vt,
                               int B, int T, int Tpad, int heads, cudaStream_t stream) {
  GemmParams p{};
  const int C = heads * 64;
  p.M = B * T; p.N = 3 * C; p.K = C; p.epi = EPI_QKV; p.bias = bias;
  p.q = (__half*)q; p.k = (__half*)k; p.vt = (__half*)vt; p.T = T; p.Tpad = Tpad; p.heads = heads;
  if (Tpad < T || (Tpad % 8) != 0) return AS_ERR_BAD_ARG;
  return launch_linear(x_f16, w_f16, p, stream);
}
