// Linear layers of the ViT block on 5th-gen tensor cores:
//   Y[M,N] = X[M,K] * W[N,K]^T (+ bias) with a fused epilogue.
// Reference ops replaced: nn.Linear qkv (models/vision_transformer.py:76), proj (:84), Mlp fc1/GELU/fc2 (:40-59)
// plus the residual adds of Block.forward (:110-115).
//
// Structure (one CTA per SM, persistent; thread-block clusters of 2 CTAs = one CTA pair per 256x256 output tile, each CTA
// owning 128 rows of it):
//   warp 0      : TMA producer  -- X and W tiles (64 halves = 128 B rows, SWIZZLE_128B) into a 6-stage smem ring
//   warp 1      : MMA issuer    -- tcgen05.mma.cta_group::2.kind::f16 (M256 N256 K16 across the pair), fp32 accumulators in
//                                  TMEM, 2 accumulator buffers (2 x 256 columns): the epilogue of tile i overlaps the mainloop of i+1
//   warps 2..9  : epilogue      -- tcgen05.ld 32x32b, bias / GELU / residual / QKV head split; every 32 x 32 chunk crosses a
//                                  private XOR-swizzled smem tile so that the global stores (and residual loads) cover whole
//                                  row segments; the epilogue mode is a template parameter
#include "common.cuh"

using namespace asb;

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int EPI_STAGE_BYTES = 8 * 4096;   // one 32 x 32 fp32 tile per epilogue warp (XOR-swizzled, no padding)
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, 128 columns each)

enum { EPI_F16 = 0, EPI_GELU_F16 = 1, EPI_RESID_F32 = 2, EPI_QKV = 3, EPI_F32 = 4 };

struct GemmParams {
  int M, N, K, epi;
  int batch;                 // independent problems; X rows / W rows / out are strided per batch
  long long out_bstride;     // elements
  long long resid_bstride;   // elements
  int ldo;                   // row stride (elements) of out / resid
  float alpha;               // scales the accumulator before bias / residual
  const float* bias;
  void* out;
  const float* resid;
  __half* q;
  __half* k;
  __half* vt;
  int T, Tpad, heads;
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), nn.GELU's exact form (VT:55).  erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, far below the fp16 rounding of the stored activation); 14 instructions per element -- erff()
// made the fc1 epilogue longer than its mainloop.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752440f;
  float t;   // rcp.approx (1 ulp): __frcp_rn expands to a Newton step plus a slow-path BRANCH per element, which serialised
             // the 32 elements of a chunk (the epilogue took 4x the mainloop)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float z2 = ax * 0.84932180028801907f;               // sqrt(log2(e) / 2): exp(-z^2) = 2^(-z2^2)
  const float e = fmaf(-poly * t, ex2_approx(-z2 * z2), 1.0f);   // erf(|x| / sqrt 2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), e, hx);                            // 0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}

template <int kPair, int kEpi, bool kTN = false>     // kTN: both operands MN-major (see as_linear_tn_f16); kEpi: epilogue mode, compiled in (keeps the staged epilogue inside the register budget); kPair 2: a pair of CTAs (cluster of 2) drives ONE cta_group::2 MMA of shape 256x256; 1: stand-alone CTA, 128x256
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const GemmParams p) {
  // Paired mode: each CTA stages its own 128 rows of X and its own 128 rows of W per k-block (32 KB instead of 48 KB), the
  // tensor cores of the two SMs exchange the W halves, and every SM's shared memory sees 128 B/clk of TMA fills + MMA
  // operand reads instead of 192 B/clk -- the single-CTA shape is shared-memory-bandwidth bound near 45% tensor utilisation.
  constexpr int kStages = kPair == 2 ? 6 : 4;
  constexpr int kBBytes = kPair == 2 ? B_BYTES / 2 : B_BYTES;
  static_assert(kStages * (A_BYTES + kBBytes) + EPI_STAGE_BYTES + 1024 + 256 <= SMEM_BYTES, "smem budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * A_BYTES;
  uint8_t* epi_s = smem + kStages * (A_BYTES + kBBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_s + EPI_STAGE_BYTES);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = kPair == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8 * kPair);   // the leader's MMA thread waits for the epilogue warps of BOTH CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair == 2) tmem_alloc_2cta<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair == 2) cluster_sync_all();   // peer barriers / TMEM are set up before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cluster_id = blockIdx.x / kPair, n_clusters = gridDim.x / kPair;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_mp = (num_m + kPair - 1) / kPair;          // groups of row blocks, one block per CTA of the pair
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles_per_batch = num_mp * num_n;
  const int tiles = tiles_per_batch * p.batch;
  const int kblocks = (p.K + BK - 1) / BK;                  // (K % 64 != 0 only with kTN: TMA zero-fills the missing rows)

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < tiles; tile += n_clusters) {
        const int bi = tile / tiles_per_batch, tb = tile - bi * tiles_per_batch;
        const int m_blk = kPair * (tb / num_n) + rank, n_blk = tb % num_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (kTN) {
            // operands stored [reduction, M] / [reduction, N]: boxes of 64 reduction rows x 64 contiguous M / N elements
            // (8 KB each, 128-byte swizzle), consumed as MN-major tiles -- no transposed copy of the activations exists
            if (kPair == 2) { if (leader) mbar_expect_tx(&full[stage], 2 * (A_BYTES + kBBytes)); }
            else mbar_expect_tx(&full[stage], A_BYTES + kBBytes);
#pragma unroll
            for (int h = 0; h < A_BYTES / 8192; ++h) {
              if (kPair == 2) tma_load_3d_2cta(smem_a + stage * A_BYTES + h * 8192, &tm_a, &full[stage], m_blk * BM + h * 64, kb * BK, bi);
              else tma_load_3d(smem_a + stage * A_BYTES + h * 8192, &tm_a, &full[stage], m_blk * BM + h * 64, kb * BK, bi);
            }
#pragma unroll
            for (int h = 0; h < kBBytes / 8192; ++h) {
              const int n0 = n_blk * BN + (kPair == 2 ? rank * (BN / 2) : 0) + h * 64;
              if (kPair == 2) tma_load_3d_2cta(smem_b + stage * kBBytes + h * 8192, &tm_b, &full[stage], n0, kb * BK, bi);
              else tma_load_3d(smem_b + stage * kBBytes + h * 8192, &tm_b, &full[stage], n0, kb * BK, bi);
            }
          } else if (kPair == 2) {
            if (leader) mbar_expect_tx(&full[stage], 2 * (A_BYTES + kBBytes));     // both CTAs' boxes land on my barrier
            tma_load_3d_2cta(smem_a + stage * A_BYTES, &tm_a, &full[stage], kb * BK, m_blk * BM, bi);
            tma_load_3d_2cta(smem_b + stage * kBBytes, &tm_b, &full[stage], kb * BK, n_blk * BN + rank * (BN / 2), bi);
          } else {
            mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
            tma_load_3d(smem_a + stage * A_BYTES, &tm_a, &full[stage], kb * BK, m_blk * BM, bi);
            tma_load_3d(smem_b + stage * B_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN, bi);
            tma_load_3d(smem_b + stage * B_BYTES + B_BYTES / 2, &tm_b, &full[stage], kb * BK, n_blk * BN + BN / 2, bi);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(0, BM * kPair, BN) | (kTN ? (1u << 15) | (1u << 16) : 0u);   // MN-major A and B
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < tiles; tile += n_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_base = smem_u32(smem_a + stage * A_BYTES);
            const uint32_t b_base = smem_u32(smem_b + stage * kBBytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // K-major: 16 k = 32 bytes along the row; MN-major: 16 k = two 8-row swizzle atoms (2 KB), 64-element blocks 8 KB apart
              const uint64_t da = kTN ? umma_desc_mn_sw128(a_base + k * 2048, 8192) : umma_desc_k_sw128(a_base + k * 32);
              const uint64_t db = kTN ? umma_desc_mn_sw128(b_base + k * 2048, 8192) : umma_desc_k_sw128(b_base + k * 32);
              if (kPair == 2) mma_f16_ss_2cta(tmem_base + acc * BN, da, db, idesc, (kb | k) != 0);
              else mma_f16_ss(tmem_base + acc * BN, da, db, idesc, (kb | k) != 0);
            }
            if (kPair == 2) {
              tc_commit_2cta_mcast(&empty[stage], (uint16_t)3);            // frees the stage in both CTAs
              if (kb == kblocks - 1) tc_commit_2cta_mcast(&tfull[acc], (uint16_t)3);
            } else {
              tc_commit(&empty[stage]);
              if (kb == kblocks - 1) tc_commit(&tfull[acc]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // Thread = accumulator row (that is how tcgen05.ld hands the tile out), but a warp-wide store of 16 bytes per ROW
    // touches 32 cache lines and costs the LSU 32 cycles -- the K = 768 GEMMs were bound by exactly that.  Every 32 x 32
    // chunk therefore crosses a private, XOR-swizzled shared-memory tile and leaves (and the residual arrives) with each
    // instruction covering whole 128-byte (fp32) / 64-byte (fp16) row segments.
    const int quad = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // which 128-column half of the accumulator this warp drains
    float* st32 = reinterpret_cast<float*>(epi_s + (warp - 2) * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int C = p.heads * 64;
    const int r4 = lane >> 3, c4 = lane & 7;        // fp32 read-back: row 4i + r4, float4 column c4
    const int r8 = lane >> 2, c8 = lane & 3;        // fp16 read-back: row 8i + r8, 16-byte piece c8
    for (int tile = cluster_id; tile < tiles; tile += n_clusters) {
      const int bi = tile / tiles_per_batch, tb = tile - bi * tiles_per_batch;
      const int m_blk = kPair * (tb / num_n) + rank, n_blk = tb % num_n;
      const int m0 = m_blk * BM + quad * 32;         // first row of this warp's 32-row slab
      const int m = m0 + lane;
      const int c_beg = chalf * (BN / 64), c_end = (chalf + 1) * (BN / 64);
      const bool live = m < p.M;
      // the residual does not depend on the MMA: fetch the first chunk before the accumulator is ready and always keep
      // the next chunk's loads in flight
      float4 rnext[8];
      const float* rbase = nullptr;
      if (kEpi == EPI_RESID_F32) {
        rbase = p.resid + bi * p.resid_bstride;
        const int n0 = n_blk * BN + c_beg * 32;
        if (n0 < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mr = m0 + 4 * i + r4;
            if (mr < p.M) rnext[i] = __ldg(reinterpret_cast<const float4*>(rbase + (size_t)mr * p.ldo + n0) + c4);
          }
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      int b_idx = 0, t_idx = 0;
      if (kEpi == EPI_QKV && live) { b_idx = m / p.T; t_idx = m - b_idx * p.T; }
      size_t qk_row[4];                               // EPI_QKV: (b*heads*T + t) of the four rows this lane writes back
      if (kEpi == EPI_QKV) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int mr = min(m0 + 8 * i + r8, p.M - 1);
          const int bb = mr / p.T;
          qk_row[i] = (size_t)bb * p.heads * p.T + (mr - bb * p.T);
        }
      }
#pragma unroll 1
      for (int c = c_beg; c < c_end; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, v);
        const int n0 = n_blk * BN + c * 32;
        tc_wait_ld();
        if (n0 < p.N) {                                 // warp-uniform
          float f[32];
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
              f[4 * i] = fmaf(__uint_as_float(v[4 * i]), p.alpha, b4.x);
              f[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), p.alpha, b4.y);
              f[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), p.alpha, b4.z);
              f[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), p.alpha, b4.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
          }
          const int which = kEpi == EPI_QKV ? n0 / C : 0;
          if (kEpi == EPI_RESID_F32 || kEpi == EPI_F32) {
            // ---- fp32 out: rows of 128 bytes
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(st32 + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            __syncwarp();
            float* obase = reinterpret_cast<float*>(p.out) + bi * p.out_bstride;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + r4, mr = m0 + rr;
              float4 o = *reinterpret_cast<const float4*>(st32 + rr * 32 + ((c4 ^ (rr & 7)) << 2));
              if (kEpi == EPI_RESID_F32) {
                const float4 r = rnext[i];
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                // this slot's value is consumed: start the same row's load for the next chunk right away
                if (c + 1 < c_end && n0 + 32 < p.N && mr < p.M)
                  rnext[i] = __ldg(reinterpret_cast<const float4*>(rbase + (size_t)mr * p.ldo + n0 + 32) + c4);
              }
              if (mr < p.M) *(reinterpret_cast<float4*>(obase + (size_t)mr * p.ldo + n0) + c4) = o;
            }
            __syncwarp();
          } else if (kEpi == EPI_QKV && which == 2) {
            // ---- V^T [B,h,64,Tpad]: consecutive lanes = consecutive tokens, already 64 contiguous bytes per store
            if (live) {
              const int cc = n0 - which * C;
              const int h = cc >> 6, d0 = cc & 63;
              const size_t bh = (size_t)b_idx * p.heads + h;
              __half* dst = p.vt + (bh * 64 + d0) * (size_t)p.Tpad + t_idx;
#pragma unroll
              for (int i = 0; i < 32; ++i) dst[(size_t)i * p.Tpad] = __float2half_rn(f[i]);
            }
          } else {
            // ---- fp16 out (plain, GELU, Q / K head split): rows of 64 bytes
            if (kEpi == EPI_GELU_F16) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
            }
            uint4* st16 = reinterpret_cast<uint4*>(st32);      // row r at 64 r bytes, piece c at position c ^ ((r >> 1) & 3)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __half2 h0 = __floats2half2_rn(f[8 * i + 0], f[8 * i + 1]);
              __half2 h1 = __floats2half2_rn(f[8 * i + 2], f[8 * i + 3]);
              __half2 h2 = __floats2half2_rn(f[8 * i + 4], f[8 * i + 5]);
              __half2 h3 = __floats2half2_rn(f[8 * i + 6], f[8 * i + 7]);
              st16[lane * 4 + (i ^ ((lane >> 1) & 3))] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + r8, mr = m0 + rr;
              const uint4 o = st16[rr * 4 + (c8 ^ ((rr >> 1) & 3))];
              if (mr < p.M) {
                __half* dst;
                if (kEpi == EPI_QKV) {
                  const int cc = n0 - which * C;
                  const int h = cc >> 6, d0 = cc & 63;
                  dst = (which == 0 ? p.q : p.k) + (qk_row[i] + (size_t)h * p.T) * 64 + d0;
                } else {
                  dst = reinterpret_cast<__half*>(p.out) + bi * p.out_bstride + (size_t)mr * p.ldo + n0;
                }
                *(reinterpret_cast<uint4*>(dst) + c8) = o;
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair == 2) mbar_arrive_cluster(&tempty[acc], 0);     // the leader's MMA thread owns the accumulator hand-shake
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair == 2) cluster_sync_all();   // nobody leaves (or frees TMEM) while the pair still works on shared state
  if (warp == 1) {
    if (kPair == 2) tmem_dealloc_2cta<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

int launch_linear(const void* x, const void* w, long long x_rows_per_batch, long long w_rows_per_batch, const GemmParams& p,
                  cudaStream_t stream) {
  if (p.K % BK != 0 || p.M <= 0 || p.N <= 0 || (p.N % 32) != 0 || p.batch < 1 || (p.ldo % 4) != 0) return AS_ERR_BAD_ARG;
  CUtensorMap tm_a, tm_b;
  uint64_t dims_a[3] = {(uint64_t)p.K, (uint64_t)(p.batch > 1 ? x_rows_per_batch : p.M), (uint64_t)p.batch};
  uint64_t str_a[2] = {(uint64_t)p.K * 2, (uint64_t)x_rows_per_batch * p.K * 2};
  uint32_t box_a[3] = {BK, BM, 1};
  uint64_t dims_b[3] = {(uint64_t)p.K, (uint64_t)(p.batch > 1 ? w_rows_per_batch : p.N), (uint64_t)p.batch};
  uint64_t str_b[2] = {(uint64_t)p.K * 2, (uint64_t)w_rows_per_batch * p.K * 2};
  uint32_t box_b[3] = {BK, BN / 2, 1};     // each CTA of the cluster fetches (and multicasts) half of the W tile
  int r = as_encode_tmap(&tm_a, x, 2, 3, dims_a, str_a, box_a);
  if (r) return r;
  r = as_encode_tmap(&tm_b, w, 2, 3, dims_b, str_b, box_b);
  if (r) return r;
  static int num_sms = 0;
  if (!num_sms) {
    int dev;
    AS_CUDA(cudaGetDevice(&dev));
    AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const void* fn = nullptr;
  const bool pair = num_m > 1;
#define AS_PICK(E)                                                                                   \
  case E: fn = pair ? (const void*)linear_tcgen05_kernel<2, E> : (const void*)linear_tcgen05_kernel<1, E>; break;
  switch (p.epi) {
    AS_PICK(EPI_F16) AS_PICK(EPI_GELU_F16) AS_PICK(EPI_RESID_F32) AS_PICK(EPI_QKV) AS_PICK(EPI_F32)
    default: return AS_ERR_BAD_ARG;
  }
#undef AS_PICK
  static bool attr_done[2][8] = {};
  if (!attr_done[pair][p.epi]) {
    AS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done[pair][p.epi] = true;
  }
  void* args[] = {(void*)&tm_a, (void*)&tm_b, (void*)&p};
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (!pair) {
    const int tiles = num_n * p.batch;
    cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
    cfg.numAttrs = 0;
  } else {
    const int tiles = ((num_m + 1) / 2) * num_n * p.batch;       // cluster tiles
    const int max_clusters = num_sms / 2;
    cfg.gridDim = dim3(2 * (tiles < max_clusters ? tiles : max_clusters));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  AS_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  AS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// out [M, N] f32 = a^T b for a [R, M], b [R, N] fp16 row-major (R = reduction length, any value; M, N multiples of 64).
// The weight gradient of a Linear, dW = dY^T X (R = tokens), straight from the row-major activations: the tiles are loaded
// as 64 x 64 boxes and multiplied as MN-major operands, where the K-major kernel would need both tensors transposed first.
extern "C" int as_linear_tn_f16(const void* a_f16, const void* b_f16, float* out, int R, int M, int N, cudaStream_t stream) {
  if (R < 1 || M < 64 || N < 64 || M % 64 || N % 64) return AS_ERR_BAD_ARG;
  GemmParams p{};
  p.M = M; p.N = N; p.K = R; p.epi = EPI_F32; p.out = out; p.heads = 1; p.batch = 1; p.ldo = N; p.alpha = 1.f;
  CUtensorMap tm_a, tm_b;
  uint64_t dims_a[3] = {(uint64_t)M, (uint64_t)R, 1}, str_a[2] = {(uint64_t)M * 2, (uint64_t)R * M * 2};
  uint64_t dims_b[3] = {(uint64_t)N, (uint64_t)R, 1}, str_b[2] = {(uint64_t)N * 2, (uint64_t)R * N * 2};
  uint32_t box[3] = {64, 64, 1};
  int r = as_encode_tmap(&tm_a, a_f16, 2, 3, dims_a, str_a, box);
  if (!r) r = as_encode_tmap(&tm_b, b_f16, 2, 3, dims_b, str_b, box);
  if (r) return r;
  int dev, num_sms;
  AS_CUDA(cudaGetDevice(&dev));
  AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const bool pair = num_m > 1;
  const void* fn = pair ? (const void*)linear_tcgen05_kernel<2, EPI_F32, true> : (const void*)linear_tcgen05_kernel<1, EPI_F32, true>;
  static bool attr_done[2] = {};
  if (!attr_done[pair]) {
    AS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done[pair] = true;
  }
  void* args[] = {(void*)&tm_a, (void*)&tm_b, (void*)&p};
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (!pair) {
    cfg.gridDim = dim3(num_n < num_sms ? num_n : num_sms);
  } else {
    const int tiles = ((num_m + 1) / 2) * num_n, max_clusters = num_sms / 2;
    cfg.gridDim = dim3(2 * (tiles < max_clusters ? tiles : max_clusters));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  AS_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  AS_LAUNCH_CHECK();
  return 0;
}

// mode: 0 = fp16 out, 1 = GELU -> fp16 out, 2 = fp32 out = resid + y, 4 = fp32 out
extern "C" int as_linear_f16(const void* x_f16, const void* w_f16, const float* bias, void* out, const float* resid,
                             int M, int N, int K, int mode, cudaStream_t stream) {
  if (mode != EPI_F16 && mode != EPI_GELU_F16 && mode != EPI_RESID_F32 && mode != EPI_F32) return AS_ERR_BAD_ARG;
  if (mode == EPI_RESID_F32 && !resid) return AS_ERR_BAD_ARG;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.epi = mode; p.bias = bias; p.out = out; p.resid = resid; p.heads = 1;
  p.batch = 1; p.ldo = N; p.alpha = 1.f;
  return launch_linear(x_f16, w_f16, M, N, p, stream);
}

// x [B*T, C] fp16, w [3C, C] fp16, bias [3C] -> q,k [B,h,T,64] fp16, vt [B,h,64,Tpad] fp16 (V transposed, K-major for P*V)
extern "C" int as_qkv_proj_f16(const void* x_f16, const void* w_f16, const float* bias, void* q, void* k, void* vt,
                               int B, int T, int Tpad, int heads, cudaStream_t stream) {
  GemmParams p{};
  const int C = heads * 64;
  p.M = B * T; p.N = 3 * C; p.K = C; p.epi = EPI_QKV; p.bias = bias;
  p.q = (__half*)q; p.k = (__half*)k; p.vt = (__half*)vt; p.T = T; p.Tpad = Tpad; p.heads = heads;
  if (Tpad < T || (Tpad % 8) != 0) return AS_ERR_BAD_ARG;
  p.batch = 1; p.ldo = p.N; p.alpha = 1.f;
  return launch_linear(x_f16, w_f16, p.M, p.N, p, stream);
}

// Batched fp32-output GEMM with scaling: out[b] = resid[b] + alpha * x[b] * w[b]^T   (resid may alias out, may be NULL)
// x [batch, x_rows, K] f16 (M <= x_rows valid rows), w [batch, w_rows, K] f16 (N <= w_rows), out/resid [batch, M, ldo] f32.
// Used by the roll-out slab: split-fp16 (hi + lo) operands give fp32-level accuracy on the fp16 tensor pipe.
extern "C" int as_bgemm_f16_f32(const void* x_f16, const void* w_f16, float* out, const float* resid, int batch, int M,
                                int N, int K, int x_rows, int w_rows, int ldo, long long out_bstride, float alpha,
                                cudaStream_t stream) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.epi = resid ? EPI_RESID_F32 : EPI_F32; p.bias = nullptr; p.out = out; p.resid = resid;
  p.heads = 1; p.batch = batch; p.ldo = ldo; p.alpha = alpha; p.out_bstride = out_bstride; p.resid_bstride = out_bstride;
  if (M > x_rows || N > w_rows) return AS_ERR_BAD_ARG;
  return launch_linear(x_f16, w_f16, x_rows, w_rows, p, stream);
}
