// Multi-head self-attention core of the ViT block on tcgen05 / TMEM, fed by TMA.
// Reference ops replaced (models/vision_transformer.py): :79 q @ k^T * scale, :80 softmax, :83 attn @ v,
// and visual_transformer_det.py:236/242 attn.mean(1) (the head-mean probabilities every layer hands to the
// attention-shift head).  The [B,h,T,T] probability tensor (845 MB / image / layer at 1024^2) is never materialised.
//
// as_mhsa_fwd      : flash-style forward. One CTA = one (batch, head, 128-query tile); 2 CTAs co-reside per SM.
//                    warp 0 TMA producer (Q once, K and V^T tiles through separate 2-stage rings), warp 1 tcgen05.mma
//                    issuer (S = Q K^T into TMEM, O += P V with P read from TMEM), warps 2..5 one softmax thread per
//                    query row.  Default schedule (mhsa_fwd2_kernel<4,1>): one pass over S per tile, no row maximum in
//                    the common case, P handed back in two halves, packed fp32x2 math, a quarter of the exponentials
//                    as an FMA polynomial; mhsa_fwd_kernel is the first-generation two-pass schedule (variant 1).
//                    Outputs O [B,T,C] fp16 and the per-row log2-domain offset m and denominator l [B,h,T].
// as_attn_headmean : second pass for layers whose attention map is consumed: for a (128 x 128) tile loops the heads,
//                    recomputes S_h on tensor cores and accumulates exp2(S_h*c - m_h) / l_h in registers; writes the
//                    head-mean tile once (+ deterministic row-sum partials for the roll-out normaliser, and optionally
//                    the transposed map as a split-fp16 pair = the K-major B operand of the roll-out GEMM).
//                    Default: attn_headmean2_kernel, persistent (one CTA per SM walks a range of tiles, TMA / MMA run
//                    ahead across tile borders, per-warp staged epilogue); attn_headmean_kernel = one CTA per tile.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

using namespace asb;

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = BQ * HD * 2;   // 16 KB: 128 rows x 128 B (also one K tile, also one V^T tile = 2 x (64 x 128 B))

// ------------------------------------------------------------------ forward
constexpr int FWD_THREADS = 192;
constexpr int FWD_SMEM = 5 * TILE_BYTES + 1024 + 256;
constexpr uint32_t COL_S = 0, COL_O = 128, COL_P = 192;

struct FwdParams {
  int T, heads, nkv;
  float scale_log2;     // head_dim^-0.5 * log2(e)
  __half* o;            // [B, T, heads*64]
  float* m;             // [B, heads, T]
  float* l;             // [B, heads, T]
};

__global__ void __launch_bounds__(FWD_THREADS, 2)
mhsa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + TILE_BYTES;          // 2 stages
  uint8_t* v_s = smem + 3 * TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(q_s, &tm_q, q_full, 0, qt * BQ, bh);
      for (int j = 0; j < p.nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
        tma_load_3d(k_s + st * TILE_BYTES, &tm_k, &kv_full[st], 0, j * BKV, bh);
        tma_load_3d(v_s + st * TILE_BYTES, &tm_v, &kv_full[st], j * BKV, 0, bh);
        tma_load_3d(v_s + st * TILE_BYTES + TILE_BYTES / 2, &tm_v, &kv_full[st], j * BKV + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(0, BQ, BKV);
    constexpr uint32_t idesc_o = umma_idesc(0, BQ, HD);
    mbar_wait(q_full, 0);
    const uint32_t q_base = smem_u32(q_s);
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_base = smem_u32(k_s + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + COL_S, umma_desc_k_sw128(q_base + k * 32), umma_desc_k_sw128(k_base + k * 32), idesc_s, k != 0);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < p.nkv; ++j) {
      const int st = j & 1;
      if (j + 1 < p.nkv) issue_s(j + 1);    // S_{j+1} runs on the tensor pipe while the softmax warps finish tile j
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_base = smem_u32(v_s + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          mma_f16_ts(tmem + COL_O, tmem + COL_P + k * 8,
                     umma_desc_k_sw128(v_base + (k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32), idesc_o, (j | k) != 0);
        tc_commit(&kv_empty[st]);
        tc_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int t = qt * BQ + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < p.nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      const bool tail = kv0 + BKV > p.T;
      // pass 1: row maximum of the scaled logits (two 32-column chunks in flight per TMEM wait)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(lane_addr + COL_S + c * 32, va);
        tmem_ld_32x32(lane_addr + COL_S + c * 32 + 32, vb);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s0 = __uint_as_float(va[i]), s1 = __uint_as_float(vb[i]);
          if (tail) {
            if (kv0 + c * 32 + i >= p.T) s0 = -INFINITY;
            if (kv0 + c * 32 + 32 + i >= p.T) s1 = -INFINITY;
          }
          mx = fmaxf(mx, fmaxf(s0, s1));
        }
      }
      // lazy rescaling: the running offset m_run only moves when some row of the warp would otherwise exceed 2^8
      // (P is fp16: values up to 256 are exact enough and far from overflow); most tiles skip the O read-modify-write
      const float tmax = mx * p.scale_log2;
      const bool need = __any_sync(0xffffffffu, tmax > m_run + 8.f);
      float alpha = 1.f;
      if (need) {
        const float m_new = fmaxf(m_run, tmax);
        alpha = ex2_approx(m_run - m_new);                  // 0 on the first tile (m_run = -inf)
        m_run = m_new;
      }
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);            // P buffer and O accumulator are free again
      tc_fence_after();
      // pass 2: p = exp2(s*c - m), packed fp16 into the P columns
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(lane_addr + COL_S + c * 32, va);
        tmem_ld_32x32(lane_addr + COL_S + c * 32 + 32, vb);
        tc_wait_ld();
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // (an FMA-pipe polynomial exp2 for half of the elements was measured SLOWER on B200: 3-register FFMAs issue
          //  every other cycle per scheduler, so 9 extra FMA/ALU instructions cost more than the MUFU slot they free)
          float p0 = ex2_approx(fmaf(__uint_as_float(va[2 * i]), p.scale_log2, -m_run));
          float p1 = ex2_approx(fmaf(__uint_as_float(va[2 * i + 1]), p.scale_log2, -m_run));
          float p2 = ex2_approx(fmaf(__uint_as_float(vb[2 * i]), p.scale_log2, -m_run));
          float p3 = ex2_approx(fmaf(__uint_as_float(vb[2 * i + 1]), p.scale_log2, -m_run));
          if (tail) {
            if (kv0 + c * 32 + 2 * i >= p.T) p0 = 0.f;
            if (kv0 + c * 32 + 2 * i + 1 >= p.T) p1 = 0.f;
            if (kv0 + c * 32 + 32 + 2 * i >= p.T) p2 = 0.f;
            if (kv0 + c * 32 + 33 + 2 * i >= p.T) p3 = 0.f;
          }
          sum += (p0 + p1) + (p2 + p3);
          __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
          pk[i] = *reinterpret_cast<uint32_t*>(&h01);
          pk[16 + i] = *reinterpret_cast<uint32_t*>(&h23);
        }
        tmem_st_32x32(lane_addr + COL_P + c * 16, pk);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);                  // S fully read: the next Q K^T may overwrite it
      l_run = l_run * alpha + sum;
      if (need && j > 0) {                                  // rescale the running output (warp-uniform branch)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + COL_O + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(lane_addr + COL_O + c * 32, v);
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (p.nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const bool live = t < p.T;
    __half* dst = p.o + ((size_t)b * p.T + (live ? t : 0)) * (p.heads * HD) + h * HD;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + COL_O + c * 32, v);   // .sync.aligned: executed by every lane, stores are predicated
      tc_wait_ld();
      if (live) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __half2 hh = __floats2half2_rn(__uint_as_float(v[8 * i + 2 * e]) * inv_l, __uint_as_float(v[8 * i + 2 * e + 1]) * inv_l);
            w[e] = *reinterpret_cast<uint32_t*>(&hh);
          }
          d4[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (live) {
      p.m[(size_t)bh * p.T + t] = m_run;
      p.l[(size_t)bh * p.T + t] = l_run;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------ forward, second generation
// Same tiling and TMEM map as mhsa_fwd_kernel; what changed is the schedule of the softmax warps, which were idle 40 %
// of the time on the exp2 (MUFU) pipe that bounds this kernel:
//   * one pass over S per tile in the common case: a 32-column chunk is checked against the running offset right after
//     its TMEM load (FMNMX on the ALU pipe) and exponentiated immediately, while the next chunk's load is in flight; the
//     two-pass max / rescale path only runs on the first tile, on a ragged last tile and when some row of the warp
//     leaves the 2^8 window of the lazy rescaling
//   * S is released to the MMA warp as soon as its last chunk sits in registers, so Q K^T of the next tile overlaps the
//     second half of the exponentials instead of following them
//   * the P buffer is handed back in two halves (a commit after the first four P V MMAs): the softmax warps only wait
//     for P V of the previous tile after they have already computed half of the new probabilities
//   * K and V^T have separate rings: a K stage is free again after Q K^T, one tile earlier than V^T
//   * packed fp32x2 FMA / ADD (FFMA2 / FADD2) halve the issue slots of the logit scaling and the row sums
//   * POLY pairs of every 16-pair chunk take exp2 on the FMA pipe (packed ex2_poly2) instead of MUFU: with the FMNMX
//     chain of the window test gone (MODE 1) a quarter of the exponentials fits into the freed issue slots
constexpr int FWD2_SMEM = 5 * TILE_BYTES + 1024 + 256;

__device__ __forceinline__ uint64_t pack2(float a, float b) {
  return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32);
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float lo32f(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi32f(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }

// ex2_poly (common.cuh) for a packed pair: the range reduction and the degree-4 polynomial run as FADD2 / FFMA2
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& p0, float& p1) {
  const float x0 = fmaxf(lo32f(x2), -126.f), x1 = fmaxf(hi32f(x2), -126.f);
  const uint64_t x = pack2(x0, x1);
  const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
  const uint64_t fl = fadd2(x, magic);                       // integer part in the low mantissa bits
  const uint64_t ni = fadd2(fl, nmagic);                     // round(x)
  const uint64_t f = ffma2(ni, pack2(-1.f, -1.f), x);        // x - round(x), |f| <= 0.5
  uint64_t q = ffma2(pack2(0.009676037356257439f, 0.009676037356257439f), f, pack2(0.05592203512787819f, 0.05592203512787819f));
  q = ffma2(q, f, pack2(0.2402210682630539f, 0.2402210682630539f));
  q = ffma2(q, f, pack2(0.6931210160255432f, 0.6931210160255432f));
  q = ffma2(q, f, pack2(1.0000001192092896f, 1.0000001192092896f));
  p0 = __int_as_float(__float_as_int(lo32f(q)) + (__float_as_int(lo32f(fl)) << 23));
  p1 = __int_as_float(__float_as_int(hi32f(q)) + (__float_as_int(hi32f(fl)) << 23));
}

template <int POLY, int MODE>   // MODE 0: window test on the raw logits of every chunk (FMNMX), S released before the last
                                // chunk's exponentials; MODE 1: no max at all -- the row sum of the tile tells afterwards
                                // whether any probability left the fp16-safe range (sum < 2^15), S released after that test
__global__ void __launch_bounds__(FWD_THREADS, 2)
mhsa_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + TILE_BYTES;          // 2 stages
  uint8_t* v_s = smem + 3 * TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* pv_half = bars + 12;
  uint64_t* pv_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_half, 1);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(q_s, &tm_q, q_full, 0, qt * BQ, bh);
      auto load_k = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_3d(k_s + st * TILE_BYTES, &tm_k, &k_full[st], 0, j * BKV, bh);
      };
      load_k(0);
      for (int j = 0; j < p.nkv; ++j) {
        const int st = j & 1;
        if (j + 1 < p.nkv) load_k(j + 1);
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_3d(v_s + st * TILE_BYTES, &tm_v, &v_full[st], j * BKV, 0, bh);
        tma_load_3d(v_s + st * TILE_BYTES + TILE_BYTES / 2, &tm_v, &v_full[st], j * BKV + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(0, BQ, BKV);
    constexpr uint32_t idesc_o = umma_idesc(0, BQ, HD);
    mbar_wait(q_full, 0);
    const uint32_t q_base = smem_u32(q_s);
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_base = smem_u32(k_s + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + COL_S, umma_desc_k_sw128(q_base + k * 32), umma_desc_k_sw128(k_base + k * 32), idesc_s, k != 0);
        tc_commit(&k_empty[st]);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < p.nkv; ++j) {
      const int st = j & 1;
      if (j + 1 < p.nkv) issue_s(j + 1);
      mbar_wait(&v_full[st], (j >> 1) & 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_base = smem_u32(v_s + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          mma_f16_ts(tmem + COL_O, tmem + COL_P + k * 8,
                     umma_desc_k_sw128(v_base + (k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32), idesc_o, (j | k) != 0);
          if (k == 3) tc_commit(pv_half);         // P columns 0..31 (keys 0..63) are consumed
        }
        tc_commit(&v_empty[st]);
        tc_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int t = qt * BQ + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const uint64_t c2 = pack2(p.scale_log2, p.scale_log2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < p.nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      const bool tail = kv0 + BKV > p.T;
      bool slow = (j == 0) || tail;
      if (!slow) {
        // ---------------- fast path: one pass, chunk c+1 loading while chunk c is exponentiated
        const float lim = (m_run + 8.f) / p.scale_log2;     // raw-logit bound of the lazy-rescaling window (scale > 0)
        const uint64_t nm2 = pack2(-m_run, -m_run);
        uint64_t acc0 = 0ull, acc1 = 0ull;                  // two packed partial row sums
        uint32_t va[32], vb[32], pk[32];
        tmem_ld_32x32(lane_addr + COL_S, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&cur)[32] = (c & 1) ? vb : va;
          uint32_t (&nxt)[32] = (c & 1) ? va : vb;
          tc_wait_ld();
          if (c < 3) tmem_ld_32x32(lane_addr + COL_S + (c + 1) * 32, nxt);
          if (MODE == 0) {
            float mx = fmaxf(__uint_as_float(cur[0]), __uint_as_float(cur[1]));
            float mxb = fmaxf(__uint_as_float(cur[2]), __uint_as_float(cur[3]));
#pragma unroll
            for (int i = 4; i < 32; i += 4) {
              mx = fmaxf(mx, fmaxf(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
              mxb = fmaxf(mxb, fmaxf(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])));
            }
            if (__any_sync(0xffffffffu, fmaxf(mx, mxb) > lim)) { slow = true; break; }
            if (c == 3) {                                   // S is in registers and inside the window: release it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_empty);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x = ffma2((uint64_t)cur[2 * i] | ((uint64_t)cur[2 * i + 1] << 32), c2, nm2);
            float p0, p1;
            if (i < POLY) ex2_poly2(x, p0, p1);
            else { p0 = ex2_approx(lo32f(x)); p1 = ex2_approx(hi32f(x)); }
            if (i & 1) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1));
            __half2 hh = __floats2half2_rn(p0, p1);
            pk[(c & 1) * 16 + i] = *reinterpret_cast<uint32_t*>(&hh);
          }
          if (c & 1) {
            mbar_wait(c == 1 ? pv_half : pv_done, (j - 1) & 1);     // that half of the P buffer has been consumed (j > 0 here)
            tc_fence_after();
            tmem_st_32x32(lane_addr + COL_P + (c >> 1) * 32, pk);
          }
        }
        if (MODE == 0) {
          if (!slow) {
            const uint64_t a = fadd2(acc0, acc1);
            l_run += lo32f(a) + hi32f(a);
          }
        } else {
          const uint64_t a = fadd2(acc0, acc1);
          const float tot = lo32f(a) + hi32f(a);
          // every probability of the tile is below its row sum: sum < 2^15 means nothing came near the fp16 limit (2^16)
          if (__any_sync(0xffffffffu, !(tot < 32768.f))) {
            slow = true;                                    // S is still intact: the exact two-pass tile below redoes it
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);
            l_run += tot;
          }
        }
      }
      if (slow) {
        // ---------------- slow path: exact two-pass tile (first tile, ragged tail, offset moves).  S is still intact:
        // the fast path only releases it after every chunk passed the window test.
        tc_wait_ld();
        tc_wait_st();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          uint32_t va[32], vb[32];
          tmem_ld_32x32(lane_addr + COL_S + c * 32, va);
          tmem_ld_32x32(lane_addr + COL_S + c * 32 + 32, vb);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float s0 = __uint_as_float(va[i]), s1 = __uint_as_float(vb[i]);
            if (tail) {
              if (kv0 + c * 32 + i >= p.T) s0 = -INFINITY;
              if (kv0 + c * 32 + 32 + i >= p.T) s1 = -INFINITY;
            }
            mx = fmaxf(mx, fmaxf(s0, s1));
          }
        }
        const float tmax = mx * p.scale_log2;
        const bool need = __any_sync(0xffffffffu, tmax > m_run + 8.f);
        float alpha = 1.f;
        if (need) {
          const float m_new = fmaxf(m_run, tmax);
          alpha = ex2_approx(m_run - m_new);                  // 0 on the first tile (m_run = -inf)
          m_run = m_new;
        }
        if (j > 0) mbar_wait(pv_done, (j - 1) & 1);            // P buffer and O accumulator are free again
        tc_fence_after();
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          uint32_t va[32], vb[32];
          tmem_ld_32x32(lane_addr + COL_S + c * 32, va);
          tmem_ld_32x32(lane_addr + COL_S + c * 32 + 32, vb);
          tc_wait_ld();
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(va[2 * i]), p.scale_log2, -m_run));
            float p1 = ex2_approx(fmaf(__uint_as_float(va[2 * i + 1]), p.scale_log2, -m_run));
            float p2 = ex2_approx(fmaf(__uint_as_float(vb[2 * i]), p.scale_log2, -m_run));
            float p3 = ex2_approx(fmaf(__uint_as_float(vb[2 * i + 1]), p.scale_log2, -m_run));
            if (tail) {
              if (kv0 + c * 32 + 2 * i >= p.T) p0 = 0.f;
              if (kv0 + c * 32 + 2 * i + 1 >= p.T) p1 = 0.f;
              if (kv0 + c * 32 + 32 + 2 * i >= p.T) p2 = 0.f;
              if (kv0 + c * 32 + 33 + 2 * i >= p.T) p3 = 0.f;
            }
            sum += (p0 + p1) + (p2 + p3);
            __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
            pk[i] = *reinterpret_cast<uint32_t*>(&h01);
            pk[16 + i] = *reinterpret_cast<uint32_t*>(&h23);
          }
          tmem_st_32x32(lane_addr + COL_P + c * 16, pk);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        l_run = l_run * alpha + sum;
        if (need && j > 0) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + COL_O + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(lane_addr + COL_O + c * 32, v);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (p.nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const bool live = t < p.T;
    __half* dst = p.o + ((size_t)b * p.T + (live ? t : 0)) * (p.heads * HD) + h * HD;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + COL_O + c * 32, v);
      tc_wait_ld();
      if (live) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __half2 hh = __floats2half2_rn(__uint_as_float(v[8 * i + 2 * e]) * inv_l, __uint_as_float(v[8 * i + 2 * e + 1]) * inv_l);
            w[e] = *reinterpret_cast<uint32_t*>(&hh);
          }
          d4[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (live) {
      p.m[(size_t)bh * p.T + t] = m_run;
      p.l[(size_t)bh * p.T + t] = l_run;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------ head-mean probabilities
constexpr int HM_THREADS = 576;   // TMA warp, MMA warp, 4 x 4 math warps (each quartet covers 32 of the 128 key columns)
constexpr int HM_MATH = HM_THREADS - 64;
constexpr int HM_STAGES = 4;
constexpr int HM_SMEM_TILES = 2 * HM_STAGES * TILE_BYTES;     // HM_STAGES x (Q_h, K_h)
constexpr int HM_STAGE_LD = 129;                            // floats per staged output row (padded: conflict-free)
constexpr int HM_MAX_HEADS = 16;
constexpr int HM_SMEM = HM_SMEM_TILES + 1024 + 256 + BQ * HM_STAGE_LD * 4 + HM_MAX_HEADS * BQ * 8;

struct HmParams {
  int T, heads, ld;           // ld = row stride (floats) of the output
  float scale_log2;
  const float* m;             // [B, heads, T]
  const float* l;
  float* out;                 // [B, T, ld]
  float* rowsum_part;         // [B, T, ntile] or null
  int ntile;
  __half* t_hi;               // optional: TRANSPOSED map, split fp16 (hi + lo), scaled: [B, ldt, ldt], fully written
  __half* t_lo;
  int ldt;
  float t_scale;
  int qt0, nqt;               // query tiles [qt0, qt0 + nqt) are produced (persistent schedule only); default: all
  int per_image;              // tile order of the persistent schedule (see the kernel)
};

__global__ void __launch_bounds__(HM_THREADS, 1)
attn_headmean_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const HmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HM_SMEM_TILES);
  uint64_t* full = bars;                       // HM_STAGES
  uint64_t* empty = bars + HM_STAGES;          // HM_STAGES
  uint64_t* s_full = bars + 2 * HM_STAGES;     // 2 (TMEM S buffers)
  uint64_t* s_empty = bars + 2 * HM_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HM_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
    for (int i = 0; i < HM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], HM_MATH / 32); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int h = 0; h < p.heads; ++h) {
        const int st = h % HM_STAGES;
        mbar_wait(&empty[st], ((h / HM_STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], 2 * TILE_BYTES);
        tma_load_3d(smem + (2 * st) * TILE_BYTES, &tm_q, &full[st], 0, qt * BQ, b * p.heads + h);
        tma_load_3d(smem + (2 * st + 1) * TILE_BYTES, &tm_k, &full[st], 0, kt * BKV, b * p.heads + h);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(0, BQ, BKV);
    for (int h = 0; h < p.heads; ++h) {
      const int st = h % HM_STAGES, tb = h & 1;
      mbar_wait(&full[st], (h / HM_STAGES) & 1);
      mbar_wait(&s_empty[tb], ((h >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q_base = smem_u32(smem + (2 * st) * TILE_BYTES), k_base = smem_u32(smem + (2 * st + 1) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + tb * 128, umma_desc_k_sw128(q_base + k * 32), umma_desc_k_sw128(k_base + k * 32), idesc_s, k != 0);
        tc_commit(&empty[st]);
        tc_commit(&s_full[tb]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int cslice = (warp - 2) >> 2;                // which 32-column slice of the tile this warp accumulates
    const int row = quad * 32 + lane;
    const int t = qt * BQ + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + cslice * 32;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    // softmax row statistics of all heads for this query tile -> smem once (keeps global-load latency off the head loop)
    float2* ml_s = reinterpret_cast<float2*>(smem + HM_SMEM_TILES + 256 + BQ * HM_STAGE_LD * 4);
    for (int i = threadIdx.x - 64; i < p.heads * BQ; i += HM_MATH) {
      const int hh = i / BQ, r = i - hh * BQ, tt = qt * BQ + r;
      float2 v = make_float2(0.f, 0.f);                  // rows >= T contribute exactly 0
      if (tt < p.T) {
        const size_t si = ((size_t)b * p.heads + hh) * p.T + tt;
        v = make_float2(p.m[si], 1.f / p.l[si]);
      }
      ml_s[i] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(HM_MATH) : "memory");
    for (int h = 0; h < p.heads; ++h) {
      const int st = h & 1;
      const float2 mlv = ml_s[h * BQ + row];
      const float mrow = mlv.x, inv_l = mlv.y;
      mbar_wait(&s_full[st], (h >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32];
      tmem_ld_32x32(lane_addr + st * 128, v0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);       // S_h is in registers: the next head's MMA may overwrite the buffer
#pragma unroll
      for (int i = 0; i < 32; ++i)
        acc[i] = fmaf(ex2_approx(fmaf(__uint_as_float(v0[i]), p.scale_log2, -mrow)), inv_l, acc[i]);
    }
    // stage through smem so the global stores are row-contiguous
    float* stage = reinterpret_cast<float*>(smem + HM_SMEM_TILES + 256);
    const float inv_h = 1.f / (float)p.heads;
    float rs = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = cslice * 32 + i;
      const float v = (kt * BKV + c < p.T) ? acc[i] * inv_h : 0.f;
      rs += v;
      stage[row * HM_STAGE_LD + c] = v;
    }
    float* rs_s = reinterpret_cast<float*>(ml_s);          // the statistics are dead now: reuse for the four slice sums
    rs_s[cslice * BQ + row] = rs;                          // each thread's 32 values were added in a fixed order
    asm volatile("bar.sync 1, %0;" ::"n"(HM_MATH) : "memory");
    if (p.rowsum_part && t < p.T && cslice == 0)
      p.rowsum_part[((size_t)b * p.T + t) * p.ntile + kt] = (rs_s[row] + rs_s[BQ + row]) + (rs_s[2 * BQ + row] + rs_s[3 * BQ + row]);
    const int ncol = min(BKV, p.T - kt * BKV);
    for (int r = (warp - 2); r < BQ; r += HM_MATH / 32) {
      const int tr = qt * BQ + r;
      if (tr >= p.T) break;
      float* dst = p.out + ((size_t)b * p.T + tr) * p.ld + kt * BKV;
      for (int c = lane; c < ncol; c += 32) dst[c] = stage[r * HM_STAGE_LD + c];
    }
    if (p.t_hi) {
      // transposed tile for the roll-out GEMM (B operand, K-major): At[n = key][k = query], x = hi + lo in fp16
      for (int c = (warp - 2); c < BKV; c += HM_MATH / 32) {
        const size_t o = ((size_t)b * p.ldt + kt * BKV + c) * p.ldt + qt * BQ;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r0 = half * 64 + 2 * lane;
          const float v0 = stage[r0 * HM_STAGE_LD + c] * p.t_scale, v1 = stage[(r0 + 1) * HM_STAGE_LD + c] * p.t_scale;
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          const __half2 hi = __halves2half2(h0, h1);
          const __half2 lo = __floats2half2_rn(v0 - __half2float(h0), v1 - __half2float(h1));
          *reinterpret_cast<__half2*>(p.t_hi + o + r0) = hi;
          *reinterpret_cast<__half2*>(p.t_lo + o + r0) = lo;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------ head-mean probabilities, persistent schedule
// Same arithmetic as attn_headmean_kernel, different schedule.  The first kernel spends more time around its 12-head main
// loop than in it (one CTA per SM: TMA / statistics prologue and the staged 128 KB epilogue of every tile are exposed).
// Here one CTA per SM walks a contiguous range of (batch, query tile, key tile) triples:
//   * the TMA producer and the MMA issuer run ahead across tile borders (4-stage operand ring, FOUR S buffers = all 512
//     TMEM columns), so the next tile's first heads are already in TMEM when the math warps get there
//   * every math warp stages its 32 x 32 result in a private shared-memory tile and writes it out itself (row-major fp32
//     and the transposed split-fp16 pair as 16-byte stores covering whole 64 / 128-byte runs, one row-sum partial per
//     32-column slice): no block barrier, and the stores drain while the next tile's exponentials are running
//   * the softmax statistics are reloaded only when the (batch, query tile) changes (every ~33 tiles)
constexpr int HM2_STAGES = 4;
constexpr int HM2_NBUF = 4;
constexpr int HM2_SMEM_TILES = 2 * HM2_STAGES * TILE_BYTES;
constexpr int HM2_WSTAGE = 32 * 33;                          // floats per math warp: its 32 x 32 result, odd row stride
constexpr int HM2_SMEM = HM2_SMEM_TILES + 1024 + 256 + HM_MAX_HEADS * BQ * 8 + (HM_MATH / 32) * HM2_WSTAGE * 4;

__global__ void __launch_bounds__(HM_THREADS, 1)
attn_headmean2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const HmParams p,
                      const int n_tiles, const int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HM2_SMEM_TILES);
  uint64_t* full = bars;                         // HM2_STAGES
  uint64_t* empty = bars + HM2_STAGES;           // HM2_STAGES
  uint64_t* s_full = bars + 2 * HM2_STAGES;      // HM2_NBUF
  uint64_t* s_empty = s_full + HM2_NBUF;         // HM2_NBUF
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + HM2_NBUF);
  float2* ml_s = reinterpret_cast<float2*>(smem + HM2_SMEM_TILES + 256);
  float* stage_s = reinterpret_cast<float*>(smem + HM2_SMEM_TILES + 256 + HM_MAX_HEADS * BQ * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Contiguous tile ranges per CTA.  (Dealing the tiles round-robin -- the whole grid on a few query-tile rows of ONE image at a
  // time, so that the live Q / K set is one image's instead of the batch's -- was measured in round 2: DRAM reads do drop, but
  // the row statistics are then reloaded for every tile behind two 512-thread barriers and the pass got 30 % SLOWER
  // (3.30 -> 4.31 ms per step); the kernel is MUFU-bound, not DRAM-bound.)
  // p.per_image (env AS_HEADMEAN_ORDER=image, experiment): every image's tiles are split contiguously over the grid and the images
  // are walked one after the other, so the grid works on ONE image at a time (live set = its Q and K, 13 MB) while a CTA still
  // sweeps consecutive key tiles of a query row.  Measured: 0.643 ms vs 0.557 ms for the default order (roll-out operands only,
  // profiles/microbench_headmean_r3.txt) -- the statistics reloads cost more than the DRAM reads they save; not the default.
  const int nt = p.ntile;
  const int per_img = p.nqt * nt;
  const int img_lo = (int)((long long)blockIdx.x * per_img / gridDim.x), img_hi = (int)((long long)(blockIdx.x + 1) * per_img / gridDim.x);
  const int img_cnt = img_hi - img_lo;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int tile1 = min(n_tiles, tile0 + tiles_per_cta);
  const int my_tiles = p.per_image ? (n_tiles / per_img) * img_cnt : max(tile1 - tile0, 0);
  auto tile_of = [&](int ti) { return p.per_image ? (ti / img_cnt) * per_img + img_lo + ti % img_cnt : tile0 + ti; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
    for (int i = 0; i < HM2_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < HM2_NBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], HM_MATH / 32); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t g = 0;
      // every (q-tile, k-tile) re-reads the Q / K tiles of all heads: keep them in L2 against the output streams
      const uint64_t keep = l2_policy_evict_last();
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int tile = tile_of(ti);
        const int b = tile / (p.nqt * nt), r = tile - b * p.nqt * nt, qt = p.qt0 + r / nt, kt = r % nt;
        for (int h = 0; h < p.heads; ++h, ++g) {
          const uint32_t st = g % HM2_STAGES;
          mbar_wait(&empty[st], ((g / HM2_STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[st], 2 * TILE_BYTES);
          tma_load_3d_hint(smem + (2 * st) * TILE_BYTES, &tm_q, &full[st], 0, qt * BQ, b * p.heads + h, keep);
          tma_load_3d_hint(smem + (2 * st + 1) * TILE_BYTES, &tm_k, &full[st], 0, kt * BKV, b * p.heads + h, keep);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(0, BQ, BKV);
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)p.heads;
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t st = g % HM2_STAGES, tb = g % HM2_NBUF;
      mbar_wait(&full[st], (g / HM2_STAGES) & 1);
      mbar_wait(&s_empty[tb], ((g / HM2_NBUF) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q_base = smem_u32(smem + (2 * st) * TILE_BYTES), k_base = smem_u32(smem + (2 * st + 1) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + tb * 128, umma_desc_k_sw128(q_base + k * 32), umma_desc_k_sw128(k_base + k * 32), idesc_s, k != 0);
        tc_commit(&empty[st]);
        tc_commit(&s_full[tb]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int cslice = (warp - 2) >> 2;                // which 32-column slice of the tile this warp accumulates
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + cslice * 32;
    const float inv_h = 1.f / (float)p.heads;
    uint32_t g = 0;
    int cur_bq = -1;
    const uint64_t stream_out = l2_policy_evict_first();   // the maps are written once and read by a later kernel
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int tile = tile_of(ti);
      const int b = tile / (p.nqt * nt), r = tile - b * p.nqt * nt, qt = p.qt0 + r / nt, kt = r % nt;
      const int t = qt * BQ + row;
      if (b * nt + qt != cur_bq) {                     // new query tile: softmax row statistics of all heads -> smem
        cur_bq = b * nt + qt;
        asm volatile("bar.sync 1, %0;" ::"n"(HM_MATH) : "memory");     // nobody still reads the previous statistics
        for (int i = threadIdx.x - 64; i < p.heads * BQ; i += HM_MATH) {
          const int hh = i / BQ, rr = i - hh * BQ, tt = qt * BQ + rr;
          float2 v = make_float2(0.f, 0.f);              // rows >= T contribute exactly 0
          if (tt < p.T) {
            const size_t si = ((size_t)b * p.heads + hh) * p.T + tt;
            v = make_float2(p.m[si], 1.f / p.l[si]);
          }
          ml_s[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(HM_MATH) : "memory");
      }
      uint64_t acc2[16];                               // packed fp32 pairs: FFMA2 halves the issue slots around the MUFU
#pragma unroll
      for (int i = 0; i < 16; ++i) acc2[i] = 0ull;
      const uint64_t c2 = pack2(p.scale_log2, p.scale_log2);
      for (int h = 0; h < p.heads; ++h, ++g) {
        const uint32_t tb = g % HM2_NBUF;
        const float2 mlv = ml_s[h * BQ + row];
        const uint64_t nm2 = pack2(-mlv.x, -mlv.x), il2 = pack2(mlv.y, mlv.y);
        mbar_wait(&s_full[tb], (g / HM2_NBUF) & 1);
        tc_fence_after();
        uint32_t v0[32];
        tmem_ld_32x32(lane_addr + tb * 128, v0);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[tb]);       // S_h is in registers: a later head's MMA may overwrite the buffer
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t x = ffma2((uint64_t)v0[2 * i] | ((uint64_t)v0[2 * i + 1] << 32), c2, nm2);
          // (moving a share of these exponentials to the FMA pipe, as the forward kernel does, was measured: no gain here)
          acc2[i] = ffma2(pack2(ex2_approx(lo32f(x)), ex2_approx(hi32f(x))), il2, acc2[i]);
        }
      }
      float acc[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc[2 * i] = lo32f(acc2[i]); acc[2 * i + 1] = hi32f(acc2[i]); }
      // results: through a per-warp staging tile (no block barrier, the stores drain under the next tile's exponentials)
      // so that every global store instruction writes whole 64 / 128-byte runs
      const int c0 = kt * BKV + cslice * 32;
      float* st = stage_s + (warp - 2) * HM2_WSTAGE;
      float rs = 0.f;
      __syncwarp();                                    // the previous tile's read-back is complete
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float v = (c0 + i < p.T) ? acc[i] * inv_h : 0.f;
        rs += v;
        st[lane * 33 + i] = v;
      }
      if (t < p.T && p.rowsum_part) p.rowsum_part[((size_t)b * p.T + t) * (4 * nt) + kt * 4 + cslice] = rs;
      __syncwarp();
      if (p.out) {
        const int rr = lane >> 3, cc = (lane & 7) * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r2 = it * 4 + rr, t2 = qt * BQ + quad * 32 + r2;
          const float* sp = st + r2 * 33 + cc;
          const float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
          if (t2 < p.T) st_global_f4_hint(p.out + ((size_t)b * p.T + t2) * p.ld + c0 + cc, v, stream_out);
        }
      }
      if (p.t_hi) {
        // transposed tile for the roll-out GEMM (B operand, K-major): At[n = key][k = query], x = hi + lo in fp16; the
        // padding (key >= T or query >= T) is written as zeros (the staged values are exactly 0 there)
        const int cb = lane >> 2, r0 = (lane & 3) * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int col = it * 8 + cb;
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = st[(r0 + 2 * e) * 33 + col] * p.t_scale, v1 = st[(r0 + 2 * e + 1) * 33 + col] * p.t_scale;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half2 hi = __halves2half2(h0, h1);
            const __half2 lo = __floats2half2_rn(v0 - __half2float(h0), v1 - __half2float(h1));
            hw[e] = *reinterpret_cast<const uint32_t*>(&hi);
            lw[e] = *reinterpret_cast<const uint32_t*>(&lo);
          }
          const size_t o = ((size_t)b * p.ldt + c0 + col) * p.ldt + qt * BQ + quad * 32 + r0;
          st_global_u4_hint(p.t_hi + o, make_uint4(hw[0], hw[1], hw[2], hw[3]), stream_out);
          st_global_u4_hint(p.t_lo + o, make_uint4(lw[0], lw[1], lw[2], lw[3]), stream_out);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int encode_qk(CUtensorMap* tm, const void* base, int BH, int T) {
  uint64_t dims[3] = {HD, (uint64_t)T, (uint64_t)BH};
  uint64_t str[2] = {HD * 2, (uint64_t)T * HD * 2};
  uint32_t box[3] = {HD, BQ, 1};
  return as_encode_tmap(tm, base, 2, 3, dims, str, box);
}

int g_mhsa_variant = -1;
int mhsa_variant() {
  if (g_mhsa_variant < 0) {
    const char* e = getenv("AS_MHSA_VARIANT");
    g_mhsa_variant = e ? atoi(e) : 4;
  }
  return g_mhsa_variant;
}

}  // namespace

// Schedule of the attention forward: 1 = first-generation kernel (two passes over S per tile), 2 = single pass with a window
// test per chunk, 3 = single pass with the row-sum test, 4 (default) = 3 with a quarter of the exponentials on the FMA pipe
// (measured at B=8, T=4197, 12 heads: 0.68 / 0.60 / 0.61 / 0.55 ms).  Also: env AS_MHSA_VARIANT.
// Two further schedules were built, verified against the same tests and dropped because they did not move the time:
// eight softmax warps per CTA (two threads per query row, verdict / maximum exchanged through a 64-thread named barrier):
// 0.563 ms, and 64-key tiles with S and P double-buffered in TMEM (the softmax warps never wait for the tensor cores):
// 0.568 ms.  Neither latency nor warp count is what bounds this kernel.
extern "C" int as_mhsa_set_variant(int v) {
  if (v < 1 || v > 4) return AS_ERR_BAD_ARG;
  g_mhsa_variant = v;
  return 0;
}

extern "C" int as_mhsa_fwd(const void* q, const void* k, const void* vt, void* o, float* m, float* l, int B, int T,
                           int Tpad, int heads, cudaStream_t stream) {
  if (Tpad % BKV || Tpad < T) return AS_ERR_BAD_ARG;
  CUtensorMap tm_q, tm_k, tm_v;
  int r = encode_qk(&tm_q, q, B * heads, T);
  if (r) return r;
  r = encode_qk(&tm_k, k, B * heads, T);
  if (r) return r;
  uint64_t dims[3] = {(uint64_t)Tpad, HD, (uint64_t)B * heads};
  uint64_t str[2] = {(uint64_t)Tpad * 2, (uint64_t)Tpad * HD * 2};
  uint32_t box[3] = {64, HD, 1};
  r = as_encode_tmap(&tm_v, vt, 2, 3, dims, str, box);
  if (r) return r;
  static bool attr = false;
  if (!attr) {
    AS_CUDA(cudaFuncSetAttribute(mhsa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    AS_CUDA(cudaFuncSetAttribute(mhsa_fwd2_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD2_SMEM));
    AS_CUDA(cudaFuncSetAttribute(mhsa_fwd2_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD2_SMEM));
    AS_CUDA(cudaFuncSetAttribute(mhsa_fwd2_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD2_SMEM));
    attr = true;
  }
  FwdParams p;
  p.T = T; p.heads = heads; p.nkv = (T + BKV - 1) / BKV;
  p.scale_log2 = (float)(0.125 * 1.4426950408889634);   // head_dim^-0.5 (VT:67), head_dim = 64
  p.o = (__half*)o; p.m = m; p.l = l;
  const dim3 grid((T + BQ - 1) / BQ, heads, B);
  switch (mhsa_variant()) {
    case 1: mhsa_fwd_kernel<<<grid, FWD_THREADS, FWD_SMEM, stream>>>(tm_q, tm_k, tm_v, p); break;
    case 3: mhsa_fwd2_kernel<0, 1><<<grid, FWD_THREADS, FWD2_SMEM, stream>>>(tm_q, tm_k, tm_v, p); break;
    default: mhsa_fwd2_kernel<4, 1><<<grid, FWD_THREADS, FWD2_SMEM, stream>>>(tm_q, tm_k, tm_v, p); break;
    case 2: mhsa_fwd2_kernel<0, 0><<<grid, FWD_THREADS, FWD2_SMEM, stream>>>(tm_q, tm_k, tm_v, p); break;
  }
  AS_LAUNCH_CHECK();
  return 0;
}

// out [B,T,ld] fp32 (ld >= T), rowsum_part [B,T,rowsum_slices*ceil(T/128)] (may be null).  rowsum_slices selects the
// schedule: 4 = persistent kernel (one row-sum partial per 32-column slice), 1 = first-generation kernel (one per tile).
// Persistent schedule only: ``out`` may be null (only the transposed split-fp16 pair and the row sums are produced -- all the
// roll-out reads of a layer that is not the last, RH:1265), and ``q_row0`` > 0 restricts the work to the query tiles that
// contain rows [q_row0, T) (the roll-out reads only the point-token rows of the LAST layer, RH:2272); rows of ``out`` /
// ``rowsum_part`` before the first such tile are not written.
extern "C" int as_attn_headmean_ex(const void* q, const void* k, const float* m, const float* l, float* out, int ld,
                                   float* rowsum_part, int rowsum_slices, void* t_hi, void* t_lo, int ldt, float t_scale, int B,
                                   int T, int heads, int q_row0, cudaStream_t stream) {
  if (ld < T || heads > HM_MAX_HEADS || q_row0 < 0 || q_row0 >= T) return AS_ERR_BAD_ARG;
  if (t_hi && (!t_lo || ldt != (T + BKV - 1) / BKV * BKV)) return AS_ERR_BAD_ARG;
  if (!out && !t_hi) return AS_ERR_BAD_ARG;
  if ((!out || q_row0 > 0) && rowsum_slices != 4) return AS_ERR_BAD_ARG;
  if (q_row0 > 0 && t_hi) return AS_ERR_BAD_ARG;            // the transposed copy must be complete: it is a GEMM operand
  CUtensorMap tm_q, tm_k;
  int r = encode_qk(&tm_q, q, B * heads, T);
  if (r) return r;
  r = encode_qk(&tm_k, k, B * heads, T);
  if (r) return r;
  static bool attr = false;
  static int num_sms = 0;
  if (!attr) {
    int dev;
    AS_CUDA(cudaGetDevice(&dev));
    AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    AS_CUDA(cudaFuncSetAttribute(attn_headmean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HM_SMEM));
    AS_CUDA(cudaFuncSetAttribute(attn_headmean2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HM2_SMEM));
    attr = true;
  }
  HmParams p;
  p.T = T; p.heads = heads; p.ld = ld; p.scale_log2 = (float)(0.125 * 1.4426950408889634);
  p.m = m; p.l = l; p.out = out; p.rowsum_part = rowsum_part; p.ntile = (T + BKV - 1) / BKV;
  p.t_hi = (__half*)t_hi; p.t_lo = (__half*)t_lo; p.ldt = ldt; p.t_scale = t_scale;
  const int nt = (T + BQ - 1) / BQ;
  p.qt0 = q_row0 / BQ; p.nqt = nt - p.qt0;
  {
    static int order = -1;                   // env AS_HEADMEAN_ORDER=image: per-image tile order (experiment)
    if (order < 0) { const char* e = getenv("AS_HEADMEAN_ORDER"); order = (e && e[0] == 'i') ? 1 : 0; }
    p.per_image = order;
  }
  if (rowsum_slices == 4) {
    if (ld % 4 || ld < nt * BKV) return AS_ERR_BAD_ARG;   // float4 row stores, whole 128-column tiles
    const int n_tiles = p.nqt * nt * B;
    const int per = (n_tiles + num_sms - 1) / num_sms;
    const int grid = (n_tiles + per - 1) / per;
    attn_headmean2_kernel<<<grid, HM_THREADS, HM2_SMEM, stream>>>(tm_q, tm_k, p, n_tiles, per);
  } else if (rowsum_slices == 1) {
    attn_headmean_kernel<<<dim3(nt, nt, B), HM_THREADS, HM_SMEM, stream>>>(tm_q, tm_k, p);
  } else {
    return AS_ERR_BAD_ARG;
  }
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" int as_attn_headmean(const void* q, const void* k, const float* m, const float* l, float* out, int ld,
                                float* rowsum_part, int rowsum_slices, void* t_hi, void* t_lo, int ldt, float t_scale, int B,
                                int T, int heads, cudaStream_t stream) {
  if (!out) return AS_ERR_BAD_ARG;
  return as_attn_headmean_ex(q, k, m, l, out, ld, rowsum_part, rowsum_slices, t_hi, t_lo, ldt, t_scale, B, T, heads, 0, stream);
}
