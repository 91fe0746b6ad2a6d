// Attention backward on tcgen05 (what DDP training of the backbone needs; reference forward VT:74-86, whose
// backward torch autograd derives from the materialised [B,h,T,T] matrix).  Flash-style: nothing of size T x T is stored; the
// probabilities are recomputed from (Q, K, m, l) exactly as the head-mean pass does: P = exp2(c * QK^T - m) / l.
//
//   dV = P^T dO          dP = dO V^T          delta_i = sum_d dO_id O_id          dS = P o (dP - delta) * scale
//   dK = dS^T Q          dQ = dS K
//
// Two kernels, both deterministic (no atomics): as_mhsa_bwd_dkv keeps one 128-key tile (K_j, V_j) resident and walks the query
// tiles, accumulating dK_j / dV_j in TMEM; as_mhsa_bwd_dq keeps one 128-query tile (Q_i, dO_i) resident, walks the key tiles and
// accumulates dQ_i in TMEM.  Per tile pair: S and dP on the tensor cores into TMEM, four softmax warps (thread = query row = TMEM
// lane) turn them into P and dS (fp16, written into shared memory in the 128-byte-swizzled operand layout), and the tensor
// cores consume those tiles straight from shared memory -- as an MN-major A operand for the transposed products (P^T, dS^T),
// K-major for dS K.  All B operands are K-major: the host passes Q^T, dO^T, K^T ([BH, 64, Tpad], like the forward's V^T) next
// to the row-major tensors.
// Pipelining: S / dP are produced per 64-key HALF (two N = 64 MMA groups with their own full / free barriers), so the tensor cores
// compute the next half (or the next tile's first half) while the softmax warps work on the current one, and the transposed
// products of tile i run in the shadow of tile i+1's first half.
#include "common.cuh"

using namespace asb;

namespace {

constexpr int BT = 128, HD = 64;
constexpr int TILE = BT * HD * 2;          // 16 KB: [128 rows x 64 halves], or a transposed tile as two [64 x 64] boxes
constexpr int PTILE = BT * BT * 2;         // 32 KB: [128 q x 128 k] fp16 as two 64-column blocks of 16 KB
constexpr int BWD_THREADS = 576;           // warp 0 TMA, warp 1 MMA, warps 2-17 softmax: TMEM lane quadrant = warp & 3, 32-column chunk =
                                           // (warp - 2) >> 2.  One warp per scheduler was latency-bound (a dependent chain of ~1.5 k
                                           // instructions per tile: 6 k cycles); four independent chunks per scheduler hide it.
constexpr int SOFTMAX_WARPS = 16;
constexpr uint32_t C_S = 0, C_DP = 128, C_ACC0 = 256, C_ACC1 = 320;

struct BwdParams {
  int T, heads, nt;                        // nt = ceil(T / 128)
  float scale_log2, scale;                 // head_dim^-0.5 * log2(e), head_dim^-0.5
  const float* m;                          // [BH, T] softmax offsets of the forward (log2 domain)
  const float* l;                          // [BH, T] softmax denominators
  const float* delta;                      // [BH, T] rowsum(dO o O)
  float* out0;                             // dkv: dK [BH, T, 64];  dq: dQ [BH, T, 64]
  float* out1;                             // dkv: dV [BH, T, 64]
  __half* qkv16;                           // optional: dQ | dK | dV as fp16 rows of the qkv Linear's output [B*T, 3*heads*64]
};

// 32 gradient values of one (token, head) -> fp16, at their place in the qkv Linear's output row (VT:76: column =
// which * C + head * 64 + d), so that the projection's dX / dW GEMMs read them without a layout pass
__device__ __forceinline__ void store_qkv16(const BwdParams& p, const uint32_t (&a)[32], int bh, int t, int which, int half32) {
  const int b = bh / p.heads, h = bh - b * p.heads, C = p.heads * HD;
  uint4* dst = reinterpret_cast<uint4*>(p.qkv16 + ((size_t)b * p.T + t) * (3 * C) + which * C + h * HD + half32 * 32);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hh = __floats2half2_rn(__uint_as_float(a[8 * g + 2 * e]), __uint_as_float(a[8 * g + 2 * e + 1]));
      w[e] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    dst[g] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ int sw128p(int row, int col) {      // byte offset of (row, col) in a [rows x 64 halves] swizzled block
  return row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1);
}

// P and dS for one 32-key CHUNK of a (query tile, key tile) pair from the S / dP accumulators of this thread's query row.
// kWantP: also write P (the dkv kernel needs both tiles, the dq kernel only dS).
template <bool kWantP>
__device__ __forceinline__ void make_p_ds_chunk(uint32_t tm_row, uint8_t* p_s, uint8_t* ds_s, int ch, int row, bool row_ok, int k0,
                                                int T, float c, float scale, float mi, float inv_l, float di, uint64_t* p_free,
                                                uint32_t p_free_parity) {
  const int half = ch >> 1;
  {
    uint32_t s[32], dp[32];
    tmem_ld_32x32(tm_row + C_S + ch * 32, s);
    tmem_ld_32x32(tm_row + C_DP + ch * 32, dp);
    tc_wait_ld();
    uint32_t pw[16], dw[16];
    const bool full_tile = k0 + ch * 32 + 32 <= T;               // every column of the chunk is a real key (all tiles but the last)
    if (row_ok && full_tile) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), c, -mi)) * inv_l;
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), c, -mi)) * inv_l;
        const float d0 = p0 * (__uint_as_float(dp[2 * i]) - di) * scale, d1 = p1 * (__uint_as_float(dp[2 * i + 1]) - di) * scale;
        const __half2 ph = __floats2half2_rn(p0, p1), dh = __floats2half2_rn(d0, d1);
        pw[i] = *reinterpret_cast<const uint32_t*>(&ph);
        dw[i] = *reinterpret_cast<const uint32_t*>(&dh);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float pv[2], dv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = ch * 32 + 2 * i + e;
          const bool ok = row_ok && (k0 + col < T);
          const float p = ok ? ex2_approx(fmaf(__uint_as_float(s[2 * i + e]), c, -mi)) * inv_l : 0.f;
          pv[e] = p;
          dv[e] = p * (__uint_as_float(dp[2 * i + e]) - di) * scale;
        }
        const __half2 ph = __floats2half2_rn(pv[0], pv[1]), dh = __floats2half2_rn(dv[0], dv[1]);
        pw[i] = *reinterpret_cast<const uint32_t*>(&ph);
        dw[i] = *reinterpret_cast<const uint32_t*>(&dh);
      }
    }
    mbar_wait(p_free, p_free_parity);                          // the previous tile's P / dS have been consumed by the tensor cores
    // 32 columns = 64 bytes = four 16-byte chunks of the row's 128-byte swizzled line in block `half`
    const int cbase = (ch & 1) * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int off = half * (PTILE / 2) + sw128p(row, cbase + g * 8);
      if (kWantP) *reinterpret_cast<uint4*>(p_s + off) = make_uint4(pw[4 * g], pw[4 * g + 1], pw[4 * g + 2], pw[4 * g + 3]);
      *reinterpret_cast<uint4*>(ds_s + off) = make_uint4(dw[4 * g], dw[4 * g + 1], dw[4 * g + 2], dw[4 * g + 3]);
    }
  }
}

// ------------------------------------------------------------------ dK, dV: one key tile per CTA, loop over the query tiles
constexpr int DKV_STAGE = 4 * TILE;                                   // Q_i, dO_i, Q^T_i, dO^T_i
constexpr int DKV_SMEM = 1024 + 2 * TILE + 2 * DKV_STAGE + 2 * PTILE + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
mhsa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_qt, const __grid_constant__ CUtensorMap tm_dot, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* k_s = smem;                       // [128 k x 64 d]
  uint8_t* v_s = smem + TILE;                // [128 k x 64 d]
  uint8_t* ring = smem + 2 * TILE;           // 2 stages
  uint8_t* p_s = ring + 2 * DKV_STAGE;       // P  [128 q x 128 k]
  uint8_t* ds_s = p_s + PTILE;               // dS [128 q x 128 k]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + PTILE);
  uint64_t* kv_full = bars;                  // 1
  uint64_t* full = bars + 1;                 // 2
  uint64_t* empty = bars + 3;                // 2
  uint64_t* s_full = bars + 5;               // 2: S, dP of a 64-key half ready (commit)
  uint64_t* s_free = bars + 7;               // 2: that half read (4 warps)
  uint64_t* p_full = bars + 9;               // P, dS written (4 warps)
  uint64_t* p_free = bars + 10;              // P, dS consumed (commit)
  uint64_t* acc_full = bars + 11;            // all MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, bh = blockIdx.z * p.heads + blockIdx.y;
  const int k0 = jt * BT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_qt); tma_prefetch_desc(&tm_dot);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], SOFTMAX_WARPS / 2); }
    mbar_init(p_full, SOFTMAX_WARPS); mbar_init(p_free, 1); mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TILE);
      tma_load_3d(k_s, &tm_k, kv_full, 0, k0, bh);
      tma_load_3d(v_s, &tm_v, kv_full, 0, k0, bh);
      for (int it = 0; it < p.nt; ++it) {
        const int st = it & 1;
        mbar_wait(&empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&full[st], DKV_STAGE);
        uint8_t* d = ring + st * DKV_STAGE;
        tma_load_3d(d, &tm_q, &full[st], 0, it * BT, bh);
        tma_load_3d(d + TILE, &tm_do, &full[st], 0, it * BT, bh);
        tma_load_3d(d + 2 * TILE, &tm_qt, &full[st], it * BT, 0, bh);
        tma_load_3d(d + 2 * TILE + TILE / 2, &tm_qt, &full[st], it * BT + 64, 0, bh);
        tma_load_3d(d + 3 * TILE, &tm_dot, &full[st], it * BT, 0, bh);
        tma_load_3d(d + 3 * TILE + TILE / 2, &tm_dot, &full[st], it * BT + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_h = umma_idesc(0, BT, 64);                       // S, dP of one half: M = 128 q, N = 64 k
    constexpr uint32_t idesc_t = umma_idesc(0, BT, HD) | (1u << 15);          // dV, dK: M = 128 k (A MN-major), N = 64 d
    mbar_wait(kv_full, 0);
    // S / dP of (tile, half): Q_i K_j[half]^T and dO_i V_j[half]^T into their 64-column slots
    auto issue_half = [&](int it, int half) {
      const uint32_t q_a = smem_u32(ring + (it & 1) * DKV_STAGE), do_a = q_a + TILE;
      const uint32_t kh = smem_u32(k_s) + half * (TILE / 2), vh = smem_u32(v_s) + half * (TILE / 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + C_S + half * 64, umma_desc_k_sw128(q_a + k * 32), umma_desc_k_sw128(kh + k * 32), idesc_h, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + C_DP + half * 64, umma_desc_k_sw128(do_a + k * 32), umma_desc_k_sw128(vh + k * 32), idesc_h, k != 0);
        tc_commit(&s_full[half]);
      }
      __syncwarp();
    };
    mbar_wait(&full[0], 0);
    tc_fence_after();
    issue_half(0, 0);
    issue_half(0, 1);
    for (int it = 0; it < p.nt; ++it) {
      const int st = it & 1;
      const bool more = it + 1 < p.nt;
      if (more) {                                              // first half of the next tile as soon as this tile's first half is read
        mbar_wait(&full[st ^ 1], ((it + 1) >> 1) & 1);
        mbar_wait(&s_free[0], it & 1);
        tc_fence_after();
        issue_half(it + 1, 0);
      }
      if (more) {                                              // second half of the next tile BEFORE this tile's transposed products:
        mbar_wait(&s_free[1], it & 1);                         // its softmax warps start 512 tensor cycles earlier, and dV / dK run
        tc_fence_after();                                      // in the shadow of their exponentials
        issue_half(it + 1, 1);
      }
      mbar_wait(p_full, it & 1);
      tc_fence_after();
      const uint32_t q_a = smem_u32(ring + st * DKV_STAGE), qt_a = q_a + 2 * TILE, dot_a = q_a + 3 * TILE;
      if (elect_one()) {
        const uint32_t pa = smem_u32(p_s), da = smem_u32(ds_s);
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {                    // K = 16 query rows per MMA
          const uint32_t boff = (k >> 2) * (TILE / 2) + (k & 3) * 32;   // transposed tiles: two [64 d x 64 q] boxes
          mma_f16_ss(tmem + C_ACC1, umma_desc_mn_sw128(pa + k * 2048, PTILE / 2), umma_desc_k_sw128(dot_a + boff), idesc_t, (it | k) != 0);
          mma_f16_ss(tmem + C_ACC0, umma_desc_mn_sw128(da + k * 2048, PTILE / 2), umma_desc_k_sw128(qt_a + boff), idesc_t, (it | k) != 0);
        }
        tc_commit(p_free);
        tc_commit(&empty[st]);                                 // everything that read stage `st` has been issued before this commit
        if (!more) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3, row = quad * 32 + lane, ch = (warp - 2) >> 2, half = ch >> 1;
    const uint32_t tm_row = tmem + ((uint32_t)(quad * 32) << 16);
    float n_m = 0.f, n_l = 1.f, n_d = 0.f;                       // the next tile's row statistics, loaded one tile ahead
    if (row < p.T) { const size_t s0 = (size_t)bh * p.T + row; n_m = p.m[s0]; n_l = p.l[s0]; n_d = p.delta[s0]; }
    for (int it = 0; it < p.nt; ++it) {
      const int t = it * BT + row;
      const bool ok = t < p.T;
      const float mi = n_m, inv_l = 1.f / n_l, di = n_d;
      if (t + BT < p.T) { const size_t s1 = (size_t)bh * p.T + t + BT; n_m = p.m[s1]; n_l = p.l[s1]; n_d = p.delta[s1]; }
      mbar_wait(&s_full[half], it & 1);
      tc_fence_after();
      make_p_ds_chunk<true>(tm_row, p_s, ds_s, ch, row, ok, k0, p.T, p.scale_log2, p.scale, mi, inv_l, di, p_free, (it & 1) ^ 1);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_free[half]); mbar_arrive(p_full); }
    }
    // epilogue: thread = key row; the four chunk groups take dK[:, 0:32], dK[:, 32:64], dV[:, 0:32], dV[:, 32:64]
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int kr = k0 + row;
    {
      const int hf = ch & 1;
      uint32_t a[32];
      tmem_ld_32x32(tm_row + (ch < 2 ? C_ACC0 : C_ACC1) + hf * 32, a);
      tc_wait_ld();
      if (kr < p.T) {
        if (p.qkv16) store_qkv16(p, a, bh, kr, ch < 2 ? 1 : 2, hf);
        else {
          float4* dst = reinterpret_cast<float4*>((ch < 2 ? p.out0 : p.out1) + ((size_t)bh * p.T + kr) * HD + hf * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            dst[g] = make_float4(__uint_as_float(a[4 * g]), __uint_as_float(a[4 * g + 1]), __uint_as_float(a[4 * g + 2]), __uint_as_float(a[4 * g + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ dQ: one query tile per CTA, loop over the key tiles
constexpr int DQ_STAGE = 3 * TILE;                                    // K_j, V_j, K^T_j
constexpr int DQ_SMEM = 1024 + 2 * TILE + 2 * DQ_STAGE + PTILE + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
mhsa_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_kt, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* do_s = smem + TILE;
  uint8_t* ring = smem + 2 * TILE;
  uint8_t* ds_s = ring + 2 * DQ_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + PTILE);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = bars + 3;
  uint64_t* s_full = bars + 5;               // 2
  uint64_t* s_free = bars + 7;               // 2
  uint64_t* p_full = bars + 9;
  uint64_t* p_free = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.z * p.heads + blockIdx.y;
  const int q0 = qt * BT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kt);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], SOFTMAX_WARPS / 2); }
    mbar_init(p_full, SOFTMAX_WARPS); mbar_init(p_free, 1); mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE);
      tma_load_3d(q_s, &tm_q, q_full, 0, q0, bh);
      tma_load_3d(do_s, &tm_do, q_full, 0, q0, bh);
      for (int it = 0; it < p.nt; ++it) {
        const int st = it & 1;
        mbar_wait(&empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&full[st], DQ_STAGE);
        uint8_t* d = ring + st * DQ_STAGE;
        tma_load_3d(d, &tm_k, &full[st], 0, it * BT, bh);
        tma_load_3d(d + TILE, &tm_v, &full[st], 0, it * BT, bh);
        tma_load_3d(d + 2 * TILE, &tm_kt, &full[st], it * BT, 0, bh);
        tma_load_3d(d + 2 * TILE + TILE / 2, &tm_kt, &full[st], it * BT + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_h = umma_idesc(0, BT, 64);                       // S, dP of one half: M = 128 q, N = 64 k
    constexpr uint32_t idesc_q = umma_idesc(0, BT, HD);                       // dQ: M = 128 q, N = 64 d, both K-major
    mbar_wait(q_full, 0);
    auto issue_half = [&](int it, int half) {
      const uint32_t k_a = smem_u32(ring + (it & 1) * DQ_STAGE) + half * (TILE / 2), v_a = k_a + TILE;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + C_S + half * 64, umma_desc_k_sw128(smem_u32(q_s) + k * 32), umma_desc_k_sw128(k_a + k * 32), idesc_h, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_f16_ss(tmem + C_DP + half * 64, umma_desc_k_sw128(smem_u32(do_s) + k * 32), umma_desc_k_sw128(v_a + k * 32), idesc_h, k != 0);
        tc_commit(&s_full[half]);
      }
      __syncwarp();
    };
    mbar_wait(&full[0], 0);
    tc_fence_after();
    issue_half(0, 0);
    issue_half(0, 1);
    for (int it = 0; it < p.nt; ++it) {
      const int st = it & 1;
      const bool more = it + 1 < p.nt;
      if (more) {
        mbar_wait(&full[st ^ 1], ((it + 1) >> 1) & 1);
        mbar_wait(&s_free[0], it & 1);
        tc_fence_after();
        issue_half(it + 1, 0);
      }
      if (more) {
        mbar_wait(&s_free[1], it & 1);
        tc_fence_after();
        issue_half(it + 1, 1);
      }
      mbar_wait(p_full, it & 1);
      tc_fence_after();
      const uint32_t kt_a = smem_u32(ring + st * DQ_STAGE) + 2 * TILE;
      if (elect_one()) {
        const uint32_t da = smem_u32(ds_s);
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {                    // K = 16 keys per MMA
          const uint32_t aoff = (k >> 2) * (PTILE / 2) + (k & 3) * 32;
          const uint32_t boff = (k >> 2) * (TILE / 2) + (k & 3) * 32;
          mma_f16_ss(tmem + C_ACC0, umma_desc_k_sw128(da + aoff), umma_desc_k_sw128(kt_a + boff), idesc_q, (it | k) != 0);
        }
        tc_commit(p_free);
        tc_commit(&empty[st]);
        if (!more) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3, row = quad * 32 + lane, ch = (warp - 2) >> 2, half = ch >> 1;
    const uint32_t tm_row = tmem + ((uint32_t)(quad * 32) << 16);
    const int t = q0 + row;
    const bool ok = t < p.T;
    const size_t si = (size_t)bh * p.T + (ok ? t : 0);
    const float mi = p.m[si], inv_l = 1.f / p.l[si], di = p.delta[si];
    for (int it = 0; it < p.nt; ++it) {
      mbar_wait(&s_full[half], it & 1);
      tc_fence_after();
      make_p_ds_chunk<false>(tm_row, nullptr, ds_s, ch, row, ok, it * BT, p.T, p.scale_log2, p.scale, mi, inv_l, di, p_free, (it & 1) ^ 1);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_free[half]); mbar_arrive(p_full); }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (ch < 2) {                                               // thread = query row; chunk groups 0 / 1 take dQ[:, 0:32] / dQ[:, 32:64]
      uint32_t a[32];
      tmem_ld_32x32(tm_row + C_ACC0 + ch * 32, a);
      tc_wait_ld();
      if (ok) {
        if (p.qkv16) store_qkv16(p, a, bh, t, 0, ch);
        else {
          float4* dq = reinterpret_cast<float4*>(p.out0 + ((size_t)bh * p.T + t) * HD + ch * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            dq[g] = make_float4(__uint_as_float(a[4 * g]), __uint_as_float(a[4 * g + 1]), __uint_as_float(a[4 * g + 2]), __uint_as_float(a[4 * g + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int enc_rows(CUtensorMap* tm, const void* base, int BH, int T) {           // [BH, T, 64] fp16, box 128 rows
  uint64_t dims[3] = {HD, (uint64_t)T, (uint64_t)BH};
  uint64_t str[2] = {HD * 2, (uint64_t)T * HD * 2};
  uint32_t box[3] = {HD, BT, 1};
  return as_encode_tmap(tm, base, 2, 3, dims, str, box);
}
int enc_transposed(CUtensorMap* tm, const void* base, int BH, int Tpad) {  // [BH, 64, Tpad] fp16, box [64 rows x 64 cols]
  uint64_t dims[3] = {(uint64_t)Tpad, HD, (uint64_t)BH};
  uint64_t str[2] = {(uint64_t)Tpad * 2, (uint64_t)Tpad * HD * 2};
  uint32_t box[3] = {64, HD, 1};
  return as_encode_tmap(tm, base, 2, 3, dims, str, box);
}

}  // namespace

// Backward of as_mhsa_fwd.  q, k, v, d_o: [B, heads, T, 64] fp16 (head-major rows); qt, kt, dot: the same tensors transposed,
// [B, heads, 64, Tpad] fp16 with zero padding (Tpad = T rounded up to 128); m, l: the forward's row statistics [B, heads, T];
// delta [B, heads, T] = rowsum(dO o O).  Outputs dq, dk, dv [B, heads, T, 64] fp32 (gradients w.r.t. the UNSCALED q, k: the
// head_dim^-0.5 factor of VT:79 is applied inside).
// as_mhsa_bwd_ex: ``dqkv16`` non-null -> the three gradients are written as fp16 into one [B*T, 3*heads*64] tensor laid out
// like the qkv Linear's output (the operand of its dX / dW GEMMs) and dq / dk / dv are not touched.
extern "C" int as_mhsa_bwd_ex(const void* q, const void* k, const void* v, const void* d_o, const void* qt, const void* kt,
                              const void* dot, const float* m, const float* l, const float* delta, float* dq, float* dk,
                              float* dv, void* dqkv16, int B, int T, int Tpad, int heads, cudaStream_t stream) {
  if (Tpad % BT || Tpad < T || T < 1 || (!dqkv16 && (!dq || !dk || !dv))) return AS_ERR_BAD_ARG;
  const int BH = B * heads;
  CUtensorMap tq, tk, tv, tdo, tqt, tkt, tdot;
  int r = enc_rows(&tq, q, BH, T);
  if (!r) r = enc_rows(&tk, k, BH, T);
  if (!r) r = enc_rows(&tv, v, BH, T);
  if (!r) r = enc_rows(&tdo, d_o, BH, T);
  if (!r) r = enc_transposed(&tqt, qt, BH, Tpad);
  if (!r) r = enc_transposed(&tkt, kt, BH, Tpad);
  if (!r) r = enc_transposed(&tdot, dot, BH, Tpad);
  if (r) return r;
  static bool attr = false;
  if (!attr) {
    AS_CUDA(cudaFuncSetAttribute(mhsa_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    AS_CUDA(cudaFuncSetAttribute(mhsa_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    attr = true;
  }
  BwdParams p;
  p.T = T; p.heads = heads; p.nt = (T + BT - 1) / BT;
  p.scale = 0.125f; p.scale_log2 = (float)(0.125 * 1.4426950408889634);
  p.m = m; p.l = l; p.delta = delta; p.qkv16 = (__half*)dqkv16;
  const dim3 grid(p.nt, heads, B);
  p.out0 = dk; p.out1 = dv;
  mhsa_bwd_dkv_kernel<<<grid, BWD_THREADS, DKV_SMEM, stream>>>(tq, tk, tv, tdo, tqt, tdot, p);
  p.out0 = dq; p.out1 = nullptr;
  mhsa_bwd_dq_kernel<<<grid, BWD_THREADS, DQ_SMEM, stream>>>(tq, tk, tv, tdo, tkt, p);
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" int as_mhsa_bwd(const void* q, const void* k, const void* v, const void* d_o, const void* qt, const void* kt,
                           const void* dot, const float* m, const float* l, const float* delta, float* dq, float* dk, float* dv,
                           int B, int T, int Tpad, int heads, cudaStream_t stream) {
  return as_mhsa_bwd_ex(q, k, v, d_o, qt, kt, dot, m, l, delta, dq, dk, dv, nullptr, B, T, Tpad, heads, stream);
}
