// Attention-shift loop, second generation (as_mean_shift_v2): ONE persistent cooperative kernel for all iterations.
// Reference: cosine_shift_batch RH:830-854 + update_density_batch RH:882-908 (RH = stdroi_point_deform_attn_reppoints.py).
//
// What changed against the first persistent kernel (attnshift_meanshift_fused.cu, round 1) and why
// (profiles/l2stream_r2.txt, profiles/microbench_meanshift_r1z.txt):
//   * one CTA = 128 tokens (ONE M=128 MMA tile) instead of 256, sized so that TWO CTAs share an SM when an image has at most
//     64 seed columns (<= 112 KB shared memory, 256 TMEM columns each).  A single TMA-issuing thread per SM sustained ~75 GB/s
//     (the round-1 kernel streamed at 6 TB/s chip-wide, L2 delivers 17 TB/s to two CTAs per SM), and -- more important -- a third
//     of the round-1 time was latency-bound glue between the two streams of an iteration (group barriers, small reductions,
//     the assignment): with two CTAs of DIFFERENT images on an SM one streams while the other one waits.
//   * similarities stay in TMEM for the whole iteration (masked / scaled in place by the affinity epilogue); the column
//     statistics are butterfly reductions over the warp's registers (no [tokens x seeds] shared-memory copy): this is what
//     makes the shared-memory budget fit, and it lifts the 64-column limit: KP = 64, 128 or 256 seed columns per image.
//   * the prototype accumulators of the update live in TMEM only per 128-channel block (double buffered): any C % 128 == 0
//     (ViT-L: 1024), and the partial-prototype write-out of block b overlaps the MMAs of block b + 1.
//
// Work layout: an image's N tokens are U = ceil(N/64) units of 64; its G = ceil(U/2) CTAs take units {q, G + q} (round-robin:
// every CTA sees two distant rows of the image, box-shaped instance masks load the group evenly).  The CTAs of an image meet at
// a global counter four times per iteration; all reductions are ordered -> deterministic.
//
// Per iteration and CTA (warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = workers, thread = token = TMEM lane):
//   A  affinity   TMA streams [128 tok x 64 ch] (hi, lo) tiles and the image's seed tiles p^ (split fp16, KP rows); tcgen05
//                 hi.hi + hi.lo + lo.hi -> TMEM [128 x KP]; epilogue masks (box of the seed's instance) / scales in place,
//                 column max + density partials (sum, count over the previous assignment) -> global
//   -- barrier --  A2 tau, logit max of every seed; partial softmax denominators -> global
//   -- barrier --  B  1/Z, per token and instance arg-max seed (first wins) + weight; sparse weight tiles (weight x |f|, scaled by a
//                 per-seed power of two, split fp16) in shared memory; update D[ch, seed] += F^T[ch, tok] . W[seed, tok] on tcgen05
//                 from the same token tiles addressed MN-major, per 128-channel block; partial prototypes -> global
//   -- barrier --  C  ordered sum of the G partials -> new prototypes (fp32, returned) + normalised split-fp16 seeds p^
//   -- barrier --
// and one more (unmasked) affinity pass for the returned similarity maps.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>
#include <algorithm>

using namespace asb;

namespace {

constexpr int TOKC = 128;                // tokens per CTA
constexpr int STAGE = 32768;             // [128 tok x 64 ch] hi + lo  ==  [64 tok x 128 ch] hi + lo
constexpr int V2_THREADS = 192;
constexpr float OP_SCALE = 1024.f;

template <int KP> struct Cfg {
  static constexpr int NA = KP == 128 ? 3 : 2;               // token ring stages
  static constexpr int BTILE = KP * 256;                     // seed tile of one 64-channel block: hi KP*128 B + lo KP*128 B
  static constexpr int NBUF = KP == 256 ? 1 : 2;             // update accumulator buffers in TMEM
  static constexpr int TM_COLS = KP == 64 ? 256 : 512;
  static constexpr int MAXOBJ = KP == 64 ? 8 : 16;
  static constexpr int MISC = MAXOBJ * TOKC * 5 + TOKC * 4 + 4 * KP * 3 * 4 + KP * 6 * 4 + 2 * KP + 512;
  static constexpr int SMEM = 1024 + NA * STAGE + 2 * BTILE + MISC;
};

struct Box2 { int r0, r1, c0, c1; };
__device__ __forceinline__ Box2 patch_box2(const float* roi, int hp, int wp) {   // box2mask(rois // 16), RH:303-309
  Box2 b;
  b.c0 = (int)floorf(roi[0] / 16.f); b.r0 = (int)floorf(roi[1] / 16.f);
  b.c1 = (int)(floorf(roi[2] / 16.f) + 1.f); b.r1 = (int)(floorf(roi[3] / 16.f) + 1.f);
  b.c0 = max(0, min(b.c0, wp)); b.c1 = max(0, min(b.c1, wp));
  b.r0 = max(0, min(b.r0, hp)); b.r1 = max(0, min(b.r1, hp));
  return b;
}
__device__ __forceinline__ bool in_box2(const Box2& b, int n, int wp) {
  const int r = n / wp, c = n - r * wp;
  return r >= b.r0 && r < b.r1 && c >= b.c0 && c < b.c1;
}

struct V2Params {
  int n_img, N, C, hp, wp, S, G, n_shift, clamp0, n_tot;
  float tt0, temp;
  const int* img_first; const int* img_nobj; const float* rois;
  const float* den;            // [n_img][N]   |f| (clamped at 1e-8)
  float* proto;                // [n_tot][S][C] in/out
  __half* phat_hi;             // [n_img][KP][C] normalised seeds * 2^10, split fp16 (scratch; read back through TMA)
  __half* phat_lo;
  float* sim_out;              // [n_tot][S][N]
  int* trace;                  // [n_shift][n_tot][N] or null
  float* colmax_part;          // [n_img][G][KP]
  float* dens_part;            // [n_img][G][KP][2]
  float* z_part;               // [n_img][G][KP]
  float* proto_part;           // [n_img][G][KP][C]
  unsigned* bar;               // [n_img] monotonic group-barrier counters
  unsigned* sched;             // [2 + #SMs] work tickets of the two halves of the image groups, CTAs seen per SM
  unsigned long long* dbg;     // optional [grid][16] accumulated ns per phase
  unsigned stagger_ns;         // start delay of the second half of the image groups (see the kernel)
};

__device__ __forceinline__ void group_barrier2(unsigned* ctr, unsigned target) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const uint64_t t0 = global_timer_ns();
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      __nanosleep(64);
      if ((++spins & 0xff) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  tc_fence_after();
}
__device__ __forceinline__ void workers_sync2() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// byte offset of element (row, col) inside a K-major [rows x 64 halves] tile with the 128-byte swizzle
__device__ __forceinline__ int sw128b(int row, int col) { return row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1); }

__device__ __forceinline__ float sum_partials2(const float* base, size_t stride, int G) {
  float acc = 0.f;
  for (int g0 = 0; g0 < G; g0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (g0 + i < G) ? __ldcg(base + (size_t)(g0 + i) * stride) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i];
  }
  return acc;
}

// Column reductions of a [32 lanes x 32 columns] register tile: afterwards lane L holds the reduction of column L over the
// warp's 32 lanes.  Each step halves the number of live columns per lane (31 shuffles in total), fixed order: deterministic.
template <bool kMax>
__device__ __forceinline__ float warp_col_reduce(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? a[i + n / 2] : a[i];
      const float send = up ? a[i] : a[i + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      a[i] = kMax ? fmaxf(keep, recv) : keep + recv;
    }
  }
  return a[0];
}

template <int KP>
// KP = 64: two CTAs per SM.  Registers are granted per 4-warp granule, so the 6 warps of a CTA count as 8: the cap must be
// 65536 / (2 x 256) = 128 per thread (launch bounds of 256 threads), not the 168 that 192 threads would suggest.
__global__ void __launch_bounds__(KP == 64 ? 256 : V2_THREADS, KP == 64 ? 2 : 1)
mean_shift_v2_kernel(const __grid_constant__ CUtensorMap tm_hi64, const __grid_constant__ CUtensorMap tm_lo64,
                     const __grid_constant__ CUtensorMap tm_phi, const __grid_constant__ CUtensorMap tm_plo,
                     const V2Params p) {
  using K = Cfg<KP>;
  constexpr int NA = K::NA, BTILE = K::BTILE, NBUF = K::NBUF, MAXOBJ = K::MAXOBJ;
  constexpr int NCH = KP / 32;                                  // 32-column chunks of a similarity row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                        // NA stages
  uint8_t* bw = smem + NA * STAGE;                             // phase A: 2 seed tiles; phase B: the 2 units' weight tiles
  uint8_t* misc = bw + 2 * BTILE;
  float* w_s = reinterpret_cast<float*>(misc);                 // [MAXOBJ][TOKC] weight of the assigned seed (0 outside the box)
  float* den_s = w_s + MAXOBJ * TOKC;                          // [TOKC]
  float* red_s = den_s + TOKC;                                 // [4 warps][3][KP]
  float* st_s = red_s + 4 * 3 * KP;                            // [KP][4] 1/tt, column max, 1/Z, tau
  float* sc_s = st_s + KP * 4;                                 // [KP][2] weight scale 2^k of the seed, 2^-k / OP_SCALE
  int8_t* idx_s = reinterpret_cast<int8_t*>(sc_s + KP * 2);    // [MAXOBJ][TOKC] assigned seed of the previous iteration
  uint8_t* cj_s = reinterpret_cast<uint8_t*>(idx_s + MAXOBJ * TOKC);       // [KP] instance of a seed column (255: unused column)
  uint8_t* cs_s = cj_s + KP;                                   // [KP] seed index of the column inside its instance
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(cs_s + KP) + 7) & ~(uintptr_t)7);
  uint64_t* a_full = bars;             // NA (<= 3)
  uint64_t* a_empty = bars + 3;        // NA
  uint64_t* b_full = bars + 6;         // 2
  uint64_t* b_empty = bars + 8;        // 2
  uint64_t* acc_full = bars + 10;      // 1
  uint64_t* w_full = bars + 11;        // 1 (weight tiles built: 4 worker warps arrive)
  uint64_t* upd_full = bars + 12;      // NBUF (accumulator of a channel block complete)
  uint64_t* upd_free = bars + 14;      // NBUF (accumulator drained: 4 worker warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3;                                   // TMEM lane quadrant this warp may access
  const bool worker = warp >= 2;
  const int tl = quad * 32 + lane;                             // worker: local token = TMEM lane = row of the MMA tile
  // Which (image group, rank) a CTA works on is decided at run time from WHERE it landed: the first CTA to arrive on an SM takes
  // a ticket of the first half of the image groups, the second one a ticket of the second half (falling back to the other half
  // when its own is exhausted), and the second half starts `stagger_ns` late.  So on every SM one CTA streams tokens (TMA /
  // tensor pipe) while the other is in the latency-bound part of its iteration (statistics, assignment, group barriers);
  // without the offset all images march in lock step and the co-resident CTAs only get in each other's way.
  const int groups = gridDim.x / p.G;
  const int groups_a = (groups + 1) / 2;
  __shared__ int s_ticket[2];
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned slot = atomicAdd(p.sched + 2 + smid, 1u);
    int half = (int)(slot & 1u);
    const int need[2] = {groups_a * p.G, (groups - groups_a) * p.G};
    int t = (int)atomicAdd(p.sched + half, 1u);
    if (t >= need[half]) { half ^= 1; t = (int)atomicAdd(p.sched + half, 1u); }
    s_ticket[0] = half; s_ticket[1] = t;
  }
  __syncthreads();
  const int grp = (s_ticket[0] ? groups_a : 0) + s_ticket[1] / p.G, q = s_ticket[1] % p.G;
  const int kblocks = p.C / 64;
  const int cblocks = p.C / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi64); tma_prefetch_desc(&tm_lo64);
    tma_prefetch_desc(&tm_phi); tma_prefetch_desc(&tm_plo);
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < NBUF; ++i) { mbar_init(&upd_full[i], 1); mbar_init(&upd_free[i], 4); }
    mbar_init(acc_full, 1);
    mbar_init(w_full, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<K::TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_row = tmem + ((uint32_t)(quad * 32) << 16);   // this warp's lanes
  constexpr uint32_t TM_UPD = KP;                                 // update accumulators follow the similarities

  int a_stage = 0; uint32_t a_phase = 0;        // producer + MMA: token ring
  uint32_t bcount = 0;                          // producer + MMA: seed tile uses so far
  uint32_t acount = 0;                          // workers: affinity passes so far
  uint32_t ucount = 0;                          // MMA + workers: channel blocks of the update so far
  uint32_t wcount = 0;                          // MMA: update phases so far

  if (p.stagger_ns && groups >= 2 && grp >= groups_a) {
    if (threadIdx.x == 0) {
      const uint64_t t0 = global_timer_ns();
      while (global_timer_ns() - t0 < p.stagger_ns) __nanosleep(256);
    }
    __syncthreads();
  }
  uint64_t t_prev = global_timer_ns();
  auto mark = [&](int k) {
    if (p.dbg && threadIdx.x == 64) { const uint64_t t = global_timer_ns(); p.dbg[blockIdx.x * 16 + k] += t - t_prev; t_prev = t; }
  };
  auto tok = [&](int t) { return ((t >> 6) * p.G + q) * 64 + (t & 63); };

  for (int img = grp; img < p.n_img; img += groups) {
    const int nobj = p.img_nobj[img], o0 = p.img_first[img];
    const int kb_cols = nobj * p.S;
    unsigned* ctr = p.bar + img;
    unsigned bar_target = 0;
    const int n_tok = tok(tl);                                 // global token of this worker thread
    const bool valid = worker && n_tok < p.N;
    float fmax_cta = 1.f;
    unsigned boxmask = 0;                                      // bit j: token inside instance j's box
    if (worker) {
      const float dn = valid ? p.den[(size_t)img * p.N + n_tok] : 1.f;
      den_s[tl] = dn;
      for (int j = 0; j < MAXOBJ; ++j) { idx_s[j * TOKC + tl] = -1; w_s[j * TOKC + tl] = 0.f; }
      for (int col = tl; col < KP; col += TOKC) {
        const int cj = col < kb_cols ? col / p.S : 255;
        cj_s[col] = (uint8_t)cj;
        cs_s[col] = (uint8_t)(col < kb_cols ? col - cj * p.S : 0);
      }
      if (valid)
        for (int j = 0; j < nobj; ++j)
          if (in_box2(patch_box2(p.rois + 4 * (o0 + j), p.hp, p.wp), n_tok, p.wp)) boxmask |= 1u << j;
      const float wm = warp_max(valid ? dn : 0.f);
      if (lane == 0) red_s[quad] = wm;
      workers_sync2();
      fmax_cta = fmaxf(fmaxf(red_s[0], red_s[1]), fmaxf(red_s[2], red_s[3]));
      fmax_cta = fmaxf(fmax_cta, 1e-8f);
      workers_sync2();
    }

    // new prototypes (ordered sum of the group's partials) and their normalised split-fp16 copy, rows q, q+G, ...
    auto phase_c = [&](bool from_partials) {
      if (worker) {
        for (int r = q; r < kb_cols; r += p.G) {
          float v[8];                                           // channels tl*4 + e*512 + {0..3}: float4 per 512-channel half
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          if (from_partials) {
            for (int g0 = 0; g0 < p.G; g0 += 8) {
              float4 t[2][8];
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int gg = 0; gg < 8; ++gg) {
                  const int c = tl * 4 + h * 512, g = g0 + gg;
                  t[h][gg] = (c < p.C && g < p.G)
                                 ? __ldcg(reinterpret_cast<const float4*>(p.proto_part + (((size_t)img * p.G + g) * KP + r) * p.C + c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int gg = 0; gg < 8; ++gg) {
                  v[h * 4 + 0] += t[h][gg].x; v[h * 4 + 1] += t[h][gg].y; v[h * 4 + 2] += t[h][gg].z; v[h * 4 + 3] += t[h][gg].w;
                }
            }
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = tl * 4 + h * 512;
              if (c < p.C) {
                const float4 t = *reinterpret_cast<const float4*>(p.proto + ((size_t)o0 * p.S + r) * p.C + c);
                v[h * 4 + 0] = t.x; v[h * 4 + 1] = t.y; v[h * 4 + 2] = t.z; v[h * 4 + 3] = t.w;
              }
            }
          }
          float ss = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
          if (from_partials) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = tl * 4 + h * 512;
              if (c < p.C)
                *reinterpret_cast<float4*>(p.proto + ((size_t)o0 * p.S + r) * p.C + c) = make_float4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
            }
          }
          ss = warp_sum(ss);
          if (lane == 0) red_s[quad] = ss;
          workers_sync2();
          const float tot = (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]);
          const float nrm = fmaxf(sqrtf(tot), 1e-8f);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = tl * 4 + h * 512;
            if (c < p.C) {
              __half hi4[4], lo4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = v[h * 4 + e] / nrm * OP_SCALE;
                hi4[e] = __float2half_rn(x);
                lo4[e] = __float2half_rn(x - __half2float(hi4[e]));
              }
              const size_t o = ((size_t)img * KP + r) * p.C + c;
              *reinterpret_cast<uint2*>(p.phat_hi + o) = *reinterpret_cast<uint2*>(hi4);
              *reinterpret_cast<uint2*>(p.phat_lo + o) = *reinterpret_cast<uint2*>(lo4);
            }
          }
          workers_sync2();
        }
        fence_proxy_async_all();                                // the next reader of phat_hi/lo is another CTA's TMA
      }
    };
    phase_c(false);
    bar_target += p.G;
    group_barrier2(ctr, bar_target);
    mark(0);

    for (int it = 0; it <= p.n_shift; ++it) {
      const bool last = (it == p.n_shift);                      // extra pass: unmasked similarities for the output
      // ================================================================ phase A: affinity
      if (warp == 0) {
        if (lane == 0) {
          fence_proxy_async_all();
          for (int kb = 0; kb < kblocks; ++kb) {
            const uint32_t bb = bcount & 1;
            mbar_wait(&b_empty[bb], ((bcount >> 1) & 1) ^ 1);
            mbar_expect_tx(&b_full[bb], BTILE);
            tma_load_3d(bw + bb * BTILE, &tm_phi, &b_full[bb], kb * 64, 0, img);
            tma_load_3d(bw + bb * BTILE + KP * 128, &tm_plo, &b_full[bb], kb * 64, 0, img);
            ++bcount;
            mbar_wait(&a_empty[a_stage], a_phase ^ 1);
            mbar_expect_tx(&a_full[a_stage], STAGE);
#pragma unroll
            for (int h = 0; h < 2; ++h) {                       // the 128-row operand tile = the CTA's two 64-token units
              const int row0 = (h * p.G + q) * 64;
              tma_load_3d(ring + a_stage * STAGE + h * 8192, &tm_hi64, &a_full[a_stage], kb * 64, row0, img);
              tma_load_3d(ring + a_stage * STAGE + 16384 + h * 8192, &tm_lo64, &a_full[a_stage], kb * 64, row0, img);
            }
            if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
          }
        }
      } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(0, 128, KP);
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint32_t bb = bcount & 1;
          mbar_wait(&b_full[bb], (bcount >> 1) & 1);
          mbar_wait(&a_full[a_stage], a_phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = smem_u32(ring + a_stage * STAGE), a_lo = a_hi + 16384;
            const uint32_t b_hi = smem_u32(bw + bb * BTILE), b_lo = b_hi + KP * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, (kb | k) != 0);
              mma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_lo + k * 32), idesc, 1);
              mma_f16_ss(tmem, umma_desc_k_sw128(a_lo + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, 1);
            }
            tc_commit(&a_empty[a_stage]);
            tc_commit(&b_empty[bb]);
            if (kb == kblocks - 1) tc_commit(acc_full);
          }
          __syncwarp();
          if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
          ++bcount;
        }
      } else {
        // ---- epilogue: mask / scale the similarities in place (TMEM), column statistics from the registers
        mbar_wait(acc_full, acount & 1);
        ++acount;
        tc_fence_after();
        mark(10);
        const float alpha = 1.f / (OP_SCALE * OP_SCALE);
        float* hv_s = w_s;                                      // value at the previously assigned seed, per (instance, token): w_s is dead here
        int hits = 0;                                           // hit columns seen so far = instance index of the next one
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          if (ch * 32 >= kb_cols) break;
          uint32_t raw[32];
          tmem_ld_32x32(tm_row + ch * 32, raw);
          // per-chunk bit masks instead of per-column bookkeeping (no branches, no counters in the 32-column loop):
          //   vmask: columns that exist; onmask: columns whose instance's box holds this token (every column on the last pass);
          //   hitmask: the column of each instance this token was assigned to in the previous iteration
          const int nc = kb_cols - ch * 32;
          const unsigned vmask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
          unsigned onmask = 0, hitmask = 0;
          if (valid) {
            if (last) onmask = vmask;
            for (int jj = 0; jj < nobj; ++jj) {
              const int c0 = jj * p.S - ch * 32, c1 = c0 + p.S;                  // the instance's columns, chunk-relative
              if (!last && ((boxmask >> jj) & 1u) && c1 > 0 && c0 < 32) {
                const unsigned lo = c0 <= 0 ? 0xffffffffu : (0xffffffffu << c0);
                const unsigned hi = c1 >= 32 ? 0xffffffffu : ((1u << c1) - 1u);
                onmask |= lo & hi;
              }
              if (!last && it > 0) {
                const int c = c0 + (int)idx_s[jj * TOKC + tl];
                if ((unsigned)c < 32u) hitmask |= 1u << c;
              }
            }
            onmask &= vmask;
          }
          const unsigned tvmask = valid ? vmask : 0u;
          tc_wait_ld();
          unsigned mymax = 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float v = ((onmask >> c) & 1u) ? __uint_as_float(raw[c]) * alpha : 0.f;
            raw[c] = __float_as_uint(v);
            if (!last) {
              // column maximum over the warp's 32 tokens: one REDUX on an order-preserving integer image of the float
              const unsigned b = __float_as_uint(v);
              const unsigned u = ((tvmask >> c) & 1u) ? (b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u)) : 0u;
              const unsigned m = __reduce_max_sync(0xffffffffu, u);
              if (lane == c) mymax = m;
              // a valid token has exactly one hit column per instance, in instance order: the k-th hit belongs to instance k
              if ((hitmask >> c) & 1u) hv_s[(hits + __popc(hitmask & ((1u << c) - 1u))) * TOKC + tl] = v;
            }
          }
          hits += __popc(hitmask);
          if (last) {
            // returned maps [o][s][n]: one coalesced row of this CTA's tokens per seed
            if (valid) {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const int col = ch * 32 + c;
                if (col < kb_cols) {
                  float v = __uint_as_float(raw[c]);
                  if (p.clamp0) v = fmaxf(v, 0.f);
                  p.sim_out[((size_t)o0 * p.S + col) * p.N + n_tok] = v;
                }
              }
            }
          } else {
            tmem_st_32x32(tm_row + ch * 32, raw);
            red_s[quad * KP + ch * 32 + lane] = __uint_as_float(mymax);     // still the integer image; decoded below
          }
        }
        if (!last) {
          tc_wait_st();
          tc_fence_before();
          workers_sync2();
          // per column: max over the 4 warps, and the density partials = ordered sum / count over this CTA's tokens of the
          // similarity at the seed the token was assigned to in the previous iteration (fixed token order: deterministic)
          for (int col = tl; col < kb_cols; col += TOKC) {
            unsigned um = __float_as_uint(red_s[col]);
#pragma unroll
            for (int w = 1; w < 4; ++w) um = max(um, __float_as_uint(red_s[w * KP + col]));
            const float mx = um == 0u ? -FLT_MAX : __uint_as_float(um ^ ((um >> 31) ? 0x80000000u : 0xffffffffu));
            float sv = 0.f, cv = 0.f;
            if (it > 0) {
              const int jc = col / p.S, sc = col - jc * p.S;
#pragma unroll 8
              for (int t = 0; t < TOKC; ++t) {
                const bool hit = idx_s[jc * TOKC + t] == sc;
                sv += hit ? hv_s[jc * TOKC + t] : 0.f;
                cv += hit ? 1.f : 0.f;
              }
            }
            const size_t pi = ((size_t)img * p.G + q) * KP + col;
            p.colmax_part[pi] = mx; p.dens_part[pi * 2] = sv; p.dens_part[pi * 2 + 1] = cv;
          }
        }
      }
      mark(1);
      if (last) break;
      bar_target += p.G;
      group_barrier2(ctr, bar_target);
      mark(2);
      // ================================================================ phase A2: per-seed statistics, partial Z
      if (worker) {
        for (int col = tl; col < kb_cols; col += TOKC) {
          const size_t p0 = (size_t)img * p.G * KP + col;
          float mx = -FLT_MAX;
          for (int g0 = 0; g0 < p.G; g0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (g0 + i < p.G) ? __ldcg(p.colmax_part + p0 + (size_t)(g0 + i) * KP) : -FLT_MAX;
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, v[i]);
          }
          float tt, tau = 0.f;
          if (it == 0) tt = p.tt0;
          else {
            const float tot = sum_partials2(p.dens_part + p0 * 2, (size_t)KP * 2, p.G);
            const float cnt = sum_partials2(p.dens_part + p0 * 2 + 1, (size_t)KP * 2, p.G);
            tau = fmaxf(1.f - (cnt >= 1.f ? tot / cnt : 0.f), 1e-10f);       // RH:883-885, 908
            tt = p.temp * tau;
          }
          st_s[col * 4] = 1.f / tt; st_s[col * 4 + 1] = mx; st_s[col * 4 + 3] = tau;
        }
        workers_sync2();
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          if (ch * 32 >= kb_cols) break;
          uint32_t raw[32];
          tmem_ld_32x32(tm_row + ch * 32, raw);
          tc_wait_ld();
          float e[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int col = ch * 32 + c;
            // logit - max logit as (v - max) / tt: exactly 0 at the maximum even when tt is ~1e-11 (tau clamped at 1e-10)
            e[c] = (valid && col < kb_cols) ? expf((__uint_as_float(raw[c]) - st_s[col * 4 + 1]) * st_s[col * 4]) : 0.f;
          }
          // the exponentials replace the similarities in TMEM: the assignment only needs weight = e / Z
#pragma unroll
          for (int c = 0; c < 32; ++c) raw[c] = __float_as_uint(e[c]);
          tmem_st_32x32(tm_row + ch * 32, raw);
          const float z = warp_col_reduce<false>(e, lane);
          red_s[quad * KP + ch * 32 + lane] = z;
        }
        tc_wait_st();
        tc_fence_before();
        workers_sync2();
        for (int col = tl; col < kb_cols; col += TOKC)
          p.z_part[((size_t)img * p.G + q) * KP + col] = (red_s[col] + red_s[KP + col]) + (red_s[2 * KP + col] + red_s[3 * KP + col]);
      }
      mark(3);
      bar_target += p.G;
      group_barrier2(ctr, bar_target);
      mark(4);
      // ================================================================ phase B: assign + update
      if (warp == 0) {
        if (lane == 0) {
          // channel blocks in DESCENDING order: the affinity pass before this one finished with the last channels and the one
          // after it starts with the first, so consecutive passes meet in whatever part of the token set the L2 still holds
          for (int cb = cblocks - 1; cb >= 0; --cb)
            for (int u = 0; u < 2; ++u) {
              mbar_wait(&a_empty[a_stage], a_phase ^ 1);
              mbar_expect_tx(&a_full[a_stage], STAGE);
              uint8_t* dst = ring + a_stage * STAGE;
              const int row0 = (u * p.G + q) * 64;
              tma_load_3d(dst, &tm_hi64, &a_full[a_stage], (2 * cb) * 64, row0, img);
              tma_load_3d(dst + 8192, &tm_hi64, &a_full[a_stage], (2 * cb + 1) * 64, row0, img);
              tma_load_3d(dst + 16384, &tm_lo64, &a_full[a_stage], (2 * cb) * 64, row0, img);
              tma_load_3d(dst + 24576, &tm_lo64, &a_full[a_stage], (2 * cb + 1) * 64, row0, img);
              if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
            }
        }
      } else if (warp == 1) {
        // D[channel (M = 128: two 64-channel tiles, LBO apart), seed (N = KP)] += F^T . W over the 64 tokens of the unit
        constexpr uint32_t idesc_u = umma_idesc(0, 128, KP) | (1u << 15);       // A is MN-major (channels contiguous)
        mbar_wait(w_full, wcount & 1);
        ++wcount;
        tc_fence_after();
        for (int cb = cblocks - 1; cb >= 0; --cb) {
          const uint32_t buf = ucount % NBUF, use = ucount / NBUF;
          mbar_wait(&upd_free[buf], (use & 1) ^ 1);
          tc_fence_after();
          for (int u = 0; u < 2; ++u) {
            mbar_wait(&a_full[a_stage], a_phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_hi = smem_u32(ring + a_stage * STAGE), a_lo = a_hi + 16384;
              const uint32_t w_hi = smem_u32(bw + u * BTILE), w_lo = w_hi + KP * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k) {                     // 16 tokens per MMA = two 8-token swizzle atoms (SBO = 1024 B)
                const uint64_t dah = umma_desc_mn_sw128(a_hi + k * 2048, 8192), dal = umma_desc_mn_sw128(a_lo + k * 2048, 8192);
                const uint64_t dwh = umma_desc_k_sw128(w_hi + k * 32), dwl = umma_desc_k_sw128(w_lo + k * 32);
                mma_f16_ss(tmem + TM_UPD + buf * KP, dah, dwh, idesc_u, (u | k) != 0);
                mma_f16_ss(tmem + TM_UPD + buf * KP, dah, dwl, idesc_u, 1);
                mma_f16_ss(tmem + TM_UPD + buf * KP, dal, dwh, idesc_u, 1);
              }
              tc_commit(&a_empty[a_stage]);
              if (u == 1) tc_commit(&upd_full[buf]);
            }
            __syncwarp();
            if (++a_stage == NA) { a_stage = 0; a_phase ^= 1; }
          }
          ++ucount;
        }
      } else {
        for (int col = tl; col < kb_cols; col += TOKC)
          st_s[col * 4 + 2] = 1.f / sum_partials2(p.z_part + (size_t)img * p.G * KP + col, KP, p.G);     // fixed order
        workers_sync2();
        tc_fence_after();
        {
          float best = -1.f;
          int bi = 0;
#pragma unroll 1
          for (int ch = 0; ch < NCH; ++ch) {
            if (ch * 32 >= kb_cols) break;
            uint32_t raw[32];
            tmem_ld_32x32(tm_row + ch * 32, raw);
            const int nc = kb_cols - ch * 32;
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              // branch-free running arg-max per instance (first maximum wins, like torch.argmax): restart at the instance's first
              // column, and write the running result after EVERY column -- the value left behind by the instance's last column is
              // the final one (no per-instance flush branch in the unrolled loop)
              const int col = min(ch * 32 + c, kb_cols - 1);
              const int sidx = cs_s[col], jv = cj_s[col];
              const float w = __uint_as_float(raw[c]) * st_s[col * 4 + 2];       // e (written by the statistics phase) / Z
              const bool first = sidx == 0;
              const float bprev = first ? -1.f : best;
              const int iprev = first ? 0 : bi;
              const bool better = w > bprev;
              best = better ? w : bprev;
              bi = better ? sidx : iprev;
              if (c < nc) {
                idx_s[jv * TOKC + tl] = valid ? (int8_t)bi : (int8_t)-1;          // density of the next iteration counts every token
                w_s[jv * TOKC + tl] = ((boxmask >> jv) & 1u) ? best : 0.f;        // masked tokens are zero vectors: they add nothing
                if (p.trace && valid && sidx == p.S - 1) p.trace[((size_t)it * p.n_tot + o0 + jv) * p.N + n_tok] = bi;
              }
            }
          }
        }
        tc_fence_before();
        // per-seed power-of-two scale: weight x |f| <= (1/Z) x fmax < 2^10 after scaling, so the fp16 hi + lo split keeps
        // ~22 bits of the weights that matter
        for (int col = tl; col < KP; col += TOKC) {
          float sc = 1.f;
          if (col < kb_cols) sc = exp2f((float)(9 - ilogbf(st_s[col * 4 + 2] * fmax_cta)));
          sc_s[col * 2] = sc;
          sc_s[col * 2 + 1] = 1.f / (sc * OP_SCALE);            // f = (hi + lo) * |f| / 2^10: |f| is folded into the weight
        }
        for (int i = tl; i < 2 * BTILE / 16; i += TOKC) reinterpret_cast<uint4*>(bw)[i] = make_uint4(0, 0, 0, 0);
        workers_sync2();
        {
          const int u = tl >> 6, tcol = tl & 63;
          for (int j = 0; j < nobj; ++j) {
            const float wv = w_s[j * TOKC + tl];
            if (wv != 0.f) {
              const int r = j * p.S + idx_s[j * TOKC + tl];
              const float v = wv * den_s[tl] * sc_s[r * 2];
              const __half h = __float2half_rn(v);
              uint8_t* dst = bw + u * BTILE + sw128b(r, tcol);
              *reinterpret_cast<__half*>(dst) = h;
              *reinterpret_cast<__half*>(dst + KP * 128) = __float2half_rn(v - __half2float(h));
            }
          }
        }
        fence_proxy_async();                                    // generic-proxy stores -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(w_full);
        mark(5);
        // ---- update epilogue per channel block: TMEM -> partial prototypes [seed][channel] (lane = channel: coalesced rows)
        for (int cb = cblocks - 1; cb >= 0; --cb) {
          const uint32_t buf = ucount % NBUF, use = ucount / NBUF;
          mbar_wait(&upd_full[buf], use & 1);
          tc_fence_after();
          float* dst = p.proto_part + ((size_t)img * p.G + q) * KP * p.C + cb * 128 + tl;
#pragma unroll 1
          for (int ch = 0; ch < NCH; ++ch) {
            if (ch * 32 >= kb_cols) break;
            uint32_t raw[32];
            tmem_ld_32x32(tm_row + TM_UPD + buf * KP + ch * 32, raw);
            float scl[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) scl[c] = sc_s[(ch * 32 + c) * 2 + 1];
            tc_wait_ld();
            const int ncol = min(32, kb_cols - ch * 32);
            float* d = dst + (size_t)(ch * 32) * p.C;           // running pointer: one 64-bit add per row instead of a wide multiply
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (c < ncol) *d = __uint_as_float(raw[c]) * scl[c];
              d += p.C;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&upd_free[buf]);
          ++ucount;
        }
      }
      mark(6);
      bar_target += p.G;
      group_barrier2(ctr, bar_target);
      mark(7);
      // ================================================================ phase C: new prototypes
      phase_c(true);
      mark(8);
      bar_target += p.G;
      group_barrier2(ctr, bar_target);
      mark(9);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<K::TM_COLS>(tmem);
}

__global__ void v2_split_tokens(const float* __restrict__ feats, long long fstride, int N, int C, __half* __restrict__ hi,
                                __half* __restrict__ lo, float* __restrict__ den) {
  const int img = blockIdx.y;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* f = feats + img * fstride + (long long)n * C;
  const int lane = lane_id();
  float ss = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 t = *reinterpret_cast<const float4*>(f + c);
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
  if (lane == 0) den[(size_t)img * N + n] = nrm;
  const size_t o = ((size_t)img * N + n) * C;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 t = *reinterpret_cast<const float4*>(f + c);       // second read hits L1 / L2
    const float x[4] = {t.x, t.y, t.z, t.w};
    __half h4[4], l4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = x[e] / nrm * OP_SCALE;
      h4[e] = __float2half_rn(v);
      l4[e] = __float2half_rn(v - __half2float(h4[e]));
    }
    *reinterpret_cast<uint2*>(hi + o + c) = *reinterpret_cast<uint2*>(h4);
    *reinterpret_cast<uint2*>(lo + o + c) = *reinterpret_cast<uint2*>(l4);
  }
}

int kp_for(int kmax, int max_obj) {
  if (kmax <= 64 && max_obj <= Cfg<64>::MAXOBJ) return 64;
  if (kmax <= 128 && max_obj <= Cfg<128>::MAXOBJ) return 128;
  if (kmax <= 256 && max_obj <= Cfg<256>::MAXOBJ) return 256;
  return 0;
}

int g_v2_plain_launches = 0;

template <int KP>
int launch_v2(const CUtensorMap* tm, V2Params& p, int num_sms, int G, cudaStream_t stream) {
  const size_t smem = Cfg<KP>::SMEM;
  AS_CUDA(cudaFuncSetAttribute(mean_shift_v2_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // without the carve-out preference the driver sizes the shared-memory partition for ONE CTA of this size and the second
  // CTA of the pair never becomes resident (occupancy 1: seen as a grid of 128 in the first ncu capture)
  AS_CUDA(cudaFuncSetAttribute(mean_shift_v2_kernel<KP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // Resident CTAs per SM from the hardware limits: shared memory (228 KB per SM, 1 KB reserved per CTA), registers (granted per
  // 4-warp granule and per SM sub-partition) and TMEM (512 columns).  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for
  // the two-CTA variant although the hardware limits are 2 / 2 (ncu launch__occupancy_limit_{registers,shared_mem} of this very
  // kernel); cudaLaunchCooperativeKernel applies the same estimate, so when it refuses the grid the kernel is launched plainly --
  // the grid never exceeds what is resident at once, the stream order guarantees the SMs are free, and the group barrier traps
  // after 4 s instead of hanging if that assumption were ever violated.
  cudaFuncAttributes fa;
  AS_CUDA(cudaFuncGetAttributes(&fa, mean_shift_v2_kernel<KP>));
  const int warps = (V2_THREADS + 31) / 32;
  const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
  const int warps_per_subpart = regs_per_warp ? 16384 / regs_per_warp : 64;
  int occ = (4 * warps_per_subpart) / warps;
  occ = std::min(occ, (int)((228 * 1024) / (smem + fa.sharedSizeBytes + 1024)));
  occ = std::min(occ, 512 / Cfg<KP>::TM_COLS);
  int groups = occ * num_sms / G;
  if (groups < 1) return AS_ERR_BAD_ARG;
  if (groups > p.n_img) groups = p.n_img;
  void* args[] = {(void*)&tm[0], (void*)&tm[1], (void*)&tm[2], (void*)&tm[3], (void*)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)mean_shift_v2_kernel<KP>, dim3(groups * G), dim3(V2_THREADS), args, smem, stream);
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    (void)cudaGetLastError();
    g_v2_plain_launches++;
    mean_shift_v2_kernel<KP><<<groups * G, V2_THREADS, smem, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return (int)e;
  return 0;
}

}  // namespace

static unsigned long long* g_v2_dbg = nullptr;
// profiling aid: device buffer of [grid][16] uint64 that receives the accumulated nanoseconds per phase (null = off)
extern "C" void as_mean_shift_v2_debug(unsigned long long* buf) { g_v2_dbg = buf; }

// diagnostics: resident CTAs per SM the runtime grants the KP = 64 variant (2 expected), registers, static shared memory
extern "C" int as_mean_shift_v2_occupancy(int* regs, int* static_smem, int* dyn_smem) {
  const size_t smem = Cfg<64>::SMEM;
  cudaFuncSetAttribute(mean_shift_v2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mean_shift_v2_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, mean_shift_v2_kernel<64>) != cudaSuccess) return -1;
  if (regs) *regs = a.numRegs;
  if (static_smem) *static_smem = (int)a.sharedSizeBytes;
  if (dyn_smem) *dyn_smem = (int)smem;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mean_shift_v2_kernel<64>, V2_THREADS, smem) != cudaSuccess) return -2;
  return occ + 100 * g_v2_plain_launches;     // + 100 x launches that had to bypass the cooperative launch's estimate
}

// 0 when the kernel cannot take the problem (the caller then uses as_mean_shift_tc / as_mean_shift)
extern "C" int as_mean_shift_v2_supported(int N, int C, int kmax, int max_obj) {
  return (C % 128 == 0 && C <= 1024 && N >= 1 && kp_for(kmax, max_obj) != 0) ? 1 : 0;
}

extern "C" size_t as_mean_shift_v2_workspace(int n_img, int N, int C, int kmax, int max_obj) {
  const int KP = kp_for(kmax, max_obj);
  if (!KP) return 0;
  const int U = (N + 63) / 64, G = (U + 1) / 2;
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 255) & ~(size_t)255; };
  add((size_t)n_img * N * C * 2); add((size_t)n_img * N * C * 2);      // hi, lo
  add((size_t)n_img * N * 4);                                          // den
  add((size_t)n_img * KP * C * 2); add((size_t)n_img * KP * C * 2);    // phat hi, lo
  add((size_t)n_img * G * KP * 4); add((size_t)n_img * G * KP * 8); add((size_t)n_img * G * KP * 4);
  add((size_t)n_img * G * KP * C * 4);                                 // proto partials
  add((size_t)n_img * 4);                                              // barrier counters
  add((size_t)(2 + 1024) * 4);                                         // scheduling tickets + per-SM arrival counters
  return b;
}

// Same contract as as_mean_shift_tc.  Requires as_mean_shift_v2_supported(...) and enough co-resident CTAs for one image
// (ceil(ceil(N/64)/2) <= 2 x #SMs for <= 64 seed columns, <= #SMs otherwise); AS_ERR_BAD_ARG otherwise.
extern "C" int as_mean_shift_v2(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                                const int* img_first, const int* img_nobj, int kmax, int max_obj, const float* rois,
                                int n_tot, int S, float* proto, float* sim_out, int n_shift, double tau0, double temp,
                                int clamp0, int* trace, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  const int KP = kp_for(kmax, max_obj);
  if (!as_mean_shift_v2_supported(N, C, kmax, max_obj) || hp * wp != N || kmax < 1 || S < 1 || S > 127) return AS_ERR_BAD_ARG;
  if (workspace_bytes < as_mean_shift_v2_workspace(n_img, N, C, kmax, max_obj)) return AS_ERR_BAD_ARG;
  const int U = (N + 63) / 64, G = (U + 1) / 2;
  int dev, num_sms;
  AS_CUDA(cudaGetDevice(&dev));
  AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t x) { char* r = base + off; off += (x + 255) & ~(size_t)255; return r; };
  __half* hi = (__half*)take((size_t)n_img * N * C * 2);
  __half* lo = (__half*)take((size_t)n_img * N * C * 2);
  V2Params p{};
  p.den = (float*)take((size_t)n_img * N * 4);
  p.phat_hi = (__half*)take((size_t)n_img * KP * C * 2);
  p.phat_lo = (__half*)take((size_t)n_img * KP * C * 2);
  p.colmax_part = (float*)take((size_t)n_img * G * KP * 4);
  p.dens_part = (float*)take((size_t)n_img * G * KP * 8);
  p.z_part = (float*)take((size_t)n_img * G * KP * 4);
  p.proto_part = (float*)take((size_t)n_img * G * KP * C * 4);
  p.bar = (unsigned*)take((size_t)n_img * 4);
  p.sched = (unsigned*)take((size_t)(2 + 1024) * 4);
  p.n_img = n_img; p.N = N; p.C = C; p.hp = hp; p.wp = wp; p.S = S; p.G = G; p.n_shift = n_shift; p.clamp0 = clamp0;
  p.tt0 = (float)(temp * tau0); p.temp = (float)temp;
  p.dbg = g_v2_dbg;
  {
    static int stagger = -1;                 // env AS_MS_STAGGER_NS: start offset of the second half of the image groups
    if (stagger < 0) {
      const char* e = getenv("AS_MS_STAGGER_NS");
      stagger = e ? atoi(e) : 8000;
    }
    p.stagger_ns = (unsigned)stagger;
  }
  p.img_first = img_first; p.img_nobj = img_nobj; p.rois = rois; p.proto = proto; p.sim_out = sim_out; p.trace = trace; p.n_tot = n_tot;

  AS_CUDA(cudaMemsetAsync(p.bar, 0, (size_t)n_img * 4, stream));
  AS_CUDA(cudaMemsetAsync(p.sched, 0, (size_t)(2 + 1024) * 4, stream));
  AS_CUDA(cudaMemsetAsync(p.phat_hi, 0, (size_t)n_img * KP * C * 2, stream));      // rows past an image's seed count stay zero
  AS_CUDA(cudaMemsetAsync(p.phat_lo, 0, (size_t)n_img * KP * C * 2, stream));
  v2_split_tokens<<<dim3((N + 7) / 8, n_img), 256, 0, stream>>>(feats, feat_img_stride, N, C, hi, lo, (float*)p.den);

  CUtensorMap tm[4];
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)n_img};
  uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)N * C * 2};
  uint32_t box64[3] = {64, 64, 1};
  int r = as_encode_tmap(&tm[0], hi, 2, 3, dims, str, box64);
  if (!r) r = as_encode_tmap(&tm[1], lo, 2, 3, dims, str, box64);
  uint64_t pdims[3] = {(uint64_t)C, (uint64_t)KP, (uint64_t)n_img};
  uint64_t pstr[2] = {(uint64_t)C * 2, (uint64_t)KP * C * 2};
  uint32_t pbox[3] = {64, (uint32_t)KP, 1};
  if (!r) r = as_encode_tmap(&tm[2], p.phat_hi, 2, 3, pdims, pstr, pbox);
  if (!r) r = as_encode_tmap(&tm[3], p.phat_lo, 2, 3, pdims, pstr, pbox);
  if (r) return r;
  if (KP == 64) r = launch_v2<64>(tm, p, num_sms, G, stream);
  else if (KP == 128) r = launch_v2<128>(tm, p, num_sms, G, stream);
  else r = launch_v2<256>(tm, p, num_sms, G, stream);
  if (r) return r;
  AS_LAUNCH_CHECK();
  return 0;
}
