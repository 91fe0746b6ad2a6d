// Attention-shift loop as ONE persistent cooperative kernel (as_mean_shift_fused).
// Reference: cosine_shift_batch RH:830-854 + update_density_batch RH:882-908 (RH = stdroi_point_deform_attn_reppoints.py).
//
// Layout of the work: an image's N tokens are split over G = ceil(N/256) CTAs (one per SM, 256 tokens = two 128-row MMA
// tiles each); the CTAs of an image form a group that synchronises through a global counter (4 group barriers per
// iteration, no kernel launches, no host).  Token features are read as the split-fp16 pair written once by
// tc_split_tokens (f^ * 2^10 = hi + lo; 4 bytes / element like fp32, and 100 MB at bs8 1024^2 stays L2 resident).
//
// Per iteration and CTA:
//   A  affinity    TMA streams the CTA's token tiles (hi, lo) through a 3-stage ring and the image's seed tiles p^ (split
//                  fp16, written by phase C of the previous iteration) through a 2-stage one; tcgen05 accumulates
//                  hi.hi + hi.lo + lo.hi (M128 N64 K16) for both 128-token tiles in TMEM; epilogue -> box-masked
//                  similarities in shared memory (kept for phase B), column max + density partials of the previous
//                  assignment -> global
//   -- group barrier --
//   A2 statistics  every CTA combines the partials (fixed order): tau, logit max; partial softmax denominators -> global
//   -- group barrier --
//   B  assign      Z; per token and instance the arg-max seed of the softmax weight (first wins) and its weight
//      update      on the tensor cores as well: D[channel, seed] += F^T[channel, token] . W[seed, token] per 64-token unit
//                  and 128-channel block.  F^T is the SAME shared-memory token tile the affinity pass uses, addressed as an
//                  MN-major operand (tokens = K); W is the sparse weight tile (one non-zero per token and instance: weight
//                  x |f|, scaled by a per-seed power of two and split into fp16 hi + lo) the workers scatter into shared
//                  memory after the assignment.  hi.hi + hi.lo + lo.hi as in the affinity; the 6 x [128 x 64] fp32
//                  accumulators stay in TMEM (384 columns) while the token tiles stream through a 4-stage ring; partials
//                  -> global.  Cost independent of the box sizes (the former gather over listed tokens was not).
//   -- group barrier --
//   C  reduce      ordered sum of the G partials -> new prototypes, normalised split-fp16 copy p^ for the next affinity
//                  (generic stores, fence.proxy.async + the group barrier make them visible to the other CTAs' TMA)
//   -- group barrier --
// and one more affinity pass (unmasked) for the returned similarity maps.  All reductions are ordered: deterministic.
#include <cstdlib>
#include "common.cuh"
#include <float.h>

using namespace asb;

namespace {

constexpr int TOK = 256;                 // tokens per CTA
constexpr int LDK = 64;                  // seed columns per image (n_obj * S <= 64)
constexpr int MAXOBJ = 8;
constexpr int SIM_LD = LDK + 1;
constexpr int A_STAGE = 32768;           // hi 16 KB + lo 16 KB of a [128 x 64] fp16 tile
constexpr int B_TILE = 16384;            // hi 8 KB + lo 8 KB of a [64 x 64] seed tile
constexpr int RING = 3 * A_STAGE + 2 * B_TILE;     // 128 KB: 3 A stages + 2 B tiles (phase A) = 4 update stages of 32 KB (phase B)
constexpr int U_STAGE = 32768;           // [64 tokens x 128 channels] hi + lo: four 8 KB boxes
constexpr int U_STAGES = RING / U_STAGE;
constexpr int W_TILE = 16384;            // [64 seeds x 64 tokens] weights, hi 8 KB + lo 8 KB
constexpr uint32_t TM_UPD = 128;         // first TMEM column of the update accumulators (affinity: 0..127)
constexpr int FUSED_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2..9 workers
constexpr float OP_SCALE = 1024.f;

struct Box { int r0, r1, c0, c1; };
__device__ __forceinline__ Box patch_box(const float* roi, int hp, int wp) {   // box2mask(rois // 16), RH:303-309
  Box b;
  b.c0 = (int)floorf(roi[0] / 16.f); b.r0 = (int)floorf(roi[1] / 16.f);
  b.c1 = (int)(floorf(roi[2] / 16.f) + 1.f); b.r1 = (int)(floorf(roi[3] / 16.f) + 1.f);
  b.c0 = max(0, min(b.c0, wp)); b.c1 = max(0, min(b.c1, wp));
  b.r0 = max(0, min(b.r0, hp)); b.r1 = max(0, min(b.r1, hp));
  return b;
}
__device__ __forceinline__ bool in_box(const Box& b, int n, int wp) {
  const int r = n / wp, c = n - r * wp;
  return r >= b.r0 && r < b.r1 && c >= b.c0 && c < b.c1;
}

struct FusedParams {
  int n_img, N, C, hp, wp, S, G, n_shift, clamp0;
  int cluster;                 // the G CTAs of an image form one thread-block cluster: group barriers are barrier.cluster
  int blocked;                 // token copies stored as 8 x 8 patch blocks (hp, wp multiples of 8): a 64-token unit is one block
  float tt0, temp;
  const int* img_first; const int* img_nobj; const float* rois;
  const float* den;            // [n_img][N]   |f| (clamped at 1e-8)
  float* proto;                // [n_tot][S][C] in/out
  __half* phat_hi;             // [n_img][LDK][C] normalised seeds * 2^10, split fp16 (scratch; read back through TMA)
  __half* phat_lo;
  float* sim_out;              // [n_tot][S][N]
  int* trace;                  // [n_shift][n_tot][N] or null
  int n_tot;
  float* colmax_part;          // [n_img][G][LDK]
  float* dens_part;            // [n_img][G][LDK][2]
  float* z_part;               // [n_img][G][LDK]
  float* proto_part;           // [n_img][G][LDK][C]
  unsigned* bar;               // [n_img] monotonic group-barrier counters
  unsigned long long* dbg;     // optional [grid][16] accumulated ns per phase (as_mean_shift_fused_debug)
};

__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target) {
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
      if ((++spins & 0xff) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}
// One image = one group of G CTAs.  When the group is a thread-block cluster (G in {2, 4, 8, 16}) the hardware cluster barrier
// (release / acquire at cluster scope: the partials in global memory written before it are visible to the peers after it)
// replaces the global counter + polling loop.
__device__ __forceinline__ void group_sync(int cluster, unsigned* ctr, unsigned target) {
  if (cluster) { __syncwarp(); cluster_sync_all(); }
  else group_barrier(ctr, target);
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// byte offset of element (row, col) inside a K-major [rows x 64 halves] tile with the 128-byte swizzle
__device__ __forceinline__ int sw128(int row, int col) { return row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1); }

// ordered sum of G strided partials, loads issued in batches of 8 so that their latencies overlap
__device__ __forceinline__ float sum_partials(const float* base, size_t stride, int G) {
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

__global__ void __launch_bounds__(FUSED_THREADS, 1)
mean_shift_fused_kernel(const __grid_constant__ CUtensorMap tm_hi64, const __grid_constant__ CUtensorMap tm_lo64,
                        const __grid_constant__ CUtensorMap tm_phi, const __grid_constant__ CUtensorMap tm_plo,
                        const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                       // RING bytes
  float* sims_s = reinterpret_cast<float*>(smem + RING);      // [TOK][SIM_LD]
  uint8_t* wt_s = reinterpret_cast<uint8_t*>(sims_s);         // phase B: 4 weight tiles (one per 64-token unit) alias it
  uint8_t* misc = smem + RING + TOK * SIM_LD * 4;
  float* w_s = reinterpret_cast<float*>(misc);                // [MAXOBJ][TOK] weight of the assigned seed (0 outside the box)
  float* den_s = w_s + MAXOBJ * TOK;                          // [TOK]
  float* red_s = den_s + TOK;                                 // [256][3]
  float* st_s = red_s + 256 * 3;                              // [LDK][4] 1/tt, column max, 1/Z, tau
  float* sc_s = st_s + LDK * 4;                               // [LDK][2] weight scale 2^k of the seed, 2^-k / OP_SCALE
  int8_t* idx_s = reinterpret_cast<int8_t*>(sc_s + LDK * 2);  // [MAXOBJ][TOK] assigned seed of the previous iteration
  uint64_t* bars = reinterpret_cast<uint64_t*>(idx_s + MAXOBJ * TOK);
  bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~(uintptr_t)7);
  uint64_t* a_full = bars;           // 3
  uint64_t* a_empty = bars + 3;      // 3
  uint64_t* b_full = bars + 6;       // 2
  uint64_t* b_empty = bars + 8;      // 2
  uint64_t* acc_full = bars + 10;    // 1
  uint64_t* w_full = bars + 11;      // 1 (weight tiles built: 8 worker warps arrive)
  uint64_t* u_full = bars + 12;      // U_STAGES (phase B token stages)
  uint64_t* u_empty = bars + 16;     // U_STAGES
  uint64_t* upd_done = bars + 20;    // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  uint8_t* ok_s = reinterpret_cast<uint8_t*>(bars + 26);       // [TOK] is the CTA's local token a token of the image?
  int* unit_act_s = reinterpret_cast<int*>(bars + 22);         // [4] does the 64-token unit hold a token inside any instance box?

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wt = threadIdx.x - 64;                            // worker thread id 0..255 (negative for warps 0, 1)
  const int groups = gridDim.x / p.G;
  const int grp = blockIdx.x / p.G, q = blockIdx.x % p.G;     // group id, rank inside the group
  const int kblocks = p.C / 64;
  const int cblocks = p.C / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi64); tma_prefetch_desc(&tm_lo64);
    tma_prefetch_desc(&tm_phi); tma_prefetch_desc(&tm_plo);
    for (int i = 0; i < 3; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < U_STAGES; ++i) { mbar_init(&u_full[i], 1); mbar_init(&u_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(w_full, 8);
    mbar_init(upd_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // pipeline state that survives across passes / images (registers of the respective warps)
  int a_stage = 0; uint32_t a_phase = 0;        // producer + MMA: A ring
  uint32_t bcount = 0;                          // producer + MMA: B tile uses so far
  uint32_t acount = 0;                          // workers: affinity passes so far
  uint32_t ucount = 0;                          // producer + MMA: phase B stages so far
  uint32_t wcount = 0;                          // MMA + workers: update phases so far
  uint32_t dcount = 0;                          // workers: update phases that issued MMAs (completed phases of upd_done)

  uint64_t t_prev = global_timer_ns();
  auto mark = [&](int k) {
    if (p.dbg && wt == 0) { const uint64_t t = global_timer_ns(); p.dbg[blockIdx.x * 16 + k] += t - t_prev; t_prev = t; }
  };
  if (grp >= groups) {                          // surplus CTAs (grid not a multiple of G): nothing to do
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
    return;
  }

  for (int img = grp; img < p.n_img; img += groups) {
    const int nobj = p.img_nobj[img], o0 = p.img_first[img];
    const int kb_cols = nobj * p.S;
    // the CTA's 256 tokens are four 64-token units taken round-robin over the group (unit u of CTA q = global unit u*G + q):
    // every CTA sees all parts of the image, so box-shaped instance masks load the CTAs of a group evenly
    // With blocked storage a unit is an 8 x 8 block of patches, so that whole units fall outside every instance box and are
    // skipped (below); tok() is the token's index in the image either way.
    const int wb = p.wp >> 3;
    // unit u of CTA q = global unit u*G + (q + u*rot) % G: the rotation spreads a CTA's four blocks over the block columns,
    // so a box loads the CTAs of a group evenly
    const int rot = p.blocked ? p.G / 4 + 1 : 0;
    auto unit_of = [&](int u) { return u * p.G + (q + u * rot) % p.G; };
    auto tok = [&](int tl) {
      const int unit = unit_of(tl >> 6), t = tl & 63;
      if (!p.blocked) return unit * 64 + t;
      const int by = unit / wb, bx = unit - by * wb;
      return (by * 8 + (t >> 3)) * p.wp + bx * 8 + (t & 7);
    };
    unsigned* ctr = p.bar + img;
    unsigned bar_target = 0;
    float fmax_cta = 1.f;
    // per-thread token indices of the image (index arithmetic hoisted out of the passes): n_wt for local token wt, n_tl / box
    // membership bits for the token whose affinity row this thread reads back from TMEM
    int n_wt = 0, n_tl = 0;
    unsigned inmask_tl = 0, inmask_wt = 0;
    if (wt >= 0) {
      n_wt = tok(wt);
      n_tl = tok(((warp - 2) >> 2) * 128 + (warp & 3) * 32 + lane);
      if (n_tl < p.N)
        for (int j = 0; j < nobj; ++j)
          if (in_box(patch_box(p.rois + 4 * (o0 + j), p.hp, p.wp), n_tl, p.wp)) inmask_tl |= 1u << j;
      if (n_wt < p.N)
        for (int j = 0; j < nobj; ++j)
          if (in_box(patch_box(p.rois + 4 * (o0 + j), p.hp, p.wp), n_wt, p.wp)) inmask_wt |= 1u << j;
      ok_s[wt] = n_wt < p.N;
    }
    if (wt >= 0) {
      const int n = n_wt;
      const float dn = n < p.N ? p.den[(size_t)img * p.N + n] : 1.f;
      den_s[wt] = dn;
      for (int j = 0; j < MAXOBJ; ++j) { idx_s[j * TOK + wt] = -1; w_s[j * TOK + wt] = 0.f; }
      // A 64-token unit none of whose tokens lies in any instance box contributes exact zeros to the masked affinity and to the
      // update (the reference multiplies those tokens by zero, RH:1824): its token tiles are neither loaded nor multiplied in the
      // masked passes.  (The last, unmasked pass reads every unit.)
      if (wt < 4) unit_act_s[wt] = 0;
      workers_sync();
      if (__any_sync(0xffffffffu, inmask_wt != 0) && lane == 0) atomicOr(&unit_act_s[wt >> 6], 1);
      // largest token norm of this CTA (bounds weight x |f| for the power-of-two scaling of the update's weight tiles)
      const float wm = warp_max(dn);
      if (lane == 0) red_s[warp - 2] = wm;
      workers_sync();
      float fm = red_s[0];
      for (int k = 1; k < 8; ++k) fm = fmaxf(fm, red_s[k]);
      fmax_cta = fm;
      workers_sync();
    }
    // new prototypes (ordered sum of the group's partials) and their normalised split-fp16 copy, rows q, q+G, ...
    auto phase_c = [&](bool from_partials) {
      if (wt >= 0) {
        for (int r0 = q; r0 < kb_cols; r0 += 4 * p.G) {       // four rows at a time: their partial loads overlap
          float v[4][4], ss[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[i][e] = 0.f;
          if (from_partials) {
            // ordered sum over the group's partials; the loads of all (row, channel) pairs of a step are in flight together
            for (int g0 = 0; g0 < p.G; g0 += 4) {
              float t[4][4][4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                  for (int gg = 0; gg < 4; ++gg) {
                    const int r = r0 + i * p.G, c = wt + e * 256, g = g0 + gg;
                    t[i][e][gg] = (r < kb_cols && c < p.C && g < p.G)
                                      ? __ldcg(p.proto_part + (((size_t)img * p.G + g) * LDK + r) * p.C + c) : 0.f;
                  }
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                  for (int gg = 0; gg < 4; ++gg) v[i][e] += t[i][e][gg];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int r = r0 + i * p.G, c = wt + e * 256;
                if (r < kb_cols && c < p.C) v[i][e] = p.proto[((size_t)o0 * p.S + r) * p.C + c];
              }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ss[i] = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) ss[i] += v[i][e] * v[i][e];
          }
          if (from_partials) {                                // stores only after every partial load has been issued
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int r = r0 + i * p.G, c = wt + e * 256;
                if (r < kb_cols && c < p.C) p.proto[((size_t)o0 * p.S + r) * p.C + c] = v[i][e];
              }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ss[i] = warp_sum(ss[i]);
            if (lane == 0) red_s[i * 8 + warp - 2] = ss[i];
          }
          workers_sync();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * p.G;
            float tot = 0.f;
            for (int k = 0; k < 8; ++k) tot += red_s[i * 8 + k];
            const float nrm = fmaxf(sqrtf(tot), 1e-8f);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = wt + e * 256;
              if (r < kb_cols && c < p.C) {
                const float x = v[i][e] / nrm * OP_SCALE;
                const __half h = __float2half_rn(x);
                const size_t o = ((size_t)img * LDK + r) * p.C + c;
                p.phat_hi[o] = h;
                p.phat_lo[o] = __float2half_rn(x - __half2float(h));
              }
            }
          }
          workers_sync();
        }
        fence_proxy_async_all();                              // the next reader of phat_hi/lo is another CTA's TMA
      }
    };
    phase_c(false);
    bar_target += p.G;
    group_sync(p.cluster, ctr, bar_target);
    mark(0);
    const unsigned umask = (unit_act_s[0] ? 1u : 0u) | (unit_act_s[1] ? 2u : 0u) | (unit_act_s[2] ? 4u : 0u) | (unit_act_s[3] ? 8u : 0u);
    auto uact = [&](int u) { return ((umask >> u) & 1u) != 0; };
    const bool any_unit = umask != 0;

    for (int it = 0; it <= p.n_shift; ++it) {
      const bool last = (it == p.n_shift);                    // extra pass: unmasked similarities for the output
      // ================================================================ phase A: affinity
      if (warp == 0) {
        if (lane == 0) {
          fence_proxy_async_all();
          for (int kb = 0; kb < kblocks && (last || any_unit); ++kb) {
            const uint32_t bb = bcount & 1;
            mbar_wait(&b_empty[bb], ((bcount >> 1) & 1) ^ 1);
            mbar_expect_tx(&b_full[bb], B_TILE);
            tma_load_3d(ring + 3 * A_STAGE + bb * B_TILE, &tm_phi, &b_full[bb], kb * 64, 0, img);
            tma_load_3d(ring + 3 * A_STAGE + bb * B_TILE + 8192, &tm_plo, &b_full[bb], kb * 64, 0, img);
            ++bcount;
            for (int mt = 0; mt < 2; ++mt) {
              const bool a0 = last || uact(2 * mt), a1 = last || uact(2 * mt + 1);
              if (!a0 && !a1) continue;                       // the whole 128-row tile is outside every box
              mbar_wait(&a_empty[a_stage], a_phase ^ 1);
              mbar_expect_tx(&a_full[a_stage], (a0 ? 16384 : 0) + (a1 ? 16384 : 0));
              for (int h = 0; h < 2; ++h) {                   // a 128-row operand tile = two 64-token units
                if (!(h ? a1 : a0)) continue;                 // stale rows of a skipped unit are masked to zero by the epilogue
                const int row0 = unit_of(2 * mt + h) * 64;
                tma_load_3d(ring + a_stage * A_STAGE + h * 8192, &tm_hi64, &a_full[a_stage], kb * 64, row0, img);
                tma_load_3d(ring + a_stage * A_STAGE + 16384 + h * 8192, &tm_lo64, &a_full[a_stage], kb * 64, row0, img);
              }
              if (++a_stage == 3) { a_stage = 0; a_phase ^= 1; }
            }
          }
        }
      } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(0, 128, LDK);
        const bool t0 = last || uact(0) || uact(1), t1 = last || uact(2) || uact(3);     // which 128-row tiles take part
        const int mt_last = t1 ? 1 : 0;
        for (int kb = 0; kb < kblocks && (t0 || t1); ++kb) {
          const uint32_t bb = bcount & 1;
          mbar_wait(&b_full[bb], (bcount >> 1) & 1);
          for (int mt = 0; mt < 2; ++mt) {
            if (!(mt ? t1 : t0)) continue;
            mbar_wait(&a_full[a_stage], a_phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_hi = smem_u32(ring + a_stage * A_STAGE), a_lo = a_hi + 16384;
              const uint32_t b_hi = smem_u32(ring + 3 * A_STAGE + bb * B_TILE), b_lo = b_hi + 8192;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma_f16_ss(tmem + mt * LDK, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, (kb | k) != 0);
                mma_f16_ss(tmem + mt * LDK, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(b_lo + k * 32), idesc, 1);
                mma_f16_ss(tmem + mt * LDK, umma_desc_k_sw128(a_lo + k * 32), umma_desc_k_sw128(b_hi + k * 32), idesc, 1);
              }
              tc_commit(&a_empty[a_stage]);
              if (mt == mt_last) tc_commit(&b_empty[bb]);
              if (mt == mt_last && kb == kblocks - 1) tc_commit(acc_full);
            }
            __syncwarp();
            if (++a_stage == 3) { a_stage = 0; a_phase ^= 1; }
          }
          ++bcount;
        }
      } else {
        // ---- epilogue: TMEM -> box-masked similarities in shared memory (token-major)
        {
          const int quad = warp & 3, mt = (warp - 2) >> 2;
          const int tl = mt * 128 + quad * 32 + lane;
          // bit j: token inside instance j's box (all ones on the last pass)
          unsigned inmask = n_tl < p.N ? (last ? 0xffffffffu : inmask_tl) : 0u;
          const bool pass_ran = last || any_unit;             // (uniform over the CTA) were any MMAs issued in this pass?
          const bool tile_ran = last || uact(2 * mt) || uact(2 * mt + 1);
          if (pass_ran) mbar_wait(acc_full, acount & 1);
          tc_fence_after();
          mark(10);
          uint32_t v0[32], v1[32];
          if (tile_ran) {
            tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + mt * LDK, v0);
            tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + mt * LDK + 32, v1);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v0[i] = v1[i] = 0u;
            inmask = 0;
          }
          const float alpha = 1.f / (OP_SCALE * OP_SCALE);
          float* row = sims_s + tl * SIM_LD;
          int j = 0, s = 0;
#pragma unroll
          for (int col = 0; col < LDK; ++col) {
            const float raw = __uint_as_float(col < 32 ? v0[col & 31] : v1[col & 31]);
            row[col] = (col < kb_cols && ((inmask >> j) & 1u)) ? raw * alpha : 0.f;
            if (++s == p.S) { s = 0; ++j; }
          }
          tc_fence_before();
        }
        if (last || any_unit) ++acount;                       // acc_full completed a phase only if the pass issued MMAs
        workers_sync();
        if (last) {
          // returned maps [o][s][n]: one coalesced row of this CTA's tokens per seed
          const int n = n_wt;
          if (n < p.N)
            for (int col = 0; col < kb_cols; ++col) {
              float v = sims_s[wt * SIM_LD + col];
              if (p.clamp0) v = fmaxf(v, 0.f);
              p.sim_out[((size_t)o0 * p.S + col) * p.N + n] = v;
            }
        } else {
          // column statistics over this CTA's tokens: max, density partials of the previous assignment
          const int col = wt & 63, g4 = wt >> 6;
          float mx = -FLT_MAX, sv = 0.f, cv = 0.f;
          if (col < kb_cols) {
            const int j = col / p.S, s = col - j * p.S;
#pragma unroll 8
            for (int t = g4; t < TOK; t += 4) {
              const bool ok = ok_s[t] != 0;
              const float v = sims_s[t * SIM_LD + col];
              mx = ok ? fmaxf(mx, v) : mx;
              const bool hit = ok && it > 0 && idx_s[j * TOK + t] == s;
              sv += hit ? v : 0.f; cv += hit ? 1.f : 0.f;
            }
          }
          red_s[wt * 3] = mx; red_s[wt * 3 + 1] = sv; red_s[wt * 3 + 2] = cv;
          workers_sync();
          if (g4 == 0 && col < kb_cols) {
            for (int g = 1; g < 4; ++g) { mx = fmaxf(mx, red_s[(g * 64 + col) * 3]); sv += red_s[(g * 64 + col) * 3 + 1]; cv += red_s[(g * 64 + col) * 3 + 2]; }
            const size_t pi = ((size_t)img * p.G + q) * LDK + col;
            p.colmax_part[pi] = mx; p.dens_part[pi * 2] = sv; p.dens_part[pi * 2 + 1] = cv;
          }
        }
      }
      mark(1);
      if (last) break;
      bar_target += p.G;
      group_sync(p.cluster, ctr, bar_target);
      mark(2);
      // ================================================================ phase A2: per-seed statistics, partial Z
      if (wt >= 0) {
        const int col = wt & 63, g4 = wt >> 6;
        if (g4 == 0 && col < kb_cols) {
          const size_t p0 = (size_t)img * p.G * LDK + col;
          float mx = -FLT_MAX;
          for (int g0 = 0; g0 < p.G; g0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (g0 + i < p.G) ? __ldcg(p.colmax_part + p0 + (size_t)(g0 + i) * LDK) : -FLT_MAX;
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, v[i]);
          }
          float tt, tau = 0.f;
          if (it == 0) tt = p.tt0;
          else {
            const float tot = sum_partials(p.dens_part + p0 * 2, (size_t)LDK * 2, p.G);
            const float cnt = sum_partials(p.dens_part + p0 * 2 + 1, (size_t)LDK * 2, p.G);
            tau = fmaxf(1.f - (cnt >= 1.f ? tot / cnt : 0.f), 1e-10f);       // RH:883-885, 908
            tt = p.temp * tau;
          }
          st_s[col * 4] = 1.f / tt; st_s[col * 4 + 1] = mx; st_s[col * 4 + 3] = tau;
        }
        workers_sync();
        float z = 0.f;
        if (col < kb_cols) {
          // logit - max logit as (v - max) / tt: exactly 0 at the maximum even when tt is ~1e-11 (tau clamped at 1e-10)
          const float inv_tt = st_s[col * 4], cmax = st_s[col * 4 + 1];
#pragma unroll 8
          for (int t = g4; t < TOK; t += 4) {
            const float e = expf((sims_s[t * SIM_LD + col] - cmax) * inv_tt);
            sims_s[t * SIM_LD + col] = e;                       // the assignment only needs weight = e / Z: no second exp
            z += ok_s[t] ? e : 0.f;
          }
        }
        red_s[wt] = z;
        workers_sync();
        if (g4 == 0 && col < kb_cols) {
          for (int g = 1; g < 4; ++g) z += red_s[g * 64 + col];
          p.z_part[((size_t)img * p.G + q) * LDK + col] = z;
        }
      }
      mark(3);
      bar_target += p.G;
      group_sync(p.cluster, ctr, bar_target);
      mark(4);
      // ================================================================ phase B: assign + update
      if (warp == 0) {
        if (lane == 0) {                                      // the ring is free: the update's token tiles start streaming now
          // channel blocks in DESCENDING order: the affinity pass before this one finished with the last channels and the
          // one after it starts with the first, so consecutive passes meet in whatever part of the 100 MB token set the
          // L2 still holds
          for (int cb = cblocks - 1; cb >= 0; --cb)
            for (int u = 0; u < 4; ++u) {
              if (!uact(u)) continue;                         // every weight of the unit is zero
              const uint32_t st = ucount % U_STAGES;
              mbar_wait(&u_empty[st], ((ucount / U_STAGES) & 1) ^ 1);
              mbar_expect_tx(&u_full[st], U_STAGE);
              uint8_t* dst = ring + st * U_STAGE;
              const int row0 = unit_of(u) * 64;
              tma_load_3d(dst, &tm_hi64, &u_full[st], (2 * cb) * 64, row0, img);
              tma_load_3d(dst + 8192, &tm_hi64, &u_full[st], (2 * cb + 1) * 64, row0, img);
              tma_load_3d(dst + 16384, &tm_lo64, &u_full[st], (2 * cb) * 64, row0, img);
              tma_load_3d(dst + 24576, &tm_lo64, &u_full[st], (2 * cb + 1) * 64, row0, img);
              ++ucount;
            }
        }
      } else if (warp == 1) {
        // D[channel (M = 128: two 64-channel tiles, LBO apart), seed (N = 64)] += F^T . W over the 64 tokens of the unit
        constexpr uint32_t idesc_u = umma_idesc(0, 128, LDK) | (1u << 15);      // A is MN-major (channels contiguous)
        mbar_wait(w_full, wcount & 1);
        tc_fence_after();
        const int u_first = uact(0) ? 0 : uact(1) ? 1 : uact(2) ? 2 : 3;
        const int u_last = uact(3) ? 3 : uact(2) ? 2 : uact(1) ? 1 : 0;
        for (int cb = cblocks - 1; cb >= 0; --cb)
          for (int u = 0; u < 4; ++u) {
            if (!uact(u)) continue;
            const uint32_t st = ucount % U_STAGES;
            mbar_wait(&u_full[st], (ucount / U_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_hi = smem_u32(ring + st * U_STAGE), a_lo = a_hi + 16384;
              const uint32_t w_hi = smem_u32(wt_s + u * W_TILE), w_lo = w_hi + 8192;
#pragma unroll
              for (int k = 0; k < 4; ++k) {                   // 16 tokens per MMA = two 8-token swizzle atoms (SBO = 1024 B)
                const uint64_t dah = umma_desc_mn_sw128(a_hi + k * 2048, 8192), dal = umma_desc_mn_sw128(a_lo + k * 2048, 8192);
                const uint64_t dwh = umma_desc_k_sw128(w_hi + k * 32), dwl = umma_desc_k_sw128(w_lo + k * 32);
                mma_f16_ss(tmem + TM_UPD + cb * LDK, dah, dwh, idesc_u, u != u_first || k != 0);
                mma_f16_ss(tmem + TM_UPD + cb * LDK, dah, dwl, idesc_u, 1);
                mma_f16_ss(tmem + TM_UPD + cb * LDK, dal, dwh, idesc_u, 1);
              }
              tc_commit(&u_empty[st]);
              if (u == u_last && cb == 0) tc_commit(upd_done);
            }
            __syncwarp();
            ++ucount;
          }
        ++wcount;
      } else if (wt >= 0) {
        if (wt < kb_cols) st_s[wt * 4 + 2] = 1.f / sum_partials(p.z_part + (size_t)img * p.G * LDK + wt, LDK, p.G);    // fixed order
        workers_sync();
        {
          const int n = n_wt;
          const float* row = sims_s + wt * SIM_LD;
          for (int j = 0; j < nobj; ++j) {
            float best = -1.f;
            int bi = 0;
            for (int s = 0; s < p.S; s += 4) {                // four independent exp chains in flight
              float w4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int col = j * p.S + min(s + i, p.S - 1);
                w4[i] = row[col] * st_s[col * 4 + 2];            // e (written by the statistics phase) / Z
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (s + i < p.S && w4[i] > best) { best = w4[i]; bi = s + i; }     // first maximum wins (torch.argmax)
            }
            const bool valid = n < p.N;
            const bool inside = (inmask_wt >> j) & 1u;
            idx_s[j * TOK + wt] = valid ? bi : -1;            // density of the next iteration counts every token
            w_s[j * TOK + wt] = inside ? best : 0.f;          // masked tokens are zero vectors: they add nothing
            if (p.trace && valid) p.trace[((size_t)it * p.n_tot + o0 + j) * p.N + n] = bi;
          }
        }
        // per-seed power-of-two scale: weight x |f| <= (1/Z) x fmax < 2^10 after scaling, so the fp16 hi + lo split keeps
        // ~22 bits of the weights that matter
        if (wt < LDK) {
          float sc = 1.f;
          if (wt < kb_cols) sc = exp2f((float)(9 - ilogbf(st_s[wt * 4 + 2] * fmax_cta)));
          sc_s[wt * 2] = sc;
          sc_s[wt * 2 + 1] = 1.f / (sc * OP_SCALE);           // f = (hi + lo) * |f| / 2^10: |f| is folded into the weight
        }
        workers_sync();                                       // sims_s is dead from here: it becomes the four weight tiles
        for (int i = wt; i < 4 * W_TILE / 16; i += 256) reinterpret_cast<uint4*>(wt_s)[i] = make_uint4(0, 0, 0, 0);
        workers_sync();
        {
          const int u = wt >> 6, col = wt & 63;
          for (int j = 0; j < nobj; ++j) {
            const float wv = w_s[j * TOK + wt];
            if (wv != 0.f) {
              const int r = j * p.S + idx_s[j * TOK + wt];
              const float v = wv * den_s[wt] * sc_s[r * 2];
              const __half h = __float2half_rn(v);
              uint8_t* dst = wt_s + u * W_TILE + sw128(r, col);
              *reinterpret_cast<__half*>(dst) = h;
              *reinterpret_cast<__half*>(dst + 8192) = __float2half_rn(v - __half2float(h));
            }
          }
        }
        fence_proxy_async();                                  // generic-proxy stores -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(w_full);
        mark(5);
        // ---- update epilogue: TMEM -> partial prototypes [seed][channel] (lane = channel: coalesced rows)
        ++wcount;
        if (any_unit) { mbar_wait(upd_done, dcount & 1); ++dcount; }
        tc_fence_after();
        {
          const int quad = warp & 3, half = (warp - 2) >> 2;
          for (int cb = half; cb < cblocks; cb += 2) {
            uint32_t v0[32], v1[32];
            if (any_unit) {
              tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + TM_UPD + cb * LDK, v0);
              tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + TM_UPD + cb * LDK + 32, v1);
              tc_wait_ld();
            } else {                                          // no token of this CTA lies in a box: its partial is zero
#pragma unroll
              for (int i = 0; i < 32; ++i) v0[i] = v1[i] = 0u;
            }
            float* dst = p.proto_part + ((size_t)img * p.G + q) * LDK * p.C + cb * 128 + quad * 32 + lane;
#pragma unroll
            for (int col = 0; col < LDK; ++col) {                 // running pointer: one 64-bit add per row, no wide multiply
              if (col < kb_cols) *dst = __uint_as_float(col < 32 ? v0[col & 31] : v1[col & 31]) * sc_s[col * 2 + 1];
              dst += p.C;
            }
          }
          tc_fence_before();
        }
      }
      mark(6);
      bar_target += p.G;
      group_sync(p.cluster, ctr, bar_target);
      mark(7);
      // ================================================================ phase C: new prototypes
      phase_c(true);
      mark(8);
      bar_target += p.G;
      group_sync(p.cluster, ctr, bar_target);
      mark(9);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace

// den / split tokens are produced by the same kernels as the tc variant
extern "C" size_t as_mean_shift_fused_workspace(int n_img, int N, int C) {
  const int G = (N + TOK - 1) / TOK;
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 255) & ~(size_t)255; };
  add((size_t)n_img * N * C * 2); add((size_t)n_img * N * C * 2);      // hi, lo
  add((size_t)n_img * N * 4);                                          // den
  add((size_t)n_img * LDK * C * 2); add((size_t)n_img * LDK * C * 2);  // phat hi, lo
  add((size_t)n_img * G * LDK * 4); add((size_t)n_img * G * LDK * 8); add((size_t)n_img * G * LDK * 4);
  add((size_t)n_img * G * LDK * C * 4);                                // proto partials
  add((size_t)n_img * 4);                                              // barrier counters
  return b;
}

static unsigned long long* g_fused_dbg = nullptr;
static int g_fused_clusters = -1;            // co-resident clusters reported for the last cluster launch (diagnostics)
#ifndef AS_MS_CLUSTER_DEFAULT
#define AS_MS_CLUSTER_DEFAULT 0
#endif
// profiling aid: device buffer of [grid][16] uint64 that receives the accumulated nanoseconds per phase (null = off)
extern "C" void as_mean_shift_fused_debug(unsigned long long* buf) { g_fused_dbg = buf; }
// diagnostics: co-resident clusters the driver reported for the last cluster launch (-1: no cluster launch so far)
extern "C" int as_mean_shift_fused_occupancy(void) { return g_fused_clusters; }

namespace {
__global__ void fused_split_tokens(const float* __restrict__ feats, long long fstride, int N, int C, int wp, int blocked,
                                   __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ den) {
  const int img = blockIdx.y;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* f = feats + img * fstride + (long long)n * C;
  const int lane = lane_id();
  float ss = 0.f;
  for (int c = lane * 4; c < C; c += 128) {                    // C % 128 == 0: float4 loads, 8-byte stores
    const float4 t = *reinterpret_cast<const float4*>(f + c);
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
  if (lane == 0) den[(size_t)img * N + n] = nrm;
  int pos = n;                                                 // row of the token in the split copies
  if (blocked) {
    const int y = n / wp, x = n - y * wp;
    pos = ((y >> 3) * (wp >> 3) + (x >> 3)) * 64 + (y & 7) * 8 + (x & 7);
  }
  const size_t o = ((size_t)img * N + pos) * C;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 t = *reinterpret_cast<const float4*>(f + c);
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
}  // namespace

// Same contract as as_mean_shift_tc; requires C % 128 == 0, C <= 768 (the update accumulators of all channels live in TMEM),
// kmax <= 64 and ceil(N/256) <= #SMs.
extern "C" int as_mean_shift_fused(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                                   const int* obj_img, const int* img_first, const int* img_nobj, int kmax, const float* rois,
                                   int n_tot, int S, float* proto, float* sim_out, int n_shift, double tau0, double temp,
                                   int clamp0, int* trace, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  (void)obj_img;
  if (n_tot <= 0) return 0;
  const int G = (N + TOK - 1) / TOK;
  int dev, num_sms;
  AS_CUDA(cudaGetDevice(&dev));
  AS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (C % 128 || C > 768 || hp * wp != N || kmax > LDK || kmax < 1 || G > num_sms || (kmax + S - 1) / S > MAXOBJ) return AS_ERR_BAD_ARG;
  if (workspace_bytes < as_mean_shift_fused_workspace(n_img, N, C)) return AS_ERR_BAD_ARG;
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t x) { char* r = base + off; off += (x + 255) & ~(size_t)255; return r; };
  __half* hi = (__half*)take((size_t)n_img * N * C * 2);
  __half* lo = (__half*)take((size_t)n_img * N * C * 2);
  FusedParams p{};
  p.den = (float*)take((size_t)n_img * N * 4);
  p.phat_hi = (__half*)take((size_t)n_img * LDK * C * 2);
  p.phat_lo = (__half*)take((size_t)n_img * LDK * C * 2);
  p.colmax_part = (float*)take((size_t)n_img * G * LDK * 4);
  p.dens_part = (float*)take((size_t)n_img * G * LDK * 8);
  p.z_part = (float*)take((size_t)n_img * G * LDK * 4);
  p.proto_part = (float*)take((size_t)n_img * G * LDK * C * 4);
  p.bar = (unsigned*)take((size_t)n_img * 4);
  p.n_img = n_img; p.N = N; p.C = C; p.hp = hp; p.wp = wp; p.S = S; p.G = G; p.n_shift = n_shift; p.clamp0 = clamp0;
  p.tt0 = (float)(temp * tau0); p.temp = (float)temp;
  p.dbg = g_fused_dbg;
  p.img_first = img_first; p.img_nobj = img_nobj; p.rois = rois; p.proto = proto; p.sim_out = sim_out; p.trace = trace; p.n_tot = n_tot;

  AS_CUDA(cudaMemsetAsync(p.bar, 0, (size_t)n_img * 4, stream));
  AS_CUDA(cudaMemsetAsync(p.phat_hi, 0, (size_t)n_img * LDK * C * 2, stream));      // rows past an image's seed count stay zero
  AS_CUDA(cudaMemsetAsync(p.phat_lo, 0, (size_t)n_img * LDK * C * 2, stream));
  p.blocked = (hp % 8 == 0 && wp % 8 == 0) ? 1 : 0;
  if (const char* e = getenv("AS_MS_BLOCKED")) p.blocked = p.blocked && atoi(e) != 0;      // measurement switch (row units)
  fused_split_tokens<<<dim3((N + 7) / 8, n_img), 256, 0, stream>>>(feats, feat_img_stride, N, C, wp, p.blocked, hi, lo, (float*)p.den);

  CUtensorMap tm[4];
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)n_img};
  uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)N * C * 2};
  uint32_t box64[3] = {64, 64, 1};
  int r = as_encode_tmap(&tm[0], hi, 2, 3, dims, str, box64);
  if (!r) r = as_encode_tmap(&tm[1], lo, 2, 3, dims, str, box64);
  uint64_t pdims[3] = {(uint64_t)C, (uint64_t)LDK, (uint64_t)n_img};
  uint64_t pstr[2] = {(uint64_t)C * 2, (uint64_t)LDK * C * 2};
  if (!r) r = as_encode_tmap(&tm[2], p.phat_hi, 2, 3, pdims, pstr, box64);
  if (!r) r = as_encode_tmap(&tm[3], p.phat_lo, 2, 3, pdims, pstr, box64);
  if (r) return r;

  const size_t smem = 1024 + RING + (size_t)TOK * SIM_LD * 4 + /* w */ MAXOBJ * TOK * 4 + /* idx */ MAXOBJ * TOK + /* den */ TOK * 4 +
                      /* red */ 256 * 3 * 4 + /* st, sc */ LDK * 6 * 4 + /* barriers */ 256 + 512;
  AS_CUDA(cudaFuncSetAttribute(mean_shift_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int groups = num_sms / G;
  if (groups > n_img) groups = n_img;
  const int grid = groups * G;
  // cluster launch when the group size allows it (env AS_MS_CLUSTER=0: global-counter barriers under a cooperative launch)
  static int use_cluster = -1;
  if (use_cluster < 0) { const char* e = getenv("AS_MS_CLUSTER"); use_cluster = e ? atoi(e) : AS_MS_CLUSTER_DEFAULT; }
  p.cluster = use_cluster && (G == 2 || G == 4 || G == 8 || G == 16);
  void* args[] = {&tm[0], &tm[1], &tm[2], &tm[3], &p};
  if (p.cluster) {
    if (G > 8) AS_CUDA(cudaFuncSetAttribute(mean_shift_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FUSED_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)mean_shift_fused_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
    g_fused_clusters = max_clusters;
    if (max_clusters >= 1) {
      // fewer co-resident clusters than images: the images are dealt over the clusters that exist (no inter-cluster dependency)
      if (max_clusters < groups) { groups = max_clusters; cfg.gridDim = dim3(groups * G); }
      AS_CUDA(cudaLaunchKernelExC(&cfg, (const void*)mean_shift_fused_kernel, args));
      AS_LAUNCH_CHECK();
      return 0;
    }
    p.cluster = 0;                               // this GPU cannot place a cluster of G such CTAs: counter barriers
  }
  AS_CUDA(cudaLaunchCooperativeKernel((const void*)mean_shift_fused_kernel, dim3(grid), dim3(FUSED_THREADS), args, smem, stream));
  AS_LAUNCH_CHECK();
  return 0;
}
