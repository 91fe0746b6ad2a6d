// Attention-shift loop on device (no host sync between iterations).
// Replaces: mean_shift_grid_prototype (RH:1778-1840: seed grid + feature gather),
//           cosine_shift_batch (RH:830-854) and update_density_batch (RH:882-908)
// RH = mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py of the reference.
//
// Data layout: token-major features [img][N][C] fp32 (= last_feat[:,1:] as the backbone stores it, no transpose),
// instances are a flat list (obj -> image index), the per-instance box mask is evaluated on the fly
// (the reference materialises n_obj masked copies of the feature map, RH:1824).
//
// Per iteration (all launches on one stream, deterministic reductions, no atomics on floats):
//   ms_sim      : sim[o,s,n] = <p^[o,s], f[n]> / max(|f[n]|,eps) (0 outside the box), row max, and the density
//                 partial sums of the PREVIOUS iteration's assignment (same prototypes, same tokens)
//   ms_stats    : tau[o,s] (density), logits' max and softmax denominator Z[o,s]
//   ms_assign   : per token the arg-max seed of softmax weight (first wins) and its weight
//   ms_update   : partial new prototypes sum_n w * f[n] per token tile
//   ms_finish   : ordered reduction of the partials, new p and p^ = p / max(|p|,eps)
// then one unmasked ms_sim for the returned similarity maps.
#include "common.cuh"
#include <float.h>

using namespace asb;

namespace {

constexpr int SIM_TOK = 128;   // tokens per CTA in ms_sim (one thread per token)
constexpr int SIM_SG = 16;     // seeds per CTA in ms_sim
constexpr int SIM_CK = 32;     // channel chunk staged in smem
constexpr int UPD_TOK = 256;   // tokens per CTA in ms_update

struct Box { int r0, r1, c0, c1; };

// box2mask(rois // 16) of RH:303-309: rows [int(y1//16), int(y2//16 + 1)), cols likewise, python-slice clipped
__device__ __forceinline__ Box patch_box(const float* roi, int hp, int wp) {
  Box b;
  b.c0 = (int)floorf(roi[0] / 16.f);
  b.r0 = (int)floorf(roi[1] / 16.f);
  b.c1 = (int)(floorf(roi[2] / 16.f) + 1.f);
  b.r1 = (int)(floorf(roi[3] / 16.f) + 1.f);
  b.c0 = max(0, min(b.c0, wp)); b.c1 = max(0, min(b.c1, wp));
  b.r0 = max(0, min(b.r0, hp)); b.r1 = max(0, min(b.r1, hp));
  return b;
}
__device__ __forceinline__ bool in_box(const Box& b, int n, int wp) {
  const int r = n / wp, c = n - r * wp;
  return r >= b.r0 && r < b.r1 && c >= b.c0 && c < b.c1;
}

// order-preserving float <-> uint encoding for atomicMax
__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];   // fixed order
  return t;
}

// den[img][n] = max(||f||_2, 1e-8)   (F.cosine_similarity clamps each norm separately, torch >= 1.12)
__global__ void ms_token_norm(const float* __restrict__ feats, long long fstride, int N, int C, float* __restrict__ den) {
  const int img = blockIdx.y;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float4* f = reinterpret_cast<const float4*>(feats + img * fstride + (long long)n * C);
  float ss = 0.f;
  for (int c = lane_id(); c < C / 4; c += 32) {
    float4 v = __ldg(f + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if (lane_id() == 0) den[(long long)img * N + n] = fmaxf(sqrtf(ss), 1e-8f);
}

// phat = p / max(||p||, 1e-8), one CTA per (obj, seed) row
__global__ void ms_proto_norm(const float* __restrict__ p, float* __restrict__ phat, int C) {
  __shared__ float red[32];
  const float* src = p + (long long)blockIdx.x * C;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) ss += src[c] * src[c];
  const float nrm = fmaxf(sqrtf(block_sum(ss, red)), 1e-8f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) phat[(long long)blockIdx.x * C + c] = src[c] / nrm;
}

// grid (ceil(N/128), ceil(S/16), n_tot)
__global__ void __launch_bounds__(SIM_TOK)
ms_sim(const float* __restrict__ feats, long long fstride, const float* __restrict__ den, const int* __restrict__ obj_img,
       const float* __restrict__ rois, const float* __restrict__ phat, int N, int C, int hp, int wp, int S, int masked,
       int clamp0, float* __restrict__ sim, unsigned* __restrict__ rowmax, const int* __restrict__ idx_prev,
       float* __restrict__ dens_part /*[n_tot][tiles][S][2]*/) {
  __shared__ float tok_s[SIM_CK][SIM_TOK + 1];
  __shared__ __align__(16) float ph_s[SIM_SG][SIM_CK];
  __shared__ float red_s[SIM_TOK / 32][SIM_SG][2];
  const int o = blockIdx.z, sg = blockIdx.y * SIM_SG, tile = blockIdx.x;
  const int img = obj_img[o];
  const int n0 = tile * SIM_TOK;
  const int tid = threadIdx.x;
  const int n = n0 + tid;
  const float* fimg = feats + img * fstride;
  float acc[SIM_SG];
#pragma unroll
  for (int s = 0; s < SIM_SG; ++s) acc[s] = 0.f;

  for (int c0 = 0; c0 < C; c0 += SIM_CK) {
    // stage tokens: coalesced rows of SIM_CK floats, stored transposed
    for (int i = tid; i < SIM_TOK * (SIM_CK / 4); i += SIM_TOK) {
      const int t = i / (SIM_CK / 4), q = i - t * (SIM_CK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + t < N && c0 + 4 * q < C) v = __ldg(reinterpret_cast<const float4*>(fimg + (long long)(n0 + t) * C + c0) + q);
      tok_s[4 * q + 0][t] = v.x; tok_s[4 * q + 1][t] = v.y; tok_s[4 * q + 2][t] = v.z; tok_s[4 * q + 3][t] = v.w;
    }
    for (int i = tid; i < SIM_SG * SIM_CK; i += SIM_TOK) {
      const int s = i / SIM_CK, c = i - s * SIM_CK;
      ph_s[s][c] = (sg + s < S && c0 + c < C) ? phat[((long long)o * S + sg + s) * C + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < SIM_CK; c += 4) {
      const float t0 = tok_s[c][tid], t1 = tok_s[c + 1][tid], t2 = tok_s[c + 2][tid], t3 = tok_s[c + 3][tid];
#pragma unroll
      for (int s = 0; s < SIM_SG; ++s) {
        const float4 p = *reinterpret_cast<const float4*>(&ph_s[s][c]);
        acc[s] = fmaf(t0, p.x, acc[s]);
        acc[s] = fmaf(t1, p.y, acc[s]);
        acc[s] = fmaf(t2, p.z, acc[s]);
        acc[s] = fmaf(t3, p.w, acc[s]);
      }
    }
    __syncthreads();
  }

  const bool valid = n < N;
  bool inside = true;
  if (masked) {
    const Box b = patch_box(rois + 4 * o, hp, wp);
    inside = valid && in_box(b, n, wp);
  }
  const float d = valid ? den[(long long)img * N + n] : 1.f;
  const int prev = (idx_prev && valid) ? idx_prev[(long long)o * N + n] : -1;
  const int w = tid >> 5;
#pragma unroll
  for (int s = 0; s < SIM_SG; ++s) {
    float v = inside ? acc[s] / d : 0.f;
    if (clamp0) v = fmaxf(v, 0.f);
    if (valid && sg + s < S) sim[((long long)o * S + sg + s) * N + n] = v;
    if (rowmax) {
      const float m = warp_max(valid ? v : -FLT_MAX);
      if (lane_id() == 0) red_s[w][s][0] = m;
    }
  }
  if (rowmax) {
    __syncthreads();
    if (tid < SIM_SG && sg + tid < S) {
      float m = red_s[0][tid][0];
      for (int i = 1; i < SIM_TOK / 32; ++i) m = fmaxf(m, red_s[i][tid][0]);
      atomicMax(rowmax + (long long)o * S + sg + tid, enc_f(m));
    }
    __syncthreads();
  }
  if (idx_prev) {
    // density partials of the previous assignment: sum of sim over tokens assigned to seed s, and their count
#pragma unroll
    for (int s = 0; s < SIM_SG; ++s) {
      const bool mine = (prev == sg + s);
      const float v = inside ? acc[s] / d : 0.f;
      const float sv = warp_sum(mine ? v : 0.f);
      const float cv = warp_sum(mine ? 1.f : 0.f);
      if (lane_id() == 0) { red_s[w][s][0] = sv; red_s[w][s][1] = cv; }
    }
    __syncthreads();
    if (tid < SIM_SG && sg + tid < S) {
      float sv = 0.f, cv = 0.f;
      for (int i = 0; i < SIM_TOK / 32; ++i) { sv += red_s[i][tid][0]; cv += red_s[i][tid][1]; }
      float* dp = dens_part + (((long long)o * gridDim.x + tile) * S + sg + tid) * 2;
      dp[0] = sv; dp[1] = cv;
    }
  }
}

// grid (S, n_tot): tau, logit max, softmax denominator
__global__ void ms_stats(const float* __restrict__ sim, const unsigned* __restrict__ rowmax,
                         const float* __restrict__ dens_part, int tiles, int N, int S, int first, float tt0, float temp,
                         float* __restrict__ stat /*[n_tot][S][4] = tt, lmax, Z, tau*/) {
  __shared__ float red[32];
  const int s = blockIdx.x, o = blockIdx.y;
  float tt, tau;
  if (first) {
    tt = tt0; tau = 0.f;
  } else {
    float tot = 0.f, cnt = 0.f;
    for (int t = 0; t < tiles; ++t) {
      const float* dp = dens_part + (((long long)o * tiles + t) * S + s) * 2;
      tot += dp[0]; cnt += dp[1];
    }
    const float dsty = 1.f - (cnt >= 1.f ? tot / cnt : 0.f);
    tau = fmaxf(dsty, 1e-10f);
    tt = temp * tau;
  }
  const float lmax = dec_f(rowmax[(long long)o * S + s]) / tt;
  const float* row = sim + ((long long)o * S + s) * N;
  float z = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) z += expf(row[n] / tt - lmax);
  z = block_sum(z, red);
  if (threadIdx.x == 0) {
    float* st = stat + ((long long)o * S + s) * 4;
    st[0] = tt; st[1] = lmax; st[2] = z; st[3] = tau;
  }
}

// grid (ceil(N/256), n_tot): hard assignment = argmax_s softmax weight (first maximum wins, like torch.argmax)
__global__ void ms_assign(const float* __restrict__ sim, const float* __restrict__ stat, int N, int S,
                          int* __restrict__ idx, float* __restrict__ wsel, int* __restrict__ trace) {
  extern __shared__ float st_s[];
  const int o = blockIdx.y;
  for (int i = threadIdx.x; i < S * 4; i += blockDim.x) st_s[i] = stat[(long long)o * S * 4 + i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float best = -1.f;
  int bi = 0;
  for (int s = 0; s < S; ++s) {
    const float w = expf(sim[((long long)o * S + s) * N + n] / st_s[4 * s] - st_s[4 * s + 1]) / st_s[4 * s + 2];
    if (w > best) { best = w; bi = s; }
  }
  idx[(long long)o * N + n] = bi;
  wsel[(long long)o * N + n] = best;
  if (trace) trace[(long long)o * N + n] = bi;
}

// grid (tiles_u, n_tot), dynamic smem S*CC floats; thread owns channels c = chunk + tid + j*blockDim
__global__ void __launch_bounds__(256)
ms_update(const float* __restrict__ feats, long long fstride, const int* __restrict__ obj_img,
          const float* __restrict__ rois, const int* __restrict__ idx, const float* __restrict__ wsel, int N, int C,
          int hp, int wp, int S, int CC, float* __restrict__ part /*[n_tot][tiles_u][S][C]*/) {
  extern __shared__ float acc_s[];
  __shared__ int idx_s[UPD_TOK];
  __shared__ float w_s[UPD_TOK];
  const int o = blockIdx.y, tile = blockIdx.x;
  const int img = obj_img[o];
  const int n0 = tile * UPD_TOK;
  const Box b = patch_box(rois + 4 * o, hp, wp);
  for (int t = threadIdx.x; t < UPD_TOK; t += blockDim.x) {
    const int n = n0 + t;
    const bool use = n < N && in_box(b, n, wp);
    idx_s[t] = use ? idx[(long long)o * N + n] : -1;
    w_s[t] = use ? wsel[(long long)o * N + n] : 0.f;
  }
  const float* fimg = feats + img * fstride;
  float* dst = part + ((long long)o * gridDim.x + tile) * S * C;
  for (int cbase = 0; cbase < C; cbase += CC) {
    const int cw = min(CC, C - cbase);
    for (int i = threadIdx.x; i < S * CC; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    for (int t = 0; t < UPD_TOK; ++t) {
      const int row = idx_s[t];
      const float w = w_s[t];
      if (row < 0 || w == 0.f) continue;   // block-uniform
      const float* f = fimg + (long long)(n0 + t) * C + cbase;
      for (int c = threadIdx.x; c < cw; c += blockDim.x) acc_s[row * CC + c] = fmaf(w, __ldg(f + c), acc_s[row * CC + c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * cw; i += blockDim.x) {
      const int s = i / cw, c = i - s * cw;
      dst[(long long)s * C + cbase + c] = acc_s[s * CC + c];
    }
    __syncthreads();
  }
}

// grid (n_tot*S): ordered reduction over tiles, write p and phat
__global__ void ms_finish(const float* __restrict__ part, int tiles, int S, int C, float* __restrict__ proto,
                          float* __restrict__ phat) {
  __shared__ float red[32];
  const int row = blockIdx.x, o = row / S, s = row - o * S;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int t = 0; t < tiles; ++t) v += part[(((long long)o * tiles + t) * S + s) * C + c];
    proto[(long long)row * C + c] = v;
    ss += v * v;
  }
  const float nrm = fmaxf(sqrtf(block_sum(ss, red)), 1e-8f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) phat[(long long)row * C + c] = proto[(long long)row * C + c] / nrm;
}

__global__ void ms_fill_u32(unsigned* p, unsigned v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Deterministic seed grid of RH:1786-1810 + feature gather RH:1810.  One CTA per instance.
__global__ void __launch_bounds__(256)
ms_grid_seeds(const float* __restrict__ maps /*[n_tot][N]*/, float thr, const float* __restrict__ feats, long long fstride,
              const int* __restrict__ obj_img, const float* __restrict__ rois, int N, int C, int wp, int S,
              int* __restrict__ seed_tok /*[n_tot][S]*/, float* __restrict__ proto /*[n_tot][S][C]*/) {
  extern __shared__ int list_s[];          // compacted positive token ids (N ints) + S seed ids
  __shared__ int cnt_s[256];
  __shared__ int total_s;
  const int o = blockIdx.x;
  const float* m = maps + (long long)o * N;
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(N, beg + per);
  int c = 0;
  for (int n = beg; n < end; ++n) c += (m[n] >= thr);
  cnt_s[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { const int t = cnt_s[i]; cnt_s[i] = run; run += t; }
    total_s = run;
  }
  __syncthreads();
  int pos = cnt_s[threadIdx.x];
  for (int n = beg; n < end; ++n) if (m[n] >= thr) list_s[pos++] = n;
  __syncthreads();
  const int num = total_s;
  int* seeds = list_s + N;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    int tok;
    if (num >= S) {
      tok = list_s[j * (num / S)];
    } else if (num > 0) {                       // fill_in_idx (RH:1147-1155) unrolled into a closed form
      const int r1 = num * max(1, S / num);
      const int jj = j < r1 ? j : j - r1;
      tok = list_s[jj % num];
    } else {                                    // box centre (RH:1799-1800)
      const float* r = rois + 4 * o;
      const int cx = (int)floorf((r[0] + r[2]) / 32.f), cy = (int)floorf((r[1] + r[3]) / 32.f);
      tok = cy * wp + cx;
    }
    seeds[j] = tok;
    seed_tok[(long long)o * S + j] = tok;
  }
  __syncthreads();
  const float* fimg = feats + obj_img[o] * fstride;
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) {
    const int j = i / C, cc = i - j * C;
    proto[((long long)o * S + j) * C + cc] = fimg[(long long)seeds[j] * C + cc];
  }
}

struct Ws {
  float *den, *phat, *dens, *stat, *wsel, *part;
  unsigned* rowmax;
  int* idx;
  int tiles_s, tiles_u;
  size_t bytes;
};
Ws carve(void* base, int n_img, int n_tot, int S, int N, int C) {
  Ws w;
  w.tiles_s = (N + SIM_TOK - 1) / SIM_TOK;
  w.tiles_u = (N + UPD_TOK - 1) / UPD_TOK;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return (char*)base + o; };
  w.den = (float*)take((size_t)n_img * N * 4);
  w.phat = (float*)take((size_t)n_tot * S * C * 4);
  w.dens = (float*)take((size_t)n_tot * w.tiles_s * S * 2 * 4);
  w.stat = (float*)take((size_t)n_tot * S * 4 * 4);
  w.wsel = (float*)take((size_t)n_tot * N * 4);
  w.idx = (int*)take((size_t)n_tot * N * 4);
  w.rowmax = (unsigned*)take((size_t)n_tot * S * 4);
  w.part = (float*)take((size_t)n_tot * w.tiles_u * S * C * 4);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" size_t as_mean_shift_workspace(int n_img, int n_tot, int S, int N, int C) {
  return carve(nullptr, n_img, n_tot, S, N, C).bytes;
}

// Seeds: maps [n_tot,N] (1/0 foreground on the patch grid) -> seed token ids [n_tot,S] and prototypes [n_tot,S,C]
extern "C" int as_grid_seeds(const float* maps, float thr, const float* feats, long long feat_img_stride,
                             const int* obj_img, const float* rois, int n_tot, int N, int C, int wp, int S,
                             int* seed_tok, float* proto, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  const size_t smem = (size_t)(N + S) * 4;
  if (smem > 200 * 1024) return AS_ERR_BAD_ARG;
  AS_CUDA(cudaFuncSetAttribute(ms_grid_seeds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ms_grid_seeds<<<n_tot, 256, smem, stream>>>(maps, thr, feats, feat_img_stride, obj_img, rois, N, C, wp, S, seed_tok, proto);
  AS_LAUNCH_CHECK();
  return 0;
}

// The mean-shift loop.  proto [n_tot,S,C] in/out, sim [n_tot,S,N] out (vs unmasked tokens; clamped at 0 when clamp0),
// trace (optional) [n_shift,n_tot,N] int32 hard assignments, tau_out (optional) [n_tot,S] last bandwidths.
extern "C" int as_mean_shift(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                             const int* obj_img, const float* rois, int n_tot, int S, float* proto, float* sim,
                             int n_shift, double tau0, double temp, int clamp0, int* trace, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  if (C % 4 || hp * wp != N) return AS_ERR_BAD_ARG;
  Ws w = carve(workspace, n_img, n_tot, S, N, C);
  if (workspace_bytes < w.bytes) return AS_ERR_BAD_ARG;
  int CC = C;
  while ((size_t)S * CC * 4 > 160 * 1024) CC = (CC + 1) / 2;
  const size_t upd_smem = (size_t)S * CC * 4;
  AS_CUDA(cudaFuncSetAttribute(ms_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  const dim3 gsim(w.tiles_s, (S + SIM_SG - 1) / SIM_SG, n_tot);
  ms_token_norm<<<dim3((N + 7) / 8, n_img), 256, 0, stream>>>(feats, feat_img_stride, N, C, w.den);
  ms_proto_norm<<<n_tot * S, 256, 0, stream>>>(proto, w.phat, C);
  const float tt0 = (float)(temp * tau0);   // python computes temp*tau in double, torch divides by float32(that)
  for (int it = 0; it < n_shift; ++it) {
    ms_fill_u32<<<(n_tot * S + 255) / 256, 256, 0, stream>>>(w.rowmax, 0u, n_tot * S);
    ms_sim<<<gsim, SIM_TOK, 0, stream>>>(feats, feat_img_stride, w.den, obj_img, rois, w.phat, N, C, hp, wp, S, 1, 0, sim,
                                         w.rowmax, it ? w.idx : nullptr, w.dens);
    ms_stats<<<dim3(S, n_tot), 256, 0, stream>>>(sim, w.rowmax, w.dens, w.tiles_s, N, S, it == 0, tt0, (float)temp, w.stat);
    ms_assign<<<dim3((N + 255) / 256, n_tot), 256, S * 16, stream>>>(sim, w.stat, N, S, w.idx, w.wsel,
                                                                     trace ? trace + (size_t)it * n_tot * N : nullptr);
    ms_update<<<dim3(w.tiles_u, n_tot), 256, upd_smem, stream>>>(feats, feat_img_stride, obj_img, rois, w.idx, w.wsel, N, C,
                                                                hp, wp, S, CC, w.part);
    ms_finish<<<n_tot * S, 256, 0, stream>>>(w.part, w.tiles_u, S, C, proto, w.phat);
  }
  ms_sim<<<gsim, SIM_TOK, 0, stream>>>(feats, feat_img_stride, w.den, obj_img, rois, w.phat, N, C, hp, wp, S, 0, clamp0, sim,
                                       nullptr, nullptr, nullptr);
  AS_LAUNCH_CHECK();
  return 0;
}

// Generic cosine maps: sim[g,s,n] = cos(protos[g,s,:], feats[grp_img[g], n, :])  (F.cosine_similarity semantics).
// Used for the seed-prototype maps (RH:339), the refinement maps (RH:696) and the part maps (RH:297-301).
// Cosine maps for a handful of seeds per group (refinement loop, part maps): one WARP per token.  The warp reads the
// token row once (coalesced float4, the next row already in flight), takes its norm on the way, and dots it with every
// seed of the group held in shared memory; CW_TPW consecutive tokens per warp, results parked in lane = token.
// grid (ceil(N / CW_TOK), ceil(S / 16), G), 256 threads; C % 128 == 0, C <= 1024.
constexpr int CW_SEEDS = 16;
constexpr int CW_TPW = 8;                              // tokens per warp
constexpr int CW_TOK = 8 * CW_TPW;                     // tokens per CTA
__global__ void __launch_bounds__(256)
cos_warp_rows(const float* __restrict__ feats, long long fstride, const int* __restrict__ grp_img, const float* __restrict__ phat,
              int N, int C, int S, int clamp0, float* __restrict__ sim) {
  extern __shared__ float4 ps4[];                    // [CW_SEEDS][C / 4]
  const int g = blockIdx.z, s0 = blockIdx.y * CW_SEEDS;
  const int ns = min(CW_SEEDS, S - s0);
  const int c4n = C / 4, per = C / 128;              // float4 per row; float4 per lane
  for (int i = threadIdx.x; i < ns * c4n; i += blockDim.x)
    ps4[i] = reinterpret_cast<const float4*>(phat + ((size_t)g * S + s0) * C)[i];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int n0 = blockIdx.x * CW_TOK + warp * CW_TPW;
  const float* fimg = feats + grp_img[g] * fstride;
  float4 fn[8];
  auto fetch = [&](int n) {
    const float4* row = reinterpret_cast<const float4*>(fimg + (size_t)min(n, N - 1) * C);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) fn[i] = __ldg(row + i * 32 + lane);
  };
  fetch(n0);
  __syncthreads();
  float res[CW_SEEDS];
#pragma unroll
  for (int s = 0; s < CW_SEEDS; ++s) res[s] = 0.f;
#pragma unroll 1
  for (int t = 0; t < CW_TPW; ++t) {
    float4 fv[8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) {
        fv[i] = fn[i];
        ss += fv[i].x * fv[i].x + fv[i].y * fv[i].y + fv[i].z * fv[i].z + fv[i].w * fv[i].w;
      }
    if (t + 1 < CW_TPW) fetch(n0 + t + 1);
    const float d = fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
#pragma unroll
    for (int s = 0; s < CW_SEEDS; ++s) {
      if (s >= ns) break;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < per) {
          const float4 pv = ps4[s * c4n + i * 32 + lane];
          a = fmaf(fv[i].x, pv.x, a); a = fmaf(fv[i].y, pv.y, a); a = fmaf(fv[i].z, pv.z, a); a = fmaf(fv[i].w, pv.w, a);
        }
      a = warp_sum(a);
      if (lane == t) res[s] = a / d;
    }
  }
  const int n = n0 + lane;
  if (lane < CW_TPW && n < N) {
#pragma unroll
    for (int s = 0; s < CW_SEEDS; ++s) {
      if (s >= ns) break;
      sim[((size_t)g * S + s0 + s) * N + n] = clamp0 ? fmaxf(res[s], 0.f) : res[s];
    }
  }
}

extern "C" size_t as_cosine_maps_workspace(int n_img, int G, int S, int N, int C) {
  return (((size_t)n_img * N * 4 + 255) & ~(size_t)255) + (size_t)G * S * C * 4;
}
extern "C" int as_cosine_maps(const float* feats, long long feat_img_stride, int n_img, int N, int C, const int* grp_img,
                              const float* protos, int G, int S, float* sim, int clamp0, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
  if (G <= 0) return 0;
  if (C % 4 || workspace_bytes < as_cosine_maps_workspace(n_img, G, S, N, C)) return AS_ERR_BAD_ARG;
  float* den = (float*)workspace;
  float* phat = (float*)((char*)workspace + (((size_t)n_img * N * 4 + 255) & ~(size_t)255));
  ms_proto_norm<<<G * S, 256, 0, stream>>>(protos, phat, C);
  if (C % 128 == 0 && C <= 1024) {
    const size_t smem = (size_t)CW_SEEDS * C * 4;
    static bool attr = false;
    if (!attr) {
      AS_CUDA(cudaFuncSetAttribute(cos_warp_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SEEDS * 1024 * 4));
      attr = true;
    }
    cos_warp_rows<<<dim3((N + CW_TOK - 1) / CW_TOK, (S + CW_SEEDS - 1) / CW_SEEDS, G), 256, smem, stream>>>(
        feats, feat_img_stride, grp_img, phat, N, C, S, clamp0, sim);
  } else {
    ms_token_norm<<<dim3((N + 7) / 8, n_img), 256, 0, stream>>>(feats, feat_img_stride, N, C, den);
    ms_sim<<<dim3((N + SIM_TOK - 1) / SIM_TOK, (S + SIM_SG - 1) / SIM_SG, G), SIM_TOK, 0, stream>>>(
        feats, feat_img_stride, den, grp_img, nullptr, phat, N, C, N, 1, S, 0, clamp0, sim, nullptr, nullptr, nullptr);
  }
  AS_LAUNCH_CHECK();
  return 0;
}
