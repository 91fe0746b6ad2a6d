// Attention-shift loop, tensor-core variant (as_mean_shift_tc).  Same algorithm and outputs as as_mean_shift
// (cosine_shift_batch RH:830-854 + update_density_batch RH:882-908, RH = stdroi_point_deform_attn_reppoints.py), but
//  * the token-to-seed affinity  sim[n, s] = <f^[n], p^[s]>  of a whole image (all its instances' seeds at once) is a
//    batched tcgen05 GEMM on split-fp16 operands: f^ * 2^10 = hi + lo in fp16, three MMAs (hi.hi + hi.lo + lo.hi) with
//    fp32 accumulation give ~2^-22 relative accuracy -- the fp32 CUDA-core product was FFMA-bound at 72 FLOP/B;
//  * similarities are kept token-major [img][N][LD] so every later phase reads a token's whole row contiguously;
//  * the per-seed statistics are reduced by the consumers themselves (no extra launches), the update kernel keeps
//    8 token loads in flight and only touches its shared-memory accumulator row when the assigned seed changes.
// Per iteration: split_protos -> 3 x bgemm -> colstats -> zpart -> assign -> update -> finish   (no host sync anywhere).
#include "common.cuh"
#include <float.h>

using namespace asb;

extern "C" int as_bgemm_f16_f32(const void* x_f16, const void* w_f16, float* out, const float* resid, int batch, int M,
                                int N, int K, int x_rows, int w_rows, int ldo, long long out_bstride, float alpha,
                                cudaStream_t stream);

namespace {

constexpr float OP_SCALE = 1024.f;          // 2^10: normalised components (<= 1) as split fp16 away from the subnormals
constexpr int CS_TOK = 64;                  // tokens per CTA in the column-statistics / Z / assign kernels (small: many CTAs hide latency)
constexpr int UP_TOK = 256;                 // tokens per CTA in the update kernel

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
__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ void split_store(float v, __half* hi, __half* lo, size_t i) {
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

// f^ = f / max(|f|, 1e-8) (F.cosine_similarity normalises each operand first), scaled and split.  warp per token
__global__ void tc_split_tokens(const float* __restrict__ feats, long long fstride, int N, int C, __half* __restrict__ hi,
                                __half* __restrict__ lo) {
  const int img = blockIdx.y;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* f = feats + img * fstride + (long long)n * C;
  float ss = 0.f;
  for (int c = lane_id(); c < C; c += 32) ss += f[c] * f[c];
  const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
  const size_t o = ((size_t)img * N + n) * C;
  for (int c = lane_id(); c < C; c += 32) split_store(f[c] / nrm * OP_SCALE, hi, lo, o + c);
}

// seeds of image `img` as rows [0, nobj*S) of its LD-row operand block (remaining rows zero).  grid (LD, n_img)
__global__ void tc_split_protos(const float* __restrict__ proto, const int* __restrict__ img_first, const int* __restrict__ img_nobj,
                                int S, int C, int LD, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float red[8];
  const int r = blockIdx.x, img = blockIdx.y;
  const size_t o = ((size_t)img * LD + r) * C;
  if (r >= img_nobj[img] * S) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) { hi[o + c] = __float2half_rn(0.f); lo[o + c] = __float2half_rn(0.f); }
    return;
  }
  const float* p = proto + ((size_t)img_first[img] * S + r) * C;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) ss += p[c] * p[c];
  ss = warp_sum(ss);
  if (lane_id() == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float nrm = fmaxf(sqrtf(tot), 1e-8f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) split_store(p[c] / nrm * OP_SCALE, hi, lo, o + c);
}

// masked similarity of (token n, column s) of an image: 0 outside the instance's box
struct ImgCtx {
  const float* rois; const int* img_first; const int* img_nobj;
  int S, hp, wp, N, LD;
};
// column statistics over a 256-token tile: max of the masked similarity, and the density partials of the previous
// assignment.  grid (tiles, n_img), 256 threads = (256 / LD) token groups x LD columns (LD power of two in [32, 256])
__global__ void __launch_bounds__(256)
tc_colstats(const float* __restrict__ sim, ImgCtx c, const int* __restrict__ idx_prev, unsigned* __restrict__ colmax,
            float* __restrict__ dens_part /*[n_img][tiles][LD][2]*/) {
  __shared__ float red[256][3];
  const int img = blockIdx.y, tile = blockIdx.x;
  const int col = threadIdx.x % c.LD, grp = threadIdx.x / c.LD, ngrp = 256 / c.LD;
  const int kb = c.img_nobj[img] * c.S;
  float mx = -FLT_MAX, sv = 0.f, cv = 0.f;
  if (col < kb) {
    const int o = c.img_first[img] + col / c.S, s = col % c.S;
    const Box b = patch_box(c.rois + 4 * o, c.hp, c.wp);
    const float* sp = sim + (size_t)img * c.N * c.LD + col;
    for (int t = grp; t < CS_TOK; t += ngrp) {
      const int n = tile * CS_TOK + t;
      if (n >= c.N) break;
      const float v = in_box(b, n, c.wp) ? sp[(size_t)n * c.LD] : 0.f;
      mx = fmaxf(mx, v);
      if (idx_prev && idx_prev[(size_t)o * c.N + n] == s) { sv += v; cv += 1.f; }
    }
  }
  red[threadIdx.x][0] = mx; red[threadIdx.x][1] = sv; red[threadIdx.x][2] = cv;
  __syncthreads();
  if (grp == 0 && col < kb) {
    for (int g = 1; g < ngrp; ++g) {              // fixed order
      mx = fmaxf(mx, red[g * c.LD + col][0]); sv += red[g * c.LD + col][1]; cv += red[g * c.LD + col][2];
    }
    atomicMax(colmax + (size_t)img * c.LD + col, enc_f(mx));
    float* dp = dens_part + (((size_t)img * gridDim.x + tile) * c.LD + col) * 2;
    dp[0] = sv; dp[1] = cv;
  }
}

// per-column temperature and logit max, recomputed identically by every CTA that needs them
__device__ __forceinline__ void column_stats(const ImgCtx& c, int img, int col, const unsigned* colmax, const float* dens_part,
                                            int tiles, int first, float tt0, float temp, float& tt, float& lmax, float& tau) {
  if (first) { tt = tt0; tau = 0.f; }
  else {
    float tot = 0.f, cnt = 0.f;
    for (int t = 0; t < tiles; ++t) {
      const float* dp = dens_part + (((size_t)img * tiles + t) * c.LD + col) * 2;
      tot += dp[0]; cnt += dp[1];
    }
    tau = fmaxf(1.f - (cnt >= 1.f ? tot / cnt : 0.f), 1e-10f);       // RH:883-885, 908
    tt = temp * tau;
  }
  lmax = dec_f(colmax[(size_t)img * c.LD + col]) / tt;
}

// partial softmax denominators.  grid (tiles, n_img)
__global__ void __launch_bounds__(256)
tc_zpart(const float* __restrict__ sim, ImgCtx c, const unsigned* __restrict__ colmax, const float* __restrict__ dens_part,
         int first, float tt0, float temp, float* __restrict__ z_part /*[n_img][tiles][LD]*/, float* __restrict__ stat /*[n_img][LD][4]*/) {
  __shared__ float red[256];
  const int img = blockIdx.y, tile = blockIdx.x, tiles = gridDim.x;
  const int col = threadIdx.x % c.LD, grp = threadIdx.x / c.LD, ngrp = 256 / c.LD;
  const int kb = c.img_nobj[img] * c.S;
  float z = 0.f;
  if (col < kb) {
    float tt, lmax, tau;
    column_stats(c, img, col, colmax, dens_part, tiles, first, tt0, temp, tt, lmax, tau);
    if (tile == 0 && grp == 0) { float* st = stat + ((size_t)img * c.LD + col) * 4; st[0] = tt; st[1] = lmax; st[3] = tau; }
    const int o = c.img_first[img] + col / c.S;
    const Box b = patch_box(c.rois + 4 * o, c.hp, c.wp);
    const float* sp = sim + (size_t)img * c.N * c.LD + col;
    for (int t = grp; t < CS_TOK; t += ngrp) {
      const int n = tile * CS_TOK + t;
      if (n >= c.N) break;
      const float v = in_box(b, n, c.wp) ? sp[(size_t)n * c.LD] : 0.f;
      z += expf(v / tt - lmax);
    }
  }
  red[threadIdx.x] = z;
  __syncthreads();
  if (grp == 0 && col < kb) {
    for (int g = 1; g < ngrp; ++g) z += red[g * c.LD + col];
    z_part[((size_t)img * tiles + tile) * c.LD + col] = z;
  }
}

// hard assignment per (instance, token).  grid (tiles, n_img), CS_TOK threads = one per token, the tile's rows staged in smem
__global__ void __launch_bounds__(CS_TOK)
tc_assign(const float* __restrict__ sim, ImgCtx c, const float* __restrict__ stat, const float* __restrict__ z_part, int tiles,
          int* __restrict__ idx, float* __restrict__ wsel, int* __restrict__ trace) {
  extern __shared__ float sm[];
  float* rows = sm;                                   // [256][LD + 1]
  float* st_s = sm + CS_TOK * (c.LD + 1);             // [LD][3] = tt, lmax, Z
  const int img = blockIdx.y, tile = blockIdx.x;
  const int kb = c.img_nobj[img] * c.S;
  for (int col = threadIdx.x; col < c.LD; col += blockDim.x) {
    float z = 0.f;
    if (col < kb) for (int t = 0; t < tiles; ++t) z += z_part[((size_t)img * tiles + t) * c.LD + col];     // fixed order
    st_s[3 * col] = stat[((size_t)img * c.LD + col) * 4];
    st_s[3 * col + 1] = stat[((size_t)img * c.LD + col) * 4 + 1];
    st_s[3 * col + 2] = z;
  }
  const float* sp = sim + ((size_t)img * c.N + (size_t)tile * CS_TOK) * c.LD;
  for (int i = threadIdx.x; i < CS_TOK * c.LD; i += blockDim.x) {
    const int t = i / c.LD, col = i - t * c.LD;
    rows[t * (c.LD + 1) + col] = (tile * CS_TOK + t < c.N) ? sp[i] : 0.f;
  }
  __syncthreads();
  const int n = tile * CS_TOK + threadIdx.x;
  if (n >= c.N) return;
  const float* my = rows + threadIdx.x * (c.LD + 1);
  for (int j = 0; j < c.img_nobj[img]; ++j) {
    const int o = c.img_first[img] + j;
    const Box b = patch_box(c.rois + 4 * o, c.hp, c.wp);
    const bool inside = in_box(b, n, c.wp);
    float best = -1.f;
    int bi = 0;
    for (int s = 0; s < c.S; ++s) {
      const int col = j * c.S + s;
      const float v = inside ? my[col] : 0.f;
      const float w = expf(v / st_s[3 * col] - st_s[3 * col + 1]) / st_s[3 * col + 2];
      if (w > best) { best = w; bi = s; }          // first maximum wins (torch.argmax)
    }
    idx[(size_t)o * c.N + n] = bi;
    wsel[(size_t)o * c.N + n] = best;
    if (trace) trace[(size_t)o * c.N + n] = bi;
  }
}

// partial new prototypes: part[o][tile][s][c] = sum over the tile's in-box tokens assigned to s of w * f.
// grid (tiles_u, n_tot); thread owns channels c = tid + j*256; 8 token loads in flight; the shared-memory accumulator
// row is only touched when the assigned seed changes (tokens are visited in raster order, assignments are coherent)
template <int J>
__global__ void __launch_bounds__(256)
tc_update(const float* __restrict__ feats, long long fstride, const int* __restrict__ obj_img, const float* __restrict__ rois,
          const int* __restrict__ idx, const float* __restrict__ wsel, int N, int C, int hp, int wp, int S,
          float* __restrict__ part) {
  extern __shared__ float acc_s[];                   // [S][C]
  __shared__ int idx_s[UP_TOK];
  __shared__ float w_s[UP_TOK];
  __shared__ int list_s[UP_TOK];
  __shared__ int cnt_s;
  const int o = blockIdx.y, tile = blockIdx.x, n0 = tile * UP_TOK;
  const Box b = patch_box(rois + 4 * o, hp, wp);
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) acc_s[i] = 0.f;
  for (int t = threadIdx.x; t < UP_TOK; t += blockDim.x) {
    const int n = n0 + t;
    const bool use = n < N && in_box(b, n, wp);
    idx_s[t] = use ? idx[(size_t)o * N + n] : -1;
    w_s[t] = use ? wsel[(size_t)o * N + n] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                            // compact the contributing tokens, order preserved
    int k = 0;
    for (int t = 0; t < UP_TOK; ++t) if (idx_s[t] >= 0 && w_s[t] != 0.f) list_s[k++] = t;
    cnt_s = k;
  }
  __syncthreads();
  const int cnt = cnt_s;
  const float* fimg = feats + obj_img[o] * fstride;
  float run[J];
  int cur = -1;
#pragma unroll
  for (int j = 0; j < J; ++j) run[j] = 0.f;
  for (int k0 = 0; k0 < cnt; k0 += 8) {
    float v[8][J];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = list_s[min(k0 + q, cnt - 1)];
      const float* f = fimg + (size_t)(n0 + t) * C;
#pragma unroll
      for (int j = 0; j < J; ++j) { const int ch = threadIdx.x + j * 256; v[q][j] = ch < C ? __ldg(f + ch) : 0.f; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (k0 + q >= cnt) break;
      const int t = list_s[k0 + q];
      const int row = idx_s[t];
      const float w = w_s[t];
      if (row != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int j = 0; j < J; ++j) { const int ch = threadIdx.x + j * 256; if (ch < C) acc_s[cur * C + ch] += run[j]; }
        }
        cur = row;
#pragma unroll
        for (int j = 0; j < J; ++j) run[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < J; ++j) run[j] = fmaf(w, v[q][j], run[j]);
    }
  }
  if (cur >= 0) {
#pragma unroll
    for (int j = 0; j < J; ++j) { const int ch = threadIdx.x + j * 256; if (ch < C) acc_s[cur * C + ch] += run[j]; }
  }
  __syncthreads();
  float* dst = part + ((size_t)o * gridDim.x + tile) * S * C;
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) dst[i] = acc_s[i];
}

// ordered reduction of the partials -> new prototype row.  grid (n_tot*S)
__global__ void tc_finish(const float* __restrict__ part, int tiles, int S, int C, float* __restrict__ proto) {
  const int row = blockIdx.x, o = row / S, s = row - o * S;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int t = 0; t < tiles; ++t) v += part[(((size_t)o * tiles + t) * S + s) * C + c];
    proto[(size_t)row * C + c] = v;
  }
}

// token-major [img][N][LD] -> seed-major [n_tot][S][N] output (optionally clamped at 0, RH:1840)
__global__ void tc_transpose_out(const float* __restrict__ sim, const int* __restrict__ obj_img, const int* __restrict__ img_first,
                                 int S, int N, int LD, int clamp0, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int o = blockIdx.z, img = obj_img[o], j = o - img_first[img];
  const int s0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, s = s0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && s < S) ? sim[((size_t)img * N + n) * LD + j * S + s] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = s0 + r, n = n0 + threadIdx.x;
    if (s < S && n < N) {
      float v = tile[threadIdx.x][r];
      if (clamp0) v = fmaxf(v, 0.f);
      out[((size_t)o * S + s) * N + n] = v;
    }
  }
}
__global__ void tc_fill_u32(unsigned* p, unsigned v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int round_ld(int kmax) {
  int ld = 32;
  while (ld < kmax) ld <<= 1;
  return ld;
}

struct Ws2 {
  __half *fhi, *flo, *phi, *plo;
  float *sim, *dens, *zpart, *stat, *wsel, *part;
  unsigned* colmax;
  int* idx;
  int tiles, tiles_u, LD;
  size_t bytes;
};
Ws2 carve2(void* base, int n_img, int n_tot, int S, int N, int C, int kmax) {
  Ws2 w;
  w.LD = round_ld(kmax);
  w.tiles = (N + CS_TOK - 1) / CS_TOK;
  w.tiles_u = (N + UP_TOK - 1) / UP_TOK;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return (char*)base + o; };
  w.fhi = (__half*)take((size_t)n_img * N * C * 2);
  w.flo = (__half*)take((size_t)n_img * N * C * 2);
  w.phi = (__half*)take((size_t)n_img * w.LD * C * 2);
  w.plo = (__half*)take((size_t)n_img * w.LD * C * 2);
  w.sim = (float*)take((size_t)n_img * N * w.LD * 4);
  w.dens = (float*)take((size_t)n_img * w.tiles * w.LD * 2 * 4);
  w.zpart = (float*)take((size_t)n_img * w.tiles * w.LD * 4);
  w.stat = (float*)take((size_t)n_img * w.LD * 4 * 4);
  w.wsel = (float*)take((size_t)n_tot * N * 4);
  w.idx = (int*)take((size_t)n_tot * N * 4);
  w.colmax = (unsigned*)take((size_t)n_img * w.LD * 4);
  w.part = (float*)take((size_t)n_tot * w.tiles_u * S * C * 4);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" size_t as_mean_shift_tc_workspace(int n_img, int n_tot, int S, int N, int C, int kmax) {
  return carve2(nullptr, n_img, n_tot, S, N, C, kmax).bytes;
}

// Instances must be grouped by image (img_first[i] = first instance of image i, img_nobj[i] = its count; device arrays);
// kmax = max_i img_nobj[i] * S (host value).  Other arguments as as_mean_shift.
extern "C" int as_mean_shift_tc(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                                const int* obj_img, const int* img_first, const int* img_nobj, int kmax, const float* rois,
                                int n_tot, int S, float* proto, float* sim_out, int n_shift, double tau0, double temp,
                                int clamp0, int* trace, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  if (C % 64 || hp * wp != N || kmax > 256 || kmax < 1 || C > 256 * 4 || (size_t)S * C * 4 > 200 * 1024) return AS_ERR_BAD_ARG;
  Ws2 w = carve2(workspace, n_img, n_tot, S, N, C, kmax);
  if (workspace_bytes < w.bytes) return AS_ERR_BAD_ARG;
  const int LD = w.LD;
  ImgCtx ctx{rois, img_first, img_nobj, S, hp, wp, N, LD};
  const float tt0 = (float)(temp * tau0);
  const float alpha = 1.f / (OP_SCALE * OP_SCALE);
  const size_t upd_smem = (size_t)S * C * 4;
  const size_t asg_smem = ((size_t)CS_TOK * (LD + 1) + 3 * LD) * 4;
  const int J = (C + 255) / 256;
  if (J == 1) AS_CUDA(cudaFuncSetAttribute(tc_update<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  else if (J == 2) AS_CUDA(cudaFuncSetAttribute(tc_update<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  else if (J == 3) AS_CUDA(cudaFuncSetAttribute(tc_update<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  else AS_CUDA(cudaFuncSetAttribute(tc_update<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  AS_CUDA(cudaFuncSetAttribute(tc_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asg_smem));

  tc_split_tokens<<<dim3((N + 7) / 8, n_img), 256, 0, stream>>>(feats, feat_img_stride, N, C, w.fhi, w.flo);
  auto affinity = [&]() -> int {
    tc_split_protos<<<dim3(LD, n_img), 128, 0, stream>>>(proto, img_first, img_nobj, S, C, LD, w.phi, w.plo);
    const long long bs = (long long)N * LD;
    int r = as_bgemm_f16_f32(w.fhi, w.phi, w.sim, nullptr, n_img, N, LD, C, N, LD, LD, bs, alpha, stream);
    if (r) return r;
    r = as_bgemm_f16_f32(w.fhi, w.plo, w.sim, w.sim, n_img, N, LD, C, N, LD, LD, bs, alpha, stream);
    if (r) return r;
    return as_bgemm_f16_f32(w.flo, w.phi, w.sim, w.sim, n_img, N, LD, C, N, LD, LD, bs, alpha, stream);
  };
  const dim3 gt(w.tiles, n_img);
  for (int it = 0; it < n_shift; ++it) {
    int r = affinity();
    if (r) return r;
    tc_fill_u32<<<(n_img * LD + 255) / 256, 256, 0, stream>>>(w.colmax, 0u, n_img * LD);
    tc_colstats<<<gt, 256, 0, stream>>>(w.sim, ctx, it ? w.idx : nullptr, w.colmax, w.dens);
    tc_zpart<<<gt, 256, 0, stream>>>(w.sim, ctx, w.colmax, w.dens, it == 0, tt0, (float)temp, w.zpart, w.stat);
    tc_assign<<<gt, CS_TOK, asg_smem, stream>>>(w.sim, ctx, w.stat, w.zpart, w.tiles, w.idx, w.wsel,
                                            trace ? trace + (size_t)it * n_tot * N : nullptr);
    const dim3 gu(w.tiles_u, n_tot);
    if (J == 1) tc_update<1><<<gu, 256, upd_smem, stream>>>(feats, feat_img_stride, obj_img, rois, w.idx, w.wsel, N, C, hp, wp, S, w.part);
    else if (J == 2) tc_update<2><<<gu, 256, upd_smem, stream>>>(feats, feat_img_stride, obj_img, rois, w.idx, w.wsel, N, C, hp, wp, S, w.part);
    else if (J == 3) tc_update<3><<<gu, 256, upd_smem, stream>>>(feats, feat_img_stride, obj_img, rois, w.idx, w.wsel, N, C, hp, wp, S, w.part);
    else tc_update<4><<<gu, 256, upd_smem, stream>>>(feats, feat_img_stride, obj_img, rois, w.idx, w.wsel, N, C, hp, wp, S, w.part);
    tc_finish<<<n_tot * S, 256, 0, stream>>>(w.part, w.tiles_u, S, C, proto);
  }
  int r = affinity();                                   // returned maps: against the UNMASKED tokens (RH:849)
  if (r) return r;
  tc_transpose_out<<<dim3((N + 31) / 32, (S + 31) / 32, n_tot), dim3(32, 8), 0, stream>>>(w.sim, obj_img, img_first, S, N, LD,
                                                                                          clamp0, sim_out);
  AS_LAUNCH_CHECK();
  return 0;
}
