// Instance-map side of attention shift: seed sampling support, affinity refinement, full-resolution fusion,
// pseudo masks, mask-head point candidates, and the eroded / down-sampled foreground map that seeds the mean shift.
// Reference (RH = stdroi_point_deform_attn_reppoints.py):
//   norm_attns RH:329-333, sample_point_grid RH:343-371 (counts + k-th candidate; the RNG stays on the host),
//   get_point_cos_similarity_map RH:335-341, get_refined_similarity RH:668-707,
//   get_cosine_similarity_refined_map RH:1000-1019 (+ normalize_map / decouple_instance RH:1037-1046),
//   pseudo masks RH:2356-2358, get_mask_points_single_box_cos_map_fg_bg RH:433-461 (+ corrosion RH:1182-1187),
//   get_semantic_centers head RH:2011-2020 (corrosion_batch RH:145-146 + bilinear down-sampling).
// Full-resolution maps are produced from the [hp, wp] affinity maps with on-the-fly x16 bilinear interpolation
// (upsample.cuh); only the API outputs (map_cos_fg / map_cos_bg / masks) are ever written at H x W.
#include "common.cuh"
#include "upsample.cuh"
#include <float.h>

using namespace asb;

namespace {

__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  return m;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
__device__ __forceinline__ int block_sum_i(int v, int* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------ normalised CAM predicates (seed sampling)
struct NormCtx {
  const float* low;      // [n_tot, N] selected-layer CAM of every instance
  const float* mm;       // [n_tot, 2] min / max of its x16 up-sampling
  const int* item_kind;  // 0: an < thr (bg), 1: an >= thr (fg), 2: mean over the image's instances < thr (bg supp)
  const int* item_a;     // kind 0/1: instance; kind 2: first instance of the image
  const int* item_b;     // kind 2: one past the last instance
  const float* item_thr;
  int hp, wp;
};
__device__ __forceinline__ float norm_value(const NormCtx& c, int o, int y, int x) {
  const float v = up16(c.low + (size_t)o * c.hp * c.wp, c.hp, c.wp, y, x);
  const float mn = c.mm[2 * o];
  return __fdiv_rn(__fsub_rn(v, mn), __fsub_rn(c.mm[2 * o + 1], mn));      // RH:333 (no epsilon)
}
// value the item's predicate compares with its threshold (kind 2: mean over the image's instances, torch: sum / n)
__device__ __forceinline__ float norm_item_value(const NormCtx& c, int item, int y, int x) {
  if (c.item_kind[item] != 2) return norm_value(c, c.item_a[item], y, x);
  float s = norm_value(c, c.item_a[item], y, x);
  for (int o = c.item_a[item] + 1; o < c.item_b[item]; ++o) s = __fadd_rn(s, norm_value(c, o, y, x));
  return __fdiv_rn(s, (float)(c.item_b[item] - c.item_a[item]));
}
__device__ __forceinline__ bool norm_pred(const NormCtx& c, int item, int y, int x) {
  const float v = norm_item_value(c, item, y, x), thr = c.item_thr[item];
  return c.item_kind[item] == 1 ? v >= thr : v < thr;
}
// grid (H, n_items).  rowcnt [n_levels][n_items][H]: level l counts with the threshold doubled l times (the reference doubles
// a background threshold until enough candidates exist, RH:360-364: one pass answers the first n_levels rounds of that
// loop; foreground thresholds never move, their levels repeat level 0)
constexpr int NORM_MAX_LEVELS = 4;
__global__ void norm_rowcount(NormCtx c, int n_levels, int* __restrict__ rowcnt) {
  __shared__ int red[8];
  const int y = blockIdx.x, item = blockIdx.y, W = c.wp * 16;
  const bool fg = c.item_kind[item] == 1;
  const float thr = c.item_thr[item];
  int cnt[NORM_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < NORM_MAX_LEVELS; ++l) cnt[l] = 0;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const float v = norm_item_value(c, item, y, x);
    float t = thr;
#pragma unroll
    for (int l = 0; l < NORM_MAX_LEVELS; ++l) {
      cnt[l] += fg ? (v >= thr) : (v < t);
      t = __fmul_rn(t, 2.f);
    }
  }
#pragma unroll
  for (int l = 0; l < NORM_MAX_LEVELS; ++l) {
    if (l >= n_levels) break;
    const int tot = block_sum_i(cnt[l], red);
    if (threadIdx.x == 0) rowcnt[((size_t)l * gridDim.y + item) * gridDim.x + y] = tot;
  }
}
// one warp per selection: the k-th pixel (row-major) of item that satisfies the predicate
__global__ void norm_select(NormCtx c, const int* __restrict__ rowcnt, const int* __restrict__ sel_item,
                            const int* __restrict__ sel_k, int n_sel, int* __restrict__ out_xy) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_sel) return;
  const int item = sel_item[s];
  int k = sel_k[s];
  const int H = c.hp * 16, W = c.wp * 16, lane = lane_id();
  const int* rc = rowcnt + (size_t)item * H;
  int y = -1;
  for (int y0 = 0; y0 < H && y < 0; y0 += 32) {
    const int v = (y0 + lane < H) ? rc[y0 + lane] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int tot = __shfl_sync(0xffffffffu, inc, 31);
    if (k < tot) {
      const unsigned hit = __ballot_sync(0xffffffffu, k < inc);
      const int l = __ffs(hit) - 1;
      const int before = __shfl_sync(0xffffffffu, inc - v, l);
      y = y0 + l;
      k -= before;
    } else {
      k -= tot;
    }
  }
  int xo = -1;
  if (y >= 0) {
    for (int x0 = 0; x0 < W && xo < 0; x0 += 32) {
      const bool p = (x0 + lane < W) && norm_pred(c, item, y, x0 + lane);
      const unsigned b = __ballot_sync(0xffffffffu, p);
      const int tot = __popc(b);
      if (k < tot) {
        unsigned bb = b;
        for (int i = 0; i < k; ++i) bb &= bb - 1;
        xo = x0 + __ffs(bb) - 1;
      } else {
        k -= tot;
      }
    }
  }
  if (lane == 0) { out_xy[2 * s] = xo; out_xy[2 * s + 1] = (xo >= 0) ? y : -1; }
}

// proto[g][c] = mean over P points of the point's patch feature (RH:337-338);  pts [G,P,2] (x,y) pixels
__global__ void seed_proto(const float* __restrict__ feats, long long fstride, const int* __restrict__ row_img,
                           const int* __restrict__ pts, int P, int C, int hp, int wp, float* __restrict__ proto) {
  const int g = blockIdx.x;
  const float* fimg = feats + row_img[g] * fstride;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) {
      const int x = pts[((size_t)g * P + p) * 2], y = pts[((size_t)g * P + p) * 2 + 1];
      const int iy = max(0, min(y / 16, hp - 1)), ix = max(0, min(x / 16, wp - 1));
      s += fimg[(size_t)(iy * wp + ix) * C + c];
    }
    proto[(size_t)g * C + c] = s / (float)P;
  }
}

// ------------------------------------------------------------------ affinity refinement (RH:689-703)
// one CTA per row: zero entries below tau * rowmax (in place), return the weight sum
__global__ void refine_threshold(float* __restrict__ cur, int N, float tau, float* __restrict__ wsum) {
  __shared__ float red[8];
  float* row = cur + (size_t)blockIdx.x * N;
  float mx = -FLT_MAX;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, row[n]);
  mx = block_max(mx, red);
  const float thr = __fmul_rn(mx, tau);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float v = row[n];
    if (v < thr) { v = 0.f; row[n] = 0.f; }
    s += v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) wsum[blockIdx.x] = s;
}
// partial weighted sums: part[g][tile][s][c] = sum_{n in tile} w[g,s,n] * f[n,c].   grid (tiles, G), smem S*CC floats
constexpr int WS_TOK = 32;     // short tiles: the token loop of a CTA is a chain of dependent loads, parallelism comes from the grid
__global__ void __launch_bounds__(256)
weighted_sum_partial(const float* __restrict__ feats, long long fstride, const int* __restrict__ grp_img,
                     const float* __restrict__ w, int N, int C, int S, int CC, float* __restrict__ part) {
  extern __shared__ float acc_s[];
  const int g = blockIdx.y, tile = blockIdx.x, n0 = tile * WS_TOK;
  const float* fimg = feats + grp_img[g] * fstride;
  const float* wg = w + (size_t)g * S * N;
  float* dst = part + ((size_t)g * gridDim.x + tile) * S * C;
  for (int cbase = 0; cbase < C; cbase += CC) {
    const int cw = min(CC, C - cbase);
    for (int i = threadIdx.x; i < S * CC; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    for (int t = 0; t < WS_TOK && n0 + t < N; ++t) {
      const float* f = fimg + (size_t)(n0 + t) * C + cbase;
      for (int s = 0; s < S; ++s) {
        const float ws = wg[(size_t)s * N + n0 + t];       // block-uniform
        if (ws == 0.f) continue;
        for (int c = threadIdx.x; c < cw; c += blockDim.x) acc_s[s * CC + c] = fmaf(ws, __ldg(f + c), acc_s[s * CC + c]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * cw; i += blockDim.x) {
      const int s = i / cw, c = i - s * cw;
      dst[(size_t)s * C + cbase + c] = acc_s[s * CC + c];
    }
    __syncthreads();
  }
}
// centroid[row][c] = (ordered sum of partials) / clamp(wsum[row], 1e-8)
__global__ void weighted_sum_finish(const float* __restrict__ part, const float* __restrict__ wsum, int tiles, int S, int C,
                                    float* __restrict__ out) {
  const int row = blockIdx.x, g = row / S, s = row - g * S;
  const float den = fmaxf(wsum[row], 1e-8f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int t = 0; t < tiles; ++t) v += part[(((size_t)g * tiles + t) * S + s) * C + c];
    out[(size_t)row * C + c] = v / den;
  }
}
// per image group g (rows: [0,n) fg instances, then n_extra rows that also compete -- the bg supplement(s) -- then, in the
// first-round layout, n bg instances):
//   fg rows *= box mask (in place, RH:699 / RH:740); optionally emit fg (winner-take-all over rows 0..n+n_extra-1,
//   RH:700-703 / RH:741-743) and bg (rows n+1..2n; first round only, bg_out may be null).
__global__ void refine_select(float* __restrict__ cur, int S, int N, int wp, const int* __restrict__ grp_first,
                              const int* __restrict__ grp_nobj, const float* __restrict__ rois, int emit, int n_extra,
                              float* __restrict__ fg_out, float* __restrict__ bg_out) {
  const int g = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int nobj = grp_nobj[g], o0 = grp_first[g];
  float* cg = cur + (size_t)g * S * N;
  const int r = n / wp, cidx = n - r * wp;
  float best = -FLT_MAX;
  int bi = 0;
  for (int j = 0; j < nobj + n_extra; ++j) {
    float v = cg[(size_t)j * N + n];
    if (j < nobj) {
      const float* roi = rois + 4 * (o0 + j);
      const int c0 = (int)floorf(roi[0] / 16.f), r0 = (int)floorf(roi[1] / 16.f);
      const int c1 = (int)(floorf(roi[2] / 16.f) + 1.f), r1 = (int)(floorf(roi[3] / 16.f) + 1.f);
      const bool in = r >= r0 && r < r1 && cidx >= c0 && cidx < c1;
      v = in ? v : __fmul_rn(v, 0.f);
      cg[(size_t)j * N + n] = v;
    }
    if (v > best) { best = v; bi = j; }
  }
  if (emit) {
    for (int j = 0; j < nobj; ++j) {
      fg_out[(size_t)(o0 + j) * N + n] = (bi == j) ? cg[(size_t)j * N + n] : 0.f;
      if (bg_out) bg_out[(size_t)(o0 + j) * N + n] = cg[(size_t)(nobj + 1 + j) * N + n];
    }
  }
}

// ------------------------------------------------------------------ full-resolution fusion (RH:1010-1019)
struct FuseCtx {
  const float* fg_low;   // [n_tot, N]
  const float* bg_low;
  int hp, wp;
};
__device__ __forceinline__ void fuse_vals(const FuseCtx& c, int o, int y, int x, float& fused, float& bg) {
  const Tap ty = tap_up16(y, c.hp), tx = tap_up16(x, c.wp);
  const size_t b0 = (size_t)o * c.hp * c.wp + ty.i0 * c.wp, b1 = (size_t)o * c.hp * c.wp + ty.i1 * c.wp;
  const float fg = lerp2(c.fg_low[b0 + tx.i0], c.fg_low[b0 + tx.i1], c.fg_low[b1 + tx.i0], c.fg_low[b1 + tx.i1], ty, tx);
  bg = lerp2(c.bg_low[b0 + tx.i0], c.bg_low[b0 + tx.i1], c.bg_low[b1 + tx.i0], c.bg_low[b1 + tx.i1], ty, tx);
  fused = __fmul_rn(__fsub_rn(1.f, bg), fg);
}
// stats[o] = {max fused, max bg, max bg2} as order-preserving uints
__global__ void __launch_bounds__(256) fuse_max1(FuseCtx c, unsigned* __restrict__ stats) {
  __shared__ float red[8];
  const int o = blockIdx.y, H = c.hp * 16, W = c.wp * 16;
  float mf = -FLT_MAX, mb = -FLT_MAX;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    float f, b;
    fuse_vals(c, o, p / W, p % W, f, b);
    mf = fmaxf(mf, f); mb = fmaxf(mb, b);
  }
  mf = block_max(mf, red);
  mb = block_max(mb, red);
  if (threadIdx.x == 0) { atomicMax(stats + 3 * o, enc_f(mf)); atomicMax(stats + 3 * o + 1, enc_f(mb)); }
}
__device__ __forceinline__ float bg2_val(float fused, float bg, float mf, float mb) {
  const float bgn = __fdiv_rn(bg, __fadd_rn(mb, 1e-8f));
  const float fgn = __fdiv_rn(fused, __fadd_rn(mf, 1e-8f));
  const float fake = __fsub_rn(1.f, __fadd_rn(__fmul_rn(fgn, 0.5f), __fmul_rn(bgn, 0.5f)));
  return __fadd_rn(bgn, fake);
}
__global__ void __launch_bounds__(256) fuse_max2(FuseCtx c, unsigned* __restrict__ stats) {
  __shared__ float red[8];
  const int o = blockIdx.y, H = c.hp * 16, W = c.wp * 16;
  const float mf = dec_f(stats[3 * o]), mb = dec_f(stats[3 * o + 1]);
  float m2 = -FLT_MAX;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    float f, b;
    fuse_vals(c, o, p / W, p % W, f, b);
    m2 = fmaxf(m2, bg2_val(f, b, mf, mb));
  }
  m2 = block_max(m2, red);
  if (threadIdx.x == 0) atomicMax(stats + 3 * o + 2, enc_f(m2));
}
__global__ void __launch_bounds__(256)
fuse_write(FuseCtx c, const unsigned* __restrict__ stats, float mask_thr, float* __restrict__ map_fg,
           float* __restrict__ map_bg, unsigned char* __restrict__ mask) {
  const int o = blockIdx.y, H = c.hp * 16, W = c.wp * 16;
  const float mf = dec_f(stats[3 * o]), mb = dec_f(stats[3 * o + 1]), m2 = dec_f(stats[3 * o + 2]);
  const float df = fmaxf(mf, 1e-8f), d2 = fmaxf(m2, 1e-8f);
  const float top = __fmul_rn(__fdiv_rn(mf, df), mask_thr);      // rowmax(map_fg) * pos_mask_thr (RH:2356)
  const size_t base = (size_t)o * H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    float f, b;
    fuse_vals(c, o, p / W, p % W, f, b);
    const float vf = __fdiv_rn(f, df);
    map_fg[base + p] = vf;
    if (map_bg) map_bg[base + p] = __fdiv_rn(bg2_val(f, b, mf, mb), d2);
    if (mask) mask[base + p] = vf > top;
  }
}

// ------------------------------------------------------------------ mask-head point candidates (RH:433-461)
struct Crop { int x0, y0, x1, y1; };
__device__ __forceinline__ Crop crop_of(const float* roi, int H, int W) {   // rois.int() then python slicing
  Crop c;
  c.x0 = min(max((int)roi[0], 0), W); c.y0 = min(max((int)roi[1], 0), H);
  c.x1 = min(max((int)roi[2], 0), W); c.y1 = min(max((int)roi[3], 0), H);
  return c;
}
// cstat[o] = {max fg in crop, max bg in crop}
__global__ void __launch_bounds__(256)
crop_max(const float* __restrict__ map_fg, const float* __restrict__ map_bg, const float* __restrict__ rois, int H, int W,
         unsigned* __restrict__ cstat) {
  __shared__ float red[8];
  const int o = blockIdx.y;
  const Crop c = crop_of(rois + 4 * o, H, W);
  const int cw = max(c.x1 - c.x0, 0), ch = max(c.y1 - c.y0, 0);
  float mf = -FLT_MAX, mb = -FLT_MAX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cw * ch; i += gridDim.x * blockDim.x) {
    const size_t p = (size_t)o * H * W + (size_t)(c.y0 + i / cw) * W + c.x0 + i % cw;
    mf = fmaxf(mf, map_fg[p]); mb = fmaxf(mb, map_bg[p]);
  }
  mf = block_max(mf, red);
  mb = block_max(mb, red);
  if (threadIdx.x == 0) { atomicMax(cstat + 2 * o, enc_f(mf)); atomicMax(cstat + 2 * o + 1, enc_f(mb)); }
}
// horizontal pass of the k x k erosion of (fg > max*pos_thr), restricted to the crop (outside = ignored)
__global__ void __launch_bounds__(256)
crop_erode_rows(const float* __restrict__ map_fg, const float* __restrict__ rois, const unsigned* __restrict__ cstat,
                float pos_thr, int k, int H, int W, unsigned char* __restrict__ tmp) {
  const int o = blockIdx.y;
  const Crop c = crop_of(rois + 4 * o, H, W);
  const int cw = max(c.x1 - c.x0, 0), ch = max(c.y1 - c.y0, 0), r = k / 2;
  const float thr = __fmul_rn(dec_f(cstat[2 * o]), pos_thr);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cw * ch; i += gridDim.x * blockDim.x) {
    const int y = c.y0 + i / cw, x = c.x0 + i % cw;
    const float* row = map_fg + (size_t)o * H * W + (size_t)y * W;
    bool all = true;
    for (int xx = max(x - r, c.x0); xx <= min(x + r, c.x1 - 1) && all; ++xx) all = row[xx] > thr;
    tmp[(size_t)o * H * W + (size_t)y * W + x] = all;
  }
}
// vertical pass + per-row candidate counts.  grid (H, n_tot): rowcnt[o][y] = {#pos, #neg} of crop row y
__global__ void __launch_bounds__(256)
crop_erode_cols_count(const unsigned char* __restrict__ tmp, const float* __restrict__ map_bg, const float* __restrict__ rois,
                      const unsigned* __restrict__ cstat, float neg_thr, int k, int H, int W,
                      unsigned char* __restrict__ pos, int* __restrict__ rowcnt) {
  __shared__ int red[8];
  const int o = blockIdx.y, y = blockIdx.x;
  const Crop c = crop_of(rois + 4 * o, H, W);
  int np = 0, nn = 0;
  if (y >= c.y0 && y < c.y1) {
    const int r = k / 2;
    const float thr = __fmul_rn(dec_f(cstat[2 * o + 1]), neg_thr);
    for (int x = c.x0 + threadIdx.x; x < c.x1; x += blockDim.x) {
      bool all = true;
      for (int yy = max(y - r, c.y0); yy <= min(y + r, c.y1 - 1) && all; ++yy) all = tmp[(size_t)o * H * W + (size_t)yy * W + x];
      pos[(size_t)o * H * W + (size_t)y * W + x] = all;
      np += all;
      nn += map_bg[(size_t)o * H * W + (size_t)y * W + x] > thr;
    }
  }
  np = block_sum_i(np, red);
  nn = block_sum_i(nn, red);
  if (threadIdx.x == 0) { rowcnt[((size_t)o * H + y) * 2] = np; rowcnt[((size_t)o * H + y) * 2 + 1] = nn; }
}
// one warp per selection (o, kind 0 pos / 1 neg, k): k-th candidate of the crop in row-major order -> (x, y) image coords
__global__ void crop_select(const unsigned char* __restrict__ pos, const float* __restrict__ map_bg,
                            const float* __restrict__ rois, const unsigned* __restrict__ cstat, float neg_thr,
                            const int* __restrict__ rowcnt, const int* __restrict__ sel_obj, const int* __restrict__ sel_kind,
                            const int* __restrict__ sel_k, int n_sel, int H, int W, int* __restrict__ out_xy) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_sel) return;
  const int o = sel_obj[s], kind = sel_kind[s], lane = lane_id();
  int k = sel_k[s];
  const Crop c = crop_of(rois + 4 * o, H, W);
  const float thr = __fmul_rn(dec_f(cstat[2 * o + 1]), neg_thr);
  int y = -1;
  for (int y0 = c.y0; y0 < c.y1 && y < 0; y0 += 32) {
    const int v = (y0 + lane < c.y1) ? rowcnt[((size_t)o * H + y0 + lane) * 2 + kind] : 0;
    int inc = v;
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const int tot = __shfl_sync(0xffffffffu, inc, 31);
    if (k < tot) {
      const int l = __ffs(__ballot_sync(0xffffffffu, k < inc)) - 1;
      k -= __shfl_sync(0xffffffffu, inc - v, l);
      y = y0 + l;
    } else {
      k -= tot;
    }
  }
  int xo = -1;
  if (y >= 0) {
    for (int x0 = c.x0; x0 < c.x1 && xo < 0; x0 += 32) {
      const int x = x0 + lane;
      bool p = false;
      if (x < c.x1) p = kind == 0 ? (bool)pos[(size_t)o * H * W + (size_t)y * W + x] : map_bg[(size_t)o * H * W + (size_t)y * W + x] > thr;
      const unsigned b = __ballot_sync(0xffffffffu, p);
      const int tot = __popc(b);
      if (k < tot) {
        unsigned bb = b;
        for (int i = 0; i < k; ++i) bb &= bb - 1;
        xo = x0 + __ffs(bb) - 1;
      } else {
        k -= tot;
      }
    }
  }
  if (lane == 0) { out_xy[2 * s] = xo; out_xy[2 * s + 1] = xo >= 0 ? y : -1; }
}

// ------------------------------------------------------------------ eroded + down-sampled fg map (RH:2011-2020)
// fg_low[o][patch] = bilinear-down( erode_k( map_fg > thr ) ): scale 16 -> taps at pixels 16p+7, 16p+8, weights .5/.5
__global__ void erode_down(const float* __restrict__ map_fg, int H, int W, float thr, int k, float* __restrict__ fg_low,
                           float* __restrict__ seed_map) {
  const int o = blockIdx.y, hp = H / 16, wp = W / 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hp * wp) return;
  const int py = i / wp, px = i - py * wp, r = k / 2;
  const float* m = map_fg + (size_t)o * H * W;
  float e[2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int y = 16 * py + 7 + a, x = 16 * px + 7 + b;
      bool all = true;
      for (int yy = max(y - r, 0); yy <= min(y + r, H - 1) && all; ++yy)
        for (int xx = max(x - r, 0); xx <= min(x + r, W - 1) && all; ++xx) all = m[(size_t)yy * W + xx] > thr;
      e[a][b] = all ? 1.f : 0.f;
    }
  const float t0 = __fmaf_rn(e[0][0], 0.5f, __fmul_rn(e[0][1], 0.5f));
  const float t1 = __fmaf_rn(e[1][0], 0.5f, __fmul_rn(e[1][1], 0.5f));
  const float v = __fmaf_rn(t0, 0.5f, __fmul_rn(t1, 0.5f));
  fg_low[(size_t)o * hp * wp + i] = v;
  seed_map[(size_t)o * hp * wp + i] = v > thr ? 1.f : 0.f;
}

}  // namespace

// ---- seed sampling support -------------------------------------------------------------------------------------
extern "C" int as_norm_rowcount(const float* low, const float* minmax, const int* item_kind, const int* item_a,
                                const int* item_b, const float* item_thr, int n_items, int hp, int wp, int n_levels,
                                int* rowcnt, cudaStream_t stream) {
  if (n_items <= 0) return 0;
  if (n_levels < 1 || n_levels > NORM_MAX_LEVELS) return AS_ERR_BAD_ARG;
  NormCtx c{low, minmax, item_kind, item_a, item_b, item_thr, hp, wp};
  norm_rowcount<<<dim3(hp * 16, n_items), 256, 0, stream>>>(c, n_levels, rowcnt);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_norm_select(const float* low, const float* minmax, const int* item_kind, const int* item_a,
                              const int* item_b, const float* item_thr, int hp, int wp, const int* rowcnt,
                              const int* sel_item, const int* sel_k, int n_sel, int* out_xy, cudaStream_t stream) {
  if (n_sel <= 0) return 0;
  NormCtx c{low, minmax, item_kind, item_a, item_b, item_thr, hp, wp};
  norm_select<<<(n_sel + 3) / 4, 128, 0, stream>>>(c, rowcnt, sel_item, sel_k, n_sel, out_xy);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_seed_proto(const float* feats, long long feat_img_stride, const int* row_img, const int* pts, int G,
                             int P, int C, int hp, int wp, float* proto, cudaStream_t stream) {
  if (G <= 0) return 0;
  seed_proto<<<G, 256, 0, stream>>>(feats, feat_img_stride, row_img, pts, P, C, hp, wp, proto);
  AS_LAUNCH_CHECK();
  return 0;
}

// ---- refinement pieces -----------------------------------------------------------------------------------------
extern "C" int as_refine_threshold(float* cur, int rows, int N, float tau, float* wsum, cudaStream_t stream) {
  if (rows <= 0) return 0;
  refine_threshold<<<rows, 256, 0, stream>>>(cur, N, tau, wsum);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" size_t as_weighted_centroid_workspace(int G, int S, int N, int C) {
  return (size_t)G * ((N + WS_TOK - 1) / WS_TOK) * S * C * 4;
}
// out[g,s,:] = sum_n w[g,s,n] f[img(g),n,:] / clamp(wsum[g,s], 1e-8)
extern "C" int as_weighted_centroid(const float* feats, long long feat_img_stride, const int* grp_img, const float* w,
                                    const float* wsum, int G, int S, int N, int C, float* out, void* workspace,
                                    size_t workspace_bytes, cudaStream_t stream) {
  if (G <= 0) return 0;
  if (workspace_bytes < as_weighted_centroid_workspace(G, S, N, C)) return AS_ERR_BAD_ARG;
  int CC = C;
  while ((size_t)S * CC * 4 > 160 * 1024) CC = (CC + 1) / 2;
  const size_t smem = (size_t)S * CC * 4;
  AS_CUDA(cudaFuncSetAttribute(weighted_sum_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (N + WS_TOK - 1) / WS_TOK;
  weighted_sum_partial<<<dim3(tiles, G), 256, smem, stream>>>(feats, feat_img_stride, grp_img, w, N, C, S, CC, (float*)workspace);
  weighted_sum_finish<<<G * S, 256, 0, stream>>>((const float*)workspace, wsum, tiles, S, C, out);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_refine_select(float* cur, int G, int S, int N, int wp, const int* grp_first, const int* grp_nobj,
                                const float* rois, int emit, int n_extra, float* fg_out, float* bg_out, cudaStream_t stream) {
  if (G <= 0) return 0;
  if (n_extra < 0 || (emit && !fg_out)) return AS_ERR_BAD_ARG;
  refine_select<<<dim3((N + 255) / 256, G), 256, 0, stream>>>(cur, S, N, wp, grp_first, grp_nobj, rois, emit, n_extra, fg_out,
                                                              bg_out);
  AS_LAUNCH_CHECK();
  return 0;
}

// ---- full resolution -------------------------------------------------------------------------------------------
// map_fg / map_bg [n_tot,H,W] fp32 (map_bg optional), mask [n_tot,H,W] uint8 (optional); stats scratch n_tot*3 uint32
extern "C" int as_fuse_instance_maps(const float* fg_low, const float* bg_low, int n_tot, int hp, int wp, float mask_thr,
                                     float* map_fg, float* map_bg, unsigned char* mask, void* stats_scratch,
                                     cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  FuseCtx c{fg_low, bg_low, hp, wp};
  unsigned* st = (unsigned*)stats_scratch;
  AS_CUDA(cudaMemsetAsync(st, 0, (size_t)n_tot * 3 * 4, stream));
  const dim3 grid(128, n_tot);
  fuse_max1<<<grid, 256, 0, stream>>>(c, st);
  fuse_max2<<<grid, 256, 0, stream>>>(c, st);
  fuse_write<<<grid, 256, 0, stream>>>(c, st, mask_thr, map_fg, map_bg, mask);
  AS_LAUNCH_CHECK();
  return 0;
}

// pos [n_tot,H,W] uint8 eroded-foreground candidates, rowcnt [n_tot,H,2]; scratch: tmp [n_tot,H,W] uint8 + cstat n_tot*2 uint32
extern "C" size_t as_mask_candidates_workspace(int n_tot, int H, int W) { return (size_t)n_tot * H * W + (size_t)n_tot * 8 + 256; }
extern "C" int as_mask_candidates(const float* map_fg, const float* map_bg, const float* rois, int n_tot, int H, int W,
                                  float pos_thr, float neg_thr, int corr_size, unsigned char* pos, int* rowcnt,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  if (workspace_bytes < as_mask_candidates_workspace(n_tot, H, W)) return AS_ERR_BAD_ARG;
  unsigned* cstat = (unsigned*)workspace;
  unsigned char* tmp = (unsigned char*)workspace + (((size_t)n_tot * 8 + 255) / 256) * 256;
  AS_CUDA(cudaMemsetAsync(cstat, 0, (size_t)n_tot * 8, stream));
  crop_max<<<dim3(64, n_tot), 256, 0, stream>>>(map_fg, map_bg, rois, H, W, cstat);
  crop_erode_rows<<<dim3(128, n_tot), 256, 0, stream>>>(map_fg, rois, cstat, pos_thr, corr_size, H, W, tmp);
  crop_erode_cols_count<<<dim3(H, n_tot), 256, 0, stream>>>(tmp, map_bg, rois, cstat, neg_thr, corr_size, H, W, pos, rowcnt);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_mask_select(const unsigned char* pos, const float* map_bg, const float* rois, const void* workspace,
                              float neg_thr, const int* rowcnt, const int* sel_obj, const int* sel_kind, const int* sel_k,
                              int n_sel, int H, int W, int* out_xy, cudaStream_t stream) {
  if (n_sel <= 0) return 0;
  crop_select<<<(n_sel + 3) / 4, 128, 0, stream>>>(pos, map_bg, rois, (const unsigned*)workspace, neg_thr, rowcnt, sel_obj,
                                                  sel_kind, sel_k, n_sel, H, W, out_xy);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_erode_downsample(const float* map_fg, int n_tot, int H, int W, float thr, int corr_size, float* fg_low,
                                   float* seed_map, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  const int N = (H / 16) * (W / 16);
  erode_down<<<dim3((N + 127) / 128, n_tot), 128, 0, stream>>>(map_fg, H, W, thr, corr_size, fg_low, seed_map);
  AS_LAUNCH_CHECK();
  return 0;
}
