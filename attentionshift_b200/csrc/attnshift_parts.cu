// Part discovery after the mean shift: seed filtering, greedy prototype merging, part centres.
// Reference (RH): filter_maps RH:265-275, merge_maps RH:278-294, cal_similarity RH:297-301 (through as_cosine_maps),
// get_center_coord_with_feat RH:222-262.  All ragged results are written into fixed-size padded arrays plus counts /
// validity flags, so nothing here needs a host round trip; the host reads the flags once at the end of the batch.
#include "common.cuh"
#include <float.h>

using namespace asb;

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// keep[o,s] = mean of fg_low over {sim > 0.8} >= pos_thr      (grid n_tot*S)
__global__ void filter_score(const float* __restrict__ sim, const float* __restrict__ fg_low, int S, int N, float pos_thr,
                             int* __restrict__ keep, float* __restrict__ score_out) {
  __shared__ float red[8];
  const int row = blockIdx.x, o = row / S;
  const float* sr = sim + (size_t)row * N;
  const float* fr = fg_low + (size_t)o * N;
  float num = 0.f, den = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float f = sr[n] > 0.8f ? 1.f : 0.f;
    num += fr[n] * f;
    den += f;
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) {
    const float sc = num / fmaxf(den, 1e-6f);
    keep[row] = sc >= pos_thr;
    if (score_out) score_out[row] = sc;
  }
}

// one CTA per instance.  smem: phat [S][C] is too big in general -> cos matrix computed by warps straight from global.
__global__ void __launch_bounds__(256)
merge_protos(const float* __restrict__ proto, const int* __restrict__ keep, int S, int C, float thr,
             float* __restrict__ merged, int* __restrict__ n_merged) {
  extern __shared__ float sm[];
  float* cosm = sm;                       // [S][S]
  float* nrm = sm + S * S;                // [S]
  int* kept = reinterpret_cast<int*>(nrm + S);   // [S]
  float* wrow = reinterpret_cast<float*>(kept + S);  // [S][S] group membership weights
  __shared__ int nk_s, ng_s;
  const int o = blockIdx.x;
  const float* P = proto + (size_t)o * S * C;
  if (threadIdx.x == 0) {
    int nk = 0;
    for (int s = 0; s < S; ++s) if (keep[o * S + s]) kept[nk++] = s;
    nk_s = nk;
  }
  __syncthreads();
  const int nk = nk_s;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = lane_id();
  for (int i = warp; i < nk; i += nw) {
    const float* p = P + (size_t)kept[i] * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss += p[c] * p[c];
    ss = warp_sum(ss);
    if (lane == 0) nrm[i] = fmaxf(sqrtf(ss), 1e-8f);
  }
  __syncthreads();
  for (int e = warp; e < nk * nk; e += nw) {
    const int i = e / nk, j = e - i * nk;
    const float* a = P + (size_t)kept[i] * C;
    const float* b = P + (size_t)kept[j] * C;
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d += (a[c] / nrm[i]) * (b[c] / nrm[j]);
    d = warp_sum(d);
    if (lane == 0) cosm[i * S + j] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // live[i][j] = (j >= i && cos >= thr); walk rows in order; a used row zeroes the ROWS of its members (RH:286-292)
    int ng = 0;
    for (int i = 0; i < nk; ++i) for (int j = 0; j < nk; ++j) wrow[i * S + j] = (j >= i && cosm[i * S + j] >= thr) ? 1.f : 0.f;
    for (int i = 0; i < nk; ++i) {
      float tot = 0.f;
      for (int j = 0; j < nk; ++j) tot += wrow[i * S + j];
      if (tot > 0.f) {
        float* g = cosm + ng * S;                 // reuse: group weights, tail slot = weight sum
        float tmp[64];
        for (int j = 0; j < nk; ++j) tmp[j] = wrow[i * S + j];
        for (int j = 0; j < nk; ++j) if (tmp[j] > 0.f) for (int q = 0; q < nk; ++q) wrow[j * S + q] = 0.f;
        for (int j = 0; j < nk; ++j) g[j] = tmp[j];
        nrm[ng] = tot;
        ++ng;
      }
    }
    ng_s = ng;
    n_merged[o] = ng;
  }
  __syncthreads();
  const int ng = ng_s;
  for (int e = threadIdx.x; e < ng * C; e += blockDim.x) {
    const int g = e / C, c = e - g * C;
    float v = 0.f;
    for (int j = 0; j < nk; ++j) {
      const float w = cosm[g * S + j];
      if (w > 0.f) v += P[(size_t)kept[j] * C + c];
    }
    merged[((size_t)o * S + g) * C + c] = v / (nrm[g] + 1e-8f);
  }
  for (int e = ng * C + threadIdx.x; e < S * C; e += blockDim.x) merged[(size_t)o * S * C + e] = 0.f;
}

// per (instance, part): max, mean coordinates of the arg-max set, area(sim > 0.9).  grid (S, n_tot)
__global__ void part_stats(const float* __restrict__ pmap, const int* __restrict__ n_parts, int S, int N, int wp,
                           float* __restrict__ stat /*[n_tot][S][4] = mean_row, mean_col, area, max*/) {
  __shared__ float red[8];
  const int s = blockIdx.x, o = blockIdx.y;
  if (s >= n_parts[o]) return;
  const float* m = pmap + ((size_t)o * S + s) * N;
  float mx = -FLT_MAX;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, m[n]);
  mx = warp_max(mx);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  float sr = 0.f, sc = 0.f, cnt = 0.f, area = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float v = m[n];
    if (v >= mx) { sr += (float)(n / wp); sc += (float)(n % wp); cnt += 1.f; }
    if (v > 0.9f) area += 1.f;
  }
  sr = block_sum(sr, red); sc = block_sum(sc, red); cnt = block_sum(cnt, red); area = block_sum(area, red);
  if (threadIdx.x == 0) {
    float* st = stat + ((size_t)o * S + s) * 4;
    st[0] = sr / cnt; st[1] = sc / cnt; st[2] = area; st[3] = mx;
  }
}
// per instance: parts ordered by area (descending, stable), first KP kept; centre = (mean + 0.5) * 16; box test; feature gather
__global__ void part_centers(const float* __restrict__ stat, const int* __restrict__ n_parts, const float* __restrict__ rois,
                             const float* __restrict__ feats, long long fstride, const int* __restrict__ obj_img, int S, int C,
                             int wp, int KP, float* __restrict__ centers /*[n_tot][KP][2]*/, int* __restrict__ valid,
                             int* __restrict__ part_id, float* __restrict__ cfeat /*[n_tot][KP][C]*/) {
  __shared__ int order[64];
  __shared__ int tok_s[8];
  __shared__ int ok_s[8];
  const int o = blockIdx.x;
  const int np = n_parts[o];
  if (threadIdx.x == 0) {
    for (int i = 0; i < np; ++i) order[i] = i;
    for (int i = 1; i < np; ++i) {                 // stable insertion sort, descending area
      const int v = order[i];
      const float a = stat[((size_t)o * S + v) * 4 + 2];
      int j = i - 1;
      while (j >= 0 && stat[((size_t)o * S + order[j]) * 4 + 2] < a) { order[j + 1] = order[j]; --j; }
      order[j + 1] = v;
    }
    const float* r = rois + 4 * o;
    for (int k = 0; k < KP; ++k) {
      int ok = 0, pid = -1, tok = 0;
      float cx = -1.f, cy = -1.f;
      if (k < np) {
        pid = order[k];
        const float mr = stat[((size_t)o * S + pid) * 4], mc = stat[((size_t)o * S + pid) * 4 + 1];
        cx = (mc + 0.5f) * 16.f; cy = (mr + 0.5f) * 16.f;
        ok = (cx >= r[0]) && (cx <= r[2]) && (cy >= r[1]) && (cy <= r[3]);
        tok = (int)mr * wp + (int)mc;
      }
      centers[((size_t)o * KP + k) * 2] = cx; centers[((size_t)o * KP + k) * 2 + 1] = cy;
      valid[o * KP + k] = ok; part_id[o * KP + k] = pid;
      tok_s[k] = tok; ok_s[k] = ok;
    }
  }
  __syncthreads();
  const float* fimg = feats + obj_img[o] * fstride;
  for (int e = threadIdx.x; e < KP * C; e += blockDim.x) {
    const int k = e / C, c = e - k * C;
    cfeat[((size_t)o * KP + k) * C + c] = ok_s[k] ? fimg[(size_t)tok_s[k] * C + c] : 0.f;
  }
}

}  // namespace

extern "C" int as_filter_seeds(const float* sim, const float* fg_low, int n_tot, int S, int N, float pos_thr, int* keep,
                               float* score, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  filter_score<<<n_tot * S, 256, 0, stream>>>(sim, fg_low, S, N, pos_thr, keep, score);
  AS_LAUNCH_CHECK();
  return 0;
}
extern "C" int as_merge_prototypes(const float* proto, const int* keep, int n_tot, int S, int C, float thr, float* merged,
                                   int* n_merged, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  if (S > 64) return AS_ERR_BAD_ARG;
  const size_t smem = ((size_t)2 * S * S + 2 * S) * 4;
  merge_protos<<<n_tot, 256, smem, stream>>>(proto, keep, S, C, thr, merged, n_merged);
  AS_LAUNCH_CHECK();
  return 0;
}
// stat scratch [n_tot,S,4] floats
extern "C" int as_part_centers(const float* pmap, const int* n_parts, const float* rois, const float* feats,
                               long long feat_img_stride, const int* obj_img, int n_tot, int S, int N, int C, int wp, int KP,
                               float* centers, int* valid, int* part_id, float* cfeat, float* stat_scratch,
                               cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  if (KP > 8 || S > 64) return AS_ERR_BAD_ARG;
  part_stats<<<dim3(S, n_tot), 256, 0, stream>>>(pmap, n_parts, S, N, wp, stat_scratch);
  part_centers<<<n_tot, 256, 0, stream>>>(stat_scratch, n_parts, rois, feats, feat_img_stride, obj_img, S, C, wp, KP, centers,
                                          valid, part_id, cfeat);
  AS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ RoIAlign for the MIL layer selection
// RH:2953-2972 pools every per-layer pseudo box with mmcv's RoIAlign (configs/mae/attnshift_voc12aug.py:64-68: output 7 x 7,
// sampling_ratio 0 = adaptive, stride 16, aligned=True) before MAEBoxHeadMIL scores it.  Same arithmetic as the published
// RoIAlign (aligned: box * scale - 0.5; per bin ceil(roi / 7) x ceil(roi / 7) bilinear samples, average), on the TOKEN-MAJOR
// feature map [n_img, hp*wp, C] (no [B,C,Hp,Wp] transpose), writing [n_roi, 49, C] fp32 -- the row layout the head's LayerNorm /
// decoder_embed GEMM consume (MIL:146-150 flattens to exactly this).  One CTA per (bin, roi), threads over channels.
namespace {
__global__ void __launch_bounds__(128) roi_align_tokens_kernel(const float* __restrict__ feats, long long fstride, const float* __restrict__ rois,
                                                               int hp, int wp, int C, int pooled, float scale, float* __restrict__ out) {
  const int bin = blockIdx.x, r = blockIdx.y;
  const int ph = bin / pooled, pw = bin - ph * pooled;
  const float* roi = rois + 5 * r;
  const int img = (int)roi[0];
  const float x0 = roi[1] * scale - 0.5f, y0 = roi[2] * scale - 0.5f, x1 = roi[3] * scale - 0.5f, y1 = roi[4] * scale - 0.5f;
  const float rw = x1 - x0, rh = y1 - y0;
  const float bw = rw / (float)pooled, bh = rh / (float)pooled;
  const int gh = max((int)ceilf(rh / (float)pooled), 1) , gw = max((int)ceilf(rw / (float)pooled), 1);
  const float cnt = (float)max(gh * gw, 1);
  const float* f = feats + img * fstride;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int iy = 0; iy < gh; ++iy) {
      float y = y0 + ph * bh + (iy + 0.5f) * bh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        float x = x0 + pw * bw + (ix + 0.5f) * bw / (float)gw;
        float yy = y;
        if (yy < -1.f || yy > (float)hp || x < -1.f || x > (float)wp) continue;      // outside the map: contributes 0
        yy = fmaxf(yy, 0.f);
        x = fmaxf(x, 0.f);
        int yl = (int)yy, xl = (int)x, yh, xh;
        if (yl >= hp - 1) { yh = yl = hp - 1; yy = (float)yl; } else yh = yl + 1;
        if (xl >= wp - 1) { xh = xl = wp - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = yy - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
        acc += hy * hx * f[(size_t)(yl * wp + xl) * C + c] + hy * lx * f[(size_t)(yl * wp + xh) * C + c] +
               ly * hx * f[(size_t)(yh * wp + xl) * C + c] + ly * lx * f[(size_t)(yh * wp + xh) * C + c];
      }
    }
    out[((size_t)r * pooled * pooled + bin) * C + c] = acc / cnt;
  }
}
}  // namespace

// feats [n_img, hp*wp, C] f32 token-major (image stride feat_img_stride floats), rois [n_roi, 5] = (image, x1, y1, x2, y2) in
// pixels, spatial scale 1/stride -> out [n_roi, pooled*pooled, C] f32.
extern "C" int as_roi_align_tokens(const float* feats, long long feat_img_stride, const float* rois, int n_roi, int hp, int wp,
                                   int C, int pooled, float spatial_scale, float* out, cudaStream_t stream) {
  if (n_roi <= 0) return 0;
  if (pooled < 1 || pooled > 32 || hp < 1 || wp < 1 || C < 1) return AS_ERR_BAD_ARG;
  roi_align_tokens_kernel<<<dim3(pooled * pooled, n_roi), 128, 0, stream>>>(feats, feat_img_stride, rois, hp, wp, C, pooled, spatial_scale, out);
  AS_LAUNCH_CHECK();
  return 0;
}
