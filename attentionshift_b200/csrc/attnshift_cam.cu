// CAM slicing, x16 bilinear up-sampling evaluated on the fly, and CAM -> pseudo box (connected components).
// Reference: RH:2272-2275 (slice point-token rows of the roll-out, bilinear x16, align_corners=False) and
// get_bbox_from_cam_fast RH:60-116 (min-max normalise, binarise at seed_thr, cc_torch labelling, keep components with
// area >= seed_multiple * largest, joint extent, mirror-expand around the GT point).
// The up-sampled [7, n_gt, H, W] CAM stack (29 MB per instance at 1024^2) is never written: every consumer
// re-evaluates the 4-tap interpolation from the [hp, wp] map (16 KB) with the exact arithmetic of
// torch's CPU kernel:  t = fma(v0, wx0, v1*wx1);  out = fma(t0, wy0, t1*wy1).
// Connected components: 8-connectivity union-find (label = smallest pixel index of the component); only the
// partition matters to the caller.  cc_torch itself is absent from the reference tree (parity unpinned, see oracle).
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

// cams[l, o, n] = rows[img(o), l, point(o), 1 + n]      rows: [B, L, n_rows, T]
__global__ void cam_gather(const float* __restrict__ rows, const int* __restrict__ obj_img, const int* __restrict__ obj_pt,
                           int L, int n_rows, int T /* row stride */, int N, int n_tot, float* __restrict__ cams) {
  const int o = blockIdx.y, l = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  cams[((size_t)l * n_tot + o) * N + n] = rows[(((size_t)obj_img[o] * L + l) * n_rows + obj_pt[o]) * T + 1 + n];
}

__global__ void fill_u32(unsigned* p, unsigned v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// min / max of the up-sampled map.  grid (chunks, n_maps); mm[m] = {enc(min) as ~, enc(max)}
__global__ void __launch_bounds__(256)
cam_minmax(const float* __restrict__ lows, int hp, int wp, unsigned* __restrict__ mm) {
  extern __shared__ float low_s[];
  const int m = blockIdx.y;
  const int N = hp * wp, H = hp * 16, W = wp * 16;
  for (int i = threadIdx.x; i < N; i += blockDim.x) low_s[i] = lows[(size_t)m * N + i];
  __syncthreads();
  float mn = FLT_MAX, mx = -FLT_MAX;
  const int total = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const float v = up16(low_s, hp, wp, p / W, p % W);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __shared__ float r0[8], r1[8];
  if (lane_id() == 0) { r0[threadIdx.x >> 5] = mn; r1[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, r0[i]); mx = fmaxf(mx, r1[i]); }
    atomicMin(mm + 2 * m, enc_f(mn));
    atomicMax(mm + 2 * m + 1, enc_f(mx));
  }
}

__global__ void minmax_decode(const unsigned* __restrict__ mm, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dec_f(mm[i]);
}

// ---------------------------------------------------------------- union-find CCL
__device__ __forceinline__ int uf_find(volatile int* L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// ---- run-based labelling.  A row's maximal foreground runs are the units: every pixel points at its run head, only
// run heads carry areas / extents, and unions are issued only where a new adjacency can appear (8-connectivity):
//   up fg            -> union(p, up)       unless left is fg (left already met `up` as its up-right neighbour)
//   up bg, upleft fg -> union(p, upleft)   unless left is fg (left already met it as its `up`)
//   up bg, upright fg-> union(p, upright)  always
// label[p] = head index if foreground else -1;  runlen[head] = run length.
// fg = (v - mn) / clamp(mx - mn, 1e-6) >= thr      (RH:63-66).   grid (H, n_maps), 256 threads, 4 pixels / thread
__global__ void __launch_bounds__(256)
ccl_init_runs(const float* __restrict__ lows, const float* __restrict__ mmf, int hp, int wp, float thr,
              int* __restrict__ labels, int* __restrict__ runlen) {
  __shared__ unsigned bits[128 + 1];              // foreground bitmap of the row (W <= 4096)
  const int m = blockIdx.y, y = blockIdx.x;
  const int H = hp * 16, W = wp * 16;
  const int nw = (W + 31) / 32;
  const float* low = lows + (size_t)m * hp * wp;
  const float mn = mmf[2 * m], den = fmaxf(mmf[2 * m + 1] - mn, 1e-6f);
  int* lab = labels + (size_t)m * H * W + (size_t)y * W;
  int* rl = runlen + (size_t)m * H * W + (size_t)y * W;
  const int lane = lane_id();
  for (int x0 = (threadIdx.x >> 5) * 32; x0 < nw * 32; x0 += blockDim.x) {
    const int x = x0 + lane;
    const bool fg = x < W && ((up16(low, hp, wp, y, x) - mn) / den >= thr);
    const unsigned b = __ballot_sync(0xffffffffu, fg);
    if (lane == 0) bits[x0 >> 5] = b;
  }
  if (threadIdx.x == 0) bits[nw] = 0u;             // sentinel: background past the row end
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int w = x >> 5;
    const int bpos = x & 31;
    if (!((bits[w] >> bpos) & 1u)) { lab[x] = -1; continue; }
    // nearest background bit to the left -> run start
    unsigned inv = ~bits[w] & ((1u << bpos) - 1u);
    int start;
    if (inv) {
      start = (w << 5) + 32 - __clz(inv);
    } else {
      int ww = w - 1;
      while (ww >= 0 && bits[ww] == 0xffffffffu) --ww;
      start = ww < 0 ? 0 : (ww << 5) + 32 - __clz(~bits[ww]);
    }
    lab[x] = (int)((size_t)y * W + start);
    if (start == x) {                               // run head: nearest background bit to the right -> run end
      inv = ~bits[w] & ~((bpos == 31) ? 0xffffffffu : ((2u << bpos) - 1u));
      int end;
      if (inv) {
        end = (w << 5) + __ffs(inv) - 1;
      } else {
        int ww = w + 1;
        while (bits[ww] == 0xffffffffu) ++ww;        // bits[nw] == 0 terminates
        end = (ww << 5) + __ffs(~bits[ww]) - 1;
      }
      rl[x] = min(end, W) - x;
    }
  }
}

__global__ void ccl_merge_runs(int* __restrict__ labels, int H, int W) {
  int* lab = labels + (size_t)blockIdx.y * H * W;
  const int total = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    if (p < W || lab[p] < 0) continue;
    const int x = p % W;
    const bool left = x > 0 && lab[p - 1] >= 0;
    const bool up = lab[p - W] >= 0;
    if (up) {
      if (!left) uf_union(lab, p, p - W);
    } else {
      if (!left && x > 0 && lab[p - W - 1] >= 0) uf_union(lab, p, p - W - 1);
      if (x + 1 < W && lab[p - W + 1] >= 0) uf_union(lab, p, p - W + 1);
    }
  }
}

// per run head: root, area accumulation (one atomic per run)
__global__ void ccl_run_area(int* __restrict__ labels, const int* __restrict__ runlen, int* __restrict__ area, int H, int W) {
  int* lab = labels + (size_t)blockIdx.y * H * W;
  const int* rl = runlen + (size_t)blockIdx.y * H * W;
  int* ar = area + (size_t)blockIdx.y * H * W;
  const int total = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int len = rl[p];
    if (len <= 0) continue;                 // not a run head
    const int r = uf_find(lab, p);
    atomicAdd(&ar[r], len);
  }
}
__global__ void ccl_max_area(const int* __restrict__ area, int H, int W, int* __restrict__ max_area) {
  const int* ar = area + (size_t)blockIdx.y * H * W;
  const int total = H * W;
  int best = 0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) best = max(best, ar[p]);
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane_id() == 0 && best > 0) atomicMax(max_area + blockIdx.y, best);
}
// extent of the kept components from the run heads; ext[m] = {xmin, ymin, xmax, ymax}
__global__ void ccl_run_extent(int* __restrict__ labels, const int* __restrict__ runlen, const int* __restrict__ area,
                               const int* __restrict__ max_area, float ratio, int H, int W, int* __restrict__ ext) {
  const int m = blockIdx.y;
  int* lab = labels + (size_t)m * H * W;
  const int* rl = runlen + (size_t)m * H * W;
  const int* ar = area + (size_t)m * H * W;
  const float need = ratio * (float)max_area[m];
  const int total = H * W;
  int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int len = rl[p];
    if (len <= 0) continue;
    const int r = uf_find(lab, p);
    if ((float)ar[r] >= need) {
      const int y = p / W, x = p - y * W;
      x0 = min(x0, x); x1 = max(x1, x + len - 1); y0 = min(y0, y); y1 = max(y1, y);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane_id() == 0 && x1 >= 0) {
    atomicMin(ext + 4 * m, x0); atomicMin(ext + 4 * m + 1, y0);
    atomicMax(ext + 4 * m + 2, x1); atomicMax(ext + 4 * m + 3, y1);
  }
}
// optional (tests / visualisation): the reference's remained_label_masks
__global__ void ccl_keep_mask(int* __restrict__ labels, const int* __restrict__ area, const int* __restrict__ max_area,
                              float ratio, int H, int W, unsigned char* __restrict__ keep_mask) {
  const int m = blockIdx.y;
  int* lab = labels + (size_t)m * H * W;
  const int* ar = area + (size_t)m * H * W;
  const float need = ratio * (float)max_area[m];
  const int total = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    bool keep = false;
    if (lab[p] >= 0) keep = (float)ar[uf_find(lab, p)] >= need;
    keep_mask[(size_t)m * total + p] = keep;
  }
}
__global__ void ext_init(int* ext, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ext[i] = (i & 3) < 2 ? INT_MAX : -1;
}
// RH:97-115 'expand': mirror the far side of the kept extent around the GT point, clip to the image
__global__ void cam_expand_box(const int* __restrict__ ext, const float* __restrict__ points /*[n_tot,2] (x,y)*/,
                               int n_tot, int n_maps, float img_w, float img_h, float* __restrict__ boxes) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_maps) return;
  const int o = m % n_tot;
  float* b = boxes + 4 * m;
  if (ext[4 * m + 2] < 0) { b[0] = 0.f; b[1] = 0.f; b[2] = 1.f; b[3] = 1.f; return; }
  const float px0 = (float)ext[4 * m], py0 = (float)ext[4 * m + 1], px1 = (float)ext[4 * m + 2], py1 = (float)ext[4 * m + 3];
  const float xc = points[2 * o], yc = points[2 * o + 1];
  float x0, x1, y0, y1;
  if (fabsf(xc - px0) > fabsf(xc - px1)) { x0 = px0; x1 = xc * 2.f - x0; x1 = x1 < img_w ? x1 : img_w; }
  else { x1 = px1; x0 = xc * 2.f - x1; x0 = x0 > 0.f ? x0 : 0.f; }
  if (fabsf(yc - py0) > fabsf(yc - py1)) { y0 = py0; y1 = yc * 2.f - y0; y1 = y1 < img_h ? y1 : img_h; }
  else { y1 = py1; y0 = yc * 2.f - y1; y0 = y0 > 0.f ? y0 : 0.f; }
  b[0] = x0; b[1] = y0; b[2] = x1; b[3] = y1;
}

}  // namespace

extern "C" int as_cam_gather(const float* rows, const int* obj_img, const int* obj_pt, int L, int n_rows, int T, int N,
                             int n_tot, float* cams, cudaStream_t stream) {
  if (n_tot <= 0) return 0;
  cam_gather<<<dim3((N + 255) / 256, n_tot, L), 256, 0, stream>>>(rows, obj_img, obj_pt, L, n_rows, T, N, n_tot, cams);
  AS_LAUNCH_CHECK();
  return 0;
}

// lows [n_maps, hp*wp] -> minmax [n_maps, 2] of the x16 bilinear up-sampling.  scratch: n_maps*2 uint32
extern "C" int as_cam_minmax(const float* lows, int n_maps, int hp, int wp, float* minmax, void* scratch,
                             cudaStream_t stream) {
  if (n_maps <= 0) return 0;
  unsigned* mm = (unsigned*)scratch;
  fill_u32<<<(2 * n_maps + 255) / 256, 256, 0, stream>>>(mm, 0u, 2 * (size_t)n_maps);
  // min slots must start at UINT_MAX
  cudaMemset2DAsync(mm, 8, 0xff, 4, n_maps, stream);
  const size_t smem = (size_t)hp * wp * 4;
  AS_CUDA(cudaFuncSetAttribute(cam_minmax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cam_minmax<<<dim3(32, n_maps), 256, smem, stream>>>(lows, hp, wp, mm);
  minmax_decode<<<(2 * n_maps + 255) / 256, 256, 0, stream>>>(mm, minmax, 2 * n_maps);
  AS_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t as_cam_bbox_workspace(int n_maps, int H, int W) {
  return (size_t)n_maps * H * W * 12 + (size_t)n_maps * 5 * 4 + 1024;
}

// boxes [n_maps,4] for maps ordered [layer][instance] (n_maps = L * n_tot); points [n_tot,2].
// keep_mask (optional) [n_maps,H,W] uint8 = the reference's remained_label_masks.
extern "C" int as_cam_bbox(const float* lows, const float* minmax, const float* points, int n_maps, int n_tot, int hp,
                           int wp, float cam_thr, float area_ratio, float* boxes, unsigned char* keep_mask,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_maps <= 0) return 0;
  const int H = hp * 16, W = wp * 16;
  if (W > 4096 || workspace_bytes < as_cam_bbox_workspace(n_maps, H, W)) return AS_ERR_BAD_ARG;
  int* labels = (int*)workspace;
  int* area = labels + (size_t)n_maps * H * W;
  int* runlen = area + (size_t)n_maps * H * W;
  int* max_area = runlen + (size_t)n_maps * H * W;
  int* ext = max_area + n_maps;
  AS_CUDA(cudaMemsetAsync(area, 0, ((size_t)2 * n_maps * H * W + n_maps) * 4, stream));   // area, runlen, max_area
  ext_init<<<(4 * n_maps + 255) / 256, 256, 0, stream>>>(ext, 4 * n_maps);
  const dim3 grid(74, n_maps);
  ccl_init_runs<<<dim3(H, n_maps), 256, 0, stream>>>(lows, minmax, hp, wp, cam_thr, labels, runlen);
  ccl_merge_runs<<<grid, 256, 0, stream>>>(labels, H, W);
  ccl_run_area<<<grid, 256, 0, stream>>>(labels, runlen, area, H, W);
  ccl_max_area<<<grid, 256, 0, stream>>>(area, H, W, max_area);
  ccl_run_extent<<<grid, 256, 0, stream>>>(labels, runlen, area, max_area, area_ratio, H, W, ext);
  if (keep_mask) ccl_keep_mask<<<grid, 256, 0, stream>>>(labels, area, max_area, area_ratio, H, W, keep_mask);
  cam_expand_box<<<(n_maps + 127) / 128, 128, 0, stream>>>(ext, points, n_tot, n_maps, (float)W, (float)H, boxes);
  AS_LAUNCH_CHECK();
  return 0;
}
