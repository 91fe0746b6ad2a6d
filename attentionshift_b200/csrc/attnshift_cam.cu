// CAM slicing, x16 bilinear up-sampling evaluated on the fly, and CAM -> pseudo box (connected components).
// Reference: RH:2272-2275 (slice point-token rows of the roll-out, bilinear x16, align_corners=False) and
// get_bbox_from_cam_fast RH:60-116 (min-max normalise, binarise at seed_thr, cc_torch labelling, keep components with
// area >= seed_multiple * largest, joint extent, mirror-expand around the GT point).
// The up-sampled [7, n_gt, H, W] CAM stack (29 MB per instance at 1024^2) is never written: every consumer
// re-evaluates the 4-tap interpolation from the [hp, wp] map (16 KB) with the exact arithmetic of
// torch's CPU kernel:  t = fma(v0, wx0, v1*wx1);  out = fma(t0, wy0, t1*wy1).
// Connected components: 8-connectivity union-find over the foreground RUNS of each row (label = smallest run index of
// the component); only the partition matters to the caller.  cc_torch itself is absent from the reference tree (parity unpinned, see oracle).
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

// ---------------------------------------------------------------- connected components on run lists
// The binarised map is a x16 bilinear up-sampling: a row has a handful of foreground runs, so the maps are never
// labelled pixel by pixel.  Two kernels:
//   ccl_bitmap  (fully parallel)  fg bit of every pixel -> bitmap [n_maps][H][W/32] (22 MB for 168 maps of 1024^2)
//   ccl_runs    (one CTA per map) bitmap -> compact run list in row order (thread = row, one block scan),
//               8-connectivity union-find over the RUNS with the labels in shared memory (a run meets the runs of the
//               previous row whose columns overlap [start-1, end+1]; their index range is a bitmap rank: prefix count +
//               popcount), areas by one atomic per distinct root and warp, the largest area, the joint
//               extent of the kept components, and the mirror-expanded box (RH:97-115) -- everything the pixel-based
//               version spread over five full-resolution passes (~8 GB of traffic per batch).
// label = smallest run index of the component; only the partition matters to the caller.
// fg = (v - mn) / clamp(mx - mn, 1e-6) >= thr      (RH:63-66)
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

constexpr int BM_ROWS = 32;       // rows per CTA of ccl_bitmap
// grid (ceil(H / BM_ROWS), n_maps), 256 threads: thread = column (strided), the column's horizontal tap lives in registers
__global__ void __launch_bounds__(256)
ccl_bitmap(const float* __restrict__ lows, const float* __restrict__ mmf, int hp, int wp, float thr,
           unsigned* __restrict__ bits) {
  extern __shared__ float low_rows[];              // the source rows this CTA touches
  const int m = blockIdx.y, y0 = blockIdx.x * BM_ROWS;
  const int H = hp * 16, W = wp * 16, nw = (W + 31) / 32;
  const float* low = lows + (size_t)m * hp * wp;
  const float mn = mmf[2 * m], den = fmaxf(mmf[2 * m + 1] - mn, 1e-6f), rden = __frcp_rn(den);
  const int y1 = min(y0 + BM_ROWS, H);
  const int r_lo = tap_up16(y0, hp).i0, r_hi = tap_up16(y1 - 1, hp).i1;
  const int nr = r_hi - r_lo + 1;
  for (int i = threadIdx.x; i < nr * wp; i += blockDim.x) low_rows[i] = low[r_lo * wp + i];
  __syncthreads();
  for (int x0 = (threadIdx.x >> 5) * 32; x0 < nw * 32; x0 += blockDim.x) {
    const int x = x0 + lane_id();
    const Tap tx = tap_up16(min(x, W - 1), wp);
    for (int y = y0; y < y1; ++y) {
      const Tap ty = tap_up16(y, hp);
      const float* ra = low_rows + (ty.i0 - r_lo) * wp;
      const float* rb = low_rows + (ty.i1 - r_lo) * wp;
      const float v = lerp2(ra[tx.i0], ra[tx.i1], rb[tx.i0], rb[tx.i1], ty, tx);
      // (v - mn) / den >= thr with the IEEE division only where a reciprocal multiply (<= 2e-7 relative off) cannot decide
      const float num = v - mn, t = num * rden;
      bool fg = t >= thr;
      if (fabsf(t - thr) <= 1e-6f * fmaxf(1.f, fabsf(t))) fg = num / den >= thr;
      const unsigned b = __ballot_sync(0xffffffffu, fg && x < W);
      if (lane_id() == 0) bits[((size_t)m * H + y) * nw + (x0 >> 5)] = b;
    }
  }
}

constexpr int CR_THREADS = 1024;
constexpr int CR_SMEM_RUNS = 24576;              // labels of up to this many runs live in shared memory (96 KB, 2 CTAs / SM)

__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot /* smem [33] */, int* total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_tot[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;                      // exclusive warp offsets
    if (lane == 31) warp_tot[32] = winc;
  }
  __syncthreads();
  const int r = warp_tot[warp] + inc - v;
  *total = warp_tot[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ unsigned run_starts(const unsigned* row_bits, int wx) {
  const unsigned word = row_bits[wx];
  const unsigned prev = wx > 0 ? (row_bits[wx - 1] >> 31) : 0u;
  return word & ~((word << 1) | prev);
}

// one CTA per map
__global__ void __launch_bounds__(CR_THREADS)
ccl_runs(const unsigned* __restrict__ bits_all, int H, int W, int cap, unsigned short* __restrict__ rs_all,
         unsigned short* __restrict__ re_all, unsigned short* __restrict__ ry_all, int* __restrict__ label_all,
         int* __restrict__ area_all, int* __restrict__ wordoff_all, float ratio, const float* __restrict__ points, int n_tot,
         float* __restrict__ boxes, unsigned char* __restrict__ keep_mask) {
  extern __shared__ int label_s[];                 // CR_SMEM_RUNS ints
  __shared__ int scan_s[33];
  __shared__ int red_s[5][32];
  const int m = blockIdx.x, tid = threadIdx.x, lane = lane_id();
  const int nw = (W + 31) / 32;
  const unsigned* bits = bits_all + (size_t)m * H * nw;
  unsigned short* rs = rs_all + (size_t)m * cap;
  unsigned short* re = re_all + (size_t)m * cap;
  unsigned short* ry = ry_all + (size_t)m * cap;
  int* area = area_all + (size_t)m * cap;
  int* word_off = wordoff_all + (size_t)m * ((size_t)H * nw + 1);   // runs that start before word (y, wx), row-major

  // ---- phase 1: run list in (row, column) order; thread = row (one block scan per 1024 rows)
  int base = 0;
  for (int y0 = 0; y0 < H; y0 += CR_THREADS) {
    const int y = y0 + tid;
    const unsigned* row = bits + (size_t)y * nw;
    int cnt = 0;
    if (y < H)
      for (int wx = 0; wx < nw; ++wx) cnt += __popc(run_starts(row, wx));
    int tot;
    int off = base + block_excl_scan(cnt, scan_s, &tot);
    if (y < H) {
      for (int wx = 0; wx < nw; ++wx) {
        word_off[(size_t)y * nw + wx] = off;
        const unsigned word = row[wx];
        unsigned starts = run_starts(row, wx);
        while (starts) {
          const int b = __ffs(starts) - 1;
          starts &= starts - 1;
          // end of the run that starts at bit b: first zero at or after b, possibly in a later word of the row
          const unsigned inv = ~word & ~((1u << b) - 1u);
          int end;
          if (inv) {
            end = (wx << 5) + __ffs(inv) - 2;
          } else {
            int ww = wx + 1;
            unsigned nxt = 0;
            while (ww < nw && (nxt = row[ww]) == 0xffffffffu) ++ww;
            end = (ww < nw) ? (ww << 5) + __ffs(~nxt) - 2 : nw * 32 - 1;
          }
          rs[off] = (unsigned short)((wx << 5) + b);
          re[off] = (unsigned short)min(end, W - 1);
          ry[off] = (unsigned short)y;
          ++off;
        }
      }
    }
    base += tot;
  }
  const int n_runs = base;
  int* label = n_runs <= CR_SMEM_RUNS ? label_s : label_all + (size_t)m * cap;
  for (int i = tid; i < n_runs; i += CR_THREADS) { label[i] = i; area[i] = 0; }
  __syncthreads();

  // ---- phase 2: unions with the overlapping runs of the previous row (8-connectivity: columns [start-1, end+1]).
  // Their index range follows from the bitmap: rank of a column = runs started up to it = word_off + popcount, all loads
  // independent.  Runs are visited in row order, 1024 at a time, and compressed after every batch, so the chains a
  // later find() has to walk stay a couple of links long.
  for (int i0 = 0; i0 < n_runs; i0 += CR_THREADS) {
    const int i = i0 + tid;
    if (i < n_runs) {
      const int y = ry[i];
      if (y > 0) {
        const unsigned* prow = bits + (size_t)(y - 1) * nw;
        const int* poff = word_off + (size_t)(y - 1) * nw;
        const int s = (int)rs[i] - 1, e = min((int)re[i] + 1, W - 1);
        const int we = e >> 5;
        const int hi = poff[we] + __popc(run_starts(prow, we) & (0xffffffffu >> (31 - (e & 31))));     // starts at columns <= e
        int lo = poff[0];
        if (s >= 0) {
          const int ws = s >> 5;
          lo = poff[ws] + __popc(run_starts(prow, ws) & (0xffffffffu >> (31 - (s & 31))));               // starts at columns <= s
          if ((prow[ws] >> (s & 31)) & 1u) --lo;    // the run that covers column s reaches into the window
        }
        for (int j = lo; j < hi; ++j) uf_union(label, i, j);
      }
    }
    __syncthreads();
    if (i < n_runs) label[i] = uf_find(label, i);
    __syncthreads();
  }
  // final flattening: label = root
  for (int i = tid; i < n_runs; i += CR_THREADS) {
    const int r = uf_find(label, i);
    label[i] = r;
  }
  __syncthreads();

  // ---- phase 3: areas (runs of a warp mostly share their root: one atomic per distinct root and warp), largest area
  for (int i0 = 0; i0 < n_runs; i0 += CR_THREADS) {
    const int i = i0 + tid;
    const bool on = i < n_runs;
    const int root = on ? label[i] : -1;
    const int len = on ? (int)re[i] - (int)rs[i] + 1 : 0;
    unsigned todo = __ballot_sync(0xffffffffu, on);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int r0 = __shfl_sync(0xffffffffu, root, leader);
      const bool mine = on && root == r0;
      const int sum = warp_sum(mine ? len : 0);
      if (lane == leader) atomicAdd(&area[r0], sum);
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
  __syncthreads();
  int best = 0;
  for (int i = tid; i < n_runs; i += CR_THREADS) best = max(best, area[i]);
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) red_s[0][tid >> 5] = best;
  __syncthreads();
  best = red_s[0][lane];
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  const float need = ratio * (float)best;

  // ---- phase 4: joint extent of the kept components (area >= ratio * largest, RH:73-83)
  int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
  for (int i = tid; i < n_runs; i += CR_THREADS) {
    if ((float)area[label[i]] >= need) {
      x0 = min(x0, (int)rs[i]); x1 = max(x1, (int)re[i]);
      y0 = min(y0, (int)ry[i]); y1 = max(y1, (int)ry[i]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) { red_s[1][tid >> 5] = x0; red_s[2][tid >> 5] = y0; red_s[3][tid >> 5] = x1; red_s[4][tid >> 5] = y1; }
  __syncthreads();
  if (tid < 32) {
    x0 = red_s[1][tid]; y0 = red_s[2][tid]; x1 = red_s[3][tid]; y1 = red_s[4][tid];
    for (int o = 16; o > 0; o >>= 1) {
      x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
      x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (tid == 0) {
      // 'expand' (RH:97-115): mirror the far side of the kept extent around the GT point, clip to the image
      float* b = boxes + 4 * m;
      if (x1 < 0) { b[0] = 0.f; b[1] = 0.f; b[2] = 1.f; b[3] = 1.f; }
      else {
        const int o = m % n_tot;
        const float img_w = (float)W, img_h = (float)H;
        const float px0 = (float)x0, py0 = (float)y0, px1 = (float)x1, py1 = (float)y1;
        const float xc = points[2 * o], yc = points[2 * o + 1];
        float bx0, bx1, by0, by1;
        if (fabsf(xc - px0) > fabsf(xc - px1)) { bx0 = px0; bx1 = xc * 2.f - bx0; bx1 = bx1 < img_w ? bx1 : img_w; }
        else { bx1 = px1; bx0 = xc * 2.f - bx1; bx0 = bx0 > 0.f ? bx0 : 0.f; }
        if (fabsf(yc - py0) > fabsf(yc - py1)) { by0 = py0; by1 = yc * 2.f - by0; by1 = by1 < img_h ? by1 : img_h; }
        else { by1 = py1; by0 = yc * 2.f - by1; by0 = by0 > 0.f ? by0 : 0.f; }
        b[0] = bx0; b[1] = by0; b[2] = bx1; b[3] = by1;
      }
    }
  }
  // ---- optional (tests / visualisation): the reference's remained_label_masks
  if (keep_mask) {
    unsigned char* km = keep_mask + (size_t)m * H * W;
    for (size_t i = tid; i < (size_t)H * W; i += CR_THREADS) km[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_runs; i += CR_THREADS)
      if ((float)area[label[i]] >= need)
        for (int x = rs[i]; x <= (int)re[i]; ++x) km[(size_t)ry[i] * W + x] = 1;
  }
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

// run capacity per map: a row of W pixels holds at most ceil(W / 2) runs
static size_t ccl_cap(int H, int W) { return (size_t)H * ((W + 1) / 2); }

extern "C" size_t as_cam_bbox_workspace(int n_maps, int H, int W) {
  const size_t cap = ccl_cap(H, W);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return al((size_t)n_maps * H * ((W + 31) / 32) * 4) + 3 * al((size_t)n_maps * cap * 2) + 2 * al((size_t)n_maps * cap * 4) +
         al((size_t)n_maps * ((size_t)H * ((W + 31) / 32) + 1) * 4) + 1024;
}

// boxes [n_maps,4] for maps ordered [layer][instance] (n_maps = L * n_tot); points [n_tot,2].
// keep_mask (optional) [n_maps,H,W] uint8 = the reference's remained_label_masks.
extern "C" int as_cam_bbox(const float* lows, const float* minmax, const float* points, int n_maps, int n_tot, int hp,
                           int wp, float cam_thr, float area_ratio, float* boxes, unsigned char* keep_mask,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_maps <= 0) return 0;
  const int H = hp * 16, W = wp * 16;
  if (W > 65535 || H > 65535 || workspace_bytes < as_cam_bbox_workspace(n_maps, H, W)) return AS_ERR_BAD_ARG;
  const size_t cap = ccl_cap(H, W);
  const int nw = (W + 31) / 32;
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t x) { char* r = base + off; off += (x + 255) & ~(size_t)255; return r; };
  unsigned* bits = (unsigned*)take((size_t)n_maps * H * nw * 4);
  unsigned short* rs = (unsigned short*)take((size_t)n_maps * cap * 2);
  unsigned short* re = (unsigned short*)take((size_t)n_maps * cap * 2);
  unsigned short* ry = (unsigned short*)take((size_t)n_maps * cap * 2);
  int* label = (int*)take((size_t)n_maps * cap * 4);
  int* area = (int*)take((size_t)n_maps * cap * 4);
  int* word_off = (int*)take((size_t)n_maps * ((size_t)H * nw + 1) * 4);
  const size_t smem = (size_t)(BM_ROWS / 16 + 3) * wp * 4;
  ccl_bitmap<<<dim3((H + BM_ROWS - 1) / BM_ROWS, n_maps), 256, smem, stream>>>(lows, minmax, hp, wp, cam_thr, bits);
  static bool attr = false;
  if (!attr) {
    AS_CUDA(cudaFuncSetAttribute(ccl_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_RUNS * 4));
    attr = true;
  }
  ccl_runs<<<n_maps, CR_THREADS, CR_SMEM_RUNS * 4, stream>>>(bits, H, W, (int)cap, rs, re, ry, label, area, word_off, area_ratio,
                                                             points, n_tot, boxes, keep_mask);
  AS_LAUNCH_CHECK();
  return 0;
}
