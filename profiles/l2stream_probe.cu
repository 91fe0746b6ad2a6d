// How fast can TMA stream an L2-resident token set into shared memory on a B200?  (design input for the mean-shift kernel)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I attentionshift_b200/csrc profiles/l2stream_probe.cu \
//        attentionshift_b200/csrc/tma_host.cu -o profiles/l2stream_probe.bin
// Every CTA walks its 64-token units (round-robin over the grid) for every 64-channel block, loading hi and lo boxes of
// 8 KB each into a ring of `stages` x 32 KB (= 2 units x hi/lo, like the affinity pass); a consumer warp frees each stage
// as soon as it has landed.  Reports GB/s per configuration.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace asb;

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                                                    int n_img, int units_per_img, int kblocks, int stages, int passes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * 32768);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int total_pairs = n_img * units_per_img / 2;       // a stage = two consecutive units of one image
  int st = 0; uint32_t ph = 0;
  for (int pass = 0; pass < passes; ++pass)
    for (int kb = 0; kb < kblocks; ++kb)
      for (int pr = blockIdx.x; pr < total_pairs; pr += gridDim.x) {
        const int img = pr / (units_per_img / 2), u = (pr % (units_per_img / 2)) * 2;
        if (warp == 0) {
          if (lane == 0) {
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], 32768);
            uint8_t* dst = smem + (size_t)st * 32768;
            tma_load_3d(dst, &tm_hi, &full[st], kb * 64, u * 64, img);
            tma_load_3d(dst + 8192, &tm_hi, &full[st], kb * 64, (u + 1) * 64, img);
            tma_load_3d(dst + 16384, &tm_lo, &full[st], kb * 64, u * 64, img);
            tma_load_3d(dst + 24576, &tm_lo, &full[st], kb * 64, (u + 1) * 64, img);
          }
        } else {
          mbar_wait(&full[st], ph);
          if (lane == 0) mbar_arrive(&empty[st]);
        }
        if (++st == stages) { st = 0; ph ^= 1; }
      }
}

int main() {
  const int N = 4096, C = 768;
  const int max_img = 8;
  __half *hi, *lo;
  cudaMalloc(&hi, (size_t)max_img * N * C * 2);
  cudaMalloc(&lo, (size_t)max_img * N * C * 2);
  cudaMemset(hi, 0, (size_t)max_img * N * C * 2);
  cudaMemset(lo, 0, (size_t)max_img * N * C * 2);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int n_img, grid, stages; };
  std::vector<Cfg> cfgs = {{8, 128, 3}, {8, 128, 6}, {8, 148, 3}, {8, 148, 6}, {8, 256, 3}, {8, 296, 3}, {8, 296, 2}, {8, 444, 2},
                           {4, 128, 3}, {4, 148, 6}, {4, 296, 3}, {2, 148, 6}, {2, 296, 3}, {1, 148, 6}, {1, 296, 3}};
  for (const Cfg& c : cfgs) {
    CUtensorMap tm[2];
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)c.n_img};
    uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)N * C * 2};
    uint32_t box[3] = {64, 64, 1};
    if (as_encode_tmap(&tm[0], hi, 2, 3, dims, str, box) || as_encode_tmap(&tm[1], lo, 2, 3, dims, str, box)) { printf("tmap failed\n"); return 1; }
    const size_t smem = 1024 + (size_t)c.stages * 32768 + 256;
    const int passes = 20;
    stream_kernel<<<c.grid, 64, smem>>>(tm[0], tm[1], c.n_img, N / 64, C / 64, c.stages, 3);     // warm
    cudaEventRecord(e0);
    stream_kernel<<<c.grid, 64, smem>>>(tm[0], tm[1], c.n_img, N / 64, C / 64, c.stages, passes);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)passes * c.n_img * N * C * 4;
    printf("n_img %d (%.0f MB)  grid %3d  stages %d (%3zu KB smem/CTA): %7.1f GB/s   %.2f us / pass\n", c.n_img, c.n_img * N * C * 4 / 1e6,
           c.grid, c.stages, smem / 1024, bytes / ms / 1e6, ms * 1e3 / passes);
  }
  return 0;
}
