// Host-side helper (no device code): the first k raw 32-bit outputs of at::mt19937 seeded like
// torch.Generator().manual_seed(seed) -- what torch.randint / torch.randperm consume on the CPU (RH:368 randint, RH:447
// randperm run on the default CPU generator in the reference).  The attention-shift host wrapper draws the random seed /
// mask-point indices of a whole batch from these in one vectorised step instead of one torch call per instance.
#include <stdint.h>
#include <vector>

extern "C" int as_mt19937_draws(const uint32_t* seeds, int n_keys, int k, uint32_t* out) {
  if (n_keys < 0 || k < 0 || k > 624) return 10001;
  constexpr int N = 624, M = 397;
  std::vector<uint32_t> st(N);
  for (int key = 0; key < n_keys; ++key) {
    st[0] = seeds[key];
    for (int j = 1; j < N; ++j) st[j] = 1812433253u * (st[j - 1] ^ (st[j - 1] >> 30)) + (uint32_t)j;
    auto twist = [](uint32_t u, uint32_t v) {
      return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
    };
    // first generation of the state (only the first k words are needed; word i reads old words i, i+1 and i+M, or the
    // already regenerated word i+M-N once i >= N-M)
    for (int i = 0; i < k; ++i) {
      const uint32_t a = st[i], b = st[(i + 1) % N];
      const uint32_t c = st[(i + M) % N];        // for i >= N-M this index is < i: already the NEW value, as in next_state()
      st[i] = c ^ twist(a, (i + 1 < N) ? b : st[0]);
    }
    for (int i = 0; i < k; ++i) {
      uint32_t y = st[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      out[(size_t)key * k + i] = y;
    }
  }
  return 0;
}
