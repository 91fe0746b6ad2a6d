// Device-side timing slots that also work inside CUDA graphs.
// bench.py wants the device time of every C-ABI call of a step.  torch.cuda.Event pairs cannot be used around calls that are
// captured into a CUDA graph (the backbone forward is replayed as one graph), so the library keeps its own pool of event
// pairs: recorded with cudaEventRecordExternal while the stream is capturing (they become event-record nodes and are
// re-recorded by every replay), plainly otherwise.  as_timer_elapsed reads a pair after the stream has been synchronised.
#include "common.cuh"

namespace {
constexpr int MAX_SLOTS = 8192;
cudaEvent_t g_ev[2 * MAX_SLOTS];
bool g_made[MAX_SLOTS];
}  // namespace

extern "C" int as_timer_slots(void) { return MAX_SLOTS; }

// which: 0 = start of the slot's interval, 1 = its end
extern "C" int as_timer_record(int slot, int which, cudaStream_t stream) {
  if (slot < 0 || slot >= MAX_SLOTS || (which != 0 && which != 1)) return AS_ERR_BAD_ARG;
  if (!g_made[slot]) {
    AS_CUDA(cudaEventCreate(&g_ev[2 * slot]));
    AS_CUDA(cudaEventCreate(&g_ev[2 * slot + 1]));
    g_made[slot] = true;
  }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  AS_CUDA(cudaStreamIsCapturing(stream, &st));
  const unsigned flags = st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
  AS_CUDA(cudaEventRecordWithFlags(g_ev[2 * slot + which], stream, flags));
  return 0;
}

extern "C" int as_timer_elapsed(int slot, float* ms) {
  if (slot < 0 || slot >= MAX_SLOTS || !g_made[slot] || !ms) return AS_ERR_BAD_ARG;
  AS_CUDA(cudaEventElapsedTime(ms, g_ev[2 * slot], g_ev[2 * slot + 1]));
  return 0;
}
