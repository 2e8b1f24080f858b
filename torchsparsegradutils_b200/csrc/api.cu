// api.cu -- library-wide C-ABI entry points (version, error strings, launch accounting).
#include "common.cuh"

namespace tsgu {
unsigned long long g_launches = 0;
thread_local int g_sm_margin = 0;
}

extern "C" int tsgu_set_sm_margin(int sms) {
  const int prev = tsgu::g_sm_margin;
  tsgu::g_sm_margin = sms < 0 ? 0 : sms;
  return prev;
}

// ---- host mailbox: a few words of device state made visible to the host WITHOUT a copy-engine transfer.  A new
// pattern needs a handful of scalars on the host (longest row, verdict of the window plan, padded size); as a
// cudaMemcpy those reads queue behind whatever bulk D2H traffic the application has in flight (measured: +1 ms per
// read under a result stream of 68 MB per item).  A kernel store into mapped pinned memory does not.
namespace {
__global__ void publish_kernel(const int* __restrict__ src, volatile int* __restrict__ dst, int words) {
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
}  // namespace

extern "C" int tsgu_mailbox_create(size_t bytes, void** host_ptr, void** dev_ptr) {
  if (!host_ptr || !dev_ptr || bytes == 0) return TSGU_ERR_SHAPE;
  void* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
  if (e != cudaSuccess) return (int)e;
  void* d = nullptr;
  e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) {
    cudaFreeHost(h);
    return (int)e;
  }
  *host_ptr = h;
  *dev_ptr = d;
  return 0;
}

extern "C" int tsgu_mailbox_destroy(void* host_ptr) {
  return host_ptr ? (int)cudaFreeHost(host_ptr) : 0;
}

extern "C" int tsgu_publish(const void* src, void* mailbox_dev, size_t bytes, void* stream) {
  if ((bytes & 3) || ((uintptr_t)src & 3) || ((uintptr_t)mailbox_dev & 3)) return TSGU_ERR_SHAPE;
  if (bytes == 0) return 0;
  publish_kernel<<<1, 64, 0, tsgu::as_stream(stream)>>>((const int*)src, (volatile int*)mailbox_dev, (int)(bytes / 4));
  tsgu::count_launch();
  return tsgu::launch_status();
}

extern "C" int tsgu_version(void) { return TSGU_ABI_VERSION; }

extern "C" int64_t tsgu_launch_count(void) {
  return (int64_t)__atomic_load_n(&tsgu::g_launches, __ATOMIC_RELAXED);
}

extern "C" const char* tsgu_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case TSGU_ERR_DTYPE: return "tsgu: unknown value/index dtype";
    case TSGU_ERR_SHAPE: return "tsgu: invalid shape or size argument";
    case TSGU_ERR_WORKSPACE: return "tsgu: workspace missing or too small";
    case TSGU_ERR_ALGO: return "tsgu: unknown algorithm selector";
    case TSGU_ERR_RANGE: return "tsgu: sizes do not fit the requested index dtype";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "tsgu: unknown error";
}
