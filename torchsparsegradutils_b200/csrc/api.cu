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
