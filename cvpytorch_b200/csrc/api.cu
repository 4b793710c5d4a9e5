// Library-level pieces of the C ABI: error text, version, launch counter, driver entry point lookup.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "internal.h"

namespace cvb {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

}  // namespace cvb

extern "C" const char* cvb_last_error_string(void) { return cvb::g_err; }
extern "C" int cvb_version(void) { return 100; /* 0.1.0 */ }
extern "C" int64_t cvb_launch_count(void) { return cvb::g_launches.load(std::memory_order_relaxed); }
