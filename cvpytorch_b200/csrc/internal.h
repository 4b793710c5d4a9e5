// Internal helpers shared by the translation units of libcvb200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

#include "../../include/cvb200.h"

namespace cvb {

// thread-local error text returned by cvb_last_error_string()
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define CVB_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::cvb::set_error(CVB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define CVB_REQUIRE(cond, ...)                                       \
  do {                                                               \
    if (!(cond)) return ::cvb::set_error(CVB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// NMS score histogram (shared by nms.cu and the decode kernel): bin = float_bits(score) >> 17, per image
constexpr int kNmsBins = 16384;
constexpr int kNmsCap = 65536;   // candidate key capacity per image (phase B)
constexpr int kNmsCapA = 4096;   // phase-A candidate capacity = one shared-memory sort chunk

// NMS workspace layout (bytes): [hist B*kNmsBins u32][6 slots of B u32][keysB B*kNmsCap u64][keysA B*kNmsCapA u64][rowmax B*A f32]
// rowmax[b][anchor] = best score (obj*cls, > conf) of the anchor row, 0 when it has no candidate: written by the pass that
// builds the histogram (decode kernel or the NMS count pass) and lets the emit passes skip rows below the threshold bin.
inline size_t nms_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t nms_ws_head_bytes(int B) { return nms_align_up((size_t)B * kNmsBins * 4, 256) + 6 * nms_align_up((size_t)B * 4, 256); }
inline size_t nms_ws_rowmax_offset(int B) { return nms_ws_head_bytes(B) + (size_t)B * (kNmsCap + kNmsCapA) * 8; }

#ifdef __CUDACC__
// sigmoid with two MUFU ops (ex2, rcp), |rel err| ~ 2^-22 (the reference's own CPU sigmoid is ~1 ulp; parity budget 1e-3).  Shared by
// the YOLOv5 / YOLOX / FCOS decode kernels so that every pass that recomputes a score gets the identical value.
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
#endif

}  // namespace cvb
