// Hardware probe (not part of the library): can ONE TMA-loaded halo image in shared memory feed all nine taps of a 3x3 conv
// through row-shifted UMMA shared-memory descriptors?
//
// Layout under test: halo image [18 rows][16 px][32 ch fp16] = 64-byte rows, 64B swizzle (as written by one TMA box), output
// tile 8 x 16 pixels (m = y*8 + x).  Tap (dy,dx) uses descriptor start = base + ((dy+1)*16 + (dx+1))*64 B with stride byte
// offset (8-row group pitch) = 16 px * 64 B = 1024 B.  B = 32x32 identity, so D[m][n] must equal X[y0+y+dy][x0+x+dx][n]
// (zero outside the image).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../cvpytorch_b200/csrc -lcuda
// Prints PASS/FAIL per (tile, tap).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cvb;

constexpr int H = 20, W = 24, C = 32;
constexpr int HALO_ROWS = 18, HALO_PITCH = 16;
constexpr int A_BYTES = HALO_ROWS * HALO_PITCH * C * 2;  // 18432

template <int SWZ>
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  constexpr uint64_t layout = SWZ == 128 ? 2ull : (SWZ == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, int x0, int y0, int dy, int dx, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // halo image
  uint8_t* sB = smem + 20480;         // 32 x 64 B identity (SW64), 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 20480 + 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  // B[n][k] = (n == k), K-major rows of 64 B, 64B swizzle: 16-byte chunk c of row n lives at chunk c ^ ((n >> 1) & 3)
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    const int chunk = (k / 8) ^ ((n >> 1) & 3);
    reinterpret_cast<__half*>(sB + n * 64 + chunk * 16)[k % 8] = __float2half(n == k ? 1.0f : 0.0f);
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], A_BYTES);
    tma_load_3d(&tmA, &bars[0], sA, 0, x0 - 1, y0 - 1);
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16_f32(128, 32);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)(((dy + 1) * HALO_PITCH + (dx + 1)) * 64);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t da = make_desc_sbo<64>(a0 + ks * 32, HALO_PITCH * 64);
      const uint64_t db = make_kmajor_desc<64>(smem_u32(sB) + ks * 32);
      umma_f16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, 2);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  const int m = warp * 32 + lane;
  for (int n = 0; n < 32; ++n) out[m * 32 + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

int main() {
  std::vector<__half> hx((size_t)H * W * C);
  auto X = [&](int y, int x, int c) -> float { return (c & 1) ? (float)c : (float)(y * W + x + 1); };
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) hx[((size_t)y * W + x) * C + c] = __float2half(X(y, x, c));
  __half* dx_;
  cudaMalloc(&dx_, hx.size() * 2);
  cudaMemcpy(dx_, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  float* dout;
  cudaMalloc(&dout, 128 * 32 * 4);

  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  CUtensorMap tm;
  const cuuint64_t dims[3] = {C, W, H};
  const cuuint64_t strides[2] = {C * 2, (cuuint64_t)W * C * 2};
  const cuuint32_t box[3] = {C, HALO_PITCH, HALO_ROWS};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dx_, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    return 2;
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  int fails = 0;
  const int tiles[3][2] = {{0, 0}, {8, 0}, {16, 4}};  // (x0, y0): left/top border, interior, right/bottom border
  std::vector<float> ho(128 * 32);
  for (auto& t : tiles)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        cudaMemset(dout, 0xFF, 128 * 32 * 4);
        probe_kernel<<<1, 128, 32768>>>(tm, t[0], t[1], dy, dx, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel error: %s\n", cudaGetErrorString(e));
          return 3;
        }
        cudaMemcpy(ho.data(), dout, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            const int y = t[1] + m / 8 + dy, x = t[0] + m % 8 + dx;
            const float want = (y < 0 || y >= H || x < 0 || x >= W) ? 0.0f : X(y, x, n);
            if (ho[m * 32 + n] != want) {
              if (first < 0) first = m * 32 + n;
              ++bad;
            }
          }
        printf("tile(%2d,%2d) tap(%2d,%2d): %s", t[0], t[1], dy, dx, bad ? "FAIL" : "PASS");
        if (bad) printf("  %d mismatches, first at m=%d n=%d got %g", bad, first / 32, first % 32, ho[first]);
        printf("\n");
        fails += bad != 0;
      }
  printf(fails ? "PROBE FAILED (%d cases)\n" : "PROBE PASSED\n", fails);
  return fails ? 1 : 0;
}
