// Micro-benchmark (not part of the library): L2 -> shared-memory TMA fill bandwidth of the whole chip, the quantity that bounds the
// 3x3 conv layers (nine shifted boxes of an L2-resident tensor per tile) and the weight-tile re-reads of every non-resident layer.
//
//   mode 0  every CTA streams DISTINCT 16 KB boxes (128 rows x 128 B) of an L2-resident buffer   (activation tiles)
//   mode 1  every CTA streams the SAME sequence of boxes                                          (weight tiles, unicast)
//   mode 2  clusters of 2: each CTA loads HALF of the box and multicasts it to both CTAs          (weight tiles, multicast)
//   mode 3  clusters of 2, both CTAs load the same full boxes, no multicast                       (control for mode 2)
// One CTA per SM, one issuing thread, ring of 8 stages; reports delivered bytes (what lands in shared memory) per second and per
// SM clock.  `buf_mb` selects the footprint (<= 64 keeps it L2 resident; 1024 streams from HBM for comparison).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../cvpytorch_b200/csrc -I../include l2fill_bench.cu -o l2fill_bench
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cvb;

constexpr int kStages = 8;
constexpr int kBoxRows = 128;
constexpr int kBoxBytes = kBoxRows * 128;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

struct Params {
  int mode, iters, rows_total;
};

template <bool CLUSTER>
__device__ __forceinline__ void body(const CUtensorMap& tm, const CUtensorMap& tmh, Params p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kBoxBytes);
  uint64_t* empty = full + kStages;  // multicast: "both CTAs consumed the stage" (2 arrivals, local + remote)
  const uint32_t rank = CLUSTER ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);
    }
    fence_barrier_init();
  }
  if (CLUSTER) cluster_sync_all();
  else __syncthreads();
  if (threadIdx.x == 0) {
    const int tiles_total = p.rows_total / kBoxRows;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t phase_full = 0, phase_empty = 0;
    while (done < p.iters) {
      while (issued < p.iters && issued - done < kStages) {
        const int s = issued % kStages;
        int tile;
        if (p.mode == 0) tile = (int)(((long long)blockIdx.x * p.iters + issued) % tiles_total);
        else if (p.mode == 1) tile = issued % tiles_total;
        else tile = (int)(((long long)(blockIdx.x >> 1) * 7919 + issued) % tiles_total);
        if (p.mode == 2) {
          if (issued >= kStages) {  // the peer must also have consumed the previous contents of this stage
            mbar_wait(&empty[s], ((issued / kStages) - 1) & 1, 10 + s);
          }
          mbar_expect_tx(&full[s], kBoxBytes);
          tma_load_2d_mc(&tmh, &full[s], smem + s * kBoxBytes + rank * (kBoxBytes / 2), 0, tile * kBoxRows + (int)rank * (kBoxRows / 2), 3);
        } else {
          mbar_expect_tx(&full[s], kBoxBytes);
          tma_load_2d(&tm, &full[s], smem + s * kBoxBytes, 0, tile * kBoxRows);
        }
        ++issued;
      }
      const int s = done % kStages;
      mbar_wait(&full[s], (done / kStages) & 1, 100 + s);
      if (p.mode == 2) {  // tell both CTAs this stage is free here
        const uint32_t local = smem_u32(&empty[s]);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(local) : "memory");
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank ^ 1u));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
      ++done;
    }
    (void)phase_full;
    (void)phase_empty;
    cycles[blockIdx.x] = clock64() - t0;
  }
  if (CLUSTER) cluster_sync_all();
}

__global__ void __launch_bounds__(64) fill_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmh, Params p,
                                                  long long* cycles) {
  body<false>(tm, tmh, p, cycles);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64) fill_kernel_c2(const __grid_constant__ CUtensorMap tm,
                                                                              const __grid_constant__ CUtensorMap tmh, Params p, long long* cycles) {
  body<true>(tm, tmh, p, cycles);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  const int smem_bytes = kStages * kBoxBytes + 2 * kStages * 8 + 64;
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(fill_kernel_c2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* dcy;
  cudaMalloc(&dcy, sms * sizeof(long long));
  std::vector<long long> h(sms);
  printf("SMs %d, max clock %d MHz\n", sms, khz / 1000);
  printf("%6s %5s | %10s %12s %12s\n", "buf_MB", "mode", "ms", "TB/s", "B/clk/SM");
  for (int buf_mb : {16, 48, 1024}) {
    const size_t bytes = (size_t)buf_mb << 20;
    void* buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) {
      printf("alloc %d MB failed\n", buf_mb);
      continue;
    }
    cudaMemset(buf, 1, bytes);
    const int rows_total = (int)(bytes / 128);
    CUtensorMap tm, tmh;
    auto encode = [&](CUtensorMap* m, int box_rows) {
      const cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
      const cuuint64_t strides[1] = {128};
      const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
      const cuuint32_t es[2] = {1, 1};
      return reinterpret_cast<EncodeFn>(fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (encode(&tm, kBoxRows) != CUDA_SUCCESS || encode(&tmh, kBoxRows / 2) != CUDA_SUCCESS) {
      printf("encode failed\n");
      return 2;
    }
    for (int mode = 0; mode < 4; ++mode) {
      Params p{mode, 4096, rows_total};
      const int grid = sms & ~1;
      float best_ms = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode >= 2) fill_kernel_c2<<<grid, 64, smem_bytes>>>(tm, tmh, p, dcy);
        else fill_kernel<<<grid, 64, smem_bytes>>>(tm, tmh, p, dcy);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed (mode %d): %s\n", mode, cudaGetErrorString(e));
          return 1;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
      }
      cudaMemcpy(h.data(), dcy, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double delivered = (double)grid * p.iters * kBoxBytes;
      printf("%6d %5d | %10.3f %12.2f %12.1f\n", buf_mb, mode, best_ms, delivered / (best_ms * 1e-3) / 1e12, (double)p.iters * kBoxBytes / (double)mx);
    }
    cudaFree(buf);
  }
  return 0;
}
