// Micro-benchmark (not part of the library): cycles per tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of
//   N            32 / 64 / 128 / 256
//   swizzle      64 B rows (BLOCK_K 32) or 128 B rows (BLOCK_K 64)
//   A layout     contiguous 8-row groups (standard tile) or groups 16 pixels apart with a row-shifted start (the halo-tile layout
//                of DESIGN.md 4.1 that turned out slower in the conv kernel)
//   accumulators one chain (every MMA accumulates into the same TMEM columns) or two alternating accumulators
//   operands     the same smem tiles every time, or four tiles in rotation (distinct shared-memory lines)
// One CTA per SM, one issuing thread, `iters` MMAs back to back, one commit at the end; time = clock64 around issue + completion.
// Purpose: decide whether the small-N conv layers are bound by a per-instruction floor, by A-operand delivery from shared memory,
// or by the accumulate dependency -- the open question behind the "MMA count, not bytes" observation in profiles/README.md.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../cvpytorch_b200/csrc -I../include mma_bench.cu -o mma_bench
// Run:   ./mma_bench            (prints a table; no inputs)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cvb;

struct Params {
  int n;           // MMA N
  int swz;         // 64 or 128
  int halo;        // 1: row-shifted start + stride byte offset of 16 rows
  int two_acc;     // 1: alternate between two accumulators
  int rotate;      // 1: rotate over 4 operand tiles
  int iters;
};

template <int SWZ>
__device__ __forceinline__ uint64_t desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  return make_kmajor_desc_sbo<SWZ>(smem_addr, sbo_bytes);
}

__global__ void __launch_bounds__(128) mma_bench_kernel(Params p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: 4 A regions of 48 KB (enough for the halo pitch), 4 B regions of 32 KB, barrier + tmem slot at the end
  uint8_t* sA = smem;
  uint8_t* sB = smem + 4 * 40960;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 40960 + 4 * 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (4 * 40960 + 4 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16_f32(128, p.n);
    const uint32_t row_bytes = (uint32_t)p.swz;
    const uint32_t std_sbo = 8u * row_bytes;
    const uint32_t halo_sbo = 16u * row_bytes;   // groups one 16-pixel image row apart
    const uint32_t halo_shift = row_bytes;        // start shifted by one pixel (unaligned 8-row group)
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const int tile = p.rotate ? (it & 3) : 0;
      const uint32_t a = smem_u32(sA + tile * 40960) + (p.halo ? halo_shift : 0u);
      const uint32_t b = smem_u32(sB + tile * 16384);
      const uint64_t da = p.swz == 128 ? desc_sbo<128>(a, p.halo ? halo_sbo : std_sbo) : desc_sbo<64>(a, p.halo ? halo_sbo : std_sbo);
      const uint64_t db = p.swz == 128 ? make_kmajor_desc<128>(b) : make_kmajor_desc<64>(b);
      const uint32_t d = tmem + (uint32_t)((p.two_acc && (it & 1)) ? 256 : 0);
      umma_f16(d, da, db, idesc, it >= (p.two_acc ? 2 : 1) ? 1u : 0u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0, 1);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem_bytes = 4 * 40960 + 4 * 16384 + 64;
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* dcy;
  cudaMalloc(&dcy, sms * sizeof(long long));
  std::vector<long long> h(sms);
  printf("%5s %4s %5s %7s %7s | %10s %10s\n", "N", "swz", "halo", "two_acc", "rotate", "cyc/MMA", "MAC/cyc/SM");
  const int ns[4] = {32, 64, 128, 256};
  for (int swz : {64, 128})
    for (int n : ns)
      for (int halo = 0; halo < 2; ++halo)
        for (int two = 0; two < 2; ++two)
          for (int rot = 0; rot < 2; ++rot) {
            if (two && n > 256) continue;
            Params p{n, swz, halo, two, rot, 4096};
            for (int rep = 0; rep < 2; ++rep) {  // first repetition warms up
              mma_bench_kernel<<<sms, 128, smem_bytes>>>(p, dcy);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("kernel failed: %s\n", cudaGetErrorString(e));
                return 1;
              }
            }
            cudaMemcpy(h.data(), dcy, sms * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (long long v : h) mx = v > mx ? v : mx;
            const double cyc = (double)mx / p.iters;
            printf("%5d %4d %5d %7d %7d | %10.1f %10.0f\n", n, swz, halo, two, rot, cyc, 128.0 * n * 16 / cyc);
          }
  return 0;
}
