// Micro-benchmark (not part of the library): does the operand ADDRESS PATTERN of the conv kernel's "halo" loader slow tcgen05.mma down?
// One CTA per SM, one issuing thread, the exact MMA sequence of one 64->64 3x3 tile (9 taps x 4 K steps x {N=128 pair MMA, N=64 cross
// MMA}) repeated `tiles` times, no TMA traffic, no epilogue.  Variants toggle, one at a time, what differs from the classic loader:
//   bit 0  A view start moves with the tap (row offsets inside one 184-row slot, unaligned to the 8-row swizzle atom)
//   bit 1  B tile moves with the tap (nine resident 16 KB tiles instead of one)
//   bit 2  stride between 8-row groups = 10 rows (1280 B) instead of 8 rows (1024 B)
//   bit 3  A_lo plane 23 KB after A_hi (instead of directly after the 16 KB tile)
//   bit 4  commit to an mbarrier after every tap (classic) instead of once per tile
//   bit 5  eight more warps poll an mbarrier (mbar_wait) for the whole run, like the conv kernel's epilogue warps waiting for an accumulator
//   bit 6  same, but the pollers __nanosleep(64) between polls
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../cvpytorch_b200/csrc -I../include mma_bench2.cu -o mma_bench2
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace cvb;

struct P {
  int variant, tiles;
};

__global__ void __launch_bounds__(384) k(P p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // A region: 2 planes x 24 KB; B region: 9 x 16 KB; barriers at the end
  uint8_t* sA = smem;
  uint8_t* sB = smem + 49152;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 9 * 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (49152 + 9 * 16384) / 4; i += blockDim.x) {
    // variant bit 7: pseudo-random fp16 operands in [-2, 2) instead of the constant 1.0 (does the MMA rate depend on the DATA?)
    uint32_t v = 0x3c003c00u;
    if (p.variant & 128) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83ff83ffu) | 0x3c003c00u;  // sign + mantissa random, exponent of 1.0
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
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
  if (threadIdx.x >= 128 && (p.variant & 96)) {  // pollers: wait until the issuing thread is done
    if (p.variant & 64) {
      while (!mbar_try_wait(&bar[0], 0)) __nanosleep(64);
    } else {
      mbar_wait(&bar[0], 0, 7);
    }
  }
  if (threadIdx.x < 32 && elect_one()) {
    const bool va = p.variant & 1, vb = p.variant & 2, vs = p.variant & 4, vl = p.variant & 8, vc = p.variant & 16;
    const uint32_t sbo = vs ? 1280u : 1024u;
    const uint32_t lo_off = vl ? 23552u : 16384u;
    constexpr uint32_t ID128 = make_idesc_f16_f32(128, 128), ID64 = make_idesc_f16_f32(128, 64);
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int tile = 0; tile < p.tiles; ++tile) {
      for (int t = 0; t < 9; ++t) {
        const uint32_t arow = va ? (uint32_t)((t / 3) * (vs ? 10 : 8) + (vs ? t % 3 : 0)) : 0u;
        const uint32_t a = smem_u32(sA) + arow * 128u;
        const uint32_t b = smem_u32(sB) + (vb ? (uint32_t)t * 16384u : 0u);
        const uint64_t dah = make_kmajor_desc_sbo<128>(a, sbo);
        const uint64_t dal = make_kmajor_desc_sbo<128>(a + lo_off, sbo);
        const uint64_t dbh = make_kmajor_desc<128>(b);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_f16(tmem, dah + 2 * ks, dbh + 2 * ks, ID128, (t | ks) ? 1u : 0u);
          umma_f16(tmem + 64, dal + 2 * ks, dbh + 2 * ks, ID64, 1u);
        }
        if (vc) umma_commit(&bar[1]);
      }
    }
    umma_commit(&bar[0]);
    mbar_wait(&bar[0], ph, 1);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem_bytes = 49152 + 9 * 16384 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* d;
  cudaMalloc(&d, sms * 8);
  std::vector<long long> h(sms);
  printf("variant bits: 1=A moves 2=B moves 4=sbo 1280 8=lo plane +23K 16=commit per tap\n%8s | %10s\n", "variant", "cyc/MMA");
  for (int v : {0, 15, 128, 143, 159}) {
    P p{v, 64};
    for (int rep = 0; rep < 2; ++rep) {
      k<<<sms, 384, smem_bytes>>>(p, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("variant %d failed: %s\n", v, cudaGetErrorString(e));
        return 1;
      }
    }
    cudaMemcpy(h.data(), d, sms * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long x : h) mx = x > mx ? x : mx;
    printf("%8d | %10.1f\n", v, (double)mx / (p.tiles * 72.0));
  }
  return 0;
}
