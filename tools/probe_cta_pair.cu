// Hardware probe (not part of the library; to be run at the start of the next round): the minimal CTA-pair (cta_group::2) MMA
// protocol the conv kernel would adopt.
//
//   cluster of 2 CTAs; each CTA loads ITS 128 rows of A (K = 32 fp16, 64-byte rows, 64B swizzle) and ITS half of B (N/2 = 32 of
//   N = 64 rows) by TMA, both signalling the LEADER's mbarrier (.cta_group::2 TMA form, barrier address with the peer bit cleared);
//   the leader's elected thread issues ONE accumulation chain of two tcgen05.mma.cta_group::2 (K steps of 16) with M = 256, N = 64;
//   tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11) releases a "done" barrier in BOTH CTAs; each CTA reads its own
//   128 TMEM lanes and stores its 128 x 64 block of D.  Host check: D == A * B^T exactly (small integers).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../cvpytorch_b200/csrc -I../include probe_cta_pair.cu -o probe_cta_pair -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cvb;

constexpr int M = 256, N = 64, K = 32;
constexpr int A_HALF_BYTES = 128 * K * 2;      // 8192
constexpr int B_HALF_BYTES = (N / 2) * K * 2;  // 2048

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's shared memory, the transaction bytes are counted on the LEADER's barrier
// (same offset, peer bit 24 of the shared::cluster address cleared -- the convention of CUTLASS' SM100_TMA_2SM_LOAD).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                  const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;             // this CTA's 128 rows of A
  uint8_t* sB = smem + 8192;      // this CTA's N/2 rows of B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8192 + 2048);  // [0] operands landed (leader's copy is used), [1] MMA done
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 0) {
    tmem_alloc_pair(slot, 64);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_expect_tx(&bars[0], 2 * (A_HALF_BYTES + B_HALF_BYTES));  // both CTAs' shares land on the leader's barrier
    tma_load_2d_pair(&tmA, &bars[0], sA, 0, (int)rank * 128);
    tma_load_2d_pair(&tmB, &bars[0], sB, 0, (int)rank * (N / 2));
    if (rank == 0) {
      mbar_wait(&bars[0], 0, 1);
      tc_fence_after();
      // instruction descriptor: M = 256 across the pair, N = 64
      const uint32_t idesc = make_idesc_f16_f32(256, N);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t da = make_kmajor_desc<64>(smem_u32(sA) + ks * 32);
        const uint64_t db = make_kmajor_desc<64>(smem_u32(sB) + ks * 32);
        umma_f16_pair(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit_pair(&bars[1]);
    }
  }
  mbar_wait(&bars[1], 0, 2);
  tc_fence_after();
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32), v);
    tmem_ld_wait();
    const int row = (int)rank * 128 + warp * 32 + lane;
    for (int n = 0; n < 32; ++n) out[row * N + half * 32 + n] = __uint_as_float(v[n]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 64);
}

int main() {
  std::vector<__half> ha((size_t)M * K), hb((size_t)N * K);
  auto A = [](int m, int k) { return (float)((m * 7 + k * 3) % 11 - 5); };
  auto Bv = [](int n, int k) { return (float)((n * 5 + k) % 7 - 3); };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) ha[(size_t)m * K + k] = __float2half(A(m, k));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hb[(size_t)n * K + k] = __float2half(Bv(n, k));
  __half *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, (size_t)M * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xFF, (size_t)M * N * 4);

  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  auto encode = [&](CUtensorMap* tm, void* base, int rows, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)K, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    return reinterpret_cast<EncodeFn>(fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap tmA, tmB;
  if (encode(&tmA, da, M, 128) != CUDA_SUCCESS || encode(&tmB, db, N, N / 2) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 2;
  }
  cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe_pair_kernel<<<2, 128, 16384>>>(tmA, tmB, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel error: %s\n", cudaGetErrorString(e));
    return 3;
  }
  std::vector<float> ho((size_t)M * N);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += A(m, k) * Bv(n, k);
      if (ho[(size_t)m * N + n] != ref) {
        if (first < 0) first = m * N + n;
        ++bad;
      }
    }
  if (bad) printf("PROBE FAILED: %d mismatches, first at m=%d n=%d got %g\n", bad, first / N, first % N, ho[first]);
  else printf("PROBE PASSED: cta_group::2 MMA, M=256 N=64 K=32\n");
  return bad != 0;
}
