// Training-step kernels for the YOLOX C3 block: conv forward / backward-data / backward-weight as bf16 implicit
// GEMMs on tcgen05 + TMA, and the BatchNorm(batch statistics) + SiLU forward / backward element-wise passes around them.
//
// Replaces, for one  BaseConv = nn.Conv2d -> nn.BatchNorm2d -> SiLU  in training mode (src/models/modules/yolox_modules.py:35-55) and
// its autograd backward (what trainer.py:177-207 runs through cuDNN / ATen):
//   forward   y = conv(x, W)                         tconv_kernel            (M = 128 pixels, N = cout tile, K = taps x cin)
//             batch mean / var of y                  bn_stats_kernel + bn_finalize_kernel
//             a = silu(gamma * (y - mean) * rstd + beta)     bn_silu_fwd_kernel
//   backward  dz = da * silu'(z);  sum(dz), sum(dz * xhat)   bn_silu_bwd_reduce_kernel   (or dz straight from the consumer's dgrad epilogue)
//             dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))   bn_silu_bwd_apply_kernel
//             dx = conv(dy, W rotated by 180 degrees, in/out swapped)   tconv_kernel again (optionally with the PRODUCER's SiLU' in the epilogue)
//             dW[co][tap][ci] = sum_pixels dy[p][co] * x[p + tap][ci]    twgrad_kernel   (K = pixels: both operands MN-major in shared memory)
//
// Layout: activations and gradients are NHWC bf16, dense (pitch == C, C a multiple of 64); weights bf16 [cout][kh*kw][cin] (K-major rows),
// fp32 master weights / fp32 weight gradients outside.  One TMA box {64 channels, TW, TH, NB} with TW*TH*NB == 128 pixels is at the same
// time a K-major A tile of the forward GEMM (row = pixel, 128 bytes of channels) and an MN-major operand of the weight-gradient GEMM
// (K = pixel rows); out-of-bounds pixels are zero-filled by TMA = the convolution padding, and contribute nothing to dW.
#include <cuda_bf16.h>

#include <cstring>
#include <new>

#include "internal.h"
#include "ptx.cuh"

namespace cvb {

constexpr int kTThreads = 192;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (one per TMEM lane quarter)
constexpr int kTStagesFwd = 4;
constexpr int kTStagesWg = 3;

// instruction descriptor, kind::f16 with bf16 operands and fp32 accumulation; a_mn / b_mn = 1: operand is MN-major in shared memory
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// MN-major operand, 128-byte swizzle: rows are K indices (128 bytes = 64 MN elements each), 8-row groups SBO apart, the next 64 MN
// elements (panel) LBO apart  (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t make_mnmajor_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

struct alignas(64) TConvArgs {
  CUtensorMap tmA[4];  // activations {C, W, H, B}, box {64, TW, TH, NB}; stride 2: one map per (row parity, column parity) of the input
  CUtensorMap tmB;     // weights {K = taps * cin, cout}, box {64, BN}
  int tiles_w, tiles_h, tiles_b;
  int TW, TH, NB, H, W, B;   // H, W: the tile grid (output pixels of this launch)
  int cin, cout, taps;
  // tap t reads input map tap_map[t] at (w + tap_dw[t], h + tap_dh[t]) and the weight K block tap_wblk[t]
  int8_t tap_dh[9], tap_dw[9], tap_map[9], tap_wblk[9];
  // output pixel of grid point (h, w): (h * os + ooh, w * os + oow) of an out_H x out_W tensor (stride-2 backward-data writes one parity class per launch)
  int out_H, out_W, os, ooh, oow;
  int stages;  // ring depth = min(4, K iterations): short-K 1x1 layers take less shared memory, so more CTAs share an SM
  __nv_bfloat16* out;
  const __nv_bfloat16* y_prev;  // SiLU' epilogue: conv output of the layer that produced this conv's input, and its BN scale / shift
  const float* s_prev;
  const float* t_prev;
};

// D[128 pixels, BN couts] = sum over (tap, 64-channel chunk) of A[pixels, 64] * W[couts, 64]^T, bf16 x bf16 -> fp32 in TMEM.
// SILU_BWD: the result is d(loss)/d(a_prev) with a_prev = silu(z_prev), z_prev = y_prev * s + t: the epilogue multiplies by silu'(z_prev),
// so the producer's activation gradient never makes a round trip through HBM.
template <int BN, bool SILU_BWD>
__global__ void __launch_bounds__(kTThreads, 1) tconv_kernel(const __grid_constant__ TConvArgs a) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t IDESC = make_idesc_bf16(128, BN, 0, 0);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int STAGES = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTStagesFwd;
  uint64_t* tfull = empty + kTStagesFwd;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  float* st_s = reinterpret_cast<float*>(tmem_slot + 2);  // [BN] scale, [BN] shift of the producer's BN (SILU_BWD)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA[0]);
    tma_prefetch_desc(&a.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  const int n0 = blockIdx.y * BN;
  if (SILU_BWD) {
    for (int i = threadIdx.x; i < BN; i += kTThreads) {
      st_s[i] = a.s_prev[n0 + i];
      st_s[BN + i] = a.t_prev[n0 + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mt = blockIdx.x;
  const int wt = mt % a.tiles_w, t2 = mt / a.tiles_w;
  const int ht = t2 % a.tiles_h, bt = t2 / a.tiles_h;
  const int w0 = wt * a.TW, h0 = ht * a.TH, b0 = bt * a.NB;
  const int chunks = a.cin / 64;
  const int k_iters = a.taps * chunks;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tap = 0; tap < a.taps; ++tap) {
        const int dh = a.tap_dh[tap], dw = a.tap_dw[tap];
        const CUtensorMap* mapA = &a.tmA[a.tap_map[tap]];
        const int kb = a.tap_wblk[tap] * a.cin;
        for (int ck = 0; ck < chunks; ++ck) {
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
          uint8_t* sb = smem + stage * STAGE;
          tma_load_4d(mapA, &full[stage], sb, ck * 64, w0 + dw, h0 + dh, b0);
          tma_load_2d(&a.tmB, &full[stage], sb + A_BYTES, kb + ck * 64, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < k_iters; ++i) {
        mbar_wait(&full[stage], phase, 300 + stage);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE);
        const uint64_t da = make_kmajor_desc<128>(sa);
        const uint64_t db = make_kmajor_desc<128>(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, IDESC, (i | k) ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row % a.TW, r2 = row / a.TW;
    const int th = r2 % a.TH, nb = r2 / a.TH;
    const int ow = w0 + tw, oh = h0 + th, ob = b0 + nb;
    const bool valid = ow < a.W && oh < a.H && ob < a.B;
    const size_t pix = ((size_t)ob * a.out_H + (oh * a.os + a.ooh)) * a.out_W + (ow * a.os + a.oow);
    mbar_wait(tfull, 0, 400);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
        uint4 o[4];
        __nv_bfloat162* ob2 = reinterpret_cast<__nv_bfloat162*>(o);
        if (SILU_BWD) {
          const uint4* yp = reinterpret_cast<const uint4*>(a.y_prev + pix * a.cout + n0 + c);
          uint4 yv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) yv[j] = __ldg(yp + j);
          const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(yv);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 yf = __bfloat1622float2(y2[j]);
            const float z0 = fmaf(yf.x, st_s[c + 2 * j], st_s[BN + c + 2 * j]);
            const float z1 = fmaf(yf.y, st_s[c + 2 * j + 1], st_s[BN + c + 2 * j + 1]);
            const float g0 = 1.0f / (1.0f + __expf(-z0)), g1 = 1.0f / (1.0f + __expf(-z1));
            // silu'(z) = sigmoid(z) * (1 + z * (1 - sigmoid(z)))
            const float d0 = __uint_as_float(v[2 * j]) * (g0 * fmaf(z0, 1.0f - g0, 1.0f));
            const float d1 = __uint_as_float(v[2 * j + 1]) * (g1 * fmaf(z1, 1.0f - g1, 1.0f));
            ob2[j] = __floats2bfloat162_rn(d0, d1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) ob2[j] = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        }
        uint4* dst = reinterpret_cast<uint4*>(a.out + pix * a.cout + n0 + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = o[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

struct alignas(64) TWgradArgs {
  CUtensorMap tmX[4];  // x  {cin, W, H, B},  box {64, TW, TH, NB}; stride 2: one map per input parity class
  CUtensorMap tmD;     // dy {cout, Wo, Ho, B}, box {64, TW, TH, NB}
  int tiles_w, tiles_h, tiles_b;  // over the OUTPUT pixels (dy)
  int TW, TH, NB;
  int cin, cout, taps;
  int nblocks;  // cout / (64 * NPAN): output-channel blocks of one launch (cout = 512 runs as two 256-wide blocks), folded into blockIdx.z
  int8_t tap_dh[9], tap_dw[9], tap_map[9];
  int splits;
  float* dw;  // [cout][taps][cin] fp32, accumulated with atomics (zeroed by the caller)
};

// dW^T block [128 cin rows, cout columns] of one filter tap += sum over this CTA's pixel tiles of x_tile^T * dy_tile  (K = 128 pixels per
// stage, 8 MMAs of K = 16).  Both operands are the NHWC tiles exactly as TMA lands them: MN-major (channels contiguous), 64-channel panels.
template <int NPAN>  // cout / 64
__global__ void __launch_bounds__(kTThreads, 1) twgrad_kernel(const __grid_constant__ TWgradArgs a) {
  constexpr int PANEL = 128 * 128;  // 128 pixel rows x 64 channels bf16
  constexpr int X_BYTES = 2 * PANEL;
  constexpr int D_BYTES = NPAN * PANEL;
  constexpr int STAGE = X_BYTES + D_BYTES;
  constexpr int N = NPAN * 64;
  constexpr int STAGES = NPAN == 4 ? 2 : kTStagesWg;  // 96 KB stages at cout = 256
  constexpr uint32_t IDESC = make_idesc_bf16(128, N, 1, 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmX[0]);
    tma_prefetch_desc(&a.tmD);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, N < 32 ? 32 : N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int split = blockIdx.x, tap = blockIdx.y, c0 = ((int)blockIdx.z / a.nblocks) * 128, n0 = ((int)blockIdx.z % a.nblocks) * N;
  const int m_pan = (a.cin - c0) >= 128 ? 2 : 1;  // valid 64-channel panels of this cin block
  const int dh = a.tap_dh[tap], dw_ = a.tap_dw[tap];
  const CUtensorMap* mapX = &a.tmX[a.tap_map[tap]];
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_b;
  const int per = (m_tiles + a.splits - 1) / a.splits;
  const int t_begin = split * per, t_end = min(m_tiles, t_begin + per);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = t_begin; mt < t_end; ++mt) {
        const int wt = mt % a.tiles_w, t2 = mt / a.tiles_w;
        const int ht = t2 % a.tiles_h, bt = t2 / a.tiles_h;
        const int w0 = wt * a.TW, h0 = ht * a.TH, b0 = bt * a.NB;
        mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
        mbar_expect_tx(&full[stage], (uint32_t)(m_pan * PANEL + D_BYTES));
        uint8_t* sb = smem + stage * STAGE;
        for (int p = 0; p < m_pan; ++p) tma_load_4d(mapX, &full[stage], sb + p * PANEL, c0 + p * 64, w0 + dw_, h0 + dh, b0);
        for (int p = 0; p < NPAN; ++p) tma_load_4d(&a.tmD, &full[stage], sb + X_BYTES + p * PANEL, n0 + p * 64, w0, h0, b0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int mt = t_begin; mt < t_end; ++mt) {
        mbar_wait(&full[stage], phase, 300 + stage);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // K = 16 pixel rows = 2048 bytes per MMA
          const uint64_t da = make_mnmajor_desc_sw128(sa + k * 2048, PANEL, 1024);
          const uint64_t db = make_mnmajor_desc_sw128(sa + X_BYTES + k * 2048, PANEL, 1024);
          umma_f16(tmem_base, da, db, IDESC, (first && k == 0) ? 0u : 1u);
        }
        first = false;
        umma_commit(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else if (t_begin < t_end) {
    const int q = warp & 3;
    const int row = q * 32 + lane;  // cin index inside the block
    const bool valid = row < m_pan * 64 && c0 + row < a.cin;
    mbar_wait(tfull, 0, 400);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j)  // lanes = consecutive cin: one coalesced 128-byte reduction per (cout, tap)
          atomicAdd(a.dw + ((size_t)(n0 + c + j) * a.taps + tap) * a.cin + c0 + row, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, N < 32 ? 32 : N);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm + SiLU element-wise passes
// NHWC bf16, C a multiple of 8.  A thread owns one 8-channel vector column and walks pixels; per-channel partial sums are combined in
// shared memory, then one atomicAdd per channel and CTA.
constexpr int kEwThreads = 256;

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// Per-channel reductions over pixels: a CTA owns a contiguous pixel range, thread (v, pl) = (8-channel vector, pixel lane) accumulates in
// registers, the pixel lanes are combined through shared memory ([lane][2C] partials, no atomics) and 2C global atomicAdds leave the CTA.
constexpr int kRedSmemFloats = kEwThreads * 16;  // np * 2C == 256 * 16 for every C

__device__ __forceinline__ void red_finish(float (&a1)[8], float (&a2)[8], int v, int pl, int np, bool active, int C, float* s_part, float* __restrict__ sums) {
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s_part[pl * 2 * C + v * 8 + i] = a1[i];
      s_part[pl * 2 * C + C + v * 8 + i] = a2[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kEwThreads) {
    float t = 0.0f;
    for (int l = 0; l < np; ++l) t += s_part[l * 2 * C + c];
    atomicAdd(&sums[c], t);
  }
}

// sums[c] += sum_p y[p][c], sums[C + c] += sum_p y[p][c]^2
__global__ void __launch_bounds__(kEwThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long npix, int C, float* __restrict__ sums) {
  __shared__ float s_part[kRedSmemFloats];
  const int cv = C / 8;
  const int v = threadIdx.x % cv;
  const int pl = threadIdx.x / cv, np = kEwThreads / cv;
  const bool active = pl < np;
  const long long chunk = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(npix, p0 + chunk);
  float a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a1[i] = a2[i] = 0.0f;
  if (active) {
    long long p = p0 + pl;
    for (; p + np < p1; p += 2 * np) {  // two independent 16-byte loads in flight
      float f[8], h[8];
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(y + p * C) + v);
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(y + (p + np) * C) + v);
      bf16x8_to_float(u0, f);
      bf16x8_to_float(u1, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a1[i] += f[i] + h[i];
        a2[i] = fmaf(f[i], f[i], fmaf(h[i], h[i], a2[i]));
      }
    }
    if (p < p1) {
      float f[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(y + p * C) + v), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a1[i] += f[i];
        a2[i] = fmaf(f[i], f[i], a2[i]);
      }
    }
  }
  red_finish(a1, a2, v, pl, np, active, C, s_part, sums);
}

// batch statistics -> mean, rstd, folded scale / shift; running statistics updated like nn.BatchNorm2d (unbiased variance, momentum)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, long long npix, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ stat /* [4][C]: mean, rstd, scale, shift */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)npix;
  const double mean = (double)sums[c] / n;
  double var = (double)sums[C + c] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float s = gamma[c] * rstd;
  stat[c] = (float)mean;
  stat[C + c] = rstd;
  stat[2 * C + c] = s;
  stat[3 * C + c] = beta[c] - (float)mean * s;
  if (running_mean != nullptr) {
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * n / (n > 1.0 ? n - 1.0 : 1.0));
  }
}

// a = silu(y * scale + shift)
__global__ void __launch_bounds__(kEwThreads) bn_silu_fwd_kernel(const __nv_bfloat16* __restrict__ y, long long nvec, int C, const float* __restrict__ stat,
                                                                 __nv_bfloat16* __restrict__ out) {
  const int cv = C / 8;
  const long long stride = (long long)gridDim.x * kEwThreads;
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < nvec; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nvec;
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(y) + i);
    const uint4 u1 = two ? __ldg(reinterpret_cast<const uint4*>(y) + i2) : u0;  // two independent 16-byte loads in flight
    {
      const int c0 = (int)(i % cv) * 8;
      float f[8];
      bf16x8_to_float(u0, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(f[k], __ldg(stat + 2 * C + c0 + k), __ldg(stat + 3 * C + c0 + k));
        f[k] = z / (1.0f + __expf(-z));
      }
      reinterpret_cast<uint4*>(out)[i] = float_to_bf16x8(f);
    }
    if (two) {
      const int c0 = (int)(i2 % cv) * 8;
      float f[8];
      bf16x8_to_float(u1, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(f[k], __ldg(stat + 2 * C + c0 + k), __ldg(stat + 3 * C + c0 + k));
        f[k] = z / (1.0f + __expf(-z));
      }
      reinterpret_cast<uint4*>(out)[i2] = float_to_bf16x8(f);
    }
  }
}

__device__ __forceinline__ float silu_grad(float z) {
  const float g = 1.0f / (1.0f + __expf(-z));
  return g * fmaf(z, 1.0f - g, 1.0f);
}

// sums[c] += sum_p dz, sums[C + c] += sum_p dz * xhat   (dz = g * silu'(z) when `g` is d(loss)/d(a); dz = g when g_is_dz)
__global__ void __launch_bounds__(kEwThreads) bn_silu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y, long long npix,
                                                                        int C, const float* __restrict__ stat, int g_is_dz, float* __restrict__ sums) {
  __shared__ float s_part[kRedSmemFloats];
  const int cv = C / 8;
  const int v = threadIdx.x % cv;
  const int pl = threadIdx.x / cv, np = kEwThreads / cv;
  const bool active = pl < np;
  const long long chunk = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(npix, p0 + chunk);
  float a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a1[i] = a2[i] = 0.0f;
  if (active) {
    float mean[8], rstd[8], sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = stat[v * 8 + i];
      rstd[i] = stat[C + v * 8 + i];
      sc[i] = stat[2 * C + v * 8 + i];
      sh[i] = stat[3 * C + v * 8 + i];
    }
    for (long long p = p0 + pl; p < p1; p += np) {
      float fy[8], fg[8];
      const uint4 uy = __ldg(reinterpret_cast<const uint4*>(y + p * C) + v);
      const uint4 ug = __ldg(reinterpret_cast<const uint4*>(g + p * C) + v);
      bf16x8_to_float(uy, fy);
      bf16x8_to_float(ug, fg);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dz = g_is_dz ? fg[i] : fg[i] * silu_grad(fmaf(fy[i], sc[i], sh[i]));
        a1[i] += dz;
        a2[i] = fmaf(dz, (fy[i] - mean[i]) * rstd[i], a2[i]);
      }
    }
  }
  red_finish(a1, a2, v, pl, np, active, C, s_part, sums);
}

// dy = gamma * rstd * (dz - sum_dz / N - xhat * sum_dzx / N)
__global__ void __launch_bounds__(kEwThreads) bn_silu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y, long long nvec, int C,
                                                                       const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ sums,
                                                                       float inv_n, int g_is_dz, __nv_bfloat16* __restrict__ dy) {
  const int cv = C / 8;
  const long long stride = (long long)gridDim.x * kEwThreads;
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < nvec; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nvec;
    // four independent 16-byte loads in flight
    const uint4 uy0 = __ldg(reinterpret_cast<const uint4*>(y) + i), ug0 = __ldg(reinterpret_cast<const uint4*>(g) + i);
    const uint4 uy1 = two ? __ldg(reinterpret_cast<const uint4*>(y) + i2) : uy0, ug1 = two ? __ldg(reinterpret_cast<const uint4*>(g) + i2) : ug0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long ii = h ? i2 : i;
      const int c0 = (int)(ii % cv) * 8;
      float fy[8], fg[8];
      bf16x8_to_float(h ? uy1 : uy0, fy);
      bf16x8_to_float(h ? ug1 : ug0, fg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        const float mean = __ldg(stat + c), rstd = __ldg(stat + C + c);
        const float dz = g_is_dz ? fg[k] : fg[k] * silu_grad(fmaf(fy[k], __ldg(stat + 2 * C + c), __ldg(stat + 3 * C + c)));
        const float xhat = (fy[k] - mean) * rstd;
        fg[k] = __ldg(gamma + c) * rstd * (dz - __ldg(sums + c) * inv_n - xhat * __ldg(sums + C + c) * inv_n);
      }
      reinterpret_cast<uint4*>(dy)[ii] = float_to_bf16x8(fg);
    }
  }
}

// fp32 master weights [cout][cin][kh][kw] (nn.Conv2d) -> bf16 [cout][kh*kw][cin] (forward / wgrad layout) and the backward-data operand
// bf16 [cin][kh*kw][cout] with the taps rotated by 180 degrees
__global__ void pack_train_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wb) {
  const long long n = (long long)cout * cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int ci = (int)((i / taps) % cin);
    const int co = (int)(i / ((long long)taps * cin));
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    wf[((size_t)co * taps + t) * cin + ci] = v;
    wb[((size_t)ci * taps + (taps - 1 - t)) * cout + co] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host
static int encode_tiled(CUtensorMap* m, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(CVB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(CVB_ERR_CUDA, "cuTensorMapEncodeTiled (training) failed: %d", (int)r);
  return CVB_OK;
}

// pixel box TW x TH x NB with exactly 128 pixels (powers of two) that wastes the fewest rows on out-of-bounds pixels
static void choose_box128(int B, int H, int W, int* TW, int* TH, int* NB) {
  double best = -1.0;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; tw * th <= 128; th *= 2) {
      const int nb = 128 / (tw * th);
      const double tiles = (double)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(B, nb);
      const double util = (double)B * H * W / (tiles * 128.0) + 1e-6 * tw;
      if (util > best) {
        best = util;
        *TW = tw;
        *TH = th;
        *NB = nb;
      }
    }
}

static int act_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int TW, int TH, int NB) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  const cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)NB};
  return encode_tiled(m, 4, base, dims, str, box);
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_train_pack_weights(const float* w, int32_t cout, int32_t cin, int32_t k, void* w_fwd, void* w_bwd, void* stream) {
  CVB_REQUIRE(w && w_fwd && w_bwd && cout > 0 && cin > 0 && (k == 1 || k == 3), "train_pack_weights: bad argument");
  const long long n = (long long)cout * cin * k * k;
  pack_train_weights_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, as_stream(stream)>>>(
      w, cout, cin, k * k, static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_bwd));
  CVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CVB_OK;
}

// tensor map of one parity class (py, px) of an NHWC tensor: pixels (2i + py, 2j + px); (0, 0) with step 1 is the plain map
static int act_map_strided(CUtensorMap* m, const void* base, int B, int H, int W, int C, int step, int py, int px, int TW, int TH, int NB) {
  const int Wp = (W - px + step - 1) / step, Hp = (H - py + step - 1) / step;
  if (Wp <= 0 || Hp <= 0) return act_map(m, base, B, H, W, C, TW, TH, NB);  // never addressed (no pixel of this parity): any valid map
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)B};
  const cuuint64_t str[3] = {(cuuint64_t)C * 2 * step, (cuuint64_t)C * 2 * W * step, (cuuint64_t)C * 2 * W * H};
  const cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)NB};
  return encode_tiled(m, 4, static_cast<const uint8_t*>(base) + ((size_t)py * W + px) * C * 2, dims, str, box);
}

// stride-2 3x3 / pad 1: tap (ky, kx) reads input pixel (2h + ky - 1, 2w + kx - 1) = parity class ((ky + 1) & 1, (kx + 1) & 1) at index
// (h + (ky == 0 ? -1 : 0), w + (kx == 0 ? -1 : 0)) of that class
static void s2_taps(int8_t* dh, int8_t* dw, int8_t* map) {
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int t = ky * 3 + kx;
      dh[t] = (int8_t)(ky == 0 ? -1 : 0);
      dw[t] = (int8_t)(kx == 0 ? -1 : 0);
      map[t] = (int8_t)((((ky + 1) & 1) << 1) | ((kx + 1) & 1));
    }
}

static int launch_tconv(TConvArgs& a, const void* w_packed, int wblocks, const void* y_prev, const float* bn_stat_prev, void* stream) {
  choose_box128(a.B, a.H, a.W, &a.TW, &a.TH, &a.NB);
  a.tiles_w = ceil_div(a.W, a.TW);
  a.tiles_h = ceil_div(a.H, a.TH);
  a.tiles_b = ceil_div(a.B, a.NB);
  a.y_prev = static_cast<const __nv_bfloat16*>(y_prev);
  a.s_prev = bn_stat_prev ? bn_stat_prev + 2 * (size_t)a.cout : nullptr;
  a.t_prev = bn_stat_prev ? bn_stat_prev + 3 * (size_t)a.cout : nullptr;
  const int bn = a.cout % 128 == 0 ? 128 : 64;
  {
    const long long K = (long long)wblocks * a.cin;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)a.cout};
    const cuuint64_t str[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)bn};
    const int rc = encode_tiled(&a.tmB, 2, w_packed, dims, str, box);
    if (rc != CVB_OK) return rc;
  }
  const void* fn;
  if (bn == 128) fn = y_prev ? reinterpret_cast<const void*>(&tconv_kernel<128, true>) : reinterpret_cast<const void*>(&tconv_kernel<128, false>);
  else fn = y_prev ? reinterpret_cast<const void*>(&tconv_kernel<64, true>) : reinterpret_cast<const void*>(&tconv_kernel<64, false>);
  const int k_iters = a.taps * (a.cin / 64);
  a.stages = k_iters < kTStagesFwd ? k_iters : kTStagesFwd;
  const int smem = a.stages * (128 * 128 + bn * 128) + 256 + 2 * bn * 4;
  CVB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* kargs[1] = {&a};
  const dim3 grid((unsigned)(a.tiles_w * a.tiles_h * a.tiles_b), (unsigned)(a.cout / bn));
  CVB_CHECK_CUDA(cudaLaunchKernel(fn, grid, dim3(kTThreads), kargs, (size_t)smem, as_stream(stream)));
  count_launch();
  return CVB_OK;
}

extern "C" int cvb_train_conv(const void* x, int32_t B, int32_t H, int32_t W, int32_t cin, const void* w_packed, int32_t cout, int32_t k, int32_t stride, void* out,
                              const void* y_prev, const float* bn_stat_prev, void* stream) {
  CVB_REQUIRE(x && w_packed && out, "train_conv: null tensor");
  CVB_REQUIRE((k == 1 || k == 3) && cin % 64 == 0 && cout % 64 == 0 && B > 0 && H > 0 && W > 0, "train_conv: k in {1,3}, channels multiples of 64 (cin=%d cout=%d k=%d)", cin, cout, k);
  CVB_REQUIRE(stride == 1 || (stride == 2 && k == 3), "train_conv: stride 1, or stride 2 with k = 3");
  CVB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "train_conv: pointers must be 16-byte aligned");
  CVB_REQUIRE((y_prev == nullptr) == (bn_stat_prev == nullptr), "train_conv: the SiLU' epilogue needs both y_prev and its BN statistics");
  TConvArgs a;
  memset(&a, 0, sizeof(a));
  const int Ho = (H + 2 * (k / 2) - k) / stride + 1, Wo = (W + 2 * (k / 2) - k) / stride + 1;
  a.H = Ho;
  a.W = Wo;
  a.B = B;
  a.cin = cin;
  a.cout = cout;
  a.taps = k * k;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out_H = Ho;
  a.out_W = Wo;
  a.os = 1;
  choose_box128(B, Ho, Wo, &a.TW, &a.TH, &a.NB);
  int rc = CVB_OK;
  if (stride == 1) {
    for (int t = 0; t < k * k; ++t) {
      a.tap_dh[t] = (int8_t)(t / k - k / 2);
      a.tap_dw[t] = (int8_t)(t % k - k / 2);
      a.tap_map[t] = 0;
      a.tap_wblk[t] = (int8_t)t;
    }
    rc = act_map(&a.tmA[0], x, B, H, W, cin, a.TW, a.TH, a.NB);
  } else {
    s2_taps(a.tap_dh, a.tap_dw, a.tap_map);
    for (int t = 0; t < 9; ++t) a.tap_wblk[t] = (int8_t)t;
    for (int py = 0; py < 2 && rc == CVB_OK; ++py)
      for (int px = 0; px < 2 && rc == CVB_OK; ++px) rc = act_map_strided(&a.tmA[py * 2 + px], x, B, H, W, cin, 2, py, px, a.TW, a.TH, a.NB);
  }
  if (rc != CVB_OK) return rc;
  return launch_tconv(a, w_packed, k * k, y_prev, bn_stat_prev, stream);
}

// Backward-data of the 3x3 / stride 2 / pad 1 convolution: dx[2a + pi, 2b + pj] gets dy[a + dh, b + dw] through the taps whose parity matches
// (pi = 0: ky = 1; pi = 1: ky = 0 (dh = +1) and ky = 2 (dh = 0); same for columns) -- four dense stride-1 sub-convolutions over dy, one per
// parity class of dx, each writing its class with a strided store.  w_bwd = the [cin][9][cout] operand of cvb_train_pack_weights (taps rotated).
extern "C" int cvb_train_conv_dgrad_s2(const void* dy, int32_t B, int32_t Ho, int32_t Wo, int32_t cout, const void* w_bwd, int32_t cin, int32_t H, int32_t W, void* dx,
                                       const void* y_prev, const float* bn_stat_prev, void* stream) {
  CVB_REQUIRE(dy && w_bwd && dx, "train_conv_dgrad_s2: null tensor");
  CVB_REQUIRE(cin % 64 == 0 && cout % 64 == 0 && B > 0 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1, "train_conv_dgrad_s2: bad geometry");
  CVB_REQUIRE((y_prev == nullptr) == (bn_stat_prev == nullptr), "train_conv_dgrad_s2: the SiLU' epilogue needs both y_prev and its BN statistics");
  for (int pi = 0; pi < 2; ++pi)
    for (int pj = 0; pj < 2; ++pj) {
      const int Hs = (H - pi + 1) / 2, Ws = (W - pj + 1) / 2;  // pixels of this parity class
      if (Hs <= 0 || Ws <= 0) continue;
      TConvArgs a;
      memset(&a, 0, sizeof(a));
      a.H = Hs;
      a.W = Ws;
      a.B = B;
      a.cin = cout;  // the GEMM's K: dy channels
      a.cout = cin;  // the GEMM's N: dx channels
      a.out = static_cast<__nv_bfloat16*>(dx);
      a.out_H = H;
      a.out_W = W;
      a.os = 2;
      a.ooh = pi;
      a.oow = pj;
      int nt = 0;
      for (int ky = 0; ky < 3; ++ky) {
        if (((ky + 1) & 1) != pi) continue;  // rows of parity pi are reached by ky with 2*oh + ky - 1 = ih
        for (int kx = 0; kx < 3; ++kx) {
          if (((kx + 1) & 1) != pj) continue;
          a.tap_dh[nt] = (int8_t)(ky == 0 ? 1 : 0);
          a.tap_dw[nt] = (int8_t)(kx == 0 ? 1 : 0);
          a.tap_map[nt] = 0;
          a.tap_wblk[nt] = (int8_t)(8 - (ky * 3 + kx));  // w_bwd stores tap t of the forward filter at block 8 - t
          ++nt;
        }
      }
      a.taps = nt;
      choose_box128(B, Hs, Ws, &a.TW, &a.TH, &a.NB);
      int rc = act_map(&a.tmA[0], dy, B, Ho, Wo, cout, a.TW, a.TH, a.NB);
      if (rc != CVB_OK) return rc;
      rc = launch_tconv(a, w_bwd, 9, y_prev, bn_stat_prev, stream);
      if (rc != CVB_OK) return rc;
    }
  return CVB_OK;
}

extern "C" int cvb_train_conv_wgrad(const void* x, const void* dy, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t k, int32_t stride, float* dw,
                                    void* stream) {
  CVB_REQUIRE(x && dy && dw, "train_conv_wgrad: null tensor");
  CVB_REQUIRE((k == 1 || k == 3) && cin % 64 == 0 && (cout == 64 || cout == 128 || cout % 256 == 0), "train_conv_wgrad: k in {1,3}, cin %% 64 == 0, cout in {64, 128, multiples of 256}");
  CVB_REQUIRE(stride == 1 || (stride == 2 && k == 3), "train_conv_wgrad: stride 1, or stride 2 with k = 3");
  CVB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "train_conv_wgrad: pointers must be 16-byte aligned");
  TWgradArgs a;
  memset(&a, 0, sizeof(a));
  const int Ho = (H + 2 * (k / 2) - k) / stride + 1, Wo = (W + 2 * (k / 2) - k) / stride + 1;
  choose_box128(B, Ho, Wo, &a.TW, &a.TH, &a.NB);
  a.tiles_w = ceil_div(Wo, a.TW);
  a.tiles_h = ceil_div(Ho, a.TH);
  a.tiles_b = ceil_div(B, a.NB);
  a.cin = cin;
  a.cout = cout;
  a.taps = k * k;
  a.dw = dw;
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_b;
  const int cblocks = ceil_div(cin, 128);
  const int npan = cout >= 256 ? 4 : cout / 64;
  a.nblocks = cout / (64 * npan);
  int splits = (2 * num_sms()) / (a.taps * cblocks * a.nblocks);  // ~two waves of CTAs: (split, tap, cin block x cout block)
  if (splits < 1) splits = 1;
  if (splits > m_tiles) splits = m_tiles;
  a.splits = splits;
  int rc = CVB_OK;
  if (stride == 1) {
    for (int t = 0; t < k * k; ++t) {
      a.tap_dh[t] = (int8_t)(t / k - k / 2);
      a.tap_dw[t] = (int8_t)(t % k - k / 2);
      a.tap_map[t] = 0;
    }
    rc = act_map(&a.tmX[0], x, B, H, W, cin, a.TW, a.TH, a.NB);
  } else {
    s2_taps(a.tap_dh, a.tap_dw, a.tap_map);
    for (int py = 0; py < 2 && rc == CVB_OK; ++py)
      for (int px = 0; px < 2 && rc == CVB_OK; ++px) rc = act_map_strided(&a.tmX[py * 2 + px], x, B, H, W, cin, 2, py, px, a.TW, a.TH, a.NB);
  }
  if (rc == CVB_OK) rc = act_map(&a.tmD, dy, B, Ho, Wo, cout, a.TW, a.TH, a.NB);
  if (rc != CVB_OK) return rc;
  const void* fn = npan == 1 ? reinterpret_cast<const void*>(&twgrad_kernel<1>) : (npan == 2 ? reinterpret_cast<const void*>(&twgrad_kernel<2>) : reinterpret_cast<const void*>(&twgrad_kernel<4>));
  const int stages = npan == 4 ? 2 : kTStagesWg;
  const int smem = stages * ((2 + npan) * 128 * 128) + 256;
  CVB_REQUIRE(smem <= 232448, "train_conv_wgrad: cout=%d needs %d bytes of shared memory", cout, smem);
  CVB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* kargs[1] = {&a};
  const dim3 grid((unsigned)splits, (unsigned)a.taps, (unsigned)(cblocks * a.nblocks));
  CVB_CHECK_CUDA(cudaLaunchKernel(fn, grid, dim3(kTThreads), kargs, (size_t)smem, as_stream(stream)));
  count_launch();
  return CVB_OK;
}

// reduction kernels: `np` pixel lanes per CTA, at least ~4 pixels per thread, at most four CTAs per SM (2C global atomics per CTA)
static int red_grid(long long npix, int np) {
  long long g = (npix + (long long)np * 4 - 1) / ((long long)np * 4);
  const long long cap = (long long)num_sms() * 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int ew_grid(long long work_items) {
  long long g = (work_items + kEwThreads - 1) / kEwThreads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int cvb_train_bn_stats(const void* y, int64_t npix, int32_t C, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                  float* running_var, float* sums_scratch, float* stat, void* stream) {
  CVB_REQUIRE(y && gamma && beta && sums_scratch && stat && npix > 0 && C % 8 == 0 && C <= 2048, "train_bn_stats: bad argument");
  cudaStream_t st = as_stream(stream);
  CVB_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * (size_t)C * sizeof(float), st));
  const int np = kEwThreads / (C / 8) > 0 ? kEwThreads / (C / 8) : 1;
  CVB_REQUIRE(C / 8 <= kEwThreads, "train_bn_stats: C too large");
  bn_stats_kernel<<<red_grid(npix, np), kEwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(y), npix, C, sums_scratch);
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums_scratch, npix, C, gamma, beta, eps, momentum, running_mean, running_var, stat);
  CVB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return CVB_OK;
}

extern "C" int cvb_train_bn_silu_fwd(const void* y, int64_t npix, int32_t C, const float* stat, void* out, void* stream) {
  CVB_REQUIRE(y && stat && out && C % 8 == 0, "train_bn_silu_fwd: bad argument");
  const long long nvec = npix * (C / 8);
  bn_silu_fwd_kernel<<<ew_grid(nvec), kEwThreads, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y), nvec, C, stat, static_cast<__nv_bfloat16*>(out));
  CVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CVB_OK;
}

extern "C" int cvb_train_bn_silu_bwd(const void* g, int32_t g_is_dz, const void* y, int64_t npix, int32_t C, const float* stat, const float* gamma, float* sums /* [2][C]: out = dbeta, dgamma */,
                                     void* dy, void* stream) {
  CVB_REQUIRE(g && y && stat && gamma && sums && dy && C % 8 == 0 && C / 8 <= kEwThreads, "train_bn_silu_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  CVB_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  const int np = kEwThreads / (C / 8);
  bn_silu_bwd_reduce_kernel<<<red_grid(npix, np), kEwThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(y), npix, C, stat, g_is_dz, sums);
  const long long nvec = npix * (C / 8);
  bn_silu_bwd_apply_kernel<<<ew_grid(nvec), kEwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(y), nvec, C, stat, gamma, sums,
                                                                 (float)(1.0 / (double)npix), g_is_dz, static_cast<__nv_bfloat16*>(dy));
  CVB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return CVB_OK;
}
