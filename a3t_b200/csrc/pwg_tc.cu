// ParallelWaveGAN gated residual block on the tensor cores (tcgen05 / TMEM / TMA), one launch per block.
// Arithmetic of espnet2/gan_tts/wavenet/residual_block.py:114-169 as parallel_wavegan.py:136-173 chains it:
//   h = Conv1d_{64->128, k=3, dilation d}(x) + Conv1x1_{80->128}(c) + b1;   g = tanh(h[:64]) * sigmoid(h[64:])
//   o = Conv1x1_{64->128}(g) + b2;   x <- (o[:64] + x) * sqrt(0.5);   skip += o[64:]
//
// The waveform gate is fp32 in the reference and the parity gate is atol 1e-4 on the waveform after 30 blocks, so the
// two contractions run as SPLIT-fp16 GEMMs: every operand is stored as hi + lo fp16 planes (22 mantissa bits) and
// each product is three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): ~2^-21 relative error per product
// at three times the bf16 cost, still ~30x the CUDA-core kernel this replaces (pwg.cu).
//
// Orientation: M = 128 consecutive samples of one utterance, N = 128 channels.
//   MMA 1: D1[sample, ch] = sum_k X[sample, k] W1[ch, k],  K = 3 taps x 64 channels + 80 aux (+48 zero) = 320:
//          A = activation planes stored CHANNELS-LAST (B, T, 64): a 128-sample x 64-channel tile is 16 KB of
//          consecutive 128-byte rows, K-major, straight from the planes by TMA; the tap shift (+-d samples) is the
//          TMA row coordinate (any d: rows are 128-byte aligned) and its zero fill is the conv's zero padding;
//          B = W1 (K-major, stored chunk-major so that every 128 x 64 chunk is contiguous).  Five 64-wide K chunks
//          through a 2-stage ring (W1 is re-streamed from L2 per tile).
//   gate : thread = sample reads its h row from TMEM (columns c and c + 64 in the same thread), writes g as the
//          K-major A operand of MMA 2 (one 128-byte swizzled row per sample and plane).
//   MMA 2: D2[sample, ch] = sum_k g[sample, k] W2[ch, k], K = 64, W2 resident in shared memory.
//   epilogue: thread = sample: its row of the channels-last x planes is one contiguous 64-byte run per plane (16-byte
//          vector loads / stores); skip stays channel-major fp32, lanes = consecutive samples => coalesced.
// TMEM: D1 and D2 double buffered (4 x 128 columns).  Warps: 8 epilogue (2 per TMEM lane quarter), TMA producer, MMA.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace a3t {
namespace pwgtc {
using namespace tc;

constexpr int EPI_WARPS = 8;
// two MMA issuers: MMA 1 of the next tile waits on TMA data for a whole tile period; MMA 2 of the current tile must not
// queue behind it in one thread's program order
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1, MMA2_WARP = EPI_WARPS + 2;
constexpr int NUM_THREADS = 32 * (EPI_WARPS + 3);
constexpr int KCH = 5;                       // K chunks of 64: taps 0..2, aux 0..63, aux 64..79 (+ zero fill)
constexpr uint32_t TILE16 = 16384;           // one 128 x 64 fp16 operand tile
constexpr uint32_t STAGE = 4 * TILE16;       // A hi, A lo, B hi, B lo
constexpr int SMEM_BYTES = 1024 + 2 * STAGE + 4 * TILE16 + 1024 + 256;

struct Params {
  const __half* xh; const __half* xl;        // input planes (B, T, 64) channels-last
  __half* yh; __half* yl;                    // output planes
  float* skip;                               // (B, 64, T) fp32
  const float* b1; const float* b2;          // (128) each
  int B; int64_t T;
  int dil, first;
  int passes;                                // 3: hi*hi + hi*lo + lo*hi;  2: weights as single fp16 (x_hi + x_lo) * w_hi
};

__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2 / (1 + exp(2x)): ex2.approx (2 ulp) + division; absolute error ~1e-7 (tanh.approx's 5e-4 would not do)
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, 1.f + e);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ void split16(float v, unsigned short& hi, unsigned short& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
pwg_resblock_tc_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                       const __grid_constant__ CUtensorMap tmCh, const __grid_constant__ CUtensorMap tmCl,
                       const __grid_constant__ CUtensorMap tmW1h, const __grid_constant__ CUtensorMap tmW1l,
                       const __grid_constant__ CUtensorMap tmW2h, const __grid_constant__ CUtensorMap tmW2l,
                       const __grid_constant__ Params p) {
  A3T_PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sRing = base, sG = sRing + 2 * STAGE, sW2 = sG + 2 * TILE16, sBias = sW2 + 2 * TILE16, sBar = sBias + 1024;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 16u + 8u * s; };
  auto d1_full = [&](int i) { return sBar + 32u + 8u * i; };
  auto d1_empty = [&](int i) { return sBar + 48u + 8u * i; };
  auto d2_full = [&](int i) { return sBar + 64u + 8u * i; };
  auto d2_empty = [&](int i) { return sBar + 80u + 8u * i; };
  const uint32_t g_full = sBar + 96, g_empty = sBar + 104, w2_full = sBar + 112, tmem_slot = sBar + 120;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (int)((p.T + 127) / 128);
  const int ntiles = p.B * tiles_per_b;

  if (warp == PRODUCER_WARP && lane == 0) {
    const CUtensorMap* maps[8] = {&tmXh, &tmXl, &tmCh, &tmCl, &tmW1h, &tmW1l, &tmW2h, &tmW2l};
    for (int i = 0; i < 8; i++) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)maps[i]) : "memory");
    for (int i = 0; i < 2; i++) {
      mbar_init(full(i), 1); mbar_init(empty(i), 1);
      mbar_init(d1_full(i), 1); mbar_init(d1_empty(i), EPI_WARPS);
      mbar_init(d2_full(i), 1); mbar_init(d2_empty(i), EPI_WARPS);
    }
    mbar_init(g_full, EPI_WARPS); mbar_init(g_empty, 1); mbar_init(w2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 256) {   // biases to shared memory: [0,128) = b1, [128,256) = b2
    const float v = threadIdx.x < 128 ? p.b1[threadIdx.x] : p.b2[threadIdx.x - 128];
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + 4 * threadIdx.x), "f"(v) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  A3T_PDL_WAIT();

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      mbar_expect_tx(w2_full, p.passes >= 3 ? 2 * TILE16 : TILE16);
      tma_load_4d(sW2, &tmW2h, w2_full, 0, 0, 0, 0);
      if (p.passes >= 3) tma_load_4d(sW2 + TILE16, &tmW2l, w2_full, 0, 0, 0, 0);
      int step = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_b;
        const int t0 = (tile - b * tiles_per_b) * 128;
        for (int kc = 0; kc < KCH; kc++, step++) {
          const int s = step & 1;
          if (step >= 2) mbar_wait(empty(s), ((step >> 1) & 1) ^ 1);
          const uint32_t st = sRing + s * STAGE;
          mbar_expect_tx(full(s), p.passes >= 3 ? STAGE : STAGE - TILE16);
          // A: 128 samples (rows) x 64 k (channels of one tap / aux channels), one 16 KB box per plane
          const CUtensorMap* mh = kc < 3 ? &tmXh : &tmCh;
          const CUtensorMap* ml = kc < 3 ? &tmXl : &tmCl;
          const int row0 = kc < 3 ? t0 + (kc - 1) * p.dil : t0;
          const int col0 = kc == 4 ? 64 : 0;
          tma_load_4d(st, mh, full(s), col0, row0, b, 0);
          tma_load_4d(st + TILE16, ml, full(s), col0, row0, b, 0);
          tma_load_4d(st + 2 * TILE16, &tmW1h, full(s), 0, 0, kc, 0);
          if (p.passes >= 3) tma_load_4d(st + 3 * TILE16, &tmW1l, full(s), 0, 0, kc, 0);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      // kind::f16 with fp16 operands (format 0), fp32 accumulate, both operands K-major
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int step = 0;
      auto mma1 = [&](int i) {
        const int bi = i & 1;
        if (i >= 2) mbar_wait(d1_empty(bi), ((i >> 1) & 1) ^ 1);
        const uint32_t d = tmem_base + 128u * bi;
        for (int kc = 0; kc < KCH; kc++, step++) {
          const int s = step & 1;
          mbar_wait(full(s), (step >> 1) & 1);
          tc_fence_after();
          const uint32_t st = sRing + s * STAGE;
          const uint64_t ah = make_smem_desc(st, 16), al = make_smem_desc(st + TILE16, 16);
          const uint64_t bh = make_smem_desc(st + 2 * TILE16, 16), bl = make_smem_desc(st + 3 * TILE16, 16);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t ka = (uint64_t)((k * 32) >> 4), kb = ka;
            umma_bf16(d, ah + ka, bh + kb, idesc1, (kc | k) ? 1u : 0u);
            if (p.passes >= 3) umma_bf16(d, ah + ka, bl + kb, idesc1, 1u);
            umma_bf16(d, al + ka, bh + kb, idesc1, 1u);
          }
          umma_commit(empty(s));
        }
        umma_commit(d1_full(bi));
      };
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) my_tiles++;
      for (int i = 0; i < my_tiles; i++) mma1(i);
    }
  } else if (warp == MMA2_WARP) {
    if (elect_one()) {
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t dGh = make_smem_desc(sG, 16), dGl = make_smem_desc(sG + TILE16, 16);
      const uint64_t dW2h = make_smem_desc(sW2, 16), dW2l = make_smem_desc(sW2 + TILE16, 16);
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) my_tiles++;
      mbar_wait(w2_full, 0);
      for (int i = 0; i < my_tiles; i++) {
        const int bi = i & 1;
        mbar_wait(g_full, i & 1);
        if (i >= 2) mbar_wait(d2_empty(bi), ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + 256u + 128u * bi;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint64_t kk = (uint64_t)((k * 32) >> 4);
          umma_bf16(d, dGh + kk, dW2h + kk, idesc2, k ? 1u : 0u);
          if (p.passes >= 3) umma_bf16(d, dGh + kk, dW2l + kk, idesc2, 1u);
          umma_bf16(d, dGl + kk, dW2h + kk, idesc2, 1u);
        }
        umma_commit(g_empty);
        umma_commit(d2_full(bi));
      }
    }
  } else {
    // ===================================== gate + epilogue warps ==============================
    const int q = warp & 3, hh = warp >> 2;          // TMEM lane quarter, channel half (channels 32 hh .. +31 of each 64-block)
    const int row = q * 32 + lane;
    const uint32_t lane_t = ((uint32_t)(q * 32) << 16);
    const uint32_t sw = (uint32_t)(row & 7);
    const float rs = 0.70710678118654752440f;         // math.sqrt(0.5)
    // software pipeline: gate(i + 1) runs before epilogue(i), so MMA 2 of tile i has a whole gate phase to finish
    auto gate = [&](int i) {
      const int bi = i & 1;
      // ---- gate: g = tanh(h_a + b1_a) * sigmoid(h_b + b1_b) for channels 32 hh .. +31
      mbar_wait(d1_full(bi), (i >> 1) & 1);
      tc_fence_after();
      uint32_t ha[32], hb[32];
      tmem_ld32(tmem_base + 128u * bi + lane_t + 32u * hh, ha);
      tmem_ld32(tmem_base + 128u * bi + lane_t + 64u + 32u * hh, hb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty(bi));
      uint32_t gh[16], gl[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        unsigned short h0, l0, h1, l1;
        float ba0, bb0, ba1, bb1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ba0), "=f"(ba1) : "r"(sBias + 4 * (32 * hh + c)) : "memory");
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bb0), "=f"(bb1) : "r"(sBias + 4 * (64 + 32 * hh + c)) : "memory");
        split16(fast_tanh(__uint_as_float(ha[c]) + ba0) * fast_sigmoid(__uint_as_float(hb[c]) + bb0), h0, l0);
        split16(fast_tanh(__uint_as_float(ha[c + 1]) + ba1) * fast_sigmoid(__uint_as_float(hb[c + 1]) + bb1), h1, l1);
        gh[c >> 1] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        gl[c >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
      }
      if (i > 0) mbar_wait(g_empty, (i - 1) & 1);   // MMA 2 of the previous tile has read the g tiles
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t off = row * 128 + (((4 * hh + u) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sG + off), "r"(gh[4 * u]), "r"(gh[4 * u + 1]),
                     "r"(gh[4 * u + 2]), "r"(gh[4 * u + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sG + TILE16 + off), "r"(gl[4 * u]), "r"(gl[4 * u + 1]),
                     "r"(gl[4 * u + 2]), "r"(gl[4 * u + 3])
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_full);
    };
    auto epilogue = [&](int i, int tile) {
      const int bi = i & 1;
      const int b = tile / tiles_per_b;
      const int64_t t = (int64_t)(tile - b * tiles_per_b) * 128 + row;
      const bool t_ok = t < p.T;
      // ---- residual inputs of my 32 channels (independent of MMA 2: issued before waiting for it): 64 contiguous
      // bytes per plane in the channels-last row of this sample
      const int64_t xoff = ((int64_t)b * p.T + t) * 64 + 32 * hh;
      uint4 xhv[4], xlv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        xhv[u] = t_ok ? __ldg(reinterpret_cast<const uint4*>(p.xh + xoff) + u) : make_uint4(0u, 0u, 0u, 0u);
        xlv[u] = t_ok ? __ldg(reinterpret_cast<const uint4*>(p.xl + xoff) + u) : make_uint4(0u, 0u, 0u, 0u);
      }
      // the skip accumulators of my 32 channels too: all 32 loads in flight before the first store (a load after a
      // store through the same pointer cannot be hoisted by the compiler: 32 serialised DRAM round trips otherwise)
      float* const sk = p.skip + ((int64_t)b * 64 + 32 * hh) * p.T + t;
      float skv[32];
#pragma unroll
      for (int c = 0; c < 32; c++) skv[c] = (t_ok && !p.first) ? __ldcs(sk + (int64_t)c * p.T) : 0.f;
      // ---- epilogue: x <- (o_res + x) sqrt(.5) as hi / lo planes, skip += o_skip
      mbar_wait(d2_full(bi), (i >> 1) & 1);
      tc_fence_after();
      uint32_t orr[32], osk[32];
      tmem_ld32(tmem_base + 256u + 128u * bi + lane_t + 32u * hh, orr);
      tmem_ld32(tmem_base + 256u + 128u * bi + lane_t + 64u + 32u * hh, osk);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(bi));
      if (t_ok) {
        const uint32_t* xhw = reinterpret_cast<const uint32_t*>(xhv);
        const uint32_t* xlw = reinterpret_cast<const uint32_t*>(xlv);
        uint32_t yhw[16], ylw[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float b2r0, b2r1, b2s0, b2s1;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b2r0), "=f"(b2r1) : "r"(sBias + 4 * (128 + 32 * hh + c)) : "memory");
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b2s0), "=f"(b2s1) : "r"(sBias + 4 * (192 + 32 * hh + c)) : "memory");
          const uint32_t wh = xhw[c >> 1], wl = xlw[c >> 1];
          const float x0 = __half2float(__ushort_as_half((unsigned short)(wh & 0xFFFFu))) + __half2float(__ushort_as_half((unsigned short)(wl & 0xFFFFu)));
          const float x1 = __half2float(__ushort_as_half((unsigned short)(wh >> 16))) + __half2float(__ushort_as_half((unsigned short)(wl >> 16)));
          unsigned short h0, l0, h1, l1;
          split16((__uint_as_float(orr[c]) + b2r0 + x0) * rs, h0, l0);
          split16((__uint_as_float(orr[c + 1]) + b2r1 + x1) * rs, h1, l1);
          yhw[c >> 1] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          ylw[c >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
          const float os0 = __uint_as_float(osk[c]) + b2s0, os1 = __uint_as_float(osk[c + 1]) + b2s1;
          float* sp = sk + (int64_t)c * p.T;
          __stcs(sp, skv[c] + os0);
          __stcs(sp + p.T, skv[c + 1] + os1);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          reinterpret_cast<uint4*>(p.yh + xoff)[u] = make_uint4(yhw[4 * u], yhw[4 * u + 1], yhw[4 * u + 2], yhw[4 * u + 3]);
          reinterpret_cast<uint4*>(p.yl + xoff)[u] = make_uint4(ylw[4 * u], ylw[4 * u + 1], ylw[4 * u + 2], ylw[4 * u + 3]);
        }
      }
    };
    int i = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, i++) {
      gate(i);
      if (prev_tile >= 0) epilogue(i - 1, prev_tile);
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(i - 1, prev_tile);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// fp32 channel-major (B, C, T) -> channels-last fp16 hi / lo planes (B, T, C); C % 8 == 0.  Thread = (b, t, group of 8
// channels): the eight loads are coalesced across lanes (consecutive t), the two stores are 16-byte vectors.
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, int B, int C, int64_t T) {
  A3T_PDL_TRIGGER();
  const int groups = C / 8;
  const int64_t n = (int64_t)B * groups * T;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx % T;
    const int64_t r = idx / T;
    const int g = (int)(r % groups);
    const int b = (int)(r / groups);
    const float* s0 = src + ((int64_t)b * C + 8 * g) * T + t;
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      unsigned short h0, l0, h1, l1;
      split16(s0[(int64_t)(2 * e) * T], h0, l0);
      split16(s0[(int64_t)(2 * e + 1) * T], h1, l1);
      hw[e] = (uint32_t)h0 | ((uint32_t)h1 << 16);
      lw[e] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    const int64_t o = ((int64_t)b * T + t) * C + 8 * g;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

static bool plane_map(CUtensorMap* m, const void* base, int64_t T, int C, int B) {
  const int64_t dims[4] = {C, T, B, 1}, str[3] = {C, T * C, T * C};
  const int box[4] = {64, 128, 1, 1};
  return encode_map(m, base, dims, str, box);
}
static bool weight_map(CUtensorMap* m, const void* base, int chunks) {   // (chunks, 128, 64) contiguous
  const int64_t dims[4] = {64, 128, chunks, 1}, str[3] = {64, 128 * 64, (int64_t)chunks * 128 * 64};
  const int box[4] = {64, 128, 1, 1};
  return encode_map(m, base, dims, str, box);
}

}  // namespace pwgtc
}  // namespace a3t

using namespace a3t;

extern "C" int a3t_pwg_split_planes(const float* src, void* hi, void* lo, int B, int C, int64_t T, void* stream) {
  A3T_REQUIRE(src && hi && lo && C > 0 && (C % 8) == 0, "pwg_split_planes: bad arguments (C must be a multiple of 8)");
  if (B == 0 || T == 0) return A3T_OK;
  int64_t blocks = ((int64_t)B * (C / 8) * T + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pwgtc::split_planes_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__half*)hi, (__half*)lo, B, C, T);
  return check_launch("pwg_split_planes");
}

extern "C" int a3t_pwg_resblock_tc(const void* xh, const void* xl, const void* ch, const void* cl, const void* w1h,
                                   const void* w1l, const void* w2h, const void* w2l, const float* b1, const float* b2,
                                   void* yh, void* yl, float* skip, int B, int64_t T, int dil, int flags, void* stream) {
  using namespace pwgtc;
  A3T_REQUIRE(xh && xl && ch && cl && w1h && w1l && w2h && w2l && b1 && b2 && yh && yl && skip, "pwg_resblock_tc: null pointer");
  A3T_REQUIRE(T < ((int64_t)1 << 31) && dil >= 1, "pwg_resblock_tc: bad sizes");
  A3T_REQUIRE(xh != yh && xl != yl, "pwg_resblock_tc: output planes must not alias the input (neighbouring tiles read the halo)");
  if (B == 0 || T == 0) return A3T_OK;
  CUtensorMap tmXh, tmXl, tmCh, tmCl, tmW1h, tmW1l, tmW2h, tmW2l;
  A3T_REQUIRE(plane_map(&tmXh, xh, T, 64, B) && plane_map(&tmXl, xl, T, 64, B) && plane_map(&tmCh, ch, T, 80, B) &&
                  plane_map(&tmCl, cl, T, 80, B) && weight_map(&tmW1h, w1h, KCH) && weight_map(&tmW1l, w1l, KCH) &&
                  weight_map(&tmW2h, w2h, 1) && weight_map(&tmW2l, w2l, 1),
              "pwg_resblock_tc: tensor map");
  Params p;
  p.xh = (const __half*)xh; p.xl = (const __half*)xl; p.yh = (__half*)yh; p.yl = (__half*)yl; p.skip = skip; p.b1 = b1; p.b2 = b2;
  p.B = B; p.T = T; p.dil = dil; p.first = flags & 1; p.passes = (flags & 2) ? 2 : 3;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(pwg_resblock_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("pwg_resblock_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return A3T_ERR_CUDA;
    }
    attr = true;
  }
  const int64_t ntiles = (int64_t)B * ((T + 127) / 128);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  pwg_resblock_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmXh, tmXl, tmCh, tmCl, tmW1h, tmW1l, tmW2h, tmW2l, p);
  return check_launch("pwg_resblock_tc");
}
