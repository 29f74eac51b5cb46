// tcgen05 tensor-core GEMM for sm_100a: the dense contractions of the A3T Conformer step.
//
//   * operands bf16, staged in shared memory by TMA (4-D tensor maps, 128-byte swizzle, zero OOB
//     fill — the OOB fill is what implements the 1-D conv "same" padding and ragged edges),
//   * fp32 accumulators in TMEM (2 x 256 columns, double buffered against the epilogue),
//   * one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16, cta_group::1),
//   * warp-specialised CTAs (1 per SM) that keep taking work items until none is left: warps 0..7 = epilogue
//     (tcgen05.ld -> bias / ReLU / mask / dropout / residual in registers -> swizzled staging rows -> TMA tensor
//     store), warp 8 = TMA producer, warp 9 = MMA issuer and TMEM owner.  The order of work items is dynamic:
//     cluster launch control (the grid is one CTA or CTA pair per item, running CTAs cancel pending ones and take
//     over their index), so the kernel uses whatever SMs are free when another kernel holds some.
//
// Modes (include/a3t_b200.h): PLAIN (any of the four operand-major combinations, batched),
// CONV (implicit 1-D conv: K loop over (tap, channel block), A rows shifted by tap - pad inside
// their sequence) and WGRAD (reduction over rows of every sequence, both operands MN-major,
// optional split-K with fp32 atomics).
//
// Replaces torch.nn.Linear / Conv1d / bmm calls of the reference: transformer/attention.py:55-96,
// :185-202, transformer/multi_layer_conv.py:61-62, conformer/convolution.py:28-54,
// tacotron2/decoder.py:189-238 and their autograd backward.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace a3t {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
// The warp scheduler favours the highest warp id among eligible warps: the two single-lane control warps
// (TMA producer, MMA issuer) sit ABOVE the eight instruction-heavy epilogue warps so they are never starved.
constexpr int PRODUCER_WARP = NUM_EPI_WARPS, MMA_WARP = NUM_EPI_WARPS + 1;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // one 32x32 fp32 transposition buffer per epilogue warp
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns per accumulator stage
constexpr int SMEM_BYTES_MAX = 227 * 1024;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int MAX_STAGES = 8;
constexpr int CLC_SLOTS = 4;   // ring of cluster-launch-control responses (work items fetched ahead)
constexpr int NUM_BARS = 2 * MAX_STAGES + 5 + NUM_EPI_WARPS + 3 * CLC_SLOTS;
constexpr int BAR_BYTES = ((8 * NUM_BARS + 15) / 16) * 16 + 16 * CLC_SLOTS + 16;

struct Params {
  A3tGemmDesc d;
  const float* bias;
  const float* res;
  const void* mask;
  const unsigned long long* seed;
  void* C;
  uint32_t idesc;
  int block_n;          // UMMA N (multiple of 16, <= 256)
  int stages;
  int a_mn, b_mn;       // 1 = operand is MN-major in global/shared memory
  int m_tiles_per_seq;  // CONV: tiles per sequence, else all M tiles
  int m_tiles;          // total M tiles (per batch entry)
  int m_units;          // scheduling units along M: m_tiles (1-CTA) or ceil(m_tiles / 2) (CTA pairs)
  int n_tiles_per_tap;  // WGRAD: N tiles per tap, else all N tiles
  int n_step;           // column step of consecutive N tiles (block_n; 64 when a WGRAD tile holds all taps of 64 channels)
  int wg_alltaps;       // WGRAD: the N tile is (tap, 64 channels) for every tap -> interleaved (c, tap) output rows
  int n_tiles;          // total N tiles
  int k_iters;          // total K iterations of one output tile
  int splits;           // split-K factor (atomics when > 1)
  int cblocks;          // CONV: ceil(cin / 64);  WGRAD: ceil(seq / 64)
  int num_work;         // m_tiles * n_tiles * batch * splits
  int a_c2, a_c3, b_c2, b_c3;  // 0 when that batch coordinate is pinned (stride 0 / size 1)
  int vec_c, vec_r, vec_m;     // 16-byte vector access allowed for C / residual / mask
  int clc;                     // 1 = dynamic tile order through cluster launch control (grid = one CTA / pair per work item,
                               // running CTAs cancel pending ones and take over their index); 0 = static round robin
  int dbg;                     // tuning experiments: 1 = no loads, MMAs free-run; 2 = loads run, MMAs do not wait; 3 = loads only; 5 = loads only, no epilogue; 6 = no epilogue (timing only, garbage results)
};

// Weight gradient of a 3-tap conv: the accumulator row holds [tap][64 channels]; the output row is
// [channel][tap].  Chunk K = floats 32K .. 32K+31 of the interleaved 192-float row (compile-time mapping).
template <int K>
__device__ __forceinline__ void wgrad3_chunk(uint32_t taddr_row, float alpha, uint32_t (&out)[32]) {
  constexpr int CS = (32 * K) / 3;  // first channel touched by this chunk
  uint32_t a0[16], a1[16], a2[16];
  tmem_ld16(taddr_row + CS, a0);
  tmem_ld16(taddr_row + 64 + CS, a1);
  tmem_ld16(taddr_row + 128 + CS, a2);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const int f = 32 * K + i, c = f / 3 - CS, tp = f % 3;
    const uint32_t v = tp == 0 ? a0[c] : (tp == 1 ? a1[c] : a2[c]);
    out[i] = __float_as_uint(alpha * __uint_as_float(v));
  }
}

struct Work {
  int z, b1, b2;
  int seq_idx, m0;   // m0: row inside the sequence (CONV) or global row
  int tap_n, n0;     // WGRAD: tap of the N tile and channel offset; else n0 = column offset
  int k_begin, k_end;
};

// w enumerates (z, m unit, n tile, split); in CTA-pair mode an M unit is two M tiles, one per CTA rank
// (the second tile of the last unit may not exist: its loads are out of bounds = zeros, nothing is stored)
__device__ __forceinline__ Work decode_work(const Params& p, int w, int rank = 0, int per_unit = 1) {
  Work t;
  int split = w % p.splits;
  int r = w / p.splits;
  int n_t = r % p.n_tiles;
  r /= p.n_tiles;
  int m_t = (r % p.m_units) * per_unit + rank;
  t.z = r / p.m_units;
  t.b1 = t.z / p.d.batch2;
  t.b2 = t.z - t.b1 * p.d.batch2;
  t.seq_idx = m_t / p.m_tiles_per_seq;
  t.m0 = (m_t - t.seq_idx * p.m_tiles_per_seq) * BLOCK_M;
  t.tap_n = n_t / p.n_tiles_per_tap;
  t.n0 = (n_t - t.tap_n * p.n_tiles_per_tap) * p.n_step;
  int per = (p.k_iters + p.splits - 1) / p.splits;
  t.k_begin = split * per;
  t.k_end = min(p.k_iters, t.k_begin + per);
  return t;
}

// ---------------------------------------------------------------------------------------------
// epilogue helper: 4 consecutive columns of one output row (lane layout after the smem transpose:
// 8 lanes cover 32 consecutive columns of a row, a warp instruction touches 4 rows x 128 B)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue4(const Params& p, const Drop& dr, float4 a, int n4, int nlim, int64_t crow,
                                          int64_t rrow, unsigned long long drow, int tap_n) {
  const A3tGemmDesc& d = p.d;
  float v[4] = {d.alpha * a.x, d.alpha * a.y, d.alpha * a.z, d.alpha * a.w};
  const bool full = n4 + 4 <= nlim;
  const int ncol = (d.mode == A3T_GEMM_WGRAD) ? tap_n * d.cin + n4 : n4;  // logical column (bias / dropout index)
  if (p.bias) {
    if (full) {
      float4 b0 = __ldg((const float4*)(p.bias + ncol));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n4 + j < nlim) v[j] += __ldg(p.bias + ncol + j);
    }
  }
  if (d.relu) {
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
  }
  // element offset of column n4 inside C (and the mask, which shares C's strides)
  const int64_t coff = (d.mode == A3T_GEMM_WGRAD) ? crow + (int64_t)tap_n * d.sc_tap + (int64_t)n4 * d.sc_n
                                                  : crow + (int64_t)n4 * d.sc_n;
  if (p.mask) {
    if (full && p.vec_m) {
      float mm[4];
      if (d.dtype_mask == A3T_BF16) {
        uint2 mv = __ldg((const uint2*)((const __nv_bfloat16*)p.mask + coff));
        const __nv_bfloat16* mb = (const __nv_bfloat16*)&mv;
#pragma unroll
        for (int j = 0; j < 4; j++) mm[j] = __bfloat162float(mb[j]);
      } else {
        float4 m0 = __ldg((const float4*)((const float*)p.mask + coff));
        mm[0] = m0.x; mm[1] = m0.y; mm[2] = m0.z; mm[3] = m0.w;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = (mm[j] != 0.f) ? v[j] * d.mask_scale : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n4 + j < nlim) {
          float mv = load_as_f32(p.mask, d.dtype_mask, coff + (int64_t)j * d.sc_n);
          v[j] = (mv != 0.f) ? v[j] * d.mask_scale : 0.f;
        }
    }
  }
  if (dr.on) {
    const unsigned long long idx0 = drow + (unsigned long long)ncol;
    if ((idx0 & 3ull) == 0) {
      bool kp[4];
      drop_keep4(dr, drop_fold(idx0), kp);
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = kp[j] ? v[j] * dr.inv_keep : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = drop_keep(dr, idx0 + j) ? v[j] * dr.inv_keep : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] *= d.out_scale;
  if (p.res) {
    const float* rp = p.res + rrow + (int64_t)n4 * d.sr_n;
    if (full && p.vec_r) {
      float4 r0 = __ldg((const float4*)rp);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n4 + j < nlim) v[j] += __ldg(rp + (int64_t)j * d.sr_n);
    }
  }
  if (p.splits > 1) {  // split-K partial sums (fp32 output, zero-initialised by the launcher)
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (n4 + j < nlim) atomicAdd((float*)p.C + coff + (int64_t)j * d.sc_n, v[j]);
    return;
  }
  if (full && p.vec_c) {
    if (d.dtype_c == A3T_BF16) {
      __nv_bfloat162 h[2] = {__floats2bfloat162_rn(v[0], v[1]), __floats2bfloat162_rn(v[2], v[3])};
      *(uint2*)((__nv_bfloat16*)p.C + coff) = *(const uint2*)h;
    } else {
      *(float4*)((float*)p.C + coff) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (n4 + j < nlim) store_from_f32(p.C, d.dtype_c, coff + (int64_t)j * d.sc_n, v[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// EPI: -1 = generic epilogue (any strides, WGRAD scatter, split-K atomics); otherwise a bit set
// EPI_MASK | EPI_DROP | EPI_RES | EPI_BF16 selecting the specialised vectorised epilogue.
constexpr int EPI_MASK = 1, EPI_DROP = 2, EPI_RES = 4, EPI_BF16 = 8;
constexpr int EPI_WGRAD = 16;  // fp32 weight-gradient scatter (N, C, taps), optional split-K atomics
constexpr int EPI_WGRAD_TMA = 17;  // fp32 weight gradient through TMA tensor store / reduce-add (taps 1 or 3)

template <int EPI, bool CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ Params p) {
  A3T_PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA-pair mode: each CTA stages its own 128 A rows and HALF of the B tile; the leader's MMAs read both
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int per_unit = CTA2 ? 2 : 1;
  const int group = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // scheduling group (CTA or pair)
  const int ngroups = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int b_rows = CTA2 ? p.block_n / 2 : p.block_n;                      // B rows (N extent) staged by this CTA
  const uint32_t b_stage_bytes = (uint32_t)b_rows * BLOCK_K * 2;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const uint32_t epi_base = smem_base + p.stages * stage_bytes;   // per-warp transposition buffers
  const uint32_t bar_base = epi_base + NUM_EPI_WARPS * EPI_STAGE_BYTES;  // 8-byte mbarriers after them
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
  auto res_bar = [&](int w) { return bar_base + 8u * (2 * MAX_STAGES + 5 + w); };  // one per epilogue warp
  // cluster launch control: response ring + full (response landed) / empty (consumers done) / armed (pair mode: the
  // peer CTA has armed its own full barrier, the leader may multicast the next response) barriers
  constexpr int CLC_BAR0 = 2 * MAX_STAGES + 5 + NUM_EPI_WARPS;
  auto clc_full = [&](int s) { return bar_base + 8u * (CLC_BAR0 + s); };
  auto clc_empty = [&](int s) { return bar_base + 8u * (CLC_BAR0 + CLC_SLOTS + s); };
  auto clc_armed = [&](int s) { return bar_base + 8u * (CLC_BAR0 + 2 * CLC_SLOTS + s); };
  const uint32_t clc_resp = bar_base + ((8u * NUM_BARS + 15u) & ~15u);
  const bool clc = p.clc != 0;
  // next work item of this role.  Static order: stride ngroups.  CLC: wait for the response of query #it (issued by
  // the producer thread one item ahead), release the slot (MMA thread: arrive; epilogue warps: lane 0 after a warp
  // sync; the producer threads do not arrive -- they are the ones waiting on `empty`).
  auto next_work = [&](int& w, int& it, int who) -> bool {   // who: 0 = producer, 1 = single thread, 2 = whole warp
    if (!clc) {
      w += ngroups;
      return w < p.num_work;
    }
    const int slot = it & (CLC_SLOTS - 1);
    mbar_wait(clc_full(slot), (uint32_t)(it / CLC_SLOTS) & 1u);
    uint32_t x;
    const bool ok = clc_read(clc_resp + 16u * slot, x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot is rewritten through the async proxy
    if (who == 1) mbar_arrive(clc_empty(slot));
    if (who == 2) {
      __syncwarp();
      if (lane == 0) mbar_arrive(clc_empty(slot));
    }
    it++;
    w = (int)(CTA2 ? (x >> 1) : x);
    return ok;
  };

  if (warp == PRODUCER_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    if constexpr (EPI >= 0 && EPI != EPI_WGRAD) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
    if constexpr (EPI >= 0 && EPI < 16 && (EPI & 4)) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmR) : "memory");
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * per_unit);  // pair mode: the peer's epilogue warps arrive remotely
    }
    for (int w = 0; w < NUM_EPI_WARPS; w++) mbar_init(res_bar(w), 1);
    for (int s = 0; s < CLC_SLOTS; s++) {
      mbar_init(clc_full(s), 1);
      mbar_init(clc_empty(s), NUM_EPI_WARPS + (rank == 0 ? 1 : 0));  // epilogue warps (+ the MMA thread on the leader)
      mbar_init(clc_armed(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers must be initialised before anything targets them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  A3T_PDL_WAIT();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  const A3tGemmDesc& d = p.d;

#ifdef A3T_TUNING
  const int dbg = p.dbg;
#else
  constexpr int dbg = 0;  // timing-experiment modes (garbage results) exist in -DA3T_TUNING builds only
#endif
  if (warp == PRODUCER_WARP) {
    // ===================================== TMA producer =====================================
    // The K loop is one thread issuing dependent instructions: it is kept to a barrier wait, the TMA
    // instructions and a handful of adds.  Every tensor-map coordinate is linear in a two-level counter
    // (inner j < cblocks, outer o): CONV (j = channel block, o = tap), WGRAD (j = row block inside the
    // sequence, o = sequence), PLAIN (j = K block, never wraps) -- no division inside the loop.
    if (elect_one() && dbg != 1) {
      int s = 0;
      uint32_t ph = 0;
      const int a_boxes = p.a_mn ? BLOCK_M / 64 : 1;
      const int b_boxes = p.b_mn ? b_rows / 64 : 1;
      const int b_off = (int)rank * b_rows;  // this CTA's slice of the B tile along N
      // coordinate order: {a inner, a outer, a dim2, a dim3, b inner, b outer, b dim2, b dim3}
      int dj[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int inner = p.k_iters;
      if (d.mode == A3T_GEMM_CONV) {
        inner = p.cblocks;
        dj[0] = BLOCK_K; dj[4] = BLOCK_K;
        dw[0] = -p.cblocks * BLOCK_K; dw[1] = 1; dw[4] = d.cin - p.cblocks * BLOCK_K;
      } else if (d.mode == A3T_GEMM_WGRAD) {
        inner = p.cblocks;
        dj[1] = BLOCK_K; dj[5] = BLOCK_K;
        dw[1] = -p.cblocks * BLOCK_K; dw[2] = 1; dw[5] = -p.cblocks * BLOCK_K; dw[6] = 1;
      } else {
        dj[0] = p.a_mn ? 0 : BLOCK_K; dj[1] = p.a_mn ? BLOCK_K : 0;
        dj[4] = p.b_mn ? 0 : BLOCK_K; dj[5] = p.b_mn ? BLOCK_K : 0;
      }
      int w = group, wit = 0;
      for (;;) {
        if (clc) {
          // query #wit (answers: which item follows this one) goes out before this item's loads
          const int slot = wit & (CLC_SLOTS - 1);
          const uint32_t sph = (uint32_t)(wit / CLC_SLOTS) & 1u;
          mbar_wait(clc_empty(slot), sph ^ 1u);
          mbar_expect_tx(clc_full(slot), 16);
          if constexpr (CTA2) {
            if (rank != 0) {
              mbar_arrive_leader(clc_armed(slot));
            } else {
              mbar_wait(clc_armed(slot), sph);
              clc_try_cancel_multicast(clc_resp + 16u * slot, clc_full(slot));
            }
          } else {
            clc_try_cancel(clc_resp + 16u * slot, clc_full(slot));
          }
        }
        const Work t = decode_work(p, w, (int)rank, per_unit);
        int o0 = 0, j = t.k_begin;
        if (inner != p.k_iters) { o0 = t.k_begin / inner; j = t.k_begin - o0 * inner; }
        int c[8];
        if (d.mode == A3T_GEMM_CONV) {
          c[0] = 0; c[1] = t.m0 - d.pad; c[2] = t.seq_idx; c[3] = 0;
          c[4] = 0; c[5] = t.n0 + b_off; c[6] = 0; c[7] = 0;
        } else if (d.mode == A3T_GEMM_WGRAD) {
          c[0] = t.m0; c[1] = 0; c[2] = 0; c[3] = 0;
          c[4] = t.n0 + b_off; c[5] = t.tap_n - d.pad; c[6] = 0; c[7] = 0;
        } else {
          c[0] = p.a_mn ? t.m0 : 0; c[1] = p.a_mn ? 0 : t.m0; c[2] = t.b2 * p.a_c2; c[3] = t.b1 * p.a_c3;
          c[4] = p.b_mn ? t.n0 + b_off : 0; c[5] = p.b_mn ? 0 : t.n0 + b_off; c[6] = t.b2 * p.b_c2; c[7] = t.b1 * p.b_c3;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] += j * dj[i] + o0 * (dw[i] + inner * dj[i]);
        for (int kit = t.k_begin; kit < t.k_end; kit++) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * stage_bytes, sb = sa + A_STAGE_BYTES;
          const uint32_t fb = full_bar(s);
          if constexpr (CTA2) {
            // one transaction barrier (the leader's) collects the bytes of both CTAs
            if (rank == 0) mbar_expect_tx(fb, 2 * stage_bytes);
            for (int q = 0; q < a_boxes; q++)
              tma_load_4d_2sm(sa + q * (BLOCK_K * 128), &tmA, fb, c[0] + 64 * q, c[1], c[2], c[3]);
            for (int q = 0; q < b_boxes; q++)
              tma_load_4d_2sm(sb + q * (BLOCK_K * 128), &tmB, fb, c[4] + 64 * q, c[5], c[6], c[7]);
          } else {
            mbar_expect_tx(fb, stage_bytes);
            for (int q = 0; q < a_boxes; q++)
              tma_load_4d(sa + q * (BLOCK_K * 128), &tmA, fb, c[0] + 64 * q, c[1], c[2], c[3]);
            if (p.wg_alltaps) {  // box q = tap q of the same 64 channels: rows shifted by q
              for (int q = 0; q < b_boxes; q++)
                tma_load_4d(sb + q * (BLOCK_K * 128), &tmB, fb, c[4], c[5] + q, c[6], c[7]);
            } else {
              for (int q = 0; q < b_boxes; q++)
                tma_load_4d(sb + q * (BLOCK_K * 128), &tmB, fb, c[4] + 64 * q, c[5], c[6], c[7]);
            }
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
#pragma unroll
          for (int i = 0; i < 8; i++) c[i] += dj[i];
          if (++j == inner) {
            j = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] += dw[i];
          }
        }
        if (!next_work(w, wit, 0)) break;
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA issuer =======================================
    if (elect_one() && rank == 0) {  // pair mode: only the leader CTA issues MMAs (they span both SMs)
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      // per-instruction K advance inside a stage: K-major = 32 B along the swizzled row,
      // MN-major = 16 k-rows of 128 B.  Descriptors differ only in the 14-bit start-address field
      // (shared memory < 256 KB, so adding byte offsets >> 4 to the low word cannot carry out of it).
      const uint32_t a_kstep = (p.a_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t b_kstep = (p.b_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint64_t adesc0 = make_smem_desc(smem_base, p.a_mn ? BLOCK_K * 128 : 16);
      const uint64_t bdesc0 = make_smem_desc(smem_base + A_STAGE_BYTES, p.b_mn ? BLOCK_K * 128 : 16);
      const uint32_t stage16 = stage_bytes >> 4;
      const bool wait_full = dbg != 1 && dbg != 2, do_mma = dbg != 3 && dbg != 5;
      const uint32_t idesc = p.idesc;
      int w = group, wit = 0;
      for (;;) {
        const Work t = decode_work(p, w, 0, per_unit);
        mbar_wait(tempty_bar(as), aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
        uint32_t accum = 0;
        for (int kit = t.k_begin; kit < t.k_end; kit++) {
          if (wait_full) mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(s * stage16), bd = bdesc0 + (uint64_t)(s * stage16);
          if (do_mma) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
              if constexpr (CTA2) umma_bf16_2sm(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, accum);
              else umma_bf16(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, accum);
              accum = 1;
            }
          }
          // frees the smem stage (in both CTAs) once the MMAs above have read it
          if constexpr (CTA2) umma_commit_2sm(empty_bar(s));
          else umma_commit(empty_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTA2) umma_commit_2sm(tfull_bar(as));
        else umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aph ^= 1; }
        if (!next_work(w, wit, 1)) break;
      }
    }
  } else {
    // ===================================== epilogue =========================================
    // 8 warps: warp%4 selects the TMEM lane quarter (hardware rule), (warp-2)/4 the odd/even 32-column
    // chunks.  Each chunk: tcgen05.ld (thread = row) -> swizzled smem transpose -> lanes along the
    // row, so residual/mask loads and the C stores are coalesced 128-byte row segments.
    const int ew = warp;
    const int q = warp & 3;
    const int half = ew >> 2;
    const uint32_t stg = epi_base + ew * EPI_STAGE_BYTES;
    int as = 0;
    uint32_t aph = 0;
    const Drop dr = make_drop(d.drop_p, p.seed, d.drop_site);
    const int nchunks = (dbg >= 5) ? 0 : (p.block_n + 31) / 32;  // dbg 5/6: barrier handshake only, nothing read or stored
    const int rsub = lane >> 3, c4 = lane & 7;
    const int nlim = (d.mode == A3T_GEMM_WGRAD) ? d.cin : d.N;
    if constexpr (EPI == EPI_WGRAD_TMA) {
      // ---- weight gradient, fp32 (N, C, taps) with taps innermost: thread = row, 32-float chunks of the
      // output row staged in 128B-swizzled shared memory, then ONE TMA tensor store (or reduce-add when
      // the K range is split over CTAs) per chunk: no scalar atomics, full-line requests ----
      const float alpha = d.alpha;
      const bool reduce = p.splits > 1;
      const uint32_t stg_row = stg + lane * 128;
      const uint32_t sw = (uint32_t)(lane & 7);
      const int nch = (dbg >= 5) ? 0 : (p.wg_alltaps ? 6 : (p.block_n + 31) / 32);
      int w = group, wit = 0;
      for (;;) {
        const Work t = decode_work(p, w, (int)rank, per_unit);
        const int trow0 = t.m0 + q * 32;
        const uint32_t taddr_row = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
        const int fcol0 = p.wg_alltaps ? t.n0 * 3 : t.n0;  // first float of this tile inside the output row
        mbar_wait(tfull_bar(as), aph);
        tc_fence_after();
        for (int k = half; k < nch; k += NUM_EPI_WARPS / 4) {
          uint32_t out[32];
          if (p.wg_alltaps) {
            switch (k) {
              case 0: wgrad3_chunk<0>(taddr_row, alpha, out); break;
              case 1: wgrad3_chunk<1>(taddr_row, alpha, out); break;
              case 2: wgrad3_chunk<2>(taddr_row, alpha, out); break;
              case 3: wgrad3_chunk<3>(taddr_row, alpha, out); break;
              case 4: wgrad3_chunk<4>(taddr_row, alpha, out); break;
              default: wgrad3_chunk<5>(taddr_row, alpha, out); break;
            }
          } else {
            tmem_ld32(taddr_row + k * 32, out);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) out[i] = __float_as_uint(alpha * __uint_as_float(out[i]));
          }
          if (k + NUM_EPI_WARPS / 4 >= nch) {  // all TMEM reads of this warp are done: hand the stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; u++)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((u ^ sw) << 4)), "r"(out[4 * u]),
                         "r"(out[4 * u + 1]), "r"(out[4 * u + 2]), "r"(out[4 * u + 3])
                         : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            const int c0 = fcol0 + 32 * k;
            if (reduce)
              asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                               (uint64_t)&tmC),
                           "r"(stg), "r"(c0), "r"(trow0), "r"(0), "r"(0)
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                               (uint64_t)&tmC),
                           "r"(stg), "r"(c0), "r"(trow0), "r"(0), "r"(0)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (nch <= half) {  // this warp had no chunk: still release the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
        if (!next_work(w, wit, 2)) break;
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else
    if constexpr (EPI == EPI_WGRAD) {
      // ---- weight gradient: C[m, (tap, c)] fp32 at m*sc_m + tap*sc_tap + c*sc_n, alpha only ----
      const float alpha = d.alpha;
      const bool atomic = p.splits > 1;
      float* const Cf = (float*)p.C;
      int w = group, wit = 0;
      for (;;) {
        const Work t = decode_work(p, w, (int)rank, per_unit);
        const int row0 = t.m0 + q * 32 + rsub;
        bool waited = false, released = false;
        for (int c = half; c < nchunks; c += 2) {
          const int col = c * 32 + c4 * 4;
          const int n4 = t.n0 + col;
          const bool col_ok = col < p.block_n && n4 < d.cin;
          if (!waited) {
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            waited = true;
          }
          uint32_t acc[32];
          tmem_ld32(tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + c * 32, acc);
          tmem_ld_wait();
          if (c + 2 >= nchunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
            released = true;
          }
#pragma unroll
          for (int g4 = 0; g4 < 8; g4++)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((g4 ^ (lane & 7)) << 4)),
                         "r"(acc[4 * g4]), "r"(acc[4 * g4 + 1]), "r"(acc[4 * g4 + 2]), "r"(acc[4 * g4 + 3])
                         : "memory");
          __syncwarp();
          float* const cp0 = Cf + (int64_t)row0 * d.sc_m + (int64_t)t.tap_n * d.sc_tap + (int64_t)n4 * d.sc_n;
#pragma unroll
          for (int it = 0; it < 8; it++) {
            const int R = it * 4 + rsub;
            float4 a;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                         : "r"(stg + R * 128 + ((c4 ^ (R & 7)) << 4))
                         : "memory");
            if (col_ok && (row0 + it * 4) < d.M) {
              float* cp = cp0 + (int64_t)it * 4 * d.sc_m;
              const float v[4] = {alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w};
#pragma unroll
              for (int j = 0; j < 4; j++)
                if (n4 + j < d.cin) {
                  if (atomic) atomicAdd(cp + (int64_t)j * d.sc_n, v[j]);
                  else cp[(int64_t)j * d.sc_n] = v[j];
                }
            }
          }
          __syncwarp();
        }
        if (!waited) {
          mbar_wait(tfull_bar(as), aph);
          tc_fence_after();
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
        if (!next_work(w, wit, 2)) break;
      }
    } else
    if constexpr (EPI >= 0) {
      // ---- specialised path (PLAIN/CONV, unit column stride): thread = row.  A "superchunk" is the span
      // of columns whose output is 128 bytes per row (32 fp32 / 64 bf16): tcgen05.ld -> bias / ReLU /
      // mask / dropout / residual in registers -> 128B-swizzled staging rows in shared memory -> one TMA
      // tensor store per superchunk (the tensor map clips rows and columns outside C, which also
      // implements the sequence boundary of CONV and the phantom tile of a CTA pair).  Mask and residual
      // are read straight from global memory by the row's own thread (16-byte loads, L1 keeps the lines). ----
      constexpr bool kMask = EPI & EPI_MASK, kDrop = EPI & EPI_DROP, kRes = EPI & EPI_RES, kBf16 = EPI & EPI_BF16;
      constexpr int SC_COLS = kBf16 ? 64 : 32;
      static_assert(!(kRes && kBf16), "residual epilogue is fp32-output only");
      const float alpha = d.alpha, out_scale = d.out_scale, mask_scale = d.mask_scale;
      const bool relu = d.relu != 0;
      const float drop_mul = kDrop ? dr.inv_keep * out_scale : out_scale;
      const int row_lim = (d.mode == A3T_GEMM_CONV) ? d.seq : d.M;
      const int nsc = (dbg >= 5) ? 0 : (p.block_n + SC_COLS - 1) / SC_COLS;
      const uint32_t stg_row = stg + lane * 128;
      const uint32_t sw = (uint32_t)(lane & 7);
      uint32_t rph = 0;  // phase of this warp's residual barrier
      int w = group, wit = 0;
      for (;;) {
        const Work t = decode_work(p, w, (int)rank, per_unit);
        const int trow0 = t.m0 + q * 32;                              // first row of this warp (in sequence / matrix)
        const int row = trow0 + lane;
        const int mrow = (d.mode == A3T_GEMM_CONV) ? t.seq_idx * d.seq + row : row;  // row inside C
        const bool row_ok = row < row_lim && (d.mode != A3T_GEMM_CONV || t.seq_idx * d.seq < d.M);
        const int64_t cbase = (int64_t)t.b1 * d.sc_b1 + (int64_t)t.b2 * d.sc_b2 + (int64_t)mrow * d.sc_m;
        const unsigned long long dbase = ((unsigned long long)t.z * d.M + mrow) * (unsigned long long)d.N;
        const int cc2 = (d.mode == A3T_GEMM_CONV) ? t.seq_idx : t.b2, cc3 = (d.mode == A3T_GEMM_CONV) ? 0 : t.b1;
        bool waited = false, released = false;
        for (int sc = half; sc < nsc; sc += NUM_EPI_WARPS / 4) {
          const int n0c = t.n0 + sc * SC_COLS;
          // operands of the epilogue that live in global memory: issue the loads before waiting on TMEM
          uint4 mk[kMask ? SC_COLS / 8 : 1];
          if constexpr (kRes) {
            // residual tile (32 rows x 128 B, fp32): TMA-loaded into this warp's staging buffer while the
            // accumulator is read and the bias / dropout math runs; added in place before the tensor store
            if (lane == 0) {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous store has read the buffer
              mbar_expect_tx(res_bar(ew), 32 * 128);
              tma_load_4d(stg, &tmR, res_bar(ew), n0c, trow0, cc2, cc3);
            }
          }
          if constexpr (kMask) {
#pragma unroll
            for (int g8 = 0; g8 < SC_COLS / 8; g8++) {
              mk[g8] = make_uint4(0u, 0u, 0u, 0u);
              if (row_ok && n0c + 8 * g8 < d.N) mk[g8] = __ldg((const uint4*)((const __nv_bfloat16*)p.mask + cbase + n0c + 8 * g8));
            }
          }
          if (!waited) {
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            waited = true;
          }
          uint32_t acc[SC_COLS];
          {
            const uint32_t taddr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + sc * SC_COLS;
            tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
            if constexpr (SC_COLS == 64) tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          }
          tmem_ld_wait();
          if (sc + NUM_EPI_WARPS / 4 >= nsc) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
            released = true;
          }
#pragma unroll
          for (int g4 = 0; g4 < SC_COLS / 4; g4++) {
            const int n4 = n0c + 4 * g4;
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && n4 < d.N) bias4 = __ldg((const float4*)(p.bias + n4));  // same address in every lane
            float v[4] = {fmaf(alpha, __uint_as_float(acc[4 * g4]), bias4.x), fmaf(alpha, __uint_as_float(acc[4 * g4 + 1]), bias4.y),
                          fmaf(alpha, __uint_as_float(acc[4 * g4 + 2]), bias4.z), fmaf(alpha, __uint_as_float(acc[4 * g4 + 3]), bias4.w)};
            if (relu) {
#pragma unroll
              for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
            }
            if constexpr (kMask) {
              const uint32_t* mw = reinterpret_cast<const uint32_t*>(&mk[g4 >> 1]) + 2 * (g4 & 1);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const uint32_t bits = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xFFFFu);
                v[j] = (bits & 0x7FFFu) ? v[j] * mask_scale : 0.f;   // bf16 +-0 -> masked
              }
            }
            if constexpr (kDrop) {
              bool kp[4];  // N % 4 == 0 and n4 % 4 == 0 on this path: the element index is a multiple of 4
              drop_keep4(dr, drop_fold(dbase + (unsigned long long)n4), kp);
#pragma unroll
              for (int j = 0; j < 4; j++) v[j] = kp[j] ? v[j] * drop_mul : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < 4; j++) v[j] *= out_scale;
            }
            if constexpr (kBf16) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
              acc[2 * g4] = *reinterpret_cast<uint32_t*>(&h0);       // packed in place: unit u = acc[4u .. 4u+3]
              acc[2 * g4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
            } else {
#pragma unroll
              for (int j = 0; j < 4; j++) acc[4 * g4 + j] = __float_as_uint(v[j]);
            }
          }
          if constexpr (kRes) {
            mbar_wait(res_bar(ew), rph);  // residual rows have landed (rows / columns outside the tensor read as 0)
            rph ^= 1u;
#pragma unroll
            for (int u = 0; u < 8; u++) {
              float4 rr;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w)
                           : "r"(stg_row + ((u ^ sw) << 4))
                           : "memory");
              acc[4 * u] = __float_as_uint(__uint_as_float(acc[4 * u]) + rr.x);
              acc[4 * u + 1] = __float_as_uint(__uint_as_float(acc[4 * u + 1]) + rr.y);
              acc[4 * u + 2] = __float_as_uint(__uint_as_float(acc[4 * u + 2]) + rr.z);
              acc[4 * u + 3] = __float_as_uint(__uint_as_float(acc[4 * u + 3]) + rr.w);
            }
          } else {
            // the previous TMA store of this warp must have finished READING the staging rows
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int u = 0; u < 8; u++)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((u ^ sw) << 4)), "r"(acc[4 * u]),
                         "r"(acc[4 * u + 1]), "r"(acc[4 * u + 2]), "r"(acc[4 * u + 3])
                         : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                             (uint64_t)&tmC),
                         "r"(stg), "r"(n0c), "r"(trow0), "r"(cc2), "r"(cc3)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (!waited) {
          mbar_wait(tfull_bar(as), aph);
          tc_fence_after();
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
        if (!next_work(w, wit, 2)) break;
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
    } else {
    int w = group, wit = 0;
    for (;;) {
      const Work t = decode_work(p, w, (int)rank, per_unit);
      const int64_t cbase = (int64_t)t.b1 * d.sc_b1 + (int64_t)t.b2 * d.sc_b2;
      const int64_t rbase = (int64_t)t.b1 * d.sr_b1 + (int64_t)t.b2 * d.sr_b2;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      bool released = false;
      for (int c = half; c < nchunks; c += 2) {
        uint32_t acc[32];
        tmem_ld32(taddr + c * 32, acc);
        tmem_ld_wait();
        if (c + 2 >= nchunks) {  // this warp has read all its columns: hand the TMEM stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
          released = true;
        }
#pragma unroll
        for (int g4 = 0; g4 < 8; g4++)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((g4 ^ (lane & 7)) << 4)),
                       "r"(acc[4 * g4]), "r"(acc[4 * g4 + 1]), "r"(acc[4 * g4 + 2]), "r"(acc[4 * g4 + 3])
                       : "memory");
        __syncwarp();
        const int col = c * 32 + c4 * 4;       // column inside the tile
        const int n4 = t.n0 + col;             // column inside C (channel inside the tap for WGRAD)
        const bool col_ok = col < p.block_n && n4 < nlim;
#pragma unroll 2
        for (int it = 0; it < 8; it++) {
          const int R = it * 4 + rsub;
          float4 a;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                       : "r"(stg + R * 128 + ((c4 ^ (R & 7)) << 4))
                       : "memory");
          const int row = t.m0 + q * 32 + R;
          int m;
          bool row_ok;
          if (d.mode == A3T_GEMM_CONV) {
            row_ok = row < d.seq && t.seq_idx * d.seq < d.M;
            m = t.seq_idx * d.seq + row;
          } else {
            m = row;
            row_ok = m < d.M;
          }
          if (row_ok && col_ok)
            epilogue4(p, dr, a, n4, nlim, cbase + (int64_t)m * d.sc_m, rbase + (int64_t)m * d.sr_m,
                      ((unsigned long long)t.z * d.M + m) * (unsigned long long)d.N, t.tap_n);
        }
        __syncwarp();
      }
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as)); }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
      if (!next_work(w, wit, 2)) break;
    }
    }
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory until the very end
  else __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static inline bool mul8(int64_t v) { return (v & 7) == 0; }

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (const char* e = tune_env("A3T_TC_SMS")) {  // tuning experiments: restrict the persistent grid
      int v = atoi(e);
      if (v >= 2 && v < n) n = v;
    }
  }
  return n;
}

}  // namespace tc

int gemm_tc_launch(const A3tGemmDesc* dp, const void* A, const void* B, void* C, const float* bias,
                   const float* res, const void* mask, const unsigned long long* seed, cudaStream_t st,
                   bool probe_only) {
  using namespace tc;
  const A3tGemmDesc& d = *dp;
  // ---- qualification ------------------------------------------------------------------------
  if (d.dtype_a != A3T_BF16 || d.dtype_b != A3T_BF16) return A3T_ERR_UNSUPPORTED;
  if (d.M < 1 || d.N < 1 || d.K < 1) return A3T_ERR_UNSUPPORTED;
  if (!al16(A) || !al16(B)) return A3T_ERR_UNSUPPORTED;
  const int nbatch = d.batch1 * d.batch2;
  Params p;
  memset(&p, 0, sizeof(p));
  p.d = d;
  int64_t adims[4], astr[3], bdims[4], bstr[3];
  int abox[4] = {64, 1, 1, 1}, bbox[4] = {64, 1, 1, 1};
  p.a_c2 = p.a_c3 = p.b_c2 = p.b_c3 = 0;
  int nlim = d.N;  // extent one "N tile row" covers
  if (d.mode == A3T_GEMM_CONV) {
    if (nbatch != 1 || d.sa_k != 1 || d.sb_k != 1 || d.sb_tap != d.cin || !mul8(d.sa_m) || !mul8(d.sb_n))
      return A3T_ERR_UNSUPPORTED;
    p.a_mn = p.b_mn = 0;
    adims[0] = d.cin; adims[1] = d.seq; adims[2] = d.M / d.seq; adims[3] = 1;
    astr[0] = d.sa_m; astr[1] = (int64_t)d.seq * d.sa_m; astr[2] = astr[1];
    bdims[0] = d.K; bdims[1] = d.N; bdims[2] = 1; bdims[3] = 1;
    bstr[0] = d.sb_n; bstr[1] = d.sb_n; bstr[2] = d.sb_n;
    p.m_tiles_per_seq = ceil_div(d.seq, BLOCK_M);
    p.m_tiles = p.m_tiles_per_seq * (d.M / d.seq);
    p.cblocks = ceil_div(d.cin, BLOCK_K);
    p.k_iters = d.taps * p.cblocks;
  } else if (d.mode == A3T_GEMM_WGRAD) {
    if (nbatch != 1 || d.sa_m != 1 || d.sb_n != 1 || !mul8(d.sa_k) || !mul8(d.sb_k)) return A3T_ERR_UNSUPPORTED;
    p.a_mn = p.b_mn = 1;
    adims[0] = d.M; adims[1] = d.seq; adims[2] = d.K / d.seq; adims[3] = 1;
    astr[0] = d.sa_k; astr[1] = (int64_t)d.seq * d.sa_k; astr[2] = astr[1];
    bdims[0] = d.cin; bdims[1] = d.seq; bdims[2] = d.K / d.seq; bdims[3] = 1;
    bstr[0] = d.sb_k; bstr[1] = (int64_t)d.seq * d.sb_k; bstr[2] = bstr[1];
    p.m_tiles = ceil_div(d.M, BLOCK_M);
    p.m_tiles_per_seq = 1 << 30;  // one "sequence": tile index -> row offset, also for the phantom tile of a CTA pair
    p.cblocks = ceil_div(d.seq, BLOCK_K);
    p.k_iters = (d.K / d.seq) * p.cblocks;
    nlim = d.cin;
  } else {
    if (d.sa_k == 1 && mul8(d.sa_m)) p.a_mn = 0;
    else if (d.sa_m == 1 && mul8(d.sa_k)) p.a_mn = 1;
    else return A3T_ERR_UNSUPPORTED;
    if (d.sb_k == 1 && mul8(d.sb_n)) p.b_mn = 0;
    else if (d.sb_n == 1 && mul8(d.sb_k)) p.b_mn = 1;
    else return A3T_ERR_UNSUPPORTED;
    if (!mul8(d.sa_b1) || !mul8(d.sa_b2) || !mul8(d.sb_b1) || !mul8(d.sb_b2)) return A3T_ERR_UNSUPPORTED;
    const int64_t a_row = p.a_mn ? d.sa_k : d.sa_m, b_row = p.b_mn ? d.sb_k : d.sb_n;
    adims[0] = p.a_mn ? d.M : d.K; adims[1] = p.a_mn ? d.K : d.M;
    bdims[0] = p.b_mn ? d.N : d.K; bdims[1] = p.b_mn ? d.K : d.N;
    astr[0] = a_row; bstr[0] = b_row;
    p.a_c2 = (d.batch2 > 1 && d.sa_b2 != 0); p.a_c3 = (d.batch1 > 1 && d.sa_b1 != 0);
    p.b_c2 = (d.batch2 > 1 && d.sb_b2 != 0); p.b_c3 = (d.batch1 > 1 && d.sb_b1 != 0);
    adims[2] = p.a_c2 ? d.batch2 : 1; astr[1] = p.a_c2 ? d.sa_b2 : a_row;
    adims[3] = p.a_c3 ? d.batch1 : 1; astr[2] = p.a_c3 ? d.sa_b1 : a_row;
    bdims[2] = p.b_c2 ? d.batch2 : 1; bstr[1] = p.b_c2 ? d.sb_b2 : b_row;
    bdims[3] = p.b_c3 ? d.batch1 : 1; bstr[2] = p.b_c3 ? d.sb_b1 : b_row;
    p.m_tiles = ceil_div(d.M, BLOCK_M);
    p.m_tiles_per_seq = 1 << 30;  // one "sequence": tile index -> row offset, also for the phantom tile of a CTA pair
    p.k_iters = ceil_div(d.K, BLOCK_K);
  }
  for (int i = 0; i < 3; i++)
    if (astr[i] <= 0 || bstr[i] <= 0 || astr[i] >= ((int64_t)1 << 38) || bstr[i] >= ((int64_t)1 << 38))
      return A3T_ERR_UNSUPPORTED;
  if (d.dtype_c != A3T_F32 && d.dtype_c != A3T_BF16) return A3T_ERR_UNSUPPORTED;
  if (mask && d.dtype_mask != A3T_F32 && d.dtype_mask != A3T_BF16) return A3T_ERR_UNSUPPORTED;
  if (probe_only) return get_encode() ? A3T_OK : A3T_ERR_UNSUPPORTED;

  // ---- tile shape and split-K: minimise  waves x (K iterations per work item) x (bytes staged per K
  // iteration + a fixed per-iteration cost); the main loop is bound by the shared-memory fill rate ------
  const int sms = num_sms();
  const bool plain_epi = !bias && !res && !mask && d.drop_p == 0.f && !d.relu;
  const bool can_split = d.mode == A3T_GEMM_WGRAD && d.dtype_c == A3T_F32 && plain_epi && d.sc_tap == 1 &&
                         d.sc_n == d.taps && d.sc_m == (int64_t)d.taps * d.cin;
  // WGRAD through TMA store / reduce-add: output rows are contiguous (c, tap) runs of fp32
  // 0 = cost model decides; 1 / 2 = force single-CTA / CTA-pair tiles (d.impl == A3T_IMPL_TC_PAIR, or the
  // A3T_TC_CTA switch of tuning builds)
  const char* env_cta = tune_env("A3T_TC_CTA");
  const int cta_force = env_cta ? atoi(env_cta) : (d.impl == A3T_IMPL_TC_PAIR ? 2 : 0);
  const bool force_pair = cta_force == 2;
  const bool wg_tma = can_split && (d.taps == 1 || (d.taps == 3 && !force_pair)) && ((int64_t)d.cin * d.taps) % 4 == 0 &&
                      al16(C) && !tune_env("A3T_TC_GENERIC_EPI") && !tune_env("A3T_TC_WGRAD_ATOMIC");
  const bool wg3 = wg_tma && d.taps == 3;  // N tile = 3 taps x 64 channels (single-CTA tiles only)
  int best_bn = 0, best_split = 1, best_cta2 = 0;
  double best_cost = 1e30;
  const int cands[5] = {256, 192, 128, 64, ((nlim + 15) / 16) * 16};
  for (int cta2 = 0; cta2 <= 1; cta2++) {
    // measured on B200 (tools/bench_gemm2.py, profiles/r01_gemm_experiments.md): CTA pairs (half the B tile staged per
    // SM) gain 2-6 % on the K >= 1152 implicit-conv shapes (FFN fwd / dgrad) and lose on the small-K, batched and
    // weight-gradient shapes, so they are the default for 3-tap convolutions only
    const bool pair_default = d.mode == A3T_GEMM_CONV && d.taps >= 3 && d.K >= 1024 && p.m_tiles >= 8;
    if (cta_force ? cta_force != cta2 + 1 : (cta2 != 0) != pair_default) continue;
    if (cta2 && p.m_tiles < 2) continue;
    if (cta2 && wg3) continue;
    for (int ci = 0; ci < 5; ci++) {
      int bn = cands[ci];
      if (bn > 256 || bn < 16) continue;
      if (wg3 && bn != 192) continue;
      if (p.b_mn && (bn % (cta2 ? 128 : 64))) continue;
      if (cta2 && (bn % 32)) continue;
      int nt = wg3 ? ceil_div(nlim, 64) : ceil_div(nlim, bn) * (d.mode == A3T_GEMM_WGRAD ? d.taps : 1);
      int64_t units = (int64_t)(cta2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * nt * nbatch;
      int groups = cta2 ? sms / 2 : sms;
      int max_split = can_split ? (p.k_iters / 16 < 1 ? 1 : p.k_iters / 16) : 1;
      if (max_split > 32) max_split = 32;
      for (int sp = 1; sp <= max_split; sp++) {
        int per = (p.k_iters + sp - 1) / sp;
        if (sp > 1 && (int64_t)(sp - 1) * per >= p.k_iters) continue;  // would leave an empty split
        int64_t waves = (units * sp + groups - 1) / groups;
        // per work item: K loop (bytes staged per CTA + fixed cost) + epilogue (~ tile area)
        double kbytes = A_STAGE_BYTES + (cta2 ? bn * 64.0 : bn * 128.0) + 4096.0;
        double cost = (double)waves * ((double)per * kbytes + 128.0 * bn * (sp > 1 ? 40.0 : 20.0));
        if (cost < best_cost * 0.999) { best_cost = cost; best_bn = bn; best_split = sp; best_cta2 = cta2; }
      }
    }
  }
  if (best_bn == 0) return A3T_ERR_UNSUPPORTED;
  if (const char* e = tune_env("A3T_TC_BN")) {  // tuning experiments only
    int v = atoi(e);
    if (v >= 16 && v <= 256 && v % 16 == 0 && !(p.b_mn && v % (best_cta2 ? 128 : 64)) && !(best_cta2 && v % 32)) {
      if (!wg3) {
        best_bn = v;
        best_split = 1;
      }
    }
  }
  const bool cta2 = best_cta2 != 0;
  p.block_n = best_bn;
  p.wg_alltaps = wg3 ? 1 : 0;
  p.n_step = wg3 ? 64 : p.block_n;
  p.n_tiles_per_tap = ceil_div(nlim, p.n_step);
  p.n_tiles = p.n_tiles_per_tap * ((d.mode == A3T_GEMM_WGRAD && !wg3) ? d.taps : 1);
  p.m_units = cta2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  int64_t tiles = (int64_t)p.m_units * p.n_tiles * nbatch;
  if (tiles * best_split > (1 << 30)) return A3T_ERR_UNSUPPORTED;
  p.splits = best_split;
  p.num_work = (int)tiles * p.splits;

  const int b_rows = cta2 ? p.block_n / 2 : p.block_n;
  abox[1] = p.a_mn ? BLOCK_K : BLOCK_M;
  bbox[1] = p.b_mn ? BLOCK_K : b_rows;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_rows * BLOCK_K * 2;
  const int bar_bytes = BAR_BYTES;
  const int epi_bytes = NUM_EPI_WARPS * EPI_STAGE_BYTES;
  p.stages = (SMEM_BYTES_MAX - 1024 - bar_bytes - epi_bytes) / (int)stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (const char* e = tune_env("A3T_TC_STAGES")) {
    int v = atoi(e);
    if (v >= 2 && v < p.stages) p.stages = v;
  }
  if (p.stages < 2) return A3T_ERR_UNSUPPORTED;
  const int smem_bytes = 1024 + p.stages * stage_bytes + epi_bytes + bar_bytes;

  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
            ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)((cta2 ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
  p.bias = bias; p.res = res; p.mask = mask; p.seed = seed; p.C = C;
  const int esc = d.dtype_c == A3T_BF16 ? 8 : 4;  // elements per 16 bytes
  p.vec_c = (d.mode != A3T_GEMM_WGRAD) && d.sc_n == 1 && al16(C) && (d.sc_m % esc) == 0 && (d.sc_b1 % esc) == 0 &&
            (d.sc_b2 % esc) == 0;
  p.vec_r = res && d.sr_n == 1 && al16(res) && (d.sr_m % 4) == 0 && (d.sr_b1 % 4) == 0 && (d.sr_b2 % 4) == 0;
  const int esm = d.dtype_mask == A3T_BF16 ? 8 : 4;
  p.vec_m = mask && (d.mode != A3T_GEMM_WGRAD) && d.sc_n == 1 && al16(mask) && (d.sc_m % esm) == 0 &&
            (d.sc_b1 % esm) == 0 && (d.sc_b2 % esm) == 0;
  if (bias && ((uintptr_t)bias & 15)) return A3T_ERR_UNSUPPORTED;

  CUtensorMap tmA, tmB;
  if (!encode_map(&tmA, A, adims, astr, abox) || !encode_map(&tmB, B, bdims, bstr, bbox)) return A3T_ERR_UNSUPPORTED;

  // ---- epilogue class -------------------------------------------------------------------------
  int epi = -1;
  CUtensorMap tmC = tmA, tmR = tmA;  // placeholders unless the TMA epilogue is selected
  if (wg_tma) {
    int64_t cdims[4] = {(int64_t)d.cin * d.taps, d.M, 1, 1}, cstr[3] = {d.sc_m, d.sc_m, d.sc_m};
    int cbox[4] = {32, 32, 1, 1};
    if (encode_map(&tmC, C, cdims, cstr, cbox, 4)) epi = EPI_WGRAD_TMA;
  }
  if (epi >= 0) {
  } else if (d.mode == A3T_GEMM_WGRAD && d.dtype_c == A3T_F32 && plain_epi && !wg3 && !tune_env("A3T_TC_GENERIC_EPI")) epi = EPI_WGRAD;
  else if (d.mode != A3T_GEMM_WGRAD && p.splits == 1 && p.vec_c && (!res || (p.vec_r && d.dtype_c == A3T_F32)) &&
           (!mask || (p.vec_m && d.dtype_mask == A3T_BF16)) && (d.N % 4) == 0 && !tune_env("A3T_TC_GENERIC_EPI")) {
    // TMA tensor store of C: (N, rows, batch2 | sequence, batch1); superchunks must not straddle N tiles
    const int sc_cols = d.dtype_c == A3T_BF16 ? 64 : 32;
    const int cs = d.dtype_c == A3T_BF16 ? 2 : 4;
    if (p.block_n % sc_cols == 0 || p.n_tiles == 1) {
      int64_t cdims[4], cstr[3];
      int cbox[4] = {sc_cols, 32, 1, 1};
      cdims[0] = d.N;
      cstr[0] = d.sc_m;
      if (d.mode == A3T_GEMM_CONV) {
        cdims[1] = d.seq; cdims[2] = d.M / d.seq; cdims[3] = 1;
        cstr[1] = (int64_t)d.seq * d.sc_m; cstr[2] = cstr[1];
      } else {
        cdims[1] = d.M; cdims[2] = d.batch2; cdims[3] = d.batch1;
        cstr[1] = d.batch2 > 1 ? d.sc_b2 : d.sc_m; cstr[2] = d.batch1 > 1 ? d.sc_b1 : d.sc_m;
      }
      bool ok = true;
      for (int i = 0; i < 3; i++) ok = ok && cstr[i] > 0 && cstr[i] < ((int64_t)1 << 37) && ((cstr[i] * cs) % 16) == 0;
      if (ok && res) {  // residual: same coordinates as C, its own strides (fp32)
        int64_t rstr[3];
        rstr[0] = d.sr_m;
        if (d.mode == A3T_GEMM_CONV) { rstr[1] = (int64_t)d.seq * d.sr_m; rstr[2] = rstr[1]; }
        else { rstr[1] = d.batch2 > 1 ? d.sr_b2 : d.sr_m; rstr[2] = d.batch1 > 1 ? d.sr_b1 : d.sr_m; }
        for (int i = 0; i < 3; i++) ok = ok && rstr[i] > 0 && rstr[i] < ((int64_t)1 << 37) && ((rstr[i] * 4) % 16) == 0;
        ok = ok && encode_map(&tmR, res, cdims, rstr, cbox, 4);
      }
      if (ok && encode_map(&tmC, C, cdims, cstr, cbox, cs))
        epi = (mask ? EPI_MASK : 0) | (d.drop_p > 0.f ? EPI_DROP : 0) | (res ? EPI_RES : 0) |
              (d.dtype_c == A3T_BF16 ? EPI_BF16 : 0);
    }
  }
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
#define A3T_ROW(C2)                                                                                               \
  {gemm_tc_kernel<0, C2>,  gemm_tc_kernel<1, C2>,  gemm_tc_kernel<2, C2>,  gemm_tc_kernel<3, C2>,  gemm_tc_kernel<4, C2>,  \
   gemm_tc_kernel<5, C2>,  gemm_tc_kernel<6, C2>,  gemm_tc_kernel<7, C2>,  gemm_tc_kernel<8, C2>,  gemm_tc_kernel<9, C2>,  \
   gemm_tc_kernel<10, C2>, gemm_tc_kernel<11, C2>, gemm_tc_kernel<-1, C2>, gemm_tc_kernel<-1, C2>, gemm_tc_kernel<-1, C2>, \
   gemm_tc_kernel<-1, C2>, gemm_tc_kernel<EPI_WGRAD, C2>, gemm_tc_kernel<EPI_WGRAD_TMA, C2>, gemm_tc_kernel<-1, C2>}
  static const KernelFn table[2][19] = {A3T_ROW(false), A3T_ROW(true)};
  static bool attr_set[2][19] = {{false}};
  const int ki = epi < 0 ? 18 : epi;
  const KernelFn fn = table[cta2 ? 1 : 0][ki];
  if (!attr_set[cta2 ? 1 : 0][ki]) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_MAX);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return A3T_ERR_CUDA;
    }
    attr_set[cta2 ? 1 : 0][ki] = true;
  }
  if (p.splits > 1 && !d.c_zeroed) {
    cudaError_t e = cudaMemsetAsync(C, 0, (size_t)d.M * d.sc_m * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("gemm_tc: memset: %s", cudaGetErrorString(e));
      return A3T_ERR_CUDA;
    }
  }
  if (const char* e = tune_env("A3T_TC_DBGMODE")) p.dbg = atoi(e);
  // dynamic tile order (cluster launch control): the grid is one CTA (pair) per work item, the hardware launches as
  // many as fit and every running CTA cancels pending ones to take over their index.  Unlike a persistent grid of
  // exactly #SM CTAs this keeps its speed when another kernel (the NCCL all-reduce of finished gradient ranges)
  // occupies some SMs: no CTA waits a whole kernel duration for an SM.
  p.clc = (p.dbg == 0 && !tune_env("A3T_TC_STATIC")) ? 1 : 0;
  if (tune_env("A3T_TC_DEBUG"))
    fprintf(stderr, "gemm_tc: M=%d N=%d K=%d mode=%d cta2=%d bn=%d splits=%d stages=%d work=%d epi=%d\n", d.M, d.N, d.K,
            d.mode, (int)cta2, p.block_n, p.splits, p.stages, p.num_work, epi);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (cta2) {
    int groups = sms / 2;
    int g = (p.clc || p.num_work < groups) ? p.num_work : groups;
    cfg.gridDim = dim3(2 * g);
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = 2;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    nattr++;
  } else {
    cfg.gridDim = dim3((p.clc || p.num_work < sms) ? p.num_work : sms);
  }
  static const bool pdl = !tune_env("A3T_NO_PDL");
  if (pdl && !(p.splits > 1 && !d.c_zeroed)) {  // (the library's own memset node must not be overtaken)
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    nattr++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  {
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmC, tmR, p);
    if (e != cudaSuccess) {
      set_error("gemm_tc: launch: %s", cudaGetErrorString(e));
      return A3T_ERR_CUDA;
    }
  }
  return check_launch("gemm_tc");
}
}  // namespace a3t
