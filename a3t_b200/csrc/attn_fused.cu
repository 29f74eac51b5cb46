// Fused legacy relative-position attention for sm_100a (tcgen05 / TMEM / TMA): the score tensors AC, P and dP of
// transformer/attention.py:167-209 never reach HBM.
//
//   forward  (attn_fwd_kernel):  one CTA per (utterance, head, 128-query tile), loop over 128-key tiles:
//       S = (q+u) K^T on the tensor cores into TMEM (double buffered)  ->  + rel_shift(BD_raw) read with its skew
//       -> key-pad mask, online softmax (lazy rescale), dropout hash  ->  P (bf16) to shared memory  ->
//       O += P V on the tensor cores (accumulator resident in TMEM)  ->  ctx = O / l, lse.
//   backward (attn_bwd_kernel):  same decomposition, 64-key tiles: recompute S and P from lse, dP = dO V^T,
//       dS = P * (dropout'(dP) - delta) * scale;  dQu += dS K on the tensor cores (TMEM resident);  the three
//       operands the remaining batched contractions need are written once each: P_dropped (for dV = Pd^T dO),
//       dS (for dK = dS^T (q+u)) and dBD_raw = rel_shift^T(dS) (for d(q+v) = dBD_raw p, dp = dBD_raw^T (q+v)).
//
// BD_raw = (q+v) p^T is produced by the batched GEMM (gemm_tc.cu) in bf16 and consumed here through
// rel_shift's index map (attention.py:145-165): BD[i,j] = BD_raw[i, S-1-i+j] (j <= i), 0 (j = i+1),
// BD_raw[i+1, j-i-2] (j >= i+2).  The band a score tile needs is a parallelogram (one element of skew per query):
// it arrives by TMA as eight boxes of 16 query rows x (tile width + 24) columns (a TMA box must start 16-byte
// aligned in the innermost dimension, so each box starts up to 7 elements early), and the thread that owns a query
// (the tcgen05.ld layout) re-aligns its row window with a funnel shift.  The backward writes dS tiles by TMA tensor
// store and scatters dBD_raw through the inverse map from the staged tile.
//
// Replaces: attention.py:190-209 (matmul / rel_shift / softmax / dropout / matmul) and its autograd backward.
#include <float.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace a3t {
namespace fa {
using namespace tc;

constexpr int BM = 128;
constexpr int NUM_SM_WARPS = 16;   // softmax warps: (warp & 3) = TMEM lane quarter, (warp >> 2) = column quarter
// three independent TMA producers (K, V, bias): each chain waits only on its own consumer, so a load is issued the
// moment its buffer is free instead of queueing behind the other operands' waits
constexpr int PRODUCER_WARP = NUM_SM_WARPS, V_WARP = NUM_SM_WARPS + 1, BIAS_WARP = NUM_SM_WARPS + 2, MMA_WARP = NUM_SM_WARPS + 3;
constexpr int NUM_THREADS = 32 * (NUM_SM_WARPS + 4);
constexpr int STORE_WARP = NUM_SM_WARPS + 4;   // backward only: TMA tensor stores of dS / Pd
constexpr int NUM_THREADS_BWD = NUM_THREADS + 32;
constexpr int SMEM_MAX = 227 * 1024;

struct Params {
  const __nv_bfloat16* bd_raw;  // (B,H,S,ld) bf16
  const uint8_t* keymask;       // (B,S), 1 = valid key
  const unsigned long long* seed;
  float* lse;                   // (B,H,S) log2-domain log-sum-exp of the scaled scores
  __nv_bfloat16* ctx;           // fwd out / bwd in: (B,S,D)
  const __nv_bfloat16* dctx;    // bwd in: (B,S,D)
  float* delta;                 // bwd workspace: (B,H,S) sum_c dO O per (row, head), written by attn_delta_kernel
  __nv_bfloat16* dq;            // bwd out: dqkv4 base (B,S,4D); d(q+u) goes to columns [h*dk, (h+1)*dk)
  __nv_bfloat16* pd;            // bwd out: (B,H,S,ld) dropped probabilities
  __nv_bfloat16* dbd;           // bwd out: (B,H,S,ld) dBD_raw
  long long* trace;             // tuning builds: per-role clock64 stamps of CTA 0 (NULL = off)
  int64_t ld;                   // row pitch of every (B,H,S,S) tensor, elements
  int B, H, S, D;
  float scale, c2;              // 1/sqrt(dk); scale * log2(e)
  float drop_p;
  uint32_t site;
};

#ifdef A3T_TUNING
#define A3T_TRACE(role, idx) \
  do { if (p.trace && blockIdx.x == 0 && (idx) < 256) p.trace[(role) * 256 + (idx)] = clock64(); } while (0)
#else
#define A3T_TRACE(role, idx) do { } while (0)
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// validity bits of all keys of utterance b into shared memory (word w, bit e = key 32 w + e is a real key); one
// coalesced byte load + ballot per word, words dealt round-robin to the calling warps, loads issued back to back
__device__ __forceinline__ void build_key_bits(const uint8_t* __restrict__ km, int S, uint32_t sKB, int warp, int lane, int nwarps) {
  const int nwords = (S + 31) / 32 + 4;   // tiles read up to 128 keys past the last valid one
  for (int w0 = warp; w0 < nwords; w0 += 4 * nwarps) {
    uint8_t v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = 32 * (w0 + u * nwarps) + lane;
      v[u] = j < S ? km[j] : (uint8_t)0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t bits = __ballot_sync(0xffffffffu, v[u] != 0);
      if (lane == 0 && w0 + u * nwarps < nwords) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sKB + 4 * (w0 + u * nwarps)), "r"(bits) : "memory");
    }
  }
}

// keep flags of COLS consecutive elements starting at element index idx0: word k holds the 16-bit random values of
// elements 2k (low half) and 2k+1 (high half) -- the pair hash of common.cuh re-aligned to an odd start
template <int COLS>
__device__ __forceinline__ void drop_words(const Drop& dr, unsigned long long idx0, uint32_t (&w)[COLS / 2]) {
  const uint32_t f0 = drop_fold(idx0);
  const uint32_t pbase = f0 >> 1;
  const uint32_t sh = (f0 & 1u) * 16u;
  uint32_t hprev = drop_hash(dr, pbase);
#pragma unroll
  for (int k = 0; k < COLS / 2; k++) {
    const uint32_t hnext = drop_hash(dr, pbase + k + 1);
    w[k] = __funnelshift_r(hprev, hnext, sh);
    hprev = hnext;
  }
}

// The rel_shift bias of a (16 queries x COLS keys) sub-tile arrives by TMA as a box of 16 rows x (COLS + 24) columns of
// BD_raw: the band a tile needs is a parallelogram (one element of skew per query), and TMA wants the first column
// of a box 16-byte aligned, so the box starts `sft` (0..7) elements early: row r holds the COLS values of its query
// from element (15 - r + sft) on.  Thread = query reads its window as aligned words and re-aligns it with a funnel
// shift: word k of the result = elements (2k, 2k+1) of the thread's column slice.
__device__ __forceinline__ int box_col0(int S, int i_first, int j0, bool upper) {
  // first column query (i_first + 15) needs: rel_shift reads BD_raw[i, S-1-i+j] below the diagonal (j <= i) and
  // BD_raw[i+1, j-i-2] above it (j >= i+2; the box is then anchored one row lower)
  return upper ? j0 - (i_first + 15) - 2 : S - 1 - (i_first + 15) + j0;
}
template <int COLS, int TCOLS>   // COLS: keys per tile; TCOLS: keys per thread (column slice cs of the tile)
__device__ __forceinline__ void read_bias_window(uint32_t boxes, int sft, int lane, int cs, uint32_t (&out)[TCOLS / 2]) {
  constexpr int PITCH = (COLS + 24) * 2;
  const int r = lane & 15;
  const int e0 = 15 - r + sft + TCOLS * cs;
  const uint32_t a = boxes + (lane >> 4) * (16 * PITCH) + r * PITCH + 4 * (e0 >> 1);
  const uint32_t sh = (uint32_t)(e0 & 1) * 16u;
  uint32_t w[TCOLS / 2 + 1];
#pragma unroll
  for (int k = 0; k <= TCOLS / 2; k++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[k]) : "r"(a + 4 * k) : "memory");
#pragma unroll
  for (int k = 0; k < TCOLS / 2; k++) out[k] = __funnelshift_r(w[k], w[k + 1], sh);
}
// diagonal tile: keys j <= i take the lower band, j >= i + 2 the upper one, j == i + 1 is rel_shift's structural zero
template <int N>
__device__ __forceinline__ void merge_diag(uint32_t (&bw)[N], const uint32_t (&up)[N], int i, int j0) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const int ja = j0 + 2 * k, jb = ja + 1;
    const uint32_t lo16 = ja <= i ? (bw[k] & 0xFFFFu) : (ja == i + 1 ? 0u : (up[k] & 0xFFFFu));
    const uint32_t hi16 = jb <= i ? (bw[k] & 0xFFFF0000u) : (jb == i + 1 ? 0u : (up[k] & 0xFFFF0000u));
    bw[k] = lo16 | hi16;
  }
}

// ================================================================================================
// forward
// ================================================================================================
// Softmax warp (q, cq): q = warp & 3 is the TMEM lane quarter (hardware rule: warp w reads lanes 32 (w % 4) ...),
// cq = warp >> 2 the column quarter; thread = one query row x 32 keys of the 128-key tile.
template <int DK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmBD,
                const __grid_constant__ Params p) {
  A3T_PDL_TRIGGER();
  constexpr int NC = DK / 64;                  // 64-wide chunks of the head dimension
  constexpr uint32_t QB = NC * 16384u;         // one 128 x DK bf16 operand tile
  constexpr uint32_t BOX = 16 * 152 * 2;       // bias box of one 16-query group (128 keys + 15 of skew + 7 of alignment)
  constexpr int QCOLS = DK / 4;                // O columns per thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + QB, sV = sK + QB, sP = sV + QB, sB = sP + 32768u, sRed = sB + 8 * BOX, sKB = sRed + 4096u,
                 sBar = sKB + 1024u;
  const uint32_t q_full = sBar, k_full = sBar + 8, k_empty = sBar + 16, v_full = sBar + 24, v_empty = sBar + 32,
                 p_full = sBar + 72, p_empty = sBar + 80, o_full = sBar + 88, tmem_slot = sBar + 96, b_full = sBar + 104,
                 b_empty = sBar + 112;
  auto s_full = [&](int i) { return sBar + 40u + 8u * i; };
  auto s_empty = [&](int i) { return sBar + 56u + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, H = p.H, D = p.D;
  const int nqt = (S + BM - 1) / BM, nkt = nqt;
  const int qt = blockIdx.x % nqt, bh = blockIdx.x / nqt;
  const int h = bh % H, b = bh / H;
  const int i0 = qt * BM;

  if (warp == PRODUCER_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBD) : "memory");
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full(0), 1); mbar_init(s_full(1), 1); mbar_init(s_empty(0), NUM_SM_WARPS); mbar_init(s_empty(1), NUM_SM_WARPS);
    mbar_init(p_full, NUM_SM_WARPS); mbar_init(p_empty, 1); mbar_init(o_full, 1);
    mbar_init(b_full, 1); mbar_init(b_empty, NUM_SM_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  A3T_PDL_WAIT();
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      // operand tiles are stored as 8 KB boxes of 64 rows x 64 columns (128-byte swizzled rows)
      mbar_expect_tx(q_full, QB);
      for (int c = 0; c < NC; c++)
        for (int r = 0; r < 2; r++) tma_load_4d(sQ + c * 16384 + r * 8192, &tmQKV, q_full, h * DK + 64 * c, i0 + 64 * r, b, 0);
      for (int t = 0; t < nkt; t++) {
        const int j0 = t * 128;
        if (t > 0) mbar_wait(k_empty, (t - 1) & 1);
        mbar_expect_tx(k_full, QB);
        for (int c = 0; c < NC; c++)
          for (int r = 0; r < 2; r++)
            tma_load_4d(sK + c * 16384 + r * 8192, &tmQKV, k_full, 2 * D + h * DK + 64 * c, j0 + 64 * r, b, 0);
        A3T_TRACE(1, t);
      }
    }
  } else if (warp == V_WARP) {
    if (elect_one()) {
      for (int t = 0; t < nkt; t++) {
        const int j0 = t * 128;
        if (t > 0) mbar_wait(v_empty, (t - 1) & 1);
        mbar_expect_tx(v_full, QB);
        for (int kc = 0; kc < 2; kc++)   // V as the MN-major B operand of P V: per 64-key chunk, NC atoms of 64 head columns
          for (int a = 0; a < NC; a++)
            tma_load_4d(sV + kc * (NC * 8192) + a * 8192, &tmQKV, v_full, 3 * D + h * DK + 64 * a, j0 + 64 * kc, b, 0);
        A3T_TRACE(2, t);
      }
    }
  } else if (warp == BIAS_WARP) {
    if (elect_one()) {
      int nb = 0;   // bias loads issued
      auto load_bias = [&](int j0, bool upper) {
        if (nb > 0) mbar_wait(b_empty, (nb - 1) & 1);
        mbar_expect_tx(b_full, 8 * BOX);
        for (int g = 0; g < 8; g++) {
          const int i_first = i0 + 16 * g;
          tma_load_4d(sB + g * BOX, &tmBD, b_full, box_col0(S, i_first, j0, upper) & ~7, i_first + (upper ? 1 : 0), h, b);
        }
        A3T_TRACE(3, nb);
        nb++;
      };
      for (int t = 0; t < nkt; t++) {
        load_bias(t * 128, t > qt);              // t < qt: below the diagonal; t == qt: the lower part first
        if (t == qt) load_bias(t * 128, true);   // the diagonal tile also needs the part above the diagonal
      }
    }
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DK >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
      const uint64_t dQ = make_smem_desc(sQ, 16), dK = make_smem_desc(sK, 16), dP = make_smem_desc(sP, 16);
      const uint64_t dV = make_smem_desc(sV, 8192);
      auto issue_s = [&](int t) {
        mbar_wait(k_full, t & 1);
        if (t >= 2) mbar_wait(s_empty(t & 1), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + 128u * (t & 1);
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(d, dQ + (uint64_t)((c * 16384 + k * 32) >> 4), dK + (uint64_t)((c * 16384 + k * 32) >> 4), idesc_s,
                      (c | k) ? 1u : 0u);
        umma_commit(k_empty);
        umma_commit(s_full(t & 1));
        A3T_TRACE(4, t);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int t = 0; t < nkt; t++) {
        if (t + 1 < nkt) issue_s(t + 1);
        mbar_wait(p_full, t & 1);
        mbar_wait(v_full, t & 1);
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < 2; kc++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(tmem_O, dP + (uint64_t)((kc * 16384 + k * 32) >> 4), dV + (uint64_t)((kc * (NC * 8192) + k * 2048) >> 4),
                      idesc_o, (t | kc | k) ? 1u : 0u);
        umma_commit(v_empty);
        umma_commit(p_empty);
        A3T_TRACE(5, t);
      }
      umma_commit(o_full);
    }
  } else {
    // ===================================== softmax warps ====================================
    const int q = warp & 3, cq = warp >> 2;
    const int row = q * 32 + lane, i = i0 + row;
    const bool tr = warp == 0 && lane == 0;
    const uint8_t* km = p.keymask + (int64_t)b * S;
    build_key_bits(km, S, sKB, warp, lane, NUM_SM_WARPS);
    named_bar(5, NUM_SM_WARPS * 32);
    const Drop dr = make_drop(p.drop_p, p.seed, p.site);
    const unsigned long long drow = ((unsigned long long)bh * S + (unsigned long long)min(i, S - 1)) * (unsigned long long)S;
    const uint32_t lane_t = ((uint32_t)(q * 32) << 16);
    // running maximum m in RAW score units (before the 1/sqrt(dk) scale); l accumulates exp2((x - m) c2) * inv_keep:
    // the dropout scale is folded into the exponent, so kept probabilities need no multiply
    float m = -INFINITY, l = 0.f;
    const float c2 = p.c2;
    const float lk = dr.on ? log2f(dr.inv_keep) : 0.f;
    const float thr_raw = 8.f / c2;
    int nb = 0;   // bias boxes consumed
    auto take_bias = [&](uint32_t (&w)[16], int jt, bool upper) {
      mbar_wait(b_full, nb & 1);
      read_bias_window<128, 32>(sB + 2 * q * BOX, box_col0(S, i0, jt, upper) & 7, lane, cq, w);
      __syncwarp();
      if (lane == 0) mbar_arrive(b_empty);
      nb++;
    };
    for (int t = 0; t < nkt; t++) {
      const int j0 = t * 128 + 32 * cq;
      uint32_t bw[16], kb;
      if (tr) A3T_TRACE(0, 8 * t);
      take_bias(bw, t * 128, t > qt);
      if (t == qt) {
        uint32_t up[16];
        take_bias(up, t * 128, true);
        merge_diag<16>(bw, up, i, j0);
      } else if (j0 == i + 1) {   // first key of the tile right of the diagonal, last query of the tile
        bw[0] &= 0xFFFF0000u;
      }
      if (tr) A3T_TRACE(0, 8 * t + 1);
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kb) : "r"(sKB + (uint32_t)(j0 >> 5) * 4u) : "memory");
      mbar_wait(s_full(t & 1), (t >> 1) & 1);
      tc_fence_after();
      if (tr) A3T_TRACE(0, 8 * t + 2);
      uint32_t s[32];
      tmem_ld32(tmem_base + 128u * (t & 1) + lane_t + 32u * cq, s);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(t & 1));
      float mx = -INFINITY;
      if (kb == 0xFFFFFFFFu) {
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const float x = __uint_as_float(s[c]) + ((c & 1) ? bf_hi(bw[c >> 1]) : bf_lo(bw[c >> 1]));
          s[c] = __float_as_uint(x);
          mx = fmaxf(mx, x);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c++) {
          float x = __uint_as_float(s[c]) + ((c & 1) ? bf_hi(bw[c >> 1]) : bf_lo(bw[c >> 1]));
          x = ((kb >> c) & 1u) ? x : -INFINITY;
          s[c] = __float_as_uint(x);
          mx = fmaxf(mx, x);
        }
      }
      // the four threads of a row (warps q, q+4, q+8, q+12) agree on the tile maximum
      const uint32_t red = sRed + (uint32_t)(t & 1) * 2048u + row * 4;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + cq * 512), "f"(mx) : "memory");
      named_bar(1 + q, 128);
      if (tr) A3T_TRACE(0, 8 * t + 3);
      float m_new = m;
#pragma unroll
      for (int o = 0; o < 4; o++) {
        float other;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(red + o * 512) : "memory");
        m_new = fmaxf(m_new, other);
      }
      // lazy rescale: the running maximum only moves when it grows by more than 2^8 (P stays <= 256: exact enough in
      // bf16 / fp32 and the accumulator in TMEM is rarely touched)
      const bool resc = m_new > m + thr_raw;
      float fac = 1.f;
      if (resc) {
        fac = ex2((m - m_new) * c2);   // m = -inf on the first tile: 0
        l *= fac;
        m = m_new;
      }
      const float off = lk - ((m == -INFINITY) ? 0.f : m) * c2;
      uint32_t rw[16];
      if (dr.on) drop_words<32>(dr, drow + (unsigned long long)j0, rw);
      uint32_t pk[16];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        float p0 = ex2(fmaf(__uint_as_float(s[2 * k]), c2, off)), p1 = ex2(fmaf(__uint_as_float(s[2 * k + 1]), c2, off));
        l += p0 + p1;
        if (dr.on) {
          p0 = ((rw[k] & 0xFFFFu) >= dr.thr) ? p0 : 0.f;
          p1 = ((rw[k] >> 16) >= dr.thr) ? p1 : 0.f;
        }
        pk[k] = pack_bf16(p0, p1);
      }
      if (tr) A3T_TRACE(0, 8 * t + 4);
      if (t > 0) {
        mbar_wait(p_empty, (t - 1) & 1);   // P V of the previous tile has read the P buffer and updated O
        if (__any_sync(0xffffffffu, resc)) {
          tc_fence_after();
          const uint32_t ta = tmem_O + lane_t + (uint32_t)(cq * QCOLS);
          if constexpr (QCOLS >= 32) {
            uint32_t o[32];
            tmem_ld32(ta, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; c++) o[c] = __float_as_uint(__uint_as_float(o[c]) * fac);
            tmem_st32(ta, o);
          }
          if constexpr (QCOLS % 32 == 16) {
            uint32_t o[16];
            tmem_ld16(ta + (QCOLS - 16), o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; c++) o[c] = __float_as_uint(__uint_as_float(o[c]) * fac);
            tmem_st16(ta + (QCOLS - 16), o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }
      if (tr) A3T_TRACE(0, 8 * t + 5);
      // 32 keys = 64 bytes = chunks 4 (cq & 1) .. +3 of the row in the K-major operand tile of keys 64 (cq >> 1) ..
      const uint32_t prow = sP + (cq >> 1) * 16384 + row * 128;
      const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
      for (int u = 0; u < 4; u++)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (((4 * (cq & 1) + u) ^ sw) << 4)), "r"(pk[4 * u]),
                     "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                     : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr) A3T_TRACE(0, 8 * t + 6);
    }
    // ---- epilogue: ctx = O / l, lse ----
    const uint32_t red = sRed + (uint32_t)(nkt & 1) * 2048u + row * 4;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + cq * 512), "f"(l) : "memory");
    named_bar(1 + q, 128);
    float ltot = 0.f;
#pragma unroll
    for (int o = 0; o < 4; o++) {
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(red + o * 512) : "memory");
      ltot += other;
    }
    // ltot carries the folded dropout scale: sum of probabilities = ltot / inv_keep
    const float inv = ltot > 0.f ? dr.inv_keep / ltot : 0.f;
    mbar_wait(o_full, 0);
    tc_fence_after();
    __nv_bfloat16* crow = p.ctx + ((int64_t)b * S + i) * D + h * DK + cq * QCOLS;
    const uint32_t ta = tmem_O + lane_t + (uint32_t)(cq * QCOLS);
    auto store8 = [&](const uint32_t* o, int col) {
      uint4 v;
      v.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      v.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      v.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      v.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      *reinterpret_cast<uint4*>(crow + col) = v;
    };
    if constexpr (QCOLS >= 32) {
      uint32_t o[32];
      tmem_ld32(ta, o);
      tmem_ld_wait();
      if (i < S) {
#pragma unroll
        for (int u = 0; u < 4; u++) store8(&o[8 * u], 8 * u);
      }
    }
    if constexpr (QCOLS % 32 == 16) {
      uint32_t o[16];
      tmem_ld16(ta + (QCOLS - 16), o);
      tmem_ld_wait();
      if (i < S) {
        store8(&o[0], QCOLS - 16);
        store8(&o[8], QCOLS - 8);
      }
    }
    if (cq == 0 && i < S) p.lse[(int64_t)bh * S + i] = ltot > 0.f ? m * c2 + log2f(ltot) - lk : 1e30f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ================================================================================================
// backward
// ================================================================================================
// 64-key tiles.  Shared memory: Qu and dO resident (2 x 128 x DK), K double buffered and V single buffered (3 x 64 x DK),
// the dS operand tile and the Pd staging tile (2 x 128 x 64), the bias boxes.  TMEM: S and dP double buffered
// (4 x 64 columns), dQu (DK).  STORE_WARP issues the TMA tensor stores of dS and Pd.  Softmax warp (q, cq): one query
// row x 16 keys per thread.
template <int DK>
__global__ void __launch_bounds__(NUM_THREADS_BWD, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmBD, const __grid_constant__ CUtensorMap tmDS,
                const __grid_constant__ CUtensorMap tmPD, const __grid_constant__ Params p) {
  A3T_PDL_TRIGGER();
  constexpr int NC = DK / 64;
  constexpr uint32_t QB = NC * 16384u;         // 128 x DK
  constexpr uint32_t KB = NC * 8192u;          // 64 x DK
  constexpr uint32_t BOX = 16 * 88 * 2;        // bias box of one 16-query group (64 keys + 15 of skew + 7 of alignment)
  constexpr int QCOLS = DK / 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sDO = sQ + QB, sK = sDO + QB, sV = sK + 2 * KB, sDS = sV + KB, sPD = sDS + 16384u, sB = sPD + 16384u,
                 sRed = sDS /* used once, before the first tile */, sKB = sB + 8 * BOX, sBar = sKB + 256u;
  const uint32_t q_full = sBar, out_full = sBar + 8, out_empty = sBar + 16, o_full = sBar + 24, tmem_slot = sBar + 32,
                 v_full = sBar + 40, v_empty = sBar + 48, b_full = sBar + 56, b_empty = sBar + 64;
  auto k_full = [&](int i) { return sBar + 72u + 8u * i; };
  auto k_empty = [&](int i) { return sBar + 88u + 8u * i; };
  auto s_full = [&](int i) { return sBar + 104u + 8u * i; };
  auto s_empty = [&](int i) { return sBar + 120u + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, H = p.H, D = p.D;
  const int nqt = (S + BM - 1) / BM, nkt = (S + 63) / 64;
  const int qt = blockIdx.x % nqt, bh = blockIdx.x / nqt;
  const int h = bh % H, b = bh / H;
  const int i0 = qt * BM;
  // key tiles 2 qt and 2 qt + 1 straddle the diagonal of this query tile (both bands of BD_raw are needed)
  auto is_diag = [&](int t) { return (t >> 1) == qt; };

  if (warp == PRODUCER_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmDO) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBD) : "memory");
    mbar_init(q_full, 1); mbar_init(out_full, NUM_SM_WARPS); mbar_init(out_empty, 2 + NUM_SM_WARPS); mbar_init(o_full, 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 1); mbar_init(b_full, 1); mbar_init(b_empty, NUM_SM_WARPS);
    for (int i = 0; i < 2; i++) {
      mbar_init(k_full(i), 1); mbar_init(k_empty(i), 1); mbar_init(s_full(i), 1); mbar_init(s_empty(i), NUM_SM_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == STORE_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmDS) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmPD) : "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  A3T_PDL_WAIT();
  // TMEM columns: S buffers at 0 / 64, dP buffers at 128 / 192, dQu at 256
  const uint32_t tmem_dQ = tmem_base + 256;

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * QB);
      for (int c = 0; c < NC; c++)
        for (int r = 0; r < 2; r++) {
          tma_load_4d(sQ + c * 16384 + r * 8192, &tmQKV, q_full, h * DK + 64 * c, i0 + 64 * r, b, 0);
          tma_load_4d(sDO + c * 16384 + r * 8192, &tmDO, q_full, h * DK + 64 * c, i0 + 64 * r, b, 0);
        }
      for (int t = 0; t < nkt; t++) {
        const int st = t & 1, j0 = t * 64;
        if (t >= 2) mbar_wait(k_empty(st), ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full(st), KB);
        for (int c = 0; c < NC; c++) tma_load_4d(sK + st * KB + c * 8192, &tmQKV, k_full(st), 2 * D + h * DK + 64 * c, j0, b, 0);
        A3T_TRACE(1, t);
      }
    }
  } else if (warp == V_WARP) {
    if (elect_one()) {
      for (int t = 0; t < nkt; t++) {
        if (t > 0) mbar_wait(v_empty, (t - 1) & 1);
        mbar_expect_tx(v_full, KB);
        for (int c = 0; c < NC; c++) tma_load_4d(sV + c * 8192, &tmQKV, v_full, 3 * D + h * DK + 64 * c, t * 64, b, 0);
        A3T_TRACE(2, t);
      }
    }
  } else if (warp == BIAS_WARP) {
    if (elect_one()) {
      int nb = 0;
      auto load_bias = [&](int j0, bool upper) {
        if (nb > 0) mbar_wait(b_empty, (nb - 1) & 1);
        mbar_expect_tx(b_full, 8 * BOX);
        for (int g = 0; g < 8; g++) {
          const int i_first = i0 + 16 * g;
          tma_load_4d(sB + g * BOX, &tmBD, b_full, box_col0(S, i_first, j0, upper) & ~7, i_first + (upper ? 1 : 0), h, b);
        }
        A3T_TRACE(3, nb);
        nb++;
      };
      for (int t = 0; t < nkt; t++) {
        load_bias(t * 64, (t >> 1) > qt);
        if (is_diag(t)) load_bias(t * 64, true);
      }
    }
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_q = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DK >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
      const uint64_t dQ = make_smem_desc(sQ, 16), dDO = make_smem_desc(sDO, 16), dDS = make_smem_desc(sDS, 16);
      const uint64_t dV = make_smem_desc(sV, 16);
      auto issue_s = [&](int t) {
        const int st = t & 1;
        mbar_wait(k_full(st), (t >> 1) & 1);
        if (t >= 2) mbar_wait(s_empty(st), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dK = make_smem_desc(sK + st * KB, 16);
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(tmem_base + 64u * st, dQ + (uint64_t)((c * 16384 + k * 32) >> 4), dK + (uint64_t)((c * 8192 + k * 32) >> 4),
                      idesc_s, (c | k) ? 1u : 0u);
        mbar_wait(v_full, t & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(tmem_base + 128u + 64u * st, dDO + (uint64_t)((c * 16384 + k * 32) >> 4),
                      dV + (uint64_t)((c * 8192 + k * 32) >> 4), idesc_s, (c | k) ? 1u : 0u);
        umma_commit(v_empty);
        umma_commit(s_full(st));
        A3T_TRACE(4, t);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int t = 0; t < nkt; t++) {
        if (t + 1 < nkt) issue_s(t + 1);
        mbar_wait(out_full, t & 1);
        tc_fence_after();
        // dQu += dS (128 x 64 keys, K-major) * K (64 keys x DK: the K tile read MN-major, NC atoms 8 KB apart)
        const uint64_t dKm = make_smem_desc(sK + (t & 1) * KB, 8192);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_bf16(tmem_dQ, dDS + (uint64_t)((k * 32) >> 4), dKm + (uint64_t)((k * 2048) >> 4), idesc_q, (t | k) ? 1u : 0u);
        umma_commit(k_empty(t & 1));
        umma_commit(out_empty);
        A3T_TRACE(5, t);
      }
      umma_commit(o_full);
    }
  } else if (warp == STORE_WARP) {
    // dS (for dK = dS^T (q+u)) and Pd (for dV = Pd^T dO) leave through TMA tensor stores of the two staging tiles
    if (elect_one()) {
      for (int t = 0; t < nkt; t++) {
        mbar_wait(out_full, t & 1);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)&tmDS),
                     "r"(sDS), "r"(t * 64), "r"(i0), "r"(h), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)&tmPD),
                     "r"(sPD), "r"(t * 64), "r"(i0), "r"(h), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(out_empty);
        A3T_TRACE(6, t);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const int q = warp & 3, cq = warp >> 2;
    const int row = q * 32 + lane, i = i0 + row;
    const bool row_ok = i < S;
    const bool tr = warp == 0 && lane == 0;
    const uint8_t* km = p.keymask + (int64_t)b * S;
    build_key_bits(km, S, sKB, warp, lane, NUM_SM_WARPS);
    // dBD_raw = rel_shift^T(dS) is written from the STAGED dS tile with lanes along the key axis (rows of dBD_raw shift
    // by one element per query, so no aligned vector store exists: 2-byte elements, 64 contiguous bytes per row and
    // instruction); this warp scatters query rows 8 warp .. +7 of every tile
    const int ld = (int)p.ld;
    unsigned short* const dbd = reinterpret_cast<unsigned short*>(p.dbd + (int64_t)bh * S * p.ld);
    const int dl = ld - S - 1;                              // lower band -> upper band offset
    if (qt == 0)                                            // BD_raw[0, 0 .. S-2]: the row the reshape drops, gradient 0
      for (int j = threadIdx.x; j < S - 1; j += NUM_SM_WARPS * 32) dbd[j] = 0;
    const Drop dr = make_drop(p.drop_p, p.seed, p.site);
    const unsigned long long drow = ((unsigned long long)bh * S + (unsigned long long)min(i, S - 1)) * (unsigned long long)S;
    const uint32_t lane_t = ((uint32_t)(q * 32) << 16);
    const float c2 = p.c2;
    named_bar(5, NUM_SM_WARPS * 32);   // key bits are visible
    // delta_i = sum_c dO[i,c] O[i,c] over the head (attn_delta_kernel) and the row's log-sum-exp
    const float delta = row_ok ? p.delta[(int64_t)bh * S + i] : 0.f;
    const float lse = row_ok ? p.lse[(int64_t)bh * S + i] : 1e30f;
    // dS = P (dropout'(dP) - delta) scale with the scales folded: dS = P * fma(dP, ks, -ds) (kept) or P * (-ds) (dropped)
    const float ks = dr.inv_keep * p.scale, dsn = -delta * p.scale;
    int nb = 0;
    auto take_bias = [&](uint32_t (&w)[8], int jt, bool upper) {
      mbar_wait(b_full, nb & 1);
      read_bias_window<64, 16>(sB + 2 * q * BOX, box_col0(S, i0, jt, upper) & 7, lane, cq, w);
      __syncwarp();
      if (lane == 0) mbar_arrive(b_empty);
      nb++;
    };
    for (int t = 0; t < nkt; t++) {
      const int st = t & 1;
      const int j0 = t * 64 + 16 * cq;
      uint32_t bw[8], kb;
      if (tr) A3T_TRACE(0, 8 * t);
      take_bias(bw, t * 64, (t >> 1) > qt);
      if (is_diag(t)) {
        uint32_t up[8];
        take_bias(up, t * 64, true);
        merge_diag<8>(bw, up, i, j0);
      } else if (j0 == i + 1) {
        bw[0] &= 0xFFFF0000u;
      }
      if (tr) A3T_TRACE(0, 8 * t + 1);
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kb) : "r"(sKB + (uint32_t)(j0 >> 5) * 4u) : "memory");
      kb = (kb >> (j0 & 31)) & 0xFFFFu;
      mbar_wait(s_full(st), (t >> 1) & 1);
      tc_fence_after();
      if (tr) A3T_TRACE(0, 8 * t + 2);
      uint32_t s[16], dp[16];
      tmem_ld16(tmem_base + 64u * st + lane_t + 16u * cq, s);
      tmem_ld16(tmem_base + 128u + 64u * st + lane_t + 16u * cq, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(st));
      uint32_t rw[8];
      if (dr.on) drop_words<16>(dr, drow + (unsigned long long)j0, rw);
      uint32_t pdw[8], dsw[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float pv[2], dv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = 2 * k + e;
          const float x = __uint_as_float(s[c]) + (e ? bf_hi(bw[k]) : bf_lo(bw[k]));
          const bool valid = ((kb >> c) & 1u) != 0;
          const float P = valid ? ex2(fmaf(x, c2, -lse)) : 0.f;
          const float g = __uint_as_float(dp[c]);
          bool keep = true;
          if (dr.on) keep = (e ? (rw[k] >> 16) : (rw[k] & 0xFFFFu)) >= dr.thr;
          pv[e] = keep ? (dr.on ? P * dr.inv_keep : P) : 0.f;
          dv[e] = P * (keep ? fmaf(g, ks, dsn) : dsn);
        }
        pdw[k] = pack_bf16(pv[0], pv[1]);
        dsw[k] = pack_bf16(dv[0], dv[1]);
      }
      if (tr) A3T_TRACE(0, 8 * t + 3);
      // ---- outputs.  The staging tiles are free once the previous tile's dQu MMA and tensor stores have read them ----
      if (t > 0) mbar_wait(out_empty, (t - 1) & 1);
      if (tr) A3T_TRACE(0, 8 * t + 4);
      const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t off = row * 128 + (((2 * cq + u) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(dsw[4 * u]), "r"(dsw[4 * u + 1]),
                     "r"(dsw[4 * u + 2]), "r"(dsw[4 * u + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPD + off), "r"(pdw[4 * u]), "r"(pdw[4 * u + 1]),
                     "r"(pdw[4 * u + 2]), "r"(pdw[4 * u + 3])
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_full);
      if (tr) A3T_TRACE(0, 8 * t + 5);
      mbar_wait(out_full, t & 1);   // every warp's rows are staged
#pragma unroll
      for (int half = 0; half < 2; half++) {                // keys 32 half .. +31 of the tile
        const int j = t * 64 + 32 * half + lane;
        const bool j_ok = j < S;
        unsigned short* const pl = dbd + j;
        const uint32_t cb = (uint32_t)(2 * (32 * half + lane));   // byte offset of the key inside a 128-byte row
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int rowl = 8 * warp + r, ii = i0 + rowl;
          unsigned short v;
          asm volatile("ld.shared.u16 %0, [%1];"
                       : "=h"(v)
                       : "r"(sDS + rowl * 128 + ((((cb >> 4) ^ (uint32_t)(rowl & 7)) << 4) | (cb & 15u)))
                       : "memory");
          if (j_ok && ii < S && j != ii + 1) pl[ii * (ld - 1) + S - 1 + (j > ii ? dl : 0)] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(out_empty);
      if (tr) A3T_TRACE(0, 8 * t + 6);
    }
    // ---- epilogue: d(q+u) -> dqkv4[:, :, h*dk ...] ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    __nv_bfloat16* qrow = p.dq + ((int64_t)b * S + i) * (4 * (int64_t)D) + h * DK + cq * QCOLS;
    const uint32_t ta = tmem_dQ + lane_t + (uint32_t)(cq * QCOLS);
    auto store8 = [&](const uint32_t* o, int col) {
      uint4 v;
      v.x = pack_bf16(__uint_as_float(o[0]), __uint_as_float(o[1]));
      v.y = pack_bf16(__uint_as_float(o[2]), __uint_as_float(o[3]));
      v.z = pack_bf16(__uint_as_float(o[4]), __uint_as_float(o[5]));
      v.w = pack_bf16(__uint_as_float(o[6]), __uint_as_float(o[7]));
      *reinterpret_cast<uint4*>(qrow + col) = v;
    };
    if constexpr (QCOLS >= 32) {
      uint32_t o[32];
      tmem_ld32(ta, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int u = 0; u < 4; u++) store8(&o[8 * u], 8 * u);
      }
    }
    if constexpr (QCOLS % 32 == 16) {
      uint32_t o[16];
      tmem_ld16(ta + (QCOLS - 16), o);
      tmem_ld_wait();
      if (row_ok) {
        store8(&o[0], QCOLS - 16);
        store8(&o[8], QCOLS - 8);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// delta[b,h,i] = sum_c dO[b,i,h*dk+c] * O[b,i,h*dk+c]: one warp per (b, i), lanes along the row
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ ctx,
                                                         float* __restrict__ delta, int B, int H, int S, int D) {
  A3T_PDL_TRIGGER();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= (int64_t)B * S) return;
  const int dk = D / H;
  const uint4* a = reinterpret_cast<const uint4*>(dctx + r * D);
  const uint4* o = reinterpret_cast<const uint4*>(ctx + r * D);
  const int b = (int)(r / S), i = (int)(r - (int64_t)b * S);
  for (int h = 0; h < H; h++) {
    float acc = 0.f;
    for (int u = lane; u < dk / 8; u += 32) {
      const uint4 x = __ldg(a + h * (dk / 8) + u), y = __ldg(o + h * (dk / 8) + u);
      acc += bf_lo(x.x) * bf_lo(y.x) + bf_hi(x.x) * bf_hi(y.x) + bf_lo(x.y) * bf_lo(y.y) + bf_hi(x.y) * bf_hi(y.y) +
             bf_lo(x.z) * bf_lo(y.z) + bf_hi(x.z) * bf_hi(y.z) + bf_lo(x.w) * bf_lo(y.w) + bf_hi(x.w) * bf_hi(y.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) delta[((int64_t)b * H + h) * S + i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool qkv_map(CUtensorMap* m, const void* base, int B, int S, int row_elems) {
  const int64_t dims[4] = {row_elems, S, B, 1}, str[3] = {row_elems, (int64_t)S * row_elems, (int64_t)S * row_elems};
  const int box[4] = {64, 64, 1, 1};
  return encode_map(m, base, dims, str, box);
}

template <typename K>
static int set_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return A3T_ERR_CUDA;
  }
  return A3T_OK;
}

static bool score_map(CUtensorMap* m, const void* base, int B, int H, int S, int64_t ld, int box_cols, int box_rows, bool swz) {
  const int64_t dims[4] = {ld, S, H, B}, str[3] = {ld, (int64_t)S * ld, (int64_t)H * S * ld};
  const int box[4] = {box_cols, box_rows, 1, 1};
  return encode_map(m, base, dims, str, box, 2, swz);
}

static int launch(const void* fn, int grid, int threads, int smem, cudaStream_t st, void** args, const char* what) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) {
    set_error("%s: launch: %s", what, cudaGetErrorString(e));
    return A3T_ERR_CUDA;
  }
  return check_launch(what);
}

}  // namespace fa
}  // namespace a3t

using namespace a3t;

static long long* g_trace = nullptr;   // tuning builds only (a3t_attn_set_trace)
extern "C" int a3t_attn_set_trace(void* device_buffer) {
  g_trace = (long long*)device_buffer;
  return A3T_OK;
}

extern "C" int a3t_attn_fused_supported(int B, int H, int S, int D) {
  if (B < 1 || H < 1 || S < 1 || D < 1 || D % H) return 0;
  const int dk = D / H;
  if (dk != 64 && dk != 128 && dk != 192) return 0;
  if ((int64_t)B * H * S * S >= ((int64_t)1 << 32)) return 0;   // dropout element index stays 32-bit inside a row block
  return tc::get_encode() ? 1 : 0;
}

static int check_common(const char* what, const void* qkv4, const void* bd_raw, const uint8_t* keymask, int B, int H, int S,
                        int D, int64_t ld, float drop_p, const unsigned long long* seed) {
  A3T_REQUIRE(qkv4 && bd_raw && keymask, "%s: null pointer", what);
  A3T_REQUIRE(a3t_attn_fused_supported(B, H, S, D), "%s: unsupported shape (B=%d H=%d S=%d D=%d; dk must be 64/128/192)", what, B,
              H, S, D);
  A3T_REQUIRE(ld >= S, "%s: row pitch %lld < S", what, (long long)ld);
  A3T_REQUIRE((((uintptr_t)qkv4) & 15) == 0 && (D % 8) == 0, "%s: qkv4 must be 16-byte aligned", what);
  A3T_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed), "%s: bad dropout arguments", what);
  return A3T_OK;
}

extern "C" int a3t_relpos_attn_fwd(const void* qkv4, const void* bd_raw, int64_t ld, const uint8_t* keymask, void* ctx,
                                   float* lse, int B, int H, int S, int D, float scale, float drop_p,
                                   const unsigned long long* seed, uint32_t site, void* stream) {
  using namespace fa;
  int rc = check_common("relpos_attn_fwd", qkv4, bd_raw, keymask, B, H, S, D, ld, drop_p, seed);
  if (rc) return rc;
  A3T_REQUIRE(ctx && lse && (((uintptr_t)ctx) & 15) == 0, "relpos_attn_fwd: ctx / lse");
  const int dk = D / H;
  A3T_REQUIRE((ld % 8) == 0 && (((uintptr_t)bd_raw) & 15) == 0, "relpos_attn_fwd: bd_raw pitch %% 8 / alignment");
  CUtensorMap tm, tmBD;
  A3T_REQUIRE(qkv_map(&tm, qkv4, B, S, 4 * D) && score_map(&tmBD, bd_raw, B, H, S, ld, 152, 16, false),
              "relpos_attn_fwd: tensor map");
  Params p;
  memset(&p, 0, sizeof(p));
  p.bd_raw = (const __nv_bfloat16*)bd_raw; p.keymask = keymask; p.seed = seed; p.lse = lse; p.ctx = (__nv_bfloat16*)ctx;
  p.ld = ld; p.B = B; p.H = H; p.S = S; p.D = D; p.scale = scale; p.c2 = scale * 1.4426950408889634f; p.drop_p = drop_p;
  p.site = site;
  p.trace = g_trace;
  const int nc = dk / 64;
  A3T_REQUIRE(S <= 8000, "relpos_attn_fwd: S=%d exceeds the key-bit table", S);
  const int smem = 1024 + 3 * nc * 16384 + 32768 + 8 * 16 * 152 * 2 + 4096 + 1024 + 128;
  const int grid = B * H * ((S + BM - 1) / BM);
  void* args[3] = {&tm, &tmBD, &p};
  const void* fn = dk == 192 ? (const void*)attn_fwd_kernel<192> : dk == 128 ? (const void*)attn_fwd_kernel<128> : (const void*)attn_fwd_kernel<64>;
  static bool attr[3] = {false, false, false};
  if (!attr[nc - 1]) {
    rc = dk == 192 ? set_smem(attn_fwd_kernel<192>, SMEM_MAX, "relpos_attn_fwd")
                   : dk == 128 ? set_smem(attn_fwd_kernel<128>, SMEM_MAX, "relpos_attn_fwd") : set_smem(attn_fwd_kernel<64>, SMEM_MAX, "relpos_attn_fwd");
    if (rc) return rc;
    attr[nc - 1] = true;
  }
  return launch(fn, grid, NUM_THREADS, smem, (cudaStream_t)stream, args, "relpos_attn_fwd");
}

extern "C" int a3t_relpos_attn_bwd(const void* qkv4, const void* bd_raw, int64_t ld, const uint8_t* keymask, const void* ctx,
                                   const void* dctx, const float* lse, float* delta_ws, void* dqkv4, void* pd, void* ds, void* dbd,
                                   int B, int H, int S, int D, float scale, float drop_p, const unsigned long long* seed,
                                   uint32_t site, void* stream) {
  using namespace fa;
  int rc = check_common("relpos_attn_bwd", qkv4, bd_raw, keymask, B, H, S, D, ld, drop_p, seed);
  if (rc) return rc;
  A3T_REQUIRE(ctx && dctx && lse && delta_ws && dqkv4 && pd && ds && dbd, "relpos_attn_bwd: null pointer");
  A3T_REQUIRE(((((uintptr_t)ctx) | ((uintptr_t)dctx) | ((uintptr_t)dqkv4) | ((uintptr_t)ds)) & 15) == 0 && (ld % 8) == 0,
              "relpos_attn_bwd: 16-byte alignment / pitch multiple of 8");
  const int dk = D / H;
  A3T_REQUIRE(((((uintptr_t)bd_raw) | ((uintptr_t)pd)) & 15) == 0 && S <= 1900, "relpos_attn_bwd: alignment / S <= 1900");
  CUtensorMap tmQ, tmDO, tmBD, tmDS, tmPD;
  A3T_REQUIRE(qkv_map(&tmQ, qkv4, B, S, 4 * D) && qkv_map(&tmDO, dctx, B, S, D) &&
                  score_map(&tmBD, bd_raw, B, H, S, ld, 88, 16, false),
              "relpos_attn_bwd: tensor map");
  {
    // stores clip at column S (not at the pitch): the pad columns of a row are never written
    const int64_t dims[4] = {S, S, H, B}, str[3] = {ld, (int64_t)S * ld, (int64_t)H * S * ld};
    const int box[4] = {64, 128, 1, 1};
    A3T_REQUIRE(tc::encode_map(&tmDS, ds, dims, str, box) && tc::encode_map(&tmPD, pd, dims, str, box),
                "relpos_attn_bwd: dS / Pd tensor map");
  }
  Params p;
  memset(&p, 0, sizeof(p));
  p.bd_raw = (const __nv_bfloat16*)bd_raw; p.keymask = keymask; p.seed = seed; p.lse = const_cast<float*>(lse);
  p.ctx = (__nv_bfloat16*)const_cast<void*>(ctx); p.dctx = (const __nv_bfloat16*)dctx; p.dq = (__nv_bfloat16*)dqkv4;
  p.pd = (__nv_bfloat16*)pd; p.dbd = (__nv_bfloat16*)dbd; p.delta = delta_ws;
  attn_delta_kernel<<<(unsigned)(((int64_t)B * S + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dctx, (const __nv_bfloat16*)ctx,
                                                                                         delta_ws, B, H, S, D);
  rc = check_launch("relpos_attn_bwd(delta)");
  if (rc) return rc;
  p.ld = ld; p.B = B; p.H = H; p.S = S; p.D = D; p.scale = scale; p.c2 = scale * 1.4426950408889634f; p.drop_p = drop_p;
  p.site = site;
  p.trace = g_trace;
  const int nc = dk / 64;
  const int smem = 1024 + 2 * nc * 16384 + 3 * nc * 8192 + 2 * 16384 + 8 * 16 * 88 * 2 + 256 + 256;
  const int grid = B * H * ((S + BM - 1) / BM);
  void* args[6] = {&tmQ, &tmDO, &tmBD, &tmDS, &tmPD, &p};
  const void* fn = dk == 192 ? (const void*)attn_bwd_kernel<192> : dk == 128 ? (const void*)attn_bwd_kernel<128> : (const void*)attn_bwd_kernel<64>;
  static bool attr[3] = {false, false, false};
  if (!attr[nc - 1]) {
    rc = dk == 192 ? set_smem(attn_bwd_kernel<192>, SMEM_MAX, "relpos_attn_bwd")
                   : dk == 128 ? set_smem(attn_bwd_kernel<128>, SMEM_MAX, "relpos_attn_bwd") : set_smem(attn_bwd_kernel<64>, SMEM_MAX, "relpos_attn_bwd");
    if (rc) return rc;
    attr[nc - 1] = true;
  }
  return launch(fn, grid, NUM_THREADS_BWD, smem, (cudaStream_t)stream, args, "relpos_attn_bwd");
}
