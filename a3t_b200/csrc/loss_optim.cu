// Masked L1 loss (espnet2/tts/sedit/sedit_model.py:320-340) and the trainer glue
// (espnet2/train/trainer.py:631-675 clip + step, schedulers/noam_lr.py:58-65).
#include "common.cuh"

namespace a3t {

constexpr int L1_BLOCKS = 296;

// one warp per row; partial[blk] = (sum mask*l1, sum mask) in double
__global__ void __launch_bounds__(256) masked_l1_partial_kernel(const float* __restrict__ before,
                                                                const float* __restrict__ after,
                                                                const float* __restrict__ y,
                                                                const uint8_t* __restrict__ mask,
                                                                double* __restrict__ partial, int64_t rows, int C) {
  A3T_PDL_TRIGGER();
  __shared__ double sh[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double num = 0.0, den = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    if (!mask[r]) continue;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      float t = y[r * C + c];
      s += fabsf(before[r * C + c] - t);
      if (after) s += fabsf(after[r * C + c] - t);
    }
    s = warp_sum(s);
    num += (double)s;
    den += 1.0;
  }
  if (lane == 0) { sh[0][warp] = num; sh[1][warp] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += sh[0][w]; b += sh[1][w]; }
    partial[blockIdx.x * 2] = a;
    partial[blockIdx.x * 2 + 1] = b;
  }
}
__global__ void masked_l1_final_kernel(const double* __restrict__ partial, float* __restrict__ out, int nblk) {
  A3T_PDL_TRIGGER();
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 32) { a += partial[i * 2]; b += partial[i * 2 + 1]; }
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if (threadIdx.x == 0) {
    float den = (float)b + 1e-10f;
    out[0] = (float)a / den;
    out[1] = den;
  }
}
__global__ void __launch_bounds__(256) masked_l1_bwd_kernel(const float* __restrict__ gloss,
                                                            const float* __restrict__ before,
                                                            const float* __restrict__ after,
                                                            const float* __restrict__ y,
                                                            const uint8_t* __restrict__ mask,
                                                            const float* __restrict__ den, float* __restrict__ dbefore,
                                                            float* __restrict__ dafter, int64_t rows, int C) {
  A3T_PDL_TRIGGER();
  const float g = gloss[0] / den[0];
  int64_t n = rows * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = i / C;
    float m = mask[r] ? g : 0.f;
    float t = y[i];
    float db = before[i] - t;
    dbefore[i] = m * (db > 0.f ? 1.f : (db < 0.f ? -1.f : 0.f));
    if (dafter) {
      float da = after[i] - t;
      dafter[i] = m * (da > 0.f ? 1.f : (da < 0.f ? -1.f : 0.f));
    }
  }
}

// ---- optimizer ----------------------------------------------------------------------------
constexpr int SQ_BLOCKS = 1024;
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ g, int64_t n,
                                                             double* __restrict__ partial) {
  A3T_PDL_TRIGGER();
  __shared__ double sh[8];
  double acc = 0.0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float facc = 0.f;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = g[i];
    facc = fmaf(v, v, facc);
    if (++cnt == 64) { acc += (double)facc; facc = 0.f; cnt = 0; }
  }
  acc += (double)facc;
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sqnorm_final_kernel(const double* __restrict__ partial, double* __restrict__ sq, int nblk) {
  A3T_PDL_TRIGGER();
  double a = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 32) a += partial[i];
  a = warp_sum_d(a);
  if (threadIdx.x == 0) sq[0] = a;
}

// clip_grad_norm_ + Adam + Noam in one pass.  Every thread derives the same scalars from
// (sq, step); thread 0 of block 0 advances the step counter at the END of the kernel via a
// second tiny kernel (stream order) so all blocks see the same value.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   const double* __restrict__ sq, const int64_t* __restrict__ step_p,
                                                   float base_lr, float model_size, float warmup, float beta1,
                                                   float beta2, float eps, float max_norm, float grad_scale,
                                                   const float* __restrict__ denom) {
  A3T_PDL_TRIGGER();
  if (denom) grad_scale /= denom[0];
  // total norm of the scaled gradient
  const double norm = sqrt(sq[0]) * (double)grad_scale;
  if (!isfinite(norm)) return;  // trainer.py:640-656: skip the update
  const double step = (double)(step_p[0] + 1);
  float coef = 1.f;
  if (max_norm > 0.f) {
    double c = (double)max_norm / (norm + 1e-6);
    coef = c < 1.0 ? (float)c : 1.f;
  }
  const float gs = grad_scale * coef;
  double lr = (double)base_lr;
  if (warmup > 0.f)
    lr = (double)base_lr * pow((double)model_size, -0.5) * fmin(pow(step, -0.5), step * pow((double)warmup, -1.5));
  const double bc1 = 1.0 - pow((double)beta1, step);
  const double bc2 = 1.0 - pow((double)beta2, step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t i = tid; i < n4; i += stride) {  // 16-byte accesses: 28 B of traffic per parameter
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    float ge[4] = {g4.x, g4.y, g4.z, g4.w}, me[4] = {m4.x, m4.y, m4.z, m4.w}, ve[4] = {v4.x, v4.y, v4.z, v4.w};
    float pe[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float gi = ge[e] * gs;
      me[e] = me[e] * beta1 + gi * (1.f - beta1);
      ve[e] = ve[e] * beta2 + gi * gi * (1.f - beta2);
      float dn = sqrtf(ve[e]) * inv_sqrt_bc2 + eps;
      pe[e] = pe[e] - step_size * (me[e] / dn);
    }
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
    float gi = g[i] * gs;
    float mi = m[i] * beta1 + gi * (1.f - beta1);
    float vi = v[i] * beta2 + gi * gi * (1.f - beta2);
    m[i] = mi;
    v[i] = vi;
    float dn = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - step_size * (mi / dn);
  }
}
__global__ void step_advance_kernel(const double* __restrict__ sq, int64_t* __restrict__ step_p, float grad_scale,
                                    const float* __restrict__ denom) {
  A3T_PDL_TRIGGER();
  if (denom) grad_scale /= denom[0];
  const double norm = sqrt(sq[0]) * (double)grad_scale;
  if (isfinite(norm)) step_p[0] += 1;
}
__global__ void seed_advance_kernel(unsigned long long* seed) {
  A3T_PDL_TRIGGER();
  *seed = *seed * 6364136223846793005ull + 1442695040888963407ull;
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_masked_l1_fwd(const float* before, const float* after, const float* y, const uint8_t* mask,
                                 float* out, double* partial, int64_t rows, int C, void* stream) {
  A3T_REQUIRE(before && y && mask && out && partial, "masked_l1_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = a3t_colsum_blocks(rows);
  masked_l1_partial_kernel<<<nblk, 256, 0, st>>>(before, after, y, mask, partial, rows, C);
  int rc = check_launch("masked_l1_partial");
  if (rc) return rc;
  masked_l1_final_kernel<<<1, 32, 0, st>>>(partial, out, nblk);
  return check_launch("masked_l1_final");
}

extern "C" int a3t_masked_l1_bwd(const float* gloss, const float* before, const float* after, const float* y,
                                 const uint8_t* mask, const float* den, float* dbefore, float* dafter, int64_t rows,
                                 int C, void* stream) {
  A3T_REQUIRE(gloss && before && y && mask && den && dbefore, "masked_l1_bwd: null pointer");
  A3T_REQUIRE((after == nullptr) == (dafter == nullptr), "masked_l1_bwd: after/dafter mismatch");
  int64_t n = rows * C;
  if (n == 0) return A3T_OK;
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  masked_l1_bwd_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(gloss, before, after, y, mask, den, dbefore, dafter,
                                                                 rows, C);
  return check_launch("masked_l1_bwd");
}

extern "C" int a3t_grad_sqnorm(const float* g, int64_t n, double* sq, double* partial, void* stream) {
  A3T_REQUIRE(g && sq && partial, "grad_sqnorm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  sqnorm_partial_kernel<<<SQ_BLOCKS, 256, 0, st>>>(g, n, partial);
  int rc = check_launch("grad_sqnorm");
  if (rc) return rc;
  sqnorm_final_kernel<<<1, 32, 0, st>>>(partial, sq, SQ_BLOCKS);
  return check_launch("grad_sqnorm_final");
}

extern "C" int a3t_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const double* sq, int64_t* step,
                             float base_lr, float model_size, float warmup, float beta1, float beta2, float eps,
                             float max_norm, float grad_scale, const float* denom, void* stream) {
  A3T_REQUIRE(p && g && m && v && sq && step, "adam_step: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  adam_kernel<<<(int)b, 256, 0, st>>>(p, g, m, v, n, sq, step, base_lr, model_size, warmup, beta1, beta2, eps,
                                      max_norm, grad_scale, denom);
  int rc = check_launch("adam_step");
  if (rc) return rc;
  step_advance_kernel<<<1, 1, 0, st>>>(sq, step, grad_scale, denom);
  return check_launch("adam_step_advance");
}

extern "C" int a3t_seed_advance(unsigned long long* seed, void* stream) {
  A3T_REQUIRE(seed, "seed_advance: null pointer");
  seed_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(seed);
  return check_launch("seed_advance");
}
