// Probe training on latents: the pieces of a linear / MLP probe that are not a GEMM.
// The Linear layers run as 1x1 implicit-GEMM convolutions through tvae_conv_gemm / tvae_wgrad_gemm (+ bias in the
// epilogue) and the optimiser is tvae_adamw; what is left is the pointwise activation + dropout between layers and
// the MSE loss with its gradient (src/scripts/linear_probe_analysis.py:212-353).
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

__device__ __forceinline__ void philox_round_p(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// four uniform 32-bit words for counter (lo, hi) under `seed` (Philox4x32-10)
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr, uint32_t stream_id, uint32_t (&c)[4]) {
  c[0] = (uint32_t)ctr; c[1] = (uint32_t)(ctr >> 32); c[2] = stream_id; c[3] = 0x5052424Fu;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round_p(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// activation codes: the C ABI's 0 identity / 1 GELU (erf) / 2 ReLU / 3 SiLU, plus 4 = tanh (probe configs)
__device__ __forceinline__ float pact(float y, int act) { return act == 4 ? tanhf(y) : act_f(y, act); }
__device__ __forceinline__ float pact_grad(float y, int act) {
  if (act == 4) { const float t = tanhf(y); return 1.0f - t * t; }
  return act_grad_f(y, act);
}

// keep mask of element (row, c): one Philox call covers four consecutive channels of a row
__device__ __forceinline__ float keep_scale(uint64_t seed, uint64_t offset, long long row, int c4, int j, float p,
                                            float inv_keep, const uint32_t (&r)[4]) {
  (void)seed; (void)offset; (void)row; (void)c4;
  // uniform in [0, 1): dropped when u < p (torch semantics: an element is zeroed with probability p)
  const float u = (float)(r[j] >> 8) * (1.0f / 16777216.0f);
  return u < p ? 0.f : inv_keep;
}

// out = dropout(act(x)): x fp32 [rows][x_pitch] (bias already added by the conv epilogue), out bf16 [rows][out_pitch]
__global__ void __launch_bounds__(256)
act_dropout_fwd_kernel(const float* __restrict__ x, int x_pitch, long long rows, int C, int act, float p, uint64_t seed,
                       uint64_t offset, __nv_bfloat16* __restrict__ out, int out_pitch) {
  const int Q = (out_pitch + 3) >> 2;                      // channel quads per row, pad lanes included (zeroed)
  const long long total = rows * Q;
  const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Q;
    const int c4 = (int)(i - row * Q) << 2;
    uint32_t r[4] = {0, 0, 0, 0};
    if (p > 0.f) philox4(seed, (uint64_t)(offset + row), (uint32_t)(c4 >> 2), r);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c4 + j;
      float a = 0.f;
      if (c < C) {
        a = pact(x[row * x_pitch + c], act);
        if (p > 0.f) a *= keep_scale(seed, offset, row, c4, j, p, inv_keep, r);
      }
      v[j] = a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c4 + j < out_pitch) out[row * out_pitch + c4 + j] = __float2bfloat16(v[j]);
  }
}

// dx = da * mask * act'(x), bf16
__global__ void __launch_bounds__(256)
act_dropout_bwd_kernel(const float* __restrict__ x, int x_pitch, const __nv_bfloat16* __restrict__ da, int da_pitch,
                       long long rows, int C, int act, float p, uint64_t seed, uint64_t offset,
                       __nv_bfloat16* __restrict__ dx, int dx_pitch) {
  const int Q = (dx_pitch + 3) >> 2;
  const long long total = rows * Q;
  const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Q;
    const int c4 = (int)(i - row * Q) << 2;
    uint32_t r[4] = {0, 0, 0, 0};
    if (p > 0.f) philox4(seed, (uint64_t)(offset + row), (uint32_t)(c4 >> 2), r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c4 + j;
      if (c >= dx_pitch) continue;
      float g = 0.f;
      if (c < C) {
        g = __bfloat162float(da[row * da_pitch + c]) * pact_grad(x[row * x_pitch + c], act);
        if (p > 0.f) g *= keep_scale(seed, offset, row, c4, j, p, inv_keep, r);
      }
      dx[row * dx_pitch + c] = __float2bfloat16(g);
    }
  }
}

// sums[0] = sum (pred - y)^2, sums[1] = sum y, sums[2] = sum y^2 over the first n_valid rows (fp64, one block, fixed
// order); dpred (optional, bf16 [rows_padded][dp_pitch]) = 2 (pred - y) / n_valid on valid rows, 0 elsewhere
__global__ void __launch_bounds__(1024)
probe_mse_kernel(const float* __restrict__ pred, int pred_pitch, const float* __restrict__ y, long long y_pitch,
                 long long n_valid,
                 long long rows_padded, double* __restrict__ sums, __nv_bfloat16* __restrict__ dpred, int dp_pitch) {
  __shared__ double sh[3][1024];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const float scale = 2.0f / (float)n_valid;
  for (long long r = threadIdx.x; r < rows_padded; r += blockDim.x) {
    float g = 0.f;
    if (r < n_valid) {
      const float yr = y[r * y_pitch];
      const float d = pred[r * pred_pitch] - yr;
      const double yy = (double)yr;
      s0 += (double)d * (double)d;
      s1 += yy;
      s2 += yy * yy;
      g = d * scale;
    }
    if (dpred) {
      for (int c = 0; c < dp_pitch; ++c) dpred[r * dp_pitch + c] = __float2bfloat16(c == 0 ? g : 0.f);
    }
  }
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) sums[threadIdx.x] = sh[threadIdx.x][0];
}

inline int pw_grid(long long work) {
  long long g = (work + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (int)g;
}

}  // namespace
}  // namespace tvae

using namespace tvae;

extern "C" int32_t tvae_act_dropout_fwd(const float* x, int32_t x_pitch, int64_t rows, int32_t C, int32_t act,
                                        float p_drop, uint64_t seed, uint64_t offset, void* out_bf16,
                                        int32_t out_pitch, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out_bf16 && rows > 0 && C > 0 && x_pitch >= C && out_pitch >= C, "tvae_act_dropout_fwd: bad arguments");
  TVAE_CHECK(p_drop >= 0.f && p_drop < 1.f && act >= 0 && act <= 4, "tvae_act_dropout_fwd: bad activation / dropout");
  act_dropout_fwd_kernel<<<pw_grid(rows * ((out_pitch + 3) / 4)), 256, 0, stream>>>(
      x, x_pitch, rows, C, act, p_drop, seed, offset, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_act_dropout_bwd(const float* x, int32_t x_pitch, const void* da_bf16, int32_t da_pitch,
                                        int64_t rows, int32_t C, int32_t act, float p_drop, uint64_t seed,
                                        uint64_t offset, void* dx_bf16, int32_t dx_pitch, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && da_bf16 && dx_bf16 && rows > 0 && C > 0 && x_pitch >= C && da_pitch >= C && dx_pitch >= C,
             "tvae_act_dropout_bwd: bad arguments");
  TVAE_CHECK(p_drop >= 0.f && p_drop < 1.f && act >= 0 && act <= 4, "tvae_act_dropout_bwd: bad activation / dropout");
  act_dropout_bwd_kernel<<<pw_grid(rows * ((dx_pitch + 3) / 4)), 256, 0, stream>>>(
      x, x_pitch, reinterpret_cast<const __nv_bfloat16*>(da_bf16), da_pitch, rows, C, act, p_drop, seed, offset,
      reinterpret_cast<__nv_bfloat16*>(dx_bf16), dx_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_probe_mse(const float* pred, int32_t pred_pitch, const float* target, int64_t target_pitch,
                                  int64_t n_valid, int64_t rows_padded, double* sums, void* dpred_bf16, int32_t dp_pitch,
                                  cudaStream_t stream) {
  TVAE_ENTER(pred);
  TVAE_CHECK(pred && target && sums && n_valid > 0 && rows_padded >= n_valid && pred_pitch >= 1 && target_pitch >= 1,
             "tvae_probe_mse: bad arguments");
  TVAE_CHECK(!dpred_bf16 || dp_pitch >= 1, "tvae_probe_mse: bad dp_pitch");
  probe_mse_kernel<<<1, 1024, 0, stream>>>(pred, pred_pitch, target, target_pitch, n_valid, rows_padded, sums,
                                           reinterpret_cast<__nv_bfloat16*>(dpred_bf16), dp_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
