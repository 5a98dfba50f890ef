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

// ---------------------------------------------------------------------------------------------- probe targets
// The component fields the probes regress on (src/scripts/linear_probe_analysis.py:60-110 normalize_component, :180-190
// 4x4 nanmean pooling to the latent grid). NaN marks an invalid pixel everywhere below.

// order-preserving map float -> uint32 (negative values reversed); every non-NaN float has a distinct key
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// out[0..4] = count, sum (x - c), sum (x - c)^2, min x, max x over the non-NaN elements (fp64, fixed order)
__global__ void __launch_bounds__(1024)
nan_moments_kernel(const float* __restrict__ x, long long n, float center, double* __restrict__ out) {
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i];
    if (v == v) {
      const double d = (double)v - (double)center;
      cnt += 1.0; s1 += d; s2 += d * d;
      mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
  }
  __shared__ double sh[3][32];
  __shared__ float shm[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { sh[0][warp] = cnt; sh[1][warp] = s1; sh[2][warp] = s2; shm[0][warp] = mn; shm[1][warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    float lo = INFINITY, hi = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += sh[0][w]; b += sh[1][w]; c += sh[2][w];
      lo = fminf(lo, shm[0][w]); hi = fmaxf(hi, shm[1][w]);
    }
    out[0] = a; out[1] = b; out[2] = c; out[3] = (double)lo; out[4] = (double)hi;
  }
}

// One pass of an exact radix select over v = x (use_abs = 0) or |x - center| (use_abs = 1), NaN skipped: hist[b] +=
// number of elements whose key agrees with `prefix` on `prefix_mask` and whose byte (key >> shift) & 255 is b.
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ x, long long n, float center, int use_abs, uint32_t prefix,
                   uint32_t prefix_mask, int shift, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0u;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i];
    if (xv == xv) {
      const float v = use_abs ? fabsf(__fsub_rn(xv, center)) : xv;
      const uint32_t k = float_key(v);
      if ((k & prefix_mask) == prefix) atomicAdd(&sh[(k >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// mode 0: (x - a) / b   (zscore: a = mean, b = std + 1e-8; minmax: a = min, b = max - min + 1e-8)
// mode 1: asinh(x / b)  (b = scale + 1e-8)
// mode 2: logit(a + (1 - 2a) x) = log(p / (1 - p))
__device__ __forceinline__ float component_f(float x, int mode, float a, float b, float one_minus_2a) {
  if (mode == 0) return __fdiv_rn(__fsub_rn(x, a), b);
  if (mode == 1) return asinhf(__fdiv_rn(x, b));
  const float p = __fadd_rn(a, __fmul_rn(one_minus_2a, x));
  return logf(__fdiv_rn(p, __fsub_rn(1.0f, p)));
}

// One thread per pooled pixel: normalises its pool x pool block (optionally writing the normalised field) and takes
// the mean of the valid values (NaN when the block has none).
__global__ void __launch_bounds__(256)
component_pool_kernel(const float* __restrict__ x, int H, int W, int pitch, int pool, int mode, float a, float b,
                      float one_minus_2a, float* __restrict__ normalized, float* __restrict__ pooled) {
  const int hp = H / pool, wp = W / pool;
  const long long total = (long long)hp * wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / wp), c = (int)(i - (long long)r * wp);
    float s = 0.f;
    int cnt = 0;
    for (int dr = 0; dr < pool; ++dr) {
      const long long off = (long long)(r * pool + dr) * pitch + (long long)c * pool;
      for (int dc = 0; dc < pool; ++dc) {
        const float v = x[off + dc];
        const float f = (v == v) ? component_f(v, mode, a, b, one_minus_2a) : v;
        if (normalized) normalized[(long long)(r * pool + dr) * W + (long long)c * pool + dc] = f;
        if (f == f) { s += f; ++cnt; }
      }
    }
    pooled[i] = cnt ? __fdiv_rn(s, (float)cnt) : __int_as_float(0x7fc00000);
  }
}

// rows / columns beyond the last whole pool block only exist in the normalised field
__global__ void __launch_bounds__(256)
component_edge_kernel(const float* __restrict__ x, int H, int W, int pitch, int pool, int mode, float a, float b,
                      float one_minus_2a, float* __restrict__ normalized) {
  const int h0 = (H / pool) * pool, w0 = (W / pool) * pool;
  const long long total = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / W), c = (int)(i - (long long)r * W);
    if (r < h0 && c < w0) continue;
    const float v = x[(long long)r * pitch + c];
    normalized[i] = (v == v) ? component_f(v, mode, a, b, one_minus_2a) : v;
  }
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

extern "C" int32_t tvae_nan_moments(const float* x, int64_t n, float center, double* out5, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out5 && n > 0, "tvae_nan_moments: bad arguments");
  nan_moments_kernel<<<1, 1024, 0, stream>>>(x, n, center, out5);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_select_hist(const float* x, int64_t n, float center, int32_t use_abs, uint32_t prefix,
                                    uint32_t prefix_mask, int32_t shift, uint64_t* hist256, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && hist256 && n > 0, "tvae_select_hist: bad arguments");
  TVAE_CHECK(shift >= 0 && shift <= 24 && (shift & 7) == 0 && (prefix & ~prefix_mask) == 0u,
             "tvae_select_hist: shift must be 0, 8, 16 or 24 and prefix must lie inside prefix_mask");
  TVAE_CUDA(cudaMemsetAsync(hist256, 0, 256 * sizeof(uint64_t), stream));
  select_hist_kernel<<<pw_grid(n), 256, 0, stream>>>(x, n, center, use_abs, prefix, prefix_mask, shift,
                                                     reinterpret_cast<unsigned long long*>(hist256));
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_component_pool(const float* x, int32_t H, int32_t W, int32_t pitch, int32_t pool, int32_t mode,
                                       float a, float b, float* normalized, float* pooled, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && pooled && H > 0 && W > 0 && pitch >= W && pool >= 1 && H >= pool && W >= pool,
             "tvae_component_pool: bad arguments");
  TVAE_CHECK(mode >= 0 && mode <= 2, "tvae_component_pool: mode must be 0 (affine), 1 (asinh) or 2 (logit)");
  const float om2a = 1.0f - 2.0f * a;
  component_pool_kernel<<<pw_grid((long long)(H / pool) * (W / pool)), 256, 0, stream>>>(x, H, W, pitch, pool, mode, a, b,
                                                                                       om2a, normalized, pooled);
  TVAE_CUDA(cudaGetLastError());
  if (normalized && (H % pool || W % pool)) {
    component_edge_kernel<<<pw_grid((long long)H * W), 256, 0, stream>>>(x, H, W, pitch, pool, mode, a, b, om2a, normalized);
    TVAE_CUDA(cudaGetLastError());
  }
  return 0;
}
