// Fused latent / loss / optimiser kernels of the TEMPO-VAE hot path (all HBM- or latency-bound):
//   reparameterisation + KL (Philox or caller-supplied eps), reconstruction NLL with its gradient,
//   L2-product masked MSE, global grad-norm and the fused clip + AdamW step.
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// standard normal for element `idx` of sample `sample` under `seed` (Box-Muller on two 32-bit uniforms)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t sample, uint32_t idx) {
  uint32_t c[4] = {idx >> 1, 0u, (uint32_t)sample, (uint32_t)(sample >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = (float)c[1] * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, co;
  __sincosf(6.283185307179586f * u2, &s, &co);
  return (idx & 1) ? r * s : r * co;
}

// ---------------------------------------------------------------------------------------------- reparam + KL
// CL = 1: one block per sample (the training shapes: 8 K latent elements per sample, hundreds of samples).
// CL = 8: a cluster of eight 1024-thread blocks per sample for the whole-granule passes (ONE sample of 524 K latent
// elements: a single 256-thread block took 1.2 ms); the eight KL partials meet in rank 0 through distributed shared
// memory and are added in rank order, so the result does not depend on scheduling.
template <int CL>
__global__ void reparam_fwd_kernel(const float* __restrict__ moments, const float* __restrict__ eps, uint64_t seed,
                                   uint64_t sample_offset, int HW, int Z, __nv_bfloat16* __restrict__ z_bf16,
                                   __nv_bfloat16* __restrict__ z_lo, int z_pitch, float* __restrict__ z_nchw,
                                   float* __restrict__ eps_out, float* __restrict__ kl) {
  __shared__ double red[32];
  __shared__ double cl_part;
  const int b = blockIdx.x / CL;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int total = HW * Z;
  double acc = 0.0;
  for (int e = rank * blockDim.x + threadIdx.x; e < total; e += CL * blockDim.x) {
    const int p = e / Z, c = e - p * Z;
    const long long row = (long long)b * HW + p;
    const float mean = moments[row * 2 * Z + c];
    float lv = moments[row * 2 * Z + Z + c];
    lv = fminf(fmaxf(lv, -30.0f), 20.0f);
    const long long nchw = ((long long)b * Z + c) * HW + p;
    const float ev = eps ? eps[nchw] : philox_normal(seed, sample_offset + (uint64_t)b, (uint32_t)(c * HW + p));
    const float std = expf(0.5f * lv);
    const float z = mean + std * ev;
    if (z_bf16) z_bf16[row * z_pitch + c] = __float2bfloat16(z);
    if (z_lo) z_lo[row * z_pitch + c] = __float2bfloat16(z - __bfloat162float(__float2bfloat16(z)));
    if (z_nchw) z_nchw[nchw] = z;
    if (eps_out) eps_out[nchw] = ev;
    acc += (double)(0.5f * (mean * mean + expf(lv) - 1.0f - lv));
  }
  const double t = block_sum(acc, red);
  if (CL == 1) {
    if (threadIdx.x == 0 && kl) kl[b] = (float)t;
    return;
  }
  if (threadIdx.x == 0) cl_part = t;
  cluster_sync_all();
  if (rank == 0 && threadIdx.x == 0 && kl) {
    double sum = 0.0;
    for (int r = 0; r < CL; ++r) {
      double v;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(mapa_u32(smem_u32(&cl_part), (uint32_t)r)) : "memory");
      sum += v;
    }
    kl[b] = (float)sum;
  }
  cluster_sync_all();      // nobody leaves while rank 0 may still be reading its shared memory
}

__global__ void reparam_bwd_kernel(const float* __restrict__ moments, const float* __restrict__ dz1,
                                   const float* __restrict__ eps1, const float* __restrict__ dz2,
                                   const float* __restrict__ eps2, float kl_scale, int B, int HW, int Z,
                                   __nv_bfloat16* __restrict__ dm) {
  const long long total = (long long)B * HW * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Z;
    const int c = (int)(i - row * Z);
    const int b = (int)(row / HW), p = (int)(row - (long long)b * HW);
    const float mean = moments[row * 2 * Z + c];
    const float lv_raw = moments[row * 2 * Z + Z + c];
    const float lv = fminf(fmaxf(lv_raw, -30.0f), 20.0f);
    const float std = expf(0.5f * lv), var = expf(lv);
    const long long nchw = ((long long)b * Z + c) * HW + p;
    float g_mean = kl_scale * mean;
    float g_lv = kl_scale * 0.5f * (var - 1.0f);
    if (dz1) {
      const float d = dz1[row * Z + c];
      g_mean += d;
      g_lv += d * eps1[nchw] * 0.5f * std;
    }
    if (dz2) {
      const float d = dz2[row * Z + c];
      g_mean += d;
      g_lv += d * eps2[nchw] * 0.5f * std;
    }
    if (lv_raw < -30.0f || lv_raw > 20.0f) g_lv = 0.f;  // clamp has zero gradient outside its range
    dm[row * 2 * Z + c] = __float2bfloat16(g_mean);
    dm[row * 2 * Z + Z + c] = __float2bfloat16(g_lv);
  }
}

// ---------------------------------------------------------------------------------------------- NLL
constexpr int NLL_BLOCKS = 148 * 8;

template <int VEC>
__global__ void nll_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const float* __restrict__ xh,
                               int xh_pitch, long long P, int C, int loss_type, const float* __restrict__ logvar,
                               int batch, __nv_bfloat16* __restrict__ dxh, int dx_pitch, double* __restrict__ ws) {
  __shared__ double red[32];
  const int U = C / VEC;
  const long long total = P * U;
  const float gs = dxh ? expf(-logvar[0]) / (float)batch : 0.f;
  double a_rec = 0.0, a_sq = 0.0;
  float f_rec = 0.f, f_sq = 0.f;
  int cnt = 0;
  // (pixel, channel-vector) advance incrementally by the grid stride: no 64-bit division per element
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long sp = stride / U;
  const int sc = (int)(stride - sp * U);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long p = i / U;
  int cu = (int)(i - p * U);
  for (; i < total; i += stride, p += sp, cu += sc) {
    if (cu >= U) { cu -= U; ++p; }
    const int c = cu * VEC;
    float xv[VEC], hv[VEC], g[VEC];
    if (VEC == 4) {
      const uint2 d = *reinterpret_cast<const uint2*>(x + p * x_pitch + c);
      xv[0] = bf16_bits_to_f(d.x & 0xffffu); xv[1] = bf16_bits_to_f(d.x >> 16);
      xv[2] = bf16_bits_to_f(d.y & 0xffffu); xv[3] = bf16_bits_to_f(d.y >> 16);
      const float4 h = *reinterpret_cast<const float4*>(xh + p * xh_pitch + c);
      hv[0] = h.x; hv[1] = h.y; hv[2] = h.z; hv[3] = h.w;
    } else {
      xv[0] = __bfloat162float(x[p * x_pitch + c]);
      hv[0] = xh[p * xh_pitch + c];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float d = hv[j] - xv[j];
      const float sq = d * d;
      f_sq += sq;
      if (loss_type == 0) {
        f_rec += fabsf(d);
        g[j] = (d > 0.f) ? gs : ((d < 0.f) ? -gs : 0.f);
      } else {
        f_rec += sq;
        g[j] = 2.0f * d * gs;
      }
    }
    if (dxh) {
      if (VEC == 4) {
        uint2 o;
        o.x = pack_bf16(g[0], g[1]);
        o.y = pack_bf16(g[2], g[3]);
        *reinterpret_cast<uint2*>(dxh + p * dx_pitch + c) = o;
      } else {
        dxh[p * dx_pitch + c] = __float2bfloat16(g[0]);
      }
    }
    if (++cnt == 32) { a_rec += f_rec; a_sq += f_sq; f_rec = f_sq = 0.f; cnt = 0; }
  }
  a_rec += f_rec; a_sq += f_sq;
  const double t1 = block_sum(a_rec, red);
  const double t2 = block_sum(a_sq, red);
  if (threadIdx.x == 0) {
    ws[2 * blockIdx.x] = t1;
    ws[2 * blockIdx.x + 1] = t2;
  }
}
// Row-structured variant (C % 4 == 0): thread u of a block owns channel vector u (4 channels) and walks the block's
// pixel rows, so the per-channel sums of the gradient (= the bias gradient of the conv that produced xhat) fall out of
// the same pass: csp[block][C]. blockDim.x >= dx_pitch/4 (a multiple of 32); threads past C/4 zero the pad lanes.
__global__ void nll_fwd_rows_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const float* __restrict__ xh,
                                    int xh_pitch, long long P, int C, int loss_type, const float* __restrict__ logvar,
                                    int batch, __nv_bfloat16* __restrict__ dxh, int dx_pitch, double* __restrict__ ws,
                                    float* __restrict__ csp) {
  __shared__ double red[32];
  const int U = C >> 2, u = threadIdx.x, c = u << 2;
  const bool live = u < U;
  const float gs = dxh ? expf(-logvar[0]) / (float)batch : 0.f;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(r0 + per, P);
  double a_rec = 0.0, a_sq = 0.0;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int RU = 4;      // rows in flight per thread
  for (long long r = r0; r < r1; r += RU) {
    float f_rec = 0.f, f_sq = 0.f;
    const int nr = (int)min((long long)RU, r1 - r);
    if (live) {
      uint2 d[RU];
      float4 h[RU];
#pragma unroll
      for (int k = 0; k < RU; ++k)
        if (k < nr) {
          d[k] = *reinterpret_cast<const uint2*>(x + (r + k) * x_pitch + c);
          h[k] = *reinterpret_cast<const float4*>(xh + (r + k) * xh_pitch + c);
        }
#pragma unroll
      for (int k = 0; k < RU; ++k)
        if (k < nr) {
          const float xv[4] = {bf16_bits_to_f(d[k].x & 0xffffu), bf16_bits_to_f(d[k].x >> 16),
                               bf16_bits_to_f(d[k].y & 0xffffu), bf16_bits_to_f(d[k].y >> 16)};
          const float hv[4] = {h[k].x, h[k].y, h[k].z, h[k].w};
          float g[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float df = hv[j] - xv[j];
            const float sq = df * df;
            f_sq += sq;
            if (loss_type == 0) {
              f_rec += fabsf(df);
              g[j] = (df > 0.f) ? gs : ((df < 0.f) ? -gs : 0.f);
            } else {
              f_rec += sq;
              g[j] = 2.0f * df * gs;
            }
            cs[j] += g[j];
          }
          if (dxh) {
            uint2 o;
            o.x = pack_bf16(g[0], g[1]);
            o.y = pack_bf16(g[2], g[3]);
            *reinterpret_cast<uint2*>(dxh + (r + k) * dx_pitch + c) = o;
          }
        }
    } else if (dxh && c + 4 <= dx_pitch) {
      for (int k = 0; k < nr; ++k) *reinterpret_cast<uint2*>(dxh + (r + k) * dx_pitch + c) = make_uint2(0u, 0u);
    }
    a_rec += f_rec; a_sq += f_sq;
  }
  if (csp && live)
    *reinterpret_cast<float4*>(csp + (long long)blockIdx.x * C + c) = make_float4(cs[0], cs[1], cs[2], cs[3]);
  const double t1 = block_sum(a_rec, red);
  const double t2 = block_sum(a_sq, red);
  if (threadIdx.x == 0) {
    ws[2 * blockIdx.x] = t1;
    ws[2 * blockIdx.x + 1] = t2;
  }
}
// out[c] = sum over blocks of csp[block][c], fixed order: 32 channels x 8 block lanes per CTA
__global__ void __launch_bounds__(256)
nll_colsum_final_kernel(const float* __restrict__ csp, int nblocks, int C, float* __restrict__ out) {
  __shared__ float sa[8][33];
  const int cx = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f;
  if (c < C)
    for (int b = bl; b < nblocks; b += 8) a += csp[(long long)b * C + c];
  sa[bl][cx] = a;
  __syncthreads();
  if (bl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += sa[l][cx];
    out[c] = t;
  }
}
__global__ void nll_final_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ sums) {
  __shared__ double red[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += ws[2 * i]; b += ws[2 * i + 1]; }
  const double t1 = block_sum(a, red);
  const double t2 = block_sum(b, red);
  if (threadIdx.x == 0) { sums[0] = t1; sums[1] = t2; sums[2] = 0.0; }
}

// Per-sample reconstruction error sums (evaluation metrics of src/scripts/evaluate_reconstruction.py:23-42):
// part[n][chunk] = (sum |x - xhat|, sum (x - xhat)^2) over the chunk's pixels; recon_metrics_final adds the chunks.
constexpr int RM_CHUNKS = 16;
__global__ void __launch_bounds__(256)
recon_metrics_kernel(const __nv_bfloat16* __restrict__ x, int x_pitch, const float* __restrict__ xh, int xh_pitch,
                     int HW, int C, double* __restrict__ part) {
  __shared__ double red[32];
  const int n = blockIdx.y;
  const int per = (HW + RM_CHUNKS - 1) / RM_CHUNKS;
  const int p0 = blockIdx.x * per, p1 = min(p0 + per, HW);
  double a1 = 0.0, a2 = 0.0;
  for (int p = p0; p < p1; ++p) {
    const long long row = (long long)n * HW + p;
    float f1 = 0.f, f2 = 0.f;
    for (int c = threadIdx.x; c < C; c += 256) {
      const float d = xh[row * xh_pitch + c] - __bfloat162float(x[row * x_pitch + c]);
      f1 += fabsf(d);
      f2 = fmaf(d, d, f2);
    }
    a1 += f1; a2 += f2;
  }
  const double t1 = block_sum(a1, red);
  const double t2 = block_sum(a2, red);
  if (threadIdx.x == 0) {
    part[2 * (n * RM_CHUNKS + blockIdx.x)] = t1;
    part[2 * (n * RM_CHUNKS + blockIdx.x) + 1] = t2;
  }
}
__global__ void recon_metrics_final_kernel(const double* __restrict__ part, int N, double count, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < RM_CHUNKS; ++k) { a += part[2 * (n * RM_CHUNKS + k)]; b += part[2 * (n * RM_CHUNKS + k) + 1]; }
  out[2 * n] = (float)(a / count);       // MAE
  out[2 * n + 1] = (float)(b / count);   // MSE
}

// loss / metric scalars of AutoencoderKL.get_loss (src/model.py:660-668), one thread:
//   out[0] = loss, out[1] = nll_loss, out[2] = kl_loss (already * kl_weight), out[3] = pixel_mse,
//   out[4] = d(loss)/d(logvar)
__global__ void vae_loss_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ kl, int B,
                                         const float* __restrict__ logvar, double n_elem, float kl_weight,
                                         float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double lv = (double)logvar[0];
  const double inv_var = exp(-lv);
  double kls = 0.0;
  for (int b = 0; b < B; ++b) kls += (double)kl[b];
  const double nll = (sums[0] * inv_var + lv * n_elem) / (double)B;
  const double klw = (double)kl_weight * kls / (double)B;
  out[0] = (float)(nll + klw);
  out[1] = (float)nll;
  out[2] = (float)klw;
  out[3] = (float)(sums[1] / n_elem);
  out[4] = (float)((n_elem - sums[0] * inv_var) / (double)B);
}

// ---------------------------------------------------------------------------------------------- L2 head loss
struct L2Targets { const float* t[8]; };

__device__ __forceinline__ float pooled4x4(const float* __restrict__ tgt, int b, int i, int j, int H, int W) {
  const float* base = tgt + ((long long)b * H + 4 * i) * W + 4 * j;
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s += base[r * W + q];
  }
  return s * 0.0625f;  // NaN propagates exactly like nn.AvgPool2d
}

// one block per product
__global__ void l2head_fwd_kernel(const float* __restrict__ pred, int pitch, L2Targets tg, int B, int h, int w,
                                  double* __restrict__ out) {
  __shared__ double red[32];
  const int pr = blockIdx.x;
  const float* tgt = tg.t[pr];
  double se = 0.0, cnt = 0.0;
  if (tgt) {
    const int total = B * h * w;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int b = e / (h * w), r = e - b * h * w;
      const int i = r / w, j = r - i * w;
      const float t = pooled4x4(tgt, b, i, j, 4 * h, 4 * w);
      if (!isnan(t)) {
        const float d = pred[(long long)e * pitch + pr] - t;
        se += (double)(d * d);
        cnt += 1.0;
      }
    }
  }
  const double t1 = block_sum(se, red);
  const double t2 = block_sum(cnt, red);
  if (threadIdx.x == 0) { out[2 * pr] = t1; out[2 * pr + 1] = t2; }
}

__global__ void l2head_bwd_kernel(const float* __restrict__ pred, int pitch, L2Targets tg, int nprod, int B, int h,
                                  int w, const double* __restrict__ sums, const float* __restrict__ weights,
                                  float grad_scale, __nv_bfloat16* __restrict__ dpred, int dp_pitch) {
  const long long total = (long long)B * h * w * dp_pitch;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / dp_pitch;
    const int pr = (int)(idx - e * dp_pitch);
    float g = 0.f;
    if (pr < nprod && tg.t[pr]) {
      const double cnt = sums[2 * pr + 1];
      if (cnt > 0.0) {
        const int b = (int)(e / (h * w)), r = (int)(e - (long long)b * h * w);
        const int i = r / w, j = r - i * w;
        const float t = pooled4x4(tg.t[pr], b, i, j, 4 * h, 4 * w);
        if (!isnan(t)) g = weights[pr] * grad_scale * 2.0f * (pred[e * pitch + pr] - t) / (float)cnt;
      }
    }
    dpred[idx] = __float2bfloat16(g);
  }
}

// ---------------------------------------------------------------------------------------------- optimiser
constexpr int SUMSQ_BLOCKS = 148 * 4;

__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ ws) {
  __shared__ double red[32];
  double acc = 0.0;
  float f = 0.f;
  int cnt = 0;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    f += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (++cnt == 16) { acc += f; f = 0.f; cnt = 0; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    f += v * v;
  }
  acc += f;
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = t;
}
__global__ void sumsq_final_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ out) {
  __shared__ double red[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += ws[i];
  const double t = block_sum(a, red);
  if (threadIdx.x == 0) out[0] = t;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, const double* __restrict__ sumsq, float max_norm,
                             float grad_scale) {
  float coef = grad_scale;
  if (sumsq) {
    const float total_norm = (float)sqrt(sumsq[0]) * grad_scale;
    float c = max_norm / (total_norm + 1e-6f);
    coef *= fminf(c, 1.0f);
  }
  const float decay = 1.0f - lr * wd;
  const float step_size = lr / bc1;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w};
    const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = ga[j] * coef;
      pa[j] *= decay;
      ma[j] = beta1 * ma[j] + (1.0f - beta1) * gg;
      va[j] = beta2 * va[j] + (1.0f - beta2) * gg * gg;
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    const float gg = g[i] * coef;
    float pp = p[i] * decay;
    const float mm = beta1 * m[i] + (1.0f - beta1) * gg;
    const float vv = beta2 * v[i] + (1.0f - beta2) * gg * gg;
    pp -= step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace
}  // namespace tvae

using namespace tvae;

extern "C" int32_t tvae_reparam_fwd(const float* moments, const float* eps, uint64_t seed, uint64_t sample_offset,
                                    int32_t B, int32_t HW, int32_t Z, void* z_bf16, int32_t z_pitch, float* z_nchw,
                                    float* eps_out, float* kl, void* z_lo, cudaStream_t stream) {
  TVAE_ENTER(moments);
  TVAE_CHECK(moments, "tvae_reparam_fwd: null moments");
  TVAE_CHECK(B > 0 && HW > 0 && Z > 0, "tvae_reparam_fwd: bad shape");
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(z_bf16);
  __nv_bfloat16* zl = reinterpret_cast<__nv_bfloat16*>(z_lo);
  if ((long long)HW * Z >= 65536 && B <= 64) {     // few, large samples (whole-granule inference): 8 blocks per sample
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)B * 8u);
    cfg.blockDim = dim3(1024);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TVAE_CUDA(cudaLaunchKernelEx(&cfg, reparam_fwd_kernel<8>, moments, eps, seed, sample_offset, (int)HW, (int)Z, zb, zl,
                                 (int)z_pitch, z_nchw, eps_out, kl));
    return 0;
  }
  reparam_fwd_kernel<1><<<B, 256, 0, stream>>>(moments, eps, seed, sample_offset, HW, Z, zb, zl, z_pitch, z_nchw, eps_out, kl);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_reparam_bwd(const float* moments, const float* dz1, const float* eps1, const float* dz2,
                                    const float* eps2, float kl_scale, int32_t B, int32_t HW, int32_t Z, void* dm,
                                    cudaStream_t stream) {
  TVAE_ENTER(moments);
  TVAE_CHECK(moments && dm, "tvae_reparam_bwd: null pointer");
  TVAE_CHECK((!dz1 || eps1) && (!dz2 || eps2), "tvae_reparam_bwd: dz without eps");
  const long long total = (long long)B * HW * Z;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  reparam_bwd_kernel<<<grid, 256, 0, stream>>>(moments, dz1, eps1, dz2, eps2, kl_scale, B, HW, Z,
                                               reinterpret_cast<__nv_bfloat16*>(dm));
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_nll_workspace_bytes(int32_t C) {
  return (int64_t)NLL_BLOCKS * 2 * sizeof(double) + (int64_t)NLL_BLOCKS * (C > 0 ? (C + 3) / 4 * 4 : 0) * sizeof(float);
}

extern "C" int32_t tvae_nll_fwd(const void* x, int32_t x_pitch, const float* xhat, int32_t xh_pitch, int64_t P,
                                int32_t C, int32_t loss_type, const float* logvar, int32_t batch, void* dxhat,
                                int32_t dx_pitch, float* dx_colsum, double* sums, double* ws, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && xhat && sums && ws, "tvae_nll_fwd: null pointer");
  TVAE_CHECK(!dxhat || logvar, "tvae_nll_fwd: dxhat needs logvar");
  TVAE_CHECK(!dx_colsum || dxhat, "tvae_nll_fwd: dx_colsum needs dxhat");
  TVAE_CHECK(loss_type == 0 || loss_type == 1, "tvae_nll_fwd: loss_type must be 0 (l1) or 1 (l2)");
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dxhat);
  const bool vec = (C % 4 == 0) && (x_pitch % 4 == 0) && (xh_pitch % 4 == 0) && (!dxhat || dx_pitch % 4 == 0);
  float* csp = reinterpret_cast<float*>(ws + 2 * NLL_BLOCKS);
  const int threads = ((dxhat ? dx_pitch : C) / 4 + 31) / 32 * 32;
  if (vec && threads <= 1024 && P >= NLL_BLOCKS && (reinterpret_cast<uintptr_t>(xhat) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(dxhat) & 7) == 0) {
    nll_fwd_rows_kernel<<<NLL_BLOCKS, threads, 0, stream>>>(xp, x_pitch, xhat, xh_pitch, P, C, loss_type, logvar, batch,
                                                           dp, dx_pitch, ws, dx_colsum ? csp : nullptr);
    TVAE_CUDA(cudaGetLastError());
    if (dx_colsum) nll_colsum_final_kernel<<<(C + 31) / 32, 256, 0, stream>>>(csp, NLL_BLOCKS, C, dx_colsum);
  } else {
    if (vec)
      nll_fwd_kernel<4><<<NLL_BLOCKS, 256, 0, stream>>>(xp, x_pitch, xhat, xh_pitch, P, C, loss_type, logvar, batch, dp,
                                                       dx_pitch, ws);
    else
      nll_fwd_kernel<1><<<NLL_BLOCKS, 256, 0, stream>>>(xp, x_pitch, xhat, xh_pitch, P, C, loss_type, logvar, batch, dp,
                                                       dx_pitch, ws);
    TVAE_CUDA(cudaGetLastError());
    if (dx_colsum) {   // generic shapes: a separate pass over the gradient just written (workspace: colsum partials)
      TVAE_CHECK(tvae_colsum_workspace_bytes(P, C) <= (int64_t)NLL_BLOCKS * ((C + 3) / 4 * 4) * 4,
                 "tvae_nll_fwd: workspace too small for the column sums");
      if (tvae_colsum_bf16(dxhat, P, C, dx_pitch, dx_colsum, csp, stream) != 0) return -1;
    }
  }
  TVAE_CUDA(cudaGetLastError());
  nll_final_kernel<<<1, 256, 0, stream>>>(ws, NLL_BLOCKS, sums);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_recon_metrics_workspace_bytes(int32_t N) { return (int64_t)N * RM_CHUNKS * 2 * sizeof(double); }

extern "C" int32_t tvae_recon_metrics(const void* x, int32_t x_pitch, const float* xhat, int32_t xh_pitch, int32_t N,
                                      int32_t HW, int32_t C, float* out, double* ws, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && xhat && out && ws, "tvae_recon_metrics: null pointer");
  TVAE_CHECK(N > 0 && HW > 0 && C > 0 && x_pitch >= C && xh_pitch >= C, "tvae_recon_metrics: bad shape");
  recon_metrics_kernel<<<dim3(RM_CHUNKS, N), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_pitch, xhat,
                                                              xh_pitch, HW, C, ws);
  TVAE_CUDA(cudaGetLastError());
  recon_metrics_final_kernel<<<(N + 127) / 128, 128, 0, stream>>>(ws, N, (double)HW * C, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_l2head_loss_fwd(const float* pred, int32_t pitch, const float* const* targets, int32_t nprod,
                                        int32_t B, int32_t h, int32_t w, double* out, cudaStream_t stream) {
  TVAE_ENTER(pred);
  TVAE_CHECK(pred && targets && out, "tvae_l2head_loss_fwd: null pointer");
  TVAE_CHECK(nprod >= 1 && nprod <= 8, "tvae_l2head_loss_fwd: nprod must be in [1, 8]");
  L2Targets tg;
  for (int i = 0; i < 8; ++i) tg.t[i] = i < nprod ? targets[i] : nullptr;
  l2head_fwd_kernel<<<nprod, 1024, 0, stream>>>(pred, pitch, tg, B, h, w, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_l2head_loss_bwd(const float* pred, int32_t pitch, const float* const* targets, int32_t nprod,
                                        int32_t B, int32_t h, int32_t w, const double* sums, const float* weights,
                                        float grad_scale, void* dpred, int32_t dp_pitch, cudaStream_t stream) {
  TVAE_ENTER(pred);
  TVAE_CHECK(pred && targets && sums && weights && dpred, "tvae_l2head_loss_bwd: null pointer");
  TVAE_CHECK(nprod >= 1 && nprod <= 8 && dp_pitch >= nprod, "tvae_l2head_loss_bwd: bad nprod / pitch");
  L2Targets tg;
  for (int i = 0; i < 8; ++i) tg.t[i] = i < nprod ? targets[i] : nullptr;
  const long long total = (long long)B * h * w * dp_pitch;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  l2head_bwd_kernel<<<grid, 256, 0, stream>>>(pred, pitch, tg, nprod, B, h, w, sums, weights, grad_scale,
                                              reinterpret_cast<__nv_bfloat16*>(dpred), dp_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_sumsq_workspace_bytes(int64_t n) { (void)n; return (int64_t)SUMSQ_BLOCKS * sizeof(double); }

extern "C" int32_t tvae_sumsq(const float* g, int64_t n, double* out, double* ws, cudaStream_t stream) {
  TVAE_ENTER(g);
  TVAE_CHECK(g && out && ws, "tvae_sumsq: null pointer");
  TVAE_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "tvae_sumsq: g must be 16-byte aligned");
  sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, stream>>>(g, n, ws);
  TVAE_CUDA(cudaGetLastError());
  sumsq_final_kernel<<<1, 256, 0, stream>>>(ws, SUMSQ_BLOCKS, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_adamw(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step,
                              const double* sumsq, float max_norm, float grad_scale, cudaStream_t stream) {
  TVAE_ENTER(param);
  TVAE_CHECK(param && grad && exp_avg && exp_avg_sq, "tvae_adamw: null pointer");
  TVAE_CHECK(step >= 1, "tvae_adamw: step is 1-based");
  TVAE_CHECK(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
             "tvae_adamw: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long g4 = (n / 4 + 255) / 256;
  int grid = (int)(g4 < 1 ? 1 : (g4 > 148 * 8 ? 148 * 8 : g4));
  adamw_kernel<<<grid, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                         (float)bc1, (float)sqrt(bc2), sumsq, max_norm, grad_scale);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_vae_loss_finalize(const double* sums, const float* kl, int32_t B, const float* logvar,
                                          double n_elem, float kl_weight, float* out, cudaStream_t stream) {
  TVAE_ENTER(sums);
  TVAE_CHECK(sums && kl && logvar && out, "tvae_vae_loss_finalize: null pointer");
  vae_loss_finalize_kernel<<<1, 32, 0, stream>>>(sums, kl, B, logvar, n_elem, kl_weight, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

namespace tvae {
namespace {
// total = vae_loss + sum_p w_p * mse_p over products with at least one valid pixel (src/model_with_l2.py:161-172)
__global__ void l2head_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ weights, int nprod,
                                       const float* __restrict__ vae_scal, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double total = (double)vae_scal[0];
  for (int p = 0; p < nprod; ++p) {
    const double cnt = sums[2 * p + 1];
    if (cnt > 0.0) {
      const double mse = sums[2 * p] / cnt;
      total += (double)weights[p] * mse;
      out[1 + p] = (float)mse;
    } else {
      out[1 + p] = __int_as_float(0x7fc00000);
    }
  }
  out[0] = (float)total;
}
}  // namespace
}  // namespace tvae

extern "C" int32_t tvae_l2head_finalize(const double* sums, const float* weights, int32_t nprod, const float* vae_scal,
                                        float* out, cudaStream_t stream) {
  TVAE_ENTER(sums);
  TVAE_CHECK(sums && weights && vae_scal && out, "tvae_l2head_finalize: null pointer");
  tvae::l2head_finalize_kernel<<<1, 32, 0, stream>>>(sums, weights, nprod, vae_scal, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
