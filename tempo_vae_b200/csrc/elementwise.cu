// HBM-bound kernels of the TEMPO-VAE hot path: layout conversion, weight packing, GroupNorm(+GELU) forward and
// backward, bias-gradient column sums. All are streaming kernels with 16-byte vector accesses on the NHWC
// channel dimension; reductions are two-stage and order-fixed (bit-reproducible run to run).
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int EW_THREADS = 256;

inline int ew_grid(long long work_items, int per_block = EW_THREADS, int max_blocks = 148 * 16) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------- pack weights
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                   __nv_bfloat16* __restrict__ out_lo, int Crow, int TR, int TK, int C, int c_pad,
                                   long long s_row, long long s_col, long long s_tap) {
  const long long kp = (long long)TK * c_pad;
  const long long total = (long long)TR * Crow * kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / kp;
    const int kk = (int)(i - row * kp);
    const int tk = kk / c_pad, c = kk - tk * c_pad;
    const int tr = (int)(row / Crow), cr = (int)(row - (long long)tr * Crow);
    float v = 0.f;
    if (c < C) v = w[cr * s_row + c * s_col + (tr + tk) * s_tap];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    out[i] = hi;
    if (out_lo) out_lo[i] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

// Every stale pack of a train step in ONE launch: work block wb (PACK_CHUNK elements) belongs to descriptor d with
// block_start[d] <= wb < block_start[d + 1] (binary search), then the same element mapping as above.
constexpr int PACK_CHUNK = 2048;
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const tvae_pack_desc* __restrict__ descs, const long long* __restrict__ block_start, int n,
                            long long total_blocks) {
  for (long long wb = blockIdx.x; wb < total_blocks; wb += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (block_start[mid] <= wb) lo = mid; else hi = mid - 1;
    }
    const tvae_pack_desc d = descs[lo];
    const long long kp = (long long)d.TK * d.c_pad;
    const long long total = (long long)d.TR * d.Crow * kp;
    const long long e0 = (wb - block_start[lo]) * PACK_CHUNK;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
#pragma unroll 2
    for (int k = threadIdx.x; k < PACK_CHUNK; k += 256) {
      const long long i = e0 + k;
      if (i >= total) break;
      const long long row = i / kp;
      const int kk = (int)(i - row * kp);
      const int tk = kk / d.c_pad, c = kk - tk * d.c_pad;
      const int tr = (int)(row / d.Crow), cr = (int)(row - (long long)tr * d.Crow);
      float v = 0.f;
      if (c < d.C) v = d.w[cr * d.s_row + c * d.s_col + (tr + tk) * d.s_tap];
      out[i] = __float2bfloat16(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------- transposes
// NCHW fp32 -> NHWC bf16 (pad lanes [C, pitch) zeroed). Tile: 64 channels x 64 pixels.
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                         __nv_bfloat16* __restrict__ out_lo, int C, int HW, int pitch) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const float* xn = x + (long long)n * C * HW;
  {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
    for (int i = ty; i < 64; i += 4) {
      const int c = c0 + i, p = p0 + tx;
      tile[i][tx] = (c < C && p < HW) ? xn[(long long)c * HW + p] : 0.f;
    }
  }
  __syncthreads();
  {
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int c = c0 + 2 * cx;
#pragma unroll 4
    for (int i = py; i < 64; i += 8) {
      const int p = p0 + i;
      if (p < HW && c < pitch) {
        const long long oidx = ((long long)n * HW + p) * pitch + c;
        __nv_bfloat16* o = out + oidx;
        const float v0 = tile[2 * cx][i], v1 = tile[2 * cx + 1][i];
        if (c + 1 < pitch) {
          *reinterpret_cast<uint32_t*>(o) = pack_bf16(v0, v1);
        } else {
          o[0] = __float2bfloat16(v0);
        }
        if (out_lo) {
          out_lo[oidx] = __float2bfloat16(v0 - __bfloat162float(__float2bfloat16(v0)));
          if (c + 1 < pitch) out_lo[oidx + 1] = __float2bfloat16(v1 - __bfloat162float(__float2bfloat16(v1)));
        }
      }
    }
  }
}

// Same transpose with 16-byte accesses on both sides (HW % 4 == 0, pitch % 8 == 0, no split-bf16 output): float4 loads
// along pixels, 8-channel bf16 stores along channels; 64 channels x 64 pixels per block.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_vec_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int HW, int pitch) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const float* xn = x + (long long)n * C * HW;
  {
    const int pq = (threadIdx.x & 15) << 2, cr = threadIdx.x >> 4;
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + cr + 16 * i, p = p0 + pq;
      v[i] = (c < C && p < HW) ? *reinterpret_cast<const float4*>(xn + (long long)c * HW + p)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* t = &tile[cr + 16 * i][pq];
      t[0] = v[i].x; t[1] = v[i].y; t[2] = v[i].z; t[3] = v[i].w;
    }
  }
  __syncthreads();
  {
    const int c8 = (threadIdx.x & 7) << 3, pr = threadIdx.x >> 3;
    const int c = c0 + c8;
    if (c < pitch) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pl = pr + 32 * i, p = p0 + pl;
        if (p < HW) {
          uint4 o;
          o.x = pack_bf16(tile[c8][pl], tile[c8 + 1][pl]);
          o.y = pack_bf16(tile[c8 + 2][pl], tile[c8 + 3][pl]);
          o.z = pack_bf16(tile[c8 + 4][pl], tile[c8 + 5][pl]);
          o.w = pack_bf16(tile[c8 + 6][pl], tile[c8 + 7][pl]);
          *reinterpret_cast<uint4*>(out + ((long long)n * HW + p) * pitch + c) = o;   // lanes >= C were loaded as 0
        }
      }
    }
  }
}

// NHWC (fp32 or bf16, pitch) -> NCHW fp32
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ out, int C, int HW, int pitch) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  {
    const int cx = threadIdx.x & 63, py = threadIdx.x >> 6;
#pragma unroll 4
    for (int i = py; i < 64; i += 4) {
      const int c = c0 + cx, p = p0 + i;
      float v = 0.f;
      if (c < C && p < HW) v = (float)x[((long long)n * HW + p) * pitch + c];
      tile[i][cx] = v;
    }
  }
  __syncthreads();
  {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    float* on = out + (long long)n * C * HW;
#pragma unroll 4
    for (int i = ty; i < 64; i += 4) {
      const int c = c0 + i, p = p0 + tx;
      if (c < C && p < HW) on[(long long)c * HW + p] = tile[tx][i];
    }
  }
}

__device__ __forceinline__ float bf16_residual(float v) { return v - __bfloat162float(__float2bfloat16(v)); }

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                   __nv_bfloat16* __restrict__ out_lo, long long n) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = o;
    if (out_lo) {
      uint2 l;
      l.x = pack_bf16(bf16_residual(v.x), bf16_residual(v.y));
      l.y = pack_bf16(bf16_residual(v.z), bf16_residual(v.w));
      reinterpret_cast<uint2*>(out_lo)[i] = l;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    out[i] = __float2bfloat16(x[i]);
    if (out_lo) out_lo[i] = __float2bfloat16(bf16_residual(x[i]));
  }
}

// ---------------------------------------------------------------------------------------------- GroupNorm stats
// one block per (n, g); x fp32 [N][HW][C]
template <int VEC>
__global__ void gn_stats_kernel(const float* __restrict__ x, int HW, int C, int G, float eps,
                                float* __restrict__ stats) {
  __shared__ double red[32];
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int gs = C / G;
  const int U = gs / VEC;
  const float* base = x + (long long)n * HW * C + (long long)g * gs;
  const long long total = (long long)HW * U;
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  for (long long e = threadIdx.x; e < total; e += blockDim.x) {
    const long long p = e / U;
    const int u = (int)(e - p * U);
    const float* ptr = base + p * C + u * VEC;
    if (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(ptr);
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    } else {
      const float v = *ptr;
      s1 += v;
      s2 += v * v;
    }
    if (++cnt == 64) {  // bound the fp32 run length, continue in fp64
      d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0;
    }
  }
  d1 += s1; d2 += s2;
  const double t1 = block_sum(d1, red);
  const double t2 = block_sum(d2, red);
  if (threadIdx.x == 0) {
    const double cntd = (double)HW * gs;
    const double mean = t1 / cntd;
    double var = t2 / cntd - mean * mean;
    if (var < 0) var = 0;
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// (mean, rstd) from the conv epilogue's per-tile partial sums; fp64 combine in a fixed order
__global__ void gn_stats_finalize_kernel(const float* __restrict__ part, int spi, int NG, int G, double count, float eps,
                                         float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  const int n = i / G, g = i - n * G;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < spi; ++k) {
    const float* p = part + (((long long)n * spi + k) * G + g) * 2;
    s1 += (double)p[0];
    s2 += (double)p[1];
  }
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0) var = 0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// The same for FEW samples with MANY tiles each (whole-granule inference: one sample, 2,048 tile slots per group at
// 128 x 2048; one thread per (sample, group) walked them serially, 67 us per call): one block per (sample, group), the
// slots strided over the threads, fixed-order fp64 block reduction.
__global__ void __launch_bounds__(256)
gn_stats_finalize_wide_kernel(const float* __restrict__ part, int spi, int G, double count, float eps,
                              float* __restrict__ stats) {
  __shared__ double red1[32], red2[32];
  const int i = blockIdx.x;
  const int n = i / G, g = i - n * G;
  double s1 = 0.0, s2 = 0.0;
  for (int k = threadIdx.x; k < spi; k += blockDim.x) {
    const float* p = part + (((long long)n * spi + k) * G + g) * 2;
    s1 += (double)p[0];
    s2 += (double)p[1];
  }
  s1 = block_sum(s1, red1);
  s2 = block_sum(s2, red2);
  if (threadIdx.x == 0) {
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0) var = 0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// ---------------------------------------------------------------------------------------------- GN apply (+GELU)
template <int VEC>
__global__ void gn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, long long rows,
                                  int HW, int C, int G, int act, __nv_bfloat16* __restrict__ out,
                                  __nv_bfloat16* __restrict__ out_lo) {
  const int gs = C / G;
  const int U = C / VEC;
  const long long total = rows * U;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / U;
    const int c = (int)(i - row * U) * VEC;
    const int n = (int)(row / HW);
    const int g = c / gs;
    const float mean = stats[2 * (n * G + g)], rstd = stats[2 * (n * G + g) + 1];
    const float* xp = x + row * C + c;
    if (VEC == 8) {
      const float4 a = *reinterpret_cast<const float4*>(xp);
      const float4 b = *reinterpret_cast<const float4*>(xp + 4);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + c + 4);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = (v[j] - mean) * rstd * gm[j] + bt[j];
        v[j] = act_f(y, act);
      }
      uint4 o;
      o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
      o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(out + row * C + c) = o;
      if (out_lo) {
        uint4 l;
        l.x = pack_bf16(bf16_residual(v[0]), bf16_residual(v[1])); l.y = pack_bf16(bf16_residual(v[2]), bf16_residual(v[3]));
        l.z = pack_bf16(bf16_residual(v[4]), bf16_residual(v[5])); l.w = pack_bf16(bf16_residual(v[6]), bf16_residual(v[7]));
        *reinterpret_cast<uint4*>(out_lo + row * C + c) = l;
      }
    } else {
      float y = (*xp - mean) * rstd * gamma[c] + beta[c];
      const float av = act_f(y, act);
      out[row * C + c] = __float2bfloat16(av);
      if (out_lo) out_lo[row * C + c] = __float2bfloat16(bf16_residual(av));
    }
  }
}

// ---------------------------------------------------------------------------------------------- GN backward
// Stage A: one block per (n, g). Per-channel sums S1[n][c] = sum dy, S2[n][c] = sum dy * xhat, where
// dy = da * act'(gamma * xhat + beta); then the group means m1 = mean(dy*gamma), m2 = mean(dy*gamma*xhat).
template <int VEC>
__global__ void gn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const __nv_bfloat16* __restrict__ da, int HW, int C, int G, int act, int N,
                                     float* __restrict__ ws) {
  extern __shared__ float sm[];  // [blockDim][2*VEC]
  __shared__ float red[32];
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int gs = C / G;
  const int U = gs / VEC;            // channel units per group (<= blockDim)
  const int lanes = blockDim.x / U;  // pixel lanes
  const int u = threadIdx.x % U, lane = threadIdx.x / U;
  const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
  const int c0 = g * gs + u * VEC;
  float gm[VEC], bt[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    gm[j] = gamma[c0 + j]; bt[j] = beta[c0 + j]; s1[j] = 0.f; s2[j] = 0.f;
  }
  if (lane < lanes) {
    for (int p = lane; p < HW; p += lanes) {
      const long long off = ((long long)n * HW + p) * C + c0;
      float xv[VEC], dv[VEC];
      if (VEC == 8) {
        const float4 a = *reinterpret_cast<const float4*>(x + off);
        const float4 b = *reinterpret_cast<const float4*>(x + off + 4);
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
        const uint4 d = *reinterpret_cast<const uint4*>(da + off);
        const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dv[2 * j] = bf16_bits_to_f(dw[j] & 0xffffu);
          dv[2 * j + 1] = bf16_bits_to_f(dw[j] >> 16);
        }
      } else {
        xv[0] = x[off];
        dv[0] = __bfloat162float(da[off]);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float xh = (xv[j] - mean) * rstd;
        float dy = dv[j];
        if (act) dy *= act_grad_f(xh * gm[j] + bt[j], act);
        s1[j] += dy;
        s2[j] += dy * xh;
      }
    }
  }
  float* mine = sm + (size_t)threadIdx.x * 2 * VEC;
#pragma unroll
  for (int j = 0; j < VEC; ++j) { mine[j] = s1[j]; mine[VEC + j] = s2[j]; }
  __syncthreads();
  // fixed-order reduction over pixel lanes
  float gsum1 = 0.f, gsum2 = 0.f;
  for (int t = threadIdx.x; t < U * 2 * VEC; t += blockDim.x) {
    const int uu = t / (2 * VEC), k = t % (2 * VEC);
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[(size_t)(l * U + uu) * 2 * VEC + k];
    const int c = g * gs + uu * VEC + (k % VEC);
    // ws layout: [2][N][C] per-channel sums, then [N][G][2] group means
    ws[((long long)(k / VEC) * N + n) * C + c] = acc;
    const float gmc = gamma[c];
    if (k < VEC) gsum1 += acc * gmc; else gsum2 += acc * gmc;
  }
  const float t1 = block_sum(gsum1, red);
  const float t2 = block_sum(gsum2, red);
  if (threadIdx.x == 0) {
    const float inv = 1.0f / ((float)HW * (float)gs);
    float* gm_out = ws + 2ll * N * C + 2ll * blockIdx.x;
    gm_out[0] = t1 * inv;
    gm_out[1] = t2 * inv;
  }
}

// dgamma[c] = sum_n S2[n][c], dbeta[c] = sum_n S1[n][c]
__global__ void gn_bwd_param_kernel(const float* __restrict__ ws, int N, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    b += ws[(long long)n * C + c];
    a += ws[((long long)N + n) * C + c];
  }
  dgamma[c] = a;
  dbeta[c] = b;
}

// Stage B: dx = rstd * (dy*gamma - m1 - xhat*m2) (+ gres)
template <int VEC>
__global__ void gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ gres,
                                    const float* __restrict__ gmeans, long long rows, int HW, int C, int G, int act,
                                    __nv_bfloat16* __restrict__ dx) {
  const int gs = C / G;
  const int U = C / VEC;
  const long long total = rows * U;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / U;
    const int c = (int)(i - row * U) * VEC;
    const int n = (int)(row / HW);
    const int g = c / gs;
    const int sg = n * G + g;
    const float mean = stats[2 * sg], rstd = stats[2 * sg + 1];
    const float m1 = gmeans[2 * sg], m2 = gmeans[2 * sg + 1];
    const long long off = row * C + c;
    float xv[VEC], dv[VEC], rv[VEC];
    if (VEC == 8) {
      const float4 a = *reinterpret_cast<const float4*>(x + off);
      const float4 b = *reinterpret_cast<const float4*>(x + off + 4);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
      const uint4 d = *reinterpret_cast<const uint4*>(da + off);
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dv[2 * j] = bf16_bits_to_f(dw[j] & 0xffffu);
        dv[2 * j + 1] = bf16_bits_to_f(dw[j] >> 16);
      }
      if (gres) {
        const uint4 r = *reinterpret_cast<const uint4*>(gres + off);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rv[2 * j] = bf16_bits_to_f(rw[j] & 0xffffu);
          rv[2 * j + 1] = bf16_bits_to_f(rw[j] >> 16);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) rv[j] = 0.f;
      }
    } else {
      xv[0] = x[off];
      dv[0] = __bfloat162float(da[off]);
      rv[0] = gres ? __bfloat162float(gres[off]) : 0.f;
    }
    float o[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float gmj = gamma[c + j];
      const float xh = (xv[j] - mean) * rstd;
      float dy = dv[j];
      if (act) dy *= act_grad_f(xh * gmj + beta[c + j], act);
      o[j] = rstd * (dy * gmj - m1 - xh * m2) + rv[j];
    }
    if (VEC == 8) {
      uint4 w;
      w.x = pack_bf16(o[0], o[1]); w.y = pack_bf16(o[2], o[3]);
      w.z = pack_bf16(o[4], o[5]); w.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + off) = w;
    } else {
      dx[off] = __float2bfloat16(o[0]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- column sums
// block (bx, by): channel-unit tile by, row chunk bx. Threads = UL unit-lanes x RL row-lanes.
template <int VEC>
__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int pitch, int UL,
                                      long long rows_per_block, float* __restrict__ ws) {
  extern __shared__ float sm[];  // [blockDim][VEC]
  const int RL = blockDim.x / UL;
  const int ul = threadIdx.x % UL, rl = threadIdx.x / UL;
  const int c = (blockIdx.y * UL + ul) * VEC;
  const long long r0 = blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (c < C && rl < RL) {
    for (long long r = r0 + rl; r < r1; r += RL) {
      const __nv_bfloat16* p = x + r * pitch + c;
      if (VEC == 8) {
        const uint4 d = *reinterpret_cast<const uint4*>(p);
        const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += bf16_bits_to_f(dw[j] & 0xffffu);
          acc[2 * j + 1] += bf16_bits_to_f(dw[j] >> 16);
        }
      } else if (VEC == 4) {
        const uint2 d = *reinterpret_cast<const uint2*>(p);
        acc[0] += bf16_bits_to_f(d.x & 0xffffu); acc[1] += bf16_bits_to_f(d.x >> 16);
        acc[2] += bf16_bits_to_f(d.y & 0xffffu); acc[3] += bf16_bits_to_f(d.y >> 16);
      } else {
        acc[0] += __bfloat162float(*p);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[(size_t)threadIdx.x * VEC + j] = acc[j];
  __syncthreads();
  for (int t = threadIdx.x; t < UL * VEC; t += blockDim.x) {
    const int uu = t / VEC, j = t % VEC;
    const int cc = (blockIdx.y * UL + uu) * VEC + j;
    if (cc < C) {
      float s = 0.f;
      for (int l = 0; l < RL; ++l) s += sm[(size_t)(l * UL + uu) * VEC + j];
      ws[(long long)blockIdx.x * C + cc] = s;
    }
  }
}
// out[c] = sum over the row blocks of ws[block][c], in a fixed order: 32 channels x 32 block lanes per CTA -- lane l adds
// blocks l, l + 32, ... (up to 19 dependent adds instead of 592: the one-thread-per-channel version of this kernel took
// 23 us per call, 15 calls per train step), then the 32 lane sums are added in lane order.
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ ws, int nblocks, int C, float* __restrict__ out) {
  __shared__ float sa[32][33];
  const int cl = threadIdx.x & 31, l = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (int b = l; b < nblocks; b += 32) s += ws[(long long)b * C + c];
  sa[l][cl] = s;
  __syncthreads();
  if (l == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sa[k][cl];
    out[c] = t;
  }
}

inline int colsum_blocks(long long rows) {
  long long b = (rows + 255) / 256;
  if (b > 592) b = 592;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace tvae

using namespace tvae;

extern "C" int32_t tvae_pack_weight(const float* w, void* out, int32_t Crow, int32_t TR, int32_t TK, int32_t C,
                                    int32_t c_pad, int64_t s_row, int64_t s_col, int64_t s_tap, void* out_lo,
                                    cudaStream_t stream) {
  TVAE_ENTER(w);
  TVAE_CHECK(w && out, "tvae_pack_weight: null pointer");
  TVAE_CHECK(TR == 1 || TK == 1, "tvae_pack_weight: one of TR, TK must be 1");
  TVAE_CHECK(c_pad >= C, "tvae_pack_weight: c_pad < C");
  const long long total = (long long)TR * Crow * TK * c_pad;
  pack_weight_kernel<<<ew_grid(total), EW_THREADS, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out),
                                                               reinterpret_cast<__nv_bfloat16*>(out_lo), Crow, TR, TK, C,
                                                               c_pad, s_row, s_col, s_tap);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_pack_chunk_elems(void) { return PACK_CHUNK; }

extern "C" int32_t tvae_pack_weights_batched(const tvae_pack_desc* descs_device, const int64_t* block_start_device,
                                             int32_t n, int64_t total_blocks, cudaStream_t stream) {
  TVAE_ENTER(descs_device);
  TVAE_CHECK(descs_device && block_start_device && n > 0 && total_blocks > 0, "tvae_pack_weights_batched: bad arguments");
  long long grid = total_blocks;
  if (grid > 148 * 16) grid = 148 * 16;
  pack_weights_batched_kernel<<<(int)grid, 256, 0, stream>>>(descs_device,
                                                            reinterpret_cast<const long long*>(block_start_device), n,
                                                            total_blocks);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_nchw_f32_to_nhwc_bf16(const float* x, void* out, int32_t N, int32_t C, int32_t HW,
                                              int32_t pitch, void* out_lo, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out, "tvae_nchw_f32_to_nhwc_bf16: null pointer");
  TVAE_CHECK(pitch >= C && pitch % 2 == 0, "tvae_nchw_f32_to_nhwc_bf16: bad pitch");
  dim3 grid((HW + 63) / 64, (pitch + 63) / 64, N);
  if (!out_lo && HW % 4 == 0 && pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    nchw_to_nhwc_bf16_vec_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), C, HW, pitch);
  else
    nchw_to_nhwc_bf16_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out),
                                                       reinterpret_cast<__nv_bfloat16*>(out_lo), C, HW, pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int32_t tvae_nhwc_f32_to_nchw_f32(const float* x, float* out, int32_t N, int32_t C, int32_t HW,
                                             int32_t pitch, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out, "tvae_nhwc_f32_to_nchw_f32: null pointer");
  dim3 grid((HW + 63) / 64, (C + 63) / 64, N);
  nhwc_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(x, out, C, HW, pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int32_t tvae_nhwc_bf16_to_nchw_f32(const void* x, float* out, int32_t N, int32_t C, int32_t HW,
                                              int32_t pitch, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out, "tvae_nhwc_bf16_to_nchw_f32: null pointer");
  dim3 grid((HW + 63) / 64, (C + 63) / 64, N);
  nhwc_to_nchw_kernel<__nv_bfloat16>
      <<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, C, HW, pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int32_t tvae_f32_to_bf16(const float* x, void* out, int64_t n, void* out_lo, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out, "tvae_f32_to_bf16: null pointer");
  f32_to_bf16_kernel<<<ew_grid(n / 4 + 1), EW_THREADS, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out),
                                                                   reinterpret_cast<__nv_bfloat16*>(out_lo), n);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

// channels-last fp32 rows (pitch in_pitch) -> channels-last bf16 rows (pitch out_pitch), pad lanes [C, out_pitch) zeroed;
// a pair of channels per thread (out_pitch is even), 64-bit loads when the input rows allow it
__global__ void nhwc_f32_to_bf16_kernel(const float* __restrict__ x, long long in_pitch, long long rows, int C,
                                        __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo,
                                        int out_pitch, int vec2) {
  const int U = out_pitch >> 1;
  const long long total = rows * U;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long sp = stride / U;
  const int su = (int)(stride - sp * U);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long r = i / U;
  int u = (int)(i - r * U);
  for (; i < total; i += stride, r += sp, u += su) {
    if (u >= U) { u -= U; ++r; }
    const int c = u << 1;
    float v0 = 0.f, v1 = 0.f;
    const float* src = x + r * in_pitch + c;
    if (c + 1 < C) {
      if (vec2) {
        const float2 t = *reinterpret_cast<const float2*>(src);
        v0 = t.x; v1 = t.y;
      } else {
        v0 = src[0]; v1 = src[1];
      }
    } else if (c < C) {
      v0 = src[0];
    }
    *reinterpret_cast<uint32_t*>(out + r * out_pitch + c) = pack_bf16(v0, v1);
    if (out_lo)
      *reinterpret_cast<uint32_t*>(out_lo + r * out_pitch + c) = pack_bf16(bf16_residual(v0), bf16_residual(v1));
  }
}

// z = clamp((log(max(rad, min_rad)) - mean[c]) / (std[c] + 1e-8), lo, hi): raw radiance rows [rows][C] -> z-scored
// log-radiance, as fp32 rows (same pitch C) and/or as the bf16 operand rows the conv kernels read (pitch out_pitch)
__global__ void normalize_radiance_kernel(const float* __restrict__ rad, const float* __restrict__ mean,
                                          const float* __restrict__ stdv, long long rows, int C, float min_rad,
                                          float lo, float hi, float* __restrict__ out_f32,
                                          __nv_bfloat16* __restrict__ out_bf16, int out_pitch) {
  const int U = (out_bf16 ? out_pitch : C + (C & 1)) >> 1;     // channel pairs per row (pad lanes of the bf16 rows incl.)
  const long long total = rows * U;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long sp = stride / U;
  const int su = (int)(stride - sp * U);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long r = i / U;
  int u = (int)(i - r * U);
  for (; i < total; i += stride, r += sp, u += su) {
    if (u >= U) { u -= U; ++r; }
    const int c = u << 1;
    float z[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (c + j < C) {
        const float v = logf(fmaxf(rad[r * C + c + j], min_rad));
        z[j] = fminf(fmaxf((v - mean[c + j]) / (stdv[c + j] + 1e-8f), lo), hi);
        if (out_f32) out_f32[r * C + c + j] = z[j];
      }
    }
    if (out_bf16) *reinterpret_cast<uint32_t*>(out_bf16 + r * out_pitch + c) = pack_bf16(z[0], z[1]);
  }
}

extern "C" int32_t tvae_normalize_radiance(const float* rad, const float* mean, const float* stdv, int64_t rows,
                                           int32_t C, float min_radiance, float clip_min, float clip_max,
                                           float* out_f32, void* out_bf16, int32_t out_pitch, cudaStream_t stream) {
  TVAE_ENTER(rad);
  TVAE_CHECK(rad && mean && stdv && (out_f32 || out_bf16), "tvae_normalize_radiance: null pointer");
  TVAE_CHECK(!out_bf16 || (out_pitch >= C && out_pitch % 2 == 0), "tvae_normalize_radiance: bad out_pitch");
  const long long pairs = rows * ((out_bf16 ? out_pitch : C + 1) / 2);
  normalize_radiance_kernel<<<ew_grid(pairs), EW_THREADS, 0, stream>>>(
      rad, mean, stdv, rows, C, min_radiance, clip_min, clip_max, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16),
      out_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_nhwc_f32_to_nhwc_bf16(const float* x, int64_t in_pitch, int64_t rows, int32_t C, void* out,
                                              int32_t out_pitch, void* out_lo, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out, "tvae_nhwc_f32_to_nhwc_bf16: null pointer");
  TVAE_CHECK(in_pitch >= C && out_pitch >= C && out_pitch % 2 == 0, "tvae_nhwc_f32_to_nhwc_bf16: bad pitch");
  const int vec2 = (in_pitch % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  nhwc_f32_to_bf16_kernel<<<ew_grid(rows * (out_pitch / 2)), EW_THREADS, 0, stream>>>(
      x, in_pitch, rows, C, reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<__nv_bfloat16*>(out_lo), out_pitch,
      vec2);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_gn_stats(const float* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, float* stats,
                                 cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && stats, "tvae_gn_stats: null pointer");
  TVAE_CHECK(G > 0 && C % G == 0, "tvae_gn_stats: C %% G != 0");
  const int gs = C / G;
  if (gs % 4 == 0 && C % 4 == 0)
    gn_stats_kernel<4><<<N * G, 512, 0, stream>>>(x, HW, C, G, eps, stats);
  else
    gn_stats_kernel<1><<<N * G, 512, 0, stream>>>(x, HW, C, G, eps, stats);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_gn_stats_finalize(const float* part, int32_t spi, int32_t N, int32_t G, double count, float eps,
                                          float* stats, cudaStream_t stream) {
  TVAE_ENTER(part);
  TVAE_CHECK(part && stats && spi > 0 && N > 0 && G > 0, "tvae_gn_stats_finalize: bad arguments");
  const int NG = N * G;
  if (spi >= 256 && NG <= 1024)    // few samples, many tile slots each: a block per (sample, group)
    gn_stats_finalize_wide_kernel<<<NG, 256, 0, stream>>>(part, spi, G, count, eps, stats);
  else
    gn_stats_finalize_kernel<<<(NG + 127) / 128, 128, 0, stream>>>(part, spi, NG, G, count, eps, stats);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_gn_act_fwd(const void* xv, int32_t x_is_bf16, const float* stats, const float* gamma,
                                   const float* beta, int32_t N, int32_t HW, int32_t C, int32_t G, int32_t act,
                                   void* out, void* out_lo, cudaStream_t stream) {
  return tvae_gn_act_fwd2(xv, x_is_bf16, stats, gamma, beta, N, HW, C, G, act, out, out_lo, nullptr, stream);
}

extern "C" int32_t tvae_gn_act_fwd2(const void* xv, int32_t x_is_bf16, const float* stats, const float* gamma,
                                    const float* beta, int32_t N, int32_t HW, int32_t C, int32_t G, int32_t act,
                                    void* out, void* out_lo, void* act_grad, cudaStream_t stream) {
  TVAE_ENTER(xv);
  TVAE_CHECK(xv && stats && gamma && beta && out, "tvae_gn_act_fwd: null pointer");
  TVAE_CHECK(G > 0 && C % G == 0, "tvae_gn_act_fwd: C %% G != 0");
  const long long rows = (long long)N * HW;
  const int gs = C / G;
  TVAE_CHECK(!act_grad || (gn_fast_ok(C, G) && out_lo == nullptr && act != 0),
             "tvae_gn_act_fwd2: act_grad needs an activation, the fast-path geometry and no split-bf16 output");
  if (gn_fast_ok(C, G) && out_lo == nullptr) {   // the split-bf16 ("fp32 mode") output uses the generic kernel
    gn_act_fwd_fast(xv, x_is_bf16 != 0, stats, gamma, beta, N, HW, C, G, act, reinterpret_cast<__nv_bfloat16*>(out),
                    reinterpret_cast<__nv_bfloat16*>(act_grad), stream);
    TVAE_CUDA(cudaGetLastError());
    return 0;
  }
  TVAE_CHECK(!x_is_bf16, "tvae_gn_act_fwd: a bf16 input needs the fast-path geometry (C/8 dividing 256, groups of whole octets)");
  const float* x = reinterpret_cast<const float*>(xv);
  if (gs % 8 == 0)
    gn_act_fwd_kernel<8><<<ew_grid(rows * (C / 8)), EW_THREADS, 0, stream>>>(
        x, stats, gamma, beta, rows, HW, C, G, act, reinterpret_cast<__nv_bfloat16*>(out),
        reinterpret_cast<__nv_bfloat16*>(out_lo));
  else
    gn_act_fwd_kernel<1><<<ew_grid(rows * C), EW_THREADS, 0, stream>>>(x, stats, gamma, beta, rows, HW, C, G, act,
                                                                        reinterpret_cast<__nv_bfloat16*>(out),
                                                                        reinterpret_cast<__nv_bfloat16*>(out_lo));
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_gn_bwd_workspace_bytes(int32_t N, int32_t HW, int32_t C, int32_t G) {
  if (gn_fast_ok(C, G)) return gn_bwd_fast_ws_floats(N, HW, C, G) * 4;
  // generic kernels: sums + group means, then the partials of a column-sum pass over dx (dx_colsum)
  return (2ll * N * C + 2ll * N * G) * 4 + (int64_t)colsum_blocks((long long)N * HW) * C * 4;
}

extern "C" int32_t tvae_gn_act_bwd(const void* xv, int32_t x_is_bf16, const float* stats, const float* gamma,
                                   const float* beta, const void* da, const void* gres, int32_t N, int32_t HW, int32_t C, int32_t G,
                                   int32_t act, void* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws,
                                   cudaStream_t stream) {
  return tvae_gn_act_bwd2(xv, x_is_bf16, stats, gamma, beta, da, gres, nullptr, N, HW, C, G, act, dx, dgamma, dbeta,
                          dx_colsum, ws, stream);
}

extern "C" int32_t tvae_gn_act_bwd2(const void* xv, int32_t x_is_bf16, const float* stats, const float* gamma,
                                    const float* beta, const void* da, const void* gres, const void* act_grad, int32_t N,
                                    int32_t HW, int32_t C, int32_t G, int32_t act, void* dx, float* dgamma, float* dbeta,
                                    float* dx_colsum, float* ws, cudaStream_t stream) {
  TVAE_ENTER(xv);
  TVAE_CHECK(xv && stats && gamma && beta && da && dx && dgamma && dbeta && ws, "tvae_gn_act_bwd: null pointer");
  TVAE_CHECK(G > 0 && C % G == 0, "tvae_gn_act_bwd: C %% G != 0");
  const int gs = C / G;
  const long long rows = (long long)N * HW;
  const __nv_bfloat16* dap = reinterpret_cast<const __nv_bfloat16*>(da);
  const __nv_bfloat16* grp = reinterpret_cast<const __nv_bfloat16*>(gres);
  __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
  const float* gmeans = ws + 2ll * N * C;
  if (gn_fast_ok(C, G)) {
    gn_act_bwd_fast(xv, x_is_bf16 != 0, stats, gamma, beta, dap, grp, reinterpret_cast<const __nv_bfloat16*>(act_grad), N,
                    HW, C, G, act, dxp, dgamma, dbeta, dx_colsum, ws, stream);
    TVAE_CUDA(cudaGetLastError());
    return 0;
  }
  TVAE_CHECK(!x_is_bf16, "tvae_gn_act_bwd: a bf16 input needs the fast-path geometry");
  const float* x = reinterpret_cast<const float*>(xv);
  if (gs % 8 == 0 && gs / 8 <= 256) {
    const int threads = 256;
    gn_bwd_reduce_kernel<8><<<N * G, threads, threads * 16 * sizeof(float), stream>>>(x, stats, gamma, beta, dap, HW, C,
                                                                                     G, act, N, ws);
    TVAE_CUDA(cudaGetLastError());
    gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ws, N, C, dgamma, dbeta);
    gn_bwd_apply_kernel<8><<<ew_grid(rows * (C / 8)), EW_THREADS, 0, stream>>>(x, stats, gamma, beta, dap, grp, gmeans,
                                                                              rows, HW, C, G, act, dxp);
  } else {
    TVAE_CHECK(gs <= 256, "tvae_gn_act_bwd: group size %d > 256 needs a multiple of 8", gs);
    const int threads = 256;
    gn_bwd_reduce_kernel<1><<<N * G, threads, threads * 2 * sizeof(float), stream>>>(x, stats, gamma, beta, dap, HW, C,
                                                                                    G, act, N, ws);
    TVAE_CUDA(cudaGetLastError());
    gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ws, N, C, dgamma, dbeta);
    gn_bwd_apply_kernel<1><<<ew_grid(rows * C), EW_THREADS, 0, stream>>>(x, stats, gamma, beta, dap, grp, gmeans, rows,
                                                                        HW, C, G, act, dxp);
  }
  TVAE_CUDA(cudaGetLastError());
  if (dx_colsum)   // generic shapes: a separate pass over dx
    return tvae_colsum_bf16(dx, rows, C, C, dx_colsum, ws + 2ll * N * C + 2ll * N * G, stream);
  return 0;
}

extern "C" int32_t tvae_gn_set_bwd_fused(int32_t on, int32_t group_mb) {
  gn_set_bwd_fused(on, group_mb);
  return 0;
}

extern "C" int64_t tvae_colsum_workspace_bytes(int64_t rows, int32_t C) {
  return (int64_t)colsum_blocks(rows) * C * 4;
}

extern "C" int32_t tvae_colsum_bf16(const void* x, int64_t rows, int32_t C, int32_t pitch, float* out, float* ws,
                                    cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out && ws, "tvae_colsum_bf16: null pointer");
  const int nb = colsum_blocks(rows);
  const long long rpb = (rows + nb - 1) / nb;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const int vec = (C % 8 == 0 && pitch % 8 == 0) ? 8 : ((C % 4 == 0 && pitch % 4 == 0) ? 4 : 1);
  const int U = (C + vec - 1) / vec;
  int UL = 1;
  while (UL < U && UL < 256) UL <<= 1;  // power of two <= 256 so it divides the block
  dim3 grid(nb, (U + UL - 1) / UL);
  const size_t smem = 256 * vec * sizeof(float);
  if (vec == 8) colsum_partial_kernel<8><<<grid, 256, smem, stream>>>(xp, rows, C, pitch, UL, rpb, ws);
  else if (vec == 4) colsum_partial_kernel<4><<<grid, 256, smem, stream>>>(xp, rows, C, pitch, UL, rpb, ws);
  else colsum_partial_kernel<1><<<grid, 256, smem, stream>>>(xp, rows, C, pitch, UL, rpb, ws);
  TVAE_CUDA(cudaGetLastError());
  colsum_final_kernel<<<(C + 31) / 32, 1024, 0, stream>>>(ws, nb, C, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
