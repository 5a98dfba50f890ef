// Bandwidth-tuned GroupNorm(+activation) kernels for the common case (group size a multiple of 8 channels,
// C/8 a divisor of 256): no integer division in the inner loops, per-thread constants hoisted (scale/shift per
// channel, statistics per group), two rows in flight per thread, 16/32-byte vector accesses, fast exact GELU.
// Selected by the tvae_gn_* entry points in elementwise.cu; the generic kernels there remain the fallback.
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {

namespace {

__device__ __forceinline__ float act_fast(float y, int act) { return act == 1 ? gelu_fast(y) : act_f(y, act); }
__device__ __forceinline__ float act_grad_fast(float y, int act) { return act == 1 ? gelu_grad_fast(y) : act_grad_f(y, act); }

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 d = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = bf16_bits_to_f(w[j] & 0xffffu);
    v[2 * j + 1] = bf16_bits_to_f(w[j] >> 16);
  }
}
// 8 consecutive channels of the GroupNorm input: fp32 (residual stream) or bf16 (a conv output that only feeds a norm)
__device__ __forceinline__ void load8x(const float* p, float (&v)[8]) { load8(p, v); }
__device__ __forceinline__ void load8x(const __nv_bfloat16* p, float (&v)[8]) { load8_bf16(p, v); }
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
  o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// grid (row chunks, N); 256 threads = (256/U) row lanes x U channel-octets, U = C/8
template <typename TX>
__global__ void __launch_bounds__(256)
gn_act_fwd_fast_kernel(const TX* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, int HW, int C, int G, int act, int rpb,
                       __nv_bfloat16* __restrict__ out) {
  const int U = C >> 3, lanes = 256 / U;
  const int u = threadIdx.x % U, lane = threadIdx.x / U;
  const int c = u << 3, n = blockIdx.y;
  const int g = c / (C / G);
  const float mean = stats[2 * (n * G + g)], rstd = stats[2 * (n * G + g) + 1];
  float sc[8], sh[8];
  load8(gamma + c, sc);
  load8(beta + c, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] *= rstd;
    sh[j] = fmaf(-mean, sc[j], sh[j]);
  }
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(r0 + rpb, HW);
  const long long base = (long long)n * HW * C + c;
  int r = r0 + lane;
  for (; r + lanes < r1; r += 2 * lanes) {
    float a[8], b[8];
    load8x(x + base + (long long)r * C, a);
    load8x(x + base + (long long)(r + lanes) * C, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = act_fast(fmaf(a[j], sc[j], sh[j]), act);
      b[j] = act_fast(fmaf(b[j], sc[j], sh[j]), act);
    }
    store8_bf16(out + base + (long long)r * C, a);
    store8_bf16(out + base + (long long)(r + lanes) * C, b);
  }
  if (r < r1) {
    float a[8];
    load8x(x + base + (long long)r * C, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = act_fast(fmaf(a[j], sc[j], sh[j]), act);
    store8_bf16(out + base + (long long)r * C, a);
  }
}

// dx = rstd * (dy*gamma - m1 - xhat*m2) (+ gres), dy = da * act'(gamma*xhat + beta)
// cs_part (optional): per-block column sums of dx (before its bf16 rounding), [n][chunk][C] -- dx is the output
// gradient of the conv that produced x, so its column sums are that conv's bias gradient and the separate pass over
// dx (tvae_colsum_bf16) disappears.
template <typename TX>
__global__ void __launch_bounds__(256)
gn_bwd_apply_fast_kernel(const TX* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const __nv_bfloat16* __restrict__ da,
                         const __nv_bfloat16* __restrict__ gres, const float* __restrict__ gmeans, int HW, int C, int G,
                         int act, int rpb, __nv_bfloat16* __restrict__ dx, float* __restrict__ cs_part) {
  __shared__ float cs_sm[256 * 8];
  const int U = C >> 3, lanes = 256 / U;
  const int u = threadIdx.x % U, lane = threadIdx.x / U;
  const int c = u << 3, n = blockIdx.y;
  const int sg = n * G + c / (C / G);
  const float mean = stats[2 * sg], rstd = stats[2 * sg + 1];
  // everything that multiplies x is folded into per-thread constants, so an element costs one FMA for the activation
  // argument y = x*sc + sh, one for the mean terms t = x*ta + tb (= rstd*(m1 + xhat*m2)), act', and two for dx
  const float ta = gmeans[2 * sg + 1] * rstd * rstd;
  const float tb = gmeans[2 * sg] * rstd - mean * ta;
  float sc[8], sh[8], gr[8], cs[8];
  load8(gamma + c, sc);
  load8(beta + c, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gr[j] = sc[j] * rstd;
    sc[j] = gr[j];
    sh[j] = fmaf(-mean, sc[j], sh[j]);
    cs[j] = 0.f;
  }
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(r0 + rpb, HW);
  const long long base = (long long)n * HW * C + c;
  for (int r = r0 + lane; r < r1; r += lanes) {
    const long long off = base + (long long)r * C;
    float xv[8], dv[8], rv[8];
    load8x(x + off, xv);
    load8_bf16(da + off, dv);
    if (gres) {
      load8_bf16(gres + off, rv);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) rv[j] = 0.f;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dy = dv[j];
      if (act) dy *= act_grad_fast(fmaf(xv[j], sc[j], sh[j]), act);
      o[j] = fmaf(dy, gr[j], rv[j]) - fmaf(xv[j], ta, tb);
      cs[j] += o[j];
    }
    store8_bf16(dx + off, o);
  }
  if (cs_part) {
#pragma unroll
    for (int j = 0; j < 8; ++j) cs_sm[threadIdx.x * 8 + j] = cs[j];
    __syncthreads();
    float* out = cs_part + ((long long)n * gridDim.x + blockIdx.x) * C;
    for (int t = threadIdx.x; t < U * 8; t += 256) {
      const int uu = t >> 3, k = t & 7;
      float acc = 0.f;
      for (int l = 0; l < lanes; ++l) acc += cs_sm[(l * U + uu) * 8 + k];   // fixed order
      out[(uu << 3) + k] = acc;
    }
  }
}

// column sums of a [rows][C] fp32 matrix in two fixed-order stages: slice sums, then the sum of the slices
constexpr int CS_SLICES = 32;
__global__ void __launch_bounds__(256)
colsum_rows_slice_kernel(const float* __restrict__ part, int rows, int C, float* __restrict__ slices) {
  __shared__ float sa[8][33];
  const int cx = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int per = (rows + CS_SLICES - 1) / CS_SLICES;
  const int r0 = blockIdx.y * per, r1 = min(r0 + per, rows);
  float a = 0.f;
  if (c < C)
    for (int r = r0 + rl; r < r1; r += 8) a += part[(long long)r * C + c];
  sa[rl][cx] = a;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += sa[l][cx];
    slices[(long long)blockIdx.y * C + c] = t;
  }
}
__global__ void colsum_rows_final_kernel(const float* __restrict__ slices, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
#pragma unroll 8
  for (int s = 0; s < CS_SLICES; ++s) t += slices[(long long)s * C + c];
  out[c] = t;
}

// Stage A1: grid (row chunks, N) like the apply kernel (full 2 KB rows => long DRAM bursts): per-channel partial sums
// of dy and dy*xhat over the chunk's rows -> part[n][chunk][2][C].
template <typename TX>
__global__ void __launch_bounds__(256, 4)
gn_bwd_rowsum_fast_kernel(const TX* __restrict__ x, const float* __restrict__ stats,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const __nv_bfloat16* __restrict__ da, int HW, int C, int G, int act, int rpb,
                          float* __restrict__ part) {
  extern __shared__ float sm[];  // [256][16]
  const int U = C >> 3, lanes = 256 / U;
  const int u = threadIdx.x % U, lane = threadIdx.x / U;
  const int c = u << 3, n = blockIdx.y;
  const int sg = n * G + c / (C / G);
  const float mean = stats[2 * sg], rstd = stats[2 * sg + 1];
  // y = x*sc + sh (sc = gamma*rstd, sh = beta - mean*sc); sum(dy*xhat) = rstd*(sum(dy*x) - mean*sum(dy)) is formed
  // once per thread at the end, so the loop body is one FMA for y, act', and two accumulations per element
  float sc[8], sh[8], s1[8], s2[8];
  load8(gamma + c, sc);
  load8(beta + c, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] *= rstd;
    sh[j] = fmaf(-mean, sc[j], sh[j]);
    s1[j] = 0.f; s2[j] = 0.f;
  }
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(r0 + rpb, HW);
  const long long base = (long long)n * HW * C + c;
  for (int r = r0 + lane; r < r1; r += lanes) {
    float xa[8], da_[8];
    load8x(x + base + (long long)r * C, xa);
    load8_bf16(da + base + (long long)r * C, da_);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dy = da_[j];
      if (act) dy *= act_grad_fast(fmaf(xa[j], sc[j], sh[j]), act);
      s1[j] += dy;
      s2[j] = fmaf(dy, xa[j], s2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s2[j] = (s2[j] - mean * s1[j]) * rstd;     // sum dy*x  ->  sum dy*xhat
  float* mine = sm + (size_t)threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[j] = s1[j]; mine[8 + j] = s2[j]; }
  __syncthreads();
  float* out = part + ((long long)n * gridDim.x + blockIdx.x) * 2 * C;
  for (int t = threadIdx.x; t < U * 16; t += 256) {
    const int uu = t >> 4, k = t & 15;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[(size_t)(l * U + uu) * 16 + k];   // fixed order
    out[(k >> 3) * C + (uu << 3) + (k & 7)] = acc;
  }
}

// Stage A2: one block per sample: S1/S2[n][c] = sum over chunks; group means m1 = mean(dy*gamma), m2 = mean(dy*gamma*xhat)
__global__ void __launch_bounds__(256)
gn_bwd_finalize_fast_kernel(const float* __restrict__ part, const float* __restrict__ gamma, int chunks, int HW, int C,
                            int G, int N, float* __restrict__ ws, const float* __restrict__ raw_stats) {
  extern __shared__ float sm[];  // [2][C]
  const int n = blockIdx.x;
  const float* pn = part + (long long)n * chunks * 2 * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < chunks; ++k) {
      a += pn[(long long)k * 2 * C + c];
      b += pn[(long long)k * 2 * C + C + c];
    }
    if (raw_stats) {   // partials hold sum(dy * x): convert to sum(dy * xhat)
      const int sg = n * G + c / (C / G);
      b = (b - raw_stats[2 * sg] * a) * raw_stats[2 * sg + 1];
    }
    ws[(long long)n * C + c] = a;                     // S1 (sum dy)
    ws[((long long)N + n) * C + c] = b;               // S2 (sum dy*xhat)
    const float gmc = gamma[c];
    sm[c] = a * gmc;
    sm[C + c] = b * gmc;
  }
  __syncthreads();
  const int gs = C / G;
  for (int g = threadIdx.x; g < G; g += 256) {
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < gs; ++j) { t1 += sm[g * gs + j]; t2 += sm[C + g * gs + j]; }
    const float inv = 1.0f / ((float)HW * (float)gs);
    float* gm_out = ws + 2ll * N * C + 2ll * (n * G + g);
    gm_out[0] = t1 * inv;
    gm_out[1] = t2 * inv;
  }
}

// dgamma[c] = sum_n S2[n][c], dbeta[c] = sum_n S1[n][c]; block = 32 channels x 8 sample lanes
__global__ void __launch_bounds__(256)
gn_bwd_param_fast_kernel(const float* __restrict__ ws, int N, int C, float* __restrict__ dgamma,
                         float* __restrict__ dbeta) {
  __shared__ float sa[8][33], sb[8][33];
  const int cx = threadIdx.x & 31, nl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int n = nl; n < N; n += 8) {
      b += ws[(long long)n * C + c];
      a += ws[((long long)N + n) * C + c];
    }
  }
  sa[nl][cx] = a;
  sb[nl][cx] = b;
  __syncthreads();
  if (nl == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) { ta += sa[l][cx]; tb += sb[l][cx]; }
    dgamma[c] = ta;
    dbeta[c] = tb;
  }
}

inline int rows_per_block(int HW) { return HW < 128 ? HW : 128; }

}  // namespace

bool gn_fast_ok(int C, int G) {
  if (G <= 0 || C % G) return false;
  const int gs = C / G;
  if (gs % 8) return false;
  const int U = C / 8;
  return U >= 1 && U <= 256 && (256 % U) == 0 && (256 % (gs / 8)) == 0;
}

int gn_act_fwd_fast(const void* x, bool x_bf16, const float* stats, const float* gamma, const float* beta, int N,
                    int HW, int C, int G, int act, __nv_bfloat16* out, cudaStream_t stream) {
  const int rpb = rows_per_block(HW);
  dim3 grid((HW + rpb - 1) / rpb, N);
  if (x_bf16)
    gn_act_fwd_fast_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), stats, gamma, beta, HW,
                                                     C, G, act, rpb, out);
  else
    gn_act_fwd_fast_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), stats, gamma, beta, HW, C, G,
                                                     act, rpb, out);
  return 0;
}

long long gn_bwd_fast_ws_floats(int N, int HW, int C, int G) {
  const int rpb = rows_per_block(HW);
  const long long chunks = (HW + rpb - 1) / rpb;
  return 2ll * N * C + 2ll * N * G + (long long)N * chunks * 2 * C + (long long)CS_SLICES * C;
}

int gn_act_bwd_fast(const void* xv, bool x_bf16, const float* stats, const float* gamma, const float* beta,
                    const __nv_bfloat16* da, const __nv_bfloat16* gres, int N, int HW, int C, int G, int act,
                    __nv_bfloat16* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws,
                    cudaStream_t stream) {
  const int rpb = rows_per_block(HW);
  const int chunks = (HW + rpb - 1) / rpb;
  dim3 grid(chunks, N);
  float* part = ws + 2ll * N * C + 2ll * N * G;
  const float* x = reinterpret_cast<const float*>(xv);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv);
  if (x_bf16)
    gn_bwd_rowsum_fast_kernel<<<grid, 256, 256 * 16 * sizeof(float), stream>>>(xb, stats, gamma, beta, da, HW, C, G,
                                                                              act, rpb, part);
  else
    gn_bwd_rowsum_fast_kernel<<<grid, 256, 256 * 16 * sizeof(float), stream>>>(x, stats, gamma, beta, da, HW, C, G,
                                                                              act, rpb, part);
  gn_bwd_finalize_fast_kernel<<<N, 256, 2 * C * sizeof(float), stream>>>(part, gamma, chunks, HW, C, G, N, ws, nullptr);
  gn_bwd_param_fast_kernel<<<(C + 31) / 32, 256, 0, stream>>>(ws, N, C, dgamma, dbeta);
  // the row-sum partials are consumed by now (stream order): their region is reused for the column sums of dx
  if (x_bf16)
    gn_bwd_apply_fast_kernel<<<grid, 256, 0, stream>>>(xb, stats, gamma, beta, da, gres, ws + 2ll * N * C, HW, C, G, act,
                                                       rpb, dx, dx_colsum ? part : nullptr);
  else
    gn_bwd_apply_fast_kernel<<<grid, 256, 0, stream>>>(x, stats, gamma, beta, da, gres, ws + 2ll * N * C, HW, C, G, act,
                                                       rpb, dx, dx_colsum ? part : nullptr);
  if (dx_colsum) {
    float* slices = part + (long long)N * chunks * 2 * C;
    colsum_rows_slice_kernel<<<dim3((C + 31) / 32, CS_SLICES), 256, 0, stream>>>(part, N * chunks, C, slices);
    colsum_rows_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(slices, C, dx_colsum);
  }
  return 0;
}

}  // namespace tvae
