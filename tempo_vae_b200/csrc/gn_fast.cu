// Bandwidth-tuned GroupNorm(+activation) kernels for the common case (group size a multiple of 8 channels,
// C/8 a divisor of 256): no integer division in the inner loops, per-thread constants hoisted (scale/shift per
// channel, statistics per group), two rows in flight per thread, 16/32-byte vector accesses, fast exact GELU.
// Selected by the tvae_gn_* entry points in elementwise.cu; the generic kernels there remain the fallback.
#include <stdlib.h>
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
// value and derivative of the activation in one evaluation (GELU: both come out of the same Phi / phi pair)
__device__ __forceinline__ void act_and_grad_fast(float y, int act, float& a, float& g) {
  if (act == 1) {
    float q, pe;
    gelu_q(y, q, pe);
    a = gelu_from_q(y, q);                                    // the same value gelu_fast produces, bit for bit
    g = fmaf(y, pe, (y >= 0.f) ? 1.0f - q : q);
  } else {
    a = act_f(y, act);
    g = act_grad_f(y, act);
  }
}

// GP: also store act'(y) (bf16) for the backward pass, which then never evaluates the activation again
template <typename TX, bool GP>
__global__ void __launch_bounds__(256)
gn_act_fwd_fast_kernel(const TX* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, int HW, int C, int G, int act, int rpb,
                       __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ gp_out) {
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
  if (GP) {
    for (; r < r1; r += lanes) {
      float a[8], g[8];
      load8x(x + base + (long long)r * C, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) act_and_grad_fast(fmaf(a[j], sc[j], sh[j]), act, a[j], g[j]);
      store8_bf16(out + base + (long long)r * C, a);
      store8_bf16(gp_out + base + (long long)r * C, g);
    }
  } else {
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
}

// dx = rstd * (dy*gamma - m1 - xhat*m2) (+ gres), dy = da * act'(gamma*xhat + beta)
// cs_part (optional): per-block column sums of dx (before its bf16 rounding), [n][chunk][C] -- dx is the output
// gradient of the conv that produced x, so its column sums are that conv's bias gradient and the separate pass over
// dx (tvae_colsum_bf16) disappears.
// FROM_DY: the row-sum pass already left dy = da * act'(y) (bf16) in the dx buffer; this pass reads it back element by
// element (each thread overwrites exactly what it read), so act' is evaluated once per element instead of twice and
// `da` is not read again.
template <typename TX, bool FROM_DY>
__global__ void __launch_bounds__(256)
gn_bwd_apply_fast_kernel(const TX* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const __nv_bfloat16* __restrict__ da,
                         const __nv_bfloat16* __restrict__ gres, const float* __restrict__ gmeans, int HW, int C, int G,
                         int act, int rpb, __nv_bfloat16* dx, float* __restrict__ cs_part) {
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
    load8_bf16((FROM_DY ? dx : da) + off, dv);
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
      if (!FROM_DY && act) dy *= act_grad_fast(fmaf(xv[j], sc[j], sh[j]), act);
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
                          float* __restrict__ part, __nv_bfloat16* __restrict__ dy_out,
                          const __nv_bfloat16* __restrict__ gp) {
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
    if (gp) {            // act'(y) saved by the forward pass: no activation arithmetic here at all
      float gv[8];
      load8_bf16(gp + base + (long long)r * C, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dy = da_[j] * gv[j];
        s1[j] += dy;
        s2[j] = fmaf(dy, xa[j], s2[j]);
        da_[j] = dy;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dy = da_[j];
        if (act) dy *= act_grad_fast(fmaf(xa[j], sc[j], sh[j]), act);
        s1[j] += dy;
        s2[j] = fmaf(dy, xa[j], s2[j]);
        da_[j] = dy;
      }
    }
    if (dy_out) store8_bf16(dy_out + base + (long long)r * C, da_);   // consumed in place by the apply pass (FROM_DY)
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


// ---------------------------------------------------------------------------------------------------------------
// GroupNorm backward in ONE pass over HBM (round 2). The two-kernel version above reads x and da twice (row sums, then
// apply): 16 B/element for an fp32 x with a residual-branch gradient against 10 B for a single pass. Here one
// persistent kernel (one CTA per SM, all co-resident: cooperative launch) walks the batch in groups of a few samples
// that fit the 126 MB L2:
//   phase A units (128 rows of one sample) produce the per-channel partial sums exactly like gn_bwd_rowsum_fast_kernel;
//   the LAST unit of a sample to finish (atomic ticket) reduces that sample's partials in fixed order (bit-reproducible
//   whoever does it), publishes S1/S2 and the group means and raises the sample's flag;
//   phase B units of the same group then re-read x and da -- served by the L2, they were loaded microseconds ago -- and
//   write dx.
// Every CTA walks the same static unit list A(0), [A(1), B(0)], [A(2), B(1)], ..., B(last) strided by the grid size: a B
// unit only ever waits for A units that sit EARLIER in every CTA's list (no deadlock with co-resident CTAs, no grid-wide
// barrier), and by the time a CTA reaches B(g) that group's A units were handed out a whole group earlier.
// Data path: a producer warp streams 32 KB stages (8-32 rows of x, da and, in phase B, gres) into a 6-deep shared-memory
// ring with cp.async.bulk + mbarrier transaction counts; eight consumer warps compute from shared memory. A first
// version with register-staged global loads moved the right number of bytes (ncu: 5.39 GB instead of 8.67 GB per
// [256,64,64,512] call) but was latency-bound at one or two CTAs per SM (2.1 ms against 1.58 ms for the two kernels):
// only ~100 KB in flight per SM and a chain of dependent round trips (partials, fence, ticket, flag) per unit. The ring
// keeps ~190 KB in flight per SM whatever the consumers are doing. Second-read operands and the streams (gres, dx) carry
// evict-first hints so they do not push the group out of L2.
constexpr int GNR_CONSUMERS = 512;
constexpr int GNR_THREADS = GNR_CONSUMERS + 32;
constexpr int GNR_UNIT_ROWS = 128;
constexpr int GNR_STAGE_BYTES = 32 * 1024;
constexpr int GNR_STAGES = 6;
constexpr size_t GNR_SMEM = (size_t)GNR_STAGES * GNR_STAGE_BYTES + GNR_CONSUMERS * 16 * sizeof(float) + 256 + 128 /*align*/;
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull;   // createpolicy.fractional.L2::evict_first, 1.0
constexpr unsigned long long L2_EVICT_NORMAL = 0x1000000000000000ull;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                          unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(GNR_CONSUMERS) : "memory"); }
__device__ __forceinline__ void store8_bf16_stream(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
  o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  __stcs(reinterpret_cast<uint4*>(p), o);
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct GnUnit { int n, k, phase; };
// unit t of the software-pipelined static list (groups padded to S samples; n >= N means "skip")
__device__ __forceinline__ GnUnit gn_unit(long long t, int S, int Kc, int groups) {
  const long long UG = (long long)S * Kc;
  int g, phase;
  long long local;
  if (t < UG) { g = 0; phase = 0; local = t; }
  else {
    const long long tt = t - UG;
    const int slot = (int)(tt / (2 * UG)) + 1;
    const long long r = tt - (long long)(slot - 1) * 2 * UG;
    if (slot == groups) { g = groups - 1; phase = 1; local = r; }
    else if (r < UG) { g = slot; phase = 0; local = r; }
    else { g = slot - 1; phase = 1; local = r - UG; }
  }
  GnUnit u;
  u.n = g * S + (int)(local / Kc);
  u.k = (int)(local % Kc);
  u.phase = phase;
  return u;
}

template <typename TX>
__global__ void __launch_bounds__(GNR_THREADS, 1)
gn_bwd_ring_kernel(const TX* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const __nv_bfloat16* __restrict__ da,
                   const __nv_bfloat16* __restrict__ gres, int N, int HW, int C, int G, int act, int S, int rs,
                   float* __restrict__ part, float* __restrict__ ws, int* __restrict__ tickets, int* __restrict__ flags,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ cs_part) {
  extern __shared__ uint8_t gnr_raw[];
  const uint32_t raw_addr = smem_u32(gnr_raw);
  uint8_t* smem = gnr_raw + ((128u - (raw_addr & 127u)) & 127u);
  float* red = reinterpret_cast<float*>(smem + GNR_STAGES * GNR_STAGE_BYTES);            // [GNR_CONSUMERS][16]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GNR_STAGES * GNR_STAGE_BYTES + GNR_CONSUMERS * 16 * sizeof(float));
  uint64_t* empty_bar = full_bar + GNR_STAGES;
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GNR_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], GNR_CONSUMERS / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int Kc = HW / GNR_UNIT_ROWS;
  const int groups = (N + S - 1) / S;
  const long long total = 2ll * groups * S * Kc;
  const int nsub = GNR_UNIT_ROWS / rs;
  const uint32_t x_bytes = (uint32_t)rs * C * sizeof(TX), h_bytes = (uint32_t)rs * C * 2;

  if (warp == GNR_CONSUMERS / 32) {
    // ------------------------------------------------------------------ producer: one lane streams the ring
    if ((threadIdx.x & 31) == 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const GnUnit un = gn_unit(t, S, Kc, groups);
        if (un.n >= N) continue;
        const long long row0 = (long long)un.n * HW + (long long)un.k * GNR_UNIT_ROWS;
        const bool with_res = un.phase && gres != nullptr;
        const unsigned long long pol = un.phase ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
        for (int sub = 0; sub < nsub; ++sub) {
          mbar_wait(&empty_bar[stage], ph ^ 1, 11);
          uint8_t* st = smem + stage * GNR_STAGE_BYTES;
          const long long e0 = (row0 + (long long)sub * rs) * C;
          mbar_arrive_expect_tx(&full_bar[stage], x_bytes + h_bytes + (with_res ? h_bytes : 0u));
          bulk_load(st, x + e0, x_bytes, &full_bar[stage], pol);
          bulk_load(st + x_bytes, da + e0, h_bytes, &full_bar[stage], pol);
          if (with_res) bulk_load(st + x_bytes + h_bytes, gres + e0, h_bytes, &full_bar[stage], L2_EVICT_FIRST);
          if (++stage == GNR_STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers (8 warps)
  const int tid = threadIdx.x;
  const int U = C >> 3, lanes = GNR_CONSUMERS / U;
  const int u = tid % U, lane = tid / U;
  const int c = u << 3;
  const int gs = C / G;
  const int gidx = c / gs;
  const int rpt = rs / lanes;                 // rows per thread per stage
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  int stage = 0;
  uint32_t ph = 0;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const GnUnit un = gn_unit(t, S, Kc, groups);
    if (un.n >= N) continue;
    const int n = un.n, k = un.k;
    const int sg = n * G + gidx;
    const float mean = stats[2 * sg], rstd = stats[2 * sg + 1];
    float sc[8], sh[8];
    load8(gamma + c, sc);
    load8(beta + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] *= rstd;
      sh[j] = fmaf(-mean, sc[j], sh[j]);
    }
    const long long row0 = (long long)n * HW + (long long)k * GNR_UNIT_ROWS;
    if (!un.phase) {
      // ---------------------------------------------------------------- phase A: per-channel sums of dy and dy*x
      float s1[8], s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
      for (int sub = 0; sub < nsub; ++sub) {
        mbar_wait(&full_bar[stage], ph, 12);
        const uint8_t* st = smem + stage * GNR_STAGE_BYTES;
        const TX* xs = reinterpret_cast<const TX*>(st);
        const __nv_bfloat16* ds = reinterpret_cast<const __nv_bfloat16*>(st + x_bytes);
        for (int q = 0; q < rpt; ++q) {
          const int r = lane + q * lanes;
          float xv[8], dv[8];
          load8x(xs + (size_t)r * C + c, xv);
          load8_bf16(ds + (size_t)r * C + c, dv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float dy = dv[j];
            if (act) dy *= act_grad_fast(fmaf(xv[j], sc[j], sh[j]), act);
            s1[j] += dy;
            s2[j] = fmaf(dy, xv[j], s2[j]);
          }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == GNR_STAGES) { stage = 0; ph ^= 1; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s2[j] = (s2[j] - mean * s1[j]) * rstd;     // sum dy*x  ->  sum dy*xhat
      float* mine = red + (size_t)tid * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { mine[j] = s1[j]; mine[8 + j] = s2[j]; }
      consumer_bar();
      float* out = part + ((long long)n * Kc + k) * 2 * C;
      for (int q = tid; q < U * 16; q += GNR_CONSUMERS) {
        const int uu = q >> 4, kk = q & 15;
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += red[(size_t)(l * U + uu) * 16 + kk];   // fixed order
        __stcg(out + (kk >> 3) * C + (uu << 3) + (kk & 7), acc);
      }
      consumer_bar();
      if (tid == 0) {
        __threadfence();                                     // release the CTA's partials (cumulative over the barrier)
        s_last = (atomicAdd(tickets + n, 1) == Kc - 1);
      }
      consumer_bar();
      if (s_last) {
        // -------------------------------------------------------------- the sample's last unit: finalize (fixed order)
        __threadfence();
        const float* pn = part + (long long)n * Kc * 2 * C;
        float* smf = red;             // [2][C]
        for (int cc = tid; cc < 2 * C; cc += GNR_CONSUMERS) {
          float a = 0.f;
#pragma unroll 8
          for (int q = 0; q < Kc; ++q) a += __ldcg(pn + (long long)q * 2 * C + cc);
          const int ch = cc < C ? cc : cc - C;
          ws[((long long)(cc < C ? 0 : N) + n) * C + ch] = a;               // S1 (sum dy) | S2 (sum dy*xhat)
          smf[cc] = a * gamma[ch];
        }
        consumer_bar();
        for (int gg = tid; gg < G; gg += GNR_CONSUMERS) {
          float t1 = 0.f, t2 = 0.f;
          for (int j = 0; j < gs; ++j) { t1 += smf[gg * gs + j]; t2 += smf[C + gg * gs + j]; }
          const float inv = 1.0f / ((float)HW * (float)gs);
          float* gm_out = ws + 2ll * N * C + 2ll * (n * G + gg);
          __stcg(gm_out, t1 * inv);
          __stcg(gm_out + 1, t2 * inv);
        }
        consumer_bar();
        if (tid == 0) {
          __threadfence();
          st_release_gpu(flags + n, 1);
        }
      }
      consumer_bar();            // `red` is reused by the next unit
    } else {
      // ---------------------------------------------------------------- phase B: dx (x and da come from L2)
      if (tid == 0) {
        long long t0 = clock64();
        while (ld_acquire_gpu(flags + n) == 0) {
          __nanosleep(64);
          if (clock64() - t0 > TVAE_WAIT_TIMEOUT_CYCLES) {
            printf("tvae: gn_bwd_ring flag wait timeout sample=%d block=%d\n", n, (int)blockIdx.x);
            __trap();
          }
        }
      }
      consumer_bar();
      const float* gmeans = ws + 2ll * N * C;
      const float m1 = __ldcg(gmeans + 2 * sg), m2 = __ldcg(gmeans + 2 * sg + 1);
      const float ta = m2 * rstd * rstd;
      const float tb = m1 * rstd - mean * ta;
      for (int sub = 0; sub < nsub; ++sub) {
        mbar_wait(&full_bar[stage], ph, 13);
        const uint8_t* st = smem + stage * GNR_STAGE_BYTES;
        const TX* xs = reinterpret_cast<const TX*>(st);
        const __nv_bfloat16* ds = reinterpret_cast<const __nv_bfloat16*>(st + x_bytes);
        const __nv_bfloat16* gsm = reinterpret_cast<const __nv_bfloat16*>(st + x_bytes + h_bytes);
        for (int q = 0; q < rpt; ++q) {
          const int r = lane + q * lanes;
          float xv[8], dv[8], rv[8], o[8];
          load8x(xs + (size_t)r * C + c, xv);
          load8_bf16(ds + (size_t)r * C + c, dv);
          if (gres) {
            load8_bf16(gsm + (size_t)r * C + c, rv);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float dy = dv[j];
            if (act) dy *= act_grad_fast(fmaf(xv[j], sc[j], sh[j]), act);
            o[j] = fmaf(dy, sc[j], rv[j]) - fmaf(xv[j], ta, tb);
            cs[j] += o[j];
          }
          store8_bf16_stream(dx + (row0 + (long long)sub * rs + r) * C + c, o);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == GNR_STAGES) { stage = 0; ph ^= 1; }
      }
    }
  }
  if (cs_part) {       // column sums of dx over every unit this CTA applied (static schedule => reproducible)
    consumer_bar();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 8 + j] = cs[j];
    consumer_bar();
    float* out = cs_part + (long long)blockIdx.x * C;
    for (int q = tid; q < U * 8; q += GNR_CONSUMERS) {
      const int uu = q >> 3, kk = q & 7;
      float acc = 0.f;
      for (int l = 0; l < lanes; ++l) acc += red[(l * U + uu) * 8 + kk];   // fixed order
      out[(uu << 3) + kk] = acc;
    }
  }
}

constexpr int GNF_MAX_GRID = 1024;

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
                    int HW, int C, int G, int act, __nv_bfloat16* out, __nv_bfloat16* gp_out, cudaStream_t stream) {
  const int rpb = rows_per_block(HW);
  dim3 grid((HW + rpb - 1) / rpb, N);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  const float* xf = reinterpret_cast<const float*>(x);
  if (gp_out) {
    if (x_bf16) gn_act_fwd_fast_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(xb, stats, gamma, beta, HW, C, G, act, rpb, out, gp_out);
    else gn_act_fwd_fast_kernel<float, true><<<grid, 256, 0, stream>>>(xf, stats, gamma, beta, HW, C, G, act, rpb, out, gp_out);
  } else {
    if (x_bf16) gn_act_fwd_fast_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(xb, stats, gamma, beta, HW, C, G, act, rpb, out, nullptr);
    else gn_act_fwd_fast_kernel<float, false><<<grid, 256, 0, stream>>>(xf, stats, gamma, beta, HW, C, G, act, rpb, out, nullptr);
  }
  return 0;
}

long long gn_bwd_fast_ws_floats(int N, int HW, int C, int G) {
  const int rpb = rows_per_block(HW);
  const long long chunks = (HW + rpb - 1) / rpb;
  const long long two_pass = 2ll * N * C + 2ll * N * G + (long long)N * chunks * 2 * C + (long long)CS_SLICES * C;
  const long long kc = (HW + GNR_UNIT_ROWS - 1) / GNR_UNIT_ROWS;
  const long long fused = 2ll * N * C + 2ll * N * G + (long long)N * kc * 2 * C + (long long)(GNF_MAX_GRID + CS_SLICES) * C +
                          2ll * N + 64;
  return two_pass > fused ? two_pass : fused;
}

// 0: always the two-pass kernels; 1: the single-pass persistent kernel for tensors that do not fit the L2; 2: the single
// pass whatever the size (tests). TVAE_GN_BWD_FUSED / tvae_gn_set_bwd_fused. group_mb: megabytes of (x + da) per
// L2-resident group.
static int g_gn_fused = -1;
static int g_gn_group_mb = 24;
static int g_gn_store_dy = 1;       // TVAE_GN_STORE_DY=0: the apply pass recomputes act' from da (round-1 behaviour)
static void gn_read_env() {
  if (g_gn_fused >= 0) return;
  const char* d = getenv("TVAE_GN_STORE_DY");
  if (d && d[0] == '0') g_gn_store_dy = 0;
  // Default OFF. Measured on B200, [256,64,64,512] (profiles/gn_bwd_single_pass_r2.md): the single pass moves 5.39 GB
  // instead of 8.67 GB (ncu dram bytes -- the L2 hand-over works) but takes 2.05 ms against 1.59 ms for the two kernels:
  // it is bound by instruction issue, not by bytes (same time for 5.4 GB fp32+gres and 3.2 GB bf16 inputs; 2.73 ms with
  // 8 consumer warps, 2.05 ms with 16), because GELU' is evaluated in both phases (~50 instructions per element) and one
  // CTA per SM leaves 16 warps to hide the MUFU / shared-memory latencies where the two-pass kernels have 64.
  const char* e = getenv("TVAE_GN_BWD_FUSED");
  g_gn_fused = (e && e[0] == '1') ? 1 : 0;
  const char* m = getenv("TVAE_GN_GROUP_MB");
  if (m && atoi(m) > 0) g_gn_group_mb = atoi(m);
}
void gn_set_bwd_fused(int on, int group_mb) {
  gn_read_env();
  g_gn_fused = on == 2 ? 2 : (on ? 1 : 0);
  if (group_mb > 0) g_gn_group_mb = group_mb;
}

// rows per ring stage: the largest power of two whose x + da + gres rows fit GNR_STAGE_BYTES (at most 32)
static int gn_ring_rows(int C, int xbytes) {
  int rs = 32;
  while (rs > 1 && (long long)rs * C * (xbytes + 4) > GNR_STAGE_BYTES) rs >>= 1;
  return rs;
}
static bool gn_ring_ok(int HW, int C, int xbytes) {
  if (C % 8 || GNR_CONSUMERS % (C / 8) || 2 * C > GNR_CONSUMERS * 16 || HW % GNR_UNIT_ROWS) return false;
  const int rs = gn_ring_rows(C, xbytes), lanes = GNR_CONSUMERS / (C / 8);
  return rs >= lanes && rs % lanes == 0 && GNR_UNIT_ROWS % rs == 0 && (long long)rs * C * (xbytes + 4) <= GNR_STAGE_BYTES;
}

template <typename TX>
static int launch_gn_bwd_ring(const TX* x, const float* stats, const float* gamma, const float* beta,
                              const __nv_bfloat16* da, const __nv_bfloat16* gres, int N, int HW, int C, int G, int act,
                              __nv_bfloat16* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws,
                              cudaStream_t stream) {
  static PerDeviceOnce attr_set;
  if (attr_set.pending()) {
    if (cudaFuncSetAttribute(gn_bwd_ring_kernel<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GNR_SMEM) != cudaSuccess)
      return -1;
    attr_set.mark();
  }
  const int Kc = HW / GNR_UNIT_ROWS;
  const long long per_sample = (long long)HW * C * (sizeof(TX) + 2);
  long long S = ((long long)g_gn_group_mb << 20) / per_sample;
  if (S < 1) S = 1;
  if (S > N) S = N;
  int grid = num_sms();                       // one CTA per SM, all co-resident (cooperative launch)
  const long long units = 2ll * N * Kc;
  if (grid > units) grid = (int)units;
  if (grid > GNF_MAX_GRID) grid = GNF_MAX_GRID;
  float* part = ws + 2ll * N * C + 2ll * N * G;
  float* cs_part = part + (long long)N * Kc * 2 * C;
  float* slices = cs_part + (long long)GNF_MAX_GRID * C;
  int* tickets = reinterpret_cast<int*>(slices + (long long)CS_SLICES * C);
  int* flags = tickets + N;
  if (cudaMemsetAsync(tickets, 0, 2ull * N * sizeof(int), stream) != cudaSuccess) return -1;
  int Si = (int)S;
  int rs = gn_ring_rows(C, (int)sizeof(TX));
  float* cs_arg = dx_colsum ? cs_part : nullptr;
  void* args[] = {(void*)&x, (void*)&stats, (void*)&gamma, (void*)&beta, (void*)&da, (void*)&gres, (void*)&N, (void*)&HW,
                  (void*)&C, (void*)&G, (void*)&act, (void*)&Si, (void*)&rs, (void*)&part, (void*)&ws, (void*)&tickets,
                  (void*)&flags, (void*)&dx, (void*)&cs_arg};
  if (cudaLaunchCooperativeKernel((const void*)gn_bwd_ring_kernel<TX>, dim3(grid), dim3(GNR_THREADS), args, GNR_SMEM,
                                  stream) != cudaSuccess)
    return -1;
  gn_bwd_param_fast_kernel<<<(C + 31) / 32, 256, 0, stream>>>(ws, N, C, dgamma, dbeta);
  if (dx_colsum) {
    colsum_rows_slice_kernel<<<dim3((C + 31) / 32, CS_SLICES), 256, 0, stream>>>(cs_part, grid, C, slices);
    colsum_rows_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(slices, C, dx_colsum);
  }
  return 0;
}

int gn_act_bwd_fast(const void* xv, bool x_bf16, const float* stats, const float* gamma, const float* beta,
                    const __nv_bfloat16* da, const __nv_bfloat16* gres, const __nv_bfloat16* gp, int N, int HW, int C,
                    int G, int act, __nv_bfloat16* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws,
                    cudaStream_t stream) {
  const int rpb = rows_per_block(HW);
  const int chunks = (HW + rpb - 1) / rpb;
  dim3 grid(chunks, N);
  float* part = ws + 2ll * N * C + 2ll * N * G;
  const float* x = reinterpret_cast<const float*>(xv);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv);
  gn_read_env();
  // Tensors that fit the L2 anyway get their second pass from it with the two-kernel version; the persistent kernel
  // is for the ones that do not (>= 96 MB of x + da)
  const long long bytes = (long long)N * HW * C * ((x_bf16 ? 2 : 4) + 2);
  if (((g_gn_fused == 1 && bytes >= (96ll << 20)) || g_gn_fused == 2) && gn_ring_ok(HW, C, x_bf16 ? 2 : 4)) {
    const int rc = x_bf16 ? launch_gn_bwd_ring(xb, stats, gamma, beta, da, gres, N, HW, C, G, act, dx, dgamma, dbeta,
                                               dx_colsum, ws, stream)
                          : launch_gn_bwd_ring(x, stats, gamma, beta, da, gres, N, HW, C, G, act, dx, dgamma, dbeta,
                                               dx_colsum, ws, stream);
    if (rc == 0) return 0;
    set_error("gn_bwd_ring: cooperative launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  // dy hand-over (g_gn_store_dy, default on): with an activation the row-sum pass writes dy = da * act'(y) into dx and the
  // apply pass starts from it. +2 B/element of traffic, -16 instructions/element: inside the train step these kernels run
  // at the power-capped SM clock (1.3-1.5 GHz) and are issue-bound, not bandwidth-bound.
  const bool hand_over = g_gn_store_dy != 0 && act != 0;
  __nv_bfloat16* dy_out = hand_over ? dx : nullptr;
  // act'(y) saved by the forward pass (tvae_gn_act_fwd2) is only usable together with the hand-over: the apply pass
  // must not need act' either
  const __nv_bfloat16* gp_in = hand_over ? gp : nullptr;
  if (x_bf16)
    gn_bwd_rowsum_fast_kernel<<<grid, 256, 256 * 16 * sizeof(float), stream>>>(xb, stats, gamma, beta, da, HW, C, G,
                                                                              act, rpb, part, dy_out, gp_in);
  else
    gn_bwd_rowsum_fast_kernel<<<grid, 256, 256 * 16 * sizeof(float), stream>>>(x, stats, gamma, beta, da, HW, C, G,
                                                                              act, rpb, part, dy_out, gp_in);
  gn_bwd_finalize_fast_kernel<<<N, 256, 2 * C * sizeof(float), stream>>>(part, gamma, chunks, HW, C, G, N, ws, nullptr);
  gn_bwd_param_fast_kernel<<<(C + 31) / 32, 256, 0, stream>>>(ws, N, C, dgamma, dbeta);
  // the row-sum partials are consumed by now (stream order): their region is reused for the column sums of dx
  float* csp = dx_colsum ? part : nullptr;
  const float* gm = ws + 2ll * N * C;
  if (x_bf16 && hand_over)
    gn_bwd_apply_fast_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(xb, stats, gamma, beta, da, gres, gm, HW, C, G, act, rpb, dx, csp);
  else if (x_bf16)
    gn_bwd_apply_fast_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(xb, stats, gamma, beta, da, gres, gm, HW, C, G, act, rpb, dx, csp);
  else if (hand_over)
    gn_bwd_apply_fast_kernel<float, true><<<grid, 256, 0, stream>>>(x, stats, gamma, beta, da, gres, gm, HW, C, G, act, rpb, dx, csp);
  else
    gn_bwd_apply_fast_kernel<float, false><<<grid, 256, 0, stream>>>(x, stats, gamma, beta, da, gres, gm, HW, C, G, act, rpb, dx, csp);
  if (dx_colsum) {
    float* slices = part + (long long)N * chunks * 2 * C;
    colsum_rows_slice_kernel<<<dim3((C + 31) / 32, CS_SLICES), 256, 0, stream>>>(part, N * chunks, C, slices);
    colsum_rows_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(slices, C, dx_colsum);
  }
  return 0;
}

}  // namespace tvae
