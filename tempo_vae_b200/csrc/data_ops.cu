// Data-side HBM-bound kernels: device-side tile gather, fused
// crop + flip + rot90 + log/z-score/clip tile extraction straight from a raw radiance granule, per-channel spectrum
// statistics, and the batch statistics the trainer prints at step 0. Streaming kernels: coalesced along the channel
// dimension, fixed-order two-stage reductions (bit-reproducible).
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------ gather rows
// dst[j] = src[idx[j]] for rows of row_bytes (multiple of 16) bytes: blockIdx.y = output row, 16-byte accesses.
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx,
                                                          long long row_vec, long long n_src, uint4* __restrict__ dst) {
  const long long j = blockIdx.y;
  long long r = idx[j];
  if (r < 0 || r >= n_src) r = 0;                              // defensive: host validates the range
  const uint4* s = src + r * row_vec;
  uint4* d = dst + j * row_vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread
  for (; i + 3 * stride < row_vec; i += 4 * stride) {
    const uint4 a = __ldg(s + i), b = __ldg(s + i + stride), c = __ldg(s + i + 2 * stride), e = __ldg(s + i + 3 * stride);
    d[i] = a; d[i + stride] = b; d[i + 2 * stride] = c; d[i + 3 * stride] = e;
  }
  for (; i < row_vec; i += stride) d[i] = __ldg(s + i);
}

// ------------------------------------------------------------------------------------------------ tile extraction
// One output tile pixel row per (block.y); threads run along channel pairs. spec[t] = {row0, col0, flags, k}:
// flags bit 0 = flip along the mirror axis (torch.flip(tile, dims=[0])), bit 1 = flip along the track axis (dims=[1]),
// k = number of 90-degree rotations (torch.rot90(tile, k, dims=[0, 1])), applied in that order like
// src/scripts/prepare_tempo_tiles.py:36-52. The kernel walks the OUTPUT and inverts the chain to find its source
// pixel: rot90 by k, then the flips.
__device__ __forceinline__ void source_pixel(int oi, int oj, int T, int flags, int k, int* si, int* sj) {
  int i, j;                                   // position in the tile before the rotation
  switch (k & 3) {
    case 1: i = oj; j = T - 1 - oi; break;            // out[i,j] = x[j, T-1-i]
    case 2: i = T - 1 - oi; j = T - 1 - oj; break;    // out[i,j] = x[T-1-i, T-1-j]
    case 3: i = T - 1 - oj; j = oi; break;            // out[i,j] = x[T-1-j, i]
    default: i = oi; j = oj; break;
  }
  if (flags & 2) j = T - 1 - j;               // second flip (dims=[1]) undone first
  if (flags & 1) i = T - 1 - i;
  *si = i; *sj = j;
}

__global__ void __launch_bounds__(256)
extract_tiles_kernel(const float* __restrict__ rad, int M, int NT, int C, const int* __restrict__ spec, int T,
                     const float* __restrict__ mean, const float* __restrict__ stdv, float min_rad, float lo, float hi,
                     int normalize, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int out_pitch) {
  const int t = blockIdx.y;
  const int4 sp = *reinterpret_cast<const int4*>(spec + 4 * t);
  const int U = (out_bf16 ? out_pitch : C + (C & 1)) >> 1;         // channel pairs per output row (bf16 pad lanes incl.)
  const long long total = (long long)T * T * U;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long spx = stride / U;
  const int su = (int)(stride - spx * U);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long px = i / U;
  int u = (int)(i - px * U);
  for (; i < total; i += stride, px += spx, u += su) {
    if (u >= U) { u -= U; ++px; }
    const int oi = (int)(px / T), oj = (int)(px - (long long)oi * T);
    int si, sj;
    source_pixel(oi, oj, T, sp.z, sp.w, &si, &sj);
    const float* src = rad + ((long long)(sp.x + si) * NT + (sp.y + sj)) * C;
    const long long orow = (long long)t * T * T + px;
    const int c = u << 1;
    float z[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (c + q < C) {
        float v = __ldg(src + c + q);
        if (normalize) {
          v = logf(fmaxf(v, min_rad));
          v = fminf(fmaxf((v - mean[c + q]) / (stdv[c + q] + 1e-8f), lo), hi);
        }
        z[q] = v;
        if (out_f32) out_f32[orow * C + c + q] = v;
      }
    }
    if (out_bf16) *reinterpret_cast<uint32_t*>(out_bf16 + orow * out_pitch + c) = pack2_bf16(z[0], z[1]);
  }
}

// ------------------------------------------------------------------------------------------------ spectrum statistics
// Per-channel sum and sum of squares of v = log(max(rad, min_rad)) over `rows` pixels (np.log(np.clip(...)) of
// src/scripts/compute_tempo_stats.py:68-71). Block b owns a contiguous pixel range; thread = channel (coalesced rows).
// fp32 partial sums over 32 pixels are folded into fp64 accumulators; block partials go to the workspace and are
// reduced in block order by the accumulate kernel.
constexpr int STATS_ROWS_PER_BLOCK = 512;
__global__ void __launch_bounds__(256)
spectrum_partial_kernel(const float* __restrict__ rad, long long rows, int C, float min_rad, int take_log,
                        double* __restrict__ part /* [blocks][2][C] */) {
  const long long r0 = (long long)blockIdx.x * STATS_ROWS_PER_BLOCK;
  const long long r1 = r0 + STATS_ROWS_PER_BLOCK < rows ? r0 + STATS_ROWS_PER_BLOCK : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (long long rb = r0; rb < r1; rb += 32) {
      float fs = 0.f, fq = 0.f;
      const long long re = rb + 32 < r1 ? rb + 32 : r1;
#pragma unroll 8
      for (long long r = rb; r < re; ++r) {
        float v = __ldg(rad + r * C + c);
        if (take_log) v = logf(fmaxf(v, min_rad));
        fs += v;
        fq = fmaf(v, v, fq);
      }
      s += (double)fs;
      q += (double)fq;
    }
    part[((long long)blockIdx.x * 2 + 0) * C + c] = s;
    part[((long long)blockIdx.x * 2 + 1) * C + c] = q;
  }
}
__global__ void spectrum_accum_kernel(const double* __restrict__ part, int blocks, int C, double* __restrict__ acc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < blocks; ++b) {
    s += part[((long long)b * 2 + 0) * C + c];
    q += part[((long long)b * 2 + 1) * C + c];
  }
  acc[c] += s;
  acc[C + c] += q;
}
__global__ void spectrum_finalize_kernel(const double* __restrict__ acc, double n, int C, float* __restrict__ mean,
                                         float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[c] / n;
  double var = acc[C + c] / n - m * m;          // population variance (np.std, ddof = 0)
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  stdv[c] = (float)sqrt(var);
}

// ------------------------------------------------------------------------------------------------ batch statistics
// min / max / sum / sum of squares over the C valid channels of `rows` channels-contiguous rows (fp32 or bf16, any row
// pitch) -- or over a flat fp32 array when rows = 1. Two stages, fixed order.
template <typename T>
__device__ __forceinline__ float load_as_f32(const T* p);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
batch_stats_partial_kernel(const T* __restrict__ x, long long rows, long long C, long long pitch,
                           double* __restrict__ part /* [blocks][4] */) {
  float mn = INFINITY, mx = -INFINITY;
  double s = 0.0, q = 0.0;
  const long long total = rows * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / C, c = i - r * C;
    const float v = load_as_f32<T>(x + r * pitch + c);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    s += (double)v;
    q += (double)v * (double)v;
  }
  __shared__ double sh[4][256];
  sh[0][threadIdx.x] = (double)mn; sh[1][threadIdx.x] = (double)mx; sh[2][threadIdx.x] = s; sh[3][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] = fmin(sh[0][threadIdx.x], sh[0][threadIdx.x + o]);
      sh[1][threadIdx.x] = fmax(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
      sh[3][threadIdx.x] += sh[3][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) part[blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void batch_stats_final_kernel(const double* __restrict__ part, int blocks, double n, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mn = INFINITY, mx = -INFINITY, s = 0.0, q = 0.0;
  for (int b = 0; b < blocks; ++b) {
    mn = fmin(mn, part[b * 4 + 0]); mx = fmax(mx, part[b * 4 + 1]); s += part[b * 4 + 2]; q += part[b * 4 + 3];
  }
  const double mean = s / n;
  double var = n > 1.0 ? (q - n * mean * mean) / (n - 1.0) : 0.0;        // unbiased, like Tensor.std()
  if (var < 0.0) var = 0.0;
  out[0] = (float)mn; out[1] = (float)mx; out[2] = (float)mean; out[3] = (float)sqrt(var);
}

constexpr int BSTAT_BLOCKS = 148 * 8;

}  // namespace
}  // namespace tvae

using namespace tvae;

extern "C" int32_t tvae_gather_rows(const void* src, int64_t n_src, int64_t row_bytes, const int64_t* idx, int32_t n,
                                    void* dst, cudaStream_t stream) {
  TVAE_ENTER(src);
  TVAE_CHECK(src && idx && dst, "tvae_gather_rows: null pointer");
  TVAE_CHECK(row_bytes > 0 && row_bytes % 16 == 0, "tvae_gather_rows: row_bytes must be a positive multiple of 16");
  TVAE_CHECK((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0,
             "tvae_gather_rows: src and dst must be 16-byte aligned");
  if (n <= 0) return 0;
  TVAE_CHECK(n <= 65535, "tvae_gather_rows: at most 65535 rows per call");
  const long long row_vec = row_bytes / 16;
  long long bx = (row_vec + 256 * 4 - 1) / (256 * 4);
  const long long cap = (148LL * 8 + n - 1) / n;                // ~8 resident CTAs per SM over the whole grid
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  gather_rows_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<const long long*>(idx), row_vec, n_src,
      reinterpret_cast<uint4*>(dst));
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_extract_tiles(const float* rad, int32_t M, int32_t NT, int32_t C, const int32_t* spec,
                                      int32_t n_tiles, int32_t T, const float* mean, const float* stdv,
                                      float min_radiance, float clip_min, float clip_max, float* out_f32,
                                      void* out_bf16, int32_t out_pitch, cudaStream_t stream) {
  TVAE_ENTER(rad);
  TVAE_CHECK(rad && spec && (out_f32 || out_bf16), "tvae_extract_tiles: null pointer");
  TVAE_CHECK((mean == nullptr) == (stdv == nullptr), "tvae_extract_tiles: mean and std go together");
  TVAE_CHECK(T > 0 && T <= M && T <= NT, "tvae_extract_tiles: tile %d does not fit a %d x %d granule", T, M, NT);
  TVAE_CHECK(!out_bf16 || (out_pitch >= C && out_pitch % 2 == 0), "tvae_extract_tiles: bad out_pitch");
  if (n_tiles <= 0) return 0;
  TVAE_CHECK(n_tiles <= 65535, "tvae_extract_tiles: at most 65535 tiles per call");
  const long long U = ((out_bf16 ? out_pitch : C + 1) / 2);
  const long long work = (long long)T * T * U;
  long long bx = (work + 255) / 256;
  const long long cap = (148LL * 16 + n_tiles - 1) / n_tiles;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  extract_tiles_kernel<<<dim3((unsigned)bx, (unsigned)n_tiles), 256, 0, stream>>>(
      rad, M, NT, C, spec, T, mean, stdv, min_radiance, clip_min, clip_max, mean != nullptr ? 1 : 0, out_f32,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), out_pitch);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_spectrum_stats_workspace_bytes(int64_t rows, int32_t C) {
  const long long blocks = (rows + STATS_ROWS_PER_BLOCK - 1) / STATS_ROWS_PER_BLOCK;
  return blocks * 2 * (long long)C * (long long)sizeof(double);
}

extern "C" int32_t tvae_spectrum_stats_accum(const float* rad, int64_t rows, int32_t C, float min_radiance,
                                             int32_t take_log, double* acc, double* workspace, cudaStream_t stream) {
  TVAE_ENTER(rad);
  TVAE_CHECK(rad && acc && workspace, "tvae_spectrum_stats_accum: null pointer");
  if (rows <= 0) return 0;
  const long long blocks = (rows + STATS_ROWS_PER_BLOCK - 1) / STATS_ROWS_PER_BLOCK;
  TVAE_CHECK(blocks <= 0x7fffffffLL, "tvae_spectrum_stats_accum: too many rows for one call");
  spectrum_partial_kernel<<<(unsigned)blocks, 256, 0, stream>>>(rad, rows, C, min_radiance, take_log, workspace);
  TVAE_CUDA(cudaGetLastError());
  spectrum_accum_kernel<<<(C + 127) / 128, 128, 0, stream>>>(workspace, (int)blocks, C, acc);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_spectrum_stats_finalize(const double* acc, int64_t total_rows, int32_t C, float* mean,
                                                float* stdv, cudaStream_t stream) {
  TVAE_ENTER(acc);
  TVAE_CHECK(acc && mean && stdv && total_rows > 0, "tvae_spectrum_stats_finalize: bad arguments");
  spectrum_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(acc, (double)total_rows, C, mean, stdv);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t tvae_batch_stats_workspace_bytes(void) { return BSTAT_BLOCKS * 4 * (int64_t)sizeof(double); }

extern "C" int32_t tvae_batch_stats(const void* x, int32_t is_bf16, int64_t rows, int64_t C, int64_t pitch, float* out,
                                    double* workspace, cudaStream_t stream) {
  TVAE_ENTER(x);
  TVAE_CHECK(x && out && workspace && rows > 0 && C > 0 && pitch >= C, "tvae_batch_stats: bad arguments");
  const long long total = rows * C;
  long long blocks = (total + 255) / 256;
  if (blocks > BSTAT_BLOCKS) blocks = BSTAT_BLOCKS;
  if (is_bf16)
    batch_stats_partial_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), rows, C, pitch, workspace);
  else
    batch_stats_partial_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(x), rows, C,
                                                                          pitch, workspace);
  TVAE_CUDA(cudaGetLastError());
  batch_stats_final_kernel<<<1, 32, 0, stream>>>(workspace, (int)blocks, (double)total, out);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
