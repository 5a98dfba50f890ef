// Mid-block self-attention core (softmax(Q^T K / sqrt(d)) V) for the 16x16-token bottleneck.
// Replaces the two einsums + softmax of AttnBlock.forward (src/model.py:128-139) and their backward.
//
// The work is tiny (0.067 GFLOP / sample forward; Appendix B of SURVEY.md) and the head layout is
// channel-interleaved (head h owns channels {d*heads + h}), so this is an fp32 SIMT flash-style kernel:
// one thread per (token, head), K/V (or Q/dO) row tiles staged through shared memory with fully coalesced
// row loads, online softmax, nothing of size T x T ever written to memory. It works for any T (whole-granule
// inference runs 16,384 tokens through the same code).
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int ROWS = 64;  // tokens per block (x heads threads)

// ---------------------------------------------------------------------------------------------- forward
template <int HD>
__global__ void attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                int pitch, int T, int heads, int KT, float scale, __nv_bfloat16* __restrict__ out_bf16,
                                float* __restrict__ out_f32, float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int C = HD * heads;
  float* sK = sm;                  // [KT][C]
  float* sV = sm + (size_t)KT * C; // [KT][C]
  const int b = blockIdx.y;
  const int h = threadIdx.x % heads;
  const int tq = blockIdx.x * ROWS + threadIdx.x / heads;
  const bool active = tq < T;
  const long long rowq = (long long)b * T + (active ? tq : 0);

  float qr[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    qr[d] = q[rowq * pitch + d * heads + h] * scale;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < T; k0 += KT) {
    const int kt = min(KT, T - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * C; i += blockDim.x) {
      const int r = i / C, c = i - r * C;
      const long long row = (long long)b * T + k0 + r;
      sK[i] = k[row * pitch + c];
      sV[i] = v[row * pitch + c];
    }
    __syncthreads();
    for (int r = 0; r < kt; ++r) {
      const float* kr = sK + r * C + h;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s += qr[d] * kr[d * heads];
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn);
      const float p = __expf(s - mn);
      l = l * corr + p;
      const float* vr = sV + r * C + h;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = o[d] * corr + p * vr[d * heads];
      m = mn;
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    const long long row = (long long)b * T + tq;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      const float val = o[d] * inv;
      if (out_bf16) out_bf16[row * C + d * heads + h] = __float2bfloat16(val);
      if (out_f32) out_f32[row * C + d * heads + h] = val;
    }
    if (lse) lse[((long long)b * heads + h) * T + tq] = m + __logf(l);
  }
}

// ---------------------------------------------------------------------------------------------- backward: dQ
template <int HD>
__global__ void attn_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                   const float* __restrict__ v, int pitch, const float* __restrict__ o,
                                   const float* __restrict__ dout, const float* __restrict__ lse, int T, int heads,
                                   int KT, float scale, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dsum) {
  extern __shared__ float sm[];
  const int C = HD * heads;
  float* sK = sm;
  float* sV = sm + (size_t)KT * C;
  const int b = blockIdx.y;
  const int h = threadIdx.x % heads;
  const int tq = blockIdx.x * ROWS + threadIdx.x / heads;
  const bool active = tq < T;
  const long long rowq = (long long)b * T + (active ? tq : 0);

  float qr[HD], dor[HD], dq[HD];
  float D = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    qr[d] = q[rowq * pitch + d * heads + h] * scale;
    dor[d] = dout[rowq * C + d * heads + h];
    D += dor[d] * o[rowq * C + d * heads + h];
    dq[d] = 0.f;
  }
  const float L = lse[((long long)b * heads + h) * T + (active ? tq : 0)];
  if (active) dsum[((long long)b * heads + h) * T + tq] = D;

  for (int k0 = 0; k0 < T; k0 += KT) {
    const int kt = min(KT, T - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * C; i += blockDim.x) {
      const int r = i / C, c = i - r * C;
      const long long row = (long long)b * T + k0 + r;
      sK[i] = k[row * pitch + c];
      sV[i] = v[row * pitch + c];
    }
    __syncthreads();
    for (int r = 0; r < kt; ++r) {
      const float* kr = sK + r * C + h;
      const float* vr = sV + r * C + h;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s += qr[d] * kr[d * heads];
        dp += dor[d] * vr[d * heads];
      }
      const float p = __expf(s - L);
      const float ds = p * (dp - D);
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] += ds * kr[d * heads];
    }
  }
  if (active) {
    const long long row = (long long)b * T + tq;
#pragma unroll
    for (int d = 0; d < HD; ++d) dqkv[row * 3 * C + d * heads + h] = __float2bfloat16(dq[d] * scale);
  }
}

// ---------------------------------------------------------------------------------------------- backward: dK, dV
template <int HD>
__global__ void attn_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                    const float* __restrict__ v, int pitch, const float* __restrict__ dout,
                                    const float* __restrict__ lse, const float* __restrict__ dsum, int T, int heads,
                                    int QT, float scale, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ float sm[];
  const int C = HD * heads;
  float* sQ = sm;                        // [QT][C]
  float* sDO = sm + (size_t)QT * C;      // [QT][C]
  float* sL = sDO + (size_t)QT * C;      // [QT][heads]
  float* sD = sL + (size_t)QT * heads;   // [QT][heads]
  const int b = blockIdx.y;
  const int h = threadIdx.x % heads;
  const int tk = blockIdx.x * ROWS + threadIdx.x / heads;
  const bool active = tk < T;
  const long long rowk = (long long)b * T + (active ? tk : 0);

  float kr[HD], vr[HD], dk[HD], dv[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    kr[d] = k[rowk * pitch + d * heads + h];
    vr[d] = v[rowk * pitch + d * heads + h];
    dk[d] = 0.f;
    dv[d] = 0.f;
  }
  for (int q0 = 0; q0 < T; q0 += QT) {
    const int qt = min(QT, T - q0);
    __syncthreads();
    for (int i = threadIdx.x; i < qt * C; i += blockDim.x) {
      const int r = i / C, c = i - r * C;
      const long long row = (long long)b * T + q0 + r;
      sQ[i] = q[row * pitch + c];
      sDO[i] = dout[row * C + c];
    }
    for (int i = threadIdx.x; i < qt * heads; i += blockDim.x) {
      const int r = i / heads, hh = i - r * heads;
      sL[i] = lse[((long long)b * heads + hh) * T + q0 + r];
      sD[i] = dsum[((long long)b * heads + hh) * T + q0 + r];
    }
    __syncthreads();
    for (int r = 0; r < qt; ++r) {
      const float* qr = sQ + r * C + h;
      const float* dor = sDO + r * C + h;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s += qr[d * heads] * kr[d];
        dp += dor[d * heads] * vr[d];
      }
      const float p = __expf(s * scale - sL[r * heads + h]);
      const float ds = p * (dp - sD[r * heads + h]);
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        dv[d] += p * dor[d * heads];
        dk[d] += ds * qr[d * heads];
      }
    }
  }
  if (active) {
    const long long row = (long long)b * T + tk;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      dqkv[row * 3 * C + C + d * heads + h] = __float2bfloat16(dk[d] * scale);
      dqkv[row * 3 * C + 2 * C + d * heads + h] = __float2bfloat16(dv[d]);
    }
  }
}

int tile_rows(int C, int arrays) {
  // rows of C floats per staged array so that `arrays` of them fit in ~64 KB
  int r = (64 * 1024) / (arrays * C * 4);
  int t = 1;
  while (t * 2 <= r && t < 64) t *= 2;
  return t;
}

}  // namespace
}  // namespace tvae

using namespace tvae;

#define TVAE_HD_DISPATCH(HDV, ...)                                       \
  switch (HDV) {                                                         \
    case 4: { constexpr int HD = 4; __VA_ARGS__; break; }                \
    case 8: { constexpr int HD = 8; __VA_ARGS__; break; }                \
    case 16: { constexpr int HD = 16; __VA_ARGS__; break; }              \
    case 32: { constexpr int HD = 32; __VA_ARGS__; break; }              \
    case 64: { constexpr int HD = 64; __VA_ARGS__; break; }              \
    default: tvae::set_error("attention: unsupported head dim %d (supported: 4, 8, 16, 32, 64)", HDV); return -1; \
  }

extern "C" int32_t tvae_attn_fwd(const float* q, const float* k, const float* v, int32_t pitch, int32_t B, int32_t T,
                                 int32_t C, int32_t heads, void* out_bf16, float* out_f32, float* lse,
                                 cudaStream_t stream) {
  TVAE_ENTER(q);
  TVAE_CHECK(q && k && v && (out_bf16 || out_f32), "tvae_attn_fwd: null pointer");
  TVAE_CHECK(heads > 0 && C % heads == 0 && ROWS * heads <= 1024, "tvae_attn_fwd: bad heads");
  const int hd = C / heads;
  const float scale = 1.0f / sqrtf((float)hd);
  const int KT = tile_rows(C, 2);
  const size_t smem = (size_t)2 * KT * C * sizeof(float);
  dim3 grid((T + ROWS - 1) / ROWS, B);
  const int threads = ROWS * heads;
  TVAE_HD_DISPATCH(hd, {
    TVAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<HD><<<grid, threads, smem, stream>>>(q, k, v, pitch, T, heads, KT, scale,
                                                        reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, lse);
  });
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_attn_bwd(const float* q, const float* k, const float* v, int32_t pitch, const float* o,
                                 const float* d_out, const float* lse, int32_t B, int32_t T, int32_t C, int32_t heads,
                                 void* dqkv_bf16, float* workspace, cudaStream_t stream) {
  TVAE_ENTER(q);
  TVAE_CHECK(q && k && v && o && d_out && lse && dqkv_bf16 && workspace, "tvae_attn_bwd: null pointer");
  TVAE_CHECK(heads > 0 && C % heads == 0 && ROWS * heads <= 1024, "tvae_attn_bwd: bad heads");
  const int hd = C / heads;
  const float scale = 1.0f / sqrtf((float)hd);
  dim3 grid((T + ROWS - 1) / ROWS, B);
  const int threads = ROWS * heads;
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
  {
    const int KT = tile_rows(C, 2);
    const size_t smem = (size_t)2 * KT * C * sizeof(float);
    TVAE_HD_DISPATCH(hd, {
      TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attn_bwd_dq_kernel<HD><<<grid, threads, smem, stream>>>(q, k, v, pitch, o, d_out, lse, T, heads, KT, scale, dp,
                                                             workspace);
    });
    TVAE_CUDA(cudaGetLastError());
  }
  {
    const int QT = tile_rows(C, 2);
    const size_t smem = ((size_t)2 * QT * C + (size_t)2 * QT * heads) * sizeof(float);
    TVAE_HD_DISPATCH(hd, {
      TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attn_bwd_dkv_kernel<HD><<<grid, threads, smem, stream>>>(q, k, v, pitch, d_out, lse, workspace, T, heads, QT,
                                                              scale, dp);
    });
    TVAE_CUDA(cudaGetLastError());
  }
  return 0;
}
