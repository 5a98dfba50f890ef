// Tensor-core version of the mid-block self-attention core for head dimension 32 (the shipped configuration:
// 128 channels, 4 channel-interleaved heads; src/model.py:128-139): flash-style, nothing T x T is materialised.
//
// mma.sync.m16n8k8 TF32 (fp32 accumulate): the work is ~0.2 GFLOP/sample per train step — far too small and too
// irregular (T = 256 tokens, d = 32) for a tcgen05/TMEM pipeline to pay off, but large enough that the fp32 SIMT
// kernels of attention.cu cost 6 ms/step at B = 256 (they are bound by the shared-memory -> register data path).
// TF32 keeps 10 mantissa bits on Q, K, V, dO, P (relative error ~5e-4 on the logits); the exact SIMT kernels
// remain for other head sizes and for the fp32 mode.
//
// One warp owns 16 query rows (forward, dQ) or 16 key rows (dK/dV). The row-side operand lives in registers as A
// fragments; the column-side operand is staged in shared memory as [row][36 floats] (pitch 36 => both B-fragment
// access patterns below are bank-conflict free). Scores are produced 64 columns at a time (8 MMA n-tiles), turned
// into probabilities in registers, and fed straight back as the A operand of the second GEMM: the C-fragment
// column order (2t, 2t+1) differs from the A-fragment order (t, t+4), which is absorbed by reading the B operand
// (V / K / dO / Q rows) in the matching permuted order — no shuffles.
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int HD = 32;          // head dimension
constexpr int SP = 36;          // smem row pitch (floats)
constexpr int TILE = 256;       // column-side rows staged per outer iteration
constexpr int WARPS = 8;        // 128 row-side tokens per CTA
constexpr int CHUNK = 64;       // columns per score chunk
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage rows [row0, row0+rows) of one head (channels d*heads + h) as tf32 into dst[r][SP]; rows >= valid are zero
__device__ __forceinline__ void stage_head_rows(uint32_t* __restrict__ dst, const float* __restrict__ src, int pitch,
                                                long long row0, int rows_valid, int rows_pad, int heads, int h) {
  for (int i = threadIdx.x; i < rows_pad * HD; i += blockDim.x) {
    const int r = i >> 5, d = i & 31;
    float v = 0.f;
    if (r < rows_valid) v = src[(row0 + r) * pitch + d * heads + h];
    dst[r * SP + d] = tf32(v);
  }
}

// A fragments (4 k-steps over d) of 16 rows starting at `row0`, scaled; rows >= T read as zero
__device__ __forceinline__ void load_row_frags(uint32_t (&a)[4][4], const float* __restrict__ src, int pitch,
                                               long long base_row, int row0, int T, int heads, int h, float scale,
                                               int g, int t) {
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int d0 = ks * 8 + t, d1 = d0 + 4;
    a[ks][0] = tf32(r0 < T ? src[(base_row + r0) * pitch + d0 * heads + h] * scale : 0.f);
    a[ks][1] = tf32(r1 < T ? src[(base_row + r1) * pitch + d0 * heads + h] * scale : 0.f);
    a[ks][2] = tf32(r0 < T ? src[(base_row + r0) * pitch + d1 * heads + h] * scale : 0.f);
    a[ks][3] = tf32(r1 < T ? src[(base_row + r1) * pitch + d1 * heads + h] * scale : 0.f);
  }
}

// c[j] (16 x 8 scores of n-tile j) = A(16 x 32) . B^T, B rows n0 + j*8 + g of the staged tile
__device__ __forceinline__ void scores_chunk(float (&c)[8][4], const uint32_t (&a)[4][4], const uint32_t* __restrict__ sB,
                                             int n0, int g, int t) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
    const uint32_t* row = sB + (n0 + j * 8 + g) * SP + t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_tf32(c[j], a[ks], row[ks * 8], row[ks * 8 + 4]);
  }
}

// acc(16 x 32) += P(16 x 64, C-fragment layout in p[j]) . B(64 x 32), B rows n0 + j*8 + {2t, 2t+1}
__device__ __forceinline__ void accumulate_chunk(float (&acc)[4][4], const float (&p)[8][4],
                                                 const uint32_t* __restrict__ sB, int n0, int g, int t) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a[4] = {tf32(p[j][0]), tf32(p[j][2]), tf32(p[j][1]), tf32(p[j][3])};
    const uint32_t* r0 = sB + (n0 + j * 8 + 2 * t) * SP + g;
    const uint32_t* r1 = r0 + SP;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) mma_tf32(acc[nd], a, r0[nd * 8], r1[nd * 8]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int pitch,
                   int T, int heads, float scale, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                   float* __restrict__ lse) {
  extern __shared__ uint32_t smu[];
  uint32_t* sK = smu;
  uint32_t* sV = smu + TILE * SP;
  const int b = blockIdx.y / heads, h = blockIdx.y % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * (WARPS * 16) + warp * 16;
  const long long base = (long long)b * T;
  const int C = HD * heads;

  uint32_t aq[4][4];
  load_row_frags(aq, q, pitch, base, row0, T, heads, h, scale * LOG2E, g, t);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[4][4];
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;

  for (int k0 = 0; k0 < T; k0 += TILE) {
    const int kt = min(TILE, T - k0);
    const int kt_pad = (kt + CHUNK - 1) / CHUNK * CHUNK;
    __syncthreads();
    stage_head_rows(sK, k, pitch, base + k0, kt, kt_pad, heads, h);
    stage_head_rows(sV, v, pitch, base + k0, kt, kt_pad, heads, h);
    __syncthreads();
    for (int kc = 0; kc < kt_pad; kc += CHUNK) {
      float s[8][4];
      scores_chunk(s, aq, sK, kc, g, t);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kc + j * 8 + 2 * t;
        if (key >= kt) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= kt) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
      const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
        s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
        ps0 += s[j][0] + s[j][1];
        ps1 += s[j][2] + s[j][3];
      }
      l0 = l0 * c0 + ps0;
      l1 = l1 * c1 + ps1;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) { o[nd][0] *= c0; o[nd][1] *= c0; o[nd][2] *= c1; o[nd][3] *= c1; }
      accumulate_chunk(o, s, sV, kc, g, t);
    }
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    const int d = nd * 8 + 2 * t;
    if (r0 < T) {
      const long long o0 = (base + r0) * C + (long long)d * heads + h;
      const float v0 = o[nd][0] * i0, v1 = o[nd][1] * i0;
      if (out_f32) { out_f32[o0] = v0; out_f32[o0 + heads] = v1; }
      if (out_bf16) { out_bf16[o0] = __float2bfloat16(v0); out_bf16[o0 + heads] = __float2bfloat16(v1); }
    }
    if (r1 < T) {
      const long long o1 = (base + r1) * C + (long long)d * heads + h;
      const float v0 = o[nd][2] * i1, v1 = o[nd][3] * i1;
      if (out_f32) { out_f32[o1] = v0; out_f32[o1 + heads] = v1; }
      if (out_bf16) { out_bf16[o1] = __float2bfloat16(v0); out_bf16[o1 + heads] = __float2bfloat16(v1); }
    }
  }
  if (lse && t == 0) {
    const float ln2 = 0.6931471805599453f;
    if (r0 < T) lse[((long long)b * heads + h) * T + r0] = (m0 + log2f(l0)) * ln2;
    if (r1 < T) lse[((long long)b * heads + h) * T + r1] = (m1 + log2f(l1)) * ln2;
  }
}

// ---------------------------------------------------------------------------------------------- backward: dQ
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dq_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int pitch,
                      const float* __restrict__ o, const float* __restrict__ dout, const float* __restrict__ lse,
                      int T, int heads, float scale, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dsum) {
  extern __shared__ uint32_t smu[];
  uint32_t* sK = smu;
  uint32_t* sV = smu + TILE * SP;
  const int b = blockIdx.y / heads, h = blockIdx.y % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * (WARPS * 16) + warp * 16;
  const long long base = (long long)b * T;
  const int C = HD * heads;
  const int r0 = row0 + g, r1 = row0 + g + 8;

  uint32_t aq[4][4], ado[4][4];
  load_row_frags(aq, q, pitch, base, row0, T, heads, h, scale * LOG2E, g, t);
  load_row_frags(ado, dout, C, base, row0, T, heads, h, 1.0f, g, t);
  // D = rowsum(dO * O): each thread owns 8 of the 32 d's of its two rows
  float D0 = 0.f, D1 = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int d0 = ks * 8 + t, d1 = d0 + 4;
    if (r0 < T) {
      D0 += dout[(base + r0) * C + d0 * heads + h] * o[(base + r0) * C + d0 * heads + h];
      D0 += dout[(base + r0) * C + d1 * heads + h] * o[(base + r0) * C + d1 * heads + h];
    }
    if (r1 < T) {
      D1 += dout[(base + r1) * C + d0 * heads + h] * o[(base + r1) * C + d0 * heads + h];
      D1 += dout[(base + r1) * C + d1 * heads + h] * o[(base + r1) * C + d1 * heads + h];
    }
  }
  D0 = quad_sum(D0);
  D1 = quad_sum(D1);
  const float L0 = (r0 < T ? lse[((long long)b * heads + h) * T + r0] : 0.f) * LOG2E;
  const float L1 = (r1 < T ? lse[((long long)b * heads + h) * T + r1] : 0.f) * LOG2E;
  if (t == 0) {
    if (r0 < T) dsum[((long long)b * heads + h) * T + r0] = D0;
    if (r1 < T) dsum[((long long)b * heads + h) * T + r1] = D1;
  }
  float dq[4][4];
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;

  for (int k0 = 0; k0 < T; k0 += TILE) {
    const int kt = min(TILE, T - k0);
    const int kt_pad = (kt + CHUNK - 1) / CHUNK * CHUNK;
    __syncthreads();
    stage_head_rows(sK, k, pitch, base + k0, kt, kt_pad, heads, h);
    stage_head_rows(sV, v, pitch, base + k0, kt, kt_pad, heads, h);
    __syncthreads();
    for (int kc = 0; kc < kt_pad; kc += CHUNK) {
      float s[8][4], dp[8][4];
      scores_chunk(s, aq, sK, kc, g, t);
      scores_chunk(dp, ado, sV, kc, g, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kc + j * 8 + 2 * t;
        const bool ok0 = key < kt, ok1 = key + 1 < kt;
        const float p0 = ok0 ? exp2f(s[j][0] - L0) : 0.f, p1 = ok1 ? exp2f(s[j][1] - L0) : 0.f;
        const float p2 = ok0 ? exp2f(s[j][2] - L1) : 0.f, p3 = ok1 ? exp2f(s[j][3] - L1) : 0.f;
        s[j][0] = p0 * (dp[j][0] - D0); s[j][1] = p1 * (dp[j][1] - D0);
        s[j][2] = p2 * (dp[j][2] - D1); s[j][3] = p3 * (dp[j][3] - D1);
      }
      accumulate_chunk(dq, s, sK, kc, g, t);
    }
  }
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    const int d = nd * 8 + 2 * t;
    if (r0 < T) {
      const long long o0 = (base + r0) * 3 * C + (long long)d * heads + h;
      dqkv[o0] = __float2bfloat16(dq[nd][0] * scale);
      dqkv[o0 + heads] = __float2bfloat16(dq[nd][1] * scale);
    }
    if (r1 < T) {
      const long long o1 = (base + r1) * 3 * C + (long long)d * heads + h;
      dqkv[o1] = __float2bfloat16(dq[nd][2] * scale);
      dqkv[o1 + heads] = __float2bfloat16(dq[nd][3] * scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward: dK, dV
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dkv_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       int pitch, const float* __restrict__ dout, const float* __restrict__ lse,
                       const float* __restrict__ dsum, int T, int heads, float scale,
                       __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ uint32_t smu[];
  uint32_t* sQ = smu;
  uint32_t* sDO = smu + TILE * SP;
  float* sL = reinterpret_cast<float*>(smu + 2 * TILE * SP);
  float* sD = sL + TILE;
  const int b = blockIdx.y / heads, h = blockIdx.y % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * (WARPS * 16) + warp * 16;      // key rows of this warp
  const long long base = (long long)b * T;
  const int C = HD * heads;
  const int r0 = row0 + g, r1 = row0 + g + 8;

  uint32_t ak[4][4], av[4][4];
  load_row_frags(ak, k, pitch, base, row0, T, heads, h, scale * LOG2E, g, t);
  load_row_frags(av, v, pitch, base, row0, T, heads, h, 1.0f, g, t);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    dk[nd][0] = dk[nd][1] = dk[nd][2] = dk[nd][3] = 0.f;
    dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.f;
  }
  for (int q0 = 0; q0 < T; q0 += TILE) {
    const int qt = min(TILE, T - q0);
    const int qt_pad = (qt + CHUNK - 1) / CHUNK * CHUNK;
    __syncthreads();
    stage_head_rows(sQ, q, pitch, base + q0, qt, qt_pad, heads, h);
    stage_head_rows(sDO, dout, C, base + q0, qt, qt_pad, heads, h);
    for (int i = threadIdx.x; i < qt_pad; i += blockDim.x) {
      sL[i] = i < qt ? lse[((long long)b * heads + h) * T + q0 + i] * LOG2E : 0.f;
      sD[i] = i < qt ? dsum[((long long)b * heads + h) * T + q0 + i] : 0.f;
    }
    __syncthreads();
    for (int qc = 0; qc < qt_pad; qc += CHUNK) {
      float st[8][4], dpt[8][4];
      scores_chunk(st, ak, sQ, qc, g, t);        // S^T: rows = keys, columns = queries
      scores_chunk(dpt, av, sDO, qc, g, t);      // dP^T
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int qi = qc + j * 8 + 2 * t;
        const bool ok0 = qi < qt, ok1 = qi + 1 < qt;
        const float La = sL[qi], Lb = sL[qi + 1], Da = sD[qi], Db = sD[qi + 1];
        const float p0 = ok0 ? exp2f(st[j][0] - La) : 0.f, p1 = ok1 ? exp2f(st[j][1] - Lb) : 0.f;
        const float p2 = ok0 ? exp2f(st[j][2] - La) : 0.f, p3 = ok1 ? exp2f(st[j][3] - Lb) : 0.f;
        st[j][0] = p0; st[j][1] = p1; st[j][2] = p2; st[j][3] = p3;
        dpt[j][0] = p0 * (dpt[j][0] - Da); dpt[j][1] = p1 * (dpt[j][1] - Db);
        dpt[j][2] = p2 * (dpt[j][2] - Da); dpt[j][3] = p3 * (dpt[j][3] - Db);
      }
      accumulate_chunk(dv, st, sDO, qc, g, t);   // dV += P^T dO
      accumulate_chunk(dk, dpt, sQ, qc, g, t);   // dK += dS^T Q
    }
  }
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    const int d = nd * 8 + 2 * t;
    if (r0 < T) {
      const long long o0 = (base + r0) * 3 * C + (long long)d * heads + h;
      dqkv[o0 + C] = __float2bfloat16(dk[nd][0] * scale);
      dqkv[o0 + C + heads] = __float2bfloat16(dk[nd][1] * scale);
      dqkv[o0 + 2 * C] = __float2bfloat16(dv[nd][0]);
      dqkv[o0 + 2 * C + heads] = __float2bfloat16(dv[nd][1]);
    }
    if (r1 < T) {
      const long long o1 = (base + r1) * 3 * C + (long long)d * heads + h;
      dqkv[o1 + C] = __float2bfloat16(dk[nd][2] * scale);
      dqkv[o1 + C + heads] = __float2bfloat16(dk[nd][3] * scale);
      dqkv[o1 + 2 * C] = __float2bfloat16(dv[nd][2]);
      dqkv[o1 + 2 * C + heads] = __float2bfloat16(dv[nd][3]);
    }
  }
}

constexpr size_t SMEM_FWD = (size_t)2 * TILE * SP * sizeof(uint32_t);
constexpr size_t SMEM_DKV = SMEM_FWD + 2 * TILE * sizeof(float);

}  // namespace
}  // namespace tvae

using namespace tvae;

// 1 (default): tcgen05 kind::tf32 kernels of attention_sm100.cu; 0: the mma.sync kernels of this file. Same precision
// class (TF32 operands, fp32 accumulation), different summation order.
static int g_attn_tcgen05 = 1;
extern "C" int32_t tvae_attn_set_tcgen05(int32_t enable) {
  const int old = g_attn_tcgen05;
  if (enable >= 0) g_attn_tcgen05 = enable ? 1 : 0;
  return old;
}

extern "C" int32_t tvae_attn_fwd_tc(const float* q, const float* k, const float* v, int32_t pitch, int32_t B, int32_t T,
                                    int32_t C, int32_t heads, void* out_bf16, float* out_f32, float* lse,
                                    cudaStream_t stream) {
  TVAE_ENTER(q);
  TVAE_CHECK(q && k && v && (out_bf16 || out_f32), "tvae_attn_fwd_tc: null pointer");
  TVAE_CHECK(heads > 0 && C == HD * heads, "tvae_attn_fwd_tc: head dimension must be 32 (C = %d, heads = %d)", C, heads);
  if (g_attn_tcgen05) return attn_fwd_sm100(q, k, v, pitch, B, T, heads, out_bf16, out_f32, lse, stream);
  const float scale = 1.0f / sqrtf((float)HD);
  TVAE_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD));
  dim3 grid((T + WARPS * 16 - 1) / (WARPS * 16), B * heads);
  attn_fwd_tc_kernel<<<grid, WARPS * 32, SMEM_FWD, stream>>>(q, k, v, pitch, T, heads, scale,
                                                            reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, lse);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t tvae_attn_bwd_tc(const float* q, const float* k, const float* v, int32_t pitch, const float* o,
                                    const float* d_out, const float* lse, int32_t B, int32_t T, int32_t C,
                                    int32_t heads, void* dqkv_bf16, float* workspace, cudaStream_t stream) {
  TVAE_ENTER(q);
  TVAE_CHECK(q && k && v && o && d_out && lse && dqkv_bf16 && workspace, "tvae_attn_bwd_tc: null pointer");
  TVAE_CHECK(heads > 0 && C == HD * heads, "tvae_attn_bwd_tc: head dimension must be 32 (C = %d, heads = %d)", C, heads);
  if (g_attn_tcgen05) return attn_bwd_sm100(q, k, v, pitch, o, d_out, lse, B, T, heads, dqkv_bf16, workspace, stream);
  const float scale = 1.0f / sqrtf((float)HD);
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
  dim3 grid((T + WARPS * 16 - 1) / (WARPS * 16), B * heads);
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD));
  attn_bwd_dq_tc_kernel<<<grid, WARPS * 32, SMEM_FWD, stream>>>(q, k, v, pitch, o, d_out, lse, T, heads, scale, dp,
                                                               workspace);
  TVAE_CUDA(cudaGetLastError());
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DKV));
  attn_bwd_dkv_tc_kernel<<<grid, WARPS * 32, SMEM_DKV, stream>>>(q, k, v, pitch, d_out, lse, workspace, T, heads, scale,
                                                                dp);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
