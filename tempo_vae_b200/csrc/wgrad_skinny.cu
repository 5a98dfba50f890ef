// Weight gradient of the LAST FEW channels of a wide, awkward channel count (1028 = 8 x 128 + 4: encoder.conv_in's
// input channels, decoder.conv_out's output channels; src/model.py:424-431, 634-640) for 3x3 stride-1 convolutions:
//
//   out[c][tap][n] = sum_{pixel} Wd[pixel][n] * S[pixel + sign * tap][c]        c < Cs <= 4, n < Cw, tap = (ky, kx)
//
// Wd = the wide operand (dY for conv_in, x for conv_out; bf16 NHWC, Cw a multiple of 64, <= 512), S = the skinny one
// (the 4 tail channels of x resp. dY), shifted by the tap with zero padding at the image border; sign = +1 when the
// tap shifts S relative to Wd's pixel (conv_in: dW[n][c][tap] = sum dY[px][n] x[px + tap][c]), -1 when it shifts Wd
// (conv_out: dW[c][n][tap] = sum dY[px][c] x[px + tap][n] = sum x[px'][n] dY[px' - tap][c]).
//
// Why a kernel of its own: on the tcgen05 weight-gradient GEMM these 4 channels cost a whole padded 128-row M tile
// (1.3-1.8 ms per launch), or, as a 16-column N tile, re-stream the wide operand once per tap (operand-bound, 1.7 ms).
// Here the nine taps sit on the M side of a 36(48) x Cw x pixels GEMM, so the wide operand (1.07 GB at B=256) is read
// exactly ONCE: TMA boxes {64 ch, 64 px} (SWIZZLE_128B) into a 3-stage ring, B fragments by ldmatrix.trans, the A tile
// [(tap, c)][64 px] gathered from L2 one chunk ahead (8 bytes per (tap, pixel)), mma.sync.m16n8k16 bf16 with fp32
// accumulators held in registers across the CTA's whole pixel range. HBM-bound: 0.04 % of the step's flops.
// Per-CTA partials go to a workspace and are summed in fixed order (bit-reproducible) into the parameter's layout.
#include <algorithm>
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int KP = 64;                  // pixels per chunk
constexpr int NTH = 256;                // 8 warps, one 64-channel box of the wide operand each
constexpr int STAGES = 3;
constexpr int MROWS = 48;               // 9 taps x <= 4 channels = 36 rows, padded to three m16 tiles
constexpr int A_BYTES = MROWS * 128;    // [48][64 px] bf16, 128-byte rows, 16-byte chunks XOR (row & 7)
constexpr int BOX_BYTES = 64 * KP * 2;  // one {64 ch, 64 px} box
constexpr int PAIRS = 9 * KP;           // (tap, pixel) pairs per chunk
constexpr int PAIRS_PER_THREAD = (PAIRS + NTH - 1) / NTH;

struct SkinnyParams {
  const __nv_bfloat16* s;
  long long s_pitch;      // elements
  int Cs, Cw, nbox;
  int H, W;
  int h_shift, w_shift;   // log2 when H / W are powers of two, else -1
  long long pixels;
  int sign;
  int nchunks;
  float* partial;         // [grid][36][Cw]
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// the 4 (Cs) skinny channels of every (tap, pixel) pair this thread owns, for the chunk starting at pixel p0
__device__ __forceinline__ void gather_pairs(uint2 (&v)[PAIRS_PER_THREAD], const SkinnyParams& p, long long p0) {
#pragma unroll
  for (int i = 0; i < PAIRS_PER_THREAD; ++i) {
    v[i] = make_uint2(0u, 0u);
    const int idx = (int)threadIdx.x + i * NTH;
    if (idx >= PAIRS) continue;
    const int k = idx & (KP - 1), tap = idx >> 6;
    const uint32_t px = (uint32_t)p0 + (uint32_t)k;          // pixels < 2^31 (checked by the host)
    if (px >= (uint32_t)p.pixels) continue;
    int x, y;
    if (p.w_shift >= 0 && p.h_shift >= 0) {                    // power-of-two images: no division (64-bit % cost more
      x = (int)(px & (uint32_t)(p.W - 1));                     // than the chunk's MMAs in the first version)
      y = (int)((px >> p.w_shift) & (uint32_t)(p.H - 1));
    } else {
      const uint32_t row = px / (uint32_t)p.W;
      x = (int)(px - row * (uint32_t)p.W);
      y = (int)(row % (uint32_t)p.H);
    }
    const int dy = p.sign * (tap / 3 - 1), dx = p.sign * (tap % 3 - 1);
    if ((unsigned)(y + dy) >= (unsigned)p.H || (unsigned)(x + dx) >= (unsigned)p.W) continue;
    const __nv_bfloat16* src = p.s + (long long)((int)px + dy * p.W + dx) * p.s_pitch;
    if (p.Cs == 4) {
      v[i] = __ldg(reinterpret_cast<const uint2*>(src));
    } else {
      uint32_t e[4] = {0u, 0u, 0u, 0u};
      for (int c = 0; c < p.Cs; ++c) e[c] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(src) + c);
      v[i] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
    }
  }
}
__device__ __forceinline__ void scatter_pairs(const uint2 (&v)[PAIRS_PER_THREAD], uint8_t* a_tile, int Cs) {
#pragma unroll
  for (int i = 0; i < PAIRS_PER_THREAD; ++i) {
    const int idx = (int)threadIdx.x + i * NTH;
    if (idx >= PAIRS) continue;
    const int k = idx & (KP - 1), tap = idx >> 6;
    const uint32_t e[4] = {v[i].x & 0xffffu, v[i].x >> 16, v[i].y & 0xffffu, v[i].y >> 16};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < Cs) {
        const int m = tap * Cs + c;
        *reinterpret_cast<unsigned short*>(a_tile + m * 128 + ((((k >> 3) ^ (m & 7)) << 4) | ((k & 7) << 1))) =
            (unsigned short)e[c];
      }
    }
  }
}

__global__ void __launch_bounds__(NTH, 1)
wgrad_skinny_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ SkinnyParams p) {
  extern __shared__ uint8_t smraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = p.nbox * BOX_BYTES;
  uint8_t* sA = sm + STAGES * stage_bytes;                       // two A tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + 2 * A_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nlocal = ((int)blockIdx.x < p.nchunks) ? (p.nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto chunk_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&wmap);
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * A_BYTES / 16; i += NTH) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  auto issue = [&](int i) {   // thread 0 only
    const int s = i % STAGES;
    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
    const long long c = chunk_of(i);
    for (int j = 0; j < p.nbox; ++j)
      tma_load_2d(&wmap, &full_bar[s], sm + s * stage_bytes + j * BOX_BYTES, j * 64, (int)(c * KP));
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < STAGES && i < nlocal; ++i) issue(i);

  uint2 pairs[PAIRS_PER_THREAD];
  if (nlocal > 0) {
    gather_pairs(pairs, p, chunk_of(0) * KP);
    scatter_pairs(pairs, sA, p.Cs);
  }
  __syncthreads();
  if (nlocal > 1) gather_pairs(pairs, p, chunk_of(1) * KP);

  float acc[3][8][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  const bool active = warp < p.nbox;            // warp w owns channels [64 w, 64 w + 64)
  // ldmatrix lane addressing. A (row-major [m][k]): matrix i = lane / 8 -> rows (i & 1) * 8 + lane % 8, k chunk i >> 1.
  // B (.trans of [px][ch]): matrix i -> pixel rows (i & 1) * 8 + lane % 8, channel chunk (i >> 1).
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lchunk = lane >> 4;
  for (int i = 0; i < nlocal; ++i) {
    mbar_wait(&full_bar[i % STAGES], (uint32_t)((i / STAGES) & 1), 41);
    if (active) {
      const uint32_t aT = smem_u32(sA + (i & 1) * A_BYTES);
      const uint32_t bT = smem_u32(sm + (i % STAGES) * stage_bytes + warp * BOX_BYTES);
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks) {
        uint32_t a[3][4];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const int m = mt * 16 + lrow;
          ldmatrix_x4(a[mt], aT + m * 128 + (((ks * 2 + lchunk) ^ (m & 7)) << 4));
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          const int r = ks * 16 + lrow;
          ldmatrix_x4_trans(b, bT + r * 128 + (((np * 2 + lchunk) ^ (r & 7)) << 4));
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            mma_bf16(acc[mt][np * 2], a[mt], b[0], b[1]);
            mma_bf16(acc[mt][np * 2 + 1], a[mt], b[2], b[3]);
          }
        }
      }
    }
    if (i + 1 < nlocal) scatter_pairs(pairs, sA + ((i + 1) & 1) * A_BYTES, p.Cs);
    __syncthreads();      // every warp is done with W stage i % STAGES and A tile i & 1; A tile (i + 1) & 1 is complete
    if (threadIdx.x == 0 && i + STAGES < nlocal) issue(i + STAGES);
    if (i + 2 < nlocal) gather_pairs(pairs, p, chunk_of(i + 2) * KP);
  }
  if (active) {
    const int g = lane >> 2, t = lane & 3;
    const int mreal = 9 * p.Cs;
    float* dst = p.partial + (long long)blockIdx.x * 36 * p.Cw;
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = warp * 64 + nt * 8 + 2 * t;
        const int m0 = mt * 16 + g, m1 = m0 + 8;
        if (m0 < mreal) *reinterpret_cast<float2*>(dst + (long long)m0 * p.Cw + n) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        if (m1 < mreal) *reinterpret_cast<float2*>(dst + (long long)m1 * p.Cw + n) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
  }
}

// out[c * stride_c + n * stride_n + tap] (+)= sum_g partial[g][tap * Cs + c][n], g in fixed order
__global__ void wgrad_skinny_reduce_kernel(const float* __restrict__ partial, int nparts, int Cs, int Cw,
                                           float* __restrict__ grad, long long stride_c, long long stride_n,
                                           int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 9 * Cs * Cw;
  if (idx >= total) return;
  const int n = idx % Cw, m = idx / Cw;
  const int tap = m / Cs, c = m % Cs;
  float s = 0.f;
  for (int g = 0; g < nparts; ++g) s += partial[((long long)g * 36 + m) * Cw + n];
  float* o = grad + c * stride_c + n * stride_n + tap;
  *o = accumulate ? *o + s : s;
}

int grid_for(long long nchunks) { return (int)std::min<long long>(num_sms(), std::max<long long>(nchunks, 1)); }
PerDeviceOnce g_once;

}  // namespace
}  // namespace tvae

using namespace tvae;

extern "C" int64_t tvae_wgrad_skinny_workspace_bytes(int32_t Cw) {
  return (int64_t)num_sms() * 36 * (int64_t)Cw * (int64_t)sizeof(float);
}

extern "C" int32_t tvae_wgrad_skinny(const void* wide_bf16, int32_t Cw, int32_t wide_pitch, const void* skinny_bf16,
                                     int32_t Cs, int32_t skinny_pitch, int32_t N, int32_t H, int32_t W,
                                     int32_t shift_sign, float* grad, int64_t stride_c, int64_t stride_n,
                                     int32_t accumulate, float* workspace, cudaStream_t stream) {
  TVAE_ENTER(wide_bf16);
  TVAE_CHECK(wide_bf16 && skinny_bf16 && grad && workspace, "tvae_wgrad_skinny: null pointer");
  TVAE_CHECK(Cw > 0 && Cw % 64 == 0 && Cw <= 512, "tvae_wgrad_skinny: wide channels must be a multiple of 64, <= 512 (got %d)", Cw);
  TVAE_CHECK(Cs >= 1 && Cs <= 4, "tvae_wgrad_skinny: 1..4 skinny channels (got %d)", Cs);
  TVAE_CHECK(shift_sign == 1 || shift_sign == -1, "tvae_wgrad_skinny: shift_sign must be +1 or -1");
  TVAE_CHECK(wide_pitch % 8 == 0 && ((uintptr_t)wide_bf16 & 15) == 0, "tvae_wgrad_skinny: wide operand must be 16-byte aligned with pitch %% 8 == 0");
  TVAE_CHECK(Cs != 4 || (skinny_pitch % 4 == 0 && ((uintptr_t)skinny_bf16 & 7) == 0),
             "tvae_wgrad_skinny: 4 skinny channels are read as 8-byte words (pointer and pitch must allow it)");
  const long long pixels = (long long)N * H * W;
  TVAE_CHECK(pixels > 0 && pixels < (1ll << 31), "tvae_wgrad_skinny: bad pixel count");
  CUtensorMap wmap;
  const uint64_t dims[2] = {(uint64_t)Cw, (uint64_t)pixels};
  const uint64_t strides[1] = {(uint64_t)wide_pitch * 2};
  const uint32_t box[2] = {64, (uint32_t)KP};
  if (make_tmap_bf16(&wmap, wide_bf16, 2, dims, strides, box) != 0) return -1;
  SkinnyParams p;
  p.s = reinterpret_cast<const __nv_bfloat16*>(skinny_bf16);
  p.s_pitch = skinny_pitch;
  p.Cs = Cs; p.Cw = Cw; p.nbox = Cw / 64;
  p.H = H; p.W = W;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.h_shift = log2_or_neg(H); p.w_shift = log2_or_neg(W);
  p.pixels = pixels;
  p.sign = shift_sign;
  p.nchunks = (int)((pixels + KP - 1) / KP);
  p.partial = workspace;
  const int grid = grid_for(p.nchunks);
  const int smem = STAGES * p.nbox * BOX_BYTES + 2 * A_BYTES + 64 + 1024;
  if (g_once.pending()) {
    TVAE_CUDA(cudaFuncSetAttribute(wgrad_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   STAGES * 8 * BOX_BYTES + 2 * A_BYTES + 64 + 1024));
    g_once.mark();
  }
  wgrad_skinny_kernel<<<grid, NTH, smem, stream>>>(wmap, p);
  TVAE_CUDA(cudaGetLastError());
  const int total = 9 * Cs * Cw;
  wgrad_skinny_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(workspace, grid, Cs, Cw, grad, stride_c, stride_n,
                                                                     accumulate);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
