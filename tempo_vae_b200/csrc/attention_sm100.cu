// Mid-block self-attention on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators and the
// probability operand in TMEM) for head dimension 32: forward, dQ and dK/dV. Replaces the einsum / softmax / einsum of
// AttnBlock.forward (src/model.py:128-139) and its autograd backward, like attention_tc.cu (mma.sync), with the same
// C ABI, layouts and TF32 precision (10 mantissa bits on Q, K, V, dO, P, dS; fp32 accumulation and softmax).
//
// Layout: one head's slice of a token row is 32 floats = 128 bytes = exactly one 128-byte swizzle row. A tile of
// token rows staged as [row][32 floats] is
//   * a K-major operand with the tokens on M/N and the head dimension on K  (Q K^T, dO V^T, K Q^T, V dO^T) when its
//     16-byte chunks are XOR-swizzled by row & 7 (SWIZZLE_128B; K step = 32 B along the row), and
//   * an MN-major operand with the head dimension on N and the tokens on K  (P V, dS K, P^T dO, dS^T Q) when its
//     32-byte chunks are XOR-swizzled by row & 3 (SWIZZLE_128B_BASE32B, the only MN-major layout tf32 has; K step =
//     8 rows = 1 KB down) -- measured: the plain SWIZZLE_128B descriptor with the MN-major bit reads zeros.
// Nothing is ever transposed in shared memory; a tile that plays both roles (K in dQ, Q and dO in dK/dV) is stored
// twice from the same loaded registers. The first GEMM of a pair is smem x smem, the second takes its A operand (the
// probabilities / score gradients, written back over the scores by tcgen05.st) straight from TMEM.
//
// One CTA = 128 threads = 128 row-side tokens = the 128 TMEM lanes; thread i owns row i of every accumulator, so the
// softmax needs no shuffles. Column-side tokens stream through in tiles of 64. Heads are channel-interleaved in global
// memory (channel = d * heads + h, the reference's reshape(b, c_, n_heads, hw)), which no TMA box can express
// (4-byte inner extent), so tiles are staged with plain loads; the whole working set (q, k, v, dO of one sample:
// 0.5 MB) lives in L2. Several CTAs share an SM (TMEM: 128 resp. 256 of 512 columns each) to overlap one CTA's staging
// with another's MMAs and softmax.
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int HD = 32;                 // head dimension
constexpr int RT = 128;                // row-side tokens per CTA (TMEM lanes)
constexpr int CT = 64;                 // column-side tokens per tile
constexpr int NT = 128;                // threads per CTA
constexpr int ROW_TILE_BYTES = RT * 128;
constexpr int COL_TILE_BYTES = CT * 128;
constexpr int OPITCH = 33;             // floats per row of the output transposition buffer (bank-conflict free)
constexpr float LOG2E = 1.4426950408889634f;

// Round to TF32 (10 mantissa bits), ties away from zero, as two integer instructions. cvt.rna.tf32.f32 lowers to a
// five-instruction sequence with an Inf/NaN check; every operand element passes through here, and none of them is
// Inf/NaN unless the inputs already were.
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// kind::tf32, fp32 accumulate (InstrDescriptor of cute/arch/mma_sm100_desc.hpp: a/b format 2 = TF32)
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // c_format = F32
  d |= 2u << 7;                          // a_format = TF32
  d |= 2u << 10;                         // b_format = TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (descriptor layout type 1, Swizzle<2,5,2>:
// the 32-byte chunk index of a 128-byte row is XORed with row & 3; atoms of 4 K-rows, SBO between atoms)
__device__ __forceinline__ uint64_t make_smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem: 128 lanes x 8 columns] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Staging of `ROWS` token rows of head h (channels d * heads + h) starting at global row `row0`, split into the global
// loads (into registers: issued one tile AHEAD so their L2 latency hides behind the MMAs and the softmax of the
// current tile -- with two to three CTAs of four warps per SM nothing else would hide it) and the tf32 stores into a
// K-major tile (dst_k: element (r, d) at r * 128 + ((d / 4) ^ (r & 7)) * 16 + (d & 3) * 4) and / or an MN-major tile
// (dst_mn: r * 128 + ((d / 8) ^ (r & 3)) * 32 + (d & 7) * 4). Rows >= valid are zero. One warp per row, lane = d: the
// 32 loads of a row cover 32 * heads * 4 contiguous bytes, the 32 stores hit 32 different banks.
template <int ROWS>
__device__ __forceinline__ void load_tile(float (&regs)[ROWS / 4], const float* __restrict__ src, long long pitch,
                                          long long row0, int valid, int heads, int h) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = src + (row0 + warp) * pitch + lane * heads + h;
  const uint32_t step = 4u * (uint32_t)pitch;     // a tile spans < 2^31 elements: 32-bit offsets from one base pointer
  if (valid >= ROWS) {
#pragma unroll
    for (int i = 0; i < ROWS / 4; ++i) regs[i] = __ldg(p + (uint32_t)i * step);
  } else {
#pragma unroll
    for (int i = 0; i < ROWS / 4; ++i) {
      regs[i] = 0.f;
      if (warp + 4 * i < valid) regs[i] = __ldg(p + (uint32_t)i * step);
    }
  }
}
template <int ROWS>
__device__ __forceinline__ void store_tile(const float (&regs)[ROWS / 4], uint8_t* __restrict__ dst_k,
                                           uint8_t* __restrict__ dst_mn, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c16 = (uint32_t)(lane >> 2), w16 = (uint32_t)(lane & 3) << 2;
  const uint32_t c32 = (uint32_t)(lane >> 3), w32 = (uint32_t)(lane & 7) << 2;
#pragma unroll
  for (int i = 0; i < ROWS / 4; ++i) {
    const int r = warp + 4 * i;
    const uint32_t t = tf32_rna(regs[i] * scale);
    if (dst_k) *reinterpret_cast<uint32_t*>(dst_k + r * 128 + (((c16 ^ (uint32_t)(r & 7)) << 4) | w16)) = t;
    if (dst_mn) *reinterpret_cast<uint32_t*>(dst_mn + r * 128 + (((c32 ^ (uint32_t)(r & 3)) << 5) | w32)) = t;
  }
}
__device__ __forceinline__ void stage_tile(uint8_t* __restrict__ dst_k, uint8_t* __restrict__ dst_mn,
                                           const float* __restrict__ src, long long pitch, long long row0, int valid,
                                           int heads, int h, float scale) {
  float regs[RT / 4];
  load_tile<RT>(regs, src, pitch, row0, valid, heads, h);
  store_tile<RT>(regs, dst_k, dst_mn, scale);
}

// Each thread hands over its row of 32 values; they are written out as out[(grow0 + r) * opitch + d * heads + h] with
// one warp per row (lane = d), rows >= valid skipped. `buf` holds RT * OPITCH floats. Ends with the data in flight
// only from registers (no trailing barrier needed before `buf` is reused by ANOTHER call: callers sync in between).
template <typename OutT>
__device__ __forceinline__ void write_rows(float* __restrict__ buf, const float (&vals)[HD], float mul,
                                           OutT* __restrict__ out, long long grow0, long long opitch, int valid,
                                           int heads, int h) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < HD; ++d) buf[threadIdx.x * OPITCH + d] = vals[d] * mul;
  __syncthreads();
#pragma unroll 8
  for (int r = warp; r < RT; r += NT / 32) {
    if (r < valid) {
      const float v = buf[r * OPITCH + lane];
      OutT* o = out + (grow0 + r) * opitch + lane * heads + h;
      if constexpr (sizeof(OutT) == 4) *o = v;
      else *o = __float2bfloat16(v);
    }
  }
}

struct Smem {
  uint8_t* base;       // 1024-byte aligned
  uint64_t* bars;      // [0]: score MMAs done, [1]: output MMAs done
  uint32_t* tmem_ptr;
};
__device__ __forceinline__ Smem carve(uint8_t* raw, int tile_bytes) {
  Smem s;
  s.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  s.bars = reinterpret_cast<uint64_t*>(s.base + tile_bytes);
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.bars + 2);
  return s;
}

// ---------------------------------------------------------------------------------------------- forward
// TMEM: S / P [0, 64), O tile [64, 96)  -> 128 columns, up to four CTAs per SM.
constexpr int FWD_TMEM_COLS = 128;
constexpr int FWD_TILE_BYTES = ROW_TILE_BYTES + 2 * COL_TILE_BYTES;      // Q | K | V = 32 KB

__global__ void __launch_bounds__(NT)
attn_fwd_sm100_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int pitch,
                      int T, int heads, float scale, __nv_bfloat16* __restrict__ out_bf16,
                      float* __restrict__ out_f32, float* __restrict__ lse) {
  extern __shared__ uint8_t smraw[];
  const Smem sm = carve(smraw, FWD_TILE_BYTES);
  uint8_t* sQ = sm.base;
  uint8_t* sK = sQ + ROW_TILE_BYTES;
  uint8_t* sV = sK + COL_TILE_BYTES;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;   // (sample, head) on x: no 65,535 limit on B * heads
  const int row0 = blockIdx.y * RT;
  const int rvalid = min(RT, T - row0);
  const long long base = (long long)b * T;
  const int C = HD * heads;

  if (warp == 0) tmem_alloc(sm.tmem_ptr, FWD_TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    fence_mbar_init();
  }
  stage_tile(sQ, nullptr, q, pitch, base + row0, rvalid, heads, h, scale * LOG2E);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem_ptr;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc_s = make_idesc_tf32(RT, CT, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(RT, HD, 0, 1);
  const uint64_t dQ = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
  const uint64_t dK = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
  const uint64_t dVmn = make_smem_desc_mn32(smem_u32(sV), 1024, 512);

  float m = -INFINITY, l = 0.f;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  uint32_t phase = 0;
  float rk[CT / 4], rv[CT / 4];
  load_tile<CT>(rk, k, pitch, base, min(CT, T), heads, h);
  load_tile<CT>(rv, v, pitch, base, min(CT, T), heads, h);
  for (int k0 = 0; k0 < T; k0 += CT) {
    const int kvalid = min(CT, T - k0);
    store_tile<CT>(rk, sK, nullptr, 1.0f);
    store_tile<CT>(rv, nullptr, sV, 1.0f);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) umma_tf32_ss(tmem, dQ + (uint64_t)(ks * 2), dK + (uint64_t)(ks * 2), idesc_s, ks > 0);
        umma_commit(&sm.bars[0]);
      }
      __syncwarp();
    }
    if (k0 + CT < T) {   // next tile's loads fly during this tile's MMAs and softmax
      load_tile<CT>(rk, k, pitch, base + k0 + CT, min(CT, T - k0 - CT), heads, h);
      load_tile<CT>(rv, v, pitch, base + k0 + CT, min(CT, T - k0 - CT), heads, h);
    }
    mbar_wait(&sm.bars[0], phase, 31);
    tc_fence_after();
    uint32_t s[CT];
    tmem_ld32(tlane, s);
    tmem_ld32(tlane + 32, s + 32);
    tmem_ld_wait();
    float mx = -INFINITY;
    if (kvalid < CT) {
#pragma unroll
      for (int c = 0; c < CT; ++c)
        if (c >= kvalid) s[c] = 0xff800000u;   // -inf
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
    const float mn = fmaxf(m, mx);
    const float alpha = ex2_approx(m - mn);
    m = mn;
    float ps = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float p = ex2_approx(__uint_as_float(s[c]) - mn);
      ps += p;
      s[c] = tf32_rna(p);
    }
    l = l * alpha + ps;
    tmem_st32(tlane, s);
    tmem_st32(tlane + 32, s + 32);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < CT / 8; ++ks)
          umma_tf32_ts(tmem + CT, tmem + (uint32_t)(ks * 8), dVmn + (uint64_t)(ks * 64), idesc_o, ks > 0);
        umma_commit(&sm.bars[1]);
      }
      __syncwarp();
    }
    mbar_wait(&sm.bars[1], phase, 32);
    tc_fence_after();
    uint32_t ot[HD];
    tmem_ld32(tlane + CT, ot);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = fmaf(o[d], alpha, __uint_as_float(ot[d]));
    phase ^= 1;
  }
  // all MMAs have completed (bars[1] of the last tile): the tiles are free, reuse them for the transposed write-out
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, FWD_TMEM_COLS);
  float* buf = reinterpret_cast<float*>(sm.base);
  const float inv = 1.0f / l;
  if (out_f32) {
    write_rows(buf, o, inv, out_f32, base + row0, (long long)C, rvalid, heads, h);
    __syncthreads();
  }
  if (out_bf16) write_rows(buf, o, inv, out_bf16, base + row0, (long long)C, rvalid, heads, h);
  if (lse && (int)threadIdx.x < rvalid)
    lse[((long long)b * heads + h) * T + row0 + threadIdx.x] = (m + log2f(l)) * 0.6931471805599453f;
}

// ---------------------------------------------------------------------------------------------- backward: dQ
// rows = queries, columns = keys. TMEM: S / dS [0, 64), dP [64, 128), dQ [128, 160) -> 256 columns, two CTAs per SM.
constexpr int BWD_TMEM_COLS = 256;
constexpr int BWD_TILE_BYTES = 2 * ROW_TILE_BYTES + 4 * COL_TILE_BYTES;  // 64 KB
constexpr int BWD_EXTRA_BYTES = 2 * 8 + 16 + 2 * CT * 4;                 // barriers, TMEM pointer, per-column L and D

__global__ void __launch_bounds__(NT)
attn_bwd_dq_sm100_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                         int pitch, const float* __restrict__ o, const float* __restrict__ dout,
                         const float* __restrict__ lse, int T, int heads, float scale,
                         __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dsum) {
  extern __shared__ uint8_t smraw[];
  const Smem sm = carve(smraw, BWD_TILE_BYTES);
  uint8_t* sQ = sm.base;
  uint8_t* sDO = sQ + ROW_TILE_BYTES;
  uint8_t* sK = sDO + ROW_TILE_BYTES;
  uint8_t* sV = sK + COL_TILE_BYTES;
  uint8_t* sKmn = sV + COL_TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;   // (sample, head) on x: no 65,535 limit on B * heads
  const int row0 = blockIdx.y * RT;
  const int rvalid = min(RT, T - row0);
  const long long base = (long long)b * T;
  const int C = HD * heads;
  float* sDrow = reinterpret_cast<float*>(sK);   // D of this CTA's rows, parked in the K tile until the loop starts

  if (warp == 0) tmem_alloc(sm.tmem_ptr, BWD_TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    fence_mbar_init();
  }
  stage_tile(sQ, nullptr, q, pitch, base + row0, rvalid, heads, h, scale * LOG2E);
  {
    // dO tile, and D = rowsum(dO * O) from the same loads (one warp per row, lane = d)
    const float* pd = dout + (base + row0) * C + lane * heads + h;
    const float* po = o + (base + row0) * C + lane * heads + h;
    const uint32_t chunk = (uint32_t)(lane >> 2), within = (uint32_t)(lane & 3) << 2;
#pragma unroll 4
    for (int r = warp; r < RT; r += NT / 32) {
      float g = 0.f, ov = 0.f;
      if (r < rvalid) { g = __ldg(pd + (long long)r * C); ov = __ldg(po + (long long)r * C); }
      *reinterpret_cast<uint32_t*>(sDO + r * 128 + (((chunk ^ (uint32_t)(r & 7)) << 4) | within)) = tf32_rna(g);
      const float dsum_r = warp_sum(g * ov);
      if (lane == 0) sDrow[r] = dsum_r;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem_ptr;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  const float D = sDrow[threadIdx.x];
  const bool rok = (int)threadIdx.x < rvalid;
  const float L = rok ? lse[((long long)b * heads + h) * T + row0 + threadIdx.x] * LOG2E : 0.f;
  if (rok) dsum[((long long)b * heads + h) * T + row0 + threadIdx.x] = D;
  __syncthreads();                              // sDrow (inside sK) is about to be overwritten by the first key tile

  const uint32_t idesc_s = make_idesc_tf32(RT, CT, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(RT, HD, 0, 1);
  const uint64_t dQd = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
  const uint64_t dDO = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
  const uint64_t dK = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
  const uint64_t dV = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
  const uint64_t dKmn = make_smem_desc_mn32(smem_u32(sKmn), 1024, 512);

  uint32_t phase = 0;
  float rk[CT / 4], rv[CT / 4];
  load_tile<CT>(rk, k, pitch, base, min(CT, T), heads, h);
  load_tile<CT>(rv, v, pitch, base, min(CT, T), heads, h);
  for (int k0 = 0; k0 < T; k0 += CT) {
    const int kvalid = min(CT, T - k0);
    store_tile<CT>(rk, sK, sKmn, 1.0f);
    store_tile<CT>(rv, sV, nullptr, 1.0f);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) umma_tf32_ss(tmem, dQd + (uint64_t)(ks * 2), dK + (uint64_t)(ks * 2), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) umma_tf32_ss(tmem + CT, dDO + (uint64_t)(ks * 2), dV + (uint64_t)(ks * 2), idesc_s, ks > 0);
        umma_commit(&sm.bars[0]);
      }
      __syncwarp();
    }
    if (k0 + CT < T) {
      load_tile<CT>(rk, k, pitch, base + k0 + CT, min(CT, T - k0 - CT), heads, h);
      load_tile<CT>(rv, v, pitch, base + k0 + CT, min(CT, T - k0 - CT), heads, h);
    }
    mbar_wait(&sm.bars[0], phase, 33);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < CT; c0 += 32) {
      uint32_t s[32], dp[32];
      tmem_ld32(tlane + c0, s);
      tmem_ld32(tlane + CT + c0, dp);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p = ex2_approx(__uint_as_float(s[c]) - L);
        float ds = p * (__uint_as_float(dp[c]) - D);
        if (c0 + c >= kvalid) ds = 0.f;
        s[c] = tf32_rna(ds);
      }
      tmem_st32(tlane + c0, s);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < CT / 8; ++ks)
          umma_tf32_ts(tmem + 2 * CT, tmem + (uint32_t)(ks * 8), dKmn + (uint64_t)(ks * 64), idesc_o, (k0 > 0) || (ks > 0));
        umma_commit(&sm.bars[1]);
      }
      __syncwarp();
    }
    mbar_wait(&sm.bars[1], phase, 34);   // K / V tiles and the S columns are free again
    tc_fence_after();
    phase ^= 1;
  }
  uint32_t acc[HD];
  tmem_ld32(tlane + 2 * CT, acc);
  tmem_ld_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, BWD_TMEM_COLS);
  float vals[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) vals[d] = __uint_as_float(acc[d]);
  write_rows(reinterpret_cast<float*>(sm.base), vals, scale, dqkv, base + row0, (long long)3 * C, rvalid, heads, h);
}

// ---------------------------------------------------------------------------------------------- backward: dK, dV
// rows = keys, columns = queries. TMEM: S^T / P^T [0, 64), dP^T / dS^T [64, 128), dV [128, 160), dK [160, 192).
__global__ void __launch_bounds__(NT)
attn_bwd_dkv_sm100_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                          int pitch, const float* __restrict__ dout, const float* __restrict__ lse,
                          const float* __restrict__ dsum, int T, int heads, float scale,
                          __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ uint8_t smraw[];
  const Smem sm = carve(smraw, BWD_TILE_BYTES);
  uint8_t* sK = sm.base;
  uint8_t* sV = sK + ROW_TILE_BYTES;
  uint8_t* sQ = sV + ROW_TILE_BYTES;
  uint8_t* sDO = sQ + COL_TILE_BYTES;
  uint8_t* sQmn = sDO + COL_TILE_BYTES;
  uint8_t* sDOmn = sQmn + COL_TILE_BYTES;
  float* sL = reinterpret_cast<float*>(sm.tmem_ptr + 4);
  float* sD = sL + CT;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;   // (sample, head) on x: no 65,535 limit on B * heads
  const int row0 = blockIdx.y * RT;
  const int rvalid = min(RT, T - row0);
  const long long base = (long long)b * T;
  const int C = HD * heads;

  if (warp == 0) tmem_alloc(sm.tmem_ptr, BWD_TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    fence_mbar_init();
  }
  stage_tile(sK, nullptr, k, pitch, base + row0, rvalid, heads, h, scale * LOG2E);
  stage_tile(sV, nullptr, v, pitch, base + row0, rvalid, heads, h, 1.0f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem_ptr;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc_s = make_idesc_tf32(RT, CT, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(RT, HD, 0, 1);
  const uint64_t dKd = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
  const uint64_t dVd = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
  const uint64_t dQd = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
  const uint64_t dDO = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
  const uint64_t dQmn = make_smem_desc_mn32(smem_u32(sQmn), 1024, 512);
  const uint64_t dDOmn = make_smem_desc_mn32(smem_u32(sDOmn), 1024, 512);

  uint32_t phase = 0;
  float rq[CT / 4], rdo[CT / 4], rl = 0.f, rd = 0.f;
  const long long stat0 = ((long long)b * heads + h) * T;
  load_tile<CT>(rq, q, pitch, base, min(CT, T), heads, h);
  load_tile<CT>(rdo, dout, C, base, min(CT, T), heads, h);
  if ((int)threadIdx.x < min(CT, T)) { rl = lse[stat0 + threadIdx.x]; rd = dsum[stat0 + threadIdx.x]; }
  for (int q0 = 0; q0 < T; q0 += CT) {
    store_tile<CT>(rq, sQ, sQmn, 1.0f);
    store_tile<CT>(rdo, sDO, sDOmn, 1.0f);
    if ((int)threadIdx.x < CT) {
      sL[threadIdx.x] = rl * LOG2E;
      sD[threadIdx.x] = rd;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) umma_tf32_ss(tmem, dKd + (uint64_t)(ks * 2), dQd + (uint64_t)(ks * 2), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) umma_tf32_ss(tmem + CT, dVd + (uint64_t)(ks * 2), dDO + (uint64_t)(ks * 2), idesc_s, ks > 0);
        umma_commit(&sm.bars[0]);
      }
      __syncwarp();
    }
    if (q0 + CT < T) {
      const int nvalid = min(CT, T - q0 - CT);
      load_tile<CT>(rq, q, pitch, base + q0 + CT, nvalid, heads, h);
      load_tile<CT>(rdo, dout, C, base + q0 + CT, nvalid, heads, h);
      rl = rd = 0.f;
      if ((int)threadIdx.x < nvalid) { rl = lse[stat0 + q0 + CT + threadIdx.x]; rd = dsum[stat0 + q0 + CT + threadIdx.x]; }
    }
    mbar_wait(&sm.bars[0], phase, 35);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < CT; c0 += 32) {
      uint32_t s[32], dp[32];
      tmem_ld32(tlane + c0, s);
      tmem_ld32(tlane + CT + c0, dp);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p = ex2_approx(__uint_as_float(s[c]) - sL[c0 + c]);
        const float ds = p * (__uint_as_float(dp[c]) - sD[c0 + c]);
        s[c] = tf32_rna(p);
        dp[c] = tf32_rna(ds);
      }
      tmem_st32(tlane + c0, s);
      tmem_st32(tlane + CT + c0, dp);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < CT / 8; ++ks)
          umma_tf32_ts(tmem + 2 * CT, tmem + (uint32_t)(ks * 8), dDOmn + (uint64_t)(ks * 64), idesc_o, (q0 > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < CT / 8; ++ks)
          umma_tf32_ts(tmem + 2 * CT + HD, tmem + (uint32_t)(CT + ks * 8), dQmn + (uint64_t)(ks * 64), idesc_o, (q0 > 0) || (ks > 0));
        umma_commit(&sm.bars[1]);
      }
      __syncwarp();
    }
    mbar_wait(&sm.bars[1], phase, 36);
    tc_fence_after();
    phase ^= 1;
  }
  uint32_t av[HD], ak[HD];
  tmem_ld32(tlane + 2 * CT, av);
  tmem_ld32(tlane + 2 * CT + HD, ak);
  tmem_ld_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, BWD_TMEM_COLS);
  float vals[HD];
  float* buf = reinterpret_cast<float*>(sm.base);
#pragma unroll
  for (int d = 0; d < HD; ++d) vals[d] = __uint_as_float(ak[d]);
  write_rows(buf, vals, scale, dqkv + C, base + row0, (long long)3 * C, rvalid, heads, h);
  __syncthreads();
#pragma unroll
  for (int d = 0; d < HD; ++d) vals[d] = __uint_as_float(av[d]);
  write_rows(buf, vals, 1.0f, dqkv + 2 * C, base + row0, (long long)3 * C, rvalid, heads, h);
}

// Dynamic shared memory: the tiles + 1 KB alignment slack, padded so that exactly as many CTAs fit an SM (227 KB) as
// its 512 TMEM columns can serve -- a CTA that is resident but blocked in tcgen05.alloc would only hold shared memory.
constexpr int SMEM_FWD = 44 * 1024;     // 4 per SM with the 1 KB per-CTA reserve (actual need 32 KB + 1 KB + barriers)
constexpr int SMEM_BWD = 100 * 1024;    // 2 per SM (actual need 64 KB + 1 KB + barriers + per-column L, D)
static_assert(FWD_TILE_BYTES + 1024 + 64 <= SMEM_FWD && RT * OPITCH * 4 <= FWD_TILE_BYTES, "forward shared memory");
static_assert(BWD_TILE_BYTES + 1024 + BWD_EXTRA_BYTES + 64 <= SMEM_BWD && RT * OPITCH * 4 <= BWD_TILE_BYTES, "backward shared memory");

PerDeviceOnce g_attr_once;
int set_attrs() {
  if (!g_attr_once.pending()) return 0;
  TVAE_CUDA(cudaFuncSetAttribute(attn_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dq_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
  TVAE_CUDA(cudaFuncSetAttribute(attn_fwd_sm100_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dq_sm100_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  TVAE_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_sm100_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  g_attr_once.mark();
  return 0;
}


}  // namespace

int attn_fwd_sm100(const float* q, const float* k, const float* v, int pitch, int B, int T, int heads, void* out_bf16,
                   float* out_f32, float* lse, cudaStream_t stream) {
  if (int rc = set_attrs()) return rc;
  TVAE_CHECK(B > 0 && T > 0 && (T + RT - 1) / RT <= 65535, "tvae_attn_fwd_tc: bad batch / sequence length (B = %d, T = %d)", B, T);
  const float scale = 1.0f / sqrtf((float)HD);
  dim3 grid(B * heads, (T + RT - 1) / RT);
  attn_fwd_sm100_kernel<<<grid, NT, SMEM_FWD, stream>>>(q, k, v, pitch, T, heads, scale,
                                                        reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, lse);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

int attn_bwd_sm100(const float* q, const float* k, const float* v, int pitch, const float* o, const float* d_out,
                   const float* lse, int B, int T, int heads, void* dqkv_bf16, float* workspace, cudaStream_t stream) {
  if (int rc = set_attrs()) return rc;
  TVAE_CHECK(B > 0 && T > 0 && (T + RT - 1) / RT <= 65535, "tvae_attn_bwd_tc: bad batch / sequence length (B = %d, T = %d)", B, T);
  const float scale = 1.0f / sqrtf((float)HD);
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
  dim3 grid(B * heads, (T + RT - 1) / RT);
  attn_bwd_dq_sm100_kernel<<<grid, NT, SMEM_BWD, stream>>>(q, k, v, pitch, o, d_out, lse, T, heads, scale, dp, workspace);
  TVAE_CUDA(cudaGetLastError());
  attn_bwd_dkv_sm100_kernel<<<grid, NT, SMEM_BWD, stream>>>(q, k, v, pitch, d_out, lse, workspace, T, heads, scale, dp);
  TVAE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
