// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// and the C-ABI error plumbing. Everything here is inline PTX for sm_100a; there is no other target.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#ifndef TVAE_WAIT_TIMEOUT_CYCLES
// Bounded spins: a protocol bug must trap (and surface as a CUDA error), never hang the GPU.
#define TVAE_WAIT_TIMEOUT_CYCLES (4000000000LL)
#endif

namespace tvae {

// ----------------------------------------------------------------------------------------------
// error handling (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define TVAE_CHECK(cond, ...)                                                        \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      tvae::set_error(__VA_ARGS__);                                                  \
      return -1;                                                                     \
    }                                                                                \
  } while (0)
#define TVAE_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      tvae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

// ----------------------------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or
// the hint expires) instead of spinning. A spinning warp is not free: it shares an SM sub-partition scheduler with
// the TMA-producer / MMA-issuer warps and starves them of issue slots (ncu: 74 % of all samples of the first wgrad
// kernel sat on the spin branch of the idle epilogue warps).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Wait for the phase with the given parity to complete. Traps after a (very long) timeout so a
// broken pipeline becomes a reported launch failure instead of a hung box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TVAE_WAIT_TIMEOUT_CYCLES) {
      printf("tvae: mbarrier wait timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the allocated TMEM base address to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pair (cluster of 2, tcgen05 cta_group::2): one MMA spans both SMs (M = 256), each CTA stages its own 128 rows
// of A and HALF of the B tile, so the per-SM operand traffic from L2 drops by a third for a 128 x 256 x 64 block.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the bytes are counted on the barrier at
// `bar_cluster_addr` (the leader CTA's "full" barrier).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar_cluster_addr, void* smem, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint32_t bar_cluster_addr, void* smem, int c0,
                                                 int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
// one warp of EACH CTA of the pair executes these
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// issued by one thread of the LEADER CTA only; the descriptors address the same offsets in both CTAs' shared memory
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per thread per instruction. The GEMM
// epilogues own one output ROW per thread, so 128-bit stores would fill every sector in two half-sector requests.
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[16], int o) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[o]), "f"(v[o + 1]), "f"(v[o + 2]),
               "f"(v[o + 3]), "f"(v[o + 4]), "f"(v[o + 5]), "f"(v[o + 6]), "f"(v[o + 7])
               : "memory");
}
__device__ __forceinline__ void st_global_v8_b32(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float (&v)[16], int o) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[o]), "=f"(v[o + 1]), "=f"(v[o + 2]), "=f"(v[o + 3]), "=f"(v[o + 4]), "=f"(v[o + 5]),
                 "=f"(v[o + 6]), "=f"(v[o + 7])
               : "l"(p));
}

// Shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulate (InstrDescriptor in the same header)
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // c_format = F32
  d |= 1u << 7;                          // a_format = BF16
  d |= 1u << 10;                         // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15; // a_major
  d |= (uint32_t)(b_mn_major & 1) << 16; // b_major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ----------------------------------------------------------------------------------------------
// reductions
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum, result valid in thread 0. `red` must hold >= 32 entries. Deterministic.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  T r = T(0);
  if (w == 0) {
    r = (l < nw) ? red[l] : T(0);
    r = warp_sum(r);
  }
  return r;
}

// exact-erf GELU (nn.GELU default) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast exact-GELU pieces: the normal CDF Phi(y) from Abramowitz-Stegun 26.2.17 (|error| < 7.5e-8, two orders below
// fp32 activations' own rounding here), sharing ONE exponential e = exp(-y^2/2) with the normal density phi(y) that
// the derivative needs. MUFU approximations (rcp.approx, ex2.approx: ~1-2 ulp) instead of the IEEE-rounded
// __frcp_rn / range-checked __expf: 2 MUFU + ~14 FP32 instructions per element, where the first version spent 37
// (10 of them on the reciprocal's slow-path check alone) and made the GroupNorm backward passes ALU-bound.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// q = 1 - Phi(|y|); pe = phi(y) = exp(-y^2/2) / sqrt(2 pi)
__device__ __forceinline__ void gelu_q(float y, float& q, float& pe) {
  const float t = rcp_approx(fmaf(0.2316419f, fabsf(y), 1.0f));
  pe = 0.39894228040143267794f * ex2_approx(y * y * -0.72134752044448170368f);
  float poly = fmaf(t, 1.330274429f, -1.821255978f);
  poly = fmaf(t, poly, 1.781477937f);
  poly = fmaf(t, poly, -0.356563782f);
  poly = fmaf(t, poly, 0.319381530f);
  q = poly * t * pe;
}
// Phi = Phi(y); pe = phi(y)
__device__ __forceinline__ void gelu_phi(float y, float& Phi, float& pe) {
  float q;
  gelu_q(y, q, pe);
  Phi = (y >= 0.f) ? 1.0f - q : q;
}
// GELU(y) = y Phi(y) = relu(y) - |y| (1 - Phi(|y|)): one FMNMX and one FFMA after q, instead of the compare / subtract /
// select that forms Phi followed by a multiply (16 -> 14 instructions per element in the GroupNorm forward, which inside
// the power-capped step is instruction-bound)
__device__ __forceinline__ float gelu_from_q(float y, float q) { return fmaf(-fabsf(y), q, fmaxf(y, 0.f)); }
__device__ __forceinline__ float gelu_fast(float y) {
  float q, pe;
  gelu_q(y, q, pe);
  return gelu_from_q(y, q);
}
__device__ __forceinline__ float gelu_grad_fast(float y) {
  float Phi, pe;
  gelu_phi(y, Phi, pe);
  return fmaf(y, pe, Phi);
}

// activation codes of the C ABI: 0 identity, 1 GELU (exact erf), 2 ReLU, 3 SiLU  (src/model.py:333-339)
__device__ __forceinline__ float act_f(float y, int act) {
  switch (act) {
    case 1: return gelu_f(y);
    case 2: return fmaxf(y, 0.f);
    case 3: return y / (1.0f + __expf(-y));
    default: return y;
  }
}
__device__ __forceinline__ float act_grad_f(float y, int act) {
  switch (act) {
    case 1: return gelu_grad_f(y);
    case 2: return y > 0.f ? 1.f : 0.f;
    case 3: {
      const float s = 1.0f / (1.0f + __expf(-y));
      return s * (1.0f + y * (1.0f - s));
    }
    default: return 1.f;
  }
}

__device__ __forceinline__ float bf16_bits_to_f(uint32_t hi16) { return __uint_as_float(hi16 << 16); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace tvae
