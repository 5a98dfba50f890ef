// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + shift(tap), cin] * Wp[cout, tap, cin]  (+ bias, + residual)
//
// * A (activations, NHWC bf16) is never im2col'ed: every K-block is ONE 4-D TMA box
//   {64 ch, bw, bh, bn} of the activation tensor at the tap-shifted coordinate; out-of-image taps are
//   produced by TMA's out-of-bounds zero fill, so the 3x3 halo costs nothing. Stride-2 2x2 convs use one
//   strided tensor-map view per tap. The box lands in shared memory as 128 rows x 128 B with the 128-byte
//   swizzle, which is exactly the canonical K-major UMMA operand layout.
// * B (packed weights [rows][taps * c_pad] bf16, K-major) is a 2-D TMA box {64, BN}.
// * One elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) into a double-buffered fp32 TMEM
//   accumulator (2 x 256 columns); an 8-warp epilogue drains it with tcgen05.ld while the next tile's MMAs run.
// * Persistent: grid = #SMs, static round-robin over (m_tile, n_tile) with n fastest so the CTAs that share
//   an A tile run at the same time and hit L2.
//
// Replaces the nn.Conv2d / nn.ConvTranspose2d call sites of the reference:
//   src/model.py:21-42 (get_conv), :240-247 (2x2 s2 down), :270-278 (2x2 s2 transposed up),
//   and their autograd dgrad (same kernel, transposed/flipped weight pack).
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_STAGE_BYTES = 256 * BK * 2;      // 32 KB (BN <= 256)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int STAGES = 4;
// CTA-pair variant: every CTA stages its 128 A rows and half of the B tile => 32 KB per stage, 6 stages
constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES / 2;
constexpr int PAIR_STAGES = 6;
static_assert(PAIR_STAGES * PAIR_STAGE_BYTES == STAGES * STAGE_BYTES, "both variants use the same shared-memory budget");
constexpr int NTHREADS = 320;                    // warp0: TMA, warp1: MMA + TMEM alloc, warps 2-9: epilogue
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*stats*/ + 2048 /*bias*/;

struct ConvMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  // "fp32 mode" (split-bf16 emulation): low-order halves of the operands, x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
  CUtensorMap alo[4];
  CUtensorMap blo;
};

struct ConvParams {
  int m_tiles, n_tiles, bn;
  int num_kb, cblks, ntaps, last_k16;
  int nseg;   // 1: bf16 operands; 3: split-bf16 products hi*hi + hi*lo + lo*hi into the same fp32 accumulator
  int tiles_w, tiles_h, bw, bh, bnimg;
  int dh[9], dw[9], amap[9];
  // epilogue
  int m_total, n_valid;
  float* out_f32;
  int ld_f32;
  __nv_bfloat16* out_bf16;
  __nv_bfloat16* out_bf16_lo;
  int ld_bf16;
  const float* bias;
  const float* res;
  int ld_res;
  int up_mode, up_H, up_W, cout_per_tap;
  // fused GroupNorm statistics of the output: per-tile partial (sum, sum of squares) per group
  float* stats_part;   // [slots][G][2], slot = m_tile (x4 + tap for the transposed conv); NULL = off
  int gs, G;           // group size (multiple of 16, divides BN), number of groups
  int split_chunk;     // first 16-column chunk handled by the second warp of each TMEM lane quarter
  // fused reconstruction loss (decoder.conv_out under get_loss): the output never reaches HBM as fp32; see tvae_conv_args
  const __nv_bfloat16* nll_x;
  int nll_x_pitch, nll_l2, nll_batch;
  const float* nll_logvar;
  float* nll_part;     // [m_tile][n_tile][8 epilogue warps][2] = (sum of |d| or d^2, sum of d^2)
  int wide;            // bit 0/1/2: fp32 output / bf16 output / residual rows are 32-byte aligned => 256-bit accesses
  // optional per-tile timeline (tools/conv_trace.py): trace[(unit * trace_cap + tile) * 8 + j], SM clock cycles:
  // 0 MMA warp before the accumulator-free wait, 1 after it, 2 after the first operand stage arrived, 3 after the last
  // K block was issued, 4 cycles spent waiting for the other operand stages, 5 epilogue: accumulator ready, 6 epilogue
  // done, 7 producer: first load of the tile issued
  unsigned long long* trace;
  int trace_cap;
};

// PAIR = true: launched as clusters of two CTAs (one TPC). The pair owns two vertically adjacent M tiles and one
// N tile; the leader CTA's MMA thread issues tcgen05.mma.cta_group::2 with M = 256, which reads A rows 0-127 /
// 128-255 and B rows 0-bn/2 / bn/2-bn from the leader's / the peer's shared memory and writes each CTA's 128
// accumulator rows into that CTA's own TMEM. Barrier topology: "full" lives in the leader (both producers' TMA
// bytes are counted there), "empty" and "tmem full" are per CTA and signalled by a multicast commit, "tmem empty"
// lives in the leader and collects the 16 epilogue warps of both CTAs.
// LEAN = true: the epilogue of the data-gradient launches (bf16 output only: no fp32 output, residual, statistics,
// split-bf16 half or pixel-shuffle scatter) with those feature paths compiled out.
template <bool PAIR, bool LEAN, bool TRACE = false>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_gemm_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
  constexpr int STAGES = PAIR ? PAIR_STAGES : tvae::STAGES;
  constexpr int STAGE_BYTES = PAIR ? PAIR_STAGE_BYTES : tvae::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float2* stat_sm = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 256);   // [2 acc stages][4 warps][16 groups]
  float* bias_sm = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256 + 1024);  // [2 acc stages][256 columns]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    if (p.nseg > 1) {
      for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.alo[i]);
      tma_prefetch_desc(&maps.blo);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 16 : 8);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's TMA / arrives must not reach a barrier that is not initialised yet
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  auto m_tile_of = [&](int t) { return PAIR ? 2 * (t / p.n_tiles) + (int)rank : t / p.n_tiles; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the loop and one elected lane issues: with warp-uniform control flow the compiler keeps
    // coordinates and addresses in uniform registers and emits each UTMALDG once. (Under `if (lane == 0)` every
    // uniform-datapath instruction was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop.)
    {
      int stage = 0;
      uint32_t phase = 0;
      // bytes of BOTH CTAs' loads for a pair (counted on the leader's barrier)
      const uint32_t tx_bytes = (PAIR ? 2u : 1u) * A_STAGE_BYTES + (uint32_t)p.bn * (BK * 2);
      for (int t = unit0; t < total_tiles; t += nunits) {
        const int mt = m_tile_of(t), nt = t % p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.bw;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bnimg;   // past the last image for the odd tail: zero fill
        const int brow = nt * p.bn + (PAIR ? (int)rank * (p.bn >> 1) : 0);
        int kb = 0;
        if (TRACE && p.trace && rank == 0 && lane == 0) {
          const int ti = (t - unit0) / nunits;
          if (ti < p.trace_cap) p.trace[((long long)unit0 * p.trace_cap + ti) * 8 + 7] = (unsigned long long)clock64();
        }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int hh = h0 + p.dh[tap], ww = w0 + p.dw[tap];
          for (int cb = 0; cb < p.cblks; ++cb, ++kb) {
            for (int seg = 0; seg < p.nseg; ++seg) {
              const CUtensorMap* am = seg == 2 ? &maps.alo[p.amap[tap]] : &maps.a[p.amap[tap]];
              const CUtensorMap* bm = seg == 1 ? &maps.blo : &maps.b;
              mbar_wait(&empty_bar[stage], phase ^ 1, 1);
              uint8_t* sa = smem + stage * STAGE_BYTES;
              if (elect_one()) {
                if (PAIR) {
                  const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                  if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                  tma_load_4d_pair(am, fb, sa, cb * BK, ww, hh, n0);
                  tma_load_2d_pair(bm, fb, sa + A_STAGE_BYTES, kb * BK, brow);
                } else {
                  mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                  tma_load_4d(am, &full_bar[stage], sa, cb * BK, ww, hh, n0);
                  tma_load_2d(bm, &full_bar[stage], sa + A_STAGE_BYTES, kb * BK, brow);
                }
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Same structure: all lanes wait and count, one elected lane issues the MMAs and their commits. The issuing
    // thread was the bottleneck (ncu source view: ~90 SASS instructions per K block, 655 cycles per 512 cycles of
    // tensor work) as long as the whole loop lived inside a single-lane branch.
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, p.bn, 0, 0);
      // Shared-memory descriptors are advanced incrementally (start address field += stage bytes >> 4) and everything an
      // MMA needs is pinned in registers BEFORE the wait on the operand stage: the ring is usually empty when the warp
      // asks (profiles/conv_trace_r2.md), so every instruction between "stage arrived" and the first UTCHMMA is idle
      // tensor time.
      const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t da = desc0;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = unit0; t < total_tiles; t += nunits) {
        const bool tr = TRACE && p.trace != nullptr;
        long long tr0 = 0, tr1 = 0, tr2 = 0, trw = 0;
        if (tr) tr0 = clock64();
        mbar_wait(&tempty_bar[as], aphase ^ 1, 2);
        tc_fence_after();
        if (tr) tr1 = clock64();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        int kb = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int cb = 0; cb < p.cblks; ++cb, ++kb) {
            for (int seg = 0; seg < p.nseg; ++seg) {
              const uint64_t db = da + (uint64_t)(A_STAGE_BYTES >> 4);
              const int nk = (cb == p.cblks - 1) ? p.last_k16 : 4;
              const bool last = (kb == p.num_kb - 1) && (seg == p.nseg - 1);
              const uint32_t acc0 = (kb | seg) != 0 ? 1u : 0u;
              uint64_t* const fb = &full_bar[stage];
              uint64_t* const eb = &empty_bar[stage];
              asm volatile("" ::"l"(da), "l"(db), "r"(nk), "r"(acc0), "l"(fb), "l"(eb));   // materialise before the wait
              long long trc = 0;
              if (tr) trc = clock64();
              mbar_wait(fb, phase, 3);
              tc_fence_after();
              if (tr) {
                const long long now = clock64();
                if (kb == 0 && seg == 0) tr2 = now; else trw += now - trc;
              }
              if (elect_one()) {
                if (nk == 4) {
                  // +32 B per K=16 step inside the 128-B swizzle row (descriptor address is in 16-B units)
                  if (PAIR) {
                    umma_bf16_pair(d_tmem, da, db, idesc, acc0 != 0);
                    umma_bf16_pair(d_tmem, da + 2, db + 2, idesc, true);
                    umma_bf16_pair(d_tmem, da + 4, db + 4, idesc, true);
                    umma_bf16_pair(d_tmem, da + 6, db + 6, idesc, true);
                  } else {
                    umma_bf16(d_tmem, da, db, idesc, acc0 != 0);
                    umma_bf16(d_tmem, da + 2, db + 2, idesc, true);
                    umma_bf16(d_tmem, da + 4, db + 4, idesc, true);
                    umma_bf16(d_tmem, da + 6, db + 6, idesc, true);
                  }
                } else {
                  for (int k = 0; k < nk; ++k) {
                    if (PAIR) umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (acc0 | (uint32_t)k) != 0);
                    else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (acc0 | (uint32_t)k) != 0);
                  }
                }
                // commits track the MMAs of the issuing thread: same elected lane
                if (PAIR) umma_commit_pair(eb);
                else umma_commit(eb);
                if (last) {
                  if (PAIR) umma_commit_pair(&tfull_bar[as]);
                  else umma_commit(&tfull_bar[as]);
                }
              }
              __syncwarp();
              da += (uint64_t)(STAGE_BYTES >> 4);
              if (++stage == STAGES) { stage = 0; phase ^= 1; da = desc0; }
            }
          }
        }
        if (tr && lane == 0) {
          const int ti = (t - unit0) / nunits;
          if (ti < p.trace_cap) {
            unsigned long long* rec = p.trace + ((long long)unit0 * p.trace_cap + ti) * 8;
            rec[0] = tr0; rec[1] = tr1; rec[2] = tr2; rec[3] = clock64(); rec[4] = trw;
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 128 rows x 2 column halves)
    // TMEM lane quarter q = warp % 4 is a hardware rule; the two warps that share a quarter split the tile's
    // 16-column chunks between them, and each one keeps the NEXT chunk's tcgen05.ld and residual loads in flight
    // while it processes the current one: the epilogue is latency-bound (one in-order warp per scheduler), and for
    // fp32-output layers it used to be as long as the MMA window itself.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int nch = p.bn >> 4;
    const int cbeg = half ? p.split_chunk : 0, cend = half ? nch : p.split_chunk;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = unit0; t < total_tiles; t += nunits) {
      const int mt = m_tile_of(t), nt = t % p.n_tiles;
      int col0 = nt * p.bn;   // column in the (tap, cout) / cout space
      const int tap = (!LEAN && p.up_mode) ? col0 / p.cout_per_tap : 0;
      if (!LEAN && p.up_mode) col0 -= tap * p.cout_per_tap;
      if (p.bias) {
        // The tile's bias slice goes through shared memory, fetched while the MMAs of the tile are still running. (It
        // used to be four __ldg per 16-column chunk issued right before their use: 14 % of all stall samples of the
        // epilogue-bound K = 256 layers sat on that load.) Buffer `as` was last read two tiles ago, and every warp has
        // passed the barrier of the tile in between since.
        const int e = (warp - 2) * 32 + lane;
        bias_sm[as * 256 + e] = (e < p.bn && col0 + e < p.n_valid) ? __ldg(p.bias + col0 + e) : 0.f;
        asm volatile("bar.sync 2, 256;" ::: "memory");   // the 8 epilogue warps only
      }
      const long long pix = (long long)mt * BM + row;
      const bool row_ok = pix < p.m_total;
      long long opix = pix;
      if (!LEAN && p.up_mode) {
        const int hw = p.up_H * p.up_W;
        const int n = (int)(pix / hw);
        const int rem = (int)(pix - (long long)n * hw);
        const int h = rem / p.up_W, w = rem - h * p.up_W;
        opix = ((long long)n * (2 * p.up_H) + (2 * h + (tap >> 1))) * (2 * p.up_W) + (2 * w + (tap & 1));
      }
      mbar_wait(&tfull_bar[as], aphase, 4);
      tc_fence_after();
      const bool etr = TRACE && p.trace != nullptr && rank == 0 && warp == 2 && lane == 0;
      if (etr) {
        const int ti = (t - unit0) / nunits;
        if (ti < p.trace_cap) p.trace[((long long)unit0 * p.trace_cap + ti) * 8 + 5] = (unsigned long long)clock64();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
      const bool do_stats = !LEAN && p.stats_part != nullptr;
      float st1 = 0.f, st2 = 0.f;
      // group bookkeeping without a division per chunk (gs % 16 == 0 and a warp's chunk range starts on a group
      // boundary or inside ONE group -- fused_stats_ok on the host side)
      const int chunks_per_group = do_stats ? (p.gs >> 4) : 1;
      int gslot = do_stats ? (cbeg << 4) / p.gs : 0;
      int gchunk = do_stats ? cbeg - gslot * chunks_per_group : 0;
      const float* res_row = (!LEAN && p.res) ? p.res + opix * p.ld_res : nullptr;
      const bool do_nll = !LEAN && p.nll_x != nullptr;
      const __nv_bfloat16* x_row = do_nll ? p.nll_x + opix * p.nll_x_pitch : nullptr;
      const float ngs = do_nll ? expf(-__ldg(p.nll_logvar)) / (float)p.nll_batch : 0.f;
      const uint32_t ngs_pair = pack_bf16(ngs, ngs);
      float nrec = 0.f, nsq = 0.f;

      uint32_t r[16];
      float rs[16];
      auto load_res = [&](int col) {
        if (p.wide & 4) {
          ld_global_v8(res_row + col, rs, 0);
          ld_global_v8(res_row + col + 8, rs, 8);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 x4 = *reinterpret_cast<const float4*>(res_row + col + 4 * j);
            rs[4 * j] = x4.x; rs[4 * j + 1] = x4.y; rs[4 * j + 2] = x4.z; rs[4 * j + 3] = x4.w;
          }
        }
      };
      uint4 xn0 = make_uint4(0u, 0u, 0u, 0u), xn1 = xn0;
      auto load_target = [&](int col) {
        xn0 = __ldg(reinterpret_cast<const uint4*>(x_row + col));
        xn1 = __ldg(reinterpret_cast<const uint4*>(x_row + col + 8));
      };
      if (cbeg < cend) {                                    // prologue: first chunk in flight
        tmem_ld16(taddr + (uint32_t)(cbeg << 4), r);
        const int col = col0 + (cbeg << 4);
        if (res_row && row_ok && col + 16 <= p.n_valid) load_res(col);
        if (do_nll && row_ok && col + 16 <= p.n_valid) load_target(col);
      }
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c = ch << 4;
        const int col = col0 + c;
        const uint4 xq0 = xn0, xq1 = xn1;      // the loss target's 16 values of this chunk (loaded one chunk ahead)
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        float rv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) rv[j] = rs[j];
        if (ch + 1 < cend) {                                // next chunk: TMEM load + residual loads go out now
          tmem_ld16(taddr + (uint32_t)(c + 16), r);
          const int ncol = col + 16;
          if (res_row && row_ok && ncol + 16 <= p.n_valid) load_res(ncol);
          if (do_nll && row_ok && ncol + 16 <= p.n_valid) load_target(ncol);
        }
        if (row_ok && col < p.n_valid) {
          const bool full = (col + 16 <= p.n_valid);
          if (p.bias) {      // columns >= n_valid hold 0 in bias_sm
            const float4* b4 = reinterpret_cast<const float4*>(bias_sm + as * 256 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 bb = b4[j];
              v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            }
          }
          if (res_row) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += rv[j];
            } else {
              for (int j = 0; j < 16; ++j)
                if (col + j < p.n_valid) v[j] += res_row[col + j];
            }
          }
          if (do_nll) {
            // d = reconstruction - target; the bf16 output is the loss gradient wrt the reconstruction
            // (src/model.py:656-663), written right here. Kept lean on purpose: in the first version this block doubled
            // the epilogue's instruction count (both loss types evaluated and selected per element, a bounds predicate
            // per element, a two-compare sign) and the tensor pipe of decoder.conv_out dropped from 90 % to 83 %.
            __nv_bfloat16* op = p.out_bf16 + opix * p.ld_bf16 + col;
            if (full) {
              const uint32_t xw[8] = {xq0.x, xq0.y, xq0.z, xq0.w, xq1.x, xq1.y, xq1.z, xq1.w};
              uint32_t w8[8];
              if (!p.nll_l2) {
                // l1: the gradient is +-g or 0, g = exp(-logvar) / batch. The packed bf16 pair is |g| with the sign bits
                // of the two differences OR'ed in (one PRMT + one LOP3 per pair, no float->bf16 conversion); an exact
                // zero difference (gradient 0, as torch.sign gives) is patched on a rarely taken path.
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float d0 = v[2 * j] - bf16_bits_to_f(xw[j] & 0xffffu);
                  const float d1 = v[2 * j + 1] - bf16_bits_to_f(xw[j] >> 16);
                  nsq = fmaf(d0, d0, nsq);
                  nsq = fmaf(d1, d1, nsq);
                  nrec += fabsf(d0);
                  nrec += fabsf(d1);
                  uint32_t w = (__byte_perm(__float_as_uint(d0), __float_as_uint(d1), 0x7030) & 0x80008000u) | ngs_pair;
                  if (d0 * d1 == 0.f) {
                    if (d0 == 0.f) w &= 0xffff0000u;
                    if (d1 == 0.f) w &= 0x0000ffffu;
                  }
                  w8[j] = w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float d0 = v[2 * j] - bf16_bits_to_f(xw[j] & 0xffffu);
                  const float d1 = v[2 * j + 1] - bf16_bits_to_f(xw[j] >> 16);
                  const float q0 = d0 * d0, q1 = d1 * d1;
                  nsq += q0; nsq += q1;
                  nrec += q0; nrec += q1;
                  w8[j] = pack_bf16(2.0f * d0 * ngs, 2.0f * d1 * ngs);
                }
              }
              if (p.wide & 2) {
                st_global_v8_b32(op, w8);
              } else {
                *reinterpret_cast<uint4*>(op) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                *reinterpret_cast<uint4*>(op + 8) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
              }
            } else {                          // the ragged last chunk of the channel range (1028 = 64 x 16 + 4)
              for (int j = 0; j < 16; ++j) {
                if (col + j < p.n_valid) {
                  const float d = v[j] - __bfloat162float(x_row[col + j]);
                  const float sq = d * d;
                  nsq += sq;
                  nrec += p.nll_l2 ? sq : fabsf(d);
                  const float gj = p.nll_l2 ? 2.0f * d * ngs : ((d > 0.f) ? ngs : ((d < 0.f) ? -ngs : 0.f));
                  op[j] = __float2bfloat16(gj);
                } else if (col + j < p.ld_bf16) {
                  op[j] = __float2bfloat16(0.f);            // pad lanes of the gradient rows
                }
              }
            }
          }
          if (do_stats) {   // host guarantees full 16-column chunks (Cout % gs == 0, gs % 16 == 0)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              st1 += v[j];
              st2 = fmaf(v[j], v[j], st2);
            }
          }
          if (!LEAN && p.out_f32) {
            float* op = p.out_f32 + opix * p.ld_f32 + col;
            if (full && (p.wide & 1)) {
              st_global_v8(op, v, 0);
              st_global_v8(op + 8, v, 8);
            } else if (full) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int j = 0; j < 16; ++j)
                if (col + j < p.n_valid) op[j] = v[j];
            }
          }
          if (LEAN || (p.out_bf16 && !do_nll)) {
            __nv_bfloat16* op = p.out_bf16 + opix * p.ld_bf16 + col;
            if (full) {
              uint32_t w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
              if (p.wide & 2) {
                st_global_v8_b32(op, w8);
              } else {
                *reinterpret_cast<uint4*>(op) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                *reinterpret_cast<uint4*>(op + 8) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
              }
            } else {
              for (int j = 0; j < 16; ++j)
                if (col + j < p.n_valid) op[j] = __float2bfloat16(v[j]);
            }
            if (!LEAN && p.out_bf16_lo) {     // residual of the bf16 rounding, for split-bf16 consumers
              __nv_bfloat16* ol = p.out_bf16_lo + opix * p.ld_bf16 + col;
              for (int j = 0; j < 16; ++j)
                if (col + j < p.n_valid) ol[j] = __float2bfloat16(v[j] - __bfloat162float(__float2bfloat16(v[j])));
            }
          }
        }
        if (do_stats && ++gchunk == chunks_per_group) {   // a group's columns are complete: reduce over the warp's 32 rows
          const float w1 = warp_sum(st1), w2 = warp_sum(st2);
          if (lane == 0) stat_sm[(as * 4 + q) * 16 + gslot] = make_float2(w1, w2);
          st1 = 0.f; st2 = 0.f;
          gchunk = 0;
          ++gslot;
        }
      }
      if (do_nll) {
        const float w1 = warp_sum(nrec), w2 = warp_sum(nsq);
        if (lane == 0 && mt < p.m_tiles) {
          float* dst = p.nll_part + (((long long)mt * p.n_tiles + nt) * 8 + (warp - 2)) * 2;
          dst[0] = w1; dst[1] = w2;
        }
      }
      if (do_stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
        const int ng = p.bn / p.gs;
        if (half == 0 && row < ng) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {                     // fixed order => deterministic
            const float2 t2 = stat_sm[(as * 4 + w) * 16 + row];
            s1 += t2.x; s2 += t2.y;
          }
          const int gidx = col0 / p.gs + row;
          if (gidx < p.G && mt < p.m_tiles) {
            const int tapslot = p.up_mode ? (nt * p.bn) / p.cout_per_tap : 0;
            const long long slot = p.up_mode ? (long long)mt * 4 + tapslot : mt;
            float* dst = p.stats_part + (slot * p.G + gidx) * 2;
            dst[0] = s1; dst[1] = s2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (etr) {
        const int ti = (t - unit0) / nunits;
        if (ti < p.trace_cap) p.trace[((long long)unit0 * p.trace_cap + ti) * 8 + 6] = (unsigned long long)clock64();
      }
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        else mbar_arrive(&tempty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the leader's barriers and both CTAs' operands stay alive until the pair is done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// On by default. History (B=256 train step, alternating runs on one box): while the MMA-issuing thread was the
// bottleneck (single-lane branch, see the MMA issuer above) the pair schedule was 1 % slower than one CTA per SM;
// with the elected-lane issue loop it is 5 % faster in the step (116.4 vs 123.1 ms) and the bf16-output launch runs
// at 1.64 PFLOP/s (3.01 ms) against 1.1-1.2 PFLOP/s for the single-CTA schedule, which is then limited by the
// 96 B/clk/SM of operand traffic that the pair schedule cuts to 64 B/clk/SM.
int g_conv_cta_pair = 1;
// sums[0] = sum of |d| (l1) or d^2 (l2), sums[1] = sum of d^2, sums[2] = 0: the layout tvae_nll_fwd produces; the
// per-(tile, warp) fp32 partials are added in a fixed order in fp64, in two stages (128 slices, then the slices): one
// block walking all 327,680 pairs of the B=256 step took 60 us on the critical path.
constexpr int NLL_SLICES = 128;
__global__ void __launch_bounds__(256) conv_nll_slice_kernel(const float* __restrict__ part, long long n,
                                                             double* __restrict__ slices) {
  __shared__ double red[32];
  const long long per = (n + NLL_SLICES - 1) / NLL_SLICES;
  const long long i0 = blockIdx.x * per, i1 = min(i0 + per, n);
  double a = 0.0, b = 0.0;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float2 v = *reinterpret_cast<const float2*>(part + 2 * i);
    a += (double)v.x; b += (double)v.y;
  }
  const double t1 = block_sum(a, red);
  const double t2 = block_sum(b, red);
  if (threadIdx.x == 0) { slices[2 * blockIdx.x] = t1; slices[2 * blockIdx.x + 1] = t2; }
}
__global__ void __launch_bounds__(NLL_SLICES) conv_nll_final_kernel(const double* __restrict__ slices,
                                                                    double* __restrict__ sums) {
  __shared__ double red[32];
  const double t1 = block_sum(slices[2 * threadIdx.x], red);
  const double t2 = block_sum(slices[2 * threadIdx.x + 1], red);
  if (threadIdx.x == 0) { sums[0] = t1; sums[1] = t2; sums[2] = 0.0; }
}

int g_conv_lean_epilogue = 1;
unsigned long long* g_conv_trace = nullptr;
int g_conv_trace_cap = 0;

int pick_bn(int cout) {
  if (cout <= 256) return (cout + 15) / 16 * 16;
  const int nt = (cout + 255) / 256;
  const int per = (cout + nt - 1) / nt;
  return (per + 15) / 16 * 16;
}

}  // namespace

}  // namespace tvae

using namespace tvae;

extern "C" int32_t tvae_conv_gemm(const tvae_conv_args* a, cudaStream_t stream) {
  TVAE_ENTER(a ? a->x : nullptr);
  TVAE_CHECK(a != nullptr, "tvae_conv_gemm: null args");
  TVAE_CHECK(a->x && a->w, "tvae_conv_gemm: null x/w");
  TVAE_CHECK(a->out_f32 || a->out_bf16, "tvae_conv_gemm: no output");
  TVAE_CHECK(a->kind >= 0 && a->kind <= 2, "tvae_conv_gemm: bad kind %d", a->kind);
  TVAE_CHECK(a->x_pitch % 8 == 0 && a->k_pitch % 8 == 0, "tvae_conv_gemm: pitches must be multiples of 8 elements");
  TVAE_CHECK((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
             "tvae_conv_gemm: x/w must be 16-byte aligned");
  TVAE_CHECK(a->c_pad % 64 == 0 && a->c_pad >= a->C, "tvae_conv_gemm: c_pad must be a multiple of 64 and >= C");

  ConvMaps maps;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const bool split = a->x_lo != nullptr && a->w_lo != nullptr;   // split-bf16 ("fp32 mode") operands
  TVAE_CHECK((a->x_lo == nullptr) == (a->w_lo == nullptr), "tvae_conv_gemm: x_lo and w_lo must be given together");

  // geometry of the A-operand pixel grid (== output grid for kind 0/1, == input grid for kind 2)
  int gH = a->H, gW = a->W;
  if (a->kind == 1) {
    TVAE_CHECK(a->H % 2 == 0 && a->W % 2 == 0, "2x2 stride-2 conv needs even H, W");
    gH = a->H / 2; gW = a->W / 2;
  }
  TVAE_CHECK(pixel_box(gH, gW, BM, &p.bw, &p.bh, &p.bnimg),
             "tvae_conv_gemm: unsupported spatial size %dx%d (W must be a power of two < 128 or a multiple of 128)",
             gH, gW);
  p.tiles_w = gW / p.bw;
  p.tiles_h = gH / p.bh;
  const long long m_total = (long long)a->N * gH * gW;
  TVAE_CHECK(m_total < (1ll << 31), "tvae_conv_gemm: too many pixels");
  p.m_total = (int)m_total;
  p.m_tiles = (int)((m_total + BM - 1) / BM);

  p.cblks = a->c_pad / BK;
  const int c_rem = a->C - (p.cblks - 1) * BK;             // valid channels in the last block
  TVAE_CHECK(c_rem > 0, "tvae_conv_gemm: c_pad too large for C");
  p.last_k16 = (c_rem + 15) / 16;

  const uint64_t pitchB = (uint64_t)a->x_pitch * 2;
  if (a->kind == 0) {
    TVAE_CHECK(a->R == 1 || a->R == 3, "tvae_conv_gemm: R must be 1 or 3");
    p.ntaps = a->R * a->R;
    for (int t = 0; t < p.ntaps; ++t) {
      int dh = t / a->R - a->R / 2, dw = t % a->R - a->R / 2;
      if (a->flip) { dh = -dh; dw = -dw; }
      p.dh[t] = dh; p.dw[t] = dw; p.amap[t] = 0;
    }
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    uint64_t strides[3] = {pitchB, pitchB * a->W, pitchB * a->W * a->H};
    uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bnimg};
    for (int i = 0; i < 4; ++i) {
      if (make_tmap_bf16(&maps.a[i], a->x, 4, dims, strides, box)) return -3;
      if (make_tmap_bf16(&maps.alo[i], split ? a->x_lo : a->x, 4, dims, strides, box)) return -3;
    }
  } else if (a->kind == 1) {
    p.ntaps = 4;
    for (int t = 0; t < 4; ++t) {
      p.dh[t] = 0; p.dw[t] = 0; p.amap[t] = t;
      const int ty = t >> 1, tx = t & 1;
      const uint8_t* base = reinterpret_cast<const uint8_t*>(a->x) + ((size_t)ty * a->W + tx) * pitchB;
      uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)gW, (uint64_t)gH, (uint64_t)a->N};
      uint64_t strides[3] = {2 * pitchB, 2 * pitchB * a->W, pitchB * a->W * a->H};
      uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bnimg};
      if (make_tmap_bf16(&maps.a[t], base, 4, dims, strides, box)) return -3;
      const uint8_t* base_lo =
          reinterpret_cast<const uint8_t*>(split ? a->x_lo : a->x) + ((size_t)ty * a->W + tx) * pitchB;
      if (make_tmap_bf16(&maps.alo[t], base_lo, 4, dims, strides, box)) return -3;
    }
  } else {
    p.ntaps = 1;
    p.dh[0] = p.dw[0] = 0; p.amap[0] = 0;
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    uint64_t strides[3] = {pitchB, pitchB * a->W, pitchB * a->W * a->H};
    uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bnimg};
    for (int i = 0; i < 4; ++i) {
      if (make_tmap_bf16(&maps.a[i], a->x, 4, dims, strides, box)) return -3;
      if (make_tmap_bf16(&maps.alo[i], split ? a->x_lo : a->x, 4, dims, strides, box)) return -3;
    }
  }
  p.nseg = split ? 3 : 1;
  p.num_kb = p.ntaps * p.cblks;
  TVAE_CHECK(a->k_pitch >= p.num_kb * BK, "tvae_conv_gemm: k_pitch %d < taps*c_pad %d", a->k_pitch, p.num_kb * BK);

  // N tiling
  int n_cols;  // total GEMM columns
  if (a->kind == 2) {
    TVAE_CHECK(a->Cout % 16 == 0, "transposed conv needs Cout %% 16 == 0");
    int bn = a->bn;
    if (bn <= 0) {
      bn = 16;
      for (int c = 256; c >= 16; c -= 16)
        if (a->Cout % c == 0) { bn = c; break; }
    }
    TVAE_CHECK(a->Cout % bn == 0, "transposed conv: bn must divide Cout");
    p.bn = bn;
    n_cols = 4 * a->Cout;
    p.up_mode = 1; p.up_H = a->H; p.up_W = a->W; p.cout_per_tap = a->Cout;
  } else {
    p.bn = a->bn > 0 ? a->bn : pick_bn(a->Cout);
    n_cols = a->Cout;
  }
  TVAE_CHECK(p.bn % 16 == 0 && p.bn >= 16 && p.bn <= 256, "tvae_conv_gemm: bad bn %d", p.bn);
  p.n_tiles = (n_cols + p.bn - 1) / p.bn;
  // CTA pairs (cta_group::2) whenever there are at least two M tiles to pair up
  const bool pair = g_conv_cta_pair != 0 && p.m_tiles >= 2;
  p.n_valid = a->Cout;
  TVAE_CHECK(a->w_rows >= n_cols, "tvae_conv_gemm: packed weight has %d rows, need %d", a->w_rows, n_cols);
  {
    uint64_t dims[2] = {(uint64_t)a->k_pitch, (uint64_t)a->w_rows};
    uint64_t strides[1] = {(uint64_t)a->k_pitch * 2};
    uint32_t box[2] = {BK, (uint32_t)(pair ? p.bn / 2 : p.bn)};
    if (make_tmap_bf16(&maps.b, a->w, 2, dims, strides, box)) return -3;
    if (make_tmap_bf16(&maps.blo, split ? a->w_lo : a->w, 2, dims, strides, box)) return -3;
  }

  p.stats_part = a->stats_part;
  p.trace = g_conv_trace;
  p.trace_cap = g_conv_trace_cap;
  if (p.stats_part) {
    p.G = a->stats_groups;
    TVAE_CHECK(p.G > 0 && a->Cout % p.G == 0, "tvae_conv_gemm: stats_groups must divide Cout");
    p.gs = a->Cout / p.G;
    TVAE_CHECK(p.gs % 16 == 0 && p.bn % p.gs == 0 && p.bn / p.gs <= 16,
               "tvae_conv_gemm: fused statistics need a group size that is a multiple of 16 and divides the N tile");
    TVAE_CHECK(p.bnimg == 1 && (long long)gH * gW % BM == 0,
               "tvae_conv_gemm: fused statistics need at least 128 pixels per image");
  }
  {  // column split between the two epilogue warps of a lane quarter (aligned to GroupNorm groups when stats are on)
    const int nch = p.bn / 16;
    int first = (nch + 1) / 2;
    if (p.stats_part) {
      const int gpc = p.gs / 16;
      first = (first + gpc - 1) / gpc * gpc;
    }
    p.split_chunk = first > nch ? nch : first;
  }
  p.out_f32 = a->out_f32; p.ld_f32 = a->out_f32_pitch;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16); p.ld_bf16 = a->out_bf16_pitch;
  p.nll_x = reinterpret_cast<const __nv_bfloat16*>(a->nll_x);
  if (p.nll_x) {
    TVAE_CHECK(a->kind == 0 && !a->flip && p.out_bf16 && !p.out_f32 && !p.res && !p.stats_part && !a->out_bf16_lo && !split,
               "tvae_conv_gemm: the fused reconstruction loss needs a stride-1 forward conv with the bf16 output only");
    TVAE_CHECK(a->nll_logvar && a->nll_workspace && a->nll_sums && a->nll_batch > 0 && (a->nll_loss_type == 0 || a->nll_loss_type == 1),
               "tvae_conv_gemm: incomplete nll_* arguments");
    TVAE_CHECK(a->nll_x_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(a->nll_x) & 15) == 0 && a->nll_x_pitch >= a->Cout,
               "tvae_conv_gemm: nll_x must be 16-byte aligned with a pitch that is a multiple of 8 and >= Cout");
    p.nll_x_pitch = a->nll_x_pitch;
    p.nll_l2 = a->nll_loss_type;
    p.nll_batch = a->nll_batch;
    p.nll_logvar = a->nll_logvar;
    p.nll_part = a->nll_workspace;
  }
  p.out_bf16_lo = reinterpret_cast<__nv_bfloat16*>(a->out_bf16_lo);
  TVAE_CHECK(!p.out_bf16_lo || p.out_bf16, "tvae_conv_gemm: out_bf16_lo needs out_bf16");
  p.bias = a->bias;
  p.res = a->residual; p.ld_res = a->res_pitch;
  if (p.out_f32) TVAE_CHECK(p.ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0, "out_f32 alignment");
  if (p.out_bf16) TVAE_CHECK(p.ld_bf16 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0, "out_bf16 alignment");
  if (p.out_f32 && p.ld_f32 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 31) == 0) p.wide |= 1;
  if (p.out_bf16 && p.ld_bf16 % 16 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 31) == 0) p.wide |= 2;
  if (p.res && p.ld_res % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 31) == 0) p.wide |= 4;
  if (p.bias) TVAE_CHECK((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "bias must be 16-byte aligned");
  if (p.res) TVAE_CHECK(p.ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0, "residual alignment");

  {
    static bool env_read = false;            // TVAE_CONV_LEAN=0 keeps the generic epilogue (A/B measurements)
    if (!env_read) {
      const char* e = getenv("TVAE_CONV_LEAN");
      if (e) g_conv_lean_epilogue = atoi(e) != 0;
      env_read = true;
    }
  }
  const bool lean = g_conv_lean_epilogue && p.out_bf16 && !p.out_f32 && !p.res && !p.stats_part && !p.out_bf16_lo &&
                    !p.up_mode && !p.nll_x;
  if (pair) {
    static PerDeviceOnce attr_set;
    if (attr_set.pending()) {
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.mark();
    }
    const int total = (p.m_tiles + 1) / 2 * p.n_tiles;
    int grid = num_sms() / 2;
    if (grid > total) grid = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // the timeline instantiations exist for the pair schedule only; production launches never carry the clock reads
    if (p.trace && lean) TVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, true, true>, maps, p));
    else if (p.trace) TVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, false, true>, maps, p));
    else if (lean) TVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, true>, maps, p));
    else TVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, false>, maps, p));
  } else {
    static PerDeviceOnce attr_set;
    if (attr_set.pending()) {
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      TVAE_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.mark();
    }
    const int total = p.m_tiles * p.n_tiles;
    int grid = num_sms();
    if (grid > total) grid = total;
    if (lean) conv_gemm_kernel<false, true><<<grid, NTHREADS, SMEM_BYTES, stream>>>(maps, p);
    else conv_gemm_kernel<false, false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(maps, p);
  }
  TVAE_CUDA(cudaGetLastError());
  if (p.nll_x) {
    const long long npairs = (long long)p.m_tiles * p.n_tiles * 8;
    double* slices = reinterpret_cast<double*>(p.nll_part + ((2 * npairs + 3) & ~3ll));   // 16-byte aligned tail
    conv_nll_slice_kernel<<<NLL_SLICES, 256, 0, stream>>>(p.nll_part, npairs, slices);
    conv_nll_final_kernel<<<1, NLL_SLICES, 0, stream>>>(slices, a->nll_sums);
    TVAE_CUDA(cudaGetLastError());
  }
  return 0;
}

// upper bound for any N-tile choice: n_tiles <= ceil(Cout / 16)
extern "C" int64_t tvae_conv_nll_workspace_bytes(int64_t pixels, int32_t Cout) {
  return ((pixels + BM - 1) / BM) * (int64_t)((Cout + 15) / 16) * 8 * 2 * (int64_t)sizeof(float) + 16 +
         2 * NLL_SLICES * (int64_t)sizeof(double);
}

extern "C" int32_t tvae_conv_set_trace(void* device_buffer, int32_t tiles_per_unit) {
  g_conv_trace = reinterpret_cast<unsigned long long*>(device_buffer);
  g_conv_trace_cap = device_buffer ? tiles_per_unit : 0;
  return 0;
}

extern "C" int32_t tvae_conv_set_cta_pair(int32_t enable) {
  const int prev = g_conv_cta_pair;
  g_conv_cta_pair = enable ? 1 : 0;
  return prev;
}
