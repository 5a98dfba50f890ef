// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[m, tap, n] = sum_{pixel} P[pixel, m] * Q[pixel (+) tap, n]          (reduction over N*H*W pixels)
//
// P is the "dense" operand over the pixel grid (dY for Conv2d, x for ConvTranspose2d); Q is the operand that
// is tap-shifted (3x3 / 1x1, zero padded through TMA out-of-bounds fill) or tap-strided (2x2 stride 2, one
// tensor-map view per tap). Both operands are NHWC bf16, i.e. the reduction dimension (pixels) is the slow
// one: they are fed to the tensor core as MN-major SWIZZLE_128B operands ({64 ch, 64 pixels} TMA boxes,
// LBO = 8 KB between 64-channel chunks, SBO = 1 KB between 8-pixel groups).
// Split-K over pixel blocks writes fp32 partials [split][m][tap][n]; tvae_wgrad_reduce sums the splits in a
// fixed order (deterministic) and scatters into the parameter's own layout ([m][n][tap]: OIHW for Conv2d,
// [Cin][Cout][kH][kW] for ConvTranspose2d).
//
// Replaces autograd's convolution_backward (weight part) for src/model.py:21-42 call sites.
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {
namespace {

constexpr int BM = 128;
constexpr int BKP = 64;                       // pixels per K-block
constexpr int CHUNK_BYTES = 64 * BKP * 2;     // one {64 ch, 64 px} box = 8 KB
constexpr int A_STAGE_BYTES = 2 * CHUNK_BYTES;
constexpr int B_STAGE_BYTES = 4 * CHUNK_BYTES;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // 48 KB
constexpr int STAGES = 4;
// CTA-pair variant (cta_group::2, M = 256): each CTA stages its own 128 P channels and HALF of the Q tile
constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES / 2;   // 32 KB
constexpr int PAIR_STAGES = 6;
static_assert(PAIR_STAGES * PAIR_STAGE_BYTES == STAGES * STAGE_BYTES, "same shared-memory budget");
constexpr int MAX_REM_SPLITS = 16;            // split-K bound of the odd last M tile (runs as its own launch)
constexpr int NTHREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct WgradMaps {
  CUtensorMap p;
  CUtensorMap q[4];
};

struct WgradParams {
  int m_tiles, n_tiles, bn, ntaps, splits;
  int nblocks;                 // pixel blocks in total
  int tiles_w, tiles_h, bw, bh, bnimg;
  int dh[9], dw[9], qmap[9];
  int cm, cn;                  // valid channels of P / Q
  int nq;                      // 64-channel boxes of Q per stage (per CTA)
  float* part;                 // [splits][part_rows][ntaps * cn_pitch], row 0 = GEMM row part_row0
  int part_rows, part_row0;
  int cn_pitch;
  int mt0;                     // first M tile of this launch (m_tiles counts the tiles of this launch)
};

// PAIR: clusters of two CTAs, see conv_gemm.cu for the barrier topology (full/tmem-empty in the leader, empty/tmem-full
// per CTA by multicast commit). The pair owns two adjacent M tiles (P channels) of one (N tile, tap, split) unit.
template <bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_gemm_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradParams p) {
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.p);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.q[i]);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 8 : 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int mgroups = PAIR ? p.m_tiles / 2 : p.m_tiles;          // host launches an even tile count for pairs
  const int units_per_split = mgroups * p.n_tiles * p.ntaps;
  const int total_units = units_per_split * p.splits;
  auto m_tile_of = [&](int mg) { return p.mt0 + (PAIR ? 2 * mg + (int)rank : mg); };

  if (warp == 0) {
    {   // all lanes loop (warp-uniform control flow => uniform registers), one elected lane issues the TMA loads
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (PAIR ? 2u : 1u) * (uint32_t)(2 + p.nq) * CHUNK_BYTES;
      for (int u = unit0; u < total_units; u += nunits) {
        const int split = u / units_per_split;
        int r = u - split * units_per_split;
        const int tap = r % p.ntaps; r /= p.ntaps;
        const int nt = r % p.n_tiles;
        const int mt = m_tile_of(r / p.n_tiles);
        const int qch0 = nt * p.bn + (PAIR ? (int)rank * (p.bn >> 1) : 0);
        const int kb0 = (int)((long long)p.nblocks * split / p.splits);
        const int kb1 = (int)((long long)p.nblocks * (split + 1) / p.splits);
        const CUtensorMap* qm = &maps.q[p.qmap[tap]];
        // pixel-block coordinates advance incrementally (no div/mod on the producer's critical path)
        int iw = kb0 % p.tiles_w, ih = (kb0 / p.tiles_w) % p.tiles_h, in = kb0 / (p.tiles_w * p.tiles_h);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in * p.bnimg;
          if (++iw == p.tiles_w) {
            iw = 0;
            if (++ih == p.tiles_h) { ih = 0; ++in; }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1, 11);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            if (PAIR) {
              const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d_pair(&maps.p, fb, sa, mt * BM, w0, h0, n0);
              tma_load_4d_pair(&maps.p, fb, sa + CHUNK_BYTES, mt * BM + 64, w0, h0, n0);
              for (int j = 0; j < p.nq; ++j)
                tma_load_4d_pair(qm, fb, sb + j * CHUNK_BYTES, qch0 + j * 64, w0 + p.dw[tap], h0 + p.dh[tap], n0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d(&maps.p, &full_bar[stage], sa, mt * BM, w0, h0, n0);
              tma_load_4d(&maps.p, &full_bar[stage], sa + CHUNK_BYTES, mt * BM + 64, w0, h0, n0);
              for (int j = 0; j < p.nq; ++j)
                tma_load_4d(qm, &full_bar[stage], sb + j * CHUNK_BYTES, qch0 + j * 64, w0 + p.dw[tap],
                            h0 + p.dh[tap], n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {   // all lanes wait and count, one elected lane issues the MMAs and their commits (see conv_gemm.cu)
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, p.bn, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int u = unit0; u < total_units; u += nunits) {
        const int split = u / units_per_split;
        const int kb0 = (int)((long long)p.nblocks * split / p.splits);
        const int kb1 = (int)((long long)p.nblocks * (split + 1) / p.splits);
        if (kb1 <= kb0) continue;  // empty split: the epilogue writes zeros without touching TMEM
        mbar_wait(&tempty_bar[as], aphase ^ 1, 12);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, 13);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa, CHUNK_BYTES, 1024);
          const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES, CHUNK_BYTES, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BKP / 16; ++k) {
              // 16 pixels = two 8-row swizzle atoms = 2048 B further along K
              if (PAIR) umma_bf16_pair(d_tmem, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (kb > kb0) || (k > 0));
              else umma_bf16(d_tmem, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (kb > kb0) || (k > 0));
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit_pair(&tfull_bar[as]);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    const long long row_pitch = (long long)p.ntaps * p.cn_pitch;
    for (int u = unit0; u < total_units; u += nunits) {
      const int split = u / units_per_split;
      int r = u - split * units_per_split;
      const int tap = r % p.ntaps; r /= p.ntaps;
      const int nt = r % p.n_tiles;
      const int mt = m_tile_of(r / p.n_tiles);
      const int kb0 = (int)((long long)p.nblocks * split / p.splits);
      const int kb1 = (int)((long long)p.nblocks * (split + 1) / p.splits);
      const int m = mt * BM + row;
      float* orow = p.part + ((long long)split * p.part_rows + (m - p.part_row0)) * row_pitch + (long long)tap * p.cn_pitch;
      if (kb1 > kb0) {
        mbar_wait(&tfull_bar[as], aphase, 14);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
        for (int c = 0; c < p.bn; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c, v);
          tmem_ld_wait();
          const int col = nt * p.bn + c;
          if (m < p.cm && col < p.cn) {
            if (col + 16 <= p.cn) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(orow + col + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
            } else {
              for (int j = 0; j < 16; ++j)
                if (col + j < p.cn) orow[col + j] = __uint_as_float(v[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
          else mbar_arrive(&tempty_bar[as]);
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      } else {
        // empty split (more splits than pixel blocks): contribute zeros
        for (int c = 0; c < p.bn; ++c) {
          const int col = nt * p.bn + c;
          if (m < p.cm && col < p.cn) orow[col] = 0.f;
        }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int g_wgrad_cta_pair = 1;

// grad[m][n][tap] (+)= sum_s part[s][m - row0][tap][n] for the GEMM rows m in [row0, row0 + nrows); with `swapped`
// the GEMM ran with the operand roles exchanged (GEMM row = the parameter's second index): grad[n][m][tap].
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ grad, int cm, int cn,
                                    int ntaps, int cn_pitch, int splits, int accumulate, int swapped, int row0,
                                    int nrows, long long split_stride, int ld, int off) {
  const long long total = (long long)nrows * cn * ntaps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % ntaps);
    const long long rc = i / ntaps;
    int ml, n;
    long long dst;
    if (!swapped) {            // parameter [cm][cn][tap]: consecutive threads write consecutive elements
      n = (int)(rc % cn);
      ml = (int)(rc / cn);
      dst = ((long long)(row0 + ml) * ld + off + n) * ntaps + tap;
    } else {                   // parameter [cn][cm][tap]
      ml = (int)(rc % nrows);
      n = (int)(rc / nrows);
      dst = ((long long)n * ld + off + row0 + ml) * ntaps + tap;
    }
    const float* src = part + ((long long)ml * ntaps + tap) * cn_pitch + n;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += src[k * split_stride];
    grad[dst] = accumulate ? grad[dst] + s : s;
  }
}

// split-K factor that fills `slots` (SMs, or CTA pairs) best with `base` units per split
int pick_splits(int base, int slots, long long nblocks, int max_splits) {
  int best = 1;
  double best_eff = 0;
  for (int s = 1; s <= max_splits; ++s) {
    if (s > 1 && nblocks / s < 16) break;
    const long long units = (long long)base * s;
    const double eff = (double)units / (double)(((units + slots - 1) / slots) * slots);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return best;
}

int wgrad_bn(int cn) {
  return cn <= 256 ? (cn + 15) / 16 * 16 : ((cn + (cn + 255) / 256 - 1) / ((cn + 255) / 256) + 15) / 16 * 16;
}

}  // namespace

}  // namespace tvae

using namespace tvae;

// Workspace: [splits][cm][ntaps * cn_pitch] for the main launch plus [MAX_REM_SPLITS][128][ntaps * cn_pitch] for the odd
// last M tile, which runs as a second launch with its own split-K factor when the main launch uses CTA pairs.
extern "C" int64_t tvae_wgrad_workspace_bytes(int32_t cm, int32_t cn, int32_t ntaps, int32_t splits) {
  const int64_t cn_pitch = (cn + 3) / 4 * 4;
  return ((int64_t)splits * cm + (int64_t)MAX_REM_SPLITS * BM) * ntaps * cn_pitch * 4;
}

extern "C" int32_t tvae_wgrad_splits(int32_t cm, int32_t cn, int32_t ntaps, int64_t pixels) {
  const int bn = wgrad_bn(cn);
  const int m_tiles = (cm + BM - 1) / BM, n_tiles = (cn + bn - 1) / bn;
  const long long nblocks = (pixels + BKP - 1) / BKP;
  // tiny GEMMs (the 1x1 convolutions of the 16x16 level: ONE unit per split) may cut the pixel range into up to 64
  // pieces, or 15 CTAs do all the work while 133 SMs idle (46 TFLOP/s, 48 us per call measured)
  if (g_wgrad_cta_pair && m_tiles >= 2) {   // main launch: pairs of M tiles on pairs of SMs
    const int base = (m_tiles / 2) * n_tiles * ntaps;
    return pick_splits(base, num_sms() / 2, nblocks, base <= 4 ? 64 : 16);
  }
  const int base = m_tiles * n_tiles * ntaps;
  return pick_splits(base, num_sms(), nblocks, base <= 8 ? 64 : 16);
}

extern "C" int32_t tvae_wgrad_set_cta_pair(int32_t enable) {
  const int prev = g_wgrad_cta_pair;
  g_wgrad_cta_pair = enable ? 1 : 0;
  return prev;
}

extern "C" int32_t tvae_wgrad_gemm(const tvae_wgrad_args* a, cudaStream_t stream) {
  TVAE_ENTER(a ? a->p : nullptr);
  TVAE_CHECK(a && a->p && a->q && a->grad && a->workspace, "tvae_wgrad_gemm: null pointer");
  TVAE_CHECK(a->kind >= 0 && a->kind <= 2, "tvae_wgrad_gemm: bad kind");
  TVAE_CHECK(a->p_pitch % 8 == 0 && a->q_pitch % 8 == 0, "tvae_wgrad_gemm: pitches must be multiples of 8");
  TVAE_CHECK(a->splits >= 1, "tvae_wgrad_gemm: splits must be >= 1");
  TVAE_CHECK(!a->flip || a->kind == 0, "tvae_wgrad_gemm: flip (exchanged operand roles) is for stride-1 convs only");
  TVAE_CHECK(a->grad_ld == 0 || (a->grad_off >= 0 && a->grad_off + (a->flip ? a->Cm : a->Cn) <= a->grad_ld),
             "tvae_wgrad_gemm: grad_off + channels exceeds grad_ld");

  WgradMaps maps;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  // pixel grid = grid of the dense operand P: [N, gH, gW]
  const int gH = a->H, gW = a->W;
  TVAE_CHECK(pixel_box(gH, gW, BKP, &p.bw, &p.bh, &p.bnimg), "tvae_wgrad_gemm: unsupported spatial size %dx%d", gH, gW);
  p.tiles_w = gW / p.bw;
  p.tiles_h = gH / p.bh;
  // a pixel block is {bw, bh, bnimg}: blocks tile every image exactly, images are grouped by bnimg
  p.nblocks = ((a->N + p.bnimg - 1) / p.bnimg) * p.tiles_h * p.tiles_w;

  p.cm = a->Cm; p.cn = a->Cn;
  p.m_tiles = (a->Cm + BM - 1) / BM;
  p.bn = wgrad_bn(a->Cn);
  p.n_tiles = (a->Cn + p.bn - 1) / p.bn;
  p.cn_pitch = (a->Cn + 3) / 4 * 4;

  const uint64_t pp = (uint64_t)a->p_pitch * 2, qp = (uint64_t)a->q_pitch * 2;
  uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bnimg};
  {
    uint64_t dims[4] = {(uint64_t)a->Cm, (uint64_t)gW, (uint64_t)gH, (uint64_t)a->N};
    uint64_t strides[3] = {pp, pp * gW, pp * gW * gH};
    if (make_tmap_bf16(&maps.p, a->p, 4, dims, strides, box)) return -3;
  }
  if (a->kind == 0) {
    TVAE_CHECK(a->R == 1 || a->R == 3, "tvae_wgrad_gemm: R must be 1 or 3");
    p.ntaps = a->R * a->R;
    for (int t = 0; t < p.ntaps; ++t) {
      p.dh[t] = t / a->R - a->R / 2; p.dw[t] = t % a->R - a->R / 2; p.qmap[t] = 0;
      if (a->flip) { p.dh[t] = -p.dh[t]; p.dw[t] = -p.dw[t]; }
    }
    uint64_t dims[4] = {(uint64_t)a->Cn, (uint64_t)gW, (uint64_t)gH, (uint64_t)a->N};
    uint64_t strides[3] = {qp, qp * gW, qp * gW * gH};
    for (int i = 0; i < 4; ++i)
      if (make_tmap_bf16(&maps.q[i], a->q, 4, dims, strides, box)) return -3;
  } else {
    // Q lives on the 2x finer grid [N, 2gH, 2gW]; tap (ty, tx) samples pixel (2h+ty, 2w+tx)
    p.ntaps = 4;
    const int qW = 2 * gW, qH = 2 * gH;
    for (int t = 0; t < 4; ++t) {
      p.dh[t] = p.dw[t] = 0; p.qmap[t] = t;
      const int ty = t >> 1, tx = t & 1;
      const uint8_t* base = reinterpret_cast<const uint8_t*>(a->q) + ((size_t)ty * qW + tx) * qp;
      uint64_t dims[4] = {(uint64_t)a->Cn, (uint64_t)gW, (uint64_t)gH, (uint64_t)a->N};
      uint64_t strides[3] = {2 * qp, 2 * qp * qW, qp * qW * qH};
      if (make_tmap_bf16(&maps.q[t], base, 4, dims, strides, box)) return -3;
    }
  }

  const int m_tiles = p.m_tiles;
  const bool pair = g_wgrad_cta_pair != 0 && m_tiles >= 2;
  const int main_tiles = pair ? (m_tiles & ~1) : m_tiles;      // the odd last tile becomes its own single-CTA launch
  const int rem_tiles = m_tiles - main_tiles;
  const long long row_elems = (long long)p.ntaps * p.cn_pitch;
  auto reduce = [&](const float* part, int splits, int row0, int nrows, long long split_stride) {
    const long long total_out = (long long)nrows * a->Cn * p.ntaps;
    int rgrid = (int)((total_out + 255) / 256);
    if (rgrid > 148 * 16) rgrid = 148 * 16;
    const int ld = a->grad_ld > 0 ? a->grad_ld : (a->flip ? a->Cm : a->Cn);
    wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(part, a->grad, a->Cm, a->Cn, p.ntaps, p.cn_pitch, splits,
                                                  a->accumulate, a->flip, row0, nrows, split_stride, ld, a->grad_off);
  };

  // ---- main launch
  p.mt0 = 0;
  p.m_tiles = main_tiles;
  p.splits = a->splits;
  p.part = a->workspace;
  p.part_rows = a->Cm;
  p.part_row0 = 0;
  if (pair) {
    p.nq = (p.bn / 2 + 63) / 64;
    static PerDeviceOnce attr_set;
    if (attr_set.pending()) {
      TVAE_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.mark();
    }
    const int total = (main_tiles / 2) * p.n_tiles * p.ntaps * p.splits;
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
    TVAE_CUDA(cudaLaunchKernelEx(&cfg, wgrad_gemm_kernel<true>, maps, p));
  } else {
    p.nq = (p.bn + 63) / 64;
    static PerDeviceOnce attr_set;
    if (attr_set.pending()) {
      TVAE_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.mark();
    }
    const int total = main_tiles * p.n_tiles * p.ntaps * p.splits;
    int grid = num_sms();
    if (grid > total) grid = total;
    wgrad_gemm_kernel<false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(maps, p);
  }
  TVAE_CUDA(cudaGetLastError());
  {
    const int nrows = main_tiles * BM < a->Cm ? main_tiles * BM : a->Cm;
    reduce(p.part, p.splits, 0, nrows, (long long)a->Cm * row_elems);
  }
  TVAE_CUDA(cudaGetLastError());

  // ---- odd last M tile: one CTA per SM, its own split-K factor and its own workspace region
  if (rem_tiles) {
    static PerDeviceOnce attr_set;
    if (attr_set.pending()) {
      TVAE_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set.mark();
    }
    p.mt0 = main_tiles;
    p.m_tiles = 1;
    p.nq = (p.bn + 63) / 64;
    p.splits = pick_splits(p.n_tiles * p.ntaps, num_sms(), p.nblocks, MAX_REM_SPLITS);
    p.part = a->workspace + (long long)a->splits * a->Cm * row_elems;
    p.part_rows = BM;
    p.part_row0 = main_tiles * BM;
    const int total = p.n_tiles * p.ntaps * p.splits;
    int grid = num_sms();
    if (grid > total) grid = total;
    wgrad_gemm_kernel<false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(maps, p);
    TVAE_CUDA(cudaGetLastError());
    reduce(p.part, p.splits, main_tiles * BM, a->Cm - main_tiles * BM, (long long)BM * row_elems);
  }
  TVAE_CUDA(cudaGetLastError());
  return 0;
}
