// Host-side internals shared by the translation units of libtvae_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
#include "../../include/tvae.h"

namespace tvae {

void set_error(const char* fmt, ...);
int num_sms();
// bind a CUDA context to the calling thread if it has none (device = owner of device_ptr); 0 on success
int enter(const void* device_ptr);
#define TVAE_ENTER(ptr)                      \
  do {                                       \
    if (tvae::enter(ptr) != 0) return -4;    \
  } while (0)

// Kernel attributes (e.g. the dynamic shared-memory limit) are PER DEVICE: a process that drives several GPUs must set
// them once on each (a plain `static bool` set on device 0 leaves device 1 with the 48 KB default and the launch fails
// with "invalid argument"). Setting an attribute twice is harmless, so no lock: check, set, then mark.
struct PerDeviceOnce {
  bool done[64] = {};
  bool pending() const {
    int d = 0;
    return cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64 || !done[d];
  }
  void mark() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) done[d] = true;
  }
};

// Encode a bf16 tiled tensor map with 128-byte swizzle and zero out-of-bounds fill.
// dims/box are innermost-first; strides (bytes) are for dims 1..rank-1. Returns 0 on success.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

// Pixel box {bw, bh, bn} that covers `rows` consecutive pixels (raster order) of an [N, H, W] grid.
bool pixel_box(int H, int W, int rows, int* bw, int* bh, int* bn);

// bandwidth-tuned GroupNorm paths (gn_fast.cu)
bool gn_fast_ok(int C, int G);
long long gn_bwd_fast_ws_floats(int N, int HW, int C, int G);
int gn_act_fwd_fast(const void* x, bool x_bf16, const float* stats, const float* gamma, const float* beta, int N,
                    int HW, int C, int G, int act, __nv_bfloat16* out, __nv_bfloat16* gp_out, cudaStream_t stream);
int gn_act_bwd_fast(const void* x, bool x_bf16, const float* stats, const float* gamma, const float* beta, const __nv_bfloat16* da,
                    const __nv_bfloat16* gres, const __nv_bfloat16* gp, int N, int HW, int C, int G, int act,
                    __nv_bfloat16* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws, cudaStream_t stream);

// tcgen05 (kind::tf32) attention for head dimension 32 (attention_sm100.cu); arguments as tvae_attn_fwd_tc / _bwd_tc
int attn_fwd_sm100(const float* q, const float* k, const float* v, int pitch, int B, int T, int heads, void* out_bf16,
                   float* out_f32, float* lse, cudaStream_t stream);
int attn_bwd_sm100(const float* q, const float* k, const float* v, int pitch, const float* o, const float* d_out,
                   const float* lse, int B, int T, int heads, void* dqkv_bf16, float* workspace, cudaStream_t stream);

// A/B switch of the single-pass GroupNorm backward (gn_fast.cu); group_mb <= 0 keeps the current group size
void gn_set_bwd_fused(int on, int group_mb);

}  // namespace tvae
