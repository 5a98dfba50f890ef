// Host utilities: thread-local error string, SM count cache, TMA descriptor encoding.
#include <stdarg.h>
#include <stdio.h>
#include <mutex>
#include "common.cuh"
#include "tvae_internal.h"

namespace tvae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------------------------
// Thread/context binding. PyTorch's autograd worker threads set their CUDA device lazily, so an entry point can be
// called on a thread that has NO current context yet; this library's (statically linked) runtime would then bind
// device 0. Every entry point therefore calls enter(ptr): if the thread has no current context it binds the primary
// context of the device that owns `ptr` (a device pointer argument). It never switches an already-bound thread; if that
// thread's current device is not the one that owns `ptr`, the call is refused with an error.
typedef CUresult (*ctx_get_current_fn)(CUcontext*);
typedef CUresult (*ptr_get_attr_fn)(void*, CUpointer_attribute, CUdeviceptr);

static void* driver_sym(const char* name) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return sym;
}

int enter(const void* device_ptr) {
  static ctx_get_current_fn get_cur = nullptr;
  static ptr_get_attr_fn get_attr = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    get_cur = reinterpret_cast<ctx_get_current_fn>(driver_sym("cuCtxGetCurrent"));
    get_attr = reinterpret_cast<ptr_get_attr_fn>(driver_sym("cuPointerGetAttribute"));
  });
  if (!get_cur || !get_attr) return 0;  // cannot check: rely on the caller's thread state
  CUcontext cur = nullptr;
  if (get_cur(&cur) == CUDA_SUCCESS && cur != nullptr) {
    // Bound thread: the launch goes to the thread's current device. A pointer that lives on ANOTHER device would
    // fault inside the kernel (sticky error, dead process); refuse the call instead. The library still never
    // switches the device -- the caller does (tempo_vae_b200/ops.py: _on_device).
    int ord = -1, dev = -1;
    if (device_ptr != nullptr &&
        get_attr(&ord, CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL, reinterpret_cast<CUdeviceptr>(device_ptr)) == CUDA_SUCCESS &&
        ord >= 0 && cudaGetDevice(&dev) == cudaSuccess && dev != ord) {
      set_error("operand %p lives on CUDA device %d but the calling thread's current device is %d: make the "
                "operand's device current before calling (cudaSetDevice / torch.cuda.device)", device_ptr, ord, dev);
      return -1;
    }
    return 0;
  }
  int ordinal = -1;
  if (device_ptr == nullptr ||
      get_attr(&ordinal, CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL, reinterpret_cast<CUdeviceptr>(device_ptr)) != CUDA_SUCCESS ||
      ordinal < 0) {
    set_error("no CUDA context is current on this thread and the device of pointer %p cannot be determined", device_ptr);
    return -1;
  }
  if (cudaSetDevice(ordinal) != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed while binding a context to the calling thread", ordinal);
    return -1;
  }
  return 0;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(sym);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  encode_tiled_fn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return -1;
  }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] stride0=%llu base=%p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
    return -1;
  }
  return 0;
}

bool pixel_box(int H, int W, int rows, int* bw, int* bh, int* bn) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (W >= rows) {
    if (W % rows) return false;
    *bw = rows; *bh = 1; *bn = 1;
    return true;
  }
  if (!pow2(W)) return false;
  *bw = W;
  const int rem = rows / W;
  if (H >= rem) {
    if (H % rem) return false;
    *bh = rem; *bn = 1;
    return true;
  }
  if (!pow2(H)) return false;
  *bh = H;
  *bn = rem / H;
  return true;
}

}  // namespace tvae

extern "C" const char* tvae_last_error() { return tvae::g_err; }
extern "C" int32_t tvae_abi_version() { return TVAE_ABI_VERSION; }
