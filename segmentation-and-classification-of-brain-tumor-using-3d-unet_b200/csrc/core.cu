// core.cu — library-wide plumbing: last-error string, device query, TMA descriptor encode via the driver entry point.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <stdarg.h>
#include <mutex>
#include <atomic>

static thread_local char g_err[1024] = "";
std::atomic<long long> g_b3d_launches{0};   // the Flask app serves threaded (main.py:1059): count atomically

std::atomic<int> g_b3d_ordered_issue{getenv("B3D_ORDERED_ISSUE") ? atoi(getenv("B3D_ORDERED_ISSUE")) : 0};

void b3d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// SMs left free for someone else's resident CTAs (NCCL channel CTAs under data parallelism): every grid of this library is
// sized from b3d_num_sms(), and the tensor-core kernels are persistent one-CTA-per-SM grids that cannot share an SM with an
// NCCL CTA — a conv CTA that finds "its" SM taken waits for the whole collective to finish.
std::atomic<int> g_b3d_reserved_sms{getenv("B3D_RESERVED_SMS") ? atoi(getenv("B3D_RESERVED_SMS")) : 0};

int b3d_num_sms() {
  static const int sms = [] {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  const int r = g_b3d_reserved_sms.load(std::memory_order_relaxed);
  return r > 0 ? (sms - r > 8 ? sms - r : 8) : sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

int b3d_encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  if (!g_encode) {
    b3d_set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return B3D_ERR_DRIVER;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                    : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                           : CU_TENSOR_MAP_SWIZZLE_NONE)),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b3d_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu %llu box %u %u %u %u %u base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
                  rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, base);
    return B3D_ERR_DRIVER;
  }
  return B3D_OK;
}

extern "C" {
const char* b3d_last_error_string() { return g_err; }
int b3d_version() { return 100; }
long long b3d_launch_count() { return g_b3d_launches.load(); }
// 1: bit-reproducible forward / input gradients (single MMA issuer in the z-marching conv kernel); returns the old value
// grids of every later launch use (SM count - n) SMs; returns the previous reservation.  Process-wide, atomic.
int b3d_set_reserved_sms(int n) { return g_b3d_reserved_sms.exchange(n < 0 ? 0 : n); }
int b3d_set_ordered_issue(int on) { return g_b3d_ordered_issue.exchange(on < 0 ? 0 : (on > 2 ? 2 : on)); }
// ---- workspace sizes: the caller allocates every scratch buffer (nothing persistent lives in the library), these say how much ----
// split-K partial slices of b3d_conv_fprop / b3d_conv_fprop_add: up to 16 fp32 slices [split][voxels][Cout]; only the small
// (deep-level) problems ever split, so problems whose 16 slices would exceed 64 MB get none (0 = pass ws = NULL).
size_t b3d_conv_fprop_workspace_bytes(int N, int D, int H, int W, int Cout) {
  const size_t one = (size_t)N * D * H * W * Cout * 4;
  return one * 16 <= ((size_t)1 << 26) ? one * 16 : 0;
}
// b3d_convT2_dgrad: the same policy on its [voxels][Cin] output (one slice is always required)
size_t b3d_convT2_dgrad_workspace_bytes(int N, int D, int H, int W, int Cin) {
  const size_t one = (size_t)N * D * H * W * Cin * 4;
  return one * 16 <= ((size_t)1 << 26) ? one * 16 : one;
}
// b3d_conv_wgrad: the tap-major fp32 accumulator [ks^3][Cin][roundup16(Cout)] the flush kernels reduce into before the
// transposing finalize writes the reference layout; b3d_convT2_wgrad: [8][Cin][Cout]
size_t b3d_conv_wgrad_workspace_bytes(int Cin, int Cout, int ks) {
  return (size_t)ks * ks * ks * Cin * ((Cout + 15) / 16 * 16) * 4;
}
size_t b3d_convT2_wgrad_workspace_bytes(int Cin, int Cout) { return (size_t)8 * Cin * Cout * 4; }
// 0 if the current device is sm_100 (B200); negative otherwise — callers must fail loudly, there is no fallback.
int b3d_check_device() {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { b3d_set_error("no CUDA device"); return B3D_ERR_CUDA; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) { b3d_set_error("device is sm_%d%d, libb3d is built for sm_100a only", major, minor); return B3D_ERR_UNSUPPORTED; }
  return B3D_OK;
}
}
