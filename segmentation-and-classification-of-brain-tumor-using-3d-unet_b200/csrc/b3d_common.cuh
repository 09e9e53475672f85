// b3d_common.cuh — shared device/host helpers for the B200-native 3D U-Net hot path.
// Raw PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc/mma/commit/ld), fences.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// Error plumbing (never throw / exit across the C ABI)
// ---------------------------------------------------------------------------------------------
#define B3D_OK 0
#define B3D_ERR_INVALID   -1   // bad shape / argument
#define B3D_ERR_CUDA      -2   // CUDA runtime error
#define B3D_ERR_UNSUPPORTED -3 // shape outside what the sm_100a kernels cover (no fallback exists)
#define B3D_ERR_DRIVER    -4   // driver entry point / tensor-map encode failure

void b3d_set_error(const char* fmt, ...);

#define B3D_CHECK_CUDA(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      b3d_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,            \
                    cudaGetErrorString(_e));                                         \
      return B3D_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define B3D_REQUIRE(cond, ...)                                                       \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      b3d_set_error(__VA_ARGS__);                                                    \
      return B3D_ERR_INVALID;                                                        \
    }                                                                                \
  } while (0)

int b3d_num_sms();
// Debugging knobs are read from the environment ONCE per call site (function-local static: thread-safe, no getenv on the
// launch path).
#include <stdlib.h>
#define B3D_ENV_FLAG(name) ([] { static const bool v_ = getenv(name) != nullptr; return v_; }())
#define B3D_ENV_INT(name) ([] { static const int v_ = getenv(name) ? atoi(getenv(name)) : 0; return v_; }())
// b3d_set_ordered_issue(1): one MMA-issuing warp instead of two ping-pong issuers in conv_zs.cu, which makes the fp32
// accumulation order — hence the forward pass and the input gradients — bit-reproducible from run to run (default: env
// B3D_ORDERED_ISSUE, else 0).
#include <atomic>
extern std::atomic<int> g_b3d_ordered_issue;
#include <atomic>
extern std::atomic<long long> g_b3d_launches;  // kernels launched by the library (bench.py reports it)

// ---------------------------------------------------------------------------------------------
// Small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Σ over the lanes of a warp that share (lane % group), group a power of two <= 32 (32: no-op).  Every lane must call it.
// Used before block-level fp64 shared atomics so that only `group` lanes per warp contend (and the sum stays order-fixed).
__device__ __forceinline__ float warp_sum_mod(float v, int group) {
  for (int o = 16; o >= group; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 bf16 <-> 8 floats through one 16-byte vector
struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// (streaming .cs stores measured no faster on any bandwidth kernel and 1 % slower on the step: outputs are the next kernel's inputs)
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// "register-lean" streaming kernels keep their per-channel coefficients in shared memory and re-read them inside the loop:
// the volatile asm stops the compiler from hoisting the loads back into (dozens of) registers.
__device__ __forceinline__ float2 lds2v(const float* p) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ float2 bfw(unsigned w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }
__device__ __forceinline__ unsigned wbf(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&t);
}

// ---------------------------------------------------------------------------------------------
// PTX: shared-address conversion, mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (CUDA error surfaced to the caller) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {  // ~2 s at 2 GHz: far beyond any legitimate wait
        if (err_flag) atomicExch(err_flag, code);
        __threadfence_system();
        __trap();
      }
    }
  }
}

// TMA 5-D tiled load (coords innermost-first), completes on an mbarrier with tx bytes.
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Warp-uniform leader election (elect.sync): role warps keep their control flow warp-uniform so that descriptors live
// in uniform registers, and only the tcgen05 / TMA issue itself is predicated on the elected lane (the CUTLASS idiom).
// A plain `if (lane == 0)` around the loops makes ptxas emit an ELECT/R2UR "waterfall" around every UTCHMMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread l of warp w reads lane 32*(w%4)+l).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8x16B core matrices).
//   K-major : core matrix = 8 rows (M/N) x 16 B (8 bf16 of K), rows 16 B apart;
//             LBO = byte distance between the two K halves of one K=16 step, SBO = distance between 8-row groups.
//   MN-major: core matrix = 8 K-rows x 16 B (8 bf16 of M/N), K rows 16 B apart;
//             SBO = distance between 8-element M/N groups, LBO = distance between 8-row K groups.
// (cute::UMMA::SmemDescriptor bit layout: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=0.)
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t sbo_bytes) {
  return (uint64_t)(((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14)) << 32;
}
// swizzled K-major descriptor high word: layout_type 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B (0 = none)
__device__ __forceinline__ uint64_t umma_desc_hi_sw(uint32_t sbo_bytes, int layout_type) {
  return umma_desc_hi(sbo_bytes) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint32_t lo = ((addr_bytes >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  return umma_desc_hi(sbo_bytes) | lo;
}
// Instruction descriptor for kind::f16: bf16 A/B, fp32 accumulate. major: 0 = K-major, 1 = MN-major.
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N, int a_major, int b_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Host: TMA descriptor encode through the driver entry point (no -lcuda link dependency)
// ---------------------------------------------------------------------------------------------
int b3d_encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, int swizzle_bytes = 0);
