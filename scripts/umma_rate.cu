// umma_rate.cu — microbenchmark: sustained cycles per tcgen05.mma (M128 x N x K16, bf16 -> fp32) as a function of N, the
// shared-memory layout of the operands (SWIZZLE_128B/64B/32B K-major rows) and the operand source (A from SMEM vs TMEM).
// It answers the design question of SURVEY hard part 1: is a voxels-on-M / Cout-on-N implicit GEMM bound by the SMEM
// operand fetch at small N, and by how much?  One CTA per SM, one thread issues `iters` MMAs whose A start address walks
// the 27-tap pattern of a halo tile (row shifts) exactly like the conv kernels do.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate scripts/umma_rate.cu
// Run:   build/umma_rate [ss|ts]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RP {
  int N, RB, layout_type, iters, nacc, ts, same_a, commit_every;
  long long* out;
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const RP P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // fill 200 KB with small finite bf16 values
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u;
    w[i] = 0x3C003C00u | (h & 0x007F007Fu);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sA = smem_u32(smem);
    const uint32_t a_rows = 4 * 128 + 2 * 130 + 8;
    const uint32_t sB = sA + ((a_rows * P.RB + 1023) / 1024) * 1024;
    const uint32_t b_room = 200 * 1024 - (sB - sA);
    int ntapB = (int)(b_room / (uint32_t)(P.N * P.RB));
    if (ntapB > 9) ntapB = 9;
    if (ntapB < 1) ntapB = 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi = ((uint64_t)(((8u * P.RB) >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)P.layout_type << 61);
    const uint32_t rb16 = P.RB >> 4;
    const int nk16 = P.RB / 32;
    const uint32_t bb = smem_u32(&bar), bb2 = smem_u32(&bar2);
    const uint32_t phase = 0;
    const long long t0 = clock64();
    int k16 = 0, tap = 0, mb = 0;
    for (int it = 0; it < P.iters; ++it) {
      uint32_t a_row = mb * 128 + (tap / 3) * 130 + (tap % 3);
      if (P.same_a) a_row = 0;
      const uint32_t a16 = (sA >> 4) + a_row * rb16 + k16 * 2;
      const uint32_t b16 = (sB >> 4) + (uint32_t)((tap % ntapB) * P.N) * rb16 + k16 * 2;
      const uint32_t d = tmem + (uint32_t)((mb % P.nacc) * P.N);
      const uint64_t adesc = hi | (a16 & 0x3FFFu) | (1ull << 16);
      const uint64_t bdesc = hi | (b16 & 0x3FFFu) | (1ull << 16);
      const uint32_t accum = it > 0 ? 1u : 0u;
      if (P.ts) {
        const uint32_t a_t = tmem + 384 + (uint32_t)(((tap * 4 + k16) & 15) * 8);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
      }
      if (++k16 == nk16) { k16 = 0; if (++tap == 9) { tap = 0; if (++mb == 4) mb = 0; } }
      if (P.commit_every > 0 && (it % P.commit_every) == P.commit_every - 1 && it + 1 < P.iters) {
        // emulate the per-stage commit of the conv kernels (arrive on a barrier nobody waits for except at the end)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bb2) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bb) : "memory");
    // wait for the final phase: with intermediate commits the barrier flips several times; poll until the parity we expect
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bb), "r"(phase) : "memory");
    }
    const long long t1 = clock64();
    P.out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// Lean variant: what a tuned conv kernel can do.  N / row bytes are template parameters, the 9 in-plane taps x K16 steps of
// one M-block are fully unrolled (descriptor low words = base + compile-time constant), `issuers` warps issue independent
// MMA streams (own accumulators) concurrently.
// ---------------------------------------------------------------------------------------------------------------
struct LP { int blocks; int issuers; int rowpitch; long long* out; };

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, int RB>
__global__ void __launch_bounds__(128, 1) lean_kernel(const LP P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u;
    w[i] = 0x3C003C00u | (h & 0x007F007Fu);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp < P.issuers) {
    constexpr uint32_t rb16 = RB / 16;
    constexpr int NK16 = RB / 32;
    const uint32_t sA16 = smem_u32(smem) >> 4;
    const uint32_t a_rows = 4 * 128 + 2 * 130 + 8;
    const uint32_t sB16 = sA16 + (((a_rows * RB + 1023) / 1024) * 1024 >> 4);
    constexpr int NTAPB = (90 * 1024) / (N * RB) >= 9 ? 9 : ((90 * 1024) / (N * RB) >= 3 ? 3 : 1);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi = ((uint64_t)(((8u * RB) >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(RB == 128 ? 2 : (RB == 64 ? 4 : 6)) << 61);
    const uint64_t hi_lbo = hi | (1ull << 16);
    const uint32_t rp1 = (uint32_t)P.rowpitch * rb16, rp2 = 2u * rp1;
    constexpr int NACC = (512 / N) >= 4 ? 4 : (512 / N);
    const uint32_t d0 = tmem + (uint32_t)((warp % NACC) * N);
    t0 = clock64();
    for (int blk = 0; blk < P.blocks; ++blk) {
      const uint32_t abase = sA16 + (uint32_t)((blk + warp) & 3) * 128u * rb16;
      if (elect_one()) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t arow = abase + (kh == 0 ? 0u : (kh == 1 ? rp1 : rp2));
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int k = 0; k < NK16; ++k) {
              const uint32_t alo = arow + kw * rb16 + k * 2;
              const uint32_t blo = sB16 + (uint32_t)(((kh * 3 + kw) % NTAPB) * N) * rb16 + k * 2;
              mma_ss(d0, hi_lbo | alo, hi_lbo | blo, idesc, (blk | kh | kw | k) ? 1u : 0u);
            }
          }
        }
      }
      __syncwarp();
    }
    const uint32_t bb = smem_u32(&bars[warp]);
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bb) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bb), "r"(0) : "memory");
    }
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) P.out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int RB>
static void run_lean(int issuers, int grid) {
  const int blocks = 600;
  long long* d;
  CHECK(cudaMalloc(&d, sizeof(long long) * grid * 4));
  CHECK(cudaMemset(d, 0, sizeof(long long) * grid * 4));
  LP P{blocks, issuers, 130, d};
  CHECK(cudaFuncSetAttribute(lean_kernel<N, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  lean_kernel<N, RB><<<grid, 128, 201 * 1024>>>(P);
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  lean_kernel<N, RB><<<grid, 128, 201 * 1024>>>(P);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid * 4);
  CHECK(cudaMemcpy(h.data(), d, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : h) mx = std::max(mx, v);
  const double mmas = (double)blocks * 9 * (RB / 32) * issuers;  // per CTA
  const double clk = (double)mx / mmas;
  const double floor_clk = N / 2.0;
  const double tflops = 2.0 * 128 * N * 16 * mmas * grid / (ms * 1e-3) / 1e12;
  printf("LEAN unrolled N=%3d RB=%3d issuers=%d grid=%3d  clk/MMA %.1f (floor %.0f => %.0f%% of tensor peak)  kernel %.3f ms  %.0f TFLOP/s\n",
         N, RB, issuers, grid, clk, floor_clk, 100.0 * floor_clk / clk, ms, tflops);
  cudaFree(d);
}


// MN-major variant (what conv_wg2.cu issues): A = 4 groups of 32 channels shifted by 1 row each (LBO = 1 row), B = N/32
// groups shifted by `rowpitch` rows, SWIZZLE_64B rows of 64 bytes, K = 16 consecutive rows.
template <int N>
__global__ void __launch_bounds__(128, 1) lean_mn_kernel(const LP P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) { uint32_t h = (uint32_t)i * 2654435761u; w[i] = 0x3C003C00u | (h & 0x007F007Fu); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp < P.issuers) {
    constexpr uint32_t RB = 64, rb16 = 4;
    const uint32_t sA16 = smem_u32(smem) >> 4;                 // dY-like tile: rows of 64 B
    const uint32_t sB16 = sA16 + (40 * 1024 >> 4);             // X-like planes
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi = ((uint64_t)(((8u * RB) >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)4 << 61);
    const uint32_t lboA = (RB >> 4) << 16;
    const uint32_t lboB = (((uint32_t)P.rowpitch * RB) >> 4) << 16;
    const uint32_t d0 = tmem + (uint32_t)(warp * 128);
    long long t0 = clock64();
    for (int blk = 0; blk < P.blocks; ++blk) {
      const uint32_t ab = (sA16 + (uint32_t)((blk & 1) * 130) * rb16) | lboA;
      const uint32_t bb = (sB16 + (uint32_t)((blk & 1) * P.rowpitch) * rb16) | lboB;
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
            mma_ss(d0 + kd * 0, hi | (ab + j * 16 * rb16), hi | (bb + kd * (40 * 1024 >> 4) + j * 16 * rb16), idesc, (blk | j | kd) ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    const uint32_t bbar = smem_u32(&bars[warp]);
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bbar) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bbar), "r"(0) : "memory");
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) P.out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N>
static void run_lean_mn(int rowpitch, int grid) {
  const int blocks = 800;
  long long* d;
  CHECK(cudaMalloc(&d, sizeof(long long) * grid * 4));
  CHECK(cudaMemset(d, 0, sizeof(long long) * grid * 4));
  LP P{blocks, 1, rowpitch, d};
  CHECK(cudaFuncSetAttribute(lean_mn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  lean_mn_kernel<N><<<grid, 128, 201 * 1024>>>(P);
  CHECK(cudaDeviceSynchronize());
  lean_mn_kernel<N><<<grid, 128, 201 * 1024>>>(P);
  CHECK(cudaDeviceSynchronize());
  std::vector<long long> h(grid * 4);
  CHECK(cudaMemcpy(h.data(), d, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : h) mx = std::max(mx, v);
  const double clk = (double)mx / ((double)blocks * 24);
  printf("LEAN MN-major SW64 N=%3d (A groups 1 row apart, B groups %d rows apart): clk/MMA %.1f (K-major model: %.0f)\n", N, rowpitch, clk,
         std::max(N / 2.0, (4096.0 + 32.0 * N) / 128.0));
  cudaFree(d);
}

static void run(const char* name, RP P, int grid) {
  long long* d;
  CHECK(cudaMalloc(&d, sizeof(long long) * grid));
  P.out = d;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<<<grid, 128, 201 * 1024>>>(P);  // warm
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate_kernel<<<grid, 128, 201 * 1024>>>(P);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid);
  CHECK(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  const double med = (double)h[grid / 2] / P.iters;
  const double floor_clk = P.N / 2.0;  // 128*N*16 MAC / 4096 MAC/clk
  const double tflops = 2.0 * 128 * P.N * 16 * (double)P.iters * grid / (ms * 1e-3) / 1e12;
  printf("%-44s N=%3d RB=%3d nacc=%d grid=%3d  clk/MMA min %.1f med %.1f max %.1f  (floor %.0f => %.0f%% of tensor peak)  kernel %.3f ms  %.0f TFLOP/s\n",
         name, P.N, P.RB, P.nacc, grid, (double)h[0] / P.iters, med, (double)h[grid - 1] / P.iters, floor_clk,
         100.0 * floor_clk / med, ms, tflops);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const bool ts = argc > 1 && !strcmp(argv[1], "ts");
  if (argc > 1 && !strcmp(argv[1], "mn")) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run_lean_mn<96>(128, sms); run_lean_mn<96>(130, sms); run_lean_mn<96>(64, sms); run_lean_mn<48>(128, sms);
    run_lean_mn<32>(128, sms); run_lean_mn<64>(128, sms); run_lean_mn<128>(128, sms);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "lean")) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int iss : {1, 2, 4}) {
      run_lean<32, 128>(iss, sms); run_lean<64, 128>(iss, sms); run_lean<96, 128>(iss, sms); run_lean<128, 128>(iss, sms);
      run_lean<192, 128>(iss, sms); run_lean<256, 128>(iss, sms);
      run_lean<32, 64>(iss, sms); run_lean<96, 64>(iss, sms);
    }
    run_lean<32, 128>(1, 1); run_lean<96, 128>(1, 1); run_lean<256, 128>(1, 1);
    return 0;
  }
  CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  if (!ts) {
    for (int n : Ns) {
      RP P{n, 128, 2, iters, 4, 0, 0, 0, nullptr};
      if (4 * n > 512) P.nacc = 512 / n;
      run("SS SW128 tap-walk", P, sms);
    }
    for (int n : {32, 96}) {
      RP P{n, 64, 4, iters, 4, 0, 0, 0, nullptr};
      run("SS SW64 tap-walk", P, sms);
      RP Q{n, 32, 6, iters, 4, 0, 0, 0, nullptr};
      run("SS SW32 tap-walk", Q, sms);
    }
    for (int n : {32, 96, 256}) {
      RP P{n, 128, 2, iters, 1, 0, 1, 0, nullptr};
      run("SS SW128 same-A single acc", P, sms);
    }
    for (int n : {32, 96}) {
      RP P{n, 128, 2, iters, 4, 0, 0, 8, nullptr};
      run("SS SW128 tap-walk commit/8", P, sms);
    }
    for (int n : {32, 96, 256}) {
      RP P{n, 128, 2, iters, 4, 0, 0, 0, nullptr};
      if (4 * n > 512) P.nacc = 512 / n;
      run("SS SW128 tap-walk 1 CTA only", P, 1);
    }
  } else {
    for (int n : {16, 32, 64, 96, 128, 256}) {
      RP P{n, 128, 2, iters, 3, 1, 0, 0, nullptr};
      if (3 * n > 384) P.nacc = 384 / n;
      run("TS (A in TMEM) SW128 B", P, sms);
    }
  }
  return 0;
}
