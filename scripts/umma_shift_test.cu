// umma_shift_test.cu — bring-up probe: can a K-major SWIZZLED smem tile (filled by TMA) be consumed by tcgen05.mma with a
// start address shifted by an arbitrary number of ROWS (what an implicit-GEMM tap shift needs)?  Tries base_offset = 0
// and base_offset = (addr >> 7) & 7 for SWIZZLE_128B / 64B / 32B and prints the max error against a CPU reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_shift_test scripts/umma_shift_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  CUtensorMap tmA, tmB;
  int rowbytes;      // 128, 64, 32
  int layout_type;   // 2 = SW128, 4 = SW64, 6 = SW32
  int shift;         // rows
  int base_mode;     // 0: base_offset 0 ; 1: (addr>>7)&7
  int ksteps;        // rowbytes / 32
  float* out;        // [128][16]
  int a_rows;        // rows staged for A
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sA = smem_u32(smem);
  const uint32_t a_bytes = P.a_rows * P.rowbytes;
  const uint32_t sB = sA + ((a_bytes + 1023) / 1024) * 1024;
  const uint32_t b_bytes = 16 * P.rowbytes;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fb = smem_u32(&bar_full);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(a_bytes + b_bytes) : "memory");
    // A in boxes of 128 rows
    for (int r = 0; r < P.a_rows; r += 128)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sA + r * P.rowbytes), "l"(&P.tmA), "r"(fb), "r"(0), "r"(r) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"(&P.tmB), "r"(fb), "r"(0), "r"(0) : "memory");
    // wait
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(fb), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sbo = 8 * P.rowbytes;
    for (int k = 0; k < P.ksteps; ++k) {
      const uint32_t a_addr = sA + P.shift * P.rowbytes + k * 32;
      const uint32_t b_addr = sB + k * 32;
      uint32_t base = 0;
      if (P.base_mode == 1) base = (a_addr >> 7) & 7u;
      const uint64_t hi_common = ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)P.layout_type << 61);
      const uint64_t adesc = hi_common | ((uint64_t)base << 49) | ((a_addr >> 4) & 0x3FFF) | (1ull << 16);
      const uint64_t bdesc = hi_common | ((b_addr >> 4) & 0x3FFF) | (1ull << 16);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  __syncwarp();
  {
    uint32_t ok = 0;
    const uint32_t mb = smem_u32(&bar_mma);
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(mb), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) P.out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int AROWS = 384;
  const int modes[3][3] = {{128, 2, (int)CU_TENSOR_MAP_SWIZZLE_128B}, {64, 4, (int)CU_TENSOR_MAP_SWIZZLE_64B}, {32, 6, (int)CU_TENSOR_MAP_SWIZZLE_32B}};
  const int shifts[] = {0, 8, 1, 2, 3, 4, 5, 7, 9, 130, 131};
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int m = 0; m < 3; ++m) {
    const int rowbytes = modes[m][0], K = rowbytes / 2;
    std::vector<__nv_bfloat16> hA((size_t)AROWS * K), hB((size_t)16 * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(123 + m);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (float)((rand() % 13) - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dOut;
    CHECK(cudaMalloc(&dA, hA.size() * 2)); CHECK(cudaMalloc(&dB, hB.size() * 2)); CHECK(cudaMalloc(&dOut, 128 * 16 * 4));
    CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    Params P;
    memset(&P, 0, sizeof(P));
    {
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)AROWS}; cuuint64_t strides[1] = {(cuuint64_t)rowbytes};
      cuuint32_t box[2] = {(cuuint32_t)K, 128}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&P.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       (CUtensorMapSwizzle)modes[m][2], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", (int)r); return 1; }
      cuuint64_t dimsb[2] = {(cuuint64_t)K, 16}; cuuint32_t boxb[2] = {(cuuint32_t)K, 16};
      r = enc(&P.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              (CUtensorMapSwizzle)modes[m][2], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    P.rowbytes = rowbytes; P.layout_type = modes[m][1]; P.ksteps = rowbytes / 32; P.out = dOut; P.a_rows = AROWS;
    for (int base_mode = 0; base_mode < 2; ++base_mode)
      for (size_t si = 0; si < sizeof(shifts) / sizeof(int); ++si) {
        P.shift = shifts[si]; P.base_mode = base_mode;
        CHECK(cudaMemset(dOut, 0, 128 * 16 * 4));
        probe_kernel<<<1, 128, 100 * 1024>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * 16);
        CHECK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int i = 0; i < 128; ++i)
          for (int j = 0; j < 16; ++j) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)(i + P.shift) * K + k] * fB[(size_t)j * K + k];
            maxerr = fmax(maxerr, fabs(ref - out[i * 16 + j]));
          }
        printf("swizzle %3dB shift %3d base_mode %d : max err %.4f %s\n", rowbytes, P.shift, base_mode, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
  }
  return 0;
}
