// umma_mn_test.cu — bring-up probe for the weight-gradient kernel: MN-major SWIZZLED operands for tcgen05.mma.
// A tile T[position][channel] (row = RB bytes, TMA-written with SWIZZLE_64B/128B) is used as an MN-major operand:
// K = 16 consecutive positions (rows) starting at an ARBITRARY row, M (or N) = groups of GW = RB/2 channels, group g read
// from the same tile shifted by g*L rows (LBO = L*RB bytes: the "taps stacked on M/N" trick).
//   D[m][n] = sum_k TA[ra + (m/GW)*LA + k][m % GW] * TB[rb + (n/GW)*LB + k][n % GW]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_mn_test scripts/umma_mn_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  CUtensorMap tmA, tmB;
  int RB, layout_type, rows;   // tile rows staged for each operand
  int ra, LA, rb, LB, N;       // start rows, group strides (rows), N columns
  int swap_lbo_sbo;            // 1: put the group stride in SBO and the 8-row K-group stride in LBO (no-swizzle convention)
  float* out;                  // [128][N]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sA = smem_u32(smem);
  const uint32_t bytes = P.rows * P.RB;
  const uint32_t sB = sA + ((bytes + 1023) / 1024) * 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fb = smem_u32(&bar_full);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(2 * bytes) : "memory");
    for (int r = 0; r < P.rows; r += 128) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sA + r * P.RB), "l"(&P.tmA), "r"(fb), "r"(0), "r"(r) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sB + r * P.RB), "l"(&P.tmB), "r"(fb), "r"(0), "r"(r) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(fb), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // kind::f16, bf16 x bf16 -> f32, A and B MN-major (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(P.N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t kgrp = 8u * P.RB;  // 8 K rows
    auto mk = [&](uint32_t addr, uint32_t group_stride_bytes) -> uint64_t {
      uint32_t lbo = group_stride_bytes, sbo = kgrp;
      if (P.swap_lbo_sbo) { lbo = kgrp; sbo = group_stride_bytes; }
      return ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)P.layout_type << 61) |
             ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((addr >> 4) & 0x3FFF);
    };
    const uint64_t adesc = mk(sA + P.ra * P.RB, (uint32_t)P.LA * P.RB);
    const uint64_t bdesc = mk(sB + P.rb * P.RB, (uint32_t)P.LB * P.RB);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  __syncwarp();
  {
    uint32_t ok = 0;
    const uint32_t mb = smem_u32(&bar_mma);
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(mb), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int j0 = 0; j0 < P.N; j0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + j0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) P.out[(warp * 32 + lane) * P.N + j0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int modes[3][3] = {{64, 4, (int)CU_TENSOR_MAP_SWIZZLE_64B}, {128, 2, (int)CU_TENSOR_MAP_SWIZZLE_128B}, {32, 6, (int)CU_TENSOR_MAP_SWIZZLE_32B}};
  for (int m = 0; m < 3; ++m) {
    const int RB = modes[m][0], GW = RB / 2;
    const int ROWS = (RB == 128) ? 640 : 1024;
    std::vector<__nv_bfloat16> hA((size_t)ROWS * GW), hB((size_t)ROWS * GW);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(7 + m);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (float)((rand() % 13) - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = v; }
    __nv_bfloat16 *dA, *dB; float* dOut;
    CHECK(cudaMalloc(&dA, hA.size() * 2)); CHECK(cudaMalloc(&dB, hB.size() * 2)); CHECK(cudaMalloc(&dOut, 128 * 256 * 4));
    CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    Params P;
    memset(&P, 0, sizeof(P));
    cuuint64_t dims[2] = {(cuuint64_t)GW, (cuuint64_t)ROWS}; cuuint64_t strides[1] = {(cuuint64_t)RB};
    cuuint32_t box[2] = {(cuuint32_t)GW, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&P.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (CUtensorMapSwizzle)modes[m][2], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; }
    r = enc(&P.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            (CUtensorMapSwizzle)modes[m][2], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    P.RB = RB; P.layout_type = modes[m][1]; P.rows = ROWS; P.out = dOut;
    // {ra, LA, rb, LB, groupsN}
    const int cases[][5] = {{0, 16, 0, 16, 1}, {0, 16, 0, 16, 2}, {8, 16, 24, 16, 3}, {1, 1, 0, 130, 3}, {3, 1, 5, 66, 3},
                            {17, 130, 2, 1, 3}, {0, 1, 0, 34, 3}};
    for (int swap = 0; swap < 2; ++swap)
      for (auto& c : cases) {
        P.ra = c[0]; P.LA = c[1]; P.rb = c[2]; P.LB = c[3]; P.N = c[4] * GW; P.swap_lbo_sbo = swap;
        if (P.N > 256 || P.N % 16) continue;
        const int MG = 128 / GW;
        CHECK(cudaMemset(dOut, 0, 128 * 256 * 4));
        probe_kernel<<<1, 128, 200 * 1024>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * P.N);
        CHECK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int i = 0; i < 128; ++i)
          for (int j = 0; j < P.N; ++j) {
            double ref = 0;
            for (int k = 0; k < 16; ++k)
              ref += (double)fA[(size_t)(P.ra + (i / GW) * P.LA + k) * GW + i % GW] * fB[(size_t)(P.rb + (j / GW) * P.LB + k) * GW + j % GW];
            maxerr = fmax(maxerr, fabs(ref - out[i * P.N + j]));
          }
        printf("RB %3d (M groups %d) ra %2d LA %3d rb %2d LB %3d N %3d swap %d : max err %.4f %s\n", RB, MG, P.ra, P.LA, P.rb, P.LB,
               P.N, swap, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
  }
  return 0;
}
