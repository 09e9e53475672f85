// tma_probe.cu — how fast does one SM's TMA engine move a tiled box, as a function of inner-row bytes and rank?
// Every CTA (one per SM) issues `iters` box loads (2 in flight) from an L2-resident NDHWC tensor and reports clk/box.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Params { CUtensorMap tm; int rank; int iters; uint32_t bytes; int H, D; long long* out; };
__global__ void __launch_bounds__(32, 1) probe(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t half = (P.bytes + 1023) / 1024 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < P.iters + 2; ++it) {
      const int s = it & 1;
      if (it >= 2) {
        uint32_t ok = 0;
        const uint32_t par = ((it - 2) >> 1) & 1;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
      }
      if (it < P.iters) {
        const uint32_t fb = smem_u32(&bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(P.bytes) : "memory");
        const int y = ((blockIdx.x * 7 + it * 3) % (P.H - 8));
        const int z = (blockIdx.x + it) % P.D;
        if (P.rank == 5)
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(smem_u32(smem) + s * half), "l"(&P.tm), "r"(fb), "r"(0), "r"(-1), "r"(y), "r"(z), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(smem) + s * half), "l"(&P.tm), "r"(fb), "r"(0), "r"((y * 131 + z * 977) % 60000) : "memory");
      }
    }
    P.out[blockIdx.x] = clock64() - t0;
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int W = 128, H = 128, D = 32, C = 64;  // 64 MB bf16 tensor: L2 resident
  void* buf; CHECK(cudaMalloc(&buf, (size_t)W * H * D * C * 2)); CHECK(cudaMemset(buf, 1, (size_t)W * H * D * C * 2));
  long long* dout; CHECK(cudaMalloc(&dout, 148 * 8));
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int iters = 64;
  printf("%-40s %10s %10s %10s\n", "box", "bytes", "clk/box", "B/clk/SM");
  for (int rank = 5; rank >= 2; rank -= 3)
    for (int kc = 8; kc <= 64; kc *= 2)
      for (int rows = 0; rows < 2; ++rows) {
        Params P; memset(&P, 0, sizeof(P));
        const int rb = kc * 2;
        CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : rb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        char name[128];
        if (rank == 5) {
          const int BH = rows ? 6 : 3;
          cuuint64_t dims[5] = {(cuuint64_t)C, W, H, D, 1}; cuuint64_t st[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H, (cuuint64_t)C * 2 * W * H * D};
          cuuint32_t box[5] = {(cuuint32_t)kc, 130, (cuuint32_t)BH, 1, 1};
          CUresult r = enc(&P.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r) { printf("encode failed %d\n", (int)r); return 1; }
          P.bytes = (uint32_t)kc * 2 * 130 * BH; sprintf(name, "5D halo box (%dch,130,%d,1,1) rows=%d", kc, BH, 130 * BH);
        } else {
          const int R = rows ? 256 : 128;
          cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)W * H * D}; cuuint64_t st[1] = {(cuuint64_t)C * 2};
          cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)R};
          CUresult r = enc(&P.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r) { printf("encode failed %d\n", (int)r); return 1; }
          P.bytes = (uint32_t)kc * 2 * R; sprintf(name, "2D box (%dch, %d rows)", kc, R);
        }
        P.rank = rank; P.iters = iters; P.H = H; P.D = D; P.out = dout;
        probe<<<148, 32, 2 * ((P.bytes + 1023) / 1024 * 1024) + 1024>>>(P);
        CHECK(cudaDeviceSynchronize());
        probe<<<148, 32, 2 * ((P.bytes + 1023) / 1024 * 1024) + 1024>>>(P);
        CHECK(cudaDeviceSynchronize());
        long long h[148]; CHECK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148.0 * iters;
        printf("%-40s %10u %10.0f %10.2f\n", name, P.bytes, avg, P.bytes / avg);
      }
  return 0;
}
