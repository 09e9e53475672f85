// pdl_probe.cu — how much of the kernel-to-kernel gap inside a replayed CUDA graph does programmatic dependent launch
// (griddepcontrol.wait at the top of every kernel + griddepcontrol.launch_dependents, launch attribute
// cudaLaunchAttributeProgrammaticStreamSerialization) recover on B200?  The train step is one graph of ~400 dependent
// launches that alternate persistent one-CTA-per-SM tensor kernels (~200 KB of shared memory) with bandwidth kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pdl_probe scripts/pdl_probe.cu && /tmp/pdl_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// bandwidth kernel: y = x * 1.0001 over n float4 (grid-stride)
template <int EARLY>
__global__ void __launch_bounds__(256) stream_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n) {
  if (EARLY) pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    float4 v = x[i]; v.x *= 1.0001f; v.y *= 1.0001f; v.z *= 1.0001f; v.w *= 1.0001f; y[i] = v;
  }
  if (!EARLY) pdl_trigger();
}

// persistent "tensor-like" kernel: one CTA per SM, big dynamic smem, prologue that clears smem, then a per-CTA slice of work
template <int EARLY>
__global__ void __launch_bounds__(352, 1) persistent_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n, int spin) {
  extern __shared__ float4 sm[];
  if (EARLY) pdl_trigger();
  for (int i = threadIdx.x; i < 2048; i += 352) sm[i] = make_float4(0, 0, 0, 0);   // prologue (barrier init / TMEM alloc stand-in)
  __syncthreads();
  pdl_wait();
  const long long per = (n + gridDim.x - 1) / gridDim.x, b = blockIdx.x * per, e = min(n, b + per);
  for (long long i = b + threadIdx.x; i < e; i += 352) {
    float4 v = x[i];
    for (int k = 0; k < spin; ++k) v.x = v.x * 1.0001f + sm[(threadIdx.x + k) & 2047].x;
    y[i] = v;
  }
  if (!EARLY) pdl_trigger();
}

template <typename K, typename... A>
static void launch(K k, dim3 g, dim3 b, size_t smem, cudaStream_t s, int pdl, A... a) {
  cudaLaunchConfig_t c = {};
  c.gridDim = g; c.blockDim = b; c.dynamicSmemBytes = smem; c.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  c.attrs = at; c.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&c, k, a...));
}

int main() {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const size_t smem = 200 * 1024;
  CK(cudaFuncSetAttribute(persistent_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nbig = (256LL << 20) / 16;   // 256 MB buffers
  float4 *a, *b; CK(cudaMalloc(&a, nbig * 16)); CK(cudaMalloc(&b, nbig * 16)); CK(cudaMemset(a, 0, nbig * 16)); CK(cudaMemset(b, 0, nbig * 16));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int L = 200;   // launches per graph
  // sizes: (bytes of the bandwidth kernel, bytes of the persistent kernel)
  const long long sizes[3][2] = {{1 << 20, 1 << 20}, {32 << 20, 16 << 20}, {256 << 20, 64 << 20}};
  for (int sz = 0; sz < 3; ++sz) {
    for (int mode = 0; mode < 3; ++mode) {   // 0 plain, 1 PDL (trigger at the end), 2 PDL (trigger at the top)
      const long long ns = sizes[sz][0] / 16, np = sizes[sz][1] / 16;
      cudaGraph_t g; cudaGraphExec_t ge;
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < L; ++i) {
        float4* src = (i & 1) ? b : a; float4* dst = (i & 1) ? a : b;
        if (i & 1) {
          if (mode == 2) launch(persistent_kernel<1>, dim3(148), dim3(352), smem, s, 1, (const float4*)src, dst, np, 8);
          else launch(persistent_kernel<0>, dim3(148), dim3(352), smem, s, mode, (const float4*)src, dst, np, 8);
        } else {
          const int grid = (int)((ns + 255) / 256 < 148 * 8 ? (ns + 255) / 256 : 148 * 8);
          if (mode == 2) launch(stream_kernel<1>, dim3(grid), dim3(256), 0, s, 1, (const float4*)src, dst, ns);
          else launch(stream_kernel<0>, dim3(grid), dim3(256), 0, s, mode, (const float4*)src, dst, ns);
        }
      }
      CK(cudaStreamEndCapture(s, &g));
      CK(cudaGraphInstantiate(&ge, g, 0));
      for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
      CK(cudaStreamSynchronize(s));
      CK(cudaEventRecord(e0, s));
      const int reps = 10;
      for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(e1, s));
      CK(cudaStreamSynchronize(s));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("stream %4lld MB / persistent %3lld MB  mode %d (%s): %.2f us per launch\n", sizes[sz][0] >> 20, sizes[sz][1] >> 20, mode,
             mode == 0 ? "plain" : mode == 1 ? "PDL, trigger at end" : "PDL, trigger at top", ms * 1000.0 / (reps * L));
      CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
  }
  return 0;
}
