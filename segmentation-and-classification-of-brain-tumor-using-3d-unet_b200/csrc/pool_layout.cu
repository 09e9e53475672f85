// pool_layout.cu — MaxPool3d(2,2)+Dropout3d forward/backward, NCDHW fp32 -> NDHWC bf16 input staging,
//                  channel sums, column (bias-gradient) sums.
//
// Reference semantics:
//   self.pool = nn.MaxPool3d(2,2); self.dropout = nn.Dropout3d(p)      /root/reference/main.py:109-110,173-174
//     out[n, z,y,x, c] = max_{a,b,c'} x[n, 2z+a, 2y+b, 2x+c', c] * mask[n,c]          (mask = Bernoulli(1-p)/(1-p), or 1)
//     backward routes dy*mask to the FIRST maximum in (a,b,c') raster order (ATen max_pool3d_with_indices tie rule).
//   images.to(device) fp32 NCDHW                                        /root/reference/training.py:287
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

static int ew_blocks2(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)b3d_num_sms() * 16;
  return (int)std::max<long long>(1, std::min(b, cap));
}

__global__ void __launch_bounds__(256) pool_fwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ mask,
                                                       bf16* __restrict__ out, long long ldo, int N, int D, int H, int W,
                                                       int C, int relu) {
  const int C8 = C >> 3;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int c0 = (int)(t % C8) * 8; t /= C8;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do); const int n = (int)(t / Do);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const long long vox = (((long long)n * D + 2 * zo + a) * H + 2 * yo + b) * W + 2 * xo + c;
          float v[8];
          unpack8(ldg16_stream(x + vox * ldx + c0), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
        }
    if (relu) {   // MaxPool(ReLU(x)) = ReLU(MaxPool(x))  (BrainTumorClassifier, main.py:307-312)
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], 0.f);
    }
    if (mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] *= mask[(long long)n * C + c0 + j];
    }
    const long long ovox = (((long long)n * Do + zo) * Ho + yo) * Wo + xo;
    stg16(out + ovox * ldo + c0, pack8(m));
  }
}

// AdaptiveAvgPool3d over NDHWC bf16 (optionally of ReLU(x)): one block per (n, output cell); a thread owns an 8-channel chunk
// and a slice of the window's voxels; partial sums are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(256) relu_adaptive_avgpool_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out,
                                                                    int D, int H, int W, int C, int OD, int OH, int OW, int relu) {
  __shared__ float part[256 * 8];
  const int n = blockIdx.y;
  int o = blockIdx.x;
  const int ox = o % OW; o /= OW;
  const int oy = o % OH; const int oz = o / OH;
  const int z0 = (oz * D) / OD, z1 = ((oz + 1) * D + OD - 1) / OD;
  const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH;
  const int x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
  const int wz = z1 - z0, wy = y1 - y0, wx = x1 - x0;
  const int nvox = wz * wy * wx;
  const int C8 = C >> 3;
  for (int cb = 0; cb < C8; cb += 256) {   // 256 chunks (2048 channels) per pass; thread t: chunk cb + t % nch, voxel lane t / nch
    const int nch = min(256, C8 - cb);
    const int lanes = 256 / nch;            // voxel lanes per chunk (>= 1)
    const int ch = threadIdx.x % nch, vl = threadIdx.x / nch;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (vl < lanes) {
      for (int v = vl; v < nvox; v += lanes) {
        const int dx = v % wx; const int t = v / wx; const int dy = t % wy; const int dz = t / wy;
        const long long vox = (((long long)n * D + z0 + dz) * H + y0 + dy) * W + x0 + dx;
        float f[8];
        unpack8(ldg16_stream(x + vox * ldx + (cb + ch) * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += relu ? fmaxf(f[j], 0.f) : f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < nch) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
      for (int l = 0; l < lanes; ++l)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += part[(l * nch + threadIdx.x) * 8 + j];
      const float inv = 1.f / (float)nvox;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (cb + threadIdx.x) * 8 + j;
        out[(((long long)n * C + c) * OD + oz) * OH * OW + oy * OW + ox] = s[j] * inv;
      }
    }
    __syncthreads();
  }
}

// dx (full-res) = dy*mask at the first arg-max of each 2x2x2 window, 0 elsewhere.  ACC: dx += ...
template <bool ACC>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ mask,
                                                       const bf16* __restrict__ dy, long long lddy, bf16* __restrict__ dx,
                                                       long long lddx, const float* __restrict__ cadd, int N, int D, int H, int W,
                                                       int C) {
  const int C8 = C >> 3;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int c0 = (int)(t % C8) * 8; t /= C8;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do); const int n = (int)(t / Do);
    float v[8][8];
    float m[8]; int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = k >> 2, b = (k >> 1) & 1, c = k & 1;
      const long long vox = (((long long)n * D + 2 * zo + a) * H + 2 * yo + b) * W + 2 * xo + c;
      unpack8(ldg16_stream(x + vox * ldx + c0), v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[k][j] > m[j]) { m[j] = v[k][j]; arg[j] = k; }
    }
    const long long ovox = (((long long)n * Do + zo) * Ho + yo) * Wo + xo;
    float g[8];
    unpack8(ldg16_stream(dy + ovox * lddy + c0), g);
    if (mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= mask[(long long)n * C + c0 + j];
    }
    float ca[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ACC && cadd) {   // per-(sample, channel) constant owed to the running gradient (the gate's channel-attention branch)
#pragma unroll
      for (int j = 0; j < 8; ++j) ca[j] = cadd[(long long)n * C + c0 + j];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = k >> 2, b = (k >> 1) & 1, c = k & 1;
      const long long vox = (((long long)n * D + 2 * zo + a) * H + 2 * yo + b) * W + 2 * xo + c;
      float o[8];
      if (ACC) unpack8(ldg16(dx + vox * lddx + c0), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float val = (arg[j] == k) ? g[j] : 0.f;
        o[j] = ACC ? o[j] + val + ca[j] : val;
      }
      stg16(dx + vox * lddx + c0, pack8(o));
    }
  }
}

// fp32 NCDHW [N][Cin][V] -> bf16 NDHWC [N][V][Cpad] (zero padded channels)
__global__ void __launch_bounds__(256) to_ndhwc_kernel(const float* __restrict__ x, bf16* __restrict__ out, long long ldo, int N,
                                                       int Cin, long long V, int Cpad) {
  const int C8 = Cpad >> 3;
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    for (int c8 = 0; c8 < C8; ++c8) {
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c8 * 8 + j;
        a[j] = (c < Cin) ? __ldg(x + ((long long)n * Cin + c) * V + v) : 0.f;
      }
      stg16(out + i * ldo + c8 * 8, pack8(a));
    }
  }
}

// bf16 NDHWC [N][V][C] (pitch ld) -> fp32 NCDHW [N][C][V]   (used for standalone module outputs / debugging)
__global__ void __launch_bounds__(256) to_ncdhw_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out, int N,
                                                       int C, long long V) {
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      float a[8];
      unpack8(ldg16_stream(x + i * ldx + c8 * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) out[((long long)n * C + c8 * 8 + j) * V + v] = a[j];
    }
  }
}

// sums[n][c] += Σ_v x[n][v][c]     (double atomics; caller zeroes)  — SE average pool and bias gradients
__global__ void __launch_bounds__(256) channel_sum_kernel(const bf16* __restrict__ x, long long ldx, double* __restrict__ sums,
                                                          long long V, int C) {
  extern __shared__ double red[];  // [C], fp64: thread arrival order cannot change the sums (forward SE mean)
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) red[c] = 0.0;
  __syncthreads();
  const int C8 = C >> 3;
  const bool fixed = (blockDim.x % C8) == 0;
  const bf16* xn = x + (long long)n * V * ldx;
  const long long total = V * C8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int myc0 = -1;
  if (fixed) {  // fixed channel chunk per thread: 4 independent 16-byte loads in flight
    const int c0 = (int)(threadIdx.x % C8) * 8;
    const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
    long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
    for (; vox + 3 * vstep < V; vox += 4 * vstep) {
      const uint4 u0 = ldg16_stream(xn + vox * ldx + c0), u1 = ldg16_stream(xn + (vox + vstep) * ldx + c0);
      const uint4 u2 = ldg16_stream(xn + (vox + 2 * vstep) * ldx + c0), u3 = ldg16_stream(xn + (vox + 3 * vstep) * ldx + c0);
      float a[8], b[8], c[8], d[8];
      unpack8(u0, a); unpack8(u1, b); unpack8(u2, c); unpack8(u3, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (a[j] + b[j]) + (c[j] + d[j]);
    }
    for (; vox < V; vox += vstep) {
      float a[8];
      unpack8(ldg16_stream(xn + vox * ldx + c0), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
    myc0 = c0;
  } else
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float a[8];
    unpack8(ldg16_stream(xn + vox * ldx + c0), a);
    if (fixed) {
      myc0 = c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[c0 + j], (double)a[j]);
    }
  }
  if (fixed) {   // fold the lanes of a warp that own the same chunk, then one fp64 atomic per warp and channel
    const int grp = C8 < 32 ? C8 : 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum_mod(acc[j], grp);
    if ((int)(threadIdx.x & 31) < grp) {
      const int c0 = (int)(threadIdx.x % C8) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[c0 + j], (double)acc[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&sums[(long long)n * C + c], red[c]);
}

extern "C" {

int b3d_pool_fwd(const void* x, long long ldx, const float* mask, void* out, long long ldo, int N, int D, int H, int W,
                 int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "pool_fwd: need even dims and C%%8==0");
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  pool_fwd_kernel<<<ew_blocks2(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, mask, (bf16*)out, ldo, N,
                                                                           D, H, W, C, 0); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// out = MaxPool3d(2)(ReLU(x))   (nn.ReLU + nn.MaxPool3d(2) of BrainTumorClassifier.features, main.py:307-312)
int b3d_relu_pool_fwd(const void* x, long long ldx, void* out, long long ldo, int N, int D, int H, int W, int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "relu_pool_fwd: need even dims and C%%8==0");
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  pool_fwd_kernel<<<ew_blocks2(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, nullptr, (bf16*)out, ldo, N,
                                                                           D, H, W, C, 1); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// out[n][c][oz][oy][ox] (fp32, the reference's NCDHW flatten order) = mean over the adaptive window of ReLU(x)
// window of output index o along an axis of length L with O outputs: [floor(o*L/O), ceil((o+1)*L/O))   (ATen)
int b3d_relu_adaptive_avgpool(const void* x, long long ldx, float* out, int N, int D, int H, int W, int C, int OD, int OH,
                              int OW, int relu, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && OD > 0 && OH > 0 && OW > 0 && OD <= D && OH <= H && OW <= W, "adaptive_avgpool: bad shape");
  dim3 grid(OD * OH * OW, N);
  relu_adaptive_avgpool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, out, D, H, W, C, OD, OH, OW, relu); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// cadd (optional, fp32 [N][C], accumulate mode only): a per-(sample, channel) constant added to dx in the same pass — the
// channel-attention branch of the attention gate contributes such a constant to the skip gradient (main.py:270-276,297).
int b3d_pool_bwd_add(const void* x, long long ldx, const float* mask, const void* dy, long long lddy, void* dx,
                     long long lddx, int accumulate, const float* cadd, int N, int D, int H, int W, int C, void* stream);

int b3d_pool_bwd(const void* x, long long ldx, const float* mask, const void* dy, long long lddy, void* dx,
                 long long lddx, int accumulate, int N, int D, int H, int W, int C, void* stream) {
  return b3d_pool_bwd_add(x, ldx, mask, dy, lddy, dx, lddx, accumulate, nullptr, N, D, H, W, C, stream);
}

int b3d_pool_bwd_add(const void* x, long long ldx, const float* mask, const void* dy, long long lddy, void* dx,
                     long long lddx, int accumulate, const float* cadd, int N, int D, int H, int W, int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "pool_bwd: need even dims and C%%8==0");
  B3D_REQUIRE(cadd == nullptr || accumulate, "pool_bwd: the channel constant needs accumulate mode");
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  if (accumulate) { pool_bwd_kernel<true><<<ew_blocks2(total, 256), 256, 0, (cudaStream_t)stream>>>( (const bf16*)x, ldx, mask, (const bf16*)dy, lddy, (bf16*)dx, lddx, cadd, N, D, H, W, C); ++g_b3d_launches; }
  else { pool_bwd_kernel<false><<<ew_blocks2(total, 256), 256, 0, (cudaStream_t)stream>>>( (const bf16*)x, ldx, mask, (const bf16*)dy, lddy, (bf16*)dx, lddx, nullptr, N, D, H, W, C); ++g_b3d_launches; }
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_to_ndhwc_bf16(const float* x, void* out, long long ldo, int N, int Cin, long long V, int Cpad, void* stream) {
  B3D_REQUIRE(Cpad % 8 == 0 && Cpad >= Cin, "to_ndhwc: bad Cpad");
  to_ndhwc_kernel<<<ew_blocks2((long long)N * V, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)out, ldo, N, Cin, V, Cpad); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_to_ncdhw_f32(const void* x, long long ldx, float* out, int N, int C, long long V, void* stream) {
  B3D_REQUIRE(C % 8 == 0, "to_ncdhw: C%%8");
  to_ncdhw_kernel<<<ew_blocks2((long long)N * V, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, out, N, C, V); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_channel_sum(const void* x, long long ldx, double* sums, int N, long long V, int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= 4096, "channel_sum: bad C");
  const int per_sample = std::max(1, std::min(ew_blocks2(V * (C / 8), 256 * 8), b3d_num_sms() * 4 / std::max(1, N)));
  dim3 grid(per_sample, N);
  channel_sum_kernel<<<grid, 256, C * sizeof(double), (cudaStream_t)stream>>>((const bf16*)x, ldx, sums, V, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
