// norm.cu — fused GroupNorm (+ReLU) (+GroupNorm'd residual add) forward and backward, NDHWC bf16, fp32/fp64 statistics.
//
// Reference semantics: nn.GroupNorm(G, C, eps=1e-5) + nn.ReLU in DoubleConv3D (/root/reference/main.py:215-233,235-242):
//     a1  = relu(GN8(y1))                       -> b3d_gn_apply(relu=1)
//     out = relu(GN8(y2)) + GN8'(r)             -> b3d_gn_apply(relu=1, residual r)     (no ReLU after the add)
// The per-(sample, group) sum / sum-of-squares come from the producing convolution's epilogue (conv_igemm.cu), so the
// forward is ONE read + ONE write of the tensor (HBM-bound: 2T resp. 3T bytes).
// Backward:  dz = dy * [relu mask];  dx = rstd * (dz*γ − mean_g(dz*γ) − x̂ * mean_g(dz*γ*x̂))
//   phase 1 (b3d_gn_bwd_reduce): per-(n,c) Σdz, Σdz·x̂        (one read of dy and y)
//   phase 2 (b3d_gn_bwd_apply) : dx                           (one read of dy and y, one write)
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define GN_THREADS 256
#define GN_MAXC 2048

__device__ __forceinline__ void gn_mean_rstd(const double* __restrict__ stats, int n, int G, int g, double m, float eps,
                                             float& mean, float& rstd) {
  const double s = stats[((long long)n * G + g) * 2], q = stats[((long long)n * G + g) * 2 + 1];
  const double mu = s / m;
  double var = q / m - mu * mu;
  if (var < 0) var = 0;
  mean = (float)mu;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// Group terms of the GroupNorm backward, block-cooperative: grp[2g], grp[2g+1] = Σ_{k in group g} γ_k · sums[n][k][0|1]
// (fp64).  One warp per group, lanes strided over the group's channels.  The apply kernels used to run this sum serially
// per channel in every block (C/G dependent fp64 loads deep: ~10 us of prologue at the 512-channel levels, where the whole
// tensor is 1 MB).  Ends with a __syncthreads().  G <= GN_MAXG.
#define GN_MAXG 32
__device__ __forceinline__ void gn_group_terms(const float* __restrict__ gamma, const double* __restrict__ sums, int n, int C,
                                               int G, double* grp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int cpg = C / G;
  for (int g = warp; g < G; g += nw) {
    double sb = 0, sc = 0;
    for (int k = g * cpg + lane; k < (g + 1) * cpg; k += 32) {
      const double ga = (double)gamma[k];
      sb += ga * sums[((long long)n * C + k) * 2];
      sc += ga * sums[((long long)n * C + k) * 2 + 1];
    }
    sb = warp_sum_d(sb); sc = warp_sum_d(sc);
    if (lane == 0) { grp[2 * g] = sb; grp[2 * g + 1] = sc; }
  }
  __syncthreads();
}

// out = act(GN(y)) [+ GN_r(r) | + r]
template <bool RELU, int RES>  // RES: 0 none, 1 GroupNorm'd residual, 2 plain residual
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(
    const bf16* __restrict__ y, long long ldy, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, int G, const bf16* __restrict__ r, long long ldr, const double* __restrict__ stats_r,
    const float* __restrict__ gamma_r, const float* __restrict__ beta_r, int Gr, bf16* __restrict__ out, long long ldo,
    long long V, int C, float eps) {
  extern __shared__ float sm[];
  float* sc = sm; float* sh = sm + C; float* scr = sm + 2 * C; float* shr = sm + 3 * C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd(stats, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    const float s = gamma[c] * rstd;
    sc[c] = s; sh[c] = beta[c] - mean * s;
    if (RES == 1) {
      gn_mean_rstd(stats_r, n, Gr, c / (C / Gr), (double)(C / Gr) * (double)V, eps, mean, rstd);
      const float s2 = gamma_r[c] * rstd;
      scr[c] = s2; shr[c] = beta_r[c] - mean * s2;
    }
  }
  __syncthreads();
  const int C8 = C >> 3;
  const long long total = V * C8;
  const bf16* yn = y + (long long)n * V * ldy;
  const bf16* rn = (RES != 0) ? r + (long long)n * V * ldr : nullptr;
  bf16* on = out + (long long)n * V * ldo;
  if ((blockDim.x % C8) == 0) {
    // the thread's 8-channel chunk never changes (stride is a multiple of C8): keep the coefficients in registers
    // (the shared-memory path below has 4-way bank conflicts: lanes read floats 8 apart)
    const int c0 = (int)(threadIdx.x % C8) * 8;
    float k1[8], k2[8], k3[8], k4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { k1[j] = sc[c0 + j]; k2[j] = sh[c0 + j]; k3[j] = (RES == 1) ? scr[c0 + j] : 0.f; k4[j] = (RES == 1) ? shr[c0 + j] : 0.f; }
    const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
    long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
    for (; vox + vstep < V; vox += 2 * vstep) {  // two independent voxels per iteration
      float a[8], a2[8], b[8], b2[8];
      const uint4 u0 = ldg16_stream(yn + vox * ldy + c0), u1 = ldg16_stream(yn + (vox + vstep) * ldy + c0);
      uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
      if (RES != 0) { w0 = ldg16_stream(rn + vox * ldr + c0); w1 = ldg16_stream(rn + (vox + vstep) * ldr + c0); }
      unpack8(u0, a); unpack8(u1, a2);
      if (RES != 0) { unpack8(w0, b); unpack8(w1, b2); }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(a[j], k1[j], k2[j]), t2 = fmaf(a2[j], k1[j], k2[j]);
        if (RELU) { t = fmaxf(t, 0.f); t2 = fmaxf(t2, 0.f); }
        if (RES == 1) { t += fmaf(b[j], k3[j], k4[j]); t2 += fmaf(b2[j], k3[j], k4[j]); }
        if (RES == 2) { t += b[j]; t2 += b2[j]; }
        a[j] = t; a2[j] = t2;
      }
      stg16(on + vox * ldo + c0, pack8(a));
      stg16(on + (vox + vstep) * ldo + c0, pack8(a2));
    }
    if (vox < V) {
      float a[8], b[8];
      unpack8(ldg16_stream(yn + vox * ldy + c0), a);
      if (RES != 0) unpack8(ldg16_stream(rn + vox * ldr + c0), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(a[j], k1[j], k2[j]);
        if (RELU) t = fmaxf(t, 0.f);
        if (RES == 1) t += fmaf(b[j], k3[j], k4[j]);
        if (RES == 2) t += b[j];
        a[j] = t;
      }
      stg16(on + vox * ldo + c0, pack8(a));
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float a[8];
    unpack8(ldg16_stream(yn + vox * ldy + c0), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(a[j], sc[c0 + j], sh[c0 + j]);
      if (RELU) t = fmaxf(t, 0.f);
      a[j] = t;
    }
    if (RES != 0) {
      float b[8];
      unpack8(ldg16_stream(rn + vox * ldr + c0), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += (RES == 1) ? fmaf(b[j], scr[c0 + j], shr[c0 + j]) : b[j];
    }
    stg16(on + vox * ldo + c0, pack8(a));
  }
}

// phase 1: sums[n][c][0] += Σ_v dz ; sums[n][c][1] += Σ_v dz*xhat        (dz = dy * relu-mask)
template <bool RELU>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_bwd_reduce_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ y, long long ldy,
    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, int G,
    double* __restrict__ sums, long long V, int C, float eps) {
  extern __shared__ float sm[];
  float* s_mean = sm; float* s_rstd = sm + C; float* s_g = sm + 2 * C; float* s_b = sm + 3 * C;
  double* red = reinterpret_cast<double*>(sm + 4 * C);  // [2][C] block accumulators (fp64: arrival order cannot change dX)
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd(stats, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    s_mean[c] = mean; s_rstd[c] = rstd; s_g[c] = gamma[c]; s_b[c] = beta[c];
    red[c] = 0.0; red[C + c] = 0.0;
  }
  __syncthreads();
  const int C8 = C >> 3;
  // thread owns a fixed 8-channel chunk when blockDim % C8 == 0 (C8 <= 256 and power of two in practice); otherwise
  // fall back to per-element shared atomics.
  const bool fixed = (blockDim.x % C8) == 0;
  const bf16* dyn = dy + (long long)n * V * lddy;
  const bf16* yn = y + (long long)n * V * ldy;
  const long long total = V * C8;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
  int myc0 = -1;
  if (fixed) {
    const int c0 = (int)(threadIdx.x % C8) * 8;
    float ka[8], kb[8], kg[8], kbe[8];  // xhat = x*ka + kb ; relu test on xhat*gamma + beta
#pragma unroll
    for (int j = 0; j < 8; ++j) { ka[j] = s_rstd[c0 + j]; kb[j] = -s_mean[c0 + j] * s_rstd[c0 + j]; kg[j] = s_g[c0 + j]; kbe[j] = s_b[c0 + j]; }
    const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
    long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
    for (; vox < V; vox += 2 * vstep) {
      const bool two = vox + vstep < V;
      const long long vox2 = two ? vox + vstep : vox;
      float d[8], x[8], d2[8], x2[8];
      const uint4 ud = ldg16_stream(dyn + vox * lddy + c0), ux = ldg16_stream(yn + vox * ldy + c0);
      const uint4 ud2 = ldg16_stream(dyn + vox2 * lddy + c0), ux2 = ldg16_stream(yn + vox2 * ldy + c0);
      unpack8(ud, d); unpack8(ux, x); unpack8(ud2, d2); unpack8(ux2, x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(x[j], ka[j], kb[j]), xh2 = fmaf(x2[j], ka[j], kb[j]);
        float dz = d[j], dz2 = two ? d2[j] : 0.f;
        if (RELU && fmaf(xh, kg[j], kbe[j]) <= 0.f) dz = 0.f;
        if (RELU && fmaf(xh2, kg[j], kbe[j]) <= 0.f) dz2 = 0.f;
        a1[j] += dz + dz2; a2[j] = fmaf(dz, xh, fmaf(dz2, xh2, a2[j]));
      }
    }
    myc0 = c0;
  } else
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float d[8], x[8];
    unpack8(ldg16_stream(dyn + vox * lddy + c0), d);
    unpack8(ldg16_stream(yn + vox * ldy + c0), x);
    if (fixed) {
      myc0 = c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - s_mean[c0 + j]) * s_rstd[c0 + j];
        float dz = d[j];
        if (RELU && fmaf(xh, s_g[c0 + j], s_b[c0 + j]) <= 0.f) dz = 0.f;
        a1[j] += dz; a2[j] += dz * xh;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - s_mean[c0 + j]) * s_rstd[c0 + j];
        float dz = d[j];
        if (RELU && fmaf(xh, s_g[c0 + j], s_b[c0 + j]) <= 0.f) dz = 0.f;
        atomicAdd(&red[c0 + j], (double)dz); atomicAdd(&red[C + c0 + j], (double)(dz * xh));
      }
    }
  }
  if (fixed) {   // lanes l, l + C8, ... of a warp own the same chunk: fold them with shuffles, one fp64 atomic per warp and channel
    const int grp = C8 < 32 ? C8 : 32;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a1[j] = warp_sum_mod(a1[j], grp); a2[j] = warp_sum_mod(a2[j], grp); }
    if (lane < grp) {
      const int c0 = (int)(threadIdx.x % C8) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&red[c0 + j], (double)a1[j]); atomicAdd(&red[C + c0 + j], (double)a2[j]); }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&sums[((long long)n * C + c) * 2], red[c]);
    atomicAdd(&sums[((long long)n * C + c) * 2 + 1], red[C + c]);
  }
}

// phase 2: dx = A_c*dz − B_g − xhat*Cg ; optionally accumulates into dx (dx += ...)
template <bool RELU, bool ACC>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_bwd_apply_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ y, long long ldy,
    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, int G,
    const double* __restrict__ sums, bf16* __restrict__ dx, long long lddx, long long V, int C, float eps) {
  extern __shared__ float sm[];
  float* s_mean = sm; float* s_rstd = sm + C; float* s_g = sm + 2 * C; float* s_b = sm + 3 * C;
  float* s_B = sm + 4 * C; float* s_C = sm + 5 * C;  // per channel copies of the group terms
  const int n = blockIdx.y;
  const int cpg = C / G;
  const double m = (double)cpg * (double)V;
  __shared__ double s_grp[2 * GN_MAXG];
  gn_group_terms(gamma, sums, n, C, G, s_grp);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    const int g = c / cpg;
    gn_mean_rstd(stats, n, G, g, m, eps, mean, rstd);
    const double sb = s_grp[2 * g], sc2 = s_grp[2 * g + 1];
    s_mean[c] = mean; s_rstd[c] = rstd; s_g[c] = gamma[c]; s_b[c] = beta[c];
    s_B[c] = (float)(sb / m) * rstd; s_C[c] = (float)(sc2 / m) * rstd;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const long long total = V * C8;
  const bf16* dyn = dy + (long long)n * V * lddy;
  const bf16* yn = y + (long long)n * V * ldy;
  bf16* dxn = dx + (long long)n * V * lddx;
  if ((blockDim.x % C8) == 0) {
    const int c0 = (int)(threadIdx.x % C8) * 8;
    float ka[8], kb[8], kg[8], kbe[8], kA[8], kB[8], kC[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ka[j] = s_rstd[c0 + j]; kb[j] = -s_mean[c0 + j] * s_rstd[c0 + j]; kg[j] = s_g[c0 + j]; kbe[j] = s_b[c0 + j];
      kA[j] = s_g[c0 + j] * s_rstd[c0 + j]; kB[j] = s_B[c0 + j]; kC[j] = s_C[c0 + j];
    }
    const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
    long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
    for (; vox < V; vox += 2 * vstep) {
      const bool two = vox + vstep < V;
      const long long vox2 = two ? vox + vstep : vox;
      float d[8], x[8], o[8], d2[8], x2[8], o2[8];
      const uint4 ud = ldg16_stream(dyn + vox * lddy + c0), ux = ldg16_stream(yn + vox * ldy + c0);
      const uint4 ud2 = ldg16_stream(dyn + vox2 * lddy + c0), ux2 = ldg16_stream(yn + vox2 * ldy + c0);
      if (ACC) { unpack8(ldg16(dxn + vox * lddx + c0), o); unpack8(ldg16(dxn + vox2 * lddx + c0), o2); }
      unpack8(ud, d); unpack8(ux, x); unpack8(ud2, d2); unpack8(ux2, x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(x[j], ka[j], kb[j]), xh2 = fmaf(x2[j], ka[j], kb[j]);
        float dz = d[j], dz2 = d2[j];
        if (RELU && fmaf(xh, kg[j], kbe[j]) <= 0.f) dz = 0.f;
        if (RELU && fmaf(xh2, kg[j], kbe[j]) <= 0.f) dz2 = 0.f;
        const float v = dz * kA[j] - kB[j] - xh * kC[j], v2 = dz2 * kA[j] - kB[j] - xh2 * kC[j];
        o[j] = ACC ? o[j] + v : v; o2[j] = ACC ? o2[j] + v2 : v2;
      }
      stg16(dxn + vox * lddx + c0, pack8(o));
      if (two) stg16(dxn + vox2 * lddx + c0, pack8(o2));
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float d[8], x[8], o[8];
    unpack8(ldg16_stream(dyn + vox * lddy + c0), d);
    unpack8(ldg16_stream(yn + vox * ldy + c0), x);
    if (ACC) unpack8(ldg16(dxn + vox * lddx + c0), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - s_mean[c0 + j]) * s_rstd[c0 + j];
      float dz = d[j];
      if (RELU && fmaf(xh, s_g[c0 + j], s_b[c0 + j]) <= 0.f) dz = 0.f;
      const float v = dz * s_g[c0 + j] * s_rstd[c0 + j] - s_B[c0 + j] - xh * s_C[c0 + j];
      o[j] = ACC ? o[j] + v : v;
    }
    stg16(dxn + vox * lddx + c0, pack8(o));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Dual GroupNorm backward of a residual block's tail:  out = relu(GN_a(ya)) + GN_b(yb)   (main.py:238-240)
// Both branches receive the same upstream gradient dy; the separate kernels read it four times (2 x reduce, 2 x apply), the
// dual ones twice: 8 instead of 10 tensor passes per block.  Fixed 8-channel chunk per thread (blockDim % (C/8) == 0),
// algebraically folded per-channel coefficients so that both branches' constants fit in registers:
//   x̂ = x*ka + kb ;  relu test: x*S + T <= 0 ;  dx = dz*P - Q - x*R   with P = γ*rstd, R = ka*Cg, Q = Bg + kb*Cg
//   (Bg, Cg = rstd * mean_g(γ·Σdz), rstd * mean_g(γ·Σdz·x̂): the group terms of the standard formula).
// ------------------------------------------------------------------------------------------------------------------
template <int OCC>
__global__ void __launch_bounds__(GN_THREADS, OCC) gn_bwd_dual_reduce_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ ya, long long ldya, const double* __restrict__ stats_a,
    const float* __restrict__ gamma_a, const float* __restrict__ beta_a, const bf16* __restrict__ yb, long long ldyb,
    const double* __restrict__ stats_b, int G, double* __restrict__ sums_a, double* __restrict__ sums_b, long long V, int C,
    float eps) {
  extern __shared__ float sm[];
  float* c_ka = sm; float* c_kb = sm + C; float* c_S = sm + 2 * C; float* c_T = sm + 3 * C;
  float* c_kab = sm + 4 * C; float* c_kbb = sm + 5 * C;
  double* red = reinterpret_cast<double*>(sm + 6 * C);   // [4][C]: Σdz_a, Σdz_a·x̂a, Σdy, Σdy·x̂b
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd(stats_a, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    c_ka[c] = rstd; c_kb[c] = -mean * rstd;
    { const float sa = gamma_a[c] * rstd; c_S[c] = sa; c_T[c] = beta_a[c] - mean * sa; }   // the forward's own expression (gn_apply_kernel)
    gn_mean_rstd(stats_b, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    c_kab[c] = rstd; c_kbb[c] = -mean * rstd;
    red[c] = 0.0; red[C + c] = 0.0; red[2 * C + c] = 0.0; red[3 * C + c] = 0.0;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy;
  const bf16* yan = ya + (long long)n * V * ldya;
  const bf16* ybn = yb + (long long)n * V * ldyb;
  float ka[8], kb[8], kS[8], kT[8], kab[8], kbb[8], a1[8], a2[8], a3[8], a4[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ka[j] = c_ka[c0 + j]; kb[j] = c_kb[c0 + j]; kS[j] = c_S[c0 + j]; kT[j] = c_T[c0 + j];
    kab[j] = c_kab[c0 + j]; kbb[j] = c_kbb[c0 + j];
    a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; a4[j] = 0.f;
  }
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += 2 * vstep) {
    const bool two = vox + vstep < V;
    const long long vox2 = two ? vox + vstep : vox;
    const uint4 ud = ldg16_stream(dyn + vox * lddy + c0), ua = ldg16_stream(yan + vox * ldya + c0), ub = ldg16_stream(ybn + vox * ldyb + c0);
    const uint4 ud2 = ldg16_stream(dyn + vox2 * lddy + c0), ua2 = ldg16_stream(yan + vox2 * ldya + c0), ub2 = ldg16_stream(ybn + vox2 * ldyb + c0);
    float d[8], x[8], z[8];
    unpack8(ud, d); unpack8(ua, x); unpack8(ub, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = (fmaf(x[j], kS[j], kT[j]) <= 0.f) ? 0.f : d[j];
      a1[j] += dz; a2[j] = fmaf(dz, fmaf(x[j], ka[j], kb[j]), a2[j]);
      a3[j] += d[j]; a4[j] = fmaf(d[j], fmaf(z[j], kab[j], kbb[j]), a4[j]);
    }
    if (two) {
      unpack8(ud2, d); unpack8(ua2, x); unpack8(ub2, z);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dz = (fmaf(x[j], kS[j], kT[j]) <= 0.f) ? 0.f : d[j];
        a1[j] += dz; a2[j] = fmaf(dz, fmaf(x[j], ka[j], kb[j]), a2[j]);
        a3[j] += d[j]; a4[j] = fmaf(d[j], fmaf(z[j], kab[j], kbb[j]), a4[j]);
      }
    }
  }
  const int grp = C8 < 32 ? C8 : 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = warp_sum_mod(a1[j], grp); a2[j] = warp_sum_mod(a2[j], grp);
    a3[j] = warp_sum_mod(a3[j], grp); a4[j] = warp_sum_mod(a4[j], grp);
  }
  if ((int)(threadIdx.x & 31) < grp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[c0 + j], (double)a1[j]); atomicAdd(&red[C + c0 + j], (double)a2[j]);
      atomicAdd(&red[2 * C + c0 + j], (double)a3[j]); atomicAdd(&red[3 * C + c0 + j], (double)a4[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&sums_a[((long long)n * C + c) * 2], red[c]);
    atomicAdd(&sums_a[((long long)n * C + c) * 2 + 1], red[C + c]);
    atomicAdd(&sums_b[((long long)n * C + c) * 2], red[2 * C + c]);
    atomicAdd(&sums_b[((long long)n * C + c) * 2 + 1], red[3 * C + c]);
  }
}

template <int OCC>
__global__ void __launch_bounds__(GN_THREADS, OCC) gn_bwd_dual_apply_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ ya, long long ldya, const double* __restrict__ stats_a,
    const float* __restrict__ gamma_a, const float* __restrict__ beta_a, const double* __restrict__ sums_a,
    const bf16* __restrict__ yb, long long ldyb, const double* __restrict__ stats_b, const float* __restrict__ gamma_b,
    const double* __restrict__ sums_b, int G, bf16* __restrict__ dxa, long long lddxa, bf16* __restrict__ dxb, long long lddxb,
    long long V, int C, float eps) {
  extern __shared__ float sm[];
  float* c_S = sm; float* c_T = sm + C; float* c_Pa = sm + 2 * C; float* c_Qa = sm + 3 * C; float* c_Ra = sm + 4 * C;
  float* c_Pb = sm + 5 * C; float* c_Qb = sm + 6 * C; float* c_Rb = sm + 7 * C;
  const int n = blockIdx.y;
  const int cpg = C / G;
  const double m = (double)cpg * (double)V;
  __shared__ double s_grp[4 * GN_MAXG];
  gn_group_terms(gamma_a, sums_a, n, C, G, s_grp);
  gn_group_terms(gamma_b, sums_b, n, C, G, s_grp + 2 * GN_MAXG);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float mean, rstd;
    gn_mean_rstd(stats_a, n, G, g, m, eps, mean, rstd);
    double sb = s_grp[2 * g], sc2 = s_grp[2 * g + 1];
    float Bg = (float)(sb / m) * rstd, Cg = (float)(sc2 / m) * rstd, kb = -mean * rstd;
    { const float sa = gamma_a[c] * rstd; c_S[c] = sa; c_T[c] = beta_a[c] - mean * sa; }   // the forward's own expression
    c_Pa[c] = gamma_a[c] * rstd; c_Qa[c] = fmaf(kb, Cg, Bg); c_Ra[c] = rstd * Cg;
    gn_mean_rstd(stats_b, n, G, g, m, eps, mean, rstd);
    sb = s_grp[2 * GN_MAXG + 2 * g]; sc2 = s_grp[2 * GN_MAXG + 2 * g + 1];
    Bg = (float)(sb / m) * rstd; Cg = (float)(sc2 / m) * rstd; kb = -mean * rstd;
    c_Pb[c] = gamma_b[c] * rstd; c_Qb[c] = fmaf(kb, Cg, Bg); c_Rb[c] = rstd * Cg;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy;
  const bf16* yan = ya + (long long)n * V * ldya;
  const bf16* ybn = yb + (long long)n * V * ldyb;
  bf16* dxan = dxa + (long long)n * V * lddxa;
  bf16* dxbn = dxb + (long long)n * V * lddxb;
  float kS[8], kT[8], Pa[8], Qa[8], Ra[8], Pb[8], Qb[8], Rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    kS[j] = c_S[c0 + j]; kT[j] = c_T[c0 + j]; Pa[j] = c_Pa[c0 + j]; Qa[j] = c_Qa[c0 + j]; Ra[j] = c_Ra[c0 + j];
    Pb[j] = c_Pb[c0 + j]; Qb[j] = c_Qb[c0 + j]; Rb[j] = c_Rb[c0 + j];
  }
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += 2 * vstep) {
    const bool two = vox + vstep < V;
    const long long vox2 = two ? vox + vstep : vox;
    const uint4 ud = ldg16_stream(dyn + vox * lddy + c0), ua = ldg16_stream(yan + vox * ldya + c0), ub = ldg16_stream(ybn + vox * ldyb + c0);
    const uint4 ud2 = ldg16_stream(dyn + vox2 * lddy + c0), ua2 = ldg16_stream(yan + vox2 * ldya + c0), ub2 = ldg16_stream(ybn + vox2 * ldyb + c0);
    float d[8], x[8], z[8], oa[8], ob[8];
    unpack8(ud, d); unpack8(ua, x); unpack8(ub, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = (fmaf(x[j], kS[j], kT[j]) <= 0.f) ? 0.f : d[j];
      oa[j] = fmaf(dz, Pa[j], -fmaf(x[j], Ra[j], Qa[j]));
      ob[j] = fmaf(d[j], Pb[j], -fmaf(z[j], Rb[j], Qb[j]));
    }
    stg16(dxan + vox * lddxa + c0, pack8(oa));
    stg16(dxbn + vox * lddxb + c0, pack8(ob));
    if (two) {
      unpack8(ud2, d); unpack8(ua2, x); unpack8(ub2, z);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dz = (fmaf(x[j], kS[j], kT[j]) <= 0.f) ? 0.f : d[j];
        oa[j] = fmaf(dz, Pa[j], -fmaf(x[j], Ra[j], Qa[j]));
        ob[j] = fmaf(d[j], Pb[j], -fmaf(z[j], Rb[j], Qb[j]));
      }
      stg16(dxan + vox2 * lddxa + c0, pack8(oa));
      stg16(dxbn + vox2 * lddxb + c0, pack8(ob));
    }
  }
}

// ---- register-lean variants of the two dual kernels --------------------------------------------------------------------
// The per-channel coefficients stay in shared memory and are re-read (volatile 8-byte loads, two channels at a time) inside
// the loop instead of living in 48-64 registers per thread: ~60 registers -> four resident blocks per SM instead of two,
// i.e. twice the loads in flight for a kernel that is purely latency x bandwidth bound (ncu round 2: 126 registers, 25 % warps
// active, 4.6 TB/s).  Arithmetic expressions are the same as in the kernels above (bit-identical results).
// Coefficient layout: [kind][chunk][8 + 2 pad] floats — the lanes of a warp (different chunks) read 8-byte words 40 bytes
// apart, which spreads 16 chunks over all 32 banks (a plain [kind][channel] layout is C/32-way conflicted: 16-way at C = 512).
// GN_KS(C) = floats per kind.
#define GN_KS(C) ((C) + ((C) >> 2))
__device__ __forceinline__ int kidx(int kind, int c, int C) { return kind * GN_KS(C) + (c >> 3) * 10 + (c & 7); }

template <int UB>
__global__ void __launch_bounds__(GN_THREADS, 4) gn_bwd_dual_apply_lean_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ ya, long long ldya, const double* __restrict__ stats_a,
    const float* __restrict__ gamma_a, const float* __restrict__ beta_a, const double* __restrict__ sums_a,
    const bf16* __restrict__ yb, long long ldyb, const double* __restrict__ stats_b, const float* __restrict__ gamma_b,
    const double* __restrict__ sums_b, int G, bf16* __restrict__ dxa, long long lddxa, bf16* __restrict__ dxb, long long lddxb,
    long long V, int C, float eps) {
  extern __shared__ float sm[];
  constexpr int c_S = 0; constexpr int c_T = 1; constexpr int c_Pa = 2; constexpr int c_Qa = 3; constexpr int c_Ra = 4;
  constexpr int c_Pb = 5; constexpr int c_Qb = 6; constexpr int c_Rb = 7;
  const int n = blockIdx.y;
  const int cpg = C / G;
  const double m = (double)cpg * (double)V;
  __shared__ double s_grp[4 * GN_MAXG];
  gn_group_terms(gamma_a, sums_a, n, C, G, s_grp);
  gn_group_terms(gamma_b, sums_b, n, C, G, s_grp + 2 * GN_MAXG);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float mean, rstd;
    gn_mean_rstd(stats_a, n, G, g, m, eps, mean, rstd);
    double sb = s_grp[2 * g], sc2 = s_grp[2 * g + 1];
    float Bg = (float)(sb / m) * rstd, Cg = (float)(sc2 / m) * rstd, kb = -mean * rstd;
    { const float sa = gamma_a[c] * rstd; sm[kidx(c_S, c, C)] = sa; sm[kidx(c_T, c, C)] = beta_a[c] - mean * sa; }
    sm[kidx(c_Pa, c, C)] = gamma_a[c] * rstd; sm[kidx(c_Qa, c, C)] = fmaf(kb, Cg, Bg); sm[kidx(c_Ra, c, C)] = rstd * Cg;
    gn_mean_rstd(stats_b, n, G, g, m, eps, mean, rstd);
    sb = s_grp[2 * GN_MAXG + 2 * g]; sc2 = s_grp[2 * GN_MAXG + 2 * g + 1];
    Bg = (float)(sb / m) * rstd; Cg = (float)(sc2 / m) * rstd; kb = -mean * rstd;
    sm[kidx(c_Pb, c, C)] = gamma_b[c] * rstd; sm[kidx(c_Qb, c, C)] = fmaf(kb, Cg, Bg); sm[kidx(c_Rb, c, C)] = rstd * Cg;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy + c0;
  const bf16* yan = ya + (long long)n * V * ldya + c0;
  const bf16* ybn = yb + (long long)n * V * ldyb + c0;
  bf16* dxan = dxa + (long long)n * V * lddxa + c0;
  bf16* dxbn = dxb + (long long)n * V * lddxb + c0;
  const float* kc = sm + (int)(threadIdx.x % C8) * 10;   // padded coefficient layout (kidx)
  const int KS = GN_KS(C);
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += UB * vstep) {
    unsigned d[UB][4], x[UB][4], z[UB][4];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const long long vu = (vox + u * vstep < V) ? vox + u * vstep : vox;
      const uint4 t0 = ldg16_stream(dyn + vu * lddy), t1 = ldg16_stream(yan + vu * ldya), t2 = ldg16_stream(ybn + vu * ldyb);
      d[u][0] = t0.x; d[u][1] = t0.y; d[u][2] = t0.z; d[u][3] = t0.w;
      x[u][0] = t1.x; x[u][1] = t1.y; x[u][2] = t1.z; x[u][3] = t1.w;
      z[u][0] = t2.x; z[u][1] = t2.y; z[u][2] = t2.z; z[u][3] = t2.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 S = lds2v(kc + 0 * KS + 2 * q), T = lds2v(kc + 1 * KS + 2 * q), Pa = lds2v(kc + 2 * KS + 2 * q),
                   Qa = lds2v(kc + 3 * KS + 2 * q), Ra = lds2v(kc + 4 * KS + 2 * q), Pb = lds2v(kc + 5 * KS + 2 * q),
                   Qb = lds2v(kc + 6 * KS + 2 * q), Rb = lds2v(kc + 7 * KS + 2 * q);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const float2 dd = bfw(d[u][q]), xx = bfw(x[u][q]), zz = bfw(z[u][q]);
        const float dz0 = (fmaf(xx.x, S.x, T.x) <= 0.f) ? 0.f : dd.x;
        const float dz1 = (fmaf(xx.y, S.y, T.y) <= 0.f) ? 0.f : dd.y;
        // d[] / x[] words are dead from here on: reuse them for the packed results
        d[u][q] = wbf(fmaf(dz0, Pa.x, -fmaf(xx.x, Ra.x, Qa.x)), fmaf(dz1, Pa.y, -fmaf(xx.y, Ra.y, Qa.y)));
        x[u][q] = wbf(fmaf(dd.x, Pb.x, -fmaf(zz.x, Rb.x, Qb.x)), fmaf(dd.y, Pb.y, -fmaf(zz.y, Rb.y, Qb.y)));
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const long long vu = vox + u * vstep;
      if (vu < V) {
        stg16(dxan + vu * lddxa, make_uint4(d[u][0], d[u][1], d[u][2], d[u][3]));
        stg16(dxbn + vu * lddxb, make_uint4(x[u][0], x[u][1], x[u][2], x[u][3]));
      }
    }
  }
}

template <int UB>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_bwd_dual_reduce_lean_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ ya, long long ldya, const double* __restrict__ stats_a,
    const float* __restrict__ gamma_a, const float* __restrict__ beta_a, const bf16* __restrict__ yb, long long ldyb,
    const double* __restrict__ stats_b, int G, double* __restrict__ sums_a, double* __restrict__ sums_b, long long V, int C,
    float eps) {
  extern __shared__ float sm[];
  constexpr int c_ka = 0; constexpr int c_kb = 1; constexpr int c_S = 2; constexpr int c_T = 3;
  constexpr int c_kab = 4; constexpr int c_kbb = 5;
  double* red = reinterpret_cast<double*>(sm + 6 * GN_KS(C));   // [4][C]: Σdz_a, Σdz_a·x̂a, Σdy, Σdy·x̂b
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd(stats_a, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    sm[kidx(c_ka, c, C)] = rstd; sm[kidx(c_kb, c, C)] = -mean * rstd;
    { const float sa = gamma_a[c] * rstd; sm[kidx(c_S, c, C)] = sa; sm[kidx(c_T, c, C)] = beta_a[c] - mean * sa; }
    gn_mean_rstd(stats_b, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    sm[kidx(c_kab, c, C)] = rstd; sm[kidx(c_kbb, c, C)] = -mean * rstd;
    red[c] = 0.0; red[C + c] = 0.0; red[2 * C + c] = 0.0; red[3 * C + c] = 0.0;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy + c0;
  const bf16* yan = ya + (long long)n * V * ldya + c0;
  const bf16* ybn = yb + (long long)n * V * ldyb + c0;
  const float* kc = sm + (int)(threadIdx.x % C8) * 10;   // padded coefficient layout (kidx)
  const int KS = GN_KS(C);
  float a1[8], a2[8], a3[8], a4[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; a4[j] = 0.f; }
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += UB * vstep) {
    unsigned d[UB][4], x[UB][4], z[UB][4];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const bool ok = vox + u * vstep < V;
      const long long vu = ok ? vox + u * vstep : vox;
      uint4 t0 = ldg16_stream(dyn + vu * lddy);
      const uint4 t1 = ldg16_stream(yan + vu * ldya), t2 = ldg16_stream(ybn + vu * ldyb);
      if (!ok) t0 = make_uint4(0u, 0u, 0u, 0u);   // a zero gradient contributes nothing to any of the four sums
      d[u][0] = t0.x; d[u][1] = t0.y; d[u][2] = t0.z; d[u][3] = t0.w;
      x[u][0] = t1.x; x[u][1] = t1.y; x[u][2] = t1.z; x[u][3] = t1.w;
      z[u][0] = t2.x; z[u][1] = t2.y; z[u][2] = t2.z; z[u][3] = t2.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 ka = lds2v(kc + 0 * KS + 2 * q), kb = lds2v(kc + 1 * KS + 2 * q), S = lds2v(kc + 2 * KS + 2 * q),
                   T = lds2v(kc + 3 * KS + 2 * q), kab = lds2v(kc + 4 * KS + 2 * q), kbb = lds2v(kc + 5 * KS + 2 * q);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const float2 dd = bfw(d[u][q]), xx = bfw(x[u][q]), zz = bfw(z[u][q]);
        const float dz0 = (fmaf(xx.x, S.x, T.x) <= 0.f) ? 0.f : dd.x;
        const float dz1 = (fmaf(xx.y, S.y, T.y) <= 0.f) ? 0.f : dd.y;
        a1[2 * q] += dz0; a2[2 * q] = fmaf(dz0, fmaf(xx.x, ka.x, kb.x), a2[2 * q]);
        a3[2 * q] += dd.x; a4[2 * q] = fmaf(dd.x, fmaf(zz.x, kab.x, kbb.x), a4[2 * q]);
        a1[2 * q + 1] += dz1; a2[2 * q + 1] = fmaf(dz1, fmaf(xx.y, ka.y, kb.y), a2[2 * q + 1]);
        a3[2 * q + 1] += dd.y; a4[2 * q + 1] = fmaf(dd.y, fmaf(zz.y, kab.y, kbb.y), a4[2 * q + 1]);
      }
    }
  }
  const int grp = C8 < 32 ? C8 : 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = warp_sum_mod(a1[j], grp); a2[j] = warp_sum_mod(a2[j], grp);
    a3[j] = warp_sum_mod(a3[j], grp); a4[j] = warp_sum_mod(a4[j], grp);
  }
  if ((int)(threadIdx.x & 31) < grp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[c0 + j], (double)a1[j]); atomicAdd(&red[C + c0 + j], (double)a2[j]);
      atomicAdd(&red[2 * C + c0 + j], (double)a3[j]); atomicAdd(&red[3 * C + c0 + j], (double)a4[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&sums_a[((long long)n * C + c) * 2], red[c]);
    atomicAdd(&sums_a[((long long)n * C + c) * 2 + 1], red[C + c]);
    atomicAdd(&sums_b[((long long)n * C + c) * 2], red[2 * C + c]);
    atomicAdd(&sums_b[((long long)n * C + c) * 2 + 1], red[3 * C + c]);
  }
}

// register-lean single-branch kernels (fixed 8-channel chunk per thread, i.e. blockDim % (C/8) == 0): same scheme as the dual
// ones above.  UB voxels in flight per thread; launch bounds follow the register budget of UB.
template <bool RELU, int UB>
__global__ void __launch_bounds__(GN_THREADS, UB >= 4 ? 3 : 4) gn_bwd_reduce_lean_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ y, long long ldy,
    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, int G,
    double* __restrict__ sums, long long V, int C, float eps) {
  extern __shared__ float sm[];
  constexpr int c_ka = 0; constexpr int c_kb = 1; constexpr int c_kg = 2; constexpr int c_kbe = 3;
  double* red = reinterpret_cast<double*>(sm + 4 * GN_KS(C));
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd(stats, n, G, c / (C / G), (double)(C / G) * (double)V, eps, mean, rstd);
    sm[kidx(c_ka, c, C)] = rstd; sm[kidx(c_kb, c, C)] = -mean * rstd; sm[kidx(c_kg, c, C)] = gamma[c]; sm[kidx(c_kbe, c, C)] = beta[c];
    red[c] = 0.0; red[C + c] = 0.0;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy + c0;
  const bf16* yn = y + (long long)n * V * ldy + c0;
  const float* kc = sm + (int)(threadIdx.x % C8) * 10;   // padded coefficient layout (kidx)
  const int KS = GN_KS(C);
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += UB * vstep) {
    unsigned d[UB][4], x[UB][4];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const bool ok = vox + u * vstep < V;
      const long long vu = ok ? vox + u * vstep : vox;
      uint4 t0 = ldg16_stream(dyn + vu * lddy);
      const uint4 t1 = ldg16_stream(yn + vu * ldy);
      if (!ok) t0 = make_uint4(0u, 0u, 0u, 0u);
      d[u][0] = t0.x; d[u][1] = t0.y; d[u][2] = t0.z; d[u][3] = t0.w;
      x[u][0] = t1.x; x[u][1] = t1.y; x[u][2] = t1.z; x[u][3] = t1.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 ka = lds2v(kc + 0 * KS + 2 * q), kb = lds2v(kc + 1 * KS + 2 * q);
      float2 kg = make_float2(0.f, 0.f), kbe = kg;
      if (RELU) { kg = lds2v(kc + 2 * KS + 2 * q); kbe = lds2v(kc + 3 * KS + 2 * q); }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const float2 dd = bfw(d[u][q]), xx = bfw(x[u][q]);
        const float xh0 = fmaf(xx.x, ka.x, kb.x), xh1 = fmaf(xx.y, ka.y, kb.y);
        float dz0 = dd.x, dz1 = dd.y;
        if (RELU && fmaf(xh0, kg.x, kbe.x) <= 0.f) dz0 = 0.f;
        if (RELU && fmaf(xh1, kg.y, kbe.y) <= 0.f) dz1 = 0.f;
        a1[2 * q] += dz0; a2[2 * q] = fmaf(dz0, xh0, a2[2 * q]);
        a1[2 * q + 1] += dz1; a2[2 * q + 1] = fmaf(dz1, xh1, a2[2 * q + 1]);
      }
    }
  }
  const int grp = C8 < 32 ? C8 : 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = warp_sum_mod(a1[j], grp); a2[j] = warp_sum_mod(a2[j], grp); }
  if ((int)(threadIdx.x & 31) < grp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&red[c0 + j], (double)a1[j]); atomicAdd(&red[C + c0 + j], (double)a2[j]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&sums[((long long)n * C + c) * 2], red[c]);
    atomicAdd(&sums[((long long)n * C + c) * 2 + 1], red[C + c]);
  }
}

template <bool RELU, bool ACC, int UB>
__global__ void __launch_bounds__(GN_THREADS, 4) gn_bwd_apply_lean_kernel(
    const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ y, long long ldy,
    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, int G,
    const double* __restrict__ sums, bf16* __restrict__ dx, long long lddx, long long V, int C, float eps) {
  extern __shared__ float sm[];
  constexpr int c_ka = 0; constexpr int c_kb = 1; constexpr int c_kg = 2; constexpr int c_kbe = 3;
  constexpr int c_B = 4; constexpr int c_C = 5;
  const int n = blockIdx.y;
  const int cpg = C / G;
  const double m = (double)cpg * (double)V;
  __shared__ double s_grp[2 * GN_MAXG];
  gn_group_terms(gamma, sums, n, C, G, s_grp);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    const int g = c / cpg;
    gn_mean_rstd(stats, n, G, g, m, eps, mean, rstd);
    const double sb = s_grp[2 * g], sc2 = s_grp[2 * g + 1];
    sm[kidx(c_ka, c, C)] = rstd; sm[kidx(c_kb, c, C)] = -mean * rstd; sm[kidx(c_kg, c, C)] = gamma[c]; sm[kidx(c_kbe, c, C)] = beta[c];
    sm[kidx(c_B, c, C)] = (float)(sb / m) * rstd; sm[kidx(c_C, c, C)] = (float)(sc2 / m) * rstd;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int c0 = (int)(threadIdx.x % C8) * 8;
  const bf16* dyn = dy + (long long)n * V * lddy + c0;
  const bf16* yn = y + (long long)n * V * ldy + c0;
  bf16* dxn = dx + (long long)n * V * lddx + c0;
  const float* kc = sm + (int)(threadIdx.x % C8) * 10;   // padded coefficient layout (kidx)
  const int KS = GN_KS(C);
  const long long vstep = ((long long)gridDim.x * blockDim.x) / C8;
  long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / C8;
  for (; vox < V; vox += UB * vstep) {
    unsigned d[UB][4], x[UB][4], o[UB][4];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const long long vu = (vox + u * vstep < V) ? vox + u * vstep : vox;
      const uint4 t0 = ldg16_stream(dyn + vu * lddy), t1 = ldg16_stream(yn + vu * ldy);
      d[u][0] = t0.x; d[u][1] = t0.y; d[u][2] = t0.z; d[u][3] = t0.w;
      x[u][0] = t1.x; x[u][1] = t1.y; x[u][2] = t1.z; x[u][3] = t1.w;
      if (ACC) { const uint4 t2 = ldg16(dxn + vu * lddx); o[u][0] = t2.x; o[u][1] = t2.y; o[u][2] = t2.z; o[u][3] = t2.w; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 ka = lds2v(kc + 0 * KS + 2 * q), kb = lds2v(kc + 1 * KS + 2 * q), kg = lds2v(kc + 2 * KS + 2 * q),
                   kbe = lds2v(kc + 3 * KS + 2 * q), kB = lds2v(kc + 4 * KS + 2 * q), kC = lds2v(kc + 5 * KS + 2 * q);
      const float kA0 = kg.x * ka.x, kA1 = kg.y * ka.y;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const float2 dd = bfw(d[u][q]), xx = bfw(x[u][q]);
        const float xh0 = fmaf(xx.x, ka.x, kb.x), xh1 = fmaf(xx.y, ka.y, kb.y);
        float dz0 = dd.x, dz1 = dd.y;
        if (RELU && fmaf(xh0, kg.x, kbe.x) <= 0.f) dz0 = 0.f;
        if (RELU && fmaf(xh1, kg.y, kbe.y) <= 0.f) dz1 = 0.f;
        float v0 = dz0 * kA0 - kB.x - xh0 * kC.x, v1 = dz1 * kA1 - kB.y - xh1 * kC.y;
        if (ACC) { const float2 oo = bfw(o[u][q]); v0 += oo.x; v1 += oo.y; }
        d[u][q] = wbf(v0, v1);
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const long long vu = vox + u * vstep;
      if (vu < V) stg16(dxn + vu * lddx, make_uint4(d[u][0], d[u][1], d[u][2], d[u][3]));
    }
  }
}

// dgamma[c] = Σ_n sums[n][c][1], dbeta[c] = Σ_n sums[n][c][0]
__global__ void gn_param_grad_kernel(const double* __restrict__ sums, int N, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0, b = 0;
  for (int n = 0; n < N; ++n) { b += sums[((long long)n * C + c) * 2]; a += sums[((long long)n * C + c) * 2 + 1]; }
  if (accumulate) { dgamma[c] += (float)a; dbeta[c] += (float)b; }
  else { dgamma[c] = (float)a; dbeta[c] = (float)b; }
}

// out = a + b (bf16, pitched) — gradient accumulation of two branches
__global__ void __launch_bounds__(GN_THREADS) add_kernel(const bf16* __restrict__ a, long long lda,
                                                         const bf16* __restrict__ b, long long ldb, bf16* __restrict__ o,
                                                         long long ldo, long long V, int C) {
  const int C8 = C >> 3;
  const long long total = V * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float x[8], y[8];
    unpack8(ldg16_stream(a + vox * lda + c0), x);
    unpack8(ldg16_stream(b + vox * ldb + c0), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    stg16(o + vox * ldo + c0, pack8(x));
  }
}

static int ew_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)b3d_num_sms() * 16;
  return (int)std::max<long long>(1, std::min(b, cap));
}

extern "C" {

int b3d_gn_apply(const void* y, long long ldy, const double* stats, const float* gamma, const float* beta, int G,
                 int relu, int res_mode, const void* r, long long ldr, const double* stats_r, const float* gamma_r,
                 const float* beta_r, int Gr, void* out, long long ldo, int N, long long V, int C, float eps,
                 void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= GN_MAXC && C % G == 0, "gn_apply: bad C=%d G=%d", C, G);
  B3D_REQUIRE(ldy % 8 == 0 && ldo % 8 == 0, "gn_apply: pitches must be multiples of 8");
  B3D_REQUIRE(res_mode >= 0 && res_mode <= 2, "gn_apply: bad res_mode");
  if (res_mode == 1) B3D_REQUIRE(C % Gr == 0, "gn_apply: bad Gr");
  dim3 grid(ew_blocks(V * (C / 8), GN_THREADS * 2), N);
  const size_t smem = 4 * (size_t)C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(RL, RS)                                                                                              \
  gn_apply_kernel<RL, RS><<<grid, GN_THREADS, smem, st>>>((const bf16*)y, ldy, stats, gamma, beta, G, (const bf16*)r, ldr, \
                                                          stats_r, gamma_r, beta_r, Gr, (bf16*)out, ldo, V, C, eps)
  if (relu) { if (res_mode == 0) LAUNCH(true, 0); else if (res_mode == 1) LAUNCH(true, 1); else LAUNCH(true, 2); }
  else { if (res_mode == 0) LAUNCH(false, 0); else if (res_mode == 1) LAUNCH(false, 1); else LAUNCH(false, 2); }
#undef LAUNCH
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// sums: double [N][C][2], must be zeroed by the caller before the call (accumulated with atomics)
int b3d_gn_bwd_reduce(const void* dy, long long lddy, const void* y, long long ldy, const double* stats,
                      const float* gamma, const float* beta, int G, int relu, double* sums, int N, long long V, int C,
                      float eps, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= GN_MAXC && C % G == 0, "gn_bwd_reduce: bad C=%d G=%d", C, G);
  // exactly one resident wave (3 CTAs per SM): 4 per SM ran 1.33 waves, the last one a third full
  const int per_sample = std::max(1, std::min(ew_blocks(V * (C / 8), GN_THREADS * 8), b3d_num_sms() * 3 / std::max(1, N)));
  dim3 grid(per_sample, N);
  const size_t smem = (4 * (size_t)GN_KS(C) + 4 * (size_t)C) * sizeof(float);   // 4 padded coefficient arrays + 2C doubles
  static const cudaError_t attr = [] {   // one-time, thread-safe; up to 64 KB at C = GN_MAXC (the wide model's 2048-channel concat buffers)
    cudaError_t e = cudaFuncSetAttribute(gn_bwd_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * GN_MAXC * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gn_bwd_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * GN_MAXC * (int)sizeof(float));
  }();
  B3D_CHECK_CUDA(attr);
  cudaStream_t st = (cudaStream_t)stream;
#define RARGS (const bf16*)dy, lddy, (const bf16*)y, ldy, stats, gamma, beta, G, sums, V, C, eps
  // B3D_GN_LEAN: bit 0 = lean apply, bit 1 = lean reduce (register-lean kernels, fixed-chunk shapes only)
  static const int lean = getenv("B3D_GN_LEAN") ? atoi(getenv("B3D_GN_LEAN")) : 3;
  if ((lean & 2) && (GN_THREADS % (C / 8)) == 0) {
    static const cudaError_t attr2 = [] {
      cudaError_t e = cudaFuncSetAttribute(gn_bwd_reduce_lean_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * GN_MAXC * (int)sizeof(float));
      if (e != cudaSuccess) return e;
      return cudaFuncSetAttribute(gn_bwd_reduce_lean_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * GN_MAXC * (int)sizeof(float));
    }();
    B3D_CHECK_CUDA(attr2);
    if (relu) gn_bwd_reduce_lean_kernel<true, 4><<<grid, GN_THREADS, smem, st>>>(RARGS);
    else gn_bwd_reduce_lean_kernel<false, 4><<<grid, GN_THREADS, smem, st>>>(RARGS);
  } else if (relu) gn_bwd_reduce_kernel<true><<<grid, GN_THREADS, smem, st>>>(RARGS);
  else gn_bwd_reduce_kernel<false><<<grid, GN_THREADS, smem, st>>>(RARGS);
  ++g_b3d_launches;
#undef RARGS
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gn_bwd_apply(const void* dy, long long lddy, const void* y, long long ldy, const double* stats,
                     const float* gamma, const float* beta, int G, int relu, const double* sums, void* dx,
                     long long lddx, int accumulate, int N, long long V, int C, float eps, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= GN_MAXC && C % G == 0 && G <= GN_MAXG, "gn_bwd_apply: bad C=%d G=%d", C, G);
  dim3 grid(ew_blocks(V * (C / 8), GN_THREADS * 2), N);
  const size_t smem = 6 * (size_t)GN_KS(C) * sizeof(float);   // six (padded, see kidx) coefficient arrays
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(RL, AC)                                                                                               \
  gn_bwd_apply_kernel<RL, AC><<<grid, GN_THREADS, smem, st>>>((const bf16*)dy, lddy, (const bf16*)y, ldy, stats, gamma, \
                                                              beta, G, sums, (bf16*)dx, lddx, V, C, eps)
#define LAUNCHL(RL, AC)                                                                                                  \
  gn_bwd_apply_lean_kernel<RL, AC, 2><<<grid, GN_THREADS, smem, st>>>((const bf16*)dy, lddy, (const bf16*)y, ldy, stats, gamma, \
                                                                      beta, G, sums, (bf16*)dx, lddx, V, C, eps)
  static const int lean = getenv("B3D_GN_LEAN") ? atoi(getenv("B3D_GN_LEAN")) : 3;
  static const cudaError_t attr = [] {   // 48 KB of coefficients at C = 2048 plus the static group-term scratch: opt in once
    const int mx = 6 * GN_KS(GN_MAXC) * (int)sizeof(float);
    cudaError_t e = cudaSuccess;
#define SETA(K) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)
    SETA((gn_bwd_apply_kernel<true, true>)); SETA((gn_bwd_apply_kernel<true, false>));
    SETA((gn_bwd_apply_kernel<false, true>)); SETA((gn_bwd_apply_kernel<false, false>));
    SETA((gn_bwd_apply_lean_kernel<true, true, 2>)); SETA((gn_bwd_apply_lean_kernel<true, false, 2>));
    SETA((gn_bwd_apply_lean_kernel<false, true, 2>)); SETA((gn_bwd_apply_lean_kernel<false, false, 2>));
#undef SETA
    return e;
  }();
  B3D_CHECK_CUDA(attr);
  if ((lean & 1) && (GN_THREADS % (C / 8)) == 0) {
    if (relu) { if (accumulate) LAUNCHL(true, true); else LAUNCHL(true, false); }
    else { if (accumulate) LAUNCHL(false, true); else LAUNCHL(false, false); }
  } else
  if (relu) { if (accumulate) LAUNCH(true, true); else LAUNCH(true, false); }
  else { if (accumulate) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
#undef LAUNCHL
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// Dual backward of  out = relu(GN_a(ya)) + GN_b(yb)  w.r.t. ya and yb (both branches see dy).  sums_a / sums_b: zeroed
// double [N][C][2], left holding (Σdz, Σdz·x̂) per channel for b3d_gn_param_grad.  Returns 1 (nothing launched) when the
// shape does not suit the fixed-chunk kernels — the caller then uses b3d_gn_bwd_reduce / b3d_gn_bwd_apply per branch.
int b3d_gn_bwd_dual(const void* dy, long long lddy, const void* ya, long long ldya, const double* stats_a,
                    const float* gamma_a, const float* beta_a, const void* yb, long long ldyb, const double* stats_b,
                    const float* gamma_b, int G, double* sums_a, double* sums_b, void* dxa, long long lddxa, void* dxb,
                    long long lddxb, int N, long long V, int C, float eps, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C % G == 0 && G <= GN_MAXG, "gn_bwd_dual: bad C=%d G=%d", C, G);
  const int C8 = C / 8;
  if (C > 512 || (GN_THREADS % C8) != 0) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int per_sample = std::max(1, std::min(ew_blocks(V * C8, GN_THREADS * 8), b3d_num_sms() * 2 / std::max(1, N)));   // one resident wave
  // OCC = resident blocks per SM the register allocation is capped for: 2 (121/126 registers, no spills; ncu: 25 % warps
  // active, 4.3-4.6 TB/s, profiles/ncu_r1_gn_bwd_dual.txt) or 3 (80 registers, ~150 bytes of spills).  Measured at
  // 2x128^3x32 (scripts/gn_dual_once.py): OCC 2 0.485 ms, OCC 3 0.635 ms — the spills cost more than the extra warps hide,
  // so 2 is the default; B3D_GN_DUAL_OCC=3 keeps the experiment reproducible.
  static const int occ = getenv("B3D_GN_DUAL_OCC") ? atoi(getenv("B3D_GN_DUAL_OCC")) : 2;
  const size_t smem_r = (6 * (size_t)GN_KS(C)) * sizeof(float) + 4 * (size_t)C * sizeof(double), smem_a = 8 * (size_t)GN_KS(C) * sizeof(float);
  const dim3 grid_r(per_sample, N), grid_a(ew_blocks(V * C8, GN_THREADS * 2), N);
#define DUAL_ARGS_R (const bf16*)dy, lddy, (const bf16*)ya, ldya, stats_a, gamma_a, beta_a, (const bf16*)yb, ldyb, stats_b, G, sums_a, sums_b, V, C, eps
#define DUAL_ARGS_A (const bf16*)dy, lddy, (const bf16*)ya, ldya, stats_a, gamma_a, beta_a, sums_a, (const bf16*)yb, ldyb, stats_b, gamma_b, sums_b, G, (bf16*)dxa, lddxa, (bf16*)dxb, lddxb, V, C, eps
  // B3D_GN_DUAL_LEAN: bit 0 = lean apply kernel, bit 1 = lean reduce kernel (see the kernels' header)
  static const int lean = getenv("B3D_GN_DUAL_LEAN") ? atoi(getenv("B3D_GN_DUAL_LEAN")) : 3;
  if (lean & 2) {
    const dim3 grid_l(std::max(1, std::min(ew_blocks(V * C8, GN_THREADS * 8), b3d_num_sms() * 3 / std::max(1, N))), N);
    gn_bwd_dual_reduce_lean_kernel<2><<<grid_l, GN_THREADS, smem_r, st>>>(DUAL_ARGS_R);
  } else if (occ == 3) gn_bwd_dual_reduce_kernel<3><<<grid_r, GN_THREADS, smem_r, st>>>(DUAL_ARGS_R);
  else gn_bwd_dual_reduce_kernel<2><<<grid_r, GN_THREADS, smem_r, st>>>(DUAL_ARGS_R);
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  if (lean & 1) gn_bwd_dual_apply_lean_kernel<2><<<grid_a, GN_THREADS, smem_a, st>>>(DUAL_ARGS_A);
  else if (occ == 3) gn_bwd_dual_apply_kernel<3><<<grid_a, GN_THREADS, smem_a, st>>>(DUAL_ARGS_A);
  else gn_bwd_dual_apply_kernel<2><<<grid_a, GN_THREADS, smem_a, st>>>(DUAL_ARGS_A);
  ++g_b3d_launches;
#undef DUAL_ARGS_R
#undef DUAL_ARGS_A
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gn_param_grad(const double* sums, int N, int C, float* dgamma, float* dbeta, int accumulate, void* stream) {
  gn_param_grad_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, N, C, dgamma, dbeta, accumulate); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_add_bf16(const void* a, long long lda, const void* b, long long ldb, void* o, long long ldo, long long V, int C,
                 void* stream) {
  B3D_REQUIRE(C % 8 == 0, "add: C must be a multiple of 8");
  add_kernel<<<ew_blocks(V * (C / 8), GN_THREADS * 2), GN_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)o, ldo, V, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
