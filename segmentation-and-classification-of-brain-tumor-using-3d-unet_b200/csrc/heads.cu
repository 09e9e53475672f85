// heads.cu — the 4-class output heads of the U-Net (all HBM-bound, CUDA cores):
//   * deep-supervision head: Conv1x1(f_i -> K)+bias on an encoder skip, then F.interpolate(trilinear,
//     align_corners=False) to full resolution                      /root/reference/main.py:137-140,164-171
//   * final head: BatchNorm3d(F2) (batch stats in train, running stats in eval, momentum 0.1) -> ReLU -> Conv1x1(F2 -> K)
//     applied to the output of final_conv.0                         /root/reference/main.py:129-134,198
// Logits are produced as fp32 NCDHW (what the reference returns and what the loss reads).  K (classes) is fixed to 4
// (BraTS labels 0..3, main.py:336 / train_model.py:174); other values are rejected loudly.
#include "loss_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define KCLS 4

static int hd_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)b3d_num_sms() * 16;
  return (int)std::max<long long>(1, std::min(b, cap));
}

// ------------------------------------------------------------------------------------------------
// deep-supervision 1x1 head: l[n][v][k] = b[k] + Σ_c skip[n][v][c] * W[k][c]      (fp32 float4 per voxel)
// lanes of a warp = (voxel sub-index, 8-channel chunk); partial dots are reduced with shuffles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) ds_head_fwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                          const float* __restrict__ b, float4* __restrict__ out,
                                                          long long NV, int C) {
  extern __shared__ float sw[];  // [K][C]
  for (int i = threadIdx.x; i < KCLS * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int C8 = C >> 3;
  const int lanes_c = C8 < 32 ? C8 : 32;       // lanes cooperating on one voxel
  const int vpw = 32 / lanes_c;                // voxels per warp iteration
  const int lane = threadIdx.x & 31;
  const int lc = lane % lanes_c, lv = lane / lanes_c;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (C8 <= 32) {
    // few channels (levels 0-2: the launches that move hundreds of MB): every lane owns ONE 16-byte chunk per voxel, so four
    // voxel groups are loaded back to back before the first FMA — 4 independent loads in flight per lane instead of 1
    constexpr int UF = 4;
    float wk[KCLS][8];
#pragma unroll
    for (int k = 0; k < KCLS; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) wk[k][j] = (lc < C8) ? sw[k * C + lc * 8 + j] : 0.f;
    for (long long v0 = warp_id * vpw * UF; v0 < NV; v0 += nwarps * vpw * UF) {
      uint4 raw[UF];
#pragma unroll
      for (int u = 0; u < UF; ++u) {
        const long long v = v0 + u * vpw + lv;
        raw[u] = (v < NV && lc < C8) ? ldg16_stream(x + v * ldx + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < UF; ++u) {
        const long long v = v0 + u * vpw + lv;
        float a[8], acc[KCLS] = {0.f, 0.f, 0.f, 0.f};
        unpack8(raw[u], a);
#pragma unroll
        for (int k = 0; k < KCLS; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(a[j], wk[k][j], acc[k]);
        for (int o = lanes_c >> 1; o > 0; o >>= 1)
#pragma unroll
          for (int k = 0; k < KCLS; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lc == 0 && v < NV) out[v] = make_float4(acc[0] + b[0], acc[1] + b[1], acc[2] + b[2], acc[3] + b[3]);
      }
    }
    return;
  }
  for (long long v0 = warp_id * vpw; v0 < NV; v0 += nwarps * vpw) {
    const long long v = v0 + lv;
    float acc[KCLS] = {0.f, 0.f, 0.f, 0.f};
    if (v < NV) {
      for (int c8 = lc; c8 < C8; c8 += lanes_c) {
        float a[8];
        unpack8(ldg16_stream(x + v * ldx + c8 * 8), a);
#pragma unroll
        for (int k = 0; k < KCLS; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(a[j], sw[k * C + c8 * 8 + j], acc[k]);
      }
    }
    for (int o = lanes_c >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < KCLS; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lc == 0 && v < NV) out[v] = make_float4(acc[0] + b[0], acc[1] + b[1], acc[2] + b[2], acc[3] + b[3]);
  }
}

// backward of the 1x1 head: dl is PLANAR fp32 [N][K][Vs] (CL = false) or channel-last float4 [N][Vs] (CL = true, what the
// fused deep-supervision loss produces); dskip[n][v][c] (+)= Σ_k dl_k W[k][c];
// dW[k][c] += Σ dl_k skip_c ; db[k] += Σ dl_k   (fp32 atomics into caller-zeroed buffers)
template <bool ACC, bool CL>
__global__ void __launch_bounds__(256, 2) ds_head_bwd_kernel(const float* __restrict__ dl, const bf16* __restrict__ x,
                                                          long long ldx, const float* __restrict__ w, bf16* __restrict__ dx,
                                                          long long lddx, float* __restrict__ dW, float* __restrict__ db,
                                                          int N, long long Vs, int C) {
  extern __shared__ float sw[];  // [K][C] weights, then [K][C] dW accum, then [K] db accum
  float* sdw = sw + KCLS * C;
  float* sdb = sdw + KCLS * C;
  for (int i = threadIdx.x; i < KCLS * C; i += blockDim.x) { sw[i] = w[i]; sdw[i] = 0.f; }
  if (threadIdx.x < KCLS) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int C8 = C >> 3;
  const int lanes_c = C8 < 32 ? C8 : 32;
  const int vpw = 32 / lanes_c;
  const int lane = threadIdx.x & 31;
  const int lc = lane % lanes_c, lv = lane / lanes_c;
  const long long NV = (long long)N * Vs;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nchunks = (C8 + lanes_c - 1) / lanes_c;  // chunks per lane (1 unless C > 256)
  float gw[KCLS][8];
#pragma unroll
  for (int k = 0; k < KCLS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) gw[k][j] = 0.f;
  float gb[KCLS] = {0.f, 0.f, 0.f, 0.f};
  if (nchunks == 1 && CL) {
    // every lane owns one 16-byte chunk per voxel: UB voxel groups per iteration, all their loads (logit gradient, skip,
    // running dskip) issued before the first FMA; the 4 x 8 weights are re-read from shared memory two channels at a time
    // (register-lean, see b3d_common.cuh) so that four groups fit in the register file
    constexpr int UB = 4;
    const float* wq = sw + lc * 8;
    for (long long v0 = warp_id * vpw * UB; v0 < NV; v0 += nwarps * vpw * UB) {
      unsigned xa[UB][4], xo[UB][4];
      float4 g4[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const long long v = v0 + u * vpw + lv;
        const bool ok = v < NV && lc < C8;
        g4[u] = ok ? __ldg(reinterpret_cast<const float4*>(dl) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint4 t = ok ? ldg16_stream(x + v * ldx + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        xa[u][0] = t.x; xa[u][1] = t.y; xa[u][2] = t.z; xa[u][3] = t.w;
        if (ACC) {
          const uint4 t2 = ok ? ldg16(dx + v * lddx + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
          xo[u][0] = t2.x; xo[u][1] = t2.y; xo[u][2] = t2.z; xo[u][3] = t2.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 wv[KCLS];
#pragma unroll
        for (int k = 0; k < KCLS; ++k) wv[k] = (lc < C8) ? lds2v(wq + k * C + 2 * q) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const float g[KCLS] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
          const float2 a = bfw(xa[u][q]);
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int k = 0; k < KCLS; ++k) { s0 = fmaf(g[k], wv[k].x, s0); s1 = fmaf(g[k], wv[k].y, s1); }
          if (ACC) { const float2 o = bfw(xo[u][q]); s0 += o.x; s1 += o.y; }
          xo[u][q] = wbf(s0, s1);
#pragma unroll
          for (int k = 0; k < KCLS; ++k) { gw[k][2 * q] = fmaf(g[k], a.x, gw[k][2 * q]); gw[k][2 * q + 1] = fmaf(g[k], a.y, gw[k][2 * q + 1]); }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const long long v = v0 + u * vpw + lv;
        if (!(v < NV && lc < C8)) continue;
        stg16(dx + v * lddx + lc * 8, make_uint4(xo[u][0], xo[u][1], xo[u][2], xo[u][3]));
        if (lc == 0) { gb[0] += g4[u].x; gb[1] += g4[u].y; gb[2] += g4[u].z; gb[3] += g4[u].w; }
      }
    }
  } else
  for (long long v0 = warp_id * vpw; v0 < NV; v0 += nwarps * vpw) {
    const long long v = v0 + lv;
    if (v >= NV) continue;
    const long long n = v / Vs, vs = v - n * Vs;
    float g[KCLS];
    if (CL) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(dl) + v);
      g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    } else {
#pragma unroll
      for (int k = 0; k < KCLS; ++k) g[k] = __ldg(dl + (n * KCLS + k) * Vs + vs);
    }
    if (lc == 0) {
#pragma unroll
      for (int k = 0; k < KCLS; ++k) gb[k] += g[k];
    }
    for (int c8 = lc; c8 < C8; c8 += lanes_c) {
      float a[8], o[8];
      unpack8(ldg16_stream(x + v * ldx + c8 * 8), a);
      if (ACC) unpack8(ldg16(dx + v * lddx + c8 * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < KCLS; ++k) s = fmaf(g[k], sw[k * C + c8 * 8 + j], s);
        o[j] = ACC ? o[j] + s : s;
      }
      stg16(dx + v * lddx + c8 * 8, pack8(o));
      if (nchunks == 1) {
#pragma unroll
        for (int k = 0; k < KCLS; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) gw[k][j] = fmaf(g[k], a[j], gw[k][j]);
      } else {
#pragma unroll
        for (int k = 0; k < KCLS; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(&sdw[k * C + c8 * 8 + j], g[k] * a[j]);
      }
    }
  }
  const bool p2 = (lanes_c & (lanes_c - 1)) == 0;
  if (nchunks == 1) {   // lanes lc, lc + lanes_c, ... of a warp own the same chunk: fold them with shuffles, then one shared atomic per warp
#pragma unroll
    for (int k = 0; k < KCLS; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) gw[k][j] = p2 ? warp_sum_mod(gw[k][j], lanes_c) : gw[k][j];
    if ((!p2 || lane < lanes_c) && lc < C8) {
#pragma unroll
      for (int k = 0; k < KCLS; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sdw[k * C + lc * 8 + j], gw[k][j]);
    }
  }
  if (lc == 0) {
#pragma unroll
    for (int k = 0; k < KCLS; ++k) atomicAdd(&sdb[k], gb[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KCLS * C; i += blockDim.x) atomicAdd(&dW[i], sdw[i]);
  if (threadIdx.x < KCLS) atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// trilinear upsample (align_corners=False) of low-res logits float4[N][Dl][Hl][Wl] to planar fp32 [N][K][D][H][W]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trilinear_up_fwd_kernel(const float4* __restrict__ lo, float* __restrict__ out, int N,
                                                               int Dl, int Hl, int Wl, int D, int H, int W) {
  const long long V = (long long)D * H * W;
  const long long total = (long long)N * V;
  const float sd = (float)Dl / D, shh = (float)Hl / H, sww = (float)Wl / W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H); t /= H;
    const int z = (int)(t % D); const int n = (int)(t / D);
    int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
    lerp_src(z, sd, Dl, z0, z1, lz); lerp_src(y, shh, Hl, y0, y1, ly); lerp_src(x, sww, Wl, x0, x1, lx);
    const float4* b = lo + (long long)n * Dl * Hl * Wl;
#define AT(zz, yy, xx) __ldg(b + ((long long)(zz) * Hl + (yy)) * Wl + (xx))
    const float4 v000 = AT(z0, y0, x0), v001 = AT(z0, y0, x1), v010 = AT(z0, y1, x0), v011 = AT(z0, y1, x1);
    const float4 v100 = AT(z1, y0, x0), v101 = AT(z1, y0, x1), v110 = AT(z1, y1, x0), v111 = AT(z1, y1, x1);
#undef AT
    const float wz0 = 1.f - lz, wy0 = 1.f - ly, wx0 = 1.f - lx;
#define MIX(f)                                                                                         \
  (wz0 * (wy0 * (wx0 * v000.f + lx * v001.f) + ly * (wx0 * v010.f + lx * v011.f)) +                    \
   lz * (wy0 * (wx0 * v100.f + lx * v101.f) + ly * (wx0 * v110.f + lx * v111.f)))
    const long long vv = ((long long)z * H + y) * W + x;
    float* o = out + (long long)n * KCLS * V + vv;
    o[0] = MIX(x); o[V] = MIX(y); o[2 * V] = MIX(z); o[3 * V] = MIX(w);
#undef MIX
  }
}

// adjoint of 1-D linear upsampling along one axis of a planar fp32 tensor [outer][Lout][inner] -> [outer][Lin][inner]
__global__ void __launch_bounds__(256) lerp_adjoint_kernel(const float* __restrict__ in, float* __restrict__ out, long long outer,
                                                           int Lout, int Lin, long long inner) {
  const long long total = outer * Lin * inner;
  const float scale = (float)Lin / Lout;
  const int s = Lout / Lin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const long long q = t % inner; t /= inner;
    const int li = (int)(t % Lin); const long long ou = t / Lin;
    const float* src = in + ou * Lout * inner + q;
    float acc = 0.f;
    const int o_lo = max(0, s * (li - 1)), o_hi = min(Lout, s * (li + 2));
    for (int o = o_lo; o < o_hi; ++o) {
      int i0, i1; float l1;
      lerp_src(o, scale, Lin, i0, i1, l1);
      float wgt = 0.f;
      if (i0 == li) wgt += 1.f - l1;
      if (i1 == li) wgt += l1;
      if (wgt != 0.f) acc = fmaf(wgt, __ldg(src + (long long)o * inner), acc);
    }
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// final head:  logits[n][k][v] = b2[k] + Σ_c W2[k][c] * relu(BN(h[n][v][c]))
// bn: [0..F2) = mean, [F2..2F2) = rstd  (computed by final_bn_prepare from batch or running statistics)
// ------------------------------------------------------------------------------------------------
__global__ void final_bn_prepare_kernel(const double* __restrict__ stats, double count, int train, float* running_mean,
                                        float* running_var, long long* num_batches, float momentum, float eps,
                                        float* __restrict__ bn, int F2, int update_running) {
  const int c = threadIdx.x;
  if (c >= F2) return;
  float mean, var;
  if (train) {
    const double mu = stats[2 * c] / count;
    double v = stats[2 * c + 1] / count - mu * mu;
    if (v < 0) v = 0;
    mean = (float)mu; var = (float)v;
    if (update_running) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(v * count / (count - 1.0));
      if (c == 0 && num_batches) *num_batches += 1;
    }
  } else {
    mean = running_mean[c]; var = running_var[c];
  }
  bn[c] = mean;
  bn[F2 + c] = rsqrtf(var + eps);
}

template <int F2>
__global__ void __launch_bounds__(256) final_head_fwd_kernel(const bf16* __restrict__ h, long long ldh, const float* __restrict__ bn,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             float* __restrict__ out, int N, long long V) {
  __shared__ float s_sc[F2], s_sh[F2], s_w[KCLS * F2], s_b[KCLS];
  if (threadIdx.x < F2) {
    const float sc = gamma[threadIdx.x] * bn[F2 + threadIdx.x];
    s_sc[threadIdx.x] = sc; s_sh[threadIdx.x] = beta[threadIdx.x] - bn[threadIdx.x] * sc;
  }
  for (int i = threadIdx.x; i < KCLS * F2; i += blockDim.x) s_w[i] = w2[i];
  if (threadIdx.x < KCLS) s_b[threadIdx.x] = b2[threadIdx.x];
  __syncthreads();
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    float acc[KCLS];
#pragma unroll
    for (int k = 0; k < KCLS; ++k) acc[k] = s_b[k];
#pragma unroll
    for (int c8 = 0; c8 < F2 / 8; ++c8) {
      float a[8];
      unpack8(ldg16_stream(h + i * ldh + c8 * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float r = fmaxf(fmaf(a[j], s_sc[c8 * 8 + j], s_sh[c8 * 8 + j]), 0.f);
#pragma unroll
        for (int k = 0; k < KCLS; ++k) acc[k] = fmaf(r, s_w[k * F2 + c8 * 8 + j], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < KCLS; ++k) out[(n * KCLS + k) * V + v] = acc[k];
  }
}

// backward phase 1: red[0..F2) Σdz, [F2..2F2) Σdz*xhat, then dW2 [K][F2], then db2 [K]   (double atomics, caller zeroes)
// F2/8 lanes share a voxel, each owning one 16-byte chunk of h: 52 accumulators per thread instead of 100 (the thread-per-voxel
// version needed 160 registers -> one CTA per SM, one voxel in flight per thread: 1.3 TB/s), four voxel groups in flight.
template <int F2>
__global__ void __launch_bounds__(256, 2) final_head_bwd_reduce_kernel(const float* __restrict__ dl, const bf16* __restrict__ h,
                                                                    long long ldh, const float* __restrict__ bn,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    const float* __restrict__ w2, double* __restrict__ red, int N,
                                                                    long long V) {
  constexpr int NR = 2 * F2 + KCLS * F2 + KCLS;
  constexpr int LPV = F2 / 8;          // lanes per voxel
  constexpr int VPW = 32 / LPV;        // voxels per warp pass
  constexpr int UF = 2;
  __shared__ double s_red[NR];   // fp64: warp arrival order cannot change the BatchNorm-backward sums
  for (int i = threadIdx.x; i < NR; i += blockDim.x) s_red[i] = 0.0;
  const int lane = threadIdx.x & 31;
  const int lc = lane % LPV, lv = lane / LPV;
  __shared__ __align__(16) float s_w[LPV * KCLS * 8];   // [chunk][class][8]: read back as float4 broadcasts
  for (int i = threadIdx.x; i < LPV * KCLS * 8; i += blockDim.x) {
    const int c8 = i / (KCLS * 8), k = (i / 8) % KCLS, j = i & 7;
    s_w[i] = w2[k * F2 + c8 * 8 + j];
  }
  float k_mean[8], k_rstd[8], k_g[8], k_bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lc * 8 + j;
    k_mean[j] = bn[c]; k_rstd[j] = bn[F2 + c]; k_g[j] = gamma[c]; k_bt[j] = beta[c];
  }
  __syncthreads();
  const float4* wl = reinterpret_cast<const float4*>(s_w + lc * KCLS * 8);
  float a_dz[8], a_dzx[8], a_w[KCLS][8], a_b[KCLS];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a_dz[j] = 0.f; a_dzx[j] = 0.f; }
#pragma unroll
  for (int k = 0; k < KCLS; ++k) {
    a_b[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a_w[k][j] = 0.f;
  }
  const long long total = (long long)N * V;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i0 = warp_id * VPW * UF; i0 < total; i0 += nwarps * VPW * UF) {
    uint4 raw[UF];
    float g[UF][KCLS];
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      const long long i = i0 + u * VPW + lv;
      const bool ok = i < total;
      const long long n = ok ? i / V : 0, v = ok ? i - n * V : 0;
      raw[u] = ok ? ldg16_stream(h + i * ldh + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < KCLS; ++k) g[u][k] = ok ? __ldg(dl + (n * KCLS + k) * V + v) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      const long long i = i0 + u * VPW + lv;
      if (i >= total) continue;
      float a[8];
      unpack8(raw[u], a);
      if (lc == 0) {
#pragma unroll
        for (int k = 0; k < KCLS; ++k) a_b[k] += g[u][k];
      }
      float dr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < KCLS; ++k) {
        const float4 w0 = wl[2 * k], w1 = wl[2 * k + 1];
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) dr[j] = fmaf(g[u][k], wv[j], dr[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (a[j] - k_mean[j]) * k_rstd[j];
        const float hn = fmaf(xh, k_g[j], k_bt[j]);
        const float r = fmaxf(hn, 0.f);
#pragma unroll
        for (int k = 0; k < KCLS; ++k) a_w[k][j] = fmaf(g[u][k], r, a_w[k][j]);
        const float dz = hn > 0.f ? dr[j] : 0.f;
        a_dz[j] += dz; a_dzx[j] += dz * xh;
      }
    }
  }
  // fold the lanes that own the same chunk (stride LPV), then one shared fp64 atomic per warp and value
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s1 = warp_sum_mod(a_dz[j], LPV), s2 = warp_sum_mod(a_dzx[j], LPV);
    if (lane < LPV) { atomicAdd(&s_red[lc * 8 + j], (double)s1); atomicAdd(&s_red[F2 + lc * 8 + j], (double)s2); }
  }
#pragma unroll
  for (int k = 0; k < KCLS; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sv = warp_sum_mod(a_w[k][j], LPV);
      if (lane < LPV) atomicAdd(&s_red[2 * F2 + k * F2 + lc * 8 + j], (double)sv);
    }
    const float sb = warp_sum(a_b[k]);
    if (lane == 0) atomicAdd(&s_red[2 * F2 + KCLS * F2 + k], (double)sb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NR; i += blockDim.x) atomicAdd(&red[i], s_red[i]);
}

// backward phase 2: dh = BN-backward(dz) as bf16 [N][V][F2]; train uses batch statistics, eval the running ones
template <int F2>
__global__ void __launch_bounds__(256) final_head_bwd_apply_kernel(const float* __restrict__ dl, const bf16* __restrict__ h,
                                                                   long long ldh, const float* __restrict__ bn,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   const float* __restrict__ w2, const double* __restrict__ red,
                                                                   int train, bf16* __restrict__ dh, long long lddh, int N,
                                                                   long long V) {
  __shared__ float s_mean[F2], s_rstd[F2], s_g[F2], s_bt[F2], s_w[KCLS * F2], s_m1[F2], s_m2[F2];
  const double cnt = (double)N * (double)V;
  if (threadIdx.x < F2) {
    const int c = threadIdx.x;
    s_mean[c] = bn[c]; s_rstd[c] = bn[F2 + c]; s_g[c] = gamma[c]; s_bt[c] = beta[c];
    s_m1[c] = train ? (float)(red[c] / cnt) : 0.f;
    s_m2[c] = train ? (float)(red[F2 + c] / cnt) : 0.f;
  }
  for (int i = threadIdx.x; i < KCLS * F2; i += blockDim.x) s_w[i] = w2[i];
  __syncthreads();
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    float g[KCLS];
#pragma unroll
    for (int k = 0; k < KCLS; ++k) g[k] = __ldg(dl + (n * KCLS + k) * V + v);
#pragma unroll
    for (int c8 = 0; c8 < F2 / 8; ++c8) {
      float a[8], o[8];
      unpack8(ldg16_stream(h + i * ldh + c8 * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c8 * 8 + j;
        const float xh = (a[j] - s_mean[c]) * s_rstd[c];
        const float hn = fmaf(xh, s_g[c], s_bt[c]);
        float dr = 0.f;
#pragma unroll
        for (int k = 0; k < KCLS; ++k) dr = fmaf(g[k], s_w[k * F2 + c], dr);
        const float dz = hn > 0.f ? dr : 0.f;
        o[j] = s_g[c] * s_rstd[c] * (dz - s_m1[c] - xh * s_m2[c]);
      }
      stg16(dh + i * lddh + c8 * 8, pack8(o));
    }
  }
}

// unpack the reduction buffer into parameter gradients (fp32)
__global__ void final_head_param_grad_kernel(const double* __restrict__ red, int F2, float* dgamma, float* dbeta, float* dW2,
                                             float* db2) {
  const int i = threadIdx.x;
  if (i < F2) { dbeta[i] = (float)red[i]; dgamma[i] = (float)red[F2 + i]; }
  for (int j = i; j < KCLS * F2; j += blockDim.x) dW2[j] = (float)red[2 * F2 + j];
  if (i < KCLS) db2[i] = (float)red[2 * F2 + KCLS * F2 + i];
}

extern "C" {

int b3d_ds_head_fwd(const void* x, long long ldx, const float* w, const float* b, float* out, long long NV, int C, int K,
                    void* stream) {
  B3D_REQUIRE(K == KCLS, "ds_head: only %d output classes supported (got %d)", KCLS, K);
  B3D_REQUIRE(C % 8 == 0 && C <= 2048, "ds_head: bad C");
  ds_head_fwd_kernel<<<hd_blocks(NV * 8, 256), 256, KCLS * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, w, b, (float4*)out, NV, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

static int ds_head_bwd_launch(const float* dl, bool channel_last, const void* x, long long ldx, const float* w, void* dx,
                              long long lddx, int accumulate, float* dW, float* db, int N, long long Vs, int C, int K,
                              void* stream) {
  B3D_REQUIRE(K == KCLS, "ds_head: only %d output classes supported (got %d)", KCLS, K);
  B3D_REQUIRE(C % 8 == 0 && C <= 2048, "ds_head: bad C");
  const size_t smem = (2 * KCLS * C + KCLS) * sizeof(float);
  const int blocks = std::min(hd_blocks((long long)N * Vs * 8, 256), b3d_num_sms() * 2);   // one resident wave (128 registers: 2 CTAs per SM)
  cudaStream_t st = (cudaStream_t)stream;
#define DSB(A, L) ds_head_bwd_kernel<A, L><<<blocks, 256, smem, st>>>(dl, (const bf16*)x, ldx, w, (bf16*)dx, lddx, dW, db, N, Vs, C)
  if (accumulate) { if (channel_last) DSB(true, true); else DSB(true, false); }
  else { if (channel_last) DSB(false, true); else DSB(false, false); }
#undef DSB
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_ds_head_bwd(const float* dl, const void* x, long long ldx, const float* w, void* dx, long long lddx,
                    int accumulate, float* dW, float* db, int N, long long Vs, int C, int K, void* stream) {
  return ds_head_bwd_launch(dl, false, x, ldx, w, dx, lddx, accumulate, dW, db, N, Vs, C, K, stream);
}

// same, with the logit gradient channel-last: dl float [N][Vs][4] (16-byte aligned)
int b3d_ds_head_bwd_cl(const float* dl, const void* x, long long ldx, const float* w, void* dx, long long lddx,
                       int accumulate, float* dW, float* db, int N, long long Vs, int C, int K, void* stream) {
  return ds_head_bwd_launch(dl, true, x, ldx, w, dx, lddx, accumulate, dW, db, N, Vs, C, K, stream);
}

int b3d_trilinear_up_fwd(const float* lo, float* out, int N, int Dl, int Hl, int Wl, int D, int H, int W, int K,
                         void* stream) {
  B3D_REQUIRE(K == KCLS, "trilinear_up: only %d classes supported", KCLS);
  trilinear_up_fwd_kernel<<<hd_blocks((long long)N * D * H * W, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)lo, out, N, Dl, Hl, Wl, D, H, W); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// dup planar [N][K][D][H][W] -> dlo planar [N][K][Dl][Hl][Wl]; tmp must hold N*K*D*H*Wl + N*K*D*Hl*Wl floats
int b3d_trilinear_up_bwd(const float* dup, float* dlo, float* tmp, int N, int Dl, int Hl, int Wl, int D, int H, int W,
                         int K, void* stream) {
  B3D_REQUIRE(D % Dl == 0 && H % Hl == 0 && W % Wl == 0, "trilinear_up_bwd: integer scale factors only");
  cudaStream_t st = (cudaStream_t)stream;
  const long long NK = (long long)N * K;
  float* t1 = tmp;
  float* t2 = tmp + NK * D * H * Wl;
  long long tot = NK * D * H * Wl;
  lerp_adjoint_kernel<<<hd_blocks(tot, 256), 256, 0, st>>>(dup, t1, NK * D * H, W, Wl, 1); ++g_b3d_launches;
  tot = NK * D * Hl * Wl;
  lerp_adjoint_kernel<<<hd_blocks(tot, 256), 256, 0, st>>>(t1, t2, NK * D, H, Hl, Wl); ++g_b3d_launches;
  tot = NK * Dl * Hl * Wl;
  lerp_adjoint_kernel<<<hd_blocks(tot, 256), 256, 0, st>>>(t2, dlo, NK, D, Dl, (long long)Hl * Wl); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_final_bn_prepare(const double* stats, double count, int train, float* running_mean, float* running_var,
                         long long* num_batches, float momentum, float eps, float* bn, int F2, int update_running,
                         void* stream) {
  B3D_REQUIRE(F2 <= 64, "final_bn_prepare: F2 too large");
  final_bn_prepare_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(stats, count, train, running_mean, running_var, num_batches,
                                                             momentum, eps, bn, F2, update_running); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

#define F2_DISPATCH(F2v, CALL8, CALL16, CALL32)                                        \
  if (F2v == 8) { CALL8; } else if (F2v == 16) { CALL16; } else if (F2v == 32) { CALL32; } \
  else { b3d_set_error("final head: F2=%d unsupported (8,16,32)", F2v); return B3D_ERR_UNSUPPORTED; }

int b3d_final_head_fwd(const void* h, long long ldh, const float* bn, const float* gamma, const float* beta,
                       const float* w2, const float* b2, float* out, int N, long long V, int F2, int K, void* stream) {
  B3D_REQUIRE(K == KCLS, "final_head: only %d output classes supported (got %d)", KCLS, K);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = hd_blocks((long long)N * V, 256);
  F2_DISPATCH(F2,
    (final_head_fwd_kernel<8><<<blocks, 256, 0, st>>>((const bf16*)h, ldh, bn, gamma, beta, w2, b2, out, N, V)),
    (final_head_fwd_kernel<16><<<blocks, 256, 0, st>>>((const bf16*)h, ldh, bn, gamma, beta, w2, b2, out, N, V)),
    (final_head_fwd_kernel<32><<<blocks, 256, 0, st>>>((const bf16*)h, ldh, bn, gamma, beta, w2, b2, out, N, V))); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// red: double [2*F2 + K*F2 + K], caller zeroes
int b3d_final_head_bwd(const float* dl, const void* h, long long ldh, const float* bn, const float* gamma,
                       const float* beta, const float* w2, double* red, int train, void* dh, long long lddh,
                       float* dgamma, float* dbeta, float* dW2, float* db2, int N, long long V, int F2, int K,
                       void* stream) {
  B3D_REQUIRE(K == KCLS, "final_head: only %d output classes supported (got %d)", KCLS, K);
  cudaStream_t st = (cudaStream_t)stream;
  const int rblocks = std::min(hd_blocks((long long)N * V, 256), b3d_num_sms() * 2);
  const int blocks = hd_blocks((long long)N * V, 256);
  F2_DISPATCH(F2,
    (final_head_bwd_reduce_kernel<8><<<rblocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, N, V)),
    (final_head_bwd_reduce_kernel<16><<<rblocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, N, V)),
    (final_head_bwd_reduce_kernel<32><<<rblocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, N, V))); ++g_b3d_launches;
  F2_DISPATCH(F2,
    (final_head_bwd_apply_kernel<8><<<blocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, train, (bf16*)dh, lddh, N, V)),
    (final_head_bwd_apply_kernel<16><<<blocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, train, (bf16*)dh, lddh, N, V)),
    (final_head_bwd_apply_kernel<32><<<blocks, 256, 0, st>>>(dl, (const bf16*)h, ldh, bn, gamma, beta, w2, red, train, (bf16*)dh, lddh, N, V))); ++g_b3d_launches;
  final_head_param_grad_kernel<<<1, 128, 0, st>>>(red, F2, dgamma, dbeta, dW2, db2); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
