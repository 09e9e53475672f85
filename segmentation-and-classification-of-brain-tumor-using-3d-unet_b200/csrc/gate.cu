// gate.cu — AttentionGate3D (spatial attention + squeeze-excite channel attention), NDHWC bf16, HBM-bound kernels.
//
// Reference: /root/reference/main.py:244-299
//     g1 = GN4(conv1(g)+bg) ; x1 = GN4(conv1(x)+bx)                       (the two 1x1 convs run on tcgen05: conv_igemm.cu,
//     q  = relu(g1 + x1)                                                    their epilogues give the GN4 statistics)
//     ψr = conv1(q; wψ)+bψ   [1 channel] ; ψ = sigmoid(GN(1,1)(ψr))        -> gate_psi_fwd (+ Σψr, Σψr² per sample)
//     ca = sigmoid(W2·relu(W1·mean_v(x)+b1)+b2)                            -> channel_sum (pool_layout.cu) + gate_se_fwd
//     out = x * ψ * ca                                                      -> gate_apply_fwd (writes the concat slice)
// Backward follows SURVEY App. A2 with the same three global reductions.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

static int gt_blocks_per_sample(long long work, int threads, int N, int per_sm = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = std::max(1, b3d_num_sms() * per_sm / std::max(1, N));
  return (int)std::max<long long>(1, std::min(b, cap));
}

__device__ __forceinline__ void mean_rstd_from(const double* st, double m, float eps, float& mean, float& rstd) {
  const double mu = st[0] / m;
  double var = st[1] / m - mu * mu;
  if (var < 0) var = 0;
  mean = (float)mu; rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// ------------------------------------------------------------------------------------------------
// ψr[n][v] = bψ + Σ_c wψ[c] * relu(GN4(g1r)[c] + GN4(x1r)[c]) ;  stats_psi[n] += (Σψr, Σψr²)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_psi_fwd_kernel(
    const bf16* __restrict__ g1r, const bf16* __restrict__ x1r, const double* __restrict__ st_g,
    const double* __restrict__ st_x, const float* __restrict__ gam_g, const float* __restrict__ bet_g,
    const float* __restrict__ gam_x, const float* __restrict__ bet_x, const float* __restrict__ wpsi, float bpsi_unused,
    const float* __restrict__ bpsi, float* __restrict__ psi_raw, double* __restrict__ st_psi, long long V, int F, float eps) {
  extern __shared__ float sm[];
  float* scg = sm; float* scx = sm + F; float* sh = sm + 2 * F; float* wp = sm + 3 * F;
  __shared__ double s_red[2];   // fp64: warp arrival order cannot change the forward statistics
  const int n = blockIdx.y;
  const int cpg = F / 4;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float mg, rg, mx, rx;
    mean_rstd_from(st_g + ((long long)n * 4 + c / cpg) * 2, (double)cpg * V, eps, mg, rg);
    mean_rstd_from(st_x + ((long long)n * 4 + c / cpg) * 2, (double)cpg * V, eps, mx, rx);
    const float a = gam_g[c] * rg, b = gam_x[c] * rx;
    scg[c] = a; scx[c] = b; sh[c] = bet_g[c] - mg * a + bet_x[c] - mx * b; wp[c] = wpsi[c];
  }
  if (threadIdx.x < 2) s_red[threadIdx.x] = 0.0;
  __syncthreads();
  const float bp = bpsi[0];
  const int F8 = F >> 3;
  const int lanes_c = F8 < 32 ? F8 : 32;
  const int vpw = 32 / lanes_c;
  const int lane = threadIdx.x & 31;
  const int lc = lane % lanes_c, lv = lane / lanes_c;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bf16* gn_ = g1r + (long long)n * V * F;
  const bf16* xn_ = x1r + (long long)n * V * F;
  float s1 = 0.f, s2 = 0.f;
  if (F8 <= 32) {
    // every lane owns one fixed 8-channel chunk: four voxel groups in flight, the chunk's 32 constants re-read two channels at
    // a time from a padded shared tile (36-float chunk stride: lanes of different chunks hit different banks)
    __shared__ __align__(16) float cst[32 * 36];   // [chunk][kind: scg scx sh wp][8] (+4 pad)
    for (int i = threadIdx.x; i < F8 * 32; i += blockDim.x) {
      const int c8 = i >> 5, kind = (i >> 3) & 3, j = i & 7, c = c8 * 8 + j;
      cst[c8 * 36 + (i & 31)] = kind == 0 ? scg[c] : kind == 1 ? scx[c] : kind == 2 ? sh[c] : wp[c];
    }
    __syncthreads();
    const float* cb = cst + (lc < F8 ? lc : 0) * 36;
    constexpr int UP = 4;
    for (long long v0 = warp_id * vpw * UP; v0 < V; v0 += nwarps * vpw * UP) {
      unsigned wa[UP][4], wb[UP][4];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const long long v = v0 + u * vpw + lv;
        const bool ok = v < V && lc < F8;
        const uint4 ra = ok ? ldg16_stream(gn_ + v * F + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        const uint4 rb = ok ? ldg16_stream(xn_ + v * F + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        wa[u][0] = ra.x; wa[u][1] = ra.y; wa[u][2] = ra.z; wa[u][3] = ra.w;
        wb[u][0] = rb.x; wb[u][1] = rb.y; wb[u][2] = rb.z; wb[u][3] = rb.w;
      }
      float acc[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) acc[u] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 kg = lds2v(cb + 0 * 8 + 2 * q), kx = lds2v(cb + 1 * 8 + 2 * q), ks = lds2v(cb + 2 * 8 + 2 * q),
                     kw = lds2v(cb + 3 * 8 + 2 * q);
#pragma unroll
        for (int u = 0; u < UP; ++u) {
          const float2 a = bfw(wa[u][q]), b = bfw(wb[u][q]);
          const float q0 = fmaxf(fmaf(a.x, kg.x, fmaf(b.x, kx.x, ks.x)), 0.f);
          const float q1 = fmaxf(fmaf(a.y, kg.y, fmaf(b.y, kx.y, ks.y)), 0.f);
          acc[u] = fmaf(q0, kw.x, acc[u]);      // same channel order as the generic path below
          acc[u] = fmaf(q1, kw.y, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const long long v = v0 + u * vpw + lv;
        float t = (v < V && lc < F8) ? acc[u] : 0.f;
        for (int o = lanes_c >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lc == 0 && v < V) {
          const float p = t + bp;
          psi_raw[(long long)n * V + v] = p;
          s1 += p; s2 += p * p;
        }
      }
    }
  } else
  for (long long v0 = warp_id * vpw; v0 < V; v0 += nwarps * vpw) {
    const long long v = v0 + lv;
    float acc = 0.f;
    if (v < V) {
      for (int c8 = lc; c8 < F8; c8 += lanes_c) {
        float a[8], b[8];
        unpack8(ldg16_stream(gn_ + v * F + c8 * 8), a);
        unpack8(ldg16_stream(xn_ + v * F + c8 * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c8 * 8 + j;
          const float q = fmaxf(fmaf(a[j], scg[c], fmaf(b[j], scx[c], sh[c])), 0.f);
          acc = fmaf(q, wp[c], acc);
        }
      }
    }
    for (int o = lanes_c >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lc == 0 && v < V) {
      const float p = acc + bp;
      psi_raw[(long long)n * V + v] = p;
      s1 += p; s2 += p * p;
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { atomicAdd(&s_red[0], (double)s1); atomicAdd(&s_red[1], (double)s2); }
  __syncthreads();
  if (threadIdx.x < 2) atomicAdd(&st_psi[(long long)n * 2 + threadIdx.x], s_red[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite: ca[n][c] = sigmoid(b2[c] + Σ_j W2[c][j] * relu(b1[j] + Σ_k W1[j][k] * mean[n][k]))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) gate_se_fwd_kernel(const double* __restrict__ xsum, long long V, const float* __restrict__ w1,
                                   const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                   float* __restrict__ ca, float* __restrict__ zbuf, float* __restrict__ meanbuf, int C) {
  // one CTA per sample; every mat-vec row is owned by a WARP (lanes stride the contiguous row, shuffle reduction), so the
  // weight reads are coalesced and all 8 warps work (the first version let C/8 threads walk rows with stride-C loads)
  extern __shared__ float sm[];
  float* mean = sm; float* z = sm + C;
  const int n = blockIdx.x, R = C / 8;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    mean[c] = (float)(xsum[(long long)n * C + c] / (double)V);
    meanbuf[(long long)n * C + c] = mean[c];
  }
  __syncthreads();
  // four rows per warp pass, all loads of the pass in flight before the first use: the kernel is pure DRAM latency
  // (one CTA per sample streams 2 x C*C/8 floats), so memory-level parallelism is the only thing that matters
  for (int j0 = wrp * 4; j0 < R; j0 += nw * 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < C; k += 32) {
      float wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = (j0 + i < R) ? __ldg(w1 + (long long)(j0 + i) * C + k) : 0.f;
      const float mk = mean[k];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = fmaf(wv[i], mk, a[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float s = warp_sum(a[i]);
      if (lane == 0 && j0 + i < R) { const float v = fmaxf(s + b1[j0 + i], 0.f); z[j0 + i] = v; zbuf[(long long)n * R + j0 + i] = v; }
    }
  }
  __syncthreads();
  // second mat-vec: one THREAD per output channel, its row of W2 (R contiguous floats) read with independent loads
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* wr = w2 + (long long)c * R;
    float a = 0.f;
#pragma unroll 8
    for (int j = 0; j < R; ++j) a = fmaf(__ldg(wr + j), z[j], a);
    ca[(long long)n * C + c] = 1.f / (1.f + __expf(-(a + b2[c])));
  }
}

// out = x * sigmoid(GN1(ψr)) * ca
__global__ void __launch_bounds__(256) gate_apply_fwd_kernel(const bf16* __restrict__ x, long long ldx,
                                                             const float* __restrict__ psi_raw, const double* __restrict__ st_psi,
                                                             const float* __restrict__ gpsi, const float* __restrict__ bpsi_n,
                                                             const float* __restrict__ ca, bf16* __restrict__ out, long long ldo,
                                                             long long V, int C, float eps) {
  extern __shared__ float sca[];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sca[c] = ca[(long long)n * C + c];
  float mu, rstd;
  mean_rstd_from(st_psi + (long long)n * 2, (double)V, eps, mu, rstd);
  const float a = gpsi[0] * rstd, b = bpsi_n[0] - mu * a;
  __syncthreads();
  const int C8 = C >> 3;
  const long long total = V * C8;
  const bf16* xn = x + (long long)n * V * ldx;
  bf16* on = out + (long long)n * V * ldo;
  const float* pn = psi_raw + (long long)n * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    const float s = 1.f / (1.f + __expf(-fmaf(__ldg(pn + vox), a, b)));
    float v[8];
    unpack8(ldg16_stream(xn + vox * ldx + c0), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s * sca[c0 + j];
    stg16(on + vox * ldo + c0, pack8(v));
  }
}

// ------------------------------------------------------------------------------------------------
// backward 1: dx_direct = dout*s*ca ; dψn[v] = (Σ_c dout*x*ca) * s(1−s) ; dca[n][c] += Σ_v dout*x*s ;
//             st_dpsi[n] += (Σ dψn, Σ dψn*x̂ψ)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_apply_bwd_kernel(
    const bf16* __restrict__ dout, long long lddo, const bf16* __restrict__ x, long long ldx,
    const float* __restrict__ psi_raw, const double* __restrict__ st_psi, const float* __restrict__ gpsi,
    const float* __restrict__ bpsi_n, const float* __restrict__ ca, bf16* __restrict__ dx, long long lddx,
    float* __restrict__ dpsin, double* __restrict__ dca, double* __restrict__ st_dpsi, long long V, int C, float eps) {
  extern __shared__ float sm[];
  float* sca = sm; double* sdca = reinterpret_cast<double*>(sm + C);   // fp64 block accumulators: order-independent
  __shared__ double s_red[2];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { sca[c] = ca[(long long)n * C + c]; sdca[c] = 0.0; }
  if (threadIdx.x < 2) s_red[threadIdx.x] = 0.0;
  float mu, rstd;
  mean_rstd_from(st_psi + (long long)n * 2, (double)V, eps, mu, rstd);
  const float a = gpsi[0] * rstd, b = bpsi_n[0] - mu * a;
  __syncthreads();
  const int C8 = C >> 3;
  const int lanes_c = C8 < 32 ? C8 : 32;
  const int vpw = 32 / lanes_c;
  const int lane = threadIdx.x & 31;
  const int lc = lane % lanes_c, lv = lane / lanes_c;
  const int nchunks = (C8 + lanes_c - 1) / lanes_c;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bf16* don = dout + (long long)n * V * lddo;
  const bf16* xn = x + (long long)n * V * ldx;
  bf16* dxn = dx + (long long)n * V * lddx;
  float acc_ca[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_ca[j] = 0.f;
  float r1 = 0.f, r2 = 0.f;
  if (nchunks == 1) {
    // every lane owns one 16-byte chunk per voxel: UB voxel groups per iteration with all their loads issued first (one
    // group at a time left 32 KB in flight per SM and ran at 48 % of the copy bandwidth, scripts/ew_bench.py round 2)
    constexpr int UB = 4;
    float cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cc[j] = sca[lc * 8 + j];
    const float* prn = psi_raw + (long long)n * V;
    for (long long v0 = warp_id * vpw * UB; v0 < V; v0 += nwarps * vpw * UB) {
      uint4 ud[UB], ux[UB];
      float pr[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const long long v = v0 + u * vpw + lv;
        const bool ok = v < V;
        pr[u] = ok ? __ldg(prn + v) : 0.f;
        ud[u] = ok ? ldg16_stream(don + v * lddo + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        ux[u] = ok ? ldg16_stream(xn + v * ldx + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const long long v = v0 + u * vpw + lv;
        const float s = 1.f / (1.f + __expf(-fmaf(pr[u], a, b)));
        const float xh = (pr[u] - mu) * rstd;
        float d[8], xv[8], o[8];
        unpack8(ud[u], d); unpack8(ux[u], xv);
        float ds = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dxv = d[j] * xv[j];
          ds = fmaf(dxv, cc[j], ds);
          o[j] = d[j] * s * cc[j];
          acc_ca[j] = fmaf(dxv, s, acc_ca[j]);
        }
        if (v < V) stg16(dxn + v * lddx + lc * 8, pack8(o));
        for (int o2 = lanes_c >> 1; o2 > 0; o2 >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o2);
        if (lc == 0 && v < V) {
          const float dpn = ds * s * (1.f - s);
          dpsin[(long long)n * V + v] = dpn;
          r1 += dpn; r2 += dpn * xh;
        }
      }
    }
  } else
  for (long long v0 = warp_id * vpw; v0 < V; v0 += nwarps * vpw) {
    const long long v = v0 + lv;
    float ds = 0.f, s = 0.f, xh = 0.f;
    if (v < V) {
      const float pr = __ldg(psi_raw + (long long)n * V + v);
      s = 1.f / (1.f + __expf(-fmaf(pr, a, b)));
      xh = (pr - mu) * rstd;
      for (int c8 = lc; c8 < C8; c8 += lanes_c) {
        float d[8], xv[8], o[8];
        unpack8(ldg16_stream(don + v * lddo + c8 * 8), d);
        unpack8(ldg16_stream(xn + v * ldx + c8 * 8), xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float cc = sca[c8 * 8 + j];
          const float dxv = d[j] * xv[j];
          ds = fmaf(dxv, cc, ds);
          o[j] = d[j] * s * cc;
          if (nchunks == 1) acc_ca[j] = fmaf(dxv, s, acc_ca[j]);
          else atomicAdd(&sdca[c8 * 8 + j], (double)(dxv * s));
        }
        stg16(dxn + v * lddx + c8 * 8, pack8(o));
      }
    }
    for (int o = lanes_c >> 1; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
    if (lc == 0 && v < V) {
      const float dpn = ds * s * (1.f - s);
      dpsin[(long long)n * V + v] = dpn;
      r1 += dpn; r2 += dpn * xh;
    }
  }
  if (nchunks == 1) {   // lanes with the same lc own the same chunk: fold them first (lanes_c is a power of two)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_ca[j] = warp_sum_mod(acc_ca[j], lanes_c);
    if (lane < lanes_c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sdca[lc * 8 + j], (double)acc_ca[j]);
    }
  }
  r1 = warp_sum(r1); r2 = warp_sum(r2);
  if (lane == 0) { atomicAdd(&s_red[0], (double)r1); atomicAdd(&s_red[1], (double)r2); }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&dca[(long long)n * C + c], sdca[c]);
  if (threadIdx.x < 2) atomicAdd(&st_dpsi[(long long)n * 2 + threadIdx.x], s_red[threadIdx.x]);
}

// SE backward (one block per sample).  Param grads are accumulated with fp32 atomics into caller-zeroed buffers.
// xadd[n][c] = d(mean_c)/V : the constant the SE branch adds to every voxel of dx.
__global__ void __launch_bounds__(512) gate_se_bwd_kernel(const double* __restrict__ dca, const float* __restrict__ ca, const float* __restrict__ zbuf,
                                   const float* __restrict__ meanbuf, const float* __restrict__ w1, const float* __restrict__ w2,
                                   long long V, float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                                   float* __restrict__ db2, float* __restrict__ xadd, int C) {
  extern __shared__ float sm[];
  const int R = C / 8;
  float* dp2 = sm; float* dp1 = sm + C; float* z = sm + C + R; float* mean = sm + C + 2 * R;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float cc = ca[(long long)n * C + c];
    dp2[c] = (float)dca[(long long)n * C + c] * cc * (1.f - cc);
    mean[c] = meanbuf[(long long)n * C + c];
    atomicAdd(&db2[c], dp2[c]);
  }
  for (int j = threadIdx.x; j < R; j += blockDim.x) z[j] = zbuf[(long long)n * R + j];
  __syncthreads();
  for (int i = threadIdx.x; i < C * R; i += blockDim.x) {
    const int c = i / R, j = i - c * R;
    atomicAdd(&dW2[i], dp2[c] * z[j]);
  }
  // dp1[j] = Σ_c W2[c][j] dp2[c]: warps split the rows c, lanes the (contiguous) columns j; the per-warp partials are summed
  // in FIXED order (this feeds an input gradient: no order-dependent atomics)
  float* part = mean + C;   // [nw][R]
  for (int j0 = 0; j0 < R; j0 += 32) {
    const int j = j0 + lane;
    float a = 0.f;
    if (j < R) {
#pragma unroll 8
      for (int c = wrp; c < C; c += nw) a = fmaf(__ldg(w2 + (long long)c * R + j), dp2[c], a);
      part[wrp * R + j] = a;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < nw; ++w) acc += part[w * R + j];
    const float v = z[j] > 0.f ? acc : 0.f;
    dp1[j] = v;
    atomicAdd(&db1[j], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    atomicAdd(&dW1[i], dp1[j] * mean[c]);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
#pragma unroll 8
    for (int j = 0; j < R; ++j) a = fmaf(__ldg(w1 + (long long)j * C + c), dp1[j], a);
    xadd[(long long)n * C + c] = a / (float)V;
  }
}

// ------------------------------------------------------------------------------------------------
// backward 2: dψr = GN1-backward(dψn) ; dz[c] = dψr * wψ[c] * [q_c>0]  (bf16 [N][V][F]) ;
//   sums_g[n][c] += (Σdz, Σdz*x̂g) ; sums_x[n][c] += (Σdz, Σdz*x̂x) ; dwψ[c] += Σ dψr*q_c ; dbψ += Σ dψr
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) gate_psi_bwd_kernel(
    const float* __restrict__ dpsin, const float* __restrict__ psi_raw, const double* __restrict__ st_psi,
    const double* __restrict__ st_dpsi, const float* __restrict__ gpsi, const bf16* __restrict__ g1r,
    const bf16* __restrict__ x1r, const double* __restrict__ st_g, const double* __restrict__ st_x,
    const float* __restrict__ gam_g, const float* __restrict__ bet_g, const float* __restrict__ gam_x,
    const float* __restrict__ bet_x, const float* __restrict__ wpsi, bf16* __restrict__ dz, double* __restrict__ sums_g,
    double* __restrict__ sums_x, float* __restrict__ dwpsi, float* __restrict__ dbpsi, long long V, int F, float eps) {
  extern __shared__ float sm[];
  float* mg = sm; float* rg = sm + F; float* mx = sm + 2 * F; float* rx = sm + 3 * F;
  float* gg = sm + 4 * F; float* gx = sm + 5 * F; float* sh = sm + 6 * F; float* wp = sm + 7 * F;
  double* red = reinterpret_cast<double*>(sm + 8 * F);  // [4][F]: Σdz, Σdz*x̂g, Σdz*x̂x, Σdψr*q
  __shared__ double s_db;
  const int n = blockIdx.y;
  const int cpg = F / 4;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float m1, r1, m2, r2;
    mean_rstd_from(st_g + ((long long)n * 4 + c / cpg) * 2, (double)cpg * V, eps, m1, r1);
    mean_rstd_from(st_x + ((long long)n * 4 + c / cpg) * 2, (double)cpg * V, eps, m2, r2);
    mg[c] = m1; rg[c] = r1; mx[c] = m2; rx[c] = r2; gg[c] = gam_g[c]; gx[c] = gam_x[c];
    sh[c] = bet_g[c] + bet_x[c]; wp[c] = wpsi[c];
    red[c] = 0.0; red[F + c] = 0.0; red[2 * F + c] = 0.0; red[3 * F + c] = 0.0;
  }
  if (threadIdx.x == 0) s_db = 0.0;
  float mu, rstd;
  mean_rstd_from(st_psi + (long long)n * 2, (double)V, eps, mu, rstd);
  const float gp = gpsi[0];
  const float m1p = gp * (float)(st_dpsi[(long long)n * 2] / (double)V);
  const float m2p = gp * (float)(st_dpsi[(long long)n * 2 + 1] / (double)V);
  __syncthreads();
  const int F8 = F >> 3;
  const int lanes_c = F8 < 32 ? F8 : 32;
  const int vpw = 32 / lanes_c;
  const int lane = threadIdx.x & 31;
  const int lc = lane % lanes_c, lv = lane / lanes_c;
  const int nchunks = (F8 + lanes_c - 1) / lanes_c;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bf16* gn_ = g1r + (long long)n * V * F;
  const bf16* xn_ = x1r + (long long)n * V * F;
  bf16* dzn = dz + (long long)n * V * F;
  float a0[8], a1[8], a2[8], a3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; }
  float db = 0.f;
  if (nchunks == 1) {
    // every lane owns ONE fixed 8-channel chunk.  Its 64 constants are re-read per voxel from shared memory as 16 float4
    // broadcasts (a chunk-major copy, instead of 64 scalar reads), two voxel groups are in flight, and the arithmetic keeps the
    // association of the generic path (the ReLU mask must agree with the forward's)
    // chunk stride 68 floats, not 64: lanes of different chunks would otherwise all hit the same four banks (an F/8-way
    // conflict on every one of the 16 float4 reads per voxel — the 64^3 / 32^3 gates ran 3-8x over their byte share)
    __shared__ __align__(16) float cst[32 * 68];   // [chunk][kind: mg rg gg mx rx gx sh wp][8] (+4 pad)
    for (int i = threadIdx.x; i < F8 * 64; i += blockDim.x) {
      const int c8 = i >> 6, kind = (i >> 3) & 7, j = i & 7, c = c8 * 8 + j;
      cst[c8 * 68 + (i & 63)] = kind == 0 ? mg[c] : kind == 1 ? rg[c] : kind == 2 ? gg[c] : kind == 3 ? mx[c] : kind == 4 ? rx[c]
             : kind == 5 ? gx[c] : kind == 6 ? sh[c] : wp[c];
    }
    __syncthreads();
    const float* cb = cst + (lc < F8 ? lc : 0) * 68;
    // four voxel groups in flight; the 64 constants of the lane's chunk are re-read two channels at a time (volatile 8-byte
    // loads) so that they never occupy more than 16 registers: 2 CTAs per SM with 8 x 16-byte loads per thread in flight
    constexpr int UP = 4;
    for (long long v0 = warp_id * vpw * UP; v0 < V; v0 += nwarps * vpw * UP) {
      unsigned wa[UP][4], wb[UP][4];
      float prv[UP], dpv[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const long long v = v0 + u * vpw + lv;
        const bool ok = v < V && lc < F8;
        const uint4 ra = ok ? ldg16_stream(gn_ + v * F + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        const uint4 rb = ok ? ldg16_stream(xn_ + v * F + lc * 8) : make_uint4(0u, 0u, 0u, 0u);
        wa[u][0] = ra.x; wa[u][1] = ra.y; wa[u][2] = ra.z; wa[u][3] = ra.w;
        wb[u][0] = rb.x; wb[u][1] = rb.y; wb[u][2] = rb.z; wb[u][3] = rb.w;
        prv[u] = ok ? __ldg(psi_raw + (long long)n * V + v) : 0.f;
        dpv[u] = ok ? __ldg(dpsin + (long long)n * V + v) : 0.f;
      }
      float dpr[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const long long v = v0 + u * vpw + lv;
        const bool ok = v < V && lc < F8;
        const float xh = (prv[u] - mu) * rstd;
        dpr[u] = ok ? rstd * (gp * dpv[u] - m1p - xh * m2p) : 0.f;   // 0: an absent voxel adds nothing to any sum
        if (lc == 0) db += dpr[u];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 kmg = lds2v(cb + 0 * 8 + 2 * q), krg = lds2v(cb + 1 * 8 + 2 * q), kgg = lds2v(cb + 2 * 8 + 2 * q),
                     kmx = lds2v(cb + 3 * 8 + 2 * q), krx = lds2v(cb + 4 * 8 + 2 * q), kgx = lds2v(cb + 5 * 8 + 2 * q),
                     ksh = lds2v(cb + 6 * 8 + 2 * q), kwp = lds2v(cb + 7 * 8 + 2 * q);
#pragma unroll
        for (int u = 0; u < UP; ++u) {
          const float2 a = bfw(wa[u][q]), b = bfw(wb[u][q]);
          const float xg0 = (a.x - kmg.x) * krg.x, xx0 = (b.x - kmx.x) * krx.x;
          const float xg1 = (a.y - kmg.y) * krg.y, xx1 = (b.y - kmx.y) * krx.y;
          const float q0 = fmaf(xg0, kgg.x, fmaf(xx0, kgx.x, ksh.x)), q1 = fmaf(xg1, kgg.y, fmaf(xx1, kgx.y, ksh.y));
          const float dz0 = q0 > 0.f ? dpr[u] * kwp.x : 0.f, dz1 = q1 > 0.f ? dpr[u] * kwp.y : 0.f;
          wa[u][q] = wbf(dz0, dz1);
          a0[2 * q] += dz0; a1[2 * q] = fmaf(dz0, xg0, a1[2 * q]); a2[2 * q] = fmaf(dz0, xx0, a2[2 * q]);
          a3[2 * q] = fmaf(dpr[u], fmaxf(q0, 0.f), a3[2 * q]);
          a0[2 * q + 1] += dz1; a1[2 * q + 1] = fmaf(dz1, xg1, a1[2 * q + 1]); a2[2 * q + 1] = fmaf(dz1, xx1, a2[2 * q + 1]);
          a3[2 * q + 1] = fmaf(dpr[u], fmaxf(q1, 0.f), a3[2 * q + 1]);
        }
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const long long v = v0 + u * vpw + lv;
        if (v < V && lc < F8) stg16(dzn + v * F + lc * 8, make_uint4(wa[u][0], wa[u][1], wa[u][2], wa[u][3]));
      }
    }
  } else
  for (long long v0 = warp_id * vpw; v0 < V; v0 += nwarps * vpw) {
    const long long v = v0 + lv;
    if (v >= V) continue;
    const float pr = __ldg(psi_raw + (long long)n * V + v);
    const float xh = (pr - mu) * rstd;
    const float dpr = rstd * (gp * __ldg(dpsin + (long long)n * V + v) - m1p - xh * m2p);
    if (lc == 0) db += dpr;
    for (int c8 = lc; c8 < F8; c8 += lanes_c) {
      float a[8], b[8], o[8];
      unpack8(ldg16_stream(gn_ + v * F + c8 * 8), a);
      unpack8(ldg16_stream(xn_ + v * F + c8 * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c8 * 8 + j;
        const float xg = (a[j] - mg[c]) * rg[c], xx = (b[j] - mx[c]) * rx[c];
        const float q = fmaf(xg, gg[c], fmaf(xx, gx[c], sh[c]));
        const float dzv = q > 0.f ? dpr * wp[c] : 0.f;
        o[j] = dzv;
        const float qq = fmaxf(q, 0.f);
        if (nchunks == 1) {
          a0[j] += dzv; a1[j] = fmaf(dzv, xg, a1[j]); a2[j] = fmaf(dzv, xx, a2[j]); a3[j] = fmaf(dpr, qq, a3[j]);
        } else {
          atomicAdd(&red[c], (double)dzv); atomicAdd(&red[F + c], (double)(dzv * xg)); atomicAdd(&red[2 * F + c], (double)(dzv * xx));
          atomicAdd(&red[3 * F + c], (double)(dpr * qq));
        }
      }
      stg16(dzn + v * F + c8 * 8, pack8(o));
    }
  }
  if (nchunks == 1) {   // fold the lanes that own the same chunk (lanes_c is a power of two), then one atomic per warp
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = warp_sum_mod(a0[j], lanes_c); a1[j] = warp_sum_mod(a1[j], lanes_c);
      a2[j] = warp_sum_mod(a2[j], lanes_c); a3[j] = warp_sum_mod(a3[j], lanes_c);
    }
    if (lane < lanes_c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[lc * 8 + j], (double)a0[j]); atomicAdd(&red[F + lc * 8 + j], (double)a1[j]);
        atomicAdd(&red[2 * F + lc * 8 + j], (double)a2[j]); atomicAdd(&red[3 * F + lc * 8 + j], (double)a3[j]);
      }
    }
  }
  db = warp_sum(db);
  if (lane == 0) atomicAdd(&s_db, (double)db);
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    atomicAdd(&sums_g[((long long)n * F + c) * 2], red[c]);
    atomicAdd(&sums_g[((long long)n * F + c) * 2 + 1], red[F + c]);
    atomicAdd(&sums_x[((long long)n * F + c) * 2], red[c]);
    atomicAdd(&sums_x[((long long)n * F + c) * 2 + 1], red[2 * F + c]);
    atomicAdd(&dwpsi[c], (float)red[3 * F + c]);
  }
  if (threadIdx.x == 0) atomicAdd(dbpsi, (float)s_db);
}

// dγψ = Σ_n Σ_v dψn*x̂ψ ; dβψ = Σ_n Σ_v dψn
__global__ void gate_psi_norm_grad_kernel(const double* __restrict__ st_dpsi, int N, float* dgpsi, float* dbpsi_n) {
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int n = 0; n < N; ++n) { b += st_dpsi[2 * n]; a += st_dpsi[2 * n + 1]; }
    dgpsi[0] = (float)a; dbpsi_n[0] = (float)b;
  }
}

// dx[n][v][c] += xadd[n][c]   (SE-branch contribution, constant per channel)
__global__ void __launch_bounds__(256) add_channel_const_kernel(bf16* __restrict__ dx, long long lddx, const float* __restrict__ xadd,
                                                                long long V, int C) {
  extern __shared__ float sa[];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sa[c] = xadd[(long long)n * C + c];
  __syncthreads();
  const int C8 = C >> 3;
  const long long total = V * C8;
  bf16* dxn = dx + (long long)n * V * lddx;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const int c0 = (int)(i - vox * C8) * 8;
    float v[8];
    unpack8(ldg16(dxn + vox * lddx + c0), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += sa[c0 + j];
    stg16(dxn + vox * lddx + c0, pack8(v));
  }
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" {

int b3d_gate_psi_fwd(const void* g1r, const void* x1r, const double* st_g, const double* st_x, const float* gam_g,
                     const float* bet_g, const float* gam_x, const float* bet_x, const float* wpsi, const float* bpsi,
                     float* psi_raw, double* st_psi, int N, long long V, int F, float eps, void* stream) {
  B3D_REQUIRE(F % 8 == 0 && F % 4 == 0 && (pow2(F / 8) || (F / 8) % 32 == 0) && F <= 1024, "gate_psi_fwd: F=%d unsupported", F);
  dim3 grid(gt_blocks_per_sample(V * 8, 256, N), N);
  gate_psi_fwd_kernel<<<grid, 256, 4 * F * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)g1r, (const bf16*)x1r, st_g, st_x, gam_g, bet_g, gam_x, bet_x, wpsi, 0.f, bpsi, psi_raw, st_psi, V, F, eps); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gate_se_fwd(const double* xsum, long long V, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* ca, float* zbuf, float* meanbuf, int N, int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= 4096, "gate_se_fwd: bad C");
  gate_se_fwd_kernel<<<N, 512, (C + C / 8) * sizeof(float), (cudaStream_t)stream>>>(xsum, V, w1, b1, w2, b2, ca, zbuf, meanbuf, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gate_apply_fwd(const void* x, long long ldx, const float* psi_raw, const double* st_psi, const float* gpsi,
                       const float* bpsi_n, const float* ca, void* out, long long ldo, int N, long long V, int C,
                       float eps, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= 4096, "gate_apply_fwd: bad C");
  dim3 grid(gt_blocks_per_sample(V * (C / 8), 512, N), N);
  gate_apply_fwd_kernel<<<grid, 256, C * sizeof(float), (cudaStream_t)stream>>>((const bf16*)x, ldx, psi_raw, st_psi, gpsi, bpsi_n,
                                                                               ca, (bf16*)out, ldo, V, C, eps); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gate_apply_bwd(const void* dout, long long lddo, const void* x, long long ldx, const float* psi_raw,
                       const double* st_psi, const float* gpsi, const float* bpsi_n, const float* ca, void* dx,
                       long long lddx, float* dpsin, double* dca, double* st_dpsi, int N, long long V, int C, float eps,
                       void* stream) {
  B3D_REQUIRE(C % 8 == 0 && (pow2(C / 8) || (C / 8) % 32 == 0) && C <= 4096, "gate_apply_bwd: C=%d unsupported", C);
  // (fewer, longer CTAs at the small levels measured slower: 46.7 -> 77.3 us at 64^3 — the kernel wants thread-level parallelism)
  dim3 grid(gt_blocks_per_sample(V * 8, 256, N), N);
  gate_apply_bwd_kernel<<<grid, 256, 3 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)dout, lddo, (const bf16*)x, ldx, psi_raw, st_psi, gpsi, bpsi_n, ca, (bf16*)dx, lddx, dpsin, dca, st_dpsi, V, C, eps); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gate_se_bwd(const double* dca, const float* ca, const float* zbuf, const float* meanbuf, const float* w1,
                    const float* w2, long long V, float* dW1, float* db1, float* dW2, float* db2, float* xadd, int N, int C,
                    void* stream) {
  gate_se_bwd_kernel<<<N, 512, (2 * C + 18 * (C / 8)) * sizeof(float), (cudaStream_t)stream>>>(dca, ca, zbuf, meanbuf, w1, w2, V, dW1,
                                                                                             db1, dW2, db2, xadd, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_gate_psi_bwd(const float* dpsin, const float* psi_raw, const double* st_psi, const double* st_dpsi,
                     const float* gpsi, const void* g1r, const void* x1r, const double* st_g, const double* st_x,
                     const float* gam_g, const float* bet_g, const float* gam_x, const float* bet_x, const float* wpsi,
                     void* dz, double* sums_g, double* sums_x, float* dwpsi, float* dbpsi, float* dgpsi, float* dbpsi_n,
                     int N, long long V, int F, float eps, void* stream) {
  B3D_REQUIRE(F % 8 == 0 && (pow2(F / 8) || (F / 8) % 32 == 0) && F <= 1024, "gate_psi_bwd: F=%d unsupported", F);
  // >= 2 iterations per warp (a warp iteration covers 4 groups of 32 / min(F/8, 32) voxels), capped at 8 CTAs per SM: every CTA
  // ends with 4F fp64 atomics on the same addresses, which dominated the 32^3 level (50 us for a 14 us byte share) when all
  // 592 CTAs per sample were launched for 2048 warp iterations.  (Level 0 reaches the cap; a single resident wave of long CTAs
  // measured 30 % slower there.)
  const int vpw_h = 32 / std::min(F / 8, 32);
  dim3 grid(gt_blocks_per_sample(V, 256 / 32 * vpw_h * 4 * 2, N), N);
  static const cudaError_t attr = cudaFuncSetAttribute(gate_psi_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * (int)sizeof(float));   // one-time, thread-safe
  B3D_CHECK_CUDA(attr);
  gate_psi_bwd_kernel<<<grid, 256, 16 * F * sizeof(float), (cudaStream_t)stream>>>(
      dpsin, psi_raw, st_psi, st_dpsi, gpsi, (const bf16*)g1r, (const bf16*)x1r, st_g, st_x, gam_g, bet_g, gam_x, bet_x, wpsi,
      (bf16*)dz, sums_g, sums_x, dwpsi, dbpsi, V, F, eps); ++g_b3d_launches;
  gate_psi_norm_grad_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st_dpsi, N, dgpsi, dbpsi_n); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_add_channel_const(void* dx, long long lddx, const float* xadd, int N, long long V, int C, void* stream) {
  B3D_REQUIRE(C % 8 == 0 && C <= 4096, "add_channel_const: bad C");
  dim3 grid(gt_blocks_per_sample(V * (C / 8), 512, N), N);
  add_channel_const_kernel<<<grid, 256, C * sizeof(float), (cudaStream_t)stream>>>((bf16*)dx, lddx, xadd, V, C); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
