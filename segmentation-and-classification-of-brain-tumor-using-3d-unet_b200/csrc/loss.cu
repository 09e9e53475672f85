// loss.cu — softmax-Dice / focal / cross-entropy / boundary / Tversky losses and the Dice / voxel-count metrics.
//
// Reference semantics (4 classes, fp32 logits NCDHW, int64 targets):
//   losses.CombinedLoss3D     /root/reference/losses.py:7-75     0.5*dice(1e-5) + 0.3*focal(0.25,2) + 0.2*boundary
//   losses.TverskyLoss3D      /root/reference/losses.py:77-97
//   training.CombinedLoss     /root/reference/training.py:517-566  0.5*dice(1e-6) + 0.3*CE + 0.2*focal(1,2)
//   calculate_dice_score      /root/reference/training.py:351-364  (argmax + integer counts -> confusion histogram)
//   voxel counts              /root/reference/main.py:470-474,588-591
// All of it is HBM-bound: per output tensor the forward reads logits+targets once (softmax, all reductions), the
// boundary term is a 7-point stencil pass over the stored softmax, and the backward is one more pass writing dlogits
// (math: SURVEY App. A7).  No host synchronisation: every scalar stays in a device accumulator until the caller reads it.
#include "loss_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

static int ls_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)b3d_num_sms() * 8;
  return (int)std::max<long long>(1, std::min(b, cap));
}

// pass 1: p = softmax(logits) stored planar; acc[n] += I,P,T,Σce,Σfocal
__global__ void __launch_bounds__(256) loss_softmax_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                           float* __restrict__ prob, double* __restrict__ acc, long long V,
                                                           LossCfg cfg) {
  __shared__ double s_acc[14];   // fp64: warp arrival order cannot change the loss sums
  if (threadIdx.x < 14) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  const int n = blockIdx.y;
  const float* ln = logits + (long long)n * KC * V;
  float* pn = prob + (long long)n * KC * V;
  const long long* tn = target + (long long)n * V;
  float aI[KC] = {0, 0, 0, 0}, aP[KC] = {0, 0, 0, 0}, aT[KC] = {0, 0, 0, 0}, ace = 0.f, afo = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    float z[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) z[c] = __ldg(ln + c * V + v);
    const int t = (int)__ldg(tn + v);
    const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    float e[KC], s = 0.f;
#pragma unroll
    for (int c = 0; c < KC; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
    const float inv = 1.f / s;
    const float lse = m + logf(s);
    float zt = 0.f, pt = 0.f;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float p = e[c] * inv;
      pn[c * V + v] = p;
      aP[c] += p;
      if (c == t) { aI[c] += p; aT[c] += 1.f; zt = z[c]; pt = p; }
    }
    const float ce = lse - zt;
    ace += ce;
    const float ptx = expf(-ce);
    (void)pt;
    afo += cfg.f_alpha * focal_pow(1.f - ptx, cfg.f_gamma) * ce;
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    const float a = warp_sum(aI[c]), b = warp_sum(aP[c]), d = warp_sum(aT[c]);
    if (lane == 0) { atomicAdd(&s_acc[c], (double)a); atomicAdd(&s_acc[4 + c], (double)b); atomicAdd(&s_acc[8 + c], (double)d); }
  }
  ace = warp_sum(ace); afo = warp_sum(afo);
  if (lane == 0) { atomicAdd(&s_acc[12], (double)ace); atomicAdd(&s_acc[13], (double)afo); }
  __syncthreads();
  if (threadIdx.x < 14) atomicAdd(&acc[(long long)n * ACC_STRIDE + threadIdx.x], s_acc[threadIdx.x]);
}

// pass 2: E = B(p) − B(onehot) (forward differences, zero at the far face), stored planar; acc[n][14] += ΣE²
__global__ void __launch_bounds__(256) loss_boundary_kernel(const float* __restrict__ prob, const long long* __restrict__ target,
                                                            float* __restrict__ E, double* __restrict__ acc, int D, int H, int W) {
  __shared__ double s_acc;
  if (threadIdx.x == 0) s_acc = 0.0;
  __syncthreads();
  const long long V = (long long)D * H * W;
  const int n = blockIdx.y;
  const float* pn = prob + (long long)n * KC * V;
  float* En = E + (long long)n * KC * V;
  const long long* tn = target + (long long)n * V;
  float a = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(v % W);
    const int y = (int)((v / W) % H);
    const int z = (int)(v / ((long long)W * H));
    const bool hz = z + 1 < D, hy = y + 1 < H, hx = x + 1 < W;
    const long long vz = v + (long long)H * W, vy = v + W, vx = v + 1;
    const int t = (int)__ldg(tn + v);
    const int tz = hz ? (int)__ldg(tn + vz) : -1, ty = hy ? (int)__ldg(tn + vy) : -1, tx = hx ? (int)__ldg(tn + vx) : -1;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float p = __ldg(pn + c * V + v);
      float bp = 0.f, bo = 0.f;
      if (hz) { bp += fabsf(__ldg(pn + c * V + vz) - p); bo += ((tz == c) != (t == c)) ? 1.f : 0.f; }
      if (hy) { bp += fabsf(__ldg(pn + c * V + vy) - p); bo += ((ty == c) != (t == c)) ? 1.f : 0.f; }
      if (hx) { bp += fabsf(__ldg(pn + c * V + vx) - p); bo += ((tx == c) != (t == c)) ? 1.f : 0.f; }
      const float e = bp - bo;
      En[c * V + v] = e;
      a = fmaf(e, e, a);
    }
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc, (double)a);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&acc[(long long)n * ACC_STRIDE + 14], s_acc);
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// pass 2, 4 voxels (consecutive x) per thread with 128-bit loads / stores (W % 4 == 0)
__global__ void __launch_bounds__(256) loss_boundary_vec4_kernel(const float* __restrict__ prob, const long long* __restrict__ target,
                                                                 float* __restrict__ E, double* __restrict__ acc, int D, int H, int W) {
  __shared__ double s_acc;
  if (threadIdx.x == 0) s_acc = 0.0;
  __syncthreads();
  const long long HW = (long long)H * W, V = (long long)D * HW;
  const int n = blockIdx.y;
  const float* pn = prob + (long long)n * KC * V;
  float* En = E + (long long)n * KC * V;
  const long long* tn = target + (long long)n * V;
  float a = 0.f;
  const long long V4 = V / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < V4; q += (long long)gridDim.x * blockDim.x) {
    const long long v = q * 4;
    const int x0 = (int)(v % W);
    const int y = (int)((v / W) % H);
    const int z = (int)(v / HW);
    const bool hz = z + 1 < D, hy = y + 1 < H, hxp = x0 + 4 < W;
    int t[4], tz[4], ty[4];
    {
      const longlong2 ta = __ldg(reinterpret_cast<const longlong2*>(tn + v)), tb = __ldg(reinterpret_cast<const longlong2*>(tn + v + 2));
      t[0] = (int)ta.x; t[1] = (int)ta.y; t[2] = (int)tb.x; t[3] = (int)tb.y;
      if (hz) { const longlong2 a2 = __ldg(reinterpret_cast<const longlong2*>(tn + v + HW)), b2 = __ldg(reinterpret_cast<const longlong2*>(tn + v + HW + 2));
                tz[0] = (int)a2.x; tz[1] = (int)a2.y; tz[2] = (int)b2.x; tz[3] = (int)b2.y; }
      if (hy) { const longlong2 a2 = __ldg(reinterpret_cast<const longlong2*>(tn + v + W)), b2 = __ldg(reinterpret_cast<const longlong2*>(tn + v + W + 2));
                ty[0] = (int)a2.x; ty[1] = (int)a2.y; ty[2] = (int)b2.x; ty[3] = (int)b2.y; }
    }
    const int txp = hxp ? (int)__ldg(tn + v + 4) : -1;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float* pc = pn + c * V + v;
      const float4 p4 = ld4(pc);
      const float p[4] = {p4.x, p4.y, p4.z, p4.w};
      float bp[4] = {0.f, 0.f, 0.f, 0.f}, bo[4] = {0.f, 0.f, 0.f, 0.f};
      if (hz) { const float4 q4 = ld4(pc + HW); const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { bp[i] += fabsf(qv[i] - p[i]); bo[i] += ((tz[i] == c) != (t[i] == c)) ? 1.f : 0.f; } }
      if (hy) { const float4 q4 = ld4(pc + W); const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { bp[i] += fabsf(qv[i] - p[i]); bo[i] += ((ty[i] == c) != (t[i] == c)) ? 1.f : 0.f; } }
      const float pxp = hxp ? __ldg(pc + 4) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < 3 || hxp) {
          const float nb = (i < 3) ? p[i < 3 ? i + 1 : 3] : pxp;
          const int tn1 = (i < 3) ? t[i < 3 ? i + 1 : 3] : txp;
          bp[i] += fabsf(nb - p[i]); bo[i] += ((tn1 == c) != (t[i] == c)) ? 1.f : 0.f;
        }
      }
      float e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { e[i] = bp[i] - bo[i]; a = fmaf(e[i], e[i], a); }
      *reinterpret_cast<float4*>(En + c * V + v) = make_float4(e[0], e[1], e[2], e[3]);
    }
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc, (double)a);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&acc[(long long)n * ACC_STRIDE + 14], s_acc);
}

// values[0..5] = total, dice, focal, boundary, ce, tversky     (tiny, one thread)
__global__ void loss_finalize_kernel(const double* __restrict__ acc, int N, long long V, LossCfg cfg, float* __restrict__ values) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double dice = 0, tv = 0, ce = 0, fo = 0, bd = 0;
  for (int n = 0; n < N; ++n) {
    const double* a = acc + (long long)n * ACC_STRIDE;
    for (int c = 0; c < KC; ++c) {
      const double I = a[c], P = a[4 + c], T = a[8 + c];
      dice += (2.0 * I + cfg.smooth) / (P + T + cfg.smooth);
      tv += (I + cfg.tv_smooth) / (I + cfg.tv_alpha * (P - I) + cfg.tv_beta * (T - I) + cfg.tv_smooth);
    }
    ce += a[12]; fo += a[13]; bd += a[14];
  }
  const double nv = (double)N * (double)V;
  const double dl = 1.0 - dice / (N * KC), tl = 1.0 - tv / (N * KC);
  const double cel = ce / nv, fl = fo / nv, bl = bd / (nv * KC);
  values[0] = (float)(cfg.w_dice * dl + cfg.w_focal * fl + cfg.w_ce * cel + cfg.w_boundary * bl + cfg.w_tv * tl);
  values[1] = (float)dl; values[2] = (float)fl; values[3] = (float)bl; values[4] = (float)cel; values[5] = (float)tl;
}

// backward: dlogits = gscale[0]*wscale * d(total)/d(logits)
__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ E,
                                                       const long long* __restrict__ target, const double* __restrict__ acc,
                                                       const float* __restrict__ gscale, float wscale, float* __restrict__ dlogits,
                                                       int N, int D, int H, int W, LossCfg cfg) {
  __shared__ float s_d1[KC], s_d2[KC], s_t1[KC], s_t2[KC], s_t3[KC];
  const long long V = (long long)D * H * W;
  const int n = blockIdx.y;
  if (threadIdx.x < KC) {
    const double* a = acc + (long long)n * ACC_STRIDE;
    const int c = threadIdx.x;
    const double I = a[c], P = a[4 + c], T = a[8 + c];
    const double U = P + T + cfg.smooth;
    // d(dice loss)/dp = -(1/NC) * [2*o*U - (2I+s)] / U^2  = o*s_d1 + s_d2
    s_d1[c] = (float)(-(2.0 / U) / (N * KC));
    s_d2[c] = (float)(((2.0 * I + cfg.smooth) / (U * U)) / (N * KC));
    // tversky: R = (I+s)/Dn, Dn = (1-a-b)I + aP + bT + s ; dR/dp = [o*Dn - (I+s)*((1-a-b)*o + a)]/Dn^2
    const double Dn = (1.0 - cfg.tv_alpha - cfg.tv_beta) * I + cfg.tv_alpha * P + cfg.tv_beta * T + cfg.tv_smooth;
    s_t1[c] = (float)(-(1.0 / Dn) / (N * KC));                                                       // * o
    s_t2[c] = (float)(((I + cfg.tv_smooth) * (1.0 - cfg.tv_alpha - cfg.tv_beta) / (Dn * Dn)) / (N * KC));  // * o
    s_t3[c] = (float)(((I + cfg.tv_smooth) * cfg.tv_alpha / (Dn * Dn)) / (N * KC));                  // const
  }
  __syncthreads();
  const float g = gscale ? gscale[0] * wscale : wscale;
  const float inv_nv = 1.f / ((float)N * (float)V);
  const float wb = cfg.w_boundary * 2.f * inv_nv / KC;
  const float* pn = prob + (long long)n * KC * V;
  const float* En = E ? E + (long long)n * KC * V : nullptr;
  const long long* tn = target + (long long)n * V;
  float* dn = dlogits + (long long)n * KC * V;
  const bool use_b = (cfg.w_boundary != 0.f) && En != nullptr;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const int t = (int)__ldg(tn + v);
    float p[KC], G[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      p[c] = __ldg(pn + c * V + v);
      const float o = (c == t) ? 1.f : 0.f;
      G[c] = cfg.w_dice * (o * s_d1[c] + s_d2[c]) + cfg.w_tv * (o * (s_t1[c] + s_t2[c]) + s_t3[c]);
    }
    if (use_b) {
      const int x = (int)(v % W);
      const int y = (int)((v / W) % H);
      const int z = (int)(v / ((long long)W * H));
      const long long st[3] = {(long long)H * W, (long long)W, 1};
      const bool hp[3] = {z + 1 < D, y + 1 < H, x + 1 < W};
      const bool hm[3] = {z > 0, y > 0, x > 0};
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float ev = __ldg(En + c * V + v);
        float gb = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (hp[a]) {  // term |p(v+e) - p(v)| at u = v : d/dp(v) = -sign
            const float d = __ldg(pn + c * V + v + st[a]) - p[c];
            gb -= (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * ev;
          }
          if (hm[a]) {  // term at u = v-e : d/dp(v) = +sign(p(v) - p(v-e)) * E(v-e)
            const float d = p[c] - __ldg(pn + c * V + v - st[a]);
            gb += (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * __ldg(En + c * V + v - st[a]);
          }
        }
        G[c] = fmaf(wb, gb, G[c]);
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < KC; ++c) dot = fmaf(G[c], p[c], dot);
    const float pt = fmaxf(p[t < 0 ? 0 : (t >= KC ? KC - 1 : t)], 1e-38f);
    const float ce = -logf(pt);
    const float om = 1.f - pt;
    float fprime = 0.f;
    if (cfg.w_focal != 0.f) {
      const float gm = cfg.f_gamma;
      const float t1 = focal_pow(om, gm);
      const float t2 = (gm == 0.f) ? 0.f : gm * focal_pow(om, gm - 1.f) * pt * ce;
      fprime = cfg.w_focal * cfg.f_alpha * (t1 + t2) * inv_nv;
    }
    const float lin = fprime + cfg.w_ce * inv_nv;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float o = (c == t) ? 1.f : 0.f;
      dn[c * V + v] = g * (p[c] * (G[c] - dot) + lin * (p[c] - o));
    }
  }
}


// backward, 4 voxels (consecutive x) per thread: the 7-point boundary stencil is served by 128-bit loads (W % 4 == 0)

__global__ void __launch_bounds__(256) loss_bwd_vec4_kernel(const float* __restrict__ prob, const float* __restrict__ E,
                                                            const long long* __restrict__ target, const double* __restrict__ acc,
                                                            const float* __restrict__ gscale, float wscale,
                                                            float* __restrict__ dlogits, int N, int D, int H, int W, LossCfg cfg) {
  __shared__ float s_d1[KC], s_d2[KC], s_t1[KC], s_t2[KC], s_t3[KC];
  const long long V = (long long)D * H * W;
  const int n = blockIdx.y;
  if (threadIdx.x < KC) {
    const double* a = acc + (long long)n * ACC_STRIDE;
    const int c = threadIdx.x;
    const double I = a[c], P = a[4 + c], T = a[8 + c];
    const double U = P + T + cfg.smooth;
    s_d1[c] = (float)(-(2.0 / U) / (N * KC));
    s_d2[c] = (float)(((2.0 * I + cfg.smooth) / (U * U)) / (N * KC));
    const double Dn = (1.0 - cfg.tv_alpha - cfg.tv_beta) * I + cfg.tv_alpha * P + cfg.tv_beta * T + cfg.tv_smooth;
    s_t1[c] = (float)(-(1.0 / Dn) / (N * KC));
    s_t2[c] = (float)(((I + cfg.tv_smooth) * (1.0 - cfg.tv_alpha - cfg.tv_beta) / (Dn * Dn)) / (N * KC));
    s_t3[c] = (float)(((I + cfg.tv_smooth) * cfg.tv_alpha / (Dn * Dn)) / (N * KC));
  }
  __syncthreads();
  const float g = gscale ? gscale[0] * wscale : wscale;
  const float inv_nv = 1.f / ((float)N * (float)V);
  const float wb = cfg.w_boundary * 2.f * inv_nv / KC;
  const float* pn = prob + (long long)n * KC * V;
  const float* En = E ? E + (long long)n * KC * V : nullptr;
  const long long* tn = target + (long long)n * V;
  float* dn = dlogits + (long long)n * KC * V;
  const bool use_b = (cfg.w_boundary != 0.f) && En != nullptr;
  const long long HW = (long long)H * W;
  const long long V4 = V / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < V4; q += (long long)gridDim.x * blockDim.x) {
    const long long v = q * 4;
    const int x0 = (int)(v % W);
    const int y = (int)((v / W) % H);
    const int z = (int)(v / HW);
    const longlong2 ta = __ldg(reinterpret_cast<const longlong2*>(tn + v));
    const longlong2 tb = __ldg(reinterpret_cast<const longlong2*>(tn + v + 2));
    const int t[4] = {(int)ta.x, (int)ta.y, (int)tb.x, (int)tb.y};
    float p[KC][4], G[KC][4];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float4 pc = ld4(pn + c * V + v);
      p[c][0] = pc.x; p[c][1] = pc.y; p[c][2] = pc.z; p[c][3] = pc.w;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float o = (c == t[i]) ? 1.f : 0.f;
        G[c][i] = cfg.w_dice * (o * s_d1[c] + s_d2[c]) + cfg.w_tv * (o * (s_t1[c] + s_t2[c]) + s_t3[c]);
      }
    }
    if (use_b) {
      const bool hzp = z + 1 < D, hzm = z > 0, hyp = y + 1 < H, hym = y > 0, hxp = x0 + 4 < W, hxm = x0 > 0;
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float* pc = pn + c * V + v;
        const float* ec = En + c * V + v;
        const float4 e4 = ld4(ec);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        float gb[4] = {0.f, 0.f, 0.f, 0.f};
        // z and y axes: whole-vector neighbours
        if (hzp) { const float4 q4 = ld4(pc + HW); const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) gb[i] -= sgnf(qv[i] - p[c][i]) * ev[i]; }
        if (hzm) { const float4 q4 = ld4(pc - HW); const float4 f4 = ld4(ec - HW);
          const float qv[4] = {q4.x, q4.y, q4.z, q4.w}; const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) gb[i] += sgnf(p[c][i] - qv[i]) * fv[i]; }
        if (hyp) { const float4 q4 = ld4(pc + W); const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) gb[i] -= sgnf(qv[i] - p[c][i]) * ev[i]; }
        if (hym) { const float4 q4 = ld4(pc - W); const float4 f4 = ld4(ec - W);
          const float qv[4] = {q4.x, q4.y, q4.z, q4.w}; const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) gb[i] += sgnf(p[c][i] - qv[i]) * fv[i]; }
        // x axis: neighbours inside the vector, plus one scalar on each side
        const float pxp = hxp ? __ldg(pc + 4) : 0.f;
        const float pxm = hxm ? __ldg(pc - 1) : 0.f;
        const float exm = hxm ? __ldg(ec - 1) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < 3 || hxp) { const float nb = (i < 3) ? p[c][i < 3 ? i + 1 : 3] : pxp; gb[i] -= sgnf(nb - p[c][i]) * ev[i]; }
          if (i > 0 || hxm) { const float nb = (i > 0) ? p[c][i > 0 ? i - 1 : 0] : pxm; const float en = (i > 0) ? ev[i > 0 ? i - 1 : 0] : exm;
            gb[i] += sgnf(p[c][i] - nb) * en; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) G[c][i] = fmaf(wb, gb[i], G[c][i]);
      }
    }
    float out[KC][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < KC; ++c) dot = fmaf(G[c][i], p[c][i], dot);
      float ptv = p[0][i];
#pragma unroll
      for (int c = 1; c < KC; ++c) ptv = (t[i] == c || (c == KC - 1 && t[i] >= KC)) ? p[c][i] : ptv;
      const float pt = fmaxf(ptv, 1e-38f);
      const float ce = -logf(pt);
      const float om = 1.f - pt;
      float fprime = 0.f;
      if (cfg.w_focal != 0.f) {
        const float gm = cfg.f_gamma;
        const float t1 = focal_pow(om, gm);
        const float t2 = (gm == 0.f) ? 0.f : gm * focal_pow(om, gm - 1.f) * pt * ce;
        fprime = cfg.w_focal * cfg.f_alpha * (t1 + t2) * inv_nv;
      }
      const float lin = fprime + cfg.w_ce * inv_nv;
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float o = (c == t[i]) ? 1.f : 0.f;
        out[c][i] = g * (p[c][i] * (G[c][i] - dot) + lin * (p[c][i] - o));
      }
    }
#pragma unroll
    for (int c = 0; c < KC; ++c)
      *reinterpret_cast<float4*>(dn + c * V + v) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
  }
}

// ------------------------------------------------------------------------------------------------
// metrics: argmax mask (first maximum wins, like torch.argmax) + 4x4 confusion histogram H[pred][true] (int64)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                        unsigned char* __restrict__ mask, unsigned long long* __restrict__ hist,
                                                        int N, long long V) {
  __shared__ unsigned int s_h[KC * KC];
  if (threadIdx.x < KC * KC) s_h[threadIdx.x] = 0u;
  __syncthreads();
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    const float* l = logits + n * KC * V + v;
    float best = __ldg(l);
    int arg = 0;
#pragma unroll
    for (int c = 1; c < KC; ++c) {
      const float z = __ldg(l + c * V);
      if (z > best || (z != z && best == best)) { best = z; arg = c; }
    }
    if (mask) mask[i] = (unsigned char)arg;
    if (target) {
      const int t = (int)__ldg(target + i);
      if (t >= 0 && t < KC) atomicAdd(&s_h[arg * KC + t], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < KC * KC && target) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}

// per-class counts [4] and per-slice (LAST axis, as seg[:, :, z] in main.py:588-591) tumour counts [W] of a u8 mask [D][H][W]
__global__ void __launch_bounds__(256) voxel_count_kernel(const unsigned char* __restrict__ mask, long long V, int W,
                                                          unsigned long long* __restrict__ cls, unsigned long long* __restrict__ slices) {
  extern __shared__ unsigned int s_cnt[];  // [4 + W]
  for (int i = threadIdx.x; i < KC + W; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    const int m = mask[i];
    if (m < KC) atomicAdd(&s_cnt[m], 1u);
    if (m > 0) atomicAdd(&s_cnt[KC + (int)(i % W)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KC + W; i += blockDim.x) {
    if (s_cnt[i] == 0u) continue;
    if (i < KC) atomicAdd(&cls[i], (unsigned long long)s_cnt[i]);
    else atomicAdd(&slices[i - KC], (unsigned long long)s_cnt[i]);
  }
}

int b3d_launch_loss_finalize(const double* acc, int N, long long V, const LossCfg& cfg, float* values, cudaStream_t st) {
  loss_finalize_kernel<<<1, 32, 0, st>>>(acc, N, V, cfg, values); ++g_b3d_launches;
  return B3D_OK;
}

extern "C" {

// cfg: 11 floats {w_dice, smooth, w_focal, f_alpha, f_gamma, w_ce, w_boundary, w_tv, tv_alpha, tv_beta, tv_smooth}
// prob / E: fp32 scratch [N][4][V] each (E only touched when w_boundary != 0); acc: double [N][16], zeroed here;
// values: float[6] = total, dice, focal, boundary, ce, tversky.
int b3d_loss_fwd(const float* logits, const long long* target, const float* cfg11, float* prob, float* E, double* acc,
                 float* values, int N, int K, int D, int H, int W, void* stream) {
  B3D_REQUIRE(K == KC, "loss: only %d classes supported (got %d)", KC, K);
  LossCfg cfg;
  memcpy(&cfg, cfg11, sizeof(cfg));
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)D * H * W;
  B3D_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * ACC_STRIDE * N, st));
  dim3 grid(std::max(1, ls_blocks(V, 256) / std::max(1, N)), N);
  loss_softmax_kernel<<<grid, 256, 0, st>>>(logits, target, prob, acc, V, cfg); ++g_b3d_launches;
  if (cfg.w_boundary != 0.f) {
    if (W % 4 == 0 && (((uintptr_t)prob | (uintptr_t)E | (uintptr_t)target) & 15) == 0) {
      dim3 g4(std::max(1, ls_blocks(V / 4, 256) / std::max(1, N)), N);
      loss_boundary_vec4_kernel<<<g4, 256, 0, st>>>(prob, target, E, acc, D, H, W); ++g_b3d_launches;
    } else {
      loss_boundary_kernel<<<grid, 256, 0, st>>>(prob, target, E, acc, D, H, W); ++g_b3d_launches;
    }
  }
  loss_finalize_kernel<<<1, 32, 0, st>>>(acc, N, V, cfg, values); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// gscale: optional device float (upstream gradient of the scalar loss); wscale: host weight (deep-supervision weight)
int b3d_loss_bwd(const float* prob, const float* E, const long long* target, const double* acc, const float* cfg11,
                 const float* gscale, float wscale, float* dlogits, int N, int K, int D, int H, int W, void* stream) {
  B3D_REQUIRE(K == KC, "loss: only %d classes supported (got %d)", KC, K);
  LossCfg cfg;
  memcpy(&cfg, cfg11, sizeof(cfg));
  const long long V = (long long)D * H * W;
  if (W % 4 == 0 && (((uintptr_t)prob | (uintptr_t)dlogits | (uintptr_t)E) & 15) == 0 && ((uintptr_t)target & 15) == 0) {
    dim3 grid(std::max(1, ls_blocks(V / 4, 256) / std::max(1, N)), N);
    loss_bwd_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prob, E, target, acc, gscale, wscale, dlogits, N, D, H, W, cfg); ++g_b3d_launches;
  } else {
    dim3 grid(std::max(1, ls_blocks(V, 256) / std::max(1, N)), N);
    loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prob, E, target, acc, gscale, wscale, dlogits, N, D, H, W, cfg); ++g_b3d_launches;
  }
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// hist: uint64[16] (zeroed by caller), mask: optional uint8 [N][V]
int b3d_confusion(const float* logits, const long long* target, unsigned char* mask, unsigned long long* hist, int N,
                  int K, long long V, void* stream) {
  B3D_REQUIRE(K == KC, "confusion: only %d classes supported (got %d)", KC, K);
  confusion_kernel<<<ls_blocks((long long)N * V, 256), 256, 0, (cudaStream_t)stream>>>(logits, target, mask, hist, N, V); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

int b3d_voxel_counts(const unsigned char* mask, long long V, int W, unsigned long long* cls, unsigned long long* slices,
                     void* stream) {
  B3D_REQUIRE(W <= 4096, "voxel_counts: W too large");
  voxel_count_kernel<<<ls_blocks(V, 256), 256, (KC + W) * sizeof(unsigned int), (cudaStream_t)stream>>>(mask, V, W, cls, slices); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
