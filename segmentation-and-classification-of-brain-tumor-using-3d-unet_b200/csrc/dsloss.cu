// dsloss.cu — fused deep-supervision loss (SURVEY K-dsloss, App. A5 + A7).
//
// Reference semantics: UNet3D up-samples every deep-supervision head to full resolution with
// F.interpolate(mode="trilinear", align_corners=False) (/root/reference/main.py:164-171) and DeepSupervisionLoss3D evaluates
// CombinedLoss3D (dice + focal + boundary, /root/reference/losses.py:17-75) on each map (/root/reference/losses.py:107-126).
// Materialised, every output costs a [N,4,D,H,W] fp32 map plus softmax / boundary scratch (~1 GB of HBM traffic per output
// at 2x128^3).  Here the loss reads the LOW-RES 4-channel logits (float4 per voxel, L2 resident) and interpolates in
// registers:
//   pass 1 (dsloss_fwd):  a CTA owns an 8x8x32 full-resolution tile, computes softmax(interp(lo)) once per tile point
//                         (+1 halo) into shared memory, and accumulates every reduction of App. A7 (I, P, T, Σce, Σfocal,
//                         ΣE²) — no full-resolution tensor is written;
//   pass 2 (dsloss_bwd):  recomputes the softmax tile (±1 halo) and the boundary residual E in shared memory, forms
//                         d(loss)/d(up-sampled logits) per voxel and applies the ADJOINT of the trilinear interpolation
//                         inside the tile (separable, shared memory), so only d(loss)/d(low-res logits) leaves the CTA
//                         (fp64 atomics: fixed-order partials, arrival order moves the sum by ~1e-16).
// Scale 1 (the level-0 head) skips the interpolation and stores the gradient directly.  Targets are uint8 (one conversion
// per loss call instead of four int64 reads per output).  Everything is HBM / L2-bound CUDA-core work.
#include "loss_common.cuh"
#include "b3d_internal.h"
#include <algorithm>
#include <mutex>

#define TZ 8
#define TY 8
#define TX 32
#define NTHREADS 256

__device__ __forceinline__ float4 ld_lo(const float4* p) { return __ldg(p); }

// up-sampled logits at full-resolution voxel (z, y, x): the arithmetic of trilinear_up_fwd_kernel (heads.cu), term for term
template <int S>
__device__ __forceinline__ float4 logit_at(const float4* __restrict__ lo, int Dl, int Hl, int Wl, int z, int y, int x) {
  if (S == 1) return ld_lo(lo + ((long long)z * Hl + y) * Wl + x);
  const float sc = 1.f / (float)S;
  int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
  lerp_src(z, sc, Dl, z0, z1, lz); lerp_src(y, sc, Hl, y0, y1, ly); lerp_src(x, sc, Wl, x0, x1, lx);
#define AT(zz, yy, xx) ld_lo(lo + ((long long)(zz) * Hl + (yy)) * Wl + (xx))
  const float4 v000 = AT(z0, y0, x0), v001 = AT(z0, y0, x1), v010 = AT(z0, y1, x0), v011 = AT(z0, y1, x1);
  const float4 v100 = AT(z1, y0, x0), v101 = AT(z1, y0, x1), v110 = AT(z1, y1, x0), v111 = AT(z1, y1, x1);
#undef AT
  const float wz0 = 1.f - lz, wy0 = 1.f - ly, wx0 = 1.f - lx;
#define MIX(f)                                                                                         \
  (wz0 * (wy0 * (wx0 * v000.f + lx * v001.f) + ly * (wx0 * v010.f + lx * v011.f)) +                    \
   lz * (wy0 * (wx0 * v100.f + lx * v101.f) + ly * (wx0 * v110.f + lx * v111.f)))
  return make_float4(MIX(x), MIX(y), MIX(z), MIX(w));
#undef MIX
}

// Low-res sub-tile staging (S > 1): the 8 taps of every interpolated point come from shared memory instead of 8 dependent
// L2 round trips per point.  A tile of T full-res positions plus halo touches at most T/S + 3 low-res cells per axis.
template <int S> struct LoTile {
  static constexpr int CZ = TZ / S + 3, CY = TY / S + 3, CX = TX / S + 3, N = (S > 1) ? CZ * CY * CX : 1;
};

template <int S>
__device__ __forceinline__ void stage_lo(float4* __restrict__ slo, const float4* __restrict__ lon, int Dl, int Hl, int Wl, int zmin,
                                         int ymin, int xmin, int& bz, int& by, int& bx) {
  using LT = LoTile<S>;
  const float sc = 1.f / (float)S;
  int i1; float l1;
  lerp_src(zmin, sc, Dl, bz, i1, l1); lerp_src(ymin, sc, Hl, by, i1, l1); lerp_src(xmin, sc, Wl, bx, i1, l1);
  for (int i = threadIdx.x; i < LT::N; i += NTHREADS) {
    const int cx = i % LT::CX, cy = (i / LT::CX) % LT::CY, cz = i / (LT::CX * LT::CY);
    const int gz = min(bz + cz, Dl - 1), gy = min(by + cy, Hl - 1), gx = min(bx + cx, Wl - 1);
    slo[i] = ld_lo(lon + ((long long)gz * Hl + gy) * Wl + gx);
  }
}

// logit_at<S> served from the staged sub-tile (same arithmetic, term for term)
template <int S>
__device__ __forceinline__ float4 logit_tile(const float4* __restrict__ slo, int bz, int by, int bx, int Dl, int Hl, int Wl, int z,
                                             int y, int x) {
  using LT = LoTile<S>;
  const float sc = 1.f / (float)S;
  int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
  lerp_src(z, sc, Dl, z0, z1, lz); lerp_src(y, sc, Hl, y0, y1, ly); lerp_src(x, sc, Wl, x0, x1, lx);
  z0 -= bz; z1 -= bz; y0 -= by; y1 -= by; x0 -= bx; x1 -= bx;
#define AT(zz, yy, xx) slo[((zz) * LT::CY + (yy)) * LT::CX + (xx)]
  const float4 v000 = AT(z0, y0, x0), v001 = AT(z0, y0, x1), v010 = AT(z0, y1, x0), v011 = AT(z0, y1, x1);
  const float4 v100 = AT(z1, y0, x0), v101 = AT(z1, y0, x1), v110 = AT(z1, y1, x0), v111 = AT(z1, y1, x1);
#undef AT
  const float wz0 = 1.f - lz, wy0 = 1.f - ly, wx0 = 1.f - lx;
#define MIX(f)                                                                                         \
  (wz0 * (wy0 * (wx0 * v000.f + lx * v001.f) + ly * (wx0 * v010.f + lx * v011.f)) +                    \
   lz * (wy0 * (wx0 * v100.f + lx * v101.f) + ly * (wx0 * v110.f + lx * v111.f)))
  return make_float4(MIX(x), MIX(y), MIX(z), MIX(w));
#undef MIX
}

// softmax of 4 logits exactly as loss_softmax_kernel computes it; also returns ce = lse - z_t and the focal term for class t
__device__ __forceinline__ float4 softmax4(const float4 zz, int t, float& ce) {
  const float z[KC] = {zz.x, zz.y, zz.z, zz.w};
  const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
  float e[KC], s = 0.f;
#pragma unroll
  for (int c = 0; c < KC; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
  const float inv = 1.f / s;
  const float lse = m + logf(s);
  float zt = z[0];
#pragma unroll
  for (int c = 1; c < KC; ++c) zt = (t == c) ? z[c] : zt;
  ce = lse - zt;
  return make_float4(e[0] * inv, e[1] * inv, e[2] * inv, e[3] * inv);
}

__device__ __forceinline__ float comp(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

__global__ void __launch_bounds__(256) target_u8_kernel(const long long* __restrict__ t, unsigned char* __restrict__ o, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long v = __ldg(t + i);
    o[i] = (v < 0 || v >= KC) ? (unsigned char)255 : (unsigned char)v;
  }
}

// -------------------------------------------------------------------------------------------------------------------------
// pass 1: all reductions.  acc[n][0..3] I, [4..7] P, [8..11] T, [12] Σce, [13] Σfocal, [14] ΣE²
// -------------------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(NTHREADS) dsloss_fwd_kernel(const float4* __restrict__ lo, const unsigned char* __restrict__ tgt,
                                                              double* __restrict__ acc, int N, int D, int H, int W, LossCfg cfg) {
  constexpr int PZ = TZ + 1, PY = TY + 1, PX = TX + 1, NP = PZ * PY * PX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sp = reinterpret_cast<float4*>(smem_raw);          // [NP] softmax tile
  float4* slo = sp + NP;                                      // [LoTile<S>::N] staged low-res logits
  unsigned char* st = reinterpret_cast<unsigned char*>(slo + LoTile<S>::N);   // [NP] labels
  __shared__ double s_acc[15];
  const int Dl = D / S, Hl = H / S, Wl = W / S;
  const int tiles_x = W / TX, tiles_y = H / TY, tiles_z = D / TZ;
  const long long tiles_per_n = (long long)tiles_x * tiles_y * tiles_z;
  const long long ntiles = tiles_per_n * N;
  const long long V = (long long)D * H * W;
  float a[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) a[i] = 0.f;
  int cur_n = -1;
  const int lane = threadIdx.x & 31;

  auto flush = [&](int n) {
    if (threadIdx.x < 15) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 15; ++i) {
      const float s = warp_sum(a[i]);
      if (lane == 0) atomicAdd(&s_acc[i], (double)s);
      a[i] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x < 15) atomicAdd(&acc[(long long)n * ACC_STRIDE + threadIdx.x], s_acc[threadIdx.x]);
    __syncthreads();
  };

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = (int)(tile / tiles_per_n);
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = n;
    }
    long long r = tile - (long long)n * tiles_per_n;
    const int x0 = (int)(r % tiles_x) * TX; r /= tiles_x;
    const int y0 = (int)(r % tiles_y) * TY;
    const int z0 = (int)(r / tiles_y) * TZ;
    const float4* lon = lo + (long long)n * Dl * Hl * Wl;
    const unsigned char* tn = tgt + (long long)n * V;
    __syncthreads();   // the previous tile's stencil reads are done
    int bz = 0, by = 0, bx = 0;
    if (S > 1) { stage_lo<S>(slo, lon, Dl, Hl, Wl, z0, y0, x0, bz, by, bx); __syncthreads(); }
#pragma unroll 2
    for (int i = threadIdx.x; i < NP; i += NTHREADS) {
      const int dx = i % PX, dy = (i / PX) % PY, dz = i / (PX * PY);
      const int gz = z0 + dz, gy = y0 + dy, gx = x0 + dx;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      int t = 255;
      if (gz < D && gy < H && gx < W) {
        t = tn[((long long)gz * H + gy) * W + gx];
        float ce;
        p = softmax4(S > 1 ? logit_tile<S>(slo, bz, by, bx, Dl, Hl, Wl, gz, gy, gx) : logit_at<S>(lon, Dl, Hl, Wl, gz, gy, gx), t, ce);
        if (dz < TZ && dy < TY && dx < TX) {   // interior point: the per-voxel sums
          a[4] += p.x; a[5] += p.y; a[6] += p.z; a[7] += p.w;
          if (t < KC) {
            const float pt = comp(p, t);
            a[0] += (t == 0) ? pt : 0.f; a[1] += (t == 1) ? pt : 0.f; a[2] += (t == 2) ? pt : 0.f; a[3] += (t == 3) ? pt : 0.f;
            a[8] += (t == 0) ? 1.f : 0.f; a[9] += (t == 1) ? 1.f : 0.f; a[10] += (t == 2) ? 1.f : 0.f; a[11] += (t == 3) ? 1.f : 0.f;
            a[12] += ce;
            const float ptx = expf(-ce);
            a[13] += cfg.f_alpha * focal_pow(1.f - ptx, cfg.f_gamma) * ce;
          }
        }
      }
      sp[i] = p;
      st[i] = (unsigned char)t;
    }
    __syncthreads();
    if (cfg.w_boundary != 0.f) {
#pragma unroll
      for (int k = 0; k < TZ; ++k) {
        const int dx = lane, dy = threadIdx.x >> 5, dz = k;
        const int i = (dz * PY + dy) * PX + dx;
        const bool hz = z0 + dz + 1 < D, hy = y0 + dy + 1 < H, hx = x0 + dx + 1 < W;
        const float4 p = sp[i], pz = sp[i + PY * PX], py = sp[i + PX], px = sp[i + 1];
        const int t = st[i], tz = st[i + PY * PX], ty = st[i + PX], tx = st[i + 1];
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          const float pc = comp(p, c);
          float bp = 0.f, bo = 0.f;
          if (hz) { bp += fabsf(comp(pz, c) - pc); bo += ((tz == c) != (t == c)) ? 1.f : 0.f; }
          if (hy) { bp += fabsf(comp(py, c) - pc); bo += ((ty == c) != (t == c)) ? 1.f : 0.f; }
          if (hx) { bp += fabsf(comp(px, c) - pc); bo += ((tx == c) != (t == c)) ? 1.f : 0.f; }
          const float e = bp - bo;
          a[14] = fmaf(e, e, a[14]);
        }
      }
    }
  }
  if (cur_n >= 0) flush(cur_n);
}

// -------------------------------------------------------------------------------------------------------------------------
// pass 2: d(total)/d(low-res logits).  S == 1: dlo is float4 [N][V] (plain stores); S > 1: dlo is double [N][Vl][4] (atomics,
// zeroed by the caller).
// -------------------------------------------------------------------------------------------------------------------------
template <int S>
struct BwdSmem {
  static constexpr int PZ = TZ + 2, PY = TY + 2, PX = TX + 2, NP = PZ * PY * PX;        // softmax tile, halo -1 .. +1
  static constexpr int EZ = TZ + 1, EY = TY + 1, EX = TX + 1, NE = EZ * EY * EX;        // boundary residual, halo -1 .. 0
  static constexpr int LZ = TZ / S + 2, LY = TY / S + 2, LX = TX / S + 2;               // low-res cells a tile touches
  static constexpr size_t bytes = (size_t)(NP + NE) * sizeof(float4) + NP;
};

template <int S>
__global__ void __launch_bounds__(NTHREADS) dsloss_bwd_kernel(const float4* __restrict__ lo, const unsigned char* __restrict__ tgt,
                                                              const double* __restrict__ acc, const float* __restrict__ gscale,
                                                              float wscale, void* __restrict__ dlo_out, int N, int D, int H, int W,
                                                              LossCfg cfg) {
  using SM = BwdSmem<S>;
  constexpr int PY = SM::PY, PX = SM::PX, NP = SM::NP, EY = SM::EY, EX = SM::EX, NE = SM::NE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sp = reinterpret_cast<float4*>(smem_raw);
  float4* sE = sp + NP;
  unsigned char* st = reinterpret_cast<unsigned char*>(sE + NE);
  __shared__ float s_d1[KC], s_d2[KC], s_t1[KC], s_t2[KC], s_t3[KC];
  const int Dl = D / S, Hl = H / S, Wl = W / S;
  const int tiles_x = W / TX, tiles_y = H / TY, tiles_z = D / TZ;
  const long long tiles_per_n = (long long)tiles_x * tiles_y * tiles_z;
  const long long ntiles = tiles_per_n * N;
  const long long V = (long long)D * H * W;
  const float g = gscale ? gscale[0] * wscale : wscale;
  const float inv_nv = 1.f / ((float)N * (float)V);
  const float wb = cfg.w_boundary * 2.f * inv_nv / KC;
  const bool use_b = cfg.w_boundary != 0.f;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  int cur_n = -1;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = (int)(tile / tiles_per_n);
    __syncthreads();   // everything of the previous tile (adjoint buffers alias sp / sE) is done
    if (n != cur_n) {
      cur_n = n;
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
    }
    long long r = tile - (long long)n * tiles_per_n;
    const int x0 = (int)(r % tiles_x) * TX; r /= tiles_x;
    const int y0 = (int)(r % tiles_y) * TY;
    const int z0 = (int)(r / tiles_y) * TZ;
    const float4* lon = lo + (long long)n * Dl * Hl * Wl;
    const unsigned char* tn = tgt + (long long)n * V;
    // ---- softmax tile with halo -1 .. +1 (index 0 <-> global coordinate origin - 1); the staged low-res tile aliases sE
    int bz = 0, by = 0, bx = 0;
    float4* slo = sE;
    static_assert(LoTile<S>::N <= NE, "staged low-res tile must fit the E tile");
    if (S > 1) { stage_lo<S>(slo, lon, Dl, Hl, Wl, max(z0 - 1, 0), max(y0 - 1, 0), max(x0 - 1, 0), bz, by, bx); __syncthreads(); }
#pragma unroll 2
    for (int i = threadIdx.x; i < NP; i += NTHREADS) {
      const int dx = i % PX, dy = (i / PX) % PY, dz = i / (PX * PY);
      const int gz = z0 + dz - 1, gy = y0 + dy - 1, gx = x0 + dx - 1;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      int t = 255;
      if (gz >= 0 && gy >= 0 && gx >= 0 && gz < D && gy < H && gx < W) {
        t = tn[((long long)gz * H + gy) * W + gx];
        float ce;
        p = softmax4(S > 1 ? logit_tile<S>(slo, bz, by, bx, Dl, Hl, Wl, gz, gy, gx) : logit_at<S>(lon, Dl, Hl, Wl, gz, gy, gx), t, ce);
      }
      sp[i] = p;
      st[i] = (unsigned char)t;
    }
    __syncthreads();
    // ---- boundary residual E = B(p) - B(onehot) at the tile voxels and their -1 neighbours (index 0 <-> origin - 1)
    if (use_b) {
      for (int i = threadIdx.x; i < NE; i += NTHREADS) {
        const int dx = i % EX, dy = (i / EX) % EY, dz = i / (EX * EY);
        const int gz = z0 + dz - 1, gy = y0 + dy - 1, gx = x0 + dx - 1;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gz >= 0 && gy >= 0 && gx >= 0) {
          const int j = (dz * PY + dy) * PX + dx;
          const bool hz = gz + 1 < D, hy = gy + 1 < H, hx = gx + 1 < W;
          const float4 p = sp[j], pz = sp[j + PY * PX], py = sp[j + PX], px = sp[j + 1];
          const int t = st[j], tz = st[j + PY * PX], ty = st[j + PX], tx = st[j + 1];
          float ev[KC];
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const float pc = comp(p, c);
            float bp = 0.f, bo = 0.f;
            if (hz) { bp += fabsf(comp(pz, c) - pc); bo += ((tz == c) != (t == c)) ? 1.f : 0.f; }
            if (hy) { bp += fabsf(comp(py, c) - pc); bo += ((ty == c) != (t == c)) ? 1.f : 0.f; }
            if (hx) { bp += fabsf(comp(px, c) - pc); bo += ((tx == c) != (t == c)) ? 1.f : 0.f; }
            ev[c] = bp - bo;
          }
          e = make_float4(ev[0], ev[1], ev[2], ev[3]);
        }
        sE[i] = e;
      }
    }
    __syncthreads();
    // ---- per-voxel gradient w.r.t. the up-sampled logits (the arithmetic of loss_bwd_kernel)
    float4 dzv[TZ];
#pragma unroll
    for (int k = 0; k < TZ; ++k) {
      const int dx = lane, dy = wrp, dz = k;
      const int gz = z0 + dz, gy = y0 + dy, gx = x0 + dx;
      const int j = ((dz + 1) * PY + (dy + 1)) * PX + (dx + 1);
      const int je = ((dz + 1) * EY + (dy + 1)) * EX + (dx + 1);
      const int t = st[j];
      const float4 p4 = sp[j];
      const float p[KC] = {p4.x, p4.y, p4.z, p4.w};
      float G[KC];
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float o = (c == t) ? 1.f : 0.f;
        G[c] = cfg.w_dice * (o * s_d1[c] + s_d2[c]) + cfg.w_tv * (o * (s_t1[c] + s_t2[c]) + s_t3[c]);
      }
      if (use_b) {
        const int sj[3] = {PY * PX, PX, 1};
        const int se[3] = {EY * EX, EX, 1};
        const bool hp[3] = {gz + 1 < D, gy + 1 < H, gx + 1 < W};
        const bool hm[3] = {gz > 0, gy > 0, gx > 0};
        const float4 ev4 = sE[je];
        float gb[KC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          if (hp[ax]) {
            const float4 q = sp[j + sj[ax]];
#pragma unroll
            for (int c = 0; c < KC; ++c) gb[c] -= sgnf(comp(q, c) - p[c]) * comp(ev4, c);
          }
          if (hm[ax]) {
            const float4 q = sp[j - sj[ax]];
            const float4 em = sE[je - se[ax]];
#pragma unroll
            for (int c = 0; c < KC; ++c) gb[c] += sgnf(p[c] - comp(q, c)) * comp(em, c);
          }
        }
#pragma unroll
        for (int c = 0; c < KC; ++c) G[c] = fmaf(wb, gb[c], G[c]);
      }
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < KC; ++c) dot = fmaf(G[c], p[c], dot);
      float ptv = p[0];
#pragma unroll
      for (int c = 1; c < KC; ++c) ptv = (t == c || (c == KC - 1 && t >= KC)) ? p[c] : ptv;
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
      float o4[KC];
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float o = (c == t) ? 1.f : 0.f;
        o4[c] = g * (p[c] * (G[c] - dot) + lin * (p[c] - o));
      }
      dzv[k] = make_float4(o4[0], o4[1], o4[2], o4[3]);
      if (S == 1)
        reinterpret_cast<float4*>(dlo_out)[(long long)n * V + ((long long)gz * H + gy) * W + gx] = dzv[k];
    }
    if constexpr (S > 1) {
    // ---- adjoint of the trilinear interpolation inside the tile (separable: x, then y, then z), buffers alias sp / sE
    constexpr int LZ = SM::LZ, LY = SM::LY, LX = SM::LX;
    float4* sdz = sE;                       // [TZ][TY][TX]           (NE >= TZ*TY*TX)
    float4* b1 = sp;                        // [TZ][TY][LX]
    float4* b2 = sp + TZ * TY * LX;         // [TZ][LY][LX]
    float4* b3 = b2 + TZ * LY * LX;         // [LZ][LY][LX]
    static_assert(TZ * TY * LX + TZ * LY * LX + LZ * LY * LX <= NP, "adjoint buffers must fit the softmax tile");
    static_assert(TZ * TY * TX <= NE, "gradient tile must fit the E tile");
    __syncthreads();                        // every thread finished reading sp / sE
#pragma unroll
    for (int k = 0; k < TZ; ++k) sdz[(k * TY + wrp) * TX + lane] = dzv[k];
    __syncthreads();
    const float sc = 1.f / (float)S;
    const int lx0 = x0 / S - 1, ly0 = y0 / S - 1, lz0 = z0 / S - 1;   // low-res index of local cell 0 (may be -1: unused)
    for (int i = threadIdx.x; i < TZ * TY * LX; i += NTHREADS) {
      const int jl = i % LX, row = i / LX;
      const int cell = lx0 + jl;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cell >= 0 && cell < Wl) {
        const int o_lo = max(x0, S * cell - S / 2 - 1), o_hi = min(x0 + TX, S * cell + S + S / 2 + 1);
        for (int o = o_lo; o < o_hi; ++o) {
          int i0, i1; float l1;
          lerp_src(o, sc, Wl, i0, i1, l1);
          float wgt = 0.f;
          if (i0 == cell) wgt += 1.f - l1;
          if (i1 == cell) wgt += l1;
          if (wgt != 0.f) {
            const float4 v = sdz[row * TX + (o - x0)];
            s4.x = fmaf(wgt, v.x, s4.x); s4.y = fmaf(wgt, v.y, s4.y); s4.z = fmaf(wgt, v.z, s4.z); s4.w = fmaf(wgt, v.w, s4.w);
          }
        }
      }
      b1[i] = s4;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TZ * LY * LX; i += NTHREADS) {
      const int jl = i % LX, yl = (i / LX) % LY, z = i / (LX * LY);
      const int cell = ly0 + yl;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cell >= 0 && cell < Hl) {
        const int o_lo = max(y0, S * cell - S / 2 - 1), o_hi = min(y0 + TY, S * cell + S + S / 2 + 1);
        for (int o = o_lo; o < o_hi; ++o) {
          int i0, i1; float l1;
          lerp_src(o, sc, Hl, i0, i1, l1);
          float wgt = 0.f;
          if (i0 == cell) wgt += 1.f - l1;
          if (i1 == cell) wgt += l1;
          if (wgt != 0.f) {
            const float4 v = b1[(z * TY + (o - y0)) * LX + jl];
            s4.x = fmaf(wgt, v.x, s4.x); s4.y = fmaf(wgt, v.y, s4.y); s4.z = fmaf(wgt, v.z, s4.z); s4.w = fmaf(wgt, v.w, s4.w);
          }
        }
      }
      b2[i] = s4;
    }
    __syncthreads();
    double* dlo = reinterpret_cast<double*>(dlo_out) + (long long)n * Dl * Hl * Wl * KC;
    for (int i = threadIdx.x; i < LZ * LY * LX; i += NTHREADS) {
      const int jl = i % LX, yl = (i / LX) % LY, zl = i / (LX * LY);
      const int cz = lz0 + zl, cy = ly0 + yl, cx = lx0 + jl;
      if (cz < 0 || cz >= Dl || cy < 0 || cy >= Hl || cx < 0 || cx >= Wl) continue;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const int o_lo = max(z0, S * cz - S / 2 - 1), o_hi = min(z0 + TZ, S * cz + S + S / 2 + 1);
      for (int o = o_lo; o < o_hi; ++o) {
        int i0, i1; float l1;
        lerp_src(o, sc, Dl, i0, i1, l1);
        float wgt = 0.f;
        if (i0 == cz) wgt += 1.f - l1;
        if (i1 == cz) wgt += l1;
        if (wgt != 0.f) {
          const float4 v = b2[((o - z0) * LY + yl) * LX + jl];
          s4.x = fmaf(wgt, v.x, s4.x); s4.y = fmaf(wgt, v.y, s4.y); s4.z = fmaf(wgt, v.z, s4.z); s4.w = fmaf(wgt, v.w, s4.w);
        }
      }
      double* d = dlo + (((long long)cz * Hl + cy) * Wl + cx) * KC;
      if (s4.x != 0.f) atomicAdd(d + 0, (double)s4.x);
      if (s4.y != 0.f) atomicAdd(d + 1, (double)s4.y);
      if (s4.z != 0.f) atomicAdd(d + 2, (double)s4.z);
      if (s4.w != 0.f) atomicAdd(d + 3, (double)s4.w);
    }
    (void)b3;
    }
  }
}

template <int S>
static int launch_fwd(const float4* lo, const unsigned char* tgt, double* acc, int N, int D, int H, int W, const LossCfg& cfg,
                      cudaStream_t st) {
  constexpr int NP = (TZ + 1) * (TY + 1) * (TX + 1);
  constexpr size_t smem = (size_t)(NP + LoTile<S>::N) * sizeof(float4) + NP;
  static const cudaError_t attr = cudaFuncSetAttribute(dsloss_fwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr != cudaSuccess) { b3d_set_error("dsloss_fwd: cannot raise the shared-memory limit: %s", cudaGetErrorString(attr)); return B3D_ERR_CUDA; }
  const long long ntiles = (long long)N * (D / TZ) * (H / TY) * (W / TX);
  const int blocks = (int)std::min<long long>(ntiles, (long long)b3d_num_sms() * 3);
  dsloss_fwd_kernel<S><<<blocks, NTHREADS, smem, st>>>(lo, tgt, acc, N, D, H, W, cfg); ++g_b3d_launches;
  return B3D_OK;
}

template <int S>
static int launch_bwd(const float4* lo, const unsigned char* tgt, const double* acc, const float* gscale, float wscale, void* dlo,
                      int N, int D, int H, int W, const LossCfg& cfg, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(dsloss_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdSmem<S>::bytes);
  });
  if (attr_err != cudaSuccess) { b3d_set_error("dsloss_bwd: cannot raise the shared-memory limit: %s", cudaGetErrorString(attr_err)); return B3D_ERR_CUDA; }
  const long long ntiles = (long long)N * (D / TZ) * (H / TY) * (W / TX);
  const int blocks = (int)std::min<long long>(ntiles, (long long)b3d_num_sms() * 2);
  dsloss_bwd_kernel<S><<<blocks, NTHREADS, BwdSmem<S>::bytes, st>>>(lo, tgt, acc, gscale, wscale, dlo, N, D, H, W, cfg); ++g_b3d_launches;
  return B3D_OK;
}

extern "C" {

// int64 class indices -> uint8 (values outside [0,4) become 255 = "no class": they contribute to no sum, like a label the
// one-hot never matches).  losses.py:20 / training.py:103 supply int64.
int b3d_target_u8(const long long* target, unsigned char* out, long long count, void* stream) {
  const int blocks = (int)std::min<long long>((count + 255) / 256, (long long)b3d_num_sms() * 8);
  target_u8_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(target, out, count); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// lo: low-res logits float4 [N][D/S][H/S][W/S]; target_u8 [N][D][H][W]; acc double [N][16] (zeroed here); values float[6].
// Replaces F.interpolate (main.py:165-170) + CombinedLoss3D.forward (losses.py:63-75) for one deep-supervision output.
int b3d_dsloss_fwd(const float* lo, const unsigned char* target_u8, const float* cfg11, double* acc, float* values, int N,
                   int scale, int D, int H, int W, void* stream) {
  B3D_REQUIRE(D % 32 == 0 && H % 32 == 0 && W % 32 == 0, "dsloss: D,H,W must be multiples of 32 (got %d,%d,%d)", D, H, W);
  LossCfg cfg;
  memcpy(&cfg, cfg11, sizeof(cfg));
  cudaStream_t st = (cudaStream_t)stream;
  B3D_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * ACC_STRIDE * N, st));
  const float4* l4 = (const float4*)lo;
  int rc = B3D_OK;
  switch (scale) {
    case 1: rc = launch_fwd<1>(l4, target_u8, acc, N, D, H, W, cfg, st); break;
    case 2: rc = launch_fwd<2>(l4, target_u8, acc, N, D, H, W, cfg, st); break;
    case 4: rc = launch_fwd<4>(l4, target_u8, acc, N, D, H, W, cfg, st); break;
    case 8: rc = launch_fwd<8>(l4, target_u8, acc, N, D, H, W, cfg, st); break;
    default: b3d_set_error("dsloss: scale %d unsupported (1,2,4,8)", scale); return B3D_ERR_UNSUPPORTED;
  }
  if (rc != B3D_OK) return rc;
  b3d_launch_loss_finalize(acc, N, (long long)D * H * W, cfg, values, st);
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// dlo: scale 1 -> float [N][V][4] (every element written); scale > 1 -> double [N][Vl][4], zeroed HERE, accumulated with
// fp64 atomics.  gscale: optional device float (upstream gradient of the scalar), wscale: host weight.
int b3d_dsloss_bwd(const float* lo, const unsigned char* target_u8, const double* acc, const float* cfg11, const float* gscale,
                   float wscale, void* dlo, int N, int scale, int D, int H, int W, void* stream) {
  B3D_REQUIRE(D % 32 == 0 && H % 32 == 0 && W % 32 == 0, "dsloss: D,H,W must be multiples of 32 (got %d,%d,%d)", D, H, W);
  LossCfg cfg;
  memcpy(&cfg, cfg11, sizeof(cfg));
  cudaStream_t st = (cudaStream_t)stream;
  const float4* l4 = (const float4*)lo;
  int rc = B3D_OK;
  if (scale > 1)
    B3D_CHECK_CUDA(cudaMemsetAsync(dlo, 0, sizeof(double) * KC * (size_t)N * (D / scale) * (H / scale) * (W / scale), st));
  switch (scale) {
    case 1: rc = launch_bwd<1>(l4, target_u8, acc, gscale, wscale, dlo, N, D, H, W, cfg, st); break;
    case 2: rc = launch_bwd<2>(l4, target_u8, acc, gscale, wscale, dlo, N, D, H, W, cfg, st); break;
    case 4: rc = launch_bwd<4>(l4, target_u8, acc, gscale, wscale, dlo, N, D, H, W, cfg, st); break;
    case 8: rc = launch_bwd<8>(l4, target_u8, acc, gscale, wscale, dlo, N, D, H, W, cfg, st); break;
    default: b3d_set_error("dsloss: scale %d unsupported (1,2,4,8)", scale); return B3D_ERR_UNSUPPORTED;
  }
  if (rc != B3D_OK) return rc;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
