// loss_common.cuh — pieces shared by loss.cu (losses on materialised fp32 logits) and dsloss.cu (fused deep-supervision loss).
#pragma once
#include "b3d_common.cuh"

#define KC 4
#define ACC_STRIDE 16  // per sample: I[4] P[4] T[4] Σce Σfocal ΣE² pad

struct LossCfg {
  float w_dice, smooth, w_focal, f_alpha, f_gamma, w_ce, w_boundary, w_tv, tv_alpha, tv_beta, tv_smooth;
};

__device__ __forceinline__ float focal_pow(float base, float gamma) {
  if (gamma == 2.f) return base * base;
  if (gamma == 1.f) return base;
  if (gamma == 0.f) return 1.f;
  return powf(fmaxf(base, 0.f), gamma);
}

__device__ __forceinline__ float sgnf(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// source index / weight of F.interpolate(mode="trilinear", align_corners=False) along one axis (main.py:165-170)
__device__ __forceinline__ void lerp_src(int o, float scale, int nin, int& i0, int& i1, float& l1) {
  float src = scale * (o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > nin - 1) i0 = nin - 1;
  i1 = i0 + ((i0 < nin - 1) ? 1 : 0);
  l1 = src - (float)i0;
}

// values[0..5] = total, dice, focal, boundary, ce, tversky from the per-sample accumulators (loss.cu)
int b3d_launch_loss_finalize(const double* acc, int N, long long V, const LossCfg& cfg, float* values, cudaStream_t st);
