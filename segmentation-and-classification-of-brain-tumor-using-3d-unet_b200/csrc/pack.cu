// pack.cu — conv-weight packing (fp32 reference layout -> bf16 tcgen05 operand layouts) and the fused optimizer step.
//
// The conv kernels read a bf16 copy of every conv weight, packed [ntaps][rows][Kp] (K-major rows, one swizzled TMA box per K
// chunk), once per layout a layer needs (fprop + dgrad).  The fp32 nn.Parameter in the reference layout stays the source of
// truth (state_dict compatibility, /root/reference/training.py:398-404), so every optimizer step is followed by a re-pack.
//
//   mode 0 conv fprop : W[co][ci][t]          -> k = ci, row = co, tap = t
//   mode 1 conv dgrad : W[co][ci][t]          -> k = co, row = ci, tap = ntaps-1-t       (flipped + transposed)
//   (3x3x3: the packed tap index runs (kh,kw,kd) with kd fastest, so the three kd taps of one (kh,kw) are adjacent row
//    blocks — conv_zs.cu multiplies them in ONE N = 3*Cout MMA)
//   mode 2 convT fprop: Wt[ci][co][t8]        -> k = ci, row = t8*Cout+co, tap 0
//   mode 3 convT dgrad: Wt[ci][co][t8]        -> k = t8*Cout+co, row = ci, tap 0
//
// Fused optimizer step (replaces torch.optim.AdamW(lr, weight_decay=1e-4, betas=(0.9, 0.999)),
// /root/reference/training.py:186-191, stepped at :296-304, lr scheduled by CosineAnnealingWarmRestarts :194-196,252):
//   adamw_pack_multi_kernel : ONE launch over every packed conv weight — a CTA owns a 16 x 16 x taps tile, reads p, g, m, v
//                             once (128-bit loads), applies decoupled-weight-decay Adam in fp32, writes p, m, v back and emits
//                             the tile into BOTH bf16 packed layouts (32-byte runs) from shared memory: 32 B per parameter
//                             instead of 28 B (AdamW) + 8 B (re-pack) in two passes and ~50 launches;
//   adamw_flat_multi_kernel : ONE launch over every other parameter (norm scales, biases, fp32 1x1 heads).
// lr and the step count live in device memory (CUDA-graph replay + LR schedulers keep working).
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

// ---------------------------------------------------------------------------------------------------------------------
// generic (single layout) pack — used for one-off packs
// ---------------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int mode, int Cout, int Cin,
                                   int ntaps, int Kp, int rows) {
  const int ptaps = (mode >= 2) ? 1 : ntaps;
  const long long total = (long long)ptaps * rows * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    long long r = i / Kp;
    const int row = (int)(r % rows);
    int t = (int)(r / rows);
    if (ntaps == 27 && mode < 2) t = (t % 3) * 9 + t / 3;  // packed tap order (kh,kw,kd) -> reference order (kd,kh,kw)
    float val = 0.f;
    if (mode == 0) {
      if (k < Cin && row < Cout) val = w[((long long)row * Cin + k) * ntaps + t];
    } else if (mode == 1) {
      if (k < Cout && row < Cin) val = w[((long long)k * Cin + row) * ntaps + (ntaps - 1 - t)];
    } else if (mode == 2) {
      const int t8 = row / Cout, co = row - t8 * Cout;
      if (k < Cin && t8 < ntaps) val = w[((long long)k * Cout + co) * ntaps + t8];
    } else {
      const int t8 = k / Cout, co = k - t8 * Cout;
      if (t8 < ntaps && row < Cin) val = w[((long long)row * Cout + co) * ntaps + t8];
    }
    out[i] = __float2bfloat16(val);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// AdamW (decoupled weight decay), the arithmetic of torch's fused kernel in fp32
// ---------------------------------------------------------------------------------------------------------------------
// The scalar constants are formed in DOUBLE exactly as torch does on the host (1 - beta, lr * wd, bias corrections, lr / bc1)
// and rounded to fp32 once; the per-element arithmetic uses IEEE division / square root (the library is built with
// --use_fast_math, whose approximate forms would cost 2 ulp per step against torch.optim.AdamW).
struct AdamConsts {
  float lr_wd;        // lr * weight_decay
  float beta2, omb1, omb2;   // beta2, 1 - beta1, 1 - beta2
  float step_size;    // lr / (1 - beta1^t)
  float bc2_sqrt;     // sqrt(1 - beta2^t)
  float eps;
};

__device__ __forceinline__ AdamConsts adam_consts(const float* __restrict__ lr_p, const float* __restrict__ step_p, double beta1,
                                                  double beta2, double eps, double wd) {
  __shared__ AdamConsts s_c;
  if (threadIdx.x == 0) {
    const double lr = (double)*lr_p, t = (double)*step_p;
    AdamConsts c;
    c.lr_wd = (float)(lr * wd); c.beta2 = (float)beta2; c.omb1 = (float)(1.0 - beta1); c.omb2 = (float)(1.0 - beta2);
    c.eps = (float)eps;
    c.step_size = (float)(lr / (1.0 - pow(beta1, t)));
    c.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, t));
    s_c = c;
  }
  __syncthreads();
  return s_c;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamConsts& c) {
  p = __fsub_rn(p, __fmul_rn(c.lr_wd, p));                        // param.mul_(1 - lr * wd)  (same value to 1 ulp)
  m = __fmaf_rn(c.omb1, __fsub_rn(g, m), m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = __fmaf_rn(__fmul_rn(c.omb2, g), g, __fmul_rn(c.beta2, v));   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
  p = __fsub_rn(p, __fmul_rn(c.step_size, __fdiv_rn(m, denom)));   // param.addcdiv_(exp_avg, denom, value = -step_size)
}

// ---------------------------------------------------------------------------------------------------------------------
// Both packed copies of one weight from ONE coalesced read.  A CTA owns a 16 x 16 tile of the two leading source dimensions
// with all taps: it reads 16 runs of 16*T contiguous floats into shared memory (optionally stepping AdamW on the way) and
// writes 32-byte runs (16 bf16 along K) into each packed layout.  src[(a*B + b)*T + t]:
//   conv  (a = co, b = ci): fprop[tp][co][ci]                 dgrad[tp][ci][co] holding tap ntaps-1-t   (tp = packed tap order)
//   convT (a = ci, b = co): fprop[t8*Cout + co][ci]           dgrad[ci][t8*Cout + co]
// ---------------------------------------------------------------------------------------------------------------------
template <bool ADAM>
__device__ __forceinline__ void pack_tile(float* __restrict__ tile, float* __restrict__ w, const float* __restrict__ g,
                                          float* __restrict__ m, float* __restrict__ v, const AdamConsts& ac, bf16* __restrict__ out_f,
                                          bf16* __restrict__ out_d, bool convt, int A, int B, int T, int Kp_f, int rows_f, int Kp_d,
                                          int rows_d, int a0, int b0) {
  const int run = 16 * T, pitch = run + 1;
  const bool vec = (b0 + 16 <= B) && (((long long)B * T) % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0) &&
                   (!ADAM || (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0));
  if (vec) {   // whole, 16-byte aligned runs
    const int run4 = run / 4;
    for (int i = threadIdx.x; i < 16 * run4; i += blockDim.x) {
      const int al = i / run4, r4 = i - al * run4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a0 + al < A) {
        const long long off = ((long long)(a0 + al) * B + b0) * T + 4 * r4;
        x = *reinterpret_cast<const float4*>(w + off);
        if (ADAM) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(g + off));
          float4 mm = *reinterpret_cast<const float4*>(m + off), vv = *reinterpret_cast<const float4*>(v + off);
          adam1(x.x, gg.x, mm.x, vv.x, ac); adam1(x.y, gg.y, mm.y, vv.y, ac);
          adam1(x.z, gg.z, mm.z, vv.z, ac); adam1(x.w, gg.w, mm.w, vv.w, ac);
          *reinterpret_cast<float4*>(w + off) = x;
          *reinterpret_cast<float4*>(m + off) = mm;
          *reinterpret_cast<float4*>(v + off) = vv;
        }
      }
      float* d = tile + al * pitch + 4 * r4;
      d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
    }
  } else {
    for (int i = threadIdx.x; i < 16 * run; i += blockDim.x) {
      const int al = i / run, r = i - al * run;
      const int bl = r / T;
      float x = 0.f;
      if (a0 + al < A && b0 + bl < B) {
        const long long off = ((long long)(a0 + al) * B + b0) * T + r;
        x = w[off];
        if (ADAM) {
          float mm = m[off], vv = v[off];
          adam1(x, g[off], mm, vv, ac);
          w[off] = x; m[off] = mm; v[off] = vv;
        }
      }
      tile[al * pitch + r] = x;
    }
  }
  __syncthreads();
  // work item = (tap, line, half): 8 bf16 = 16 bytes; a line is 16 elements along the packed K index
  const int items = T * 16 * 2;
  for (int i = threadIdx.x; i < 2 * items; i += blockDim.x) {
    const bool second = i >= items;            // false: K runs along b (fprop of conv / dgrad of convT); true: K runs along a
    int j = second ? i - items : i;
    const int half = j & 1; j >>= 1;
    const int line = j & 15; const int tp = j >> 4;
    float o[8];
    if (!convt) {
      int t = tp;
      if (T == 27) t = (tp % 3) * 9 + tp / 3;   // packed tap order (kh,kw,kd) -> reference order (kd,kh,kw)
      if (!second) {   // fprop[tp][co = a0+line][ci = b0 + 8*half ..]
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tile[line * pitch + (8 * half + e) * T + t];
        const int co = a0 + line, ci = b0 + 8 * half;
        if (co < rows_f && ci < Kp_f) stg16(out_f + ((long long)tp * rows_f + co) * Kp_f + ci, pack8(o));
      } else {         // dgrad[tp][ci = b0+line][co = a0 + 8*half ..], tap flipped
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tile[(8 * half + e) * pitch + line * T + (T - 1 - t)];
        const int ci = b0 + line, co = a0 + 8 * half;
        if (ci < rows_d && co < Kp_d) stg16(out_d + ((long long)tp * rows_d + ci) * Kp_d + co, pack8(o));
      }
    } else {
      const int Cout = B;
      if (!second) {   // dgrad[ci = a0+line][k = t8*Cout + co .. 8 consecutive co]
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tile[line * pitch + (8 * half + e) * T + tp];
        const int ci = a0 + line, co = b0 + 8 * half;
        if (ci < rows_d && co < Cout) {
          if (co + 8 <= Cout) stg16(out_d + (long long)ci * Kp_d + (long long)tp * Cout + co, pack8(o));
          else for (int e = 0; e < 8 && co + e < Cout; ++e) out_d[(long long)ci * Kp_d + (long long)tp * Cout + co + e] = __float2bfloat16(o[e]);
        }
      } else {         // fprop[row = t8*Cout + co = b0+line][k = ci = a0 + 8*half ..]
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tile[(8 * half + e) * pitch + line * T + tp];
        const int co = b0 + line, ci = a0 + 8 * half;
        if (co < Cout && ci < Kp_f) stg16(out_f + ((long long)tp * Cout + co) * Kp_f + ci, pack8(o));
      }
    }
  }
}

template <bool CONVT>
__global__ void __launch_bounds__(256) pack_pair_kernel(const float* __restrict__ w, bf16* __restrict__ out_f,
                                                        bf16* __restrict__ out_d, int A, int B, int T, int Kp_f, int rows_f,
                                                        int Kp_d, int rows_d) {
  extern __shared__ float tile[];   // [16 a][16 b][T] (+1 pad per a-row to spread banks)
  AdamConsts ac = {};
  pack_tile<false>(tile, const_cast<float*>(w), nullptr, nullptr, nullptr, ac, out_f, out_d, CONVT, A, B, T, Kp_f, rows_f, Kp_d,
                   rows_d, blockIdx.y * 16, blockIdx.x * 16);
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-tensor fused AdamW
// table row (16 x int64) of a packed conv weight:
//   0 w  1 g  2 m  3 v  4 out_f  5 out_d  6 A  7 B  8 T  9 convT  10 Kp_f  11 rows_f  12 Kp_d  13 rows_d  14 first tile  15 tiles_b
// table row (8 x int64) of a flat tensor:  0 w  1 g  2 m  3 v  4 numel  5 first block  6,7 unused
// ---------------------------------------------------------------------------------------------------------------------
#define ADAM_FLAT_CHUNK 4096   // elements per CTA of the flat kernel
#define ADAM_MAXP 64           // packed tensors per launch  (table passed BY VALUE: 8 KB of kernel parameters)
#define ADAM_MAXF 256          // flat tensors per launch    (16 KB)
// The tables travel as kernel PARAMETERS (CUDA >= 12.1: up to 32 KB), not through device memory: no host->device copy per
// step, nothing to keep alive, and a captured CUDA graph stores them by value in the kernel node.
struct PackTab { long long r[ADAM_MAXP][16]; };
struct FlatTab { long long r[ADAM_MAXF][8]; };

__global__ void __launch_bounds__(256) adamw_pack_multi_kernel(const __grid_constant__ PackTab T_, int ntensors, long long tile_base,
                                                               const float* __restrict__ lr_p, const float* __restrict__ step_p,
                                                               double beta1, double beta2, double eps, double wd) {
  extern __shared__ float tile[];
  const long long bid = tile_base + blockIdx.x;
  int lo = 0, hi = ntensors - 1;            // last tensor whose first tile <= bid
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (T_.r[mid][14] <= bid) lo = mid; else hi = mid - 1;
  }
  const long long* r = T_.r[lo];
  const int local = (int)(bid - r[14]);
  const int tiles_b = (int)r[15];
  const int a0 = (local / tiles_b) * 16, b0 = (local % tiles_b) * 16;
  const AdamConsts ac = adam_consts(lr_p, step_p, beta1, beta2, eps, wd);
  pack_tile<true>(tile, reinterpret_cast<float*>(r[0]), reinterpret_cast<const float*>(r[1]), reinterpret_cast<float*>(r[2]),
                  reinterpret_cast<float*>(r[3]), ac, reinterpret_cast<bf16*>(r[4]), reinterpret_cast<bf16*>(r[5]), r[9] != 0, (int)r[6],
                  (int)r[7], (int)r[8], (int)r[10], (int)r[11], (int)r[12], (int)r[13], a0, b0);
}

__global__ void __launch_bounds__(256) adamw_flat_multi_kernel(const __grid_constant__ FlatTab T_, int ntensors, long long block_base,
                                                               const float* __restrict__ lr_p, const float* __restrict__ step_p,
                                                               double beta1, double beta2, double eps, double wd) {
  const long long bid = block_base + blockIdx.x;
  int lo = 0, hi = ntensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (T_.r[mid][5] <= bid) lo = mid; else hi = mid - 1;
  }
  const long long* r = T_.r[lo];
  float* w = reinterpret_cast<float*>(r[0]);
  const float* g = reinterpret_cast<const float*>(r[1]);
  float* m = reinterpret_cast<float*>(r[2]);
  float* v = reinterpret_cast<float*>(r[3]);
  const long long numel = r[4];
  const long long beg = (bid - r[5]) * ADAM_FLAT_CHUNK;
  const long long end = beg + ADAM_FLAT_CHUNK < numel ? beg + ADAM_FLAT_CHUNK : numel;
  const AdamConsts ac = adam_consts(lr_p, step_p, beta1, beta2, eps, wd);
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    float p = w[i], mm = m[i], vv = v[i];
    adam1(p, g[i], mm, vv, ac);
    w[i] = p; m[i] = mm; v[i] = vv;
  }
}

extern "C" {

int b3d_pack_weight(int mode, const float* w, int Cout, int Cin, int ntaps, void* out, int Kp, int rows,
                    void* stream) {
  B3D_REQUIRE(mode >= 0 && mode <= 3, "pack_weight: bad mode %d", mode);
  B3D_REQUIRE(Kp % 16 == 0 && rows > 0, "pack_weight: bad Kp/rows");
  const int ptaps = (mode >= 2) ? 1 : ntaps;
  const long long total = (long long)ptaps * rows * Kp;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (bf16*)out, mode, Cout, Cin, ntaps, Kp, rows); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// Both packed copies of a weight in one launch (modes 0+1 for nn.Conv3d, 2+3 for nn.ConvTranspose3d k2 s2); the layouts
// are those of b3d_pack_weight.  convT: the dgrad buffer (K = 8*Cout, which need not be a multiple of 16 wide per tap) must be
// zero-initialised by the caller when 8*Cout < Kp_d (never the case for Cout % 2 == 0).
int b3d_pack_weight_pair(int convT, const float* w, int Cout, int Cin, int ntaps, void* out_fprop, void* out_dgrad,
                         void* stream) {
  B3D_REQUIRE(ntaps >= 1 && ntaps <= 27, "pack_weight_pair: bad ntaps %d", ntaps);
  const int r16i = (Cin + 15) / 16 * 16, r16o = (Cout + 15) / 16 * 16;
  const size_t smem = (size_t)16 * (16 * ntaps + 1) * sizeof(float);
  if (!convT) {
    dim3 grid(r16i / 16, r16o / 16);   // x: ci tiles (b), y: co tiles (a)
    pack_pair_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(w, (bf16*)out_fprop, (bf16*)out_dgrad, Cout, Cin, ntaps,
                                                                      r16i, r16o, r16o, r16i); ++g_b3d_launches;
  } else {
    B3D_REQUIRE(ntaps == 8 && Cout % 8 == 0, "pack_weight_pair: ConvTranspose3d needs 8 taps and Cout %% 8 == 0");
    dim3 grid((Cout + 15) / 16, r16i / 16);   // x: co tiles (b), y: ci tiles (a)
    pack_pair_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(w, (bf16*)out_fprop, (bf16*)out_dgrad, Cin, Cout, 8, r16i,
                                                                     8 * Cout, 8 * Cout, r16i); ++g_b3d_launches;
  }
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// One AdamW step (torch.optim.AdamW semantics: decoupled weight decay, bias correction with step count *step) over
//  * `n_pack` packed conv weights: HOST table int64 [n_pack][16] (layout above; column 14 = first tile, ascending from 0),
//    `total_tiles` CTAs — parameters, moments AND both bf16 packed copies are written in the same pass;
//  * `n_flat` other tensors: HOST table int64 [n_flat][8] (column 5 = first block), `total_blocks` CTAs of 4096 elements.
// The tables are copied into kernel parameters (chunks of 64 / 256 tensors per launch).  lr, step: device floats (step already
// incremented for this update).
int b3d_adamw_step(const long long* pack_table, int n_pack, long long total_tiles, const long long* flat_table, int n_flat,
                   long long total_blocks, const float* lr, const float* step, double beta1, double beta2, double eps,
                   double weight_decay, void* stream) {
  B3D_REQUIRE(total_tiles < (1ll << 31) && total_blocks < (1ll << 31), "adamw_step: too many tiles");
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n_pack; first += ADAM_MAXP) {
    const int cnt = std::min(ADAM_MAXP, n_pack - first);
    PackTab T;
    memset(&T, 0, sizeof(T));
    int max_taps = 1;
    for (int i = 0; i < cnt; ++i) {
      memcpy(T.r[i], pack_table + (long long)(first + i) * 16, 16 * sizeof(long long));
      B3D_REQUIRE(T.r[i][8] >= 1 && T.r[i][8] <= 27, "adamw_step: bad tap count %lld", T.r[i][8]);
      max_taps = std::max(max_taps, (int)T.r[i][8]);
    }
    const long long base = T.r[0][14];
    const long long end = (first + cnt < n_pack) ? pack_table[(long long)(first + cnt) * 16 + 14] : total_tiles;
    if (end <= base) continue;
    const size_t smem = (size_t)16 * (16 * max_taps + 1) * sizeof(float);
    adamw_pack_multi_kernel<<<(unsigned)(end - base), 256, smem, st>>>(T, cnt, base, lr, step, beta1, beta2, eps, weight_decay);
    ++g_b3d_launches;
  }
  for (int first = 0; first < n_flat; first += ADAM_MAXF) {
    const int cnt = std::min(ADAM_MAXF, n_flat - first);
    FlatTab T;
    memset(&T, 0, sizeof(T));
    for (int i = 0; i < cnt; ++i) memcpy(T.r[i], flat_table + (long long)(first + i) * 8, 8 * sizeof(long long));
    const long long base = T.r[0][5];
    const long long end = (first + cnt < n_flat) ? flat_table[(long long)(first + cnt) * 8 + 5] : total_blocks;
    if (end <= base) continue;
    adamw_flat_multi_kernel<<<(unsigned)(end - base), 256, 0, st>>>(T, cnt, base, lr, step, beta1, beta2, eps, weight_decay);
    ++g_b3d_launches;
  }
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
