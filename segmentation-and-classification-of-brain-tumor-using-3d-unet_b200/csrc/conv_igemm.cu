// conv_igemm.cu — implicit-GEMM 3-D convolution on tcgen05 tensor cores (sm_100a), NDHWC bf16.
//
// Replaces, for the hot path, what the reference dispatches to cuDNN / oneDNN for
//   nn.Conv3d(k=3,pad=1)            /root/reference/main.py:130,216,219   (fprop; dgrad = same kernel, flipped+transposed taps)
//   nn.Conv3d(k=1)                  /root/reference/main.py:229,252,258   (residual / gate projections)
//   nn.ConvTranspose3d(k=2,s=2)     /root/reference/main.py:121           (GEMM + pixel-shuffle epilogue; dgrad = 8 strided K-maps)
//
// Design (B200-first, not a cuDNN translation):
//   * voxels on UMMA-M (128 rows), output channels on UMMA-N, input channels x taps on K.
//   * one TMA box load stages a HALO TILE of the activation ([BD][BH][BW] voxels x 8-channel chunks) in shared memory in
//     the SWIZZLE_NONE "interleaved" K-major layout  [chunk][voxel][8 ch = 16 B].  In that layout a tap shift (kd,kh,kw)
//     is just +((kd*BH+kh)*BW+kw)*16 B on the UMMA descriptor start address, so all 27 taps re-use the same staged bytes
//     (no im2col, no per-tap reload from L2).  M-blocks are 128 consecutive positions of the halo-pitched linear index;
//     rows that land in the halo gap are computed and discarded (128/130 useful at W=128).
//   * weights are staged per K-chunk by ONE TMA box from the packed [K/8][tap][Cout][8] tensor.
//   * accumulators live in TMEM (MB x BN fp32 columns, double-buffered when they fit) ; a single elected thread issues
//     tcgen05.mma; 4 epilogue warps drain TMEM with tcgen05.ld, add bias, emit bf16 NDHWC (or fp32 split-K atomics, or the
//     ConvTranspose pixel-shuffle scatter) and the per-(sample,group) sum / sum-of-squares the following GroupNorm /
//     BatchNorm needs (so normalisation is a single read+write pass afterwards).
//   * persistent CTAs (one per SM), warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
//     warps 2-5 = epilogue.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>
#include <math.h>

struct alignas(64) IgemmParams {
  CUtensorMap tmA[8];
  CUtensorMap tmW;
  int N, D, H, W;        // output-space extent covered by tiles
  int Cout;              // real number of output columns (per tap for pixel shuffle: P.ps_cout)
  int halo, ks;          // halo = ks/2 ; ks = 3 or 1
  int TD, TH, TW, BD, BH, BW, box_vox;
  int RB, layout_type;   // smem row bytes (= KC*2: 32/64/128) and the matching UMMA swizzle layout type (6/4/2)
  int MB, BN, n_blocks, KC, k_chunks, chunks_per_map, stages, acc_bufs, acc_stride, tmem_cols;
  int tiles_x, tiles_y, tiles_z, num_tiles, num_items, ksplit, chunks_per_split;
  int row_mode, xblocks;   // row_mode: every M-block is one 128-wide run of a W row (no halo-gap rows)
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
  int mode;              // 0 bf16 store, 1 fp32 atomic accumulate, 2 pixel-shuffle bf16 store
  bf16* out; long long ld_out;
  int ps_cout;           // pixel shuffle: channels per tap
  const float* bias;
  float* out_f32; long long ld_f32;
  double* stats; int cpg; int stats_groups; int stats_batch;
  int* err;
};

#define IGEMM_THREADS 192

// first halo-pitched linear position covered by M-block mb
__device__ __forceinline__ int mblock_base(const IgemmParams& P, int mb) {
  if (!P.row_mode) return mb * 128;
  const int xb = mb % P.xblocks;
  const int t = mb / P.xblocks;
  const int py = t % P.TH, pz = t / P.TH;
  return (pz * P.BH + py) * P.BW + xb * 128;
}

__global__ void __launch_bounds__(IGEMM_THREADS, 1) igemm_kernel(const __grid_constant__ IgemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + S * P.a_stage_bytes;
  uint8_t* aux = smem + (size_t)S * (P.a_stage_bytes + P.b_stage_bytes);
  const uint32_t full0 = smem_u32(aux);
  const uint32_t empty0 = full0 + 8 * S;
  const uint32_t tfull0 = empty0 + 8 * S;
  const uint32_t tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * S + 32);
  float* s_stats = reinterpret_cast<float*>(aux + 16 * S + 48);  // [64]

  if (threadIdx.x == 0) {
    if (sA & 127u) { if (P.err) atomicExch(P.err, 9); __trap(); }  // TMA destinations need 128 B alignment
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_stats[threadIdx.x - 64] = 0.f;
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int items_per_tile = P.ksplit * P.n_blocks;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      for (int m = 0; m * P.chunks_per_map < P.k_chunks; ++m) tma_prefetch_desc(&P.tmA[m]);
      tma_prefetch_desc(&P.tmW);
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
        const int tile = item / items_per_tile;
        const int rem = item - tile * items_per_tile;
        const int split = rem / P.n_blocks, nblk = rem - split * P.n_blocks;
        int t = tile;
        const int tx = t % P.tiles_x; t /= P.tiles_x;
        const int ty = t % P.tiles_y; t /= P.tiles_y;
        const int tz = t % P.tiles_z; const int n = t / P.tiles_z;
        const int x0 = tx * P.TW - P.halo, y0 = ty * P.TH - P.halo, z0 = tz * P.TD - P.halo;
        const int kc0 = split * P.chunks_per_split;
        const int kc1 = min(P.k_chunks, kc0 + P.chunks_per_split);
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(empty0 + 8 * s, ph ^ 1, P.err, 1);
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, P.a_tx_bytes + P.b_tx_bytes);
          const int map = kc / P.chunks_per_map;
          const int cbase = (kc - map * P.chunks_per_map) * P.KC;
          tma_load_5d(sA + s * P.a_stage_bytes, &P.tmA[map], fb, cbase, x0, y0, z0, n);
          tma_load_3d(sB + s * P.b_stage_bytes, &P.tmW, fb, kc * P.KC, nblk * P.BN, 0);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, P.BN, 0, 0);
      // K-major swizzled operands: rows of RB bytes, 8-row swizzle atoms (SBO = 8*RB), LBO unused (=1)
      const uint64_t hi = umma_desc_hi_sw(8u * P.RB, P.layout_type);
      const uint32_t rb16 = (uint32_t)P.RB >> 4;  // row pitch in 16-byte units
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++it) {
        const int rem = item % items_per_tile;
        const int split = rem / P.n_blocks;
        const int kc0 = split * P.chunks_per_split;
        const int kc1 = min(P.k_chunks, kc0 + P.chunks_per_split);
        const int a = it % P.acc_bufs;
        const uint32_t aph = (uint32_t)(it / P.acc_bufs) & 1u;
        mbar_wait(tempty0 + 8 * a, aph ^ 1, P.err, 2);
        tc_fence_after();
        const uint32_t dbase = tmem_base + a * P.acc_stride;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(full0 + 8 * s, ph, P.err, 3);
          tc_fence_after();
          const uint32_t a16 = (sA + s * P.a_stage_bytes) >> 4;
          const uint32_t b16 = (sB + s * P.b_stage_bytes) >> 4;
          for (int mb = 0; mb < P.MB; ++mb) {
            const uint32_t d = dbase + mb * P.BN;
            const int mbase = mblock_base(P, mb);
            int tap = 0;
            for (int kd = 0; kd < P.ks; ++kd)
              for (int kh = 0; kh < P.ks; ++kh)
                for (int kw = 0; kw < P.ks; ++kw, ++tap) {
                  // a tap shift is a ROW shift of the start address; the swizzle XOR uses absolute smem address bits
                  // (scripts/umma_shift_test.cu), so any row offset is legal with base_offset = 0
                  const uint32_t aoff = a16 + (uint32_t)(mbase + (kd * P.BH + kh) * P.BW + kw) * rb16;
                  // packed tap order is (kh,kw,kd) for 3x3x3 (shared with conv_zs.cu, which stacks kd on N)
                  const int tapB = (P.ks == 3) ? (kh * 3 + kw) * 3 + kd : tap;
                  const uint32_t boff = b16 + (uint32_t)(tapB * P.BN) * rb16;
                  for (int k16 = 0; k16 < P.KC / 16; ++k16) {
                    const uint32_t alo = ((aoff + k16 * 2) & 0x3FFFu) | (1u << 16);
                    const uint32_t blo = ((boff + k16 * 2) & 0x3FFFu) | (1u << 16);
                    umma_bf16_ss(d, hi | alo, hi | blo, idesc, (kc > kc0 || tap > 0 || k16 > 0) ? 1u : 0u);
                  }
                }
          }
          umma_commit(empty0 + 8 * s);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
        umma_commit(tfull0 + 8 * a);
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127
    const int plane = P.BH * P.BW;
    const int seglen = P.cpg < 16 ? P.cpg : 16;
    int it = 0;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++it) {
      const int tile = item / items_per_tile;
      const int rem = item - tile * items_per_tile;
      const int nblk = rem % P.n_blocks;
      int t = tile;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y; t /= P.tiles_y;
      const int tz = t % P.tiles_z; const int n = t / P.tiles_z;
      const int a = it % P.acc_bufs;
      const uint32_t aph = (uint32_t)(it / P.acc_bufs) & 1u;
      mbar_wait(tfull0 + 8 * a, aph, P.err, 4);
      tc_fence_after();
      const int n0 = nblk * P.BN;
      for (int mb = 0; mb < P.MB; ++mb) {
        const int p = mblock_base(P, mb) + row;
        const int pz = p / plane, pr = p - pz * plane;
        const int py = pr / P.BW, px = pr - py * P.BW;
        const int z = tz * P.TD + pz, y = ty * P.TH + py, x = tx * P.TW + px;
        const bool valid = (pz < P.TD) && (py < P.TH) && (px < P.TW) && (z < P.D) && (y < P.H) && (x < P.W);
        long long vox;
        int ch_base = 0;
        if (P.mode == 2) {
          const int t8 = n0 / P.ps_cout;
          ch_base = n0 - t8 * P.ps_cout;
          const int oz = 2 * z + (t8 >> 2), oy = 2 * y + ((t8 >> 1) & 1), ox = 2 * x + (t8 & 1);
          vox = (((long long)n * (2 * P.D) + oz) * (2 * P.H) + oy) * (2 * P.W) + ox;
        } else {
          ch_base = n0;
          vox = (((long long)n * P.D + z) * P.H + y) * P.W + x;
        }
        const uint32_t trow = tmem_base + a * P.acc_stride + mb * P.BN + ((uint32_t)(q * 32) << 16);
        for (int j0 = 0; j0 < P.BN; j0 += 16) {
          uint32_t r[16];
          tmem_ld16(trow + j0, r);
          tmem_ld_wait();
          float v[16];
          const int c0 = ch_base + j0;
          const int cmax = (P.mode == 2 ? P.ps_cout : P.Cout);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float f = __uint_as_float(r[j]);
            if (P.bias != nullptr && c0 + j < cmax) f += __ldg(P.bias + c0 + j);
            v[j] = f;
          }
          if (P.stats != nullptr) {
            for (int sg = 0; sg < 16; sg += seglen) {
              float s1 = 0.f, s2 = 0.f;
              if (valid) {
                for (int j = sg; j < sg + seglen; ++j)
                  if (c0 + j < cmax) { s1 += v[j]; s2 += v[j] * v[j]; }
              }
              s1 = warp_sum(s1); s2 = warp_sum(s2);
              if (lane == 0) {
                const int g = (n0 + j0 + sg) / P.cpg - n0 / P.cpg;  // group slot local to this n-block
                atomicAdd(&s_stats[2 * g], s1);
                atomicAdd(&s_stats[2 * g + 1], s2);
              }
            }
          }
          if (valid) {
            if (P.mode == 1) {
              float* o = P.out_f32 + vox * P.ld_f32 + c0;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < cmax) atomicAdd(o + j, v[j]);
            } else {
              bf16* o = P.out + vox * P.ld_out + c0;
              if (c0 + 8 <= cmax) stg16(o, pack8(v));
              if (c0 + 16 <= cmax) stg16(o + 8, pack8(v + 8));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * a);
      if (P.stats != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int groups_blk = (n0 + P.BN - 1) / P.cpg - n0 / P.cpg + 1;
        if (et < 2 * groups_blk) {
          const int g = n0 / P.cpg + (et >> 1);
          if (g < P.stats_groups) {
            const float val = s_stats[et];
            const int ns = P.stats_batch ? 0 : n;
            atomicAdd(P.stats + ((long long)ns * P.stats_groups + g) * 2 + (et & 1), (double)val);
          }
          s_stats[et] = 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, P.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------
// split-K finalize: fp32 workspace [V][Cout] -> (+bias) -> bf16 out (pitch ld_out) + GroupNorm/BatchNorm partial sums
// ---------------------------------------------------------------------------------------------
__global__ void igemm_finalize_kernel(const float* __restrict__ ws, long long V, int Cout, long long vox_per_sample,
                                      const float* __restrict__ bias, bf16* __restrict__ out, long long ld_out,
                                      double* __restrict__ stats, int cpg, int stats_groups, int stats_batch) {
  __shared__ float s_acc[2 * 64];
  const int chunks = Cout / 8;
  const long long total = V * chunks;
  // each block handles a contiguous run of voxels of ONE sample (host guarantees blockDim*iters divides evenly enough)
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const long long per_block = (total + gridDim.x - 1) / gridDim.x;
  const long long beg = per_block * blockIdx.x;
  const long long end = min(total, beg + per_block);
  int cur_n = -1;
  for (long long base = beg; base < end; base += blockDim.x) {
    const long long idx = base + threadIdx.x;
    const bool act = idx < end;
    long long vox = 0; int c8 = 0;
    float v[8];
    if (act) {
      vox = idx / chunks; c8 = (int)(idx - vox * chunks);
      const float4 a = *reinterpret_cast<const float4*>(ws + vox * Cout + c8 * 8);
      const float4 b = *reinterpret_cast<const float4*>(ws + vox * Cout + c8 * 8 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      if (bias)
        for (int j = 0; j < 8; ++j) v[j] += bias[c8 * 8 + j];
      stg16(out + vox * ld_out + c8 * 8, pack8(v));
    }
    if (stats) {
      // sample index of the first element of this iteration; flush when the block crosses a sample boundary
      const int n_first = (int)((base / chunks) / vox_per_sample);
      const int n_last = (int)(((min(end, base + (long long)blockDim.x) - 1) / chunks) / vox_per_sample);
      if (cur_n < 0) cur_n = n_first;
      for (int nn = n_first; nn <= n_last; ++nn) {
        if (nn != cur_n) {
          __syncthreads();
          for (int i = threadIdx.x; i < 2 * (Cout / cpg) && i < 128; i += blockDim.x) {
            const int ns = stats_batch ? 0 : cur_n;
            atomicAdd(stats + ((long long)ns * stats_groups + (i >> 1)) * 2 + (i & 1), (double)s_acc[i]);
            s_acc[i] = 0.f;
          }
          __syncthreads();
          cur_n = nn;
        }
        if (act && (int)(vox / vox_per_sample) == nn) {
          if (cpg >= 8) {
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
            const int g = (c8 * 8) / cpg;
            atomicAdd(&s_acc[2 * g], s1); atomicAdd(&s_acc[2 * g + 1], s2);
          } else {
            for (int j0 = 0; j0 < 8; j0 += cpg) {
              float s1 = 0.f, s2 = 0.f;
              for (int j = j0; j < j0 + cpg; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
              const int g = (c8 * 8 + j0) / cpg;
              atomicAdd(&s_acc[2 * g], s1); atomicAdd(&s_acc[2 * g + 1], s2);
            }
          }
        }
      }
    }
  }
  if (stats && cur_n >= 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (Cout / cpg) && i < 128; i += blockDim.x) {
      const int ns = stats_batch ? 0 : cur_n;
      atomicAdd(stats + ((long long)ns * stats_groups + (i >> 1)) * 2 + (i & 1), (double)s_acc[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Weight packing: reference fp32 layouts -> bf16 [ntaps][rows][Kp]  (K-major rows; one swizzled TMA box per K chunk)
//   mode 0 conv fprop : W[co][ci][t]          -> k = ci, row = co, tap = t
//   mode 1 conv dgrad : W[co][ci][t]          -> k = co, row = ci, tap = ntaps-1-t       (flipped + transposed)
//   (3x3x3: the packed tap index runs (kh,kw,kd) with kd fastest, so the three kd taps of one (kh,kw) are adjacent row
//    blocks — conv_zs.cu multiplies them in ONE N = 3*Cout MMA)
//   mode 2 convT fprop: Wt[ci][co][t8]        -> k = ci, row = t8*Cout+co, tap 0
//   mode 3 convT dgrad: Wt[ci][co][t8]        -> k = t8*Cout+co, row = ci, tap 0
// ---------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------
// Host side: planner + launcher
// ---------------------------------------------------------------------------------------------
static const int kSmemBudget = 227 * 1024 - 2048;

struct IgemmPlan {
  int TD, TH, TW, KC, BN, MB, stages, acc_bufs, ksplit;
  double cost;
  bool ok;
};

static IgemmPlan plan_igemm(int N, int D, int H, int W, int chan_per_map, int nmaps, int CoutPad, int ks, int mode,
                            int num_sms, bool allow_split) {
  const int halo = ks / 2, ntaps = ks * ks * ks;
  const int Ktotal = chan_per_map * nmaps;
  IgemmPlan best; best.ok = false; best.cost = 1e30;
  const int bn_top = (mode == 2) ? std::min(CoutPad / 8, 256) : (CoutPad <= 256 ? CoutPad : 256);
  const int bn_cands[6] = {bn_top, 256, 128, 64, 32, 16};
  const int env_td = getenv("B3D_TD") ? atoi(getenv("B3D_TD")) : 0;
  const int env_th = getenv("B3D_TH") ? atoi(getenv("B3D_TH")) : 0;
  const int env_kc = getenv("B3D_KC") ? atoi(getenv("B3D_KC")) : 0;
  const int env_bn = getenv("B3D_BN") ? atoi(getenv("B3D_BN")) : 0;
  int TW = W;
  if (W + 2 * halo > 256) TW = 128;
  for (int TD = 1; TD <= 16 && TD <= D; TD *= 2)
    for (int TH = 1; TH <= 64 && TH <= H; TH *= 2)
      for (int KC = 64; KC >= 16; KC /= 2)
        for (int bi = 0; bi < 6; ++bi) {
          const int BN = bn_cands[bi];
          if (bi > 0 && (BN >= bn_cands[0])) continue;
          if (BN > 256 || BN % 16) continue;
          if (mode == 2 && (BN > CoutPad / 8 || (CoutPad / 8) % BN)) continue;  // pixel shuffle: n-block inside one tap
          if (chan_per_map % KC) continue;
          if (env_td && TD != env_td) continue;
          if (env_th && TH != env_th) continue;
          if (env_kc && KC != env_kc) continue;
          if (env_bn && BN != env_bn) continue;
          const int BD = TD + 2 * halo, BH = TH + 2 * halo, BW = TW + 2 * halo;
          if (BD > 256 || BH > 256 || BW > 256) continue;
          const long long box_vox = (long long)BD * BH * BW;
          const long long RB = KC * 2;
          const long long span = (long long)(TD - 1) * BH * BW + (long long)(TH - 1) * BW + TW;
          const bool row_mode = (TW % 128 == 0);
          const int MB = row_mode ? TD * TH * (TW / 128) : (int)((span + 127) / 128);
          if (MB * BN > 512) continue;
          const int acc_bufs = (2 * MB * BN <= 512) ? 2 : 1;
          const long long a = (box_vox * RB + 1023) / 1024 * 1024, b = ((long long)ntaps * BN * RB + 1023) / 1024 * 1024;
          const int k_chunks = Ktotal / KC;
          int stages = (int)std::min<long long>(4, kSmemBudget / (a + b));
          if (stages < 1) continue;
          if (stages < 2) continue;
          const long long maxoff = (row_mode ? span : (long long)MB * 128) + (long long)(ks - 1) * (BH * BW + BW + 1);
          const long long over = std::max<long long>(0, maxoff - box_vox) * RB;
          if (over > (long long)stages * b) continue;  // garbage rows must still read inside our smem
          const int n_blocks = (CoutPad + BN - 1) / BN;
          const long long tiles = (long long)N * ((D + TD - 1) / TD) * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
          long long items = tiles * n_blocks;
          int ksplit = 1;
          if (allow_split && items < num_sms && k_chunks > 1) {
            ksplit = (int)std::min<long long>(k_chunks, (num_sms + items - 1) / items);
            const int cps = (k_chunks + ksplit - 1) / ksplit;
            ksplit = (k_chunks + cps - 1) / cps;
          }
          const int cps = (k_chunks + ksplit - 1) / ksplit;
          const double clk_per_mma = std::max(BN / 2.0, 32.0 + BN / 4.0);
          const double mma_clk = (double)MB * ntaps * (KC / 16) * cps * clk_per_mma;
          const double load_clk = (double)(a + b) * cps / 40.0;
          const double epi_clk = (double)MB * (BN / 16) * 150.0;
          double item_clk = std::max(mma_clk, load_clk);
          item_clk = (acc_bufs == 2) ? std::max(item_clk, epi_clk) : item_clk + epi_clk;
          const long long its = items * ksplit;
          const long long waves = (its + num_sms - 1) / num_sms;
          double cost = (double)waves * item_clk + 3000.0;
          if (ksplit > 1) cost += 4000.0;
          if (stages < 3) cost *= 1.05;
          if (cost < best.cost) {
            best.ok = true; best.cost = cost; best.TD = TD; best.TH = TH; best.TW = TW; best.KC = KC; best.BN = BN;
            best.MB = MB; best.stages = stages; best.acc_bufs = acc_bufs; best.ksplit = ksplit;
          }
        }
  return best;
}

static bool g_smem_attr_set = false;

// Generic launcher.  `x_maps`: nmaps activation views (base pointer + dims + byte strides), each with `chan_per_map`
// channels (multiple of 16).  Output space extent is (N,D,H,W).
struct ActView {
  const void* base;
  long long sW, sH, sD, sN;  // byte strides of the W,H,D,N dims (channel stride = 2 bytes)
  int W, H, D, N;            // extents addressable through this view
  int C;                     // channels addressable (multiple of 8)
};

static int run_igemm(const ActView* views, int nmaps, int chan_per_map, const bf16* wpack, int w_rows, int ks,
                     int N, int D, int H, int W, int Cout, int mode, bf16* out, long long ld_out, int ps_cout,
                     const float* bias, double* stats, int cpg, int stats_groups, int stats_batch, float* ws,
                     size_t ws_bytes, int* err_flag, cudaStream_t stream) {
  const int ntaps = ks * ks * ks, halo = ks / 2;
  const int num_sms = b3d_num_sms();
  const int CoutPad = w_rows;  // rows in the packed weight tensor (multiple of 16)
  B3D_REQUIRE(chan_per_map % 16 == 0, "igemm: channels per K-map (%d) must be a multiple of 16", chan_per_map);
  B3D_REQUIRE(CoutPad % 16 == 0, "igemm: packed weight rows (%d) must be a multiple of 16", CoutPad);
  B3D_REQUIRE(nmaps >= 1 && nmaps <= 8, "igemm: nmaps out of range");
  if (stats) {
    B3D_REQUIRE(cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8 || cpg % 16 == 0,
                "igemm: channels-per-group %d unsupported (need 1,2,4,8 or a multiple of 16)", cpg);
  }
  const bool allow_split = (mode == 0) && ws != nullptr;
  IgemmPlan pl = plan_igemm(N, D, H, W, chan_per_map, nmaps, CoutPad, ks, mode, num_sms, allow_split);
  if (!pl.ok) {
    b3d_set_error("igemm: no tile plan for N=%d D=%d H=%d W=%d K=%dx%d Cout=%d ks=%d", N, D, H, W, nmaps, chan_per_map,
                  CoutPad, ks);
    return B3D_ERR_UNSUPPORTED;
  }
  if (stats) B3D_REQUIRE(pl.BN / cpg <= 32 && (pl.BN % cpg == 0 || cpg % pl.BN == 0), "igemm: BN %d / cpg %d unsupported", pl.BN, cpg);

  IgemmParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.D = D; P.H = H; P.W = W; P.Cout = Cout; P.halo = halo; P.ks = ks;
  P.TD = pl.TD; P.TH = pl.TH; P.TW = pl.TW;
  P.BD = pl.TD + 2 * halo; P.BH = pl.TH + 2 * halo; P.BW = pl.TW + 2 * halo;
  P.box_vox = P.BD * P.BH * P.BW;
  P.row_mode = (pl.TW % 128 == 0) ? 1 : 0; P.xblocks = pl.TW / 128;
  P.MB = pl.MB; P.BN = pl.BN; P.n_blocks = (CoutPad + pl.BN - 1) / pl.BN; P.KC = pl.KC;
  P.chunks_per_map = chan_per_map / pl.KC; P.k_chunks = P.chunks_per_map * nmaps;
  P.stages = pl.stages; P.acc_bufs = pl.acc_bufs; P.acc_stride = pl.MB * pl.BN;
  int cols = P.acc_bufs * P.acc_stride, tc = 32;
  while (tc < cols) tc *= 2;
  P.tmem_cols = tc;
  P.tiles_x = (W + P.TW - 1) / P.TW; P.tiles_y = (H + P.TH - 1) / P.TH; P.tiles_z = (D + P.TD - 1) / P.TD;
  P.num_tiles = N * P.tiles_x * P.tiles_y * P.tiles_z;
  P.ksplit = pl.ksplit; P.chunks_per_split = (P.k_chunks + pl.ksplit - 1) / pl.ksplit;
  P.num_items = P.num_tiles * P.n_blocks * P.ksplit;
  P.RB = P.KC * 2; P.layout_type = (P.RB == 128) ? 2 : (P.RB == 64 ? 4 : 6);
  P.a_tx_bytes = (uint32_t)P.box_vox * P.RB; P.b_tx_bytes = (uint32_t)ntaps * P.BN * P.RB;
  P.a_stage_bytes = (P.a_tx_bytes + 1023u) / 1024u * 1024u; P.b_stage_bytes = (P.b_tx_bytes + 1023u) / 1024u * 1024u;
  P.mode = mode; P.out = out; P.ld_out = ld_out; P.ps_cout = ps_cout; P.bias = bias;
  P.stats = stats; P.cpg = cpg > 0 ? cpg : 16; P.stats_groups = stats_groups; P.stats_batch = stats_batch;
  P.err = err_flag;

  const bool split = P.ksplit > 1;
  const long long V = (long long)N * D * H * W;
  if (split) {
    B3D_REQUIRE(ws_bytes >= (size_t)V * Cout * 4, "igemm: split-K workspace too small (%zu < %lld)", ws_bytes,
                V * Cout * 4);
    B3D_REQUIRE(Cout % 8 == 0, "igemm: split-K needs Cout %% 8 == 0");
    B3D_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)V * Cout * 4, stream));
    P.mode = 1; P.out_f32 = ws; P.ld_f32 = Cout; P.stats = nullptr; P.bias = nullptr;
  }

  for (int m = 0; m < nmaps; ++m) {
    const ActView& v = views[m];
    uint64_t dims[5] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.D, (uint64_t)v.N};
    uint64_t strides[4] = {(uint64_t)v.sW, (uint64_t)v.sH, (uint64_t)v.sD, (uint64_t)v.sN};
    uint32_t box[5] = {(uint32_t)P.KC, (uint32_t)P.BW, (uint32_t)P.BH, (uint32_t)P.BD, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmA[m], v.base, 5, dims, strides, box, P.RB);
    if (rc) return rc;
  }
  {
    const int Kp = chan_per_map * nmaps;
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)CoutPad, (uint64_t)ntaps};
    uint64_t strides[2] = {(uint64_t)Kp * 2, (uint64_t)CoutPad * Kp * 2};
    uint32_t box[3] = {(uint32_t)P.KC, (uint32_t)P.BN, (uint32_t)ntaps};
    int rc = b3d_encode_tmap_bf16(&P.tmW, wpack, 3, dims, strides, box, P.RB);
    if (rc) return rc;
  }
  const size_t smem = (size_t)P.stages * (P.a_stage_bytes + P.b_stage_bytes) + 16 * P.stages + 48 + 64 * 4 + 128 + 1024;
  if (!g_smem_attr_set) {
    B3D_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_smem_attr_set = true;
  }
  B3D_REQUIRE(smem <= 227 * 1024, "igemm: smem %zu too large", smem);
  const int grid = std::min(P.num_items, num_sms);
  if (getenv("B3D_VERBOSE"))
    fprintf(stderr,
            "[b3d] igemm N%d D%d H%d W%d K=%dx%d Cout=%d(ks%d mode%d) tile %dx%dx%d KC%d BN%d MB%d stages%d acc%d "
            "split%d items%d smem%zu tmem%d\n",
            N, D, H, W, nmaps, chan_per_map, Cout, ks, mode, P.TD, P.TH, P.TW, P.KC, P.BN, P.MB, P.stages, P.acc_bufs,
            P.ksplit, P.num_items, smem, P.tmem_cols);
  igemm_kernel<<<grid, IGEMM_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  if (split) {
    const int chunks = Cout / 8;
    const long long total = V * chunks;
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms * 8);
    // keep each block inside as few samples as possible
    igemm_finalize_kernel<<<blocks, 256, 0, stream>>>(ws, V, Cout, (long long)D * H * W, bias, out, ld_out, stats,
                                                      cpg > 0 ? cpg : 16, stats_groups, stats_batch); ++g_b3d_launches;
    B3D_CHECK_CUDA(cudaGetLastError());
  }
  return B3D_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
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

// Stride-1 "same" convolution, ks in {1,3}.  x: NDHWC bf16 with voxel pitch ldx (elements), Cin channels used (mult of 16).
// wpack: packed [Cin/8][ks^3][rows][8] (rows = roundup16(Cout)).  y: NDHWC bf16 pitch ldy.  stats: optional double
// [N or 1][groups][2] accumulated (+=) with sum / sum of squares of the (bias-added, fp32) outputs.
int b3d_conv_fprop(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
                   int N, int D, int H, int W, int Cin, int Cout, int ks, double* stats, int groups, int stats_batch,
                   void* ws, size_t ws_bytes, int* err_flag, void* stream) {
  B3D_REQUIRE(ks == 1 || ks == 3, "conv_fprop: ks must be 1 or 3");
  B3D_REQUIRE(Cin % 16 == 0, "conv_fprop: Cin (%d) must be a multiple of 16 (pad the activation)", Cin);
  B3D_REQUIRE(Cout % 8 == 0, "conv_fprop: Cout (%d) must be a multiple of 8", Cout);
  B3D_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0, "conv_fprop: pitches must be multiples of 8 elements");
  int n = N, d = D, h = H, w = W;
  if (ks == 1) {  // pointwise: flatten all voxels to a [rows][256] plane so tiles are dense
    long long V = (long long)N * D * H * W;
    int ww = 256;
    while (ww > 1 && V % ww) ww /= 2;
    long long hh = V / ww;
    int dd = 1;
    while (hh > 32768 && hh % 2 == 0) { hh /= 2; dd *= 2; }
    n = 1; d = dd; h = (int)hh; w = ww;
  }
  ActView v;
  v.base = x; v.C = Cin; v.W = w; v.H = h; v.D = d; v.N = n;
  v.sW = ldx * 2; v.sH = v.sW * w; v.sD = v.sH * h; v.sN = v.sD * d;
  const int cpg = (stats && groups > 0) ? Cout / groups : 0;
  if (ks == 3) {  // large planes, few output channels: z-marching kd-stacked kernel (conv_zs.cu)
    const int rc = b3d_try_zs(x, ldx, wpack, w_rows, bias, y, ldy, N, D, H, W, Cin, Cout, stats, cpg, groups, stats_batch,
                              err_flag, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  if (stats && ks == 1 && !stats_batch && N > 1) {
    // per-sample statistics need the sample index: keep N as the outer dim, flatten inside a sample
    long long V = (long long)D * H * W;
    int ww = 256;
    while (ww > 1 && V % ww) ww /= 2;
    long long hh = V / ww;
    int dd = 1;
    while (hh > 32768 && hh % 2 == 0) { hh /= 2; dd *= 2; }
    n = N; d = dd; h = (int)hh; w = ww;
    v.W = w; v.H = h; v.D = d; v.N = n;
    v.sW = ldx * 2; v.sH = v.sW * w; v.sD = v.sH * h; v.sN = v.sD * d;
  }
  return run_igemm(&v, 1, Cin, (const bf16*)wpack, w_rows, ks, n, d, h, w, Cout, 0, (bf16*)y, ldy, 0, bias, stats, cpg,
                   groups, stats_batch, (float*)ws, ws_bytes, err_flag, (cudaStream_t)stream);
}

// ConvTranspose3d(k=2,s=2) forward: x [N,D,H,W,Cin] -> y [N,2D,2H,2W,Cout] (pitch ldy), bias added.
// wpack: mode-2 packed [Cin/8][1][8*Cout][8].
int b3d_convT2_fprop(const void* x, long long ldx, const void* wpack, const float* bias, void* y, long long ldy, int N,
                     int D, int H, int W, int Cin, int Cout, int* err_flag, void* stream) {
  B3D_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "convT2_fprop: Cin/Cout must be multiples of 16");
  ActView v;
  v.base = x; v.C = Cin; v.W = W; v.H = H; v.D = D; v.N = N;
  v.sW = ldx * 2; v.sH = v.sW * W; v.sD = v.sH * H; v.sN = v.sD * D;
  return run_igemm(&v, 1, Cin, (const bf16*)wpack, 8 * Cout, 1, N, D, H, W, 8 * Cout, 2, (bf16*)y, ldy, Cout, bias,
                   nullptr, 0, 0, 0, nullptr, 0, err_flag, (cudaStream_t)stream);
}

// ConvTranspose3d(k=2,s=2) data gradient: dy [N,2D,2H,2W,Cout] (pitch lddy) -> dx [N,D,H,W,Cin] (pitch lddx).
// wpack: mode-3 packed [(8*Cout)/8][1][roundup16(Cin)][8].  K is gathered through 8 strided TMA views of dy.
int b3d_convT2_dgrad(const void* dy, long long lddy, const void* wpack, int w_rows, void* dx, long long lddx, int N, int D,
                     int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes, int* err_flag, void* stream) {
  B3D_REQUIRE(Cout % 16 == 0 && Cin % 8 == 0, "convT2_dgrad: Cout must be a multiple of 16, Cin of 8");
  ActView v[8];
  const long long pW = lddy * 2, pH = pW * (2 * W), pD = pH * (2 * H), pN = pD * (2 * D);
  for (int t8 = 0; t8 < 8; ++t8) {
    const int a = t8 >> 2, b = (t8 >> 1) & 1, c = t8 & 1;
    v[t8].base = (const char*)dy + a * pD + b * pH + c * pW;
    v[t8].C = Cout; v[t8].W = W; v[t8].H = H; v[t8].D = D; v[t8].N = N;
    v[t8].sW = 2 * pW; v[t8].sH = 2 * pH; v[t8].sD = 2 * pD; v[t8].sN = pN;
  }
  return run_igemm(v, 8, Cout, (const bf16*)wpack, w_rows, 1, N, D, H, W, Cin, 0, (bf16*)dx, lddx, 0, nullptr, nullptr, 0,
                   0, 0, (float*)ws, ws_bytes, err_flag, (cudaStream_t)stream);
}

}  // extern "C"
