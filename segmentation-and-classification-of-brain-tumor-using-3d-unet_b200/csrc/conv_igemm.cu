// conv_igemm.cu — implicit-GEMM 3-D convolution on tcgen05 tensor cores (sm_100a), NDHWC bf16.
//
// Replaces, for the hot path, what the reference dispatches to cuDNN / oneDNN for
//   nn.Conv3d(k=3,pad=1)            /root/reference/main.py:130,216,219   (fprop; dgrad = same kernel, flipped+transposed taps)
//   nn.Conv3d(k=1)                  /root/reference/main.py:229,252,258   (residual / gate projections)
//   nn.ConvTranspose3d(k=2,s=2)     /root/reference/main.py:121           (GEMM + pixel-shuffle epilogue; dgrad = 8 strided K-maps)
//
// Design (B200-first, not a cuDNN translation):
//   * voxels on UMMA-M (128 rows), output channels on UMMA-N, input channels x taps on K.
//   * one TMA box load stages a HALO TILE of the activation ([TD][TH+2][TW+2] voxels x KC channels) in shared memory as
//     K-major swizzled rows (one voxel = one row of KC*2 bytes).  A (kh,kw) tap is a ROW shift of the UMMA descriptor start
//     address inside that tile (the swizzle XOR uses absolute smem address bits — scripts/umma_shift_test.cu), so the 9
//     in-plane taps re-use the same staged bytes (no im2col); the kd taps are separate pipeline steps (box shifted in z).
//   * an M-block is 16 groups of 8 consecutive x positions.  "patch" tiles (TH = 16a, TW = 8b) put the groups one tile row
//     apart (descriptor SBO = row pitch) so every M row is a real voxel; "linear" tiles (small planes, 1x1x1 convs) use 128
//     consecutive halo-pitched positions and discard the rows that fall into the halo gap.
//   * weights: ONE TMA box per step from the packed [(kh,kw)][kd][Cout][K] tensor (the 9 (kh,kw) taps of this kd).
//   * accumulators live in TMEM (MB x BN fp32 columns, double-buffered when they fit); the MMA loop is templated on
//     (BN, KC) and unrolled so the single issuing thread spends a few uniform-datapath instructions per tcgen05.mma
//     (scripts/umma_rate.cu: a naive loop costs ~300 clk per MMA, the MMA itself 40-128); 4 (8 for BN <= 32) epilogue warps drain TMEM,
//     add bias, emit bf16 NDHWC with 32-byte stores (or fp32 split-K atomics, or the ConvTranspose pixel-shuffle scatter)
//     and accumulate the GroupNorm / BatchNorm sum / sum-of-squares in registers (flushed once per sample / channel block).
//   * persistent CTAs (one per SM), warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
//     warps 2-5 (2-9 for BN <= 32) = epilogue.
//   * 3x3x3 layers with <= 64 output channels take the z-marching kd-stacked kernel in conv_zs.cu instead.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>
#include <atomic>
#include <math.h>
#include <type_traits>

// producer warp, MMA warp, then 4 epilogue warps — or 8 (two per TMEM lane quadrant) for the narrow n-blocks of the streaming
// launches (BN <= 32: no spills under the 168-register cap of a 320-thread CTA; BN >= 64 would spill up to 596 bytes)
#define IG_EPI_WARPS(BN) ((BN) <= 32 ? 8 : 4)
#define IG_THREADS(BN) (64 + 32 * IG_EPI_WARPS(BN))
#define IG_MAXMB 32
#define IG_MAXSTAGES 8

struct alignas(64) IgParams {
  CUtensorMap tmA[8];
  CUtensorMap tmW;
  int N, D, H, W;        // output-space extent covered by tiles
  int Cout;              // real number of output columns (pixel shuffle: 8 * ps_cout)
  int halo, ks, KD;      // ks = 3 or 1 ; KD = ks (a pipeline step is one (K chunk, kd) pair)
  int TD, TH, TW, BH, BW;
  int MB, GS;            // M-blocks per tile ; M-block = 16 groups of 8 consecutive x positions, GS positions apart
  int mb_base[IG_MAXMB]; // first position (halo-pitched linear index inside the staged box) of each M-block
  int n_blocks, k_chunks, chunks_per_map, stages, acc_bufs, acc_stride, tmem_cols;
  int tiles_x, tiles_y, tiles_z, num_tiles, num_items, ksplit, steps_total, steps_per_split;
  uint32_t w_off, stage_bytes, a_tx, w_tx;
  int mode;              // 0 bf16 store, 1 fp32 partial store into this split's workspace slice (split-K), 2 pixel-shuffle bf16 store
  bf16* out; long long ld_out;
  int ps_cout;           // pixel shuffle: channels per tap
  const float* bias;
  const bf16* addend; long long ld_add;   // mode 0: out = conv + addend (gradient accumulation of two branches), may alias out
  float* out_f32; long long ld_f32; long long slice_f32;  // split-K partials: [ksplit][V][ld_f32]
  double* stats; int cpg; int stats_groups; int stats_batch;
  int* err;
};

__device__ __forceinline__ void stg32(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// BN: output columns per item (UMMA N) ; KC: channels per pipeline step (smem row = KC*2 bytes = swizzle span)
template <int BN, int KC>
__global__ void __launch_bounds__(IG_THREADS(BN), 1) igemm_kernel(const __grid_constant__ IgParams P) {
  constexpr int RB = KC * 2;
  constexpr int NK16 = KC / 16;
  constexpr uint32_t rb16 = RB / 16;
  constexpr int LT = (RB == 128) ? 2 : (RB == 64 ? 4 : 6);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t sA = smem_u32(smem);
  uint8_t* aux = smem + (size_t)S * P.stage_bytes;
  const uint32_t full0 = smem_u32(aux);
  const uint32_t empty0 = full0 + 8 * IG_MAXSTAGES;
  const uint32_t tfull0 = empty0 + 8 * IG_MAXSTAGES;
  const uint32_t tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * IG_MAXSTAGES + 32);
  double* s_stats = reinterpret_cast<double*>(aux + 16 * IG_MAXSTAGES + 48);  // [64], fp64: warp arrival order cannot change the sums

  if (threadIdx.x == 0) {
    if (sA & 1023u) { if (P.err) atomicExch(P.err, 9); __trap(); }
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, IG_EPI_WARPS(BN)); }
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_stats[threadIdx.x - 64] = 0.0;
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int items_per_tile = P.ksplit * P.n_blocks;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      for (int m = 0; m * P.chunks_per_map < P.k_chunks; ++m) tma_prefetch_desc(&P.tmA[m]);
      tma_prefetch_desc(&P.tmW);
    }
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
      const int st0 = split * P.steps_per_split;
      const int st1 = min(P.steps_total, st0 + P.steps_per_split);
      int kc = st0 / P.KD, kd = st0 - kc * P.KD;
      for (int st = st0; st < st1; ++st) {
        mbar_wait(empty0 + 8 * s, ph ^ 1, P.err, 1);
        if (elect_one()) {
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, P.a_tx + P.w_tx);
          const int map = kc / P.chunks_per_map;
          const int cbase = (kc - map * P.chunks_per_map) * KC;
          const uint32_t dst = sA + s * P.stage_bytes;
          tma_load_5d(dst, &P.tmA[map], fb, cbase, x0, y0, z0 + kd, n);
          tma_load_4d(dst + P.w_off, &P.tmW, fb, kc * KC, nblk * BN, kd, 0);
        }
        __syncwarp();
        if (++kd == P.KD) { kd = 0; ++kc; }
        if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform control flow, one elected lane issues) =======================
    const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    const uint64_t hiA = umma_desc_hi_sw((uint32_t)P.GS * RB, LT) | (1ull << 16);
    const uint64_t hiB = umma_desc_hi_sw(8u * RB, LT) | (1ull << 16);
    const uint32_t row_a = (uint32_t)P.BW * rb16;          // one tile row (kh step) in 16-byte units
    const uint32_t row_b = (uint32_t)(P.ks * BN) * rb16;   // one kh step of the weight stage
    uint32_t s = 0, ph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++it) {
      const int rem = item % items_per_tile;
      const int split = rem / P.n_blocks;
      const int st0 = split * P.steps_per_split;
      const int st1 = min(P.steps_total, st0 + P.steps_per_split);
      const int a = it % P.acc_bufs;
      const uint32_t aph = (uint32_t)(it / P.acc_bufs) & 1u;
      mbar_wait(tempty0 + 8 * a, aph ^ 1, P.err, 2);
      tc_fence_after();
      const uint32_t dbase = tmem_base + a * P.acc_stride;
      for (int st = st0; st < st1; ++st) {
        mbar_wait(full0 + 8 * s, ph, P.err, 3);
        tc_fence_after();
        const uint32_t a16 = (sA + s * P.stage_bytes) >> 4;
        const uint32_t b16 = a16 + (P.w_off >> 4);
        const uint32_t first = (st == st0) ? 0u : 1u;
        for (int mb = 0; mb < P.MB; ++mb) {
          const uint32_t d = dbase + mb * BN;
          uint32_t arow = a16 + (uint32_t)P.mb_base[mb] * rb16;
          uint32_t brow = b16;
          if (elect_one()) {
            for (int kh = 0; kh < P.ks; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                if (kw < P.ks) {
#pragma unroll
                  for (int k = 0; k < NK16; ++k) {
                    const uint32_t alo = arow + (uint32_t)(kw * rb16 + k * 2);
                    const uint32_t blo = brow + (uint32_t)(kw * BN * rb16 + k * 2);
                    umma_bf16_ss(d, hiA | alo, hiB | blo, idesc, (kh | kw | k) ? 1u : first);
                  }
                }
              }
              arow += row_a;
              brow += row_b;
            }
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(empty0 + 8 * s);
        __syncwarp();
        if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull0 + 8 * a);
      __syncwarp();
    }
  } else {
    // ======================= epilogue (warps 2..9 -> TMEM lane quadrants 2,3,0,1,2,3,0,1) =======================
    // Two warps per lane quadrant, taking alternate M-blocks of the item: the drain of one M-block is a serial
    // tcgen05.ld -> convert -> store chain, and four warps could not keep the streaming (1x1x1, ConvTranspose) launches fed
    // (they ran at ~50 % of the HBM floor with the planner already on the best plan: scripts/pw_sweep.py, round 2).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int MB_STEP = IG_EPI_WARPS(BN) / 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0 .. 32 * IG_EPI_WARPS - 1
    const int plane = P.BH * P.BW;
    const int cpg = P.cpg;
    // statistics granule (channels per per-thread accumulator): 16, 4 or 1 — the host guarantees BN / granule <= 16
    const int gran = cpg >= 16 ? 16 : (cpg >= 4 ? 4 : 1);
    float a1[16], a2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
    int cur_n = -1, cur_n0 = 0;
    auto flush_stats = [&]() {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i * gran < BN) {
          const float s1 = warp_sum(a1[i]), s2 = warp_sum(a2[i]);
          if (lane == 0) {
            const int gl = (cur_n0 + i * gran) / cpg - cur_n0 / cpg;  // group slot local to this n-block
            atomicAdd(&s_stats[2 * gl], (double)s1);
            atomicAdd(&s_stats[2 * gl + 1], (double)s2);
          }
        }
        a1[i] = 0.f; a2[i] = 0.f;
      }
      asm volatile("bar.sync 1, %0;" :: "n"(32 * IG_EPI_WARPS(BN)) : "memory");
      const int groups_blk = (cur_n0 + BN - 1) / cpg - cur_n0 / cpg + 1;
      if (et < 2 * groups_blk) {
        const int g = cur_n0 / cpg + (et >> 1);
        if (g < P.stats_groups) {
          const int ns = P.stats_batch ? 0 : cur_n;
          atomicAdd(P.stats + ((long long)ns * P.stats_groups + g) * 2 + (et & 1), s_stats[et]);
        }
        s_stats[et] = 0.0;
      }
      asm volatile("bar.sync 1, %0;" :: "n"(32 * IG_EPI_WARPS(BN)) : "memory");
    };
    int it = 0;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++it) {
      const int tile = item / items_per_tile;
      const int rem = item - tile * items_per_tile;
      const int nblk = rem % P.n_blocks;
      const int split = rem / P.n_blocks;
      int t = tile;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y; t /= P.tiles_y;
      const int tz = t % P.tiles_z; const int n = t / P.tiles_z;
      const int a = it % P.acc_bufs;
      const uint32_t aph = (uint32_t)(it / P.acc_bufs) & 1u;
      const int n0 = nblk * BN;
      if (P.stats != nullptr && (n != cur_n || n0 != cur_n0)) {
        if (cur_n >= 0) flush_stats();
        cur_n = n; cur_n0 = n0;
      }
      mbar_wait(tfull0 + 8 * a, aph, P.err, 4);
      tc_fence_after();
      const int cmax = (P.mode == 2 ? P.ps_cout : P.Cout);
      int t8 = 0, ch_base = n0;
      if (P.mode == 2) { t8 = n0 / P.ps_cout; ch_base = n0 - t8 * P.ps_cout; }
      // every column of this n-block is a real channel (pixel shuffle: n-blocks cover whole taps of ps_cout % 16 == 0 channels)
      const bool full = (P.mode == 2) ? (P.ps_cout % 16 == 0) : (ch_base + BN <= cmax);
      // specialised bodies: one (mode, statistics granule) combination runs, without per-element predicates
      auto body = [&](auto MODE_c, auto GRAN_c, auto FULL_c) {
        constexpr int MODE = decltype(MODE_c)::value, GRAN = decltype(GRAN_c)::value;
        constexpr bool FULL = decltype(FULL_c)::value;
        for (int mb = half; mb < P.MB; mb += MB_STEP) {
          const int p = P.mb_base[mb] + (row >> 3) * P.GS + (row & 7);
          const int pz = p / plane, pr = p - pz * plane;
          const int py = pr / P.BW, px = pr - py * P.BW;
          const int z = tz * P.TD + pz, y = ty * P.TH + py, x = tx * P.TW + px;
          const bool valid = (pz < P.TD) && (py < P.TH) && (px < P.TW) && (z < P.D) && (y < P.H) && (x < P.W);
          long long vox = (((long long)n * P.D + z) * P.H + y) * P.W + x;   // MODE 2: recomputed per 32-column chunk (tap)
          const uint32_t trow = tmem_base + a * P.acc_stride + mb * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int j0 = 0; j0 < BN; j0 += 32) {
            constexpr int NC = (BN >= 32) ? 32 : 16;   // columns per TMEM round trip
            uint32_t r[NC];
            tmem_ld16(trow + j0, r);
            if (NC == 32) tmem_ld16(trow + j0 + 16, r + 16);
            tmem_ld_wait();
            float v[NC];
            int c0 = ch_base + j0;
            if (MODE == 2) {   // pixel shuffle: this chunk of columns belongs to tap t8j of the 2x2x2 up-sampling
              const int t8j = (n0 + j0) / P.ps_cout;
              c0 = (n0 + j0) - t8j * P.ps_cout;
              const int oz = 2 * z + (t8j >> 2), oy = 2 * y + ((t8j >> 1) & 1), ox = 2 * x + (t8j & 1);
              vox = (((long long)n * (2 * P.D) + oz) * (2 * P.H) + oy) * (2 * P.W) + ox;
            }
#pragma unroll
            for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(r[j]);
            if (P.bias != nullptr) {
              if (FULL) {
#pragma unroll
                for (int j = 0; j < NC; j += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + c0 + j));
                  v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < NC; ++j) if (c0 + j < cmax) v[j] += __ldg(P.bias + c0 + j);
              }
            }
            if (MODE == 0 && P.addend != nullptr && valid) {   // fused branch-gradient add (replaces a separate add pass)
              const bf16* ap = P.addend + vox * P.ld_add + c0;
#pragma unroll
              for (int h = 0; h < NC / 8; ++h) {
                if (FULL || c0 + 8 * h + 8 <= cmax) {
                  float t8v[8];
                  unpack8(ldg16(ap + 8 * h), t8v);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * h + j] += t8v[j];
                }
              }
            }
            if (valid) {
              if (GRAN == 16) {
#pragma unroll
                for (int h = 0; h < NC / 16; ++h) {
                  float s1 = 0.f, s2 = 0.f;
#pragma unroll
                  for (int j = 0; j < 16; ++j) { s1 += v[16 * h + j]; s2 = fmaf(v[16 * h + j], v[16 * h + j], s2); }
                  a1[(j0 / 16 + h) & 15] += s1; a2[(j0 / 16 + h) & 15] += s2;
                }
              } else if (GRAN == 4) {
                if (j0 < 64) {
#pragma unroll
                  for (int qd = 0; qd < NC / 4; ++qd) {
                    const int ai = (j0 / 4 + qd) & 15;
                    a1[ai] += (v[4 * qd] + v[4 * qd + 1]) + (v[4 * qd + 2] + v[4 * qd + 3]);
                    a2[ai] += (v[4 * qd] * v[4 * qd] + v[4 * qd + 1] * v[4 * qd + 1]) +
                              (v[4 * qd + 2] * v[4 * qd + 2] + v[4 * qd + 3] * v[4 * qd + 3]);
                  }
                }
              } else if (GRAN == 1) {
                if (j0 == 0) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { a1[j] += v[j]; a2[j] = fmaf(v[j], v[j], a2[j]); }
                }
              }
              if (MODE == 1) {
                float* o = P.out_f32 + split * P.slice_f32 + vox * P.ld_f32 + c0;
                if (FULL) {
                  const uint4* pv = reinterpret_cast<const uint4*>(v);
#pragma unroll
                  for (int h = 0; h < NC / 8; ++h) stg32(o + 8 * h, pv[2 * h], pv[2 * h + 1]);
                } else {
#pragma unroll
                  for (int j = 0; j < NC; ++j) if (c0 + j < cmax) o[j] = v[j];
                }
              } else {
                bf16* o = P.out + vox * P.ld_out + c0;
                if (FULL && ((reinterpret_cast<uintptr_t>(o) & 31) == 0)) {
#pragma unroll
                  for (int h = 0; h < NC / 16; ++h) stg32(o + 16 * h, pack8(v + 16 * h), pack8(v + 16 * h + 8));
                } else {
#pragma unroll
                  for (int h = 0; h < NC / 8; ++h) if (c0 + 8 * h + 8 <= cmax) stg16(o + 8 * h, pack8(v + 8 * h));
                }
              }
            }
          }
        }
      };
      using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
      using I4 = std::integral_constant<int, 4>; using I16 = std::integral_constant<int, 16>;
      if (!full) {
        if (P.mode == 1) body(I1{}, I0{}, std::false_type{});
        else if (P.mode == 2) body(I2{}, I0{}, std::false_type{});
        else if (P.stats == nullptr) body(I0{}, I0{}, std::false_type{});
        else if (gran == 16) body(I0{}, I16{}, std::false_type{});
        else if (gran == 4) body(I0{}, I4{}, std::false_type{});
        else body(I0{}, I1{}, std::false_type{});
      } else if (P.mode == 1) body(I1{}, I0{}, std::true_type{});
      else if (P.mode == 2) body(I2{}, I0{}, std::true_type{});
      else if (P.stats == nullptr) body(I0{}, I0{}, std::true_type{});
      else if (gran == 16) body(I0{}, I16{}, std::true_type{});
      else if (gran == 4) body(I0{}, I4{}, std::true_type{});
      else body(I0{}, I1{}, std::true_type{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * a);
    }
    if (P.stats != nullptr && cur_n >= 0) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, P.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------
// split-K finalize: fp32 workspace [V][Cout] -> (+bias) -> bf16 out (pitch ld_out) + GroupNorm/BatchNorm partial sums
// ---------------------------------------------------------------------------------------------
__global__ void igemm_finalize_kernel(const float* __restrict__ ws, int nsplit, long long V, int Cout, long long vox_per_sample,
                                      const float* __restrict__ bias, bf16* __restrict__ out, long long ld_out,
                                      double* __restrict__ stats, int cpg, int stats_groups, int stats_batch) {
  __shared__ double s_acc[2 * 64];   // fp64: thread arrival order cannot change the sums
  const int chunks = Cout / 8;
  const long long total = V * chunks;
  // each block handles a contiguous run of voxels of ONE sample (host guarantees blockDim*iters divides evenly enough)
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_acc[i] = 0.0;
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
      float4 a = *reinterpret_cast<const float4*>(ws + vox * Cout + c8 * 8);
      float4 b = *reinterpret_cast<const float4*>(ws + vox * Cout + c8 * 8 + 4);
      for (int sp = 1; sp < nsplit; ++sp) {
        const float* w2 = ws + (long long)sp * V * Cout + vox * Cout + c8 * 8;
        const float4 a2 = *reinterpret_cast<const float4*>(w2);
        const float4 b2 = *reinterpret_cast<const float4*>(w2 + 4);
        a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w; b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
      }
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
            atomicAdd(stats + ((long long)ns * stats_groups + (i >> 1)) * 2 + (i & 1), s_acc[i]);
            s_acc[i] = 0.0;
          }
          __syncthreads();
          cur_n = nn;
        }
        if (act && (int)(vox / vox_per_sample) == nn) {
          if (cpg >= 8) {
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
            const int g = (c8 * 8) / cpg;
            atomicAdd(&s_acc[2 * g], (double)s1); atomicAdd(&s_acc[2 * g + 1], (double)s2);
          } else {
            for (int j0 = 0; j0 < 8; j0 += cpg) {
              float s1 = 0.f, s2 = 0.f;
              for (int j = j0; j < j0 + cpg; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
              const int g = (c8 * 8 + j0) / cpg;
              atomicAdd(&s_acc[2 * g], (double)s1); atomicAdd(&s_acc[2 * g + 1], (double)s2);
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
      atomicAdd(stats + ((long long)ns * stats_groups + (i >> 1)) * 2 + (i & 1), s_acc[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: planner + launcher
// ---------------------------------------------------------------------------------------------
static const int kSmemBudget = 227 * 1024 - 2048;
static std::atomic<int> g_plan_override[6];   // TD, TH, TW, KC, BN, verbose  (0 = planner's choice)
static thread_local char g_last_plan[256] = "";

struct IgemmPlan {
  int TD, TH, TW, KC, BN, MB, GS, stages, acc_bufs, ksplit, patch;
  long long over;  // bytes the garbage rows of the last M-block may read past the end of a stage
  double cost;
  bool ok;
};

// measured (scripts/umma_rate.cu): clocks per M128 x N x K16 MMA = max(N/2, (4096 + 32 N) / 128) — SMEM operand fetch bound
static inline double mma_clk(int n) { return std::max(n / 2.0, (4096.0 + 32.0 * n) / 128.0) + 4.0; }

static IgemmPlan plan_igemm(int N, int D, int H, int W, int chan_per_map, int nmaps, int CoutPad, int ks, int mode,
                            int num_sms, int max_split, int stat_gran) {
  const int halo = ks / 2, KD = ks, KHW = ks * ks;
  const int Ktotal = chan_per_map * nmaps;
  IgemmPlan best; best.ok = false; best.cost = 1e30;
  // tuning overrides: environment (read once) or b3d_set_plan_override (runtime, for in-process sweeps: scripts/plan_sweep.py)
  const int env_td = g_plan_override[0].load() ? g_plan_override[0].load() : B3D_ENV_INT("B3D_TD");
  const int env_th = g_plan_override[1].load() ? g_plan_override[1].load() : B3D_ENV_INT("B3D_TH");
  const int env_tw = g_plan_override[2].load() ? g_plan_override[2].load() : B3D_ENV_INT("B3D_TW");
  const int env_kc = g_plan_override[3].load() ? g_plan_override[3].load() : B3D_ENV_INT("B3D_KC");
  const int env_bn = g_plan_override[4].load() ? g_plan_override[4].load() : B3D_ENV_INT("B3D_BN");
  const int bn_cands[5] = {256, 128, 64, 32, 16};
  for (int patch = B3D_ENV_FLAG("B3D_NOPATCH") ? 0 : 1; patch >= 0; --patch) {
    // patch mode: tile = TD x (16 a) x (8 b), M-block = 16 rows x 8 columns (no wasted M rows)
    // linear mode: tile rows are TW wide (whole W when it fits), M-block = 128 consecutive halo-pitched positions
    for (int TD = 1; TD <= 16 && TD <= std::max(D, 1); TD *= 2)
      for (int TH = patch ? 16 : 1; TH <= 64; TH *= 2)
        for (int TWi = 0; TWi < 6; ++TWi) {
          int TW;
          if (patch) { TW = 8 << TWi; if (TW > 128) continue; if (H < 16 || W < 8) continue; if (TW / 2 >= W && TWi > 0) continue; }
          else { if (TWi > 0) continue; TW = (W + 2 * halo > 256) ? 128 : W; }
          if (TH / 2 >= H && TH > (patch ? 16 : 1)) continue;
          if (env_td && TD != env_td) continue;
          if (env_th && TH != env_th) continue;
          if (env_tw && TW != env_tw) continue;
          const int BH = TH + 2 * halo, BW = TW + 2 * halo;
          if (BH > 256 || BW > 256 || TD > 256) continue;
          const long long box_vox = (long long)TD * BH * BW;
          int MB;
          long long maxrow;  // one past the last box position any M-block row can touch (taps included)
          if (patch) {
            MB = TD * (TH / 16) * (TW / 8);
            maxrow = box_vox;
          } else {
            const long long span = (long long)(TD - 1) * BH * BW + (long long)(TH - 1) * BW + TW;
            MB = (int)((span + 127) / 128);
            maxrow = (long long)MB * 128 + (long long)(ks - 1) * (BW + 1);
          }
          if (MB > IG_MAXMB) continue;
          for (int KC = 64; KC >= 16; KC /= 2)
            for (int bi = 0; bi < 5; ++bi) {
              const int BN = bn_cands[bi];
              if (BN > CoutPad && BN != 16) continue;
              if (BN > CoutPad) continue;
              if (CoutPad % BN) continue;
              if (mode == 2) {  // pixel shuffle: an n-block lies inside one tap, or covers whole taps of >= 32 channels
                const int ps = CoutPad / 8;
                const bool inside = (BN <= ps && ps % BN == 0), whole = (BN % ps == 0 && ps % 32 == 0);
                if (!inside && !whole) continue;
              }
              if (chan_per_map % KC) continue;
              if (env_kc && KC != env_kc) continue;
              if (env_bn && BN != env_bn) continue;
              if (stat_gran && BN / stat_gran > 16) continue;
              if (MB * BN > 512) continue;
              const int acc_bufs = (2 * MB * BN <= 512) ? 2 : 1;
              const long long RB = KC * 2;
              const long long a = (box_vox * RB + 1023) / 1024 * 1024, b = ((long long)KHW * BN * RB + 1023) / 1024 * 1024;
              int stages = (int)std::min<long long>(IG_MAXSTAGES, kSmemBudget / (a + b));
              if (stages < 2) continue;
              const long long over = std::max<long long>(0, maxrow - box_vox) * RB;
              if ((a + b) * stages + over > kSmemBudget) continue;  // garbage rows must still read inside our allocation
              const int k_chunks = Ktotal / KC;
              const int steps_total = k_chunks * KD;
              const int n_blocks = CoutPad / BN;
              const long long tiles = (long long)N * ((D + TD - 1) / TD) * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
              const long long items = tiles * n_blocks;
              int ksplit = 1;
              if (max_split > 1 && items * 2 <= num_sms && steps_total > 1 && !B3D_ENV_FLAG("B3D_NOSPLIT")) {
                ksplit = (int)std::min<long long>(std::min(steps_total, max_split), num_sms / items);
                const int sps = (steps_total + ksplit - 1) / ksplit;
                ksplit = (steps_total + sps - 1) / sps;
              }
              const int sps = (steps_total + ksplit - 1) / ksplit;
              const double mclk = (double)MB * KHW * (KC / 16) * sps * mma_clk(BN);
              const double load_clk = (double)(a + b) * sps / 28.0;  // ~28 B/clk/SM of L2->SMEM when every SM streams
              const double epi_clk = (double)MB * (BN / 16) * (ksplit > 1 ? 100.0 : 70.0) + 300.0;
              // every pipeline step is a barrier round trip of the issuing thread (~350 clk measured as the gap between KC = 16 / 4
              // stages and KC = 32 / 2 stages plans of equal MMA and load volume: scripts/plan_sweep.py, round 2)
              double item_clk = std::max(mclk, load_clk) + 350.0 * sps;
              item_clk = (acc_bufs == 2) ? std::max(item_clk, epi_clk) : item_clk + epi_clk;
              const long long its = items * ksplit;
              const long long waves = (its + num_sms - 1) / num_sms;
              double cost = (double)waves * item_clk + 3000.0;
              if (ksplit > 1) cost += 6000.0 + (double)N * D * H * W * CoutPad * 4.0 * (ksplit + 1) / (num_sms * 20.0);  // finalize pass
              if (stages < 3 && KC < 32) cost *= 1.15;   // two stages are enough when one stage carries >= 2 K16 steps per tap
              if (cost < best.cost) {
                best.ok = true; best.cost = cost; best.TD = TD; best.TH = TH; best.TW = TW; best.KC = KC; best.BN = BN;
                best.MB = MB; best.GS = patch ? BW : 8; best.stages = stages; best.acc_bufs = acc_bufs; best.ksplit = ksplit;
                best.patch = patch; best.over = over;
              }
            }
        }
  }
  return best;
}

// Generic launcher.  `x_maps`: nmaps activation views (base pointer + dims + byte strides), each with `chan_per_map`
// channels (multiple of 16).  Output space extent is (N,D,H,W).
struct ActView {
  const void* base;
  long long sW, sH, sD, sN;  // byte strides of the W,H,D,N dims (channel stride = 2 bytes)
  int W, H, D, N;            // extents addressable through this view
  int C;                     // channels addressable (multiple of 8)
};

template <int BN, int KC>
static int ig_launch(const IgParams& P, size_t smem, int grid, cudaStream_t stream) {
  static const cudaError_t attr = cudaFuncSetAttribute(igemm_kernel<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // one-time, thread-safe
  B3D_CHECK_CUDA(attr);
  igemm_kernel<BN, KC><<<grid, IG_THREADS(BN), smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

static int run_igemm(const ActView* views, int nmaps, int chan_per_map, const bf16* wpack, int w_rows, int ks,
                     int N, int D, int H, int W, int Cout, int mode, bf16* out, long long ld_out, int ps_cout,
                     const float* bias, double* stats, int cpg, int stats_groups, int stats_batch, float* ws,
                     size_t ws_bytes, int* err_flag, cudaStream_t stream, const bf16* addend = nullptr, long long ld_add = 0) {
  const int halo = ks / 2, KD = ks, KHW = ks * ks;
  const int num_sms = b3d_num_sms();
  const int CoutPad = w_rows;  // rows in the packed weight tensor (multiple of 16)
  B3D_REQUIRE(chan_per_map % 16 == 0, "igemm: channels per K-map (%d) must be a multiple of 16", chan_per_map);
  B3D_REQUIRE(CoutPad % 16 == 0, "igemm: packed weight rows (%d) must be a multiple of 16", CoutPad);
  B3D_REQUIRE(nmaps >= 1 && nmaps <= 8, "igemm: nmaps out of range");
  int stat_gran = 0;
  if (stats) {
    B3D_REQUIRE(cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8 || cpg % 16 == 0,
                "igemm: channels-per-group %d unsupported (need 1,2,4,8 or a multiple of 16)", cpg);
    stat_gran = cpg >= 16 ? 16 : (cpg >= 4 ? 4 : 1);
  }
  int max_split = 1;  // split-K partials are plain stores into per-split workspace slices [ksplit][V][Cout] fp32
  if (mode == 0 && ws != nullptr && Cout % 8 == 0 && addend == nullptr) max_split = (int)std::min<size_t>(64, ws_bytes / ((size_t)N * D * H * W * Cout * 4));
  IgemmPlan pl = plan_igemm(N, D, H, W, chan_per_map, nmaps, CoutPad, ks, mode, num_sms, max_split, stat_gran);
  if (!pl.ok) {
    b3d_set_error("igemm: no tile plan for N=%d D=%d H=%d W=%d K=%dx%d Cout=%d ks=%d", N, D, H, W, nmaps, chan_per_map,
                  CoutPad, ks);
    return B3D_ERR_UNSUPPORTED;
  }
  if (stats) B3D_REQUIRE(pl.BN / cpg <= 32 && (pl.BN % cpg == 0 || cpg % pl.BN == 0), "igemm: BN %d / cpg %d unsupported", pl.BN, cpg);

  IgParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.D = D; P.H = H; P.W = W; P.Cout = Cout; P.halo = halo; P.ks = ks; P.KD = KD;
  P.TD = pl.TD; P.TH = pl.TH; P.TW = pl.TW;
  P.BH = pl.TH + 2 * halo; P.BW = pl.TW + 2 * halo;
  P.MB = pl.MB; P.GS = pl.GS;
  if (pl.patch) {
    int m = 0;
    for (int pz = 0; pz < P.TD; ++pz)
      for (int pa = 0; pa < P.TH / 16; ++pa)
        for (int pb = 0; pb < P.TW / 8; ++pb) P.mb_base[m++] = (pz * P.BH + pa * 16) * P.BW + pb * 8;
  } else {
    for (int m = 0; m < P.MB; ++m) P.mb_base[m] = m * 128;
  }
  const int BN = pl.BN, KC = pl.KC, RB = KC * 2;
  P.n_blocks = CoutPad / BN;
  P.chunks_per_map = chan_per_map / KC; P.k_chunks = P.chunks_per_map * nmaps;
  P.stages = pl.stages; P.acc_bufs = pl.acc_bufs; P.acc_stride = pl.MB * BN;
  int cols = P.acc_bufs * P.acc_stride, tc = 32;
  while (tc < cols) tc *= 2;
  P.tmem_cols = tc;
  P.tiles_x = (W + P.TW - 1) / P.TW; P.tiles_y = (H + P.TH - 1) / P.TH; P.tiles_z = (D + P.TD - 1) / P.TD;
  P.num_tiles = N * P.tiles_x * P.tiles_y * P.tiles_z;
  P.steps_total = P.k_chunks * KD;
  P.ksplit = pl.ksplit; P.steps_per_split = (P.steps_total + pl.ksplit - 1) / pl.ksplit;
  P.num_items = P.num_tiles * P.n_blocks * P.ksplit;
  const long long box_vox = (long long)P.TD * P.BH * P.BW;
  P.a_tx = (uint32_t)(box_vox * RB); P.w_tx = (uint32_t)(KHW * BN * RB);
  P.w_off = (P.a_tx + 1023u) / 1024u * 1024u;
  P.stage_bytes = P.w_off + (P.w_tx + 1023u) / 1024u * 1024u;
  P.mode = mode; P.out = out; P.ld_out = ld_out; P.ps_cout = ps_cout; P.bias = bias; P.addend = addend; P.ld_add = ld_add;
  P.stats = stats; P.cpg = cpg > 0 ? cpg : 16; P.stats_groups = stats_groups; P.stats_batch = stats_batch;
  P.err = err_flag;

  const bool split = P.ksplit > 1;
  const long long V = (long long)N * D * H * W;
  if (split) {
    P.mode = 1; P.out_f32 = ws; P.ld_f32 = Cout; P.slice_f32 = V * Cout; P.stats = nullptr; P.bias = nullptr;
  }

  for (int m = 0; m < nmaps; ++m) {
    const ActView& v = views[m];
    uint64_t dims[5] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.D, (uint64_t)v.N};
    uint64_t strides[4] = {(uint64_t)v.sW, (uint64_t)v.sH, (uint64_t)v.sD, (uint64_t)v.sN};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)P.BW, (uint32_t)P.BH, (uint32_t)P.TD, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmA[m], v.base, 5, dims, strides, box, RB);
    if (rc) return rc;
  }
  {
    // packed weights [(kh,kw)][kd][rows][Kp] (kd fastest of the taps): a pipeline step takes the KHW taps of one kd
    const int Kp = chan_per_map * nmaps;
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)CoutPad, (uint64_t)KD, (uint64_t)KHW};
    uint64_t strides[3] = {(uint64_t)Kp * 2, (uint64_t)CoutPad * Kp * 2, (uint64_t)KD * CoutPad * Kp * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)BN, 1, (uint32_t)KHW};
    int rc = b3d_encode_tmap_bf16(&P.tmW, wpack, 4, dims, strides, box, RB);
    if (rc) return rc;
  }
  const size_t smem = (size_t)P.stages * P.stage_bytes + std::max<size_t>((size_t)pl.over, 16 * IG_MAXSTAGES + 48 + 64 * 8 + 128) + 1024;
  B3D_REQUIRE(smem <= 227 * 1024, "igemm: smem %zu too large", smem);
  const int grid = std::min(P.num_items, num_sms);
  snprintf(g_last_plan, sizeof(g_last_plan), "tile %dx%dx%d %s KC%d BN%d MB%d stages%d acc%d split%d items%d", P.TD, P.TH, P.TW,
           pl.patch ? "patch" : "linear", KC, BN, P.MB, P.stages, P.acc_bufs, P.ksplit, P.num_items);
  if (B3D_ENV_FLAG("B3D_VERBOSE"))
    fprintf(stderr,
            "[b3d] igemm N%d D%d H%d W%d K=%dx%d Cout=%d(ks%d mode%d) tile %dx%dx%d %s KC%d BN%d MB%d stages%d acc%d "
            "split%d items%d smem%zu tmem%d\n",
            N, D, H, W, nmaps, chan_per_map, Cout, ks, mode, P.TD, P.TH, P.TW, pl.patch ? "patch" : "linear", KC, BN, P.MB,
            P.stages, P.acc_bufs, P.ksplit, P.num_items, smem, P.tmem_cols);
  int rc = B3D_ERR_UNSUPPORTED;
#define IG_CASE(B, K) if (BN == B && KC == K) rc = ig_launch<B, K>(P, smem, grid, stream);
  IG_CASE(16, 16) IG_CASE(16, 32) IG_CASE(16, 64)
  IG_CASE(32, 16) IG_CASE(32, 32) IG_CASE(32, 64)
  IG_CASE(64, 16) IG_CASE(64, 32) IG_CASE(64, 64)
  IG_CASE(128, 16) IG_CASE(128, 32) IG_CASE(128, 64)
  IG_CASE(256, 16) IG_CASE(256, 32) IG_CASE(256, 64)
#undef IG_CASE
  if (rc) return rc;
  if (split) {
    const int chunks = Cout / 8;
    const long long total = V * chunks;
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms * 8);
    // keep each block inside as few samples as possible
    igemm_finalize_kernel<<<blocks, 256, 0, stream>>>(ws, P.ksplit, V, Cout, (long long)D * H * W, bias, out, ld_out, stats,
                                                      cpg > 0 ? cpg : 16, stats_groups, stats_batch); ++g_b3d_launches;
    B3D_CHECK_CUDA(cudaGetLastError());
  }
  return B3D_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

// Stride-1 "same" convolution, ks in {1,3}.  x: NDHWC bf16 with voxel pitch ldx (elements), Cin channels used (mult of 16).
// wpack: packed [Cin/8][ks^3][rows][8] (rows = roundup16(Cout)).  y: NDHWC bf16 pitch ldy.  stats: optional double
// [N or 1][groups][2] accumulated (+=) with sum / sum of squares of the (bias-added, fp32) outputs.
int b3d_conv_fprop_add(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, const void* addend,
                       long long ld_add, void* y, long long ldy, int N, int D, int H, int W, int Cin, int Cout, int ks,
                       double* stats, int groups, int stats_batch, void* ws, size_t ws_bytes, int* err_flag, void* stream) {
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
  if (addend != nullptr) B3D_REQUIRE(ks == 1 && ld_add % 8 == 0, "conv_fprop_add: the fused addend is implemented for 1x1x1 convolutions");
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
                   groups, stats_batch, (float*)ws, ws_bytes, err_flag, (cudaStream_t)stream, (const bf16*)addend, ld_add);
}

// y = conv1x1(x) + addend with the ADDEND ON THE TENSOR CORE: K = [x channels | addend channels], B = [W ; I].  The addend
// tile arrives through the same TMA pipeline as x (prefetched `stages` tiles ahead) instead of 64-byte per-thread loads in
// the epilogue, whose DRAM latency the 4 epilogue warps cannot hide (profiles/ncu_r1_igemm_pw_add_16_32.txt: 42 % of the stall
// samples).  bf16 addend x 1.0 accumulated in fp32 = exactly the epilogue add.  wpack_aug: bf16 [rows][Cin + Cout] =
// [dgrad-packed W | identity]; y may alias addend (in-place gradient accumulation: every tile reads only rows it later writes).
int b3d_conv1_add_mma(const void* x, long long ldx, const void* addend, long long ld_add, const void* wpack_aug, int w_rows,
                      void* y, long long ldy, int N, int D, int H, int W, int Cin, int Cout, int* err_flag, void* stream) {
  B3D_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0 && w_rows == Cout, "conv1_add_mma: Cin/Cout must be multiples of 16");
  B3D_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ld_add % 8 == 0, "conv1_add_mma: pitches must be multiples of 8 elements");
  int a = Cin, b = Cout;
  while (b) { const int t = a % b; a = b; b = t; }
  const int cpm = a;                       // channels per K-map = gcd(Cin, Cout)
  const int nx = Cin / cpm, na = Cout / cpm;
  B3D_REQUIRE(cpm % 16 == 0 && nx + na <= 8, "conv1_add_mma: %d + %d channels need more than 8 K-maps", Cin, Cout);
  long long V = (long long)N * D * H * W;   // pointwise: flatten all voxels to a [rows][256] plane so tiles are dense
  int ww = 256;
  while (ww > 1 && V % ww) ww /= 2;
  long long hh = V / ww;
  int dd = 1;
  while (hh > 32768 && hh % 2 == 0) { hh /= 2; dd *= 2; }
  ActView v[8];
  for (int m = 0; m < nx + na; ++m) {
    const bool isx = m < nx;
    const long long ld = isx ? ldx : ld_add;
    v[m].base = (const char*)(isx ? x : addend) + (long long)(isx ? m : m - nx) * cpm * 2;
    v[m].C = cpm; v[m].W = ww; v[m].H = (int)hh; v[m].D = dd; v[m].N = 1;
    v[m].sW = ld * 2; v[m].sH = v[m].sW * ww; v[m].sD = v[m].sH * hh; v[m].sN = v[m].sD * dd;
  }
  return run_igemm(v, nx + na, cpm, (const bf16*)wpack_aug, w_rows, 1, 1, dd, (int)hh, ww, Cout, 0, (bf16*)y, ldy, 0, nullptr,
                   nullptr, 0, 0, 0, nullptr, 0, err_flag, (cudaStream_t)stream);
}

// Tuning hooks (not used by the product path): force tile-plan parameters of the implicit-GEMM planner (0 = planner's choice)
// and read back the plan of this thread's last launch — scripts/plan_sweep.py sweeps plans in one process with them.
int b3d_set_plan_override(int td, int th, int tw, int kc, int bn) {
  g_plan_override[0] = td; g_plan_override[1] = th; g_plan_override[2] = tw; g_plan_override[3] = kc; g_plan_override[4] = bn;
  return B3D_OK;
}
const char* b3d_last_plan() { return g_last_plan; }

int b3d_conv_fprop(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
                   int N, int D, int H, int W, int Cin, int Cout, int ks, double* stats, int groups, int stats_batch,
                   void* ws, size_t ws_bytes, int* err_flag, void* stream) {
  return b3d_conv_fprop_add(x, ldx, wpack, w_rows, bias, nullptr, 0, y, ldy, N, D, H, W, Cin, Cout, ks, stats, groups, stats_batch,
                            ws, ws_bytes, err_flag, stream);
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
