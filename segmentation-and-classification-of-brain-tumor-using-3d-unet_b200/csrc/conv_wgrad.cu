// conv_wgrad.cu — weight gradients of the 3-D convolutions on tcgen05 tensor cores (sm_100a).
//
//   dW[tap][ci][co] = Σ_{n,z,y,x} X[n, z+kd-1, y+kh-1, x+kw-1][ci] * dY[n,z,y,x][co]
// (what autograd's convolution_backward computes for nn.Conv3d / nn.ConvTranspose3d in /root/reference/main.py:121,130,
//  216,219,229,252,258 — 52 % of the reference's CPU step time).
//
// GEMM view: D[M = input channels (x stacked taps)][N = output channels] += A[M][K] * B[K][N] with K = voxels.
// Both operands are "MN-major" for the UMMA (the contracted index, voxels, is the slow one in NDHWC), staged by TMA as
//     X  slot : [row][8-channel chunk][x position][8 ch = 16 B]          (box (8, BW, CI8, BH, 1) of a 5-D view whose
//     dY slot : [row][8-channel chunk][x position][8 ch = 16 B]           chunk dimension has a 16-byte stride)
// so that  * a K=16 step is 16 consecutive x positions of one row (LBO = 128 B),
//          * the next 8-channel chunk is SBO = BW*16 B away, and — because chunks of consecutive ROWS are also SBO
//            apart — M = 128 can hold G = 128/ci_blk vertically shifted copies of the tile: for Cin <= 32 the three kh
//            taps are STACKED on M (rows y, y+1, y+2 (+1 unused)), which fills the 128-row tensor-core tile that a 32-wide
//            channel block alone would leave 75 % empty,
//          * kw shifts are +16 B on the A start address, kd shifts select another X plane of a 3-deep rolling ring.
// Accumulators (one [128 x BN] fp32 tile per (kd,kw) chain) stay in TMEM across ALL tiles a CTA processes for the same
// output block and are flushed once with fp32 atomics into dwacc[tap][ci][co]; wgrad_finalize permutes into the
// reference's [co][ci][kd][kh][kw] layout.
// "Plane mode" (W not a multiple of 16: the 8^3 / 4^3 levels) linearises a whole halo plane as the K run instead of rows.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

struct alignas(64) WgradParams {
  CUtensorMap tmX;
  CUtensorMap tmY[8];
  int variant;      // 0: kh stacked on M, 9 (kd,kw) chains ; 1: key=(kd,kh), 3 kw chains ; 2: pointwise, 1 chain
  int plane_mode;   // 0: K runs along rows ; 1: K runs along a linearised halo plane
  int N, D, H, W, halo;
  int TH, BH, BW, TZ;
  int ci_blk, CI8, G, BN, CO8, n_cib, n_cob, nmapsY;
  int nchains, xspan, R;
  uint32_t x_slot_bytes, y_slot_bytes, x_tx_bytes, y_tx_bytes;
  int XP, YP;       // plane mode: padded positions per chunk in the X / dY slot
  int ksteps;
  int tiles_y, tiles_z, items_per_key, num_keys, num_items, items_per_cta;
  float* dwacc;
  int Cin, Cout, Cout_pad, Cin_pad;
  int tmem_cols;
  int exclusive;    // every key is processed by exactly one CTA: flush with plain stores (no memset, no atomics)
  int* err;
};

#define WG_THREADS 192

struct KeyInfo { int cib, cob, kd, kh, t8; };

__device__ __forceinline__ KeyInfo decode_key(const WgradParams& P, int key) {
  KeyInfo k; k.kd = -1; k.kh = -1; k.t8 = 0;
  if (P.variant == 0) { k.cob = key % P.n_cob; k.cib = key / P.n_cob; }
  else if (P.variant == 1) { k.kh = key % 3; key /= 3; k.kd = key % 3; key /= 3; k.cob = key % P.n_cob; k.cib = key / P.n_cob; }
  else { k.t8 = key % P.nmapsY; key /= P.nmapsY; k.cob = key % P.n_cob; k.cib = key / P.n_cob; }
  return k;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sX = smem_u32(smem);
  const uint32_t sY = sX + P.R * P.x_slot_bytes;
  const uint32_t y_end = sY + 2 * P.y_slot_bytes;
  // aux region sits after a guard gap (garbage M rows may read past the last slot)
  const uint32_t aux_off = (uint32_t)(P.R * P.x_slot_bytes + 2 * P.y_slot_bytes);
  uint8_t* aux = smem + aux_off;
  const uint32_t xfull0 = smem_u32(aux);            // [R]
  const uint32_t xempty0 = xfull0 + 8 * 4;          // [R] (R <= 4)
  const uint32_t yfull0 = xempty0 + 8 * 4;          // [2]
  const uint32_t yempty0 = yfull0 + 16;             // [2]
  const uint32_t accfull = yempty0 + 16;
  const uint32_t accempty = accfull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 8 * 4 * 2 + 32 + 16);
  volatile uint32_t* s_started = reinterpret_cast<volatile uint32_t*>(aux + 8 * 4 * 2 + 32 + 16 + 8);  // chains that received MMAs
  (void)y_end;

  if (threadIdx.x == 0) {
    if (sX & 127u) { if (P.err) atomicExch(P.err, 19); __trap(); }
    for (int i = 0; i < 4; ++i) { mbar_init(xfull0 + 8 * i, 1); mbar_init(xempty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(yfull0 + 8 * i, 1); mbar_init(yempty0 + 8 * i, 1); }
    mbar_init(accfull, 1); mbar_init(accempty, 4);
    mbar_fence_init();
  }
  if (P.plane_mode) {
    // zero the staging slots once: tails beyond the TMA boxes are read as K elements and must be finite zeros
    uint4* p = reinterpret_cast<uint4*>(smem);
    const uint32_t n16 = aux_off / 16;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int item_beg = blockIdx.x * P.items_per_cta;
  const int item_end = min(P.num_items, item_beg + P.items_per_cta);
  const int zoff0 = (P.variant == 0) ? -P.halo : 0;

  if (warp == 0) {
    // ======================= TMA producer =======================
    // The whole warp walks the schedule (uniform control flow).  In plane mode a plane is CI8 / CO8 boxes of one 16-byte channel
    // chunk each (they sit at a padded pitch in shared memory): lane c issues the box of chunk c, so a plane costs one issue
    // slot instead of 16-32 serial TMA issues on one thread — at the 4^3 / 8^3 levels (8 tiny planes per key) that serial
    // issue, not the flush, bounded the kernel (ncu: 44 % of the stall samples in the accumulator-full wait of the flush warps).
    {
      if (lane == 0) tma_prefetch_desc(&P.tmX);
      uint32_t Q = 0, T = 0;  // running X-plane / dY-plane fill counters
      for (int item = item_beg; item < item_end; ++item) {
        const int key = item / P.items_per_key;
        int r = item - key * P.items_per_key;
        const KeyInfo k = decode_key(P, key);
        const int ty = r % P.tiles_y; r /= P.tiles_y;
        const int tz = r % P.tiles_z; const int n = r / P.tiles_z;
        const int y0 = ty * P.TH, z0 = tz * P.TZ;
        const int nz = min(P.TZ, P.D - z0);
        const int zoff = (P.variant == 1) ? (k.kd - 1) : zoff0;
        const int yoff = P.plane_mode ? -P.halo : ((P.variant == 0) ? -P.halo : (P.variant == 1 ? k.kh - 1 : 0));
        const int nplanes = nz + P.xspan - 1;
        for (int s = 0; s < nplanes; ++s) {
          {  // X plane s
            const uint32_t slot = Q % P.R, ph = (Q / P.R) & 1u;
            mbar_wait(xempty0 + 8 * slot, ph ^ 1, P.err, 11);
            const int z = z0 + s + zoff;
            const uint32_t fb = xfull0 + 8 * slot;
            if (z >= 0 && z < P.D) {
              if (lane == 0) mbar_expect_tx(fb, P.x_tx_bytes);
              __syncwarp();
              if (P.plane_mode) {
                for (int c8 = lane; c8 < P.CI8; c8 += 32)  // one box per chunk: chunks sit at the padded pitch XP
                  tma_load_5d(sX + slot * P.x_slot_bytes + c8 * P.XP * 16, &P.tmX, fb, 0, -P.halo, y0 + yoff,
                              k.cib * P.CI8 + c8, n * P.D + z);
              } else if (lane == 0)
                tma_load_5d(sX + slot * P.x_slot_bytes, &P.tmX, fb, 0, -P.halo, k.cib * P.CI8, y0 + yoff, n * P.D + z);
            } else if (lane == 0) {
              mbar_arrive(fb);  // out-of-volume plane: nothing to load, the consumer skips it
            }
            ++Q;
          }
          if (s >= P.xspan - 1) {  // dY plane t
            const int t = s - (P.xspan - 1);
            const uint32_t slot = T & 1u, ph = (T >> 1) & 1u;
            mbar_wait(yempty0 + 8 * slot, ph ^ 1, P.err, 12);
            const uint32_t fb = yfull0 + 8 * slot;
            if (lane == 0) mbar_expect_tx(fb, P.y_tx_bytes);
            __syncwarp();
            if (P.plane_mode) {
              for (int c8 = lane; c8 < P.CO8; c8 += 32)
                tma_load_5d(sY + slot * P.y_slot_bytes + c8 * P.YP * 16, &P.tmY[k.t8], fb, 0, 0, y0, k.cob * P.CO8 + c8,
                            n * P.D + z0 + t);
            } else if (lane == 0)
              tma_load_5d(sY + slot * P.y_slot_bytes, &P.tmY[k.t8], fb, 0, 0, k.cob * P.CO8, y0, n * P.D + z0 + t);
            ++T;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform control flow, one elected lane issues) =======================
    {
      const uint32_t idesc = umma_idesc_bf16(128, P.BN, 1, 1);
      const uint32_t lbo = (128u >> 4) << 16;
      const uint64_t hiA = umma_desc_hi(P.plane_mode ? (uint32_t)P.XP * 16u : (uint32_t)P.BW * 16u);
      const uint64_t hiB = umma_desc_hi(P.plane_mode ? (uint32_t)P.YP * 16u : (uint32_t)P.W * 16u);
      uint32_t Q = 0, T = 0, started = 0, flushes = 0;
      int cur_key = -1;
      const int nkw = (P.halo ? 3 : 1);
      for (int item = item_beg; item < item_end; ++item) {
        const int key = item / P.items_per_key;
        int r = item - key * P.items_per_key;
        const KeyInfo k = decode_key(P, key);
        const int ty = r % P.tiles_y; r /= P.tiles_y;
        const int tz = r % P.tiles_z;
        const int y0 = ty * P.TH, z0 = tz * P.TZ;
        const int nz = min(P.TZ, P.D - z0);
        const int rows = min(P.TH, P.H - y0);
        if (key != cur_key) {
          if (cur_key >= 0) {
            if (elect_one()) {
              *s_started = started;
              __threadfence_block();
              umma_commit(accfull);
            }
            __syncwarp();
            mbar_wait(accempty, flushes & 1u, P.err, 13);
            tc_fence_after();
            ++flushes;
          }
          cur_key = key; started = 0;
        }
        const int zoff = (P.variant == 1) ? (k.kd - 1) : zoff0;
        const uint32_t Q0 = Q;
        for (int t = 0; t < nz; ++t) {
          const uint32_t yslot = T & 1u, yph = (T >> 1) & 1u;
          mbar_wait(yfull0 + 8 * yslot, yph, P.err, 14);
          const uint32_t yb16 = ((sY + yslot * P.y_slot_bytes) >> 4) | lbo;
          for (int a = 0; a < P.xspan; ++a) {  // a = kd for variant 0
            const uint32_t q = Q0 + t + a;
            const uint32_t xslot = q % P.R, xph = (q / P.R) & 1u;
            mbar_wait(xfull0 + 8 * xslot, xph, P.err, 15);
            tc_fence_after();
            const int z = z0 + t + a + zoff;
            if (z >= 0 && z < P.D) {
              const uint32_t xb16 = ((sX + xslot * P.x_slot_bytes) >> 4) | lbo;
              // chains touched by this plane: variant 0 -> a*3 + kw, else kw
              const int chain0 = (P.variant == 0) ? a * 3 : 0;
              const uint32_t d0 = tmem_base + chain0 * P.BN;
              const uint32_t acc0 = (started >> chain0) & 1u, acc1 = (started >> (chain0 + 1)) & 1u,
                             acc2 = (started >> (chain0 + 2)) & 1u;
              if (elect_one()) {
                uint32_t c0 = acc0, c1 = acc1, c2 = acc2;
                if (!P.plane_mode) {
                  for (int y = 0; y < rows; ++y) {
                    uint32_t alo = xb16 + (uint32_t)(y * P.CI8) * P.BW;
                    uint32_t blo = yb16 + (uint32_t)(y * P.CO8) * P.W;
                    if (nkw == 1) {
                      for (int j = 0; j < P.ksteps; ++j) {
                        umma_bf16_ss(d0, hiA | alo, hiB | blo, idesc, c0); c0 = 1u;
                        alo += 16; blo += 16;
                      }
                    } else {
                      for (int j = 0; j < P.ksteps; ++j) {
                        umma_bf16_ss(d0, hiA | alo, hiB | blo, idesc, c0); c0 = 1u;
                        umma_bf16_ss(d0 + P.BN, hiA | (alo + 1), hiB | blo, idesc, c1); c1 = 1u;
                        umma_bf16_ss(d0 + 2 * P.BN, hiA | (alo + 2), hiB | blo, idesc, c2); c2 = 1u;
                        alo += 16; blo += 16;
                      }
                    }
                  }
                } else {
                  const uint32_t khoff = (P.variant == 1) ? (uint32_t)(k.kh * P.BW) : 0u;
                  uint32_t alo = xb16 + khoff, blo = yb16;
                  for (int j = 0; j < P.ksteps; ++j) {
                    umma_bf16_ss(d0, hiA | alo, hiB | blo, idesc, c0); c0 = 1u;
                    if (nkw == 3) {
                      umma_bf16_ss(d0 + P.BN, hiA | (alo + 1), hiB | blo, idesc, c1); c1 = 1u;
                      umma_bf16_ss(d0 + 2 * P.BN, hiA | (alo + 2), hiB | blo, idesc, c2); c2 = 1u;
                    }
                    alo += 16; blo += 16;
                  }
                }
              }
              __syncwarp();
              const bool any = P.plane_mode ? (P.ksteps > 0) : (rows > 0 && P.ksteps > 0);
              if (any) started |= (nkw == 3 ? 7u : 1u) << chain0;
            }
            if (a == 0) {  // plane t (+zoff) is not needed by later output planes
              if (elect_one()) umma_commit(xempty0 + 8 * xslot);
              __syncwarp();
            }
          }
          if (elect_one()) umma_commit(yempty0 + 8 * yslot);
          __syncwarp();
          ++T;
        }
        // planes nz .. nz+xspan-2 of this item were only partially consumed: release them
        for (int a = 1; a < P.xspan; ++a) {
          const uint32_t q = Q0 + nz - 1 + a;
          if (elect_one()) umma_commit(xempty0 + 8 * (q % P.R));
          __syncwarp();
        }
        Q = Q0 + nz + P.xspan - 1;
      }
      if (cur_key >= 0) {
        if (elect_one()) {
          *s_started = started;
          __threadfence_block();
          umma_commit(accfull);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue: flush TMEM accumulators with fp32 atomics =======================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int cur_key = -1;
    uint32_t flushes = 0;
    for (int item = item_beg; item <= item_end; ++item) {
      const int key = (item < item_end) ? item / P.items_per_key : -2;
      if (key == cur_key) continue;
      if (cur_key >= 0) {
        mbar_wait(accfull, flushes & 1u, P.err, 16);
        tc_fence_after();
        const uint32_t started = *s_started;  // chains without any MMA hold stale TMEM contents: skip them
        const KeyInfo k = decode_key(P, cur_key);
        const int khs = m / P.ci_blk;
        const int ci = k.cib * P.ci_blk + (m - khs * P.ci_blk);
        int kh = 0;
        bool row_ok = ci < P.Cin;
        if (P.variant == 0) { kh = khs; row_ok = row_ok && (kh < 3); }
        else { kh = k.kh; row_ok = row_ok && (khs == 0); }
        for (int chain = 0; chain < P.nchains; ++chain) {
          int tap;
          if (P.variant == 0) tap = ((chain / 3) * 3 + kh) * 3 + (chain % 3);
          else if (P.variant == 1) tap = (k.kd * 3 + kh) * 3 + chain;
          else tap = k.t8;
          float* dst = P.dwacc + ((long long)tap * P.Cin_pad + ci) * P.Cout_pad + k.cob * P.BN;
          const uint32_t trow = tmem_base + chain * P.BN + ((uint32_t)(q * 32) << 16);
          for (int j0 = 0; j0 < P.BN; j0 += 16) {
            uint32_t rr[16];
            tmem_ld16(trow + j0, rr);
            tmem_ld_wait();
            if (P.exclusive) {
              if (row_ok && k.cob * P.BN + j0 + 16 <= P.Cout_pad) {
                const bool live = (started >> chain) & 1u;   // chains that never received an MMA hold stale TMEM: store zeros
                float4* d4 = reinterpret_cast<float4*>(dst + j0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  d4[j] = live ? make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                                             __uint_as_float(rr[4 * j + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            } else if (row_ok && ((started >> chain) & 1u)) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (k.cob * P.BN + j0 + j < P.Cout_pad) atomicAdd(dst + j0 + j, __uint_as_float(rr[j]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty);
        ++flushes;
      }
      cur_key = key;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, P.tmem_cols); }
}

// dwacc fp32 [taps][Cin_pad][Cout_pad] -> reference layouts (fp32), optionally accumulating into an existing gradient
//   mode 0: conv   dW[co][ci][tap]           mode 1: convT  dWt[ci][co][t8]
// One block per (FIN_CIB input channels, 32 output channels): coalesced 128-byte reads of dwacc rows, smem transpose, and on
// the write side runs of FIN_CIB * ntaps consecutive floats per output channel (mode 0: 864 B, sector aligned when
// Cin % 8 == 0) or 32 * ntaps per input channel (mode 1).  One input channel per block left 108-byte runs and ran the
// 1024 x 1024 x 27 gradient at 2.3 TB/s.
#define FIN_CIB 8
// NT = taps (27 / 8 / 1): every index split below is a division by a compile-time constant when the block is full (8 input
// channels) — the generic version spent ~100 instructions per element on runtime divisions and scalar 4-byte loads and was
// issue-bound (2 TB/s on the 113 MB bottleneck gradient, profiles/overlap_probe_r2.txt context: 0.67 ms of the step is exposed
// deep-level weight-gradient work).  Loads: one 16-byte load per lane, four dwacc rows per warp instruction, all in flight.
template <int NT>
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ acc, float* __restrict__ dw, int mode,
                                                             int Cin, int Cout, int Cin_pad, int Cout_pad, int accumulate) {
  __shared__ float tile[FIN_CIB][NT][33];
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * FIN_CIB;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nci = min(FIN_CIB, Cin - ci0), nco = min(32, Cout - co0);
  const int sub = lane >> 3, l4 = (lane & 7) * 4;      // row within a group of four, first of this lane's four columns
  const bool col_ok = co0 + l4 < Cout_pad;             // Cout_pad % 16 == 0: a 16-byte chunk is inside the row or entirely outside
  static_assert(FIN_CIB == 8, "the full-block path splits a row index with r >> 3 / r & 7");
  constexpr int ROWS_FULL = FIN_CIB * NT;
  constexpr int U = (ROWS_FULL + 31) / 32;             // row groups per warp: 8 warps x 4 rows x U >= FIN_CIB * NT
  if (nci == FIN_CIB) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = (u * 8 + w) * 4 + sub;             // row r <-> (tap = r / 8, channel = r % 8): rows of one tap are adjacent
      v[u] = (r < ROWS_FULL && col_ok)
                 ? __ldcs(reinterpret_cast<const float4*>(acc + ((long long)(r >> 3) * Cin_pad + ci0 + (r & 7)) * Cout_pad + co0 + l4))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = (u * 8 + w) * 4 + sub;
      if (r < ROWS_FULL) {
        float* t = &tile[r & 7][r >> 3][l4];
        t[0] = v[u].x; t[1] = v[u].y; t[2] = v[u].z; t[3] = v[u].w;
      }
    }
  } else {
    const int nrows = nci * NT;
    for (int r = w * 4 + sub; r < nrows; r += 32) {
      const int tap = r / nci, cil = r - tap * nci;
      const float4 q = col_ok ? __ldcs(reinterpret_cast<const float4*>(acc + ((long long)tap * Cin_pad + ci0 + cil) * Cout_pad + co0 + l4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      float* t = &tile[cil][tap][l4];
      t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
    }
  }
  __syncthreads();
  if (mode == 0) {   // dW[co][ci][tap]: per output channel one run of nci * NT consecutive floats
    if (nci == FIN_CIB) {
      constexpr int RUN = FIN_CIB * NT;
      for (int idx = threadIdx.x; idx < nco * RUN; idx += 256) {
        const int col = idx / RUN, j = idx - col * RUN;
        const int cil = j / NT, tap = j - cil * NT;
        const long long o = ((long long)(co0 + col) * Cin + ci0) * NT + j;
        const float val = tile[cil][tap][col];
        dw[o] = accumulate ? dw[o] + val : val;
      }
    } else {
      const int run = nci * NT;
      for (int idx = threadIdx.x; idx < nco * run; idx += 256) {
        const int col = idx / run, j = idx - col * run;
        const int cil = j / NT, tap = j - cil * NT;
        const long long o = ((long long)(co0 + col) * Cin + ci0) * NT + j;
        const float val = tile[cil][tap][col];
        dw[o] = accumulate ? dw[o] + val : val;
      }
    }
  } else {           // dWt[ci][co][t8]: per input channel one run of nco * NT consecutive floats
    const int run = nco * NT;
    for (int idx = threadIdx.x; idx < nci * run; idx += 256) {
      const int cil = idx / run, j = idx - cil * run;
      const int col = j / NT, tap = j - col * NT;
      const long long o = ((long long)(ci0 + cil) * Cout + co0) * NT + j;
      const float val = tile[cil][tap][col];
      dw[o] = accumulate ? dw[o] + val : val;
    }
  }
}

static cudaError_t launch_wgrad_finalize(const float* acc, float* dw, int mode, int ntaps, int Cin, int Cout, int Cin_pad,
                                         int Cout_pad, int accumulate, cudaStream_t st) {
  const dim3 grid((Cout + 31) / 32, (Cin + FIN_CIB - 1) / FIN_CIB);
  if (ntaps == 27) wgrad_finalize_kernel<27><<<grid, 256, 0, st>>>(acc, dw, mode, Cin, Cout, Cin_pad, Cout_pad, accumulate);
  else if (ntaps == 8) wgrad_finalize_kernel<8><<<grid, 256, 0, st>>>(acc, dw, mode, Cin, Cout, Cin_pad, Cout_pad, accumulate);
  else if (ntaps == 1) wgrad_finalize_kernel<1><<<grid, 256, 0, st>>>(acc, dw, mode, Cin, Cout, Cin_pad, Cout_pad, accumulate);
  else return cudaErrorInvalidValue;
  ++g_b3d_launches;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static const int kWgSmemBudget = 227 * 1024 - 1024;

struct YView { const void* base; long long sW, sH, sND; };

static int run_wgrad(const void* x, long long ldx, int Cin_use, const YView* yv, int nmapsY, int N, int D, int H, int W,
                     int Cout, int ks, float* dwacc, int Cin_pad, int Cout_pad, int* err_flag, cudaStream_t stream) {
  WgradParams P;
  memset(&P, 0, sizeof(P));
  const int halo = ks / 2;
  const int num_sms = b3d_num_sms();
  B3D_REQUIRE(Cin_use % 8 == 0 && Cout % 8 == 0, "wgrad: channels must be multiples of 8");
  P.N = N; P.D = D; P.H = H; P.W = W; P.halo = halo;
  P.Cin = Cin_use; P.Cout = Cout; P.Cin_pad = Cin_pad; P.Cout_pad = Cout_pad; P.nmapsY = nmapsY;
  P.plane_mode = (W % 16 != 0) ? 1 : 0;
  P.BW = W + 2 * halo;
  B3D_REQUIRE(P.BW <= 256, "wgrad: W=%d too wide for one TMA box", W);
  if (ks == 3 && !P.plane_mode && Cin_use <= 64) {
    P.variant = 0; P.ci_blk = std::min(Cin_use, 32); P.G = 128 / P.ci_blk; P.nchains = 9; P.xspan = 3; P.R = 3;
    P.BN = (Cout_pad % 48 == 0) ? 48 : (Cout_pad >= 32 ? 32 : 16);
  } else if (ks == 3) {
    P.variant = 1; P.ci_blk = std::min(Cin_use, 128); P.G = 1; P.nchains = 3; P.xspan = 1; P.R = 2;
    P.BN = std::min(Cout_pad, 128);
  } else {
    P.variant = 2; P.ci_blk = std::min(Cin_use, 128); P.G = 1; P.nchains = 1; P.xspan = 1; P.R = 2;
    P.BN = std::min(Cout_pad, 256);
  }
  B3D_REQUIRE(Cin_use % P.ci_blk == 0, "wgrad: Cin=%d not a multiple of the channel block %d", Cin_use, P.ci_blk);
  P.CI8 = P.ci_blk / 8; P.CO8 = P.BN / 8;
  P.n_cib = Cin_use / P.ci_blk; P.n_cob = (Cout_pad + P.BN - 1) / P.BN;
  B3D_REQUIRE(P.BN % 16 == 0 && Cout_pad % 16 == 0, "wgrad: Cout_pad must be a multiple of 16");
  int cols = P.nchains * P.BN, tc = 32;
  while (tc < cols) tc *= 2;
  P.tmem_cols = tc;
  B3D_REQUIRE(tc <= 512, "wgrad: TMEM overflow");

  // tile rows (and, if even one row does not fit, a narrower N block)
  int TH;
  long long guard = 0;
  if (P.plane_mode) {
    TH = H;
    const int BH = H + 2 * halo;
    P.BH = BH;
    P.ksteps = ((H - 1) * P.BW + W + 15) / 16;
    const int maxpos = P.ksteps * 16 + (P.halo ? 2 * P.BW + 2 : 0);
    P.XP = (std::max(BH * P.BW, maxpos) + 7) / 8 * 8;
    P.YP = (std::max(H * P.BW, P.ksteps * 16) + 7) / 8 * 8;
    // one TMA box per 8-channel chunk, landing at the padded pitch (XP / YP positions); the pads stay zero
    P.x_slot_bytes = (uint32_t)(((long long)P.CI8 * P.XP * 16 + 1023) / 1024 * 1024);
    P.y_slot_bytes = (uint32_t)(((long long)P.CO8 * P.YP * 16 + 1023) / 1024 * 1024);
    P.x_tx_bytes = (uint32_t)P.CI8 * BH * P.BW * 16;
    P.y_tx_bytes = (uint32_t)P.CO8 * H * P.BW * 16;
    B3D_REQUIRE(BH <= 256 && P.CI8 <= 256, "wgrad: plane too large");
    guard = (long long)(16 - P.CI8) * P.XP * 16 + 1024;   // garbage M rows beyond the staged chunks
  } else {
    TH = 0;
    while (TH == 0) {
      for (int cand = 16; cand >= 1; cand /= 2) {
        if (cand > H && cand > 1) continue;
        const int BH = (P.variant == 0) ? cand + 2 : cand;
        const long long xs = ((long long)BH * P.CI8 * P.BW * 16 + 1023) / 1024 * 1024;
        const long long ys = ((long long)cand * P.CO8 * W * 16 + 1023) / 1024 * 1024;
        const long long over = std::max<long long>(0, (long long)((cand - 1) * P.CI8 + 16 - BH * P.CI8) * P.BW * 16);
        const long long g = std::max<long long>(0, over - 2 * ys) + 1024;
        if (P.R * xs + 2 * ys + g + 256 <= kWgSmemBudget) { TH = cand; guard = g; break; }
      }
      if (TH == 0) {
        B3D_REQUIRE(P.BN > 16, "wgrad: tile does not fit in shared memory (W=%d Cin=%d Cout=%d)", W, Cin_use, Cout);
        P.BN /= 2; P.CO8 = P.BN / 8; P.n_cob = (Cout_pad + P.BN - 1) / P.BN;
      }
    }
    P.BH = (P.variant == 0) ? TH + 2 : TH;
    P.ksteps = W / 16;
    P.x_slot_bytes = (uint32_t)(((long long)P.BH * P.CI8 * P.BW * 16 + 1023) / 1024 * 1024);
    P.y_slot_bytes = (uint32_t)(((long long)TH * P.CO8 * W * 16 + 1023) / 1024 * 1024);
    P.x_tx_bytes = (uint32_t)P.BH * P.CI8 * P.BW * 16;
    P.y_tx_bytes = (uint32_t)TH * P.CO8 * W * 16;
    B3D_REQUIRE(P.BH <= 256, "wgrad: tile too tall");
  }
  P.TH = TH;
  { int c2 = P.nchains * P.BN, t2 = 32; while (t2 < c2) t2 *= 2; P.tmem_cols = t2; }
  P.tiles_y = (H + TH - 1) / TH;
  if (P.variant == 0) P.num_keys = P.n_cib * P.n_cob;
  else if (P.variant == 1) P.num_keys = P.n_cib * P.n_cob * 9;
  else P.num_keys = P.n_cib * P.n_cob * nmapsY;
  int TZ = D;
  while (TZ > 4 && (long long)P.num_keys * N * P.tiles_y * ((D + TZ - 1) / TZ) < 2LL * num_sms) TZ /= 2;
  if (TZ < 1) TZ = 1;
  P.TZ = TZ;
  P.tiles_z = (D + TZ - 1) / TZ;
  P.items_per_key = N * P.tiles_y * P.tiles_z;
  P.num_items = P.items_per_key * P.num_keys;
  const int grid = std::min(P.num_items, num_sms);
  P.items_per_cta = (P.num_items + grid - 1) / grid;
  // many keys: give every CTA whole keys, so that the flush needs neither a zeroed accumulator nor atomics
  if (P.items_per_cta % P.items_per_key != 0 && P.num_keys * 2 >= num_sms)
    P.items_per_cta = (P.items_per_cta + P.items_per_key - 1) / P.items_per_key * P.items_per_key;
  P.exclusive = (P.items_per_cta % P.items_per_key == 0) ? 1 : 0;
  if (B3D_ENV_FLAG("B3D_WG_NOEXCL")) P.exclusive = 0;
  if (!P.exclusive)
    B3D_CHECK_CUDA(cudaMemsetAsync(dwacc, 0, (size_t)(ks == 3 ? 27 : (nmapsY > 1 ? nmapsY : 1)) * Cin_pad * Cout_pad * 4, stream));
  const int grid2 = (P.num_items + P.items_per_cta - 1) / P.items_per_cta;
  P.dwacc = dwacc; P.err = err_flag;

  // guard: garbage M rows may read beyond the last slot; `guard` keeps those reads inside the allocation
  const size_t smem = (size_t)P.R * P.x_slot_bytes + 2 * (size_t)P.y_slot_bytes + 256 + (size_t)guard;
  B3D_REQUIRE(smem <= 227 * 1024, "wgrad: smem %zu too large (N%d D%d H%d W%d Cin%d Cout%d ks%d)", smem, N, D, H, W, Cin_use,
              Cout, ks);

  const int C8tot = Cin_use / 8;
  {
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    const uint64_t sW = (uint64_t)ldx * 2, sH = sW * W, sND = sH * H;
    if (P.plane_mode) {
      dims[0] = 8; dims[1] = W; dims[2] = H; dims[3] = C8tot; dims[4] = (uint64_t)N * D;
      strides[0] = sW; strides[1] = sH; strides[2] = 16; strides[3] = sND;
      box[0] = 8; box[1] = P.BW; box[2] = P.BH; box[3] = 1; box[4] = 1;
    } else {
      dims[0] = 8; dims[1] = W; dims[2] = C8tot; dims[3] = H; dims[4] = (uint64_t)N * D;
      strides[0] = sW; strides[1] = 16; strides[2] = sH; strides[3] = sND;
      box[0] = 8; box[1] = P.BW; box[2] = P.CI8; box[3] = P.BH; box[4] = 1;
    }
    int rc = b3d_encode_tmap_bf16(&P.tmX, x, 5, dims, strides, box);
    if (rc) return rc;
  }
  for (int m = 0; m < nmapsY; ++m) {
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    const int C8o = Cout_pad / 8;
    if (P.plane_mode) {
      dims[0] = 8; dims[1] = W; dims[2] = H; dims[3] = C8o; dims[4] = (uint64_t)N * D;
      strides[0] = yv[m].sW; strides[1] = yv[m].sH; strides[2] = 16; strides[3] = yv[m].sND;
      box[0] = 8; box[1] = P.BW; box[2] = H; box[3] = 1; box[4] = 1;
    } else {
      dims[0] = 8; dims[1] = W; dims[2] = C8o; dims[3] = H; dims[4] = (uint64_t)N * D;
      strides[0] = yv[m].sW; strides[1] = 16; strides[2] = yv[m].sH; strides[3] = yv[m].sND;
      box[0] = 8; box[1] = W; box[2] = P.CO8; box[3] = TH; box[4] = 1;
    }
    int rc = b3d_encode_tmap_bf16(&P.tmY[m], yv[m].base, 5, dims, strides, box);
    if (rc) return rc;
  }
  static const cudaError_t wg_attr =   // one-time, thread-safe
      cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  B3D_CHECK_CUDA(wg_attr);
  if (B3D_ENV_FLAG("B3D_VERBOSE"))
    fprintf(stderr,
            "[b3d] wgrad N%d D%d H%d W%d Cin%d Cout%d ks%d variant%d plane%d ci_blk%d G%d BN%d chains%d TH%d TZ%d keys%d "
            "items%d (per cta %d) smem%zu tmem%d\n",
            N, D, H, W, Cin_use, Cout, ks, P.variant, P.plane_mode, P.ci_blk, P.G, P.BN, P.nchains, P.TH, P.TZ, P.num_keys,
            P.num_items, P.items_per_cta, smem, P.tmem_cols);
  wgrad_kernel<<<grid2, WG_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

extern "C" {

// Conv3d weight gradient.  x: [N,D,H,W,*] bf16 pitch ldx, first Cin channels used; dy: [N,D,H,W,Cout_pad>=Cout] pitch lddy.
// dw: fp32 [Cout][Cin_real][ks^3] (reference layout).  ws: fp32 scratch of ks^3 * Cin_pad * Cout_pad floats.
int b3d_conv_wgrad(const void* x, long long ldx, const void* dy, long long lddy, float* dw, int accumulate, int N, int D,
                   int H, int W, int Cin, int Cin_real, int Cout, int ks, float* ws, size_t ws_bytes, int* err_flag,
                   void* stream) {
  B3D_REQUIRE(ks == 1 || ks == 3, "conv_wgrad: ks must be 1 or 3");
  B3D_REQUIRE(Cin % 16 == 0 || Cin % 8 == 0, "conv_wgrad: Cin must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntaps = ks * ks * ks;
  const int Cout_pad = (Cout + 15) / 16 * 16;
  B3D_REQUIRE(lddy >= Cout_pad || Cout_pad == Cout, "conv_wgrad: dy must expose %d (padded) channels", Cout_pad);
  const size_t need = (size_t)ntaps * Cin * Cout_pad * 4;
  B3D_REQUIRE(ws_bytes >= need, "conv_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  int n = N, d = D, h = H, w = W;
  if (ks == 1) {  // pointwise: flatten every voxel of the batch into rows of up to 128 positions
    const long long V = (long long)N * D * H * W;
    int ww = 128;
    while (ww > 16 && V % ww) ww /= 2;
    if (V % ww == 0) {
      n = 1; d = 1; w = ww; h = (int)(V / ww);
      if (h > 65536) {  // keep TMA coordinates small: fold rows into planes
        int dd = 1;
        while (h > 4096 && h % 2 == 0) { h /= 2; dd *= 2; }
        d = dd;
      }
    } else {  // tiny volumes (1^3 / 2^3 levels): one short K run in plane mode, zero padded to a multiple of 16
      B3D_REQUIRE(V <= 256, "conv_wgrad: voxel count %lld is neither a multiple of 16 nor <= 256", V);
      n = 1; d = 1; h = 1; w = (int)V;
    }
  }
  YView yv;
  yv.base = dy; yv.sW = lddy * 2; yv.sH = yv.sW * w; yv.sND = yv.sH * h;
  int rc = 1;
  if (ks == 3 && Cin % 16 == 0)  // swizzled MN-major kernel (conv_wg2.cu) when the rows are a multiple of 16 wide
    rc = b3d_try_wg2(x, ldx, Cin, dy, lddy, Cout_pad, N, D, H, W, ws, Cin, err_flag, st);
  else if (ks == 1)              // streaming kernel (conv_wgp.cu)
    rc = b3d_try_wgp(x, ldx, Cin, dy, lddy, Cout_pad, (long long)N * D * H * W, 1, 0, 0, 0, ws, Cin, err_flag, st);
  if (rc < 0) return rc;
  if (rc > 0) rc = run_wgrad(x, ldx, Cin, &yv, 1, n, d, h, w, Cout, ks, ws, Cin, Cout_pad, err_flag, st);
  if (rc) return rc;
  B3D_CHECK_CUDA(launch_wgrad_finalize(ws, dw, 0, ntaps, Cin_real, Cout, Cin, Cout_pad, accumulate, st));
  return B3D_OK;
}

// ConvTranspose3d(k=2,s=2) weight gradient: x [N,D,H,W,Cin] (coarse), dy [N,2D,2H,2W,Cout] -> dWt fp32 [Cin][Cout][8]
int b3d_convT2_wgrad(const void* x, long long ldx, const void* dy, long long lddy, float* dw, int accumulate, int N, int D,
                     int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int* err_flag, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int Cout_pad = (Cout + 15) / 16 * 16;
  B3D_REQUIRE(Cout_pad == Cout, "convT2_wgrad: Cout must be a multiple of 16");
  const size_t need = (size_t)8 * Cin * Cout_pad * 4;
  B3D_REQUIRE(ws_bytes >= need, "convT2_wgrad: workspace too small");
  YView yv[8];
  const long long pW = lddy * 2, pH = pW * (2 * W), pD = pH * (2 * H);
  for (int t8 = 0; t8 < 8; ++t8) {
    const int a = t8 >> 2, b = (t8 >> 1) & 1, c = t8 & 1;
    yv[t8].base = (const char*)dy + a * pD + b * pH + c * pW;
    yv[t8].sW = 2 * pW; yv[t8].sH = 2 * pH; yv[t8].sND = 2 * pD;
  }
  // the N*D planes of the coarse grid: plane (n,z) of view t8 starts at n*(2D)*pD + 2z*pD = (n*D+z)*2*pD  -> uniform stride
  int rc = b3d_try_wgp(x, ldx, Cin, dy, lddy, Cout_pad, (long long)N * D * H * W, 8, N * D, H, W, ws, Cin, err_flag, st);
  if (rc < 0) return rc;
  if (rc > 0) rc = run_wgrad(x, ldx, Cin, yv, 8, N, D, H, W, Cout, 1, ws, Cin, Cout_pad, err_flag, st);
  if (rc) return rc;
  B3D_CHECK_CUDA(launch_wgrad_finalize(ws, dw, 1, 8, Cin, Cout, Cin, Cout_pad, accumulate, st));
  return B3D_OK;
}

}  // extern "C"
