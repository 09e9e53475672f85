// conv_zmarch.cu — input-stationary ("z-marching") variant of the tcgen05 implicit-GEMM 3x3x3 convolution for the
// full-resolution levels (large planes, few output channels: main.py:216,219,130 at level 0 — ~50 % of all conv FLOPs).
//
// The block-mode kernel (conv_igemm.cu) stages a 3-plane halo box for every output plane tile, i.e. every activation
// byte crosses L2->SMEM three times (and the packed weights once per tile).  Here a CTA marches along z:
//   * the packed weights of ALL 27 taps stay resident in shared memory for the whole kernel,
//   * each input plane tile ([TH+2][W+2] voxels x KC channels, ONE swizzled TMA box) is loaded ONCE and contributes to the
//     three output planes z-1, z, z+1 (taps kd = 2,1,0),
//   * 3 output planes are "open" at any time, each with its own TMEM accumulator set; a 4th set drains in the epilogue warps
//     while the MMA thread keeps issuing (asets = 4 when MB*BN*4 <= 512 columns, else 3).
// Tap shifts are row shifts of the UMMA start address inside the swizzled tile (legal with base_offset 0: the swizzle XOR
// uses absolute shared-memory address bits — scripts/umma_shift_test.cu).
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

struct alignas(64) ZmParams {
  CUtensorMap tmA, tmW;
  int N, D, H, W, Cout;
  int TH, BH, BW, ZL, MB, BN, KC, k_chunks, RB, layout_type, row_mode, xblocks;
  int stages, asets, acc_stride, tmem_cols;
  int tiles_y, tiles_z, num_items;
  uint32_t a_stage_bytes, a_tx_bytes, w_chunk_bytes, w_tx_bytes;
  bf16* out; long long ld_out;
  const float* bias;
  double* stats; int cpg, stats_groups, stats_batch;
  int* err;
};

#define ZM_THREADS 192
#define ZM_MAXSTAGES 6

__device__ __forceinline__ int zm_mblock_base(const ZmParams& P, int mb) {
  if (!P.row_mode) return mb * 128;
  const int xb = mb % P.xblocks;
  const int py = mb / P.xblocks;
  return py * P.BW + xb * 128;
}

__global__ void __launch_bounds__(ZM_THREADS, 1) zmarch_kernel(const __grid_constant__ ZmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t sW = smem_u32(smem);
  const uint32_t sA = sW + P.k_chunks * P.w_chunk_bytes;
  uint8_t* aux = smem + (size_t)P.k_chunks * P.w_chunk_bytes + (size_t)S * P.a_stage_bytes;
  const uint32_t full0 = smem_u32(aux);                 // [ZM_MAXSTAGES]
  const uint32_t empty0 = full0 + 8 * ZM_MAXSTAGES;     // [ZM_MAXSTAGES]
  const uint32_t tfull0 = empty0 + 8 * ZM_MAXSTAGES;    // [4]
  const uint32_t tempty0 = tfull0 + 32;                 // [4]
  const uint32_t wfull = tempty0 + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * ZM_MAXSTAGES + 64 + 8);
  float* s_stats = reinterpret_cast<float*>(aux + 16 * ZM_MAXSTAGES + 64 + 16);  // [64]

  if (threadIdx.x == 0) {
    if (sW & 1023u) { if (P.err) atomicExch(P.err, 29); __trap(); }
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_stats[threadIdx.x - 64] = 0.f;
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer (warp-uniform control flow, elected lane issues) =======================
    if (elect_one()) {
      tma_prefetch_desc(&P.tmA);
      tma_prefetch_desc(&P.tmW);
      mbar_expect_tx(wfull, P.k_chunks * P.w_tx_bytes);
      for (int kc = 0; kc < P.k_chunks; ++kc) tma_load_3d(sW + kc * P.w_chunk_bytes, &P.tmW, wfull, kc * P.KC, 0, 0);
    }
    uint32_t s = 0, ph = 0;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
      int t = item;
      const int ty = t % P.tiles_y; t /= P.tiles_y;
      const int tz = t % P.tiles_z; const int n = t / P.tiles_z;
      const int z0 = tz * P.ZL, nz = min(P.ZL, P.D - z0), y0 = ty * P.TH;
      for (int zi = z0 - 1; zi <= z0 + nz; ++zi) {
        if (zi < 0 || zi >= P.D) continue;  // zero-padding planes contribute nothing: never staged, never multiplied
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(empty0 + 8 * s, ph ^ 1, P.err, 21);
          if (elect_one()) {
            const uint32_t fb = full0 + 8 * s;
            mbar_expect_tx(fb, P.a_tx_bytes);
            tma_load_5d(sA + s * P.a_stage_bytes, &P.tmA, fb, kc * P.KC, -1, y0 - 1, zi, n);
          }
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform control flow, elected lane issues) =======================
    {
      const uint32_t idesc = umma_idesc_bf16(128, P.BN, 0, 0);
      const uint64_t hi = umma_desc_hi_sw(8u * P.RB, P.layout_type);
      const uint32_t rb16 = (uint32_t)P.RB >> 4;
      const int nk16 = P.KC / 16;
      mbar_wait(wfull, 0, P.err, 22);
      tc_fence_after();
      uint32_t s = 0, ph = 0;
      uint32_t gbase = 0;  // running index of output planes (selects accumulator set and barrier phase)
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
        int t = item;
        t /= P.tiles_y;
        const int tz = t % P.tiles_z;
        const int z0 = tz * P.ZL, nz = min(P.ZL, P.D - z0);
        for (int zi = z0 - 1; zi <= z0 + nz; ++zi) {
          if (zi < 0 || zi >= P.D) continue;
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(full0 + 8 * s, ph, P.err, 23);
            tc_fence_after();
            const uint32_t a16 = (sA + s * P.a_stage_bytes) >> 4;
            const uint32_t w16 = (sW + kc * P.w_chunk_bytes) >> 4;
            for (int kd = 0; kd < 3; ++kd) {
              const int zo = zi - kd + 1;
              if (zo < z0 || zo >= z0 + nz) continue;
              const uint32_t g = gbase + (uint32_t)(zo - z0);
              const uint32_t set = g % (uint32_t)P.asets, use = g / (uint32_t)P.asets;
              const int zi_first = (zo - 1 >= 0) ? zo - 1 : zo;
              const int zi_last = (zo + 1 < P.D) ? zo + 1 : zo;
              const bool first = (zi == zi_first) && (kc == 0);
              if (first) {
                mbar_wait(tempty0 + 8 * set, (use & 1u) ^ 1u, P.err, 24);
                tc_fence_after();
              }
              const uint32_t dbase = tmem_base + set * P.acc_stride;
              const uint32_t wkd = w16 + (uint32_t)(kd * 9 * P.BN) * rb16;
              for (int mb = 0; mb < P.MB; ++mb) {
                const uint32_t d = dbase + mb * P.BN;
                const uint32_t amb = a16 + (uint32_t)zm_mblock_base(P, mb) * rb16;
                if (elect_one()) {
                  uint32_t acc = first ? 0u : 1u;
                  uint32_t arow = amb, brow = wkd;
                  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                      const uint32_t aoff = arow + kw * rb16;
                      const uint32_t boff = brow + (uint32_t)(kw * P.BN) * rb16;
#pragma unroll
                      for (int k16 = 0; k16 < 4; ++k16) {
                        if (k16 < nk16) {
                          const uint32_t alo = ((aoff + k16 * 2) & 0x3FFFu) | (1u << 16);
                          const uint32_t blo = ((boff + k16 * 2) & 0x3FFFu) | (1u << 16);
                          umma_bf16_ss(d, hi | alo, hi | blo, idesc, acc);
                          acc = 1u;
                        }
                      }
                    }
                    arow += (uint32_t)P.BW * rb16;
                    brow += (uint32_t)(3 * P.BN) * rb16;
                  }
                }
                __syncwarp();
              }
              if (zi == zi_last && kc == P.k_chunks - 1) {
                if (elect_one()) umma_commit(tfull0 + 8 * set);
              }
            }
            if (elect_one()) umma_commit(empty0 + 8 * s);
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
          }
        }
        gbase += (uint32_t)nz;
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int seglen = P.cpg < 16 ? P.cpg : 16;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
      int t = item;
      const int ty = t % P.tiles_y; t /= P.tiles_y;
      const int tz = t % P.tiles_z; const int n = t / P.tiles_z;
      const int z0 = tz * P.ZL, nz = min(P.ZL, P.D - z0), y0 = ty * P.TH;
      for (int zo = z0; zo < z0 + nz; ++zo, ++g) {
        const uint32_t set = g % (uint32_t)P.asets, use = g / (uint32_t)P.asets;
        mbar_wait(tfull0 + 8 * set, use & 1u, P.err, 25);
        tc_fence_after();
        for (int mb = 0; mb < P.MB; ++mb) {
          const int p = zm_mblock_base(P, mb) + row;
          const int py = p / P.BW, px = p - py * P.BW;
          const int y = y0 + py;
          const bool valid = (py < P.TH) && (px < P.W) && (y < P.H);
          const long long vox = (((long long)n * P.D + zo) * P.H + y) * P.W + px;
          const uint32_t trow = tmem_base + set * P.acc_stride + mb * P.BN + ((uint32_t)(q * 32) << 16);
          for (int j0 = 0; j0 < P.BN; j0 += 16) {
            uint32_t r[16];
            tmem_ld16(trow + j0, r);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float f = __uint_as_float(r[j]);
              if (P.bias != nullptr && j0 + j < P.Cout) f += __ldg(P.bias + j0 + j);
              v[j] = f;
            }
            if (P.stats != nullptr) {
              for (int sg = 0; sg < 16; sg += seglen) {
                float s1 = 0.f, s2 = 0.f;
                if (valid) {
                  for (int j = sg; j < sg + seglen; ++j)
                    if (j0 + j < P.Cout) { s1 += v[j]; s2 += v[j] * v[j]; }
                }
                s1 = warp_sum(s1); s2 = warp_sum(s2);
                if (lane == 0) {
                  const int gi = (j0 + sg) / P.cpg;
                  atomicAdd(&s_stats[2 * gi], s1);
                  atomicAdd(&s_stats[2 * gi + 1], s2);
                }
              }
            }
            if (valid) {
              bf16* o = P.out + vox * P.ld_out + j0;
              if (j0 + 8 <= P.Cout) stg16(o, pack8(v));
              if (j0 + 16 <= P.Cout) stg16(o + 8, pack8(v + 8));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * set);
        if (P.stats != nullptr) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int groups_blk = (P.BN - 1) / P.cpg + 1;
          if (et < 2 * groups_blk) {
            const int gi = et >> 1;
            if (gi < P.stats_groups) {
              const int ns = P.stats_batch ? 0 : n;
              atomicAdd(P.stats + ((long long)ns * P.stats_groups + gi) * 2 + (et & 1), (double)s_stats[et]);
            }
            s_stats[et] = 0.f;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, P.tmem_cols); }
}

static bool g_zm_attr = false;

// Returns B3D_OK if launched, 1 if this shape is not suited to z-marching (caller falls back to the block-mode kernel),
// negative on error.
int b3d_try_zmarch(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
                   int N, int D, int H, int W, int Cin, int Cout, double* stats, int cpg, int stats_groups,
                   int stats_batch, int* err_flag, cudaStream_t stream) {
  if (getenv("B3D_NO_ZMARCH")) return 1;
  const int CoutPad = w_rows;
  if (CoutPad > 64 || W + 2 > 256 || D < 4 || H < 4) return 1;
  const int num_sms = b3d_num_sms();
  const int BN = CoutPad;
  if (stats && (cpg <= 0 || (BN / cpg > 32) || !(cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8 || cpg % 16 == 0))) return 1;
  const int BW = W + 2;
  const bool row_mode = (W % 128 == 0);
  const int budget = 227 * 1024 - 2048;
  int bTH = 0, bKC = 0, bS = 0, bAsets = 0, bMB = 0;
  double best = -1;
  for (int TH = 16; TH >= 1; TH /= 2) {
    if (TH > H) continue;
    const int BH = TH + 2;
    const int span = (TH - 1) * BW + W;
    const int MB = row_mode ? TH * (W / 128) : (span + 127) / 128;
    int asets = 0;
    if (MB * BN * 4 <= 512) asets = 4; else if (MB * BN * 3 <= 512) asets = 3; else continue;
    for (int KC = 64; KC >= 16; KC /= 2) {
      if (Cin % KC) continue;
      const int RB = KC * 2;
      const long long wch = ((long long)27 * BN * RB + 1023) / 1024 * 1024;
      const long long ast = ((long long)BH * BW * RB + 1023) / 1024 * 1024;
      const long long wtot = wch * (Cin / KC);
      // garbage M rows of the last block may read up to (MB*128 + 2*BW + 2) rows: keep that inside the A ring
      const long long maxrow = (row_mode ? span : (long long)MB * 128) + 2 * BW + 2;
      const long long over = std::max<long long>(0, maxrow - (long long)BH * BW) * RB;
      int S = (int)std::min<long long>(ZM_MAXSTAGES, (budget - wtot - over) / ast);
      if (S < 2) continue;
      const double eff = (double)(TH * W) / (MB * 128.0) * ((double)TH / BH);  // M-row efficiency x y-halo efficiency
      const double score = eff * (asets == 4 ? 1.0 : 0.9) * (S >= 3 ? 1.0 : 0.9) + 0.001 * KC / 64.0;
      if (score > best) { best = score; bTH = TH; bKC = KC; bS = S; bAsets = asets; bMB = MB; }
    }
  }
  if (best < 0) return 1;
  ZmParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.D = D; P.H = H; P.W = W; P.Cout = Cout;
  P.TH = bTH; P.BH = bTH + 2; P.BW = BW; P.MB = bMB; P.BN = BN; P.KC = bKC; P.k_chunks = Cin / bKC;
  P.RB = bKC * 2; P.layout_type = (P.RB == 128) ? 2 : (P.RB == 64 ? 4 : 6);
  P.row_mode = row_mode ? 1 : 0; P.xblocks = row_mode ? W / 128 : 1;
  P.stages = bS; P.asets = bAsets; P.acc_stride = bMB * BN;
  int cols = P.asets * P.acc_stride, tc = 32;
  while (tc < cols) tc *= 2;
  P.tmem_cols = tc;
  P.tiles_y = (H + bTH - 1) / bTH;
  int ZL = 32;
  while (ZL > 4 && (long long)N * P.tiles_y * ((D + ZL - 1) / ZL) < 3LL * num_sms) ZL /= 2;
  if (ZL > D) ZL = D;
  P.ZL = ZL; P.tiles_z = (D + ZL - 1) / ZL;
  P.num_items = N * P.tiles_y * P.tiles_z;
  P.a_tx_bytes = (uint32_t)P.BH * BW * P.RB; P.a_stage_bytes = (P.a_tx_bytes + 1023u) / 1024u * 1024u;
  P.w_tx_bytes = (uint32_t)27 * BN * P.RB; P.w_chunk_bytes = (P.w_tx_bytes + 1023u) / 1024u * 1024u;
  P.out = (bf16*)y; P.ld_out = ldy; P.bias = bias;
  P.stats = stats; P.cpg = cpg > 0 ? cpg : 16; P.stats_groups = stats_groups; P.stats_batch = stats_batch; P.err = err_flag;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t sW = (uint64_t)ldx * 2;
    uint64_t strides[4] = {sW, sW * W, sW * W * H, sW * W * H * D};
    uint32_t box[5] = {(uint32_t)P.KC, (uint32_t)BW, (uint32_t)P.BH, 1, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmA, x, 5, dims, strides, box, P.RB);
    if (rc) return rc;
    uint64_t wd[3] = {(uint64_t)Cin, (uint64_t)CoutPad, 27};
    uint64_t ws[2] = {(uint64_t)Cin * 2, (uint64_t)CoutPad * Cin * 2};
    uint32_t wb[3] = {(uint32_t)P.KC, (uint32_t)BN, 27};
    rc = b3d_encode_tmap_bf16(&P.tmW, wpack, 3, wd, ws, wb, P.RB);
    if (rc) return rc;
  }
  const long long maxrow = (row_mode ? (long long)(bTH - 1) * BW + W : (long long)bMB * 128) + 2 * BW + 2;
  const long long over = std::max<long long>(0, maxrow - (long long)P.BH * BW) * P.RB;
  const size_t smem = (size_t)P.k_chunks * P.w_chunk_bytes + (size_t)P.stages * P.a_stage_bytes + (size_t)over +
                      16 * ZM_MAXSTAGES + 64 + 16 + 64 * 4 + 1024;
  if (smem > 227 * 1024) return 1;
  if (!g_zm_attr) {
    B3D_CHECK_CUDA(cudaFuncSetAttribute(zmarch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_zm_attr = true;
  }
  const int grid = std::min(P.num_items, num_sms);
  if (getenv("B3D_VERBOSE"))
    fprintf(stderr, "[b3d] zmarch N%d D%d H%d W%d Cin%d Cout%d TH%d KC%d BN%d MB%d stages%d asets%d ZL%d items%d smem%zu tmem%d\n",
            N, D, H, W, Cin, Cout, P.TH, P.KC, P.BN, P.MB, P.stages, P.asets, P.ZL, P.num_items, smem, P.tmem_cols);
  zmarch_kernel<<<grid, ZM_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}
