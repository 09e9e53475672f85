// conv_wgp.cu — weight gradient of the 1x1x1 convolutions (and, per tap, of ConvTranspose3d k2 s2) as a STREAMING kernel
// on tcgen05 tensor cores (sm_100a).
//
//   dW[ci][co] = sum_v X[v][ci] * dY[v][co]                v = every voxel of the batch
// (autograd's convolution_backward for the residual / attention-gate nn.Conv3d(k=1) of /root/reference/main.py:229,252,258
//  and, with dY read through a stride-2 view per tap, for nn.ConvTranspose3d(k=2,s=2), main.py:121.)
//
// Roofline: HBM.  Per voxel the kernel must read (Cin + Cout) * 2 bytes and does 2 * Cin * Cout FLOP — 21 FLOP/B at
// (64, 32) — so the only job is to keep TMA loads in flight; the tensor pipe idles at ~30 %.
// GEMM view per key = (block of <= 128 output channels) x (block of <= 256 input channels): D[M = co][N = ci] += A * B with
// K = voxels.  Both operands are MN-major for the UMMA (in NDHWC the contracted index, the voxel, is the slow one):
//   * a staged tile is [channel block][KT voxels][GW channels] written by ONE swizzled TMA box of the 3-D view
//     {GW channels, V voxels (pitch ld), C/GW blocks (stride GW*2 B)}; GW = 32 (64-byte rows, SWIZZLE_64B) or 16 (32 B);
//   * the channel blocks are the descriptor's MN groups (LBO = KT rows), 8 voxels are one K group (SBO = 8 rows), and a
//     K = 16 step advances the start address by 16 rows — the scheme verified in scripts/umma_mn_test.cu and used by
//     conv_wg2.cu, without the tap shifts;
//   * M is always 128: groups beyond the real channel blocks read whatever follows the tile (inside the allocation) and
//     produce accumulator rows that are never read back.
// Tiles past the end of the volume are zero-filled by TMA and contribute nothing, so V needs no padding.
// Work = keys x K tiles, split evenly over the CTAs; accumulators persist in TMEM over a CTA's tiles of one key and are
// flushed with fp32 reductions into dwacc[ci][co] (wgrad_finalize_kernel in conv_wgrad.cu permutes to the reference layout).
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define WGP_THREADS 192   // warp 0 TMA, warp 1 MMA issuer, warps 2-5 flush
#define WGP_MAXS 8

struct alignas(64) WgpParams {
  CUtensorMap tmX;
  CUtensorMap tmY[8];        // pointwise: [0] only; transposed conv: one stride-2 view of dY per tap
  int KT, S;                 // voxels per stage, stages
  int NB;                    // N of the MMA = input channels per key
  int n_cib, n_cob;          // key grid
  int ydims;                 // rank of the dY map: 3 (pointwise) or 5 (transposed conv, stride-2 view)
  int TR, Wc, Hc;            // transposed conv: a K tile is TR rows of Wc coarse voxels (TR divides Hc)
  uint32_t x_tile, y_tile, x_tx, y_tx;
  long long tiles;           // K tiles per key
  long long total_steps;     // keys * tiles
  float* dwacc; int Cin, Cin_pad, Cout_pad;
  int tmem_cols;
  int* err;
};

template <int GWX, int GWY>
__global__ void __launch_bounds__(WGP_THREADS, 1) wgp_kernel(const __grid_constant__ WgpParams P) {
  constexpr int RBX = GWX * 2, RBY = GWY * 2;
  constexpr uint32_t rx16 = RBX / 16, ry16 = RBY / 16;
  constexpr int LTX = (RBX == 64) ? 4 : 6, LTY = (RBY == 64) ? 4 : 6;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = (uint32_t)P.S;
  const uint32_t sY = smem_u32(smem);
  const uint32_t sX = sY + S * P.y_tile;
  // aux sits after the tiles and the guard gap that keeps the unused M groups' reads inside the allocation
  const uint32_t te_a = S * (P.y_tile + P.x_tile), te_b = (S - 1) * P.y_tile + 256u * (uint32_t)P.KT;
  const uint32_t tiles_end = te_a > te_b ? te_a : te_b;
  uint8_t* aux = smem + ((tiles_end + 1023u) & ~1023u);
  const uint32_t full0 = smem_u32(aux);          // [8]
  const uint32_t empty0 = full0 + 64;            // [8]
  const uint32_t accfull = empty0 + 64;
  const uint32_t accempty = accfull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 160);

  if (threadIdx.x == 0) {
    if (sY & 1023u) { if (P.err) atomicExch(P.err, 49); __trap(); }
    for (int i = 0; i < WGP_MAXS; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(accfull, 1); mbar_init(accempty, 4);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long T = P.total_steps;
  const long long lo = T * blockIdx.x / gridDim.x, hi = T * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) { tma_prefetch_desc(&P.tmX); tma_prefetch_desc(&P.tmY[0]); }
    uint32_t q = 0;
    for (long long pos = lo; pos < hi; ++pos, ++q) {
      const int key = (int)(pos / P.tiles);
      const long long tile = pos - (long long)key * P.tiles;
      const int cob = key % P.n_cob, cib = key / P.n_cob;   // for the transposed conv "cob" also enumerates the 8 taps
      const uint32_t slot = q % S, ph = (q / S) & 1u;
      mbar_wait(empty0 + 8 * slot, ph ^ 1, P.err, 41);
      if (elect_one()) {
        const uint32_t fb = full0 + 8 * slot;
        mbar_expect_tx(fb, P.x_tx + P.y_tx);
        tma_load_3d(sX + slot * P.x_tile, &P.tmX, fb, 0, (int)(tile * P.KT), cib * (P.NB / GWX));
        if (P.ydims == 3) {
          tma_load_3d(sY + slot * P.y_tile, &P.tmY[0], fb, 0, (int)(tile * P.KT), cob * (128 / GWY));
        } else {
          // transposed conv: key's "cob" = tap * blocks + channel block; the tile is TR rows of one coarse plane
          const int nblk = P.Cout_pad / GWY > 128 / GWY ? 128 / GWY : P.Cout_pad / GWY;
          const int ncb = (P.Cout_pad / GWY + nblk - 1) / nblk;
          const int t8 = cob / ncb, cb = cob - t8 * ncb;
          const int rows_per_plane = P.Hc / P.TR;
          const long long plane = tile / rows_per_plane;
          const int r0 = (int)(tile - plane * rows_per_plane) * P.TR;
          // per-tap 5-D view {GWY, Wc (stride 2 voxels), Hc (stride 2 rows), N*Dc (stride 2 planes), channel blocks}
          tma_load_5d(sY + slot * P.y_tile, &P.tmY[t8], fb, 0, 0, r0, (int)plane, cb * nblk);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform control flow, one elected lane issues) =======================
    const uint32_t idesc = umma_idesc_bf16(128, P.NB, 1, 1);
    const uint64_t hiA = umma_desc_hi_sw(8u * RBY, LTY);
    const uint64_t hiB = umma_desc_hi_sw(8u * RBX, LTX);
    const uint32_t LBOY = ((((uint32_t)P.KT * RBY) >> 4) & 0x3FFFu) << 16;   // channel blocks: one tile of KT rows apart
    const uint32_t LBOX = ((((uint32_t)P.KT * RBX) >> 4) & 0x3FFFu) << 16;
    const int ksteps = P.KT / 16;
    uint32_t q = 0, flushes = 0, acc = 0;
    int cur_key = -1;
    for (long long pos = lo; pos < hi; ++pos, ++q) {
      const int key = (int)(pos / P.tiles);
      if (key != cur_key) {
        if (cur_key >= 0) {
          if (elect_one()) umma_commit(accfull);
          __syncwarp();
          mbar_wait(accempty, flushes & 1u, P.err, 43);
          tc_fence_after();
          ++flushes;
        }
        cur_key = key; acc = 0;
      }
      const uint32_t slot = q % S, ph = (q / S) & 1u;
      mbar_wait(full0 + 8 * slot, ph, P.err, 42);
      tc_fence_after();
      const uint32_t ya = ((sY + slot * P.y_tile) >> 4) | LBOY;
      const uint32_t xa = ((sX + slot * P.x_tile) >> 4) | LBOX;
      if (elect_one()) {
        uint32_t c = acc;
#pragma unroll 4
        for (int j = 0; j < ksteps; ++j) {
          umma_bf16_ss(tmem_base, hiA | (ya + (uint32_t)j * 16u * ry16), hiB | (xa + (uint32_t)j * 16u * rx16), idesc, c);
          c = 1u;
        }
        umma_commit(empty0 + 8 * slot);
      }
      __syncwarp();
      acc = 1u;
    }
    if (cur_key >= 0) {
      if (elect_one()) umma_commit(accfull);
      __syncwarp();
    }
  } else {
    // ======================= flush: TMEM accumulators of a finished key -> fp32 reductions into dwacc =======================
    const int qd = warp & 3;
    const int m = qd * 32 + lane;          // accumulator row = output channel within the key's 128-channel block
    uint32_t flushes = 0;
    int cur_key = -1;
    long long pos = lo;
    while (true) {
      const int key = (pos < hi) ? (int)(pos / P.tiles) : -2;
      if (key != cur_key && cur_key >= 0) {
        mbar_wait(accfull, flushes & 1u, P.err, 44);
        tc_fence_after();
        const int cobk = cur_key % P.n_cob, cib = cur_key / P.n_cob;
        int tap = 0, co;
        if (P.ydims == 3) {
          co = cobk * 128 + m;
        } else {
          const int nblk = P.Cout_pad / GWY > 128 / GWY ? 128 / GWY : P.Cout_pad / GWY;
          const int ncb = (P.Cout_pad / GWY + nblk - 1) / nblk;
          tap = cobk / ncb;
          co = (cobk - tap * ncb) * nblk * GWY + m;
          if (m >= nblk * GWY) co = P.Cout_pad;   // rows of the unused M groups
        }
        const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
        for (int j0 = 0; j0 < P.NB; j0 += 16) {
          uint32_t rr[16];
          tmem_ld16(trow + j0, rr);
          tmem_ld_wait();
          const int ci0 = cib * P.NB + j0;
          if (co < P.Cout_pad && ci0 < P.Cin) {
            float* dst = P.dwacc + ((long long)tap * P.Cin_pad + ci0) * P.Cout_pad + co;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + (long long)j * P.Cout_pad, __uint_as_float(rr[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty);
        ++flushes;
      }
      if (pos >= hi) break;
      cur_key = key;
      // jump to the first step of the next key (or the end of this CTA's range)
      const long long next = ((long long)key + 1) * P.tiles;
      pos = next < hi ? next : hi;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

template <int GWX, int GWY>
static int wgp_launch(const WgpParams& P, size_t smem, int grid, cudaStream_t stream) {
  static const cudaError_t attr = cudaFuncSetAttribute(wgp_kernel<GWX, GWY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // one-time, thread-safe
  B3D_CHECK_CUDA(attr);
  wgp_kernel<GWX, GWY><<<grid, WGP_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// Returns B3D_OK if launched, 1 if the shape is not suited (caller uses the generic kernel), negative on error.
// Pointwise (taps == 1): x [V][Cin] pitch ldx, dy [V][Cout_pad] pitch lddy, dwacc fp32 [Cin_pad][Cout_pad].
// Transposed conv (taps == 8): x coarse [N*Dc*Hc*Wc][Cin], dy fine [N,2Dc,2Hc,2Wc][Cout_pad], dwacc [8][Cin_pad][Cout_pad],
//   tap t8 = (a,b,c) reads dy[n, 2z+a, 2y+b, 2x+c].
int b3d_try_wgp(const void* x, long long ldx, int Cin, const void* dy, long long lddy, int Cout_pad, long long V, int taps,
                int ND, int Hc, int Wc, float* dwacc, int Cin_pad, int* err_flag, cudaStream_t stream) {
  if (B3D_ENV_FLAG("B3D_NO_WGP")) return 1;
  if (Cin % 16 || Cout_pad % 16 || V < 16 || (ldx * 2) % 16 || (lddy * 2) % 16) return 1;
  if (taps != 1 && taps != 8) return 1;
  const int GWX = (Cin % 32 == 0) ? 32 : 16, GWY = (Cout_pad % 32 == 0) ? 32 : 16;
  const int RBX = GWX * 2, RBY = GWY * 2;
  WgpParams P;
  memset(&P, 0, sizeof(P));
  const int nxb_total = Cin / GWX, nyb_total = Cout_pad / GWY;
  const int nxb = std::min(nxb_total, 256 / GWX), nyb = std::min(nyb_total, 128 / GWY);
  P.NB = nxb * GWX;
  P.n_cib = (nxb_total + nxb - 1) / nxb;
  const int ncb = (nyb_total + nyb - 1) / nyb;
  P.n_cob = ncb * taps;
  P.ydims = (taps == 1) ? 3 : 5;
  // K tile: as many voxels as fit ~48 KB per stage (<= 256, the TMA box limit), 4 stages
  const int bytes_per_voxel = nxb * RBX + nyb * RBY;
  int KT = 256;
  while (KT > 16 && (long long)KT * bytes_per_voxel > 48 * 1024) KT /= 2;
  if (taps == 8) {
    // a tile is TR whole rows of Wc coarse voxels; TR must divide Hc and TR * Wc must be a multiple of 16
    if (Wc > 256 || (long long)ND * Hc * Wc != V) return 1;
    int TR = std::max(1, KT / Wc);
    while (TR > 1 && Hc % TR) --TR;
    if ((TR * Wc) % 16 || (long long)TR * Wc * bytes_per_voxel > 56 * 1024 || TR > 256) return 1;
    KT = TR * Wc;
    P.TR = TR; P.Wc = Wc; P.Hc = Hc;
    P.tiles = (long long)ND * (Hc / TR);
  } else {
    while (KT > 16 && KT / 2 >= V) KT /= 2;
    P.tiles = (V + KT - 1) / KT;
  }
  P.KT = KT;
  P.x_tile = (uint32_t)(((size_t)nxb * KT * RBX + 1023) / 1024 * 1024);
  P.y_tile = (uint32_t)(((size_t)nyb * KT * RBY + 1023) / 1024 * 1024);
  // tiles must be exactly [block][KT][GW] with no padding between blocks for the LBO arithmetic: KT*RB is a multiple of 512
  P.x_tx = (uint32_t)((size_t)nxb * KT * RBX); P.y_tx = (uint32_t)((size_t)nyb * KT * RBY);
  const size_t budget = (size_t)227 * 1024 - 1024 - 256 - 1024;
  int S = WGP_MAXS;
  auto need = [&](int s) {
    const size_t a = (size_t)s * (P.x_tile + P.y_tile), b = (size_t)(s - 1) * P.y_tile + (size_t)256 * KT;
    return (std::max(a, b) + 1023) / 1024 * 1024;
  };
  while (S > 2 && need(S) > budget) --S;
  if (need(S) > budget) return 1;
  P.S = S;
  P.total_steps = (long long)P.n_cib * P.n_cob * P.tiles;
  P.dwacc = dwacc; P.Cin = Cin; P.Cin_pad = Cin_pad; P.Cout_pad = Cout_pad; P.err = err_flag;
  { int t = 32; while (t < P.NB) t *= 2; P.tmem_cols = t; }
  {
    uint64_t dims[3] = {(uint64_t)GWX, (uint64_t)V, (uint64_t)nxb_total};
    uint64_t strides[2] = {(uint64_t)ldx * 2, (uint64_t)RBX};
    uint32_t box[3] = {(uint32_t)GWX, (uint32_t)KT, (uint32_t)nxb};
    int rc = b3d_encode_tmap_bf16(&P.tmX, x, 3, dims, strides, box, RBX);
    if (rc) return rc;
  }
  if (taps == 1) {
    uint64_t dims[3] = {(uint64_t)GWY, (uint64_t)V, (uint64_t)nyb_total};
    uint64_t strides[2] = {(uint64_t)lddy * 2, (uint64_t)RBY};
    uint32_t box[3] = {(uint32_t)GWY, (uint32_t)KT, (uint32_t)nyb};
    int rc = b3d_encode_tmap_bf16(&P.tmY[0], dy, 3, dims, strides, box, RBY);
    if (rc) return rc;
  } else {
    // fine grid [N*2Dc][2Hc][2Wc][Cout_pad]: tap (a,b,c) is the stride-2 view starting at plane a, row b, voxel c
    const uint64_t pW = (uint64_t)lddy * 2, pH = pW * (2 * Wc), pD = pH * (2 * Hc);
    for (int t8 = 0; t8 < 8; ++t8) {
      const int a = t8 >> 2, b = (t8 >> 1) & 1, c = t8 & 1;
      uint64_t dims[5] = {(uint64_t)GWY, (uint64_t)Wc, (uint64_t)Hc, (uint64_t)ND, (uint64_t)nyb_total};
      uint64_t strides[4] = {2 * pW, 2 * pH, 2 * pD, (uint64_t)RBY};
      uint32_t box[5] = {(uint32_t)GWY, (uint32_t)Wc, (uint32_t)P.TR, 1, (uint32_t)nyb};
      int rc = b3d_encode_tmap_bf16(&P.tmY[t8], (const char*)dy + a * pD + b * pH + c * pW, 5, dims, strides, box, RBY);
      if (rc) return rc;
    }
  }
  const size_t smem = need(S) + 256 + 1024;
  B3D_CHECK_CUDA(cudaMemsetAsync(dwacc, 0, (size_t)taps * Cin_pad * Cout_pad * 4, stream));
  const int num_sms = b3d_num_sms();
  // at least ~4 K tiles per CTA so that a flush (up to 128 x 256 fp32 reductions) is amortised
  const int grid = (int)std::max<long long>(1, std::min<long long>(num_sms, P.total_steps / 4));
  if (B3D_ENV_FLAG("B3D_VERBOSE"))
    fprintf(stderr, "[b3d] wgp V%lld Cin%d Cout%d GWX%d GWY%d NB%d KT%d S%d keys%d tiles%lld grid%d smem%zu\n", V, Cin, Cout_pad,
            GWX, GWY, P.NB, KT, S, P.n_cib * P.n_cob, P.tiles, grid, smem);
  if (GWX == 32 && GWY == 32) return wgp_launch<32, 32>(P, smem, grid, stream);
  if (GWX == 32 && GWY == 16) return wgp_launch<32, 16>(P, smem, grid, stream);
  if (GWX == 16 && GWY == 32) return wgp_launch<16, 32>(P, smem, grid, stream);
  return wgp_launch<16, 16>(P, smem, grid, stream);
}
