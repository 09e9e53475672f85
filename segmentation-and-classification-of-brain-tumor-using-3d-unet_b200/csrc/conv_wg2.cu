// conv_wg2.cu — 3x3x3 weight gradient on tcgen05 tensor cores, swizzled MN-major operands (sm_100a).
//
//   dW[kd,kh,kw][ci][co] = sum_{n,z,y,x} X[n, z+kd-1, y+kh-1, x+kw-1][ci] * dY[n,z,y,x][co]
// (autograd's convolution_backward for nn.Conv3d(k=3,pad=1), /root/reference/main.py:130,216,219).
//
// GEMM view per key = (32- or 16-channel block of ci) x (block of co):  D[M][N] += A[M][K] * B[K][N], K = voxels.
// Both operands are MN-major for the UMMA (in NDHWC the contracted index, the voxel, is the slow one): a staged tile is
// simply [position][channel block] rows written by ONE swizzled TMA box, K = 16 consecutive x positions of one row.
//   * A = dY tile (x-halo of 1, zero filled): M = 4 (8) groups of 32 (16) channels, group g read from the SAME tile shifted
//     by g positions (descriptor LBO = one row): groups 0..2 are the three kw taps (kw = 2 - g), the rest is never read back.
//   * B = X plane tile (y-halo of 1): N = 3 groups of channels, group h shifted by h tile rows (LBO = W rows): the kh taps.
//   * the three kd taps are three accumulator chains in TMEM fed from a 4-deep ring of X planes (each plane is loaded
//     once per column and used for the output planes z-1, z, z+1).
// One MMA (M128 x N96 x K16, 56 clk by the SMEM operand-fetch law measured in scripts/umma_rate.cu) therefore produces 9
// taps x 32 x 32 products: 64 % of the tensor peak is the bound of this formulation (75 % M rows useful x 86 % fetch bound).
// scripts/umma_mn_test.cu verified the descriptor semantics used here (swizzled MN-major: LBO = group stride, SBO = 8-row
// K-group stride, arbitrary start row).
// Accumulators persist in TMEM over all tiles a CTA processes for a key and are flushed once with fp32 reductions into
// dwacc[tap][ci][co] (wgrad_finalize_kernel in conv_wgrad.cu permutes to the reference layout).
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define WG2_THREADS 224   // warp 0 TMA, warps 1 and 6 MMA issuers (ping-pong over output planes), warps 2-5 flush
#define WG2_MAXR 8    // X plane ring slots (3 live + prefetch)
#define WG2_MAXRY 4   // dY plane ring slots

struct alignas(64) Wg2Params {
  CUtensorMap tmX, tmY;
  int N, D, H, W;
  int TH, TW, tiles_y, tiles_x, columns;   // columns = N * tiles_y * tiles_x
  int R, RY;                  // ring depths
  int n_cib, n_cob;
  uint32_t x_slot, y_slot, x_tx, y_tx;
  long long total_steps;      // nkeys * columns * D
  float* dwacc; int Cin_pad, Cout_pad;
  int* err;
};

struct Wg2Seg { int key, col, za, zb; };

__device__ __forceinline__ bool wg2_next_seg(long long& pos, long long hi, int D, int columns, Wg2Seg& s) {
  if (pos >= hi) return false;
  const long long c = pos / D;
  s.key = (int)(c / columns);
  s.col = (int)(c - (long long)s.key * columns);
  s.za = (int)(pos - c * D);
  const long long rem = hi - pos;
  s.zb = (int)((rem < (long long)(D - s.za)) ? s.za + rem : D);
  pos += s.zb - s.za;
  return true;
}

template <int GWX, int GWY, int KSTEPS>
__global__ void __launch_bounds__(WG2_THREADS, 1) wg2_kernel(const __grid_constant__ Wg2Params P) {
  constexpr int RBX = GWX * 2, RBY = GWY * 2;
  constexpr uint32_t rx16 = RBX / 16, ry16 = RBY / 16;
  constexpr int LTX = (RBX == 64) ? 4 : 6, LTY = (RBY == 64) ? 4 : 6;
  constexpr int NB = 3 * GWX;            // N of one MMA: 3 kh groups
  constexpr int TCOLS = (3 * NB <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sX = smem_u32(smem);
  const uint32_t R = (uint32_t)P.R, RY = (uint32_t)P.RY;
  const uint32_t sY = sX + R * P.x_slot;
  uint8_t* aux = smem + (size_t)R * P.x_slot + (size_t)RY * P.y_slot + 1024;  // 1 KB guard: the unused M groups read past the slot
  const uint32_t xfull0 = smem_u32(aux);        // [8]
  const uint32_t xempty0 = xfull0 + 64;         // [8]
  const uint32_t yfull0 = xempty0 + 64;         // [4]
  const uint32_t yempty0 = yfull0 + 32;         // [4]
  const uint32_t accfull = yempty0 + 32;
  const uint32_t accempty = accfull + 8;
  const uint32_t hs0 = accempty + 8;            // [2] "my plane is in the tensor queue" hand-shake of the two issuer warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 224);
  volatile uint32_t* s_started = reinterpret_cast<volatile uint32_t*>(aux + 232);

  if (threadIdx.x == 0) {
    if (sX & 1023u) { if (P.err) atomicExch(P.err, 39); __trap(); }
    for (int i = 0; i < WG2_MAXR; ++i) { mbar_init(xfull0 + 8 * i, 1); mbar_init(xempty0 + 8 * i, 1); }
    for (int i = 0; i < WG2_MAXRY; ++i) { mbar_init(yfull0 + 8 * i, 1); mbar_init(yempty0 + 8 * i, 1); }
    mbar_init(accfull, 1); mbar_init(accempty, 4);
    mbar_init(hs0, 1); mbar_init(hs0 + 8, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long T = P.total_steps;
  const long long lo = T * blockIdx.x / gridDim.x, hi = T * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) { tma_prefetch_desc(&P.tmX); tma_prefetch_desc(&P.tmY); }
    uint32_t Q = 0, Tn = 0;  // running X-plane / dY-plane fill counters
    long long pos = lo;
    Wg2Seg sg;
    while (wg2_next_seg(pos, hi, P.D, P.columns, sg)) {
      const int cob = sg.key % P.n_cob, cib = sg.key / P.n_cob;
      int tcol = sg.col;
      const int tx = tcol % P.tiles_x; tcol /= P.tiles_x;
      const int ty = tcol % P.tiles_y; const int n = tcol / P.tiles_y;
      const int y0 = ty * P.TH, x0 = tx * P.TW;
      const int L = sg.zb - sg.za;
      for (int i = 0; i < L + 2; ++i) {
        {  // X plane i  (z = za - 1 + i)
          const uint32_t slot = Q % R, ph = (Q / R) & 1u;
          mbar_wait(xempty0 + 8 * slot, ph ^ 1, P.err, 31);
          const int zi = sg.za - 1 + i;
          if (elect_one()) {
            const uint32_t fb = xfull0 + 8 * slot;
            if (zi >= 0 && zi < P.D) {
              mbar_expect_tx(fb, P.x_tx);
              tma_load_5d(sX + slot * P.x_slot, &P.tmX, fb, cib * GWX, x0, y0 - 1, zi, n);
            } else {
              mbar_arrive(fb);  // plane outside the volume: nothing to load, the consumer skips it
            }
          }
          __syncwarp();
          ++Q;
        }
        if (i >= 2) {  // dY plane z = za + i - 2
          const uint32_t slot = Tn % RY, ph = (Tn / RY) & 1u;
          mbar_wait(yempty0 + 8 * slot, ph ^ 1, P.err, 32);
          if (elect_one()) {
            const uint32_t fb = yfull0 + 8 * slot;
            mbar_expect_tx(fb, P.y_tx);
            tma_load_5d(sY + slot * P.y_slot, &P.tmY, fb, cob * GWY, x0 - 1, y0, sg.za + i - 2, n);
          }
          __syncwarp();
          ++Tn;
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ======================= MMA issuers (warp-uniform control flow, one elected lane issues) =======================
    const uint32_t idesc = umma_idesc_bf16(128, NB, 1, 1);
    // MN-major swizzled descriptors: LBO = stride between channel groups (= tap shift), SBO = 8 K rows
    const uint64_t hiA = umma_desc_hi_sw(8u * RBY, LTY);
    const uint64_t hiB = umma_desc_hi_sw(8u * RBX, LTX);
    const uint32_t LBOY = ((uint32_t)RBY >> 4) << 16;                        // kw groups: 1 row apart
    const uint32_t LBOX = ((((uint32_t)P.TW * RBX) >> 4) & 0x3FFFu) << 16;   // kh groups: one tile row apart
    const uint32_t BWY = (uint32_t)P.TW + 2;
    const uint32_t role = (warp == 1) ? 0u : 1u;
    uint32_t Q0 = 0, Tn = 0, flushes = 0;
    uint32_t pc = 0;                         // running output-plane count: planes alternate between the two issuer warps
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0;   // per-chain "accumulate" flags (tracked identically by both warps)
    int cur_key = -1;
    long long pos = lo;
    Wg2Seg sg;
    while (wg2_next_seg(pos, hi, P.D, P.columns, sg)) {
      if (sg.key != cur_key) {
        if (cur_key >= 0) {
          // the warp that issued the last plane commits; the tensor pipe is in-order, so the other warp's MMAs are done too
          if (((pc - 1u) & 1u) == role && elect_one()) {
            *s_started = acc0 | (acc1 << 1) | (acc2 << 2);
            __threadfence_block();
            umma_commit(accfull);
          }
          __syncwarp();
          mbar_wait(accempty, flushes & 1u, P.err, 33);
          tc_fence_after();
          ++flushes;
        }
        cur_key = sg.key; acc0 = acc1 = acc2 = 0;
      }
      const int ty = (sg.col / P.tiles_x) % P.tiles_y;
      const int rows = min(P.TH, P.H - ty * P.TH);
      const int L = sg.zb - sg.za;
      for (int t = 0; t < L; ++t, ++pc, ++Tn) {
        const bool h0 = (sg.za - 1 + t >= 0), h1 = true, h2 = (sg.za + 1 + t < P.D);
        if ((pc & 1u) == role) {
          const uint32_t yslot = Tn % RY, yph = (Tn / RY) & 1u;
          mbar_wait(yfull0 + 8 * yslot, yph, P.err, 34);
          uint32_t xb[3];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const uint32_t q = Q0 + t + kd;
            const uint32_t slot = q % R;
            mbar_wait(xfull0 + 8 * slot, (q / R) & 1u, P.err, 35);
            xb[kd] = ((sX + slot * P.x_slot) >> 4) | LBOX;
          }
          if (pc > 0) mbar_wait(hs0 + 8 * (role ^ 1u), ((pc - 1u) >> 1) & 1u, P.err, 37);  // previous plane is in the queue
          tc_fence_after();
          const uint32_t yb = ((sY + yslot * P.y_slot) >> 4) | LBOY;
          if (elect_one()) {
            uint32_t c0 = acc0, c1 = acc1, c2 = acc2;
            if (h0 && h2) {   // interior plane: straight-line issue, descriptors = uniform base + compile-time offsets
              for (int y = 0; y < rows; ++y) {
                const uint32_t alo = yb + (uint32_t)y * BWY * ry16;
                const uint32_t boff = (uint32_t)(y * P.TW) * rx16;
                const uint32_t b0 = xb[0] + boff, b1 = xb[1] + boff, b2 = xb[2] + boff;
#pragma unroll
                for (int j = 0; j < KSTEPS; ++j) {
                  umma_bf16_ss(tmem_base, hiA | (alo + j * 16 * ry16), hiB | (b0 + j * 16 * rx16), idesc, c0);
                  umma_bf16_ss(tmem_base + NB, hiA | (alo + j * 16 * ry16), hiB | (b1 + j * 16 * rx16), idesc, c1);
                  umma_bf16_ss(tmem_base + 2 * NB, hiA | (alo + j * 16 * ry16), hiB | (b2 + j * 16 * rx16), idesc, c2);
                  c0 = c1 = c2 = 1u;
                }
              }
            } else {          // first / last plane of the volume: one or two chains have no input plane
              for (int y = 0; y < rows; ++y) {
                uint32_t alo = yb + (uint32_t)y * BWY * ry16;
                uint32_t boff = (uint32_t)(y * P.TW) * rx16;
                for (int j = 0; j < KSTEPS; ++j) {
                  if (h0) { umma_bf16_ss(tmem_base, hiA | alo, hiB | (xb[0] + boff), idesc, c0); c0 = 1u; }
                  if (h1) { umma_bf16_ss(tmem_base + NB, hiA | alo, hiB | (xb[1] + boff), idesc, c1); c1 = 1u; }
                  if (h2) { umma_bf16_ss(tmem_base + 2 * NB, hiA | alo, hiB | (xb[2] + boff), idesc, c2); c2 = 1u; }
                  alo += 16u * ry16;
                  boff += 16u * rx16;
                }
              }
            }
            mbar_arrive(hs0 + 8 * role);                 // this plane is in the tensor queue
            umma_commit(xempty0 + 8 * ((Q0 + t) % R));   // plane z-1 is not needed by later output planes
            umma_commit(yempty0 + 8 * yslot);
          }
          __syncwarp();
        }
        if (rows > 0) {  // warp-uniform copy of the per-chain accumulate flags
          if (h0) acc0 = 1u;
          if (h1) acc1 = 1u;
          if (h2) acc2 = 1u;
        }
      }
      // X planes L and L+1 of this segment were only partially consumed: release them (by the warp that issued last)
      if (((pc - 1u) & 1u) == role && elect_one()) {
        umma_commit(xempty0 + 8 * ((Q0 + L) % R));
        umma_commit(xempty0 + 8 * ((Q0 + L + 1) % R));
      }
      __syncwarp();
      Q0 += (uint32_t)L + 2;
    }
    if (cur_key >= 0) {
      if (((pc - 1u) & 1u) == role && elect_one()) {
        *s_started = acc0 | (acc1 << 1) | (acc2 << 2);
        __threadfence_block();
        umma_commit(accfull);
      }
      __syncwarp();
    }
  } else {
    // ======================= epilogue: flush the TMEM accumulators of a finished key with fp32 reductions =======================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int g = m / GWY, co_l = m - g * GWY;   // M row -> (kw group, output channel)
    const bool row_ok = g < 3;
    int cur_key = -1;
    uint32_t flushes = 0;
    long long pos = lo;
    Wg2Seg sg;
    bool more = wg2_next_seg(pos, hi, P.D, P.columns, sg);
    while (true) {
      const int key = more ? sg.key : -2;
      if (key != cur_key && cur_key >= 0) {
        mbar_wait(accfull, flushes & 1u, P.err, 36);
        tc_fence_after();
        const uint32_t started = *s_started;  // chains without any MMA hold stale TMEM contents: skip them
        const int cob = cur_key % P.n_cob, cib = cur_key / P.n_cob;
        const int co = cob * GWY + co_l;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
          for (int j0 = 0; j0 < NB; j0 += 16) {
            uint32_t rr[16];
            tmem_ld16(trow + kd * NB + j0, rr);
            tmem_ld_wait();
            if (row_ok && ((started >> kd) & 1u) && co < P.Cout_pad) {
              const int kh = j0 / GWX;
              const int ci0 = cib * GWX + (j0 - kh * GWX);
              const int tap = (kd * 3 + kh) * 3 + (2 - g);
              float* dst = P.dwacc + ((long long)tap * P.Cin_pad + ci0) * P.Cout_pad + co;
#pragma unroll
              for (int j = 0; j < 16; ++j) atomicAdd(dst + (long long)j * P.Cout_pad, __uint_as_float(rr[j]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty);
        ++flushes;
      }
      if (!more) break;
      cur_key = key;
      more = wg2_next_seg(pos, hi, P.D, P.columns, sg);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TCOLS); }
}

template <int GWX, int GWY, int KSTEPS>
static int wg2_launch1(const Wg2Params& P, size_t smem, int grid, cudaStream_t stream) {
  static const cudaError_t attr = cudaFuncSetAttribute(wg2_kernel<GWX, GWY, KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // one-time, thread-safe
  B3D_CHECK_CUDA(attr);
  wg2_kernel<GWX, GWY, KSTEPS><<<grid, WG2_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}
template <int GWX, int GWY>
static int wg2_launch(const Wg2Params& P, size_t smem, int grid, cudaStream_t stream) {
  switch (P.TW / 16) {
    case 8: return wg2_launch1<GWX, GWY, 8>(P, smem, grid, stream);
    case 4: return wg2_launch1<GWX, GWY, 4>(P, smem, grid, stream);
    case 2: return wg2_launch1<GWX, GWY, 2>(P, smem, grid, stream);
    case 1: return wg2_launch1<GWX, GWY, 1>(P, smem, grid, stream);
  }
  b3d_set_error("wg2: unsupported tile width %d", P.TW);
  return B3D_ERR_UNSUPPORTED;
}

// Returns B3D_OK if launched, 1 if the shape is not suited (caller uses the generic kernel), negative on error.
// dwacc: zeroed fp32 [27][Cin_pad][Cout_pad]; x exposes Cin (multiple of 16) channels, dy Cout_pad (multiple of 16).
int b3d_try_wg2(const void* x, long long ldx, int Cin, const void* dy, long long lddy, int Cout_pad, int N, int D, int H, int W,
                float* dwacc, int Cin_pad, int* err_flag, cudaStream_t stream) {
  if (B3D_ENV_FLAG("B3D_NO_WG2")) return 1;
  if (W % 16 || W < 16 || W + 2 > 256 || Cin % 16 || Cout_pad % 16) return 1;
  const int GWX = (Cin % 32 == 0) ? 32 : 16, GWY = (Cout_pad % 32 == 0) ? 32 : 16;
  const int RBX = GWX * 2, RBY = GWY * 2;
  // tile = TH rows x TW columns (TW a multiple of 16 dividing W).  Shared memory holds R X-plane tiles of (TH+2) x TW and
  // RY dY-plane tiles of TH x (TW+2).  Deep rings matter more than tall tiles: a plane tile is consumed in ~50 clk per
  // MMA while a TMA box needs a few thousand clocks of latency under load, so we want >= 2 planes of prefetch.
  const size_t budget = (size_t)227 * 1024 - 1024 - (1024 + 256 + 1024);
  int TH = 0, TW = 0, R = 0, RY = 0;
  size_t xs = 0, ys = 0;
  double best = -1;
  const int env_th = B3D_ENV_INT("B3D_WG2_TH");
  const int env_tw = B3D_ENV_INT("B3D_WG2_TW");
  const int env_r = B3D_ENV_INT("B3D_WG2_R");
  // measured (B3D_WG2_* sweeps, 2x128^3 32x32): full-width rows, the tallest tile and the minimal ring (R=4, RY=2) win —
  // the per-plane-tile fixed cost matters more than prefetch depth
  for (int tw = std::min(W, 128); tw >= 16; tw /= 2) {
    if (tw % 16 || W % tw || (tw != 128 && tw != 64 && tw != 32 && tw != 16)) continue;
    if (env_tw && tw != env_tw) continue;
    for (int th = std::min(H, 32); th >= 1; --th) {
      if (env_th && th != env_th) continue;
      const size_t a = ((size_t)(th + 2) * tw * RBX + 1023) / 1024 * 1024;
      const size_t b = ((size_t)th * (tw + 2) * RBY + 1023) / 1024 * 1024;
      const int r = env_r ? env_r : 4;
      const int ry = std::min(WG2_MAXRY, r - 2);
      if (r * a + ry * b > budget) continue;
      const double halo = (double)(th + 2) / th;
      const double waste = (double)(((H + th - 1) / th) * th) / H;
      const double score = (double)th * tw / (halo * waste);
      if (score > best) { best = score; TH = th; TW = tw; R = r; RY = ry; xs = a; ys = b; }
    }
  }
  if (TH < 1) return 1;
  Wg2Params P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.D = D; P.H = H; P.W = W; P.TH = TH; P.TW = TW; P.R = R; P.RY = RY;
  P.tiles_y = (H + TH - 1) / TH; P.tiles_x = W / TW; P.columns = N * P.tiles_y * P.tiles_x;
  P.n_cib = Cin / GWX; P.n_cob = Cout_pad / GWY;
  P.x_slot = (uint32_t)xs; P.y_slot = (uint32_t)ys;
  P.x_tx = (uint32_t)((TH + 2) * TW * RBX); P.y_tx = (uint32_t)(TH * (TW + 2) * RBY);
  P.total_steps = (long long)P.n_cib * P.n_cob * P.columns * D;
  P.dwacc = dwacc; P.Cin_pad = Cin_pad; P.Cout_pad = Cout_pad; P.err = err_flag;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t sW = (uint64_t)ldx * 2;
    uint64_t strides[4] = {sW, sW * W, sW * W * H, sW * W * H * D};
    uint32_t box[5] = {(uint32_t)GWX, (uint32_t)TW, (uint32_t)(TH + 2), 1, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmX, x, 5, dims, strides, box, RBX);
    if (rc) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cout_pad, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t sW = (uint64_t)lddy * 2;
    uint64_t strides[4] = {sW, sW * W, sW * W * H, sW * W * H * D};
    uint32_t box[5] = {(uint32_t)GWY, (uint32_t)(TW + 2), (uint32_t)TH, 1, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmY, dy, 5, dims, strides, box, RBY);
    if (rc) return rc;
  }
  const size_t smem = (size_t)R * xs + (size_t)RY * ys + 1024 + 256 + 1024;
  B3D_CHECK_CUDA(cudaMemsetAsync(dwacc, 0, (size_t)27 * Cin_pad * Cout_pad * 4, stream));
  const int num_sms = b3d_num_sms();
  const int grid = (int)std::min<long long>(num_sms, P.total_steps);
  if (B3D_ENV_FLAG("B3D_VERBOSE"))
    fprintf(stderr, "[b3d] wg2 N%d D%d H%d W%d Cin%d Cout%d GWX%d GWY%d TH%d TW%d R%d RY%d keys%d cols%d grid%d smem%zu\n", N, D, H,
            W, Cin, Cout_pad, GWX, GWY, TH, TW, R, RY, P.n_cib * P.n_cob, P.columns, grid, smem);
  if (GWX == 32 && GWY == 32) return wg2_launch<32, 32>(P, smem, grid, stream);
  if (GWX == 32 && GWY == 16) return wg2_launch<32, 16>(P, smem, grid, stream);
  if (GWX == 16 && GWY == 32) return wg2_launch<16, 32>(P, smem, grid, stream);
  return wg2_launch<16, 16>(P, smem, grid, stream);
}
