// conv_zs.cu — "z-marching, kd-stacked" tcgen05 implicit-GEMM 3x3x3 convolution for the wide, shallow levels of the
// U-Net (main.py:216,219,130 at levels 0/1: few output channels, millions of voxels — ~3/4 of all conv FLOPs), fprop and
// dgrad (dgrad = same kernel on flipped/transposed packed weights).
//
// Why this shape (measured, scripts/umma_rate.cu -> profiles/umma_rate_r1.txt): one M128 x N x K16 tcgen05.mma costs
// max(N/2, (4096 + 32 N)/128) clocks — the SMEM operand fetch runs at 128 B/clk, so with voxels on M and Cout = 32 on N
// the tensor pipe can never exceed 40 %.  Stacking the three kd taps of one (kh,kw) on N (N = 3*Cout: the same A tile
// feeds the accumulators of the three output planes z-1, z, z+1) raises that bound to 86 % (Cout 32) / 100 % (Cout 64) and
// issues 3x fewer instructions from the single MMA thread.
//
// Structure:
//   * a CTA owns a 16(y) x 8(x) voxel column and marches along z.  The M-block is that 16x8 patch: the UMMA A descriptor
//     walks 16 groups of 8 consecutive x positions with SBO = one tile row, so every M row is a real voxel (no halo-gap
//     rows) and every (kh,kw) tap is a start-address shift inside the staged tile.
//   * each input plane tile (18 x 10 voxels x KC channels, ONE swizzled TMA box) is staged once and multiplied against
//     the resident weights of all 27 taps: 9 x KC/16 stacked MMAs per plane.
//   * accumulators: a ring of R = 512/COUT plane slots in TMEM.  Output plane g lives in slot R-1-(g mod R), so the
//     targets (z+1, z, z-1) <-> (kd 0,1,2) are ascending, contiguous columns except when the window wraps (2 of R steps:
//     two MMAs).  Slots are zeroed by the epilogue after draining, so every MMA accumulates.
//   * work = columns x D plane-steps, split evenly over the SMs (a CTA may finish one column and start the next).
//   * warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue (TMEM -> regs -> bias/statistics -> bf16 NDHWC).
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define ZS_THREADS 352   // warp 0 TMA, warps 1 and 6 MMA issuers (ping-pong), warps 2-5 and 7-10 epilogue (even / odd planes)
#define ZS_MAXSTAGES 8
#define ZS_Q 8            // depth of the scout -> issuer record queue
#define ZS_BW 10
#define ZS_BH 18
#define ZS_BOX 180

struct alignas(64) ZsParams {
  CUtensorMap tmA, tmW;
  int N, D, H, W, Cout;
  int tiles_x, tiles_y, n_blocks, tiles_per_nb;  // column = nb * tiles_per_nb + ((n * tiles_y + ty) * tiles_x + tx)
  int k_chunks, stages;
  int solo;               // 1: warp 1 issues every plane (bit-reproducible accumulation order), warp 6 idles
  long long total_steps;  // columns * D output planes
  bf16* out; long long ld_out;
  const float* bias;
  double* stats; int cpg, stats_groups, stats_batch;
  int* err;
};

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct ZsSeg { int col, za, zb; };

// CTA b of G handles output plane-steps [b*T/G, (b+1)*T/G); a segment is the part of that range inside one column.
__device__ __forceinline__ bool zs_next_seg(long long& pos, long long hi, int D, ZsSeg& s) {
  if (pos >= hi) return false;
  s.col = (int)(pos / D);
  s.za = (int)(pos - (long long)s.col * D);
  const long long rem = hi - pos;
  s.zb = (int)((rem < (long long)(D - s.za)) ? s.za + rem : D);
  pos += s.zb - s.za;
  return true;
}

template <int COUT, int KC>
__global__ void __launch_bounds__(ZS_THREADS, 1) zs_kernel(const __grid_constant__ ZsParams P) {
  constexpr int RB = KC * 2;                 // smem row bytes = swizzle span
  constexpr int NK16 = KC / 16;
  constexpr int R = 512 / COUT;              // ring slots
  constexpr uint32_t rb16 = RB / 16;
  constexpr uint32_t A_TX = ZS_BOX * RB;
  constexpr uint32_t A_STAGE = (A_TX + 1023u) / 1024u * 1024u;
  constexpr uint32_t W_TX = 27u * COUT * RB;
  constexpr uint32_t W_CHUNK = (W_TX + 1023u) / 1024u * 1024u;
  constexpr int LT = (RB == 128) ? 2 : (RB == 64 ? 4 : 6);

  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t sW = smem_u32(smem);
  const uint32_t sA = sW + P.k_chunks * W_CHUNK;
  uint8_t* aux = smem + (size_t)P.k_chunks * W_CHUNK + (size_t)S * A_STAGE;
  const uint32_t full0 = smem_u32(aux);                  // [ZS_MAXSTAGES]
  const uint32_t empty0 = full0 + 8 * ZS_MAXSTAGES;      // [ZS_MAXSTAGES]
  const uint32_t tfull0 = empty0 + 8 * ZS_MAXSTAGES;     // [32]
  const uint32_t tempty0 = tfull0 + 8 * 32;              // [32]
  const uint32_t wfull = tempty0 + 8 * 32;
  const uint32_t wfree = wfull + 8;
  const uint32_t hs0 = wfree + 8;                        // [2] issue hand-shake between the two MMA warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * ZS_MAXSTAGES + 16 * 32 + 32);
  // fp64 so that the order in which the 8 epilogue warps add their (fixed-order fp32) partial sums cannot change the result
  double* s_stats = reinterpret_cast<double*>(aux + 16 * ZS_MAXSTAGES + 16 * 32 + 48);  // [2 * 64]
  // scout -> issuer queue (P.solo == 2): ZS_Q records of 12 words + ready / free barriers
  uint32_t* q_rec = reinterpret_cast<uint32_t*>(aux + 16 * ZS_MAXSTAGES + 16 * 32 + 48 + 128 * 8);
  const uint32_t rdy0 = smem_u32(q_rec + ZS_Q * 12);
  const uint32_t fre0 = rdy0 + 8 * ZS_Q;

  if (threadIdx.x == 0) {
    if (sW & 1023u) { if (P.err) atomicExch(P.err, 29); __trap(); }
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < R; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
    mbar_init(wfull, 1);
    mbar_init(wfree, P.solo ? 1 : 2);   // committed by every warp that issues MMAs
    mbar_init(hs0, 1); mbar_init(hs0 + 8, 1);
    for (int i = 0; i < ZS_Q; ++i) { mbar_init(rdy0 + 8 * i, 1); mbar_init(fre0 + 8 * i, 1); }
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192) s_stats[threadIdx.x - 64] = 0.0;
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 2 && warp < 6) {  // zero the whole accumulator ring once (TMEM is not cleared by allocation)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 512; c += 16) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long T = P.total_steps;
  const long long lo = T * blockIdx.x / gridDim.x, hi = T * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmW); }
    uint32_t s = 0, ph = 0;
    int cur_nb = -1;
    uint32_t wloads = 0;
    long long pos = lo;
    ZsSeg sg;
    while (zs_next_seg(pos, hi, P.D, sg)) {
      const int nb = sg.col / P.tiles_per_nb;
      int t = sg.col - nb * P.tiles_per_nb;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y; const int n = t / P.tiles_y;
      if (nb != cur_nb) {  // (re)load the resident weights of this output-channel block
        if (wloads > 0) mbar_wait(wfree, (wloads - 1) & 1u, P.err, 30);
        if (elect_one()) {
          mbar_expect_tx(wfull, P.k_chunks * W_TX);
          for (int kc = 0; kc < P.k_chunks; ++kc) tma_load_3d(sW + kc * W_CHUNK, &P.tmW, wfull, kc * KC, nb * COUT, 0);
        }
        __syncwarp();
        cur_nb = nb; ++wloads;
      }
      const int zi0 = sg.za > 0 ? sg.za - 1 : 0;
      const int zi1 = sg.zb < P.D ? sg.zb : P.D - 1;
      for (int zi = zi0; zi <= zi1; ++zi) {
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(empty0 + 8 * s, ph ^ 1, P.err, 21);
          if (elect_one()) {
            const uint32_t fb = full0 + 8 * s;
            mbar_expect_tx(fb, A_TX);
            tma_load_5d(sA + s * A_STAGE, &P.tmA, fb, kc * KC, tx * 8 - 1, ty * 16 - 1, zi, n);
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (P.solo == 2 && warp == 6) {
    // ======================= scout (ordered-issue mode 2) =======================
    // Does everything of the issue loop EXCEPT issuing: the barrier waits (stage full, accumulator slot drained + zeroed,
    // weights resident), the ring arithmetic and the descriptor bases, and hands the single issuing thread one 12-word record
    // per (plane, K chunk).  tcgen05.mma is only ordered within a thread, so ONE issuer makes the fp32 accumulation order — and
    // with it the forward pass — bit-reproducible; the scout gives that thread the overlap the second issuer provided.
    const uint32_t idesc1 = umma_idesc_bf16(128, COUT, 0, 0);
    const uint32_t idesc2 = umma_idesc_bf16(128, 2 * COUT, 0, 0);
    const uint32_t idesc3 = umma_idesc_bf16(128, 3 * COUT, 0, 0);
    constexpr uint32_t RM = R - 1;
    constexpr uint32_t LBO1 = 1u << 16;
    uint32_t s = 0, ph = 0, g0 = 0, q = 0, qph = 0;
    int cur_nb = -1;
    uint32_t wloads = 0;
    long long pos = lo;
    ZsSeg sg;
    auto wait_fresh = [&](uint32_t g) { mbar_wait(tempty0 + 8 * (g & RM), (((g / R) & 1u) ^ 1u), P.err, 24); };
    auto post = [&](uint32_t a16, uint32_t w16, uint32_t w16b, uint32_t d1, uint32_t d2, uint32_t id1, uint32_t id2, uint32_t flags) {
      mbar_wait(fre0 + 8 * q, qph ^ 1u, P.err, 27);
      if (lane == 0) {
        uint32_t* r = q_rec + q * 12;
        r[0] = a16; r[1] = w16; r[2] = w16b; r[3] = d1; r[4] = d2; r[5] = id1; r[6] = id2; r[7] = flags;
        mbar_arrive(rdy0 + 8 * q);   // release: the record (and every wait above) is visible to the issuer
      }
      __syncwarp();
      if (++q == ZS_Q) { q = 0; qph ^= 1u; }
    };
    while (zs_next_seg(pos, hi, P.D, sg)) {
      const int nb = sg.col / P.tiles_per_nb;
      if (nb != cur_nb) {
        mbar_wait(wfull, wloads & 1u, P.err, 22);
        cur_nb = nb; ++wloads;
      }
      const int L = sg.zb - sg.za;
      const int i_min = sg.za > 0 ? 0 : 1;
      const int i_max = sg.zb < P.D ? L + 1 : L;
      bool release_w = false;
      if (P.n_blocks > 1) {
        long long p2 = pos; ZsSeg nx;
        release_w = zs_next_seg(p2, hi, P.D, nx) && nx.col / P.tiles_per_nb != nb;
      }
      for (int i = i_min; i <= i_max; ++i) {
        const int o_hi = i < L - 1 ? i : L - 1;
        const int o_lo = i - 2 > 0 ? i - 2 : 0;
        const int cnt = o_hi - o_lo + 1;
        const int kd0 = i - o_hi;
        if (i == i_min) { for (int o = o_lo; o <= o_hi; ++o) wait_fresh(g0 + (uint32_t)o); }
        else if (i <= L - 1) wait_fresh(g0 + (uint32_t)i);
        const uint32_t m = (g0 + (uint32_t)o_hi) & RM;
        const uint32_t p0 = RM - m;
        const int run1 = (cnt <= (int)m + 1) ? cnt : (int)m + 1;
        const int run2 = cnt - run1;
        const uint32_t d1 = tmem_base + p0 * COUT;
        const uint32_t d2 = tmem_base;
        const uint32_t id1 = run1 == 3 ? idesc3 : (run1 == 2 ? idesc2 : idesc1);
        const uint32_t id2 = run2 == 2 ? idesc2 : idesc1;
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(full0 + 8 * s, ph, P.err, 23);
          const uint32_t a16 = ((sA + s * A_STAGE) >> 4) | LBO1;
          const uint32_t w16 = (((sW + kc * W_CHUNK) >> 4) + (uint32_t)(kd0 * COUT) * rb16) | LBO1;
          const uint32_t w16b = w16 + (uint32_t)(run1 * COUT) * rb16;
          uint32_t flags = (run2 > 0 ? 1u : 0u) | (s << 4);
          if (kc == P.k_chunks - 1) {
            uint32_t nc = 0, slots = 0;
            if (i == i_max) { for (int o = o_lo; o <= o_hi; ++o) { slots |= ((g0 + (uint32_t)o) & RM) << (5 * nc); ++nc; } }
            else if (i >= 2) { slots = (g0 + (uint32_t)(i - 2)) & RM; nc = 1; }
            flags |= (nc << 8) | (slots << 10);
            if (i == i_max && release_w) flags |= 4u;
          }
          post(a16, w16, w16b, d1, d2, id1, id2, flags);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
      g0 += (uint32_t)L;
    }
    post(0u, 0u, 0u, 0u, 0u, 0u, 0u, 8u);   // END
  } else if (P.solo == 2 && warp == 1) {
    // ======================= the single issuer (ordered-issue mode 2) =======================
    const uint64_t hiA = umma_desc_hi_sw((uint32_t)ZS_BW * RB, LT);
    const uint64_t hiB = umma_desc_hi_sw(8u * RB, LT);
    uint32_t q = 0, qph = 0;
    for (;;) {
      mbar_wait(rdy0 + 8 * q, qph, P.err, 28);
      tc_fence_after();
      const uint32_t* r = q_rec + q * 12;
      const uint32_t a16 = r[0], w16 = r[1], w16b = r[2], d1 = r[3], d2 = r[4], id1 = r[5], id2 = r[6], flags = r[7];
      if (flags & 8u) break;
      if (elect_one()) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int k = 0; k < NK16; ++k) {
              const uint32_t alo = a16 + (uint32_t)((kh * ZS_BW + kw) * rb16 + k * 2);
              const uint32_t blo = w16 + (uint32_t)(((kh * 3 + kw) * 3 * COUT) * rb16 + k * 2);
              umma_bf16_ss(d1, hiA | alo, hiB | blo, id1, 1u);
            }
          }
        }
        if (flags & 1u) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < NK16; ++k) {
                const uint32_t alo = a16 + (uint32_t)((kh * ZS_BW + kw) * rb16 + k * 2);
                const uint32_t blo = w16b + (uint32_t)(((kh * 3 + kw) * 3 * COUT) * rb16 + k * 2);
                umma_bf16_ss(d2, hiA | alo, hiB | blo, id2, 1u);
              }
            }
          }
        }
        umma_commit(empty0 + 8 * ((flags >> 4) & 15u));
        const uint32_t nc = (flags >> 8) & 3u;
        for (uint32_t c = 0; c < nc; ++c) umma_commit(tfull0 + 8 * ((flags >> (10 + 5 * c)) & 31u));
        if (flags & 4u) umma_commit(wfree);
        mbar_arrive(fre0 + 8 * q);   // the record has been consumed
      }
      __syncwarp();
      if (++q == ZS_Q) { q = 0; qph ^= 1u; }
    }
  } else if (warp == 1 || warp == 6) {
    // ======================= MMA issuers (warp-uniform control flow, one elected lane issues) =======================
    // The per-plane bookkeeping below is deliberately lean (32-bit ring arithmetic, one barrier wait and one commit in
    // the steady state): the tensor pipe only queues a few MMAs, so every scalar instruction between two planes is idle
    // tensor time (profiles/ncu_zs_r1: 408 instructions per plane before this rewrite = 60 % idle).
    const uint64_t hiA = umma_desc_hi_sw((uint32_t)ZS_BW * RB, LT);   // groups of 8 x-positions, SBO = one tile row
    const uint64_t hiB = umma_desc_hi_sw(8u * RB, LT);
    const uint32_t idesc1 = umma_idesc_bf16(128, COUT, 0, 0);
    const uint32_t idesc2 = umma_idesc_bf16(128, 2 * COUT, 0, 0);
    const uint32_t idesc3 = umma_idesc_bf16(128, 3 * COUT, 0, 0);
    constexpr uint32_t RM = R - 1;
    constexpr uint32_t LBO1 = 1u << 16;   // descriptor LBO field (unused for swizzled K-major, conventionally 1)
    uint32_t s = 0, ph = 0;
    uint32_t g0 = 0;  // running count of output planes of this CTA (selects ring slot and barrier phase)
    uint32_t pc = 0;  // running count of input planes (alternates between the two issuer warps)
    const uint32_t role = (warp == 1) ? 0u : 1u;
    int cur_nb = -1;
    uint32_t wloads = 0;
    long long pos = (P.solo && role == 1u) ? hi : lo;   // ordered-issue mode: warp 6 idles
    ZsSeg sg;
    auto wait_fresh = [&](uint32_t g) {   // the epilogue has drained + zeroed the slot that plane g re-uses
      mbar_wait(tempty0 + 8 * (g & RM), (((g / R) & 1u) ^ 1u), P.err, 24);
    };
    while (zs_next_seg(pos, hi, P.D, sg)) {
      const int nb = sg.col / P.tiles_per_nb;
      if (nb != cur_nb) {
        mbar_wait(wfull, wloads & 1u, P.err, 22);
        tc_fence_after();
        cur_nb = nb; ++wloads;
      }
      const int L = sg.zb - sg.za;
      const int i_min = sg.za > 0 ? 0 : 1;            // plane index i <-> input plane z = za - 1 + i
      const int i_max = sg.zb < P.D ? L + 1 : L;
      // Two issuer warps alternate planes: while one is blocked feeding its plane's MMAs into the (shallow) tensor queue,
      // the other does the barrier waits / ring arithmetic of the next plane, then issues right behind it.  The tensor
      // pipe executes in issue order, so accumulations into shared ring slots stay ordered; hs[r] = "warp r has issued".
      for (int i = i_min; i <= i_max; ++i, ++pc) {
        if (!P.solo && (pc & 1u) != role) {   // the other warp's plane: only keep the stage ring position in step
          for (int kc = 0; kc < P.k_chunks; ++kc) if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
          continue;
        }
        const int o_hi = i < L - 1 ? i : L - 1;         // newest target (smallest kd)
        const int o_lo = i - 2 > 0 ? i - 2 : 0;         // oldest target
        const int cnt = o_hi - o_lo + 1;                // 1..3 targets
        const int kd0 = i - o_hi;                       // kd of the newest target
        // targets that receive their first contribution from this plane
        if (i == i_min) { for (int o = o_lo; o <= o_hi; ++o) wait_fresh(g0 + (uint32_t)o); }
        else if (i <= L - 1) wait_fresh(g0 + (uint32_t)i);
        // column runs: target o_hi - j sits at ring position R-1-((g_hi - j) mod R) = p0 + j until it wraps to 0
        const uint32_t m = (g0 + (uint32_t)o_hi) & RM;
        const uint32_t p0 = RM - m;
        const int run1 = (cnt <= (int)m + 1) ? cnt : (int)m + 1;
        const int run2 = cnt - run1;                     // wrapped part, starts at ring position 0
        const uint32_t d1 = tmem_base + p0 * COUT;
        const uint32_t d2 = tmem_base;
        const uint32_t id1 = run1 == 3 ? idesc3 : (run1 == 2 ? idesc2 : idesc1);
        const uint32_t id2 = run2 == 2 ? idesc2 : idesc1;
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(full0 + 8 * s, ph, P.err, 23);
          if (!P.solo && kc == 0 && pc > 0) mbar_wait(hs0 + 8 * (role ^ 1u), ((pc - 1u) >> 1) & 1u, P.err, 26);  // previous plane issued
          tc_fence_after();
          const uint32_t a16 = ((sA + s * A_STAGE) >> 4) | LBO1;
          const uint32_t w16 = (((sW + kc * W_CHUNK) >> 4) + (uint32_t)(kd0 * COUT) * rb16) | LBO1;
          if (elect_one()) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                for (int k = 0; k < NK16; ++k) {
                  const uint32_t alo = a16 + (uint32_t)((kh * ZS_BW + kw) * rb16 + k * 2);
                  const uint32_t blo = w16 + (uint32_t)(((kh * 3 + kw) * 3 * COUT) * rb16 + k * 2);
                  umma_bf16_ss(d1, hiA | alo, hiB | blo, id1, 1u);
                }
              }
            }
            if (run2 > 0) {
              const uint32_t w16b = w16 + (uint32_t)(run1 * COUT) * rb16;
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                  for (int k = 0; k < NK16; ++k) {
                    const uint32_t alo = a16 + (uint32_t)((kh * ZS_BW + kw) * rb16 + k * 2);
                    const uint32_t blo = w16b + (uint32_t)(((kh * 3 + kw) * 3 * COUT) * rb16 + k * 2);
                    umma_bf16_ss(d2, hiA | alo, hiB | blo, id2, 1u);
                  }
                }
              }
            }
            if (kc == P.k_chunks - 1) mbar_arrive(hs0 + 8 * role);   // this plane is in the tensor queue
            umma_commit(empty0 + 8 * s);
            // targets whose last contribution was this plane are complete
            if (kc == P.k_chunks - 1) {
              if (i == i_max) { for (int o = o_lo; o <= o_hi; ++o) umma_commit(tfull0 + 8 * ((g0 + (uint32_t)o) & RM)); }
              else if (i >= 2) umma_commit(tfull0 + 8 * ((g0 + (uint32_t)(i - 2)) & RM));
            }
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
      g0 += (uint32_t)L;
      // the next segment may bring other weights: tell the producer when this segment's MMAs no longer read them
      if (P.n_blocks > 1) {
        long long p2 = pos; ZsSeg nx;
        if (zs_next_seg(p2, hi, P.D, nx) && nx.col / P.tiles_per_nb != nb) {
          if (elect_one()) umma_commit(wfree);
          __syncwarp();
        }
      }
    }
  } else {
    // ======================= epilogue: warps 2..5 drain even output planes, warps 7..10 odd ones =======================
    // (TMEM lane quadrant = warp % 4; two groups because one plane's drain + bf16 store costs more than its MMAs at small Cin)
    const int q = warp & 3;
    const uint32_t grp = (warp >= 7) ? 1u : 0u;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;   // M row -> voxel of the 16x8 patch
    const int et = grp ? 9999 : (int)threadIdx.x - 64;   // group-0 threads 0..127 publish the statistics
    const int cpg = P.cpg;
    const bool fine = (cpg < 4);             // statistics granule: 4 channels (cpg % 4 == 0) or 1 channel
    constexpr int NACC = COUT / 4 > 16 ? COUT / 4 : 16;
    float a1[NACC], a2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool has_stats = P.stats != nullptr, has_bias = P.bias != nullptr;
    uint32_t g0 = 0;
    long long pos = lo;
    ZsSeg sg;
    while (zs_next_seg(pos, hi, P.D, sg)) {
      const int nb = sg.col / P.tiles_per_nb;
      int t = sg.col - nb * P.tiles_per_nb;
      const int tx = t % P.tiles_x; t /= P.tiles_x;
      const int ty = t % P.tiles_y; const int n = t / P.tiles_y;
      const int y = ty * 16 + yl, x = tx * 8 + xl;
      const bool valid = (y < P.H) && (x < P.W);
      const int c_base = nb * COUT;
      const bool full = (c_base + COUT <= P.Cout);
      const int L = sg.zb - sg.za;
      bf16* op0 = P.out + ((((long long)n * P.D + sg.za) * P.H + y) * P.W + x) * P.ld_out + c_base;
      const long long plane_stride = (long long)P.H * P.W * P.ld_out;
      for (int o = 0; o < L; ++o) {
        const uint32_t g = g0 + (uint32_t)o;
        if ((g & 1u) != grp) continue;
        const uint32_t slot = g % R, use = g / R;
        const uint32_t pcol = (uint32_t)(R - 1 - slot) * COUT;
        mbar_wait(tfull0 + 8 * slot, use & 1u, P.err, 25);
        tc_fence_after();
        uint32_t r[COUT];
#pragma unroll
        for (int j0 = 0; j0 < COUT; j0 += 16) tmem_ld16(lane_base + pcol + j0, r + j0);
        tmem_ld_wait();
#pragma unroll
        for (int j0 = 0; j0 < COUT; j0 += 16) tmem_st16_zero(lane_base + pcol + j0);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * slot);
        bf16* op = op0 + (long long)o * plane_stride;
#pragma unroll
        for (int j0 = 0; j0 < COUT; j0 += 16) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j0 + j]);
          if (has_bias) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + c_base + j0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (c_base + j0 + j < P.Cout) v[j] += __ldg(P.bias + c_base + j0 + j);
            }
          }
          if (valid) {
            if (has_stats) {
              if (!fine) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                  a1[j0 / 4 + qd] += (v[4 * qd] + v[4 * qd + 1]) + (v[4 * qd + 2] + v[4 * qd + 3]);
                  a2[j0 / 4 + qd] += (v[4 * qd] * v[4 * qd] + v[4 * qd + 1] * v[4 * qd + 1]) +
                                     (v[4 * qd + 2] * v[4 * qd + 2] + v[4 * qd + 3] * v[4 * qd + 3]);
                }
              } else if (j0 == 0) {  // per-channel granule: only the first 16 channels of a block (COUT == 16)
#pragma unroll
                for (int j = 0; j < 16; ++j) { a1[j] += v[j]; a2[j] = fmaf(v[j], v[j], a2[j]); }
              }
            }
            if (full && ((reinterpret_cast<uintptr_t>(op + j0) & 31) == 0)) {
              const uint4 u0 = pack8(v), u1 = pack8(v + 8);
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + j0), "r"(u0.x), "r"(u0.y), "r"(u0.z),
                           "r"(u0.w), "r"(u1.x), "r"(u1.y), "r"(u1.z), "r"(u1.w) : "memory");
            } else {
              if (c_base + j0 + 8 <= P.Cout) stg16(op + j0, pack8(v));
              if (c_base + j0 + 16 <= P.Cout) stg16(op + j0 + 8, pack8(v + 8));
            }
          }
        }
      }
      g0 += (uint32_t)L;
      // flush this segment's statistics (one sample, one output-channel block)
      if (has_stats) {
        const int gran = fine ? 1 : 4;
        const int nacc = fine ? 16 : COUT / 4;
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
          if (i < nacc) {
            const float s1 = warp_sum(a1[i]), s2 = warp_sum(a2[i]);
            if (lane == 0) {
              const int gl = (i * gran) / cpg;  // group local to this channel block
              atomicAdd(&s_stats[2 * gl], (double)s1);
              atomicAdd(&s_stats[2 * gl + 1], (double)s2);
            }
          }
          a1[i] = 0.f; a2[i] = 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int groups_blk = (COUT + cpg - 1) / cpg;
        if (et < 2 * groups_blk) {
          const int gi = c_base / cpg + (et >> 1);
          if (gi < P.stats_groups) {
            const int ns = P.stats_batch ? 0 : n;
            atomicAdd(P.stats + ((long long)ns * P.stats_groups + gi) * 2 + (et & 1), s_stats[et]);
          }
          s_stats[et] = 0.0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int COUT, int KC>
static int zs_launch(const ZsParams& P, size_t smem, int grid, cudaStream_t stream) {
  static const cudaError_t attr =   // function-local static: initialised once, thread-safe
      cudaFuncSetAttribute(zs_kernel<COUT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  B3D_CHECK_CUDA(attr);
  zs_kernel<COUT, KC><<<grid, ZS_THREADS, smem, stream>>>(P); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// Returns B3D_OK if launched, 1 if this shape is not suited (caller uses the block-mode kernel), negative on error.
// wpack: [27 taps in (kh,kw,kd) order][w_rows][Cin] bf16 (b3d_pack_weight modes 0/1).
int b3d_try_zs(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
               int N, int D, int H, int W, int Cin, int Cout, double* stats, int cpg, int stats_groups, int stats_batch,
               int* err_flag, cudaStream_t stream) {
  if (B3D_ENV_FLAG("B3D_NO_ZS")) return 1;
  const int CoutPad = w_rows;
  if (CoutPad % 16 || Cin % 16) return 1;
  if (D < 2 || (long long)H * W < 256) return 1;
  const bool big_ok = (CoutPad <= 64) || (CoutPad == 128 && Cin <= 64);
  if (!big_ok && !B3D_ENV_FLAG("B3D_ZS_ALL")) return 1;
  const int KC = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  const int RB = KC * 2;
  const int k_chunks = Cin / KC;
  const uint32_t a_stage = (uint32_t)(ZS_BOX * RB + 1023) / 1024 * 1024;
  const size_t aux_bytes = 16 * ZS_MAXSTAGES + 16 * 32 + 32 + 128 * 8 + 64 + ZS_Q * 64;
  const size_t budget = 227 * 1024 - 1024 - aux_bytes;
  int COUT = 0, stages = 0;
  const int forced = B3D_ENV_INT("B3D_ZS_COUT");
  const int cands[3] = {64, 32, 16};
  for (int ci = 0; ci < 3; ++ci) {
    const int c = cands[ci];
    if (CoutPad % c) continue;
    if (forced && c != forced) continue;
    const size_t wbytes = (size_t)k_chunks * (((size_t)27 * c * RB + 1023) / 1024 * 1024);
    if (wbytes + 3 * (size_t)a_stage > budget) continue;
    COUT = c;
    stages = (int)std::min<size_t>(ZS_MAXSTAGES, (budget - wbytes) / a_stage);
    break;
  }
  if (!COUT) return 1;
  if (stats) {
    if (cpg <= 0) return 1;
    const bool ok4 = (cpg % 4 == 0) && (COUT % cpg == 0 || cpg % COUT == 0);
    const bool ok1 = (cpg < 4) && COUT == 16 && (16 % cpg == 0);
    if (!ok4 && !ok1) return 1;
    if (cpg % COUT == 0 && cpg > COUT) { /* several channel blocks share one group: handled by gi = c_base / cpg */ }
  }
  ZsParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.D = D; P.H = H; P.W = W; P.Cout = Cout;
  P.tiles_x = (W + 7) / 8; P.tiles_y = (H + 15) / 16; P.n_blocks = CoutPad / COUT;
  P.tiles_per_nb = N * P.tiles_x * P.tiles_y;
  P.k_chunks = k_chunks; P.stages = stages;
  P.solo = g_b3d_ordered_issue.load();   // 0: two ping-pong issuers, 1: one issuer, 2: one issuer fed by a scout warp
  P.total_steps = (long long)P.tiles_per_nb * P.n_blocks * D;
  P.out = (bf16*)y; P.ld_out = ldy; P.bias = bias;
  P.stats = stats; P.cpg = cpg > 0 ? cpg : 16; P.stats_groups = stats_groups; P.stats_batch = stats_batch; P.err = err_flag;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t sW = (uint64_t)ldx * 2;
    uint64_t strides[4] = {sW, sW * W, sW * W * H, sW * W * H * D};
    uint32_t box[5] = {(uint32_t)KC, ZS_BW, ZS_BH, 1, 1};
    int rc = b3d_encode_tmap_bf16(&P.tmA, x, 5, dims, strides, box, RB);
    if (rc) return rc;
    uint64_t wd[3] = {(uint64_t)Cin, (uint64_t)CoutPad, 27};
    uint64_t ws[2] = {(uint64_t)Cin * 2, (uint64_t)CoutPad * Cin * 2};
    uint32_t wb[3] = {(uint32_t)KC, (uint32_t)COUT, 27};
    rc = b3d_encode_tmap_bf16(&P.tmW, wpack, 3, wd, ws, wb, RB);
    if (rc) return rc;
  }
  const size_t wbytes = (size_t)k_chunks * (((size_t)27 * COUT * RB + 1023) / 1024 * 1024);
  const size_t smem = wbytes + (size_t)stages * a_stage + aux_bytes + 1024;
  if (smem > 227 * 1024) return 1;
  const int num_sms = b3d_num_sms();
  const int grid = (int)std::min<long long>(num_sms, P.total_steps);
  if (B3D_ENV_FLAG("B3D_VERBOSE"))
    fprintf(stderr, "[b3d] zs N%d D%d H%d W%d Cin%d Cout%d COUT%d KC%d chunks%d stages%d nblk%d cols%d grid%d smem%zu\n", N, D, H,
            W, Cin, Cout, COUT, KC, k_chunks, stages, P.n_blocks, P.tiles_per_nb * P.n_blocks, grid, smem);
#define ZS_CASE(C, K) if (COUT == C && KC == K) return zs_launch<C, K>(P, smem, grid, stream);
  ZS_CASE(16, 16) ZS_CASE(16, 32) ZS_CASE(16, 64)
  ZS_CASE(32, 16) ZS_CASE(32, 32) ZS_CASE(32, 64)
  ZS_CASE(64, 16) ZS_CASE(64, 32) ZS_CASE(64, 64)
#undef ZS_CASE
  return 1;
}
