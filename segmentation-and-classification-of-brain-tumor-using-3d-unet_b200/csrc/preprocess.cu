// preprocess.cu — the input pipeline in front of the hot path on the GPU.
//
// Reference semantics (BraTSDataset, /root/reference/training.py):
//   _preprocess_image         :117-132  p1, p99 = np.percentile(image, (1, 99)); clip; (x - mean) / (std + 1e-8);
//                                       ndimage.zoom(order=1) to 128^3; float32
//   _preprocess_segmentation  :134-146  seg[seg == 4] = 3; ndimage.zoom(order=0); uint8
//   _apply_augmentations      :148-172  rot90 in the (D,H) plane, flips, Gaussian noise, intensity scale
// On the host this is a sort (np.percentile), three full passes and two scipy zooms per modality — ~1 s per case, while a
// B200 consumes a case in ~9 ms.  Here:
//   * exact order statistics by 3-pass radix select on monotone 32-bit keys (11 + 11 + 10 bits, four ranks at once: the two
//     neighbours of each percentile), shared-memory histograms, every decision on the device (no host round trip);
//   * clip + moments in one pass (fp64 accumulation);
//   * clip + z-score + trilinear zoom fused in one gather pass, scipy's NI_ZoomShift coordinate rule (o * (in-1)/(out-1)),
//     arithmetic in double like the reference; nearest-neighbour zoom + label remap for the mask (floor(c + 0.5));
//   * one gather pass for rot90 / flips / noise (Philox4x32-10 + Box-Muller, counter = element index) / scale.
// All of it is HBM/L2-bound CUDA-core work: a 240x240x155 modality is 36 MB and stays L2 resident across the passes.
#include "b3d_common.cuh"
#include "b3d_internal.h"
#include <algorithm>

#define PP_THREADS 256
#define PP_BINS 2048
#define PP_NT 4   // ranks selected simultaneously

// state: [0..3] key prefix found so far, [4..7] remaining rank inside that prefix (as unsigned long long pairs below)
struct SelState {
  unsigned int prefix[PP_NT];
  unsigned long long rank[PP_NT];
};

__device__ __forceinline__ unsigned int f2key(float x) {
  const unsigned int u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
  const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

static int pp_blocks(long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + PP_THREADS * 4 - 1) / (PP_THREADS * 4), (long long)b3d_num_sms() * 8));
}

// pass p of the radix select: histogram of the next digit of every element that matches each rank's prefix so far
// digits: pass 0 = bits 31..21 (11), pass 1 = bits 20..10 (11), pass 2 = bits 9..0 (10)
__global__ void __launch_bounds__(PP_THREADS) pp_hist_kernel(const float* __restrict__ x, long long n, int pass,
                                                             const SelState* __restrict__ st, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[PP_NT * PP_BINS];
  for (int i = threadIdx.x; i < PP_NT * PP_BINS; i += PP_THREADS) sh[i] = 0u;
  __syncthreads();
  const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
  const unsigned int dmask = pass == 2 ? 1023u : 2047u;
  const int hi_shift = pass == 0 ? 32 : (pass == 1 ? 21 : 10);
  unsigned int pre[PP_NT];
#pragma unroll
  for (int t = 0; t < PP_NT; ++t) pre[t] = st->prefix[t];
  // ranks that share a prefix share a histogram (pass 0: all of them)
  bool same[PP_NT];
  same[0] = false;
#pragma unroll
  for (int t = 1; t < PP_NT; ++t) same[t] = (pass == 0) || (pre[t] == pre[t - 1]);
  for (long long i = (long long)blockIdx.x * PP_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * PP_THREADS) {
    const unsigned int k = f2key(__ldg(x + i));
    const unsigned int hi = pass == 0 ? 0u : (k >> hi_shift);
    const unsigned int d = (k >> shift) & dmask;
#pragma unroll
    for (int t = 0; t < PP_NT; ++t)
      if (!same[t] && (pass == 0 || hi == pre[t])) atomicAdd(&sh[t * PP_BINS + d], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PP_NT * PP_BINS; i += PP_THREADS)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void pp_init_kernel(SelState* __restrict__ st, SelState h) { if (threadIdx.x == 0 && blockIdx.x == 0) *st = h; }

// one CTA: walk each rank's histogram to the bin that contains it; extend the prefix, reduce the rank; clear the histogram
__global__ void pp_scan_kernel(SelState* __restrict__ st, unsigned int* __restrict__ hist, int pass) {
  __shared__ unsigned int pre_in[PP_NT];
  if (threadIdx.x < PP_NT) pre_in[threadIdx.x] = st->prefix[threadIdx.x];
  __syncthreads();
  if (threadIdx.x < PP_NT) {
    const int t = threadIdx.x;
    int src = t;   // the histogram this rank was counted in (ranks with equal prefixes share the first one)
    while (src > 0 && (pass == 0 || pre_in[src] == pre_in[src - 1])) --src;
    const int bits = pass == 2 ? 10 : 11;
    const int nb = 1 << bits;
    unsigned long long r = st->rank[t], cum = 0;
    int b = 0;
    for (; b < nb; ++b) {
      const unsigned long long c = hist[src * PP_BINS + b];
      if (cum + c > r) break;
      cum += c;
    }
    if (b >= nb) b = nb - 1;
    st->prefix[t] = (pre_in[t] << bits) | (unsigned int)b;
    st->rank[t] = r - cum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PP_NT * PP_BINS; i += blockDim.x) hist[i] = 0u;
}

// stats[0] = p1, [1] = p99 (np.percentile 'linear': lerp of the two neighbouring order statistics), [2..3] reserved for moments
__global__ void pp_percentile_kernel(const SelState* __restrict__ st, double frac_lo, double frac_hi, double* __restrict__ stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double a0 = (double)key2f(st->prefix[0]), a1 = (double)key2f(st->prefix[1]);
  const double b0 = (double)key2f(st->prefix[2]), b1 = (double)key2f(st->prefix[3]);
  stats[0] = a0 + (a1 - a0) * frac_lo;
  stats[1] = b0 + (b1 - b0) * frac_hi;
  stats[2] = 0.0; stats[3] = 0.0;
}

// Σ clip(x), Σ clip(x)^2 in fp64 -> stats[2], stats[3]
__global__ void __launch_bounds__(PP_THREADS) pp_moments_kernel(const float* __restrict__ x, long long n, double* __restrict__ stats) {
  __shared__ double s1[PP_THREADS / 32], s2[PP_THREADS / 32];
  const double lo = stats[0], hi = stats[1];
  double a = 0.0, b = 0.0;
  for (long long i = (long long)blockIdx.x * PP_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * PP_THREADS) {
    const double v = fmin(fmax((double)__ldg(x + i), lo), hi);
    a += v; b += v * v;
  }
  a = warp_sum_d(a); b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < PP_THREADS / 32; ++w) { ta += s1[w]; tb += s2[w]; }   // fixed order
    atomicAdd(&stats[2], ta); atomicAdd(&stats[3], tb);
  }
}

// stats -> [0] p1 [1] p99 [2] mean [3] std (population, np.std)
__global__ void pp_moments_finalize_kernel(double* __restrict__ stats, double n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double mean = stats[2] / n;
  double var = stats[3] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2] = mean; stats[3] = sqrt(var);
}

__device__ __forceinline__ void zoom_coord(int o, double zoom, int nin, int& i0, int& i1, double& f) {
  const double c = (double)o * zoom;
  i0 = (int)floor(c);
  if (i0 > nin - 1) i0 = nin - 1;
  i1 = i0 + 1 < nin ? i0 + 1 : nin - 1;
  f = c - (double)i0;
}

// out[oz][oy][ox] = trilinear( (clip(x) - mean) / (std + 1e-8) ), scipy.ndimage.zoom(order=1) coordinates
__global__ void __launch_bounds__(PP_THREADS) pp_zoom_norm_kernel(const float* __restrict__ x, int D, int H, int W,
                                                                  const double* __restrict__ stats, float* __restrict__ out, int OD,
                                                                  int OH, int OW) {
  const double lo = stats[0], hi = stats[1], mean = stats[2], inv = 1.0 / (stats[3] + 1e-8);
  const double zd = OD > 1 ? (double)(D - 1) / (double)(OD - 1) : 0.0, zh = OH > 1 ? (double)(H - 1) / (double)(OH - 1) : 0.0,
               zw = OW > 1 ? (double)(W - 1) / (double)(OW - 1) : 0.0;
  const long long total = (long long)OD * OH * OW;
  const bool same = (D == OD && H == OH && W == OW);
  for (long long i = (long long)blockIdx.x * PP_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * PP_THREADS) {
    if (same) {
      out[i] = (float)((fmin(fmax((double)__ldg(x + i), lo), hi) - mean) * inv);
      continue;
    }
    long long t = i;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH); const int oz = (int)(t / OH);
    int z0, z1, y0, y1, x0, x1; double fz, fy, fx;
    zoom_coord(oz, zd, D, z0, z1, fz); zoom_coord(oy, zh, H, y0, y1, fy); zoom_coord(ox, zw, W, x0, x1, fx);
#define TAP(zz, yy, xx) ((fmin(fmax((double)__ldg(x + ((long long)(zz) * H + (yy)) * W + (xx)), lo), hi) - mean) * inv)
    // the oracle's separable order: axis 0 (D), then 1 (H), then 2 (W)
    const double a00 = TAP(z0, y0, x0) * (1.0 - fz) + TAP(z1, y0, x0) * fz, a01 = TAP(z0, y0, x1) * (1.0 - fz) + TAP(z1, y0, x1) * fz;
    const double a10 = TAP(z0, y1, x0) * (1.0 - fz) + TAP(z1, y1, x0) * fz, a11 = TAP(z0, y1, x1) * (1.0 - fz) + TAP(z1, y1, x1) * fz;
#undef TAP
    const double b0 = a00 * (1.0 - fy) + a10 * fy, b1 = a01 * (1.0 - fy) + a11 * fy;
    out[i] = (float)(b0 * (1.0 - fx) + b1 * fx);
  }
}

// label map: value 4 -> 3, nearest-neighbour zoom (floor(c + 0.5)); output uint8 or int64
template <typename OutT>
__global__ void __launch_bounds__(PP_THREADS) pp_zoom_label_kernel(const float* __restrict__ seg, int D, int H, int W,
                                                                   OutT* __restrict__ out, int OD, int OH, int OW) {
  const double zd = OD > 1 ? (double)(D - 1) / (double)(OD - 1) : 0.0, zh = OH > 1 ? (double)(H - 1) / (double)(OH - 1) : 0.0,
               zw = OW > 1 ? (double)(W - 1) / (double)(OW - 1) : 0.0;
  const long long total = (long long)OD * OH * OW;
  const bool same = (D == OD && H == OH && W == OW);
  for (long long i = (long long)blockIdx.x * PP_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * PP_THREADS) {
    long long src = i;
    if (!same) {
      long long t = i;
      const int ox = (int)(t % OW); t /= OW;
      const int oy = (int)(t % OH); const int oz = (int)(t / OH);
      const int z = min((int)floor((double)oz * zd + 0.5), D - 1), y = min((int)floor((double)oy * zh + 0.5), H - 1),
                xx = min((int)floor((double)ox * zw + 0.5), W - 1);
      src = ((long long)z * H + y) * W + xx;
    }
    float v = __ldg(seg + src);
    if (v == 4.f) v = 3.f;
    out[i] = (OutT)v;
  }
}

// ---- augmentation -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0,
                                              unsigned int k1, unsigned int* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// out image [C][D][H][W] and labels [D][H][W]: out = flip(rot90_k(in)) ; image = (image + N(0, noise_std)) * scale
template <typename LabT>
__global__ void __launch_bounds__(PP_THREADS) pp_augment_kernel(const float* __restrict__ img, const LabT* __restrict__ lab,
                                                                float* __restrict__ oimg, LabT* __restrict__ olab, int C, int D, int H,
                                                                int W, int k, int fd, int fh, int fw, float noise_std, float scale,
                                                                unsigned long long seed) {
  const long long V = (long long)D * H * W;
  for (long long i = (long long)blockIdx.x * PP_THREADS + threadIdx.x; i < V; i += (long long)gridDim.x * PP_THREADS) {
    long long t = i;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H); int d = (int)(t / H);
    if (fd) d = D - 1 - d;           // undo the flips (applied after the rotation in the reference)
    if (fh) h = H - 1 - h;
    if (fw) w = W - 1 - w;
    int sd = d, sh = h;              // undo np.rot90(k, axes=(0,1)): out[i][j] = in[j][n-1-i] (k=1), in[n-1-i][n-1-j] (2), in[n-1-j][i] (3)
    if (k == 1) { sd = h; sh = H - 1 - d; }
    else if (k == 2) { sd = D - 1 - d; sh = H - 1 - h; }
    else if (k == 3) { sd = D - 1 - h; sh = d; }
    const long long src = ((long long)sd * H + sh) * W + w;
    if (lab != nullptr) olab[i] = lab[src];
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise_std != 0.f) {
      for (int c0 = 0; c0 < C; c0 += 4) {
        unsigned int r[4];
        philox4x32_10((unsigned int)i, (unsigned int)(i >> 32), (unsigned int)c0, 0x3d1f5a7u, (unsigned int)seed,
                      (unsigned int)(seed >> 32), r);
        // two Box-Muller pairs -> four standard normals
        const float u0 = ((float)r[0] + 0.5f) * 2.3283064e-10f, u1 = ((float)r[1] + 0.5f) * 2.3283064e-10f;
        const float u2 = ((float)r[2] + 0.5f) * 2.3283064e-10f, u3 = ((float)r[3] + 0.5f) * 2.3283064e-10f;
        const float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u2));
        float s0, c0f, s1, c1f;
        sincosf(6.2831853f * u1, &s0, &c0f); sincosf(6.2831853f * u3, &s1, &c1f);
        nz[0] = m0 * c0f; nz[1] = m0 * s0; nz[2] = m1 * c1f; nz[3] = m1 * s1;
        for (int c = c0; c < C && c < c0 + 4; ++c)
          oimg[(long long)c * V + i] = (__ldg(img + (long long)c * V + src) + noise_std * nz[c - c0]) * scale;
      }
    } else {
      for (int c = 0; c < C; ++c) oimg[(long long)c * V + i] = __ldg(img + (long long)c * V + src) * scale;
    }
  }
}

extern "C" {

// Exact np.percentile(x, (q_lo, q_hi)) (method 'linear'), then mean / population std of clip(x, p_lo, p_hi) — all on the device.
// x: fp32 [n]; work: >= 4*2048*4 + 64 bytes of device scratch; stats: double[4] <- p_lo, p_hi, mean, std.
// Replaces np.percentile / np.clip / np.mean / np.std of _preprocess_image (training.py:119-124).
int b3d_clip_stats(const float* x, long long n, double q_lo, double q_hi, void* work, size_t work_bytes, double* stats,
                   void* stream) {
  B3D_REQUIRE(n >= 1, "clip_stats: empty volume");
  B3D_REQUIRE(work_bytes >= sizeof(unsigned int) * PP_NT * PP_BINS + sizeof(SelState), "clip_stats: work buffer too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* hist = (unsigned int*)work;
  SelState* sel = (SelState*)((char*)work + sizeof(unsigned int) * PP_NT * PP_BINS);
  SelState h;
  double frac[2];
  const double qs[2] = {q_lo, q_hi};
  for (int j = 0; j < 2; ++j) {
    const double r = qs[j] / 100.0 * (double)(n - 1);
    const long long lo = (long long)floor(r);
    const long long hi = std::min<long long>(lo + 1, n - 1);
    h.prefix[2 * j] = 0u; h.prefix[2 * j + 1] = 0u;
    h.rank[2 * j] = (unsigned long long)lo; h.rank[2 * j + 1] = (unsigned long long)hi;
    frac[j] = r - (double)lo;
  }
  B3D_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned int) * PP_NT * PP_BINS, st));
  pp_init_kernel<<<1, 32, 0, st>>>(sel, h); ++g_b3d_launches;   // the 48-byte initial state travels as a kernel parameter
  const int blocks = pp_blocks(n);
  for (int pass = 0; pass < 3; ++pass) {
    pp_hist_kernel<<<blocks, PP_THREADS, 0, st>>>(x, n, pass, sel, hist); ++g_b3d_launches;
    pp_scan_kernel<<<1, 256, 0, st>>>(sel, hist, pass); ++g_b3d_launches;
  }
  pp_percentile_kernel<<<1, 32, 0, st>>>(sel, frac[0], frac[1], stats); ++g_b3d_launches;
  pp_moments_kernel<<<blocks, PP_THREADS, 0, st>>>(x, n, stats); ++g_b3d_launches;
  pp_moments_finalize_kernel<<<1, 32, 0, st>>>(stats, (double)n); ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// out = zoom_order1( (clip(x, p_lo, p_hi) - mean) / (std + 1e-8) ) with stats = {p_lo, p_hi, mean, std} from b3d_clip_stats.
// Replaces the clip / z-score / ndimage.zoom(order=1) / astype(float32) chain of _preprocess_image (training.py:120-132).
int b3d_zoom_normalize(const float* x, int D, int H, int W, const double* stats, float* out, int OD, int OH, int OW, void* stream) {
  B3D_REQUIRE(D >= 1 && H >= 1 && W >= 1 && OD >= 1 && OH >= 1 && OW >= 1, "zoom_normalize: bad shape");
  pp_zoom_norm_kernel<<<pp_blocks((long long)OD * OH * OW * 4), PP_THREADS, 0, (cudaStream_t)stream>>>(x, D, H, W, stats, out, OD, OH, OW);
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// labels: seg fp32 [D][H][W] with BraTS values {0,1,2,4} -> {0,1,2,3}, ndimage.zoom(order=0) to [OD][OH][OW]; out_dtype 0 =
// uint8, 1 = int64.  Replaces _preprocess_segmentation (training.py:134-146).
int b3d_zoom_labels(const float* seg, int D, int H, int W, void* out, int out_dtype, int OD, int OH, int OW, void* stream) {
  B3D_REQUIRE(out_dtype == 0 || out_dtype == 1, "zoom_labels: out_dtype must be 0 (uint8) or 1 (int64)");
  const int blocks = pp_blocks((long long)OD * OH * OW * 4);
  if (out_dtype == 0) pp_zoom_label_kernel<unsigned char><<<blocks, PP_THREADS, 0, (cudaStream_t)stream>>>(seg, D, H, W, (unsigned char*)out, OD, OH, OW);
  else pp_zoom_label_kernel<long long><<<blocks, PP_THREADS, 0, (cudaStream_t)stream>>>(seg, D, H, W, (long long*)out, OD, OH, OW);
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

// _apply_augmentations (training.py:148-172) with the random DECISIONS passed in: rot90 by k (0..3) in the (D,H) plane, flips
// along D/H/W, additive N(0, noise_std) noise from a counter RNG (seed), intensity scale.  img fp32 [C][D][H][W]; lab (optional)
// uint8 (lab_dtype 0) or int64 (1) [D][H][W].  k odd requires D == H.
int b3d_augment(const float* img, const void* lab, int lab_dtype, float* out_img, void* out_lab, int C, int D, int H, int W, int k,
                int flip_d, int flip_h, int flip_w, float noise_std, float scale, unsigned long long seed, void* stream) {
  B3D_REQUIRE(k >= 0 && k <= 3, "augment: k must be 0..3");
  B3D_REQUIRE(!(k & 1) || D == H, "augment: a 90/270 degree rotation in the (D,H) plane needs D == H (got %d, %d)", D, H);
  B3D_REQUIRE(img != out_img, "augment: in-place operation is not supported (gather)");
  const int blocks = pp_blocks((long long)D * H * W * 4);
  if (lab_dtype == 0)
    pp_augment_kernel<unsigned char><<<blocks, PP_THREADS, 0, (cudaStream_t)stream>>>(img, (const unsigned char*)lab, out_img, (unsigned char*)out_lab, C, D, H, W, k, flip_d, flip_h, flip_w, noise_std, scale, seed);
  else
    pp_augment_kernel<long long><<<blocks, PP_THREADS, 0, (cudaStream_t)stream>>>(img, (const long long*)lab, out_img, (long long*)out_lab, C, D, H, W, k, flip_d, flip_h, flip_w, noise_std, scale, seed);
  ++g_b3d_launches;
  B3D_CHECK_CUDA(cudaGetLastError());
  return B3D_OK;
}

}  // extern "C"
