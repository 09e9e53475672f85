// b3d_internal.h — host-side declarations shared by the translation units of libb3d.so (not part of the public ABI).
#pragma once
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>
// conv_zs.cu: z-marching, kd-stacked 3x3x3 convolution for large planes / few output channels.
// Returns 0 if launched, 1 if the shape does not suit it (use the block-mode kernel), negative on error.
int b3d_try_zs(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
                   int N, int D, int H, int W, int Cin, int Cout, double* stats, int cpg, int stats_groups,
                   int stats_batch, int* err_flag, cudaStream_t stream);
// conv_wg2.cu: 3x3x3 weight gradient with swizzled MN-major operands (W a multiple of 16).  Same return convention.
int b3d_try_wg2(const void* x, long long ldx, int Cin, const void* dy, long long lddy, int Cout_pad, int N, int D, int H, int W,
                float* dwacc, int Cin_pad, int* err_flag, cudaStream_t stream);
// conv_wgp.cu: streaming weight gradient of 1x1x1 convs (taps = 1) and ConvTranspose3d k2 s2 (taps = 8).  Same return convention.
int b3d_try_wgp(const void* x, long long ldx, int Cin, const void* dy, long long lddy, int Cout_pad, long long V, int taps,
                int ND, int Hc, int Wc, float* dwacc, int Cin_pad, int* err_flag, cudaStream_t stream);
