/* b3d.h — C ABI of libb3d.so: the B200 (sm_100a) hot path of the enhanced 3D U-Net.
 *
 * Drop-in boundary.  The reference (Ruhul-sde/Segmentation-and-classification-of-brain-tumor-using-3D-UNet) is pure
 * Python/PyTorch and has no FFI of its own: what it "binds" for this path are the ATen operators its nn.Modules
 * dispatch to.  Each entry point below names the reference call site (file:line under /root/reference) whose ATen
 * work it replaces.  Host code (segmentation-...-unet_b200/ops.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers unless noted.
 * Activations are NDHWC bf16 with a voxel pitch `ld*` in elements (so channel slices of a concat buffer are
 * addressable); logits are fp32 NCDHW; targets int64.  Every call is enqueued on `stream` (a cudaStream_t) and returns
 * immediately; 0 = success, negative = error (b3d_last_error_string()).  The library never allocates persistent device
 * memory, never synchronises, never throws and never exits.  Non-sm_100 devices are an error, not a fallback.
 */
#ifndef B3D_H_
#define B3D_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------------------------- */
int b3d_version(void);
const char* b3d_last_error_string(void);             /* thread-local */
int b3d_check_device(void);                          /* 0 iff current device is compute capability 10.x */
long long b3d_launch_count(void);                    /* kernels launched by this library so far (process-wide, atomic) */
/* Issue schedule of the z-marching 3x3x3 kernel.  0 (default): two ping-pong MMA-issuing warps — fastest, but tcgen05.mma is only
 * ordered within a thread, so the fp32 accumulation order jitters (~1e-7 of the bf16 outputs differ by an ulp between runs).
 * 2: ONE issuing thread fed by a scout warp that does the barrier waits and descriptor arithmetic -> forward pass and input
 * gradients bit-reproducible run to run, ~2 % slower per step.  1: one issuing warp without the scout (~13 % slower; kept for
 * comparison).  Default from env B3D_ORDERED_ISSUE.  Returns the previous mode.  Process-wide, atomic. */
int b3d_set_ordered_issue(int on);
/* Leave n SMs free in every grid this library launches from now on (all grids are sized from the SM count; the tensor-core
 * kernels are persistent one-CTA-per-SM grids).  For data-parallel training (training.py:290-304 under one process per GPU):
 * NCCL's channel CTAs cannot share an SM with a conv CTA, so a full-width conv grid launched while an all-reduce is running
 * waits for the collective; with n = NCCL_MAX_CTAS reserved both run side by side.  Default from env B3D_RESERVED_SMS (0).
 * Returns the previous reservation.  Process-wide, atomic; grids already captured into a CUDA graph keep their size. */
int b3d_set_reserved_sms(int n);

/* ---- convolutions on tcgen05 tensor cores (conv_igemm.cu, conv_wgrad.cu) ----------------------------------- */
/* weight repack fp32 reference layout -> bf16 [K/8][taps][rows][8].
 * mode 0: nn.Conv3d fprop, 1: Conv3d dgrad (flipped+transposed), 2: ConvTranspose3d fprop, 3: ConvTranspose3d dgrad */
int b3d_pack_weight(int mode, const float* w, int Cout, int Cin, int ntaps, void* out, int Kp, int rows, void* stream);
/* both packed copies of one weight from a single read (what a training step does after every optimizer update):
 * convT = 0: modes 0 + 1 of an nn.Conv3d weight [Cout][Cin][ntaps]; convT = 1: modes 2 + 3 of an nn.ConvTranspose3d(k2,s2)
 * weight [Cin][Cout][8] (main.py:121,130,216,219,229,252,258).  Buffers sized as for b3d_pack_weight. */
int b3d_pack_weight_pair(int convT, const float* w, int Cout, int Cin, int ntaps, void* out_fprop, void* out_dgrad,
                         void* stream);
/* One optimizer step of torch.optim.AdamW (training.py:186-191 builds it; training.py:296-304 steps it) for the whole model in
 * two kernel families: packed conv weights (parameter, both moments AND both bf16 packed copies in one pass — replaces the
 * re-pack after the step) and everything else.  HOST tables (int64): pack [n][16] = w g m v out_fprop out_dgrad A B T convT
 * Kp_f rows_f Kp_d rows_d first_tile tiles_b ; flat [n][8] = w g m v numel first_block - - .  lr / step: device floats. */
int b3d_adamw_step(const long long* pack_table, int n_pack, long long total_tiles, const long long* flat_table, int n_flat,
                   long long total_blocks, const float* lr, const float* step, double beta1, double beta2, double eps,
                   double weight_decay, void* stream);
/* nn.Conv3d(k=3,pad=1) main.py:130,216,219 and nn.Conv3d(k=1) main.py:229,252,258 (forward; with mode-1 weights: the
 * data gradient autograd computes for them).  Optional (+=) GroupNorm/BatchNorm partial sums of the output. */
int b3d_conv_fprop(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, void* y, long long ldy,
                   int N, int D, int H, int W, int Cin, int Cout, int ks, double* stats, int groups, int stats_batch,
                   void* ws, size_t ws_bytes, int* err_flag, void* stream);
/* same, with a fused  y = conv(x) + addend  (1x1x1 only; addend may alias y).  Replaces the separate accumulation autograd
 * performs when two branches feed the same tensor (residual 1x1 branch main.py:229-231,240; gate projections main.py:282-283) */
int b3d_conv_fprop_add(const void* x, long long ldx, const void* wpack, int w_rows, const float* bias, const void* addend,
                       long long ld_add, void* y, long long ldy, int N, int D, int H, int W, int Cin, int Cout, int ks,
                       double* stats, int groups, int stats_batch, void* ws, size_t ws_bytes, int* err_flag, void* stream);
/* nn.ConvTranspose3d(2f,f,k=2,s=2) main.py:121,183 forward / data gradient */
/* y = conv1x1(x) + addend with the addend carried on K through an identity block of the weights ([W ; I]): the fused
 * gradient accumulation of main.py:240 (residual branch) and main.py:297 (gate branches) in backward.  wpack_aug: bf16
 * [Cout][Cin + Cout] = [mode-1 packed W | identity]; y may alias addend. */
int b3d_conv1_add_mma(const void* x, long long ldx, const void* addend, long long ld_add, const void* wpack_aug, int w_rows,
                      void* y, long long ldy, int N, int D, int H, int W, int Cin, int Cout, int* err_flag, void* stream);
/* Scratch the caller must pass as (ws, ws_bytes) to the calls above / below (the library allocates nothing; SURVEY section 8b
 * "b3d_workspace_bytes_*").  conv_fprop: split-K partial slices, 0 = pass NULL (large problems never split). */
size_t b3d_conv_fprop_workspace_bytes(int N, int D, int H, int W, int Cout);
size_t b3d_convT2_dgrad_workspace_bytes(int N, int D, int H, int W, int Cin);
size_t b3d_conv_wgrad_workspace_bytes(int Cin, int Cout, int ks);
size_t b3d_convT2_wgrad_workspace_bytes(int Cin, int Cout);
/* tuning hooks of the implicit-GEMM tile planner (0 = planner's choice); not used by the product path */
int b3d_set_plan_override(int td, int th, int tw, int kc, int bn);
const char* b3d_last_plan(void);
int b3d_convT2_fprop(const void* x, long long ldx, const void* wpack, const float* bias, void* y, long long ldy, int N,
                     int D, int H, int W, int Cin, int Cout, int* err_flag, void* stream);
int b3d_convT2_dgrad(const void* dy, long long lddy, const void* wpack, int w_rows, void* dx, long long lddx, int N, int D,
                     int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes, int* err_flag, void* stream);
/* weight gradients (autograd convolution_backward of the same modules), fp32 in the reference layouts */
int b3d_conv_wgrad(const void* x, long long ldx, const void* dy, long long lddy, float* dw, int accumulate, int N, int D,
                   int H, int W, int Cin, int Cin_real, int Cout, int ks, float* ws, size_t ws_bytes, int* err_flag,
                   void* stream);
int b3d_convT2_wgrad(const void* x, long long ldx, const void* dy, long long lddy, float* dw, int accumulate, int N, int D,
                     int H, int W, int Cin, int Cout, float* ws, size_t ws_bytes, int* err_flag, void* stream);

/* ---- GroupNorm(+ReLU)(+residual) — nn.GroupNorm/nn.ReLU in DoubleConv3D, main.py:215-233,235-242 (norm.cu) -- */
int b3d_gn_apply(const void* y, long long ldy, const double* stats, const float* gamma, const float* beta, int G,
                 int relu, int res_mode, const void* r, long long ldr, const double* stats_r, const float* gamma_r,
                 const float* beta_r, int Gr, void* out, long long ldo, int N, long long V, int C, float eps, void* stream);
int b3d_gn_bwd_reduce(const void* dy, long long lddy, const void* y, long long ldy, const double* stats,
                      const float* gamma, const float* beta, int G, int relu, double* sums, int N, long long V, int C,
                      float eps, void* stream);
int b3d_gn_bwd_apply(const void* dy, long long lddy, const void* y, long long ldy, const double* stats,
                     const float* gamma, const float* beta, int G, int relu, const double* sums, void* dx,
                     long long lddx, int accumulate, int N, long long V, int C, float eps, void* stream);
/* dual backward of a residual block's tail  out = relu(GN_a(ya)) + GN_b(yb)  (main.py:238-240): one pass over dy for both
 * branches.  Returns 1 without launching when the shape does not suit it (use the two calls above per branch). */
int b3d_gn_bwd_dual(const void* dy, long long lddy, const void* ya, long long ldya, const double* stats_a,
                    const float* gamma_a, const float* beta_a, const void* yb, long long ldyb, const double* stats_b,
                    const float* gamma_b, int G, double* sums_a, double* sums_b, void* dxa, long long lddxa, void* dxb,
                    long long lddxb, int N, long long V, int C, float eps, void* stream);
int b3d_gn_param_grad(const double* sums, int N, int C, float* dgamma, float* dbeta, int accumulate, void* stream);
int b3d_add_bf16(const void* a, long long lda, const void* b, long long ldb, void* o, long long ldo, long long V, int C,
                 void* stream);

/* ---- MaxPool3d(2,2)+Dropout3d main.py:109-110,173-174 ; input staging training.py:287 (pool_layout.cu) ------- */
int b3d_pool_fwd(const void* x, long long ldx, const float* mask, void* out, long long ldo, int N, int D, int H, int W,
                 int C, void* stream);
/* BrainTumorClassifier.features (main.py:305-316): MaxPool3d(2)(ReLU(x)) and AdaptiveAvgPool3d((OD,OH,OW))(ReLU(x)) with the
 * result in fp32 NCDHW flatten order (what `x.view(x.size(0), -1)`, main.py:326, feeds the Linear layers) */
int b3d_relu_pool_fwd(const void* x, long long ldx, void* out, long long ldo, int N, int D, int H, int W, int C, void* stream);
int b3d_relu_adaptive_avgpool(const void* x, long long ldx, float* out, int N, int D, int H, int W, int C, int OD, int OH,
                              int OW, int relu, void* stream);
int b3d_pool_bwd(const void* x, long long ldx, const float* mask, const void* dy, long long lddy, void* dx,
                 long long lddx, int accumulate, int N, int D, int H, int W, int C, void* stream);
/* same, plus an optional per-(sample, channel) constant cadd fp32 [N][C] added to dx in the same pass (accumulate mode) */
int b3d_pool_bwd_add(const void* x, long long ldx, const float* mask, const void* dy, long long lddy, void* dx,
                     long long lddx, int accumulate, const float* cadd, int N, int D, int H, int W, int C, void* stream);
int b3d_to_ndhwc_bf16(const float* x, void* out, long long ldo, int N, int Cin, long long V, int Cpad, void* stream);
int b3d_to_ncdhw_f32(const void* x, long long ldx, float* out, int N, int C, long long V, void* stream);
int b3d_channel_sum(const void* x, long long ldx, double* sums, int N, long long V, int C, void* stream);

/* ---- AttentionGate3D main.py:244-299 (gate.cu) ---------------------------------------------------------------- */
int b3d_gate_psi_fwd(const void* g1r, const void* x1r, const double* st_g, const double* st_x, const float* gam_g,
                     const float* bet_g, const float* gam_x, const float* bet_x, const float* wpsi, const float* bpsi,
                     float* psi_raw, double* st_psi, int N, long long V, int F, float eps, void* stream);
int b3d_gate_se_fwd(const double* xsum, long long V, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* ca, float* zbuf, float* meanbuf, int N, int C, void* stream);
int b3d_gate_apply_fwd(const void* x, long long ldx, const float* psi_raw, const double* st_psi, const float* gpsi,
                       const float* bpsi_n, const float* ca, void* out, long long ldo, int N, long long V, int C,
                       float eps, void* stream);
int b3d_gate_apply_bwd(const void* dout, long long lddo, const void* x, long long ldx, const float* psi_raw,
                       const double* st_psi, const float* gpsi, const float* bpsi_n, const float* ca, void* dx,
                       long long lddx, float* dpsin, double* dca, double* st_dpsi, int N, long long V, int C, float eps,
                       void* stream);
int b3d_gate_se_bwd(const double* dca, const float* ca, const float* zbuf, const float* meanbuf, const float* w1,
                    const float* w2, long long V, float* dW1, float* db1, float* dW2, float* db2, float* xadd, int N, int C,
                    void* stream);
int b3d_gate_psi_bwd(const float* dpsin, const float* psi_raw, const double* st_psi, const double* st_dpsi,
                     const float* gpsi, const void* g1r, const void* x1r, const double* st_g, const double* st_x,
                     const float* gam_g, const float* bet_g, const float* gam_x, const float* bet_x, const float* wpsi,
                     void* dz, double* sums_g, double* sums_x, float* dwpsi, float* dbpsi, float* dgpsi, float* dbpsi_n,
                     int N, long long V, int F, float eps, void* stream);
int b3d_add_channel_const(void* dx, long long lddx, const float* xadd, int N, long long V, int C, void* stream);

/* ---- output heads: deep supervision main.py:137-140,164-171 ; final_conv main.py:129-134 (heads.cu) ----------- */
int b3d_ds_head_fwd(const void* x, long long ldx, const float* w, const float* b, float* out, long long NV, int C, int K,
                    void* stream);
int b3d_ds_head_bwd(const float* dl, const void* x, long long ldx, const float* w, void* dx, long long lddx,
                    int accumulate, float* dW, float* db, int N, long long Vs, int C, int K, void* stream);
/* same, logit gradient channel-last: dl float [N][Vs][4] (what b3d_dsloss_bwd produces) */
int b3d_ds_head_bwd_cl(const float* dl, const void* x, long long ldx, const float* w, void* dx, long long lddx,
                       int accumulate, float* dW, float* db, int N, long long Vs, int C, int K, void* stream);
int b3d_trilinear_up_fwd(const float* lo, float* out, int N, int Dl, int Hl, int Wl, int D, int H, int W, int K,
                         void* stream);
int b3d_trilinear_up_bwd(const float* dup, float* dlo, float* tmp, int N, int Dl, int Hl, int Wl, int D, int H, int W,
                         int K, void* stream);
int b3d_final_bn_prepare(const double* stats, double count, int train, float* running_mean, float* running_var,
                         long long* num_batches, float momentum, float eps, float* bn, int F2, int update_running,
                         void* stream);
int b3d_final_head_fwd(const void* h, long long ldh, const float* bn, const float* gamma, const float* beta,
                       const float* w2, const float* b2, float* out, int N, long long V, int F2, int K, void* stream);
int b3d_final_head_bwd(const float* dl, const void* h, long long ldh, const float* bn, const float* gamma,
                       const float* beta, const float* w2, double* red, int train, void* dh, long long lddh,
                       float* dgamma, float* dbeta, float* dW2, float* db2, int N, long long V, int F2, int K,
                       void* stream);

/* ---- losses losses.py:7-126, training.py:517-566 ; metrics training.py:351-364, main.py:470-474 (loss.cu) ----- */
int b3d_loss_fwd(const float* logits, const long long* target, const float* cfg11, float* prob, float* E, double* acc,
                 float* values, int N, int K, int D, int H, int W, void* stream);
int b3d_loss_bwd(const float* prob, const float* E, const long long* target, const double* acc, const float* cfg11,
                 const float* gscale, float wscale, float* dlogits, int N, int K, int D, int H, int W, void* stream);
/* ---- fused deep-supervision loss (dsloss.cu): F.interpolate(trilinear) main.py:165-170 + CombinedLoss3D.forward
 *      losses.py:63-75 as evaluated per output by DeepSupervisionLoss3D.forward losses.py:107-126, computed from the LOW-RES
 *      head logits (float4 [N][D/s][H/s][W/s], s in {1,2,4,8}) without materialising the up-sampled map.
 *      target_u8: labels as uint8 [N][D][H][W] (b3d_target_u8; anything outside 0..3 -> 255 = matches no class).
 *      dlo: s == 1 float [N][V][4]; s > 1 double [N][V/s^3][4] (zeroed inside, fp64 atomics). ------------------------- */
int b3d_target_u8(const long long* target, unsigned char* out, long long count, void* stream);
int b3d_dsloss_fwd(const float* lo, const unsigned char* target_u8, const float* cfg11, double* acc, float* values, int N,
                   int scale, int D, int H, int W, void* stream);
int b3d_dsloss_bwd(const float* lo, const unsigned char* target_u8, const double* acc, const float* cfg11, const float* gscale,
                   float wscale, void* dlo, int N, int scale, int D, int H, int W, void* stream);
int b3d_confusion(const float* logits, const long long* target, unsigned char* mask, unsigned long long* hist, int N,
                  int K, long long V, void* stream);
int b3d_voxel_counts(const unsigned char* mask, long long V, int W, unsigned long long* cls, unsigned long long* slices,
                     void* stream);

/* ---- input pipeline in front of the path (preprocess.cu): BraTSDataset._preprocess_image training.py:117-132,
 *      _preprocess_segmentation training.py:134-146, _apply_augmentations training.py:148-172 ----------------------------- */
/* exact np.percentile(x, (q_lo, q_hi)) by radix select + mean / population std of the clipped values; stats <- double[4]
 * {p_lo, p_hi, mean, std}; work: >= 4*2048*4 + 64 bytes of device scratch */
int b3d_clip_stats(const float* x, long long n, double q_lo, double q_hi, void* work, size_t work_bytes, double* stats,
                   void* stream);
/* out = ndimage.zoom(order=1)( (clip(x) - mean) / (std + 1e-8) ) as float32 [OD][OH][OW] */
int b3d_zoom_normalize(const float* x, int D, int H, int W, const double* stats, float* out, int OD, int OH, int OW, void* stream);
/* label 4 -> 3 and ndimage.zoom(order=0); out_dtype 0 = uint8, 1 = int64 */
int b3d_zoom_labels(const float* seg, int D, int H, int W, void* out, int out_dtype, int OD, int OH, int OW, void* stream);
/* rot90 by k in the (D,H) plane, flips, + N(0, noise_std) (counter RNG), * scale; lab optional (lab_dtype 0 uint8 / 1 int64) */
int b3d_augment(const float* img, const void* lab, int lab_dtype, float* out_img, void* out_lab, int C, int D, int H, int W, int k,
                int flip_d, int flip_h, int flip_w, float noise_std, float scale, unsigned long long seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B3D_H_ */
