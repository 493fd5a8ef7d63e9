/*
 * sei_b200.h -- C ABI of libsei_b200.so: the B200 (sm_100a) implementation of the
 * Scale-Equivariant-Imaging per-step hot path (degradation physics forward/adjoint,
 * random scale transform, loss reductions).
 *
 * The reference (jscanvic/Scale-Equivariant-Imaging, pure Python/PyTorch) has no FFI; its
 * boundary for this path is the Python object protocol of src/physics, src/transforms.py
 * and src/losses.  Each entry point below replaces the torch
 * library call(s) cited next to it, and is what the reference-side binding in
 * INTEGRATION.md (a ctypes stub inside the reference's own classes) would bind.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  All image tensors are float32,
 *    NCHW, contiguous; `planes` = B*C independent H x W planes.
 *  - the caller owns every buffer (inputs, outputs, workspaces).  The library never
 *    allocates or frees device memory and keeps no pointer after a call returns.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL =
 *    the legacy default stream) and is CUDA-graph capturable.
 *  - return value: 0 on success; SEI_EINVAL for an unsupported argument; otherwise a
 *    positive cudaError_t.  sei_last_error() gives a thread-local message.  There is no
 *    CPU fallback: without a CUDA device every compute call fails.
 *  - pointers named *_host are read on the host during the call; all others are
 *    device pointers.
 */
#ifndef SEI_B200_H
#define SEI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SEI_ABI_VERSION 1
#define SEI_EINVAL (-22)

/* path selectors for the operators that have a tiled (TMA row-band) and a direct kernel */
#define SEI_PATH_AUTO 0
#define SEI_PATH_DIRECT 1
#define SEI_PATH_TILED 2

int sei_abi_version(void);
const char* sei_last_error(void);
/* name of the kernel variant the last call on this thread launched ("" if none) */
const char* sei_last_kernel(void);
/* number of kernels this library has launched from this process (all threads) */
long long sei_launch_count(void);
/* 0 if a CUDA device with compute capability 10.x is current; fills the out-params */
int sei_device_info(int* sm_count, int* smem_per_block_optin, int* cc_major, int* cc_minor);

/* ---- degradation physics -------------------------------------------------------------
 * Circular blur: BlurV2.A (reference src/physics/blur/__init__.py:205-223, the rfft2 /
 * multiply / irfft2 sequence) and Blur(padding="circular").A (:34-74, the per-(b,c)
 * F.conv2d loop):           y[n] = sum_i h[i] x[(n - i + k//2) mod N]  on both axes.
 * adjoint != 0: the transpose, i.e. autograd backward of A, BlurV2.A_adjoint (:225-227)
 * and conv_transpose(.., "circular") (:77-134).
 * kernel_host: kh x kw taps, row-major, double (the reference keeps the kernel in
 * float64 and casts to the image dtype per call).  If noise != NULL the deepinv
 * GaussianNoise step attached at src/physics/__init__.py:53 is fused:
 * y += sigma * noise (noise: device, same shape as y, standard normal draws).
 * Requires H >= kh and W >= kw (the reference fails otherwise as well). */
int sei_blur_circular_f32(const float* x, float* y, long long planes, int H, int W,
                          const double* kernel_host, int kh, int kw, int adjoint,
                          const float* noise, float sigma, int path, void* stream);

/* v1 operator with the non-default paddings (src/physics/blur/__init__.py:34-161).
 * mode: 0 valid, 1 circular, 2 replicate, 3 reflect, 4 zero (transpose only).
 * transpose == 0: conv(x, filter, padding);  != 0: conv_transpose(y, filter, padding).
 * H, W are the INPUT plane sizes; the output plane size is written to Ho/Wo
 * (call with out == NULL to query the size only). */
int sei_blur_padded_f32(const float* in, float* out, long long planes, int H, int W,
                        const double* filter_host, int fh, int fw, int mode, int transpose,
                        int* Ho, int* Wo, void* stream);

/* SR forward: Downsampling.A (src/physics/downsampling/__init__.py:16-19) =
 * F.interpolate(x, scale_factor=1/rate, mode="bicubic", antialias=True).
 * x: planes x H x W  ->  y: planes x floor(H/rate) x floor(W/rate).  Optional fused noise. */
int sei_down_aa_f32(const float* x, float* y, long long planes, int H, int W, int rate,
                    const float* noise, float sigma, int path, void* stream);
/* transpose of the above (autograd backward of A; true_adjoint=True, :21-31).
 * gy: planes x floor(H/rate) x floor(W/rate)  ->  gx: planes x H x W. */
int sei_down_aa_transpose_f32(const float* gy, float* gx, long long planes, int H, int W,
                              int rate, int path, void* stream);
/* Downsampling.A_adjoint with true_adjoint=False (:32-35): plain bicubic upsample,
 * F.interpolate(y, scale_factor=rate, mode="bicubic").  y: planes x h x w -> planes x h*rate x w*rate */
int sei_up_bicubic_f32(const float* y, float* x, long long planes, int h, int w, int rate,
                       void* stream);

/* Paired random crops of a batch (reference src/crop.py:15-39 CropPair, called per dataset item at
 * src/datasets/__init__.py:78-90 through torchvision TF.crop): out[b, c, i, j] = in[b, c, top[b] + i, left[b] + j],
 * zero outside the image.  top / left: DEVICE int arrays of B offsets. */
int sei_crop_batch_f32(const float* in, float* out, int B, int C, int H, int W, int h, int w,
                       const int* top, const int* left, void* stream);

/* ---- scale transform -----------------------------------------------------------------
 * padded_downsampling_transform (src/transforms.py:60-83) with mode="bicubic",
 * padding_mode="reflection", antialiased=False: builds the grid of
 * get_downsampling_grid (:27-43) analytically and applies
 * F.grid_sample(align_corners=True).  x, out: B x C x S x S (square only, like the
 * reference).  rate: B floats, center: B x 2 floats (cx, cy), both DEVICE pointers
 * (they are drawn on the device; no host sync). */
int sei_scale_transform_f32(const float* x, float* out, int B, int C, int S,
                            const float* rate, const float* center, int path, void* stream);
/* The grid_sample step of the same transform behind its anti-aliasing pre-filter
 * (src/transforms.py:63-67,76-82, antialiased=True): x is the pre-filtered image,
 * B x C x Ssrc x Ssrc (alias_free_interpolate :44-57 = sei_resize_bicubic_f32 with
 * antialias), out is B x C x S x S, the grid is the one of an S x S image. */
int sei_scale_transform_src_f32(const float* x, float* out, int B, int C, int Ssrc, int S,
                                const float* rate, const float* center, void* stream);

/* gx [B, C, Ssrc, Ssrc] = the transpose of sei_scale_transform_src_f32 applied to gout [B, C, S, S] (autograd through
 * grid_sample in the anti-aliased padded transform, reference src/transforms.py:60-83). */
int sei_scale_transform_src_backward_f32(const float* gout, float* gx, int B, int C, int Ssrc, int S, const float* rate,
                                         const float* center, void* stream);
/* transpose of the above w.r.t. x (autograd through grid_sample; only reached with
 * --no-ProposedLoss__stop_gradient).  gx is overwritten. */
int sei_scale_transform_backward_f32(const float* gout, float* gx, int B, int C, int S,
                                     const float* rate, const float* center, void* stream);
/* sample_from + sample_downsampling_parameters (src/transforms.py:5-24) given the two
 * uniform draws u_rate (B) and u_center (B x 2):
 *   rate_b = rates[floor(n_rates * u_rate_b)],  center_b = 2 * u_center_b - 1. */
int sei_scale_params_f32(const float* u_rate, const float* u_center, int B,
                         const float* rates_host, int n_rates, float* rate, float* center,
                         void* stream);

/* Fused EI re-measurement (deepinv EILoss.forward as built at
 * src/losses/__init__.py:117-122, steps x2 = T(x_net); y = physics(x2)):
 *   x2 = scale_transform(x_net),  y_out = A(x2) + sigma * noise
 * in one kernel, so x2 is written once and never re-read from HBM.
 * Deblurring: rate_sr = 1 and kernel_host != NULL (circular blur).
 * SR: rate_sr in {2,3,4} and kernel_host == NULL (antialiased bicubic decimation).
 * x_net, x2: B x C x S x S;  y_out, noise: B x C x S/rate_sr x S/rate_sr.
 * workspace: optional device scratch of sei_ei_workspace_bytes(B, S) bytes (16-byte aligned); with it the
 * resampling taps of every row and column are computed once per image by a small pre-kernel instead of once
 * per band inside the fused kernel.  NULL is allowed. */
long long sei_ei_workspace_bytes(int B, int S);
int sei_ei_remeasure_f32(const float* x_net, float* x2, float* y_out, int B, int C, int S,
                         const float* rate, const float* center,
                         const double* kernel_host, int kh, int kw, int rate_sr,
                         const float* noise, float sigma, void* workspace, void* stream);

/* ---- loss reductions -----------------------------------------------------------------
 * All reductions are deterministic (fixed-order two-stage tree, double accumulation of
 * the per-block partials).  `workspace` must hold sei_reduce_workspace_bytes() bytes. */
long long sei_reduce_workspace_bytes(void);

/* nn.MSELoss (deepinv metric.mse; EILoss / SupLoss):  out[0] = mean((a-b)^2) */
int sei_mse_f32(const float* a, const float* b, long long n, float* out, void* workspace,
                void* stream);
/* backward: ga = gscale[0] * 2/n * (a-b);  gb = -ga if gb != NULL.  gscale: device scalar */
int sei_mse_backward_f32(const float* a, const float* b, long long n, const float* gscale,
                         float* ga, float* gb, void* stream);

/* SureGaussianLoss.forward + mc_div (src/losses/sure.py:7-76) given y1 = A(x_net),
 * y2 = A(model(y + tau*b)):
 *   mse = mean_{interior(margin_mse)} (y1-y)^2
 *   div = mean_{interior(margin_div)} b*(y2-y1)/tau
 *   out[0] = mse + 2*sigma2*div - (averaged_cst ? sigma2 : sigma2/B); out[1] = mse; out[2] = div */
int sei_sure_loss_f32(const float* y1, const float* y2, const float* y, const float* b,
                      int B, int C, int H, int W, int margin_mse, int margin_div,
                      float tau, float sigma2, int averaged_cst, float* out, void* workspace,
                      void* stream);
/* backward: g1 = d loss / d y1, g2 = d loss / d y2 (scaled by gscale[0]) */
int sei_sure_loss_backward_f32(const float* y1, const float* y, const float* b,
                               int B, int C, int H, int W, int margin_mse, int margin_div,
                               float tau, float sigma2, const float* gscale,
                               float* g1, float* g2, void* stream);

/* mc_div's probe (sure.py:8-24): out = y + tau * b, where b is `draw` placed in the
 * interior (margin wide border of zeros) -- draw: B x C x (H-2m) x (W-2m) (or y-shaped if
 * margin == 0).  b_out (optional) receives the zero-bordered b. */
int sei_sure_perturb_f32(const float* y, const float* draw, int B, int C, int H, int W,
                         int margin, float tau, float* out, float* b_out, void* stream);

/* torch.roll(x, (shift_h, shift_w), (-2, -1)) per plane: the action of deepinv's Shift transform
 * (ProposedLoss__transforms=Shifts, src/losses/__init__.py:91-94) */
int sei_roll_f32(const float* in, float* out, long long planes, int H, int W, int shift_h, int shift_w,
                 void* stream);

/* deepinv GaussianNoise.forward: out = y + sigma * noise */
int sei_add_noise_f32(const float* y, const float* noise, long long n, float sigma, float* out,
                      void* stream);

/* ---- restoration CNN: dense contractions ---------------------------------------------------
 * bf16 x bf16 -> fp32-accumulate GEMM on tcgen05 tensor cores (TMEM accumulator, TMA operand tiles):
 *     D[M, N] = A[M, K] * B[N, K]^T (+ bias[N])
 * A, B: bf16 row-major with leading dimensions lda, ldb (elements, multiples of 8); D: bf16
 * (out_f32 == 0) or fp32 row-major with leading dimension ldd; bias: fp32 or NULL.
 * Replaces the cuDNN/cuBLAS calls behind nn.Conv2d(kernel_size=1) of the reference's
 * ConvolutionalModel (src/models/convolutional.py:40-42,106,143) on channels-last activations
 * (M = B*H*W pixels, K = C_in, N = C_out, B = weight); dgrad and wgrad use the same entry with the
 * operand roles permuted.  tile_n: 0 = auto, or 32 / 64 / 128 / 256. */
int sei_gemm_bf16_tn(const void* A, const void* B, void* D, const float* bias, long long M, int N, int K,
                     long long lda, long long ldb, long long ldd, int out_f32, int tile_n, void* stream);

/* D (bf16) = (A B^T) * gelu'(H), element-wise in the epilogue: the input gradient of the convolution that FOLLOWS the
 * GELU of a ConvBlock (reference src/models/convolutional.py:40-42: conv2 -> gelu -> conv3) with the GELU backward
 * fused in, H = the saved pre-activation [M, N] (bf16, row pitch ld_h).  N % 64 == 0, ldd % 8 == 0. */
int sei_gemm_bf16_tn_gelu_bwd(const void* A, const void* B, void* D, const void* H, long long M, int N, int K,
                              long long lda, long long ldb, long long ldd, long long ld_h, void* stream);

/* D (bf16) = A B^T + bias + res_scale * R, the addition done in the GEMM epilogue: a pointwise convolution whose output
 * is added to a tensor of the same shape -- ConvBlock's `return x + x1` (reference src/models/convolutional.py:43-51),
 * UNet's inner-residual and skip additions (:203-215: `x = x + xb`, `x = x + skips.pop()`).  R: bf16 [M, N] with row
 * pitch ld_r (multiple of 8); N % 8 == 0. */
int sei_gemm_bf16_tn_residual(const void* A, const void* B, void* D, const float* bias, const void* R, float res_scale,
                              long long M, int N, int K, long long lda, long long ldb, long long ldd, long long ld_r,
                              void* stream);

/* D (bf16) = A B^T + bias[n] * row_scale[m % period]: Downsample's pointwise convolution (reference
 * src/models/convolutional.py:136-150) applied AFTER the ideal resampler (the two commute); the constant image bias[n]
 * turns into bias[n] * R(1)[pixel], a per-row factor added in the GEMM epilogue.  row_scale: fp32 [period] (device). */
int sei_gemm_bf16_tn_rowscaled_bias(const void* A, const void* B, void* D, const float* bias, const float* row_scale,
                                    int period, long long M, int N, int K, long long lda, long long ldb, long long ldd,
                                    void* stream);

/* D (bf16) = A Bkn (optionally * Mult element-wise): B given as [K, N] row-major and read IN PLACE as an MN-major UMMA
 * operand by the CTA-pair kernel.  The input gradient of a pointwise convolution (autograd of nn.Conv2d(kernel_size=1),
 * reference src/models/convolutional.py:40-42,106,143): gx = gy W with W the (C_out x C_in) weight itself, so no
 * transposed copy of the weights exists.  M >= 256, N >= 256, N % 8 == 0; Mult: bf16 [M, N] (row pitch ld_m) or NULL. */
int sei_gemm_bf16_nn(const void* A, const void* Bkn, const void* Mult, void* D, long long M, int N, int K, long long lda,
                     long long ldb, long long ldd, long long ld_m, void* stream);

/* Aout (bf16) = gelu(A B^T + bias) and Dout (bf16) = gelu'(A B^T + bias), both written from the GEMM epilogue:
 * ConvBlock.conv2 followed by ConvBlock.gelu (reference src/models/convolutional.py:40-41, 46-47) without the
 * pre-activation ever reaching memory.  The pre-activation is rounded to bf16 before the GELU (the value an unfused
 * bf16 pipeline would store), erf-form GELU.  N > 32; Aout / Dout: row pitch ldo (multiple of 8). */
int sei_gemm_bf16_tn_gelu_dual(const void* A, const void* B, const float* bias, void* Aout, void* Dout, long long M, int N,
                               int K, long long lda, long long ldb, long long ldo, void* stream);

/* D (bf16) = (A B^T) * Mult element-wise in the epilogue: the input gradient of ConvBlock.conv3 (:42) times the stored
 * gelu' -- the backward of ConvBlock.gelu without a pass of its own.  Mult: bf16 [M, N], row pitch ld_m. */
int sei_gemm_bf16_tn_mul(const void* A, const void* B, const void* Mult, void* D, long long M, int N, int K,
                         long long lda, long long ldb, long long ldd, long long ld_m, void* stream);

/* D[M, N] (fp32) = A[K, M]^T * B[K, N]: both operands are read with the contraction index as their ROW (UMMA MN-major
 * shared-memory layout), so the weight gradient dL/dW = (dL/dy)^T x of a pointwise convolution needs no transposed
 * copies of the activations.  lda, ldb multiples of 8; K = pixels.  Split-K with fp32 atomics when M*N is small. */
int sei_gemm_bf16_atb(const void* A, const void* B, float* D, long long K, int M, int N,
                      long long lda, long long ldb, void* stream);
/* D += A^T B (same operands): the weight gradient accumulated in place into the parameter's fp32 gradient buffer. */
int sei_gemm_bf16_atb_accumulate(const void* A, const void* B, float* D, long long K, int M, int N,
                                 long long lda, long long ldb, void* stream);

/* Batched operator product  D_b[M, N] = A[M, K] * X_b[K, N]  (bf16, fp32 accumulation, bf16 result), A shared by
 * all batch entries and resident in shared memory, X streamed once.  Replaces the rfft2 / fftshift / mask or
 * zero-pad / irfft2 chains of the reference's IdealDownsample / IdealUpsample (src/models/convolutional.py
 * :54-92,113-133) on channels-last activations: those chains are fixed linear maps, applied here as two such
 * products (width then height; two terms because the reference shifts the half-spectrum axis).
 *   A: [ceil(M / tile_rows) * tile_rows][Kpad] row-major, zero padded; Kpad a multiple of 64;
 *      tile_rows = sei_bgemm_tile_rows(M, Kpad).
 *   batch entry bt = bo * b_inner + bi starts at X + bo*x_bo + bi*x_bi and D + bo*d_bo + bi*d_bi;
 *   row k of X_b = ko * k_inner + ki lives at ko*x_ko + ki*x_ki, row m of D_b = mo * m_inner + mi at mo*d_mo + mi*d_mi;
 *   the N index is contiguous; N and all strides are multiples of 8 elements. */
int sei_bgemm_tile_rows(int M, int Kpad);
int sei_bgemm_bf16(const void* A, const void* X, void* D, int M, int K, int N, int Kpad, int tile_rows,
                   long long batches, int b_inner, long long x_bo, long long x_bi,
                   int k_inner, long long x_ko, long long x_ki,
                   long long d_bo, long long d_bi, int m_inner, long long d_mo, long long d_mi, void* stream);

/* Channel LayerNorm of the reference's CNN (src/models/convolutional.py:21-30: swapaxes, nn.LayerNorm(C, eps), swapaxes)
 * on channels-last bf16 activations viewed as rows [T = B*H*W, C] (C % 8 == 0): statistics and arithmetic in fp32.
 * forward: y, and mean / rstd per row (saved for the backward).  backward: dx (bf16) and dgamma / dbeta (fp32, fixed
 * summation order); workspace of sei_ln_cl_backward_workspace_bytes(C) bytes (-1: channel count unsupported). */
int sei_ln_cl_forward_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                           long long T, int C, float eps, void* stream);
long long sei_ln_cl_backward_workspace_bytes(int C);
int sei_ln_cl_backward_bf16(const void* gy, const void* x, const float* mean, const float* rstd, const float* gamma,
                            void* dx, float* dgamma, float* dbeta, void* workspace, long long T, int C, void* stream);

/* out[c] (fp32) = sum over the T rows of x[T, C] (bf16, C % 8 == 0), fixed summation order: the bias gradient of the
 * reference's pointwise convolutions (autograd of nn.Conv2d bias).  workspace: sei_ln_cl_backward_workspace_bytes(C). */
int sei_colsum_bf16(const void* x, float* out, void* workspace, long long T, int C, void* stream);

/* 3x3 'same' convolution as an IMPLICIT GEMM on the tcgen05 tensor cores: UNet.in_conv / UNet.out_conv of the reference
 * (src/models/convolutional.py:175-176) and their input gradients.  x: bf16 [B, H, W, Cin] channels-last with Cin = 8
 * (3 image channels zero-padded) or 32; the nine shifted windows of a 16 x 8 pixel tile are bulk tensor copies whose
 * out-of-image coordinates are zero-filled (the 'same' padding) and are read in place as the K-chunks of the A operand.
 * wg: the weights in chunk form, bf16 [NCHP][N][8] with N = 32 (Cin = 8: NCHP = 10) or N = 16 (Cin = 32: NCHP = 36),
 * element [tap * (Cin / 8) + block][n][j] = w[n][8 * block + j][ky][kx], tap = 3 ky + kx, zero where padded.
 * out: bf16 [B, H, W, out_stride], out_stride = 32 (Cin = 8) or 4 (Cin = 32; channels >= out_valid are written as zero).
 * bias: fp32 [out_valid] or NULL. */
int sei_conv3x3_igemm_bf16(const void* x, const void* wg, const float* bias, void* out, int B, int H, int W, int Cin,
                           int out_stride, int out_valid, void* stream);

/* 3x3 convolution, stride 1, zero "same" padding, with 1..4 output channels, on a channels-last bf16 input
 * [B, H, W, Cin] (Cin a multiple of 8, <= 64): the reference's UNet.out_conv (src/models/convolutional.py:176,
 * Conv2d(hidden, in_channels, kernel_size=3, padding="same")).  w: [Cout, Cin, 3, 3] fp32, bias: [Cout] fp32 or NULL.
 * forward: y [B, H, W, 4] bf16 (channels >= Cout are zero).  backward: gy [B, H, W, 4] bf16 -> gx [B, H, W, Cin] bf16
 * (skipped when gx == NULL; Cin in {8, 16, 32, 64}), gw [Cout, Cin, 3, 3] and gb [Cout] fp32 in a fixed summation
 * order; workspace of sei_conv3x3_small_workspace_bytes(Cin, Cout) bytes. */
long long sei_conv3x3_small_workspace_bytes(int Cin, int Cout);
int sei_conv3x3_small_forward_bf16(const void* x, const float* w, const float* bias, void* y,
                                   int B, int H, int W, int Cin, int Cout, void* stream);
int sei_conv3x3_small_backward_bf16(const void* gy, const void* x, const float* w, void* gx, float* gw, float* gb,
                                    void* workspace, int B, int H, int W, int Cin, int Cout, void* stream);

/* Depthwise 7x7 convolution, padding 3, on channels-last bf16 activations [B, H, W, C] (C % 8 == 0): the reference's
 * ConvBlock.conv1 (src/models/convolutional.py:36-38, Conv2d(dim, dim, 7, padding=3, groups=dim)).
 * sei_dwconv7_cl_bf16: y = conv(x) (+ bias); wt = taps as [49][C] fp32 (tap-major).  Called with the taps flipped and
 * bias == NULL it is the input gradient.  sei_dwconv7_wgrad_cl_bf16: gw [C, 7, 7] and gb [C] fp32, fixed summation
 * order; workspace of sei_dwconv7_workspace_bytes(C) bytes (-1: channel count unsupported). */
long long sei_dwconv7_workspace_bytes(int C);
int sei_dwconv7_cl_bf16(const void* x, const float* wt, const float* bias, void* y, int B, int H, int W, int C, void* stream);

/* y = depthwise7x7(x) (+ bias) + res_scale * res, the addition done in the store of the convolution: the input
 * gradient of a ConvBlock (reference src/models/convolutional.py:43-51, `return x + x1`: autograd adds the incoming
 * gradient to the one that went through the block).  res: bf16 [B, H, W, C]. */
int sei_dwconv7_cl_residual_bf16(const void* x, const float* wt, const float* bias, const void* res, float res_scale,
                                 void* y, int B, int H, int W, int C, void* stream);
int sei_dwconv7_wgrad_cl_bf16(const void* gy, const void* x, float* gw, float* gb, void* workspace,
                              int B, int H, int W, int C, void* stream);

/* Exact (erf-form) GELU of the reference's ConvBlock (src/models/convolutional.py:41, nn.GELU()) on n bf16 elements
 * (n % 8 == 0): out = gelu(x) when gy == NULL, else out = gy * gelu'(x) (the backward of the same layer). */
int sei_gelu_bf16(const void* x, const void* gy, void* out, long long n, void* stream);

/* gx = gy * gelu'(h) on rows [T, C] and, in the same pass, gb[c] = sum_t gx[t, c]: GELU's backward (reference
 * src/models/convolutional.py:41) together with the bias gradient of the pointwise convolution in front of it (:40,
 * ConvBlock.conv2).  workspace: sei_ln_cl_backward_workspace_bytes(C) bytes. */
int sei_gelu_bwd_colsum_bf16(const void* h, const void* gy, void* gx, float* gb, void* workspace, long long T, int C,
                             void* stream);

/* One Adam update (torch.optim.Adam semantics without weight decay / amsgrad; reference demo/train.py:167-186,266) of
 * n float32 parameters: m, v updated in place, p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), with the
 * step count t read from device memory (graph-capturable).  lowp_bf16 (optional): bf16 copy of the updated
 * parameters written in the same pass (the operand the tensor-core GEMMs read). */
int sei_adam_step_f32(float* p, const float* g, float* m, float* v, void* lowp_bf16, const float* step,
                      long long n, float lr, float beta1, float beta2, float eps, void* stream);
/* out[c][r] = in[r][c] for a bf16 [rows, cols] matrix (the (K, N) weight copy of the input-gradient GEMMs). */
int sei_transpose_bf16(const void* in, void* out, int rows, int cols, void* stream);

/* Bias of a pointwise convolution pushed through an ideal resampler (Downsample applies the resampler first):
 * out[t][c] += pat[t % period] * bias[c] in place on bf16 rows [T, C]; and its gradient
 * gbias[c] = sum_t gy[t][c] * pat[t % period] (fixed order; workspace as for sei_colsum_bf16). */
int sei_bias_pattern_add_bf16(void* out, const float* pat, const float* bias, long long T, int C, int period, void* stream);
int sei_bias_pattern_grad_bf16(const void* gy, const float* pat, float* gbias, void* workspace, long long T, int C, int period,
                               void* stream);

/* The same channel LayerNorm for 1..32 channels of any count (one thread per row): the 3-channel normalisation of the SR
 * model's input stage (reference Upsample(in_channels=3), src/models/convolutional.py:95-104).  Same semantics as
 * sei_ln_cl_*; workspace of sei_ln_small_workspace_bytes(C) bytes. */
long long sei_ln_small_workspace_bytes(int C);
int sei_ln_small_forward_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                              long long T, int C, float eps, void* stream);
int sei_ln_small_backward_bf16(const void* gy, const void* x, const float* mean, const float* rstd, const float* gamma,
                               void* dx, float* dgamma, float* dbeta, void* workspace, long long T, int C, void* stream);

/* Bicubic resize by a fractional factor: y[planes, Ho, Wo] from x[planes, H, W], scale_* = 1 / scale_factor.  Replaces the
 * per-image F.interpolate(x_i, scale_factor=rate, mode="bicubic", antialias=...) loop of the reference's
 * normal_downsampling_transform (src/transforms.py:112-124); antialias selects ATen's _upsample_bicubic2d_aa weights. */
int sei_resize_bicubic_f32(const float* x, float* y, long long planes, int H, int W, int Ho, int Wo,
                           float scale_h, float scale_w, int antialias, void* stream);

/* gx[planes, H, W] = the transpose of sei_resize_bicubic_f32 applied to gy[planes, Ho, Wo]: autograd through
 * F.interpolate in normal_downsampling_transform / alias_free_interpolate (reference src/transforms.py:44-57,112-124)
 * when the EI branch keeps its gradient (src/losses/__init__.py:84-96, --no-ProposedLoss__stop_gradient). */
int sei_resize_bicubic_backward_f32(const float* gy, float* gx, long long planes, int H, int W, int Ho, int Wo,
                                    float scale_h, float scale_w, int antialias, void* stream);
/* deepinv.transform.Rotate (third-party, v0.2.0; used at src/losses/__init__.py:86-91): torchvision
 * transforms.functional.rotate(x, angle) with its defaults = grid_sample(mode="nearest", padding_mode="zeros",
 * align_corners=False) on torchvision's affine grid.  rescaled_theta: HOST pointer to the 3 x 2 fp32 matrix
 * theta^T / [W/2, H/2] torchvision builds (row-major). */
int sei_rotate_nearest_f32(const float* x, float* y, long long planes, int H, int W, const float* rescaled_theta,
                           void* stream);

/* gx = the transpose of sei_rotate_nearest_f32 applied to gy (autograd through deepinv's Rotate -> torchvision rotate ->
 * grid_sample(nearest) when the EI branch keeps its gradient). */
int sei_rotate_nearest_backward_f32(const float* gy, float* gx, long long planes, int H, int W, const float* rescaled_theta,
                                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEI_B200_H */
