/*
 * sei_oracle_impl.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's algorithms for the hot path
 * (jscanvic/Scale-Equivariant-Imaging, /root/reference).  Included twice by
 * sei_oracle.c, once with REAL=float (mimics the reference run in fp32: every
 * elementwise step is rounded to fp32 where torch rounds) and once with
 * REAL=double (matches the reference run in fp64 to ~1e-13).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call this library; the product path (scale-equivariant-imaging_b200/) never
 * does.  Parity pin: tests/test_oracle_golden.py checks every function below against the
 * .npz fixtures in tests/golden, which were produced by running the reference itself
 * (tests/golden/make_golden.py).
 *
 * Where the reference delegates to a torch library call, the library's documented
 * algorithm is restated (ATen UpSampleKernel / GridSamplerKernel semantics), anchored on
 * the reference's call sites:
 *   F.interpolate(bicubic, antialias=True)       src/physics/downsampling/__init__.py:16-19
 *   F.interpolate(bicubic)                       src/physics/downsampling/__init__.py:34
 *   F.grid_sample(bicubic, reflection, ac=True)  src/transforms.py:77-83
 *   torch.fft circular convolution               src/physics/blur/__init__.py:205-223
 */

#ifndef REAL
#error "define REAL and FN before including"
#endif

/* ------------------------------------------------------------------------------------
 * Circular blur.  BlurV2.A (src/physics/blur/__init__.py:205-223): the PSF is the kernel
 * zero-padded to HxW and rolled by -(k//2), multiplied in the Fourier domain, i.e. the
 * circular CONVOLUTION   y[n] = sum_i h[i] * x[(n - i + k//2) mod N]   on both axes.
 * Blur(padding="circular").A (same file :34-74, :190-191) is the same map for odd k.
 * adjoint != 0 gives the transpose (circular CORRELATION): BlurV2.A_adjoint (:225-227,
 * a vjp of A), autograd backward of A, and conv_transpose(..., "circular") (:77-134).
 *   xbar[m] = sum_i h[i] * ybar[(m + i - k//2) mod N]
 * ---------------------------------------------------------------------------------- */
void FN(orc_blur_circular)(const REAL* x, REAL* y, long planes, int H, int W,
                           const REAL* h, int kh, int kw, int adjoint)
{
    const int ch = kh / 2, cw = kw / 2;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; ++p) {
        const REAL* xp = x + p * (long)H * W;
        REAL* yp = y + p * (long)H * W;
        for (long i = 0; i < (long)H * W; ++i) yp[i] = 0;
        for (int n1 = 0; n1 < H; ++n1) {
            REAL* yrow = yp + (long)n1 * W;
            for (int i1 = 0; i1 < kh; ++i1) {
                int r = adjoint ? (n1 + i1 - ch) : (n1 - i1 + ch);
                r %= H; if (r < 0) r += H;
                const REAL* xrow = xp + (long)r * W;
                for (int i2 = 0; i2 < kw; ++i2) {
                    const REAL c = h[i1 * kw + i2];
                    int s = adjoint ? (i2 - cw) : (cw - i2);   /* source col = (n2 + s) mod W */
                    s %= W; if (s < 0) s += W;
                    /* two contiguous segments: n2 in [0, W-s) reads xrow[n2+s]; rest wraps */
                    const int split = W - s;
                    for (int n2 = 0; n2 < split; ++n2) yrow[n2] += c * xrow[n2 + s];
                    for (int n2 = split; n2 < W; ++n2) yrow[n2] += c * xrow[n2 + s - W];
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * v1 operator: conv / conv_transpose with padding modes (blur/__init__.py:9-161).
 * mode: 0 valid, 1 circular, 2 replicate, 3 reflect, 4 zero (conv_transpose only).
 * extend_filter (:9-31): size-1 axes become 3 (centred), even axes get one trailing zero.
 * The filter is flipped on both axes BEFORE being extended (:45-49, :94-98).
 * ---------------------------------------------------------------------------------- */
static int FN(orc_ext_size)(int n) { return n == 1 ? 3 : (n % 2 == 0 ? n + 1 : n); }

static void FN(orc_flip_extend)(const REAL* f, int fh, int fw, REAL* out, int eh, int ew)
{
    const int oh = (fh == 1) ? 1 : 0, ow = (fw == 1) ? 1 : 0;
    for (int i = 0; i < eh * ew; ++i) out[i] = 0;
    for (int a = 0; a < fh; ++a)
        for (int b = 0; b < fw; ++b)
            out[(a + oh) * ew + (b + ow)] = f[(fh - 1 - a) * fw + (fw - 1 - b)];
}

/* index of the source sample that padded coordinate m (already shifted by -pad) refers to;
 * -1 = no sample (zero) */
static int FN(orc_pad_index)(int m, int n, int mode)
{
    if (m >= 0 && m < n) return m;
    switch (mode) {
    case 1: m %= n; return m < 0 ? m + n : m;
    case 2: return m < 0 ? 0 : n - 1;
    case 3: return m < 0 ? -m : 2 * (n - 1) - m;
    default: return -1;
    }
}

void FN(orc_conv_v1_out_size)(int H, int W, int fh, int fw, int mode, int* Ho, int* Wo)
{
    const int ph = (FN(orc_ext_size)(fh) - 1) / 2, pw = (FN(orc_ext_size)(fw) - 1) / 2;
    *Ho = mode == 0 ? H - 2 * ph : H;
    *Wo = mode == 0 ? W - 2 * pw : W;
}

void FN(orc_conv_v1)(const REAL* x, REAL* y, long planes, int H, int W,
                     const REAL* filt, int fh, int fw, int mode)
{
    const int eh = FN(orc_ext_size)(fh), ew = FN(orc_ext_size)(fw);
    const int ph = (eh - 1) / 2, pw = (ew - 1) / 2;
    REAL fe[64 * 64];
    FN(orc_flip_extend)(filt, fh, fw, fe, eh, ew);
    int Ho, Wo;
    FN(orc_conv_v1_out_size)(H, W, fh, fw, mode, &Ho, &Wo);
    const int off_h = mode == 0 ? 0 : -ph, off_w = mode == 0 ? 0 : -pw;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; ++p) {
        const REAL* xp = x + p * (long)H * W;
        REAL* yp = y + p * (long)Ho * Wo;
        for (int n1 = 0; n1 < Ho; ++n1)
            for (int n2 = 0; n2 < Wo; ++n2) {
                REAL acc = 0;
                for (int a = 0; a < eh; ++a) {
                    const int r = FN(orc_pad_index)(n1 + a + off_h, H, mode);
                    if (r < 0) continue;
                    for (int b = 0; b < ew; ++b) {
                        const int c = FN(orc_pad_index)(n2 + b + off_w, W, mode);
                        if (c < 0) continue;
                        acc += fe[a * ew + b] * xp[(long)r * W + c];
                    }
                }
                yp[(long)n1 * Wo + n2] = acc;
            }
    }
}

/* conv_transpose (:77-161): full transposed convolution to (H+2ph, W+2pw), then the
 * borders are folded back onto the interior according to the padding mode. */
void FN(orc_conv_transpose_v1_out_size)(int H, int W, int fh, int fw, int mode, int* Ho, int* Wo)
{
    const int ph = (FN(orc_ext_size)(fh) - 1) / 2, pw = (FN(orc_ext_size)(fw) - 1) / 2;
    *Ho = mode == 0 ? H + 2 * ph : H;
    *Wo = mode == 0 ? W + 2 * pw : W;
}

void FN(orc_conv_transpose_v1)(const REAL* y, REAL* x, long planes, int H, int W,
                               const REAL* filt, int fh, int fw, int mode)
{
    const int eh = FN(orc_ext_size)(fh), ew = FN(orc_ext_size)(fw);
    const int ph = (eh - 1) / 2, pw = (ew - 1) / 2;
    REAL fe[64 * 64];
    FN(orc_flip_extend)(filt, fh, fw, fe, eh, ew);
    int Ho, Wo;
    FN(orc_conv_transpose_v1_out_size)(H, W, fh, fw, mode, &Ho, &Wo);
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; ++p) {
        const REAL* yp = y + p * (long)H * W;
        REAL* xp = x + p * (long)Ho * Wo;
        for (long i = 0; i < (long)Ho * Wo; ++i) xp[i] = 0;
        for (int n1 = 0; n1 < H; ++n1)
            for (int n2 = 0; n2 < W; ++n2) {
                const REAL v = yp[(long)n1 * W + n2];
                for (int a = 0; a < eh; ++a) {
                    /* full-output row n1+a; interior coordinate n1+a-ph */
                    const int r = mode == 0 ? n1 + a : FN(orc_pad_index)(n1 + a - ph, H, mode);
                    if (r < 0) continue;
                    for (int b = 0; b < ew; ++b) {
                        const int c = mode == 0 ? n2 + b : FN(orc_pad_index)(n2 + b - pw, W, mode);
                        if (c < 0) continue;
                        xp[(long)r * Wo + c] += fe[a * ew + b] * v;
                    }
                }
            }
    }
}

/* ------------------------------------------------------------------------------------
 * SR forward: Downsampling.A (src/physics/downsampling/__init__.py:16-19) =
 * F.interpolate(x, scale_factor=1/rate, mode="bicubic", antialias=True).
 * ATen separable anti-aliased resize: per output index i along an axis,
 *   scale = rate, support = 2*scale, center = scale*(i+0.5),
 *   xmin = max(0, (int)(center - support + 0.5)), xsize = min(in, (int)(center + support + 0.5)) - xmin,
 *   w_j = cubic_{a=-0.5}((j + xmin - center + 0.5)/scale), renormalised to sum 1.
 * Output size = floor(in * (1/rate)).  Horizontal pass first, then vertical.
 * ---------------------------------------------------------------------------------- */
static REAL FN(orc_aa_cubic)(REAL x)
{
    const REAL a = (REAL)-0.5;
    if (x < 0) x = -x;
    if (x < (REAL)1.0) return ((a + 2) * x - (a + 3)) * x * x + 1;
    if (x < (REAL)2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0;
}

int FN(orc_down_out_size)(int n, int rate) { return (int)floor((double)n * (1.0 / (double)rate)); }

/* weights for one axis: w[i*maxk + j], xmin[i], xsize[i]; returns maxk */
static int FN(orc_aa_weights)(int in, int out, int rate, REAL** w_out, int** xmin_out, int** xsize_out)
{
    const REAL scale = (REAL)(1.0 / (1.0 / (double)rate));
    const REAL support = (REAL)2.0 * scale;
    const int maxk = (int)ceil((double)support) * 2 + 1;
    REAL* w = (REAL*)calloc((size_t)out * maxk, sizeof(REAL));
    int* xmin = (int*)malloc(sizeof(int) * out);
    int* xsize = (int*)malloc(sizeof(int) * out);
    const REAL invscale = scale >= 1 ? (REAL)1.0 / scale : (REAL)1.0;
    for (int i = 0; i < out; ++i) {
        const REAL center = (REAL)((double)scale * (i + 0.5));
        int lo = (int)((double)center - (double)support + 0.5); if (lo < 0) lo = 0;
        int hi = (int)((double)center + (double)support + 0.5); if (hi > in) hi = in;
        xmin[i] = lo; xsize[i] = hi - lo;
        REAL total = 0;
        for (int j = 0; j < xsize[i]; ++j) {
            const REAL wj = FN(orc_aa_cubic)((REAL)(((double)(j + lo) - (double)center + 0.5) * (double)invscale));
            w[(size_t)i * maxk + j] = wj; total += wj;
        }
        if (total != 0) for (int j = 0; j < xsize[i]; ++j) w[(size_t)i * maxk + j] /= total;
    }
    *w_out = w; *xmin_out = xmin; *xsize_out = xsize;
    return maxk;
}

void FN(orc_down_aa)(const REAL* x, REAL* y, long planes, int H, int W, int rate)
{
    const int Ho = FN(orc_down_out_size)(H, rate), Wo = FN(orc_down_out_size)(W, rate);
    REAL *wh, *ww; int *hmin, *hsize, *wmin, *wsize;
    const int kh = FN(orc_aa_weights)(H, Ho, rate, &wh, &hmin, &hsize);
    const int kw = FN(orc_aa_weights)(W, Wo, rate, &ww, &wmin, &wsize);
#pragma omp parallel
    {
        REAL* tmp = (REAL*)malloc(sizeof(REAL) * (size_t)H * Wo);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            const REAL* xp = x + p * (long)H * W;
            REAL* yp = y + p * (long)Ho * Wo;
            for (int r = 0; r < H; ++r)
                for (int j = 0; j < Wo; ++j) {
                    REAL acc = 0;
                    const REAL* src = xp + (long)r * W + wmin[j];
                    const REAL* wj = ww + (size_t)j * kw;
                    for (int t = 0; t < wsize[j]; ++t) acc += wj[t] * src[t];
                    tmp[(size_t)r * Wo + j] = acc;
                }
            for (int i = 0; i < Ho; ++i) {
                REAL* yrow = yp + (long)i * Wo;
                for (int j = 0; j < Wo; ++j) yrow[j] = 0;
                for (int t = 0; t < hsize[i]; ++t) {
                    const REAL c = wh[(size_t)i * kh + t];
                    const REAL* trow = tmp + (size_t)(hmin[i] + t) * Wo;
                    for (int j = 0; j < Wo; ++j) yrow[j] += c * trow[j];
                }
            }
        }
        free(tmp);
    }
    free(wh); free(ww); free(hmin); free(hsize); free(wmin); free(wsize);
}

/* transpose of orc_down_aa: what autograd computes for A in a training step, and
 * Downsampling.A_adjoint with true_adjoint=True (downsampling/__init__.py:21-31). */
void FN(orc_down_aa_vjp)(const REAL* gy, REAL* gx, long planes, int H, int W, int rate)
{
    const int Ho = FN(orc_down_out_size)(H, rate), Wo = FN(orc_down_out_size)(W, rate);
    REAL *wh, *ww; int *hmin, *hsize, *wmin, *wsize;
    const int kh = FN(orc_aa_weights)(H, Ho, rate, &wh, &hmin, &hsize);
    const int kw = FN(orc_aa_weights)(W, Wo, rate, &ww, &wmin, &wsize);
#pragma omp parallel
    {
        REAL* tmp = (REAL*)malloc(sizeof(REAL) * (size_t)H * Wo);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            const REAL* gyp = gy + p * (long)Ho * Wo;
            REAL* gxp = gx + p * (long)H * W;
            for (size_t i = 0; i < (size_t)H * Wo; ++i) tmp[i] = 0;
            for (int i = 0; i < Ho; ++i)
                for (int t = 0; t < hsize[i]; ++t) {
                    const REAL c = wh[(size_t)i * kh + t];
                    REAL* trow = tmp + (size_t)(hmin[i] + t) * Wo;
                    const REAL* grow = gyp + (long)i * Wo;
                    for (int j = 0; j < Wo; ++j) trow[j] += c * grow[j];
                }
            for (long i = 0; i < (long)H * W; ++i) gxp[i] = 0;
            for (int r = 0; r < H; ++r)
                for (int j = 0; j < Wo; ++j) {
                    const REAL v = tmp[(size_t)r * Wo + j];
                    REAL* dst = gxp + (long)r * W + wmin[j];
                    const REAL* wj = ww + (size_t)j * kw;
                    for (int t = 0; t < wsize[j]; ++t) dst[t] += wj[t] * v;
                }
        }
        free(tmp);
    }
    free(wh); free(ww); free(hmin); free(hsize); free(wmin); free(wsize);
}

/* Keys cubic convolution coefficients with A=-0.75 (torch upsample_bicubic2d / grid_sample). */
static void FN(orc_cubic_coeffs)(REAL t, REAL c[4])
{
    const REAL A = (REAL)-0.75;
    REAL x;
    x = t + 1; c[0] = ((A * x - 5 * A) * x + 8 * A) * x - 4 * A;
    x = t;     c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    x = 1 - t; c[2] = ((A + 2) * x - (A + 3)) * x * x + 1;
    x = 2 - t; c[3] = ((A * x - 5 * A) * x + 8 * A) * x - 4 * A;
}

/* Downsampling.A_adjoint with true_adjoint=False (downsampling/__init__.py:32-35): NOT an
 * adjoint, a plain bicubic upsample F.interpolate(y, scale_factor=rate, mode="bicubic"),
 * align_corners=False: src = (dst + 0.5)/rate - 0.5, 4 taps with indices clamped. */
void FN(orc_up_bicubic)(const REAL* y, REAL* x, long planes, int h, int w, int rate)
{
    const int H = h * rate, W = w * rate;
    const REAL s = (REAL)(1.0 / (double)rate);
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; ++p) {
        const REAL* yp = y + p * (long)h * w;
        REAL* xp = x + p * (long)H * W;
        for (int i = 0; i < H; ++i) {
            const REAL ry = s * ((REAL)i + (REAL)0.5) - (REAL)0.5;
            const int iy = (int)floor((double)ry);
            REAL cy[4]; FN(orc_cubic_coeffs)(ry - (REAL)iy, cy);
            for (int j = 0; j < W; ++j) {
                const REAL rx = s * ((REAL)j + (REAL)0.5) - (REAL)0.5;
                const int ix = (int)floor((double)rx);
                REAL cx[4]; FN(orc_cubic_coeffs)(rx - (REAL)ix, cx);
                REAL acc = 0;
                for (int a = 0; a < 4; ++a) {
                    int r = iy - 1 + a; r = r < 0 ? 0 : (r > h - 1 ? h - 1 : r);
                    REAL row = 0;
                    for (int b = 0; b < 4; ++b) {
                        int c = ix - 1 + b; c = c < 0 ? 0 : (c > w - 1 ? w - 1 : c);
                        row += yp[(long)r * w + c] * cx[b];
                    }
                    acc += row * cy[a];
                }
                xp[(long)i * W + j] = acc;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Scale transform: padded_downsampling_transform (src/transforms.py:60-83) with the grid
 * of get_downsampling_grid (:27-43), square SxS images:
 *   gx[b,i,j] = (1/rate_b) * ((2/S * j - 1) - cx_b) + cx_b     (x indexes width)
 *   gy[b,i,j] = (1/rate_b) * ((2/S * i - 1) - cy_b) + cy_b
 * center[b] = (cx_b, cy_b).  Then F.grid_sample(mode="bicubic",
 * padding_mode="reflection", align_corners=True): pixel = ((g+1)/2)*(S-1), floor, Keys
 * coefficients (A=-0.75), every tap coordinate reflected about [0, S-1] and clipped.
 * ---------------------------------------------------------------------------------- */
static REAL FN(orc_grid_coord)(int idx, int S, REAL inv_rate, REAL c)
{
    const REAL u = (REAL)(2.0 / (double)S) * (REAL)idx - (REAL)1.0;   /* 2/w is a Python float */
    return inv_rate * (u - c) + c;
}

void FN(orc_scale_grid)(REAL* grid, int B, int S, const REAL* rate, const REAL* center)
{
    for (int b = 0; b < B; ++b) {
        const REAL inv_rate = (REAL)1.0 / rate[b];
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                REAL* g = grid + (((long)b * S + i) * S + j) * 2;
                g[0] = FN(orc_grid_coord)(j, S, inv_rate, center[2 * b + 0]);
                g[1] = FN(orc_grid_coord)(i, S, inv_rate, center[2 * b + 1]);
            }
    }
}

static int FN(orc_reflect_clip)(int idx, int S)
{
    /* reflect_coordinates(in, 0, 2*(S-1)) then clip to [0, S-1]; exact on integers */
    if (S == 1) return 0;
    const int span = S - 1;
    int v = idx < 0 ? -idx : idx;
    const int flips = v / span, extra = v % span;
    v = (flips % 2 == 0) ? extra : span - extra;
    return v < 0 ? 0 : (v > S - 1 ? S - 1 : v);
}

/* x: B x C x Ss x Ss (Ss == S, or the pre-filtered smaller image of the antialiased variant, src/transforms.py:63-67:
 * the grid keeps the ORIGINAL shape S and grid_sample un-normalises against the source size), out: B x C x S x S */
void FN(orc_scale_transform_src)(const REAL* x, REAL* out, int B, int C, int Ss, int S,
                                 const REAL* rate, const REAL* center)
{
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL inv_rate = (REAL)1.0 / rate[b];
            const REAL* xp = x + ((long)b * C + c) * Ss * Ss;
            REAL* op = out + ((long)b * C + c) * S * S;
            for (int i = 0; i < S; ++i) {
                const REAL gy = FN(orc_grid_coord)(i, S, inv_rate, center[2 * b + 1]);
                const REAL py = ((gy + 1) / 2) * (REAL)(Ss - 1);
                const REAL fy = (REAL)floor((double)py);
                REAL cy[4]; FN(orc_cubic_coeffs)(py - fy, cy);
                int ry[4];
                for (int a = 0; a < 4; ++a) ry[a] = FN(orc_reflect_clip)((int)fy - 1 + a, Ss);
                for (int j = 0; j < S; ++j) {
                    const REAL gx = FN(orc_grid_coord)(j, S, inv_rate, center[2 * b + 0]);
                    const REAL px = ((gx + 1) / 2) * (REAL)(Ss - 1);
                    const REAL fx = (REAL)floor((double)px);
                    REAL cx[4]; FN(orc_cubic_coeffs)(px - fx, cx);
                    int rx[4];
                    for (int t = 0; t < 4; ++t) rx[t] = FN(orc_reflect_clip)((int)fx - 1 + t, Ss);
                    REAL acc = 0;
                    for (int a = 0; a < 4; ++a) {
                        const REAL* row = xp + (long)ry[a] * Ss;
                        const REAL interp = row[rx[0]] * cx[0] + row[rx[1]] * cx[1]
                                          + row[rx[2]] * cx[2] + row[rx[3]] * cx[3];
                        acc += interp * cy[a];
                    }
                    op[(long)i * S + j] = acc;
                }
            }
        }
}

void FN(orc_scale_transform)(const REAL* x, REAL* out, int B, int C, int S,
                             const REAL* rate, const REAL* center)
{
    FN(orc_scale_transform_src)(x, out, B, C, S, S, rate, center);
}

/* transpose of orc_scale_transform w.r.t. x: what autograd computes through grid_sample when
 * ProposedLoss__stop_gradient is off (deepinv EILoss no_grad=False; src/losses/__init__.py:117-122). */
void FN(orc_scale_transform_vjp)(const REAL* gout, REAL* gx, int B, int C, int S,
                                 const REAL* rate, const REAL* center)
{
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL inv_rate = (REAL)1.0 / rate[b];
            const REAL* gp = gout + ((long)b * C + c) * S * S;
            REAL* xp = gx + ((long)b * C + c) * S * S;
            for (long i = 0; i < (long)S * S; ++i) xp[i] = 0;
            for (int i = 0; i < S; ++i) {
                const REAL gy = FN(orc_grid_coord)(i, S, inv_rate, center[2 * b + 1]);
                const REAL py = ((gy + 1) / 2) * (REAL)(S - 1);
                const REAL fy = (REAL)floor((double)py);
                REAL cy[4]; FN(orc_cubic_coeffs)(py - fy, cy);
                int ry[4];
                for (int a = 0; a < 4; ++a) ry[a] = FN(orc_reflect_clip)((int)fy - 1 + a, S);
                for (int j = 0; j < S; ++j) {
                    const REAL gxc = FN(orc_grid_coord)(j, S, inv_rate, center[2 * b + 0]);
                    const REAL px = ((gxc + 1) / 2) * (REAL)(S - 1);
                    const REAL fx = (REAL)floor((double)px);
                    REAL cx[4]; FN(orc_cubic_coeffs)(px - fx, cx);
                    const REAL g = gp[(long)i * S + j];
                    for (int a = 0; a < 4; ++a)
                        for (int t = 0; t < 4; ++t)
                            xp[(long)ry[a] * S + FN(orc_reflect_clip)((int)fx - 1 + t, S)] += g * cy[a] * cx[t];
                }
            }
        }
}

/* ------------------------------------------------------------------------------------
 * Loss assembly.
 * SureGaussianLoss.forward + mc_div (src/losses/sure.py:7-76), given the operator
 * outputs y1 = A(model(y)), y2 = A(model(y + tau*b)):
 *   div  = mean over the div-interior of b*(y2-y1)/tau      (margin_div: margin if cropped_div else 0)
 *   mse  = mean over the mse-interior of (y1-y)^2           (margin_mse = margin)
 *   loss = mse + 2*sigma^2*div - (averaged_cst ? sigma^2 : sigma^2 / B)
 * Sums are accumulated in double regardless of REAL.
 * ---------------------------------------------------------------------------------- */
double FN(orc_sure_loss)(const REAL* y1, const REAL* y2, const REAL* y, const REAL* b,
                         int B, int C, int H, int W, int margin_mse, int margin_div,
                         double tau, double sigma2, int averaged_cst,
                         double* mse_out, double* div_out)
{
    double mse = 0, div = 0;
    for (long p = 0; p < (long)B * C; ++p)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const long o = (p * H + i) * W + j;
                if (i >= margin_mse && i < H - margin_mse && j >= margin_mse && j < W - margin_mse) {
                    const REAL d = y1[o] - y[o];
                    mse += (double)(d * d);
                }
                if (i >= margin_div && i < H - margin_div && j >= margin_div && j < W - margin_div) {
                    const REAL t = b[o] * (y2[o] - y1[o]) / (REAL)tau;
                    div += (double)t;
                }
            }
    mse /= (double)B * C * (H - 2 * margin_mse) * (W - 2 * margin_mse);
    div /= (double)B * C * (H - 2 * margin_div) * (W - 2 * margin_div);
    if (mse_out) *mse_out = mse;
    if (div_out) *div_out = div;
    return mse + 2.0 * sigma2 * div - (averaged_cst ? sigma2 : sigma2 / (double)B);
}

/* nn.MSELoss() (deepinv.loss.metric.mse, used by EILoss / SupLoss; src/losses/__init__.py:17-37,117-122) */
double FN(orc_mse)(const REAL* a, const REAL* b, long n)
{
    double s = 0;
    for (long i = 0; i < n; ++i) { const REAL d = a[i] - b[i]; s += (double)(d * d); }
    return s / (double)n;
}

/* deepinv GaussianNoise (attached at src/physics/__init__.py:53): y + n*sigma with the
 * standard-normal tensor n supplied by the caller. */
void FN(orc_add_noise)(const REAL* y, const REAL* n, REAL* out, long count, REAL sigma)
{
    for (long i = 0; i < count; ++i) out[i] = y[i] + n[i] * sigma;
}
