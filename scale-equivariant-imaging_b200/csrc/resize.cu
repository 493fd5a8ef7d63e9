// resize.cu -- bicubic resize by a fractional factor: the "normal" scale transform of the reference
// (src/transforms.py:112-145 NormalDownsamplingTransform -> F.interpolate(x, scale_factor=rate, mode="bicubic",
// antialias=...), rate in {0.75, 0.5}).  A non-default variant (demo/train.py defaults to kind="padded"), so one direct
// kernel: a thread per output element, separable weights evaluated on the fly.
//   antialias = 0: ATen upsample_bicubic2d -- source coordinate scale * (dst + 0.5) - 0.5 (no clamping), Keys cubic
//                  A = -0.75 on the four neighbours, indices clamped to the image;
//   antialias = 1: ATen _upsample_bicubic2d_aa -- window [center - support, center + support) with support = 2 * scale
//                  for scale >= 1, cubic a = -0.5 stretched by 1 / scale, weights renormalised (sei::aa_axis_weights).
// scale = 1 / scale_factor, as ATen computes it when a scale factor is given.
#include "sei_common.cuh"
#include <algorithm>

namespace sei {

struct ResizeParams {
    const float* x;
    float* y;
    int H, W, Ho, Wo, aa;
    float sh, sw;
    long long total;
};

__device__ __forceinline__ void aa_axis_weights_f(int i, int in_size, float scale, float (&w)[kAaMaxTaps], int& xmin, int& xsize)
{
    const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const float center = scale * ((float)i + 0.5f);
    int lo = (int)(center - support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5f);
    if (hi > in_size) hi = in_size;
    xmin = lo;
    xsize = min(hi - lo, kAaMaxTaps);
    float total = 0.0f;
#pragma unroll
    for (int j = 0; j < kAaMaxTaps; ++j) {
        float wj = 0.0f;
        if (j < xsize) wj = aa_cubic(((float)(j + lo) - center + 0.5f) * invscale);
        w[j] = wj;
        total += wj;
    }
    if (total != 0.0f) {
#pragma unroll
        for (int j = 0; j < kAaMaxTaps; ++j) w[j] = w[j] / total;
    }
}

__global__ void __launch_bounds__(256) resize_bicubic_kernel(const __grid_constant__ ResizeParams p)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p.Wo);
        const long long t = idx / p.Wo;
        const int i = (int)(t % p.Ho);
        const float* xp = p.x + (t / p.Ho) * (long long)p.H * p.W;
        float acc = 0.f;
        if (!p.aa) {
            const float ry = p.sh * ((float)i + 0.5f) - 0.5f, rx = p.sw * ((float)j + 0.5f) - 0.5f;
            const float fy = floorf(ry), fx = floorf(rx);
            float cy[4], cx[4];
            keys_coeffs(ry - fy, cy);
            keys_coeffs(rx - fx, cx);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = min(max((int)fy - 1 + a, 0), p.H - 1);
                float row = 0.f;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int c = min(max((int)fx - 1 + b, 0), p.W - 1);
                    row = fmaf(__ldg(xp + (size_t)r * p.W + c), cx[b], row);
                }
                acc = fmaf(row, cy[a], acc);
            }
        } else {
            float wy[kAaMaxTaps], wx[kAaMaxTaps];
            int ymin, ysize, xmin, xsize;
            aa_axis_weights_f(i, p.H, p.sh, wy, ymin, ysize);
            aa_axis_weights_f(j, p.W, p.sw, wx, xmin, xsize);
            for (int a = 0; a < ysize; ++a) {
                const float* row = xp + (size_t)(ymin + a) * p.W + xmin;
                float rv = 0.f;
                for (int b = 0; b < xsize; ++b) rv = fmaf(__ldg(row + b), wx[b], rv);
                acc = fmaf(rv, wy[a], acc);
            }
        }
        p.y[idx] = acc;
    }
}


// Transposes (autograd through the transforms when the EI branch keeps its gradient, --no-ProposedLoss__stop_gradient:
// reference src/losses/__init__.py:84-96 with no_grad=False): every output gradient is scattered to the taps its
// forward value was read from, with the forward kernel's weights, by fp32 atomics into a zeroed buffer.  Not on the
// default path, so direct kernels.
__global__ void __launch_bounds__(256) resize_bicubic_backward_kernel(const __grid_constant__ ResizeParams p)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p.Wo);
        const long long t = idx / p.Wo;
        const int i = (int)(t % p.Ho);
        float* gx = p.y + (t / p.Ho) * (long long)p.H * p.W;          // (p.x: output gradient, p.y: input gradient)
        const float g = __ldg(p.x + idx);
        if (!p.aa) {
            const float ry = p.sh * ((float)i + 0.5f) - 0.5f, rx = p.sw * ((float)j + 0.5f) - 0.5f;
            const float fy = floorf(ry), fx = floorf(rx);
            float cy[4], cx[4];
            keys_coeffs(ry - fy, cy);
            keys_coeffs(rx - fx, cx);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = min(max((int)fy - 1 + a, 0), p.H - 1);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int c = min(max((int)fx - 1 + b, 0), p.W - 1);
                    atomicAdd(gx + (size_t)r * p.W + c, g * cy[a] * cx[b]);
                }
            }
        } else {
            float wy[kAaMaxTaps], wx[kAaMaxTaps];
            int ymin, ysize, xmin, xsize;
            aa_axis_weights_f(i, p.H, p.sh, wy, ymin, ysize);
            aa_axis_weights_f(j, p.W, p.sw, wx, xmin, xsize);
            for (int a = 0; a < ysize; ++a) {
                float* row = gx + (size_t)(ymin + a) * p.W + xmin;
                const float ga = g * wy[a];
                for (int b = 0; b < xsize; ++b) atomicAdd(row + b, ga * wx[b]);
            }
        }
    }
}

// Rotation by nearest-neighbour resampling: deepinv.transform.Rotate -> torchvision.transforms.functional.rotate(x, angle)
// with its defaults (NEAREST, expand=False, zero fill), i.e. F.grid_sample(x, grid, mode="nearest", padding_mode="zeros",
// align_corners=False) on the grid of torchvision's _gen_affine_grid.  The fp32 operation order of that grid (linspace
// from both ends, rescaled matrix, unnormalisation, round-half-even) is kept so that the SAME source pixel is picked;
// the six matrix entries arrive already rescaled (theta^T / [W/2, H/2], computed by the caller in fp32 like torchvision).
struct RotateParams {
    const float* x;
    float* y;
    int H, W;
    float r00, r10, r20, r01, r11, r21;      // gx = bx * r00 + by * r10 + r20 ; gy = bx * r01 + by * r11 + r21
    float x_start, x_end, x_step, y_start, y_end, y_step;
    long long total;
};

// torch.linspace: start + step * i in the first half, end - step * (steps - 1 - i) in the second
__device__ __forceinline__ float linspace_at(int i, int steps, float start, float end, float step)
{
    return i < steps / 2 ? __fadd_rn(start, __fmul_rn(step, (float)i)) : __fsub_rn(end, __fmul_rn(step, (float)(steps - 1 - i)));
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256) rotate_nearest_kernel(const __grid_constant__ RotateParams p)
{
    const int H = p.H, W = p.W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        const long long t = idx / W;
        const int i = (int)(t % H);
        const long long plane = t / H;
        const float bx = linspace_at(j, W, p.x_start, p.x_end, p.x_step), by = linspace_at(i, H, p.y_start, p.y_end, p.y_step);
        const float gx = __fadd_rn(__fadd_rn(__fmul_rn(bx, p.r00), __fmul_rn(by, p.r10)), p.r20);
        const float gy = __fadd_rn(__fadd_rn(__fmul_rn(bx, p.r01), __fmul_rn(by, p.r11)), p.r21);
        // grid_sample, align_corners=False: ((g + 1) * size - 1) / 2, then nearbyint
        const float fx = rintf(__fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 2.0f));
        const float fy = rintf(__fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 2.0f));
        const bool inside = fx >= 0.0f && fx <= (float)(W - 1) && fy >= 0.0f && fy <= (float)(H - 1);
        const long long src = plane * (long long)H * W + (long long)(int)fy * W + (int)fx;
        if (BACKWARD) {                    // p.x: output gradient, p.y: zeroed input gradient (several outputs may pick one pixel)
            if (inside) atomicAdd(p.y + src, __ldg(p.x + idx));
        } else {
            p.y[idx] = inside ? __ldg(p.x + src) : 0.0f;
        }
    }
}

}  // namespace sei

using namespace sei;

// y[planes, Ho, Wo] = bicubic resize of x[planes, H, W]; scale_h / scale_w = 1 / scale_factor per axis
extern "C" int sei_resize_bicubic_f32(const float* x, float* y, long long planes, int H, int W, int Ho, int Wo,
                                      float scale_h, float scale_w, int antialias, void* stream)
{
    SEI_REQUIRE(x && y, "null pointer argument");
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "bad shape planes=%lld %dx%d -> %dx%d", planes, H, W, Ho, Wo);
    SEI_REQUIRE(scale_h > 0.f && scale_w > 0.f, "scales must be positive");
    SEI_REQUIRE(!antialias || (4.0f * std::max(scale_h, scale_w) + 2.0f <= (float)kAaMaxTaps),
                "antialiased resize supports scale factors down to 0.29 (%d taps per axis)", kAaMaxTaps);
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    ResizeParams p;
    p.x = x; p.y = y; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.aa = antialias ? 1 : 0; p.sh = scale_h; p.sw = scale_w;
    p.total = planes * (long long)Ho * Wo;
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 32);
    resize_bicubic_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return finish_launch("resize_bicubic_kernel");
}

// gx[planes, H, W] = transpose of sei_resize_bicubic_f32 applied to gy[planes, Ho, Wo] (same scales / antialias flag)
extern "C" int sei_resize_bicubic_backward_f32(const float* gy, float* gx, long long planes, int H, int W, int Ho, int Wo,
                                               float scale_h, float scale_w, int antialias, void* stream)
{
    SEI_REQUIRE(gy && gx, "null pointer argument");
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "bad shape planes=%lld %dx%d -> %dx%d", planes, H, W, Ho, Wo);
    SEI_REQUIRE(scale_h > 0.f && scale_w > 0.f, "scales must be positive");
    SEI_REQUIRE(!antialias || (4.0f * std::max(scale_h, scale_w) + 2.0f <= (float)kAaMaxTaps),
                "antialiased resize supports scale factors down to 0.29 (%d taps per axis)", kAaMaxTaps);
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SEI_CUDA(cudaMemsetAsync(gx, 0, (size_t)planes * H * W * sizeof(float), st));
    ResizeParams p;
    p.x = gy; p.y = gx; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.aa = antialias ? 1 : 0; p.sh = scale_h; p.sw = scale_w;
    p.total = planes * (long long)Ho * Wo;
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 32);
    resize_bicubic_backward_kernel<<<grid, 256, 0, st>>>(p);
    return finish_launch("resize_bicubic_backward_kernel");
}

static int rotate_nearest_impl(const float* x, float* y, long long planes, int H, int W, const float* rescaled_theta,
                               bool backward, void* stream);

extern "C" int sei_rotate_nearest_f32(const float* x, float* y, long long planes, int H, int W, const float* rescaled_theta,
                                      void* stream)
{
    return rotate_nearest_impl(x, y, planes, H, W, rescaled_theta, false, stream);
}

// gx = transpose of sei_rotate_nearest_f32 applied to gy (every output gradient goes to the source pixel it was read from)
extern "C" int sei_rotate_nearest_backward_f32(const float* gy, float* gx, long long planes, int H, int W,
                                               const float* rescaled_theta, void* stream)
{
    return rotate_nearest_impl(gy, gx, planes, H, W, rescaled_theta, true, stream);
}

static int rotate_nearest_impl(const float* x, float* y, long long planes, int H, int W, const float* rescaled_theta,
                               bool backward, void* stream)
{
    SEI_REQUIRE(x && y && rescaled_theta, "null pointer argument");
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0, "bad shape planes=%lld %dx%d", planes, H, W);
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    RotateParams p;
    p.x = x; p.y = y; p.H = H; p.W = W;
    // rescaled_theta: HOST pointer, 3 x 2 row-major (torchvision's rescaled_theta[0])
    p.r00 = rescaled_theta[0]; p.r01 = rescaled_theta[1]; p.r10 = rescaled_theta[2]; p.r11 = rescaled_theta[3];
    p.r20 = rescaled_theta[4]; p.r21 = rescaled_theta[5];
    // linspace(-W/2 + 0.5, W/2 + 0.5 - 1, W) with the end points computed in double and rounded once, like the Python floats
    p.x_start = (float)(-(double)W * 0.5 + 0.5); p.x_end = (float)((double)W * 0.5 + 0.5 - 1.0);
    p.y_start = (float)(-(double)H * 0.5 + 0.5); p.y_end = (float)((double)H * 0.5 + 0.5 - 1.0);
    p.x_step = W > 1 ? (p.x_end - p.x_start) / (float)(W - 1) : 0.0f;
    p.y_step = H > 1 ? (p.y_end - p.y_start) / (float)(H - 1) : 0.0f;
    p.total = planes * (long long)H * W;
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 32);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (backward) {
        SEI_CUDA(cudaMemsetAsync(y, 0, (size_t)p.total * sizeof(float), st));
        rotate_nearest_kernel<true><<<grid, 256, 0, st>>>(p);
    } else {
        rotate_nearest_kernel<false><<<grid, 256, 0, st>>>(p);
    }
    return finish_launch("rotate_nearest_kernel");
}
