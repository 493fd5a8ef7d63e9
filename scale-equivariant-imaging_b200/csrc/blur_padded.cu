// blur_padded.cu -- the reference's v1 operator with its non-default paddings:
// conv / conv_transpose of src/physics/blur/__init__.py:34-161 (valid, circular, replicate,
// reflect; "zero" for the transpose only), including extend_filter (:9-31).  Not on the training
// hot path (the factory always builds padding="circular", src/physics/__init__.py:46) but part of
// the physics API surface; one thread per output element, filter taps in the constant bank.
#include "sei_common.cuh"
#include <algorithm>

namespace sei {

constexpr int kPadMaxK = 31;

struct PadParams {
    const float* in;
    float* out;
    int H, W, Ho, Wo, eh, ew, mode;
    long long total;
    float fe[kPadMaxK * kPadMaxK];   // flipped + extended filter, row-major eh x ew
};

// source index referred to by padded coordinate m (already shifted by -pad); -1 = zero
__device__ __forceinline__ int pad_index(int m, int n, int mode)
{
    if (m >= 0 && m < n) return m;
    switch (mode) {
    case 1: m %= n; return m < 0 ? m + n : m;
    case 2: return m < 0 ? 0 : n - 1;
    case 3: return m < 0 ? -m : 2 * (n - 1) - m;
    default: return -1;
    }
}

__global__ void __launch_bounds__(256) conv_padded_kernel(const __grid_constant__ PadParams p)
{
    const int ph = (p.eh - 1) / 2, pw = (p.ew - 1) / 2;
    const int off_h = p.mode == 0 ? 0 : -ph, off_w = p.mode == 0 ? 0 : -pw;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n2 = (int)(idx % p.Wo);
        const long long t = idx / p.Wo;
        const int n1 = (int)(t % p.Ho);
        const float* xp = p.in + (t / p.Ho) * (long long)p.H * p.W;
        float acc = 0.f;
        for (int a = 0; a < p.eh; ++a) {
            const int r = pad_index(n1 + a + off_h, p.H, p.mode);
            if (r < 0) continue;
            for (int b = 0; b < p.ew; ++b) {
                const int c = pad_index(n2 + b + off_w, p.W, p.mode);
                if (c < 0) continue;
                acc = fmaf(p.fe[a * p.ew + b], __ldg(xp + (size_t)r * p.W + c), acc);
            }
        }
        p.out[idx] = acc;
    }
}

// full[m1][m2] = sum_{a,b} fe[a][b] * y[m1-a][m2-b]   (the un-folded transposed convolution)
__device__ __forceinline__ float full_tconv(const PadParams& p, const float* yp, int m1, int m2)
{
    float acc = 0.f;
    const int a_lo = max(0, m1 - (p.H - 1)), a_hi = min(p.eh - 1, m1);
    const int b_lo = max(0, m2 - (p.W - 1)), b_hi = min(p.ew - 1, m2);
    for (int a = a_lo; a <= a_hi; ++a)
        for (int b = b_lo; b <= b_hi; ++b)
            acc = fmaf(p.fe[a * p.ew + b], __ldg(yp + (size_t)(m1 - a) * p.W + (m2 - b)), acc);
    return acc;
}

__global__ void __launch_bounds__(256) conv_transpose_padded_kernel(const __grid_constant__ PadParams p)
{
    const int ph = (p.eh - 1) / 2, pw = (p.ew - 1) / 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % p.Wo);
        const long long t = idx / p.Wo;
        const int r = (int)(t % p.Ho);
        const float* yp = p.in + (t / p.Ho) * (long long)p.H * p.W;
        float acc;
        if (p.mode == 0) {
            acc = full_tconv(p, yp, r, c);
        } else {
            // interior sample plus every border sample of the full output that folds onto (r, c)
            acc = 0.f;
            for (int k1 = -1; k1 < 2 * ph; ++k1) {
                const int m1 = k1 < 0 ? r + ph : (k1 < ph ? k1 : p.H + k1);   // border rows: [0,ph) and [H+ph, H+2ph)
                if (k1 >= 0 && pad_index(m1 - ph, p.H, p.mode) != r) continue;
                for (int k2 = -1; k2 < 2 * pw; ++k2) {
                    const int m2 = k2 < 0 ? c + pw : (k2 < pw ? k2 : p.W + k2);
                    if (k2 >= 0 && pad_index(m2 - pw, p.W, p.mode) != c) continue;
                    acc += full_tconv(p, yp, m1, m2);
                }
            }
        }
        p.out[idx] = acc;
    }
}

static int ext_size(int n) { return n == 1 ? 3 : (n % 2 == 0 ? n + 1 : n); }

}  // namespace sei

using namespace sei;

extern "C" int sei_blur_padded_f32(const float* in, float* out, long long planes, int H, int W,
                                   const double* filter_host, int fh, int fw, int mode, int transpose,
                                   int* Ho_out, int* Wo_out, void* stream)
{
    SEI_REQUIRE(filter_host, "null filter");
    SEI_REQUIRE(fh >= 1 && fw >= 1 && ext_size(fh) <= kPadMaxK && ext_size(fw) <= kPadMaxK, "filter %dx%d unsupported", fh, fw);
    SEI_REQUIRE(mode >= 0 && mode <= 4 && (transpose || mode != 4), "bad padding mode %d", mode);
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0, "bad shape");
    const int eh = ext_size(fh), ew = ext_size(fw), ph = (eh - 1) / 2, pw = (ew - 1) / 2;
    int Ho = H, Wo = W;
    if (mode == 0) {
        Ho = transpose ? H + 2 * ph : H - 2 * ph;
        Wo = transpose ? W + 2 * pw : W - 2 * pw;
    }
    SEI_REQUIRE(Ho > 0 && Wo > 0, "image %dx%d too small for a valid convolution with a %dx%d filter", H, W, eh, ew);
    // F.pad limits: circular needs pad <= size, reflect needs pad < size
    SEI_REQUIRE(mode != 1 || (ph <= H && pw <= W), "circular padding larger than the image");
    SEI_REQUIRE(mode != 3 || (ph < H && pw < W), "reflect padding must be smaller than the image");
    if (Ho_out) *Ho_out = Ho;
    if (Wo_out) *Wo_out = Wo;
    if (!out) return 0;
    SEI_REQUIRE(in, "null input");
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    PadParams p;
    p.in = in; p.out = out; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.eh = eh; p.ew = ew; p.mode = mode;
    p.total = planes * (long long)Ho * Wo;
    // extend_filter applied to the flipped filter: size-1 axes are centred in 3, even axes get a trailing zero
    const int oh = fh == 1 ? 1 : 0, ow = fw == 1 ? 1 : 0;
    for (int i = 0; i < eh * ew; ++i) p.fe[i] = 0.f;
    for (int a = 0; a < fh; ++a)
        for (int b = 0; b < fw; ++b)
            p.fe[(a + oh) * ew + (b + ow)] = (float)filter_host[(fh - 1 - a) * fw + (fw - 1 - b)];
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 32);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (transpose) {
        conv_transpose_padded_kernel<<<grid, 256, 0, st>>>(p);
        return finish_launch("conv_transpose_padded_kernel");
    }
    conv_padded_kernel<<<grid, 256, 0, st>>>(p);
    return finish_launch("conv_padded_kernel");
}
