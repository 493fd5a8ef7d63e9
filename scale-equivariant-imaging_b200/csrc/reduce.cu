// reduce.cu -- loss reductions and the small elementwise steps of the loss assembly
// (reference: src/losses/sure.py:7-76 mc_div / SureGaussianLoss; nn.MSELoss via deepinv
// metric.mse in EILoss / SupLoss, src/losses/__init__.py:17-37,117-122; deepinv GaussianNoise).
//
// Reductions are single-launch and deterministic: every block reduces its grid-stride slice with
// warp shuffles (fp32 per thread, fp64 from the warp level up), writes one fp64 partial per
// quantity to the workspace, and the last block to finish (atomic ticket) sums the partials in
// index order and writes the result.  The workspace's ticket is reset by that block, so a
// zero-initialised workspace can be reused by every call on the same stream.
#include "sei_common.cuh"
#include <algorithm>

namespace sei {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;
constexpr long long kRedWorkspaceBytes = 64 + 2ll * kRedMaxBlocks * 8;

struct RedWorkspace {
    unsigned int ticket;
    unsigned int pad[15];
    double partial[2][kRedMaxBlocks];
};

// finish a block's contribution; returns true in thread 0 of the last block, with totals in v
template <int NV>
__device__ __forceinline__ bool grid_finish(double (&v)[NV], RedWorkspace* ws, double* scratch, bool* s_last)
{
    block_sum<NV>(v, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) ws->partial[k][blockIdx.x] = v[k];
        __threadfence();
        const unsigned t = atomicAdd(&ws->ticket, 1u);
        *s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!*s_last) return false;
    __threadfence();
    double t[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double a = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a += __ldcg(&ws->partial[k][i]);
        t[k] = a;
    }
    __syncthreads();
    block_sum<NV>(t, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = t[k];
        ws->ticket = 0;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(kRedThreads) mse_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          long long n, float* out, RedWorkspace* ws, int vec_ok)
{
    __shared__ double scratch[64];
    __shared__ bool s_last;
    float acc = 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    if (vec_ok) {
        const long long n4 = n >> 2;
        for (long long i = tid; i < n4; i += nth) {
            const float4 u = ld_stream4(a + 4 * i), w = ld_stream4(b + 4 * i);
            const float d0 = u.x - w.x, d1 = u.y - w.y, d2 = u.z - w.z, d3 = u.w - w.w;
            acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
        }
        for (long long i = 4 * n4 + tid; i < n; i += nth) { const float d = a[i] - b[i]; acc = fmaf(d, d, acc); }
    } else {
        for (long long i = tid; i < n; i += nth) { const float d = a[i] - b[i]; acc = fmaf(d, d, acc); }
    }
    double v[1] = {(double)acc};
    if (grid_finish<1>(v, ws, scratch, &s_last)) out[0] = (float)(v[0] / (double)n);
}

__global__ void __launch_bounds__(kRedThreads) mse_backward_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   long long n, const float* __restrict__ gscale,
                                                                   float* ga, float* gb, int vec_ok)
{
    const float s = __ldg(gscale) * (float)(2.0 / (double)n);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    if (vec_ok) {
        const long long n4 = n >> 2;
        for (long long i = tid; i < n4; i += nth) {
            const float4 u = ld_stream4(a + 4 * i), w = ld_stream4(b + 4 * i);
            const float4 g = make_float4(s * (u.x - w.x), s * (u.y - w.y), s * (u.z - w.z), s * (u.w - w.w));
            st_stream4(ga + 4 * i, g);
            if (gb) st_stream4(gb + 4 * i, make_float4(-g.x, -g.y, -g.z, -g.w));
        }
        for (long long i = 4 * n4 + tid; i < n; i += nth) {
            const float g = s * (a[i] - b[i]);
            ga[i] = g;
            if (gb) gb[i] = -g;
        }
    } else {
        for (long long i = tid; i < n; i += nth) {
            const float g = s * (a[i] - b[i]);
            ga[i] = g;
            if (gb) gb[i] = -g;
        }
    }
}

struct SureParams {
    const float* y1;
    const float* y2;
    const float* y;
    const float* b;
    float* out;        // forward: out[3]; backward: unused
    float* g1;
    float* g2;
    const float* gscale;
    RedWorkspace* ws;
    int B, C, H, W, margin_mse, margin_div, averaged_cst, vec;
    float tau, sigma2;
    long long total;
};

__device__ __forceinline__ bool interior(int i, int j, int H, int W, int m) { return i >= m && i < H - m && j >= m && j < W - m; }

__global__ void __launch_bounds__(kRedThreads) sure_loss_kernel(const __grid_constant__ SureParams p)
{
    __shared__ double scratch[64];
    __shared__ bool s_last;
    float mse = 0.f, div = 0.f;
    if (p.vec) {
        // one float4 (4 consecutive columns of one row) per thread and iteration, 32-bit index math
        const unsigned CW = (unsigned)p.W >> 2, n4 = (unsigned)(p.total >> 2);
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
            const unsigned row = q / CW;
            const int j0 = (int)(q - row * CW) * 4, i = (int)(row % (unsigned)p.H);
            const size_t idx = (size_t)q * 4;
            const float4 a1 = ld_stream4(p.y1 + idx), a2 = ld_stream4(p.y2 + idx), yy = ld_stream4(p.y + idx),
                         bb = ld_stream4(p.b + idx);
            const float v1[4] = {a1.x, a1.y, a1.z, a1.w}, v2[4] = {a2.x, a2.y, a2.z, a2.w},
                        vy[4] = {yy.x, yy.y, yy.z, yy.w}, vb[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (interior(i, j0 + e, p.H, p.W, p.margin_mse)) {
                    const float d = v1[e] - vy[e];
                    mse = fmaf(d, d, mse);
                }
                if (interior(i, j0 + e, p.H, p.W, p.margin_div)) div = fmaf(vb[e], v2[e] - v1[e], div);
            }
        }
    } else {
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
             idx += (long long)gridDim.x * blockDim.x) {
            const int j = (int)(idx % p.W);
            const int i = (int)((idx / p.W) % p.H);
            const float a1 = __ldcs(p.y1 + idx);
            if (interior(i, j, p.H, p.W, p.margin_mse)) {
                const float d = a1 - __ldcs(p.y + idx);
                mse = fmaf(d, d, mse);
            }
            if (interior(i, j, p.H, p.W, p.margin_div)) div = fmaf(__ldcs(p.b + idx), __ldcs(p.y2 + idx) - a1, div);
        }
    }
    double v[2] = {(double)mse, (double)div};
    if (grid_finish<2>(v, p.ws, scratch, &s_last)) {
        const double planes = (double)p.B * p.C;
        const double n_mse = planes * (p.H - 2 * p.margin_mse) * (double)(p.W - 2 * p.margin_mse);
        const double n_div = planes * (p.H - 2 * p.margin_div) * (double)(p.W - 2 * p.margin_div);
        const double m = v[0] / n_mse, d = v[1] / ((double)p.tau * n_div);
        const double cst = p.averaged_cst ? (double)p.sigma2 : (double)p.sigma2 / (double)p.B;
        p.out[0] = (float)(m + 2.0 * (double)p.sigma2 * d - cst);
        p.out[1] = (float)m;
        p.out[2] = (float)d;
    }
}

__global__ void __launch_bounds__(kRedThreads) sure_loss_backward_kernel(const __grid_constant__ SureParams p)
{
    const double planes = (double)p.B * p.C;
    const double n_mse = planes * (p.H - 2 * p.margin_mse) * (double)(p.W - 2 * p.margin_mse);
    const double n_div = planes * (p.H - 2 * p.margin_div) * (double)(p.W - 2 * p.margin_div);
    const float g = __ldg(p.gscale);
    const float k_mse = g * (float)(2.0 / n_mse);
    const float k_div = g * (float)(2.0 * (double)p.sigma2 / ((double)p.tau * n_div));
    if (p.vec) {
        const unsigned CW = (unsigned)p.W >> 2, n4 = (unsigned)(p.total >> 2);
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
            const unsigned row = q / CW;
            const int j0 = (int)(q - row * CW) * 4, i = (int)(row % (unsigned)p.H);
            const size_t idx = (size_t)q * 4;
            const float4 a1 = ld_stream4(p.y1 + idx), yy = ld_stream4(p.y + idx), bb = ld_stream4(p.b + idx);
            const float v1[4] = {a1.x, a1.y, a1.z, a1.w}, vy[4] = {yy.x, yy.y, yy.z, yy.w}, vb[4] = {bb.x, bb.y, bb.z, bb.w};
            float o1[4], o2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = interior(i, j0 + e, p.H, p.W, p.margin_mse) ? k_mse * (v1[e] - vy[e]) : 0.f;
                const float d = interior(i, j0 + e, p.H, p.W, p.margin_div) ? k_div * vb[e] : 0.f;
                o1[e] = a - d;
                o2[e] = d;
            }
            st_stream4(p.g1 + idx, make_float4(o1[0], o1[1], o1[2], o1[3]));
            st_stream4(p.g2 + idx, make_float4(o2[0], o2[1], o2[2], o2[3]));
        }
        return;
    }
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p.W);
        const int i = (int)((idx / p.W) % p.H);
        float a = 0.f, d = 0.f;
        if (interior(i, j, p.H, p.W, p.margin_mse)) a = k_mse * (__ldcs(p.y1 + idx) - __ldcs(p.y + idx));
        if (interior(i, j, p.H, p.W, p.margin_div)) d = k_div * __ldcs(p.b + idx);
        __stcs(p.g1 + idx, a - d);
        __stcs(p.g2 + idx, d);
    }
}

__global__ void __launch_bounds__(256) sure_perturb_vec_kernel(const float* __restrict__ y, const float* __restrict__ draw,
                                                               int H, int W, int margin, float tau, long long total,
                                                               float* out, float* b_out)
{
    const int Hi = H - 2 * margin, Wi = W - 2 * margin;
    const unsigned CW = (unsigned)W >> 2, n4 = (unsigned)(total >> 2);
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
        const unsigned row = q / CW;
        const int j0 = (int)(q - row * CW) * 4;
        const unsigned plane = row / (unsigned)H;
        const int i = (int)(row - plane * (unsigned)H);
        const size_t idx = (size_t)q * 4;
        const float4 yy = ld_stream4(y + idx);
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (i >= margin && i < H - margin) {
            const float* drow = draw + ((size_t)plane * Hi + (i - margin)) * Wi - margin;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j0 + e >= margin && j0 + e < W - margin) b[e] = __ldcs(drow + j0 + e);
        }
        st_stream4(out + idx, make_float4(__fadd_rn(yy.x, __fmul_rn(b[0], tau)), __fadd_rn(yy.y, __fmul_rn(b[1], tau)),
                                          __fadd_rn(yy.z, __fmul_rn(b[2], tau)), __fadd_rn(yy.w, __fmul_rn(b[3], tau))));
        if (b_out) st_stream4(b_out + idx, make_float4(b[0], b[1], b[2], b[3]));
    }
}

__global__ void __launch_bounds__(256) sure_perturb_kernel(const float* __restrict__ y, const float* __restrict__ draw,
                                                           int H, int W, int margin, float tau, long long total,
                                                           float* out, float* b_out)
{
    const int Hi = H - 2 * margin, Wi = W - 2 * margin;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        const long long t = idx / W;
        const int i = (int)(t % H);
        float b = 0.f;
        if (interior(i, j, H, W, margin)) b = __ldcs(draw + ((t / H) * Hi + (i - margin)) * Wi + (j - margin));
        out[idx] = __fadd_rn(y[idx], __fmul_rn(b, tau));
        if (b_out) b_out[idx] = b;
    }
}

__global__ void __launch_bounds__(256) add_noise_vec_kernel(const float* __restrict__ y, const float* __restrict__ n,
                                                            long long total4, float sigma, float* out)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        const float4 a = ld_stream4(y + 4 * q), b = ld_stream4(n + 4 * q);
        st_stream4(out + 4 * q, make_float4(__fadd_rn(a.x, __fmul_rn(b.x, sigma)), __fadd_rn(a.y, __fmul_rn(b.y, sigma)),
                                            __fadd_rn(a.z, __fmul_rn(b.z, sigma)), __fadd_rn(a.w, __fmul_rn(b.w, sigma))));
    }
}

__global__ void __launch_bounds__(256) add_noise_kernel(const float* __restrict__ y, const float* __restrict__ n,
                                                        long long total, float sigma, float* out)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
        out[idx] = __fadd_rn(y[idx], __fmul_rn(n[idx], sigma));
}

// circular shift of every plane by (sy, sx): out[i][j] = in[(i - sy) mod H][(j - sx) mod W]   (torch.roll; the
// group action of deepinv's Shift transform used by the "ei-shift" ablation, src/losses/__init__.py:91-94)
__global__ void __launch_bounds__(256) roll_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                   int sy, int sx, long long total)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        const long long t = idx / W;
        const int i = (int)(t % H);
        int si = i - sy, sj = j - sx;
        si += si < 0 ? H : 0;
        sj += sj < 0 ? W : 0;
        out[idx] = __ldg(in + (t / H) * (long long)H * W + (long long)si * W + sj);
    }
}

static unsigned red_grid(long long n, int sm_count)
{
    const long long want = (n + (long long)kRedThreads * 16 - 1) / ((long long)kRedThreads * 16);
    return (unsigned)std::max<long long>(1, std::min<long long>(want, std::min(kRedMaxBlocks, sm_count * 4)));
}

static unsigned ew_grid(long long n, int sm_count)
{
    return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sm_count * 16));
}

}  // namespace sei

using namespace sei;

extern "C" long long sei_reduce_workspace_bytes(void) { return kRedWorkspaceBytes; }

extern "C" int sei_mse_f32(const float* a, const float* b, long long n, float* out, void* workspace, void* stream)
{
    SEI_REQUIRE(a && b && out && workspace, "null pointer argument");
    SEI_REQUIRE(n > 0, "empty reduction");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const int vec_ok = aligned16(a) && aligned16(b);
    mse_kernel<<<red_grid(n, dp.sm_count), kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a, b, n, out, reinterpret_cast<RedWorkspace*>(workspace), vec_ok);
    return finish_launch("mse_kernel");
}

extern "C" int sei_mse_backward_f32(const float* a, const float* b, long long n, const float* gscale,
                                    float* ga, float* gb, void* stream)
{
    SEI_REQUIRE(a && b && gscale && ga, "null pointer argument");
    SEI_REQUIRE(n > 0, "empty tensor");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const int vec_ok = aligned16(a) && aligned16(b) && aligned16(ga) && (!gb || aligned16(gb));
    mse_backward_kernel<<<ew_grid((n + 3) / 4, dp.sm_count), kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a, b, n, gscale, ga, gb, vec_ok);
    return finish_launch("mse_backward_kernel");
}

static int sure_fill(SureParams& p, int B, int C, int H, int W, int margin_mse, int margin_div, float tau, float sigma2)
{
    SEI_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad shape");
    SEI_REQUIRE(margin_mse >= 0 && margin_div >= 0 && 2 * margin_mse < H && 2 * margin_mse < W &&
                    2 * margin_div < H && 2 * margin_div < W,
                "margins (%d, %d) leave no interior in a %dx%d image", margin_mse, margin_div, H, W);
    SEI_REQUIRE(tau != 0.f, "tau must be non-zero");
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.margin_mse = margin_mse; p.margin_div = margin_div;
    p.tau = tau; p.sigma2 = sigma2;
    p.total = (long long)B * C * H * W;
    return 0;
}

extern "C" int sei_sure_loss_f32(const float* y1, const float* y2, const float* y, const float* b,
                                 int B, int C, int H, int W, int margin_mse, int margin_div,
                                 float tau, float sigma2, int averaged_cst, float* out, void* workspace, void* stream)
{
    SEI_REQUIRE(y1 && y2 && y && b && out && workspace, "null pointer argument");
    SureParams p = {};
    int rc = sure_fill(p, B, C, H, W, margin_mse, margin_div, tau, sigma2);
    if (rc) return rc;
    DeviceProps dp;
    rc = get_device_props(&dp);
    if (rc) return rc;
    p.y1 = y1; p.y2 = y2; p.y = y; p.b = b; p.out = out; p.averaged_cst = averaged_cst;
    p.ws = reinterpret_cast<RedWorkspace*>(workspace);
    p.vec = (W % 4 == 0) && p.total < (1ll << 32) && aligned16(y1) && aligned16(y2) && aligned16(y) && aligned16(b);
    sure_loss_kernel<<<red_grid(p.total, dp.sm_count), kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return finish_launch("sure_loss_kernel");
}

extern "C" int sei_sure_loss_backward_f32(const float* y1, const float* y, const float* b,
                                          int B, int C, int H, int W, int margin_mse, int margin_div,
                                          float tau, float sigma2, const float* gscale, float* g1, float* g2,
                                          void* stream)
{
    SEI_REQUIRE(y1 && y && b && gscale && g1 && g2, "null pointer argument");
    SureParams p = {};
    int rc = sure_fill(p, B, C, H, W, margin_mse, margin_div, tau, sigma2);
    if (rc) return rc;
    DeviceProps dp;
    rc = get_device_props(&dp);
    if (rc) return rc;
    p.y1 = y1; p.y = y; p.b = b; p.gscale = gscale; p.g1 = g1; p.g2 = g2;
    p.vec = (W % 4 == 0) && p.total < (1ll << 32) && aligned16(y1) && aligned16(y) && aligned16(b) && aligned16(g1) && aligned16(g2);
    sure_loss_backward_kernel<<<ew_grid(p.vec ? p.total / 4 : p.total, dp.sm_count), kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return finish_launch("sure_loss_backward_kernel");
}

extern "C" int sei_sure_perturb_f32(const float* y, const float* draw, int B, int C, int H, int W,
                                    int margin, float tau, float* out, float* b_out, void* stream)
{
    SEI_REQUIRE(y && draw && out, "null pointer argument");
    SEI_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && margin >= 0 && 2 * margin < H && 2 * margin < W, "bad shape / margin");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long total = (long long)B * C * H * W;
    if ((W % 4 == 0) && total < (1ll << 32) && aligned16(y) && aligned16(out) && (!b_out || aligned16(b_out))) {
        sure_perturb_vec_kernel<<<ew_grid(total / 4, dp.sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            y, draw, H, W, margin, tau, total, out, b_out);
        return finish_launch("sure_perturb_vec_kernel");
    }
    sure_perturb_kernel<<<ew_grid(total, dp.sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        y, draw, H, W, margin, tau, total, out, b_out);
    return finish_launch("sure_perturb_kernel");
}

extern "C" int sei_roll_f32(const float* in, float* out, long long planes, int H, int W, int shift_h, int shift_w, void* stream)
{
    SEI_REQUIRE(in && out && in != out, "null or aliased pointer argument");
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0, "bad shape");
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const int sy = ((shift_h % H) + H) % H, sx = ((shift_w % W) + W) % W;
    const long long total = planes * (long long)H * W;
    roll_kernel<<<ew_grid(total, dp.sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, H, W, sy, sx, total);
    return finish_launch("roll_kernel");
}

extern "C" int sei_add_noise_f32(const float* y, const float* noise, long long n, float sigma, float* out, void* stream)
{
    SEI_REQUIRE(y && noise && out, "null pointer argument");
    SEI_REQUIRE(n > 0, "empty tensor");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    if (n % 4 == 0 && aligned16(y) && aligned16(noise) && aligned16(out)) {
        add_noise_vec_kernel<<<ew_grid(n / 4, dp.sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, noise, n / 4, sigma, out);
        return finish_launch("add_noise_vec_kernel");
    }
    add_noise_kernel<<<ew_grid(n, dp.sm_count), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, noise, n, sigma, out);
    return finish_launch("add_noise_kernel");
}
