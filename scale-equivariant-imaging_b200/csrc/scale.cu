// scale.cu -- the random scale transform (reference: src/transforms.py:5-109).
//
// padded_downsampling_transform = get_downsampling_grid (:27-43) + F.grid_sample(bicubic,
// reflection, align_corners=True) (:77-83).  The grid is a per-image axis-aligned affine map, so
// the 4x4 bicubic gather is SEPARABLE: row taps depend only on the output row, column taps only
// on the output column.  The grid is never materialised; its coordinates are recomputed with the
// reference's exact fp32 rounding sequence (scale_src_coord).
//
// Tiled kernel (scale_band_kernel): one CTA = TH output rows x full width of one plane.  The
// source rows the band touches (a contiguous range after reflection, <= TH/rate + 4 rows) are
// staged by bulk async copies (TMA engine); a vertical 4-tap pass (128-bit shared loads) is
// followed by a horizontal 4-tap gather within shared memory.  Bands whose source range does not
// fit the staging buffer (rates < 0.5) read their taps from global memory instead.
// Direct kernel: one thread per output element, for shapes the tiled path does not take.
#include "tile_ops.cuh"
#include <algorithm>
#include <stdlib.h>

namespace sei {

constexpr int kScaleThreads = 256;

struct ScaleParams {
    const float* x;
    float* out;
    const float* rate;
    const float* center;
    int C, S, TH, nbands, SRC_MAX;
    float two_over_S;
};

template <int ST>
__global__ void __launch_bounds__(kScaleThreads, 2) scale_band_kernel(const __grid_constant__ ScaleParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int s_lo, s_hi;

    const int S = ST ? ST : p.S;
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int b = (int)(plane / p.C);
    const int r0 = band * p.TH;
    const int th = min(p.TH, S - r0);

    float* sSrc = reinterpret_cast<float*>(smem_raw);                    // [SRC_MAX][S]
    float* sTmp = sSrc + (size_t)p.SRC_MAX * S;                           // [TH][S]
    AxisTap* colT = reinterpret_cast<AxisTap*>(sTmp + (size_t)p.TH * S);  // [S]
    AxisTap* rowT = colT + S;                                             // [TH]

    const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
    const float cx = __ldg(p.center + 2 * b), cy = __ldg(p.center + 2 * b + 1);
    const float* xplane = p.x + (size_t)plane * S * S;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_lo = S;
        s_hi = -1;
    }
    __syncthreads();
    if (threadIdx.x < th) {
        AxisTap t;
        scale_axis_tap(r0 + threadIdx.x, S, p.two_over_S, inv_rate, cy, t);
        rowT[threadIdx.x] = t;
        const int lo = min(min(t.idx[0], t.idx[1]), min(t.idx[2], t.idx[3]));
        const int hi = max(max(t.idx[0], t.idx[1]), max(t.idx[2], t.idx[3]));
        atomicMin(&s_lo, lo);
        atomicMax(&s_hi, hi);
    }
    __syncthreads();
    const int lo = s_lo, nsrc = s_hi - s_lo + 1;
    const bool staged = nsrc <= p.SRC_MAX;
    if (staged && threadIdx.x == 0) {
        const uint32_t row_bytes = (uint32_t)S * 4u;
        mbar_arrive_expect_tx(&bar, (uint32_t)nsrc * row_bytes);
        bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sSrc), reinterpret_cast<const unsigned char*>(xplane),
                                S, row_bytes, lo, nsrc, &bar);
    }
    for (int j = threadIdx.x; j < S; j += kScaleThreads) {
        AxisTap t;
        scale_axis_tap(j, S, p.two_over_S, inv_rate, cx, t);
        colT[j] = t;
    }
    if (staged) mbar_wait(&bar, 0);

    // ---- vertical pass: sTmp[r][c] = sum_a wy[r][a] * src[iy[r][a]][c]
    if (staged) {
        if (threadIdx.x < th) {
#pragma unroll
            for (int a = 0; a < 4; ++a) rowT[threadIdx.x].idx[a] -= lo;
        }
        __syncthreads();
        scale_vpass<kScaleThreads, ST>(sSrc, sTmp, S, th, rowT);
    } else {
        __syncthreads();
        scale_vpass<kScaleThreads, ST>(xplane, sTmp, S, th, rowT);
    }
    __syncthreads();

    // ---- horizontal pass: out[r][j] = sum_b wx[j][b] * sTmp[r][ix[j][b]]
    float* oplane = p.out + (size_t)plane * S * S;
    const int ngrp = max(1, kScaleThreads / S);
    const int grp = threadIdx.x / S;
    if (grp < ngrp) {
        for (int j = threadIdx.x - grp * S; j < S; j += kScaleThreads) {
            const AxisTap t = colT[j];
#pragma unroll 4
            for (int r = grp; r < th; r += ngrp)
                __stcs(oplane + (size_t)(r0 + r) * S + j, scale_hgather(sTmp + r * S, t));
        }
    }
}

struct ScaleDirectParams {
    const float* x;
    float* out;
    const float* rate;
    const float* center;
    int C, S;
    float two_over_S;
    long long total;
};

__global__ void __launch_bounds__(256) scale_direct_kernel(const __grid_constant__ ScaleDirectParams p)
{
    const int S = p.S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % S);
        const long long t = idx / S;
        const int i = (int)(t % S);
        const long long plane = t / S;
        const int b = (int)(plane / p.C);
        const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        AxisTap ty, tx;
        scale_axis_tap(i, S, p.two_over_S, inv_rate, __ldg(p.center + 2 * b + 1), ty);
        scale_axis_tap(j, S, p.two_over_S, inv_rate, __ldg(p.center + 2 * b), tx);
        const float* xp = p.x + plane * (long long)S * S;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float* row = xp + (size_t)ty.idx[a] * S;
            float h = __ldg(row + tx.idx[0]) * tx.w[0];
            h = fmaf(__ldg(row + tx.idx[1]), tx.w[1], h);
            h = fmaf(__ldg(row + tx.idx[2]), tx.w[2], h);
            h = fmaf(__ldg(row + tx.idx[3]), tx.w[3], h);
            acc = fmaf(h, ty.w[a], acc);
        }
        p.out[idx] = acc;
    }
}

// transpose of the transform w.r.t. the image (autograd through grid_sample when stop_gradient is off, reference
// src/losses/__init__.py:117-122 with no_grad=False): every output gradient is scattered to its 4x4 taps with
// fp32 atomics into a zeroed buffer.  Not on the default path (stop_gradient=True), so a direct kernel.
__global__ void __launch_bounds__(256) scale_backward_kernel(const __grid_constant__ ScaleDirectParams p)
{
    const int S = p.S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % S);
        const long long t = idx / S;
        const int i = (int)(t % S);
        const long long plane = t / S;
        const int b = (int)(plane / p.C);
        const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        AxisTap ty, tx;
        scale_axis_tap(i, S, p.two_over_S, inv_rate, __ldg(p.center + 2 * b + 1), ty);
        scale_axis_tap(j, S, p.two_over_S, inv_rate, __ldg(p.center + 2 * b), tx);
        float* gx = p.out + plane * (long long)S * S;
        const float g = __ldg(p.x + idx);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float ga = g * ty.w[a];
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(gx + (size_t)ty.idx[a] * S + tx.idx[c], ga * tx.w[c]);
        }
    }
}

__global__ void scale_params_kernel(const float* u_rate, const float* u_center, int B, int n_rates,
                                    float r0, float r1, float r2, float r3, float* rate, float* center)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // sample_from (src/transforms.py:5-11): values[floor(N * U)]
    int k = (int)floorf(__fmul_rn((float)n_rates, u_rate[b]));
    k = min(max(k, 0), n_rates - 1);
    rate[b] = k == 0 ? r0 : (k == 1 ? r1 : (k == 2 ? r2 : r3));
    // center = 2 * U - 1 (:20-22)
    center[2 * b] = __fsub_rn(__fmul_rn(2.0f, u_center[2 * b]), 1.0f);
    center[2 * b + 1] = __fsub_rn(__fmul_rn(2.0f, u_center[2 * b + 1]), 1.0f);
}

// staging rows for a band of th output rows at the smallest rate the reference samples (0.5)
static int scale_src_rows(int th) { return 2 * th + 6; }

int scale_pick_band_rows(int S, int smem_optin, size_t* smem_out)
{
    const size_t budget = std::min((size_t)smem_optin, (size_t)110 * 1024);
    int best = 0;
    for (int th = 8; th <= 64; th += 8) {
        const size_t need = ((size_t)scale_src_rows(th) + th) * S * 4 + (size_t)(S + th) * sizeof(AxisTap);
        if (need <= budget) best = th;
    }
    if (const char* e = getenv("SEI_SCALE_TH")) {     // tuning override
        const int f = atoi(e);
        if (f >= 8 && f % 8 == 0 && ((size_t)scale_src_rows(f) + f) * S * 4 + (size_t)(S + f) * sizeof(AxisTap) <= (size_t)smem_optin)
            best = f;
    }
    if (best == 0) return 0;
    best = std::min(best, ((S + 7) / 8) * 8);
    *smem_out = ((size_t)scale_src_rows(best) + best) * S * 4 + (size_t)(S + best) * sizeof(AxisTap);
    return best;
}

}  // namespace sei

using namespace sei;

extern "C" int sei_scale_transform_f32(const float* x, float* out, int B, int C, int S,
                                       const float* rate, const float* center, int path, void* stream)
{
    SEI_REQUIRE(x && out && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0, "bad shape B=%d C=%d S=%d", B, C, S);
    SEI_REQUIRE(path >= SEI_PATH_AUTO && path <= SEI_PATH_TILED, "bad path %d", path);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long planes = (long long)B * C;
    size_t smem = 0;
    const int TH = (S % 4 == 0 && S >= 8 && aligned16(x) && aligned16(out)) ? scale_pick_band_rows(S, dp.smem_optin, &smem) : 0;
    const bool tiled_ok = TH > 0 && TH <= kScaleThreads && planes * ((S + TH - 1) / TH) < (1ll << 31);
    SEI_REQUIRE(path != SEI_PATH_TILED || tiled_ok, "tiled scale-transform path not available for S=%d", S);
    const float two_over_S = (float)(2.0 / (double)S);
    if (tiled_ok && path != SEI_PATH_DIRECT) {
        ScaleParams p;
        p.x = x; p.out = out; p.rate = rate; p.center = center;
        p.C = C; p.S = S; p.TH = TH; p.nbands = (S + TH - 1) / TH; p.SRC_MAX = scale_src_rows(TH);
        p.two_over_S = two_over_S;
        const unsigned grid = (unsigned)(planes * p.nbands);
        if (S == 256) {
            SEI_CUDA(allow_smem(scale_band_kernel<256>, smem));
            scale_band_kernel<256><<<grid, kScaleThreads, smem, st>>>(p);
        } else if (S == 512) {
            SEI_CUDA(allow_smem(scale_band_kernel<512>, smem));
            scale_band_kernel<512><<<grid, kScaleThreads, smem, st>>>(p);
        } else {
            SEI_CUDA(allow_smem(scale_band_kernel<0>, smem));
            scale_band_kernel<0><<<grid, kScaleThreads, smem, st>>>(p);
        }
        return finish_launch("scale_band_kernel");
    }
    ScaleDirectParams d;
    d.x = x; d.out = out; d.rate = rate; d.center = center; d.C = C; d.S = S;
    d.two_over_S = two_over_S;
    d.total = planes * (long long)S * S;
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_direct_kernel<<<grid, 256, 0, st>>>(d);
    return finish_launch("scale_direct_kernel");
}

extern "C" int sei_scale_transform_backward_f32(const float* gout, float* gx, int B, int C, int S,
                                                const float* rate, const float* center, void* stream)
{
    SEI_REQUIRE(gout && gx && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0, "bad shape B=%d C=%d S=%d", B, C, S);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    ScaleDirectParams d;
    d.x = gout; d.out = gx; d.rate = rate; d.center = center; d.C = C; d.S = S;
    d.two_over_S = (float)(2.0 / (double)S);
    d.total = (long long)B * C * S * S;
    SEI_CUDA(cudaMemsetAsync(gx, 0, (size_t)d.total * sizeof(float), st));
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_backward_kernel<<<grid, 256, 0, st>>>(d);
    return finish_launch("scale_backward_kernel");
}

extern "C" int sei_scale_params_f32(const float* u_rate, const float* u_center, int B,
                                    const float* rates_host, int n_rates, float* rate, float* center,
                                    void* stream)
{
    SEI_REQUIRE(u_rate && u_center && rates_host && rate && center, "null pointer argument");
    SEI_REQUIRE(n_rates >= 1 && n_rates <= 4, "n_rates %d unsupported (1..4)", n_rates);
    if (B <= 0) return 0;
    float r[4] = {rates_host[0], rates_host[0], rates_host[0], rates_host[0]};
    for (int i = 0; i < n_rates; ++i) r[i] = rates_host[i];
    scale_params_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        u_rate, u_center, B, n_rates, r[0], r[1], r[2], r[3], rate, center);
    return finish_launch("scale_params_kernel");
}
