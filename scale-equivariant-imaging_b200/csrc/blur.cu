// blur.cu -- circular blur A / A^T (reference: src/physics/blur/__init__.py BlurV2.A :205-223,
// Blur(padding="circular") :34-74/:164-194, adjoints :77-134/:225-227).
//
// Tiled kernel (blur_band_kernel): one CTA owns a full-width band of TH output rows of one
// plane.  The TH+2P input rows (circular in y) are contiguous runs of global memory, so they are
// staged into shared memory by 1..3 bulk async copies on the TMA engine (UBLKCP) completing on
// an mbarrier; the horizontal wrap-around is free because whole rows are resident.  The kernel is
// separable: a vertical pass (register-blocked, 8 output rows x 4 columns per thread, taps as
// constant-bank operands) writes an intermediate band to shared memory, a horizontal pass
// (4 outputs per thread from 128-bit shared loads) produces the result, adds the optional
// sigma * noise epilogue and stores with 128-bit writes.  HBM traffic = 4 B read + 4 B written
// per element (halo rows are re-read from L2, not DRAM).
//
// Direct kernel (blur_direct_kernel): any shape / any (also non-separable, even-sized) kernel,
// one thread per output element, taps in the constant bank.
#include "tile_ops.cuh"
#include <algorithm>
#include <stdlib.h>

namespace sei {


struct BlurBandParams {
    const float* x;
    float* y;
    const float* noise;
    float sigma;
    int H, W, TH, nbands;
    long long total_bands;
    float cv[kMaxK];   // correlation-form taps: y[n] = sum_t c[t] x[n + t - P]
    float ch[kMaxK];
};

// Persistent: each CTA walks bands blockIdx.x, blockIdx.x + gridDim.x, ...
//   NSTAGE = 2: two input stages, the bulk copy of band i+1 is issued before band i is filtered;
//   NSTAGE = 1: one input stage, the copy of band i+1 is issued as soon as the vertical pass of band i has consumed
//               the stage, so it overlaps the horizontal pass; the smaller footprint buys one more CTA per SM.
//   H16: horizontal pass with 16 outputs per work item (tile_ops.cuh), needs W % 16 == 0.
template <int K, bool NOISE, int NT, int WT, int NSTAGE, bool H16>
__global__ void __launch_bounds__(NT) blur_band_kernel(const __grid_constant__ BlurBandParams p)
{
    constexpr int P = K / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];

    const int H = p.H, W = WT ? WT : p.W;
    const size_t stage_floats = (size_t)(p.TH + 2 * P) * W;
    float* sIn = reinterpret_cast<float*>(smem_raw);      // [NSTAGE][TH + 2P][W]
    float* sMid = sIn + NSTAGE * stage_floats;             // [TH][pitch]
    const uint32_t row_bytes = (uint32_t)W * 4u;
    const long long total = p.total_bands;

    auto issue = [&](long long w, int buf) {
        const int band = (int)(w % p.nbands);
        const long long plane = w / p.nbands;
        const int r0 = band * p.TH;
        const int rin = min(p.TH, H - r0) + 2 * P;
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)rin * row_bytes);
        bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sIn + buf * stage_floats),
                                reinterpret_cast<const unsigned char*>(p.x + (size_t)plane * H * W), H, row_bytes,
                                r0 - P, rin, &bar[buf]);
    };

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && (long long)blockIdx.x < total) issue(blockIdx.x, 0);

    int it = 0;
    for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int buf = NSTAGE == 2 ? (it & 1) : 0;
        if (NSTAGE == 2 && threadIdx.x == 0 && w + gridDim.x < total) {
            fence_proxy_async();              // stage buf^1 was read by generic loads in iteration it-1
            issue(w + gridDim.x, buf ^ 1);
        }
        const int band = (int)(w % p.nbands);
        const long long plane = w / p.nbands;
        const int r0 = band * p.TH;
        const int th = min(p.TH, H - r0);
        mbar_wait(&bar[buf], NSTAGE == 2 ? ((it >> 1) & 1) : (it & 1));

        blur_vpass<K, NT, WT, true>(sIn + buf * stage_floats, sMid, W, th, p.cv);
        __syncthreads();
        if (NSTAGE == 1 && threadIdx.x == 0 && w + gridDim.x < total) {
            fence_proxy_async();              // the stage was read by generic loads in the vertical pass
            issue(w + gridDim.x, 0);
        }
        const size_t row0 = ((size_t)plane * H + r0) * W;
        if (H16)
            blur_hpass16<K, NT, NOISE, WT, true>(sMid, W, th, p.ch, p.y + row0, NOISE ? p.noise + row0 : nullptr, p.sigma);
        else
            blur_hpass<K, NT, NOISE, WT, true>(sMid, W, th, p.ch, p.y + row0, NOISE ? p.noise + row0 : nullptr, p.sigma);
        __syncthreads();
    }
}

struct BlurDirectParams {
    const float* x;
    float* y;
    const float* noise;
    float sigma;
    int H, W, kh, kw, adjoint;
    long long total;
    float k2[kMaxK * kMaxK];
};

__global__ void __launch_bounds__(256) blur_direct_kernel(const __grid_constant__ BlurDirectParams p)
{
    const int H = p.H, W = p.W, ch = p.kh / 2, cw = p.kw / 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n2 = (int)(idx % W);
        const long long t = idx / W;
        const int n1 = (int)(t % H);
        const float* xp = p.x + (t / H) * (long long)H * W;
        float acc = 0.f;
        for (int i1 = 0; i1 < p.kh; ++i1) {
            int r = p.adjoint ? n1 + i1 - ch : n1 - i1 + ch;
            if (r < 0) r += H;
            if (r >= H) r -= H;
            const float* xrow = xp + (size_t)r * W;
            for (int i2 = 0; i2 < p.kw; ++i2) {
                int c = p.adjoint ? n2 + i2 - cw : n2 - i2 + cw;
                if (c < 0) c += W;
                if (c >= W) c -= W;
                acc = fmaf(p.k2[i1 * p.kw + i2], __ldg(xrow + c), acc);
            }
        }
        if (p.noise) acc = fmaf(p.sigma, p.noise[idx], acc);
        p.y[idx] = acc;
    }
}

// kernel = outer(v, h) within rounding?  (both named families are exactly separable)
static bool factor_separable(const double* k, int kh, int kw, double* v, double* h)
{
    double total = 0, maxabs = 0;
    for (int i = 0; i < kh * kw; ++i) { total += k[i]; maxabs = std::max(maxabs, fabs(k[i])); }
    if (total == 0 || maxabs == 0) return false;
    for (int i = 0; i < kh; ++i) { v[i] = 0; for (int j = 0; j < kw; ++j) v[i] += k[i * kw + j]; }
    for (int j = 0; j < kw; ++j) { h[j] = 0; for (int i = 0; i < kh; ++i) h[j] += k[i * kw + j]; h[j] /= total; }
    double err = 0;
    for (int i = 0; i < kh; ++i)
        for (int j = 0; j < kw; ++j) err = std::max(err, fabs(v[i] * h[j] - k[i * kw + j]));
    return err <= 1e-9 * maxabs;
}

bool factor_separable_public(const double* k, int kh, int kw, double* v, double* h)
{
    return factor_separable(k, kh, kw, v, h);
}

static int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

struct BandConfig {
    int TH, threads, ctas_per_sm, nstage, h16;
    size_t smem;
};

template <int K, bool NOISE, int NT, int WT, int NSTAGE, bool H16>
static int launch_band_inst(const BlurBandParams& p, const BandConfig& c, int sm_count, cudaStream_t st)
{
    SEI_CUDA(allow_smem(blur_band_kernel<K, NOISE, NT, WT, NSTAGE, H16>, c.smem));
    const long long want = (long long)sm_count * c.ctas_per_sm;
    const unsigned grid = (unsigned)std::min<long long>(p.total_bands, want);
    blur_band_kernel<K, NOISE, NT, WT, NSTAGE, H16><<<grid, NT, c.smem, st>>>(p);
    return finish_launch(NOISE ? "blur_band_kernel<noise>" : "blur_band_kernel");
}

// CTA shapes: (128 threads, one stage, 16-wide horizontal items) or (256 threads, two stages, 4-wide items; any
// W % 4 == 0).  Measured on B200 (profiles/r01_blur_variants.md, 128x3x256x256): the first is faster whenever the
// kernel is long or the noise epilogue is on (Gaussian_R2 A+noise 61 vs 70 us, Gaussian_R3 57 vs 63 us), the second
// for short kernels without noise (Box_R3: 38 vs 44 us).
template <int K, bool NOISE, int WT>
static int launch_band_w(const BlurBandParams& p, const BandConfig& c, int sm_count, cudaStream_t st)
{
    if (c.h16) return launch_band_inst<K, NOISE, 128, WT, 1, true>(p, c, sm_count, st);
    return launch_band_inst<K, NOISE, 256, WT, 2, false>(p, c, sm_count, st);
}

// width-specialised instantiations for the benchmark shapes (256: cfg2/cfg4, 512: cfg5), run-time width otherwise
template <int K>
static int launch_band(const BlurBandParams& p, const BandConfig& c, int sm_count, cudaStream_t st)
{
    if (p.W == 256)
        return p.noise ? launch_band_w<K, true, 256>(p, c, sm_count, st) : launch_band_w<K, false, 256>(p, c, sm_count, st);
    if (p.W == 512)
        return p.noise ? launch_band_w<K, true, 512>(p, c, sm_count, st) : launch_band_w<K, false, 512>(p, c, sm_count, st);
    return p.noise ? launch_band_w<K, true, 0>(p, c, sm_count, st) : launch_band_w<K, false, 0>(p, c, sm_count, st);
}

// NSTAGE input stages of (th + 2P) rows + one halo-padded intermediate of th rows
static size_t blur_band_smem(int th, int K, int W, int nstage)
{
    return ((size_t)nstage * (th + 2 * (K / 2)) * W + (size_t)th * blur_mid_pitch_host(K, W)) * 4;
}

// Band height / CTA shape.  Overridable for tuning: SEI_BLUR_TH, SEI_BLUR_H16 (0/1), SEI_BLUR_CTAS.
BandConfig blur_pick_band_config(int H, int W, int K, bool noise, int smem_optin)
{
    BandConfig c = {0, 256, 1, 2, 0, 0};
    const int per_sm = 227 * 1024;
    c.h16 = (W % 16 == 0) ? env_int("SEI_BLUR_H16", (K > 7 || noise) ? 1 : 0) : 0;
    c.threads = c.h16 ? 128 : 256;
    c.nstage = c.h16 ? 1 : 2;
    int best = 0;
    for (int th = 8; th <= 16; th += 8)   // TH = 16 measured best on B200 (profiles/)
        if (blur_band_smem(th, K, W, c.nstage) + 1024 <= (size_t)smem_optin) best = th;
    const int forced = env_int("SEI_BLUR_TH", 0);
    if (forced > 0 && forced % 8 == 0 && blur_band_smem(forced, K, W, c.nstage) + 1024 <= (size_t)smem_optin) best = forced;
    if (best == 0) return c;
    c.TH = std::min(best, ((H + 7) / 8) * 8);
    c.smem = blur_band_smem(c.TH, K, W, c.nstage);
    c.ctas_per_sm = std::max(1, std::min(c.threads == 128 ? 6 : 4, (int)(per_sm / (c.smem + 1024))));
    const int fc = env_int("SEI_BLUR_CTAS", 0);
    if (fc >= 1 && fc <= c.ctas_per_sm) c.ctas_per_sm = fc;
    return c;
}

}  // namespace sei

using namespace sei;

extern "C" int sei_blur_circular_f32(const float* x, float* y, long long planes, int H, int W,
                                     const double* kernel_host, int kh, int kw, int adjoint,
                                     const float* noise, float sigma, int path, void* stream)
{
    SEI_REQUIRE(x && y && kernel_host, "null pointer argument");
    SEI_REQUIRE(planes >= 0 && H > 0 && W > 0, "bad shape planes=%lld H=%d W=%d", planes, H, W);
    SEI_REQUIRE(kh >= 1 && kw >= 1 && kh <= kMaxK && kw <= kMaxK, "kernel size %dx%d unsupported (max %d)", kh, kw, kMaxK);
    SEI_REQUIRE(H >= kh && W >= kw, "image %dx%d smaller than the %dx%d blur kernel", H, W, kh, kw);
    SEI_REQUIRE(path >= SEI_PATH_AUTO && path <= SEI_PATH_TILED, "bad path %d", path);
    SEI_REQUIRE(planes * (long long)H * W < (1ll << 40), "tensor too large");
    if (planes == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;

    double v[kMaxK], h[kMaxK];
    const bool sep = kh == kw && (kh % 2 == 1) && factor_separable(kernel_host, kh, kw, v, h);
    const bool ksupported = kh == 5 || kh == 7 || kh == 9 || kh == 13 || kh == 19;
    const int P = kh / 2;
    BandConfig cfg = {0, 256, 1, 2, 0, 0};
    if (sep && ksupported && (W % 4 == 0) && W >= 4 * ((P + 3) / 4)) cfg = blur_pick_band_config(H, W, kh, noise != nullptr, dp.smem_optin);
    const int TH = cfg.TH;
    const bool aligned = aligned16(x) && aligned16(y) && (!noise || aligned16(noise));
    const bool tiled_ok = TH > 0 && aligned && planes * ((H + TH - 1) / TH) < (1ll << 31);
    SEI_REQUIRE(path != SEI_PATH_TILED || tiled_ok,
                "tiled blur path not available (separable=%d k=%d W=%d aligned=%d)", (int)sep, kh, W, (int)aligned);

    if (tiled_ok && path != SEI_PATH_DIRECT) {
        BlurBandParams p;
        p.x = x; p.y = y; p.noise = noise; p.sigma = sigma;
        p.H = H; p.W = W; p.TH = TH; p.nbands = (H + TH - 1) / TH;
        p.total_bands = planes * p.nbands;
        for (int t = 0; t < kh; ++t) {
            // forward (convolution): c[t] = h[K-1-t]; transpose (correlation): c[t] = h[t]
            p.cv[t] = (float)(adjoint ? v[t] : v[kh - 1 - t]);
            p.ch[t] = (float)(adjoint ? h[t] : h[kh - 1 - t]);
        }
        switch (kh) {
        case 5: return launch_band<5>(p, cfg, dp.sm_count, st);
        case 7: return launch_band<7>(p, cfg, dp.sm_count, st);
        case 9: return launch_band<9>(p, cfg, dp.sm_count, st);
        case 13: return launch_band<13>(p, cfg, dp.sm_count, st);
        default: return launch_band<19>(p, cfg, dp.sm_count, st);
        }
    }

    BlurDirectParams p;
    p.x = x; p.y = y; p.noise = noise; p.sigma = sigma;
    p.H = H; p.W = W; p.kh = kh; p.kw = kw; p.adjoint = adjoint ? 1 : 0;
    p.total = planes * (long long)H * W;
    for (int i = 0; i < kh * kw; ++i) p.k2[i] = (float)kernel_host[i];
    const long long blocks = (p.total + 255) / 256;
    const unsigned grid = (unsigned)std::min<long long>(blocks, (long long)dp.sm_count * 32);
    blur_direct_kernel<<<grid, 256, 0, st>>>(p);
    return finish_launch("blur_direct_kernel");
}
