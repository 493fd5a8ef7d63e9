// fused.cu -- the EI re-measurement step as ONE kernel:
//     x2 = T(x_net)   (random scale transform, src/transforms.py:60-109)
//     y' = A(x2) + sigma * n      (physics(x2) of deepinv EILoss, built at src/losses/__init__.py:117-122)
// The reference runs T (grid build + grid_sample), A (three FFTs) and the noise add as ~15
// launches with x2 and the grid round-tripping through HBM.  Here one CTA produces a full-width
// band of TH rows of y': it stages the source rows of x_net it needs (bulk async copies on the
// TMA engine), resamples the TH + 2P rows of x2 the circular blur needs (separable 4-tap
// vertical then horizontal pass) into shared memory, writes the band's own TH rows of x2 to
// global once (they are the EI target), then runs the separable blur passes on the resident x2
// tile and stores y' with the noise epilogue.  HBM traffic: read x_net once, write x2 once, read
// noise once, write y' once = 16 B per element instead of >= 60 B unfused.
#include "tile_ops.cuh"
#include <algorithm>
#include <stdlib.h>

extern "C" int sei_scale_transform_f32(const float*, float*, int, int, int, const float*, const float*, int, void*);

namespace sei {

constexpr int kEiThreads = 256;

struct EiBlurParams {
    const float* x_net;
    float* x2;
    float* y;
    const float* noise;
    const float* rate;
    const float* center;
    const AxisTap* taps;     // optional [B][2][S]: column taps, then row taps, precomputed once per image
    float sigma;
    int C, S, TH, nbands, SRC_MAX, tmp_floats;
    float two_over_S;
    float cv[kMaxK];
    float ch[kMaxK];
};

// taps of every column ([b][0][.]) and row ([b][1][.]) of every image: B * 2 * S entries
__global__ void __launch_bounds__(256) scale_taps_kernel(const float* __restrict__ rate, const float* __restrict__ center,
                                                         int S, float two_over_S, AxisTap* __restrict__ taps, int total)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int i = idx % S, axis = (idx / S) & 1, b = idx / (2 * S);
    AxisTap t;
    scale_axis_tap(i, S, two_over_S, __fdiv_rn(1.0f, __ldg(rate + b)), __ldg(center + 2 * b + axis), t);
    taps[idx] = t;
}

// ST: compile-time image size (0 = run time).  PRE: taps come precomputed from p.taps.
template <int K, bool NOISE, int ST, bool PRE>
__global__ void __launch_bounds__(kEiThreads) ei_blur_band_kernel(const __grid_constant__ EiBlurParams p)
{
    constexpr int P = K / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int s_lo[3], s_hi[3];

    const int S = ST ? ST : p.S;
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int b = (int)(plane / p.C);
    const int r0 = band * p.TH;
    const int th = min(p.TH, S - r0);
    const int n2 = th + 2 * P;                 // rows of x2 the blur of this band reads

    float* sSrc = reinterpret_cast<float*>(smem_raw);                          // [SRC_MAX][S]  (later: x2 tile)
    float* sTmp = sSrc + (size_t)p.SRC_MAX * S;                                 // [tmp_floats]  (later: padded blur intermediate)
    AxisTap* rowT = reinterpret_cast<AxisTap*>(sTmp + p.tmp_floats);            // [TH+2P]
    AxisTap* colT = rowT + (p.TH + 2 * P);                                      // [S] (only without precomputed taps)
    float* sX2 = sSrc;
    float* sMid = sTmp;

    const float* xplane = p.x_net + (size_t)plane * S * S;
    const AxisTap* gcol = PRE ? p.taps + (size_t)b * 2 * S : nullptr;
    const AxisTap* grow = PRE ? gcol + S : nullptr;
    float inv_rate = 0.f, cx = 0.f, cy = 0.f;
    if (!PRE) {
        inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        cx = __ldg(p.center + 2 * b);
        cy = __ldg(p.center + 2 * b + 1);
    }

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 3) {
        s_lo[threadIdx.x] = S;
        s_hi[threadIdx.x] = -1;
    }
    __syncthreads();

    // taps of the n2 rows of x2 (image rows r0-P .. r0+th+P-1, circular); up to three runs of
    // consecutive image rows: wrapped-from-above, in-range, wrapped-from-below
    AxisTap my_tap;
    int my_seg = -1;
    if (threadIdx.x < n2) {
        int i = r0 - P + (int)threadIdx.x;
        my_seg = i < 0 ? 0 : (i >= S ? 2 : 1);
        i = i < 0 ? i + S : (i >= S ? i - S : i);
        if (PRE) my_tap = grow[i];
        else scale_axis_tap(i, S, p.two_over_S, inv_rate, cy, my_tap);
        const int lo = min(min(my_tap.idx[0], my_tap.idx[1]), min(my_tap.idx[2], my_tap.idx[3]));
        const int hi = max(max(my_tap.idx[0], my_tap.idx[1]), max(my_tap.idx[2], my_tap.idx[3]));
        atomicMin(&s_lo[my_seg], lo);
        atomicMax(&s_hi[my_seg], hi);
    }
    __syncthreads();
    const int n_s0 = max(0, s_hi[0] - s_lo[0] + 1), n_s1 = max(0, s_hi[1] - s_lo[1] + 1),
              n_s2 = max(0, s_hi[2] - s_lo[2] + 1);
    const int nsrc = n_s0 + n_s1 + n_s2;
    const bool staged = nsrc <= p.SRC_MAX && p.SRC_MAX > n2;
    if (staged && threadIdx.x == 0) {
        const uint32_t row_bytes = (uint32_t)S * 4u;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(xplane);
        unsigned char* dst = reinterpret_cast<unsigned char*>(sSrc);
        mbar_arrive_expect_tx(&bar, (uint32_t)nsrc * row_bytes);
        if (n_s0) bulk_load_rows_circular(dst, src, S, row_bytes, s_lo[0], n_s0, &bar);
        if (n_s1) bulk_load_rows_circular(dst + (size_t)n_s0 * row_bytes, src, S, row_bytes, s_lo[1], n_s1, &bar);
        if (n_s2) bulk_load_rows_circular(dst + (size_t)(n_s0 + n_s1) * row_bytes, src, S, row_bytes, s_lo[2], n_s2, &bar);
    }
    if (threadIdx.x < n2) {
        if (staged) {
            const int shift = my_seg == 0 ? s_lo[0] : (my_seg == 1 ? s_lo[1] - n_s0 : s_lo[2] - n_s0 - n_s1);
#pragma unroll
            for (int a = 0; a < 4; ++a) my_tap.idx[a] -= shift;
        }
        rowT[threadIdx.x] = my_tap;
    }
    if (!PRE) {
        for (int j = threadIdx.x; j < S; j += kEiThreads) {
            AxisTap t;
            scale_axis_tap(j, S, p.two_over_S, inv_rate, cx, t);
            colT[j] = t;
        }
    }
    __syncthreads();
    if (staged) mbar_wait(&bar, 0);

    // ---- T, vertical 4-tap pass over the n2 rows
    if (staged) scale_vpass<kEiThreads, ST>(sSrc, sTmp, S, n2, rowT);      // shared-memory loads (LDS)
    else scale_vpass<kEiThreads, ST>(xplane, sTmp, S, n2, rowT);            // rates < 0.5: taps straight from global
    __syncthreads();

    // ---- T, horizontal 4-tap gather -> resident x2 tile (+ the band's own rows to global)
    {
        float* x2plane = p.x2 + (size_t)plane * S * S;
        const int ngrp = max(1, kEiThreads / S);
        const int grp = threadIdx.x / S;
        if (grp < ngrp) {
            for (int j = threadIdx.x - grp * S; j < S; j += kEiThreads) {
                const AxisTap t = PRE ? gcol[j] : colT[j];
#pragma unroll 4
                for (int li = grp; li < n2; li += ngrp) {
                    const float v = scale_hgather(sTmp + li * S, t);
                    sX2[li * S + j] = v;
                    if (li >= P && li < P + th) __stcs(x2plane + (size_t)(r0 + li - P) * S + j, v);
                }
            }
        }
    }
    __syncthreads();

    // ---- A: separable circular blur of the resident x2 tile (halo-padded intermediate), noise epilogue
    blur_vpass<K, kEiThreads, ST, true>(sX2, sMid, S, th, p.cv);
    __syncthreads();
    const size_t row0 = ((size_t)plane * S + r0) * S;
    blur_hpass<K, kEiThreads, NOISE, ST, true>(sMid, S, th, p.ch, p.y + row0, NOISE ? p.noise + row0 : nullptr, p.sigma);
}

bool factor_separable_public(const double* k, int kh, int kw, double* v, double* h);

template <int K, bool NOISE, int ST, bool PRE>
static int launch_ei_blur_inst(const EiBlurParams& p, long long planes, size_t smem, cudaStream_t st)
{
    SEI_CUDA(allow_smem(ei_blur_band_kernel<K, NOISE, ST, PRE>, smem));
    ei_blur_band_kernel<K, NOISE, ST, PRE><<<(unsigned)(planes * p.nbands), kEiThreads, smem, st>>>(p);
    return finish_launch("ei_blur_band_kernel");
}

template <int K>
static int launch_ei_blur(const EiBlurParams& p, long long planes, size_t smem, cudaStream_t st)
{
    const bool pre = p.taps != nullptr;
    if (p.S == 256 && pre)
        return p.noise ? launch_ei_blur_inst<K, true, 256, true>(p, planes, smem, st)
                       : launch_ei_blur_inst<K, false, 256, true>(p, planes, smem, st);
    if (pre)
        return p.noise ? launch_ei_blur_inst<K, true, 0, true>(p, planes, smem, st)
                       : launch_ei_blur_inst<K, false, 0, true>(p, planes, smem, st);
    return p.noise ? launch_ei_blur_inst<K, true, 0, false>(p, planes, smem, st)
                   : launch_ei_blur_inst<K, false, 0, false>(p, planes, smem, st);
}

static int ei_tmp_floats(int th, int P, int S)
{
    const int n2 = th + 2 * P, pitch = blur_mid_pitch_host(2 * P + 1, S);
    return std::max(n2 * S, th * pitch);
}

// rows of the source staging buffer (which the x2 tile later re-uses): enough for rate 0.5, or -- experiment
// SEI_EI_NOSTAGE=1 -- just the x2 tile, every band then gathering its taps straight from global / L2
static int ei_src_rows(int th, int P)
{
    const int n2 = th + 2 * P;
    const char* e = getenv("SEI_EI_NOSTAGE");
    return (e && *e == '1') ? n2 : 2 * n2 + 12;
}

static size_t ei_blur_smem(int th, int P, int S)
{
    const int n2 = th + 2 * P;
    return ((size_t)ei_src_rows(th, P) * S + (size_t)ei_tmp_floats(th, P, S)) * 4 + (size_t)(S + n2) * sizeof(AxisTap);
}

}  // namespace sei

using namespace sei;

extern "C" long long sei_ei_workspace_bytes(int B, int S)
{
    return (long long)B * 2 * S * (long long)sizeof(AxisTap);
}

extern "C" int sei_ei_remeasure_f32(const float* x_net, float* x2, float* y_out, int B, int C, int S,
                                    const float* rate, const float* center,
                                    const double* kernel_host, int kh, int kw, int rate_sr,
                                    const float* noise, float sigma, void* workspace, void* stream)
{
    SEI_REQUIRE(x_net && x2 && y_out && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0, "bad shape B=%d C=%d S=%d", B, C, S);
    SEI_REQUIRE((rate_sr == 1) == (kernel_host != nullptr), "pass a blur kernel (rate_sr = 1) or an SR rate (kernel = NULL)");
    SEI_REQUIRE(rate_sr >= 1 && rate_sr <= 4, "rate_sr %d unsupported", rate_sr);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long planes = (long long)B * C;

    if (rate_sr == 1) {
        SEI_REQUIRE(kh >= 1 && kw >= 1 && kh <= kMaxK && kw <= kMaxK, "kernel size %dx%d unsupported", kh, kw);
        SEI_REQUIRE(S >= kh && S >= kw, "image %dx%d smaller than the %dx%d blur kernel", S, S, kh, kw);
        double v[kMaxK], h[kMaxK];
        const bool sep = kh == kw && (kh % 2 == 1) && factor_separable_public(kernel_host, kh, kw, v, h);
        const bool ksupported = kh == 5 || kh == 7 || kh == 9 || kh == 13 || kh == 19;
        const int P = kh / 2;
        int TH = 0;
        if (sep && ksupported && S % 4 == 0 && S >= 4 * ((P + 3) / 4) && aligned16(x_net) && aligned16(x2) &&
            aligned16(y_out) && (!noise || aligned16(noise))) {
            const size_t budget2 = std::min((size_t)dp.smem_optin, (size_t)110 * 1024);
            for (int th = 8; th <= 32; th += 8)
                if (ei_blur_smem(th, P, S) <= budget2 && th + 2 * P <= kEiThreads) TH = th;
            const int forced = getenv("SEI_EI_TH") ? atoi(getenv("SEI_EI_TH")) : 0;
            if (forced > 0 && forced % 8 == 0 && ei_blur_smem(forced, P, S) <= (size_t)dp.smem_optin) TH = forced;
            if (TH == 0 && ei_blur_smem(8, P, S) <= (size_t)dp.smem_optin) TH = 8;   // one CTA per SM
            if (TH) TH = std::min(TH, ((S + 7) / 8) * 8);
        }
        // Measured on B200 (profiles/r01_op_sweep*.md): the fused kernel is shared-memory-bandwidth bound (the 4-tap
        // gather of T, recomputed for the blur's halo rows) and takes 193 us where the two stand-alone kernels take
        // 71 + 71 us, so the pair is the default; SEI_EI_FUSED=1 (or path == SEI_PATH_TILED via the env) selects it.
        const char* fz = getenv("SEI_EI_FUSED");
        const bool want_fused = fz && *fz == '1';
        if (want_fused && TH > 0 && planes * ((S + TH - 1) / TH) < (1ll << 31)) {
            EiBlurParams p;
            p.x_net = x_net; p.x2 = x2; p.y = y_out; p.noise = noise; p.rate = rate; p.center = center;
            p.sigma = sigma; p.C = C; p.S = S; p.TH = TH; p.nbands = (S + TH - 1) / TH;
            p.SRC_MAX = ei_src_rows(TH, P);
            p.tmp_floats = ei_tmp_floats(TH, P, S);
            p.two_over_S = (float)(2.0 / (double)S);
            p.taps = reinterpret_cast<const AxisTap*>(workspace);
            if (workspace) {
                // taps of every row and column, once per image instead of once per band
                SEI_REQUIRE(aligned16(workspace), "workspace must be 16-byte aligned");
                const int total = B * 2 * S;
                scale_taps_kernel<<<(total + 255) / 256, 256, 0, st>>>(rate, center, S, p.two_over_S,
                                                                     reinterpret_cast<AxisTap*>(workspace), total);
                rc = finish_launch("scale_taps_kernel");
                if (rc) return rc;
            }
            for (int t = 0; t < kh; ++t) {
                p.cv[t] = (float)v[kh - 1 - t];
                p.ch[t] = (float)h[kh - 1 - t];
            }
            const size_t smem = ei_blur_smem(TH, P, S);
            switch (kh) {
            case 5: return launch_ei_blur<5>(p, planes, smem, st);
            case 7: return launch_ei_blur<7>(p, planes, smem, st);
            case 9: return launch_ei_blur<9>(p, planes, smem, st);
            case 13: return launch_ei_blur<13>(p, planes, smem, st);
            default: return launch_ei_blur<19>(p, planes, smem, st);
            }
        }
        // shapes the fused kernel does not take: the two stand-alone kernels back to back
        rc = sei_scale_transform_f32(x_net, x2, B, C, S, rate, center, SEI_PATH_AUTO, stream);
        if (rc) return rc;
        return sei_blur_circular_f32(x2, y_out, planes, S, S, kernel_host, kh, kw, 0, noise, sigma, SEI_PATH_AUTO, stream);
    }

    // SR: scale transform, then antialiased decimation with the noise epilogue
    rc = sei_scale_transform_f32(x_net, x2, B, C, S, rate, center, SEI_PATH_AUTO, stream);
    if (rc) return rc;
    return sei_down_aa_f32(x2, y_out, planes, S, S, rate_sr, noise, sigma, SEI_PATH_AUTO, stream);
}
