// down.cu -- SR physics: antialiased bicubic decimation A and its transpose A^T
// (reference: src/physics/downsampling/__init__.py:16-19 F.interpolate(bicubic, antialias=True);
//  autograd backward / true adjoint :21-31; plain bicubic upsample :32-35).
//
// A is separable: per axis, output i reads the 4*rate inputs [rate*i - OFF, rate*i - OFF + 4*rate)
// with fixed polyphase weights; the first/last two outputs of an axis have truncated,
// renormalised windows (ATen semantics, see sei::aa_axis_weights).
//
// Forward tiled kernel (down_band_kernel): one CTA = TH output rows x full width of one plane.
// The rate*TH + 3*rate input rows are streamed through a double-buffered shared-memory stage by
// bulk async copies (TMA engine, mbarrier completion) in chunks of CH rows; each chunk is
// filtered horizontally (decimating by rate) into a resident intermediate of width W/rate, and
// a final vertical pass produces the band.  Input is read from HBM once (+ 3*rate halo rows per
// band from L2).  Transpose tiled kernel (down_t_band_kernel): one CTA = TH input-resolution
// rows; the ~TH/rate + 4 gradient rows it depends on are staged by one bulk copy; vertical then
// horizontal 4-tap polyphase passes.  Direct kernels cover every other shape.
#include "sei_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace sei {

constexpr int kDownThreads = 256;

// interior window start: xmin(i) = rate*i - aa_off(rate), aa_off = ceil(1.5*rate - 0.5) = 3, 4, 6 for rate 2, 3, 4
__host__ __device__ constexpr int aa_off(int rate) { return (3 * rate) / 2; }

struct DownParams {
    const float* x;      // fwd: input planes H x W ; transpose: gy planes Ho x Wo
    float* y;            // fwd: output planes Ho x Wo ; transpose: gx planes H x W
    const float* noise;
    float sigma;
    int H, W, Ho, Wo;
    int TH, nbands, CH;
    float wint[kAaMaxTaps];   // interior weights
};

struct BorderTab {      // weights of the 4 border outputs of one axis: indices 0, 1, n-2, n-1
    float w[4][kAaMaxTaps];
    int xmin[4];
    int xsize[4];
};

// forward kernel parameters: border tables are computed on the host once per launch (the first version computed them
// in 8 threads of every CTA while the other 248 waited at a barrier: 35 % of its stall samples)
struct DownFwdParams {
    DownParams d;
    BorderTab colTab, rowTab;
};

__device__ __forceinline__ int border_slot(int i, int n) { return i < 2 ? i : (i >= n - 2 ? i - (n - 4) : -1); }

// WT: compile-time INPUT width (0 = run time).  Interior outputs (all but the first / last two of an axis) use the
// fixed polyphase taps from the constant bank in branch-free loops; the few border outputs are handled by separate
// small loops with per-CTA weight tables, so no warp ever executes both paths (the first version did, in every warp).
template <int R, int WT>
__global__ void __launch_bounds__(kDownThreads) down_band_kernel(const __grid_constant__ DownFwdParams fp)
{
    constexpr int T = 4 * R, OFF = aa_off(R);
    constexpr int NT = kDownThreads;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    const DownParams& p = fp.d;
    const BorderTab& colTab = fp.colTab;
    const BorderTab& rowTab = fp.rowTab;

    const int H = p.H, W = WT ? WT : p.W, Ho = p.Ho, Wo = W / R, CW = Wo >> 2;
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int i0 = band * p.TH;
    const int th = min(p.TH, Ho - i0);
    const int in_lo = max(0, R * i0 - OFF);
    const int in_hi = min(H, R * (i0 + th - 1) - OFF + T);
    const int nin = in_hi - in_lo;
    const int CH = p.CH;
    const int nchunks = (nin + CH - 1) / CH;

    float* sTmp = reinterpret_cast<float*>(smem_raw);                  // [R*TH + 3R][Wo]
    float* sStage = sTmp + (R * p.TH + 3 * R) * Wo;                     // [2][CH][W]
    const unsigned char* xplane = reinterpret_cast<const unsigned char*>(p.x + (size_t)plane * H * W);
    const uint32_t row_bytes = (uint32_t)W * 4u;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = min(CH, nin);
        mbar_arrive_expect_tx(&bar[0], (uint32_t)n * row_bytes);
        bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sStage), xplane, H, row_bytes, in_lo, n, &bar[0]);
    }

    constexpr int SKIP = 4 * ((OFF + 3) / 4) - OFF;
    constexpr int NV = (SKIP + 7 * R + 3) / 4;
    const int IW = CW - 2;                                             // interior 4-output groups per row
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (threadIdx.x == 0 && c + 1 < nchunks) {
            const int n = min(CH, nin - (c + 1) * CH);
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar[buf ^ 1], (uint32_t)n * row_bytes);
            bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sStage + (buf ^ 1) * CH * W), xplane, H,
                                    row_bytes, in_lo + (c + 1) * CH, n, &bar[buf ^ 1]);
        }
        mbar_wait(&bar[buf], (c >> 1) & 1);
        const int nrows = min(CH, nin - c * CH);
        const float* stage = sStage + buf * CH * W;
        float* tmp = sTmp + c * CH * Wo;
        // ---- horizontal pass, interior groups: tmp[row][j] = sum_t w[t] * in[row][R*j - OFF + t]
        for (int item = threadIdx.x; item < nrows * IW; item += NT) {
            const int r = item / IW, j4 = 1 + item - r * IW;
            const float* src = stage + r * W + 4 * R * j4 - (OFF + SKIP);
            float v[4 * NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(src + 4 * q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
            float out[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < T; ++t) a = fmaf(p.wint[t], v[SKIP + R * o + t], a);
                out[o] = a;
            }
            *reinterpret_cast<float4*>(tmp + r * Wo + 4 * j4) = make_float4(out[0], out[1], out[2], out[3]);
        }
        // ---- horizontal pass, the 8 outputs of the first and last group of every row
        for (int item = threadIdx.x; item < nrows * 8; item += NT) {
            const int r = item >> 3, e = item & 7;
            const int j = e < 4 ? e : Wo - 8 + e;
            const int k = border_slot(j, Wo);
            const float* row = stage + r * W;
            float a = 0.f;
            if (k >= 0) {
                const float* src = row + colTab.xmin[k];
                for (int t = 0; t < colTab.xsize[k]; ++t) a = fmaf(colTab.w[k][t], src[t], a);
            } else {
                const float* src = row + R * j - OFF;
#pragma unroll
                for (int t = 0; t < T; ++t) a = fmaf(p.wint[t], src[t], a);
            }
            tmp[r * Wo + j] = a;
        }
        __syncthreads();
    }

    // ---- vertical pass: y[i][j] = sum_t w[t] * sTmp[R*i - OFF + t - in_lo][j]
    // Register-blocked: a work item owns 4 consecutive output rows x 4 columns and streams the 3R + T intermediate rows
    // they share once (14 shared loads per 4 output vectors at R = 2 instead of 32; the per-row version was bound by
    // the shared-memory pipe).  Groups that contain one of the image's four border rows, or a ragged tail, take the
    // per-row path with the border weight tables.
    float* yplane = p.y + (size_t)plane * Ho * Wo;
    const float* nplane = p.noise ? p.noise + (size_t)plane * Ho * Wo : nullptr;
    const int ngrp = (th + 3) >> 2;
    for (int item = threadIdx.x; item < ngrp * CW; item += NT) {
        const int gq = item / CW, j4 = item - gq * CW;
        const int r0 = 4 * gq, nr = min(4, th - r0);
        const int i = i0 + r0;
        if (nr == 4 && i >= 2 && i + 3 < Ho - 2) {
            const int g = i * Wo + 4 * j4;
            float4 nz[4];
            if (nplane) {
#pragma unroll
                for (int o = 0; o < 4; ++o) nz[o] = ld_stream4(nplane + g + o * Wo);
            }
            float4 acc[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* src = sTmp + (R * i - OFF - in_lo) * Wo + 4 * j4;
#pragma unroll
            for (int q = 0; q < 3 * R + T; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(src + q * Wo);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int t = q - R * o;
                    if (t >= 0 && t < T) {
                        const float w = p.wint[t];
                        acc[o].x = fmaf(w, v.x, acc[o].x); acc[o].y = fmaf(w, v.y, acc[o].y);
                        acc[o].z = fmaf(w, v.z, acc[o].z); acc[o].w = fmaf(w, v.w, acc[o].w);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                if (nplane) {
                    acc[o].x = fmaf(p.sigma, nz[o].x, acc[o].x); acc[o].y = fmaf(p.sigma, nz[o].y, acc[o].y);
                    acc[o].z = fmaf(p.sigma, nz[o].z, acc[o].z); acc[o].w = fmaf(p.sigma, nz[o].w, acc[o].w);
                }
                st_stream4(yplane + g + o * Wo, acc[o]);
            }
            continue;
        }
        for (int rr = 0; rr < nr; ++rr) {
            const int ii = i + rr;
            const int g = ii * Wo + 4 * j4;
            float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nplane) nz = ld_stream4(nplane + g);
            const int k = border_slot(ii, Ho);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < 0) {
                const float* src = sTmp + (R * ii - OFF - in_lo) * Wo + 4 * j4;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const float4 v = *reinterpret_cast<const float4*>(src + t * Wo);
                    const float w = p.wint[t];
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
            } else {
                const float* src = sTmp + (rowTab.xmin[k] - in_lo) * Wo + 4 * j4;
                for (int t = 0; t < rowTab.xsize[k]; ++t) {
                    const float4 v = *reinterpret_cast<const float4*>(src + t * Wo);
                    const float w = rowTab.w[k][t];
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
            }
            if (nplane) {
                acc.x = fmaf(p.sigma, nz.x, acc.x); acc.y = fmaf(p.sigma, nz.y, acc.y);
                acc.z = fmaf(p.sigma, nz.z, acc.z); acc.w = fmaf(p.sigma, nz.w, acc.w);
            }
            st_stream4(yplane + g, acc);
        }
    }
}

// ------------------------------------------------------------------ transpose, tiled
// contributors of input-resolution index m along one axis: up to NQ (output index, weight) pairs
constexpr int kNQ = 6;
struct Contrib {
    int idx[kNQ];
    float w[kNQ];
};

SEI_HD void aa_contributors(int m, int in_size, int out_size, int rate, Contrib& c)
{
    const int off = aa_off(rate);
    const int ic = (m + off) / rate;
    int n = 0;
#pragma unroll 1
    for (int i = ic - 4; i <= ic + 1; ++i) {
        int idx = 0;
        float wv = 0.f;
        if (i >= 0 && i < out_size) {
            float w[kAaMaxTaps];
            int xmin, xsize;
            aa_axis_weights(i, in_size, rate, w, xmin, xsize);
            const int t = m - xmin;
            if (t >= 0 && t < xsize) {
                float sel = 0.f;
#pragma unroll
                for (int q = 0; q < kAaMaxTaps; ++q) sel = (q == t) ? w[q] : sel;
                idx = i;
                wv = sel;
            }
        }
        c.idx[n] = idx;
        c.w[n] = wv;
        ++n;
    }
}

constexpr int kBorderCols = 20;   // border columns per side handled through contributor tables
constexpr int kBorderRows = 16;   // border rows per side

// transpose kernel parameters: contributor tables of the border rows / columns, computed on the host
struct DownTParams {
    DownParams d;
    Contrib colL[kBorderCols], colR[kBorderCols], rowLo[kBorderRows], rowHi[kBorderRows];
};

// gx = A^T gy.  Interior rows / columns of gx receive exactly four outputs each, (m+OFF)/R - q with tap
// (m+OFF)%R + R*q, q = 0..3 (a 4-tap polyphase upsampler); rows / columns near the border go through
// contributor tables.  Interior and border are separate loops (no divergent warps).
template <int R, int WT>
__global__ void __launch_bounds__(kDownThreads) down_t_band_kernel(const __grid_constant__ DownTParams tp)
{
    constexpr int OFF = aa_off(R);
    constexpr int NT = kDownThreads;
    constexpr bool kStaticPhase = (4 % R) == 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const DownParams& p = tp.d;

    const int H = p.H, W = WT ? WT : p.W, Ho = p.Ho, Wo = W / R, CWo = Wo >> 2;
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int m0 = band * p.TH;
    const int th = min(p.TH, H - m0);
    const int g_lo = max(0, (m0 + OFF) / R - 4);
    const int g_hi = min(Ho, (m0 + th - 1 + OFF) / R + 2);
    const int ng = g_hi - g_lo;
    const int GR = p.TH / R + 8;                     // allocated gradient rows

    float* sG = reinterpret_cast<float*>(smem_raw);             // [GR][Wo]
    float* sTmp = sG + GR * Wo;                                  // [TH][Wo]

    const int NL = min(W, 5 * R - OFF);                          // columns [0, NL) are border
    const int NR0 = max(NL, R * (Wo - 2) - OFF);                 // columns [NR0, W) are border
    const int ML = min(H, 5 * R - OFF), MR0 = max(ML, R * (Ho - 2) - OFF);   // same for rows
    const int n4_lo = (NL + 3) >> 2, n4_hi = max(n4_lo, NR0 >> 2);         // float4 groups that are fully interior

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && ng > 0) {
        const uint32_t row_bytes = (uint32_t)Wo * 4u;
        mbar_arrive_expect_tx(&bar, (uint32_t)ng * row_bytes);
        bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sG),
                                reinterpret_cast<const unsigned char*>(p.x + (size_t)plane * Ho * Wo), Ho, row_bytes,
                                g_lo, ng, &bar);
    }
    if (ng > 0) mbar_wait(&bar, 0);

    // ---- vertical pass: sTmp[m][j] = sum_q w_q * gy[i_q][j]   (rows are warp-uniform)
    for (int item = threadIdx.x; item < th * CWo; item += NT) {
        const int r = item / CWo, j4 = item - r * CWo;
        const int m = m0 + r;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m >= ML && m < MR0) {
            const int ph = (m + OFF) % R, ib = (m + OFF) / R;
            const float* src = sG + (ib - g_lo) * Wo + 4 * j4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(src - q * Wo);
                const float w = p.wint[ph + R * q];
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
        } else {
            const Contrib& c = m < ML ? tp.rowLo[m] : tp.rowHi[m - MR0];
#pragma unroll
            for (int q = 0; q < kNQ; ++q) {
                const float w = c.w[q];
                if (w != 0.f) {
                    const float4 v = *reinterpret_cast<const float4*>(sG + (c.idx[q] - g_lo) * Wo + 4 * j4);
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
            }
        }
        *reinterpret_cast<float4*>(sTmp + r * Wo + 4 * j4) = acc;
    }
    __syncthreads();

    // ---- horizontal pass, interior float4 groups: gx[m][n] = sum_q wint[ph + R q] * sTmp[m][jb - q]
    float* gplane = p.y + (size_t)plane * H * W;
    const int IWn = n4_hi - n4_lo;
    for (int item = threadIdx.x; item < th * IWn; item += NT) {
        const int r = item / IWn, n4 = n4_lo + item - r * IWn;
        const float* row = sTmp + r * Wo;
        float out[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int n = 4 * n4 + o;
            const int ph = kStaticPhase ? (o + OFF) % R : (n + OFF) % R;
            const int jb = kStaticPhase ? (4 / R) * n4 + (o + OFF) / R : (n + OFF) / R;
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) a = fmaf(p.wint[ph + R * q], row[jb - q], a);
            out[o] = a;
        }
        st_stream4(gplane + (size_t)(m0 + r) * W + 4 * n4, make_float4(out[0], out[1], out[2], out[3]));
    }
    // ---- horizontal pass, border columns
    const int nbl = 4 * n4_lo, nbr = W - 4 * n4_hi;
    for (int item = threadIdx.x; item < th * (nbl + nbr); item += NT) {
        const int r = item / (nbl + nbr), e = item - r * (nbl + nbr);
        const int n = e < nbl ? e : 4 * n4_hi + (e - nbl);
        const Contrib& c = e < nbl ? tp.colL[e] : tp.colR[e - nbl];
        const float* row = sTmp + r * Wo;
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kNQ; ++q)
            if (c.w[q] != 0.f) a = fmaf(c.w[q], row[c.idx[q]], a);
        gplane[(size_t)(m0 + r) * W + n] = a;
    }
}

// ------------------------------------------------------------------ direct kernels (any shape)
struct DownDirectParams {
    const float* x;
    float* y;
    const float* noise;
    float sigma;
    int H, W, Ho, Wo, rate;
    long long total;
};

__global__ void __launch_bounds__(128) down_direct_kernel(const __grid_constant__ DownDirectParams p)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p.Wo);
        const long long t = idx / p.Wo;
        const int i = (int)(t % p.Ho);
        const float* xp = p.x + (t / p.Ho) * (long long)p.H * p.W;
        float wy[kAaMaxTaps], wx[kAaMaxTaps];
        int ymin, ysize, xmin, xsize;
        aa_axis_weights(i, p.H, p.rate, wy, ymin, ysize);
        aa_axis_weights(j, p.W, p.rate, wx, xmin, xsize);
        float acc = 0.f;
#pragma unroll 1
        for (int a = 0; a < ysize; ++a) {
            const float* row = xp + (size_t)(ymin + a) * p.W + xmin;
            float h = 0.f;
#pragma unroll
            for (int b = 0; b < kAaMaxTaps; ++b)
                if (b < xsize) h = fmaf(wx[b], __ldg(row + b), h);
            float wsel = 0.f;
#pragma unroll
            for (int q = 0; q < kAaMaxTaps; ++q) wsel = (q == a) ? wy[q] : wsel;
            acc = fmaf(wsel, h, acc);
        }
        if (p.noise) acc = fmaf(p.sigma, p.noise[idx], acc);
        p.y[idx] = acc;
    }
}

// transpose: x = gy planes (Ho x Wo), y = gx planes (H x W)
__global__ void __launch_bounds__(128) down_t_direct_kernel(const __grid_constant__ DownDirectParams p)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % p.W);
        const long long t = idx / p.W;
        const int m = (int)(t % p.H);
        const float* gp = p.x + (t / p.H) * (long long)p.Ho * p.Wo;
        Contrib cy, cx;
        aa_contributors(m, p.H, p.Ho, p.rate, cy);
        aa_contributors(n, p.W, p.Wo, p.rate, cx);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < kNQ; ++a) {
            if (cy.w[a] != 0.f) {
                const float* row = gp + (size_t)cy.idx[a] * p.Wo;
                float h = 0.f;
#pragma unroll
                for (int b = 0; b < kNQ; ++b)
                    if (cx.w[b] != 0.f) h = fmaf(cx.w[b], __ldg(row + cx.idx[b]), h);
                acc = fmaf(cy.w[a], h, acc);
            }
        }
        p.y[idx] = acc;
    }
}

// plain bicubic upsample (A = -0.75, align_corners=False, clamped taps)
struct UpParams {
    const float* y;
    float* x;
    int h, w, rate;
    long long total;
};

__global__ void __launch_bounds__(256) up_bicubic_kernel(const __grid_constant__ UpParams p)
{
    const int H = p.h * p.rate, W = p.w * p.rate;
    const float s = 1.0f / (float)p.rate;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        const long long t = idx / W;
        const int i = (int)(t % H);
        const float* yp = p.y + (t / H) * (long long)p.h * p.w;
        const float ry = s * ((float)i + 0.5f) - 0.5f, rx = s * ((float)j + 0.5f) - 0.5f;
        const float fy = floorf(ry), fx = floorf(rx);
        float cy[4], cx[4];
        keys_coeffs(ry - fy, cy);
        keys_coeffs(rx - fx, cx);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = min(max((int)fy - 1 + a, 0), p.h - 1);
            float row = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = min(max((int)fx - 1 + b, 0), p.w - 1);
                row = fmaf(__ldg(yp + (size_t)r * p.w + c), cx[b], row);
            }
            acc = fmaf(row, cy[a], acc);
        }
        p.x[idx] = acc;
    }
}

static void interior_weights(int rate, float* wint)
{
    // any interior output index has the same weights; take one far from both borders
    float w[kAaMaxTaps];
    int xmin, xsize;
    aa_axis_weights(8, 64 * rate, rate, w, xmin, xsize);
    for (int t = 0; t < kAaMaxTaps; ++t) wint[t] = w[t];
}

template <int R, int WT>
static int launch_down_w(const DownParams& p, long long planes, size_t smem, cudaStream_t st, bool transpose)
{
    const unsigned grid = (unsigned)(planes * p.nbands);
    constexpr int OFF = aa_off(R);
    if (transpose) {
        DownTParams tp;
        tp.d = p;
        const int NL = std::min(p.W, 5 * R - OFF), NR0 = std::max(NL, R * (p.Wo - 2) - OFF);
        const int ML = std::min(p.H, 5 * R - OFF), MR0 = std::max(ML, R * (p.Ho - 2) - OFF);
        const int n4_lo = (NL + 3) / 4, n4_hi = std::max(n4_lo, NR0 / 4);
        for (int t = 0; t < 4 * n4_lo; ++t) aa_contributors(t, p.W, p.Wo, R, tp.colL[t]);
        for (int t = 0; t < p.W - 4 * n4_hi; ++t) aa_contributors(4 * n4_hi + t, p.W, p.Wo, R, tp.colR[t]);
        for (int m = 0; m < ML; ++m) aa_contributors(m, p.H, p.Ho, R, tp.rowLo[m]);
        for (int m = MR0; m < p.H; ++m) aa_contributors(m, p.H, p.Ho, R, tp.rowHi[m - MR0]);
        SEI_CUDA(allow_smem(down_t_band_kernel<R, WT>, smem));
        down_t_band_kernel<R, WT><<<grid, kDownThreads, smem, st>>>(tp);
        return finish_launch("down_t_band_kernel");
    }
    DownFwdParams fp;
    fp.d = p;
    for (int k = 0; k < 4; ++k) {
        aa_axis_weights(k < 2 ? k : p.Wo - 4 + k, p.W, R, fp.colTab.w[k], fp.colTab.xmin[k], fp.colTab.xsize[k]);
        aa_axis_weights(k < 2 ? k : p.Ho - 4 + k, p.H, R, fp.rowTab.w[k], fp.rowTab.xmin[k], fp.rowTab.xsize[k]);
    }
    SEI_CUDA(allow_smem(down_band_kernel<R, WT>, smem));
    down_band_kernel<R, WT><<<grid, kDownThreads, smem, st>>>(fp);
    return finish_launch(p.noise ? "down_band_kernel<noise>" : "down_band_kernel");
}

// width-specialised for measurements of 256 x 256 (input width 256 * rate), run-time width otherwise
template <int R>
static int launch_down(const DownParams& p, long long planes, size_t smem, cudaStream_t st, bool transpose)
{
    if (p.W == 256 * R) return launch_down_w<R, 256 * R>(p, planes, smem, st, transpose);
    return launch_down_w<R, 0>(p, planes, smem, st, transpose);
}

static int down_common(const float* in, float* out, long long planes, int H, int W, int rate,
                       const float* noise, float sigma, int path, void* stream, bool transpose)
{
    SEI_REQUIRE(in && out, "null pointer argument");
    SEI_REQUIRE(rate >= 2 && rate <= 4, "rate %d unsupported (2..4)", rate);
    SEI_REQUIRE(planes >= 0 && H >= rate && W >= rate, "bad shape planes=%lld H=%d W=%d rate=%d", planes, H, W, rate);
    SEI_REQUIRE(path >= SEI_PATH_AUTO && path <= SEI_PATH_TILED, "bad path %d", path);
    if (planes == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const int Ho = (int)floor((double)H * (1.0 / (double)rate)), Wo = (int)floor((double)W * (1.0 / (double)rate));

    DownParams p;
    p.x = in; p.y = out; p.noise = noise; p.sigma = sigma;
    p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo;
    interior_weights(rate, p.wint);

    bool tiled_ok = (W % (4 * rate) == 0) && Wo >= 8 && Ho >= 4 && aligned16(in) && aligned16(out) &&
                    (!noise || aligned16(noise));
    size_t smem = 0;
    if (tiled_ok) {
        const size_t budget_dflt = getenv("SEI_DOWN_SMEM_KB") ? (size_t)atoi(getenv("SEI_DOWN_SMEM_KB")) * 1024 : (size_t)110 * 1024;
        const size_t budget = std::min((size_t)dp.smem_optin, budget_dflt);
        if (!transpose) {
            // rows per streamed chunk: one barrier round per chunk, so chunks must be large enough to amortise it
            // (x4, 4 KB rows: 4-row chunks 77 us, 8-row chunks 58 us; profiles/r01_sr_variants.md); halved until a
            // band fits next to the two stages (very wide images)
            int ch = getenv("SEI_DOWN_CH_KB") ? std::max(1, (int)((size_t)atoi(getenv("SEI_DOWN_CH_KB")) * 1024 / ((size_t)W * 4)))
                                              : std::max(8, (int)(16384 / ((size_t)W * 4)));
            int best = 0;
            const int th_max = getenv("SEI_DOWN_TH") ? atoi(getenv("SEI_DOWN_TH")) : 16;
            for (;; ch = std::max(1, ch / 2)) {
                for (int th = 4; th <= th_max; th += 4) {
                    const size_t need = ((size_t)(rate * th + 3 * rate) * Wo + (size_t)2 * ch * W) * 4;
                    if (need <= budget) best = th;
                }
                if (best > 0 || ch == 1) break;
            }
            p.CH = ch;
            p.TH = std::min(best, ((Ho + 3) / 4) * 4);
            p.nbands = p.TH ? (Ho + p.TH - 1) / p.TH : 0;
            smem = ((size_t)(rate * p.TH + 3 * rate) * Wo + (size_t)2 * p.CH * W) * 4;
        } else {
            int best = 0;
            const int th_max = getenv("SEI_DOWNT_TH") ? atoi(getenv("SEI_DOWNT_TH")) : 16;   // 16 measured best for x4 (58 vs 63 us), neutral for x2
            for (int th = 8; th <= th_max; th += 8) {
                const size_t need = ((size_t)(th / rate + 8) * Wo + (size_t)th * Wo) * 4;
                if (need <= budget) best = th;
            }
            p.TH = std::min(best, ((H + 7) / 8) * 8);
            p.CH = 0;
            p.nbands = p.TH ? (H + p.TH - 1) / p.TH : 0;
            smem = ((size_t)(p.TH / rate + 8) * Wo + (size_t)p.TH * Wo) * 4;
            // border-column tables hold kBorderCols entries per side
            // border-column tables hold kBorderCols entries per side (columns outside the interior float4 groups)
            const int NL = 5 * rate - aa_off(rate), NR0 = std::max(NL, rate * (Wo - 2) - aa_off(rate));
            const int n4_lo = (NL + 3) / 4, n4_hi = std::max(n4_lo, NR0 / 4);
            const int ML = std::min(H, NL), MR0 = std::max(ML, rate * (Ho - 2) - aa_off(rate));
            tiled_ok = tiled_ok && 4 * n4_lo <= kBorderCols && (W - 4 * n4_hi) <= kBorderCols && ML <= kBorderRows &&
                       (H - MR0) <= kBorderRows;
        }
        tiled_ok = tiled_ok && p.TH > 0 && planes * p.nbands < (1ll << 31);
    }
    SEI_REQUIRE(path != SEI_PATH_TILED || tiled_ok, "tiled SR path not available for H=%d W=%d rate=%d", H, W, rate);

    if (tiled_ok && path != SEI_PATH_DIRECT) {
        switch (rate) {
        case 2: return launch_down<2>(p, planes, smem, st, transpose);
        case 3: return launch_down<3>(p, planes, smem, st, transpose);
        default: return launch_down<4>(p, planes, smem, st, transpose);
        }
    }
    DownDirectParams d;
    d.x = in; d.y = out; d.noise = noise; d.sigma = sigma;
    d.H = H; d.W = W; d.Ho = Ho; d.Wo = Wo; d.rate = rate;
    d.total = planes * (long long)(transpose ? (long long)H * W : (long long)Ho * Wo);
    const unsigned grid = (unsigned)std::min<long long>((d.total + 127) / 128, (long long)dp.sm_count * 64);
    if (transpose) {
        down_t_direct_kernel<<<grid, 128, 0, st>>>(d);
        return finish_launch("down_t_direct_kernel");
    }
    down_direct_kernel<<<grid, 128, 0, st>>>(d);
    return finish_launch("down_direct_kernel");
}

}  // namespace sei

using namespace sei;

extern "C" int sei_down_aa_f32(const float* x, float* y, long long planes, int H, int W, int rate,
                               const float* noise, float sigma, int path, void* stream)
{
    return down_common(x, y, planes, H, W, rate, noise, sigma, path, stream, false);
}

extern "C" int sei_down_aa_transpose_f32(const float* gy, float* gx, long long planes, int H, int W,
                                         int rate, int path, void* stream)
{
    return down_common(gy, gx, planes, H, W, rate, nullptr, 0.f, path, stream, true);
}

extern "C" int sei_up_bicubic_f32(const float* y, float* x, long long planes, int h, int w, int rate, void* stream)
{
    SEI_REQUIRE(y && x, "null pointer argument");
    SEI_REQUIRE(planes >= 0 && h > 0 && w > 0 && rate >= 1 && rate <= 8, "bad arguments");
    if (planes == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    UpParams p;
    p.y = y; p.x = x; p.h = h; p.w = w; p.rate = rate;
    p.total = planes * (long long)h * rate * w * rate;
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 32);
    up_bicubic_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return finish_launch("up_bicubic_kernel");
}
