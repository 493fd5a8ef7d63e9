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


// ================================================================== round-2 kernels
// Forward, "rows" form.  The round-1 kernel filtered horizontally first: every input element went through shared
// memory about 3.7 times (TMA fill, 2x overlapping 128-bit reads of the horizontal windows, the intermediate and its
// 4R-row vertical windows), which at R = 4 is two thirds of the SM's shared-memory bandwidth at full HBM rate
// (measured 0.28 of the copy peak).  Here the VERTICAL pass runs first, straight from global memory: a thread owns one
// 4-column group, walks down the band's input rows with 128-bit read-only loads and keeps the four output rows whose
// windows cover the current input row in rotating accumulators -- every input element is loaded exactly once and
// never touches shared memory.  Only the vertically filtered band (1/R of the input) is written to shared memory; the
// horizontal decimation reads it with 128-bit loads (8 outputs per item, lanes spread over 8 rows of a pitch = 4 mod 32
// intermediate: conflict-free) into an output tile that leaves through ONE bulk async store (TMA engine); the noise
// band arrives in that tile by a bulk async load issued at kernel start.
template <int R> struct DownGeom {
    static constexpr int T = 4 * R, OFF = aa_off(R);
    static constexpr int LPAD = 4 * ((OFF + 3) / 4);          // left margin of an intermediate row (>= OFF, multiple of 4)
    static constexpr int SKIP = LPAD - OFF;
    static constexpr int NO = 8;                               // outputs per horizontal work item
    static constexpr int NV = (SKIP + (NO - 1) * R + T + 3) / 4;
    static constexpr int RING = R <= 2 ? 6 : 4;                // steps of the per-thread input ring (one step = R rows)
};

__host__ __device__ constexpr int down_v_pitch(int W, int lpad) { return blur_pad_pitch_c(lpad + W + 8); }

struct DownRowsParams {
    const float* x;
    float* y;
    const float* noise;
    float sigma;
    int H, W, Ho, Wo, nbands;
    float wint[kAaMaxTaps];
    BorderTab colTab, rowTab;
};

template <int R, int WT, int TH>
__global__ void __launch_bounds__(kDownThreads, 2) down_rows_kernel(const __grid_constant__ DownRowsParams p)
{
    using G = DownGeom<R>;
    constexpr int T = G::T, OFF = G::OFF, NT = kDownThreads;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;

    const int H = p.H, W = WT ? WT : p.W, Ho = p.Ho, Wo = W / R;
    const int pitch = down_v_pitch(W, G::LPAD);
    const int opitch = blur_pad_pitch_c(Wo);                           // padded: the 8 rows of a quarter-warp hit 8 bank groups
    float* sOut = reinterpret_cast<float*>(smem_raw);                  // [TH][opitch]  noise in, result out (in place)
    float* sV = sOut + TH * opitch;                                    // [TH][pitch]
    float4* sRing = reinterpret_cast<float4*>(sV + TH * pitch);        // [RING * R][NT] float4, one column per thread
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int i0 = band * TH, th = min(TH, Ho - i0);
    const float* __restrict__ xplane = p.x + (size_t)plane * H * W;
    const size_t oband = ((size_t)plane * Ho + i0) * Wo;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        if (p.noise) {                                                 // one bulk copy per row of the noise band
            mbar_arrive_expect_tx(&bar, (uint32_t)th * Wo * 4u);
            for (int r = 0; r < th; ++r) bulk_g2s(sOut + r * opitch, p.noise + oband + (size_t)r * Wo, (uint32_t)Wo * 4u, &bar);
        }
    }

    // ---- vertical pass, global -> registers -> sV
    const int CWin = W >> 2;
    const int nseg = CWin >= NT ? 1 : min(TH, NT / CWin);            // row segments so that every thread has a column group
    const int THs = (TH + nseg - 1) / nseg;
    for (int unit = threadIdx.x; unit < nseg * CWin; unit += NT) {
        const int seg = unit / CWin, c4 = unit - seg * CWin;
        const int o_first = seg * THs;
        const int nout = min(THs, th - o_first);
        if (nout <= 0) continue;
        const float* __restrict__ colp = xplane + 4 * c4;
        const int row0 = R * (i0 + o_first) - OFF;                    // input row of step 0, tap 0
        const int nsteps = nout + 3;
        // Input rows reach the thread through a private ring in shared memory filled by per-thread asynchronous 16-byte
        // copies (cp.async, no registers held while in flight): D - 1 steps of R rows are always outstanding, 160 - 192
        // bytes per thread, ~90 KB per SM -- what it takes to keep HBM busy (Little's law at ~1 us latency).  A thread
        // only ever reads what it copied itself, so there is no barrier: cp.async.wait_group is the whole hand-shake.
        // (Register double-buffering kept 32 - 64 bytes per thread in flight: ncu showed 60 % of the stall samples on the
        // first FMA of a step, 0.9 eligible warps per scheduler and 39 % of the DRAM bandwidth.)
        constexpr int D = G::RING;
        float4* ring = sRing + threadIdx.x;                            // [D * R][NT]
        auto issue = [&](int s) {
            if (s < nsteps) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int row = row0 + R * s + k;
                    const bool ok = (unsigned)row < (unsigned)H;
                    cp_async16(ring + ((s % D) * R + k) * NT, colp + (size_t)(ok ? row : 0) * W, ok);
                }
            }
            cp_async_commit();                                          // one group per step, empty past the end
        };
#pragma unroll
        for (int q = 0; q < D - 1; ++q) issue(q);
        float4 acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        float* vcol = sV + o_first * pitch + G::LPAD + 4 * c4;
        for (int s0 = 0; s0 < nsteps; s0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = s0 + u;
                if (s < nsteps) {
                    issue(s + D - 1);
                    cp_async_wait<D - 1>();                             // the copies of step s have landed
                    float4 cur[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) cur[k] = ring[((s % D) * R + k) * NT];
                    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);          // output row o = s starts in slot s % 4
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        float4& a = acc[(u - d) & 3];                   // output row o = s - d
#pragma unroll
                        for (int k = 0; k < R; ++k) {
                            const float w = p.wint[R * d + k];
                            a.x = fmaf(w, cur[k].x, a.x); a.y = fmaf(w, cur[k].y, a.y);
                            a.z = fmaf(w, cur[k].z, a.z); a.w = fmaf(w, cur[k].w, a.w);
                        }
                    }
                    if (s >= 3) *reinterpret_cast<float4*>(vcol + (s - 3) * pitch) = acc[(u + 1) & 3];   // o = s - 3 is complete
                }
            }
        }
        cp_async_wait<0>();
        // the image's first / last two output rows have truncated, renormalised windows: recompute this column group
        for (int o = 0; o < nout; ++o) {
            const int i = i0 + o_first + o;
            const int kb = border_slot(i, Ho);
            if (kb < 0) continue;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t = 0; t < p.rowTab.xsize[kb]; ++t) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(colp + (size_t)(p.rowTab.xmin[kb] + t) * W));
                const float w = p.rowTab.w[kb][t];
                a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
            }
            *reinterpret_cast<float4*>(vcol + o * pitch) = a;
        }
    }
    __syncthreads();
    if (p.noise) mbar_wait(&bar, 0);

    // ---- horizontal decimation: sOut[r][j] = sum_t w[t] * sV[r][R*j - OFF + t]  (+ sigma * noise)
    constexpr int NO = G::NO, NV = G::NV, SKIP = G::SKIP;
    const int nblk = Wo / NO;                                          // Wo % 8 == 0 is required by the dispatcher
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rgroups = (th + 7) >> 3, bgroups = (nblk + 3) >> 2;
    for (int wi = warp; wi < rgroups * bgroups; wi += NT / 32) {
        const int rg = wi % rgroups, bg = wi / rgroups;
        const int r = 8 * rg + (lane & 7), blk = 4 * bg + (lane >> 3);
        if (r >= th || blk >= nblk) continue;
        const float* __restrict__ src = sV + r * pitch + R * NO * blk;   // = row base + LPAD + R*j0 - LPAD
        float v[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * q);
            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
        }
        float out[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) a = fmaf(p.wint[t], v[SKIP + R * o + t], a);
            out[o] = a;
        }
        if (blk == 0 || blk == nblk - 1) {                             // the row's first / last two outputs
            const float* row = sV + r * pitch + G::LPAD;
            auto border = [&](int kb) {
                float a = 0.f;
                for (int t = 0; t < p.colTab.xsize[kb]; ++t) a = fmaf(p.colTab.w[kb][t], row[p.colTab.xmin[kb] + t], a);
                return a;
            };
            if (blk == 0) { out[0] = border(0); out[1] = border(1); }
            if (blk == nblk - 1) { out[NO - 2] = border(2); out[NO - 1] = border(3); }
        }
        float* dst = sOut + r * opitch + NO * blk;
#pragma unroll
        for (int q = 0; q < NO / 4; ++q) {
            float4 o4 = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            if (p.noise) {
                const float4 n4 = *reinterpret_cast<const float4*>(dst + 4 * q);
                o4.x = fmaf(p.sigma, n4.x, o4.x); o4.y = fmaf(p.sigma, n4.y, o4.y);
                o4.z = fmaf(p.sigma, n4.z, o4.z); o4.w = fmaf(p.sigma, n4.w, o4.w);
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = o4;
        }
    }
    fence_proxy_async();                        // the tile was written by generic stores; the bulk store reads it
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < th; ++r) bulk_s2g(p.y + oband + (size_t)r * Wo, sOut + r * opitch, (uint32_t)Wo * 4u);
        bulk_commit();
        bulk_wait_read_all();                   // shared memory must stay valid until the engine has read it
    }
}

// Persistent form of the kernel above (taken whenever one thread per 4-column group and row segment suffices, i.e.
// input widths up to 1024): a CTA walks the bands blockIdx.x, blockIdx.x + gridDim.x, ... and its threads' input rings
// never drain -- the first steps of the NEXT band are already in flight while the current band is decimated
// horizontally and stored.  A band is only 100 - 180 KB of input, about 4 us at an SM's share of the HBM bandwidth, so a
// one-band CTA spent a third of its life ramping its loads up and down (the non-persistent kernel: 0.50 - 0.55 of the
// copy peak with the same inner loops).
constexpr int kDownStreamDepth = 3;      // ring steps of the persistent kernel: two steps (8 rows at x2 / x4, 128 B per thread) in flight

template <int R, int WT, int TH>
__global__ void __launch_bounds__(kDownThreads, 2) down_stream_kernel(const __grid_constant__ DownRowsParams p, long long total_items)
{
    using G = DownGeom<R>;
    constexpr int T = G::T, OFF = G::OFF, NT = kDownThreads;
    // one step = RS = R * JO input rows (four at x2 and x4): JO output rows start per step, JO + 3 are live, each in an
    // accumulator slot (output index mod NSLOT); the step loop is unrolled by NSLOT / JO so that slots are static
    constexpr int JO = R == 2 ? 2 : 1, RS = R * JO, NSLOT = JO == 2 ? 6 : 4, UNR = NSLOT / JO, D = kDownStreamDepth;
    constexpr int JLO = -3;                                    // lowest live output of a step, relative to JO * s
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;

    const int H = p.H, W = WT ? WT : p.W, Ho = p.Ho, Wo = W / R;
    const int pitch = down_v_pitch(W, G::LPAD);
    const int opitch = blur_pad_pitch_c(Wo);
    float* sOut = reinterpret_cast<float*>(smem_raw);                  // [TH][opitch]  noise in, result out (in place)
    float* sV = sOut + TH * opitch;                                    // [TH][pitch]
    float4* sRing = reinterpret_cast<float4*>(sV + TH * pitch);        // [RING * R][NT] float4, one column per thread

    const int CWin = W >> 2;
    int nseg = 1;
    while (2 * nseg <= TH && 2 * nseg * CWin <= NT) nseg *= 2;         // row segments: a power of two that divides TH
    const int THs = TH / nseg;                                         // output rows per segment and band
    const int NS = (R * (THs - 1) + T - 1) / RS + 1;                   // steps until the segment's last output row is complete
    const bool has_unit = (int)threadIdx.x < nseg * CWin;
    const int seg = threadIdx.x / CWin, c4 = threadIdx.x - seg * CWin;
    const int o_first = seg * THs;
    const int n_my = (int)((total_items - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- issue side of the ring: (item, step) advance independently of the consuming loop
    float4* ring = sRing + threadIdx.x;
    int is_item = 0, is_step = 0, is_slot = 0, is_row0 = 0;
    const float* is_col = nullptr;
    auto is_locate = [&]() {
        if (is_item < n_my) {
            const long long w = blockIdx.x + (long long)is_item * gridDim.x;
            const int band = (int)(w % p.nbands);
            is_col = p.x + (size_t)(w / p.nbands) * H * W + 4 * c4;
            is_row0 = R * (band * TH + o_first) - OFF;                  // input row of step 0, row 0
        }
    };
    auto issue = [&]() {
        if (has_unit && is_item < n_my) {
#pragma unroll
            for (int k = 0; k < RS; ++k) {
                const int row = is_row0 + RS * is_step + k;
                const bool ok = (unsigned)row < (unsigned)H;
                cp_async16(ring + (is_slot * RS + k) * NT, is_col + (size_t)(ok ? row : 0) * W, ok);
            }
        }
        cp_async_commit();                                              // one group per step, empty past the end
        is_slot = is_slot + 1 == D ? 0 : is_slot + 1;
        if (++is_step == NS) {
            is_step = 0;
            ++is_item;
            is_locate();
        }
    };
    is_locate();
#pragma unroll
    for (int q = 0; q < D - 1; ++q) issue();

    int slot = 0;
    for (int it = 0; it < n_my; ++it) {
        const long long w = blockIdx.x + (long long)it * gridDim.x;
        const int band = (int)(w % p.nbands);
        const long long plane = w / p.nbands;
        const int i0 = band * TH, th = min(TH, Ho - i0);
        const size_t oband = ((size_t)plane * Ho + i0) * Wo;
        if (threadIdx.x == 0 && p.noise) {                              // this band's noise rows -> the output tile
            bulk_wait_read_all();                                       // the previous band's store has read the tile
            mbar_arrive_expect_tx(&bar, (uint32_t)th * Wo * 4u);
            for (int r = 0; r < th; ++r) bulk_g2s(sOut + r * opitch, p.noise + oband + (size_t)r * Wo, (uint32_t)Wo * 4u, &bar);
        }
        // ---- vertical pass of this band's segment: NS steps
        if (has_unit) {
            float4 acc[NSLOT];
#pragma unroll
            for (int q = 0; q < NSLOT; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            float* vcol = sV + o_first * pitch + G::LPAD + 4 * c4;
            for (int s0 = 0; s0 < NS; s0 += UNR) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int s = s0 + u;
                    if (s < NS) {
                        issue();
                        cp_async_wait<D - 1>();                         // the copies of this step have landed
                        float4 cur[RS];
#pragma unroll
                        for (int k = 0; k < RS; ++k) cur[k] = ring[(slot * RS + k) * NT];
                        slot = slot + 1 == D ? 0 : slot + 1;
#pragma unroll
                        for (int j = 0; j < JO; ++j) acc[(JO * u + j) % NSLOT] = make_float4(0.f, 0.f, 0.f, 0.f);   // rows that start here
#pragma unroll
                        for (int j = JLO; j < JO; ++j) {                // output row o = JO * s + j
                            float4& a = acc[(JO * u + j + NSLOT) % NSLOT];
#pragma unroll
                            for (int k = 0; k < RS; ++k) {
                                const int t = k - R * j;                // tap of input row k of this step in o's window
                                if (t >= 0 && t < T) {
                                    const float wgt = p.wint[t];
                                    a.x = fmaf(wgt, cur[k].x, a.x); a.y = fmaf(wgt, cur[k].y, a.y);
                                    a.z = fmaf(wgt, cur[k].z, a.z); a.w = fmaf(wgt, cur[k].w, a.w);
                                }
                            }
                            // o's last input row R * o + T - 1 lies in this step  <=>  0 <= R * j + T - 1 < RS
                            if (R * j + T - 1 >= 0 && R * j + T - 1 < RS) {
                                const int o = JO * s + j;
                                if (o >= 0 && o < THs) *reinterpret_cast<float4*>(vcol + o * pitch) = a;
                            }
                        }
                    }
                }
            }
            // the image's first / last two output rows have truncated, renormalised windows: recompute this column group
            const float* __restrict__ colp = p.x + (size_t)plane * H * W + 4 * c4;
            for (int o = 0; o < THs; ++o) {
                const int i = i0 + o_first + o;
                const int kb = i < Ho ? border_slot(i, Ho) : -1;
                if (kb < 0) continue;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t = 0; t < p.rowTab.xsize[kb]; ++t) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(colp + (size_t)(p.rowTab.xmin[kb] + t) * W));
                    const float wgt = p.rowTab.w[kb][t];
                    a.x = fmaf(wgt, v.x, a.x); a.y = fmaf(wgt, v.y, a.y); a.z = fmaf(wgt, v.z, a.z); a.w = fmaf(wgt, v.w, a.w);
                }
                *reinterpret_cast<float4*>(vcol + o * pitch) = a;
            }
        } else {
            for (int s = 0; s < NS; ++s) issue();                       // keep the group count in step (empty groups)
        }
        if (threadIdx.x == 0 && !p.noise) bulk_wait_read_all();         // previous band's store has read the tile
        __syncthreads();
        if (p.noise) mbar_wait(&bar, it & 1);

        // ---- horizontal decimation: sOut[r][j] = sum_t w[t] * sV[r][R*j - OFF + t]  (+ sigma * noise)
        constexpr int NO = G::NO, NV = G::NV, SKIP = G::SKIP;
        const int nblk = Wo / NO;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int rgroups = (th + 7) >> 3, bgroups = (nblk + 3) >> 2;
        for (int wi = warp; wi < rgroups * bgroups; wi += NT / 32) {
            const int rg = wi % rgroups, bg = wi / rgroups;
            const int r = 8 * rg + (lane & 7), blk = 4 * bg + (lane >> 3);
            if (r >= th || blk >= nblk) continue;
            const float* __restrict__ src = sV + r * pitch + R * NO * blk;
            float v[4 * NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * q);
                v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
            }
            float out[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < T; ++t) a = fmaf(p.wint[t], v[SKIP + R * o + t], a);
                out[o] = a;
            }
            if (blk == 0 || blk == nblk - 1) {                         // the row's first / last two outputs
                const float* row = sV + r * pitch + G::LPAD;
                auto border = [&](int kb) {
                    float a = 0.f;
                    for (int t = 0; t < p.colTab.xsize[kb]; ++t) a = fmaf(p.colTab.w[kb][t], row[p.colTab.xmin[kb] + t], a);
                    return a;
                };
                if (blk == 0) { out[0] = border(0); out[1] = border(1); }
                if (blk == nblk - 1) { out[NO - 2] = border(2); out[NO - 1] = border(3); }
            }
            float* dst = sOut + r * opitch + NO * blk;
#pragma unroll
            for (int q = 0; q < NO / 4; ++q) {
                float4 o4 = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
                if (p.noise) {
                    const float4 n4 = *reinterpret_cast<const float4*>(dst + 4 * q);
                    o4.x = fmaf(p.sigma, n4.x, o4.x); o4.y = fmaf(p.sigma, n4.y, o4.y);
                    o4.z = fmaf(p.sigma, n4.z, o4.z); o4.w = fmaf(p.sigma, n4.w, o4.w);
                }
                *reinterpret_cast<float4*>(dst + 4 * q) = o4;
            }
        }
        fence_proxy_async();                    // the tile was written by generic stores; the bulk store reads it
        __syncthreads();                        // ... and sV is free for the next band's vertical pass
        if (threadIdx.x == 0) {
            for (int r = 0; r < th; ++r) bulk_s2g(p.y + oband + (size_t)r * Wo, sOut + r * opitch, (uint32_t)Wo * 4u);
            bulk_commit();
        }
    }
    cp_async_wait<0>();
    if (threadIdx.x == 0) bulk_wait_read_all(); // shared memory must stay valid until the engine has read it
}

// Transpose, "rows" form: gx = A^T gy.  The vertical 4-tap polyphase pass walks down the band with the four gradient
// rows it needs in rotating registers (each gradient row is loaded from global memory once per band segment, 128-bit
// read-only loads, nothing staged); the horizontal pass produces 16 (12 at R = 3) consecutive outputs per item from
// three or four 128-bit shared loads (the round-1 kernel issued four scalar shared loads per OUTPUT and sat at 0.28 -
// 0.40 of the copy peak).  Rows / columns near the image border go through contributor tables.
template <int R> struct DownTGeom {
    static constexpr int OFF = aa_off(R);
    static constexpr int NO = R == 3 ? 12 : 16;                // outputs per horizontal work item
    static constexpr int NVJ = NO / R;                          // intermediate columns advanced per item
    static constexpr int NV = R == 2 ? 4 : 3;                   // 128-bit loads per item
    static constexpr int LP = 4;                                // left margin of an intermediate row
    static constexpr int SEGROWS = 4 * R;                       // rows after which the register window is back in place
};

struct DownTRowsParams {
    const float* gy;
    float* gx;
    int H, W, Ho, Wo, TM, nbands;
    float wint[kAaMaxTaps];
    Contrib colL[16], colR[16], rowLo[kBorderRows], rowHi[kBorderRows];
};

template <int R, int WT>
__global__ void __launch_bounds__(kDownThreads) down_t_rows_kernel(const __grid_constant__ DownTRowsParams p)
{
    using G = DownTGeom<R>;
    constexpr int OFF = G::OFF, NT = kDownThreads, NO = G::NO, NV = G::NV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int H = p.H, W = WT ? WT : p.W, Ho = p.Ho, Wo = W / R;
    const int pitch = blur_pad_pitch_c(G::LP + Wo + 8);
    float* sV = reinterpret_cast<float*>(smem_raw);                    // [TM][pitch]
    const int band = blockIdx.x % p.nbands;
    const long long plane = blockIdx.x / p.nbands;
    const int m0 = band * p.TM, tm = min(p.TM, H - m0);
    const float* __restrict__ gplane = p.gy + (size_t)plane * Ho * Wo;
    const int ML = min(H, 5 * R - OFF), MR0 = max(ML, R * (Ho - 2) - OFF);     // rows outside [ML, MR0) are border rows

    // ---- vertical pass: sV[m][j] = sum_q wint[ph + R q] * gy[ib - q][j],  ph = (m + OFF) % R, ib = (m + OFF) / R
    const int CWo = Wo >> 2;
    const int nsegmax = p.TM / G::SEGROWS;
    const int nseg = CWo >= NT ? 1 : min(nsegmax, NT / CWo);
    const int TMs = ((nsegmax + nseg - 1) / nseg) * G::SEGROWS;        // rows per segment, a multiple of 4R
    for (int unit = threadIdx.x; unit < nseg * CWo; unit += NT) {
        const int seg = unit / CWo, j4 = unit - seg * CWo;
        const int r_first = seg * TMs;
        const int nrows = min(TMs, tm - r_first);
        if (nrows <= 0) continue;
        const float* __restrict__ colp = gplane + 4 * j4;
        auto load_row = [&](int i) {
            return (i >= 0 && i < Ho) ? __ldg(reinterpret_cast<const float4*>(colp + (size_t)i * Wo)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // (m0 + r_first) is a multiple of 4R: the first row has phase OFF % R and window top ib0, ib0 % 4 == (OFF / R) % 4
        const int ib0 = (m0 + r_first + OFF) / R;
        constexpr int S0 = (OFF / R) & 3;                               // register slot of gradient row ib0
        float4 g[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) g[(S0 - q) & 3] = load_row(ib0 - q);
        float4 nxt = load_row(ib0 + 1);
        float* vcol = sV + r_first * pitch + G::LP + 4 * j4;
        for (int rb = 0; rb < nrows; rb += G::SEGROWS) {
#pragma unroll
            for (int u = 0; u < G::SEGROWS; ++u) {
                const int e = OFF % R + u;                              // (m + OFF) - R * ib0 - R * (rb / R)  for this row
                const int ph = e % R, adv = e / R;                      // window top = ib0 + rb / R + adv
                if (u > 0 && ph == 0) {                                 // the window advanced by one gradient row
                    g[(S0 + adv) & 3] = nxt;
                    nxt = load_row(ib0 + rb / R + adv + 1);
                }
                if (rb + u < nrows) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float w = p.wint[ph + R * q];
                        const float4 v = g[(S0 + adv - q) & 3];
                        a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
                    }
                    *reinterpret_cast<float4*>(vcol + (rb + u) * pitch) = a;
                }
            }
            // after 4R rows the window has advanced by 4 gradient rows: slot S0 holds row ib0 + rb / R + 4 again
            {
                constexpr int e_end = OFF % R + G::SEGROWS;
                if (e_end % R == 0) {                                   // the advance that falls on the first row of the next block
                    g[(S0 + e_end / R) & 3] = nxt;
                    nxt = load_row(ib0 + rb / R + e_end / R + 1);
                }
            }
        }
        // border rows of the image: contributor tables, straight from global memory
        for (int r = 0; r < nrows; ++r) {
            const int m = m0 + r_first + r;
            if (m >= ML && m < MR0) continue;
            const Contrib& c = m < ML ? p.rowLo[m] : p.rowHi[m - MR0];
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < kNQ; ++q) {
                const float w = c.w[q];
                if (w != 0.f) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(colp + (size_t)c.idx[q] * Wo));
                    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
                }
            }
            *reinterpret_cast<float4*>(vcol + r * pitch) = a;
        }
    }
    // contributor tables of the border columns -> shared memory (per-lane indexing of kernel parameters serialises in
    // the constant cache: ncu showed 26 % of the stall samples on those loads)
    Contrib* sCol = reinterpret_cast<Contrib*>(sV + p.TM * pitch);         // [2 * NO]
    for (int e = threadIdx.x; e < 2 * NO; e += NT) sCol[e] = e < NO ? p.colL[e] : p.colR[e - NO];
    __syncthreads();

    // ---- horizontal pass: gx[m][n] = sum_q wint[(n + OFF) % R + R q] * sV[m][(n + OFF) / R - q]
    float* __restrict__ oplane = p.gx + ((size_t)plane * H + m0) * W;
    const int ngrp = W / NO;
    if constexpr (R == 3) {
        // 12 outputs per item (the phase pattern of 4 outputs is not static at R = 3)
        const int ngi = ngrp - 2;                    // interior groups: 1 .. ngrp - 2
        for (int item = threadIdx.x; item < tm * ngi; item += NT) {
            const int u = 1 + item % ngi, r = item / ngi;
            const float* row = sV + r * pitch + G::LP;
            float v[4 * NV];
            const float* src = row + G::NVJ * u - 4;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * q);
                v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
            }
            float out[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                const int ph = (o + OFF) % R, jl = (o + OFF) / R + 4;
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) a = fmaf(p.wint[ph + R * q], v[jl - q], a);
                out[o] = a;
            }
            float* dst = oplane + (size_t)r * W + NO * u;
#pragma unroll
            for (int q = 0; q < NO / 4; ++q) st_stream4(dst + 4 * q, make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]));
        }
    } else {
        // R = 2, 4: one 128-bit store per item, neighbouring lanes write neighbouring vectors (512 contiguous bytes per
        // warp instruction; 16-output items left every store instruction touching half sectors: ncu, 45 % excess
        // sectors).  The 4 outputs of vector g read the intermediate columns (4 / R) g - 2 .. + NVAL - 1 of their row.
        constexpr int NVAL = R == 2 ? 6 : 5;
        constexpr int GB = NO / 4;                   // border vectors per side (handled through the tables below)
        const int ngv = W / 4 - 2 * GB;
        for (int item = threadIdx.x; item < tm * ngv; item += NT) {
            const int g = GB + item % ngv, r = item / ngv;
            const float* src = sV + r * pitch + G::LP + (4 / R) * g - 2;
            float v[NVAL];
            if constexpr (R == 2) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float2 t2 = *reinterpret_cast<const float2*>(src + 2 * q);
                    v[2 * q] = t2.x; v[2 * q + 1] = t2.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < NVAL; ++q) v[q] = src[q];
            }
            float out[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int ph = (o + OFF) % R, jl = (o + OFF) / R - OFF / R + 3;
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) a = fmaf(p.wint[ph + R * q], v[jl - q], a);
                out[o] = a;
            }
            st_stream4(oplane + (size_t)r * W + 4 * g, make_float4(out[0], out[1], out[2], out[3]));
        }
    }
    // the first and the last group of every row contain the border columns: one thread per output, contributor tables
    for (int item = threadIdx.x; item < tm * 2 * NO; item += NT) {
        const int r = item / (2 * NO), e = item - r * 2 * NO;
        const Contrib c = sCol[e];
        const int n = e < NO ? e : W - 2 * NO + e;
        const float* row = sV + r * pitch + G::LP;
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kNQ; ++q)
            if (c.w[q] != 0.f) a = fmaf(c.w[q], row[c.idx[q]], a);
        oplane[(size_t)r * W + n] = a;
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


static size_t down_rows_smem(int th, int W, int Wo, int rate)
{
    // ring: the persistent kernel (input widths up to 1024) keeps kDownStreamDepth steps of R * JO rows (JO = 2 at x2),
    // the one-band kernel RING steps of R rows
    const int lpad = 4 * ((aa_off(rate) + 3) / 4);
    const int ring_rows = W / 4 <= kDownThreads ? kDownStreamDepth * (rate == 2 ? 4 : rate) : (rate <= 2 ? 6 * rate : 4 * rate);
    return ((size_t)th * blur_pad_pitch_c(Wo) + (size_t)th * down_v_pitch(W, lpad)) * 4 + (size_t)ring_rows * kDownThreads * 16;
}

template <int R, int WT, int TH>
static int launch_down_rows_inst(const DownRowsParams& q, long long planes, cudaStream_t st)
{
    const int W = q.W, Wo = q.Wo;
    const size_t smem = down_rows_smem(TH, W, Wo, R);
    if (W / 4 <= kDownThreads) {
        DeviceProps dp;
        int rc = get_device_props(&dp);
        if (rc) return rc;
        const long long items = planes * q.nbands;
        const int per_sm = std::max(1, std::min(2, (int)((size_t)227 * 1024 / (smem + 1024))));
        const unsigned grid = (unsigned)std::min<long long>(items, (long long)dp.sm_count * per_sm);
        SEI_CUDA(allow_smem(down_stream_kernel<R, WT, TH>, smem));
        down_stream_kernel<R, WT, TH><<<grid, kDownThreads, smem, st>>>(q, items);
        return finish_launch(q.noise ? "down_rows_kernel<noise>" : "down_rows_kernel");
    }
    SEI_CUDA(allow_smem(down_rows_kernel<R, WT, TH>, smem));
    down_rows_kernel<R, WT, TH><<<(unsigned)(planes * q.nbands), kDownThreads, smem, st>>>(q);
    return finish_launch(q.noise ? "down_rows_kernel<noise>" : "down_rows_kernel");
}

template <int R>
static int launch_down_rows(DownRowsParams& q, long long planes, int th, cudaStream_t st)
{
    q.nbands = (q.Ho + th - 1) / th;
    if (th == 16) {
        if (q.W == 256 * R) return launch_down_rows_inst<R, 256 * R, 16>(q, planes, st);
        return launch_down_rows_inst<R, 0, 16>(q, planes, st);
    }
    if (q.W == 256 * R) return launch_down_rows_inst<R, 256 * R, 8>(q, planes, st);
    return launch_down_rows_inst<R, 0, 8>(q, planes, st);
}

template <int R>
static int launch_down_t_rows(DownTRowsParams& q, long long planes, size_t smem, cudaStream_t st)
{
    using G = DownTGeom<R>;
    for (int o = 0; o < G::NO; ++o) {
        aa_contributors(o, q.W, q.Wo, R, q.colL[o]);
        aa_contributors(q.W - G::NO + o, q.W, q.Wo, R, q.colR[o]);
    }
    const int ML = std::min(q.H, 5 * R - G::OFF), MR0 = std::max(ML, R * (q.Ho - 2) - G::OFF);
    for (int m = 0; m < ML; ++m) aa_contributors(m, q.H, q.Ho, R, q.rowLo[m]);
    for (int m = MR0; m < q.H; ++m) aa_contributors(m, q.H, q.Ho, R, q.rowHi[m - MR0]);
    const unsigned grid = (unsigned)(planes * q.nbands);
    if (q.W == 256 * R) {
        SEI_CUDA(allow_smem(down_t_rows_kernel<R, 256 * R>, smem));
        down_t_rows_kernel<R, 256 * R><<<grid, kDownThreads, smem, st>>>(q);
    } else {
        SEI_CUDA(allow_smem(down_t_rows_kernel<R, 0>, smem));
        down_t_rows_kernel<R, 0><<<grid, kDownThreads, smem, st>>>(q);
    }
    return finish_launch("down_t_rows_kernel");
}

static int env_int_down(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
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

    // ---- round-2 "rows" kernels (vertical pass first, straight from global memory); SEI_DOWN_V1=1 selects round 1's
    const bool v1 = env_int_down("SEI_DOWN_V1", 0) == 1;
    if (tiled_ok && !v1 && path != SEI_PATH_DIRECT && !transpose && Wo % 8 == 0 && W <= 4096) {
        DownRowsParams q;
        q.x = in; q.y = out; q.noise = noise; q.sigma = sigma; q.H = H; q.W = W; q.Ho = Ho; q.Wo = Wo;
        for (int t = 0; t < kAaMaxTaps; ++t) q.wint[t] = p.wint[t];
        for (int k = 0; k < 4; ++k) {
            aa_axis_weights(k < 2 ? k : Wo - 4 + k, W, rate, q.colTab.w[k], q.colTab.xmin[k], q.colTab.xsize[k]);
            aa_axis_weights(k < 2 ? k : Ho - 4 + k, H, rate, q.rowTab.w[k], q.rowTab.xmin[k], q.rowTab.xsize[k]);
        }
        // band height: 16 output rows at x2; 8 at x3 / x4, where the intermediate band is 4 KB per row and four CTAs
        // per SM (instead of two) hide the global-load latency of the vertical pass better (x4: 33.6 vs 42.2 us)
        int th = env_int_down("SEI_DOWN_ROWS_TH", rate >= 3 ? 8 : 16);
        th = th == 8 ? 8 : 16;
        if (down_rows_smem(th, W, Wo, rate) > (size_t)dp.smem_optin) th = 8;
        if (down_rows_smem(th, W, Wo, rate) <= (size_t)dp.smem_optin &&
            planes * ((Ho + th - 1) / th) < (1ll << 31)) {
            switch (rate) {
            case 2: return launch_down_rows<2>(q, planes, th, st);
            case 3: return launch_down_rows<3>(q, planes, th, st);
            default: return launch_down_rows<4>(q, planes, th, st);
            }
        }
    }
    if (tiled_ok && !v1 && path != SEI_PATH_DIRECT && transpose && W % (rate == 3 ? 12 : 16) == 0 && Wo >= 16 &&
        H >= 8 * rate && Ho >= 8) {
        DownTRowsParams q;
        q.gy = in; q.gx = out; q.H = H; q.W = W; q.Ho = Ho; q.Wo = Wo;
        for (int t = 0; t < kAaMaxTaps; ++t) q.wint[t] = p.wint[t];
        const int segrows = 4 * rate;
        int tm = env_int_down("SEI_DOWNT_ROWS_TM", 0);
        if (tm <= 0 || tm % segrows) {
            // every thread should own a column group: 256 threads / (Wo / 4) groups = segments of 4R rows each
            const int nseg = std::max(1, kDownThreads / std::max(1, Wo / 4));
            tm = segrows * nseg;
            while (tm < 32) tm *= 2;
        }
        const size_t smem_t = (size_t)tm * blur_pad_pitch_c(4 + Wo + 8) * 4 + 32 * sizeof(Contrib);
        const int ML = std::min(H, 5 * rate - aa_off(rate)), MR0 = std::max(ML, rate * (Ho - 2) - aa_off(rate));
        if (smem_t <= (size_t)dp.smem_optin && ML <= kBorderRows && H - MR0 <= kBorderRows &&
            planes * ((H + tm - 1) / tm) < (1ll << 31)) {
            q.TM = tm;
            q.nbands = (H + tm - 1) / tm;
            switch (rate) {
            case 2: return launch_down_t_rows<2>(q, planes, smem_t, st);
            case 3: return launch_down_t_rows<3>(q, planes, smem_t, st);
            default: return launch_down_t_rows<4>(q, planes, smem_t, st);
            }
        }
    }
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
