// tile_ops.cuh -- shared-memory tile passes reused by the stand-alone and the fused kernels.
#pragma once
#include "sei_common.cuh"

namespace sei {

constexpr int kMaxK = 31;

// ---------------------------------------------------------------- separable circular blur passes
// Template parameter WT: compile-time image width (0 = use the run-time width).  With WT > 0 every
// shared-memory address below is "base + immediate" and the work-item decode is shifts and masks; the
// ncu profile of the run-time-width version showed 58 % of its issued instructions were such index math.
// PAD: the intermediate band keeps 4*LCH halo columns on both sides of every row (circular copies, written
// by the vertical pass), so the horizontal pass needs no wrap-around logic.
template <int K> struct BlurGeom {
    static constexpr int P = K / 2;
    static constexpr int LCH = (P + 3) / 4;          // float4 chunks of halo on each side
    static constexpr int NCH = 2 * LCH + 1;
    static constexpr int LEFT = 4 * LCH;
};

// Row pitch of the halo-padded intermediate: W + 2*LEFT rounded up to 4 (mod 32) floats, so consecutive rows start
// one 16-byte bank group apart.  blur_hpass16 relies on it: a quarter-warp reads 4 rows x 2 column blocks with
// 128-bit loads and touches 8 distinct bank groups (no conflicts).
__host__ __device__ constexpr int blur_pad_pitch(int n) { return n + ((4 - n % 32) + 32) % 32; }

template <int K, bool PAD> __host__ __device__ constexpr int blur_mid_pitch(int W)
{
    return PAD ? blur_pad_pitch(W + 2 * BlurGeom<K>::LEFT) : W;
}

// host-side: pitch of the padded intermediate for a run-time tap count
inline int blur_mid_pitch_host(int K, int W) { return blur_pad_pitch(W + 8 * ((K / 2 + 3) / 4)); }

// vertical: sMid[o][c] = sum_t cv[t] * sIn[o + t][c],  o in [0, th), rows of sIn in [0, th + K - 1)
// Register-blocked: each work item owns kRH output rows x 4 columns and streams kRH + K - 1 input
// rows through a scatter-form accumulation, so every shared load feeds up to kRH * 4 FMAs and the
// taps are compile-time-indexed constant-bank operands.
template <int K, int kRH, bool PAD, bool FULL>
__device__ __forceinline__ void blur_vpass_item(const float* __restrict__ base, float* __restrict__ dst, int W, int pitch,
                                                int nvalid, int nout, int c4, int CW, const float* __restrict__ cv)
{
    using G = BlurGeom<K>;
    float4 acc[kRH];
#pragma unroll
    for (int o = 0; o < kRH; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kRH + K - 1; ++i) {
        if (FULL || i < nvalid) {
            const float4 v = *reinterpret_cast<const float4*>(base + i * W);
#pragma unroll
            for (int o = 0; o < kRH; ++o) {
                const int t = i - o;
                if (t >= 0 && t < K) {
                    const float c = cv[t];
                    acc[o].x = fmaf(c, v.x, acc[o].x); acc[o].y = fmaf(c, v.y, acc[o].y);
                    acc[o].z = fmaf(c, v.z, acc[o].z); acc[o].w = fmaf(c, v.w, acc[o].w);
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kRH; ++o)
        if (FULL || o < nout) *reinterpret_cast<float4*>(dst + o * pitch) = acc[o];
    if (PAD) {
        // circular halo copies: the first LCH chunks also go right of the row, the last LCH chunks left of it
        if (c4 < G::LCH) {
#pragma unroll
            for (int o = 0; o < kRH; ++o)
                if (FULL || o < nout) *reinterpret_cast<float4*>(dst + o * pitch + W) = acc[o];
        }
        if (c4 >= CW - G::LCH) {
#pragma unroll
            for (int o = 0; o < kRH; ++o)
                if (FULL || o < nout) *reinterpret_cast<float4*>(dst + o * pitch - W) = acc[o];
        }
    }
}

template <int K, int NT, int kRH, int WT, bool PAD>
__device__ __forceinline__ void blur_vpass_rh(const float* __restrict__ sIn, float* __restrict__ sMid, int Wrt, int th,
                                              const float* __restrict__ cv)
{
    using G = BlurGeom<K>;
    const int W = WT ? WT : Wrt;
    const int CW = W >> 2, rin = th + K - 1;
    const int pitch = blur_mid_pitch<K, PAD>(W);
    const int HL = PAD ? G::LEFT : 0;
    const int ngroups = (th + kRH - 1) / kRH;
    for (int item = threadIdx.x; item < ngroups * CW; item += NT) {
        const int g = item / CW, c4 = item - g * CW;
        const int o0 = g * kRH;
        const float* base = sIn + o0 * W + c4 * 4;
        float* dst = sMid + o0 * pitch + HL + c4 * 4;
        if (o0 + kRH <= th)
            blur_vpass_item<K, kRH, PAD, true>(base, dst, W, pitch, 0, 0, c4, CW, cv);
        else
            blur_vpass_item<K, kRH, PAD, false>(base, dst, W, pitch, rin - o0, th - o0, c4, CW, cv);
    }
}

// kRH = 8 output rows per work item (2.5 shared loads per output vector at K = 13) when that still gives
// every thread an item, otherwise 4 rows per item (more items, 4 loads per output vector)
template <int K, int NT, int WT = 0, bool PAD = false>
__device__ __forceinline__ void blur_vpass(const float* __restrict__ sIn, float* __restrict__ sMid, int Wrt, int th,
                                           const float* __restrict__ cv)
{
    const int W = WT ? WT : Wrt;
    if (((th + 7) >> 3) * (W >> 2) >= NT)
        blur_vpass_rh<K, NT, 8, WT, PAD>(sIn, sMid, Wrt, th, cv);
    else
        blur_vpass_rh<K, NT, 4, WT, PAD>(sIn, sMid, Wrt, th, cv);
}

// horizontal (circular in x): y[r][n] = sum_t ch[t] * sMid[r][(n + t - P) mod W] (+ sigma * noise)
// yrow0 / nrow0 point at the first output row of the band in global memory (row pitch W).
// one horizontal work item: 4 outputs of row r starting at column 4*c4
template <int K, bool NOISE, bool PAD>
__device__ __forceinline__ void blur_hpass_item(const float* __restrict__ row, int c4, int CW, const float (&tap)[K],
                                                float* __restrict__ ydst, const float* __restrict__ ndst, float sigma)
{
    using G = BlurGeom<K>;
    constexpr int P = G::P, LCH = G::LCH, NCH = G::NCH, LEFT = G::LEFT;
    // issue the noise load first: its DRAM latency hides behind the shared loads and the K*4 FMAs below
    // (ncu: 22 % of the fused kernel's stall samples sat on the first use of this value)
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NOISE) nz = ld_stream4(ndst);
    float v[4 * NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
        float4 t;
        if (PAD) {
            t = *reinterpret_cast<const float4*>(row + (c4 + q) * 4);      // padded column 4*(c4 - LCH + q) + LEFT
        } else {
            int cc = c4 - LCH + q;
            if (cc < 0) cc += CW;
            if (cc >= CW) cc -= CW;
            t = *reinterpret_cast<const float4*>(row + cc * 4);
        }
        v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < K; ++t) {
#pragma unroll
        for (int o = 0; o < 4; ++o) out[o] = fmaf(tap[t], v[LEFT + o + t - P], out[o]);
    }
    if (NOISE) {
        out[0] = fmaf(sigma, nz.x, out[0]); out[1] = fmaf(sigma, nz.y, out[1]);
        out[2] = fmaf(sigma, nz.z, out[2]); out[3] = fmaf(sigma, nz.w, out[3]);
    }
    st_stream4(ydst, make_float4(out[0], out[1], out[2], out[3]));
}

// horizontal (circular in x): y[r][n] = sum_t ch[t] * sMid[r][(n + t - P) mod W] (+ sigma * noise)
// yrow0 / nrow0 point at the first output row of the band in global memory (row pitch W).
// The taps are pulled into registers once per call (the constant-bank loads showed up as 17 % of the stall
// samples), and two independent work items are in flight per thread so one item's shared loads overlap the
// other's FMA chain.
template <int K, int NT, bool NOISE, int WT = 0, bool PAD = false>
__device__ __forceinline__ void blur_hpass(const float* __restrict__ sMid, int Wrt, int th, const float* __restrict__ ch,
                                           float* __restrict__ yrow0, const float* __restrict__ nrow0, float sigma)
{
    const int W = WT ? WT : Wrt;
    const int CW = W >> 2;
    const int pitch = blur_mid_pitch<K, PAD>(W);
    float tap[K];
#pragma unroll
    for (int t = 0; t < K; ++t) asm volatile("mov.f32 %0, %1;" : "=f"(tap[t]) : "f"(ch[t]));
    const int nitems = th * CW;
    int item = threadIdx.x;
    for (; item + NT < nitems; item += 2 * NT) {
        const int r0 = item / CW, c0 = item - r0 * CW;
        const int r1 = (item + NT) / CW, c1 = (item + NT) - r1 * CW;
        const int g0 = r0 * W + c0 * 4, g1 = r1 * W + c1 * 4;
        blur_hpass_item<K, NOISE, PAD>(sMid + r0 * pitch, c0, CW, tap, yrow0 + g0, NOISE ? nrow0 + g0 : nullptr, sigma);
        blur_hpass_item<K, NOISE, PAD>(sMid + r1 * pitch, c1, CW, tap, yrow0 + g1, NOISE ? nrow0 + g1 : nullptr, sigma);
    }
    if (item < nitems) {
        const int r0 = item / CW, c0 = item - r0 * CW;
        const int g0 = r0 * W + c0 * 4;
        blur_hpass_item<K, NOISE, PAD>(sMid + r0 * pitch, c0, CW, tap, yrow0 + g0, NOISE ? nrow0 + g0 : nullptr, sigma);
    }
}

// horizontal pass, 16 outputs per work item.  One item = (row r, block of 16 columns): 4 + 2*LCH 128-bit shared
// loads feed 16*K FMAs (2 loads per output vector at K = 13, against 4-5 for the 4-wide item above; the ncu
// profile of the 4-wide version showed the LSU shared-memory pipe at 71 % and the FMA pipe at 42 %).  Lanes are
// laid out 4 rows x 8 column blocks per warp (lane = 4 * block + row); with the intermediate's pitch = 4 (mod 32)
// every quarter-warp load is conflict-free.  Requires W % 16 == 0 and the padded intermediate.
// XCHG: the two lanes that own adjacent 16-column blocks of a row (lane ^ 4) swap half of their results before storing,
// so that every store (and noise load) instruction covers whole 32-byte sectors: lane E writes chunks 0 and 2 of both
// blocks, lane O chunks 1 and 3.  Without it a store instruction wrote 16 bytes at a 64-byte stride per lane and ncu
// reported 50 % excess L2 sectors for the global stores.  All 32 lanes must execute the item (warp shuffles).
template <int K, bool NOISE, bool XCHG>
__device__ __forceinline__ void blur_hpass16_item(const float* __restrict__ src, const float (&tap)[K],
                                                  float* __restrict__ ydst, const float* __restrict__ ndst, float sigma)
{
    using G = BlurGeom<K>;
    constexpr int P = G::P, LCH = G::LCH, LEFT = G::LEFT, NCH = 4 + 2 * LCH;
    const bool odd = XCHG && (threadIdx.x & 4) != 0;
    // element offset of this lane's i-th 16-byte access relative to its own block
    auto off = [&](int i) { return XCHG ? (odd ? 8 * i - 12 : 8 * i) : 4 * i; };
    float4 nz[4];
    if (NOISE) {          // own block (added before the exchange): the four 16-byte loads of a lane share L1 lines
#pragma unroll
        for (int i = 0; i < 4; ++i) nz[i] = ld_stream4(ndst + 4 * i);
    }
    float v[4 * NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(src + 4 * q);
        v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    float out[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) out[o] = 0.f;
#pragma unroll
    for (int t = 0; t < K; ++t) {
#pragma unroll
        for (int o = 0; o < 16; ++o) out[o] = fmaf(tap[t], v[LEFT + o + t - P], out[o]);
    }
    if (NOISE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[4 * i + 0] = fmaf(sigma, nz[i].x, out[4 * i + 0]); out[4 * i + 1] = fmaf(sigma, nz[i].y, out[4 * i + 1]);
            out[4 * i + 2] = fmaf(sigma, nz[i].z, out[4 * i + 2]); out[4 * i + 3] = fmaf(sigma, nz[i].w, out[4 * i + 3]);
        }
    }
    if (XCHG) {
        // E (even block) sends chunks 1, 3 and receives the partner's chunks 0, 2; O sends 0, 2 and receives 1, 3
        float rcv[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float s0 = odd ? out[e] : out[4 + e], s1 = odd ? out[8 + e] : out[12 + e];
            rcv[e] = __shfl_xor_sync(0xffffffffu, s0, 4);
            rcv[4 + e] = __shfl_xor_sync(0xffffffffu, s1, 4);
        }
        float fin[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // E: own c0, own c2, partner c0, partner c2;  O: partner c1, partner c3, own c1, own c3
            fin[e] = odd ? rcv[e] : out[e];
            fin[4 + e] = odd ? rcv[4 + e] : out[8 + e];
            fin[8 + e] = odd ? out[4 + e] : rcv[e];
            fin[12 + e] = odd ? out[12 + e] : rcv[4 + e];
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) out[o] = fin[o];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st_stream4(ydst + off(i), make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]));
}

template <int K, int NT, bool NOISE, int WT, bool TWO>
__device__ __forceinline__ void blur_hpass16(const float* __restrict__ sMid, int Wrt, int th, const float* __restrict__ ch,
                                             float* __restrict__ yrow0, const float* __restrict__ nrow0, float sigma)
{
    const int W = WT ? WT : Wrt;
    const int NB = W >> 4;
    const int pitch = blur_mid_pitch<K, true>(W);
    float tap[K];
#pragma unroll
    for (int t = 0; t < K; ++t) asm volatile("mov.f32 %0, %1;" : "=f"(tap[t]) : "f"(ch[t]));
    const int nitems = ((th + 3) >> 2) * 4 * NB;
    const bool whole = (th & 3) == 0;      // every item maps to a row of the band: no per-item row check
    // lane exchange only when every warp is full in every iteration (items per row quad = 4 * NB a multiple of 32)
    // and only without the noise epilogue: with it the exchange was measured slower (A+noise 61.7 us plain, 64.5 us with
    // paired noise loads, 68 us with the noise added before the exchange; A alone 49.3 -> 47.0 us with the exchange)
    const bool xchg = !NOISE && whole && (NB & 7) == 0;
    auto run = [&](int item, int mode) {   // mode 0: exchange, 1: plain, 2: plain with a row check
        const int j = item & 3, q = item >> 2;
        const int rg = q / NB, k = q - rg * NB;
        const int r = 4 * rg + j;
        const int g = r * W + 16 * k;
        if (mode == 0)
            blur_hpass16_item<K, NOISE, true>(sMid + r * pitch + 16 * k, tap, yrow0 + g, NOISE ? nrow0 + g : nullptr, sigma);
        else if (mode == 1 || r < th)
            blur_hpass16_item<K, NOISE, false>(sMid + r * pitch + 16 * k, tap, yrow0 + g, NOISE ? nrow0 + g : nullptr, sigma);
    };
    int item = threadIdx.x;
    if (xchg) {
        if (TWO) {
            for (; item + NT < nitems; item += 2 * NT) {
                run(item, 0);
                run(item + NT, 0);
            }
        }
        for (; item < nitems; item += NT) run(item, 0);
    } else if (whole) {
        for (; item < nitems; item += NT) run(item, 1);
    } else {
        for (; item < nitems; item += NT) run(item, 2);
    }
}

// ---------------------------------------------------------------- scale-transform taps
// pixel coordinate of output index idx along one axis, rounding exactly like the reference:
//   u = 2/S * idx - 1 ; g = 1/rate * (u - c) + c          (src/transforms.py:31-41, fp32 tensors)
//   pix = ((g + 1) / 2) * (S - 1)                          (grid_sample unnormalize, align_corners=True)
__device__ __forceinline__ float scale_src_coord(int idx, float two_over_S, float inv_rate, float c, float Sm1)
{
    const float u = __fsub_rn(__fmul_rn(two_over_S, (float)idx), 1.0f);
    const float g = __fadd_rn(__fmul_rn(inv_rate, __fsub_rn(u, c)), c);
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), Sm1);
}

struct AxisTap {
    float w[4];
    int idx[4];
};

// S: size of the SOURCE axis (the output axis enters through two_over_S only)
__device__ __forceinline__ void scale_axis_tap(int i, int S, float two_over_S, float inv_rate, float c, AxisTap& t)
{
    const float pix = scale_src_coord(i, two_over_S, inv_rate, c, (float)(S - 1));
    const float fl = floorf(pix);
    keys_coeffs(pix - fl, t.w);
    // |pix| stays far below 2^31 for any sane rate; clamp defensively so the int cast is defined
    const int base = (int)fminf(fmaxf(fl, -1.0e9f), 1.0e9f) - 1;
#pragma unroll
    for (int a = 0; a < 4; ++a) t.idx[a] = reflect_clip(base + a, S);
}

// vertical 4-tap pass: dst[r][c] = sum_a w[r][a] * src[idx[r][a]][c]   (rows of `src` have pitch S;
// idx already relative to src's first row)
template <int NT, int ST = 0>
__device__ __forceinline__ void scale_vpass(const float* __restrict__ src, float* __restrict__ dst, int Srt, int nrows,
                                            const AxisTap* __restrict__ rowT)
{
    const int S = ST ? ST : Srt;
    const int CW = S >> 2;
    for (int item = threadIdx.x; item < nrows * CW; item += NT) {
        const int r = item / CW, c4 = item - r * CW;
        const AxisTap t = rowT[r];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)t.idx[a] * S + 4 * c4);
            acc.x = fmaf(t.w[a], v.x, acc.x); acc.y = fmaf(t.w[a], v.y, acc.y);
            acc.z = fmaf(t.w[a], v.z, acc.z); acc.w = fmaf(t.w[a], v.w, acc.w);
        }
        *reinterpret_cast<float4*>(dst + (size_t)r * S + 4 * c4) = acc;
    }
}

// horizontal 4-tap gather for one (row, column): sum_b wx[b] * row[ix[b]]
__device__ __forceinline__ float scale_hgather(const float* __restrict__ row, const AxisTap& t)
{
    float a = row[t.idx[0]] * t.w[0];
    a = fmaf(row[t.idx[1]], t.w[1], a);
    a = fmaf(row[t.idx[2]], t.w[2], a);
    a = fmaf(row[t.idx[3]], t.w[3], a);
    return a;
}

}  // namespace sei
