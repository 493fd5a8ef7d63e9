// cnn_elem.cu -- bandwidth-bound layers of the restoration CNN on channels-last bf16 activations
// (reference: src/models/convolutional.py).
//
// Channel LayerNorm (reference LayerNorm :21-30 = swapaxes + nn.LayerNorm(C, eps=1e-6) + swapaxes): on channels-last
// memory a pixel's channels are one contiguous row of a [T = B*H*W, C] matrix, so no transposition is needed.
//   ln_fwd_kernel      y = (x - mean) * rstd * gamma + beta, statistics in fp32 (two-pass variance), mean / rstd saved
//   ln_bwd_dx_kernel   dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = gy * gamma
//   colsum_kernel<0>   dgamma[c] = sum_t gy * xhat, dbeta[c] = sum_t gy: every thread owns one 8-channel vector and walks
//                      down the rows; fixed-order partial sums + colsum_final_kernel (deterministic, no atomics)
//   colsum_kernel<1>   the bias gradient of the pointwise convolutions, sum_t gy
// Rows are handled by groups of G lanes (G = 4 .. 32, 16-byte vectors per lane), several rows per warp when C is
// small, so every warp-level access is a contiguous >= 512-byte run whatever C is.  VPL > 0: the row is cached in
// registers (C = 8 * G * VPL); VPL = 0: any C % 8 == 0, the row is re-read from L1/L2 for each pass.
#include "sei_common.cuh"
#include "gelu.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <stdlib.h>

namespace sei {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8])
{
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}
__device__ __forceinline__ float group_sum(float v, int G)
{
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

constexpr int kLnThreads = 256;

struct LnParams {
    const __nv_bfloat16* x;
    const __nv_bfloat16* gy;
    const float* gamma;
    const float* beta;
    __nv_bfloat16* out;      // y (forward) or dx (backward)
    float* mean;
    float* rstd;
    long long T;
    int C, G;
    float eps;
};

// GT: lanes per row as a compile-time constant (0 = run time).  ncu on the C = 32 / 128 shapes: issue slots 75 %, half of
// the instructions integer / branch (run-time trip counts of the shuffle reductions, 64-bit index arithmetic) and 31 % of
// the stall samples on the register copy that ended the one-deep prefetch -- at the power-capped clock of a training
// step an issue-bound kernel is 1.4 x slower than alone.  With GT the reductions unroll, and the prefetch alternates
// between two register buffers instead of copying.
template <int VPL, int GT = 0>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const __grid_constant__ LnParams p)
{
    constexpr int NV = VPL > 0 ? VPL : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = GT ? GT : p.G;
    const int gl = lane & (G - 1), sub = lane / G, rpw = 32 / G;
    const int nvec = p.C >> 3;
    const float invC = 1.0f / (float)p.C;
    const long long row_step = (long long)gridDim.x * (kLnThreads / 32) * rpw;
    long long rb = ((long long)blockIdx.x * (kLnThreads / 32) + warp) * rpw;
    // register-cached rows are prefetched one iteration ahead (all warps run this loop in lock step: without the
    // prefetch the loads of an iteration only start after the previous iteration's stores were issued)
    uint4 cache[NV], ahead[NV];
    auto fetch = [&](long long rbase, uint4 (&dst)[NV]) {
        const long long r = rbase + sub;
        const bool ok = r < p.T;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + (ok ? r : 0) * p.C);
#pragma unroll
        for (int i = 0; i < NV; ++i) dst[i] = ok ? __ldcs(xr + gl + i * G) : make_uint4(0, 0, 0, 0);
    };
    constexpr bool PF = VPL > 0 && VPL <= 2;       // prefetching 8 vectors per lane would halve the occupancy
    // a lane keeps the same columns for every row: their gamma / beta live in registers (the first version re-read them
    // per row, 64 B of L1 traffic per 16 B of data: the kernels ran at 0.4 of the HBM rate)
    constexpr int NP = PF ? NV : 1;
    float gam[NP][8], bet[NP][8];
    if (PF) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            load8f(p.gamma + 8 * (gl + i * G), gam[i]);
            load8f(p.beta + 8 * (gl + i * G), bet[i]);
        }
    }
    if (PF) {
        auto process = [&](long long rbase, const uint4 (&c)[NV]) {
            const long long r = rbase + sub;
            float f[NP][8];
            float sm = 0.f;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                unpack8(c[i], f[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) sm += f[i][j];
            }
            const float mu = group_sum(sm, G) * invC;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) q = fmaf(f[i][j] - mu, f[i][j] - mu, q);
            const float rs = rsqrtf(group_sum(q, G) * invC + p.eps);
            if (r < p.T) {
                uint4* yr = reinterpret_cast<uint4*>(p.out + r * p.C);
#pragma unroll
                for (int i = 0; i < NP; ++i) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[i][j] = fmaf((f[i][j] - mu) * rs, gam[i][j], bet[i][j]);
                    __stcs(yr + gl + i * G, pack8(f[i]));
                }
                if (gl == 0) {
                    p.mean[r] = mu;
                    p.rstd[r] = rs;
                }
            }
        };
        fetch(rb, cache);
        for (; rb < p.T; rb += 2 * row_step) {
            fetch(rb + row_step, ahead);             // in flight while `cache` is reduced (rows past the end: zeros)
            process(rb, cache);
            fetch(rb + 2 * row_step, cache);
            process(rb + row_step, ahead);
        }
        return;
    }
    for (; rb < p.T; rb += row_step) {
        const long long r = rb + sub;
        const bool active = r < p.T;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + (active ? r : 0) * p.C);
        if (VPL > 0) fetch(rb, cache);
        float s = 0.f;
        if (VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float f[8];
                unpack8(cache[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[j];
            }
        } else {
            for (int v = gl; v < nvec; v += G) {
                float f[8];
                unpack8(active ? __ldg(xr + v) : make_uint4(0, 0, 0, 0), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[j];
            }
        }
        const float mu = group_sum(s, G) * invC;
        float q = 0.f;
        if (VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float f[8];
                unpack8(cache[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) q = fmaf(f[j] - mu, f[j] - mu, q);
            }
        } else {
            for (int v = gl; v < nvec; v += G) {
                float f[8];
                unpack8(active ? __ldg(xr + v) : make_uint4(0, 0, 0, 0), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) q = fmaf(f[j] - mu, f[j] - mu, q);
            }
        }
        const float rs = rsqrtf(group_sum(q, G) * invC + p.eps);
        if (active) {
            uint4* yr = reinterpret_cast<uint4*>(p.out + r * p.C);
            auto emit = [&](int v, const uint4& raw) {
                float f[8], ga[8], be[8];
                unpack8(raw, f);
                load8f(p.gamma + 8 * v, ga);
                load8f(p.beta + 8 * v, be);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf((f[j] - mu) * rs, ga[j], be[j]);
                __stcs(yr + v, pack8(f));
            };
            if (VPL > 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) emit(gl + i * G, cache[i]);
            } else {
                for (int v = gl; v < nvec; v += G) emit(v, __ldg(xr + v));
            }
            if (gl == 0) {
                p.mean[r] = mu;
                p.rstd[r] = rs;
            }
        }
    }
}

template <int VPL, int GT = 0>
__global__ void __launch_bounds__(kLnThreads, 2) ln_bwd_dx_kernel(const __grid_constant__ LnParams p)
{
    constexpr int NV = VPL > 0 ? VPL : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = GT ? GT : p.G;
    const int gl = lane & (G - 1), sub = lane / G, rpw = 32 / G;
    const int nvec = p.C >> 3;
    const float invC = 1.0f / (float)p.C;
    const long long row_step = (long long)gridDim.x * (kLnThreads / 32) * rpw;
    long long rb = ((long long)blockIdx.x * (kLnThreads / 32) + warp) * rpw;
    uint4 cx[NV], cg[NV], ax[NV], ag[NV];
    auto fetch = [&](long long rbase, uint4 (&dx)[NV], uint4 (&dg)[NV]) {
        const long long r = rbase + sub;
        const bool ok = r < p.T;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + (ok ? r : 0) * p.C);
        const uint4* gr = reinterpret_cast<const uint4*>(p.gy + (ok ? r : 0) * p.C);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            dx[i] = ok ? __ldcs(xr + gl + i * G) : make_uint4(0, 0, 0, 0);
            dg[i] = ok ? __ldcs(gr + gl + i * G) : make_uint4(0, 0, 0, 0);
        }
    };
    // (prefetching the next rows, as the forward kernel does, was measured slower here: 2.7 -> 3.0 ms for C = 32 / 128)
    constexpr bool PF = false;
    // Register-cached shapes (VPL = 1, 2: C <= 512): gamma stays in registers, and the parameter gradients
    // dgamma[c] = sum_t gy * xhat, dbeta[c] = sum_t gy of the lane's columns are accumulated here -- the separate
    // colsum_kernel<0> pass re-read both operands (4.7 ms per step) -- then combined across the lanes / warps that share
    // columns and written as one partial row per CTA (p.beta doubles as the partial buffer: [gridDim.x][2][C]).
    constexpr bool FUSE = VPL > 0 && VPL <= 2;
    constexpr int NP = FUSE ? NV : 1;
    float gam[NP][8], dgam[NP][8], dbet[NP][8];
    if (FUSE) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            load8f(p.gamma + 8 * (gl + i * G), gam[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) dgam[i][j] = dbet[i][j] = 0.f;
        }
    }
    if (PF && VPL > 0 && rb < p.T) fetch(rb, cx, cg);
    for (; rb < p.T; rb += row_step) {
        const long long r = rb + sub;
        const bool active = r < p.T;
        const long long rr = active ? r : 0;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + rr * p.C);
        const uint4* gr = reinterpret_cast<const uint4*>(p.gy + rr * p.C);
        const float mu = __ldg(p.mean + rr), rs = __ldg(p.rstd + rr);
        if (PF) {
            if (VPL > 0 && rb + row_step < p.T) fetch(rb + row_step, ax, ag);
        } else if (VPL > 0) {
            fetch(rb, cx, cg);
        }
        float s1 = 0.f, s2 = 0.f;
        if (FUSE) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                float fx[8], fg[8];
                unpack8(cx[i], fx);
                unpack8(cg[i], fg);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mu) * rs, gq = fg[j] * gam[i][j];
                    s1 += gq;
                    s2 = fmaf(gq, xh, s2);
                    dgam[i][j] = fmaf(fg[j], xh, dgam[i][j]);        // (inactive rows were loaded as zeros)
                    dbet[i][j] += fg[j];
                }
            }
        }
        auto accumulate = [&](int v, const uint4& rx, const uint4& rg) {
            float fx[8], fg[8], ga[8];
            unpack8(rx, fx);
            unpack8(rg, fg);
            load8f(p.gamma + 8 * v, ga);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = fg[j] * ga[j];
                s1 += g;
                s2 = fmaf(g, (fx[j] - mu) * rs, s2);
            }
        };
        if (FUSE) {
        } else if (VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) accumulate(gl + i * G, cx[i], cg[i]);
        } else {
            for (int v = gl; v < nvec; v += G)
                accumulate(v, active ? __ldg(xr + v) : make_uint4(0, 0, 0, 0), active ? __ldg(gr + v) : make_uint4(0, 0, 0, 0));
        }
        const float m1 = group_sum(s1, G) * invC, m2 = group_sum(s2, G) * invC;
        if (active) {
            uint4* dr = reinterpret_cast<uint4*>(p.out + r * p.C);
            auto emit = [&](int v, const uint4& rx, const uint4& rg) {
                float fx[8], fg[8], ga[8];
                unpack8(rx, fx);
                unpack8(rg, fg);
                load8f(p.gamma + 8 * v, ga);
#pragma unroll
                for (int j = 0; j < 8; ++j) fx[j] = rs * (fg[j] * ga[j] - m1 - (fx[j] - mu) * rs * m2);
                __stcs(dr + v, pack8(fx));
            };
            if (FUSE) {
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    float fx[8], fg[8];
                    unpack8(cx[i], fx);
                    unpack8(cg[i], fg);
#pragma unroll
                    for (int j = 0; j < 8; ++j) fx[j] = rs * (fg[j] * gam[i][j] - m1 - (fx[j] - mu) * rs * m2);
                    __stcs(dr + gl + i * G, pack8(fx));
                }
            } else if (VPL > 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) emit(gl + i * G, cx[i], cg[i]);
            } else {
                for (int v = gl; v < nvec; v += G) emit(v, __ldg(xr + v), __ldg(gr + v));
            }
        }
        if (PF && VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                cx[i] = ax[i];
                cg[i] = ag[i];
            }
        }
    }
    if (FUSE) {
        __shared__ float red[kLnThreads / 32][1024];        // per warp: [vector][dgamma 8 | dbeta 8], nvec <= 64
        // lanes of a warp that own the same columns (different rows): xor-shuffle over the row index bits
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                for (int o = G; o < 32; o <<= 1) {
                    dgam[i][j] += __shfl_xor_sync(0xffffffffu, dgam[i][j], o);
                    dbet[i][j] += __shfl_xor_sync(0xffffffffu, dbet[i][j], o);
                }
        if (sub == 0) {
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    red[warp][(gl + i * G) * 16 + j] = dgam[i][j];
                    red[warp][(gl + i * G) * 16 + 8 + j] = dbet[i][j];
                }
        }
        __syncthreads();
        float* out = const_cast<float*>(p.beta) + (size_t)blockIdx.x * 2 * p.C;
        for (int e = threadIdx.x; e < 2 * p.C; e += kLnThreads) {
            const int which = e / p.C, c = e - which * p.C;         // 0: dgamma, 1: dbeta
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kLnThreads / 32; ++w) sum += red[w][(c >> 3) * 16 + which * 8 + (c & 7)];
            out[e] = sum;
        }
    }
}

// Column sums over the rows of [T, C] matrices.  Every thread owns one 8-channel vector (total threads is a multiple
// of nvec = C / 8, so it keeps the same channels while it strides down the rows); threads of a CTA that share a vector
// are summed through shared memory in a fixed order, and each CTA writes one partial row:
//   MODE 0 (LayerNorm parameter gradients): partial[slot][0][c] = sum gy * xhat, partial[slot][1][c] = sum gy
//   MODE 1 (bias gradient of a pointwise convolution):  partial[slot][0][c] = sum gy
// colsum_final_kernel then adds the partial rows (fixed order: deterministic, no atomics).
//   MODE 2 (bias gradient through a resampler): partial[slot][0][c] = sum gy * rowweight[r % period]  (mean = the table)
template <int MODE>
__global__ void __launch_bounds__(kLnThreads) colsum_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ gy,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ partial, long long T, int C, int period = 1)
{
    constexpr int NQ = MODE == 0 ? 2 : 1;
    __shared__ float red[kLnThreads * 8 * NQ];
    const int nvec = C >> 3;
    const long long tg = (long long)blockIdx.x * kLnThreads + threadIdx.x;
    const long long total = (long long)gridDim.x * kLnThreads;
    const int cv = (int)(tg % nvec);
    const long long rp = tg / nvec, RP = total / nvec;
    float dg[8], db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[j] = db[j] = 0.f;
    for (long long r = rp; r < T; r += RP) {
        float fg[8];
        unpack8(__ldcs(reinterpret_cast<const uint4*>(gy + r * C) + cv), fg);
        if (MODE == 0) {
            float fx[8];
            unpack8(__ldcs(reinterpret_cast<const uint4*>(x + r * C) + cv), fx);
            const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
            for (int j = 0; j < 8; ++j) dg[j] = fmaf(fg[j], (fx[j] - mu) * rs, dg[j]);
        }
        if (MODE == 2) {
            const float wr = __ldg(mean + (r % period));
#pragma unroll
            for (int j = 0; j < 8; ++j) db[j] = fmaf(fg[j], wr, db[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) db[j] += fg[j];
        }
    }
    // in-CTA reduction over the threads that own the same vector (only when a CTA spans several rows: nvec < 256)
    const int per_cta = nvec < kLnThreads ? kLnThreads / nvec : 1;      // row slots per CTA
    if (per_cta > 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[(j * NQ) * kLnThreads + threadIdx.x] = db[j];
            if (MODE == 0) red[(j * NQ + 1) * kLnThreads + threadIdx.x] = dg[j];
        }
        __syncthreads();
        if ((int)threadIdx.x < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float sb = 0.f, sg = 0.f;
                for (int k = 0; k < per_cta; ++k) {
                    sb += red[(j * NQ) * kLnThreads + threadIdx.x + k * nvec];
                    if (MODE == 0) sg += red[(j * NQ + 1) * kLnThreads + threadIdx.x + k * nvec];
                }
                db[j] = sb;
                dg[j] = sg;
            }
        }
    }
    if (per_cta == 1 || (int)threadIdx.x < nvec) {
        const long long slot = per_cta > 1 ? blockIdx.x : rp;
        float* o = partial + (size_t)slot * NQ * C + 8 * cv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {
                o[j] = dg[j];
                o[C + j] = db[j];
            } else {
                o[j] = db[j];
            }
        }
    }
}

// out[i] = sum_slot partial[slot][i], i in [0, n): 32 columns per CTA, 8 warps stride down the slots, fixed-order
// combination through shared memory
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out0,
                                                           float* __restrict__ out1, long long slots, int n, int split)
{
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < n)
        for (long long k = warp; k < slots; k += 8) s += partial[(size_t)k * n + i];
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][lane];
        if (i < split) out0[i] = t;
        else out1[i - split] = t;
    }
}

// lanes per row and cached vectors per lane for a channel count
static void ln_shape(int C, int* G, int* VPL)
{
    const int nvec = C / 8;
    int g = 1;
    while (g * 2 <= nvec && g < 32) g *= 2;
    *G = g;
    *VPL = (nvec % g == 0 && (nvec / g == 1 || nvec / g == 2 || nvec / g == 8)) ? nvec / g : 0;
}

static unsigned ln_grid(long long T, int G, int sm_count)
{
    const long long rows_per_cta = (long long)(kLnThreads / 32) * (32 / G);
    return (unsigned)std::max<long long>(1, std::min<long long>((T + rows_per_cta - 1) / rows_per_cta, (long long)sm_count * 8));
}

// threads of colsum_kernel: a multiple of nvec (see the kernel); 0 if no such launch exists
static long long colsum_threads(int C, int sm_count)
{
    const int nvec = C / 8;
    if (nvec <= kLnThreads && kLnThreads % nvec == 0) return (long long)sm_count * 4 * kLnThreads;
    if (nvec % kLnThreads == 0) {
        const int per = nvec / kLnThreads;
        return (long long)std::max(1, sm_count * 4 / per) * per * kLnThreads;
    }
    return 0;
}

// partial rows written by colsum_kernel
static long long colsum_slots(int C, long long threads)
{
    const int nvec = C / 8;
    return nvec < kLnThreads ? threads / kLnThreads : threads / nvec;
}

}  // namespace sei

using namespace sei;

extern "C" int sei_ln_cl_forward_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                      float* rstd, long long T, int C, float eps, void* stream)
{
    SEI_REQUIRE(x && gamma && beta && y && mean && rstd, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0, "bad shape T=%lld C=%d (C must be a multiple of 8)", T, C);
    SEI_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "operands must be 16-byte aligned");
    if (T == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    LnParams p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.gamma = gamma; p.beta = beta;
    p.out = static_cast<__nv_bfloat16*>(y); p.mean = mean; p.rstd = rstd; p.T = T; p.C = C; p.eps = eps;
    int VPL;
    ln_shape(C, &p.G, &VPL);
    const unsigned grid = ln_grid(T, p.G, dp.sm_count);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (VPL) {
    case 1:
        if (p.G == 4) ln_fwd_kernel<1, 4><<<grid, kLnThreads, 0, st>>>(p);
        else if (p.G == 16) ln_fwd_kernel<1, 16><<<grid, kLnThreads, 0, st>>>(p);
        else ln_fwd_kernel<1><<<grid, kLnThreads, 0, st>>>(p);
        break;
    case 2:
        if (p.G == 32) ln_fwd_kernel<2, 32><<<grid, kLnThreads, 0, st>>>(p);
        else ln_fwd_kernel<2><<<grid, kLnThreads, 0, st>>>(p);
        break;
    case 8: ln_fwd_kernel<8><<<grid, kLnThreads, 0, st>>>(p); break;
    default: ln_fwd_kernel<0><<<grid, kLnThreads, 0, st>>>(p); break;
    }
    return finish_launch("ln_fwd_kernel");
}

extern "C" long long sei_ln_cl_backward_workspace_bytes(int C)
{
    DeviceProps dp;
    if (get_device_props(&dp) || C < 8 || C % 8) return -1;
    const long long threads = colsum_threads(C, dp.sm_count);
    if (threads == 0) return -1;
    // partial rows: one per column-sum slot, or one per CTA of the fused LayerNorm backward (at most 8 per SM)
    return std::max<long long>(colsum_slots(C, threads), (long long)dp.sm_count * 8) * 2 * C * (long long)sizeof(float);
}

extern "C" int sei_ln_cl_backward_bf16(const void* gy, const void* x, const float* mean, const float* rstd,
                                       const float* gamma, void* dx, float* dgamma, float* dbeta, void* workspace,
                                       long long T, int C, void* stream)
{
    SEI_REQUIRE(gy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0, "bad shape T=%lld C=%d (C must be a multiple of 8)", T, C);
    SEI_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(dx) && aligned16(gamma) && aligned16(workspace),
                "operands must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long threads = colsum_threads(C, dp.sm_count);
    SEI_REQUIRE(threads > 0, "channel count %d unsupported by the LayerNorm parameter-gradient kernel", C);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (T == 0) {
        SEI_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, st));
        SEI_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, st));
        return 0;
    }
    LnParams p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.gy = static_cast<const __nv_bfloat16*>(gy); p.gamma = gamma;
    p.out = static_cast<__nv_bfloat16*>(dx); p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
    p.T = T; p.C = C;
    int VPL;
    ln_shape(C, &p.G, &VPL);
    const unsigned grid = ln_grid(T, p.G, dp.sm_count);
    const char* nofuse = getenv("SEI_LN_NO_FUSE");         // A/B switch: separate parameter-gradient pass
    const bool fused = (VPL == 1 || VPL == 2) && C <= 512 && !(nofuse && *nofuse == '1');
    if (fused) p.beta = static_cast<const float*>(workspace);      // partial rows [grid][2][C] (the backward needs no beta)
    switch (fused ? VPL : (VPL == 1 || VPL == 2 ? -VPL : 0)) {
    case 1:
        if (p.G == 4) ln_bwd_dx_kernel<1, 4><<<grid, kLnThreads, 0, st>>>(p);
        else if (p.G == 16) ln_bwd_dx_kernel<1, 16><<<grid, kLnThreads, 0, st>>>(p);
        else ln_bwd_dx_kernel<1><<<grid, kLnThreads, 0, st>>>(p);
        break;
    case 2:
        if (p.G == 32) ln_bwd_dx_kernel<2, 32><<<grid, kLnThreads, 0, st>>>(p);
        else ln_bwd_dx_kernel<2><<<grid, kLnThreads, 0, st>>>(p);
        break;
    default: ln_bwd_dx_kernel<0><<<grid, kLnThreads, 0, st>>>(p); break;
    }
    rc = finish_launch("ln_bwd_dx_kernel");
    if (rc) return rc;
    if (fused) {
        colsum_final_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), dgamma, dbeta, grid, 2 * C, C);
        return finish_launch("colsum_final_kernel");
    }
    colsum_kernel<0><<<(unsigned)(threads / kLnThreads), kLnThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gy), mean, rstd,
        static_cast<float*>(workspace), T, C);
    rc = finish_launch("colsum_kernel<ln>");
    if (rc) return rc;
    colsum_final_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), dgamma, dbeta,
                                                          colsum_slots(C, threads), 2 * C, C);
    return finish_launch("colsum_final_kernel");
}

// out[c] (fp32) = sum over the rows of x [T, C] (bf16): the bias gradient of a pointwise convolution.
// workspace: sei_ln_cl_backward_workspace_bytes(C) bytes.
extern "C" int sei_colsum_bf16(const void* x, float* out, void* workspace, long long T, int C, void* stream)
{
    SEI_REQUIRE(x && out && workspace, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0, "bad shape T=%lld C=%d (C must be a multiple of 8)", T, C);
    SEI_REQUIRE(aligned16(x) && aligned16(workspace), "operands must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long threads = colsum_threads(C, dp.sm_count);
    SEI_REQUIRE(threads > 0, "channel count %d unsupported by the column-sum kernel", C);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (T == 0) {
        SEI_CUDA(cudaMemsetAsync(out, 0, (size_t)C * 4, st));
        return 0;
    }
    colsum_kernel<1><<<(unsigned)(threads / kLnThreads), kLnThreads, 0, st>>>(
        nullptr, static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, static_cast<float*>(workspace), T, C);
    rc = finish_launch("colsum_kernel<bias>");
    if (rc) return rc;
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), out, out,
                                                      colsum_slots(C, threads), C, C);
    return finish_launch("colsum_final_kernel");
}

// ---------------------------------------------------------------- 3x3 convolution with a few output channels
// The network's last layer (reference UNet.out_conv, src/models/convolutional.py:176: Conv2d(hidden, in_channels = 3,
// kernel_size=3, padding="same")) maps `hidden` channels to 3.  As a GEMM it has N = 3 and needs a 9x unfolded copy
// of its input; here it is a direct CUDA-core kernel on the channels-last input (1.8 GFMA at batch 32: not
// tensor-core work), and so are its input and weight gradients.  Outputs carry 4 channels per pixel (8-byte
// pixels; channel 3 is zero when Cout = 3).
namespace sei {

constexpr int kC3MaxCin = 64, kC3MaxCout = 4;

struct Conv3Params {
    const __nv_bfloat16* x;      // [B, H, W, Cin]
    const __nv_bfloat16* gy;     // [B, H, W, 4]
    const float* w;              // [Cout, Cin, 3, 3]
    const float* bias;           // [Cout] or null
    __nv_bfloat16* y;            // forward: [B, H, W, 4]; dgrad: [B, H, W, Cin]
    float* partial;
    int B, H, W, Cin, Cout;
    long long T;
};

// shared-memory copy of the taps as ws[tap][co][ci]
__device__ __forceinline__ void conv3_load_taps(const Conv3Params& p, float* ws)
{
    for (int i = threadIdx.x; i < 9 * p.Cout * p.Cin; i += blockDim.x) {
        const int ci = i % p.Cin, co = (i / p.Cin) % p.Cout, tap = i / (p.Cin * p.Cout);
        ws[i] = __ldg(p.w + ((size_t)co * p.Cin + ci) * 9 + tap);
    }
    __syncthreads();
}

template <int COUT>
__global__ void __launch_bounds__(256) conv3_small_fwd_kernel(const __grid_constant__ Conv3Params p)
{
    extern __shared__ __align__(16) float ws[];
    conv3_load_taps(p, ws);
    const int H = p.H, W = p.W, Cin = p.Cin, nvec = Cin >> 3;
    float b0[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) b0[co] = p.bias ? __ldg(p.bias + co) : 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < p.T; t += (long long)gridDim.x * blockDim.x) {
        const int w0 = (int)(t % W);
        const long long r = t / W;
        const int h0 = (int)(r % H);
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = b0[co];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hy = h0 + ky - 1;
            if (hy < 0 || hy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int wx = w0 + kx - 1;
                if (wx < 0 || wx >= W) continue;
                const uint4* xp = reinterpret_cast<const uint4*>(p.x + (t + (long long)(ky - 1) * W + (kx - 1)) * Cin);
                const float* wt = ws + (ky * 3 + kx) * COUT * Cin;
                for (int v = 0; v < nvec; ++v) {
                    float f[8];
                    unpack8(__ldg(xp + v), f);
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        const float4 wa = *reinterpret_cast<const float4*>(wt + co * Cin + 8 * v);
                        const float4 wb = *reinterpret_cast<const float4*>(wt + co * Cin + 8 * v + 4);
                        acc[co] = fmaf(f[0], wa.x, acc[co]); acc[co] = fmaf(f[1], wa.y, acc[co]);
                        acc[co] = fmaf(f[2], wa.z, acc[co]); acc[co] = fmaf(f[3], wa.w, acc[co]);
                        acc[co] = fmaf(f[4], wb.x, acc[co]); acc[co] = fmaf(f[5], wb.y, acc[co]);
                        acc[co] = fmaf(f[6], wb.z, acc[co]); acc[co] = fmaf(f[7], wb.w, acc[co]);
                    }
                }
            }
        }
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int co = 0; co < COUT; ++co) o[co] = acc[co];
        uint2 packed;
        *reinterpret_cast<__nv_bfloat162*>(&packed.x) = __floats2bfloat162_rn(o[0], o[1]);
        *reinterpret_cast<__nv_bfloat162*>(&packed.y) = __floats2bfloat162_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(p.y + t * 4) = packed;
    }
}

// input gradient: gx[p][ci] = sum_{ky,kx,co} w[co][ci][ky][kx] * gy[p - (ky-1, kx-1)][co].  NV = Cin / 8.
template <int COUT, int NV>
__global__ void __launch_bounds__(256) conv3_small_dgrad_kernel(const __grid_constant__ Conv3Params p)
{
    extern __shared__ __align__(16) float ws[];
    conv3_load_taps(p, ws);
    constexpr int Cin = NV * 8;
    const int H = p.H, W = p.W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < p.T; t += (long long)gridDim.x * blockDim.x) {
        const int w0 = (int)(t % W);
        const long long r = t / W;
        const int h0 = (int)(r % H);
        float acc[Cin];
#pragma unroll
        for (int c = 0; c < Cin; ++c) acc[c] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hy = h0 - ky + 1;
            if (hy < 0 || hy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int wx = w0 - kx + 1;
                if (wx < 0 || wx >= W) continue;
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.gy + (t - (long long)(ky - 1) * W - (kx - 1)) * 4));
                const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                const float g[4] = {g01.x, g01.y, g23.x, g23.y};
                const float* wt = ws + (ky * 3 + kx) * COUT * Cin;
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
#pragma unroll
                    for (int c4 = 0; c4 < Cin / 4; ++c4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wt + co * Cin + 4 * c4);
                        acc[4 * c4 + 0] = fmaf(g[co], wv.x, acc[4 * c4 + 0]); acc[4 * c4 + 1] = fmaf(g[co], wv.y, acc[4 * c4 + 1]);
                        acc[4 * c4 + 2] = fmaf(g[co], wv.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(g[co], wv.w, acc[4 * c4 + 3]);
                    }
                }
            }
        }
        uint4* o = reinterpret_cast<uint4*>(p.y + t * Cin);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = acc[8 * v + j];
            o[v] = pack8(f);
        }
    }
}

// weight / bias gradient partials.  A CTA holds NG pixel groups of GS = 9 * Cin/8 threads; thread (tap, v) of a
// group accumulates gW[co][8v .. 8v+7][tap] over the group's pixels, the centre-tap / v = 0 thread also gb[co].
// partial[cta][co][ci][tap] (PyTorch weight layout), then [Cout] bias sums; combined by colsum_final_kernel.
template <int COUT>
__global__ void __launch_bounds__(256) conv3_small_wgrad_kernel(const __grid_constant__ Conv3Params p)
{
    extern __shared__ __align__(16) float red[];           // [NG][GS][COUT * 8 + COUT]
    const int H = p.H, W = p.W, Cin = p.Cin, nvec = Cin >> 3;
    const int GS = 9 * nvec, NG = blockDim.x / GS;
    const int g = threadIdx.x / GS, lt = threadIdx.x - g * GS;
    const int tap = lt / nvec, v = lt - tap * nvec;
    const int ky = tap / 3, kx = tap - ky * 3;
    float acc[COUT][8], gb[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        gb[co] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[co][j] = 0.f;
    }
    if (g < NG) {
        for (long long t = (long long)blockIdx.x * NG + g; t < p.T; t += (long long)gridDim.x * NG) {
            const int w0 = (int)(t % W);
            const long long r = t / W;
            const int h0 = (int)(r % H);
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.gy + t * 4));
            const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
            const float gv[4] = {g01.x, g01.y, g23.x, g23.y};
#pragma unroll
            for (int co = 0; co < COUT; ++co) gb[co] += gv[co];
            const int hy = h0 + ky - 1, wx = w0 + kx - 1;
            if (hy < 0 || hy >= H || wx < 0 || wx >= W) continue;
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + (t + (long long)(ky - 1) * W + (kx - 1)) * Cin) + v), f);
#pragma unroll
            for (int co = 0; co < COUT; ++co)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[co][j] = fmaf(gv[co], f[j], acc[co][j]);
        }
    }
    constexpr int PER = COUT * 8 + COUT;
    if (g < NG) {
        float* mine = red + ((size_t)g * GS + lt) * PER;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mine[co * 8 + j] = acc[co][j];
            mine[COUT * 8 + co] = gb[co];
        }
    }
    __syncthreads();
    const int nW = COUT * Cin * 9;
    float* out = p.partial + (size_t)blockIdx.x * (nW + COUT);
    for (int i = threadIdx.x; i < nW + COUT; i += blockDim.x) {
        float s = 0.f;
        if (i < nW) {
            const int tp = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
            const int l = tp * nvec + (ci >> 3);
            for (int k = 0; k < NG; ++k) s += red[((size_t)k * GS + l) * PER + co * 8 + (ci & 7)];
        } else {
            const int co = i - nW;
            const int l = 4 * nvec;                        // centre tap, vector 0
            for (int k = 0; k < NG; ++k) s += red[((size_t)k * GS + l) * PER + COUT * 8 + co];
        }
        out[i] = s;
    }
}

static int conv3_wgrad_ctas(int sm_count) { return sm_count * 4; }

}  // namespace sei

namespace sei {      // csrc/dwconv_tile.cu: shared-memory tile kernel for the 32 -> 3 channel shape
int conv3_wgrad_tile_slots(int sm_count);
int conv3_wgrad_tile(const void* g4, const void* x, float* partial, int B, int H, int W, int sm_count, cudaStream_t st);
}
static bool conv3_use_tiles(int Cin, int Cout)
{
    const char* e = getenv("SEI_CONV3_WGRAD_TILE");      // A/B switch: 0 = the direct kernel of round 1
    return Cin == 32 && Cout == 3 && !(e && *e == '0');
}

extern "C" long long sei_conv3x3_small_workspace_bytes(int Cin, int Cout)
{
    DeviceProps dp;
    if (get_device_props(&dp)) return -1;
    const int slots = std::max(conv3_wgrad_ctas(dp.sm_count), conv3_wgrad_tile_slots(dp.sm_count));
    return (long long)slots * ((long long)Cout * Cin * 9 + Cout) * (long long)sizeof(float);
}

static int conv3_check(int B, int H, int W, int Cin, int Cout)
{
    SEI_REQUIRE(B >= 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
    SEI_REQUIRE(Cin >= 8 && Cin % 8 == 0 && Cin <= kC3MaxCin, "Cin=%d unsupported (multiple of 8, <= %d)", Cin, kC3MaxCin);
    SEI_REQUIRE(Cout >= 1 && Cout <= kC3MaxCout, "Cout=%d unsupported (1..%d)", Cout, kC3MaxCout);
    return 0;
}

extern "C" int sei_conv3x3_small_forward_bf16(const void* x, const float* w, const float* bias, void* y,
                                              int B, int H, int W, int Cin, int Cout, void* stream)
{
    SEI_REQUIRE(x && w && y, "null pointer argument");
    int rc = conv3_check(B, H, W, Cin, Cout);
    if (rc) return rc;
    SEI_REQUIRE(aligned16(x) && aligned16(y), "operands must be 16-byte aligned");
    if (B == 0) return 0;
    DeviceProps dp;
    rc = get_device_props(&dp);
    if (rc) return rc;
    Conv3Params p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.w = w; p.bias = bias; p.y = static_cast<__nv_bfloat16*>(y);
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.T = (long long)B * H * W;
    const size_t smem = (size_t)9 * Cout * Cin * sizeof(float);
    const unsigned grid = (unsigned)std::min<long long>((p.T + 255) / 256, (long long)dp.sm_count * 16);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (Cout) {
    case 1: conv3_small_fwd_kernel<1><<<grid, 256, smem, st>>>(p); break;
    case 2: conv3_small_fwd_kernel<2><<<grid, 256, smem, st>>>(p); break;
    case 3: conv3_small_fwd_kernel<3><<<grid, 256, smem, st>>>(p); break;
    default: conv3_small_fwd_kernel<4><<<grid, 256, smem, st>>>(p); break;
    }
    return finish_launch("conv3_small_fwd_kernel");
}

template <int COUT>
static int conv3_launch_dgrad(const Conv3Params& p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (p.Cin) {
    case 8: conv3_small_dgrad_kernel<COUT, 1><<<grid, 256, smem, st>>>(p); break;
    case 16: conv3_small_dgrad_kernel<COUT, 2><<<grid, 256, smem, st>>>(p); break;
    case 32: conv3_small_dgrad_kernel<COUT, 4><<<grid, 256, smem, st>>>(p); break;
    case 64: conv3_small_dgrad_kernel<COUT, 8><<<grid, 256, smem, st>>>(p); break;
    default: sei::set_error("conv3x3 input gradient: Cin=%d unsupported (8, 16, 32, 64)", p.Cin); return SEI_EINVAL;
    }
    return finish_launch("conv3_small_dgrad_kernel");
}

// gx may be NULL (input gradient not needed).  gw: [Cout, Cin, 3, 3] fp32, gb: [Cout] fp32 (fixed summation order).
extern "C" int sei_conv3x3_small_backward_bf16(const void* gy, const void* x, const float* w, void* gx, float* gw,
                                               float* gb, void* workspace, int B, int H, int W, int Cin, int Cout,
                                               void* stream)
{
    SEI_REQUIRE(gy && x && w && gw && gb && workspace, "null pointer argument");
    int rc = conv3_check(B, H, W, Cin, Cout);
    if (rc) return rc;
    SEI_REQUIRE(aligned16(x) && aligned16(gy) && (!gx || aligned16(gx)) && aligned16(workspace), "operands must be 16-byte aligned");
    DeviceProps dp;
    rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nW = Cout * Cin * 9;
    if (B == 0) {
        SEI_CUDA(cudaMemsetAsync(gw, 0, (size_t)nW * 4, st));
        SEI_CUDA(cudaMemsetAsync(gb, 0, (size_t)Cout * 4, st));
        return 0;
    }
    Conv3Params p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.gy = static_cast<const __nv_bfloat16*>(gy); p.w = w;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.T = (long long)B * H * W;
    if (gx) {
        p.y = static_cast<__nv_bfloat16*>(gx);
        const size_t smem = (size_t)9 * Cout * Cin * sizeof(float);
        const unsigned grid = (unsigned)std::min<long long>((p.T + 255) / 256, (long long)dp.sm_count * 16);
        switch (Cout) {
        case 1: rc = conv3_launch_dgrad<1>(p, grid, smem, st); break;
        case 2: rc = conv3_launch_dgrad<2>(p, grid, smem, st); break;
        case 3: rc = conv3_launch_dgrad<3>(p, grid, smem, st); break;
        default: rc = conv3_launch_dgrad<4>(p, grid, smem, st); break;
        }
        if (rc) return rc;
    }
    p.partial = static_cast<float*>(workspace);
    if (conv3_use_tiles(Cin, Cout)) {
        rc = conv3_wgrad_tile(gy, x, p.partial, B, H, W, dp.sm_count, st);
        if (rc) return rc;
        colsum_final_kernel<<<(nW + Cout + 31) / 32, 256, 0, st>>>(p.partial, gw, gb, conv3_wgrad_tile_slots(dp.sm_count), nW + Cout, nW);
        return finish_launch("colsum_final_kernel");
    }
    const int GS = 9 * (Cin / 8), NG = std::max(1, 256 / GS), threads = NG * GS;
    const int ctas = conv3_wgrad_ctas(dp.sm_count);
    const size_t smem = (size_t)threads * (Cout * 8 + Cout) * sizeof(float);
    switch (Cout) {
    case 1: SEI_CUDA(allow_smem(conv3_small_wgrad_kernel<1>, smem)); conv3_small_wgrad_kernel<1><<<ctas, threads, smem, st>>>(p); break;
    case 2: SEI_CUDA(allow_smem(conv3_small_wgrad_kernel<2>, smem)); conv3_small_wgrad_kernel<2><<<ctas, threads, smem, st>>>(p); break;
    case 3: SEI_CUDA(allow_smem(conv3_small_wgrad_kernel<3>, smem)); conv3_small_wgrad_kernel<3><<<ctas, threads, smem, st>>>(p); break;
    default: SEI_CUDA(allow_smem(conv3_small_wgrad_kernel<4>, smem)); conv3_small_wgrad_kernel<4><<<ctas, threads, smem, st>>>(p); break;
    }
    rc = finish_launch("conv3_small_wgrad_kernel");
    if (rc) return rc;
    colsum_final_kernel<<<(nW + Cout + 31) / 32, 256, 0, st>>>(p.partial, gw, gb, ctas, nW + Cout, nW);
    return finish_launch("colsum_final_kernel");
}

// ---------------------------------------------------------------- depthwise 7x7 convolution, channels-last bf16
// Reference ConvBlock.conv1 (src/models/convolutional.py:36-38): Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim).
// 49 fp32 FMAs per output element and nothing to contract over: CUDA-core work whose floor is the FMA rate (3.3 GFMA
// per call at batch 32), not HBM.  A thread owns 4 channels x 8 consecutive output columns of one row and slides a
// 14-column window over each of the 7 input rows: 224 FMAs per 14 eight-byte loads and 7 tap vectors; the inputs are
// read straight from global memory (the CTA's footprint stays in L1).  The input gradient is the same kernel with the
// taps flipped (the host passes them that way); taps come as wt[49][C] fp32.
namespace sei {

constexpr int kDwRW = 8;         // output columns per work item
constexpr int kDwRH = 8;         // consecutive rows per work item

struct DwParams {
    const __nv_bfloat16* x;      // [B, H, W, C]
    const __nv_bfloat16* gy;     // wgrad only
    const float* wt;             // [49][C]
    const float* bias;           // [C] or null
    const __nv_bfloat16* res;    // [B, H, W, C] or null: y = conv(x) + bias + res_scale * res
    float res_scale;
    __nv_bfloat16* y;
    float* partial;
    int B, H, W, C, wstrips, hgroups;
    long long items;
};

__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4])
{
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

// 14-column window of one input row: x[row][w0 - 3 + j][c0 .. c0 + 3], j = 0 .. 13, zero outside the image.
// CT: compile-time channel count (0 = run time).  With CT known every load is base + immediate; the first version
// spent 2/3 of its issue slots on 64-bit address arithmetic, bounds predicates and parameter reloads (ncu: FFMA 35 %
// of the executed instructions).  INTERIOR: the strip and its halo lie inside the row, no bounds checks.
template <int CT, bool INTERIOR>
__device__ __forceinline__ void dw_load_window(const __nv_bfloat16* __restrict__ rowbase, int C, int w0, int W,
                                               float (&win)[kDwRW + 6][4])
{
    const int Cc = CT ? CT : C;
    const __nv_bfloat16* base = rowbase + (long long)(w0 - 3) * Cc;
#pragma unroll
    for (int j = 0; j < kDwRW + 6; ++j) {
        uint2 raw = make_uint2(0u, 0u);
        if (INTERIOR || (w0 + j - 3 >= 0 && w0 + j - 3 < W)) raw = __ldg(reinterpret_cast<const uint2*>(base + j * Cc));
        unpack4(raw, win[j]);
    }
}

template <int CT>
__global__ void __launch_bounds__(128) dwconv7_kernel(const __grid_constant__ DwParams p)
{
    const int H = p.H, W = p.W, C = CT ? CT : p.C, cqn = C >> 2;
    const int wstrips = p.wstrips, hgroups = p.hgroups;
    const float* __restrict__ wt = p.wt;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < p.items; it += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(it % cqn);
        long long r = it / cqn;
        const int ws = (int)(r % wstrips);
        r /= wstrips;
        const int hg = (int)(r % hgroups);
        const long long b = r / hgroups;
        const int w0 = ws * kDwRW, c0 = cq * 4;
        const bool interior = w0 >= 3 && w0 + kDwRW + 3 <= W;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
            bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        }
        // consecutive output rows by the same thread: the 7 input rows of row h are 6 of the rows of row h - 1, so
        // the CTA's working set stays in L1
        const int hend = min(H, (hg + 1) * kDwRH);
        const __nv_bfloat16* plane = p.x + (b * H) * (long long)W * C + c0;
#pragma unroll 1
        for (int h = hg * kDwRH; h < hend; ++h) {
            float acc[kDwRW][4];
#pragma unroll
            for (int j = 0; j < kDwRW; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[j][c] = bv[c];
#pragma unroll 1
            for (int ky = 0; ky < 7; ++ky) {
                const int hy = h + ky - 3;
                if (hy < 0 || hy >= H) continue;
                const __nv_bfloat16* row = plane + (long long)hy * W * C;
                float win[kDwRW + 6][4];
                if (interior) dw_load_window<CT, true>(row, C, w0, W, win);
                else dw_load_window<CT, false>(row, C, w0, W, win);
                const float* wk = wt + (size_t)(ky * 7) * C + c0;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(wk + kx * C));
#pragma unroll
                    for (int j = 0; j < kDwRW; ++j) {
                        acc[j][0] = fmaf(t.x, win[j + kx][0], acc[j][0]); acc[j][1] = fmaf(t.y, win[j + kx][1], acc[j][1]);
                        acc[j][2] = fmaf(t.z, win[j + kx][2], acc[j][2]); acc[j][3] = fmaf(t.w, win[j + kx][3], acc[j][3]);
                    }
                }
            }
            __nv_bfloat16* orow = p.y + ((b * H + h) * (long long)W + w0) * C + c0;
            if (p.res) {
                const __nv_bfloat16* rrow = p.res + ((b * H + h) * (long long)W + w0) * C + c0;
#pragma unroll
                for (int j = 0; j < kDwRW; ++j) {
                    if (interior || w0 + j < W) {
                        float rv[4];
                        unpack4(__ldg(reinterpret_cast<const uint2*>(rrow + j * C)), rv);
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[j][c] = fmaf(p.res_scale, rv[c], acc[j][c]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kDwRW; ++j) {
                if (interior || w0 + j < W) {
                    uint2 o;
                    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(acc[j][0], acc[j][1]);
                    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(acc[j][2], acc[j][3]);
                    *reinterpret_cast<uint2*>(orow + j * C) = o;
                }
            }
        }
    }
}

// weight / bias gradient partials.  blockIdx.y selects a block of CQB channel quads; a CTA holds NG groups of
// CQB * 7 threads; thread (cq, ky) of a group walks down kDwRH rows of an 8-column strip: per row one 14-column
// window of x[h + ky - 3] and 8 values of gy[h] feed 224 FMAs into gW[c][ky][0..6] for its 4 channels (28 sums; the
// ky = 3 thread also keeps the 4 bias sums).  partial[cta.x][c * 49 + ky * 7 + kx], then [C] bias sums.
template <int CT>
__global__ void __launch_bounds__(224, 2) dwconv7_wgrad_kernel(const __grid_constant__ DwParams p, int CQB, int NG)
{
    extern __shared__ __align__(16) float red[];           // [NG][CQB * 7][32]
    const int H = p.H, W = p.W, C = CT ? CT : p.C;
    const int GS = CQB * 7;
    const int g = threadIdx.x / GS, lt = threadIdx.x - g * GS;
    const int ky = lt / CQB, cql = lt - ky * CQB;
    const int cq = blockIdx.y * CQB + cql, c0 = cq * 4;
    float acc[7][4], gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[k][c] = 0.f;
    if (g < NG) {
        for (long long s = (long long)blockIdx.x * NG + g; s < p.items; s += (long long)gridDim.x * NG) {
            const int ws = (int)(s % p.wstrips);
            const long long r = s / p.wstrips;
            const int hg = (int)(r % p.hgroups);
            const long long b = r / p.hgroups;
            const int w0 = ws * kDwRW;
            const bool interior = w0 >= 3 && w0 + kDwRW + 3 <= W;
            const int hend = min(H, (hg + 1) * kDwRH);
            const __nv_bfloat16* gplane = p.gy + (b * H) * (long long)W * C + c0;
            const __nv_bfloat16* xplane = p.x + (b * H) * (long long)W * C + c0;
#pragma unroll 1
            for (int h = hg * kDwRH; h < hend; ++h) {
                const int hy = h + ky - 3;
                if (hy < 0 || hy >= H) continue;           // (never for ky = 3, which also owns the bias sums)
                const __nv_bfloat16* grow = gplane + ((long long)h * W + w0) * C;
                float gv[kDwRW][4];
#pragma unroll
                for (int j = 0; j < kDwRW; ++j) {
                    uint2 raw = make_uint2(0u, 0u);
                    if (interior || w0 + j < W) raw = __ldg(reinterpret_cast<const uint2*>(grow + j * C));
                    unpack4(raw, gv[j]);
                }
                if (ky == 3) {
#pragma unroll
                    for (int j = 0; j < kDwRW; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c) gb[c] += gv[j][c];
                }
                float win[kDwRW + 6][4];
                if (interior) dw_load_window<CT, true>(xplane + (long long)hy * W * C, C, w0, W, win);
                else dw_load_window<CT, false>(xplane + (long long)hy * W * C, C, w0, W, win);
#pragma unroll
                for (int kx = 0; kx < 7; ++kx)
#pragma unroll
                    for (int j = 0; j < kDwRW; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[kx][c] = fmaf(gv[j][c], win[j + kx][c], acc[kx][c]);
            }
        }
        float* mine = red + ((size_t)g * GS + lt) * 32;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int c = 0; c < 4; ++c) mine[kx * 4 + c] = acc[kx][c];
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[28 + c] = gb[c];
    }
    __syncthreads();
    // this CTA's slice of the partial row: channels [blockIdx.y * CQB * 4, +CQB * 4)
    float* out = p.partial + (size_t)blockIdx.x * ((size_t)C * 50);
    const int nch = CQB * 4;
    for (int i = threadIdx.x; i < nch * 50; i += blockDim.x) {
        const int cl = i / 50, e = i - cl * 50;           // e < 49: tap ky * 7 + kx; e == 49: bias
        const int cqi = cl >> 2, cc = cl & 3;
        const int c = blockIdx.y * nch + cl;
        float s = 0.f;
        if (e < 49) {
            const int kyy = e / 7, kxx = e - kyy * 7;
            for (int k = 0; k < NG; ++k) s += red[((size_t)k * GS + kyy * CQB + cqi) * 32 + kxx * 4 + cc];
            out[(size_t)c * 49 + e] = s;
        } else {
            for (int k = 0; k < NG; ++k) s += red[((size_t)k * GS + 3 * CQB + cqi) * 32 + 28 + cc];
            out[(size_t)C * 49 + c] = s;
        }
    }
}

static void dw_wgrad_shape(int C, int sm_count, int* CQB, int* NG, int* gx, int* gy)
{
    const int cqn = C / 4;
    *CQB = cqn >= 32 ? 32 : cqn;                 // cqn in {2, 4, 8, 16} for small C: a divisor of 32
    *NG = 224 / (*CQB * 7);
    *gy = cqn / *CQB;
    *gx = std::max(1, sm_count * 4 / *gy);
}

}  // namespace sei

namespace sei {      // csrc/dwconv_tile.cu: shared-memory tile kernels for channel counts that are multiples of 64
bool dwconv7_tile_supported(int C);
int dwconv7_tile_wgrad_slots(int C, int sm_count);
int dwconv7_tile_forward(const void* x, const float* wt, const float* bias, const void* res, float res_scale, void* y, int B,
                         int H, int W, int C, cudaStream_t st);
int dwconv7_tile_wgrad(const void* gy, const void* x, float* partial, int slots, int B, int H, int W, int C, cudaStream_t st);
}
static bool dwconv7_use_tiles(int C)
{
    const char* e = getenv("SEI_DWCONV_TILE");           // A/B switch: 0 = the register-window kernels of round 1
    return dwconv7_tile_supported(C) && !(e && *e == '0');
}

extern "C" long long sei_dwconv7_workspace_bytes(int C)
{
    DeviceProps dp;
    if (get_device_props(&dp) || C < 8 || C % 8) return -1;
    const int cqn = C / 4;
    long long bytes = -1;
    if (cqn % 32 == 0 || 32 % cqn == 0) {
        int CQB, NG, gx, gy;
        dw_wgrad_shape(C, dp.sm_count, &CQB, &NG, &gx, &gy);
        bytes = (long long)gx * C * 50 * (long long)sizeof(float);
    }
    if (dwconv7_tile_supported(C))
        bytes = std::max(bytes, (long long)dwconv7_tile_wgrad_slots(C, dp.sm_count) * C * 50 * (long long)sizeof(float));
    return bytes;
}

static int dwconv7_impl(const void* x, const float* wt, const float* bias, const void* res, float res_scale, void* y,
                        int B, int H, int W, int C, void* stream);

// y = depthwise7x7(x) (+ bias); wt: taps as [49][C] fp32 (flipped by the caller for the input gradient)
extern "C" int sei_dwconv7_cl_bf16(const void* x, const float* wt, const float* bias, void* y, int B, int H, int W, int C,
                                   void* stream)
{
    return dwconv7_impl(x, wt, bias, nullptr, 0.f, y, B, H, W, C, stream);
}

// y = depthwise7x7(x) (+ bias) + res_scale * res: the input gradient of a ConvBlock, whose residual branch adds the
// incoming gradient to the gradient that went through the block (autograd's accumulation of `x + x1`, reference
// src/models/convolutional.py:43-51), in the store of the convolution instead of a separate pass
extern "C" int sei_dwconv7_cl_residual_bf16(const void* x, const float* wt, const float* bias, const void* res,
                                            float res_scale, void* y, int B, int H, int W, int C, void* stream)
{
    SEI_REQUIRE(res != nullptr && aligned16(res), "res must be a 16-byte aligned [B, H, W, C] bf16 tensor");
    return dwconv7_impl(x, wt, bias, res, res_scale, y, B, H, W, C, stream);
}

static int dwconv7_impl(const void* x, const float* wt, const float* bias, const void* res, float res_scale, void* y,
                        int B, int H, int W, int C, void* stream)
{
    SEI_REQUIRE(x && wt && y, "null pointer argument");
    SEI_REQUIRE(B >= 0 && H > 0 && W > 0 && C >= 8 && C % 8 == 0, "bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
    SEI_REQUIRE(aligned16(x) && aligned16(y) && aligned16(wt) && (!bias || aligned16(bias)), "operands must be 16-byte aligned");
    if (B == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    if (dwconv7_use_tiles(C)) return dwconv7_tile_forward(x, wt, bias, res, res_scale, y, B, H, W, C, reinterpret_cast<cudaStream_t>(stream));
    DwParams p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.wt = wt; p.bias = bias; p.y = static_cast<__nv_bfloat16*>(y);
    p.res = static_cast<const __nv_bfloat16*>(res); p.res_scale = res_scale;
    p.B = B; p.H = H; p.W = W; p.C = C; p.wstrips = (W + kDwRW - 1) / kDwRW; p.hgroups = (H + kDwRH - 1) / kDwRH;
    p.items = (long long)B * p.hgroups * p.wstrips * (C / 4);
    const unsigned grid = (unsigned)std::min<long long>((p.items + 127) / 128, (long long)dp.sm_count * 64);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (C) {            // the network's widths (hidden 32, x4 per scale) get compile-time strides
    case 32: dwconv7_kernel<32><<<grid, 128, 0, st>>>(p); break;
    case 128: dwconv7_kernel<128><<<grid, 128, 0, st>>>(p); break;
    case 512: dwconv7_kernel<512><<<grid, 128, 0, st>>>(p); break;
    case 2048: dwconv7_kernel<2048><<<grid, 128, 0, st>>>(p); break;
    case 8192: dwconv7_kernel<8192><<<grid, 128, 0, st>>>(p); break;
    default: dwconv7_kernel<0><<<grid, 128, 0, st>>>(p); break;
    }
    return finish_launch("dwconv7_kernel");
}

// gw: [C, 7, 7] fp32 (PyTorch depthwise weight layout [C, 1, 7, 7]), gb: [C] fp32; fixed summation order
extern "C" int sei_dwconv7_wgrad_cl_bf16(const void* gy, const void* x, float* gw, float* gb, void* workspace,
                                         int B, int H, int W, int C, void* stream)
{
    SEI_REQUIRE(gy && x && gw && gb && workspace, "null pointer argument");
    SEI_REQUIRE(B >= 0 && H > 0 && W > 0 && C >= 8 && C % 8 == 0, "bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
    SEI_REQUIRE(sei_dwconv7_workspace_bytes(C) > 0 && (dwconv7_use_tiles(C) || (C / 4) % 32 == 0 || 32 % (C / 4) == 0),
                "channel count %d unsupported by the depthwise weight-gradient kernel", C);
    SEI_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(workspace), "operands must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (B == 0) {
        SEI_CUDA(cudaMemsetAsync(gw, 0, (size_t)C * 49 * 4, st));
        SEI_CUDA(cudaMemsetAsync(gb, 0, (size_t)C * 4, st));
        return 0;
    }
    if (dwconv7_use_tiles(C)) {
        const int slots = dwconv7_tile_wgrad_slots(C, dp.sm_count);
        rc = dwconv7_tile_wgrad(gy, x, static_cast<float*>(workspace), slots, B, H, W, C, st);
        if (rc) return rc;
        colsum_final_kernel<<<(C * 50 + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), gw, gb, slots, C * 50, C * 49);
        return finish_launch("colsum_final_kernel");
    }
    DwParams p = {};
    p.x = static_cast<const __nv_bfloat16*>(x); p.gy = static_cast<const __nv_bfloat16*>(gy);
    p.partial = static_cast<float*>(workspace);
    p.B = B; p.H = H; p.W = W; p.C = C; p.wstrips = (W + kDwRW - 1) / kDwRW; p.hgroups = (H + kDwRH - 1) / kDwRH;
    p.items = (long long)B * p.hgroups * p.wstrips;
    int CQB, NG, gx, gyb;
    dw_wgrad_shape(C, dp.sm_count, &CQB, &NG, &gx, &gyb);
    const size_t smem = (size_t)NG * CQB * 7 * 32 * sizeof(float);
    const dim3 grid(gx, gyb);
    const int threads = NG * CQB * 7;
    switch (C) {
    case 32: dwconv7_wgrad_kernel<32><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    case 128: dwconv7_wgrad_kernel<128><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    case 512: dwconv7_wgrad_kernel<512><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    case 2048: dwconv7_wgrad_kernel<2048><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    case 8192: dwconv7_wgrad_kernel<8192><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    default: dwconv7_wgrad_kernel<0><<<grid, threads, smem, st>>>(p, CQB, NG); break;
    }
    rc = finish_launch("dwconv7_wgrad_kernel");
    if (rc) return rc;
    colsum_final_kernel<<<(C * 50 + 31) / 32, 256, 0, st>>>(p.partial, gw, gb, gx, C * 50, C * 49);
    return finish_launch("colsum_final_kernel");
}

// ---------------------------------------------------------------- GELU (exact, erf form), bf16 elementwise
// Reference ConvBlock.gelu = nn.GELU() (src/models/convolutional.py:41): 0.5 x (1 + erf(x / sqrt 2)).  The library
// kernel is bound by erff's instruction count (it ran at half the HBM rate); erf is evaluated here with the
// Abramowitz-Stegun 7.1.26 rational form (|error| < 1.5e-7, far below bf16 resolution): one MUFU.RCP, one MUFU.EX2 and
// a degree-5 polynomial.  The backward kernel shares the exponential: gelu'(x) = Phi(x) + x phi(x).
namespace sei {

// Software-pipelined: the loads of the next group of kGeluUnroll vectors are issued before the current group is
// evaluated.  All threads run the same grid-stride loop in lock step, so without the prefetch the memory system idles
// while every warp computes and vice versa (the first version's time was the SUM of its memory and instruction times).
constexpr int kGeluUnroll = 4;

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec)
{
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 cur[kGeluUnroll], nxt[kGeluUnroll];
#pragma unroll
    for (int u = 0; u < kGeluUnroll; ++u)
        if (i0 + u * step < nvec) cur[u] = __ldcs(x + i0 + u * step);
    for (; i0 < nvec; i0 += step * kGeluUnroll) {
        const long long i1 = i0 + step * kGeluUnroll;
#pragma unroll
        for (int u = 0; u < kGeluUnroll; ++u)
            if (i1 + u * step < nvec) nxt[u] = __ldcs(x + i1 + u * step);
#pragma unroll
        for (int u = 0; u < kGeluUnroll; ++u) {
            if (i0 + u * step < nvec) {
                float f[8];
                unpack8(cur[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float Phi, phi;
                    gelu_parts(f[j], Phi, phi);
                    f[j] *= Phi;
                }
                __stcs(y + i0 + u * step, pack8(f));
            }
        }
#pragma unroll
        for (int u = 0; u < kGeluUnroll; ++u) cur[u] = nxt[u];
    }
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gy,
                                                       uint4* __restrict__ gx, long long nvec)
{
    constexpr int U = 2;
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 cx[U], cg[U], nx[U], ng[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (i0 + u * step < nvec) {
            cx[u] = __ldcs(x + i0 + u * step);
            cg[u] = __ldcs(gy + i0 + u * step);
        }
    for (; i0 < nvec; i0 += step * U) {
        const long long i1 = i0 + step * U;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i1 + u * step < nvec) {
                nx[u] = __ldcs(x + i1 + u * step);
                ng[u] = __ldcs(gy + i1 + u * step);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * step < nvec) {
                float f[8], g[8];
                unpack8(cx[u], f);
                unpack8(cg[u], g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float Phi, phi;
                    gelu_parts(f[j], Phi, phi);
                    g[j] *= fmaf(f[j], phi, Phi);
                }
                __stcs(gx + i0 + u * step, pack8(g));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cx[u] = nx[u];
            cg[u] = ng[u];
        }
    }
}

}  // namespace sei

// y = gelu(x) (gy == NULL) or gx = gy * gelu'(x); n bf16 elements, n % 8 == 0, 16-byte aligned
extern "C" int sei_gelu_bf16(const void* x, const void* gy, void* out, long long n, void* stream)
{
    SEI_REQUIRE(x && out, "null pointer argument");
    SEI_REQUIRE(n >= 0 && n % 8 == 0, "element count %lld must be a multiple of 8", n);
    SEI_REQUIRE(aligned16(x) && aligned16(out) && (!gy || aligned16(gy)), "operands must be 16-byte aligned");
    if (n == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long nvec = n / 8;
    const unsigned grid = (unsigned)std::min<long long>((nvec + 255) / 256, (long long)dp.sm_count * 16);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (gy)
        gelu_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(gy), static_cast<uint4*>(out), nvec);
    else
        gelu_fwd_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), nvec);
    return finish_launch(gy ? "gelu_bwd_kernel" : "gelu_fwd_kernel");
}


// gx = gy * gelu'(h) on rows [T, C] AND the column sums of gx (the bias gradient of the pointwise convolution in front
// of the GELU: ConvBlock.conv2, reference src/models/convolutional.py:40-41) in the same pass.  colsum_kernel's thread
// layout (a thread keeps one 8-channel vector and strides down the rows, so the sums stay in registers), the
// elementwise kernel's prefetch of the next rows.  Saves the separate read of the 4C-wide gradient per ConvBlock.
namespace sei {

__global__ void __launch_bounds__(kLnThreads) gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ h,
                                                                      const __nv_bfloat16* __restrict__ gy,
                                                                      __nv_bfloat16* __restrict__ gx,
                                                                      float* __restrict__ partial, long long T, int C)
{
    __shared__ float red[kLnThreads * 8];
    constexpr int U = 2;
    const int nvec = C >> 3;
    const long long tg = (long long)blockIdx.x * kLnThreads + threadIdx.x;
    const long long total = (long long)gridDim.x * kLnThreads;
    const int cv = (int)(tg % nvec);
    const long long rp = tg / nvec, RP = total / nvec;
    float db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) db[j] = 0.f;
    uint4 ch[U], cg[U], nh[U], ng[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (rp + u * RP < T) {
            ch[u] = __ldcs(reinterpret_cast<const uint4*>(h + (rp + u * RP) * C) + cv);
            cg[u] = __ldcs(reinterpret_cast<const uint4*>(gy + (rp + u * RP) * C) + cv);
        }
    for (long long r = rp; r < T; r += RP * U) {
        const long long r1 = r + RP * U;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r1 + u * RP < T) {
                nh[u] = __ldcs(reinterpret_cast<const uint4*>(h + (r1 + u * RP) * C) + cv);
                ng[u] = __ldcs(reinterpret_cast<const uint4*>(gy + (r1 + u * RP) * C) + cv);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u * RP < T) {
                float f[8], g[8];
                unpack8(ch[u], f);
                unpack8(cg[u], g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float Phi, phi;
                    gelu_parts(f[j], Phi, phi);
                    g[j] *= fmaf(f[j], phi, Phi);
                }
                const uint4 packed = pack8(g);
                __stcs(reinterpret_cast<uint4*>(gx + (r + u * RP) * C) + cv, packed);
                unpack8(packed, g);                 // the sums are those of the ROUNDED gradient, as a separate pass would see it
#pragma unroll
                for (int j = 0; j < 8; ++j) db[j] += g[j];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ch[u] = nh[u];
            cg[u] = ng[u];
        }
    }
    const int per_cta = nvec < kLnThreads ? kLnThreads / nvec : 1;
    if (per_cta > 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[j * kLnThreads + threadIdx.x] = db[j];
        __syncthreads();
        if ((int)threadIdx.x < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float sb = 0.f;
                for (int k = 0; k < per_cta; ++k) sb += red[j * kLnThreads + threadIdx.x + k * nvec];
                db[j] = sb;
            }
        }
    }
    if (per_cta == 1 || (int)threadIdx.x < nvec) {
        const long long slot = per_cta > 1 ? blockIdx.x : rp;
        float* o = partial + (size_t)slot * C + 8 * cv;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = db[j];
    }
}

}  // namespace sei

// gx = gy * gelu'(h) and gb[c] = sum_t gx[t, c]; h, gy, gx: bf16 [T, C]; workspace: sei_ln_cl_backward_workspace_bytes(C)
extern "C" int sei_gelu_bwd_colsum_bf16(const void* h, const void* gy, void* gx, float* gb, void* workspace, long long T,
                                        int C, void* stream)
{
    SEI_REQUIRE(h && gy && gx && gb && workspace, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0, "bad shape T=%lld C=%d (C must be a multiple of 8)", T, C);
    SEI_REQUIRE(aligned16(h) && aligned16(gy) && aligned16(gx) && aligned16(workspace), "operands must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long threads = colsum_threads(C, dp.sm_count);
    SEI_REQUIRE(threads > 0, "channel count %d unsupported by the column-sum kernels", C);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (T == 0) {
        SEI_CUDA(cudaMemsetAsync(gb, 0, (size_t)C * 4, st));
        return 0;
    }
    gelu_bwd_colsum_kernel<<<(unsigned)(threads / kLnThreads), kLnThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(gy), static_cast<__nv_bfloat16*>(gx),
        static_cast<float*>(workspace), T, C);
    rc = finish_launch("gelu_bwd_colsum_kernel");
    if (rc) return rc;
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), gb, gb, colsum_slots(C, threads), C, C);
    return finish_launch("colsum_final_kernel");
}

// ---------------------------------------------------------------- bias behind a resampler
// Downsample = LayerNorm -> conv1x1 -> ideal resampler; the resampler is applied BEFORE the convolution here (they
// commute), so the convolution's bias must be pushed through it: the constant image bias[c] becomes
// bias[c] * R(1)[h, w] with R(1) the resampler's response to the all-ones image (models/resample.constant_response).
//   out[t][c] += pat[t % period] * bias[c]          (in place, rows = pixels, period = Ho * Wo)
//   gbias[c]   = sum_t gy[t][c] * pat[t % period]   (colsum_kernel<2>)
namespace sei {

__global__ void __launch_bounds__(256) bias_pattern_add_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ pat,
                                                               const float* __restrict__ bias, long long T, int C, int period)
{
    const int nvec = C >> 3;
    const long long total = T * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / nvec;
        const int v = (int)(i - r * nvec);
        const float wr = __ldg(pat + (r % period));
        float f[8], b[8];
        uint4* ptr = reinterpret_cast<uint4*>(out + r * C) + v;
        unpack8(*ptr, f);
        load8f(bias + 8 * v, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(wr, b[j], f[j]);
        *ptr = pack8(f);
    }
}

}  // namespace sei

extern "C" int sei_bias_pattern_add_bf16(void* out, const float* pat, const float* bias, long long T, int C, int period,
                                         void* stream)
{
    SEI_REQUIRE(out && pat && bias, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0 && period >= 1, "bad shape T=%lld C=%d period=%d", T, C, period);
    SEI_REQUIRE(aligned16(out) && aligned16(bias), "operands must be 16-byte aligned");
    if (T == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long total = T * (C / 8);
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)dp.sm_count * 16);
    bias_pattern_add_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(out), pat, bias, T, C, period);
    return finish_launch("bias_pattern_add_kernel");
}

// gbias[c] = sum_t gy[t][c] * pat[t % period]; workspace as for sei_colsum_bf16
extern "C" int sei_bias_pattern_grad_bf16(const void* gy, const float* pat, float* gbias, void* workspace, long long T, int C,
                                          int period, void* stream)
{
    SEI_REQUIRE(gy && pat && gbias && workspace, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 8 && C % 8 == 0 && period >= 1, "bad shape T=%lld C=%d period=%d", T, C, period);
    SEI_REQUIRE(aligned16(gy) && aligned16(workspace), "operands must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long threads = colsum_threads(C, dp.sm_count);
    SEI_REQUIRE(threads > 0, "channel count %d unsupported by the column-sum kernel", C);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (T == 0) {
        SEI_CUDA(cudaMemsetAsync(gbias, 0, (size_t)C * 4, st));
        return 0;
    }
    colsum_kernel<2><<<(unsigned)(threads / kLnThreads), kLnThreads, 0, st>>>(
        nullptr, static_cast<const __nv_bfloat16*>(gy), pat, nullptr, static_cast<float*>(workspace), T, C, period);
    rc = finish_launch("colsum_kernel<pattern>");
    if (rc) return rc;
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), gbias, gbias,
                                                      colsum_slots(C, threads), C, C);
    return finish_launch("colsum_final_kernel");
}

// ---------------------------------------------------------------- channel LayerNorm for a handful of channels
// The SR model's input stage normalises over the 3 image channels (reference Upsample(in_channels=3): LayerNorm(3),
// src/models/convolutional.py:95-104): 2.1 M rows of 3 values at 512 x 512 x 8.  The library's row-wise kernels
// take 9 ms per call on such rows; one thread per row is all it needs.  Any C <= 32.
namespace sei {

constexpr int kLnSmallMaxC = 32;

__global__ void __launch_bounds__(256) ln_small_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                           float* __restrict__ mean, float* __restrict__ rstd,
                                                           long long T, int C, float eps)
{
    const float invC = 1.0f / (float)C;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < T; r += (long long)gridDim.x * blockDim.x) {
        float v[kLnSmallMaxC];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kLnSmallMaxC; ++c)
            if (c < C) {
                v[c] = __bfloat162float(x[r * C + c]);
                s += v[c];
            }
        const float mu = s * invC;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < kLnSmallMaxC; ++c)
            if (c < C) q = fmaf(v[c] - mu, v[c] - mu, q);
        const float rs = rsqrtf(q * invC + eps);
#pragma unroll
        for (int c = 0; c < kLnSmallMaxC; ++c)
            if (c < C) y[r * C + c] = __float2bfloat16_rn(fmaf((v[c] - mu) * rs, __ldg(gamma + c), __ldg(beta + c)));
        mean[r] = mu;
        rstd[r] = rs;
    }
}

// dx per row; per-CTA partial sums of dgamma / dbeta -> partial[cta][2][C] (combined by colsum_final_kernel)
__global__ void __launch_bounds__(256) ln_small_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                           float* __restrict__ partial, long long T, int C)
{
    __shared__ float red[8][2 * kLnSmallMaxC];
    const float invC = 1.0f / (float)C;
    float dg[kLnSmallMaxC], db[kLnSmallMaxC];
#pragma unroll
    for (int c = 0; c < kLnSmallMaxC; ++c) dg[c] = db[c] = 0.f;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < T; r += (long long)gridDim.x * blockDim.x) {
        const float mu = mean[r], rs = rstd[r];
        float xh[kLnSmallMaxC], g[kLnSmallMaxC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < kLnSmallMaxC; ++c)
            if (c < C) {
                xh[c] = (__bfloat162float(x[r * C + c]) - mu) * rs;
                const float go = __bfloat162float(gy[r * C + c]);
                dg[c] = fmaf(go, xh[c], dg[c]);
                db[c] += go;
                g[c] = go * __ldg(gamma + c);
                s1 += g[c];
                s2 = fmaf(g[c], xh[c], s2);
            }
        const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
        for (int c = 0; c < kLnSmallMaxC; ++c)
            if (c < C) dx[r * C + c] = __float2bfloat16_rn(rs * (g[c] - m1 - xh[c] * m2));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < kLnSmallMaxC; ++c)
        if (c < C) {
            const float a = warp_sum(dg[c]), b = warp_sum(db[c]);
            if (lane == 0) {
                red[warp][c] = a;
                red[warp][C + c] = b;
            }
        }
    __syncthreads();
    if ((int)threadIdx.x < 2 * C) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 2 * C + threadIdx.x] = s;
    }
}

static int ln_small_ctas(int sm_count) { return sm_count * 8; }

// ---- any channel count (the widths 3 * 4^s of the network without in / out convolutions: 48, 192, ...): one warp per
// row, lanes strided over the channels; fp32 two-pass statistics.  A correctness path, not a tuned one.
__global__ void __launch_bounds__(256) ln_warp_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ mean, float* __restrict__ rstd,
                                                          long long T, int C, float eps)
{
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const float invC = 1.0f / (float)C;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < T; r += nwarps) {
        const __nv_bfloat16* xr = x + r * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += __bfloat162float(xr[c]);
        const float mu = warp_sum(s) * invC;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = __bfloat162float(xr[c]) - mu;
            q = fmaf(d, d, q);
        }
        const float rs = rsqrtf(warp_sum(q) * invC + eps);
        for (int c = lane; c < C; c += 32)
            y[r * C + c] = __float2bfloat16_rn(fmaf((__bfloat162float(xr[c]) - mu) * rs, __ldg(gamma + c), __ldg(beta + c)));
        if (lane == 0) {
            mean[r] = mu;
            rstd[r] = rs;
        }
    }
}

__global__ void __launch_bounds__(256) ln_warp_bwd_dx_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                             long long T, int C)
{
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const float invC = 1.0f / (float)C;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < T; r += nwarps) {
        const float mu = mean[r], rs = rstd[r];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float xh = (__bfloat162float(x[r * C + c]) - mu) * rs;
            const float g = __bfloat162float(gy[r * C + c]) * __ldg(gamma + c);
            s1 += g;
            s2 = fmaf(g, xh, s2);
        }
        const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
        for (int c = lane; c < C; c += 32) {
            const float xh = (__bfloat162float(x[r * C + c]) - mu) * rs;
            const float g = __bfloat162float(gy[r * C + c]) * __ldg(gamma + c);
            dx[r * C + c] = __float2bfloat16_rn(rs * (g - m1 - xh * m2));
        }
    }
}

// dgamma / dbeta partials: thread = channel, CTA (x: 128-channel group, y: slab of rows) walks its rows in order;
// partial[slab][2][C], combined in fixed order by colsum_final_kernel (deterministic, no atomics)
__global__ void __launch_bounds__(128) ln_warp_param_grad_kernel(const __nv_bfloat16* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ gy,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 float* __restrict__ partial, long long T, int C)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    const long long per = (T + gridDim.y - 1) / gridDim.y;
    const long long r_lo = per * blockIdx.y, r_hi = min(T, r_lo + per);
    float dg = 0.f, db = 0.f;
    for (long long r = r_lo; r < r_hi; ++r) {
        const float go = __bfloat162float(gy[r * C + c]);
        dg = fmaf(go, (__bfloat162float(x[r * C + c]) - mean[r]) * rstd[r], dg);
        db += go;
    }
    partial[(size_t)blockIdx.y * 2 * C + c] = dg;
    partial[(size_t)blockIdx.y * 2 * C + C + c] = db;
}

static int ln_warp_slabs(long long T, int sm_count) { return (int)std::max<long long>(1, std::min<long long>(T / 64, (long long)sm_count * 4)); }

}  // namespace sei

extern "C" long long sei_ln_small_workspace_bytes(int C)
{
    DeviceProps dp;
    if (get_device_props(&dp) || C < 1 || C > 65536) return -1;
    // C <= 32: one partial row per CTA of the one-thread-per-row kernel; larger C: one per row slab (at most 4 per SM)
    return (long long)(C <= kLnSmallMaxC ? ln_small_ctas(dp.sm_count) : dp.sm_count * 4) * 2 * C * (long long)sizeof(float);
}

extern "C" int sei_ln_small_forward_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                         float* rstd, long long T, int C, float eps, void* stream)
{
    SEI_REQUIRE(x && gamma && beta && y && mean && rstd, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 1 && C <= 65536, "bad shape T=%lld C=%d", T, C);
    if (T == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    if (C > kLnSmallMaxC) {
        const unsigned gridw = (unsigned)std::min<long long>((T + 7) / 8, (long long)dp.sm_count * 8);
        ln_warp_fwd_kernel<<<gridw, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, T, C, eps);
        return finish_launch("ln_warp_fwd_kernel");
    }
    const unsigned grid = (unsigned)std::min<long long>((T + 255) / 256, (long long)ln_small_ctas(dp.sm_count));
    ln_small_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, T, C, eps);
    return finish_launch("ln_small_fwd_kernel");
}

extern "C" int sei_ln_small_backward_bf16(const void* gy, const void* x, const float* mean, const float* rstd,
                                          const float* gamma, void* dx, float* dgamma, float* dbeta, void* workspace,
                                          long long T, int C, void* stream)
{
    SEI_REQUIRE(gy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "null pointer argument");
    SEI_REQUIRE(T >= 0 && C >= 1 && C <= 65536, "bad shape T=%lld C=%d", T, C);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (T == 0) {
        SEI_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, st));
        SEI_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, st));
        return 0;
    }
    if (C > kLnSmallMaxC) {
        const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
        const __nv_bfloat16* gb = static_cast<const __nv_bfloat16*>(gy);
        const unsigned gridw = (unsigned)std::min<long long>((T + 7) / 8, (long long)dp.sm_count * 8);
        ln_warp_bwd_dx_kernel<<<gridw, 256, 0, st>>>(xb, gb, mean, rstd, gamma, static_cast<__nv_bfloat16*>(dx), T, C);
        rc = finish_launch("ln_warp_bwd_dx_kernel");
        if (rc) return rc;
        const int slabs = ln_warp_slabs(T, dp.sm_count);
        ln_warp_param_grad_kernel<<<dim3((C + 127) / 128, slabs), 128, 0, st>>>(xb, gb, mean, rstd, static_cast<float*>(workspace), T, C);
        rc = finish_launch("ln_warp_param_grad_kernel");
        if (rc) return rc;
        colsum_final_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), dgamma, dbeta, slabs, 2 * C, C);
        return finish_launch("colsum_final_kernel");
    }
    const int grid = (int)std::min<long long>((T + 255) / 256, (long long)ln_small_ctas(dp.sm_count));
    ln_small_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gy), mean,
                                              rstd, gamma, static_cast<__nv_bfloat16*>(dx), static_cast<float*>(workspace), T, C);
    rc = finish_launch("ln_small_bwd_kernel");
    if (rc) return rc;
    colsum_final_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), dgamma, dbeta, grid, 2 * C, C);
    return finish_launch("colsum_final_kernel");
}
