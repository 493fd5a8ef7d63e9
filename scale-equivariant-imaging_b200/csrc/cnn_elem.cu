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
#include <cuda_bf16.h>
#include <algorithm>

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

template <int VPL>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const __grid_constant__ LnParams p)
{
    constexpr int NV = VPL > 0 ? VPL : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = p.G;
    const int gl = lane & (G - 1), sub = lane / G, rpw = 32 / G;
    const int nvec = p.C >> 3;
    const float invC = 1.0f / (float)p.C;
    const long long row_step = (long long)gridDim.x * (kLnThreads / 32) * rpw;
    for (long long rb = ((long long)blockIdx.x * (kLnThreads / 32) + warp) * rpw; rb < p.T; rb += row_step) {
        const long long r = rb + sub;
        const bool active = r < p.T;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + (active ? r : 0) * p.C);
        uint4 cache[NV];
        float s = 0.f;
        if (VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                cache[i] = active ? __ldcs(xr + gl + i * G) : make_uint4(0, 0, 0, 0);
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

template <int VPL>
__global__ void __launch_bounds__(kLnThreads, 2) ln_bwd_dx_kernel(const __grid_constant__ LnParams p)
{
    constexpr int NV = VPL > 0 ? VPL : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = p.G;
    const int gl = lane & (G - 1), sub = lane / G, rpw = 32 / G;
    const int nvec = p.C >> 3;
    const float invC = 1.0f / (float)p.C;
    const long long row_step = (long long)gridDim.x * (kLnThreads / 32) * rpw;
    for (long long rb = ((long long)blockIdx.x * (kLnThreads / 32) + warp) * rpw; rb < p.T; rb += row_step) {
        const long long r = rb + sub;
        const bool active = r < p.T;
        const long long rr = active ? r : 0;
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + rr * p.C);
        const uint4* gr = reinterpret_cast<const uint4*>(p.gy + rr * p.C);
        const float mu = __ldg(p.mean + rr), rs = __ldg(p.rstd + rr);
        uint4 cx[NV], cg[NV];
        float s1 = 0.f, s2 = 0.f;
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
        if (VPL > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                cx[i] = active ? __ldcs(xr + gl + i * G) : make_uint4(0, 0, 0, 0);
                cg[i] = active ? __ldcs(gr + gl + i * G) : make_uint4(0, 0, 0, 0);
                accumulate(gl + i * G, cx[i], cg[i]);
            }
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
            if (VPL > 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) emit(gl + i * G, cx[i], cg[i]);
            } else {
                for (int v = gl; v < nvec; v += G) emit(v, __ldg(xr + v), __ldg(gr + v));
            }
        }
    }
}

// Column sums over the rows of [T, C] matrices.  Every thread owns one 8-channel vector (total threads is a multiple
// of nvec = C / 8, so it keeps the same channels while it strides down the rows); threads of a CTA that share a vector
// are summed through shared memory in a fixed order, and each CTA writes one partial row:
//   MODE 0 (LayerNorm parameter gradients): partial[slot][0][c] = sum gy * xhat, partial[slot][1][c] = sum gy
//   MODE 1 (bias gradient of a pointwise convolution):  partial[slot][0][c] = sum gy
// colsum_final_kernel then adds the partial rows (fixed order: deterministic, no atomics).
template <int MODE>
__global__ void __launch_bounds__(kLnThreads) colsum_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ gy,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ partial, long long T, int C)
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
#pragma unroll
        for (int j = 0; j < 8; ++j) db[j] += fg[j];
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
    case 1: ln_fwd_kernel<1><<<grid, kLnThreads, 0, st>>>(p); break;
    case 2: ln_fwd_kernel<2><<<grid, kLnThreads, 0, st>>>(p); break;
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
    return colsum_slots(C, threads) * 2 * C * (long long)sizeof(float);
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
    switch (VPL) {
    case 1: ln_bwd_dx_kernel<1><<<grid, kLnThreads, 0, st>>>(p); break;
    case 2: ln_bwd_dx_kernel<2><<<grid, kLnThreads, 0, st>>>(p); break;
    default: ln_bwd_dx_kernel<0><<<grid, kLnThreads, 0, st>>>(p); break;
    }
    rc = finish_launch("ln_bwd_dx_kernel");
    if (rc) return rc;
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
