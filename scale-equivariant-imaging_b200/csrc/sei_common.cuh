// sei_common.cuh -- shared host/device helpers of libsei_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "../../include/sei_b200.h"

namespace sei {

// ------------------------------------------------------------------ host: errors, accounting
void set_error(const char* fmt, ...);
void note_launch(const char* kernel_name);
int finish_launch(const char* kernel_name);   // cudaGetLastError -> return code (+ message)

#define SEI_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            sei::set_error(__VA_ARGS__);            \
            return SEI_EINVAL;                      \
        }                                           \
    } while (0)

#define SEI_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            sei::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                           __FILE__, __LINE__);                                          \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

struct DeviceProps {
    int sm_count;
    int smem_optin;
    int cc_major, cc_minor;
};
int get_device_props(DeviceProps* out);   // cached per device

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// opt a kernel in to `bytes` of dynamic shared memory (idempotent, cheap)
template <class K>
inline cudaError_t allow_smem(K kernel, size_t bytes)
{
    if (bytes <= 40 * 1024) return cudaSuccess;     // static shared memory counts against the 48 KB default as well
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ------------------------------------------------------------------ device: mbarrier + bulk async copy (TMA engine)
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make the barrier initialisation visible to the async proxy before any TMA targets it
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0xF4240;\n\t"   // suspend hint: 1 ms
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// wait for the phase with the given parity; a transfer that never completes (a bug) traps after a
// few seconds instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    for (int it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > 4000) __trap();
}

// generic-proxy writes to smem -> visible to the async proxy (needed before a TMA
// overwrites / reads memory that ordinary stores touched)
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one bulk global->shared copy on the TMA engine (SASS: UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one bulk shared->global copy (SASS: UBLKCP.S.G direction reversed), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// per-thread asynchronous 16-byte copy global -> shared (LDGSTS, bypassing L1); !valid: the 16 bytes are zero-filled
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(valid ? 16 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr uint32_t kBulkChunkBytes = 32768;

// Copy `nrows` consecutive rows of a row-major plane (row pitch = row_bytes, a multiple of
// 16) starting at row `row0`, rows taken modulo H (circular), into contiguous smem.
// Called by ONE thread after mbar_arrive_expect_tx(bar, nrows * row_bytes).
__device__ __forceinline__ void bulk_load_rows_circular(unsigned char* smem_dst, const unsigned char* plane,
                                                        int H, uint32_t row_bytes, int row0, int nrows,
                                                        uint64_t* bar)
{
    int r = row0 % H;
    if (r < 0) r += H;
    int remaining = nrows;
    while (remaining > 0) {
        const int n = min(remaining, H - r);
        const unsigned char* src = plane + (size_t)r * row_bytes;
        uint32_t bytes = (uint32_t)n * row_bytes;
        while (bytes > 0) {
            const uint32_t c = min(bytes, kBulkChunkBytes);
            bulk_g2s(smem_dst, src, c, bar);
            smem_dst += c;
            src += c;
            bytes -= c;
        }
        remaining -= n;
        r = 0;
    }
}

// ------------------------------------------------------------------ device: reductions
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0.  scratch: >= NV*32 doubles
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) scratch[k * 32 + warp] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = lane < nwarps ? scratch[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

__device__ __forceinline__ float4 ld_stream4(const float* p)
{
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
#endif  // __CUDACC__

// ------------------------------------------------------------------ host+device: interpolation math
#ifdef __CUDACC__
#define SEI_HD __host__ __device__ __forceinline__
#else
#define SEI_HD inline
#endif

// Keys cubic convolution coefficients, A = -0.75 (torch grid_sample / upsample_bicubic2d)
SEI_HD void keys_coeffs(float t, float c[4])
{
    const float A = -0.75f;
    float x;
    x = t + 1.0f; c[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    x = t;        c[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 1.0f - t; c[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 2.0f - t; c[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

// anti-aliasing cubic, a = -0.5 (torch _upsample_bicubic2d_aa)
SEI_HD float aa_cubic(float x)
{
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
    if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
    return 0.0f;
}

// smallest n' >= n with n' = 4 (mod 32): consecutive rows of such a pitch start one 16-byte bank group apart
__host__ __device__ constexpr int blur_pad_pitch_c(int n) { return n + ((4 - n % 32) + 32) % 32; }

constexpr int kAaMaxTaps = 16;   // 4 * rate, rate <= 4

// weights of output index i of the antialiased bicubic decimation along one axis
// (ATen _compute_indices_min_size_weights_aa with scale = rate, support = 2 * rate)
SEI_HD void aa_axis_weights(int i, int in_size, int rate, float w[kAaMaxTaps], int& xmin, int& xsize)
{
    const float scale = (float)rate;
    const float support = 2.0f * scale;
    const float invscale = 1.0f / scale;
    const float center = scale * ((float)i + 0.5f);
    int lo = (int)(center - support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5f);
    if (hi > in_size) hi = in_size;
    xmin = lo;
    xsize = hi - lo;
    float total = 0.0f;
#pragma unroll
    for (int j = 0; j < kAaMaxTaps; ++j) {
        float wj = 0.0f;
        if (j < xsize) wj = aa_cubic(((float)(j + lo) - center + 0.5f) * invscale);
        w[j] = wj;
        total += wj;
    }
    if (total != 0.0f) {
#pragma unroll
        for (int j = 0; j < kAaMaxTaps; ++j) w[j] = w[j] / total;
    }
}

// reflect an integer coordinate about [0, S-1] (grid_sample reflection, align_corners=True), then clip
SEI_HD int reflect_clip(int idx, int S)
{
    if (S == 1) return 0;
    const int span = S - 1;
    int v = idx < 0 ? -idx : idx;
    if (v > span) {
        if (v <= 2 * span) {
            v = 2 * span - v;      // one fold: every tap of the reference's rates (>= 0.5) lands here, no division
        } else {
            const int flips = v / span, extra = v - flips * span;
            v = (flips & 1) ? span - extra : extra;
        }
    }
    return v < 0 ? 0 : (v > S - 1 ? S - 1 : v);
}

}  // namespace sei
