// optim.cu -- the optimizer step of the training loop (reference: demo/train.py:11,167-186,266: torch.optim.Adam,
// optimizer.step() once per batch) as one streaming kernel per parameter tensor.
//
// torch's multi-tensor Adam ran at 28 % of the HBM roofline on the 645 M-parameter network and was followed, every
// step, by a separate fp32 -> bf16 cast of every GEMM weight and a strided transposed copy for the input-gradient
// GEMMs.  Here one pass reads p, g, m, v and writes p, m, v (28 B per parameter) and, in the same pass, the bf16
// copy the tensor-core GEMMs read; a tiled shared-memory transpose produces the (K, N) copy for dgrad.
//   m <- m + (1 - b1) (g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (torch.optim.Adam without weight decay / amsgrad; t is read from device memory so the launch is graph-capturable).
#include "sei_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace sei {

struct AdamParams {
    float* p;
    const float* g;
    float* m;
    float* v;
    __nv_bfloat16* lowp;       // optional bf16 copy of p
    const float* step;         // device scalar: the step count t (already incremented)
    long long n;
    float lr, b1, b2, eps;
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamParams a)
{
    const float t = __ldg(a.step);
    const float bc1 = 1.0f - powf(a.b1, t), bc2 = 1.0f - powf(a.b2, t);
    const float step_size = a.lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
    const float w1 = 1.0f - a.b1, w2 = 1.0f - a.b2;
    const long long nvec = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 g = __ldcs(reinterpret_cast<const float4*>(a.g) + i);
        float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
        m.x = fmaf(w1, g.x - m.x, m.x); m.y = fmaf(w1, g.y - m.y, m.y); m.z = fmaf(w1, g.z - m.z, m.z); m.w = fmaf(w1, g.w - m.w, m.w);
        v.x = fmaf(w2, g.x * g.x, a.b2 * v.x); v.y = fmaf(w2, g.y * g.y, a.b2 * v.y);
        v.z = fmaf(w2, g.z * g.z, a.b2 * v.z); v.w = fmaf(w2, g.w * g.w, a.b2 * v.w);
        p.x -= step_size * (m.x / (sqrtf(v.x) * inv_bc2_sqrt + a.eps)); p.y -= step_size * (m.y / (sqrtf(v.y) * inv_bc2_sqrt + a.eps));
        p.z -= step_size * (m.z / (sqrtf(v.z) * inv_bc2_sqrt + a.eps)); p.w -= step_size * (m.w / (sqrtf(v.w) * inv_bc2_sqrt + a.eps));
        reinterpret_cast<float4*>(a.p)[i] = p;
        reinterpret_cast<float4*>(a.m)[i] = m;
        reinterpret_cast<float4*>(a.v)[i] = v;
        if (a.lowp) {
            uint2 o;
            *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(p.x, p.y);
            *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(p.z, p.w);
            reinterpret_cast<uint2*>(a.lowp)[i] = o;
        }
    }
    // tail (n % 4 elements)
    const long long tail0 = nvec << 2;
    const long long gi = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi < a.n) {
        const float g = a.g[gi];
        float m = a.m[gi], v = a.v[gi], p = a.p[gi];
        m = fmaf(w1, g - m, m);
        v = fmaf(w2, g * g, a.b2 * v);
        p -= step_size * (m / (sqrtf(v) * inv_bc2_sqrt + a.eps));
        a.p[gi] = p; a.m[gi] = m; a.v[gi] = v;
        if (a.lowp) a.lowp[gi] = __float2bfloat16_rn(p);
    }
}

// out[c][r] = in[r][c], bf16, through a padded 64 x 64 shared-memory tile (both sides coalesced)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                             int rows, int cols)
{
    __shared__ __nv_bfloat16 tile[64][66];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
    for (int i = ty; i < 64; i += 4)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = in[(size_t)(r0 + i) * cols + c0 + tx];
    __syncthreads();
    for (int i = ty; i < 64; i += 4)
        if (c0 + i < cols && r0 + tx < rows) out[(size_t)(c0 + i) * rows + r0 + tx] = tile[tx][i];
}

}  // namespace sei

using namespace sei;

extern "C" int sei_adam_step_f32(float* p, const float* g, float* m, float* v, void* lowp_bf16, const float* step,
                                 long long n, float lr, float beta1, float beta2, float eps, void* stream)
{
    SEI_REQUIRE(p && g && m && v && step, "null pointer argument");
    SEI_REQUIRE(n >= 0, "bad element count %lld", n);
    SEI_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "p, g, m, v must be 16-byte aligned");
    SEI_REQUIRE(!lowp_bf16 || (reinterpret_cast<uintptr_t>(lowp_bf16) & 7u) == 0, "the bf16 copy must be 8-byte aligned");
    if (n == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    AdamParams a;
    a.p = p; a.g = g; a.m = m; a.v = v; a.lowp = static_cast<__nv_bfloat16*>(lowp_bf16); a.step = step; a.n = n;
    a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
    const long long nvec = std::max<long long>(1, n / 4);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nvec + 255) / 256, (long long)dp.sm_count * 16));
    adam_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    return finish_launch("adam_kernel");
}

extern "C" int sei_transpose_bf16(const void* in, void* out, int rows, int cols, void* stream)
{
    SEI_REQUIRE(in && out, "null pointer argument");
    SEI_REQUIRE(rows > 0 && cols > 0, "bad shape %d x %d", rows, cols);
    dim3 grid((cols + 63) / 64, (rows + 63) / 64);
    SEI_REQUIRE(grid.y <= 65535, "too many rows");
    transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), rows, cols);
    return finish_launch("transpose_bf16_kernel");
}
