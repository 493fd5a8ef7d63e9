// fma_probe.cu -- measures fp32 FMA issue throughput on sm_100a for three instruction forms:
// scalar FFMA with register operands, scalar FFMA with a constant-bank operand, packed FFMA2
// (fma.rn.f32x2).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float c[16]; };

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, const __grid_constant__ Taps t, float seed)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
    float r0 = seed * 0.999f, r1 = seed * 1.001f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], r0, r1);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], t.c[i], r1);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(r0, r0), make_float2(r1, r1));
                a[i] = v.x; a[i + 1] = v.y;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out)
{
    Taps t;
    for (int i = 0; i < 16; ++i) t.c[i] = 0.9f + 0.001f * i;
    const int blocks = 148 * 8, iters = 4096;
    probe<MODE><<<blocks, 256>>>(out, 16, t, 1.0f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, iters, t, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * 256 * iters * 16;
    printf("%-28s %8.3f ms  %7.2f TFMA/s  (%.1f FMA lanes/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA reg,reg,reg", out);
    run<1>("FFMA reg,const,reg", out);
    run<2>("FFMA2 (f32x2)", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
