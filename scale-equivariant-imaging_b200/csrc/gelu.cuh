// gelu.cuh -- exact (erf-form) GELU pieces shared by the elementwise kernels (cnn_elem.cu) and the GEMM epilogue that
// applies gelu'(h) to an input gradient (gemm.cu).  Phi = 0.5 (1 + erf(x / sqrt 2)), phi = exp(-x^2 / 2) / sqrt(2 pi):
// gelu(x) = x Phi, gelu'(x) = Phi + x phi.  erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7).
#pragma once

namespace sei {

__device__ __forceinline__ void gelu_parts(float x, float& Phi, float& phi)
{
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));     // MUFU.RCP (the rounded reciprocal costs ~8 instructions)
    const float e = __expf(-z * z);                                  // exp(-x^2 / 2)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erf_abs = fmaf(-poly * t, e, 1.0f);                  // erf(|x| / sqrt 2)
    Phi = 0.5f * (1.0f + copysignf(erf_abs, x));
    phi = 0.39894228040143268f * e;
}


}  // namespace sei
