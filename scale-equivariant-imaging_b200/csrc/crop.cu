// crop.cu -- paired random crops of a batch on the device (reference: src/crop.py CropPair through
// src/datasets/__init__.py:78-90, one torchvision TF.crop pair per dataset item inside the DataLoader loop).
// One launch crops every image of the batch at its own offset: out[b, c, i, j] = in[b, c, top[b] + i, left[b] + j],
// zero where the window leaves the image (TF.crop pads with zeros).  Offsets live in device memory, so the launch is
// graph-capturable and needs no host synchronisation.
#include "sei_common.cuh"
#include <algorithm>

namespace sei {

struct CropParams {
    const float* in;
    float* out;
    const int* top;
    const int* left;
    int C, H, W, h, w;
    long long total;      // B * C * h * w
};

__global__ void __launch_bounds__(256) crop_batch_kernel(const __grid_constant__ CropParams p)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p.w);
        long long t = idx / p.w;
        const int i = (int)(t % p.h);
        t /= p.h;
        const int c = (int)(t % p.C);
        const int b = (int)(t / p.C);
        const int r = __ldg(p.top + b) + i, q = __ldg(p.left + b) + j;
        float v = 0.f;
        if (r >= 0 && r < p.H && q >= 0 && q < p.W) v = __ldg(p.in + (((size_t)b * p.C + c) * p.H + r) * p.W + q);
        p.out[idx] = v;
    }
}

}  // namespace sei

using namespace sei;

extern "C" int sei_crop_batch_f32(const float* in, float* out, int B, int C, int H, int W, int h, int w,
                                  const int* top, const int* left, void* stream)
{
    SEI_REQUIRE(in && out && top && left, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0, "bad shape B=%d C=%d H=%d W=%d h=%d w=%d", B, C, H, W, h, w);
    if (B == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    CropParams p;
    p.in = in; p.out = out; p.top = top; p.left = left; p.C = C; p.H = H; p.W = W; p.h = h; p.w = w;
    p.total = (long long)B * C * h * w;
    const unsigned grid = (unsigned)std::min<long long>((p.total + 255) / 256, (long long)dp.sm_count * 16);
    crop_batch_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return finish_launch("crop_batch_kernel");
}
