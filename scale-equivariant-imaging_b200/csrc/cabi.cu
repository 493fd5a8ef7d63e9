// cabi.cu -- error plumbing, launch accounting and device queries of libsei_b200.
#include "sei_common.cuh"
#include <atomic>
#include <mutex>
#include <string.h>

namespace sei {

static thread_local char g_err[512] = "";
static thread_local char g_last_kernel[128] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void note_launch(const char* kernel_name)
{
    strncpy(g_last_kernel, kernel_name, sizeof(g_last_kernel) - 1);
    g_last_kernel[sizeof(g_last_kernel) - 1] = 0;
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

int finish_launch(const char* kernel_name)
{
    note_launch(kernel_name);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", kernel_name, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int get_device_props(DeviceProps* out)
{
    static std::mutex mu;
    static DeviceProps cache[64];
    static bool have[64] = {false};
    int dev = 0;
    SEI_CUDA(cudaGetDevice(&dev));
    SEI_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        DeviceProps p;
        SEI_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SEI_CUDA(cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SEI_CUDA(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        SEI_CUDA(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        cache[dev] = p;
        have[dev] = true;
    }
    *out = cache[dev];
    return 0;
}

}  // namespace sei

extern "C" {

int sei_abi_version(void) { return SEI_ABI_VERSION; }
const char* sei_last_error(void) { return sei::g_err; }
const char* sei_last_kernel(void) { return sei::g_last_kernel; }
long long sei_launch_count(void) { return sei::g_launches.load(); }

int sei_device_info(int* sm_count, int* smem_per_block_optin, int* cc_major, int* cc_minor)
{
    sei::DeviceProps p;
    int rc = sei::get_device_props(&p);
    if (rc) return rc;
    if (sm_count) *sm_count = p.sm_count;
    if (smem_per_block_optin) *smem_per_block_optin = p.smem_optin;
    if (cc_major) *cc_major = p.cc_major;
    if (cc_minor) *cc_minor = p.cc_minor;
    SEI_REQUIRE(p.cc_major == 10, "libsei_b200 is built for sm_100a only; device is sm_%d%d", p.cc_major, p.cc_minor);
    return 0;
}
}
