// Process-wide state of the octm library: error string, launch counter, device properties.
#include "common.cuh"

namespace octm {

thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

int max_optin_smem() {
    static thread_local int cached_dev = -1, cached = 227 * 1024;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace octm

extern "C" int octm_abi_version(void) { return OCTM_ABI_VERSION; }
extern "C" const char* octm_last_error(void) { return octm::g_last_error; }
extern "C" uint64_t octm_launch_count(void) { return octm::g_launches.load(std::memory_order_relaxed); }
