// lfs_b200 -- error plumbing and library-level entry points.
#include "common.cuh"
#include "raster.cuh"
#include "sort_scan.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

namespace lfs {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
    static std::atomic<int> cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

} // namespace lfs

extern "C" const char* lfs_last_error(void) { return lfs::g_err; }
extern "C" int lfs_abi_version(void) { return LFS_ABI_VERSION; }
extern "C" uint64_t lfs_launch_count(void) { return lfs::g_launches.load(std::memory_order_relaxed); }

extern "C" int lfs_set_option(const char* name, int value) {
    if (name && std::string(name) == "fwd_variant") {
        lfs::raster_options().fwd_variant = value;
        return LFS_OK;
    }
    if (name && std::string(name) == "bwd_variant") {
        lfs::raster_options().bwd_variant = value;
        return LFS_OK;
    }
    if (name && std::string(name) == "pre_bwd_split") {
        lfs::raster_options().pre_bwd_split = value;
        return LFS_OK;
    }
    if (name && std::string(name) == "exact_cull") {
        lfs::raster_options().exact_cull = value;
        return LFS_OK;
    }
    if (name && std::string(name) == "fg_variant") {
        lfs::raster_options().fg_variant = value;
        return LFS_OK;
    }
    if (name && std::string(name) == "sort_variant") {
        lfs::set_sort_variant(value);
        return LFS_OK;
    }
    lfs::set_error("set_option: unknown option '%s'", name ? name : "(null)");
    return LFS_ERR_INVALID_ARG;
}
