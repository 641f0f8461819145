// abi.cu -- version / error / device-info entry points of libmnf_b200.so.
#include "common.cuh"

#include <string.h>

#include <atomic>
#include <mutex>

namespace mnf {

char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

unsigned long long &launch_counter() {
    static unsigned long long n = 0;
    return n;
}

// launch-site tally: sites are string literals, so the pointer is the key (a few dozen sites; linear probe)
struct LaunchSite {
    std::atomic<const char *> name{nullptr};
    std::atomic<unsigned long long> count{0};
};
static LaunchSite g_sites[128];

void note_launch(const char *what) {
    for (auto &s : g_sites) {
        const char *cur = s.name.load(std::memory_order_acquire);
        if (cur == nullptr) {
            const char *expected = nullptr;
            if (s.name.compare_exchange_strong(expected, what, std::memory_order_acq_rel)) cur = what;
            else cur = expected;
        }
        if (cur == what || (cur && strcmp(cur, what) == 0)) {
            s.count.fetch_add(1, std::memory_order_relaxed);
            return;
        }
    }
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

const DeviceProps *device_props() {
    static DeviceProps props[64];
    static bool ready[64] = {false};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!ready[dev]) {
        DeviceProps p;
        if (cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
        if (cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            return nullptr;
        cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
        props[dev] = p;
        ready[dev] = true;
    }
    return &props[dev];
}

}  // namespace mnf

extern "C" {

int mnf_abi_version(void) { return MNF_ABI_VERSION; }

const char *mnf_last_error(void) { return mnf::err_buf(); }

uint64_t mnf_launch_count(void) { return mnf::launch_counter(); }

int64_t mnf_launch_stats(char *buf, int64_t size) {
    int64_t used = 0;
    if (buf && size > 0) buf[0] = 0;
    for (auto &s : mnf::g_sites) {
        const char *name = s.name.load(std::memory_order_acquire);
        if (!name) break;
        const unsigned long long c = s.count.load(std::memory_order_relaxed);
        if (c == 0) continue;
        char line[160];
        const int n = snprintf(line, sizeof line, "%s=%llu;", name, c);
        if (buf && used + n < size) memcpy(buf + used, line, (size_t)n + 1);
        used += n;
    }
    return used;
}

void mnf_launch_stats_reset(void) {
    for (auto &s : mnf::g_sites) s.count.store(0, std::memory_order_relaxed);
}

int mnf_device_info(int *sm_count, int *smem_optin, int *cc_major, int *cc_minor) {
    const mnf::DeviceProps *p = mnf::device_props();
    if (!p) return mnf::fail(MNF_E_DEVICE, "no CUDA device available");
    if (sm_count) *sm_count = p->sm_count;
    if (smem_optin) *smem_optin = p->smem_optin;
    if (cc_major) *cc_major = p->cc_major;
    if (cc_minor) *cc_minor = p->cc_minor;
    return 0;
}

}  // extern "C"
