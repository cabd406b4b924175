// nr_api.cu -- ABI bookkeeping entry points of libnr_b200.so (include/nr_b200.h).
#include <string.h>

#include <atomic>
#include <mutex>
#include <deque>

#include "nr_b200.h"
#include "nr_internal.h"

namespace nr_internal {

namespace {
struct Timer {
    const char* name;
    cudaEvent_t start, stop;
};
// Process-wide (not thread-local): PyTorch runs autograd backward on its own engine thread, and the caller that
// enabled profiling / reads the counters is the main thread.  The launch counter is bookkeeping, not control state.
std::atomic<bool> g_profiling{false};
std::mutex g_mu;
std::deque<Timer> g_timers;  // deque: push_back keeps references to earlier records valid
thread_local Timer* g_open = nullptr;
std::atomic<int> g_launches{0};  // written by the autograd thread, read by the caller's thread
}  // namespace

std::atomic<int>& launch_count() { return g_launches; }

void prof_begin(const char* name, cudaStream_t stream) {
    g_open = nullptr;
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    Timer t{name, nullptr, nullptr};
    if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return;
    cudaEventRecord(t.start, stream);
    std::lock_guard<std::mutex> lk(g_mu);
    g_timers.push_back(t);
    g_open = &g_timers.back();
}

void prof_end(cudaStream_t stream) {
    if (!g_open) return;
    cudaEventRecord(g_open->stop, stream);
    g_open = nullptr;
}

}  // namespace nr_internal

extern "C" int nr_b200_abi_version(void) { return NR_B200_ABI_VERSION; }

extern "C" const char* nr_b200_error_string(int code) {
    switch (code) {
        case NR_OK: return "ok";
        case NR_ERR_INVALID_ARG: return "invalid argument (shape, flag combination, missing pointer or struct_size)";
        case NR_ERR_WORKSPACE: return "workspace missing, too small or not 16-byte aligned";
        case NR_ERR_CUDA: return "CUDA launch/runtime error";
        case NR_ERR_UNSUPPORTED: return "size outside the supported range";
        default: return "unknown error code";
    }
}

extern "C" int nr_b200_last_launch_count(void) { return nr_internal::launch_count().load(); }

extern "C" void nr_b200_set_profiling(int enabled) { nr_internal::g_profiling.store(enabled != 0); }

extern "C" int nr_b200_read_profile(char* names, size_t names_bytes, float* ms, int max_entries) {
    using namespace nr_internal;
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    size_t off = 0;
    for (Timer& t : g_timers) {
        float v = -1.0f;
        if (cudaEventSynchronize(t.stop) == cudaSuccess) cudaEventElapsedTime(&v, t.start, t.stop);
        cudaEventDestroy(t.start);
        cudaEventDestroy(t.stop);
        const size_t len = strlen(t.name) + 1;
        if (n < max_entries && names && off + len <= names_bytes) {
            memcpy(names + off, t.name, len);  // NUL-separated list
            off += len;
            ms[n++] = v;
        }
    }
    g_timers.clear();
    return n;
}
