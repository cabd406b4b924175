// nr_api.cu -- ABI bookkeeping entry points of libnr_b200.so (include/nr_b200.h).
#include <string.h>

#include <vector>

#include "nr_b200.h"
#include "nr_internal.h"

namespace nr_internal {

namespace {
struct Timer {
    const char* name;
    cudaEvent_t start, stop;
};
thread_local bool g_profiling = false;
thread_local std::vector<Timer> g_timers;
}  // namespace

int& launch_count() {
    static thread_local int n = 0;
    return n;
}

void prof_begin(const char* name, cudaStream_t stream) {
    if (!g_profiling) return;
    Timer t{name, nullptr, nullptr};
    if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return;
    cudaEventRecord(t.start, stream);
    g_timers.push_back(t);
}

void prof_end(cudaStream_t stream) {
    if (!g_profiling || g_timers.empty()) return;
    cudaEventRecord(g_timers.back().stop, stream);
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

extern "C" int nr_b200_last_launch_count(void) { return nr_internal::launch_count(); }

extern "C" void nr_b200_set_profiling(int enabled) { nr_internal::g_profiling = enabled != 0; }

extern "C" int nr_b200_read_profile(char* names, size_t names_bytes, float* ms, int max_entries) {
    using namespace nr_internal;
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
