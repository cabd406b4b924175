// nr_internal.h -- host-side helpers shared by the translation units of libnr_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

namespace nr_internal {
// kernels launched by the last forward/backward call on this thread (nr_b200_last_launch_count)
int& launch_count();
// optional per-kernel CUDA-event timing on the launching stream (nr_b200_set_profiling / nr_b200_read_profile)
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);

struct LaunchScope {
    cudaStream_t s;
    LaunchScope(const char* name, cudaStream_t stream) : s(stream) { prof_begin(name, s); }
    ~LaunchScope() { prof_end(s); launch_count()++; }
};
}  // namespace nr_internal
