// nr_internal.h -- host-side helpers shared by the translation units of libnr_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "nr_b200.h"
#include "nr_geom.cuh"

namespace nr_internal {
// kernels launched by the last forward/backward call on this thread (nr_b200_last_launch_count)
std::atomic<int>& launch_count();
// optional per-kernel CUDA-event timing on the launching stream (nr_b200_set_profiling / nr_b200_read_profile)
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);

struct LaunchScope {
    cudaStream_t s;
    LaunchScope(const char* name, cudaStream_t stream) : s(stream) { prof_begin(name, s); }
    ~LaunchScope() { prof_end(s); launch_count()++; }
};

// FaceSrc / FaceGrad of a call from its ABI arguments; false = missing pointers for the chosen geometry form
inline bool make_face_src(uint32_t flags, const float* faces, const float* vertices, const int32_t* indices, int F, int Nv,
                          nr::FaceSrc* s) {
    s->F = F; s->Nv = 0; s->faces = nullptr; s->vertices = nullptr; s->idx = nullptr; s->idx_bstride = 0;
    if (flags & NR_FACES_INDEXED) {
        if (!vertices || !indices || Nv <= 0) return false;
        s->vertices = vertices; s->idx = indices; s->Nv = Nv;
        s->idx_bstride = (flags & NR_INDICES_SHARED) ? 0 : (long long)F * 3;
        return true;
    }
    if (!faces) return false;
    s->faces = faces;
    return true;
}
inline bool make_face_grad(uint32_t flags, float* grad_faces, float* grad_vertices, const int32_t* indices, int F, int Nv,
                           nr::FaceGrad* g) {
    g->F = F; g->Nv = 0; g->grad_faces = nullptr; g->grad_vertices = nullptr; g->idx = nullptr; g->idx_bstride = 0;
    if (flags & NR_FACES_INDEXED) {
        if (!grad_vertices || !indices || Nv <= 0) return false;
        g->grad_vertices = grad_vertices; g->idx = indices; g->Nv = Nv;
        g->idx_bstride = (flags & NR_INDICES_SHARED) ? 0 : (long long)F * 3;
        return true;
    }
    if (!grad_faces) return false;
    g->grad_faces = grad_faces;
    return true;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is issued once per (kernel instantiation, device, size high-water
// mark) instead of on every launch: `slot` is a function-local static of the launching template.
struct SmemOptIn {
    static constexpr int kMaxDevices = 64;
    std::atomic<int> bytes[kMaxDevices];
    SmemOptIn() { for (auto& b : bytes) b.store(0, std::memory_order_relaxed); }
    template <typename Kernel>
    cudaError_t ensure(Kernel kernel, size_t need) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        if ((int)need <= bytes[dev].load(std::memory_order_acquire)) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        if (e == cudaSuccess) bytes[dev].store((int)need, std::memory_order_release);
        return e;
    }
};
}  // namespace nr_internal
