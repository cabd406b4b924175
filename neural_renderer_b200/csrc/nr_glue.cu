// nr_glue.cu -- the step either side of the rasterizer: vertices_to_faces (reference
// neural_renderer/vertices_to_faces.py:4-21) as a fused gather (forward) / scatter-add (backward).
//
// The reference forms d loss / d vertex through Chainer's generic get_item backward (a scatter-add of the
// [B,F,3,3] face gradients into [B*Nv,3]); here it is one pass of fp32 vector reductions into the vertex gradient,
// reading grad_faces exactly once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nr_b200.h"
#include "nr_internal.h"

namespace {

__global__ void __launch_bounds__(256) k_v2f_gather(const float* __restrict__ vertices, const int32_t* __restrict__ faces,
                                                    int Nv, long long n_corners_per_item, float* __restrict__ out) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // corner index within the item: f*3 + k
    if (i >= n_corners_per_item) return;
    const int idx = __ldg(faces + (size_t)b * n_corners_per_item + i);
    float x = 0.f, y = 0.f, z = 0.f;
    if ((unsigned)idx < (unsigned)Nv) {
        const float* v = vertices + ((size_t)b * Nv + idx) * 3;
        x = __ldg(v); y = __ldg(v + 1); z = __ldg(v + 2);
    }
    float* o = out + ((size_t)b * n_corners_per_item + i) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

__global__ void __launch_bounds__(256) k_v2f_scatter(const float* __restrict__ grad_faces, const int32_t* __restrict__ faces,
                                                     int Nv, long long n_corners_per_item, float* __restrict__ grad_vertices) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_corners_per_item) return;
    const int idx = __ldg(faces + (size_t)b * n_corners_per_item + i);
    if ((unsigned)idx >= (unsigned)Nv) return;
    const float* g = grad_faces + ((size_t)b * n_corners_per_item + i) * 3;
    const float gx = __ldg(g), gy = __ldg(g + 1), gz = __ldg(g + 2);
    if (gx == 0.f && gy == 0.f && gz == 0.f) return;  // back faces and unseen faces carry exact zeros
    float* v = grad_vertices + ((size_t)b * Nv + idx) * 3;
    atomicAdd(v, gx); atomicAdd(v + 1, gy); atomicAdd(v + 2, gz);
}

}  // namespace

extern "C" int nr_b200_vertices_to_faces(const float* vertices, const int32_t* faces, int32_t B, int32_t Nv, int32_t Nf,
                                         float* out_faces, void* cuda_stream) {
    nr_internal::launch_count() = 0;
    if (!vertices || !faces || !out_faces || B <= 0 || Nv <= 0 || Nf <= 0 || B > 65535) return NR_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const long long n = (long long)Nf * 3;
    {
        nr_internal::LaunchScope ls("k_v2f_gather", stream);
        k_v2f_gather<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, stream>>>(vertices, faces, Nv, n, out_faces);
    }
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}

extern "C" int nr_b200_vertices_to_faces_backward(const float* grad_faces, const int32_t* faces, int32_t B, int32_t Nv,
                                                  int32_t Nf, float* grad_vertices, uint32_t flags, void* cuda_stream) {
    nr_internal::launch_count() = 0;
    if (!grad_faces || !faces || !grad_vertices || B <= 0 || Nv <= 0 || Nf <= 0 || B > 65535) return NR_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    if (!(flags & NR_GRAD_ACCUMULATE) &&
        cudaMemsetAsync(grad_vertices, 0, (size_t)B * Nv * 3 * sizeof(float), stream) != cudaSuccess)
        return NR_ERR_CUDA;
    const long long n = (long long)Nf * 3;
    {
        nr_internal::LaunchScope ls("k_v2f_scatter", stream);
        k_v2f_scatter<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, stream>>>(grad_faces, faces, Nv, n, grad_vertices);
    }
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}
