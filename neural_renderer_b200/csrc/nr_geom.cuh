// nr_geom.cuh -- where a face's three vertices come from and where their gradients go.
//
// The reference materialises faces [B,F,3,3] with vertices_to_faces (vertices_to_faces.py:16-21: vertices[faces]) in
// front of the rasterizer and scatter-adds the [B,F,3,3] face gradient back through Chainer's get_item backward.
// With NR_FACES_INDEXED both steps are folded into the rasterizer's own loads / atomics: every kernel reads a face
// through FaceSrc (either the materialised tensor or vertices + indices) and accumulates d loss / d vertex through
// FaceGrad (either grad_faces [B,F,3,3] or grad_vertices [B,Nv,3]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nr {

__device__ const float kZeroVertex[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // what an out-of-range index gathers

struct FaceSrc {
    const float* faces;      // [B,F,3,3], or nullptr when indexed
    const float* vertices;   // [B,Nv,3]
    const int32_t* idx;      // [B,F,3] (idx_bstride = 3F) or [F,3] (idx_bstride = 0)
    long long idx_bstride;
    int F, Nv;
};

// pointer to the 3 floats (x, y, z) of vertex k of face f of batch item b; the hot kernels are compiled once per
// geometry form (kIndexed) so that the materialised-faces path carries no trace of the indexed one
template <bool kIndexed>
__device__ __forceinline__ const float* face_vertex_t(const FaceSrc& s, int b, int f, int k) {
    if (!kIndexed) return s.faces + (((size_t)b * s.F + f) * 3 + k) * 3;
    const int i = __ldg(s.idx + (size_t)b * s.idx_bstride + (size_t)f * 3 + k);
    if ((unsigned)i >= (unsigned)s.Nv) return kZeroVertex;
    return s.vertices + ((size_t)b * s.Nv + i) * 3;
}
__device__ __forceinline__ const float* face_vertex(const FaceSrc& s, int b, int f, int k) {
    return s.idx == nullptr ? face_vertex_t<false>(s, b, f, k) : face_vertex_t<true>(s, b, f, k);
}

__device__ __forceinline__ void load_face(const FaceSrc& s, int b, int f, float c[9]) {
    if (s.idx == nullptr) {
        const float* v = s.faces + ((size_t)b * s.F + f) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) c[k] = __ldg(v + k);
    } else {
        const int32_t* ix = s.idx + (size_t)b * s.idx_bstride + (size_t)f * 3;
        const int i0 = __ldg(ix), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
        const float* vb = s.vertices + (size_t)b * s.Nv * 3;
        const float* v0 = (unsigned)i0 < (unsigned)s.Nv ? vb + (size_t)i0 * 3 : kZeroVertex;
        const float* v1 = (unsigned)i1 < (unsigned)s.Nv ? vb + (size_t)i1 * 3 : kZeroVertex;
        const float* v2 = (unsigned)i2 < (unsigned)s.Nv ? vb + (size_t)i2 * 3 : kZeroVertex;
        c[0] = __ldg(v0); c[1] = __ldg(v0 + 1); c[2] = __ldg(v0 + 2);
        c[3] = __ldg(v1); c[4] = __ldg(v1 + 1); c[5] = __ldg(v1 + 2);
        c[6] = __ldg(v2); c[7] = __ldg(v2 + 1); c[8] = __ldg(v2 + 2);
    }
}

struct FaceGrad {
    float* grad_faces;     // [B,F,3,3], or nullptr when indexed
    float* grad_vertices;  // [B,Nv,3]
    const int32_t* idx;
    long long idx_bstride;
    int F, Nv;
};

// where d loss / d (x, y, z) of vertex k of face f accumulates; nullptr for an out-of-range index (skipped, like
// nr_b200_vertices_to_faces_backward)
template <bool kIndexed>
__device__ __forceinline__ float* face_grad_vertex_t(const FaceGrad& g, int b, int f, int k) {
    if (!kIndexed) return g.grad_faces + (((size_t)b * g.F + f) * 3 + k) * 3;
    const int i = __ldg(g.idx + (size_t)b * g.idx_bstride + (size_t)f * 3 + k);
    if ((unsigned)i >= (unsigned)g.Nv) return nullptr;
    return g.grad_vertices + ((size_t)b * g.Nv + i) * 3;
}
__device__ __forceinline__ float* face_grad_vertex(const FaceGrad& g, int b, int f, int k) {
    return g.idx == nullptr ? face_grad_vertex_t<false>(g, b, f, k) : face_grad_vertex_t<true>(g, b, f, k);
}

}  // namespace nr
