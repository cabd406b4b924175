// nr_bbox.cuh -- per-face screen bounding boxes shared by the forward and backward passes.
//
// k_face_bbox: one thread per face.  Back faces (rasterize.py:252/:306/:540) and faces with a non-finite x/y (they
// can never win a pixel: their barycentric weights clamp to 0 and zp becomes NaN) get an empty box; every other
// face gets a conservative pixel box (8 bytes) that contains every pixel centre the reference's edge tests can
// accept and every column/row its edge scan can start from.  One union box per group of 32 consecutive faces (a warp
// of this kernel) lets the forward tiles skip whole groups.  This is the only per-face scratch either pass needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nr_geom.cuh"
#include "nr_math.cuh"

namespace {

constexpr int kChunk = 256;  // threads (faces) per CTA of k_face_bbox
constexpr int kGroup = 32;   // faces per group box: the unit the forward tiles cull and pull
constexpr float kBoxMargin = 1.0f / 256.0f;  // pixels; covers fp32 slack of to_pixel and of the edge tests

__device__ __forceinline__ int unpack_lo(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int unpack_hi(uint32_t v) { return (int)(short)(v >> 16); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// Conservative pixel box of a face; false = the face can never win a pixel (back-facing, rasterize.py:252/:306/:540, a
// non-finite x/y, or entirely off screen).
__device__ __forceinline__ bool face_pixel_box(float x0, float y0, float x1, float y1, float x2, float y2, int S, int& xlo,
                                               int& xhi, int& ylo, int& yhi) {
    const bool finite = isfinite(x0) && isfinite(y0) && isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2);
    if (!finite || nr::backside(x0, y0, x1, y1, x2, y2)) return false;
    const float fS = (float)S;
    // to_pixel is monotone in its argument, so the box of the pixel-space vertices is the image of the box
    const float pxmin = nr::to_pixel(fminf(x0, fminf(x1, x2)), fS), pxmax = nr::to_pixel(fmaxf(x0, fmaxf(x1, x2)), fS);
    const float pymin = nr::to_pixel(fminf(y0, fminf(y1, y2)), fS), pymax = nr::to_pixel(fmaxf(y0, fmaxf(y1, y2)), fS);
    const float lim = (float)(S - 1);
    const float fx0 = fmaxf(floorf(pxmin - kBoxMargin), 0.0f), fx1 = fminf(ceilf(pxmax + kBoxMargin), lim);
    const float fy0 = fmaxf(floorf(pymin - kBoxMargin), 0.0f), fy1 = fminf(ceilf(pymax + kBoxMargin), lim);
    if (!(fx0 <= fx1 && fy0 <= fy1)) return false;
    xlo = (int)fx0; xhi = (int)fx1; ylo = (int)fy0; yhi = (int)fy1;
    return true;
}

// ------------------------------------------------------------------------------------------------ k_face_bbox
__global__ void __launch_bounds__(kChunk) k_face_bbox(const nr::FaceSrc src, int F, int S, int ngroups,
                                                      uint2* __restrict__ bbox, uint2* __restrict__ group_bbox) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * kChunk + threadIdx.x;
    int xlo = 1, xhi = 0, ylo = 1, yhi = 0;  // empty
    if (f < F) {
        const float *v0 = nr::face_vertex(src, b, f, 0), *v1 = nr::face_vertex(src, b, f, 1), *v2 = nr::face_vertex(src, b, f, 2);
        const float x0 = __ldg(v0), y0 = __ldg(v0 + 1), x1 = __ldg(v1), y1 = __ldg(v1 + 1), x2 = __ldg(v2), y2 = __ldg(v2 + 1);
        int bx0, bx1, by0, by1;
        if (face_pixel_box(x0, y0, x1, y1, x2, y2, S, bx0, bx1, by0, by1)) { xlo = bx0; xhi = bx1; ylo = by0; yhi = by1; }
        bbox[(size_t)b * F + f] = make_uint2(pack16(xlo, xhi), pack16(ylo, yhi));
    }
    // union box of the warp's 32 faces (empty faces do not contribute; an all-empty group gets an empty box)
    const bool ne = xlo <= xhi;
    int cxlo = ne ? xlo : 32767, cxhi = ne ? xhi : -1, cylo = ne ? ylo : 32767, cyhi = ne ? yhi : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cxlo = min(cxlo, __shfl_xor_sync(0xffffffffu, cxlo, o));
        cxhi = max(cxhi, __shfl_xor_sync(0xffffffffu, cxhi, o));
        cylo = min(cylo, __shfl_xor_sync(0xffffffffu, cylo, o));
        cyhi = max(cyhi, __shfl_xor_sync(0xffffffffu, cyhi, o));
    }
    const int g = (blockIdx.x * kChunk + threadIdx.x) >> 5;
    if ((threadIdx.x & 31) == 0 && g < ngroups)
        group_bbox[(size_t)b * ngroups + g] = make_uint2(pack16(cxlo, cxhi), pack16(cylo, cyhi));
}


inline size_t nr_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// workspace layout used by both passes: [B*F] face boxes, then [B*ngroups] group boxes
inline size_t bbox_workspace_bytes(int B, int F) {
    if (B <= 0 || F <= 0) return 16;
    const size_t ngroups = ((size_t)F + kGroup - 1) / kGroup;
    return nr_align_up((size_t)B * F * sizeof(uint2), 256) + nr_align_up((size_t)B * ngroups * sizeof(uint2), 256);
}

}  // namespace
