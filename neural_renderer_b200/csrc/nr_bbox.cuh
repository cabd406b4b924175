// nr_bbox.cuh -- per-face screen bounding boxes shared by the forward and backward passes.
//
// face_pixel_box: back faces (rasterize.py:252/:306/:540) and faces with a non-finite x/y (they can never win a pixel:
// their barycentric weights clamp to 0 and zp becomes NaN) get no box; every other face gets a conservative pixel box
// that contains every pixel centre the reference's edge tests can accept (forward: <true>, with the thin-face margin)
// or every column / row its edge scan can start from (backward: <false>).  The forward evaluates it per face inside
// k_raster_faces, the backward inside the counting pass of k_strip_bin (8 bytes per face in the workspace).
// k_face_bbox is the stand-alone kernel of the same box, used only by the global-atomics binning path of rasters with
// more than 2048 strips per axis.
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

// How far outside its vertex box a face can still pass the reference's three edge tests (rasterize.py:309-311).
// Each test compares two fp32 products of fp32 differences (three roundings a side), so a pixel centre p is accepted
// by edge k whenever its exact signed distance to that edge's line is >= -delta, delta <= 4.4 u D in the rasterizer's
// input coordinates (u = 2^-24, D = sqrt(2) + max |vertex| bounds |p - v_k|).  The accepted set is therefore inside
// the triangle grown by delta along every edge normal, whose corners sit delta / sin(theta_v / 2) <= 2 delta L^2 / |det|
// beyond the vertices (theta_v the interior angle, L the longest edge, det twice the area).  For an ordinary face that
// is 1e-6 of a pixel; for a needle whose long edges meet at 1e-4 rad it is more than a pixel, and such a needle DOES win
// pixels beyond its tip in the reference (tests/test_gpu_parity.py::test_needle_faces).  The margin below carries a
// 3.6x safety factor for the rounding of L^2 / |det| itself and is capped at kThinMarginCap pixels: thinner than about
// 7e-5 rad (or exactly collinear, which never wins: every weight is NaN) the bound is not attained any more.
constexpr float kThinMarginCap = 8.0f;

__device__ __forceinline__ float thin_face_margin(float x0, float y0, float x1, float y1, float x2, float y2, float fS) {
    const float ex0 = __fsub_rn(x1, x0), ey0 = __fsub_rn(y1, y0), ex1 = __fsub_rn(x2, x1), ey1 = __fsub_rn(y2, y1);
    const float ex2 = __fsub_rn(x0, x2), ey2 = __fsub_rn(y0, y2);
    const float l2 = fmaxf(__fmaf_rn(ex0, ex0, ey0 * ey0), fmaxf(__fmaf_rn(ex1, ex1, ey1 * ey1), __fmaf_rn(ex2, ex2, ey2 * ey2)));
    const float det = fabsf(__fmaf_rn(ex0, ey1, -(ey0 * ex1)));
    const float m = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fmaxf(fabsf(x2), fabsf(y2))));
    const float D = 1.4143f * (1.0f + m);
    const float ext = __fdividef(16.0f * 5.9604645e-8f * D * fS * l2, det);  // pixels; inf / NaN for a collinear face
    return (ext < kThinMarginCap) ? ext : kThinMarginCap;
}

// Conservative pixel box of a face; false = the face can never win a pixel (back-facing, rasterize.py:252/:306/:540, a
// non-finite x/y, or entirely off screen).  kCoverage: the box must hold every pixel the forward's edge tests can
// accept (adds the thin-face margin); the backward only needs the columns / rows its edge scan can start from.
template <bool kCoverage = false>
__device__ __forceinline__ bool face_pixel_box(float x0, float y0, float x1, float y1, float x2, float y2, int S, int& xlo,
                                               int& xhi, int& ylo, int& yhi) {
    const bool finite = isfinite(x0) && isfinite(y0) && isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2);
    if (!finite || nr::backside(x0, y0, x1, y1, x2, y2)) return false;
    const float fS = (float)S;
    // to_pixel is monotone in its argument, so the box of the pixel-space vertices is the image of the box
    const float pxmin = nr::to_pixel(fminf(x0, fminf(x1, x2)), fS), pxmax = nr::to_pixel(fmaxf(x0, fmaxf(x1, x2)), fS);
    const float pymin = nr::to_pixel(fminf(y0, fminf(y1, y2)), fS), pymax = nr::to_pixel(fmaxf(y0, fmaxf(y1, y2)), fS);
    const float lim = (float)(S - 1);
    const float margin = kCoverage ? kBoxMargin + thin_face_margin(x0, y0, x1, y1, x2, y2, fS) : kBoxMargin;
    const float fx0 = fmaxf(floorf(pxmin - margin), 0.0f), fx1 = fminf(ceilf(pxmax + margin), lim);
    const float fy0 = fmaxf(floorf(pymin - margin), 0.0f), fy1 = fminf(ceilf(pymax + margin), lim);
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
