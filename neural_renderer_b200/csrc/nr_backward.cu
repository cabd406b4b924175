// nr_backward.cu -- backward pass for sm_100a.
//
// Replaces Rasterize.backward_gpu (reference neural_renderer/rasterize.py:849-889):
//
//   k_edge_scan       K5, the approximate-gradient image scan (rasterize.py:528-748).  The reference runs one thread
//                     per face that walks image columns/rows straight out of global memory.  Here a CTA owns a strip
//                     of W image lines (columns for axis 0, rows for axis 1) of one batch item, stages the strip's
//                     pixels once in shared memory as 32-byte records {I_rgb, dL/dI_rgb, dL/dalpha, face index}
//                     (transposed for axis 0 so that every scan walks contiguous shared memory), culls faces against
//                     the strip with the forward pass's chunk / face boxes, compacts the surviving
//                     (face, edge, line) scan tasks into a shared queue, and runs one task per lane.  Each task
//                     reproduces the reference's discrete decisions exactly (crossing pixel floor/ceil, the
//                     `face_index_map == fn` gates, the in-scan limit) and accumulates the same
//                     -relu(dI . dL/dI) / dist terms; only the summation order differs (fp32 atomics into grad_faces).
//   k_texture_grad    K6 (rasterize.py:760-792): the 8 trilinear weights/indices are recomputed from the saved
//                     weight/depth maps with the forward expression tree instead of being stored (64 B/pixel in the
//                     reference) and scattered with float atomics.
//   k_depth_grad      K7 (rasterize.py:805-847): analytic d zp / d(x, y, z) of the winning face.
//
// Upstream gradients arrive in API layout (planar, image orientation, pooled by 2x2 when anti-aliasing): the
// backward of rasterize_rgbad's transpose / flip / average pooling (rasterize.py:953-969) is folded into the loads.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "nr_b200.h"
#include "nr_bbox.cuh"
#include "nr_internal.h"
#include "nr_math.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kFaceQueue = 512;               // surviving faces per round (9-bit queue slot in a task word)
constexpr int kTaskQueue = 16384;             // (face, edge, line) slots expanded per round
constexpr int kMaxLines = 16;                 // W upper bound: keeps kTaskQueue / (3 W) >= kThreads faces per round
constexpr int kStripBytes = 64 * 1024;        // shared memory budget for the staged strip

struct BwdParams {
    const float* faces;
    const int32_t* fim;
    const float* wmap;
    const float* dmap;
    const float* rgb;
    const float* g_rgb;
    const float* g_alpha;
    const float* g_depth;
    const uint2* bbox;
    const uint2* chunk_bbox;
    float* grad_faces;
    float* grad_textures;
    int B, F, S, ts, nchunks;
    int W;          // lines per strip (power of two)
    int pitch;      // records per staged line (S, +1 padding for axis 0)
    uint32_t flags;
    float eps, two_over_S, tex_cmp, tex_val;
};

// upstream gradient of raster pixel (row, col) of plane `pl` -- folds the 2x2 average-pooling backward
__device__ __forceinline__ float load_grad(const float* g, bool aa, int S, size_t img_plane_index, int row, int col) {
    if (!aa) return __ldg(g + img_plane_index * (size_t)S * S + (size_t)row * S + col);
    const int H = S >> 1;
    return 0.25f * __ldg(g + img_plane_index * (size_t)H * H + (size_t)(row >> 1) * H + (col >> 1));
}

// ------------------------------------------------------------------------------------------------ k_edge_scan
// record layout (8 words): [I0 I1 I2 g0 | g1 g2 fim galpha]; alpha itself is (fim >= 0)
template <bool kRGB, bool kALPHA>
__global__ void __launch_bounds__(kThreads) k_edge_scan(const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* rec = reinterpret_cast<float4*>(smem_raw);  // [W][pitch][2]
    __shared__ int s_faceq[kFaceQueue];
    __shared__ uint16_t s_taskq[kTaskQueue];
    __shared__ int s_nface, s_ntask;

    const int tid = threadIdx.x, lane = tid & 31;
    const int axis = blockIdx.y, b = blockIdx.z;
    const int S = p.S, W = p.W, pitch = p.pitch;
    const int l0 = blockIdx.x * W;
    const int nlines = min(W, S - l0);
    const bool aa = (p.flags & NR_ANTI_ALIASING) != 0;
    const size_t plane = (size_t)S * S;

    // ---- 1. stage the strip (image orientation in global memory: raster row y is stored at row S-1-y)
    for (int i = tid; i < nlines * S; i += kThreads) {
        int line, d1, x, y;
        if (axis == 0) { line = i % nlines; d1 = i / nlines; x = l0 + line; y = d1; }   // columns: d0 = x, d1 = y
        else           { line = i / S;      d1 = i % S;      x = d1;        y = l0 + line; }
        const int row = S - 1 - y;
        const size_t o = (size_t)row * S + x;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int fi = __ldg(p.fim + (size_t)b * plane + o);
        r1.z = __int_as_float(fi);
        if (kRGB) {
            const float* rm = p.rgb + (size_t)b * 3 * plane + o;
            r0.x = __ldg(rm); r0.y = __ldg(rm + plane); r0.z = __ldg(rm + 2 * plane);
            if (p.g_rgb) {
                r0.w = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 0, row, x);
                r1.x = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 1, row, x);
                r1.y = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 2, row, x);
            }
        }
        if (kALPHA && p.g_alpha) r1.w = load_grad(p.g_alpha, aa, S, (size_t)b, row, x);
        rec[((size_t)line * pitch + d1) * 2 + 0] = r0;
        rec[((size_t)line * pitch + d1) * 2 + 1] = r1;
    }
    if (tid == 0) { s_nface = 0; s_ntask = 0; }
    __syncthreads();

    const uint2* bbox = p.bbox + (size_t)b * p.F;
    const uint2* cbox = p.chunk_bbox + (size_t)b * p.nchunks;
    const float fS = (float)S;
    const int lhi = l0 + nlines - 1;

    // per-task geometry, recomputed from the face exactly as rasterize.py:545-575 does
    struct Edge {
        float p00, p01, p10, p11, p20, p21;
        int dir, d0_from, d0_to, pi0, pi1;
    };
    auto edge_setup = [&](int f, int e, Edge& E) {
        const float* v = p.faces + ((size_t)b * p.F + f) * 9;
        const int pi0 = e, pi1 = (e + 1) % 3, pi2 = (e + 2) % 3;
        const int a = axis, c = 1 - axis;
        E.p00 = nr::to_pixel(__ldg(v + 3 * pi0 + a), fS); E.p01 = nr::to_pixel(__ldg(v + 3 * pi0 + c), fS);
        E.p10 = nr::to_pixel(__ldg(v + 3 * pi1 + a), fS); E.p11 = nr::to_pixel(__ldg(v + 3 * pi1 + c), fS);
        E.p20 = nr::to_pixel(__ldg(v + 3 * pi2 + a), fS); E.p21 = nr::to_pixel(__ldg(v + 3 * pi2 + c), fS);
        const bool lt = E.p00 < E.p10;
        E.dir = (axis == 0) ? (lt ? -1 : 1) : (lt ? 1 : -1);
        // (int)max(ceil(min(p0,p1)), 0.) and (int)min(max(p0,p1), is - 1.): truncating conversions (NaN -> 0)
        E.d0_from = __float2int_rz(fmaxf(ceilf(fminf(E.p00, E.p10)), 0.0f));
        E.d0_to = __float2int_rz(fminf(fmaxf(E.p00, E.p10), (float)(S - 1)));
        E.pi0 = pi0; E.pi1 = pi1;
    };

    // faces are queued until the next batch of kThreads could overflow the face queue or the slot expansion
    static_assert(kTaskQueue / (3 * kMaxLines) >= kThreads && kFaceQueue >= 2 * kThreads, "queue sizing");
    const int cap_faces = min(kFaceQueue, kTaskQueue / (3 * nlines));
    int nface = 0;  // uniform across the CTA
    for (int base = 0; base < p.F; base += kThreads) {
        // ---- 2a. cull faces against the strip (chunk box, then face box, on the d0 axis only: scans run to the border)
        const int f = base + tid;
        bool pass = false;
        {
            const uint2 cb = __ldg(cbox + (base / kChunk));  // kThreads == kChunk: one chunk per iteration
            const uint32_t cv = (axis == 0) ? cb.x : cb.y;
            const bool chunk_hit = !(unpack_lo(cv) > lhi || unpack_hi(cv) < l0);
            if (chunk_hit && f < p.F) {
                const uint2 bb = __ldg(bbox + f);
                const uint32_t v = (axis == 0) ? bb.x : bb.y;
                pass = (unpack_lo(bb.x) <= unpack_hi(bb.x)) && !(unpack_lo(v) > lhi || unpack_hi(v) < l0);
            }
        }
        const int cnt = __syncthreads_count(pass);
        const bool last = base + kThreads >= p.F;
        if (cnt) {
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_nface, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (pass) s_faceq[pos + __popc(m & ((1u << lane) - 1u))] = f;
            }
            nface += cnt;
        }
        if (nface == 0 || (!last && nface + kThreads <= cap_faces)) continue;
        __syncthreads();

        // ---- 2b. expand (face, edge, line) slots, keep the ones whose line is inside the edge's span
        for (int i = tid; i < nface * 3 * nlines; i += kThreads) {
            const int line = i % nlines, qe = i / nlines;
            const int e = qe % 3, q = qe / 3;
            Edge E;
            edge_setup(s_faceq[q], e, E);
            const int d0 = l0 + line;
            if (d0 >= E.d0_from && d0 <= E.d0_to) {
                const int t = atomicAdd(&s_ntask, 1);
                s_taskq[t] = (uint16_t)((q << 7) | (e << 5) | line);
            }
        }
        __syncthreads();
        const int ntask = s_ntask;

        // ---- 3. one scan task per lane
        for (int t = tid; t < ntask; t += kThreads) {
            const uint32_t tk = s_taskq[t];
            const int line = tk & 31, e = (tk >> 5) & 3, q = tk >> 7;
            const int fn = s_faceq[q];
            Edge E;
            edge_setup(fn, e, E);
            const int d0 = l0 + line;
            const float fd0 = (float)d0;
            const float slope = __fdiv_rn(__fsub_rn(E.p11, E.p01), __fsub_rn(E.p10, E.p00));
            const float d1_cross = __fmaf_rn(__fsub_rn(fd0, E.p00), slope, E.p01);
            const int d1_in = __float2int_rz(E.dir > 0 ? floorf(d1_cross) : ceilf(d1_cross));
            const int d1_out = d1_in + E.dir;
            if (d1_in < 0 || d1_in >= S || d1_out < 0 || d1_out >= S) continue;
            const float4* lrec = rec + (size_t)line * pitch * 2;
            const float4 in0 = lrec[d1_in * 2], in1 = lrec[d1_in * 2 + 1];
            const float4 out0 = lrec[d1_out * 2], out1 = lrec[d1_out * 2 + 1];
            const float a_in = (__float_as_int(in1.z) >= 0) ? 1.0f : 0.0f;
            const float a_out = (__float_as_int(out1.z) >= 0) ? 1.0f : 0.0f;
            const bool has0 = (E.p10 != fd0), has1 = (E.p00 != fd0);
            const float len = __fsub_rn(E.p10, E.p00);
            const float k0 = __fdiv_rn(len, __fsub_rn(E.p10, fd0)) * p.two_over_S;
            const float k1 = __fdiv_rn(len, __fsub_rn(fd0, E.p00)) * p.two_over_S;
            float acc0 = 0.0f, acc1 = 0.0f;

            auto visit = [&](int d1, float ra, float r0, float r1, float r2) {
                // ra / r0..r2: the reference pixel's alpha / colour (in-pixel for the out-scan, out-pixel for the in-scan)
                const float4 c0 = lrec[d1 * 2], c1 = lrec[d1 * 2 + 1];
                float dg = 0.0f;
                if (kALPHA) {
                    const float a = (__float_as_int(c1.z) >= 0) ? 1.0f : 0.0f;
                    dg = __fmaf_rn(__fsub_rn(a, ra), c1.w, dg);
                }
                if (kRGB) {
                    dg = __fmaf_rn(__fsub_rn(c0.x, r0), c0.w, dg);
                    dg = __fmaf_rn(__fsub_rn(c0.y, r1), c1.x, dg);
                    dg = __fmaf_rn(__fsub_rn(c0.z, r2), c1.y, dg);
                }
                if (dg <= 0.0f) return;
                const float tt = __fsub_rn((float)d1, d1_cross);
                if (has0) {
                    float dist = tt * k0;
                    dist = (0.0f < dist) ? dist + p.eps : dist - p.eps;
                    acc0 -= __fdividef(dg, dist);
                }
                if (has1) {
                    float dist = tt * k1;
                    dist = (0.0f < dist) ? dist + p.eps : dist - p.eps;
                    acc1 -= __fdividef(dg, dist);
                }
            };

            // out-scan: only when the inside pixel shows this face (rasterize.py:604-659)
            if (__float_as_int(in1.z) == fn) {
                const int lim = (E.dir > 0) ? S - 1 : 0;
                const int from = max(min(d1_out, lim), 0), to = min(max(d1_out, lim), S - 1);
                for (int d1 = from; d1 <= to; d1++) visit(d1, a_in, in0.x, in0.y, in0.z);
            }
            // in-scan: towards the opposite edge, pixels that show this face (rasterize.py:662-730)
            {
                float ba, bb, ea, eb;
                if (__fmul_rn(__fsub_rn(fd0, E.p00), __fsub_rn(fd0, E.p20)) < 0.0f) { ba = E.p00; bb = E.p01; ea = E.p20; eb = E.p21; }
                else { ba = E.p20; bb = E.p21; ea = E.p10; eb = E.p11; }
                const float cross2 = __fmaf_rn(__fsub_rn(fd0, ba), __fdiv_rn(__fsub_rn(eb, bb), __fsub_rn(ea, ba)), bb);
                const int lim = __float2int_rz(E.dir > 0 ? ceilf(cross2) : floorf(cross2));
                const int from = max(min(d1_in, lim), 0), to = min(max(d1_in, lim), S - 1);
                for (int d1 = from; d1 <= to; d1++) {
                    if (__float_as_int(lrec[d1 * 2 + 1].z) != fn) continue;
                    visit(d1, a_out, out0.x, out0.y, out0.z);
                }
            }
            float* gf = p.grad_faces + ((size_t)b * p.F + fn) * 9 + (1 - axis);
            if (acc0 != 0.0f) atomicAdd(gf + 3 * E.pi0, acc0);
            if (acc1 != 0.0f) atomicAdd(gf + 3 * E.pi1, acc1);
        }
        __syncthreads();
        if (tid == 0) { s_nface = 0; s_ntask = 0; }
        nface = 0;
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------- k_texture_grad
__global__ void __launch_bounds__(256) k_texture_grad(const __grid_constant__ BwdParams p) {
    const int S = p.S;
    const size_t plane = (size_t)S * S;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel within the image (image orientation)
    const int b = blockIdx.y;
    if (i >= plane) return;
    const int fn = __ldg(p.fim + (size_t)b * plane + i);
    if (fn < 0) return;
    const int row = (int)(i / S), col = (int)(i % S);
    const bool aa = (p.flags & NR_ANTI_ALIASING) != 0;
    const float g0 = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 0, row, col);
    const float g1 = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 1, row, col);
    const float g2 = load_grad(p.g_rgb, aa, S, (size_t)b * 3 + 2, row, col);
    const float* wm = p.wmap + (size_t)b * 3 * plane + i;
    const float w[3] = {__ldg(wm), __ldg(wm + plane), __ldg(wm + 2 * plane)};
    const float zp = __ldg(p.dmap + (size_t)b * plane + i);
    const float* v = p.faces + ((size_t)((p.flags & NR_TEX_Z_BATCH0) ? 0 : b) * p.F + fn) * 9;
    const int ts = p.ts;
    const nr::TexCoord tc = nr::texture_coords(w, zp, __ldg(v + 2), __ldg(v + 5), __ldg(v + 8), ts, p.tex_cmp, p.tex_val);
    float* gt = p.grad_textures + ((size_t)b * p.F + fn) * (size_t)(ts * ts * ts) * 3;
#pragma unroll
    for (int pn = 0; pn < 8; pn++) {
        const float cw = nr::corner_weight(tc, pn);
        float* t = gt + nr::corner_index(tc, pn, ts) * 3;
        atomicAdd(t + 0, cw * g0);
        atomicAdd(t + 1, cw * g1);
        atomicAdd(t + 2, cw * g2);
    }
}

// ----------------------------------------------------------------------------------------------- k_depth_grad
__global__ void __launch_bounds__(256) k_depth_grad(const __grid_constant__ BwdParams p) {
    const int S = p.S;
    const size_t plane = (size_t)S * S;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= plane) return;
    const int fn = __ldg(p.fim + (size_t)b * plane + i);
    if (fn < 0) return;
    const int row = (int)(i / S), col = (int)(i % S);
    const bool aa = (p.flags & NR_ANTI_ALIASING) != 0;
    const float g = load_grad(p.g_depth, aa, S, (size_t)b, row, col);
    const float* v = p.faces + ((size_t)b * p.F + fn) * 9;
    float c[9];
#pragma unroll
    for (int k = 0; k < 9; k++) c[k] = __ldg(v + k);
    const float fS = (float)S;
    float inv[9];
    nr::face_inverse(nr::to_pixel(c[0], fS), nr::to_pixel(c[1], fS), nr::to_pixel(c[3], fS), nr::to_pixel(c[4], fS),
                     nr::to_pixel(c[6], fS), nr::to_pixel(c[7], fS), inv);
    const float* wm = p.wmap + (size_t)b * 3 * plane + i;
    const float w[3] = {__ldg(wm), __ldg(wm + plane), __ldg(wm + 2 * plane)};
    const float depth = __ldg(p.dmap + (size_t)b * plane + i);
    const float depth2 = depth * depth;
    const float z[3] = {c[2], c[5], c[8]};
    float* gf = p.grad_faces + ((size_t)b * p.F + fn) * 9;
    // rasterize.py:824-827  d zp / d z_k = w_k * zp^2 / z_k^2
#pragma unroll
    for (int k = 0; k < 3; k++) atomicAdd(gf + 3 * k + 2, __fdiv_rn((g * w[k]) * depth2, z[k] * z[k]));
    // rasterize.py:830-837  tmp_l = -sum_v inv[v][l] / z_v ;  d zp / d (x,y)_k = -g * tmp_l * w_k * zp^2 * is / 2
    float tmp[2];
#pragma unroll
    for (int l = 0; l < 2; l++)
        tmp[l] = ((0.0f - __fdiv_rn(inv[l], z[0])) - __fdiv_rn(inv[3 + l], z[1])) - __fdiv_rn(inv[6 + l], z[2]);
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 2; l++) atomicAdd(gf + 3 * k + l, (((((-g) * tmp[l]) * w[k]) * depth2) * fS) * 0.5f);
}

inline float float_le(double d) {
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}

template <bool R, bool A>
int launch_edge_scan(const BwdParams& p, int nstrips, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_scan<R, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return NR_ERR_CUDA;
    nr_internal::LaunchScope ls("k_edge_scan", stream);
    k_edge_scan<R, A><<<dim3(nstrips, 2, p.B), kThreads, smem, stream>>>(p);
    return NR_OK;
}

}  // namespace

extern "C" size_t nr_b200_backward_workspace_bytes(int32_t B, int32_t F, int32_t S, int32_t ts, uint32_t flags) {
    (void)S; (void)ts; (void)flags;
    return bbox_workspace_bytes(B, F);
}

extern "C" int nr_b200_backward(const nr_b200_backward_args* a, void* cuda_stream) {
    nr_internal::launch_count() = 0;
    if (!a || a->struct_size != sizeof(nr_b200_backward_args)) return NR_ERR_INVALID_ARG;
    const int B = a->batch_size, F = a->num_faces, S = a->raster_size, ts = a->texture_size;
    const uint32_t flags = a->flags;
    if (B <= 0 || F <= 0 || S <= 0) return NR_ERR_INVALID_ARG;
    if (!a->faces || !a->face_index_map || !a->weight_map || !a->depth_map || !a->grad_faces) return NR_ERR_INVALID_ARG;
    const bool rgb = (flags & NR_RETURN_RGB) != 0, alpha = (flags & NR_RETURN_ALPHA) != 0, depth = (flags & NR_RETURN_DEPTH) != 0;
    if (rgb && (!a->rgb_map || !a->grad_textures || ts < 2)) return NR_ERR_INVALID_ARG;
    if ((flags & NR_ANTI_ALIASING) && (S & 1)) return NR_ERR_INVALID_ARG;
    if (S > 32767 || B > 65535) return NR_ERR_UNSUPPORTED;
    const size_t need = nr_b200_backward_workspace_bytes(B, F, S, ts, flags);
    if (!a->workspace || a->workspace_bytes < need || ((uintptr_t)a->workspace & 15)) return NR_ERR_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    if (!(flags & NR_GRAD_ACCUMULATE)) {
        nr_internal::prof_begin("memset_grads", stream);
        if (cudaMemsetAsync(a->grad_faces, 0, (size_t)B * F * 9 * sizeof(float), stream) != cudaSuccess) return NR_ERR_CUDA;
        if (rgb && cudaMemsetAsync(a->grad_textures, 0, (size_t)B * F * ts * ts * ts * 3 * sizeof(float), stream) != cudaSuccess)
            return NR_ERR_CUDA;
        nr_internal::prof_end(stream);
    }

    BwdParams p{};
    p.faces = a->faces; p.fim = a->face_index_map; p.wmap = a->weight_map; p.dmap = a->depth_map; p.rgb = a->rgb_map;
    p.g_rgb = rgb ? a->grad_rgb : nullptr; p.g_alpha = alpha ? a->grad_alpha : nullptr; p.g_depth = depth ? a->grad_depth : nullptr;
    p.grad_faces = a->grad_faces; p.grad_textures = a->grad_textures;
    p.B = B; p.F = F; p.S = S; p.ts = ts;
    p.flags = flags;
    p.eps = (float)a->eps;
    p.two_over_S = 2.0f / (float)S;
    const double tmax = (double)(ts - 1) - a->eps;
    p.tex_cmp = float_le(tmax);
    p.tex_val = (float)tmax;

    // K5 runs when an rgb or alpha gradient exists (rasterize.py:523); without upstream gradients it contributes 0
    const bool need_scan = (rgb && p.g_rgb) || (alpha && p.g_alpha);
    if (need_scan) {
        const int nchunks = (F + kChunk - 1) / kChunk;
        uint2* bbox = (uint2*)a->workspace;
        uint2* cbox = (uint2*)((char*)a->workspace + nr_align_up((size_t)B * F * sizeof(uint2), 256));
        {
            nr_internal::LaunchScope ls("k_face_bbox", stream);
            k_face_bbox<<<dim3(nchunks, B), kChunk, 0, stream>>>(a->faces, F, S, nchunks, bbox, cbox);
        }
        p.bbox = bbox; p.chunk_bbox = cbox; p.nchunks = nchunks;
        int W = kMaxLines;
        while (W > 1 && (size_t)W * (S + 1) * 32 > (size_t)kStripBytes) W >>= 1;
        p.W = W;
        p.pitch = S + 1;  // +1 record: lines start on different bank groups (transposed stores of axis 0)
        const size_t smem = (size_t)W * p.pitch * 32;
        if (smem > 200 * 1024) return NR_ERR_UNSUPPORTED;
        const int nstrips = (S + W - 1) / W;
        int rc;
        const bool use_rgb = rgb && p.g_rgb, use_alpha = alpha && p.g_alpha;
        if (use_rgb && use_alpha) rc = launch_edge_scan<true, true>(p, nstrips, smem, stream);
        else if (use_rgb) rc = launch_edge_scan<true, false>(p, nstrips, smem, stream);
        else rc = launch_edge_scan<false, true>(p, nstrips, smem, stream);
        if (rc != NR_OK) return rc;
    }
    const dim3 pgrid((unsigned)(((size_t)S * S + 255) / 256), B);
    if (rgb && p.g_rgb) {
        nr_internal::LaunchScope ls("k_texture_grad", stream);
        k_texture_grad<<<pgrid, 256, 0, stream>>>(p);
    }
    if (depth && p.g_depth) {
        nr_internal::LaunchScope ls("k_depth_grad", stream);
        k_depth_grad<<<pgrid, 256, 0, stream>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}
