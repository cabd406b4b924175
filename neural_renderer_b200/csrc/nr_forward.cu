// nr_forward.cu -- forward rasterization for sm_100a.
//
// Replaces Rasterize.forward_gpu (reference neural_renderer/rasterize.py:467-513: K1 :242-277, K2 :281-359,
// K4 :372-438, alpha/background :440-465) and the transpose / flip / 2x2 pooling of rasterize_rgbad (:953-969).
//
// The reference tests every face against every pixel (B*S*S*F face tests).  Here:
//
//   k_face_bbox   one thread per face: back-face / non-finite cull and a conservative pixel bounding box
//                 (8 bytes per face) plus one union box per group of 32 consecutive faces -- the only scratch the
//                 pass needs.
//   k_raster_tile one CTA per 64x32 screen tile (the z-tile lives in shared memory as 64-bit keys
//                 ordered-zp << 32 | face << 10 | record slot):
//                   1. every thread culls one 32-face group box against the tile (the survivors' indices are queued
//                      in shared memory); warps then pull queued groups and cull their faces by face box;
//                   2. survivors (one per lane) get their exact K1 inverse computed once into a per-tile record
//                      table {inv[9], z[3]} and a small sweep record (vertices, clipped box);
//                   3. row-span rasterization: for a fixed pixel row every edge test of the reference,
//                      r_k < (xp - x_k) * dy_k, is monotone in x, so the covered pixels of a row form one interval
//                      whose ends are found by binary search WITH THE REFERENCE'S OWN EXPRESSIONS (identical
//                      coverage, O(log width) tests per row).  Lanes = rows of the group's survivors, flattened
//                      over faces, 32 rows per pass;
//                   4. the pixels of the 32 spans are flattened again (prefix sum) and evaluated 32 fragments at a
//                      time: exact barycentric / perspective-depth expression, then a shared-memory 64-bit min --
//                      lexicographic (zp, fn) minimum == the reference's strict `<` over ascending face index;
//                   5. after one barrier every thread resolves pixels: the winner's record comes from the table
//                      (weights re-evaluated with the same expression tree), its ts^3 texture is sampled (K4), the
//                      background composited, and all maps are streamed out as planar, row-flipped (image
//                      orientation) coalesced rows; with anti-aliasing each thread owns a 2x2 quad and also emits
//                      the pooled API pixel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "nr_b200.h"
#include "nr_math.cuh"
#include "nr_internal.h"
#include "nr_bbox.cuh"

namespace {

// k_raster_tile<kAA, kTL2, kThreads>: tile side 2^kTL2 pixels (32 or 64), kThreads per CTA
constexpr int kRing = 32;      // per-warp scratch of the current group's survivors: x0 y0 x1 y1 | x2 y2 fnrec box
constexpr int kRingWords = 8;
constexpr int kFwdTileLog2Default = 6, kFwdThreadsDefault = 256;
constexpr int kTabMax = 512;   // per-tile table of survivor records {inv[9], z[3]} that fragments and the shade pass share
constexpr int kTabWords = 12;
constexpr int kLiveCap = 1024;  // groups culled against the tile per pass (the survivors' indices are queued in smem)
constexpr uint32_t kNoRec = 1023;  // z-keys carry (face index << 10 | table slot); 1023 = "not in the table"

struct FwdParams {
    nr::FaceSrc src;
    size_t tex_bstride;  // cubes per batch item in `textures` (0 with NR_TEX_SHARED)
    const float* textures;
    const float* bg_batch;
    const float* face_light;
    const uint2* bbox;
    const uint2* group_bbox;
    int32_t* fim;
    float* wmap;
    float* dmap;
    float* rgb;
    float* alpha;
    float* out_rgb;
    float* out_alpha;
    float* out_depth;
    int B, F, S, ts, ngroups;
    int tw_log2, th_log2, tiles_x;
    uint32_t flags;
    float near_lo, far_cmp, far_val, tex_cmp, tex_val;
    float bg[3];
};

// ---------------------------------------------------------------------------------------------- k_raster_tile
template <int kTL2, int kThreads>
struct __align__(16) TileShared {
    static constexpr int kTWL2 = kTL2 == 65 ? 6 : kTL2, kTHL2 = kTL2 == 65 ? 5 : kTL2;
    static constexpr int kTilePix = 1 << (kTWL2 + kTHL2);
    static constexpr int kTab = kTilePix >= 4096 ? kTabMax : kTabMax / 2;
    static constexpr int kWarps = kThreads / 32;
    unsigned long long zbuf[kTilePix];        // (ordered zp bits << 32 | face index << 10 | table slot), ~0 = empty
    float tab[kTab][kTabWords];               // survivor records {inv[9], z0, z1, z2} of this tile
    float ring[kWarps][kRing][kRingWords];    // per-warp sweep records
    float xp[64];
    float yp[64];
    int rowpre[kWarps][32];                   // per-warp: first row number of each survivor of the current group
    int spanpre[kWarps][32];                  // per-warp: first fragment number of each row span of the current pass
    int live[kLiveCap];                       // groups of the current pass whose box overlaps the tile
    int live_count;
    int next_group;
    int tab_count;
};

// {inv[9], z[3]} of face fn of batch item b, straight from global memory (survivors beyond the table's capacity)
__device__ __forceinline__ void face_record(const FwdParams& p, int b, int fn, float inv[9], float z[3]) {
    float c[9];
    nr::load_face(p.src, b, fn, c);
    const float fS = (float)p.S;
    nr::face_inverse(nr::to_pixel(c[0], fS), nr::to_pixel(c[1], fS), nr::to_pixel(c[3], fS), nr::to_pixel(c[4], fS),
                     nr::to_pixel(c[6], fS), nr::to_pixel(c[7], fS), inv);
    z[0] = c[2]; z[1] = c[5]; z[2] = c[8];
}

//@phase shade (resolve helper)
struct Shaded {
    int fim;
    float w0, w1, w2, depth, r, g, b, alpha;
};

__device__ __forceinline__ Shaded shade_pixel(const FwdParams& p, const float (*tab)[kTabWords], int b,
                                              unsigned long long key, int xi, int yi, float bgr, float bgg, float bgb) {
    Shaded o;
    if (key == ~0ull) {
        o.fim = -1; o.w0 = o.w1 = o.w2 = 0.0f; o.depth = p.far_val; o.r = bgr; o.g = bgg; o.b = bgb; o.alpha = 0.0f;
        return o;
    }
    const uint32_t fnrec = (uint32_t)(key & 0xFFFFFFFFull);
    const int fn = (int)(fnrec >> 10);
    const uint32_t rec = fnrec & 1023u;
    const float zp = nr::ordered_to_float((uint32_t)(key >> 32));
    float inv[9], z[3], w[3];
    if (rec != kNoRec) {
        const float4* t4 = reinterpret_cast<const float4*>(tab[rec]);
        const float4 a = t4[0], bb = t4[1], cc = t4[2];
        inv[0] = a.x; inv[1] = a.y; inv[2] = a.z; inv[3] = a.w; inv[4] = bb.x; inv[5] = bb.y; inv[6] = bb.z; inv[7] = bb.w;
        inv[8] = cc.x; z[0] = cc.y; z[1] = cc.z; z[2] = cc.w;
    } else {
        face_record(p, b, fn, inv, z);
    }
    (void)nr::weights_and_depth(inv, (float)xi, (float)yi, z[0], z[1], z[2], w);
    o.fim = fn; o.w0 = w[0]; o.w1 = w[1]; o.w2 = w[2]; o.depth = zp; o.alpha = 1.0f;
    o.r = o.g = o.b = 0.0f;
    if (p.flags & NR_RETURN_RGB) {
        float z0 = z[0], z1 = z[1], z2 = z[2];
        if (p.flags & NR_TEX_Z_BATCH0) {  // rasterize.py:389 -- vertex depths of batch item 0
            z0 = __ldg(nr::face_vertex(p.src, 0, fn, 0) + 2);
            z1 = __ldg(nr::face_vertex(p.src, 0, fn, 1) + 2);
            z2 = __ldg(nr::face_vertex(p.src, 0, fn, 2) + 2);
        }
        const int ts = p.ts;
        const nr::TexCoord tc = nr::texture_coords(w, zp, z0, z1, z2, ts, p.tex_cmp, p.tex_val);
        // NR_TEX_FILL_BACK: the reversed copy of face f - F/2 samples that face's cube with reversed axes
        int cube = fn, ncubes = p.F;
        bool rev = false;
        if (p.flags & NR_TEX_FILL_BACK) {
            ncubes = p.F >> 1;
            if (fn >= ncubes) { cube = fn - ncubes; rev = true; }
        }
        const float* tex = p.textures + ((size_t)b * p.tex_bstride + cube) * (size_t)(ts * ts * ts) * 3;
        float l0 = 1.0f, l1 = 1.0f, l2 = 1.0f;
        const bool lit = p.face_light != nullptr;
        if (lit) {
            const float* lp = p.face_light + ((size_t)b * p.F + fn) * 3;
            l0 = __ldg(lp); l1 = __ldg(lp + 1); l2 = __ldg(lp + 2);
        }
        float r = 0.0f, g = 0.0f, bl = 0.0f;
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            const float cw = nr::corner_weight(tc, pn);
            const float* t = tex + (rev ? nr::corner_index_rev(tc, pn, ts) : nr::corner_index(tc, pn, ts)) * 3;
            float t0 = __ldg(t + 0), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
            if (lit) {  // lighting.py:52 texel * light, rounded like the materialised product
                t0 = __fmul_rn(t0, l0); t1 = __fmul_rn(t1, l1); t2 = __fmul_rn(t2, l2);
            }
            r = __fmaf_rn(cw, t0, r);
            g = __fmaf_rn(cw, t1, g);
            bl = __fmaf_rn(cw, t2, bl);
        }
        o.r = r; o.g = g; o.b = bl;
    }
    return o;
}

//@phase prologue
template <bool kAA, int kTL2, int kThreads>
__global__ void __launch_bounds__(kThreads, kTL2 == 65 ? 1024 / kThreads : 1) k_raster_tile(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Shared = TileShared<kTL2, kThreads>;
    Shared& sm = *reinterpret_cast<Shared*>(smem_raw);
    constexpr int kTab = Shared::kTab;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int tile = blockIdx.x;
    const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
    const int tx0 = (tile % p.tiles_x) << p.tw_log2, ty0 = (tile / p.tiles_x) << p.th_log2;
    const int tx1 = min(tx0 + tw, p.S) - 1, ty1 = min(ty0 + th, p.S) - 1;  // inclusive
    const int npix = tw * th;

    for (int i = tid; i < npix; i += kThreads) sm.zbuf[i] = ~0ull;
    if (tid < 64) {
        // rasterize.py:291-292  xp = (2*xi + 1 - is) / is, evaluated in double and rounded to float
        const double dS = (double)p.S;
        sm.xp[tid] = (float)((double)(2 * (tx0 + tid) + 1 - p.S) / dS);
        sm.yp[tid] = (float)((double)(2 * (ty0 + tid) + 1 - p.S) / dS);
    }
    if (tid == 0) sm.tab_count = 0;

    // ------------------------------------------------------------------ raster phase
    const int ngroups = p.ngroups;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int gbase = 0; gbase < ngroups; gbase += kLiveCap) {
        //@phase group cull
        // ---- group cull, one group (32 consecutive faces) per thread: the groups whose union box overlaps the tile
        //      are queued in shared memory, so the warps below only ever touch faces near the tile
        if (tid == 0) { sm.live_count = 0; sm.next_group = 0; }
        __syncthreads();
        {
            const uint2* gbox = p.group_bbox + (size_t)b * ngroups;
            const int gend = min(gbase + kLiveCap, ngroups);
            for (int g0 = gbase + (warp << 5); g0 < gend; g0 += kThreads) {  // warp-uniform trip count
                const int g = g0 + lane;
                bool hit = false;
                if (g < gend) {
                    const uint2 cb = __ldg(gbox + g);
                    hit = !(unpack_lo(cb.x) > tx1 || unpack_hi(cb.x) < tx0 || unpack_lo(cb.y) > ty1 || unpack_hi(cb.y) < ty0);
                }
                const uint32_t hm = __ballot_sync(0xffffffffu, hit);
                if (hm != 0u) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sm.live_count, __popc(hm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (hit) sm.live[base + __popc(hm & lt_mask)] = g;
                }
            }
        }
        __syncthreads();
        const int nlive = sm.live_count;
        //@phase group pull + face cull + survivor records
        // ---- warp-autonomous: pull live groups
        const uint2* bbox = p.bbox + (size_t)b * p.F;
        float(*ring)[kRingWords] = sm.ring[warp];
        int* rowpre = sm.rowpre[warp];
        int* spanpre = sm.spanpre[warp];
        unsigned long long* zbuf = sm.zbuf;
        const float fS = (float)p.S;

        while (true) {
            int gi = 0;
            if (lane == 0) gi = atomicAdd(&sm.next_group, 1);
            gi = __shfl_sync(0xffffffffu, gi, 0);
            if (gi >= nlive) break;
            const int g = sm.live[gi];
            const int f = (g << 5) + lane;
            bool pass = false;
            int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
            if (f < p.F) {
                const uint2 bb = __ldg(bbox + f);
                bx0 = max(unpack_lo(bb.x), tx0); bx1 = min(unpack_hi(bb.x), tx1);
                by0 = max(unpack_lo(bb.y), ty0); by1 = min(unpack_hi(bb.y), ty1);
                pass = (bx0 <= bx1) && (by0 <= by1);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m == 0u) continue;
            const int nsurv = __popc(m);
            int tbase = 0;
            if (lane == 0) tbase = atomicAdd(&sm.tab_count, nsurv);
            tbase = __shfl_sync(0xffffffffu, tbase, 0);
            if (pass) {
                // survivor set-up, one face per lane: sweep record (vertices, clipped box) into the warp's scratch ring,
                // exact K1 inverse into the tile's record table
                const int rank = __popc(m & lt_mask);
                const int trec = tbase + rank;
                const uint32_t rec = trec < kTab ? (uint32_t)trec : kNoRec;
                float c[9];
                nr::load_face(p.src, b, f, c);
                if (rec != kNoRec) {
                    float inv[9];
                    nr::face_inverse(nr::to_pixel(c[0], fS), nr::to_pixel(c[1], fS), nr::to_pixel(c[3], fS),
                                     nr::to_pixel(c[4], fS), nr::to_pixel(c[6], fS), nr::to_pixel(c[7], fS), inv);
                    float4* t4 = reinterpret_cast<float4*>(sm.tab[rec]);
                    t4[0] = make_float4(inv[0], inv[1], inv[2], inv[3]);
                    t4[1] = make_float4(inv[4], inv[5], inv[6], inv[7]);
                    t4[2] = make_float4(inv[8], c[2], c[5], c[8]);
                }
                float4* r4 = reinterpret_cast<float4*>(ring[rank]);
                r4[0] = make_float4(c[0], c[1], c[3], c[4]);
                const uint32_t box = (uint32_t)(bx0 - tx0) | ((uint32_t)(bx1 - tx0) << 8) | ((uint32_t)(by0 - ty0) << 16) |
                                     ((uint32_t)(by1 - ty0) << 24);
                r4[1] = make_float4(c[6], c[7], __uint_as_float(((uint32_t)f << 10) | rec), __uint_as_float(box));
            }
            //@phase row spans (prefix, owner search, edge binary searches)
            // ---- row-span rasterization.  For a fixed pixel row each edge test  r_k < (xp - x_k) * dy_k  is monotone in
            //      x (xp increases with x; rounded subtraction and multiplication are monotone), so the pixels that
            //      pass all three tests form one interval [lo, hi].  Its ends are found by binary search with the
            //      reference's own expressions -- the coverage is identical, but a row costs O(log width) tests.
            //      Lanes = rows of the group's survivors (flattened over faces), 32 rows per pass; the pixels of the
            //      32 spans are then flattened again and evaluated 32 fragments at a time.
            const int h = pass ? (by1 - by0 + 1) : 0;
            int incl = h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int nrows = __shfl_sync(0xffffffffu, incl, 31);
            rowpre[lane] = incl - h;
            __syncwarp();
            for (int base = 0; base < nrows; base += 32) {
                const int r = base + lane;
                int lo = 1, hi = 0;
                uint32_t fnrec = 0, pix_row = 0;
                if (r < nrows) {
                    // owner = last lane whose first row is <= r  (upper_bound - 1 over the non-decreasing prefix)
                    int a = 0, bnd = 32;
#pragma unroll
                    for (int it = 0; it < 5; it++) {
                        const int mid = (a + bnd) >> 1;
                        if (rowpre[mid] <= r) a = mid; else bnd = mid;
                    }
                    const float4* r4 = reinterpret_cast<const float4*>(ring[__popc(m & ((1u << a) - 1u))]);
                    const float4 q0 = r4[0], q1 = r4[1];
                    const float x0 = q0.x, y0 = q0.y, x1 = q0.z, y1 = q0.w, x2 = q1.x, y2 = q1.y;
                    fnrec = __float_as_uint(q1.z);
                    const uint32_t box = __float_as_uint(q1.w);
                    const int ly = (int)((box >> 16) & 0xFF) + (r - rowpre[a]);
                    pix_row = (uint32_t)(ly << p.tw_log2);
                    const float yp = sm.yp[ly];
                    const float xk[3] = {x0, x1, x2};
                    const float dyk[3] = {__fsub_rn(y1, y0), __fsub_rn(y2, y1), __fsub_rn(y0, y2)};
                    const float rk[3] = {__fmul_rn(__fsub_rn(yp, y0), __fsub_rn(x1, x0)),
                                         __fmul_rn(__fsub_rn(yp, y1), __fsub_rn(x2, x1)),
                                         __fmul_rn(__fsub_rn(yp, y2), __fsub_rn(x0, x2))};
                    lo = box & 0xFF; hi = (box >> 8) & 0xFF;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float xe = xk[k], dy = dyk[k], rr = rk[k];
                        // out(x) = rr < (xp(x) - xe) * dy is non-decreasing in x for dy >= 0 (constant for dy == 0) and
                        // non-increasing for dy < 0: find the first x where out(x) != (dy < 0)
                        const bool neg = dy < 0.0f;
                        int a2 = lo, b2 = hi + 1;
                        while (a2 < b2) {
                            const int mid = (a2 + b2) >> 1;
                            const bool out = rr < __fmul_rn(__fsub_rn(sm.xp[mid], xe), dy);
                            if (out != neg) b2 = mid; else a2 = mid + 1;
                        }
                        if (neg) lo = a2; else hi = a2 - 1;
                    }
                }
                //@phase fragments (flatten, weights + depth, z-min)
                // flatten the 32 spans into fragments
                const int n = max(hi - lo + 1, 0);
                int sincl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, sincl, o);
                    if (lane >= o) sincl += t;
                }
                const int nfrag = __shfl_sync(0xffffffffu, sincl, 31);
                spanpre[lane] = sincl - n;
                __syncwarp();
                for (int fb = 0; fb < nfrag; fb += 32) {
                    const int i = fb + lane;
                    // all lanes take part in the shuffles; lanes past the end evaluate nothing
                    int a = 0, bnd = 32;
#pragma unroll
                    for (int it = 0; it < 5; it++) {
                        const int mid = (a + bnd) >> 1;
                        if (spanpre[mid] <= i) a = mid; else bnd = mid;
                    }
                    const uint32_t o_fnrec = __shfl_sync(0xffffffffu, fnrec, a);
                    const uint32_t o_row = __shfl_sync(0xffffffffu, pix_row, a);
                    const int o_lo = __shfl_sync(0xffffffffu, lo, a);
                    if (i < nfrag) {
                        const int pix = (int)o_row + o_lo + (i - spanpre[a]);
                        const uint32_t rec = o_fnrec & 1023u;
                        float inv[9], z[3];
                        if (rec != kNoRec) {
                            const float4* t4 = reinterpret_cast<const float4*>(sm.tab[rec]);
                            const float4 aa = t4[0], bb = t4[1], cc = t4[2];
                            inv[0] = aa.x; inv[1] = aa.y; inv[2] = aa.z; inv[3] = aa.w; inv[4] = bb.x; inv[5] = bb.y; inv[6] = bb.z;
                            inv[7] = bb.w; inv[8] = cc.x; z[0] = cc.y; z[1] = cc.z; z[2] = cc.w;
                        } else {
                            face_record(p, b, (int)(o_fnrec >> 10), inv, z);  // tile with more than kTab survivors
                        }
                        const int lx = pix & (tw - 1), ly = pix >> p.tw_log2;
                        float w[3];
                        const float zp = nr::weights_and_depth(inv, (float)(tx0 + lx), (float)(ty0 + ly), z[0], z[1], z[2], w);
                        // rasterize.py:331 + :334 against the initial depth_min = far; NaN fails both (never wins)
                        if (zp > p.near_lo && zp < p.far_cmp) {
                            const unsigned long long key = ((unsigned long long)nr::float_to_ordered(zp) << 32) | o_fnrec;
                            unsigned long long* addr = zbuf + pix;
                            if (key < *reinterpret_cast<volatile unsigned long long*>(addr)) atomicMin(addr, key);
                        }
                    }
                }
                __syncwarp();
            }
            __syncwarp();  // the scratch ring / prefix arrays are rewritten by the next group
        }
        __syncthreads();  // every fragment of this pass is in the z-tile; the group queue may be rebuilt
    }

    //@phase resolve + stores
    // ------------------------------------------------------------------ resolve + shade + stream out
    const int S = p.S;
    float bgr = p.bg[0], bgg = p.bg[1], bgb = p.bg[2];
    if (p.flags & NR_BG_PER_BATCH) {
        bgr = __ldg(p.bg_batch + 3 * b + 0); bgg = __ldg(p.bg_batch + 3 * b + 1); bgb = __ldg(p.bg_batch + 3 * b + 2);
    }
    const size_t plane = (size_t)S * S;
    const bool want_rgb = (p.flags & NR_RETURN_RGB) != 0;
    if (!kAA) {
        for (int pix = tid; pix < npix; pix += kThreads) {
            const int lx = pix & (tw - 1), ly = pix >> p.tw_log2;
            const int xi = tx0 + lx, yi = ty0 + ly;
            if (xi >= S || yi >= S) continue;
            const Shaded s = shade_pixel(p, sm.tab, b, sm.zbuf[pix], xi, yi, bgr, bgg, bgb);
            const size_t o = (size_t)b * plane + (size_t)(S - 1 - yi) * S + xi;  // image orientation
            p.fim[o] = s.fim;
            p.dmap[o] = s.depth;
            float* wm = p.wmap + (size_t)b * 3 * plane + (size_t)(S - 1 - yi) * S + xi;
            wm[0] = s.w0; wm[plane] = s.w1; wm[2 * plane] = s.w2;
            if (p.alpha) p.alpha[o] = s.alpha;
            if (want_rgb) {
                float* rm = p.rgb + (size_t)b * 3 * plane + (size_t)(S - 1 - yi) * S + xi;
                rm[0] = s.r; rm[plane] = s.g; rm[2 * plane] = s.b;
            }
        }
    } else {
        const int qw_log2 = p.tw_log2 - 1;
        const int nquad = npix >> 2;
        const int H = S >> 1;
        const size_t oplane = (size_t)H * H;
        for (int q = tid; q < nquad; q += kThreads) {
            const int lx = (q & ((1 << qw_log2) - 1)) << 1, ly = (q >> qw_log2) << 1;
            const int xi = tx0 + lx, yi = ty0 + ly;
            if (xi >= S || yi >= S) continue;  // S is even: quads are entirely in or out
            // the four pixels are shaded one after the other (keeps the register footprint of a single pixel) in
            // image order: top-left, top-right, bottom-left, bottom-right; top row = raster row yi + 1
            float sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f, sd = 0.f;
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                const int dx = k & 1, dy = 1 - (k >> 1);
                const Shaded s = shade_pixel(p, sm.tab, b, sm.zbuf[((ly + dy) << p.tw_log2) + lx + dx], xi + dx, yi + dy, bgr, bgg, bgb);
                const size_t o = (size_t)b * plane + (size_t)(S - 1 - (yi + dy)) * S + xi + dx;
                p.fim[o] = s.fim;
                p.dmap[o] = s.depth;
                float* wm = p.wmap + (size_t)b * 3 * plane + (size_t)(S - 1 - (yi + dy)) * S + xi + dx;
                wm[0] = s.w0; wm[plane] = s.w1; wm[2 * plane] = s.w2;
                if (p.alpha) p.alpha[o] = s.alpha;
                if (want_rgb) {
                    float* rm = p.rgb + (size_t)b * 3 * plane + (size_t)(S - 1 - (yi + dy)) * S + xi + dx;
                    rm[0] = s.r; rm[plane] = s.g; rm[2 * plane] = s.b;
                }
                sr += s.r; sg += s.g; sb += s.b; sa += s.alpha; sd += s.depth;
            }
            const size_t oo = (size_t)(H - 1 - (yi >> 1)) * H + (xi >> 1);
            if (want_rgb && p.out_rgb) {
                float* orgb = p.out_rgb + (size_t)b * 3 * oplane + oo;
                orgb[0] = sr * 0.25f; orgb[oplane] = sg * 0.25f; orgb[2 * oplane] = sb * 0.25f;
            }
            if (p.out_alpha) p.out_alpha[(size_t)b * oplane + oo] = sa * 0.25f;
            if (p.out_depth) p.out_depth[(size_t)b * oplane + oo] = sd * 0.25f;
        }
    }
}

inline float float_le(double d) {  // largest float <= d
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}
inline float float_ge(double d) {  // smallest float >= d
    float f = (float)d;
    if ((double)f < d) f = nextafterf(f, INFINITY);
    return f;
}

}  // namespace

extern "C" size_t nr_b200_forward_workspace_bytes(int32_t B, int32_t F, int32_t S, int32_t ts, uint32_t flags) {
    (void)S; (void)ts; (void)flags;
    return bbox_workspace_bytes(B, F);
}

extern "C" int nr_b200_forward(const nr_b200_forward_args* a, void* cuda_stream) {
    nr_internal::launch_count() = 0;
    if (!a || a->struct_size != sizeof(nr_b200_forward_args)) return NR_ERR_INVALID_ARG;
    const int B = a->batch_size, F = a->num_faces, S = a->raster_size, ts = a->texture_size;
    const uint32_t flags = a->flags;
    if (B <= 0 || F <= 0 || S <= 0) return NR_ERR_INVALID_ARG;
    if (!(flags & (NR_RETURN_RGB | NR_RETURN_ALPHA | NR_RETURN_DEPTH))) return NR_ERR_INVALID_ARG;  // rasterize.py:25-27
    if (!a->face_index_map || !a->weight_map || !a->depth_map) return NR_ERR_INVALID_ARG;
    nr::FaceSrc src{};
    if (!nr_internal::make_face_src(flags, a->faces, a->vertices, a->face_indices, F, a->num_vertices, &src)) return NR_ERR_INVALID_ARG;
    if (flags & NR_RETURN_RGB) {
        if (!a->textures || !a->rgb_map || ts < 2) return NR_ERR_INVALID_ARG;
        if ((flags & NR_TEX_FILL_BACK) && (F & 1)) return NR_ERR_INVALID_ARG;
        if ((flags & NR_BG_PER_BATCH) && !a->background_batch) return NR_ERR_INVALID_ARG;
    }
    if ((flags & NR_ANTI_ALIASING) && (S & 1)) return NR_ERR_INVALID_ARG;
    if (S > 32767 || B > 65535) return NR_ERR_UNSUPPORTED;
    if (F > (1 << 22)) return NR_ERR_UNSUPPORTED;  // z-keys hold a 22-bit face index next to the 10-bit table slot
    const size_t need = nr_b200_forward_workspace_bytes(B, F, S, ts, flags);
    if (!a->workspace || a->workspace_bytes < need || ((uintptr_t)a->workspace & 15)) return NR_ERR_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    const int nchunks = (F + kChunk - 1) / kChunk, ngroups = (F + kGroup - 1) / kGroup;
    uint2* bbox = (uint2*)a->workspace;
    uint2* cbox = (uint2*)((char*)a->workspace + nr_align_up((size_t)B * F * sizeof(uint2), 256));

    {
        nr_internal::LaunchScope ls("k_face_bbox", stream);
        k_face_bbox<<<dim3(nchunks, B), kChunk, 0, stream>>>(src, F, S, ngroups, bbox, cbox);
    }

    FwdParams p{};
    p.src = src;
    p.tex_bstride = (flags & NR_TEX_SHARED) ? 0 : ((flags & NR_TEX_FILL_BACK) ? (size_t)F / 2 : (size_t)F);
    p.textures = a->textures; p.bg_batch = a->background_batch;
    p.face_light = (flags & NR_RETURN_RGB) ? a->face_light : nullptr;
    p.bbox = bbox; p.group_bbox = cbox;
    p.fim = a->face_index_map; p.wmap = a->weight_map; p.dmap = a->depth_map; p.rgb = a->rgb_map; p.alpha = a->alpha_map;
    p.out_rgb = a->out_rgb; p.out_alpha = a->out_alpha; p.out_depth = a->out_depth;
    p.B = B; p.F = F; p.S = S; p.ts = ts; p.ngroups = ngroups;
    int tl = kFwdTileLog2Default, threads = kFwdThreadsDefault;
    bool force_square = false;
#ifdef NR_B200_TUNING  // experiment builds only: NR_B200_FWD_TILE (5|6), NR_B200_FWD_THREADS (128|256)
    if (const char* env = getenv("NR_B200_FWD_TILE")) { tl = atoi(env) <= 5 ? 5 : 6; force_square = atoi(env) == 6; }
    if (const char* env = getenv("NR_B200_FWD_THREADS")) threads = atoi(env) == 128 ? 128 : 256;
#endif
    // shrink the tile for small rasters so that it is not mostly padding (the kernels are compiled for 32 / 64 pixel
    // tiles; smaller rasters run the 32-pixel variant with the tile clipped to the image)
    while (tl > 3 && (1 << (tl - 1)) >= S) tl--;
    tl = tl >= 6 ? 6 : 5;
    // default: 64 wide x 32 tall tiles (4 CTAs of 256 threads per SM)
    const bool wide = (tl == 6) && !force_square;
    p.tw_log2 = tl; p.th_log2 = wide ? 5 : tl;
    p.tiles_x = (S + (1 << tl) - 1) >> tl;
    const int tiles_y = (S + (1 << p.th_log2) - 1) >> p.th_log2;
    p.flags = flags;
    p.near_lo = float_le(a->near_);
    p.far_cmp = fminf(float_ge(a->far_), (float)a->far_);
    p.far_val = (float)a->far_;
    const double tmax = (double)(ts - 1) - a->eps;
    p.tex_cmp = float_le(tmax);
    p.tex_val = (float)tmax;
    p.bg[0] = a->background[0]; p.bg[1] = a->background[1]; p.bg[2] = a->background[2];

    cudaError_t e = cudaSuccess;
    const dim3 grid(p.tiles_x * tiles_y, B);
    const bool aa = (flags & NR_ANTI_ALIASING) != 0;
#define NR_LAUNCH_TILE(AA, TL2, T)                                                                                         \
    do {                                                                                                                   \
        const size_t smem = sizeof(TileShared<TL2, T>);                                                                    \
        static nr_internal::SmemOptIn optin;                                                                               \
        e = optin.ensure(k_raster_tile<AA, TL2, T>, smem);                                                                 \
        if (e != cudaSuccess) return NR_ERR_CUDA;                                                                          \
        nr_internal::LaunchScope ls("k_raster_tile", stream);                                                              \
        k_raster_tile<AA, TL2, T><<<grid, T, smem, stream>>>(p);                                                           \
    } while (0)
    if (wide) {
        if (threads == 128) { if (aa) NR_LAUNCH_TILE(true, 65, 128); else NR_LAUNCH_TILE(false, 65, 128); }
        else                { if (aa) NR_LAUNCH_TILE(true, 65, 256); else NR_LAUNCH_TILE(false, 65, 256); }
    } else if (tl >= 6) {
        if (threads == 128) { if (aa) NR_LAUNCH_TILE(true, 6, 128); else NR_LAUNCH_TILE(false, 6, 128); }
        else                { if (aa) NR_LAUNCH_TILE(true, 6, 256); else NR_LAUNCH_TILE(false, 6, 256); }
    } else {
        if (threads == 128) { if (aa) NR_LAUNCH_TILE(true, 5, 128); else NR_LAUNCH_TILE(false, 5, 128); }
        else                { if (aa) NR_LAUNCH_TILE(true, 5, 256); else NR_LAUNCH_TILE(false, 5, 256); }
    }
#undef NR_LAUNCH_TILE
    e = cudaGetLastError();
    return e == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}
