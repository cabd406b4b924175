/*
 * nr_b200.h -- C ABI of the B200-native differentiable mesh rasterizer.
 *
 * This is the drop-in boundary for the hot path of hiroharu-kato/neural_renderer:
 * the `Rasterize` function object (neural_renderer/rasterize.py:19-897) and the
 * post-processing owned by `rasterize_rgbad` (rasterize.py:945-969).  The
 * reference has no FFI of its own -- CuPy hands raw device pointers to
 * JIT-compiled kernels (`chainer.cuda.elementwise(...)(arrays)`,
 * rasterize.py:236, :277, :359, :435, :745, :789, :844) -- so the entry points
 * below are what a binding for that path would call instead:
 *
 *   nr_b200_forward    replaces Rasterize.forward_gpu   (rasterize.py:467-513: K1 :242, K2 :281, K4 :372,
 *                      alpha/background :440-465) fused with the transpose / vertical flip / 2x2 average
 *                      pooling of rasterize_rgbad (rasterize.py:953-969)
 *   nr_b200_backward   replaces Rasterize.backward_gpu  (rasterize.py:849-889: K5 :528, K6 :760, K7 :805)
 *                      fused with the backward of that same post-processing
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer is a DEVICE pointer owned by the caller (torch allocator,
 *     cudaMalloc, ...) unless it says "host"; the stream is a cudaStream_t passed as void*; launches are
 *     asynchronous on that stream; nothing is allocated behind the caller's back (scratch = explicit workspace);
 *   - return value 0 = NR_OK, negative = error (nr_b200_error_string); never throws, no global state, re-entrant;
 *   - all images are planar, row-major, in IMAGE orientation: row 0 is the TOP row (the reference's
 *     `[:, ::-1, :]` flip is folded in), i.e. raster row yi (NDC y up) is stored at row S-1-yi;
 *   - S = raster size = image_size, or 2*image_size with NR_ANTI_ALIASING (then the API images are the 2x2
 *     means, size S/2, and the raster-resolution maps are still written because the backward needs them);
 *   - float32 / int32 throughout, C-contiguous.
 */
#ifndef NR_B200_H_
#define NR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NR_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define NR_B200_API __attribute__((visibility("default")))
#else
#define NR_B200_API
#endif

/* error codes */
#define NR_OK 0
#define NR_ERR_INVALID_ARG (-1)
#define NR_ERR_WORKSPACE (-2)
#define NR_ERR_CUDA (-3)
#define NR_ERR_UNSUPPORTED (-4)

/* flags */
#define NR_RETURN_RGB 1u      /* Rasterize(return_rgb=True): needs textures                                   */
#define NR_RETURN_ALPHA 2u    /* Rasterize(return_alpha=True)                                                   */
#define NR_RETURN_DEPTH 4u    /* Rasterize(return_depth=True)                                                   */
#define NR_ANTI_ALIASING 8u   /* rasterize_rgbad(anti_aliasing=True): raster_size = 2 * image_size            */
#define NR_BG_PER_BATCH 16u   /* background_color given as [B,3] (rasterize.py:464-465) in `background_batch`  */
#define NR_TEX_Z_BATCH0 32u   /* reproduce rasterize.py:389: the texture sampler reads vertex depths of batch  */
                              /* item 0 (reference-exact; clear it for per-item depths)                        */
#define NR_GRAD_ACCUMULATE 64u /* backward: add into grad_faces / grad_textures instead of zero-filling first */
#define NR_TEX_FILL_BACK 0x400u /* Renderer.fill_back without materialising the doubled texture tensor            */
                                /* (renderer.py:78-80): F is even, faces [F/2, F) are the reversed copies of      */
                                /* [0, F/2); `textures` / `grad_textures` hold F/2 cubes and face f >= F/2 samples */
                                /* cube f - F/2 with its three texture axes reversed (permute(0,1,4,3,2,5))        */

/* ABI 3 */
#define NR_FACES_INDEXED 0x800u   /* geometry is `vertices` [B,Nv,3] + `face_indices` (vertices_to_faces.py:16-21 folded   */
                                  /* into the rasterizer): `faces` is ignored, no [B,F,3,3] tensor exists on either pass;   */
                                  /* the backward scatters d loss / d vertex straight into `grad_vertices`                  */
#define NR_INDICES_SHARED 0x1000u /* face_indices is [F,3] and serves every batch item (else [B,F,3])                       */
#define NR_TEX_SHARED 0x2000u     /* ONE set of texture cubes [F,ts,ts,ts,3] serves every batch item (a shared mesh seen   */
                                  /* from B viewpoints, mesh.py:29-34); grad_textures [F,...] is the sum over the items     */
#define NR_BWD_PART_TEXTURES 0x4000u /* backward: only the part that produces grad_textures / grad_face_light (K6)         */
#define NR_BWD_PART_FACES 0x8000u    /* backward: only the part that produces grad_faces / grad_vertices (K5, K7)           */
                                     /* neither bit = both parts; a caller that wants to start a collective on the texture   */
                                     /* gradient while the edge scan runs calls TEXTURES first, then FACES                   */

#define NR_FWD_STAGE_TEXTURES 0x10000u /* forward, RGB without anti-aliasing: stage the texture cubes of every pixel row in     */
                                       /* shared memory with bulk asynchronous copies (cp.async.bulk / TMA, one mbarrier per CTA)  */
                                       /* before sampling.  Same pixels; measured SLOWER than the direct gather on B200            */
                                       /* (DESIGN.md section 4), hence opt-in.  Ignored when the cube size is not a multiple of   */
                                       /* 16 bytes or `textures` is not 16-byte aligned.                                           */

typedef struct nr_b200_forward_args {
    uint32_t struct_size; /* sizeof(nr_b200_forward_args), for ABI evolution */
    uint32_t flags;
    int32_t batch_size;   /* B */
    int32_t num_faces;    /* F */
    int32_t raster_size;  /* S (already doubled when NR_ANTI_ALIASING) */
    int32_t texture_size; /* ts (>= 2) when NR_RETURN_RGB, else ignored */
    double near_;         /* reject zp <= near   (compared in double, like the pasted literal, rasterize.py:331) */
    double far_;          /* reject far <= zp; uncovered depth = (float)far (rasterize.py:296, :480)             */
    double eps;           /* texture-coordinate clamp `ts - 1 - eps` (rasterize.py:402)                          */
    float background[3];  /* uniform background colour (host values)                                             */
    float _pad0;
    const float *faces;            /* [B,F,3,3]  x,y in NDC [-1,1], z = camera depth                            */
    const float *textures;         /* [B,F,ts,ts,ts,3] ([F,...] with NR_TEX_SHARED) or NULL                     */
    const float *background_batch; /* [B,3] device, only with NR_BG_PER_BATCH                                   */
    /* raster-resolution maps, saved for the backward pass (all required unless noted) */
    int32_t *face_index_map; /* [B,S,S]   -1 where empty                                                      */
    float *weight_map;       /* [B,3,S,S] barycentric weights of the winning face, 0 where empty              */
    float *depth_map;        /* [B,S,S]   zp, (float)far where empty; IS the depth image when !ANTI_ALIASING  */
    float *rgb_map;          /* [B,3,S,S] post-background colour; required with NR_RETURN_RGB; IS the rgb     */
                             /*           image when !ANTI_ALIASING                                           */
    float *alpha_map;        /* [B,S,S]   0/1; optional (NULL ok); IS the alpha image when !ANTI_ALIASING     */
    /* API images at S/2, only with NR_ANTI_ALIASING (each may be NULL if not wanted) */
    float *out_rgb;   /* [B,3,S/2,S/2] */
    float *out_alpha; /* [B,S/2,S/2]   */
    float *out_depth; /* [B,S/2,S/2]   */
    void *workspace; /* nr_b200_forward_workspace_bytes() bytes, 16-byte aligned */
    size_t workspace_bytes;
    /* ABI 2: per-face RGB light factor of lighting.py:29-52 applied at sample time -- every texel is multiplied by
     * face_light[b,f,:] before the trilinear blend, bit-identical to sampling the materialised `textures * light`
     * product (lighting.py:52).  NULL = unlit. */
    const float *face_light; /* [B,F,3] or NULL */
    /* ABI 3: indexed geometry, only with NR_FACES_INDEXED (then `faces` may be NULL).  Indices outside [0, Nv) read
     * a vertex of zeros, like nr_b200_vertices_to_faces. */
    const float *vertices;       /* [B,Nv,3] x,y in NDC, z = camera depth */
    const int32_t *face_indices; /* [B,F,3], or [F,3] with NR_INDICES_SHARED */
    int32_t num_vertices;        /* Nv */
    int32_t _pad1;
} nr_b200_forward_args;

typedef struct nr_b200_backward_args {
    uint32_t struct_size;
    uint32_t flags; /* same flag set as the forward call that produced the maps */
    int32_t batch_size, num_faces, raster_size, texture_size;
    double eps; /* edge-distance epsilon (rasterize.py:650) and texture clamp epsilon -- the reference uses one value */
    const float *faces;    /* [B,F,3,3] as given to the forward call */
    const float *textures; /* may be NULL: only its shape matters for the backward pass */
    const int32_t *face_index_map; /* saved maps from nr_b200_forward */
    const float *weight_map;
    const float *depth_map;
    const float *rgb_map;
    /* upstream gradients in API layout (size S, or S/2 with NR_ANTI_ALIASING); NULL = zeros */
    const float *grad_rgb;   /* [B,3,H,W] */
    const float *grad_alpha; /* [B,H,W]   */
    const float *grad_depth; /* [B,H,W]   */
    float *grad_faces;    /* [B,F,3,3]                 */
    float *grad_textures; /* [B,F,ts,ts,ts,3] ([B,F/2,...] with NR_TEX_FILL_BACK, no B with NR_TEX_SHARED) or NULL */
    void *workspace;
    size_t workspace_bytes;
    /* ABI 2: lighting folded into the sampler.  grad_textures receives the gradient of the UNLIT textures
     * (weights scaled by face_light); grad_face_light [B,F,3] (may be NULL) receives sum over pixels of
     * grad_rgb * unlit sample and needs `textures`. */
    const float *face_light; /* [B,F,3] as given to the forward call, or NULL */
    float *grad_face_light;  /* [B,F,3] or NULL */
    /* ABI 3: indexed geometry as in the forward call; with NR_FACES_INDEXED the face gradient is reduced into
     * grad_vertices [B,Nv,3] (zero-filled first unless NR_GRAD_ACCUMULATE) and grad_faces may be NULL. */
    const float *vertices;
    const int32_t *face_indices;
    float *grad_vertices; /* [B,Nv,3] */
    int32_t num_vertices;
    int32_t _pad1;
} nr_b200_backward_args;

/* ABI version of the loaded library (== NR_B200_ABI_VERSION it was built with). */
NR_B200_API int nr_b200_abi_version(void);
NR_B200_API const char *nr_b200_error_string(int code);

/* Scratch sizes (bytes).  Pure host arithmetic; safe to call without a GPU. */
NR_B200_API size_t nr_b200_forward_workspace_bytes(int32_t batch_size, int32_t num_faces, int32_t raster_size, int32_t texture_size,
                                       uint32_t flags);
NR_B200_API size_t nr_b200_backward_workspace_bytes(int32_t batch_size, int32_t num_faces, int32_t raster_size,
                                        int32_t texture_size, uint32_t flags);

NR_B200_API int nr_b200_forward(const nr_b200_forward_args *args, void *cuda_stream);
NR_B200_API int nr_b200_backward(const nr_b200_backward_args *args, void *cuda_stream);

/* vertices_to_faces (reference vertices_to_faces.py:4-21), the step either side of the rasterizer:
 *   forward   out_faces[b,f,k,:] = vertices[b, faces[b,f,k], :]           ([B,Nv,3] x [B,Nf,3] int32 -> [B,Nf,3,3])
 *   backward  grad_vertices[b, faces[b,f,k], :] += grad_faces[b,f,k,:]   (zero-filled first unless NR_GRAD_ACCUMULATE)
 * Out-of-range indices gather zeros / are skipped. */
NR_B200_API int nr_b200_vertices_to_faces(const float *vertices, const int32_t *faces, int32_t batch_size,
                                          int32_t num_vertices, int32_t num_faces, float *out_faces, void *cuda_stream);
NR_B200_API int nr_b200_vertices_to_faces_backward(const float *grad_faces, const int32_t *faces, int32_t batch_size,
                                                   int32_t num_vertices, int32_t num_faces, float *grad_vertices,
                                                   uint32_t flags, void *cuda_stream);

/* Camera pipeline of Renderer (reference look_at.py:30-44 / look.py:29-43, then perspective.py:10-18) as one
 * per-vertex kernel each way:
 *   d = vertices[b,v,:] - eye[b];   o = rot[b] * d   (rows of rot = camera x, y, z axes; rot NULL = identity,
 *   eye NULL = origin);   with NR_CAM_PERSPECTIVE:  out = (o.x / o.z / width[b], o.y / o.z / width[b], o.z)
 * rot [B,9], eye [B,3], width [B] are device arrays; with NR_CAM_SHARED they hold ONE camera used by every item.
 * The backward writes grad_vertices [B,Nv,3] (may be NULL) and accumulates, per camera, grad_rot [.,9], grad_eye
 * [.,3], grad_width [.] (each may be NULL; zero-filled first unless NR_GRAD_ACCUMULATE). */
#define NR_CAM_PERSPECTIVE 0x100u
#define NR_CAM_SHARED 0x200u
NR_B200_API int nr_b200_camera_transform(const float *vertices, const float *rot, const float *eye, const float *width,
                                         int32_t batch_size, int32_t num_vertices, uint32_t flags, float *out,
                                         void *cuda_stream);
NR_B200_API int nr_b200_camera_transform_backward(const float *vertices, const float *rot, const float *eye,
                                                  const float *width, const float *grad_out, int32_t batch_size,
                                                  int32_t num_vertices, uint32_t flags, float *grad_vertices,
                                                  float *grad_rot, float *grad_eye, float *grad_width, void *cuda_stream);

/* Per-face light factor of lighting.py:29-51, straight from vertices and face indices:
 *   n = normalize(cross(v0 - v1, v2 - v1))  (x / (|x| + 1e-5), like chainer.functions.normalize)
 *   face_light[b,f,:] = ambient + directional * max(n . direction, 0)
 * `faces` is [B,Nf,3] int32, or [Nf,3] with NR_INDICES_SHARED (ABI 3).
 * light_params [B,9] (or [1,9] with NR_CAM_SHARED) = {intensity_ambient * color_ambient (3),
 * intensity_directional * color_directional (3), direction (3)}, device memory.  The factor is consumed by
 * nr_b200_forward_args.face_light; the backward turns d loss / d face_light (nr_b200_backward_args.grad_face_light)
 * into d loss / d vertices (zero-filled first unless NR_GRAD_ACCUMULATE). */
NR_B200_API int nr_b200_face_lighting(const float *vertices, const int32_t *faces, const float *light_params,
                                      int32_t batch_size, int32_t num_vertices, int32_t num_faces, uint32_t flags,
                                      float *face_light, void *cuda_stream);
NR_B200_API int nr_b200_face_lighting_backward(const float *vertices, const int32_t *faces, const float *light_params,
                                               const float *grad_face_light, int32_t batch_size, int32_t num_vertices,
                                               int32_t num_faces, uint32_t flags, float *grad_vertices, void *cuda_stream);

/* Texture baking of load_obj (reference load_obj.py:88-137): every texel (a, b, c) of the ts^3 cube of face f is the
 * bilinear sample of `image` [H,W,3] (rows already flipped, load_obj.py:82) at the UV position with barycentric
 * coordinates (a, b, c) / (a + b + c) of `uv_faces` [F,3,2]; faces with is_update[f] == 0 (is_update may be NULL =
 * all faces) keep their cube.  Texel (0,0,0) becomes NaN exactly as in the reference (0/0).  `textures`
 * [F,ts,ts,ts,3] is updated in place. */
NR_B200_API int nr_b200_bake_textures(const float *image, const float *uv_faces, const int32_t *is_update,
                                      int32_t num_faces, int32_t texture_size, int32_t image_height,
                                      int32_t image_width, float *textures, void *cuda_stream);

/* Number of kernels the last forward/backward call on this thread launched (for launch accounting). */
NR_B200_API int nr_b200_last_launch_count(void);

/* Optional per-kernel timing: when enabled (per host thread) every kernel launch of forward/backward is bracketed
 * by CUDA events on the launching stream.  nr_b200_read_profile synchronises on them, writes up to max_entries
 * durations (milliseconds) to `ms` and the kernel names as a NUL-separated list to `names`, clears the record and
 * returns the number of entries.  Used by bench.py for the roofline of the dominant kernel; off by default. */
NR_B200_API void nr_b200_set_profiling(int enabled);
NR_B200_API int nr_b200_read_profile(char *names, size_t names_bytes, float *ms, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* NR_B200_H_ */
