/*
 * rtx_b200.h — C-ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary for byvfx/go-raytracing's per-pixel path-tracing loop.
 * The reference has no FFI; the seam is cut at the body of BucketRenderer.renderPass
 * (reference rt/bucket_renderer.go:170-214): everything above it keeps its Go signature,
 * everything below becomes the calls declared here. INTEGRATION.md shows the cgo stub.
 *
 * Conventions
 *   - plain C types only; every call returns int32_t (RTX_OK == 0, < 0 = error enum);
 *   - all host buffers are caller-owned and only borrowed for the duration of the call
 *     (cgo pointer rules: no Go pointer is retained after return);
 *   - a context is NOT thread-safe; the library selects its CUDA device on every call;
 *   - no exceptions / abort cross the boundary; rtx_last_error() gives the message;
 *   - there is NO CPU fallback: with no usable CUDA device rtx_create fails.
 *
 * Geometry arrays are float64 (the reference computes in float64, rt/vec3.go:8-10).
 */
#ifndef RTX_B200_H
#define RTX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTX_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------------------ */
enum {
    RTX_OK = 0,
    RTX_ERR_INVALID = -1,     /* bad argument / malformed description            */
    RTX_ERR_CUDA = -2,        /* CUDA runtime failure (message in rtx_last_error) */
    RTX_ERR_UNSUPPORTED = -3, /* scene uses a construct outside the device path   */
    RTX_ERR_STATE = -4,       /* call order violated (e.g. render before upload)  */
    RTX_ERR_NOMEM = -5
};

/* ---- enums mirrored from the reference's concrete types ---------------------------- */
/* Materials: rt/material.go:33 (Lambertian) :86 (Metal) :146 (Dielectric) :202 (DiffuseLight) :243 (Isotropic) */
enum { RTX_MAT_LAMBERTIAN = 0, RTX_MAT_METAL = 1, RTX_MAT_DIELECTRIC = 2, RTX_MAT_DIFFUSE_LIGHT = 3, RTX_MAT_ISOTROPIC = 4 };
/* Textures: rt/texture.go:9 (SolidColor) :13 (CheckerTexture) */
enum { RTX_TEX_SOLID = 0, RTX_TEX_CHECKER = 1, RTX_TEX_NOISE = 2, RTX_TEX_IMAGE = 3 };
/* Hittables: rt/sphere.go:6, rt/quad.go:5, rt/triangle.go:8, rt/plane.go:5, rt/hittable_list.go:3 (Box = list of 6 quads,
 * rt/primitives.go:5), rt/bvh.go:13 (mesh BVH returned by LoadOBJ, rt/obj_loader.go:109) */
enum { RTX_GEOM_SPHERE = 0, RTX_GEOM_QUAD = 1, RTX_GEOM_TRIANGLE = 2, RTX_GEOM_PLANE = 3, RTX_GEOM_LIST = 4, RTX_GEOM_MESH = 5,
       RTX_GEOM_CIRCLE = 6 /* a primitive like 0-3 (rt/circle.go); numbered after the groups to keep ABI 1 values */ };
/* Instance wrappers: rt/transform.go:78 (Translate) :113 (RotateY) :360 (Scale) */
enum { RTX_XF_TRANSLATE = 0, RTX_XF_ROTATE_Y = 1, RTX_XF_SCALE = 2 };

/* ---- scene description ------------------------------------------------------------- */
/*
 * A flattened copy of the Go scene graph handed to NewBucketRenderer (world Hittable,
 * rt/bucket_renderer.go:54). `entries` are world.Objects in INSERTION order
 * (rt/hittable_list.go:16-19). Each entry is one geometry (a primitive, a HittableList of
 * primitives, or a mesh BVH) optionally wrapped — outermost first — by a chain of
 * Translate / RotateY / Scale instances and optionally by a constant-density Volume
 * (rt/volume.go:10).  Identifiers reported back by rtx_trace_closest are
 * (entry index, primitive index inside the entry's geometry).
 */
typedef struct rtx_scene_desc {
    uint32_t abi_version; /* RTX_ABI_VERSION */
    /* 1 when the reference would traverse NewBVHNodeFromList(world) (main.go:77), 0 when it
     * would traverse the HittableList linearly (Camera.Render). Only affects how exact ties in t
     * are resolved (see entry_rank). */
    int32_t world_is_bvh;

    /* textures */
    int32_t n_textures;
    const int32_t* tex_type;      /* [n_textures] RTX_TEX_*                                   */
    const double* tex_color;      /* [3*n] SolidColor.Albedo (rt/texture.go:10)               */
    const double* tex_inv_scale;  /* [n]   CheckerTexture.invScale = 1/scale (rt/texture.go:49) */
    const int32_t* tex_even;      /* [n]   texture id of .even (checker only)                 */
    const int32_t* tex_odd;       /* [n]   texture id of .odd                                 */

    /* materials */
    int32_t n_materials;
    const int32_t* mat_type;      /* [n] RTX_MAT_*                                            */
    const int32_t* mat_tex;       /* [n] texture id for Lambertian / DiffuseLight / Isotropic; -1 otherwise */
    const double* mat_albedo;     /* [3*n] Metal.Albedo                                       */
    const double* mat_fuzz;       /* [n] Metal.Fuzz (already clamped <= 1, rt/material.go:92) */
    const double* mat_ior;        /* [n] Dielectric.RefractionIndex                           */

    /* primitives */
    int32_t n_spheres;
    const double* sph_center;     /* [3*n] Sphere.Center.orig (rt/sphere.go:7)                */
    const double* sph_velocity;   /* [3*n] Sphere.Center.dir = center2-center1 (rt/sphere.go:27) */
    const double* sph_radius;     /* [n]   raw constructor radius; the library clamps max(0,r) for Hit (rt/sphere.go:18) */
    const int32_t* sph_mat;       /* [n] */

    int32_t n_quads;
    const double* quad_q;         /* [3*n] Quad.Q  */
    const double* quad_u;         /* [3*n] Quad.u  */
    const double* quad_v;         /* [3*n] Quad.v  (normal, D, w are re-derived exactly as rt/quad.go:16-33) */
    const int32_t* quad_mat;      /* [n] */

    int32_t n_tris;
    const double* tri_v0;         /* [3*n] */
    const double* tri_v1;         /* [3*n] */
    const double* tri_v2;         /* [3*n] */
    const int32_t* tri_mat;       /* [n] */
    const int32_t* tri_rank;      /* optional [n]: test-order rank of a mesh triangle inside its mesh's Go BVH
                                     (DFS leaf order of rt/bvh.go:120-217); NULL = library derives a canonical one */

    int32_t n_planes;
    const double* plane_point;    /* [3*n] */
    const double* plane_normal;   /* [3*n] already unit (rt/plane.go:15) */
    const int32_t* plane_mat;     /* [n] */

    /* groups: HittableList of primitives (Box) or triangle mesh */
    int32_t n_groups;
    const int32_t* group_kind;    /* [n] RTX_GEOM_LIST or RTX_GEOM_MESH */
    const int32_t* group_begin;   /* [n] LIST: first index into list_item_*; MESH: first triangle */
    const int32_t* group_count;   /* [n] */
    int32_t n_list_items;
    const int32_t* list_item_kind;  /* [n_list_items] RTX_GEOM_SPHERE..RTX_GEOM_PLANE or RTX_GEOM_CIRCLE */
    const int32_t* list_item_index; /* [n_list_items] index into that primitive array */

    /* instance transform ops, referenced by entries as ranges, OUTERMOST FIRST */
    int32_t n_xforms;
    const int32_t* xf_type;       /* [n] RTX_XF_* */
    const double* xf_a;           /* [3*n] TRANSLATE: Offset; ROTATE_Y: (SinTheta, CosTheta, 0); SCALE: Factor */
    const double* xf_b;           /* [3*n] SCALE: InvFactor (rt/transform.go:368); unused otherwise */

    /* volumes (rt/volume.go:10-15) */
    int32_t n_volumes;
    const double* vol_neg_inv_density; /* [n] */
    const int32_t* vol_mat;            /* [n] Isotropic material id */

    /* world entries, insertion order */
    int32_t n_entries;
    const int32_t* entry_geom_kind;  /* [n] RTX_GEOM_*                                      */
    const int32_t* entry_geom_index; /* [n] primitive index (kinds 0-3, 6) or group index (4,5) */
    const int32_t* entry_xf_begin;   /* [n] */
    const int32_t* entry_xf_count;   /* [n] */
    const int32_t* entry_volume;     /* [n] -1 or volume index: entry is Volume{boundary = this geometry} */
    const int32_t* entry_rank;       /* optional [n]: test-order rank in the Go BVH; NULL = canonical */

    /* Camera.Lights in order (rt/camera.go:38, :502-505). light_quad[i] = index into the quad arrays, or -1
     * when the registered light is not a *Quad (sampleAreaLight returns black, rt/camera.go:616-619). */
    int32_t n_lights;
    const int32_t* light_quad;

    /* HDRI environment (rt/hdri.go:13-26). env_width == 0 means Camera.Environment == nil. */
    int32_t env_width, env_height;
    const double* env_rgb;          /* [3*w*h] decoded linear pixels, row-major, y = 0 top (rt/image_loader.go:374-382) */
    double env_rotation;            /* radians (rt/hdri.go:51) */
    int32_t env_importance_sampling; /* HDRIEnvironment.useImportanceSampling */

    /* ---- ABI 2 ---- */
    /* Circle (rt/circle.go:5-31): a disk. Entries / list items of kind RTX_GEOM_CIRCLE index these arrays. */
    int32_t n_circles;
    const double* circle_center;  /* [3*n] */
    const double* circle_normal;  /* [3*n] already unit (NewCircle normalises, rt/circle.go:16) */
    const double* circle_radius;  /* [n] */
    const int32_t* circle_mat;    /* [n] */
    /* Perlin tables of NoiseTexture (rt/noise.go:8-28). The reference fills them from Go's auto-seeded global source; the
     * caller passes the tables it drew, so that every consumer of the scene sees the same noise. A texture of type
     * RTX_TEX_NOISE uses tex_inv_scale[i] as NoiseTexture.scale (not inverted) and tex_even[i] as its table index. */
    int32_t n_perlin;
    const double* perlin_vec;     /* [n][256][3] unit vectors (Perlin.randvec) */
    const int32_t* perlin_perm;   /* [n][3][256] permX, permY, permZ */
    /* Images of ImageTexture (rt/image_texture.go, rt/image_loader.go:44-73): ImageLoader.data as the reference holds it,
     * i.e. AFTER its load-time LinearToGamma (sqrt) of the 8-bit channels. A texture of type RTX_TEX_IMAGE uses
     * tex_even[i] as its image index; it is looked up with the hit's (u, v) (sphere / quad / triangle / circle). */
    int32_t n_images;
    const int32_t* image_width;   /* [n] */
    const int32_t* image_height;  /* [n] */
    const int64_t* image_offset;  /* [n] first pixel of image i in image_rgb */
    const double* image_rgb;      /* [3 * total pixels] row-major, y = 0 top */
} rtx_scene_desc;

/* Camera public fields (rt/camera.go:18-40). The library re-derives Initialize() (rt/camera.go:286-344)
 * in float64 in the same operation order, unless has_derived != 0, in which case the caller's
 * post-Initialize state is used verbatim (Go's math.Tan may differ from libm in the last ulp). */
typedef struct rtx_camera_desc {
    double aspect_ratio;
    int32_t image_width;
    int32_t samples_per_pixel;
    int32_t max_depth;
    double vfov;
    double look_from[3], look_at[3], vup[3];
    double defocus_angle, focus_dist;
    double look_from2[3], look_at2[3];
    int32_t camera_motion, free_camera;
    double forward[3];
    double background[3];
    int32_t use_sky_gradient;
    int32_t phantom_hdri;
    /* optional post-Initialize state */
    int32_t has_derived;
    int32_t image_height;
    double center[3], pixel00_loc[3], pixel_delta_u[3], pixel_delta_v[3];
    double u[3], v[3], w[3];
    double defocus_radius; /* FocusDist*tan(rad(DefocusAngle/2)), recomputed per ray in rt/camera.go:356 */
    double viewport_width, viewport_height;
} rtx_camera_desc;

/* Counters of the last rtx_render_pass (and cumulative since the last upload). */
typedef struct rtx_stats {
    uint64_t paths;            /* camera samples  == SamplesComputed (rt/bucket_renderer.go:272) */
    uint64_t extension_rays;   /* closest-hit scene queries (rt/camera.go:451)                   */
    uint64_t shadow_rays;      /* NEE scene queries (rt/camera.go:582, :639)                     */
    uint64_t nodes_visited;    /* wide-BVH nodes fetched (extension + shadow)                    */
    uint64_t tri_tests, sphere_tests, quad_tests, plane_tests;
    uint64_t wavefront_iterations;
    uint64_t kernel_launches;  /* kernels of this library launched during the pass               */
    double ms_generate, ms_extend, ms_shade, ms_connect, ms_total; /* CUDA-event times, last pass */
    /* scene structure */
    uint32_t tlas_nodes, blas_nodes, n_entries, n_tris;
    /* last rtx_scene_upload: where the mesh hierarchies were built and how long it took (CUDA events / host clock) */
    uint32_t blas_depth;       /* deepest mesh hierarchy, in 4-wide levels                                   */
    uint32_t bvh_on_device;    /* 1: built by the device builder (replaces NewBVHNodeFromList, rt/bvh.go:64) */
    double ms_bvh_build;       /* device time of all mesh builds (0 with the host builder)                   */
    double ms_scene_upload;    /* host wall time of the whole rtx_scene_upload call                          */
    /* ---- appended in round 2 (the fields above keep their offsets) ---- */
    double ms_tail;            /* last pass: device time from the launch that generated the pass's last camera path to the end
                                  of the pass — the drain iterations in which the stream only shrinks                        */
    double ms_reduce;          /* multi-device contexts: the NCCL sum-reduce of the accumulation buffers after the last pass */
    double ms_resolve;         /* last rtx_resolve_rgba8: kernel + device-to-host copy (CUDA events)                         */
    uint64_t tail_iterations;  /* wavefront iterations of the last pass after its last camera path was generated            */
    uint32_t n_devices;        /* 1, or the device count of an rtx_create_multi context                                      */
    uint32_t checked_build;    /* 1 when the library was compiled with RTX_CHECKED (librtx_b200_checked.so): the trace kernels assert their
                                  slot hand-over protocol and every index at run time (the stand-in for compute-sanitizer)          */
    uint64_t checked_violations;   /* failed assertions of the checked build since the library was loaded; must be 0            */
    uint64_t checked_by_kind[8];   /* by kind, see csrc/rtx_trace.cuh                                                            */
} rtx_stats;

typedef struct rtx_ctx rtx_ctx;

/* ---- lifecycle ---------------------------------------------------------------------- */
/* Replaces: NewBucketRenderer's allocation half (rt/bucket_renderer.go:54-74). One context = one GPU. */
int32_t rtx_create(int32_t device_id, rtx_ctx** out);
/* The same for N GPUs of one box behind ONE context: the reference is one process whose renderPass fans the
 * work out to its worker goroutines (rt/bucket_renderer.go:193-213, numWorkers from main.go:83-93); here renderPass fans the pass
 * out to `n` devices. Every call on the returned context acts on all of them: rtx_scene_upload / rtx_camera_set replicate the
 * scene, rtx_render_pass gives device g the sample slice [sample_base + g*spp/n, sample_base + (g+1)*spp/n) on a host thread of
 * its own and then sums the accumulation buffers onto device_ids[0] with ONE ncclReduce over NVLink (NCCL is loaded with
 * dlopen("libnccl.so.2") here, so single-GPU users never need it); resolve, statistics and the batch entry points read
 * device_ids[0]. n == 1 is rtx_create. Device ids must be distinct. */
int32_t rtx_create_multi(const int32_t* device_ids, int32_t n, rtx_ctx** out);
int32_t rtx_device_count(void);   /* CUDA devices visible to the process (0 when there is none: rtx_create then fails) */
/* Waits for the context's own streams only; a caller-owned stream passed to rtx_set_stream must still be alive or must have been
 * replaced with NULL before it was destroyed. */
int32_t rtx_destroy(rtx_ctx* ctx);
const char* rtx_last_error(const rtx_ctx* ctx); /* ctx may be NULL: last create error of this thread */
int32_t rtx_abi_version(void);

/* Replaces: NewBVHNodeFromList (rt/bvh.go:64) + the pointer graph the Go Hit methods walk.
 * Copies everything; builds the wide BVHs and the HDRI distribution (rt/hdri.go:145-224). */
int32_t rtx_scene_upload(rtx_ctx* ctx, const rtx_scene_desc* scene);
/* Replaces: Camera.Initialize (rt/camera.go:286-344). */
int32_t rtx_camera_set(rtx_ctx* ctx, const rtx_camera_desc* cam);
int32_t rtx_image_size(const rtx_ctx* ctx, int32_t* width, int32_t* height);

/* ---- the hot path -------------------------------------------------------------------- */
/* Replaces: BucketRenderer.renderPass → renderBucketWithQuality → GetRay/RayColor
 * (rt/bucket_renderer.go:170-301, rt/camera.go:368-518). Renders samples
 * [sample_base, sample_base+spp) of every pixel at depth `max_depth` and ADDS their linear
 * radiance into the context's accumulation buffer (so sample slices can be summed across GPUs).
 * `camera_max_depth` is Camera.MaxDepth, needed for the phantom-HDRI primary test
 * `depth == c.MaxDepth` (rt/camera.go:456). Blocking. */
int32_t rtx_render_pass(rtx_ctx* ctx, int32_t spp, int32_t max_depth, int32_t camera_max_depth,
                        uint64_t seed, uint32_t sample_base);
/* Zero the accumulation buffer (start of a pass: each Go pass overwrites the framebuffer, :291-300). */
int32_t rtx_accum_clear(rtx_ctx* ctx);
/* Enable per-pixel sum-of-squares accumulation (level-2 statistical parity). Default off. */
int32_t rtx_accum_enable_moments(rtx_ctx* ctx, int32_t enable);

/* Device pointer of the accumulation buffer: float[4*W*H] = (sum R, sum G, sum B, sample count) per pixel,
 * and (moments on) a second float[4*W*H] of squared sums. For the multi-GPU reduce (NCCL over NVLink):
 * the caller all-reduces / reduces these buffers across ranks in place. */
int32_t rtx_accum_device_ptr(rtx_ctx* ctx, void** sum_dev, void** sumsq_dev, int64_t* n_floats);

/* Replaces: the scale / LinearToGamma / clamp / uint8 pack of renderBucketWithQuality
 * (rt/bucket_renderer.go:275-285, rt/utils.go:85-90) and the framebuffer.Set loop (:291-300).
 * Writes row-major RGBA8, stride 4*W, A = 255 into framebuffer.Pix. total_spp = divisor. */
int32_t rtx_resolve_rgba8(rtx_ctx* ctx, int32_t total_spp, uint8_t* pix, int64_t nbytes);
/* Linear radiance moments: sum_rgb[3*W*H], sumsq_rgb[3*W*H] (may be NULL), n[W*H] (may be NULL). */
int32_t rtx_resolve_accum(rtx_ctx* ctx, float* sum_rgb, float* sumsq_rgb, uint32_t* n);

/* Level-1 parity entry. Replaces: world.Hit(r, Interval{tmin,tmax}, rec) (rt/hittable.go:16) for a batch.
 * rays = n x 7 doubles (origin xyz, direction xyz — NOT normalised —, time). Volumes are transparent here
 * (their Hit draws random numbers, rt/volume.go:66). Outputs (each may be NULL): entry_id / prim_id = -1 on miss;
 * t; normal[3n] (against the ray, rt/hittable.go:20-30); front[n]; uv[2n]; p[3n]. */
int32_t rtx_trace_closest(rtx_ctx* ctx, const double* rays, int64_t n, double tmin, double tmax,
                          int32_t* entry_id, int32_t* prim_id, double* t, double* normal,
                          uint8_t* front, double* uv, double* p);
/* Camera.GetRay for explicit sample parameters (rt/camera.go:368-435): for each k, pixel (ij[2k],ij[2k+1]),
 * square offset sq[2k..], unit-disk point disk[2k..], time tm[k] -> rays_out[7k..]. Runs the device ray-gen code. */
int32_t rtx_camera_rays(rtx_ctx* ctx, const int32_t* ij, const double* sq, const double* disk,
                        const double* tm, int64_t n, double* rays_out);

/* HDRI importance sampling (rt/hdri.go:228-297) on the device for explicit (xi1, xi2):
 * dir[3n], emission[3n], pdf[n]; and PDF(dir) for given directions. For chi-square / KAT tests. */
int32_t rtx_hdri_sample(rtx_ctx* ctx, const double* xi, int64_t n, double* dir, double* emission, double* pdf);
int32_t rtx_hdri_pdf(rtx_ctx* ctx, const double* dir, int64_t n, double* pdf);
int32_t rtx_hdri_lookup(rtx_ctx* ctx, const double* dir, int64_t n, double* rgb); /* Environment.Sample, rt/hdri.go:120 */
int32_t rtx_hdri_total_power(const rtx_ctx* ctx, double* total_power);            /* rt/hdri.go:325 */

/* The test order the last rtx_scene_upload derived ON THE DEVICE for its mesh triangles (desc.tri_rank == NULL, "bvh_device" = 1):
 * rank[t] = position of triangle t (scene order) among the leaves of the tree NewBVHNode would build over its mesh
 * (rt/bvh.go:69-217, stable sort), counted per mesh. n must be the scene's n_tris. RTX_ERR_STATE when the caller gave the ranks
 * itself. Parity tests compare it with the host mirror's tree. */
int32_t rtx_mesh_test_order(rtx_ctx* ctx, int32_t* rank, int64_t n);

/* Run all subsequent work of this context on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores
 * the context's own stream). Lets the caller order the library's kernels with its own (e.g. the NCCL reduce of the
 * accumulation buffers issued by torch.distributed) and time them with events on that stream. */
int32_t rtx_set_stream(rtx_ctx* ctx, void* cuda_stream);

int32_t rtx_get_stats(rtx_ctx* ctx, rtx_stats* out);
/* Tunables: "pool_paths" (upper limit of in-flight path slots), "overlap_connect" (shadow rays on a second stream beside the next
 * wavefront iteration, default 1; the pass is complete on the context's stream when rtx_render_pass returns either way), "count_stats" (per-ray traversal counters: bit 0 extension rays, bit 1 shadow
 * rays), "time_kernels" (CUDA-event time per kernel kind, default 1), "lean" (kernel variants compiled per scene vocabulary, default 1; 0 = all-features kernels), "fuse_tree" (hierarchy worlds shade inside the persistent trace
 * kernel, default 0), "pretest_bare" (bare primitives beside a mesh are tested at pool entry instead of through the TLAS, default 0;
 * takes effect at the next rtx_scene_upload), "tlas_flat_max", "shade_direct", "tri_pretest", "simple_below", "fuse_drain" (DESIGN.md sections 4 and 6), "drop_caches",
 * "pixel_major" (order of the camera paths of a pass: 1 = tiles of 32 neighbouring pixels, all their samples consecutively - the default;
 * 2 = all samples of one pixel consecutively; 0 = sample-major. The image does not depend on it beyond float32 summation order).
 * Returns RTX_ERR_INVALID for unknown keys. */
int32_t rtx_set_option(rtx_ctx* ctx, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* RTX_B200_H */
