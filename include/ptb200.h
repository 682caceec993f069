/* ptb200.h — C ABI of the B200-native radiance loop for maurock/small-pathtracer.
 *
 * The reference has no FFI: its "API" is the set of C++ types and the triple loop
 * in one translation unit (reference src/smallpt.cpp).  This header is the single
 * process/device boundary the new build introduces.  Every entry point cites the
 * reference code it replaces (file:line relative to the reference repo root).
 *
 * Conventions: return 0 = OK, negative = error (see pt_status); no C++ exceptions
 * cross this ABI; the caller owns all host arrays; the library copies on upload and
 * owns all device memory; one host thread drives a context; pt_render is synchronous
 * on return.  Clamp, gamma and the PPM writer stay on the host (src/smallpt.cpp:314-321,
 * :548-551) so output bytes come from the reference's own formula.
 *
 * There is NO CPU fallback behind these symbols: every compute entry fails with
 * PT_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- POD mirrors */

/* Vec, src/smallpt.cpp:24-62 (three doubles; positions, directions and RGB). */
typedef struct pt_vec3 { double x, y, z; } pt_vec3;

/* Refl_t, src/smallpt.cpp:72-74 (enum values 0,1,2). */
enum { PT_DIFF = 0, PT_SPEC = 1, PT_REFR = 2 };

/* Sphere, src/smallpt.cpp:223-228 — field order of :225-227. */
typedef struct pt_sphere {
    double  rad;
    pt_vec3 p, e, c;
    int     refl;
    int     _pad;
} pt_sphere;

/* Plane kinds.  XZ/XY/YZ are the reference's Rectangle_xz (:92-124), Rectangle_xy
 * (:137-167), Rectangle_yz (:180-210): bounded, axis aligned, NO epsilon, in-plane hit
 * coordinates rounded through float.  TILTED is the README's "tilted planes"
 * (README.md:19), absent from the reference source (parity unpinned): bounded
 * parallelogram patch given by a point, a unit normal and two in-plane unit axes. */
enum { PT_PLANE_XZ = 0, PT_PLANE_XY = 1, PT_PLANE_YZ = 2, PT_PLANE_TILTED = 3 };

typedef struct pt_plane {
    int     kind;          /* PT_PLANE_* */
    int     refl;          /* PT_DIFF / PT_SPEC / PT_REFR */
    /* axis-aligned kinds: constructor argument order of the reference classes:
     *   XZ: a=x, b=z, k=y   (Rectangle_xz(x1,x2,z1,z2,y,...)   :97)
     *   XY: a=x, b=y, k=z   (Rectangle_xy(x1,x2,y1,y2,z,...)   :142)
     *   YZ: a=y, b=z, k=x   (Rectangle_yz(y1,y2,z1,z2,x,...)   :185) */
    double  a1, a2, b1, b2, k;
    /* tilted kind: hit = o + d*tau, tau = n.(p0-o)/(n.d);
     * accept iff |(hit-p0).s| <= hs && |(hit-p0).t| <= ht && tau > eps (eps = 1e-4). */
    pt_vec3 p0, n, s, t;
    double  hs, ht;
    pt_vec3 e, c;          /* emission, colour */
} pt_plane;

/* Camera members, src/smallpt.cpp:281-284 (filled by the host Camera ctor :262-275). */
typedef struct pt_camera {
    pt_vec3 origin, lower_left_corner, horizontal, vertical;
} pt_camera;

/* Light descriptor for PT_MODE_NEE_REF_RECT: replaces the literals of
 * src/smallpt.cpp:365-367 (32, 36, 63, 36, 81.6), :467 (id 6) and :471 (1296). */
typedef struct pt_light {
    int    id;             /* scene object id of the light (reference: 6)            */
    int    _pad;
    double x0, xw;         /* x_l = x0 + xw*xi   (reference: 32, 36)                  */
    double z0, zw;         /* z_l = z0 + zw*xi   (reference: 63, 36)                  */
    double y;              /* sampled plane      (reference: 81.6)                    */
    double area;           /* PDF area           (reference: 1296)                    */
} pt_light;

/* Scene = ordered object table (the reference's `Hitable *rect[NUMBER_OBJ]`, :287-311).
 * Object id = position in `order`.  order[i] >= 0 selects planes[order[i]];
 * order[i] < 0 selects spheres[~order[i]].  order == NULL means: all planes in array
 * order, then all spheres in array order. */
typedef struct pt_scene {
    const pt_sphere *spheres; int n_spheres;
    const pt_plane  *planes;  int n_planes;
    const int       *order;
    pt_camera        camera;
    pt_light         light;
} pt_scene;

#define PT_MAX_OBJECTS 16000

/* Integrators.  COS / UNI = src/smallpt.cpp:474-477 arm with the cosine (:337-348) or
 * uniform (:351-360, weight 1 — no 2cos factor, as in the reference) sampler;
 * NEE_REF_RECT = :464-473 (hard-wired rectangular light sampling, reference-faithful);
 * NEE_CONE_SPHERE = solid-angle cone sampling toward every emissive sphere + shadow
 * ray (north-star; not in the reference source — parity unpinned). */
enum { PT_MODE_NEE_REF_RECT = 0, PT_MODE_COS = 1, PT_MODE_UNI = 2, PT_MODE_NEE_CONE_SPHERE = 3 };

/* RNG / arithmetic engines.
 * PT_ENGINE_FP32_PHILOX  : production wavefront path tracer, FP32, Philox4x32-10 keyed by
 *                          (pixel, sample, bounce).
 * PT_ENGINE_FP64_ERAND48 : validation mode — replays the reference's per-row erand48 Xi
 *                          stream (src/smallpt.cpp:530, src/utilities.h:26-51), one thread
 *                          per row, FP64, no FMA contraction. */
enum { PT_ENGINE_FP32_PHILOX = 0, PT_ENGINE_FP64_ERAND48 = 1 };

/* sin/cos used by random_scattering in the FP64 engine:
 * PT_SINCOS_LIBM = CUDA libm sin/cos (last-bit differences vs glibc are possible),
 * PT_SINCOS_DET  = the deterministic +,-,*,/-only sincos of include/ptb200_detmath.h,
 *                  bit-identical on host and device (oracle patch P6). */
enum { PT_SINCOS_LIBM = 0, PT_SINCOS_DET = 1 };

typedef struct pt_render_params {
    int      width, height;   /* src/smallpt.cpp:507 */
    int      spp;             /* src/smallpt.cpp:508 (`samps`) */
    int      mode;            /* PT_MODE_* */
    int      engine;          /* PT_ENGINE_* */
    int      sincos;          /* PT_SINCOS_* (FP64 engine only) */
    uint64_t seed;            /* Philox key (FP32 engine only) */
    /* Row-tile sharding: tile k (tile_rows rows) belongs to rank k % world.
     * rank=0, world=1 renders the whole image.  Pixels of foreign tiles stay zero. */
    int      tile_rows;       /* 0 = default (8) */
    int      rank, world;
    int      max_depth;       /* safety cap on path length; 0 = default (4096); at most 8000 */
    int      queue_capacity;  /* wavefront queue slots; 0 = default (sized to L2) */
    int      collect_stats;   /* 1 = also accumulate per-pixel sum of squares */
    int      bounces_per_launch; /* FP32 engine: bounces a path slot advances per kernel launch; 0 = default (512) */
    /* Progressive / checkpointable accumulation (FP32 engine).  accumulate = 1 renders samples
     * [sample_offset, sample_offset + spp) ON TOP of what the context already holds (same image size, mode, seed and
     * sharding) instead of starting from zero.  Samples are Philox streams keyed by their index and the accumulators
     * are integers, so any split of the sample range gives the bit-identical image; pt_readback divides by the total. */
    int      sample_offset;
    int      accumulate;
    /* Multi-GPU assembly without a gather (pt_render_into): 1 = write ONLY the rows this rank owns and leave the rest
     * of the target untouched, so that all ranks can render straight into ONE image — rank 0's buffer, opened by the
     * other processes through pt_ipc_open and written over NVLink peer memory by the resolve kernel. */
    int      owned_rows_only;
    /* NOT the reference's behaviour (default 0).  The reference's rectangles have no epsilon (src/smallpt.cpp:106): a
     * bounce that starts an ulp behind its own rectangle hits it again at a tiny t and leaks out of the box (SURVEY
     * Appendix C #2; 0.03-0.5 miss events per path).  robust_eps = 1 makes the FP32 engine's rectangles require
     * t > 1e-4 like its spheres do, which removes those leaks (and the energy they lose). */
    int      robust_eps;
} pt_render_params;

typedef struct pt_stats {
    uint64_t paths;             /* (pixel, sample) camera paths traced by this context   */
    uint64_t rays_camera;       /* unique closest-hit queries by kind                    */
    uint64_t rays_scatter;
    uint64_t rays_shadow;
    uint64_t shaded_vertices;   /* surface interactions shaded (bounces)                 */
    uint64_t miss_events;       /* closest-hit queries that missed every object          */
    uint64_t truncated;         /* paths cut by max_depth                                */
    uint64_t kernel_launches;   /* CUDA kernels launched by the last pt_render           */
    uint64_t iterations;        /* wavefront iterations of the last pt_render            */
    uint32_t max_depth_seen;
    uint32_t specialised;       /* 1 = the last FP32 render ran the scene-specialised (NVRTC) build of k_bounce */
    double   render_ms;         /* device time of the last pt_render (CUDA events)       */
    double   main_kernel_ms;    /* FP32 engine: device time of the k_bounce launches that still generated camera paths */
    uint64_t queue_slots_io;    /* path records read + written through the wavefront queues */
    /* phases of the last FP32 render, from %globaltimer stamps taken by the kernels themselves (they add up to
     * render_ms minus the memsets / constant uploads in front of the first launch) */
    double   tail_ms;           /* k_bounce launches after generation was exhausted (the survivors of long paths) */
    double   resolve_ms;        /* fixed point -> FP64 sums, incl. the peer-memory stores of owned_rows_only */
    uint64_t tail_launches;     /* how many of `iterations` belong to tail_ms */
    /* collect_stats = 1 only (the reference's sole counters are path_length and the progress print,
     * src/smallpt.cpp:529,543): why paths ended, and how many were alive at each depth */
    uint64_t term_roulette;     /* ended by Russian roulette on a surface with albedo > 0 (:449, depth > 5)            */
    uint64_t term_emitter;      /* ended on a surface with albedo 0 (the `!p` arm of :448: light sources)              */
    uint64_t term_light_sample; /* NEE_REF_RECT: finished in place along a visible light sample (:466-473)             */
    uint64_t dropped_contributions; /* radiance contributions that were NaN, negative or clamped (>= 6e10): not added  */
    uint64_t spawned_branches;  /* second REFR branches spawned while depth <= 2 (:494-495)                             */
    uint64_t live_at_depth[64]; /* [d] = paths that shaded a vertex at depth d+1 (last bucket: depth >= 64)             */
    uint64_t split_refusals;    /* REFR vertices at depth <= 2 that took one arm because their warp's stack was full (expected: 0) */
    uint64_t accel_structure;   /* 0 = brute force over every primitive; 1 = uniform grid over the small spheres (pt_set_acceleration) */
} pt_stats;

typedef enum pt_status {
    PT_OK = 0,
    PT_ERR_ARG = -1,
    PT_ERR_NO_DEVICE = -2,
    PT_ERR_CUDA = -3,
    PT_ERR_OOM = -4,
    PT_ERR_STATE = -5
} pt_status;

typedef struct pt_ctx pt_ctx;

/* ---------------------------------------------------------------- entry points */

/* Create a context on the current CUDA device (or `device` >= 0) and upload the scene
 * table + camera.  Replaces the global scene literal and Camera construction feeding the
 * loop, src/smallpt.cpp:287-311 and :521.
 * *ctx must be NULL to create a context.  If *ctx is an existing context, its scene is
 * REPLACED in place (host -> device copy of the new tables) and its device buffers (queues,
 * accumulators) are kept — the cheap path for rendering many scenes or frames. */
int pt_scene_upload(pt_ctx **ctx, const pt_scene *scene, int device);

/* Render: the whole triple loop rows x pixels x samples, src/smallpt.cpp:528-541, including
 * radiance() (:419-496 with the dead RL block :424-442 removed) and everything below it.
 * Accumulates into the context's device accumulation buffers (zeroed first). */
int pt_render(pt_ctx *ctx, const pt_render_params *params);

/* Single-process multi-GPU render: ctxs[0..n) are contexts of the SAME scene on n different devices of one node.
 * Row tile k goes to device k % n; the devices render concurrently (one host thread each) and their resolve kernels
 * store the rows straight into ctxs[0]'s image over NVLink peer memory.  Afterwards pt_readback(ctxs[0], ...) returns
 * the whole image and the summed statistics (render_ms = the slowest device).  params->rank/world are ignored. */
int pt_render_multi(pt_ctx **ctxs, int n, const pt_render_params *params);

/* Same, but accumulate into caller-owned DEVICE memory (w*h*3 doubles, row-major, top row
 * first, zeroed by the call) so a host runtime (torch.distributed) can gather row tiles
 * with NCCL without a copy.  `stream` is a cudaStream_t (0 = the context's own stream). */
int pt_render_into(pt_ctx *ctx, const pt_render_params *params, void *dev_rgb_sum, void *stream);

/* Read back per-pixel UNCLAMPED mean radiance: rgb_mean[w*h*3] doubles, row-major, top row
 * first (the order of `c[i]`, src/smallpt.cpp:538).  The caller applies clamp (:538) and
 * toInt (:319-321).  rgb_sumsq (optional, may be NULL) receives per-pixel per-channel sums
 * of squared sample radiance when collect_stats was set.  stats may be NULL. */
int pt_readback(pt_ctx *ctx, double *rgb_mean, double *rgb_sumsq, pt_stats *stats);

/* Peer-memory plumbing for the fused resolve + gather (one process per GPU on one node).  pt_device_alloc returns a
 * cudaMalloc'ed buffer on the context's device (IPC handles name whole allocations, so a framework's sub-allocated
 * tensor will not do); pt_ipc_export / pt_ipc_open wrap cudaIpcGetMemHandle / cudaIpcOpenMemHandle (peer access is
 * enabled lazily); handle = 64 opaque bytes to ship to the other processes by any means. */
int pt_device_alloc(pt_ctx *ctx, size_t bytes, void **dev_ptr);
int pt_device_free(pt_ctx *ctx, void *dev_ptr);
int pt_ipc_export(pt_ctx *ctx, const void *dev_ptr, unsigned char handle[64]);
int pt_ipc_open(pt_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int pt_ipc_close(pt_ctx *ctx, void *dev_ptr);

/* Zero-copy variant: returns the same per-pixel mean image in page-locked host memory OWNED BY THE CONTEXT (one DMA
 * from the device, no host-side copy).  The pointer stays valid until the next pt_readback_view / pt_destroy on this
 * context; NULL on error (pt_last_error).  A host loop such as src/smallpt.cpp:538 can read it in place. */
const double *pt_readback_view(pt_ctx *ctx, pt_stats *stats);

/* Multi-process host-side assembly: the rows THIS rank owns (row tiles k % world == rank of the last pt_render), as
 * means, into the caller's full-size host image (w*h*3 doubles) — one DMA per row tile, every rank over its own PCIe
 * link.  When all ranks map the same image (POSIX shared memory) the picture of src/smallpt.cpp:538 assembles in
 * host memory without funnelling through one GPU.  pt_host_register page-locks caller-owned memory for those DMAs
 * (cudaHostRegister); unregistered memory works too, through the driver's staging. */
int pt_host_register(pt_ctx *ctx, void *host_ptr, size_t bytes);
int pt_host_unregister(pt_ctx *ctx, void *host_ptr);
int pt_readback_owned(pt_ctx *ctx, double *host_image, pt_stats *stats);

/* Resume from a checkpoint: load per-pixel SUMS (and optionally sums of squares) of `spp_done` samples — what
 * pt_accum_download returned earlier, possibly in another process — into the context's accumulators, so that a
 * following pt_render with accumulate = 1 and sample_offset = spp_done continues the image.  Values must be the
 * library's own sums (multiples of 2^-24): they are restored exactly. */
int pt_accum_upload(pt_ctx *ctx, int width, int height, const double *rgb_sum, const double *rgb_sumsq, int spp_done);

/* Checkpoint: per-pixel SUMS of sample radiance (not means) accumulated so far, and the number of samples in them. */
int pt_accum_download(pt_ctx *ctx, double *rgb_sum, double *rgb_sumsq, int *spp_done);

/* Device pointer of the context-owned accumulation buffer (w*h*3 doubles: per-pixel SUM of
 * sample radiance) of the last pt_render, for zero-copy gathers. */
void *pt_accum_device_ptr(pt_ctx *ctx);

/* Closest-hit query for unit tests — `intersect(Ray,t,id)`, src/smallpt.cpp:323-335, plus
 * hittingPoint's miss rule (:371-377).  rays_od = n*6 doubles (o.xyz, d.xyz; d must be
 * unit).  precision: 64 = FP64 reference-compat arithmetic, 32 = the FP32 production
 * intersector.  t_out[i] = 1e20 and id_out[i] = -1 on a miss. */
int pt_debug_intersect(pt_ctx *ctx, const double *rays_od, int n, int precision,
                       double *t_out, int *id_out);

/* Run `n_threads` independent erand48 streams (src/utilities.h:45-51) on the device:
 * seeds = n_threads*3 uint16, out = n_threads*draws doubles. KAT entry. */
int pt_debug_erand48(pt_ctx *ctx, const uint16_t *seeds, int n_threads, int draws, double *out);

/* Philox4x32-10 on the device: ctr = n*4 uint32, key = n*2 uint32, out = n*4 uint32. KAT entry. */
int pt_debug_philox(pt_ctx *ctx, const uint32_t *ctr, const uint32_t *key, int n, uint32_t *out);

/* FFMA-only microbenchmark: measured FP32 issue peak of this device in TFLOP/s
 * (FMA = 2 FLOPs), the denominator of the FP32 roofline. */
int pt_debug_ffma_peak(pt_ctx *ctx, double *tflops, double *sm_clock_mhz);

/* Scene specialisation of the FP32 engine.  The reference's scene is a source literal (src/smallpt.cpp:287-311), so
 * its compiler folds every plane constant; pt_render does the same for an uploaded scene by compiling, with NVRTC,
 * a build of the bounce kernel in which the scene's rectangle constants and primitive counts are immediates (cached
 * per scene, mode and render layout - one GPU or several, row blocks, sample runs - for the life of the process and on
 * disk; ~0.5 s the first time, outside the timed region).
 * mode: 0 = never (generic kernel, scene in __constant__ memory); 1 (default; the environment variable PTB200_JIT
 * overrides the default) = renders of >= 2^25 paths wait for the build, smaller ones never wait: once they have
 * spent 300 ms of GPU time (PTB200_JIT_BG_MS) in the generic kernel the build of their (scene, mode) runs on a host
 * thread and is used once it is there; 2 = always, waiting.
 * The two builds perform the same operations in the same order (bit-identical images), so which one ran shows only
 * in pt_stats.specialised and in the time. */
int pt_set_specialisation(pt_ctx *ctx, int mode);

/* Acceleration structure of the FP32 engine for scenes beyond brute force.  The measured contract stays the
 * brute-force loop of src/smallpt.cpp:323-335 (every primitive, every ray): up to 512 small spheres are scanned that way.
 * mode: 0 = brute force only (a scene with more than 512 small spheres is refused by the FP32 engine);
 *       1 (default) = a uniform grid over the small spheres when there are more than 512 of them;
 *       2 = always the grid (test mode: the same ids, t and images as brute force on scenes that fit both).
 * Rectangles, huge spheres and tilted planes are always tested one by one.  The FP64 engine is always brute force (it is
 * what the grid's hit ids are checked against).  Takes effect at the next pt_scene_upload. */
int pt_set_acceleration(pt_ctx *ctx, int mode);

/* The context's statistics record as it stands (pt_readback needs a finished render; the debug entries only set
 * `specialised`). Test entry. */
int pt_debug_stats(pt_ctx *ctx, pt_stats *stats);

/* Host-only (no device needed): the specialisation header generated for `scene` and render mode `mode`, and the
 * size of the sm_100a cubin NVRTC builds from it.  spec_out/cubin_bytes/seconds may be NULL.  Bits 8 and up of `mode`,
 * when not zero, are 1 + the layout flags (pt_plan_info.layout_flags) of the render the module is for. Test entry. */
int pt_debug_specialise(const pt_scene *scene, int mode, char *spec_out, size_t spec_cap, size_t *cubin_bytes,
                        double *seconds);

/* Host-only (no device needed): how the FP32 engine would lay out a render of `scene` with `params` on a GPU of `sm_count`
 * SMs - rows and pixels this rank owns, row blocks, path slots in flight, sample-run length (1 = single samples), path
 * indices in total, and the layout flags a scene-specialised module is built for (bit 0: blocks of >= 32 pixels, 1: all
 * index divisions by multiply-shift, 2: one GPU owns every row, 3: one row block, 4: sample runs).  Test entry. */
typedef struct pt_plan_info {
    uint64_t owned_rows, owned_pixels, row_blocks, block_rows, path_slots, run_length, path_indices;
    uint32_t layout_flags, splits_refr_paths;
} pt_plan_info;
int pt_debug_plan(const pt_scene *scene, const pt_render_params *params, int sm_count, pt_plan_info *out);

void        pt_destroy(pt_ctx *ctx);
const char *pt_last_error(pt_ctx *ctx);   /* ctx may be NULL: last error of a failed upload */
const char *pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
