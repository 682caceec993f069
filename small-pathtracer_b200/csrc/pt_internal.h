// pt_internal.h — types shared by the translation units of libptb200.so (not part of the ABI).
#ifndef PT_INTERNAL_H
#define PT_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ptb200.h"
#include "pt_scene_dev.h"

// ------------------------------------------------------------------ FP64 engine scene (global memory)

struct DevObj64 {
    int    type, refl;
    // sphere: g = {rad, p.x, p.y, p.z, -}; rectangles: g = {a1, a2, b1, b2, k} (reference ctor order)
    double g[5];
    double e[3], c[3];
    // tilted plane only
    double p0[3], n[3], s[3], t[3], hs, ht;
};

// ------------------------------------------------------------------ context
struct PtJitKernel;
#define PT_STAGE_LANES 3
struct PtStageLane {                  // one read-back pipeline: its own stream, two pinned blocks of PT_STAGE_ELEMS doubles
    cudaStream_t stream = nullptr;
    double *h_stage = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
#define PT_JIT_BACKGROUND_AFTER_MS 300.0   /* jit_mode 1, small renders: GPU time in the generic kernel before the background build starts */
#ifndef PT_JIT_SPH_IMM_MAX
#define PT_JIT_SPH_IMM_MAX 256          /* specialised build: sphere scan tables up to this size become immediates */
#endif
#define PT_STAGE_ELEMS ((size_t)1 << 17)   /* 1 MB staging blocks */
#define PT_JIT_MIN_PATHS (1ull << 25)   /* jit_mode 1: renders at least this big WAIT for the specialised build (NVRTC ~0.5 s once, or the disk cache); smaller ones get it in the background from their second render on */

struct pt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0, l2_bytes = 0;
    // scene
    std::vector<DevObj64> objs;
    pt_camera cam{};
    pt_light light{};
    DevObj64 *d_objs = nullptr;
    int n_alloc = 0;                   // objects d_objs / d_mats can hold
    SceneF32 *h_scene32 = nullptr;     // host staging copy (heap; copied to __constant__ before FP32 launches)
    MatF32 *d_mats = nullptr;          // indexed by code
    float4 *d_sphf = nullptr;          // SceneF32::sphf mirrored in global memory (staged into shared memory by k_bounce)
    int n_codes_alloc = 0;
    // acceleration structure (pt_set_acceleration): uniform grid over the small spheres
    int accel_mode = 1;
    GridDev grid{};                    // device pointers + geometry (grid.n == 0: brute force)
    std::vector<unsigned int> h_grid_start, h_grid_items;
    std::vector<float4> h_grid_sph;
    unsigned int *d_grid_start = nullptr, *d_grid_items = nullptr;
    float4 *d_grid_sph = nullptr;
    size_t grid_start_cap = 0, grid_items_cap = 0, grid_sph_cap = 0;
    bool fp32_ok = false;              // scene fits the FP32 constant layout
    std::string fp32_why;
    // render state
    pt_render_params last{};
    bool rendered = false;
    bool rendered_fp32 = false;                        // the accumulators hold an FP32-engine image (accumulate = 1 may follow)
    double *d_sum = nullptr, *d_sumsq = nullptr;       // context-owned accumulation (w*h*3)
    double *d_sum_ext = nullptr;                       // caller-owned target of pt_render_into
    size_t accum_elems = 0;
    unsigned long long *d_fix = nullptr, *d_fixsq = nullptr;  // FP32 engine fixed-point accumulators
    size_t fix_elems = 0;
    long long accum_spp = 0;                           // samples per pixel held by the accumulators (progressive renders add up)
    bool fix_has_sq = false;                           // d_fixsq holds the sums of squares of those samples
    // wavefront queues (FP32 engine)
    float4 *q[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int q_capacity = 0;
    uint4 *d_warp_chunk = nullptr;                     // per warp: path indices reserved but not yet traced
    float4 *d_spawn = nullptr;                         // per-warp stacks of spawned REFR branches (scenes with REFR only)
    int spawn_warps = 0;
    unsigned int *d_counts = nullptr;                  // per-iteration live counts etc.
    int counts_len = 0;
    size_t counts_dirty = 0;
    LaunchRec *d_launch_rec = nullptr;                 // one record per k_bounce launch (phase split of pt_stats)
    unsigned long long *d_stamps = nullptr;            // %globaltimer at the start of the resolve kernel and at the end of the render
    unsigned int *h_pinned = nullptr;                  // 2 pinned words for the termination check
    cudaEvent_t ev_batch[2] = {nullptr, nullptr};
    DevStats *h_stats = nullptr;                       // pinned
    cudaEvent_t ev_rb = nullptr;                       // orders the read-back lanes after the context's stream
    PtStageLane lane[PT_STAGE_LANES];                  // pt_readback: device -> pinned staging -> caller's buffer
    double *d_mean = nullptr;                          // per-pixel mean (sum / spp), produced on the device at readback
    size_t mean_elems = 0;
    double *h_view = nullptr;                          // pinned host image handed out by pt_readback_view
    size_t view_elems = 0;
    DevStats *d_stats = nullptr;
    pt_stats stats{};
    // scene specialisation: 0 = generic kernel only, 1 = specialise renders of >= PT_JIT_MIN_PATHS paths, 2 = always
    int jit_mode = 1;
    PtJitKernel *jit = nullptr;                        // kernel chosen for the render in flight (nullptr = generic)
    int jit_flags = -1;                                // ... and the PT_RF_* flags it was asked for (-1: none)
    std::string jit_note;
    std::string err;
};

int pt_fail(pt_ctx *ctx, int code, const std::string &msg);

// scene-specialised kernels (pt_jit.cu)
struct PtJitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    cudaKernel_t kern_isect = nullptr;     // k_intersect_jit (pt_debug_intersect through the specialised closest_hit)
    double compile_seconds = 0;
};
// render_flags: PT_RF_* bits of the render the module is built for (regeneration branches as compile-time constants), -1 = none
std::string pt_jit_spec(const SceneF32 &S, int mode, bool stats, bool with_intersect = false, int render_flags = -1);
int pt_jit_block(const SceneF32 &S);      // threads per block of the scene's specialised module
int pt_jit_compile(const std::string &spec, std::vector<char> &cubin, std::string &log, double *seconds);
PtJitKernel *pt_jit_get(pt_ctx *ctx, int mode, bool stats, bool with_intersect = false, bool wait = true, int render_flags = -1);
void pt_jit_account(pt_ctx *ctx, int mode, bool stats, double ms, int render_flags = -1);
int pt_jit_build(const std::string &spec, std::vector<char> &cubin, std::string &log, double *seconds, bool *from_disk);
#define PT_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return pt_fail(ctx, e_ == cudaErrorMemoryAllocation ? PT_ERR_OOM : PT_ERR_CUDA,     \
                           std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

// engines (each in its own translation unit; the FP64 one is compiled with -fmad=false)
int pt_fp64_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s);
int pt_fp64_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s);
int pt_fp64_erand48(pt_ctx *ctx, const uint16_t *d_seeds, int n_threads, int draws, double *d_out, cudaStream_t s);

// How the FP32 engine lays a render out (pt_fp32_plan): row blocks, slots in flight, sample runs, and the PT_RF_* flags the
// regeneration step of k_bounce branches on (a scene-specialised module is built for one combination of them)
#define PT_RF_WRAP_ONCE 1      /* a block has at least 32 pixels */
#define PT_RF_MAGIC 2          /* every index -> row division is a multiply-shift */
#define PT_RF_WORLD1 4         /* one GPU owns every row */
#define PT_RF_ONE_BLOCK 8      /* no row blocks */
#define PT_RF_RUNS 16          /* sample runs */
struct Fp32Plan {
    long long owned_rows = 0;
    unsigned long long owned_pixels = 0, n_blk = 1, blk_rows = 0, blk_pixels = 0, spp_runs = 0, total = 0;
    int cap = 0;
    unsigned int run_shift = 0;
    bool want_spawn = false, use_magic = false, pix_magic = false;
    int flags = 0;
};
void pt_fp32_plan(const pt_ctx *ctx, const pt_render_params *p, Fp32Plan &pl);
int pt_fp32_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s);
int pt_fp32_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s);
int pt_fp32_philox(pt_ctx *ctx, const uint32_t *d_ctr, const uint32_t *d_key, int n, uint32_t *d_out, cudaStream_t s, int width);
int pt_fp32_ffma_peak(pt_ctx *ctx, double *tflops, double *mhz);

#endif
