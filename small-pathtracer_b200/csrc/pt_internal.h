// pt_internal.h — types shared by the translation units of libptb200.so (not part of the ABI).
#ifndef PT_INTERNAL_H
#define PT_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ptb200.h"

// ------------------------------------------------------------------ FP64 engine scene (global memory)
enum { OT_SPHERE = 0, OT_XZ = 1, OT_XY = 2, OT_YZ = 3, OT_TILT = 4 };

struct DevObj64 {
    int    type, refl;
    // sphere: g = {rad, p.x, p.y, p.z, -}; rectangles: g = {a1, a2, b1, b2, k} (reference ctor order)
    double g[5];
    double e[3], c[3];
    // tilted plane only
    double p0[3], n[3], s[3], t[3], hs, ht;
};

// ------------------------------------------------------------------ FP32 engine scene
#define PT_MAX_OBJ        512     // objects the FP32 constant-memory layout holds
#define PT_MAX_HUGE       64
#define PT_MAX_TILT       64
#define PT_HUGE_RADIUS    100.0   // spheres at least this big take the FP64 c-term path

#define PT_SPH_KAPPA      3.814697265625e-6f   /* 2^-18: relative slack of the conservative sphere scan */

#define PT_RECT_SLOTS     16      // rectangles per axis class tested by fully unrolled, constant-operand code

// The FP32 engine addresses objects by CODE (position in the class-sorted layout), not by scene id:
//   [0, 48)                         unrolled rectangle slots, code = axis*16 + k   (axis: XZ 0, XY 1, YZ 2)
//   [48, 48 + n_ovf)                overflow rectangles (generic loop), axis by axis
//   [code_sph0, +n_sph)             small spheres        [code_huge0, +n_huge)  huge spheres
//   [code_tilt0, +n_tilt)           tilted planes
// Within a class codes ascend with scene id, so "lowest id wins ties" (src/smallpt.cpp:328) holds per class.
struct SceneF32 {                 // lives in __constant__ memory: every access is warp-uniform
    int   n_slot[3];              // rectangles in the unrolled slots of each axis class (<= PT_RECT_SLOTS)
    int   ovf_begin[4];           // [axis] .. [axis+1): overflow entries of rect_a / rect_b2
    int   n_sph, n_huge, n_tilt;
    int   code_sph0, code_huge0, code_tilt0, n_codes;
    int   code_obj0;              // code of scene object 0 (where a missed ray "lands", :373-374)
    // NEE_REF_RECT light (src/smallpt.cpp:365-367,467,471)
    int   light_code;
    float lx0, lxw, lz0, lzw, ly, larea;
    int   n_lights;               // emissive spheres for NEE_CONE_SPHERE
    int   light_sph_code[32];
    float4 slot_a[3][PT_RECT_SLOTS];   // k, a1, a2 - a1, b1   (one 128-bit uniform load)
    float  slot_b2[3][PT_RECT_SLOTS];  // b2 - b1
    float4 rect_a[PT_MAX_OBJ];    // overflow rectangles: k, a1, a2, b1
    float  rect_b2[PT_MAX_OBJ];   //                      b2
    float4 sph[PT_MAX_OBJ];       // c.x, c.y, c.z, rad^2
    // conservative scan form of the same spheres (see closest_hit): centres relative to sph_c, w = |c'|^2 - rad^2;
    // padded to a multiple of 4 with entries that can never pass (w = 3e38)
    float4 sphf[PT_MAX_OBJ + 4];
    float  sph_c[3];              // translation that centres the small spheres around the origin
    float  sph_kM2;               // PT_SPH_KAPPA * max_i (|c'_i| + rad_i)^2
    int    n_sph4;                // n_sph rounded up to a multiple of 4
    double huge[PT_MAX_HUGE][4];  // c.x, c.y, c.z, rad^2 in FP64
    float4 tilt[PT_MAX_TILT][4];  // {n.xyz, n.p0} {s.xyz, s.p0} {t.xyz, t.p0} {hs, ht, -, -}
};

struct MatF32 {                   // global memory, indexed by CODE (divergent index, so NOT constant memory)
    float4 c_refl;                // c.xyz, refl (int bits)
    float4 e_type;                // e.xyz, type (int bits)
    float4 geom;                  // sphere: centre.xyz, 1/rad ; rect: k_hi, k_lo (k = hi + lo), -, - ; tilted: n.xyz
    float4 aux;                   // tilted: p0.xyz ; .w = scene id (int bits)
};

struct DevStats {                 // device-side counters (unsigned long long for atomicAdd)
    unsigned long long paths, rays_camera, rays_scatter, rays_shadow, shaded, misses, truncated;
    unsigned int max_depth_seen, pad;
};

// ------------------------------------------------------------------ context
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
    bool fp32_ok = false;              // scene fits the FP32 constant layout
    std::string fp32_why;
    // render state
    pt_render_params last{};
    bool rendered = false;
    double *d_sum = nullptr, *d_sumsq = nullptr;       // context-owned accumulation (w*h*3)
    double *d_sum_ext = nullptr;                       // caller-owned target of pt_render_into
    size_t accum_elems = 0;
    unsigned long long *d_fix = nullptr, *d_fixsq = nullptr;  // FP32 engine fixed-point accumulators
    size_t fix_elems = 0;
    // wavefront queues (FP32 engine)
    float4 *q[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int q_capacity = 0;
    uint4 *d_warp_chunk = nullptr;                     // per warp: path indices reserved but not yet traced
    unsigned int *d_counts = nullptr;                  // per-iteration live counts etc.
    int counts_len = 0;
    size_t counts_dirty = 0;
    unsigned int *h_pinned = nullptr;                  // 2 pinned words for the termination check
    cudaEvent_t ev_batch[2] = {nullptr, nullptr};
    DevStats *h_stats = nullptr;                       // pinned
    double *h_stage = nullptr;                         // pinned staging for pt_readback
    size_t stage_elems = 0;
    DevStats *d_stats = nullptr;
    pt_stats stats{};
    std::string err;
};

int pt_fail(pt_ctx *ctx, int code, const std::string &msg);
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

int pt_fp32_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s);
int pt_fp32_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s);
int pt_fp32_philox(pt_ctx *ctx, const uint32_t *d_ctr, const uint32_t *d_key, int n, uint32_t *d_out, cudaStream_t s);
int pt_fp32_ffma_peak(pt_ctx *ctx, double *tflops, double *mhz);

#endif
