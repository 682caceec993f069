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

struct SceneF32 {                 // lives in __constant__ memory: every access is warp-uniform
    int   n_obj;
    int   rect_begin[4];          // [axis] .. [axis+1): entries of rect_* for XZ(0), XY(1), YZ(2)
    int   n_sph, n_huge, n_tilt;
    // NEE_REF_RECT light (src/smallpt.cpp:365-367,467,471)
    int   light_id;
    float lx0, lxw, lz0, lzw, ly, larea;
    int   n_lights;               // emissive spheres for NEE_CONE_SPHERE
    int   light_sph[32];          // object ids
    float4 rect_a[PT_MAX_OBJ];    // k, a1, a2, b1
    float2 rect_b[PT_MAX_OBJ];    // b2, id (int bits)
    float4 sph[PT_MAX_OBJ];       // c.x, c.y, c.z, rad^2
    int    sph_id[PT_MAX_OBJ];
    double huge[PT_MAX_HUGE][4];  // c.x, c.y, c.z, rad^2 in FP64
    int    huge_id[PT_MAX_HUGE];
    float4 tilt[PT_MAX_TILT][4];  // {n.xyz, n.p0} {s.xyz, s.p0} {t.xyz, t.p0} {hs, ht, id, -}
};

struct MatF32 {                   // global memory (indexed by the hit id: divergent, so NOT constant)
    float4 c_refl;                // c.xyz, refl (int bits)
    float4 e_type;                // e.xyz, type (int bits)
    float4 geom;                  // sphere: centre.xyz, 1/rad ; rect: k_hi, k_lo (k = hi + lo), -, - ; tilted: n.xyz
    float4 aux;                   // tilted: p0.xyz
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
    MatF32 *d_mats = nullptr;
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
