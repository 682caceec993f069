// pt_scene_dev.h — the device-visible PODs of the FP32 engine (constant-memory scene image, material table,
// counters).  No system headers: this file is compiled by nvcc (through pt_internal.h) and by NVRTC (as part of the
// scene-specialised kernel source, pt_jit.cu).
#ifndef PT_SCENE_DEV_H
#define PT_SCENE_DEV_H

#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
// the few constants of include/ptb200.h the kernels use
enum { PT_DIFF = 0, PT_SPEC = 1, PT_REFR = 2 };
enum { PT_MODE_NEE_REF_RECT = 0, PT_MODE_COS = 1, PT_MODE_UNI = 2, PT_MODE_NEE_CONE_SPHERE = 3 };
#endif

enum { OT_SPHERE = 0, OT_XZ = 1, OT_XY = 2, OT_YZ = 3, OT_TILT = 4 };

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
    float light_e[3], light_c[3]; // emission and albedo of that light (a black-bodied light ends the path in place)
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
    int    refl_mask;             // bit r set: some object has material r (PT_DIFF / PT_SPEC / PT_REFR)
    double huge[PT_MAX_HUGE][4];  // c.x, c.y, c.z, rad^2 in FP64 (the FP64 c-term of round 1: -DPT_HUGE_FP64)
    float  hugef[PT_MAX_HUGE][4]; // the same centres rounded to FP32 (for b = (c - o).d)
    // re-centred FP32 form of the same spheres (see closest_hit): G = centre - huge_c and K = |G|^2 - rad^2, each as a
    // two-float (hi + lo): {G_hi.xyz, G_lo.xyz, K_hi, K_lo}
    float  hugeg[PT_MAX_HUGE][8];
    float  huge_c[3];             // reference point near the rays' origins (centre of the non-huge objects)
    float4 tilt[PT_MAX_TILT][4];  // {n.xyz, n.p0} {s.xyz, s.p0} {t.xyz, t.p0} {hs, ht, -, -}
};

// Uniform grid over the small spheres (opt-in, see pt_set_acceleration): CSR cell lists in global memory.
struct GridDev {
    float lo[3], cell[3], inv_cell[3];   // lower corner, cell size, 1 / cell size
    int   res[3];                        // cells per axis
    const unsigned int *start;           // res.x * res.y * res.z + 1 offsets into items
    const unsigned int *items;           // sphere indices, ascending within a cell
    const float4 *sph;                   // {centre, r^2} by sphere index (code = code_sph0 + index)
    int   n;                             // spheres in the grid (0 = no grid)
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
    // collect_stats only
    unsigned long long term_roulette, term_emitter, term_light_sample, dropped, spawned, split_refused;
    unsigned long long live_at_depth[64];
};

// One record per k_bounce launch, written by the launch itself (block 0, thread 0): when it started and in which phase.
struct LaunchRec {
    unsigned long long t_start;   // %globaltimer, ns
    unsigned int exhausted;       // 1 = every camera path had been handed out when this launch started (tail)
    unsigned int n_in;            // live path slots it found in its input queue
};
#define PT_MAX_LAUNCH_RECS 4096

#endif
