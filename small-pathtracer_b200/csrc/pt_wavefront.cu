// pt_wavefront.cu — production engine (PT_ENGINE_FP32_PHILOX): an FP32 wavefront path tracer for sm_100a.
//
// What it replaces (reference src/smallpt.cpp): the triple loop :528-541, radiance() :419-496,
// intersect()/hittingPoint() :323-335/:371-377, the per-primitive intersect()/normal() members
// (:102-124, :145-167, :188-210, :229-253), random_scattering() :337-360 and light_sampling() :363-369.
//
// Organisation (one kernel launch = one bounce of every live path):
//   * path state lives in SoA float4 queues (3 x 16 B per path: origin+pixel, direction+sample,
//     throughput+depth/prev/E); loads and stores are 128-bit and fully coalesced; the queues are sized to
//     stay resident in the 126 MB L2 between bounces;
//   * the scene sits in __constant__ memory, sorted by primitive class, so the intersection loops have
//     warp-uniform operands and no memory traffic;
//   * k_bounce fuses extend (closest hit), shade (emission, Russian roulette, light sampling + shadow
//     ray, BSDF sampling), REGENERATION (a lane whose path ended starts the next camera path: ray
//     generation with uniform sub-pixel jitter) and COMPACTION (warp ballot + block prefix sum + one
//     atomic per block) of the survivors into the output queue;
//   * randomness is Philox4x32-10 keyed by (pixel, sample, vertex): the image does not depend on queue
//     order, chunking or the number of GPUs;
//   * radiance is accumulated per pixel in 64-bit fixed point (2^-24) with integer atomics, which are
//     associative: the result is bit-reproducible run to run and across shardings.
//
// FP32 numerics: rectangles keep the reference's no-epsilon rule and its (k-o)/d, o+d*t
// forms for the winning hit so the self-hit "leak" statistics carry over; spheres use the
// perpendicular-distance discriminant; spheres with radius >= PT_HUGE_RADIUS get the c = |o-p|^2-r^2 term
// in FP64 and conjugate roots; the sphere a ray starts on is solved exactly (roots {0, 2b}).
#include <math_constants.h>

#include "pt_internal.h"
#include "pt_rng.cuh"

__constant__ SceneF32 c_scene;

namespace {

#define PT_PI_F 3.14159265358979323846f
#define PT_INV_PI_F 0.31830988618379067154f
#define PT_EPS_F 1e-4f
#define PT_DEPTH_DEAD 0xFFFFu
#define PT_FIX_SCALE 16777216.0f        /* 2^24 */
#define PT_FIX_INV 5.9604644775390625e-8 /* 2^-24 */
#ifndef PT_BLOCK
#define PT_BLOCK 256
#endif
#ifndef PT_BLOCKS_PER_SM
#define PT_BLOCKS_PER_SM (1024 / PT_BLOCK)   /* 4 blocks of 256: 64 registers/thread keep the bounce loop's state out of local memory */
#endif
#ifndef PT_DEFAULT_WAVES
#define PT_DEFAULT_WAVES 6
#endif
#ifndef PT_DEFAULT_ITERS
#define PT_DEFAULT_ITERS 32         /* bounces per launch while camera paths are being generated */
#endif
#ifndef PT_DEFAULT_ITERS_TAIL
#define PT_DEFAULT_ITERS_TAIL 2     /* ... once generation is exhausted (compaction pays in the tail) ... */
#endif
#ifndef PT_DEFAULT_ITERS_DRAIN
#define PT_DEFAULT_ITERS_DRAIN 16   /* ... and once the survivors no longer fill the GPU (launch latency dominates) */
#endif

struct KParams {
    float4 *qin[4];
    float4 *qout[4];
    const unsigned int *n_in;          // live count of the input queue (this launch)
    unsigned int *n_out;               // survivors (next launch), zero before the launch
    unsigned long long *gen_counter;   // next path index to hand out (advanced PT chunk by chunk, one atomic per warp)
    uint4 *warp_chunk;                 // per warp: {base lo, base hi, left, -} of the path indices it still holds
    int iters, iters_tail, iters_drain; // bounces per launch: generating / generation exhausted / survivors < drain_below
    unsigned int drain_below;
    unsigned int chunk;                // path indices a warp reserves per atomic (>= 32)
    unsigned long long total_paths;
    unsigned int owned_pixels;
    double inv_owned_pixels;
    int w, h, spp, tile_rows, rank, world, max_depth;
    unsigned long long magic_w, magic_tile;   // ceil(2^40 / w), ceil(2^40 / tile_rows): exact n / d for n < 2^24, d < 2^16
    int use_magic;                     // owned_pixels < 2^24
    float cam_o[3], cam_base[3], cam_h[3], cam_v[3];   // origin, llc - origin, horizontal, vertical
    float inv_w, inv_h;
    unsigned int seed_lo, seed_hi;
    unsigned long long *fix;           // w*h*3 fixed-point sums
    unsigned long long *fixsq;         // w*h*3 fixed-point sums of squares (STATS only)
    const MatF32 *mats;
    const float4 *sphf;                // SceneF32::sphf in global memory
    DevStats *stats;
};

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float b) { return f3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 fma3(F3 d, float t, F3 o) { return f3(fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z)); }
// single-MUFU approximations (the .ftz forms: without it every call drags a denormal-rescaling sequence along)
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ F3 normalize3(F3 a) { return a * rsqrt_fast(dot3(a, a)); }

// ---------------------------------------------------------------------------------------------- extend
// One rectangle of the reference (:102-112 / :145-155 / :188-198): t = (k - o_a) / d_a, the two in-plane
// coordinates against [a1,a2] x [b1,b2], NO epsilon, t == 0 and t < 0 are misses.  AX and K are compile-time, so
// every scene constant is a constant-bank operand of the arithmetic instruction itself (no loads, no loop).
// Slots are visited in DESCENDING k with `t <= best`, which keeps the lowest id on ties like the strict `<`
// of the ascending loop at :328.
template <int AX, int K>
__device__ __forceinline__ void rect_slot(float oa, float ia, float ou, float du, float ov, float dv, unsigned int &best_bits, int &code)
{
    // slot_a = {k, a1, a2 - a1, b1}, slot_b2 = b2 - b1.  The six float compares (ALU pipe, half rate) become three
    // unsigned-integer compares: for w = u - a1, `a1 <= u <= a2` is `bits(w) <= bits(a2 - a1)` (a negative w has the
    // sign bit set and compares high); `0 < t <= best` is `bits(t') <= bits(best)` with t' = t - denorm_min folded
    // into the FMA (t = 0 becomes negative; t < 0 and NaN compare high; any other t is unchanged by rounding).
    // The update is a pair of predicated moves, which ptxas can place on the FMA pipe.
    const float4 ra = c_scene.slot_a[AX][K];
    const float t = fmaf(ra.x - oa, ia, -1.401298464e-45f);
    const float wu = fmaf(du, t, ou) - ra.y, wv = fmaf(dv, t, ov) - ra.w;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.u32 p, %2, %3;\n\t"
        "setp.le.and.u32 p, %4, %5, p;\n\t"
        "setp.le.and.u32 p, %6, %7, p;\n\t"
        "@p mov.b32 %0, %6;\n\t"
        "@p mov.b32 %1, %8;\n\t}"
        : "+r"(best_bits), "+r"(code)
        : "r"(__float_as_uint(wu)), "r"(__float_as_uint(ra.z)), "r"(__float_as_uint(wv)), "r"(__float_as_uint(c_scene.slot_b2[AX][K])),
          "r"(__float_as_uint(t)), "r"(best_bits), "n"(AX * PT_RECT_SLOTS + K));
}

#define PT_SLOT_CASE(K) case K + 1: rect_slot<AX, K>(oa, ia, ou, du, ov, dv, best_bits, code); /* fall through */

template <int AX>
__device__ __forceinline__ void rects_axis(float oa, float ia, float ou, float du, float ov, float dv, float &best, int &code)
{
    unsigned int best_bits = __float_as_uint(best);
    switch (c_scene.n_slot[AX]) {          // warp-uniform jump into the unrolled sequence (Duff's device)
        PT_SLOT_CASE(15) PT_SLOT_CASE(14) PT_SLOT_CASE(13) PT_SLOT_CASE(12) PT_SLOT_CASE(11) PT_SLOT_CASE(10) PT_SLOT_CASE(9)
        PT_SLOT_CASE(8) PT_SLOT_CASE(7) PT_SLOT_CASE(6) PT_SLOT_CASE(5) PT_SLOT_CASE(4) PT_SLOT_CASE(3) PT_SLOT_CASE(2)
        PT_SLOT_CASE(1) PT_SLOT_CASE(0)
    default: break;
    }
    best = __uint_as_float(best_bits);
    // overflow rectangles of this axis class (more than PT_RECT_SLOTS): generic loop, ascending, strict <
    for (int i = c_scene.ovf_begin[AX]; i < c_scene.ovf_begin[AX + 1]; i++) {
        const float4 ra = c_scene.rect_a[i];
        const float t = (ra.x - oa) * ia;
        const float u = fmaf(du, t, ou), v = fmaf(dv, t, ov);
        const bool ok = !(u < ra.y) && !(u > ra.z) && !(v < ra.w) && !(v > c_scene.rect_b2[i]) && (t > 0.f) && (t < best);
        if (ok) { best = t; code = 3 * PT_RECT_SLOTS + i; }
    }
}

// intersect(Ray,t,id), :323-335.  prev = code of the object the ray starts on (-1: none).
// Returns best t (1e20f on a miss) and the winner's code (-1 on a miss).
#ifdef PT_NOINLINE_HIT
#define PT_HIT_INLINE __noinline__
#else
#define PT_HIT_INLINE __forceinline__
#endif
__device__ PT_HIT_INLINE void closest_hit(F3 o, F3 d, int prev, const float4 *__restrict__ s_sphf, float &t_out, int &code_out)
{
    float best = 1e20f;
    int code = -1;
    const float ix = rcp_fast(d.x), iy = rcp_fast(d.y), iz = rcp_fast(d.z);
    rects_axis<0>(o.y, iy, o.x, d.x, o.z, d.z, best, code);   // XZ: plane y
    rects_axis<1>(o.z, iz, o.x, d.x, o.y, d.y, best, code);   // XY: plane z
    rects_axis<2>(o.x, ix, o.y, d.y, o.z, d.z, best, code);   // YZ: plane x

    // Sphere::intersect, :229-239, eps = 1e-4, in two stages.
    // Scan (every sphere; LDS.128 + 7 FFMA + compare + mask bit): with centres c' and the origin o' relative to sph_c,
    //   b = c'.d - o'.d,   det - |o'|^2 = b^2 - (|c'|^2 - r^2 - 2 c'.o')
    // needs no per-sphere subtraction; its FP32 cancellation error is bounded by a few ulp of (|c'| + |o'|)^2, so the
    // test `det >= -kappa ((max|c'| + r)^2 + |o'|^2)` is CONSERVATIVE (kappa = 2^-18, ~5x the bound): it never rejects
    // a sphere the exact test would accept.  Candidates of 32 spheres are collected in a per-lane bit mask, branch-free.
    // Exact stage (rare): every lane pops ITS candidates (lowest index first; lanes work on different spheres at the
    // same time, so the warp runs max-over-lanes iterations, typically 1-2 per 32 spheres): the perpendicular-distance
    // discriminant det = r^2 - |op - b d|^2 on the un-translated data.
    const int ns4 = c_scene.n_sph4;
    if (ns4 > 0) {
        const F3 oc = f3(o.x - c_scene.sph_c[0], o.y - c_scene.sph_c[1], o.z - c_scene.sph_c[2]);
        const float O2 = dot3(oc, oc), OD = dot3(oc, d);
        const float thr = fmaf(O2, 1.f - PT_SPH_KAPPA, -c_scene.sph_kM2);
        const F3 o2 = oc * -2.f;
        const int prev_s = prev - c_scene.code_sph0;
        const float4 *s_sphx = s_sphf + (PT_MAX_OBJ + 4);     // exact data {centre, r^2} behind the scan table
#define PT_SPH_SCAN(IDX, BIT)                                                                               \
        {                                                                                                   \
            const float4 s = s_sphf[IDX];                                                                   \
            const float b = fmaf(s.x, d.x, fmaf(s.y, d.y, fmaf(s.z, d.z, -OD)));                            \
            const float c = fmaf(s.x, o2.x, fmaf(s.y, o2.y, fmaf(s.z, o2.z, s.w)));                         \
            if (fmaf(b, b, -c) >= thr) mask |= (BIT);                                                       \
        }
#pragma unroll 1
        for (int base = 0; base < ns4; base += 32) {
            unsigned int mask = 0u;
            if (base + 32 <= ns4) {
#pragma unroll
                for (int k = 0; k < 32; k++) PT_SPH_SCAN(base + k, 1u << k)
            } else {
#pragma unroll 1
                for (int g = 0; base + g < ns4; g += 4) {
#pragma unroll
                    for (int k = 0; k < 4; k++) PT_SPH_SCAN(base + g + k, (1u << k) << g)
                }
            }
            while (mask) {
                const int i = base + __ffs(mask) - 1;
                mask &= mask - 1u;
                const float4 s = s_sphx[i];
                F3 op = f3(s.x - o.x, s.y - o.y, s.z - o.z);
                float b = dot3(op, d);
                F3 l = f3(fmaf(-b, d.x, op.x), fmaf(-b, d.y, op.y), fmaf(-b, d.z, op.z));
                float dd = s.w - dot3(l, l);
                float sq = sqrt_fast(fmaxf(dd, 0.f));
                float t0 = b - sq, t1 = b + sq;
                float tt = t0 > PT_EPS_F ? t0 : t1;
                if (i == prev_s) tt = b + b;          // origin on this sphere: roots are exactly {0, 2b}
                // ascending i with strict < keeps the lowest id on ties (:328)
                if (dd >= 0.f && tt > PT_EPS_F && tt < best) { best = tt; code = c_scene.code_sph0 + i; }
            }
        }
#undef PT_SPH_SCAN
    }
    // Huge spheres (the 1e5-radius walls of the sphere-era scene): c in FP64, conjugate roots in FP32.
    const int nh = c_scene.n_huge;
    if (nh > 0) {
        const double ox = (double)o.x, oy = (double)o.y, oz = (double)o.z;
        const int prev_h = prev - c_scene.code_huge0;
        for (int i = 0; i < nh; i++) {
            double px = c_scene.huge[i][0] - ox, py = c_scene.huge[i][1] - oy, pz = c_scene.huge[i][2] - oz;
            double c64 = fma(px, px, fma(py, py, fma(pz, pz, -c_scene.huge[i][3])));
            float c = (i == prev_h) ? 0.f : (float)c64;
            float b = dot3(f3((float)px, (float)py, (float)pz), d);
            float det = fmaf(b, b, -c);
            if (det >= 0.f) {
                float q = b + copysignf(sqrtf(det), b);
                float ta = q, tb = c * rcp_fast(q);
                float lo = fminf(ta, tb), hi = fmaxf(ta, tb);
                float tt = lo > PT_EPS_F ? lo : hi;
                if (tt > PT_EPS_F && tt < best) { best = tt; code = c_scene.code_huge0 + i; }
            }
        }
    }
    // Tilted bounded planes, eps = 1e-4.
    const int nt = c_scene.n_tilt;
    for (int i = 0; i < nt; i++) {
        const float4 pn = c_scene.tilt[i][0], ps = c_scene.tilt[i][1], pt = c_scene.tilt[i][2], pe = c_scene.tilt[i][3];
        float denom = fmaf(pn.x, d.x, fmaf(pn.y, d.y, pn.z * d.z));
        float num = pn.w - fmaf(pn.x, o.x, fmaf(pn.y, o.y, pn.z * o.z));
        float tau = num * rcp_fast(denom);
        F3 hp = fma3(d, tau, o);
        float a = fmaf(ps.x, hp.x, fmaf(ps.y, hp.y, ps.z * hp.z)) - ps.w;
        float b = fmaf(pt.x, hp.x, fmaf(pt.y, hp.y, pt.z * hp.z)) - pt.w;
        bool ok = (fabsf(a) <= pe.x) && (fabsf(b) <= pe.y) && (tau > PT_EPS_F) && (tau < best);
        if (ok) { best = tau; code = c_scene.code_tilt0 + i; }
    }
    t_out = best;
    code_out = code;
}

// hittingPoint (:371-377) for the winning object, with t refined once (the loop's t is rcp/approx-sqrt based).
// Rectangles: t and the plane coordinate are recomputed with IEEE division and separate multiply/add — the
// reference's own forms — so that the distribution of "exactly on / just in front of / just behind the plane"
// (which drives its self-hit leaks) carries over; the plane constant is a two-float (hi + lo) so t keeps
// ~1e-7 relative accuracy even where FP32 cannot represent k (81.6, 81.5).
// Small spheres: one Newton step on |o + d t - c|^2 = r^2 from the hit point (numbers near the surface are
// small, so the residual is accurate where the quadratic's coefficients were not).
__device__ __forceinline__ void refine_hit(F3 o, F3 d, float &t, int type, float4 geom, float4 aux, F3 &x)
{
    if (type >= OT_XZ && type <= OT_YZ) {
        // one IEEE division for the three axis classes: pick the plane's axis component first
        const float oa = type == OT_XZ ? o.y : type == OT_XY ? o.z : o.x;
        const float da = type == OT_XZ ? d.y : type == OT_XY ? d.z : d.x;
        t = __fdiv_rn((geom.x - oa) + geom.y, da);
        const float xa = __fadd_rn(oa, __fmul_rn(da, t));     // the reference's o + d*t, no FMA (:375)
        x = fma3(d, t, o);
        if (type == OT_XZ) x.y = xa; else if (type == OT_XY) x.z = xa; else x.x = xa;
        return;
    }
    if (type == OT_TILT)       // n.(p0 - o) / n.d with the difference taken first and IEEE division
        t = __fdiv_rn(dot3(f3(geom.x, geom.y, geom.z), f3(aux.x - o.x, aux.y - o.y, aux.z - o.z)), dot3(f3(geom.x, geom.y, geom.z), d));
    else if (geom.w > (1.f / (float)PT_HUGE_RADIUS)) {        // small sphere
        const float rad = 1.f / geom.w;
        F3 r = f3(fmaf(d.x, t, o.x) - geom.x, fmaf(d.y, t, o.y) - geom.y, fmaf(d.z, t, o.z) - geom.z);
        const float f = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, -rad * rad)));
        const float g = 2.f * dot3(r, d);
        if (fabsf(g) > 1e-3f * rad) t -= f / g;
    }
    x = fma3(d, t, o);
}

// random_scattering: cosine-weighted (:337-348) or uniform (:351-360; weight 1 as in the reference)
template <bool UNIFORM> __device__ __forceinline__ F3 sample_hemisphere(F3 w, float xi1, float xi2)
{
    float sn, cs;
    __sincosf(fmaf(2.f * PT_PI_F, xi1, -PT_PI_F), &sn, &cs);     // angle in [-pi, pi): same distribution as 2*pi*xi
    F3 u = fabsf(w.x) > .1f ? f3(w.z, 0.f, -w.x) : f3(0.f, -w.z, w.y);   // (0,1,0) x w  or  (1,0,0) x w
    u = normalize3(u);
    F3 v = f3(w.y * u.z - w.z * u.y, w.z * u.x - w.x * u.z, w.x * u.y - w.y * u.x);
    float ru, rw;
    if (UNIFORM) { ru = sqrt_fast(xi2 * (2.f - xi2)); rw = 1.f - xi2; }
    else { ru = sqrt_fast(xi2); rw = sqrt_fast(1.f - xi2); }
    float a = cs * ru, b = sn * ru;
    return f3(fmaf(u.x, a, fmaf(v.x, b, w.x * rw)), fmaf(u.y, a, fmaf(v.y, b, w.y * rw)), fmaf(u.z, a, fmaf(v.z, b, w.z * rw)));
}

__device__ __forceinline__ void accum_add(unsigned long long *fix, unsigned int pix, F3 v)
{
    // 64-bit fixed point (2^-24): integer adds are associative => order-independent, reproducible sums
    unsigned long long *p = fix + (size_t)pix * 3;
    // (v > 0) also drops NaN; the clamp keeps a single firefly from overflowing 64 bits
    atomicAdd(p + 0, (unsigned long long)__float2ull_rn((v.x > 0.f ? fminf(v.x, 6.0e10f) : 0.f) * PT_FIX_SCALE));
    atomicAdd(p + 1, (unsigned long long)__float2ull_rn((v.y > 0.f ? fminf(v.y, 6.0e10f) : 0.f) * PT_FIX_SCALE));
    atomicAdd(p + 2, (unsigned long long)__float2ull_rn((v.z > 0.f ? fminf(v.z, 6.0e10f) : 0.f) * PT_FIX_SCALE));
}

__device__ __forceinline__ unsigned int pack_state(int depth, int prev, int E)
{
    return (unsigned int)depth | ((unsigned int)(prev + 1) << 16) | ((unsigned int)E << 31);
}

// ---------------------------------------------------------------------------------------------- the bounce kernel
// One launch = up to P.iters bounces of every path slot, state in registers:
//   load (queue) -> [ regenerate dead lanes -> bounce ] x iters -> compact survivors (queue).
// A lane whose path ended takes the next camera path at once, so lanes stay busy without a trip through the
// queue; the queue and the block-wide compaction run once per launch and matter in the tail of a render, when
// generation is exhausted and the survivors of long paths are repacked into dense warps.
// Path indices are handed out in chunks: a warp reserves P.chunk consecutive indices with ONE atomic and keeps
// what it has not used in P.warp_chunk between launches, so regeneration needs no block-wide step.
template <int MODE, bool STATS>
__global__ void __launch_bounds__(PT_BLOCK, PT_BLOCKS_PER_SM) k_bounce(const KParams P)
{
    constexpr unsigned int NW = PT_BLOCK / 32;
    __shared__ unsigned int s_warp[NW][8];   // per warp: alive, shadow, miss | trunc << 16, shaded, scatter, max depth
    __shared__ unsigned int s_base_out;
    __shared__ float4 s_sphf[2 * (PT_MAX_OBJ + 4)];    // sphere scan table, then the exact {centre, r^2} table

    const unsigned int tid = blockIdx.x * PT_BLOCK + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    // every load of the prologue is independent of the others: one L2 round trip, not three
    const unsigned int n_in = *P.n_in;
    const unsigned long long gen_seen = *(volatile unsigned long long *)P.gen_counter;
    const uint4 wc = P.warp_chunk[tid >> 5];
    const float4 qa = P.qin[0][tid], qb = P.qin[1][tid], qc = P.qin[2][tid];
    float4 ql = make_float4(0.f, 0.f, 0.f, 0.f);
    if (STATS) ql = P.qin[3][tid];
    // stage the sphere scan table (16 B per sphere) in shared memory
    if (c_scene.n_sph4 > 0) {
#pragma unroll 1
        for (int i = threadIdx.x; i < c_scene.n_sph4; i += PT_BLOCK) {
            s_sphf[i] = P.sphf[i];
            s_sphf[PT_MAX_OBJ + 4 + i] = P.sphf[PT_MAX_OBJ + 4 + i];
        }
        __syncthreads();
    }

    F3 o = f3(0, 0, 0), d = f3(0, 0, 1), T = f3(0, 0, 0), L = f3(0, 0, 0);
    unsigned int pix = 0, smp = 0;
    int depth = 0, prev = -1, E = 1;
    bool alive = false;
    unsigned int n_shadow = 0, n_miss = 0, n_trunc = 0, n_shaded = 0, n_scatter = 0, my_depth = 0;
    if (tid < n_in) {
        const unsigned int st = __float_as_uint(qc.w);
        o = f3(qa.x, qa.y, qa.z); pix = __float_as_uint(qa.w);
        d = f3(qb.x, qb.y, qb.z); smp = __float_as_uint(qb.w);
        T = f3(qc.x, qc.y, qc.z);
        depth = (int)(st & 0xFFFFu); prev = (int)((st >> 16) & 0x7FFFu) - 1; E = (int)(st >> 31);
        if (STATS) L = f3(ql.x, ql.y, ql.z);
        alive = true;
    }
    unsigned long long ck_base = (unsigned long long)wc.x | ((unsigned long long)wc.y << 32);
    unsigned int ck_left = wc.z;
    bool exhausted = gen_seen >= P.total_paths;          // monotonic counter: a stale read only delays the switch
    const int iters = !exhausted ? P.iters : (n_in < P.drain_below ? P.iters_drain : P.iters_tail);

    for (int it = 0;; it++) {
    // ---- regeneration: lanes without a path take the warp's next path indices (:528-536)
    const unsigned int b_want = __ballot_sync(0xffffffffu, !alive);
    if (b_want != 0u && (ck_left > 0u || !exhausted)) {            // warp-uniform
        const unsigned int n = __popc(b_want), r = __popc(b_want & lt);
        unsigned long long g = 0;
        bool got = false;
        if (ck_left < n && !exhausted) {
            // refill: the rest of the old chunk goes to the first lanes, the new chunk to the others
            unsigned long long g0 = 0;
            if (lane == 0) g0 = atomicAdd(P.gen_counter, (unsigned long long)P.chunk);
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            const unsigned int fresh = g0 >= P.total_paths ? 0u : (unsigned int)min((unsigned long long)P.chunk, P.total_paths - g0);
            if (fresh < P.chunk) exhausted = true;
            if (r < ck_left) { g = ck_base + r; got = true; }
            else if (r - ck_left < fresh) { g = g0 + (r - ck_left); got = true; }
            const unsigned int used = min(n - ck_left, fresh);
            ck_base = g0 + used; ck_left = fresh - used;
        } else {
            if (r < ck_left) { g = ck_base + r; got = true; }
            const unsigned int used = min(n, ck_left);
            ck_base += used; ck_left -= used;
        }
        if (!alive && got) {
            // path g -> (sample, owned pixel): pixel-major inside a sample so neighbouring lanes are neighbouring pixels
            unsigned int sidx = (unsigned int)__double2uint_rz(__ull2double_rz(g) * P.inv_owned_pixels);
            long long rr = (long long)(g - (unsigned long long)sidx * P.owned_pixels);
            if (rr < 0) { sidx--; rr += P.owned_pixels; }
            else if (rr >= (long long)P.owned_pixels) { sidx++; rr -= P.owned_pixels; }
            const unsigned int lp = (unsigned int)rr;
            unsigned int row_local, tile;
            if (P.use_magic) {               // multiply-shift division (warp-uniform branch)
                row_local = (unsigned int)(((unsigned long long)lp * P.magic_w) >> 40);
                tile = (unsigned int)(((unsigned long long)row_local * P.magic_tile) >> 40);
            } else {
                row_local = lp / (unsigned int)P.w;
                tile = row_local / (unsigned int)P.tile_rows;
            }
            const unsigned int xpix = lp - row_local * (unsigned int)P.w;
            const unsigned int y = (tile * (unsigned int)P.world + (unsigned int)P.rank) * (unsigned int)P.tile_rows
                                 + (row_local - tile * (unsigned int)P.tile_rows);
            pix = y * (unsigned int)P.w + xpix;
            smp = sidx;
            // ray generation with uniform sub-pixel jitter (:533-536)
            const uint4 rj = philox4x32_10(pix, smp, 0u, PT_DRAW_A, P.seed_lo, P.seed_hi);
            const float u = ((float)xpix - 0.5f + u01(rj.x)) * P.inv_w;                            // :533
            const float v = ((float)(P.h - 1 - (int)y) - 0.5f + u01(rj.y)) * P.inv_h;              // :534
            F3 dc = f3(fmaf(P.cam_h[0], u, fmaf(P.cam_v[0], v, P.cam_base[0])),
                       fmaf(P.cam_h[1], u, fmaf(P.cam_v[1], v, P.cam_base[1])),
                       fmaf(P.cam_h[2], u, fmaf(P.cam_v[2], v, P.cam_base[2])));                    // :276-279
            d = normalize3(dc);                                                                    // :536
            o = f3(P.cam_o[0], P.cam_o[1], P.cam_o[2]);
            T = f3(1.f, 1.f, 1.f);
            L = f3(0.f, 0.f, 0.f);
            depth = 0; prev = -1; E = 1;
            alive = true;
        }
    }
    if (it >= iters) break;
    if (!__any_sync(0xffffffffu, alive)) break;                     // nothing left to trace in this warp

    if (alive) {
        // ---- extend: closest hit (:323-335) + hittingPoint (:371-377)
        float t; int code;
        F3 Lc;                      // radiance this vertex contributes (emission, light samples): ONE accumulation per bounce
        n_shaded++;
        closest_hit(o, d, prev, s_sphf, t, code);
        F3 x;
        int on_code;
        if (code < 0) { x = f3(0.f, 0.f, 0.f); code = c_scene.code_obj0; on_code = -1; n_miss++; }   // :373-374: continue from (0,0,0) on object 0
        else on_code = code;
        const MatF32 m = P.mats[code];
        const int type = __float_as_int(m.e_type.w), refl = __float_as_int(m.c_refl.w);
        if (on_code >= 0) refine_hit(o, d, t, type, m.geom, m.aux, x);
        // ---- normal(), :118-124 / :246-253
        F3 ng;
        if (type == OT_SPHERE) ng = f3(x.x - m.geom.x, x.y - m.geom.y, x.z - m.geom.z) * m.geom.w;
        else if (type == OT_XZ) ng = f3(0.f, 1.f, 0.f);
        else if (type == OT_XY) ng = f3(0.f, 0.f, 1.f);
        else if (type == OT_YZ) ng = f3(1.f, 0.f, 0.f);
        else ng = f3(m.geom.x, m.geom.y, m.geom.z);
        const F3 nl = dot3(ng, d) < 0.f ? ng : f3(-ng.x, -ng.y, -ng.z);
        F3 f = f3(m.c_refl.x, m.c_refl.y, m.c_refl.z);
        F3 e = f3(m.e_type.x, m.e_type.y, m.e_type.z);
        if (MODE == PT_MODE_NEE_CONE_SPHERE && !E && type == OT_SPHERE) e = f3(0.f, 0.f, 0.f);
        Lc = T * e;
        // ---- Russian roulette, :447-454.  One Philox block per vertex: x -> RR (high 16 bits) and the REFR
        // branch (low 16 bits); y, z -> the two sampling uniforms of the first decision (light point or
        // hemisphere); w and the unused low bytes of y, z, w -> the hemisphere sample after an occluded light.
        const float p = f.x > f.y && f.x > f.z ? f.x : f.y > f.z ? f.y : f.z;
        depth++;
        my_depth = max(my_depth, (unsigned int)depth);
        const uint4 ra = philox4x32_10(pix, smp, (unsigned)depth, PT_DRAW_A, P.seed_lo, P.seed_hi);
        if (depth > 5 || p == 0.f) {
            if ((float)(ra.x >> 16) * (1.f / 65536.f) < p) f = f * (1.f / p);
            else alive = false;
        }
        if (alive && depth >= P.max_depth) { alive = false; n_trunc++; }
        if (alive) {
            F3 dn;
            if (refl == PT_DIFF) {
                if (MODE == PT_MODE_NEE_REF_RECT) {
                    // light_sampling (:363-369) + shadow ray (:466-467)
                    const float xl = fmaf(c_scene.lxw, u01(ra.y), c_scene.lx0), zl = fmaf(c_scene.lzw, u01(ra.z), c_scene.lz0);
                    F3 dl = normalize3(f3(xl - x.x, c_scene.ly - x.y, zl - x.z));
                    float ts; int cs;
                    n_shadow++;
                    closest_hit(x, dl, on_code, s_sphf, ts, cs);
                    if (cs == c_scene.light_code) {
                        const float pdf_inv = fabsf(c_scene.larea * dl.y / (ts * ts));   // :471
                        const float brdf = fabsf(dot3(dl, nl) * PT_INV_PI_F);            // :472
                        T = T * f * (pdf_inv * brdf);
                        const MatF32 ml = P.mats[cs];
                        const F3 fl = f3(ml.c_refl.x, ml.c_refl.y, ml.c_refl.z);
                        if (fl.x == 0.f && fl.y == 0.f && fl.z == 0.f) {
                            // the path continues along the shadow ray and ends on the (black-bodied) light:
                            // !p => one RR draw, xi < 0 is false => return e (:448-453).  Finished in place.
                            const F3 el = f3(ml.e_type.x, ml.e_type.y, ml.e_type.z);
                            Lc = Lc + T * el;
                            n_shaded++;
                            my_depth = max(my_depth, (unsigned int)depth + 1u);
                            alive = false;
                        } else {
                            dn = dl;       // general light with albedo: keep tracing from here next bounce
                        }
                    } else {
                        const unsigned int r2bits = ((ra.y & 0xFFu) << 16) | ((ra.z & 0xFFu) << 8) | (ra.w & 0xFFu);
                        dn = sample_hemisphere<false>(nl, u01(ra.w), (float)r2bits * (1.0f / 16777216.0f));   // :468
                        T = T * f;
                    }
                    E = 1;
                } else if (MODE == PT_MODE_NEE_CONE_SPHERE) {
                    dn = sample_hemisphere<false>(nl, u01(ra.y), u01(ra.z));
                    F3 esum = f3(0.f, 0.f, 0.f);
                    for (int li = 0; li < c_scene.n_lights; li++) {
                        const int lc = c_scene.light_sph_code[li];
                        const MatF32 ml = P.mats[lc];
                        F3 sw = f3(ml.geom.x - x.x, ml.geom.y - x.y, ml.geom.z - x.z);
                        const float dist2 = dot3(sw, sw), rad = 1.f / ml.geom.w;
                        if (!(dist2 > rad * rad)) continue;
                        const uint4 rl = philox4x32_10(pix, smp, (unsigned)depth, PT_DRAW_LIGHT0 + li, P.seed_lo, P.seed_hi);
                        sw = sw * rsqrtf(dist2);
                        F3 su = normalize3(fabsf(sw.x) > .1f ? f3(sw.z, 0.f, -sw.x) : f3(0.f, -sw.z, sw.y));
                        F3 sv = f3(sw.y * su.z - sw.z * su.y, sw.z * su.x - sw.x * su.z, sw.x * su.y - sw.y * su.x);
                        const float cos_a_max = sqrtf(fmaxf(0.f, 1.f - rad * rad / dist2));
                        const float eps1 = u01(rl.x), eps2 = u01(rl.y);
                        const float cos_a = 1.f - eps1 + eps1 * cos_a_max;
                        const float sin_a = sqrtf(fmaxf(0.f, 1.f - cos_a * cos_a));
                        float sp, cp;
                        __sincosf(fmaf(2.f * PT_PI_F, eps2, -PT_PI_F), &sp, &cp);
                        F3 l = normalize3(su * (cp * sin_a) + sv * (sp * sin_a) + sw * cos_a);
                        float ts; int cs;
                        n_shadow++;
                        closest_hit(x, l, on_code, s_sphf, ts, cs);
                        if (cs == lc) {
                            const float omega = 2.f * PT_PI_F * (1.f - cos_a_max);
                            const float ldn = dot3(l, nl);
                            if (ldn > 0.f) esum = esum + f * f3(ml.e_type.x, ml.e_type.y, ml.e_type.z) * (ldn * omega * PT_INV_PI_F);
                        }
                    }
                    Lc = Lc + T * esum;
                    T = T * f;
                    E = 0;
                } else {
                    dn = sample_hemisphere<MODE == PT_MODE_UNI>(nl, u01(ra.y), u01(ra.z));   // :474-477
                    T = T * f;
                    E = 1;
                }
            } else if (refl == PT_SPEC) {                                                    // :482-483
                dn = d - ng * (2.f * dot3(ng, d));
                T = T * f;
                E = 1;
            } else {                                                                         // REFR, :485-495
                const F3 rd = d - ng * (2.f * dot3(ng, d));
                const bool into = dot3(ng, nl) > 0.f;
                const float nnt = into ? (1.f / 1.5f) : 1.5f, ddn = dot3(d, nl);
                const float cos2t = 1.f - nnt * nnt * (1.f - ddn * ddn);
                T = T * f;
                E = 1;
                if (cos2t < 0.f) dn = rd;                                                    // total internal reflection
                else {
                    const F3 td = normalize3(d * nnt - ng * ((into ? 1.f : -1.f) * (ddn * nnt + sqrtf(cos2t))));
                    const float R0 = 0.04f, c = 1.f - (into ? -ddn : dot3(td, ng));
                    const float Re = R0 + (1.f - R0) * c * c * c * c * c, Tr = 1.f - Re, Pr = .25f + .5f * Re;
                    // the reference splits into both branches while depth <= 2 (:494-495); a wavefront keeps one
                    // path per slot, so the stochastic branch (:492-493) is used at every depth (same expectation).
                    if ((float)(ra.x & 0xFFFFu) * (1.f / 65536.f) < Pr) { dn = rd; T = T * (Re / Pr); }
                    else { dn = td; T = T * (Tr / (1.f - Pr)); }
                }
            }
            if (alive) { o = x; d = dn; prev = on_code; n_scatter++; }
        }
        if (STATS) {
            L = L + Lc;
            if (!alive) {       // path finished: flush its radiance and its square
                accum_add(P.fix, pix, L);
                accum_add(P.fixsq, pix, L * L);
            }
        } else if (Lc.x > 0.f || Lc.y > 0.f || Lc.z > 0.f) accum_add(P.fix, pix, Lc);
    }

    }   // bounce loop

    // ---- compaction of the survivors into the output queue: warp ballot + block prefix sum + ONE atomic per block
    const unsigned int b_alive = __ballot_sync(0xffffffffu, alive);
    {
        const unsigned int r0 = __reduce_add_sync(0xffffffffu, n_shadow), r1 = __reduce_add_sync(0xffffffffu, n_miss | (n_trunc << 16));
        const unsigned int r2 = __reduce_add_sync(0xffffffffu, n_shaded), r3 = __reduce_add_sync(0xffffffffu, n_scatter);
        const unsigned int md = __reduce_max_sync(0xffffffffu, my_depth);
        if (lane == 0) {
            s_warp[warp][0] = __popc(b_alive); s_warp[warp][1] = r0; s_warp[warp][2] = r1; s_warp[warp][3] = r2;
            s_warp[warp][4] = r3; s_warp[warp][5] = md;
            P.warp_chunk[tid >> 5] = make_uint4((unsigned int)ck_base, (unsigned int)(ck_base >> 32), ck_left, 0u);
        }
    }
    __syncthreads();
    const unsigned int cnt = lane < NW ? s_warp[lane][0] : 0u;
    const unsigned int alive_total = __reduce_add_sync(0xffffffffu, cnt), alive_before = __reduce_add_sync(0xffffffffu, lane < warp ? cnt : 0u);
    if (warp == 0) {
        const unsigned int v1 = lane < NW ? s_warp[lane][1] : 0u, v2 = lane < NW ? s_warp[lane][2] : 0u, v3 = lane < NW ? s_warp[lane][3] : 0u;
        const unsigned int v4 = lane < NW ? s_warp[lane][4] : 0u, v5 = lane < NW ? s_warp[lane][5] : 0u;
        const unsigned int s1 = __reduce_add_sync(0xffffffffu, v1), s2 = __reduce_add_sync(0xffffffffu, v2 & 0xFFFFu);
        const unsigned int s2b = __reduce_add_sync(0xffffffffu, v2 >> 16), s3 = __reduce_add_sync(0xffffffffu, v3);
        const unsigned int s4 = __reduce_add_sync(0xffffffffu, v4), s5 = __reduce_max_sync(0xffffffffu, v5);
        if (lane == 0) {
            s_base_out = alive_total ? atomicAdd(P.n_out, alive_total) : 0u;
            if (s1) atomicAdd(&P.stats->rays_shadow, (unsigned long long)s1);
            if (s2) atomicAdd(&P.stats->misses, (unsigned long long)s2);
            if (s2b) atomicAdd(&P.stats->truncated, (unsigned long long)s2b);
            if (s3) atomicAdd(&P.stats->shaded, (unsigned long long)s3);
            if (s4) atomicAdd(&P.stats->rays_scatter, (unsigned long long)s4);
            if (s5 > *(volatile unsigned int *)&P.stats->max_depth_seen) atomicMax(&P.stats->max_depth_seen, s5);
        }
    }
    __syncthreads();
    if (alive) {
        const unsigned int slot = s_base_out + alive_before + __popc(b_alive & lt);
        P.qout[0][slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pix));
        P.qout[1][slot] = make_float4(d.x, d.y, d.z, __uint_as_float(smp));
        P.qout[2][slot] = make_float4(T.x, T.y, T.z, __uint_as_float(pack_state(depth, prev, E)));
        if (STATS) P.qout[3][slot] = make_float4(L.x, L.y, L.z, 0.f);
    }
}

// fixed point -> double sums, restricted to the rows this rank owns (foreign rows stay zero)
__global__ void k_resolve(const unsigned long long *__restrict__ fix, const unsigned long long *__restrict__ fixsq,
                          double *__restrict__ sum, double *__restrict__ sumsq, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum[i] = (double)fix[i] * PT_FIX_INV;
    if (sumsq && fixsq) sumsq[i] = (double)fixsq[i] * PT_FIX_INV;
}

// pt_debug_intersect, precision 32
__global__ void k_intersect_fp32(const double *__restrict__ rays, int n_rays, double *__restrict__ t_out, int *__restrict__ id_out,
                                 const MatF32 *__restrict__ mats, const float4 *__restrict__ sphf)
{
    __shared__ float4 s_sphf[2 * (PT_MAX_OBJ + 4)];
    for (int k = threadIdx.x; k < c_scene.n_sph4; k += blockDim.x) { s_sphf[k] = sphf[k]; s_sphf[PT_MAX_OBJ + 4 + k] = sphf[PT_MAX_OBJ + 4 + k]; }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const double *r = rays + (size_t)i * 6;
    F3 o = f3((float)r[0], (float)r[1], (float)r[2]), d = f3((float)r[3], (float)r[4], (float)r[5]);
    float t; int code;
    closest_hit(o, d, -1, s_sphf, t, code);
    int id = -1;
    if (code >= 0) {
        // report the t the shading stage uses (refined once for the winning object) and the scene id
        const MatF32 m = mats[code];
        F3 x;
        refine_hit(o, d, t, __float_as_int(m.e_type.w), m.geom, m.aux, x);
        id = __float_as_int(m.aux.w);
    }
    t_out[i] = id >= 0 ? (double)t : 1e20;
    id_out[i] = id;
}

__global__ void k_philox(const uint32_t *__restrict__ ctr, const uint32_t *__restrict__ key, int n, uint32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// FFMA-only microbenchmark: 8 independent chains per thread, register operands.
__global__ void __launch_bounds__(256) k_ffma_peak(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE> void launch_bounce(bool stats, int blocks, cudaStream_t s, const KParams &P)
{
    if (stats) k_bounce<MODE, true><<<blocks, PT_BLOCK, 0, s>>>(P);
    else k_bounce<MODE, false><<<blocks, PT_BLOCK, 0, s>>>(P);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
static int ensure_queues(pt_ctx *ctx, int capacity, bool stats)
{
    if (ctx->q_capacity >= capacity && (!stats || ctx->q[0][3])) return PT_OK;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 4; b++) { if (ctx->q[a][b]) cudaFree(ctx->q[a][b]); ctx->q[a][b] = nullptr; }
    ctx->q_capacity = 0;
    // one allocation per array keeps every array 256 B aligned for the 128-bit accesses
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < (stats ? 4 : 3); b++)
            PT_CUDA(ctx, cudaMalloc(&ctx->q[a][b], sizeof(float4) * (size_t)capacity));
    if (ctx->d_warp_chunk) cudaFree(ctx->d_warp_chunk);
    ctx->d_warp_chunk = nullptr;
    PT_CUDA(ctx, cudaMalloc(&ctx->d_warp_chunk, sizeof(uint4) * (size_t)(capacity / 32)));
    ctx->q_capacity = capacity;
    return PT_OK;
}

int pt_fp32_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s)
{
    if (!ctx->fp32_ok) return pt_fail(ctx, PT_ERR_ARG, "scene does not fit the FP32 engine: " + ctx->fp32_why);
    const int w = p->width, h = p->height;
    const int tile = p->tile_rows > 0 ? p->tile_rows : 8;
    const int world = p->world > 0 ? p->world : 1;
    const bool stats = p->collect_stats != 0;
    if (p->mode == PT_MODE_NEE_CONE_SPHERE && ctx->h_scene32->n_lights == 0)
        return pt_fail(ctx, PT_ERR_ARG, "PT_MODE_NEE_CONE_SPHERE needs at least one emissive sphere");

    // rows owned by this rank
    long long owned_rows = 0;
    const int n_tiles = (h + tile - 1) / tile;
    for (int k = p->rank; k < n_tiles; k += world) owned_rows += (k * tile + tile <= h) ? tile : (h - k * tile);
    const unsigned long long owned_pixels = (unsigned long long)owned_rows * w;
    const size_t n_acc = (size_t)w * h * 3;

    // fixed-point accumulators
    if (ctx->fix_elems < n_acc) {
        if (ctx->d_fix) cudaFree(ctx->d_fix);
        if (ctx->d_fixsq) cudaFree(ctx->d_fixsq);
        ctx->d_fix = ctx->d_fixsq = nullptr; ctx->fix_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fix, n_acc * sizeof(unsigned long long)));
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fixsq, n_acc * sizeof(unsigned long long)));
        ctx->fix_elems = n_acc;
    }
    PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fix, 0, n_acc * sizeof(unsigned long long), s));
    if (stats) PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fixsq, 0, n_acc * sizeof(unsigned long long), s));

    int it_total = 0;
    if (owned_pixels > 0 && p->spp > 0) {
        const unsigned long long total = owned_pixels * (unsigned long long)p->spp;
        // queue capacity = path slots in flight = threads per launch.  Default: PT_DEFAULT_WAVES full waves of resident
        // blocks (no launch ends with a partially filled wave); the queues are touched once per launch, not per bounce.
        int cap = p->queue_capacity;
        const long long wave = (long long)ctx->sm_count * PT_BLOCKS_PER_SM * PT_BLOCK;
        if (cap <= 0) cap = (int)(PT_DEFAULT_WAVES * wave);
        {   // never more slots than paths (rounded up to whole blocks)
            const unsigned long long need = (total + PT_BLOCK - 1) / PT_BLOCK * PT_BLOCK;
            if ((unsigned long long)cap > need) cap = (int)need;
        }
        cap = (cap + PT_BLOCK - 1) / PT_BLOCK * PT_BLOCK;
        int rc = ensure_queues(ctx, cap, stats);
        if (rc) return rc;

        // per-iteration live counters (n[it]); counts[0] = cap (all dead => regenerate)
        const int max_it = 1 << 20;
        if (ctx->counts_len < max_it + 4) {
            if (ctx->d_counts) cudaFree(ctx->d_counts);
            ctx->d_counts = nullptr; ctx->counts_len = 0;
            PT_CUDA(ctx, cudaMalloc(&ctx->d_counts, sizeof(unsigned int) * (size_t)(max_it + 4)));
            ctx->counts_len = max_it + 4;
            ctx->counts_dirty = (size_t)ctx->counts_len;
        }
        // layout: [0..1] gen counter (u64), [2..] n[it]
        {   // clear only the prefix the previous render dirtied (+ slack for the speculative batch)
            size_t dirty = ctx->counts_dirty + 256;
            if (dirty > (size_t)ctx->counts_len) dirty = (size_t)ctx->counts_len;
            PT_CUDA(ctx, cudaMemsetAsync(ctx->d_counts, 0, sizeof(unsigned int) * dirty, s));
        }
        PT_CUDA(ctx, cudaMemsetAsync(ctx->d_warp_chunk, 0, sizeof(uint4) * (size_t)(cap / 32), s));   // n[0] = 0: every slot starts without a path
        PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));

        KParams P{};
        P.gen_counter = (unsigned long long *)ctx->d_counts;
        P.warp_chunk = ctx->d_warp_chunk;
        P.iters = p->bounces_per_launch > 0 ? p->bounces_per_launch : PT_DEFAULT_ITERS;
        P.iters_tail = P.iters < PT_DEFAULT_ITERS_TAIL ? P.iters : PT_DEFAULT_ITERS_TAIL;
        P.iters_drain = p->bounces_per_launch > 0 ? P.iters : PT_DEFAULT_ITERS_DRAIN;
        P.drain_below = (unsigned int)(wave / 4);
        {   // indices a warp reserves per atomic: enough for about one launch, but small renders still spread over the GPU
            unsigned long long c = total / ((unsigned long long)(cap / 32) * 2ull);
            unsigned int chunk = 32;
            while (chunk < 256 && chunk * 2ull <= c) chunk *= 2;
            P.chunk = chunk;
        }
        P.total_paths = total;
        P.owned_pixels = (unsigned int)owned_pixels;
        P.inv_owned_pixels = 1.0 / (double)owned_pixels;
        P.w = w; P.h = h; P.spp = p->spp; P.tile_rows = tile; P.rank = p->rank; P.world = world;
        P.magic_w = ((1ull << 40) + (unsigned long long)w - 1) / (unsigned long long)w;
        P.magic_tile = ((1ull << 40) + (unsigned long long)tile - 1) / (unsigned long long)tile;
        P.use_magic = (owned_pixels < (1ull << 24) && w < 65536 && tile < 65536) ? 1 : 0;
        P.max_depth = p->max_depth > 0 ? p->max_depth : 4096;
        if (P.max_depth > 65000) P.max_depth = 65000;
        const pt_camera &c = ctx->cam;
        P.cam_o[0] = (float)c.origin.x; P.cam_o[1] = (float)c.origin.y; P.cam_o[2] = (float)c.origin.z;
        P.cam_base[0] = (float)(c.lower_left_corner.x - c.origin.x);
        P.cam_base[1] = (float)(c.lower_left_corner.y - c.origin.y);
        P.cam_base[2] = (float)(c.lower_left_corner.z - c.origin.z);
        P.cam_h[0] = (float)c.horizontal.x; P.cam_h[1] = (float)c.horizontal.y; P.cam_h[2] = (float)c.horizontal.z;
        P.cam_v[0] = (float)c.vertical.x; P.cam_v[1] = (float)c.vertical.y; P.cam_v[2] = (float)c.vertical.z;
        P.inv_w = 1.f / (float)w; P.inv_h = 1.f / (float)h;
        P.seed_lo = (unsigned int)p->seed; P.seed_hi = (unsigned int)(p->seed >> 32);
        P.fix = ctx->d_fix; P.fixsq = ctx->d_fixsq;
        P.mats = ctx->d_mats; P.sphf = ctx->d_sphf; P.stats = ctx->d_stats;

        const int blocks = cap / PT_BLOCK;
        unsigned int *n_it = ctx->d_counts + 2;
        // Termination check without draining the pipeline: batch k+1 is enqueued before the live count
        // after batch k is read back (pinned slot + event per parity).
        unsigned int *h_n = ctx->h_pinned;                 // pinned slots + events live in the context
        cudaEvent_t *evb = ctx->ev_batch;
        int it = 0, nb = 0, rc2 = PT_OK;
        const int batch = 4;
        bool done = false;
        while (!done) {
            for (int b = 0; b < batch && rc2 == PT_OK; b++, it++) {
                if (it >= max_it) { rc2 = pt_fail(ctx, PT_ERR_STATE, "wavefront iteration limit reached"); break; }
                for (int a = 0; a < 4; a++) { P.qin[a] = ctx->q[it & 1][a]; P.qout[a] = ctx->q[(it + 1) & 1][a]; }
                P.n_in = n_it + it; P.n_out = n_it + it + 1;
                switch (p->mode) {
                case PT_MODE_NEE_REF_RECT: launch_bounce<PT_MODE_NEE_REF_RECT>(stats, blocks, s, P); break;
                case PT_MODE_COS: launch_bounce<PT_MODE_COS>(stats, blocks, s, P); break;
                case PT_MODE_UNI: launch_bounce<PT_MODE_UNI>(stats, blocks, s, P); break;
                default: launch_bounce<PT_MODE_NEE_CONE_SPHERE>(stats, blocks, s, P); break;
                }
                ctx->stats.kernel_launches++;
            }
            if (rc2 != PT_OK) break;
            cudaMemcpyAsync(h_n + (nb & 1), n_it + it, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
            cudaEventRecord(evb[nb & 1], s);
            if (nb > 0) {
                cudaEventSynchronize(evb[(nb - 1) & 1]);
                if (h_n[(nb - 1) & 1] == 0) done = true;
            }
            nb++;
            cudaError_t e_ = cudaGetLastError();
            if (e_ != cudaSuccess) { rc2 = pt_fail(ctx, PT_ERR_CUDA, std::string("k_bounce: ") + cudaGetErrorString(e_)); break; }
        }
        if (rc2 != PT_OK) { cudaStreamSynchronize(s); return rc2; }
        ctx->stats.iterations = (uint64_t)it;
        ctx->counts_dirty = (size_t)it + 4;
        it_total = it;
    }
    k_resolve<<<(unsigned)((n_acc + 255) / 256), 256, 0, s>>>(ctx->d_fix, stats ? ctx->d_fixsq : nullptr, d_sum, stats ? d_sumsq : nullptr, n_acc);
    ctx->stats.kernel_launches++;
    PT_CUDA(ctx, cudaGetLastError());
    ctx->stats.queue_slots_io = 0;
    if (it_total > 0) {   // queue traffic of this render: launch k reads n[k] slots and writes n[k+1]
        std::vector<unsigned int> h(it_total + 1);
        PT_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->d_counts + 2, sizeof(unsigned int) * (size_t)(it_total + 1), cudaMemcpyDeviceToHost, s));
        PT_CUDA(ctx, cudaStreamSynchronize(s));
        uint64_t io = 0;
        for (int k = 0; k <= it_total; k++) io += (k == 0 || k == it_total) ? h[k] : 2ull * h[k];
        ctx->stats.queue_slots_io = io;
    }
    return PT_OK;
}

int pt_fp32_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s)
{
    if (!ctx->fp32_ok) return pt_fail(ctx, PT_ERR_ARG, "scene does not fit the FP32 engine: " + ctx->fp32_why);
    PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));
    k_intersect_fp32<<<(n + 127) / 128, 128, 0, s>>>(d_rays, n, d_t, d_id, ctx->d_mats, ctx->d_sphf);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp32_philox(pt_ctx *ctx, const uint32_t *d_ctr, const uint32_t *d_key, int n, uint32_t *d_out, cudaStream_t s)
{
    k_philox<<<(n + 127) / 128, 128, 0, s>>>(d_ctr, d_key, n, d_out);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp32_ffma_peak(pt_ctx *ctx, double *tflops, double *mhz)
{
    const int blocks = ctx->sm_count * 8, iters = 4096;
    float *d_out = nullptr;
    PT_CUDA(ctx, cudaMalloc(&d_out, sizeof(float) * blocks * 256));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k_ffma_peak<<<blocks, 256, 0, ctx->stream>>>(d_out, iters, 1.0001f, 0.0001f);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, ctx->stream);
        k_ffma_peak<<<blocks, 256, 0, ctx->stream>>>(d_out, iters, 1.0001f, 0.0001f);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    PT_CUDA(ctx, cudaGetLastError());
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * 256;
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *mhz = khz / 1000.0;
    return PT_OK;
}
