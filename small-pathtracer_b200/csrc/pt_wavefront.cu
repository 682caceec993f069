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
#define PT_BLOCKS_PER_SM (1280 / PT_BLOCK)   /* 5 blocks of 256: 48 registers/thread, measured 3 % faster than 4 */
#endif

struct KParams {
    float4 *qin[4];
    float4 *qout[4];
    const unsigned int *n_in;          // live count of the input queue (this iteration)
    unsigned int *n_out;               // survivors (next iteration), zero before the launch
    unsigned long long *gen_counter;   // next path index to generate
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
__device__ __forceinline__ F3 normalize3(F3 a) { return a * rsqrtf(dot3(a, a)); }

__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

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
__device__ PT_HIT_INLINE void closest_hit(F3 o, F3 d, int prev, float &t_out, int &code_out)
{
    float best = 1e20f;
    int code = -1;
    const float ix = rcp_fast(d.x), iy = rcp_fast(d.y), iz = rcp_fast(d.z);
    rects_axis<0>(o.y, iy, o.x, d.x, o.z, d.z, best, code);   // XZ: plane y
    rects_axis<1>(o.z, iz, o.x, d.x, o.y, d.y, best, code);   // XY: plane z
    rects_axis<2>(o.x, ix, o.y, d.y, o.z, d.z, best, code);   // YZ: plane x

    // Sphere::intersect, :229-239, eps = 1e-4.  det = r^2 - |op - b d|^2 (perpendicular distance form).
    const int ns = c_scene.n_sph;
    const int prev_s = prev - c_scene.code_sph0;
#pragma unroll 4
    for (int i = 0; i < ns; i++) {
        const float4 s = c_scene.sph[i];
        F3 op = f3(s.x - o.x, s.y - o.y, s.z - o.z);
        float b = dot3(op, d);
        F3 l = f3(fmaf(-b, d.x, op.x), fmaf(-b, d.y, op.y), fmaf(-b, d.z, op.z));
        float det = s.w - dot3(l, l);
        if (det >= 0.f) {
            float sq = sqrt_fast(det);
            float t0 = b - sq, t1 = b + sq;
            float tt = t0 > PT_EPS_F ? t0 : t1;
            if (i == prev_s) tt = b + b;          // origin on this sphere: roots are exactly {0, 2b}
            if (tt > PT_EPS_F && tt < best) { best = tt; code = c_scene.code_sph0 + i; }
        }
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
__device__ __forceinline__ float refine_t(F3 o, F3 d, float t, int type, float4 geom, float4 aux)
{
    if (type == OT_XZ) return __fdiv_rn((geom.x - o.y) + geom.y, d.y);
    if (type == OT_XY) return __fdiv_rn((geom.x - o.z) + geom.y, d.z);
    if (type == OT_YZ) return __fdiv_rn((geom.x - o.x) + geom.y, d.x);
    if (type == OT_TILT)       // n.(p0 - o) / n.d with the difference taken first and IEEE division
        return __fdiv_rn(dot3(f3(geom.x, geom.y, geom.z), f3(aux.x - o.x, aux.y - o.y, aux.z - o.z)), dot3(f3(geom.x, geom.y, geom.z), d));
    if (type == OT_SPHERE && geom.w > (1.f / (float)PT_HUGE_RADIUS)) {
        const float rad = 1.f / geom.w;
        F3 r = f3(fmaf(d.x, t, o.x) - geom.x, fmaf(d.y, t, o.y) - geom.y, fmaf(d.z, t, o.z) - geom.z);
        const float f = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, -rad * rad)));
        const float g = 2.f * dot3(r, d);
        if (fabsf(g) > 1e-3f * rad) t -= f / g;
    }
    return t;
}

__device__ __forceinline__ F3 hit_point(F3 o, F3 d, float t, int type)
{
    F3 x = fma3(d, t, o);
    if (type == OT_XZ) x.y = __fadd_rn(o.y, __fmul_rn(d.y, t));
    else if (type == OT_XY) x.z = __fadd_rn(o.z, __fmul_rn(d.z, t));
    else if (type == OT_YZ) x.x = __fadd_rn(o.x, __fmul_rn(d.x, t));
    return x;
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
template <int MODE, bool STATS>
__global__ void __launch_bounds__(PT_BLOCK, PT_BLOCKS_PER_SM) k_bounce(const KParams P)
{
    __shared__ unsigned int s_alive[PT_BLOCK / 32], s_want[PT_BLOCK / 32];
    __shared__ unsigned int s_base_out, s_regen_ok;
    __shared__ unsigned long long s_base_gen;
    __shared__ unsigned int s_stat[6];   // shadow, miss, truncated, shaded (+inline), scatter, max depth

    const unsigned int tid = blockIdx.x * PT_BLOCK + threadIdx.x;
    const unsigned int n_in = *P.n_in;
    if (blockIdx.x * PT_BLOCK >= n_in) return;          // whole block beyond the queue
    if (threadIdx.x < 6) s_stat[threadIdx.x] = 0;
    __syncthreads();
    const bool have = tid < n_in;

    F3 o = f3(0, 0, 0), d = f3(0, 0, 1), T = f3(0, 0, 0), L = f3(0, 0, 0);
    unsigned int pix = 0, smp = 0;
    int depth = 0, prev = -1, E = 1;
    bool alive = false;
    unsigned int n_shadow = 0, n_miss = 0, n_trunc = 0, n_inline = 0, n_shaded = 0, my_depth = 0;

    if (have) {
        const float4 c = P.qin[2][tid];
        const unsigned int st = __float_as_uint(c.w);
        if ((st & 0xFFFFu) != PT_DEPTH_DEAD) {
            const float4 a = P.qin[0][tid], b = P.qin[1][tid];
            o = f3(a.x, a.y, a.z); pix = __float_as_uint(a.w);
            d = f3(b.x, b.y, b.z); smp = __float_as_uint(b.w);
            T = f3(c.x, c.y, c.z);
            depth = (int)(st & 0xFFFFu); prev = (int)((st >> 16) & 0x7FFFu) - 1; E = (int)(st >> 31);
            if (STATS) { const float4 l4 = P.qin[3][tid]; L = f3(l4.x, l4.y, l4.z); }
            alive = true;
        }
    }

    if (alive) {
        // ---- extend: closest hit (:323-335) + hittingPoint (:371-377)
        float t; int code;
        n_shaded = 1;
        closest_hit(o, d, prev, t, code);
        F3 x;
        int on_code;
        if (code < 0) { x = f3(0.f, 0.f, 0.f); code = c_scene.code_obj0; on_code = -1; n_miss++; }   // :373-374: continue from (0,0,0) on object 0
        else on_code = code;
        const MatF32 m = P.mats[code];
        const int type = __float_as_int(m.e_type.w), refl = __float_as_int(m.c_refl.w);
        if (on_code >= 0) { t = refine_t(o, d, t, type, m.geom, m.aux); x = hit_point(o, d, t, type); }
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
        if (e.x > 0.f || e.y > 0.f || e.z > 0.f) {
            if (STATS) L = L + T * e; else accum_add(P.fix, pix, T * e);
        }
        // ---- Russian roulette, :447-454.  One Philox block per vertex: x -> RR (high 16 bits) and the REFR
        // branch (low 16 bits); y, z -> the two sampling uniforms of the first decision (light point or
        // hemisphere); w and the unused low bytes of y, z, w -> the hemisphere sample after an occluded light.
        const float p = f.x > f.y && f.x > f.z ? f.x : f.y > f.z ? f.y : f.z;
        depth++;
        my_depth = depth;
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
                    closest_hit(x, dl, on_code, ts, cs);
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
                            if (STATS) L = L + T * el; else accum_add(P.fix, pix, T * el);
                            n_inline++;
                            my_depth = depth + 1;
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
                        closest_hit(x, l, on_code, ts, cs);
                        if (cs == lc) {
                            const float omega = 2.f * PT_PI_F * (1.f - cos_a_max);
                            const float ldn = dot3(l, nl);
                            if (ldn > 0.f) esum = esum + f * f3(ml.e_type.x, ml.e_type.y, ml.e_type.z) * (ldn * omega * PT_INV_PI_F);
                        }
                    }
                    if (esum.x > 0.f || esum.y > 0.f || esum.z > 0.f) {
                        if (STATS) L = L + T * esum; else accum_add(P.fix, pix, T * esum);
                    }
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
            if (alive) { o = x; d = dn; prev = on_code; }
        }
        if (!alive && STATS) {
            // path finished: flush its radiance and its square
            accum_add(P.fix, pix, L);
            accum_add(P.fixsq, pix, L * L);
        }
    }

    // ---- regeneration + compaction in ONE block-wide phase.
    // Survivors are ranked by warp ballot + block prefix sum; lanes whose path ended ("want") are ranked the same
    // way and get consecutive new path indices from ONE 64-bit atomic per block; the block's output slots come from
    // ONE 32-bit atomic per block: survivors first, regenerated paths behind them.
    const bool want = have && !alive;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int b_alive = __ballot_sync(0xffffffffu, alive), b_want = __ballot_sync(0xffffffffu, want);
    {   // counters: two warp reductions + a max, folded into the same barrier phase (scatter rays = survivors)
        const unsigned int pa = n_shadow | (n_miss << 16), pb = n_trunc | ((n_inline + n_shaded) << 16);
        const unsigned int ra_ = __reduce_add_sync(0xffffffffu, pa), rb_ = __reduce_add_sync(0xffffffffu, pb);
        const unsigned int md = __reduce_max_sync(0xffffffffu, my_depth);
        if (lane == 0) {
            s_alive[warp] = __popc(b_alive); s_want[warp] = __popc(b_want);
            if (ra_ & 0xFFFFu) atomicAdd(&s_stat[0], ra_ & 0xFFFFu);
            if (ra_ >> 16) atomicAdd(&s_stat[1], ra_ >> 16);
            if (rb_ & 0xFFFFu) atomicAdd(&s_stat[2], rb_ & 0xFFFFu);
            if (rb_ >> 16) atomicAdd(&s_stat[3], rb_ >> 16);
            atomicMax(&s_stat[5], md);
        }
    }
    __syncthreads();
    unsigned int alive_before = 0, alive_total = 0, want_before = 0, want_total = 0;
#pragma unroll
    for (int i = 0; i < PT_BLOCK / 32; i++) {
        const unsigned int ca = s_alive[i], cw = s_want[i];
        alive_before += (i < (int)warp) ? ca : 0u; alive_total += ca;
        want_before += (i < (int)warp) ? cw : 0u; want_total += cw;
    }
    if (threadIdx.x == 0) {
        unsigned long long g0 = 0;
        unsigned int ok = 0;
        if (want_total) {
            g0 = atomicAdd(P.gen_counter, (unsigned long long)want_total);
            ok = g0 >= P.total_paths ? 0u : (unsigned int)min((unsigned long long)want_total, P.total_paths - g0);
        }
        s_base_gen = g0;
        s_regen_ok = ok;
        const unsigned int n_out = alive_total + ok;
        s_base_out = n_out ? atomicAdd(P.n_out, n_out) : 0u;
        if (s_stat[0]) atomicAdd(&P.stats->rays_shadow, (unsigned long long)s_stat[0]);
        if (s_stat[1]) atomicAdd(&P.stats->misses, (unsigned long long)s_stat[1]);
        if (s_stat[2]) atomicAdd(&P.stats->truncated, (unsigned long long)s_stat[2]);
        if (s_stat[3]) atomicAdd(&P.stats->shaded, (unsigned long long)s_stat[3]);
        if (alive_total) atomicAdd(&P.stats->rays_scatter, (unsigned long long)alive_total);
        if (s_stat[5] > 0) atomicMax(&P.stats->max_depth_seen, s_stat[5]);
    }
    __syncthreads();
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int rank_alive = alive_before + __popc(b_alive & lt);
    const unsigned int rank_want = want_before + __popc(b_want & lt);
    unsigned int slot = s_base_out + rank_alive;
    if (want && rank_want < s_regen_ok) {
        // ---- ray generation with uniform sub-pixel jitter (:533-536)
        // path g -> (sample, owned pixel): pixel-major inside a sample so neighbouring lanes are neighbouring pixels
        const unsigned long long g = s_base_gen + rank_want;
        unsigned int sidx = (unsigned int)__double2uint_rz(__ull2double_rz(g) * P.inv_owned_pixels);
        long long r = (long long)(g - (unsigned long long)sidx * P.owned_pixels);
        if (r < 0) { sidx--; r += P.owned_pixels; }
        else if (r >= (long long)P.owned_pixels) { sidx++; r -= P.owned_pixels; }
        const unsigned int lp = (unsigned int)r;
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
        slot = s_base_out + alive_total + rank_want;
    }
    if (alive) {
        P.qout[0][slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pix));
        P.qout[1][slot] = make_float4(d.x, d.y, d.z, __uint_as_float(smp));
        P.qout[2][slot] = make_float4(T.x, T.y, T.z, __uint_as_float(pack_state(depth, prev, E)));
        if (STATS) P.qout[3][slot] = make_float4(L.x, L.y, L.z, 0.f);
    }

}

// marks every slot of a queue as dead (depth field = 0xFFFF) so the first bounce regenerates it
__global__ void k_mark_dead(float4 *q2, unsigned int n)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q2[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(PT_DEPTH_DEAD));
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
                                 const MatF32 *__restrict__ mats)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const double *r = rays + (size_t)i * 6;
    F3 o = f3((float)r[0], (float)r[1], (float)r[2]), d = f3((float)r[3], (float)r[4], (float)r[5]);
    float t; int code;
    closest_hit(o, d, -1, t, code);
    int id = -1;
    if (code >= 0) {
        // report the t the shading stage uses (refined once for the winning object) and the scene id
        const MatF32 m = mats[code];
        t = refine_t(o, d, t, __float_as_int(m.e_type.w), m.geom, m.aux);
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

    if (owned_pixels > 0 && p->spp > 0) {
        const unsigned long long total = owned_pixels * (unsigned long long)p->spp;
        // queue capacity: default keeps both ping-pong queues inside L2 (48 B/path/queue)
        int cap = p->queue_capacity;
        if (cap <= 0) {
            // whole waves of resident blocks (4 blocks of 256 threads per SM, __launch_bounds__(256, 4)) so that no
            // launch ends with a partially filled wave, as many as keep both queues within 3/4 of L2
            const long long wave = (long long)ctx->sm_count * PT_BLOCKS_PER_SM * PT_BLOCK;
            long long budget = (long long)ctx->l2_bytes * 3 / 4;
            if (budget <= 0) budget = 64ll << 20;
            long long waves = budget / (2 * (stats ? 64 : 48)) / wave;
            if (waves < 2) waves = 2;
            cap = (int)(waves * wave);
        }
        if ((unsigned long long)cap > total) cap = (int)total;
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
        unsigned int cap_u = (unsigned int)cap;
        PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_counts + 2, &cap_u, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
        PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));
        k_mark_dead<<<(cap + 255) / 256, 256, 0, s>>>(ctx->q[0][2], cap_u);
        ctx->stats.kernel_launches++;

        KParams P{};
        P.gen_counter = (unsigned long long *)ctx->d_counts;
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
        P.mats = ctx->d_mats; P.stats = ctx->d_stats;

        const int blocks = cap / PT_BLOCK;
        unsigned int *n_it = ctx->d_counts + 2;
        // Termination check without draining the pipeline: batch k+1 is enqueued before the live count
        // after batch k is read back (pinned slot + event per parity).
        unsigned int *h_n = ctx->h_pinned;                 // pinned slots + events live in the context
        cudaEvent_t *evb = ctx->ev_batch;
        int it = 0, nb = 0, rc2 = PT_OK;
        const int batch = 32;
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
    }
    k_resolve<<<(unsigned)((n_acc + 255) / 256), 256, 0, s>>>(ctx->d_fix, stats ? ctx->d_fixsq : nullptr, d_sum, stats ? d_sumsq : nullptr, n_acc);
    ctx->stats.kernel_launches++;
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp32_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s)
{
    if (!ctx->fp32_ok) return pt_fail(ctx, PT_ERR_ARG, "scene does not fit the FP32 engine: " + ctx->fp32_why);
    PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));
    k_intersect_fp32<<<(n + 127) / 128, 128, 0, s>>>(d_rays, n, d_t, d_id, ctx->d_mats);
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
