// pt_validate.cu — FP64 validation engine (PT_ENGINE_FP64_ERAND48).
//
// Replays the reference bit by bit: per-row erand48 Xi stream (src/smallpt.cpp:530,
// src/utilities.h:26-51), FP64 arithmetic in the reference's left-to-right order, every quirk of
// SURVEY Appendix C.  COMPILED WITH -fmad=false: no FMA contraction anywhere in this file.
//
// Mapping: ONE WARP PER IMAGE ROW.  The row's stream is consumed strictly sequentially, exactly like
// the reference's single thread; all 32 lanes carry identical path state and execute the identical
// instruction stream, and only the primitive loop of intersect() (:323-335) is shared out — lane l
// tests objects l, l+32, ... and a shuffle arg-min (lowest index wins ties, like the strict `<` of
// :328) gives every lane the same (t, id).  With <= 32 objects the lane's object lives in registers,
// so the hot loop touches no memory.  Rows are scheduled dynamically by the hardware block scheduler
// (one block = one row), the analogue of `schedule(dynamic, 1)` at :526.
#include <math_constants.h>

#include "pt_internal.h"
#include "ptb200_detmath.h"

namespace {

#define PT_PI 3.14159265358979323846   /* M_PI */

struct V { double x, y, z; };
__device__ __forceinline__ V v3(double x, double y, double z) { V r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V vadd(V a, V b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }            // :31
__device__ __forceinline__ V vsub(V a, V b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }            // :34
__device__ __forceinline__ V vscale(V a, double b) { return v3(a.x * b, a.y * b, a.z * b); }           // :37
__device__ __forceinline__ V vmult(V a, V b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }           // :47
__device__ __forceinline__ V vnorm(V a) { return vscale(a, 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }  // :50
__device__ __forceinline__ double vdot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // :53
__device__ __forceinline__ V vcross(V a, V b) {                                                         // :56
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V ld3(const double *p) { return v3(p[0], p[1], p[2]); }

// erand48 (src/utilities.h:26-51): the three 16-bit limbs are one 48-bit LCG state; the value
// ldexp(x0,-48)+ldexp(x1,-32)+ldexp(x2,-16) is exactly X * 2^-48.
struct Rng48 {
    unsigned long long x;
    __device__ __forceinline__ double next()
    {
        x = (x * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
        return (double)x * (1.0 / 281474976710656.0);
    }
};

struct LaneObj { int type; double g0, g1, g2, g3, g4; };

// per-primitive intersect — Sphere :229-239, Rectangle_xz :102-112, _xy :145-155, _yz :188-198
__device__ __forceinline__ double obj_intersect(int type, double g0, double g1, double g2, double g3, double g4,
                                                const DevObj64 *full, V o, V d)
{
    switch (type) {
    case OT_SPHERE: {
        V op = vsub(v3(g1, g2, g3), o);
        double t, eps = 1e-4;
        double b = vdot(op, d);
        double det = b * b - vdot(op, op) + g0 * g0;
        if (det < 0) return 0; else det = sqrt(det);
        return (t = b - det) > eps ? t : ((t = b + det) > eps ? t : 0);
    }
    case OT_XZ: {
        double t = (g4 - o.y) / d.y;
        float x = (float)(o.x + d.x * t);
        float z = (float)(o.z + d.z * t);
        if (x < g0 || x > g1 || z < g2 || z > g3 || t < 0) return 0;
        return t;
    }
    case OT_XY: {
        double t = (g4 - o.z) / d.z;
        float x = (float)(o.x + d.x * t);
        float y = (float)(o.y + d.y * t);
        if (x < g0 || x > g1 || y < g2 || y > g3 || t < 0) return 0;
        return t;
    }
    case OT_YZ: {
        double t = (g4 - o.x) / d.x;
        float y = (float)(o.y + d.y * t);
        float z = (float)(o.z + d.z * t);
        if (y < g0 || y > g1 || z < g2 || z > g3 || t < 0) return 0;
        return t;
    }
    default: {   // tilted bounded plane (not in the reference source)
        V n = ld3(full->n), p0 = ld3(full->p0);
        double denom = vdot(n, d);
        double tau = vdot(n, vsub(p0, o)) / denom;
        V rel = vsub(vadd(o, vscale(d, tau)), p0);
        double a = vdot(rel, ld3(full->s)), b = vdot(rel, ld3(full->t));
        if (!(fabs(a) <= full->hs) || !(fabs(b) <= full->ht) || !(tau > 1e-4)) return 0;
        return tau;
    }
    }
}

// intersect(Ray,t,id), :323-335, shared out over the warp.  Returns hit?; id untouched on a miss.
__device__ __forceinline__ bool warp_intersect(const DevObj64 *objs, int n, const LaneObj &mine, int lane,
                                               V o, V d, double &t, int &id)
{
    double best = 1e20;
    int bid = 0x7fffffff;
    if (lane < n) {
        double dd = obj_intersect(mine.type, mine.g0, mine.g1, mine.g2, mine.g3, mine.g4, objs + lane, o, d);
        if (dd != 0 && dd < best) { best = dd; bid = lane; }     // `(d = ...) && d < t` — NaN is truthy but never < t
    }
    for (int i = lane + 32; i < n; i += 32) {
        const DevObj64 *ob = objs + i;
        double dd = obj_intersect(ob->type, ob->g[0], ob->g[1], ob->g[2], ob->g[3], ob->g[4], ob, o, d);
        if (dd != 0 && dd < best) { best = dd; bid = i; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bid, off);
        if (ob < best || (ob == best && oi < bid)) { best = ob; bid = oi; }
    }
    t = best;
    if (best < 1e20) { id = bid; return true; }
    return false;
}

template <bool DET> __device__ __forceinline__ void sincos_sel(double a, double *s, double *c)
{
    if (DET) pt_det_sincos(a, s, c);
    else { *c = cos(a); *s = sin(a); }
}

// random_scattering, cosine :337-348 / uniform :351-360
template <bool DET> __device__ __forceinline__ V random_scattering(int mode, V nl, Rng48 &rng)
{
    double r1 = 2 * PT_PI * rng.next();
    double r2 = rng.next();
    double sn, cs;
    sincos_sel<DET>(r1, &sn, &cs);
    V w = nl;
    V u = vnorm(vcross(fabs(w.x) > .1 ? v3(0, 1, 0) : v3(1, 0, 0), w));
    V v = vcross(w, u);
    if (mode == PT_MODE_UNI) {
        double q = sqrt(r2 * (2 - r2));
        return vnorm(vadd(vadd(vscale(vscale(u, cs), q), vscale(vscale(v, sn), q)), vscale(w, 1 - r2)));
    }
    double r2s = sqrt(r2);
    return vnorm(vadd(vadd(vscale(vscale(u, cs), r2s), vscale(vscale(v, sn), r2s)), vscale(w, sqrt(1 - r2))));
}

struct Pending { V o, d, T; int depth, E, kind; };

struct RowStats {
    unsigned long long paths, rays_camera, rays_scatter, rays_shadow, shaded, misses, truncated;
    unsigned int max_depth;
};

// radiance(), :419-496 without the dead RL block :424-442, unrolled into a loop: the recursion
// e + f (.) radiance(next) * PDF_inverse * BRDF becomes L += T (.) e; T = ((T (.) f) * PDF_inverse) * BRDF.
// (Same real number; the association differs from the recursion in the last bits only.)  The two-way
// REFR split for depth <= 2 (:494-495) uses a 4-entry stack, reflection subtree first.
template <bool DET>
__device__ V radiance(const DevObj64 *objs, int n, const LaneObj &mine, int lane, int mode, int max_depth,
                      const pt_light &light, V ro, V rd, Rng48 &rng, RowStats &st)
{
    V L = v3(0, 0, 0);
    Pending stack[4];
    int sp = 0;
    stack[sp].o = ro; stack[sp].d = rd; stack[sp].T = v3(1, 1, 1);
    stack[sp].depth = 0; stack[sp].E = 1; stack[sp].kind = 0;
    sp++;
    while (sp > 0) {
        sp--;
        V o = stack[sp].o, d = stack[sp].d, T = stack[sp].T;
        int depth = stack[sp].depth, E = stack[sp].E, kind = stack[sp].kind;
        for (;;) {
            int id = 0;                                                       // :421
            double t;
            V x;
            if (kind == 0) st.rays_camera++; else if (kind == 1) st.rays_scatter++;
            if (!warp_intersect(objs, n, mine, lane, o, d, t, id)) { x = v3(0, 0, 0); st.misses++; }   // :373-374
            else x = vadd(o, vscale(d, t));                                   // :375
            const DevObj64 *ob = objs + id;
            int type = ob->type;
            V ng;
            switch (type) {                                                   // normal(): :118-124,161-167,204-210,246-253
            case OT_SPHERE: ng = vnorm(vsub(x, v3(ob->g[1], ob->g[2], ob->g[3]))); break;
            case OT_XZ: ng = v3(0, 1, 0); break;
            case OT_XY: ng = v3(0, 0, 1); break;
            case OT_YZ: ng = v3(1, 0, 0); break;
            default: ng = ld3(ob->n); break;
            }
            V nl = vdot(ng, d) < 0 ? ng : v3(ng.x * -1, ng.y * -1, ng.z * -1);
            V f = ld3(ob->c), e = ld3(ob->e);                                 // :446
            st.shaded++;
            if (mode == PT_MODE_NEE_CONE_SPHERE && !E && type == OT_SPHERE) e = v3(0, 0, 0);
            L = vadd(L, vmult(T, e));
            double p = f.x > f.y && f.x > f.z ? f.x : f.y > f.z ? f.y : f.z;  // :447
            bool dead = false;
            if (++depth > 5 || !p) {                                          // :448
                if (rng.next() < p) f = vscale(f, 1 / p);
                else dead = true;
            }
            if ((unsigned)depth > st.max_depth) st.max_depth = depth;
            if (dead) break;
            if (depth >= max_depth) { st.truncated++; break; }
            int refl = ob->refl;
            if (refl == PT_DIFF) {                                            // :457
                double PDF_inverse = 1, BRDF = 1;
                V dn;
                if (mode == PT_MODE_NEE_REF_RECT) {                           // :464
                    double x_light = light.x0 + light.xw * rng.next();        // :365 (P2)
                    double z_light = light.z0 + light.zw * rng.next();        // :366 (P2)
                    dn = vsub(v3(x_light, light.y, z_light), x);              // :367
                    dn = vnorm(dn);                                           // :466
                    st.rays_shadow++;
                    double ts;
                    int ids = id;                                             // id is reused by :466 and stays put on a miss
                    warp_intersect(objs, n, mine, lane, x, dn, ts, ids);
                    if (ids != light.id) {                                    // :467
                        dn = random_scattering<DET>(mode, nl, rng);           // :468
                        dn = vnorm(dn);                                       // :469
                        kind = 1;
                    } else {
                        dn = vnorm(dn);
                        PDF_inverse = fabs((light.area * vdot(dn, v3(0, 1, 0))) / (ts * ts));   // :471
                        dn = vnorm(dn);
                        BRDF = fabs(vdot(dn, nl) / PT_PI);                    // :472
                        kind = 2;                                             // continues along the shadow ray
                    }
                    dn = vnorm(dn);                                           // :479
                    T = vscale(vscale(vmult(T, f), PDF_inverse), BRDF);
                    E = 1;
                } else if (mode == PT_MODE_NEE_CONE_SPHERE) {                 // not in the reference source
                    dn = random_scattering<DET>(mode, nl, rng);
                    V esum = v3(0, 0, 0);
                    for (int i = 0; i < n; i++) {
                        const DevObj64 *s = objs + i;
                        if (s->type != OT_SPHERE) continue;
                        if (s->e[0] <= 0 && s->e[1] <= 0 && s->e[2] <= 0) continue;
                        V sp_ = v3(s->g[1], s->g[2], s->g[3]);
                        double rad = s->g[0];
                        V sw = vsub(sp_, x);
                        double dist2 = vdot(sw, sw);
                        double eps1 = rng.next(), eps2 = rng.next();
                        if (!(dist2 > rad * rad)) continue;
                        sw = vscale(sw, 1 / sqrt(dist2));
                        V su = vnorm(vcross(fabs(sw.x) > .1 ? v3(0, 1, 0) : v3(1, 0, 0), sw));
                        V sv = vcross(sw, su);
                        double cos_a_max = sqrt(1 - rad * rad / dist2);
                        double cos_a = 1 - eps1 + eps1 * cos_a_max;
                        double sin_a = sqrt(1 - cos_a * cos_a);
                        double phi = 2 * PT_PI * eps2, sphi, cphi;
                        sincos_sel<DET>(phi, &sphi, &cphi);
                        V l = vnorm(vadd(vadd(vscale(su, cphi * sin_a), vscale(sv, sphi * sin_a)), vscale(sw, cos_a)));
                        double ts;
                        int ids = -1;
                        st.rays_shadow++;
                        if (warp_intersect(objs, n, mine, lane, x, l, ts, ids) && ids == i) {
                            double omega = 2 * PT_PI * (1 - cos_a_max);
                            double ldn = vdot(l, nl);
                            if (ldn > 0) esum = vadd(esum, vscale(vmult(f, vscale(ld3(s->e), ldn * omega)), 1 / PT_PI));
                        }
                    }
                    L = vadd(L, vmult(T, esum));
                    T = vmult(T, f);
                    E = 0;
                    kind = 1;
                } else {                                                      // :474-477
                    dn = random_scattering<DET>(mode, nl, rng);
                    dn = vnorm(dn);                                           // :476
                    dn = vnorm(dn);                                           // :479
                    T = vscale(vscale(vmult(T, f), PDF_inverse), BRDF);
                    E = 1;
                    kind = 1;
                }
                o = x; d = dn;
                continue;
            }
            if (refl == PT_SPEC) {                                            // :482-483
                d = vsub(d, vscale(ng, 2 * vdot(ng, d)));
                o = x; T = vmult(T, f); E = 1; kind = 1;
                continue;
            }
            // REFR, :485-495
            V refl_d = vsub(d, vscale(ng, 2 * vdot(ng, d)));
            bool into = vdot(ng, nl) > 0;
            double nc = 1, nt = 1.5, nnt = into ? nc / nt : nt / nc, ddn = vdot(d, nl), cos2t;
            if ((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0) {              // total internal reflection
                o = x; d = refl_d; T = vmult(T, f); E = 1; kind = 1;
                continue;
            }
            V tdir = vnorm(vsub(vscale(d, nnt), vscale(ng, (into ? 1 : -1) * (ddn * nnt + sqrt(cos2t)))));
            double a = nt - nc, b = nt + nc, R0 = a * a / (b * b), c = 1 - (into ? -ddn : vdot(tdir, ng));
            double Re = R0 + (1 - R0) * c * c * c * c * c, Tr = 1 - Re, P = .25 + .5 * Re, RP = Re / P, TP = Tr / (1 - P);
            E = 1; kind = 1; o = x;
            if (depth > 2) {
                if (rng.next() < P) { d = refl_d; T = vscale(vmult(T, f), RP); }
                else { d = tdir; T = vscale(vmult(T, f), TP); }
                continue;
            }
            if (sp < 4) {   // transmission later, reflection now (the oracle's fixed order)
                stack[sp].o = x; stack[sp].d = tdir; stack[sp].T = vscale(vmult(T, f), Tr);
                stack[sp].depth = depth; stack[sp].E = 1; stack[sp].kind = 1;
                sp++;
            }
            d = refl_d; T = vscale(vmult(T, f), Re);
        }
    }
    return L;
}

// The render loops, :528-541.  blockIdx.x enumerates the rows owned by this rank.
template <bool DET>
__global__ void __launch_bounds__(32) k_validate_fp64(const DevObj64 *__restrict__ objs, int n, pt_camera cam, pt_light light,
                                                      int w, int h, int samps, int mode, int max_depth,
                                                      int tile_rows, int rank, int world,
                                                      double *__restrict__ sum, double *__restrict__ sumsq, DevStats *stats)
{
    // row owned by this block: k-th owned row
    int k = blockIdx.x;
    int tiles_before = k / tile_rows;                    // owned tiles are rank, rank+world, ...
    int y = (tiles_before * world + rank) * tile_rows + (k % tile_rows);
    if (y >= h) return;
    int lane = threadIdx.x;
    LaneObj mine;
    mine.type = -1; mine.g0 = mine.g1 = mine.g2 = mine.g3 = mine.g4 = 0;
    if (lane < n) {
        const DevObj64 *ob = objs + lane;
        mine.type = ob->type; mine.g0 = ob->g[0]; mine.g1 = ob->g[1]; mine.g2 = ob->g[2]; mine.g3 = ob->g[3]; mine.g4 = ob->g[4];
    }
    Rng48 rng;
    rng.x = (unsigned long long)(unsigned short)((unsigned)y * (unsigned)y * (unsigned)y) << 32;   // Xi = {0, 0, y^3}, :530
    V origin = v3(cam.origin.x, cam.origin.y, cam.origin.z);
    V llc = v3(cam.lower_left_corner.x, cam.lower_left_corner.y, cam.lower_left_corner.z);
    V hor = v3(cam.horizontal.x, cam.horizontal.y, cam.horizontal.z);
    V ver = v3(cam.vertical.x, cam.vertical.y, cam.vertical.z);
    RowStats st;
    st.paths = st.rays_camera = st.rays_scatter = st.rays_shadow = st.shaded = st.misses = st.truncated = 0;
    st.max_depth = 0;
    for (int x = 0; x < w; x++) {
        V m = v3(0, 0, 0), sq = v3(0, 0, 0);
        for (int s = 0; s < samps; s++) {                                                     // :531
            float u = (float)(x - 0.5 + rng.next()) / (float)w;                               // :533 (P2)
            float v = (float)((h - y - 1) - 0.5 + rng.next()) / (float)h;                     // :534 (P2)
            V d = vsub(vadd(vadd(llc, vscale(hor, (double)u)), vscale(ver, (double)v)), origin);   // :276-279
            st.paths++;
            V L = radiance<DET>(objs, n, mine, lane, mode, max_depth, light, origin, vnorm(d), rng, st);   // :536
            m = vadd(m, L);
            sq = vadd(sq, vmult(L, L));
        }
        if (lane == 0) {
            size_t i = ((size_t)y * w + x) * 3;
            sum[i] = m.x; sum[i + 1] = m.y; sum[i + 2] = m.z;
            if (sumsq) { sumsq[i] = sq.x; sumsq[i + 1] = sq.y; sumsq[i + 2] = sq.z; }
        }
    }
    if (lane == 0) {
        atomicAdd(&stats->paths, st.paths);
        atomicAdd(&stats->rays_camera, st.rays_camera);
        atomicAdd(&stats->rays_scatter, st.rays_scatter);
        atomicAdd(&stats->rays_shadow, st.rays_shadow);
        atomicAdd(&stats->shaded, st.shaded);
        atomicAdd(&stats->misses, st.misses);
        atomicAdd(&stats->truncated, st.truncated);
        atomicMax(&stats->max_depth_seen, st.max_depth);
    }
}

// pt_debug_intersect, precision 64: one thread per ray, the sequential loop of :323-335.
__global__ void k_intersect_fp64(const DevObj64 *__restrict__ objs, int n, const double *__restrict__ rays, int n_rays,
                                 double *__restrict__ t_out, int *__restrict__ id_out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const double *r = rays + (size_t)i * 6;
    V o = v3(r[0], r[1], r[2]), d = v3(r[3], r[4], r[5]);
    double t = 1e20;
    int id = -1;
    for (int k = 0; k < n; k++) {
        const DevObj64 *ob = objs + k;
        double dd = obj_intersect(ob->type, ob->g[0], ob->g[1], ob->g[2], ob->g[3], ob->g[4], ob, o, d);
        if (dd != 0 && dd < t) { t = dd; id = k; }
    }
    t_out[i] = t;
    id_out[i] = id;
}

__global__ void k_erand48(const uint16_t *__restrict__ seeds, int n_threads, int draws, double *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_threads) return;
    Rng48 rng;
    rng.x = (unsigned long long)seeds[3 * i] | ((unsigned long long)seeds[3 * i + 1] << 16) | ((unsigned long long)seeds[3 * i + 2] << 32);
    for (int k = 0; k < draws; k++) out[(size_t)i * draws + k] = rng.next();
}

}  // namespace

int pt_fp64_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s)
{
    int tile = p->tile_rows > 0 ? p->tile_rows : 8;
    int world = p->world > 0 ? p->world : 1;
    int max_depth = p->max_depth > 0 ? p->max_depth : 4096;
    int n_tiles = (p->height + tile - 1) / tile;
    int owned_tiles = (n_tiles - p->rank + world - 1) / world;
    if (owned_tiles <= 0) return PT_OK;
    int rows = owned_tiles * tile;   // rows past the image end return immediately
    int n = (int)ctx->objs.size();
    if (p->sincos == PT_SINCOS_DET)
        k_validate_fp64<true><<<rows, 32, 0, s>>>(ctx->d_objs, n, ctx->cam, ctx->light, p->width, p->height, p->spp, p->mode,
                                                  max_depth, tile, p->rank, world, d_sum, d_sumsq, ctx->d_stats);
    else
        k_validate_fp64<false><<<rows, 32, 0, s>>>(ctx->d_objs, n, ctx->cam, ctx->light, p->width, p->height, p->spp, p->mode,
                                                   max_depth, tile, p->rank, world, d_sum, d_sumsq, ctx->d_stats);
    PT_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    return PT_OK;
}

int pt_fp64_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s)
{
    k_intersect_fp64<<<(n + 127) / 128, 128, 0, s>>>(ctx->d_objs, (int)ctx->objs.size(), d_rays, n, d_t, d_id);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp64_erand48(pt_ctx *ctx, const uint16_t *d_seeds, int n_threads, int draws, double *d_out, cudaStream_t s)
{
    k_erand48<<<(n_threads + 127) / 128, 128, 0, s>>>(d_seeds, n_threads, draws, d_out);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}
