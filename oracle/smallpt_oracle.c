/* smallpt_oracle.c — CPU restatement of maurock/small-pathtracer's per-pixel Monte Carlo
 * radiance loop.  TEST INFRASTRUCTURE ONLY: it is the checker for the CUDA path, never
 * the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this file bit-for-bit
 * against (a) oracle/_ref/smallpt_ref — the reference's own src/smallpt.cpp compiled here
 * with the textual patches P0-P6 of oracle/make_ref.py — and (b) oracle/_ref/librefharness.so,
 * which #includes the UNMODIFIED reference translation unit and exports its functions;
 * the committed fixtures under tests/golden/ were produced by those two.
 * Parts the reference does not contain (tilted planes, cone light sampling toward sphere
 * lights, SPEC/REFR at HEAD) are restated from SURVEY.md 8(a5b,a13) and are "parity
 * unpinned" — they are gated indirectly (tests/test_unpinned_*.py).
 *
 * All citations are file:line of the reference repo (src/smallpt.cpp unless stated).
 * Build: gcc -O2 -fopenmp -ffp-contract=off  (no -ffast-math, no -march=native).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/ptb200.h"
#include "../include/ptb200_detmath.h"

/* `real` is double in the oracle proper.  -DORACLE_FP32 builds an EXPERIMENTAL float variant
 * (liboracle_fp32.so) used only to study how FP32 rounding changes the reference's statistics. */
#ifdef ORACLE_FP32
typedef float real;
#define sqrt(x) sqrtf(x)
#define fabs(x) fabsf(x)
#else
typedef double real;
#endif

/* ------------------------------------------------------------------ Vec, :24-62 */
typedef struct { real x, y, z; } V;
static inline V v3(real x, real y, real z) { V r = { x, y, z }; return r; }
static inline V vadd(V a, V b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }          /* :31 */
static inline V vsub(V a, V b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }          /* :34 */
static inline V vscale(V a, real b) { return v3(a.x * b, a.y * b, a.z * b); }         /* :37-45 */
static inline V vmult(V a, V b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }         /* :47 */
static inline V vnorm(V a) { return vscale(a, 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); } /* :50 */
static inline real vdot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       /* :53 */
static inline V vcross(V a, V b) {                                                       /* :56 */
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V from_pt(pt_vec3 a) { return v3(a.x, a.y, a.z); }

/* ------------------------------------------------------------------ erand48, src/utilities.h:26-51 */
static inline void dorand48(unsigned short xs[3])
{
    unsigned long accu;
    unsigned short t0, t1;
    accu = 0xe66dUL * (unsigned long)xs[0] + 0x000bUL;
    t0 = (unsigned short)accu;
    accu >>= 16;
    accu += 0xe66dUL * (unsigned long)xs[1] + 0xdeecUL * (unsigned long)xs[0];
    t1 = (unsigned short)accu;
    accu >>= 16;
    accu += 0xe66dUL * xs[2] + 0xdeecUL * xs[1] + 0x0005UL * xs[0];
    xs[0] = t0; xs[1] = t1; xs[2] = (unsigned short)accu;
}
double oracle_erand48(unsigned short xs[3])
{
    dorand48(xs);
    return ldexp((double)xs[0], -48) + ldexp((double)xs[1], -32) + ldexp((double)xs[2], -16);
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------ scene table, :82-254, :287-311 */
enum { O_SPHERE = 0, O_XZ = 1, O_XY = 2, O_YZ = 3, O_TILT = 4 };
typedef struct {
    int type, refl;
    real rad; V p;                 /* sphere :225-226 */
    real a1, a2, b1, b2, k;        /* rectangles :94,139,182 */
    V p0, n, s, t; real hs, ht;    /* tilted plane */
    V e, c;
} Obj;

typedef struct { int n; Obj *o; pt_light light; pt_camera cam; } Scene;

static int build_scene(const pt_scene *in, Scene *sc)
{
    int n = in->n_spheres + in->n_planes;
    if (n <= 0 || n > PT_MAX_OBJECTS) return -1;
    sc->n = n;
    sc->o = (Obj *)calloc((size_t)n, sizeof(Obj));
    for (int i = 0; i < n; i++) {
        int ref = in->order ? in->order[i] : (i < in->n_planes ? i : ~(i - in->n_planes));
        Obj *o = &sc->o[i];
        if (ref < 0) {
            int j = ~ref;
            if (j >= in->n_spheres) return -1;
            const pt_sphere *s = &in->spheres[j];
            o->type = O_SPHERE; o->refl = s->refl; o->rad = s->rad;
            o->p = from_pt(s->p); o->e = from_pt(s->e); o->c = from_pt(s->c);
        } else {
            if (ref >= in->n_planes) return -1;
            const pt_plane *p = &in->planes[ref];
            o->type = p->kind == PT_PLANE_XZ ? O_XZ : p->kind == PT_PLANE_XY ? O_XY
                    : p->kind == PT_PLANE_YZ ? O_YZ : O_TILT;
            o->refl = p->refl;
            o->a1 = p->a1; o->a2 = p->a2; o->b1 = p->b1; o->b2 = p->b2; o->k = p->k;
            o->p0 = from_pt(p->p0); o->n = from_pt(p->n); o->s = from_pt(p->s); o->t = from_pt(p->t);
            o->hs = p->hs; o->ht = p->ht;
            o->e = from_pt(p->e); o->c = from_pt(p->c);
        }
    }
    sc->light = in->light;
    sc->cam = in->camera;
    return 0;
}

/* per-primitive intersect: returns distance, 0 if no hit */
static inline real obj_intersect(const Obj *ob, V o, V d)
{
    switch (ob->type) {
    case O_SPHERE: {                                            /* :229-239 */
        V op = vsub(ob->p, o);
        real t, eps = 1e-4;
        real b = vdot(op, d);
        real det = b * b - vdot(op, op) + ob->rad * ob->rad;
        if (det < 0) return 0; else det = sqrt(det);
        return (t = b - det) > eps ? t : ((t = b + det) > eps ? t : 0);
    }
    case O_XZ: {                                                /* :102-112 */
        real t = (ob->k - o.y) / d.y;
        float x = (float)(o.x + d.x * t);
        float z = (float)(o.z + d.z * t);
        if (x < ob->a1 || x > ob->a2 || z < ob->b1 || z > ob->b2 || t < 0) return 0;
        return t;
    }
    case O_XY: {                                                /* :145-155 */
        real t = (ob->k - o.z) / d.z;
        float x = (float)(o.x + d.x * t);
        float y = (float)(o.y + d.y * t);
        if (x < ob->a1 || x > ob->a2 || y < ob->b1 || y > ob->b2 || t < 0) return 0;
        return t;
    }
    case O_YZ: {                                                /* :188-198 */
        real t = (ob->k - o.x) / d.x;
        float y = (float)(o.y + d.y * t);
        float z = (float)(o.z + d.z * t);
        if (y < ob->a1 || y > ob->a2 || z < ob->b1 || z > ob->b2 || t < 0) return 0;
        return t;
    }
    default: {                                                  /* tilted, SURVEY 8 a5b */
        real denom = vdot(ob->n, d);
        real tau = vdot(ob->n, vsub(ob->p0, o)) / denom;
        V rel = vsub(vadd(o, vscale(d, tau)), ob->p0);
        real a = vdot(rel, ob->s), b = vdot(rel, ob->t);
        if (!(fabs(a) <= ob->hs) || !(fabs(b) <= ob->ht) || !(tau > 1e-4)) return 0;
        return tau;
    }
    }
}

/* intersect(Ray,t,id), :323-335.  id untouched on a miss. */
static inline int scene_intersect(const Scene *sc, V o, V d, real *t, int *id)
{
    real dd, inf = *t = 1e20;
    for (int i = 0; i < sc->n; i++) {
        if ((dd = obj_intersect(&sc->o[i], o, d)) && dd < *t) { *t = dd; *id = i; }
    }
    return *t < inf;
}

/* normal(), rect :118-124,161-167,204-210; sphere :246-253.  Returns nl (oriented against r.d);
 * *n_geo receives the un-oriented normal (needed by REFR). */
static inline V obj_normal(const Obj *ob, V d, V x, V *n_geo)
{
    V n;
    switch (ob->type) {
    case O_SPHERE: n = vnorm(vsub(x, ob->p)); break;
    case O_XZ: n = v3(0, 1, 0); break;
    case O_XY: n = v3(0, 0, 1); break;
    case O_YZ: n = v3(1, 0, 0); break;
    default:   n = ob->n; break;
    }
    *n_geo = n;
    return vdot(n, d) < 0 ? n : v3(n.x * -1, n.y * -1, n.z * -1);
}

/* ------------------------------------------------------------------ integrator */
typedef struct {
    uint64_t paths, rays_camera, rays_scatter, rays_shadow, shaded, misses, truncated;
    uint32_t max_depth_seen;
} Stats;

typedef struct {
    const Scene *sc;
    int mode, sincos, max_depth;
    unsigned short *Xi;
    Stats *st;
} Ctx;

static inline void ctx_sincos(const Ctx *cx, real a, real *s, real *c)
{
    if (cx->sincos == PT_SINCOS_DET) { double sd, cd; pt_det_sincos(a, &sd, &cd); *s = sd; *c = cd; }
    else { *c = cos(a); *s = sin(a); }
}

/* random_scattering, cosine :337-348 / uniform :351-360 (returns a normalised vector) */
static inline V random_scattering(const Ctx *cx, V nl)
{
    real r1 = 2 * M_PI * oracle_erand48(cx->Xi);
    real r2 = oracle_erand48(cx->Xi);
    real sn, cs;
    ctx_sincos(cx, r1, &sn, &cs);
    V w = nl;
    V u = vnorm(vcross(fabs(w.x) > .1 ? v3(0, 1, 0) : v3(1, 0, 0), w));
    V v = vcross(w, u);
    if (cx->mode == PT_MODE_UNI) {
        real q = sqrt(r2 * (2 - r2));
        return vnorm(vadd(vadd(vscale(vscale(u, cs), q), vscale(vscale(v, sn), sqrt(r2 * (2 - r2)))), vscale(w, 1 - r2)));
    } else {
        real r2s = sqrt(r2);
        return vnorm(vadd(vadd(vscale(vscale(u, cs), r2s), vscale(vscale(v, sn), r2s)), vscale(w, sqrt(1 - r2))));
    }
}

/* radiance(), :419-496 with the dead RL block :424-442 removed.
 * E is only used by PT_MODE_NEE_CONE_SPHERE (emission of sampled sphere lights is not
 * counted again on the next diffuse hit); it is 1 everywhere in the reference modes. */
static V radiance(const Ctx *cx, V ro, V rd, int depth, int E, int ray_kind)
{
    const Scene *sc = cx->sc;
    Stats *st = cx->st;
    int id = 0;
    real t;
    V x;
    if (ray_kind == 0) st->rays_camera++; else if (ray_kind == 1) st->rays_scatter++;
    if (!scene_intersect(sc, ro, rd, &t, &id)) { x = v3(0, 0, 0); st->misses++; }   /* :371-374 */
    else x = vadd(ro, vscale(rd, t));                                               /* :375 */
    const Obj *ob = &sc->o[id];
    V n;
    V nl = obj_normal(ob, rd, x, &n);                                               /* :445 */
    V f = ob->c;                                                                    /* :446 */
    V e = ob->e;
    st->shaded++;
    if (cx->mode == PT_MODE_NEE_CONE_SPHERE && !E && ob->type == O_SPHERE) e = v3(0, 0, 0);
    real p = f.x > f.y && f.x > f.z ? f.x : f.y > f.z ? f.y : f.z;                /* :447 */
    if (++depth > 5 || !p) {                                                        /* :448 */
        if (oracle_erand48(cx->Xi) < p) f = vscale(f, 1 / p);
        else { if ((uint32_t)depth > st->max_depth_seen) st->max_depth_seen = depth; return e; }
    }
    if ((uint32_t)depth > st->max_depth_seen) st->max_depth_seen = depth;
    if (depth >= cx->max_depth) { st->truncated++; return e; }
    if (ob->refl == PT_DIFF) {                                                      /* :457 */
        V d;
        real PDF_inverse = 1, BRDF = 1;
        if (cx->mode == PT_MODE_NEE_REF_RECT) {                                     /* :464 q < 1 */
            const pt_light *L = &sc->light;
            real x_light = L->x0 + L->xw * oracle_erand48(cx->Xi);                /* :365 (P2) */
            real z_light = L->z0 + L->zw * oracle_erand48(cx->Xi);                /* :366 (P2) */
            d = vsub(v3(x_light, L->y, z_light), x);                                /* :367 */
            d = vnorm(d);                                                           /* :466 */
            st->rays_shadow++;
            scene_intersect(sc, x, d, &t, &id);
            if (id != L->id) {                                                      /* :467 */
                d = random_scattering(cx, nl);                                      /* :468 */
                d = vnorm(d);                                                       /* :469 (re-trace skipped: no side effect) */
            } else {
                d = vnorm(d);
                PDF_inverse = fabs((L->area * vdot(d, v3(0, 1, 0))) / (t * t));     /* :471 */
                d = vnorm(d);
                BRDF = fabs(vdot(d, nl) / M_PI);                                    /* :472 */
            }
            d = vnorm(d);                                                           /* :479 */
            /* a visible light sample continues along the shadow ray: not a new unique ray */
            V r = radiance(cx, x, d, depth, 1, id != L->id ? 1 : 2);
            return vadd(e, vscale(vscale(vmult(f, r), PDF_inverse), BRDF));         /* :479 */
        } else if (cx->mode == PT_MODE_NEE_CONE_SPHERE) {
            /* Not in the reference source (parity unpinned): next-event estimation toward every
             * emissive sphere by sampling the cone it subtends, shadow ray, no double counting. */
            d = random_scattering(cx, nl);
            V esum = v3(0, 0, 0);
            for (int i = 0; i < sc->n; i++) {
                const Obj *s = &sc->o[i];
                if (s->type != O_SPHERE) continue;
                if (s->e.x <= 0 && s->e.y <= 0 && s->e.z <= 0) continue;
                V sw = vsub(s->p, x);
                real dist2 = vdot(sw, sw);
                real eps1 = oracle_erand48(cx->Xi), eps2 = oracle_erand48(cx->Xi);
                if (!(dist2 > s->rad * s->rad)) continue;
                sw = vscale(sw, 1 / sqrt(dist2));
                V su = vnorm(vcross(fabs(sw.x) > .1 ? v3(0, 1, 0) : v3(1, 0, 0), sw));
                V sv = vcross(sw, su);
                real cos_a_max = sqrt(1 - s->rad * s->rad / dist2);
                real cos_a = 1 - eps1 + eps1 * cos_a_max;
                real sin_a = sqrt(1 - cos_a * cos_a);
                real phi = 2 * M_PI * eps2, sp, cp;
                ctx_sincos(cx, phi, &sp, &cp);
                V l = vnorm(vadd(vadd(vscale(su, cp * sin_a), vscale(sv, sp * sin_a)), vscale(sw, cos_a)));
                real ts; int ids = -1;
                st->rays_shadow++;
                if (scene_intersect(sc, x, l, &ts, &ids) && ids == i) {
                    real omega = 2 * M_PI * (1 - cos_a_max);
                    real ldn = vdot(l, nl);
                    if (ldn > 0) esum = vadd(esum, vscale(vmult(f, vscale(s->e, ldn * omega)), 1 / M_PI));
                }
            }
            V r = radiance(cx, x, d, depth, 0, 1);
            return vadd(vadd(e, esum), vmult(f, r));
        } else {                                                                    /* :474-477 */
            d = random_scattering(cx, nl);
            d = vnorm(d);                                                           /* :476 */
            d = vnorm(d);                                                           /* :479 */
            V r = radiance(cx, x, d, depth, 1, 1);
            return vadd(e, vscale(vscale(vmult(f, r), PDF_inverse), BRDF));
        }
    } else if (ob->refl == PT_SPEC) {                                               /* :482-483 */
        V rr = vsub(rd, vscale(n, 2 * vdot(n, rd)));
        return vadd(e, vmult(f, radiance(cx, x, rr, depth, 1, 1)));
    }
    /* REFR, :485-495 (commented at HEAD; live in src/a.exe) */
    V refl_d = vsub(rd, vscale(n, 2 * vdot(n, rd)));
    int into = vdot(n, nl) > 0;
    real nc = 1, nt = 1.5, nnt = into ? nc / nt : nt / nc, ddn = vdot(rd, nl), cos2t;
    if ((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0)
        return vadd(e, vmult(f, radiance(cx, x, refl_d, depth, 1, 1)));
    V tdir = vnorm(vsub(vscale(rd, nnt), vscale(n, (into ? 1 : -1) * (ddn * nnt + sqrt(cos2t)))));
    real a = nt - nc, b = nt + nc, R0 = a * a / (b * b), c = 1 - (into ? -ddn : vdot(tdir, n));
    real Re = R0 + (1 - R0) * c * c * c * c * c, Tr = 1 - Re, P = .25 + .5 * Re, RP = Re / P, TP = Tr / (1 - P);
    if (depth > 2) {
        if (oracle_erand48(cx->Xi) < P) return vadd(e, vmult(f, vscale(radiance(cx, x, refl_d, depth, 1, 1), RP)));
        return vadd(e, vmult(f, vscale(radiance(cx, x, tdir, depth, 1, 1), TP)));
    }
    /* both branches; evaluation order fixed here: reflection first, then transmission */
    V a1 = vscale(radiance(cx, x, refl_d, depth, 1, 1), Re);
    V a2 = vscale(radiance(cx, x, tdir, depth, 1, 1), Tr);
    return vadd(e, vmult(f, vadd(a1, a2)));
}

static inline real clamp01(real x) { return x < 0 ? 0 : x > 1 ? 1 : x; }           /* :314-316 */
static inline double clamp01d(double x) { return x < 0 ? 0 : x > 1 ? 1 : x; }
int oracle_toInt(double x) { return (int)(pow(clamp01d(x), 1 / 2.2) * 255 + .5); }      /* :319-321 */

typedef struct oracle_stats {
    uint64_t paths, rays_camera, rays_scatter, rays_shadow, shaded_vertices, miss_events, truncated;
    uint32_t max_depth_seen, threads;
    double render_ms;
} oracle_stats;

/* The render loops, :528-541, with the per-row Xi seed of :530 (P4: explicit u16 cast of the
 * wrapped 32-bit cube) and every draw from Xi (P2).  rgb_clamped = c[] (:538); rgb_mean =
 * un-clamped mean; rgb_sumsq = per-channel sum over samples of radiance^2 (for the 3-sigma gate).
 * Any output pointer may be NULL.  Rows of foreign tiles (rank/world/tile_rows) are skipped. */
int oracle_render(const pt_scene *scene, const pt_render_params *prm,
                  double *rgb_clamped, double *rgb_mean, double *rgb_sumsq, oracle_stats *stats_out)
{
    Scene sc;
    if (build_scene(scene, &sc)) return -1;
    int w = prm->width, h = prm->height, samps = prm->spp;
    int max_depth = prm->max_depth > 0 ? prm->max_depth : 4096;
    int tile = prm->tile_rows > 0 ? prm->tile_rows : 8;
    int world = prm->world > 0 ? prm->world : 1, rank = prm->rank;
    V origin = from_pt(sc.cam.origin), llc = from_pt(sc.cam.lower_left_corner);
    V hor = from_pt(sc.cam.horizontal), ver = from_pt(sc.cam.vertical);
    Stats tot; memset(&tot, 0, sizeof tot);
    int nthreads = 1;
    double t0 = 0, t1 = 0;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
    t0 = omp_get_wtime();
#endif
#pragma omp parallel
    {
        Stats st; memset(&st, 0, sizeof st);
#pragma omp for schedule(dynamic, 1)                                                   /* :526 */
        for (int y = 0; y < h; y++) {                                                  /* :528 */
            if (((y / tile) % world) != rank) continue;
            unsigned short Xi[3] = { 0, 0, (unsigned short)((uint32_t)y * (uint32_t)y * (uint32_t)y) }; /* :530 */
            Ctx cx = { &sc, prm->mode, prm->sincos, max_depth, Xi, &st };
            for (int x = 0; x < w; x++) {
                V r = v3(0, 0, 0), m = v3(0, 0, 0), sq = v3(0, 0, 0);
                for (int s = 0; s < samps; s++) {                                      /* :531 */
                    float u = (float)(x - 0.5 + oracle_erand48(Xi)) / (float)w;            /* :533 */
                    float v = (float)((h - y - 1) - 0.5 + oracle_erand48(Xi)) / (float)h;  /* :534 */
                    V d = vsub(vadd(vadd(llc, vscale(hor, u)), vscale(ver, v)), origin);  /* :276-279 */
                    st.paths++;
                    V L = radiance(&cx, origin, vnorm(d), 0, 1, 0);                    /* :536 */
                    r = vadd(r, vscale(L, 1. / samps));
                    m = vadd(m, L);
                    sq = vadd(sq, vmult(L, L));
                }
                size_t i = ((size_t)y * w + x) * 3;
                if (rgb_clamped) { rgb_clamped[i] = clamp01(r.x); rgb_clamped[i + 1] = clamp01(r.y); rgb_clamped[i + 2] = clamp01(r.z); } /* :538 */
                if (rgb_mean) { rgb_mean[i] = m.x / samps; rgb_mean[i + 1] = m.y / samps; rgb_mean[i + 2] = m.z / samps; }
                if (rgb_sumsq) { rgb_sumsq[i] = sq.x; rgb_sumsq[i + 1] = sq.y; rgb_sumsq[i + 2] = sq.z; }
            }
        }
#pragma omp critical
        {
            tot.paths += st.paths; tot.rays_camera += st.rays_camera; tot.rays_scatter += st.rays_scatter;
            tot.rays_shadow += st.rays_shadow; tot.shaded += st.shaded; tot.misses += st.misses;
            tot.truncated += st.truncated;
            if (st.max_depth_seen > tot.max_depth_seen) tot.max_depth_seen = st.max_depth_seen;
        }
    }
#ifdef _OPENMP
    t1 = omp_get_wtime();
#endif
    if (stats_out) {
        stats_out->paths = tot.paths; stats_out->rays_camera = tot.rays_camera;
        stats_out->rays_scatter = tot.rays_scatter; stats_out->rays_shadow = tot.rays_shadow;
        stats_out->shaded_vertices = tot.shaded; stats_out->miss_events = tot.misses;
        stats_out->truncated = tot.truncated; stats_out->max_depth_seen = tot.max_depth_seen;
        stats_out->threads = (uint32_t)nthreads; stats_out->render_ms = (t1 - t0) * 1e3;
    }
    free(sc.o);
    return 0;
}

/* Closest-hit queries: intersect(), :323-335 + miss rule of hittingPoint, :371-374
 * (t = 1e20, id = -1 on a miss). */
int oracle_intersect(const pt_scene *scene, const double *rays_od, int n, double *t_out, int *id_out)
{
    Scene sc;
    if (build_scene(scene, &sc)) return -1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const double *r = rays_od + (size_t)i * 6;
        real t; int id = -1;
        scene_intersect(&sc, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), &t, &id);
        t_out[i] = t; id_out[i] = id;
    }
    free(sc.o);
    return 0;
}

/* Camera::Camera, :262-275 */
void oracle_camera(const pt_vec3 *lookfrom, const pt_vec3 *lookat, const pt_vec3 *vup,
                   float vfov, float aspect, pt_camera *out)
{
    float theta = vfov * M_PI / 180;
    float half_height = tanf(theta / 2);
    float half_width = aspect * half_height;
    V origin = from_pt(*lookfrom);
    V w = vnorm(vsub(from_pt(*lookat), from_pt(*lookfrom)));
    V u = vnorm(vcross(w, from_pt(*vup)));
    V v = vcross(u, w);
    V llc = vadd(vsub(vsub(origin, vscale(u, half_width)), vscale(v, half_height)), w);
    V hor = vscale(u, half_width * 2), ver = vscale(v, half_height * 2);
    out->origin = (pt_vec3){ origin.x, origin.y, origin.z };
    out->lower_left_corner = (pt_vec3){ llc.x, llc.y, llc.z };
    out->horizontal = (pt_vec3){ hor.x, hor.y, hor.z };
    out->vertical = (pt_vec3){ ver.x, ver.y, ver.z };
}

/* PPM writer, :548-551: "P3\n%d %d\n%d\n" then "%d %d %d " per pixel. rgb = c[] (w*h*3). */
int oracle_write_ppm(const char *path, const double *rgb, int w, int h)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "P3\n%d %d\n%d\n", w, h, 255);
    for (int i = 0; i < w * h; i++)
        fprintf(f, "%d %d %d ", oracle_toInt(rgb[3 * i]), oracle_toInt(rgb[3 * i + 1]), oracle_toInt(rgb[3 * i + 2]));
    fclose(f);
    return 0;
}

void oracle_det_sincos(double a, double *s, double *c) { pt_det_sincos(a, s, c); }

/* random_scattering(nl, Xi), :337-360, exported for the golden-vector test. */
void oracle_random_scattering(const double *nl, unsigned short *xi, int mode, int sincos, double *d_out)
{
    Ctx cx = { 0, mode, sincos, 0, xi, 0 };
    V d = random_scattering(&cx, v3(nl[0], nl[1], nl[2]));
    d_out[0] = d.x; d_out[1] = d.y; d_out[2] = d.z;
}

/* number of OpenMP threads used by oracle_render / oracle_intersect (0 = leave as is); returns the maximum */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}
