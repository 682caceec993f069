// pt_api.cu — the extern "C" boundary of libptb200.so (include/ptb200.h): context, scene upload,
// render dispatch, readback, debug entries.  No torch types, plain pointers and sizes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>

#include "pt_internal.h"

static thread_local std::string g_last_error;   // for failures before a context exists

// The FP32 engine's scene image `c_scene` is ONE __constant__ symbol per device (and one per cached specialised module):
// contexts that share a device take turns from the upload of the symbol to the end of the render that reads it.
static std::mutex g_dev_mutex[64];
static std::mutex &dev_mutex(int device) { return g_dev_mutex[(unsigned)device % 64u]; }

int pt_fail(pt_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

static inline bool finite3(const pt_vec3 &v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

// Build the FP32 class-sorted constant-memory image of the scene and the code-indexed material table.
static void build_scene_f32(pt_ctx *ctx, std::vector<MatF32> &mats)
{
    SceneF32 &S = *ctx->h_scene32;
    std::memset(&S, 0, sizeof S);
    const int n = (int)ctx->objs.size();
    ctx->fp32_ok = true;
    ctx->fp32_why.clear();
    mats.clear();
    ctx->grid.n = 0;
    ctx->h_grid_start.clear(); ctx->h_grid_items.clear(); ctx->h_grid_sph.clear();
    std::vector<int> code_of(n, -1);
    // rectangles: the first PT_RECT_SLOTS of each axis class go to the unrolled slots, the rest to the overflow loop
    int n_ovf = 0;
    for (int axis = 0; axis < 3; axis++) {       // XZ, XY, YZ
        S.ovf_begin[axis] = n_ovf;
        const int want = axis == 0 ? OT_XZ : axis == 1 ? OT_XY : OT_YZ;
        int k = 0;
        for (int i = 0; i < n; i++) {
            const DevObj64 &o = ctx->objs[i];
            if (o.type != want) continue;
            if (k < PT_RECT_SLOTS) {
                // {k, a1, a2 - a1, b1}, b2 - b1: widths as float differences of the float bounds, so that u == a2 (as
                // floats) still passes:  fl(a2) - fl(a1) rounded up to the next float if needed
                const float a1 = (float)o.g[0], a2 = (float)o.g[1], b1 = (float)o.g[2], b2 = (float)o.g[3];
                float wa = a2 - a1, wb = b2 - b1;
                if (a1 + wa < a2) wa = std::nextafter(wa, INFINITY);
                if (b1 + wb < b2) wb = std::nextafter(wb, INFINITY);
                S.slot_a[axis][k] = make_float4((float)o.g[4], a1, wa, b1);
                S.slot_b2[axis][k] = wb;
                code_of[i] = axis * PT_RECT_SLOTS + k;
                k++;
            } else {
                if (n_ovf < PT_MAX_OBJ) {
                    S.rect_a[n_ovf] = make_float4((float)o.g[4], (float)o.g[0], (float)o.g[1], (float)o.g[2]);
                    S.rect_b2[n_ovf] = (float)o.g[3];
                }
                code_of[i] = 3 * PT_RECT_SLOTS + n_ovf;
                n_ovf++;
            }
        }
        S.n_slot[axis] = k;
    }
    S.ovf_begin[3] = n_ovf;
    int n_small = 0, n_huge = 0, n_tilt = 0;
    for (int i = 0; i < n; i++) {
        const DevObj64 &o = ctx->objs[i];
        if (o.type == OT_SPHERE) { if (o.g[0] >= PT_HUGE_RADIUS) n_huge++; else n_small++; }
        else if (o.type == OT_TILT) n_tilt++;
    }
    // small spheres: brute force (the measured contract) up to PT_MAX_OBJ, beyond that - or on request - a uniform grid
    const bool use_grid = n_small > 0 && (ctx->accel_mode == 2 || (ctx->accel_mode == 1 && n_small > PT_MAX_OBJ));
    if (n_small > PT_MAX_OBJ && !use_grid) { ctx->fp32_ok = false; ctx->fp32_why = "more than 512 small spheres (brute force); pt_set_acceleration(ctx, 1) enables the uniform grid"; return; }
    if (n_ovf > PT_MAX_OBJ) { ctx->fp32_ok = false; ctx->fp32_why = "more than 512 rectangles beyond the 48 unrolled slots"; return; }
    if (n_huge > PT_MAX_HUGE) { ctx->fp32_ok = false; ctx->fp32_why = "more than 64 huge spheres"; return; }
    if (n_tilt > PT_MAX_TILT) { ctx->fp32_ok = false; ctx->fp32_why = "more than 64 tilted planes"; return; }
    S.code_sph0 = 3 * PT_RECT_SLOTS + n_ovf;
    S.code_huge0 = S.code_sph0 + n_small;      // (grid or scan: sphere k of the class has code code_sph0 + k either way)
    S.code_tilt0 = S.code_huge0 + n_huge;
    S.n_codes = S.code_tilt0 + n_tilt;
    for (int i = 0; i < n; i++) {
        const DevObj64 &o = ctx->objs[i];
        if (o.type == OT_SPHERE) {
            if (o.g[0] >= PT_HUGE_RADIUS) {
                double *h = S.huge[S.n_huge];
                h[0] = o.g[1]; h[1] = o.g[2]; h[2] = o.g[3]; h[3] = o.g[0] * o.g[0];
                for (int a = 0; a < 3; a++) S.hugef[S.n_huge][a] = (float)h[a];
                code_of[i] = S.code_huge0 + S.n_huge++;
            } else if (use_grid) {      // grid: the sphere table lives in global memory, the constant-memory scan tables stay empty
                code_of[i] = S.code_sph0 + (int)ctx->h_grid_sph.size();
                ctx->h_grid_sph.push_back(make_float4((float)o.g[1], (float)o.g[2], (float)o.g[3], (float)(o.g[0] * o.g[0])));
            } else {
                S.sph[S.n_sph] = make_float4((float)o.g[1], (float)o.g[2], (float)o.g[3], (float)(o.g[0] * o.g[0]));
                code_of[i] = S.code_sph0 + S.n_sph++;
            }
            if ((o.e[0] > 0 || o.e[1] > 0 || o.e[2] > 0) && S.n_lights < 32) S.light_sph_code[S.n_lights++] = code_of[i];
        } else if (o.type == OT_TILT) {
            float4 *t = S.tilt[S.n_tilt];
            auto dotp = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
            t[0] = make_float4((float)o.n[0], (float)o.n[1], (float)o.n[2], (float)dotp(o.n, o.p0));
            t[1] = make_float4((float)o.s[0], (float)o.s[1], (float)o.s[2], (float)dotp(o.s, o.p0));
            t[2] = make_float4((float)o.t[0], (float)o.t[1], (float)o.t[2], (float)dotp(o.t, o.p0));
            t[3] = make_float4((float)o.hs, (float)o.ht, 0.f, 0.f);
            code_of[i] = S.code_tilt0 + S.n_tilt++;
        }
    }
    if (use_grid) {
        // Uniform grid over the bounding box of the small spheres, about 4 cells per sphere; a sphere is listed in every cell its
        // bounding box touches (slightly padded: the traversal works in FP32).  CSR layout, indices ascending within a cell.
        const std::vector<float4> &sp = ctx->h_grid_sph;
        const int ns = (int)sp.size();
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (const float4 &q : sp) {
            const double r = std::sqrt((double)q.w), c[3] = {q.x, q.y, q.z};
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], c[a] - r); hi[a] = std::fmax(hi[a], c[a] + r); }
        }
        double ext[3], vol = 1;
        for (int a = 0; a < 3; a++) { const double pad = 1e-3 * (hi[a] - lo[a]) + 1e-3; lo[a] -= pad; hi[a] += pad; ext[a] = hi[a] - lo[a]; vol *= ext[a]; }
        const double per = std::cbrt(4.0 * ns / vol);
        GridDev &G = ctx->grid;
        size_t ncell = 1;
        for (int a = 0; a < 3; a++) {
            G.res[a] = (int)std::fmin(128.0, std::fmax(1.0, std::floor(ext[a] * per + 0.5)));
            G.lo[a] = (float)lo[a]; G.cell[a] = (float)(ext[a] / G.res[a]); G.inv_cell[a] = (float)(G.res[a] / ext[a]);
            ncell *= (size_t)G.res[a];
        }
        auto cell_range = [&](const float4 &q, int a, int &c0, int &c1) {
            const double r = std::sqrt((double)q.w) * (1.0 + 1e-5) + 1e-4, c = a == 0 ? q.x : a == 1 ? q.y : q.z;
            c0 = (int)std::floor((c - r - lo[a]) / ext[a] * G.res[a]); c1 = (int)std::floor((c + r - lo[a]) / ext[a] * G.res[a]);
            c0 = std::max(0, std::min(G.res[a] - 1, c0)); c1 = std::max(0, std::min(G.res[a] - 1, c1));
        };
        std::vector<unsigned int> &start = ctx->h_grid_start, &items = ctx->h_grid_items;
        start.assign(ncell + 1, 0u);
        for (int pass = 0; pass < 2; pass++) {
            for (int k = 0; k < ns; k++) {                       // ascending k: items of a cell come out ascending
                int x0, x1, y0, y1, z0, z1;
                cell_range(sp[k], 0, x0, x1); cell_range(sp[k], 1, y0, y1); cell_range(sp[k], 2, z0, z1);
                for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
                    const size_t cidx = ((size_t)z * G.res[1] + y) * G.res[0] + x;
                    if (pass == 0) start[cidx + 1]++;
                    else items[start[cidx]++] = (unsigned int)k;
                }
            }
            if (pass == 0) {
                for (size_t i = 0; i < ncell; i++) start[i + 1] += start[i];
                items.resize(start[ncell]);
            } else {                                             // the fill advanced every start to its end: shift back
                for (size_t i = ncell; i > 0; i--) start[i] = start[i - 1];
                start[0] = 0u;
            }
        }
        G.n = ns;
    }
    {   // conservative scan form of the small spheres: translate so the cloud of centres is centred on the origin
        // (smaller magnitudes => smaller FP32 cancellation error in b = c'.d - o'.d and c = |c'|^2 - r^2 - 2 c'.o' + |o'|^2)
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = 0; i < n; i++) {
            const DevObj64 &o = ctx->objs[i];
            if (o.type != OT_SPHERE || o.g[0] >= PT_HUGE_RADIUS) continue;
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], o.g[1 + a]); hi[a] = std::fmax(hi[a], o.g[1 + a]); }
        }
        for (int a = 0; a < 3; a++) S.sph_c[a] = S.n_sph ? (float)(0.5 * (lo[a] + hi[a])) : 0.f;
        double M2 = 0;
        for (int k = 0; k < S.n_sph; k++) {
            const float4 c = S.sph[k];               // the FP32 centre the exact test uses
            const float x = (float)((double)c.x - (double)S.sph_c[0]), y = (float)((double)c.y - (double)S.sph_c[1]),
                        z = (float)((double)c.z - (double)S.sph_c[2]);
            const double n2 = (double)x * x + (double)y * y + (double)z * z;
            S.sphf[k] = make_float4(x, y, z, (float)(n2 - (double)c.w));
            const double m = std::sqrt(n2) + std::sqrt((double)c.w);
            M2 = std::fmax(M2, m * m);
        }
        S.n_sph4 = (S.n_sph + 3) / 4 * 4;
        for (int k = S.n_sph; k < S.n_sph4; k++) S.sphf[k] = make_float4(0.f, 0.f, 0.f, 3.0e38f);
        S.sph_kM2 = (float)(PT_SPH_KAPPA * M2);
    }
    {   // re-centred form of the huge spheres: with w = o - huge_c (small) and G = centre - huge_c,
        //   |o - centre|^2 - rad^2 = |w|^2 - 2 w.G + (|G|^2 - rad^2): the constant K = |G|^2 - rad^2 is formed in FP64 here
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        int n_near = 0;
        for (int i = 0; i < n; i++) {
            const DevObj64 &o = ctx->objs[i];
            double q[3];
            if (o.type == OT_SPHERE) { if (o.g[0] >= PT_HUGE_RADIUS) continue; q[0] = o.g[1]; q[1] = o.g[2]; q[2] = o.g[3]; }
            else if (o.type == OT_TILT) { q[0] = o.p0[0]; q[1] = o.p0[1]; q[2] = o.p0[2]; }
            else {      // rectangle centre: (a, b, k) in the axis order of its class
                const double ca = 0.5 * (o.g[0] + o.g[1]), cb = 0.5 * (o.g[2] + o.g[3]), k = o.g[4];
                if (o.type == OT_XZ) { q[0] = ca; q[1] = k; q[2] = cb; } else if (o.type == OT_XY) { q[0] = ca; q[1] = cb; q[2] = k; } else { q[0] = k; q[1] = ca; q[2] = cb; }
            }
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], q[a]); hi[a] = std::fmax(hi[a], q[a]); }
            n_near++;
        }
        const double cam[3] = {ctx->cam.origin.x, ctx->cam.origin.y, ctx->cam.origin.z};
        for (int a = 0; a < 3; a++) S.huge_c[a] = (float)(n_near ? 0.5 * (lo[a] + hi[a]) : cam[a]);
        for (int k = 0; k < S.n_huge; k++) {
            double G[3], K = -S.huge[k][3];
            for (int a = 0; a < 3; a++) { G[a] = S.huge[k][a] - (double)S.huge_c[a]; K += G[a] * G[a]; }
            for (int a = 0; a < 3; a++) { S.hugeg[k][a] = (float)G[a]; S.hugeg[k][3 + a] = (float)(G[a] - (double)S.hugeg[k][a]); }
            S.hugeg[k][6] = (float)K; S.hugeg[k][7] = (float)(K - (double)S.hugeg[k][6]);
        }
    }
    for (int i = 0; i < n; i++) S.refl_mask |= 1 << ctx->objs[i].refl;
    S.code_obj0 = code_of[0];
    S.light_code = (ctx->light.id >= 0 && ctx->light.id < n) ? code_of[ctx->light.id] : -2;
    S.lx0 = (float)ctx->light.x0; S.lxw = (float)ctx->light.xw;
    S.lz0 = (float)ctx->light.z0; S.lzw = (float)ctx->light.zw;
    S.ly = (float)ctx->light.y; S.larea = (float)ctx->light.area;
    if (ctx->light.id >= 0 && ctx->light.id < n)
        for (int a = 0; a < 3; a++) { S.light_e[a] = (float)ctx->objs[ctx->light.id].e[a]; S.light_c[a] = (float)ctx->objs[ctx->light.id].c[a]; }
    // materials by code (unused slots stay zero)
    MatF32 zero;
    std::memset(&zero, 0, sizeof zero);
    mats.assign(S.n_codes, zero);
    for (int i = 0; i < n; i++) {
        const DevObj64 &o = ctx->objs[i];
        MatF32 m;
        m.c_refl = make_float4((float)o.c[0], (float)o.c[1], (float)o.c[2], 0.f);
        std::memcpy(&m.c_refl.w, &o.refl, sizeof(int));
        m.e_type = make_float4((float)o.e[0], (float)o.e[1], (float)o.e[2], 0.f);
        std::memcpy(&m.e_type.w, &o.type, sizeof(int));
        m.aux = make_float4((float)o.p0[0], (float)o.p0[1], (float)o.p0[2], 0.f);
        std::memcpy(&m.aux.w, &i, sizeof(int));
        if (o.type == OT_SPHERE) m.geom = make_float4((float)o.g[1], (float)o.g[2], (float)o.g[3], (float)(1.0 / o.g[0]));
        else if (o.type == OT_TILT) m.geom = make_float4((float)o.n[0], (float)o.n[1], (float)o.n[2], 0.f);
        else { const float khi = (float)o.g[4]; m.geom = make_float4(khi, (float)(o.g[4] - (double)khi), 0.f, 0.f); }
        mats[code_of[i]] = m;
    }
}

extern "C" {

const char *pt_version(void) { return "ptb200 0.1 (sm_100a)"; }

const char *pt_last_error(pt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

// validation + flattening of the table in id order (the reference's rect[] order, src/smallpt.cpp:287-311); no CUDA
static int flatten_scene(const pt_scene *scene, std::vector<DevObj64> &objs, std::string &why)
{
    const int n = scene->n_spheres + scene->n_planes;
    objs.resize(n);
    for (int i = 0; i < n; i++) {
        const int ref = scene->order ? scene->order[i] : (i < scene->n_planes ? i : ~(i - scene->n_planes));
        DevObj64 o;
        std::memset(&o, 0, sizeof o);
        if (ref < 0) {
            const int j = ~ref;
            if (j >= scene->n_spheres) { why = "order[] names a sphere that does not exist"; return PT_ERR_ARG; }
            const pt_sphere &s = scene->spheres[j];
            if (!(s.rad > 0) || !finite3(s.p)) { why = "bad sphere"; return PT_ERR_ARG; }
            o.type = OT_SPHERE; o.refl = s.refl;
            o.g[0] = s.rad; o.g[1] = s.p.x; o.g[2] = s.p.y; o.g[3] = s.p.z;
            o.e[0] = s.e.x; o.e[1] = s.e.y; o.e[2] = s.e.z; o.c[0] = s.c.x; o.c[1] = s.c.y; o.c[2] = s.c.z;
        } else {
            if (ref >= scene->n_planes) { why = "order[] names a plane that does not exist"; return PT_ERR_ARG; }
            const pt_plane &p = scene->planes[ref];
            if (p.kind < PT_PLANE_XZ || p.kind > PT_PLANE_TILTED) { why = "bad plane kind"; return PT_ERR_ARG; }
            o.type = p.kind == PT_PLANE_XZ ? OT_XZ : p.kind == PT_PLANE_XY ? OT_XY : p.kind == PT_PLANE_YZ ? OT_YZ : OT_TILT;
            o.refl = p.refl;
            o.g[0] = p.a1; o.g[1] = p.a2; o.g[2] = p.b1; o.g[3] = p.b2; o.g[4] = p.k;
            o.p0[0] = p.p0.x; o.p0[1] = p.p0.y; o.p0[2] = p.p0.z; o.n[0] = p.n.x; o.n[1] = p.n.y; o.n[2] = p.n.z;
            o.s[0] = p.s.x; o.s[1] = p.s.y; o.s[2] = p.s.z; o.t[0] = p.t.x; o.t[1] = p.t.y; o.t[2] = p.t.z;
            o.hs = p.hs; o.ht = p.ht;
            o.e[0] = p.e.x; o.e[1] = p.e.y; o.e[2] = p.e.z; o.c[0] = p.c.x; o.c[1] = p.c.y; o.c[2] = p.c.z;
        }
        if (o.refl < PT_DIFF || o.refl > PT_REFR) { why = "bad material"; return PT_ERR_ARG; }
        objs[i] = o;
    }
    return PT_OK;
}

// host -> device copy of the scene tables (FP64 object table, FP32 material table) + the FP32 constant image
static int upload_tables(pt_ctx *ctx)
{
    const int n = (int)ctx->objs.size();
    if (ctx->n_alloc < n) {
        if (ctx->d_objs) cudaFree(ctx->d_objs);
        ctx->d_objs = nullptr; ctx->n_alloc = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_objs, sizeof(DevObj64) * n));
        ctx->n_alloc = n;
    }
    std::vector<MatF32> mats;
    build_scene_f32(ctx, mats);
    PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_objs, ctx->objs.data(), sizeof(DevObj64) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->fp32_ok) {
        const int nc = (int)mats.size();
        if (ctx->n_codes_alloc < nc) {
            if (ctx->d_mats) cudaFree(ctx->d_mats);
            ctx->d_mats = nullptr; ctx->n_codes_alloc = 0;     // d_sphf has a fixed size and is never re-allocated
            PT_CUDA(ctx, cudaMalloc(&ctx->d_mats, sizeof(MatF32) * nc));
            ctx->n_codes_alloc = nc;
        }
        PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_mats, mats.data(), sizeof(MatF32) * nc, cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->grid.n > 0) {          // the grid's three arrays (grown, never shrunk)
            auto grow = [&](void **ptr, size_t &cap, size_t need_bytes) -> cudaError_t {
                if (cap >= need_bytes) return cudaSuccess;
                if (*ptr) cudaFree(*ptr);
                *ptr = nullptr; cap = 0;
                cudaError_t e = cudaMalloc(ptr, need_bytes);
                if (e == cudaSuccess) cap = need_bytes;
                return e;
            };
            PT_CUDA(ctx, grow((void **)&ctx->d_grid_start, ctx->grid_start_cap, ctx->h_grid_start.size() * sizeof(unsigned int)));
            PT_CUDA(ctx, grow((void **)&ctx->d_grid_items, ctx->grid_items_cap, std::max<size_t>(1, ctx->h_grid_items.size()) * sizeof(unsigned int)));
            PT_CUDA(ctx, grow((void **)&ctx->d_grid_sph, ctx->grid_sph_cap, ctx->h_grid_sph.size() * sizeof(float4)));
            PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_start, ctx->h_grid_start.data(), ctx->h_grid_start.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
            if (!ctx->h_grid_items.empty())
                PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_items, ctx->h_grid_items.data(), ctx->h_grid_items.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
            PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_sph, ctx->h_grid_sph.data(), ctx->h_grid_sph.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            ctx->grid.start = ctx->d_grid_start; ctx->grid.items = ctx->d_grid_items; ctx->grid.sph = ctx->d_grid_sph;
        }
        // global-memory mirror of the sphere scan table: k_bounce stages it in shared memory
        if (!ctx->d_sphf) PT_CUDA(ctx, cudaMalloc(&ctx->d_sphf, 2 * sizeof(float4) * (PT_MAX_OBJ + 4)));
        if (ctx->h_scene32->n_sph4 > 0) {
            // scan table, then (at PT_MAX_OBJ + 4) the exact {centre, r^2} table; padding entries are never candidates
            PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sphf, ctx->h_scene32->sphf, sizeof(float4) * ctx->h_scene32->n_sph4, cudaMemcpyHostToDevice, ctx->stream));
            PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sphf + PT_MAX_OBJ + 4, ctx->h_scene32->sph, sizeof(float4) * ctx->h_scene32->n_sph, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PT_OK;
}

int pt_scene_upload(pt_ctx **out, const pt_scene *scene, int device)
{
    if (!out || !scene) return pt_fail(nullptr, PT_ERR_ARG, "null argument");
    pt_ctx *reuse = *out;            // non-NULL: replace the scene of an existing context, keep its device buffers
    const int n = scene->n_spheres + scene->n_planes;
    if (scene->n_spheres < 0 || scene->n_planes < 0) return pt_fail(reuse, PT_ERR_ARG, "negative object count");
    if (n <= 0 || n > PT_MAX_OBJECTS) return pt_fail(reuse, PT_ERR_ARG, "object count must be in 1..16000");
    std::vector<DevObj64> objs;
    std::string why;
    if (flatten_scene(scene, objs, why) != PT_OK) return pt_fail(reuse, PT_ERR_ARG, why);
    if (reuse) {
        if (device >= 0 && device != reuse->device) return pt_fail(reuse, PT_ERR_ARG, "context lives on another device");
        PT_CUDA(reuse, cudaSetDevice(reuse->device));
        reuse->objs.swap(objs);
        reuse->cam = scene->camera;
        reuse->light = scene->light;
        reuse->rendered = false;
        return upload_tables(reuse);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return pt_fail(nullptr, PT_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) return pt_fail(nullptr, PT_ERR_ARG, "device index out of range");

    pt_ctx *ctx = new (std::nothrow) pt_ctx();
    if (!ctx) return pt_fail(nullptr, PT_ERR_OOM, "host allocation failed");
    ctx->device = device;
    ctx->objs.swap(objs);
    ctx->cam = scene->camera;
    ctx->light = scene->light;

#define UP_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            int rc_ = pt_fail(nullptr, PT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
            pt_destroy(ctx);                                                                            \
            return rc_;                                                                                 \
        }                                                                                               \
    } while (0)

    UP_CUDA(cudaSetDevice(device));
    UP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    UP_CUDA(cudaEventCreate(&ctx->ev0));
    UP_CUDA(cudaEventCreate(&ctx->ev1));
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->l2_bytes, cudaDevAttrL2CacheSize, device);
    UP_CUDA(cudaMalloc(&ctx->d_stats, sizeof(DevStats)));
    UP_CUDA(cudaMallocHost(&ctx->h_pinned, 2 * sizeof(unsigned int)));
    UP_CUDA(cudaMallocHost(&ctx->h_stats, sizeof(DevStats)));
    UP_CUDA(cudaEventCreateWithFlags(&ctx->ev_batch[0], cudaEventDisableTiming));
    UP_CUDA(cudaEventCreateWithFlags(&ctx->ev_batch[1], cudaEventDisableTiming));
#undef UP_CUDA
    if (const char *e = std::getenv("PTB200_JIT")) ctx->jit_mode = std::atoi(e) < 0 ? 0 : std::atoi(e) > 2 ? 2 : std::atoi(e);
    ctx->h_scene32 = new (std::nothrow) SceneF32;
    if (!ctx->h_scene32) { pt_destroy(ctx); return pt_fail(nullptr, PT_ERR_OOM, "host allocation failed"); }
    int rc = upload_tables(ctx);
    if (rc != PT_OK) { std::string msg = ctx->err; pt_destroy(ctx); return pt_fail(nullptr, rc, msg); }
    *out = ctx;
    return PT_OK;
}

static int render_common(pt_ctx *ctx, const pt_render_params *p, double *ext_sum, cudaStream_t ext_stream)
{
    if (!ctx || !p) return pt_fail(ctx, PT_ERR_ARG, "null argument");
    if (p->width <= 0 || p->height <= 0 || p->spp < 0 || p->width > 65535 || p->height > 65535)
        return pt_fail(ctx, PT_ERR_ARG, "width/height must be in 1..65535 and spp >= 0");
    if (p->sample_offset < 0) return pt_fail(ctx, PT_ERR_ARG, "sample_offset must be >= 0");
    if (p->accumulate && p->engine != PT_ENGINE_FP32_PHILOX)
        return pt_fail(ctx, PT_ERR_ARG, "accumulate = 1 is an FP32 engine feature (the erand48 replay consumes one sequential stream per row)");
    if (p->robust_eps && p->engine != PT_ENGINE_FP32_PHILOX)
        return pt_fail(ctx, PT_ERR_ARG, "robust_eps is an FP32 engine option (the FP64 engine replays the reference, which has no epsilon on rectangles)");
    if (p->mode < PT_MODE_NEE_REF_RECT || p->mode > PT_MODE_NEE_CONE_SPHERE) return pt_fail(ctx, PT_ERR_ARG, "bad mode");
    if (p->engine != PT_ENGINE_FP32_PHILOX && p->engine != PT_ENGINE_FP64_ERAND48) return pt_fail(ctx, PT_ERR_ARG, "bad engine");
    const int world = p->world > 0 ? p->world : 1;
    if (p->rank < 0 || p->rank >= world) return pt_fail(ctx, PT_ERR_ARG, "rank must be in [0, world)");
    // owned_rows_only writes ONLY this rank's rows of a (possibly shared) image; the statistics path resolves whole images
    if (p->owned_rows_only && p->collect_stats)
        return pt_fail(ctx, PT_ERR_ARG, "owned_rows_only does not combine with collect_stats (the per-pixel variance is resolved for the whole image)");
    if (p->mode == PT_MODE_NEE_REF_RECT && (ctx->light.id < 0 || ctx->light.id >= (int)ctx->objs.size()))
        return pt_fail(ctx, PT_ERR_ARG, "PT_MODE_NEE_REF_RECT needs pt_scene.light.id to name a scene object");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ext_stream ? ext_stream : ctx->stream;
    const size_t n_acc = (size_t)p->width * p->height * 3;
    if (ctx->accum_elems < n_acc) {
        if (ctx->d_sum) cudaFree(ctx->d_sum);
        if (ctx->d_sumsq) cudaFree(ctx->d_sumsq);
        ctx->d_sum = ctx->d_sumsq = nullptr; ctx->accum_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_sum, n_acc * sizeof(double)));
        PT_CUDA(ctx, cudaMalloc(&ctx->d_sumsq, n_acc * sizeof(double)));
        ctx->accum_elems = n_acc;
    }
    double *d_sum = ext_sum ? ext_sum : ctx->d_sum;
    ctx->d_sum_ext = ext_sum;
    if (!p->owned_rows_only) PT_CUDA(ctx, cudaMemsetAsync(d_sum, 0, n_acc * sizeof(double), s));   // owned_rows_only: other ranks write the rest
    if (p->collect_stats) PT_CUDA(ctx, cudaMemsetAsync(ctx->d_sumsq, 0, n_acc * sizeof(double), s));
    PT_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevStats), s));
    std::memset(&ctx->stats, 0, sizeof ctx->stats);
    if (p->accumulate && (!ctx->rendered_fp32 || ctx->accum_spp <= 0 || ctx->last.width != p->width || ctx->last.height != p->height))
        return pt_fail(ctx, PT_ERR_STATE, "accumulate = 1 needs a previous FP32 render (or pt_accum_upload) of the same image size");
    ctx->last = *p;
    ctx->rendered = false;
    ctx->jit = nullptr;
    if (p->engine == PT_ENGINE_FP32_PHILOX && ctx->fp32_ok && ctx->jit_mode > 0 && ctx->grid.n == 0) {      // (grid scenes run the ahead-of-time build)
        // scene-specialised kernel: built (once per scene/mode) BEFORE the timed region starts
        const unsigned long long total = (unsigned long long)p->width * p->height * (unsigned long long)p->spp / (unsigned long long)world;
        // Large renders (and jit_mode 2) wait for the build; small ones never do: their specialisation is compiled on a
        // host thread from the second render on and used once it is there.  Which build runs cannot be seen in the
        // image: the two perform the same operations in the same order (bit-identical, tested).
        // (the module is built for this render's regeneration flags too: row blocks, sample runs, one GPU or several)
        Fp32Plan pl;
        pt_fp32_plan(ctx, p, pl);
        ctx->jit_flags = std::getenv("PTB200_JIT_NO_RENDER_FLAGS") ? -1 : pl.flags;
        ctx->jit = pt_jit_get(ctx, p->mode, p->collect_stats != 0, false, ctx->jit_mode >= 2 || total >= PT_JIT_MIN_PATHS, ctx->jit_flags);
    }
    std::unique_lock<std::mutex> scene_lock(dev_mutex(ctx->device), std::defer_lock);
    if (p->engine == PT_ENGINE_FP32_PHILOX) scene_lock.lock();      // released on return, i.e. after the stream synchronised
    PT_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    int rc = p->engine == PT_ENGINE_FP64_ERAND48 ? pt_fp64_render(ctx, p, d_sum, p->collect_stats ? ctx->d_sumsq : nullptr, s)
                                                 : pt_fp32_render(ctx, p, d_sum, ctx->d_sumsq, s);
    if (rc) return rc;
    PT_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    PT_CUDA(ctx, cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
    PT_CUDA(ctx, cudaStreamSynchronize(s));
    float ms = 0.f;
    PT_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const DevStats ds = *ctx->h_stats;
    pt_stats &st = ctx->stats;
    st.render_ms = ms;
    st.rays_shadow = ds.rays_shadow; st.miss_events = ds.misses; st.truncated = ds.truncated;
    st.shaded_vertices = ds.shaded; st.max_depth_seen = ds.max_depth_seen;
    st.term_roulette = ds.term_roulette; st.term_emitter = ds.term_emitter; st.term_light_sample = ds.term_light_sample;
    st.dropped_contributions = ds.dropped; st.spawned_branches = ds.spawned; st.split_refusals = ds.split_refused;
    for (int k = 0; k < 64; k++) st.live_at_depth[k] = ds.live_at_depth[k];
    st.specialised = ctx->jit ? 1u : 0u;
    st.accel_structure = (p->engine == PT_ENGINE_FP32_PHILOX && ctx->grid.n > 0) ? 1u : 0u;
    if (p->engine == PT_ENGINE_FP32_PHILOX && ctx->fp32_ok && ctx->jit_mode == 1 && !ctx->jit)
        pt_jit_account(ctx, p->mode, p->collect_stats != 0, ms, ctx->jit_flags);      // small render, generic kernel: counts towards its background build
    if (p->engine == PT_ENGINE_FP64_ERAND48) {
        st.paths = ds.paths; st.rays_camera = ds.rays_camera; st.rays_scatter = ds.rays_scatter;
    } else {
        // every generated path is one camera ray; scatter rays are the continuing paths counted on the device
        long long owned_rows = 0;
        const int tile = p->tile_rows > 0 ? p->tile_rows : 8, n_tiles = (p->height + tile - 1) / tile;
        for (int k = p->rank; k < n_tiles; k += world) owned_rows += (k * tile + tile <= p->height) ? tile : (p->height - k * tile);
        st.paths = (uint64_t)owned_rows * p->width * p->spp;
        st.rays_camera = st.paths;
        st.rays_scatter = ds.rays_scatter;
    }
    ctx->rendered = true;
    ctx->rendered_fp32 = p->engine == PT_ENGINE_FP32_PHILOX;
    if (!ctx->rendered_fp32) ctx->accum_spp = p->spp;
    return PT_OK;
}

int pt_render(pt_ctx *ctx, const pt_render_params *p) { return render_common(ctx, p, nullptr, nullptr); }

// Single-process multi-GPU: contexts on n devices (same scene), one host thread each; every device renders its row tiles
// (tile k -> device k % n) and its resolve kernel stores them into ctxs[0]'s accumulation buffer over NVLink peer memory.
int pt_render_multi(pt_ctx **ctxs, int n, const pt_render_params *p)
{
    if (!ctxs || n <= 0 || !p) return pt_fail(nullptr, PT_ERR_ARG, "bad argument");
    for (int i = 0; i < n; i++) if (!ctxs[i]) return pt_fail(nullptr, PT_ERR_ARG, "null context");
    if (n == 1) return render_common(ctxs[0], p, nullptr, nullptr);
    pt_ctx *root = ctxs[0];
    if (p->collect_stats || p->accumulate) return pt_fail(root, PT_ERR_ARG, "pt_render_multi does not combine with collect_stats / accumulate");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (ctxs[i]->device == ctxs[j]->device) return pt_fail(root, PT_ERR_ARG, "pt_render_multi needs one context per device");
    // the root's image: (re)allocated here so that the workers can write to it
    PT_CUDA(root, cudaSetDevice(root->device));
    const size_t n_acc = (size_t)p->width * p->height * 3;
    if (p->width <= 0 || p->height <= 0) return pt_fail(root, PT_ERR_ARG, "bad image size");
    if (root->accum_elems < n_acc) {
        if (root->d_sum) cudaFree(root->d_sum);
        if (root->d_sumsq) cudaFree(root->d_sumsq);
        root->d_sum = root->d_sumsq = nullptr; root->accum_elems = 0;
        PT_CUDA(root, cudaMalloc(&root->d_sum, n_acc * sizeof(double)));
        PT_CUDA(root, cudaMalloc(&root->d_sumsq, n_acc * sizeof(double)));
        root->accum_elems = n_acc;
    }
    for (int i = 1; i < n; i++) {           // peer access device i -> root's device
        int can = 0;
        PT_CUDA(root, cudaDeviceCanAccessPeer(&can, ctxs[i]->device, root->device));
        if (!can) return pt_fail(root, PT_ERR_STATE, "no peer access between the devices");
        PT_CUDA(root, cudaSetDevice(ctxs[i]->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return pt_fail(root, PT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    std::vector<int> rc(n, PT_OK);
    std::vector<std::thread> th;
    double *target = root->d_sum;
    for (int i = 0; i < n; i++) {
        th.emplace_back([=, &rc] {
            pt_render_params q = *p;
            q.rank = i; q.world = n; q.owned_rows_only = 1;
            if (q.tile_rows <= 0) q.tile_rows = 8;
            rc[i] = render_common(ctxs[i], &q, target, nullptr);
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i] != PT_OK) return pt_fail(root, rc[i], std::string("device ") + std::to_string(ctxs[i]->device) + ": " + ctxs[i]->err);
    // the root answers pt_readback for the whole image; its statistics are the sums over the devices
    pt_stats tot = root->stats;
    for (int i = 1; i < n; i++) {
        const pt_stats &a = ctxs[i]->stats;
        tot.paths += a.paths; tot.rays_camera += a.rays_camera; tot.rays_scatter += a.rays_scatter; tot.rays_shadow += a.rays_shadow;
        tot.shaded_vertices += a.shaded_vertices; tot.miss_events += a.miss_events; tot.truncated += a.truncated;
        tot.kernel_launches += a.kernel_launches; tot.queue_slots_io += a.queue_slots_io;
        tot.iterations = std::max(tot.iterations, a.iterations);
        tot.max_depth_seen = std::max(tot.max_depth_seen, a.max_depth_seen);
        tot.render_ms = std::max(tot.render_ms, a.render_ms);
    }
    root->stats = tot;
    root->d_sum_ext = nullptr;
    root->last = *p;
    root->last.rank = 0; root->last.world = 1; root->last.owned_rows_only = 0;
    return PT_OK;
}

int pt_debug_stats(pt_ctx *ctx, pt_stats *stats)
{
    if (!ctx || !stats) return PT_ERR_ARG;
    *stats = ctx->stats;
    return PT_OK;
}

int pt_set_acceleration(pt_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return pt_fail(ctx, PT_ERR_ARG, "acceleration mode must be 0, 1 or 2");
    ctx->accel_mode = mode;
    return PT_OK;
}

int pt_set_specialisation(pt_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return pt_fail(ctx, PT_ERR_ARG, "specialisation mode must be 0, 1 or 2");
    ctx->jit_mode = mode;
    return PT_OK;
}

int pt_debug_specialise(const pt_scene *scene, int mode, char *spec_out, size_t spec_cap, size_t *cubin_bytes, double *seconds)
{
    // host only: flatten + class-sort the scene exactly as pt_scene_upload does, then run NVRTC (no device needed)
    if (!scene) return pt_fail(nullptr, PT_ERR_ARG, "null argument");
    pt_ctx tmp;
    std::string why;
    if (flatten_scene(scene, tmp.objs, why) != PT_OK) return pt_fail(nullptr, PT_ERR_ARG, why);
    tmp.cam = scene->camera;
    tmp.light = scene->light;
    std::vector<MatF32> mats;
    SceneF32 *S = new (std::nothrow) SceneF32;
    if (!S) return pt_fail(nullptr, PT_ERR_OOM, "host allocation failed");
    tmp.h_scene32 = S;
    build_scene_f32(&tmp, mats);
    int rc = PT_OK;
    if (!tmp.fp32_ok) rc = pt_fail(nullptr, PT_ERR_ARG, "scene does not fit the FP32 engine: " + tmp.fp32_why);
    else {
        // (bits 8 and up of `mode`: layout flags + 1 of the render the module is for, 0 = none)
        const std::string spec = pt_jit_spec(*S, mode & 0xFF, false, false, (mode >> 8) - 1);
        if (spec_out && spec_cap) { std::strncpy(spec_out, spec.c_str(), spec_cap - 1); spec_out[spec_cap - 1] = 0; }
        std::vector<char> cubin;
        std::string log;
        rc = pt_jit_build(spec, cubin, log, seconds, nullptr);
        if (rc != PT_OK) rc = pt_fail(nullptr, rc, "NVRTC: " + log);
        else if (cubin_bytes) *cubin_bytes = cubin.size();
    }
    tmp.h_scene32 = nullptr;
    delete S;
    return rc;
}

int pt_debug_plan(const pt_scene *scene, const pt_render_params *params, int sm_count, pt_plan_info *out)
{
    // host only: the scene's material mask as pt_scene_upload derives it, then the pure layout function of the FP32 engine
    if (!scene || !params || !out || sm_count <= 0) return pt_fail(nullptr, PT_ERR_ARG, "null argument");
    if (params->width <= 0 || params->height <= 0 || params->spp < 0 || params->width > 65535 || params->height > 65535)
        return pt_fail(nullptr, PT_ERR_ARG, "bad render size");
    pt_ctx tmp;
    std::string why;
    if (flatten_scene(scene, tmp.objs, why) != PT_OK) return pt_fail(nullptr, PT_ERR_ARG, why);
    tmp.cam = scene->camera;
    tmp.light = scene->light;
    std::vector<MatF32> mats;
    SceneF32 *S = new (std::nothrow) SceneF32;
    if (!S) return pt_fail(nullptr, PT_ERR_OOM, "host allocation failed");
    tmp.h_scene32 = S;
    build_scene_f32(&tmp, mats);
    tmp.sm_count = sm_count;
    Fp32Plan pl;
    pt_fp32_plan(&tmp, params, pl);
    out->owned_rows = (uint64_t)pl.owned_rows; out->owned_pixels = pl.owned_pixels; out->row_blocks = pl.n_blk; out->block_rows = pl.blk_rows;
    out->path_slots = (uint64_t)pl.cap; out->run_length = 1ull << pl.run_shift; out->path_indices = pl.total;
    out->layout_flags = (uint32_t)pl.flags; out->splits_refr_paths = pl.want_spawn ? 1u : 0u;
    tmp.h_scene32 = nullptr;
    delete S;
    return PT_OK;
}

int pt_render_into(pt_ctx *ctx, const pt_render_params *p, void *dev_rgb_sum, void *stream)
{
    if (!dev_rgb_sum) return pt_fail(ctx, PT_ERR_ARG, "null device buffer");
    return render_common(ctx, p, (double *)dev_rgb_sum, (cudaStream_t)stream);
}

int pt_device_alloc(pt_ctx *ctx, size_t bytes, void **dev_ptr)
{
    if (!ctx || !dev_ptr || bytes == 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaMalloc(dev_ptr, bytes));
    return PT_OK;
}

int pt_device_free(pt_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return PT_ERR_ARG;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaFree(dev_ptr));
    return PT_OK;
}

int pt_ipc_export(pt_ctx *ctx, const void *dev_ptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ctx || !dev_ptr || !handle) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    PT_CUDA(ctx, cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    std::memcpy(handle, &h, 64);
    return PT_OK;
}

int pt_ipc_open(pt_ctx *ctx, const unsigned char handle[64], void **dev_ptr)
{
    if (!ctx || !dev_ptr || !handle) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    PT_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PT_OK;
}

int pt_ipc_close(pt_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return PT_ERR_ARG;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return PT_OK;
}

// per-pixel SUM -> MEAN on the device (the division by `samps` of src/smallpt.cpp:536), so the host only copies
__global__ void k_scale(const double *__restrict__ in, double *__restrict__ out, size_t n, double scale)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * scale;
}

// the same for the rows of one rank only (row tiles k % world == rank): pt_readback_owned moves nothing else
__global__ void k_scale_owned(const double *__restrict__ in, double *__restrict__ out, unsigned long long owned_elems, int row_elems, int tile_rows, int rank, int world, double scale)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= owned_elems) return;
    const unsigned int row_local = (unsigned int)(i / (unsigned int)row_elems), e = (unsigned int)(i - (unsigned long long)row_local * (unsigned int)row_elems);
    const unsigned int tile = row_local / (unsigned int)tile_rows;
    const unsigned int y = (tile * (unsigned int)world + (unsigned int)rank) * (unsigned int)tile_rows + (row_local - tile * (unsigned int)tile_rows);
    const size_t idx = (size_t)y * (size_t)row_elems + e;
    out[idx] = in[idx] * scale;
}

// device -> caller's (pageable) buffer.  Each LANE owns a stream and two pinned staging blocks: the copy of block k+1
// over PCIe overlaps the host memcpy of block k.  One lane's memcpy (~10 GB/s) is slower than PCIe, so images of 4 MB
// and more are split over PT_STAGE_LANES lanes, each driven by its own host thread.
static cudaError_t lane_d2h(pt_ctx *ctx, int lane, double *dst, const double *d_src, size_t n)
{
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return e;
    PtStageLane &L = ctx->lane[lane];
    const size_t blk = PT_STAGE_ELEMS;
    const size_t nblk = (n + blk - 1) / blk;
    auto issue = [&](size_t k) {
        const size_t off = k * blk, cnt = std::min(blk, n - off);
        cudaError_t e2 = cudaMemcpyAsync(L.h_stage + (k & 1) * blk, d_src + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, L.stream);
        if (e2 == cudaSuccess) e2 = cudaEventRecord(L.ev[k & 1], L.stream);
        return e2;
    };
    if (nblk && (e = issue(0)) != cudaSuccess) return e;
    for (size_t k = 0; k < nblk; k++) {
        if (k + 1 < nblk && (e = issue(k + 1)) != cudaSuccess) return e;
        if ((e = cudaEventSynchronize(L.ev[k & 1])) != cudaSuccess) return e;
        const size_t off = k * blk, cnt = std::min(blk, n - off);
        std::memcpy(dst + off, L.h_stage + (k & 1) * blk, cnt * sizeof(double));
    }
    return cudaSuccess;
}

static int staged_d2h(pt_ctx *ctx, double *dst, const double *d_src, size_t n)
{
    for (int l = 0; l < PT_STAGE_LANES; l++) {          // lanes are created on first use
        PtStageLane &L = ctx->lane[l];
        if (L.stream) continue;
        PT_CUDA(ctx, cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        PT_CUDA(ctx, cudaMallocHost(&L.h_stage, 2 * PT_STAGE_ELEMS * sizeof(double)));
        PT_CUDA(ctx, cudaEventCreateWithFlags(&L.ev[0], cudaEventDisableTiming));
        PT_CUDA(ctx, cudaEventCreateWithFlags(&L.ev[1], cudaEventDisableTiming));
    }
    // the lanes' streams start after everything queued on the context's stream (the sum -> mean kernel)
    if (!ctx->ev_rb) PT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rb, cudaEventDisableTiming));
    PT_CUDA(ctx, cudaEventRecord(ctx->ev_rb, ctx->stream));
    const int lanes = n >= 4 * PT_STAGE_ELEMS ? PT_STAGE_LANES : 1;
    for (int l = 0; l < lanes; l++) PT_CUDA(ctx, cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev_rb, 0));
    if (lanes == 1) {
        cudaError_t e = lane_d2h(ctx, 0, dst, d_src, n);
        if (e != cudaSuccess) return pt_fail(ctx, PT_ERR_CUDA, std::string("pt_readback: ") + cudaGetErrorString(e));
        return PT_OK;
    }
    // contiguous parts, block-aligned
    const size_t per = ((n + lanes - 1) / lanes + PT_STAGE_ELEMS - 1) / PT_STAGE_ELEMS * PT_STAGE_ELEMS;
    cudaError_t err[PT_STAGE_LANES];
    std::thread th[PT_STAGE_LANES];
    for (int l = 0; l < lanes; l++) {
        const size_t off = std::min(n, (size_t)l * per), cnt = std::min(per, n - off);
        err[l] = cudaSuccess;
        if (l == 0 || cnt == 0) continue;
        th[l] = std::thread([=, &err] { err[l] = lane_d2h(ctx, l, dst + off, d_src + off, cnt); });
    }
    err[0] = lane_d2h(ctx, 0, dst, d_src, std::min(per, n));
    for (int l = 1; l < lanes; l++) if (th[l].joinable()) th[l].join();
    for (int l = 0; l < lanes; l++)
        if (err[l] != cudaSuccess) return pt_fail(ctx, PT_ERR_CUDA, std::string("pt_readback: ") + cudaGetErrorString(err[l]));
    return PT_OK;
}

int pt_readback(pt_ctx *ctx, double *rgb_mean, double *rgb_sumsq, pt_stats *stats)
{
    if (!ctx) return pt_fail(ctx, PT_ERR_ARG, "null context");
    if (!ctx->rendered) return pt_fail(ctx, PT_ERR_STATE, "pt_readback before a successful pt_render");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const pt_render_params &p = ctx->last;
    const size_t n = (size_t)p.width * p.height * 3;
    if (rgb_mean) {
        const double *src = ctx->d_sum_ext ? ctx->d_sum_ext : ctx->d_sum;
        if (p.spp > 0) {
            if (ctx->mean_elems < n) {
                if (ctx->d_mean) cudaFree(ctx->d_mean);
                ctx->d_mean = nullptr; ctx->mean_elems = 0;     // h_view (pt_readback_view) is independent of d_mean
                PT_CUDA(ctx, cudaMalloc(&ctx->d_mean, n * sizeof(double)));
                ctx->mean_elems = n;
            }
            k_scale<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->d_mean, n, 1.0 / (double)(ctx->accum_spp > 0 ? ctx->accum_spp : p.spp));
            PT_CUDA(ctx, cudaGetLastError());
            src = ctx->d_mean;
        }
        int rc = staged_d2h(ctx, rgb_mean, src, n);
        if (rc) return rc;
    }
    if (rgb_sumsq) {
        if (!p.collect_stats) return pt_fail(ctx, PT_ERR_STATE, "sum of squares requested but collect_stats was 0");
        int rc = staged_d2h(ctx, rgb_sumsq, ctx->d_sumsq, n);
        if (rc) return rc;
    }
    if (stats) *stats = ctx->stats;
    return PT_OK;
}

// exact double <-> 2^-24 fixed point (the library's own sums are multiples of 2^-24 below 2^40)
__global__ void k_to_fix(const double *__restrict__ in, unsigned long long *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)__double2ull_rn(in[i] * 16777216.0);
}

int pt_accum_upload(pt_ctx *ctx, int width, int height, const double *rgb_sum, const double *rgb_sumsq, int spp_done)
{
    if (!ctx || !rgb_sum || width <= 0 || height <= 0 || spp_done <= 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)width * height * 3;
    if (ctx->fix_elems < n) {
        if (ctx->d_fix) cudaFree(ctx->d_fix);
        if (ctx->d_fixsq) cudaFree(ctx->d_fixsq);
        ctx->d_fix = ctx->d_fixsq = nullptr; ctx->fix_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fix, n * sizeof(unsigned long long)));
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fixsq, n * sizeof(unsigned long long)));
        ctx->fix_elems = n;
    }
    if (ctx->accum_elems < n) {
        if (ctx->d_sum) cudaFree(ctx->d_sum);
        if (ctx->d_sumsq) cudaFree(ctx->d_sumsq);
        ctx->d_sum = ctx->d_sumsq = nullptr; ctx->accum_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_sum, n * sizeof(double)));
        PT_CUDA(ctx, cudaMalloc(&ctx->d_sumsq, n * sizeof(double)));
        ctx->accum_elems = n;
    }
    PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sum, rgb_sum, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_to_fix<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_sum, ctx->d_fix, n);
    if (rgb_sumsq) {
        PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sumsq, rgb_sumsq, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        k_to_fix<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_sumsq, ctx->d_fixsq, n);
    }
    PT_CUDA(ctx, cudaGetLastError());
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->fix_has_sq = rgb_sumsq != nullptr;
    ctx->accum_spp = spp_done;
    ctx->d_sum_ext = nullptr;
    std::memset(&ctx->last, 0, sizeof ctx->last);
    ctx->last.width = width; ctx->last.height = height; ctx->last.spp = spp_done; ctx->last.world = 1;
    ctx->last.collect_stats = rgb_sumsq ? 1 : 0;
    ctx->rendered = true;
    ctx->rendered_fp32 = true;
    return PT_OK;
}

int pt_accum_download(pt_ctx *ctx, double *rgb_sum, double *rgb_sumsq, int *spp_done)
{
    if (!ctx) return pt_fail(ctx, PT_ERR_ARG, "null context");
    if (!ctx->rendered) return pt_fail(ctx, PT_ERR_STATE, "pt_accum_download before a successful pt_render");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->last.width * ctx->last.height * 3;
    if (rgb_sum) {
        const double *src = ctx->d_sum_ext ? ctx->d_sum_ext : ctx->d_sum;
        PT_CUDA(ctx, cudaMemcpyAsync(rgb_sum, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (rgb_sumsq) {
        if (!ctx->last.collect_stats) return pt_fail(ctx, PT_ERR_STATE, "sum of squares requested but collect_stats was 0");
        PT_CUDA(ctx, cudaMemcpyAsync(rgb_sumsq, ctx->d_sumsq, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (spp_done) *spp_done = (int)ctx->accum_spp;
    return PT_OK;
}

// Zero-copy read-back: the mean image in library-owned pinned host memory (one DMA, no host memcpy).
const double *pt_readback_view(pt_ctx *ctx, pt_stats *stats)
{
    if (!ctx) { pt_fail(ctx, PT_ERR_ARG, "null context"); return nullptr; }
    if (!ctx->rendered) { pt_fail(ctx, PT_ERR_STATE, "pt_readback_view before a successful pt_render"); return nullptr; }
    auto bad = [&](cudaError_t e, const char *what) { pt_fail(ctx, PT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); return (const double *)nullptr; };
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return bad(e, "cudaSetDevice");
    const pt_render_params &p = ctx->last;
    const size_t n = (size_t)p.width * p.height * 3;
    if (ctx->view_elems < n) {
        if (ctx->h_view) cudaFreeHost(ctx->h_view);
        ctx->h_view = nullptr; ctx->view_elems = 0;
        if ((e = cudaMallocHost(&ctx->h_view, n * sizeof(double))) != cudaSuccess) return bad(e, "cudaMallocHost");
        ctx->view_elems = n;
    }
    if (ctx->mean_elems < n) {
        if (ctx->d_mean) cudaFree(ctx->d_mean);
        ctx->d_mean = nullptr; ctx->mean_elems = 0;
        if ((e = cudaMalloc(&ctx->d_mean, n * sizeof(double))) != cudaSuccess) return bad(e, "cudaMalloc");
        ctx->mean_elems = n;
    }
    const double *src = ctx->d_sum_ext ? ctx->d_sum_ext : ctx->d_sum;
    const double spp = (double)(ctx->accum_spp > 0 ? ctx->accum_spp : p.spp);
    k_scale<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->d_mean, n, spp > 0 ? 1.0 / spp : 1.0);
    if ((e = cudaMemcpyAsync(ctx->h_view, ctx->d_mean, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bad(e, "cudaMemcpyAsync");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bad(e, "cudaStreamSynchronize");
    if (stats) *stats = ctx->stats;
    return ctx->h_view;
}

int pt_host_register(pt_ctx *ctx, void *host_ptr, size_t bytes)
{
    if (!ctx || !host_ptr || bytes == 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return PT_OK;
}

int pt_host_unregister(pt_ctx *ctx, void *host_ptr)
{
    if (!ctx || !host_ptr) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaHostUnregister(host_ptr));
    return PT_OK;
}

// The rows this rank owns (row tiles k % world == rank of the last render), as means, into the caller's FULL-SIZE
// host image: one strided DMA for the rank's full tiles, each rank over its own PCIe link.  With the image in memory every rank maps
// (POSIX shared memory, registered through pt_host_register) the host-side image assembles without a gather.
int pt_readback_owned(pt_ctx *ctx, double *host_image, pt_stats *stats)
{
    if (!ctx || !host_image) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    if (!ctx->rendered) return pt_fail(ctx, PT_ERR_STATE, "pt_readback_owned before a successful pt_render");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const pt_render_params &p = ctx->last;
    if (ctx->d_sum_ext && p.owned_rows_only)
        return pt_fail(ctx, PT_ERR_STATE, "the last render stored its rows into a shared image (owned_rows_only): read that image instead");
    const size_t n = (size_t)p.width * p.height * 3;
    if (ctx->mean_elems < n) {
        if (ctx->d_mean) cudaFree(ctx->d_mean);
        ctx->d_mean = nullptr; ctx->mean_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_mean, n * sizeof(double)));
        ctx->mean_elems = n;
    }
    const double *src = ctx->d_sum_ext ? ctx->d_sum_ext : ctx->d_sum;
    const double spp = (double)(ctx->accum_spp > 0 ? ctx->accum_spp : p.spp);
    const int world = p.world > 0 ? p.world : 1, tile = p.tile_rows > 0 ? p.tile_rows : 8;
    const int n_tiles = (p.height + tile - 1) / tile;
    const size_t row = (size_t)p.width * 3;
    {   // sums -> means, this rank's rows only
        unsigned long long owned_rows = 0;
        for (int k = p.rank; k < n_tiles; k += world) owned_rows += (unsigned long long)std::min(tile, p.height - k * tile);
        const unsigned long long owned_elems = owned_rows * row;
        if (owned_elems > 0)
            k_scale_owned<<<(unsigned)((owned_elems + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->d_mean, owned_elems, (int)row, tile, p.rank, world, spp > 0 ? 1.0 / spp : 1.0);
        PT_CUDA(ctx, cudaGetLastError());
    }
    // the full tiles lie a constant world * tile rows apart: ONE strided DMA (a tile = one "row" of the 2D copy), then the
    // ragged last tile of the image if it is this rank's
    int n_full = 0, k_last = -1;
    for (int k = p.rank; k < n_tiles; k += world) { if (k * tile + tile <= p.height) n_full++; else k_last = k; }
    const size_t pitch = (size_t)world * tile * row * sizeof(double), width = (size_t)tile * row * sizeof(double), y_first = (size_t)p.rank * tile;
    size_t max_pitch = 0;
    { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMaxPitch, ctx->device); max_pitch = (size_t)v; }
    if (n_full > 0 && pitch <= max_pitch)
        PT_CUDA(ctx, cudaMemcpy2DAsync(host_image + y_first * row, pitch, ctx->d_mean + y_first * row, pitch, width, (size_t)n_full, cudaMemcpyDeviceToHost, ctx->stream));
    else
        for (int k = p.rank, i = 0; i < n_full; k += world, i++)
            PT_CUDA(ctx, cudaMemcpyAsync(host_image + (size_t)k * tile * row, ctx->d_mean + (size_t)k * tile * row, width, cudaMemcpyDeviceToHost, ctx->stream));
    if (k_last >= 0) {
        const size_t y0 = (size_t)k_last * tile;
        PT_CUDA(ctx, cudaMemcpyAsync(host_image + y0 * row, ctx->d_mean + y0 * row, ((size_t)p.height - y0) * row * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats) *stats = ctx->stats;
    return PT_OK;
}

void *pt_accum_device_ptr(pt_ctx *ctx) { return ctx ? (void *)(ctx->d_sum_ext ? ctx->d_sum_ext : ctx->d_sum) : nullptr; }

int pt_debug_intersect(pt_ctx *ctx, const double *rays_od, int n, int precision, double *t_out, int *id_out)
{
    if (!ctx || !rays_od || !t_out || !id_out || n < 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    if (precision != 32 && precision != 64) return pt_fail(ctx, PT_ERR_ARG, "precision must be 32 or 64");
    if (n == 0) return PT_OK;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    double *d_r = nullptr, *d_t = nullptr;
    int *d_id = nullptr;
    PT_CUDA(ctx, cudaMalloc(&d_r, sizeof(double) * 6 * (size_t)n));
    PT_CUDA(ctx, cudaMalloc(&d_t, sizeof(double) * (size_t)n));
    PT_CUDA(ctx, cudaMalloc(&d_id, sizeof(int) * (size_t)n));
    int rc = PT_OK;
    std::lock_guard<std::mutex> scene_lock(dev_mutex(ctx->device));
    cudaError_t e = cudaMemcpyAsync(d_r, rays_od, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        rc = precision == 64 ? pt_fp64_intersect(ctx, d_r, n, d_t, d_id, ctx->stream) : pt_fp32_intersect(ctx, d_r, n, d_t, d_id, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaMemcpyAsync(t_out, d_t, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaMemcpyAsync(id_out, d_id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_r); cudaFree(d_t); cudaFree(d_id);
    if (rc != PT_OK) return rc;
    if (e != cudaSuccess) return pt_fail(ctx, PT_ERR_CUDA, std::string("pt_debug_intersect: ") + cudaGetErrorString(e));
    return PT_OK;
}

int pt_debug_erand48(pt_ctx *ctx, const uint16_t *seeds, int n_threads, int draws, double *out)
{
    if (!ctx || !seeds || !out || n_threads <= 0 || draws <= 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    uint16_t *d_s = nullptr;
    double *d_o = nullptr;
    PT_CUDA(ctx, cudaMalloc(&d_s, sizeof(uint16_t) * 3 * (size_t)n_threads));
    PT_CUDA(ctx, cudaMalloc(&d_o, sizeof(double) * (size_t)n_threads * draws));
    cudaError_t e = cudaMemcpyAsync(d_s, seeds, sizeof(uint16_t) * 3 * (size_t)n_threads, cudaMemcpyHostToDevice, ctx->stream);
    int rc = PT_OK;
    if (e == cudaSuccess) rc = pt_fp64_erand48(ctx, d_s, n_threads, draws, d_o, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, sizeof(double) * (size_t)n_threads * draws, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_s); cudaFree(d_o);
    if (rc != PT_OK) return rc;
    if (e != cudaSuccess) return pt_fail(ctx, PT_ERR_CUDA, std::string("pt_debug_erand48: ") + cudaGetErrorString(e));
    return PT_OK;
}

static int debug_philox(pt_ctx *ctx, const uint32_t *ctr, const uint32_t *key, int n, uint32_t *out, int width)
{
    if (!ctx || !ctr || !key || !out || n <= 0) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nc = (size_t)width * n, nk = (size_t)(width / 2) * n;
    uint32_t *d_c = nullptr, *d_k = nullptr, *d_o = nullptr;
    PT_CUDA(ctx, cudaMalloc(&d_c, sizeof(uint32_t) * nc));
    PT_CUDA(ctx, cudaMalloc(&d_k, sizeof(uint32_t) * nk));
    PT_CUDA(ctx, cudaMalloc(&d_o, sizeof(uint32_t) * nc));
    cudaError_t e = cudaMemcpyAsync(d_c, ctr, sizeof(uint32_t) * nc, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_k, key, sizeof(uint32_t) * nk, cudaMemcpyHostToDevice, ctx->stream);
    int rc = PT_OK;
    if (e == cudaSuccess) rc = pt_fp32_philox(ctx, d_c, d_k, n, d_o, ctx->stream, width);
    if (rc == PT_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, sizeof(uint32_t) * nc, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == PT_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_c); cudaFree(d_k); cudaFree(d_o);
    if (rc != PT_OK) return rc;
    if (e != cudaSuccess) return pt_fail(ctx, PT_ERR_CUDA, std::string("pt_debug_philox: ") + cudaGetErrorString(e));
    return PT_OK;
}

int pt_debug_philox(pt_ctx *ctx, const uint32_t *ctr, const uint32_t *key, int n, uint32_t *out) { return debug_philox(ctx, ctr, key, n, out, 4); }

int pt_debug_ffma_peak(pt_ctx *ctx, double *tflops, double *sm_clock_mhz)
{
    if (!ctx || !tflops || !sm_clock_mhz) return pt_fail(ctx, PT_ERR_ARG, "bad argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    return pt_fp32_ffma_peak(ctx, tflops, sm_clock_mhz);
}

void pt_destroy(pt_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 4; b++) if (ctx->q[a][b]) cudaFree(ctx->q[a][b]);
    if (ctx->d_warp_chunk) cudaFree(ctx->d_warp_chunk);
    if (ctx->d_spawn) cudaFree(ctx->d_spawn);
    if (ctx->d_grid_start) cudaFree(ctx->d_grid_start);
    if (ctx->d_grid_items) cudaFree(ctx->d_grid_items);
    if (ctx->d_grid_sph) cudaFree(ctx->d_grid_sph);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->d_launch_rec) cudaFree(ctx->d_launch_rec);
    if (ctx->d_stamps) cudaFree(ctx->d_stamps);
    if (ctx->d_fix) cudaFree(ctx->d_fix);
    if (ctx->d_fixsq) cudaFree(ctx->d_fixsq);
    if (ctx->d_sum) cudaFree(ctx->d_sum);
    if (ctx->d_sumsq) cudaFree(ctx->d_sumsq);
    if (ctx->d_mean) cudaFree(ctx->d_mean);
    if (ctx->h_view) cudaFreeHost(ctx->h_view);
    if (ctx->ev_rb) cudaEventDestroy(ctx->ev_rb);
    for (int l = 0; l < PT_STAGE_LANES; l++) {
        PtStageLane &L = ctx->lane[l];
        if (L.stream) { cudaStreamSynchronize(L.stream); cudaStreamDestroy(L.stream); }
        if (L.ev[0]) cudaEventDestroy(L.ev[0]);
        if (L.ev[1]) cudaEventDestroy(L.ev[1]);
        if (L.h_stage) cudaFreeHost(L.h_stage);
    }
    if (ctx->d_objs) cudaFree(ctx->d_objs);
    if (ctx->d_mats) cudaFree(ctx->d_mats);
    if (ctx->d_sphf) cudaFree(ctx->d_sphf);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
    if (ctx->ev_batch[0]) cudaEventDestroy(ctx->ev_batch[0]);
    if (ctx->ev_batch[1]) cudaEventDestroy(ctx->ev_batch[1]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx->h_scene32;
    delete ctx;
}

}  // extern "C"
