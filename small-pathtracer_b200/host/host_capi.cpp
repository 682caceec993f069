// host_capi.cpp — C exports of the host surface (scene tables, Camera, toInt, P3 writer) so that the
// Python test/bench plumbing uses the SAME C++ definitions as the `smallpt` executable.
// No compute here: nothing in this file touches the hot path.
#include <cstring>

#include "scene_io.hpp"

using namespace smallpt_b200;

extern "C" {

int spt_scene_counts(const char *name, int *n_spheres, int *n_planes)
{
    try {
        SceneTable s = scene_by_name(name);
        *n_spheres = int(s.spheres.size());
        *n_planes = int(s.planes.size());
        return 0;
    } catch (...) { return -1; }
}

int spt_scene_fill(const char *name, pt_sphere *spheres, pt_plane *planes, int *order, pt_light *light)
{
    try {
        SceneTable s = scene_by_name(name);
        if (!s.spheres.empty()) std::memcpy(spheres, s.spheres.data(), s.spheres.size() * sizeof(pt_sphere));
        if (!s.planes.empty()) std::memcpy(planes, s.planes.data(), s.planes.size() * sizeof(pt_plane));
        std::memcpy(order, s.order.data(), s.order.size() * sizeof(int));
        *light = s.light;
        return 0;
    } catch (...) { return -1; }
}

void spt_camera(const double *lookfrom, const double *lookat, const double *vup, float vfov, float aspect, pt_camera *out)
{
    Camera cam(Vec(lookfrom[0], lookfrom[1], lookfrom[2]), Vec(lookat[0], lookat[1], lookat[2]),
               Vec(vup[0], vup[1], vup[2]), vfov, aspect);
    *out = cam.pod();
}

// The camera of the reference's main(), src/smallpt.cpp:521.
void spt_builtin_camera(int w, int h, pt_camera *out)
{
    Camera cam(LOOKFROM, Vec(50, 40, 5), Vec(0, 1, 0), 65, float(w) / float(h));
    *out = cam.pod();
}

void spt_plane_tilted(const double *p0, const double *n, const double *along, double hs, double ht,
                      const double *e, const double *c, int refl, pt_plane *out)
{
    Plane p(Vec(p0[0], p0[1], p0[2]), Vec(n[0], n[1], n[2]), Vec(along[0], along[1], along[2]), hs, ht,
            Vec(e[0], e[1], e[2]), Vec(c[0], c[1], c[2]), Refl_t(refl));
    *out = p.as_plane();
}

double spt_clamp(double x) { return clamp(x); }
int spt_toInt(double x) { return toInt(x); }

int spt_write_ppm(const char *path, const double *rgb, int w, int h)
{
    try { write_ppm(path, rgb, w, h); return 0; } catch (...) { return -1; }
}

int spt_write_ppm_binary(const char *path, const double *rgb, int w, int h)
{
    try { write_ppm_binary(path, rgb, w, h); return 0; } catch (...) { return -1; }
}

int spt_write_pfm(const char *path, const double *rgb, int w, int h)
{
    try { write_pfm(path, rgb, w, h); return 0; } catch (...) { return -1; }
}

int spt_write_raw64(const char *path, const double *data, int w, int h, int spp, const char *what)
{
    try { write_raw64(path, data, w, h, spp, what); return 0; } catch (...) { return -1; }
}

// Scene files.  spt_scene_text: built-in scene -> text (returns the length, or -1; buf may be NULL to query it).
int spt_scene_text(const char *name, char *buf, int cap)
{
    try {
        const std::string t = scene_to_text(scene_by_name(name), CameraSpec());
        if (buf && cap > 0) { std::strncpy(buf, t.c_str(), size_t(cap) - 1); buf[cap - 1] = 0; }
        return int(t.size());
    } catch (...) { return -1; }
}

static thread_local SceneFile g_parsed;
static thread_local std::string g_parse_error;

// Parse scene text; the result stays in a thread-local slot read by spt_parsed_*.  Returns the object count or -1.
int spt_parse_scene(const char *text, int *n_spheres, int *n_planes, int *has_camera)
{
    try {
        std::istringstream in(text);
        g_parsed = parse_scene(in);
        *n_spheres = int(g_parsed.table.spheres.size());
        *n_planes = int(g_parsed.table.planes.size());
        *has_camera = g_parsed.has_camera ? 1 : 0;
        return g_parsed.table.size();
    } catch (const std::exception &e) { g_parse_error = e.what(); return -1; }
}

const char *spt_parse_error(void) { return g_parse_error.c_str(); }

void spt_parsed_fill(pt_sphere *spheres, pt_plane *planes, int *order, pt_light *light, int w, int h, pt_camera *cam)
{
    const SceneTable &s = g_parsed.table;
    if (!s.spheres.empty()) std::memcpy(spheres, s.spheres.data(), s.spheres.size() * sizeof(pt_sphere));
    if (!s.planes.empty()) std::memcpy(planes, s.planes.data(), s.planes.size() * sizeof(pt_plane));
    std::memcpy(order, s.order.data(), s.order.size() * sizeof(int));
    *light = s.light;
    *cam = g_parsed.camera.make(w, h).pod();
}

}  // extern "C"
