// smallpt_b200.hpp — host-side surface of the B200 path tracer.
//
// Mirrors the reference's source-level "API" (one translation unit, src/smallpt.cpp) so that a
// scene literal and a main() written for the reference read the same here:
//   Vec (:24-62), Ray (:67-70), Refl_t (:72-74), Hitable (:82-90), Rectangle_xz/_xy/_yz
//   (:92-221), Sphere (:223-254), Camera (:256-285), clamp/toInt (:314-321), P3 writer (:548-551).
// New: Plane (the README's "tilted planes", README.md:19, absent from the source).
//
// These classes are DATA HOLDERS.  Nothing here intersects or shades: the hot path
// (src/smallpt.cpp:323-381,419-496,528-541) lives in CUDA behind include/ptb200.h, and a
// Renderer fails loudly when that library reports an error.  There is no CPU fallback.
#ifndef SMALLPT_B200_HPP
#define SMALLPT_B200_HPP

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "ptb200.h"

namespace smallpt_b200 {

struct Vec {                       // :24-62 — value semantics only; no arithmetic is needed on the host
    double x, y, z;                //          except what Camera's constructor uses
    Vec(double x_ = 0, double y_ = 0, double z_ = 0) : x(x_), y(y_), z(z_) {}
    Vec operator+(const Vec &b) const { return Vec(x + b.x, y + b.y, z + b.z); }
    Vec operator-(const Vec &b) const { return Vec(x - b.x, y - b.y, z - b.z); }
    Vec operator*(double b) const { return Vec(x * b, y * b, z * b); }
    Vec operator*(float b) const { return Vec(x * b, y * b, z * b); }
    Vec operator*(int b) const { return Vec(x * b, y * b, z * b); }
    Vec mult(const Vec &b) const { return Vec(x * b.x, y * b.y, z * b.z); }
    Vec &norm() { return *this = *this * (1 / std::sqrt(x * x + y * y + z * z)); }   // in place, like :50
    double dot(const Vec &b) const { return x * b.x + y * b.y + z * b.z; }
    Vec operator%(const Vec &b) const { return Vec(y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x); }
    pt_vec3 pod() const { return pt_vec3{x, y, z}; }
};

struct Ray {                       // :67-70
    Vec o, d;
    Ray(Vec o_, Vec d_) : o(o_), d(d_) {}
};

enum Refl_t { DIFF = PT_DIFF, SPEC = PT_SPEC, REFR = PT_REFR };   // :72-74

// :82-90.  The reference's virtuals intersect()/normal() are the hot path and are NOT host
// functions here; a Hitable only knows how to describe itself to the C ABI.
class Hitable {
public:
    virtual ~Hitable() {}
    virtual bool is_sphere() const = 0;
    virtual pt_sphere as_sphere() const { throw std::logic_error("not a sphere"); }
    virtual pt_plane as_plane() const { throw std::logic_error("not a plane"); }
};

namespace detail {
inline pt_plane rect(int kind, double a1, double a2, double b1, double b2, double k, Vec e, Vec c, Refl_t refl)
{
    pt_plane p{};
    p.kind = kind; p.refl = refl; p.a1 = a1; p.a2 = a2; p.b1 = b1; p.b2 = b2; p.k = k;
    p.e = e.pod(); p.c = c.pod();
    return p;
}
}  // namespace detail

class Rectangle_xz : public Hitable {   // :92-135
public:
    double x1, x2, z1, z2, y;
    Vec e, c;
    Refl_t refl;
    Rectangle_xz(double x1_, double x2_, double z1_, double z2_, double y_, Vec e_, Vec c_, Refl_t refl_)
        : x1(x1_), x2(x2_), z1(z1_), z2(z2_), y(y_), e(e_), c(c_), refl(refl_) {}
    bool is_sphere() const override { return false; }
    pt_plane as_plane() const override { return detail::rect(PT_PLANE_XZ, x1, x2, z1, z2, y, e, c, refl); }
};

class Rectangle_xy : public Hitable {   // :137-178
public:
    double x1, x2, y1, y2, z;
    Vec e, c;
    Refl_t refl;
    Rectangle_xy(double x1_, double x2_, double y1_, double y2_, double z_, Vec e_, Vec c_, Refl_t refl_)
        : x1(x1_), x2(x2_), y1(y1_), y2(y2_), z(z_), e(e_), c(c_), refl(refl_) {}
    bool is_sphere() const override { return false; }
    pt_plane as_plane() const override { return detail::rect(PT_PLANE_XY, x1, x2, y1, y2, z, e, c, refl); }
};

class Rectangle_yz : public Hitable {   // :180-221
public:
    double y1, y2, z1, z2, x;
    Vec e, c;
    Refl_t refl;
    Rectangle_yz(double y1_, double y2_, double z1_, double z2_, double x_, Vec e_, Vec c_, Refl_t refl_)
        : y1(y1_), y2(y2_), z1(z1_), z2(z2_), x(x_), e(e_), c(c_), refl(refl_) {}
    bool is_sphere() const override { return false; }
    pt_plane as_plane() const override { return detail::rect(PT_PLANE_YZ, y1, y2, z1, z2, x, e, c, refl); }
};

// Tilted bounded plane (README.md:19): centre p0, normal n (normalised here), an
// in-plane direction hint `along` (projected onto the plane -> s; t = n x s), half extents.
class Plane : public Hitable {
public:
    Vec p0, n, s, t;
    double hs, ht;
    Vec e, c;
    Refl_t refl;
    Plane(Vec p0_, Vec n_, Vec along, double hs_, double ht_, Vec e_, Vec c_, Refl_t refl_)
        : p0(p0_), n(n_), hs(hs_), ht(ht_), e(e_), c(c_), refl(refl_)
    {
        n.norm();
        s = along - n * along.dot(n);
        s.norm();
        t = n % s;
    }
    bool is_sphere() const override { return false; }
    pt_plane as_plane() const override
    {
        pt_plane p{};
        p.kind = PT_PLANE_TILTED; p.refl = refl;
        p.p0 = p0.pod(); p.n = n.pod(); p.s = s.pod(); p.t = t.pod(); p.hs = hs; p.ht = ht;
        p.e = e.pod(); p.c = c.pod();
        return p;
    }
};

class Sphere : public Hitable {         // :223-254
public:
    double rad;
    Vec p, e, c;
    Refl_t refl;
    Sphere(double rad_, Vec p_, Vec e_, Vec c_, Refl_t refl_) : rad(rad_), p(p_), e(e_), c(c_), refl(refl_) {}
    bool is_sphere() const override { return true; }
    pt_sphere as_sphere() const override
    {
        pt_sphere s{};
        s.rad = rad; s.p = p.pod(); s.e = e.pod(); s.c = c.pod(); s.refl = refl;
        return s;
    }
};

class Camera {                          // :256-285 — same float fov arithmetic as the reference
public:
    Camera(Vec lookfrom, Vec lookat, Vec vup, float vfov, float aspect)
    {
        float theta = vfov * M_PI / 180;
        float half_height = std::tan(theta / 2);
        float half_width = aspect * half_height;
        origin = lookfrom;
        Vec w = (lookat - lookfrom).norm();
        Vec u = (w % vup).norm();
        Vec v = (u % w);
        lower_left_corner = origin - u * half_width - v * half_height + w;
        horizontal = u * (half_width * 2);
        vertical = v * (half_height * 2);
    }
    Ray get_ray(float s, float t) const   // :276-279 (un-normalised; the device normalises, :536)
    {
        return Ray(origin, lower_left_corner + horizontal * s + vertical * t - origin);
    }
    pt_camera pod() const { return pt_camera{origin.pod(), lower_left_corner.pod(), horizontal.pod(), vertical.pod()}; }
    Vec origin, lower_left_corner, horizontal, vertical;
};

const Vec LOOKFROM = Vec(50, 40, 168);   // :65

inline double clamp(double x) { return x < 0 ? 0 : x > 1 ? 1 : x; }                    // :314-316
inline int toInt(double x) { return int(std::pow(clamp(x), 1 / 2.2) * 255 + .5); }     // :319-321

// P3 writer, :548-551 (byte-for-byte: header, "%d %d %d " per pixel, no newlines) — but checks fopen
// and closes the file.  rgb = per-pixel means (w*h*3), clamped here as at :538.
inline void write_ppm(const std::string &path, const double *rgb, int w, int h)
{
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fprintf(f, "P3\n%d %d\n%d\n", w, h, 255);
    for (int i = 0; i < w * h; i++)
        std::fprintf(f, "%d %d %d ", toInt(clamp(rgb[3 * i])), toInt(clamp(rgb[3 * i + 1])), toInt(clamp(rgb[3 * i + 2])));
    std::fclose(f);
}

// Flattened scene: the reference's `Hitable *rect[NUMBER_OBJ]` table (:287-311) in C-ABI form.
struct SceneTable {
    std::vector<pt_sphere> spheres;
    std::vector<pt_plane> planes;
    std::vector<int> order;
    pt_light light{};
    void add(const Hitable &h)
    {
        if (h.is_sphere()) { order.push_back(~int(spheres.size())); spheres.push_back(h.as_sphere()); }
        else { order.push_back(int(planes.size())); planes.push_back(h.as_plane()); }
    }
    template <size_t N> void add_all(Hitable *(&table)[N]) { for (size_t i = 0; i < N; i++) add(*table[i]); }
    int size() const { return int(order.size()); }
    // colour + material of object `id` (scene variants: the mirror / glass spheres of the sphere-era box)
    void set_material(int id, const Vec &c, Refl_t refl)
    {
        const int ref = order.at(size_t(id));
        if (ref < 0) { pt_sphere &o = spheres[size_t(~ref)]; o.c = c.pod(); o.refl = int(refl); }
        else { pt_plane &o = planes[size_t(ref)]; o.c = c.pod(); o.refl = int(refl); }
    }
    // the literals of :365-367,:467,:471
    void set_reference_light(int id = 6, double x0 = 32, double xw = 36, double z0 = 63, double zw = 36,
                             double y = 81.6, double area = 1296)
    {
        light.id = id; light.x0 = x0; light.xw = xw; light.z0 = z0; light.zw = zw; light.y = y; light.area = area;
    }
    pt_scene pod(const Camera &cam) const
    {
        pt_scene s{};
        s.spheres = spheres.data(); s.n_spheres = int(spheres.size());
        s.planes = planes.data(); s.n_planes = int(planes.size());
        s.order = order.data(); s.camera = cam.pod(); s.light = light;
        return s;
    }
};

// Built-in scenes -------------------------------------------------------------------------------
SceneTable scene_A();     // HEAD: 17 rectangles (:287-311)
SceneTable scene_B();     // sphere era: 10 spheres (recovered from src/a.exe)
SceneTable scene_G();     // scene B with a mirror (id 7) and a glass (id 8) sphere: SPEC / REFR (:481-495)
SceneTable scene_C();     // 7 rectangles + the two commented spheres of :297-298 (image_light_test.ppm)
SceneTable scene_synthetic(int n_spheres = 256, int n_tilted = 8, uint64_t seed = 12345);   // config C4
SceneTable scene_by_name(const std::string &name);

// Thin RAII wrapper over the C ABI.  n_gpus > 1: one context per device (0 .. n_gpus-1), row tiles rendered concurrently
// and assembled in device 0's image over NVLink peer memory (pt_render_multi).
class Renderer {
public:
    Renderer(const SceneTable &scene, const Camera &cam, int device = -1, int n_gpus = 1) : ctx_(nullptr)
    {
        pt_scene s = scene.pod(cam);
        if (n_gpus <= 1) { check(pt_scene_upload(&ctx_, &s, device), "pt_scene_upload"); all_.push_back(ctx_); return; }
        for (int d = 0; d < n_gpus; d++) {
            pt_ctx *c = nullptr;
            int rc = pt_scene_upload(&c, &s, d);
            if (rc != PT_OK) { std::string msg = pt_last_error(nullptr); for (pt_ctx *x : all_) pt_destroy(x); all_.clear(); ctx_ = nullptr;
                               throw std::runtime_error("pt_scene_upload on device " + std::to_string(d) + " failed: " + msg); }
            all_.push_back(c);
        }
        ctx_ = all_[0];
    }
    ~Renderer() { for (pt_ctx *c : all_) pt_destroy(c); }
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;
    void render(const pt_render_params &p)
    {
        if (all_.size() > 1) check(pt_render_multi(all_.data(), int(all_.size()), &p), "pt_render_multi");
        else check(pt_render(ctx_, &p), "pt_render");
    }
    std::vector<double> readback(int w, int h, pt_stats *stats = nullptr, std::vector<double> *sumsq = nullptr)
    {
        std::vector<double> rgb(size_t(w) * h * 3);
        if (sumsq) sumsq->resize(rgb.size());
        check(pt_readback(ctx_, rgb.data(), sumsq ? sumsq->data() : nullptr, stats), "pt_readback");
        return rgb;
    }
    // checkpoint / resume of the accumulators (per-pixel SUMS and the number of samples in them)
    std::vector<double> accum_download(int w, int h, int *spp_done)
    {
        std::vector<double> sums(size_t(w) * h * 3);
        check(pt_accum_download(ctx_, sums.data(), nullptr, spp_done), "pt_accum_download");
        return sums;
    }
    void accum_upload(int w, int h, const std::vector<double> &sums, int spp_done)
    {
        check(pt_accum_upload(ctx_, w, h, sums.data(), nullptr, spp_done), "pt_accum_upload");
    }
    pt_ctx *ctx() { return ctx_; }
private:
    void check(int rc, const char *what)
    {
        if (rc != PT_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + pt_last_error(ctx_));
    }
    pt_ctx *ctx_;
    std::vector<pt_ctx *> all_;
};

}  // namespace smallpt_b200
#endif
