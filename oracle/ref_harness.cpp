// ref_harness.cpp — exports the UNMODIFIED reference functions for pinning the oracle.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it #includes the
// reference translation unit where it lies (-I/root/reference/src at build time, see
// oracle/Makefile), renames its main(), and wraps its own functions/classes behind a C ABI
// so tests can call Camera, Rectangle_*::intersect/normal, Sphere::intersect/normal,
// intersect(), hittingPoint(), random_scattering(), erand48, clamp and toInt exactly as the
// reference compiled them.  Output goes to oracle/_ref/librefharness.so (git-ignored).
//
// Sphere is abstract at HEAD (it does not override the RL virtuals add_key/add_value,
// src/smallpt.cpp:87-88,223-254); the subclass below only supplies those two bodies.
#define main smallpt_reference_main
#include "smallpt.cpp"
#undef main

namespace {
struct SphereConcrete : public Sphere {
    SphereConcrete(double rad_, Vec p_, Vec e_, Vec c_, Refl_t refl_) : Sphere(rad_, p_, e_, c_, refl_) {}
    std::array<float, 3> add_key(Vec &) const { return {0, 0, 0}; }
    std::array<float, 3> add_value(std::array<float, 3> &) const { return {0, 0, 0}; }
};
inline Ray mk(const double *r) { return Ray(Vec(r[0], r[1], r[2]), Vec(r[3], r[4], r[5])); }
}

extern "C" {

int ref_number_obj() { return NUMBER_OBJ; }

double ref_erand48(unsigned short *xi) { return erand48(xi); }

void ref_camera(const double *lookfrom, const double *lookat, const double *vup, float vfov, float aspect, double *out12)
{
    Camera cam(Vec(lookfrom[0], lookfrom[1], lookfrom[2]), Vec(lookat[0], lookat[1], lookat[2]),
               Vec(vup[0], vup[1], vup[2]), vfov, aspect);
    const Vec *m[4] = { &cam.origin, &cam.lower_left_corner, &cam.horizontal, &cam.vertical };
    for (int i = 0; i < 4; i++) { out12[3 * i] = m[i]->x; out12[3 * i + 1] = m[i]->y; out12[3 * i + 2] = m[i]->z; }
}

void ref_camera_get_ray(const double *lookfrom, const double *lookat, const double *vup, float vfov, float aspect,
                        float s, float t, double *out6)
{
    Camera cam(Vec(lookfrom[0], lookfrom[1], lookfrom[2]), Vec(lookat[0], lookat[1], lookat[2]),
               Vec(vup[0], vup[1], vup[2]), vfov, aspect);
    Ray r = cam.get_ray(s, t);
    out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}

// rect[i]->intersect(Ray) for the built-in scene table (src/smallpt.cpp:287-311)
void ref_object_intersect(int i, const double *rays_od, int n, double *t_out)
{
    for (int k = 0; k < n; k++) t_out[k] = rect[i]->intersect(mk(rays_od + 6 * k));
}

// rect[i]->normal(r, hit, x): out = nl(3), c(3), e(3), refl
void ref_object_normal(int i, const double *ray_od, const double *x, double *out10)
{
    Hit_records hit;
    Vec xx(x[0], x[1], x[2]);
    Vec nl = rect[i]->normal(mk(ray_od), hit, xx);
    out10[0] = nl.x; out10[1] = nl.y; out10[2] = nl.z;
    out10[3] = hit.c.x; out10[4] = hit.c.y; out10[5] = hit.c.z;
    out10[6] = hit.e.x; out10[7] = hit.e.y; out10[8] = hit.e.z;
    out10[9] = (double)hit.refl;
}

// intersect(Ray,t,id), src/smallpt.cpp:323-335, on the built-in scene; id starts at -1
void ref_scene_intersect(const double *rays_od, int n, double *t_out, int *id_out)
{
    for (int k = 0; k < n; k++) {
        double t; int id = -1;
        intersect(mk(rays_od + 6 * k), t, id);
        t_out[k] = t; id_out[k] = id;
    }
}

// hittingPoint(Ray,id), src/smallpt.cpp:371-377; id starts at 0 as in radiance() (:421)
void ref_hitting_point(const double *ray_od, double *x_out, int *id_out)
{
    int id = 0;
    Vec x = hittingPoint(mk(ray_od), id);
    x_out[0] = x.x; x_out[1] = x.y; x_out[2] = x.z; *id_out = id;
}

void ref_sphere_intersect(double rad, const double *p, const double *rays_od, int n, double *t_out)
{
    SphereConcrete s(rad, Vec(p[0], p[1], p[2]), Vec(), Vec(), DIFF);
    for (int k = 0; k < n; k++) t_out[k] = s.intersect(mk(rays_od + 6 * k));
}

void ref_sphere_normal(double rad, const double *p, const double *ray_od, const double *x, double *nl_out)
{
    SphereConcrete s(rad, Vec(p[0], p[1], p[2]), Vec(), Vec(), DIFF);
    Hit_records hit;
    Vec xx(x[0], x[1], x[2]);
    Vec nl = s.normal(mk(ray_od), hit, xx);
    nl_out[0] = nl.x; nl_out[1] = nl.y; nl_out[2] = nl.z;
}

// random_scattering(nl, Xi), src/smallpt.cpp:337-348 (cosine; the only sampler live at HEAD)
void ref_random_scattering(const double *nl, unsigned short *xi, double *d_out)
{
    Vec d = random_scattering(Vec(nl[0], nl[1], nl[2]), xi);
    d_out[0] = d.x; d_out[1] = d.y; d_out[2] = d.z;
}

double ref_clamp(double x) { return clamp(x); }
int ref_toInt(double x) { return toInt(x); }

}  // extern "C"
