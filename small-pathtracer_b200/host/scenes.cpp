// scenes.cpp — built-in scene tables, written the way the reference writes its scene literal
// (`Hitable *rect[NUMBER_OBJ] = { new Rectangle_xy(...), ... }`, src/smallpt.cpp:287-311).
#include "smallpt_b200.hpp"

#include <memory>

namespace smallpt_b200 {

namespace {
template <size_t N> SceneTable flatten(Hitable *(&table)[N])
{
    SceneTable s;
    s.add_all(table);
    for (size_t i = 0; i < N; i++) delete table[i];
    return s;
}

// Host Philox4x32-10 (Salmon et al., SC'11) for the synthetic scene generator; same constants as the
// device generator (csrc/pt_rng.cuh) so CPU- and GPU-side tools enumerate the same scene.
struct Philox {
    uint32_t key[2];
    uint32_t ctr;
    explicit Philox(uint64_t seed) : key{uint32_t(seed), uint32_t(seed >> 32)}, ctr(0) {}
    void block(uint32_t out[4])
    {
        uint32_t c0 = ctr++, c1 = 0, c2 = 0, c3 = 0, k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
            uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1);
            uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    uint32_t buf[4];
    int have = 0;
    double uniform()   // [0,1) with 32 bits
    {
        if (!have) { block(buf); have = 4; }
        return buf[4 - have--] * (1.0 / 4294967296.0);
    }
};
}  // namespace

SceneTable scene_A()   // HEAD, :287-311
{
    Hitable *rect[] = {
        new Rectangle_xy(1, 99, 0, 81.6, 0, Vec(), Vec(.75, .75, .75), DIFF),     // Front
        new Rectangle_xy(1, 99, 0, 81.6, 170, Vec(), Vec(.75, .75, .75), DIFF),   // Back
        new Rectangle_yz(0, 81.6, 0, 170, 1, Vec(), Vec(.25, .75, .25), DIFF),    // Left
        new Rectangle_yz(0, 81.6, 0, 170, 99, Vec(), Vec(.75, .25, .25), DIFF),   // Right
        new Rectangle_xz(1, 99, 0, 170, 0, Vec(), Vec(.75, .75, .75), DIFF),      // Bottom
        new Rectangle_xz(1, 99, 0, 170, 81.6, Vec(), Vec(.75, .75, .75), DIFF),   // Top
        new Rectangle_xz(32, 68, 63, 96, 81.5, Vec(12, 12, 12), Vec(), DIFF),     // Light
        new Rectangle_xy(12, 42, 0, 50, 32, Vec(), Vec(1, 1, 1), DIFF),           // Tall box
        new Rectangle_xy(12, 42, 0, 50, 62, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_yz(0, 50, 32, 62, 12, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_yz(0, 50, 32, 62, 42, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_xz(12, 42, 32, 62, 50, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_xy(63, 88, 0, 25, 63, Vec(), Vec(1, 1, 1), DIFF),           // Short box
        new Rectangle_xy(63, 88, 0, 25, 88, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_yz(0, 25, 63, 88, 63, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_yz(0, 25, 63, 88, 88, Vec(), Vec(1, 1, 1), DIFF),
        new Rectangle_xz(63, 88, 63, 88, 25, Vec(), Vec(1, 1, 1), DIFF),
    };
    SceneTable s = flatten(rect);
    s.set_reference_light(6);
    return s;
}

SceneTable scene_B()   // sphere era (src/a.exe static initialiser)
{
    Hitable *rect[] = {
        new Sphere(1e5, Vec(1e5 + 1, 40.8, 81.6), Vec(), Vec(.25, .75, .25), DIFF),    // Left
        new Sphere(1e5, Vec(1e5 + 1, 40.8, 81.6), Vec(), Vec(.25, .75, .25), DIFF),    // (duplicate in the binary)
        new Sphere(1e5, Vec(-1e5 + 99, 40.8, 81.6), Vec(), Vec(.75, .25, .25), DIFF),  // Right
        new Sphere(1e5, Vec(50, 40.8, 1e5), Vec(), Vec(.75, .75, .75), DIFF),          // Back
        new Sphere(1e5, Vec(50, 40.8, -1e5 + 170), Vec(), Vec(), DIFF),                // Front
        new Sphere(1e5, Vec(50, 1e5, 81.6), Vec(), Vec(.75, .75, .75), DIFF),          // Bottom
        new Sphere(1e5, Vec(50, -1e5 + 81.6, 81.6), Vec(), Vec(.75, .75, .75), DIFF),  // Top
        new Sphere(16.5, Vec(27, 16.5, 47), Vec(), Vec(1, 1, 1) * .999, DIFF),
        new Sphere(16.5, Vec(73, 16.5, 78), Vec(), Vec(.75, .75, .75), DIFF),
        new Sphere(600, Vec(50, 681.6 - .27, 81.6), Vec(12, 12, 12), Vec(), DIFF),     // Light
    };
    SceneTable s = flatten(rect);
    s.set_reference_light(9);
    return s;
}

SceneTable scene_G()   // the sphere-era box with Beason's mirror and glass spheres (ids 7, 8): the SPEC / REFR arms of :481-495
{
    SceneTable s = scene_B();
    s.set_material(7, Vec(1, 1, 1) * .999, SPEC);
    s.set_material(8, Vec(1, 1, 1) * .999, REFR);
    return s;
}

SceneTable scene_C()   // :288-294 + the two spheres of :297-298
{
    Hitable *rect[] = {
        new Rectangle_xy(1, 99, 0, 81.6, 0, Vec(), Vec(.75, .75, .75), DIFF),
        new Rectangle_xy(1, 99, 0, 81.6, 170, Vec(), Vec(.75, .75, .75), DIFF),
        new Rectangle_yz(0, 81.6, 0, 170, 1, Vec(), Vec(.25, .75, .25), DIFF),
        new Rectangle_yz(0, 81.6, 0, 170, 99, Vec(), Vec(.75, .25, .25), DIFF),
        new Rectangle_xz(1, 99, 0, 170, 0, Vec(), Vec(.75, .75, .75), DIFF),
        new Rectangle_xz(1, 99, 0, 170, 81.6, Vec(), Vec(.75, .75, .75), DIFF),
        new Rectangle_xz(32, 68, 63, 96, 81.5, Vec(12, 12, 12), Vec(), DIFF),
        new Sphere(16.5, Vec(27, 16.5, 47), Vec(), Vec(1, 1, 1) * .999, DIFF),
        new Sphere(16.5, Vec(73, 16.5, 78), Vec(), Vec(.75, .75, .75), DIFF),
    };
    SceneTable s = flatten(rect);
    s.set_reference_light(6);
    return s;
}

// Config C4: the six bounding rectangles + rectangular light of scene A, n_tilted bounded
// tilted planes (20x20) and n_spheres random spheres: centres uniform in [1,99]x[0,81.6]x[0,170], radii
// [1,5], albedo [0.2,0.9]^3, 10 % emissive e=(12,12,12)*U, 80 % DIFF / 10 % SPEC / 10 % REFR.
SceneTable scene_synthetic(int n_spheres, int n_tilted, uint64_t seed)
{
    SceneTable s;
    Philox rng(seed);
    s.add(Rectangle_xy(1, 99, 0, 81.6, 0, Vec(), Vec(.75, .75, .75), DIFF));
    s.add(Rectangle_xy(1, 99, 0, 81.6, 170, Vec(), Vec(.75, .75, .75), DIFF));
    s.add(Rectangle_yz(0, 81.6, 0, 170, 1, Vec(), Vec(.25, .75, .25), DIFF));
    s.add(Rectangle_yz(0, 81.6, 0, 170, 99, Vec(), Vec(.75, .25, .25), DIFF));
    s.add(Rectangle_xz(1, 99, 0, 170, 0, Vec(), Vec(.75, .75, .75), DIFF));
    s.add(Rectangle_xz(1, 99, 0, 170, 81.6, Vec(), Vec(.75, .75, .75), DIFF));
    s.add(Rectangle_xz(32, 68, 63, 96, 81.5, Vec(12, 12, 12), Vec(), DIFF));
    for (int i = 0; i < n_tilted; i++) {
        Vec p0(10 + 80 * rng.uniform(), 5 + 70 * rng.uniform(), 10 + 140 * rng.uniform());
        double z = 2 * rng.uniform() - 1, phi = 2 * M_PI * rng.uniform(), r = std::sqrt(1 - z * z);
        Vec n(r * std::cos(phi), r * std::sin(phi), z);
        Vec along = std::fabs(n.x) > .5 ? Vec(0, 1, 0) : Vec(1, 0, 0);
        Vec c(.2 + .7 * rng.uniform(), .2 + .7 * rng.uniform(), .2 + .7 * rng.uniform());
        s.add(Plane(p0, n, along, 10, 10, Vec(), c, DIFF));
    }
    for (int i = 0; i < n_spheres; i++) {
        Vec p(1 + 98 * rng.uniform(), 81.6 * rng.uniform(), 170 * rng.uniform());
        double rad = 1 + 4 * rng.uniform();
        Vec c(.2 + .7 * rng.uniform(), .2 + .7 * rng.uniform(), .2 + .7 * rng.uniform());
        double em = rng.uniform(), eu = rng.uniform(), mt = rng.uniform();
        Vec e = em < .1 ? Vec(12, 12, 12) * eu : Vec();
        Refl_t refl = mt < .8 ? DIFF : mt < .9 ? SPEC : REFR;
        s.add(Sphere(rad, p, e, c, refl));
    }
    s.set_reference_light(6);
    return s;
}

SceneTable scene_by_name(const std::string &name)
{
    if (name == "A" || name == "a") return scene_A();
    if (name == "B" || name == "b") return scene_B();
    if (name == "C" || name == "c") return scene_C();
    if (name == "G" || name == "g") return scene_G();
    if (name == "synthetic" || name == "S" || name == "s" || name == "D") return scene_synthetic();
    throw std::invalid_argument("unknown scene '" + name + "' (A, B, C, G, synthetic)");
}

}  // namespace smallpt_b200
