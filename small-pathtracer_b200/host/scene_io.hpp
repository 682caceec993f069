// scene_io.hpp — data formats on either side of the hot path:
//   * image outputs next to the reference's P3 writer (src/smallpt.cpp:548-551): binary PPM (P6, same clamp + gamma,
//     :314-321), PFM (float32 linear radiance, unclamped) and a raw float64 dump with a one-line text header (used for
//     the per-pixel mean and for the sum-of-squares / variance buffer);
//   * a line-oriented text scene format, so that scenes are data instead of a source literal (:287-311).
//
// Scene file grammar (one statement per line, `#` starts a comment, numbers are C doubles, refl = DIFF | SPEC | REFR):
//   camera   lx ly lz   ax ay az   ux uy uz   vfov          lookfrom, lookat, vup, vertical fov in degrees (:262,521)
//   sphere   rad   px py pz   ex ey ez   cx cy cz   refl                        Sphere(rad,p,e,c,refl)            (:228)
//   rect_xz  x1 x2 z1 z2 y    ex ey ez   cx cy cz   refl                        Rectangle_xz(x1,x2,z1,z2,y,...)   (:97)
//   rect_xy  x1 x2 y1 y2 z    ex ey ez   cx cy cz   refl                        Rectangle_xy(x1,x2,y1,y2,z,...)   (:142)
//   rect_yz  y1 y2 z1 z2 x    ex ey ez   cx cy cz   refl                        Rectangle_yz(y1,y2,z1,z2,x,...)   (:185)
//   plane    px py pz  nx ny nz  ax ay az  hs ht   ex ey ez   cx cy cz   refl   tilted bounded plane
//   light    id  x0 xw  z0 zw  y  area               the literals of :365-367,:467,:471 for NEE_REF_RECT
// Objects get ids in file order — the index in the reference's rect[] table, which decides ties in intersect() (:328).
#ifndef SMALLPT_B200_SCENE_IO_HPP
#define SMALLPT_B200_SCENE_IO_HPP

#include <cstdint>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "smallpt_b200.hpp"

namespace smallpt_b200 {

// P6: same header fields and the same clamp (:538) + toInt (:319-321) as the P3 writer, bytes instead of text.
inline void write_ppm_binary(const std::string &path, const double *rgb, int w, int h)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fprintf(f, "P6\n%d %d\n%d\n", w, h, 255);
    std::vector<unsigned char> row(size_t(w) * 3);
    for (int y = 0; y < h; y++) {
        for (int i = 0; i < w * 3; i++) row[i] = (unsigned char)toInt(clamp(rgb[size_t(y) * w * 3 + i]));
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
}

// PFM ("PF", little-endian float32, rows bottom to top): linear, UNclamped radiance.
inline void write_pfm(const std::string &path, const double *rgb, int w, int h)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    std::vector<float> row(size_t(w) * 3);
    for (int y = h - 1; y >= 0; y--) {
        for (int i = 0; i < w * 3; i++) row[i] = (float)rgb[size_t(y) * w * 3 + i];
        std::fwrite(row.data(), sizeof(float), row.size(), f);
    }
    std::fclose(f);
}

// Raw float64 dump: "PTB200F64 <w> <h> <channels> <spp> <what>\n" then w*h*channels little-endian doubles, row 0 = top.
inline void write_raw64(const std::string &path, const double *data, int w, int h, int spp, const char *what)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fprintf(f, "PTB200F64 %d %d 3 %d %s\n", w, h, spp, what);
    std::fwrite(data, sizeof(double), size_t(w) * h * 3, f);
    std::fclose(f);
}

inline std::vector<double> read_raw64(const std::string &path, int &w, int &h, int &spp, std::string &what)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::string line, magic;
    std::getline(in, line);
    std::istringstream hs(line);
    int ch = 0;
    hs >> magic >> w >> h >> ch >> spp >> what;
    if (magic != "PTB200F64" || ch != 3 || w <= 0 || h <= 0) throw std::runtime_error(path + ": not a PTB200F64 dump");
    std::vector<double> v(size_t(w) * h * 3);
    in.read(reinterpret_cast<char *>(v.data()), std::streamsize(v.size() * sizeof(double)));
    if (!in) throw std::runtime_error(path + ": truncated");
    return v;
}

// per-pixel sample variance from the mean and the sum of squares: max(0, sumsq/n - mean^2)
inline std::vector<double> variance_from(const std::vector<double> &mean, const std::vector<double> &sumsq, int spp)
{
    std::vector<double> v(mean.size());
    for (size_t i = 0; i < v.size(); i++) { double x = sumsq[i] / spp - mean[i] * mean[i]; v[i] = x > 0 ? x : 0; }
    return v;
}

struct CameraSpec {                    // the arguments of Camera's constructor (:262), aspect comes from the image size
    Vec lookfrom = LOOKFROM, lookat = Vec(50, 40, 5), vup = Vec(0, 1, 0);
    float vfov = 65;
    Camera make(int w, int h) const { return Camera(lookfrom, lookat, vup, vfov, float(w) / float(h)); }
};

struct SceneFile {
    SceneTable table;
    CameraSpec camera;
    bool has_camera = false, has_light = false;
};

namespace detail {
inline Refl_t parse_refl(const std::string &s, int line)
{
    if (s == "DIFF" || s == "0") return DIFF;
    if (s == "SPEC" || s == "1") return SPEC;
    if (s == "REFR" || s == "2") return REFR;
    throw std::runtime_error("scene line " + std::to_string(line) + ": material must be DIFF, SPEC or REFR, got '" + s + "'");
}
inline const char *refl_name(int r) { return r == SPEC ? "SPEC" : r == REFR ? "REFR" : "DIFF"; }
}  // namespace detail

inline SceneFile parse_scene(std::istream &in)
{
    SceneFile sf;
    std::string raw;
    int ln = 0;
    while (std::getline(in, raw)) {
        ln++;
        const size_t hash = raw.find('#');
        if (hash != std::string::npos) raw.erase(hash);
        std::istringstream ls(raw);
        std::string kw;
        if (!(ls >> kw)) continue;
        auto num = [&](int n, double *out) {
            for (int i = 0; i < n; i++)
                if (!(ls >> out[i])) throw std::runtime_error("scene line " + std::to_string(ln) + ": '" + kw + "' needs more numbers");
        };
        auto refl = [&]() { std::string r; if (!(ls >> r)) throw std::runtime_error("scene line " + std::to_string(ln) + ": material missing"); return detail::parse_refl(r, ln); };
        double v[20];
        if (kw == "camera") {
            num(10, v);
            sf.camera.lookfrom = Vec(v[0], v[1], v[2]); sf.camera.lookat = Vec(v[3], v[4], v[5]); sf.camera.vup = Vec(v[6], v[7], v[8]);
            sf.camera.vfov = float(v[9]); sf.has_camera = true;
        } else if (kw == "sphere") {
            num(10, v); Refl_t r = refl();
            Sphere s(v[0], Vec(v[1], v[2], v[3]), Vec(v[4], v[5], v[6]), Vec(v[7], v[8], v[9]), r);
            sf.table.add(s);
        } else if (kw == "rect_xz" || kw == "rect_xy" || kw == "rect_yz") {
            num(11, v); Refl_t r = refl();
            const Vec e(v[5], v[6], v[7]), c(v[8], v[9], v[10]);
            if (kw == "rect_xz") { Rectangle_xz q(v[0], v[1], v[2], v[3], v[4], e, c, r); sf.table.add(q); }
            else if (kw == "rect_xy") { Rectangle_xy q(v[0], v[1], v[2], v[3], v[4], e, c, r); sf.table.add(q); }
            else { Rectangle_yz q(v[0], v[1], v[2], v[3], v[4], e, c, r); sf.table.add(q); }
        } else if (kw == "plane") {
            num(17, v); Refl_t r = refl();
            Plane q(Vec(v[0], v[1], v[2]), Vec(v[3], v[4], v[5]), Vec(v[6], v[7], v[8]), v[9], v[10], Vec(v[11], v[12], v[13]), Vec(v[14], v[15], v[16]), r);
            sf.table.add(q);
        } else if (kw == "light") {
            num(7, v);
            sf.table.set_reference_light(int(v[0]), v[1], v[2], v[3], v[4], v[5], v[6]);
            sf.has_light = true;
        } else {
            throw std::runtime_error("scene line " + std::to_string(ln) + ": unknown statement '" + kw + "'");
        }
    }
    if (sf.table.size() == 0) throw std::runtime_error("scene file holds no objects");
    if (!sf.has_light) sf.table.light.id = -1;        // PT_MODE_NEE_REF_RECT then fails loudly in pt_render
    else if (sf.table.light.id < 0 || sf.table.light.id >= sf.table.size()) throw std::runtime_error("light id names no object");
    return sf;
}

inline SceneFile load_scene_file(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    return parse_scene(in);
}

// The inverse: any SceneTable as text (17 significant digits: doubles survive the round trip bit for bit).
inline std::string scene_to_text(const SceneTable &t, const CameraSpec &cam)
{
    std::ostringstream o;
    o << std::setprecision(17);
    auto v3 = [&](const pt_vec3 &v) { o << ' ' << v.x << ' ' << v.y << ' ' << v.z; };
    o << "# small-pathtracer scene (objects in id order)\n";
    o << "camera " << cam.lookfrom.x << ' ' << cam.lookfrom.y << ' ' << cam.lookfrom.z << ' ' << cam.lookat.x << ' ' << cam.lookat.y << ' '
      << cam.lookat.z << ' ' << cam.vup.x << ' ' << cam.vup.y << ' ' << cam.vup.z << ' ' << double(cam.vfov) << "\n";
    for (int i = 0; i < t.size(); i++) {
        const int ord = t.order[i];
        if (ord < 0) {
            const pt_sphere &s = t.spheres[~ord];
            o << "sphere " << s.rad; v3(s.p); v3(s.e); v3(s.c); o << ' ' << detail::refl_name(s.refl) << "\n";
        } else {
            const pt_plane &p = t.planes[ord];
            if (p.kind == PT_PLANE_TILTED) {
                o << "plane"; v3(p.p0); v3(p.n); v3(p.s); o << ' ' << p.hs << ' ' << p.ht; v3(p.e); v3(p.c);
            } else {
                o << (p.kind == PT_PLANE_XZ ? "rect_xz " : p.kind == PT_PLANE_XY ? "rect_xy " : "rect_yz ")
                  << p.a1 << ' ' << p.a2 << ' ' << p.b1 << ' ' << p.b2 << ' ' << p.k; v3(p.e); v3(p.c);
            }
            o << ' ' << detail::refl_name(p.refl) << "\n";
        }
    }
    if (t.light.id >= 0)
        o << "light " << t.light.id << ' ' << t.light.x0 << ' ' << t.light.xw << ' ' << t.light.z0 << ' ' << t.light.zw << ' ' << t.light.y << ' '
          << t.light.area << "\n";
    return o.str();
}

}  // namespace smallpt_b200
#endif
