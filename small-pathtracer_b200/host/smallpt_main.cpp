// smallpt_main.cpp — the `smallpt` executable: the reference's main() (src/smallpt.cpp:502-557) on the
// B200 path.  Usage:  smallpt [spp] [--mode nee|cos|uni|nee-cone] [--scene A|B|C|synthetic | --scene-file f.scene]
//                             [--size WxH] [--validate] [--det-sincos] [--seed N] [--out file.ppm]
//                             [--ppm6 file] [--pfm file] [--raw64 file] [--variance file] [--dump-scene file]
//                             [--chunk N] [--checkpoint file] [--resume file]      progressive accumulation
//                             [--gpus N]                                           row tiles over N GPUs of this node
//                             [--robust-eps]            NOT the reference: rectangles require t > 1e-4 (no self-hit leaks)
//                             [--stats]                 print why paths ended and the live-path histogram (slower render)
// `spp` is argv[1] as the north star asks (the reference hard-codes samps = 16 at :508).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "scene_io.hpp"

using namespace smallpt_b200;

int main(int argc, char *argv[])
{
    int w = 512, h = 512;          // :507
    int samps = 16;                // :508
    std::string mode = "nee", scene_name = "A", out = "image.ppm", scene_file, ppm6, pfm, raw64, variance, dump_scene, checkpoint, resume;
    int chunk = 0, gpus = 1;
    bool validate = false, det = false, robust = false, want_stats = false;
    uint64_t seed = 0;
    int argi = 1;
    if (argc > 1 && argv[1][0] != '-') { samps = std::atoi(argv[1]); argi = 2; }
    for (; argi < argc; argi++) {
        std::string a = argv[argi];
        auto need = [&](const char *what) -> const char * {
            if (argi + 1 >= argc) { std::cerr << what << " needs a value\n"; std::exit(2); }
            return argv[++argi];
        };
        if (a == "--mode") mode = need("--mode");
        else if (a == "--scene") scene_name = need("--scene");
        else if (a == "--size") { if (std::sscanf(need("--size"), "%dx%d", &w, &h) != 2) { std::cerr << "--size WxH\n"; return 2; } }
        else if (a == "--validate") validate = true;
        else if (a == "--det-sincos") det = true;
        else if (a == "--robust-eps") robust = true;
        else if (a == "--stats") want_stats = true;
        else if (a == "--seed") seed = std::strtoull(need("--seed"), nullptr, 10);
        else if (a == "--out") out = need("--out");
        else if (a == "--scene-file") scene_file = need("--scene-file");
        else if (a == "--ppm6") ppm6 = need("--ppm6");
        else if (a == "--pfm") pfm = need("--pfm");
        else if (a == "--raw64") raw64 = need("--raw64");
        else if (a == "--variance") variance = need("--variance");
        else if (a == "--dump-scene") dump_scene = need("--dump-scene");
        else if (a == "--chunk") chunk = std::atoi(need("--chunk"));
        else if (a == "--gpus") gpus = std::atoi(need("--gpus"));
        else if (a == "--checkpoint") checkpoint = need("--checkpoint");
        else if (a == "--resume") resume = need("--resume");
        else { std::cerr << "unknown argument " << a << "\n"; return 2; }
    }
    if (samps <= 0 || w <= 0 || h <= 0) { std::cerr << "spp and size must be positive\n"; return 2; }
    pt_render_params p{};
    p.width = w; p.height = h; p.spp = samps; p.seed = seed;
    p.mode = mode == "cos" ? PT_MODE_COS : mode == "uni" ? PT_MODE_UNI : mode == "nee-cone" ? PT_MODE_NEE_CONE_SPHERE : PT_MODE_NEE_REF_RECT;
    p.engine = validate ? PT_ENGINE_FP64_ERAND48 : PT_ENGINE_FP32_PHILOX;
    p.sincos = det ? PT_SINCOS_DET : PT_SINCOS_LIBM;
    p.world = 1;
    p.collect_stats = (variance.empty() && !want_stats) ? 0 : 1;
    p.robust_eps = robust ? 1 : 0;
    try {
        auto t1 = std::chrono::high_resolution_clock::now();
        SceneTable scene;
        CameraSpec cam_spec;                                                           // defaults = the literals of :65,:521
        if (!scene_file.empty()) {
            SceneFile sf = load_scene_file(scene_file);
            scene = sf.table;
            if (sf.has_camera) cam_spec = sf.camera;
        } else {
            scene = scene_by_name(scene_name);
        }
        if (!dump_scene.empty()) {
            std::ofstream o(dump_scene);
            if (!o) throw std::runtime_error("cannot open " + dump_scene);
            o << scene_to_text(scene, cam_spec);
        }
        Camera cam = cam_spec.make(w, h);                                              // :521
        if (gpus > 1 && (validate || chunk > 0 || !resume.empty() || !checkpoint.empty() || p.collect_stats)) {
            std::cerr << "--gpus N combines with the plain FP32 render only\n";
            return 2;
        }
        Renderer r(scene, cam, -1, gpus);
        // Progressive accumulation: the sample range [0, samps) in chunks; every chunk continues the same image (samples are
        // Philox streams keyed by their index), a checkpoint holds the per-pixel sums and the number of samples in them.
        if ((chunk > 0 || !resume.empty() || !checkpoint.empty()) && (validate || p.collect_stats)) {
            std::cerr << "--chunk/--checkpoint/--resume apply to the FP32 engine without --variance / --stats\n";
            return 2;
        }
        int done = 0;
        double render_ms = 0;
        uint64_t paths = 0, rays_c = 0, rays_s = 0, rays_sh = 0;
        if (!resume.empty()) {
            int cw = 0, ch = 0, cspp = 0;
            std::string what;
            std::vector<double> sums = read_raw64(resume, cw, ch, cspp, what);
            if (cw != w || ch != h || what != "sum" || cspp <= 0) throw std::runtime_error(resume + ": not a checkpoint of a " + std::to_string(w) + "x" + std::to_string(h) + " image");
            r.accum_upload(w, h, sums, cspp);
            done = cspp;
        }
        pt_stats st{};
        if (done >= samps && done > 0) { p.spp = 0; }
        while (done < samps) {
            const int n = chunk > 0 ? std::min(chunk, samps - done) : samps - done;
            p.spp = n; p.sample_offset = done; p.accumulate = done > 0 ? 1 : 0;
            r.render(p);
            done += n;
            pt_stats cs{};
            pt_readback(r.ctx(), nullptr, nullptr, &cs);
            render_ms += cs.render_ms; paths += cs.paths; rays_c += cs.rays_camera; rays_s += cs.rays_scatter; rays_sh += cs.rays_shadow;
            if (!checkpoint.empty()) {
                int spp_done = 0;
                std::vector<double> sums = r.accum_download(w, h, &spp_done);
                write_raw64(checkpoint, sums.data(), w, h, spp_done, "sum");
            }
        }
        std::vector<double> sumsq;
        std::vector<double> c = r.readback(w, h, &st, p.collect_stats ? &sumsq : nullptr);
        if (render_ms > 0) { st.render_ms = render_ms; st.paths = paths; st.rays_camera = rays_c; st.rays_scatter = rays_s; st.rays_shadow = rays_sh; }
        write_ppm(out, c.data(), w, h);                                                // :548-551
        if (!ppm6.empty()) write_ppm_binary(ppm6, c.data(), w, h);
        if (!pfm.empty()) write_pfm(pfm, c.data(), w, h);
        if (!raw64.empty()) write_raw64(raw64, c.data(), w, h, samps, "mean");
        if (!variance.empty()) { std::vector<double> var = variance_from(c, sumsq, samps); write_raw64(variance, var.data(), w, h, samps, "variance"); }
        auto t2 = std::chrono::high_resolution_clock::now();
        double rays = double(st.rays_camera + st.rays_scatter + st.rays_shadow);
        std::cout << "PATHS: " << st.paths << "  RAYS: " << (uint64_t)rays << "  MAX DEPTH: " << st.max_depth_seen << std::endl;
        std::cout << "RENDER ms: " << st.render_ms << "  Mpaths/s: " << st.paths / st.render_ms * 1e-3
                  << "  Mrays/s: " << rays / st.render_ms * 1e-3 << std::endl;
        if (want_stats && !validate) {
            std::cout << "ENDED BY: roulette " << st.term_roulette << "  emitter " << st.term_emitter << "  light sample " << st.term_light_sample
                      << "  max depth " << st.truncated << "  | MISSES: " << st.miss_events << "  DROPPED CONTRIBUTIONS: " << st.dropped_contributions << std::endl;
            std::cout << "LIVE PATHS BY DEPTH:";
            for (int k = 0; k < 64 && st.live_at_depth[k]; k++) std::cout << " " << st.live_at_depth[k];
            std::cout << std::endl;
        }
        std::cout << " DURATION : " << std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count();   // :554-556
        std::cout << std::endl;
    } catch (const std::exception &e) {
        std::cerr << "smallpt: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
