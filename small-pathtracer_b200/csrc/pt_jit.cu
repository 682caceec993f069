// pt_jit.cu — scene-specialised builds of k_bounce, compiled at run time with NVRTC for sm_100a.
//
// The reference hard-codes its scene as a source literal (src/smallpt.cpp:287-311): the compiler sees every plane
// constant.  The generic k_bounce reads them from __constant__ memory (two uniform loads per rectangle, a jump-table
// dispatch per axis class, run-time loop bounds).  Here the uploaded scene is turned back into source: a small header
// of constexpr tables (rectangle slots as hex-float literals, primitive counts, light constants) is put in front of
// the SAME kernel source (pt_kernel.cuh, embedded at build time) and compiled with PT_JIT defined, so the constants
// become immediates, equal sub-expressions (o.x - a1 for slots that share a1) are computed once, and empty primitive
// classes vanish.  One module per (scene constants, mode, statistics flag), cached for the life of the process.
//
// NVRTC is loaded with dlopen: the library has no link-time dependency on it and falls back to the ahead-of-time
// generic kernel (still a GPU kernel, never a CPU path) when it is missing or the compilation fails.
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <atomic>
#include <mutex>
#include <thread>
#include <set>

#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include "pt_internal.h"
#include "pt_kernel_src.h"

namespace {

struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*Version)(int *, int *) = nullptr;
    int major = 0, minor = 0;
    bool ok = false;
    std::string why;
};

Nvrtc &nvrtc()
{
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = std::getenv("PTB200_NVRTC");
        const char *names[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) { n.why = "libnvrtc not found"; return; }
#define PT_SYM(f) *(void **)(&n.f) = dlsym(n.h, "nvrtc" #f); if (!n.f) { n.why = "nvrtc" #f " missing"; return; }
        PT_SYM(CreateProgram) PT_SYM(CompileProgram) PT_SYM(GetCUBINSize) PT_SYM(GetCUBIN)
        PT_SYM(GetProgramLogSize) PT_SYM(GetProgramLog) PT_SYM(DestroyProgram) PT_SYM(Version)
        n.Version(&n.major, &n.minor);
#undef PT_SYM
        n.ok = true;
    });
    return n;
}

void put_float(std::string &s, float v)
{
    char b[64];
    if (v != v) std::snprintf(b, sizeof b, "__int_as_float(0x7fc00000)");
    else if (v > 3.4e38f) std::snprintf(b, sizeof b, "__int_as_float(0x7f800000)");
    else if (v < -3.4e38f) std::snprintf(b, sizeof b, "__int_as_float(0xff800000)");
    else std::snprintf(b, sizeof b, "%af", (double)v);       // hex float: exact
    s += b;
}

unsigned int float_bits(float v) { unsigned int u; std::memcpy(&u, &v, 4); return u; }

std::mutex g_mu;
std::map<std::string, PtJitKernel *> g_cache;       // key = specialisation header (contains mode and stats flag)

// A build running on its own host thread (NVRTC or the disk cache: no CUDA call — the module is loaded by the
// thread that asks for the kernel).  Small renders use it: they go on with the generic kernel meanwhile.
struct PendingBuild {
    std::thread th;
    std::atomic<bool> done{false};
    int rc = PT_ERR_STATE;
    std::vector<char> cubin;
    std::string log;
    double secs = 0;
    bool from_disk = false;
};
struct PendingSet {
    std::map<std::string, PendingBuild *> m;
    ~PendingSet() { for (auto &kv : m) { if (kv.second->th.joinable()) kv.second->th.join(); delete kv.second; } }   // process exit
};
PendingSet g_pending;
std::map<std::string, double> g_spent_ms;           // GPU time small renders of a specialisation spent in the generic kernel
bool g_exit_hook = false;

// Process exit with a build in flight: wait for it BEFORE libnvrtc's own static destructors run (handlers run in
// reverse order of registration, and this one is registered after libnvrtc was loaded).
void join_pending_at_exit()
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_pending.m) if (kv.second->th.joinable()) kv.second->th.join();
}

double background_after_ms()
{
    if (const char *e = std::getenv("PTB200_JIT_BG_MS")) return std::atof(e);
    return PT_JIT_BACKGROUND_AFTER_MS;
}

}  // namespace

// Threads per block of a scene's specialised module (the host launches it with the same number).  One block of 1024 threads
// per SM (also the ahead-of-time build's: C2 9.02 -> 8.20 ms there) measured 1.7-5 % faster than four of 256 on every config (C5 263.1 -> 255.4 ms, C2 4.89 -> 4.81, C3 1.66 -> 1.60,
// an eighth of C5 35.8 -> 34.85; 128 or 64 threads: 0.7 % slower); the scenes whose warps iterate in lockstep over a long
// immediate sphere table (PT_LOCKSTEP below) do best with 512 (C4 561.6 -> 530.0 ms; 1024: 544).
int pt_jit_block(const SceneF32 &S)
{
    if (const char *e = std::getenv("PTB200_JIT_BLOCK")) return std::max(32, std::atoi(e));       // tuning aid, with -DPT_BLOCK=... in PTB200_JIT_OPTS
    if (S.n_sph4 >= 64 && S.n_sph4 <= PT_JIT_SPH_IMM_MAX) return 512;
    return 1024;
}

// The specialisation header of a scene: everything k_bounce reads through PT_SC / PT_J_SLOT.
std::string pt_jit_spec(const SceneF32 &S, int mode, bool stats, bool with_intersect, int render_flags)
{
    std::string h;
    char b[256];
    std::snprintf(b, sizeof b, "#define PT_JIT 1\n#define PT_J_MODE %d\n#define PT_J_STATS %d\n", mode, stats ? 1 : 0);
    h += b;
    if (render_flags >= 0) {
        // the render's layout (Fp32Plan::flags): what the regeneration step would otherwise test at run time in every iteration
        // (C2 +5 %, a one-eighth share of C5 +3.7 %); a module only ever runs renders with the flags it was built for
        std::snprintf(b, sizeof b, "#define PT_BAKE_WORLD1 %d\n#define PT_BAKE_RUNS %d\n", (render_flags & PT_RF_WORLD1) ? 1 : 0, (render_flags & PT_RF_RUNS) ? 1 : 0);
        h += b;
        if (render_flags & PT_RF_WRAP_ONCE) h += "#define PT_BAKE_WRAP_ONCE 1\n";
        if (render_flags & PT_RF_MAGIC) h += "#define PT_BAKE_MAGIC 1\n";
        if (render_flags & PT_RF_ONE_BLOCK) h += "#define PT_NO_ROW_BLOCKS 1\n";
    }
    if (with_intersect) h += "#define PT_J_WITH_INTERSECT 1\n";    // also build k_intersect_jit (pt_debug_intersect)
    std::snprintf(b, sizeof b, "constexpr int PT_J_NSLOT[3] = {%d, %d, %d};\n#define PT_J_NOVF %d\n", S.n_slot[0], S.n_slot[1], S.n_slot[2], S.ovf_begin[3]);
    h += b;
    h += "constexpr float PT_J_SLOT[3][16][3] = {\n";         // k, a1, b1
    for (int a = 0; a < 3; a++) {
        h += " {";
        for (int k = 0; k < PT_RECT_SLOTS; k++) {
            h += "{"; put_float(h, S.slot_a[a][k].x); h += ","; put_float(h, S.slot_a[a][k].y); h += ","; put_float(h, S.slot_a[a][k].w); h += "},";
        }
        h += "},\n";
    }
    h += "};\nconstexpr unsigned int PT_J_SLOTB[3][16][2] = {\n";   // bits(a2 - a1), bits(b2 - b1)
    for (int a = 0; a < 3; a++) {
        h += " {";
        for (int k = 0; k < PT_RECT_SLOTS; k++) {
            std::snprintf(b, sizeof b, "{0x%08xu,0x%08xu},", float_bits(S.slot_a[a][k].z), float_bits(S.slot_b2[a][k]));
            h += b;
        }
        h += "},\n";
    }
    h += "};\n";
    auto def_i = [&](const char *n, int v) { std::snprintf(b, sizeof b, "#define PT_J_%s %d\n", n, v); h += b; };
    auto def_f = [&](const char *n, float v) { h += "#define PT_J_"; h += n; h += " ("; put_float(h, v); h += ")\n"; };
    def_i("n_sph4", S.n_sph4); def_i("n_huge", S.n_huge); def_i("n_tilt", S.n_tilt); def_i("n_lights", S.n_lights);
    def_i("has_sphere", (S.n_sph > 0 || S.n_huge > 0) ? 1 : 0); def_i("has_small_sphere", S.n_sph > 0 ? 1 : 0); def_i("has_tilt", S.n_tilt > 0 ? 1 : 0);
    def_i("has_rect", (S.n_slot[0] + S.n_slot[1] + S.n_slot[2] + S.ovf_begin[3]) > 0 ? 1 : 0);
    def_i("has_spec", (S.refl_mask >> PT_SPEC) & 1); def_i("has_refr", (S.refl_mask >> PT_REFR) & 1); def_i("has_diff", (S.refl_mask >> PT_DIFF) & 1);
    def_i("code_sph0", S.code_sph0); def_i("code_huge0", S.code_huge0); def_i("code_tilt0", S.code_tilt0);
    def_i("code_obj0", S.code_obj0); def_i("light_code", S.light_code);
    def_f("lx0", S.lx0); def_f("lxw", S.lxw); def_f("lz0", S.lz0); def_f("lzw", S.lzw); def_f("ly", S.ly); def_f("larea", S.larea);
    def_f("sph_kM2", S.sph_kM2);
    def_f("light_ex", S.light_e[0]); def_f("light_ey", S.light_e[1]); def_f("light_ez", S.light_e[2]);
    def_f("light_cx", S.light_c[0]); def_f("light_cy", S.light_c[1]); def_f("light_cz", S.light_c[2]);
    // Shadow rays toward the reference's rectangular light only ask whether the light is the closest hit: when no other slot
    // rectangle comes near the light's (so no t can agree with the light's in all but the six code bits of a key) the other
    // slots compete without their codes, seven issue slots each instead of eight (rect_slot<.., SH>)
    if (mode == PT_MODE_NEE_REF_RECT && S.light_code >= 0 && S.light_code < 3 * PT_RECT_SLOTS && !std::getenv("PTB200_NO_SHADOW_RAW")) {
        auto box = [&](int a, int k, float lo[3], float hi[3]) {        // AX 0: plane y (u = x, v = z), 1: plane z (x, y), 2: plane x (y, z)
            const float4 r = S.slot_a[a][k];
            const int ax = a == 0 ? 1 : a == 1 ? 2 : 0, au = a == 2 ? 1 : 0, av = a == 1 ? 1 : 2;
            lo[ax] = hi[ax] = r.x; lo[au] = r.y; hi[au] = r.y + r.z; lo[av] = r.w; hi[av] = r.w + S.slot_b2[a][k];
        };
        const int la = S.light_code / PT_RECT_SLOTS, lk = S.light_code % PT_RECT_SLOTS;
        float llo[3], lhi[3], ext = 0.f;
        bool apart = lk < S.n_slot[la];
        if (apart) box(la, lk, llo, lhi);
        double dmin = 1e30;
        for (int a = 0; a < 3 && apart; a++)
            for (int k = 0; k < S.n_slot[a]; k++) {
                float lo[3], hi[3];
                box(a, k, lo, hi);
                for (int c = 0; c < 3; c++) ext = std::max(ext, std::max(std::fabs(lo[c]), std::fabs(hi[c])));
                if (a == la && k == lk) continue;
                double d2 = 0;
                for (int c = 0; c < 3; c++) { const double g = std::max(0.0, std::max((double)lo[c] - lhi[c], (double)llo[c] - hi[c])); d2 += g * g; }
                dmin = std::min(dmin, std::sqrt(d2));
            }
        // two hits on one ray are at least the rectangles' distance apart; a key tie needs them within 7.6e-6 of t <= the scene's extent
        if (apart && dmin > 1e-4 * (double)ext * 3.5) h += "#define PT_J_SHADOW_RAW 1\n";
    }
    // Cone sampling of many sphere lights keeps a shadow-ray loop's state live across the unrolled scan: inlined, the
    // 64-register budget spills ~1.7 KB per thread (synthetic scene: 23 Mpaths/s); as a call, 53 Mpaths/s.
    if (mode == PT_MODE_NEE_CONE_SPHERE && S.n_sph4 >= 64 && S.n_sph4 <= PT_JIT_SPH_IMM_MAX && S.n_lights > 1) h += "#define PT_NOINLINE_HIT 1\n";
    // Long immediate sphere tables make the kernel instruction-fetch bound (C4: `no_instruction` is the top stall, 4.4 per issue): the
    // warps of a block then iterate in lockstep (one block-wide barrier per bounce) and walk the straight-line scan together,
    // sharing instruction-cache lines: +9.5 % on C4 (-3 % on scene A, where it stays off).
    // ... in blocks of 512 threads, so that 16 warps share the lines they fetch: another +6 % on C4 (1024: +3 %, 128: -8 %)
    if (S.n_sph4 >= 64 && S.n_sph4 <= PT_JIT_SPH_IMM_MAX) h += "#ifndef PT_NO_LOCKSTEP\n#define PT_LOCKSTEP 1\n#endif\n";
    {
        char bl[96];
        std::snprintf(bl, sizeof bl, "#ifndef PT_BLOCK\n#define PT_BLOCK %d\n#endif\n", pt_jit_block(S));      // threads per block of this module
        h += bl;
    }
    if (S.n_sph4 > 0 && S.n_sph4 <= PT_JIT_SPH_IMM_MAX) {      // small sphere sets: the scan table as immediates
        std::snprintf(b, sizeof b, "#define PT_J_SPH_IMM %d\nconstexpr float PT_J_SPHF[%d][4] = {\n", PT_JIT_SPH_IMM_MAX, S.n_sph4);
        h += b;
        for (int k = 0; k < S.n_sph4; k++) {
            h += " {"; put_float(h, S.sphf[k].x); h += ","; put_float(h, S.sphf[k].y); h += ","; put_float(h, S.sphf[k].z); h += ","; put_float(h, S.sphf[k].w); h += "},\n";
        }
        h += "};\n";
    }
    if (S.n_tilt > 0 && S.n_tilt <= 16) {          // tilted planes: {n, n.p0} {s, s.p0} {t, t.p0} {hs, ht, -, -} as literals
        std::snprintf(b, sizeof b, "#define PT_J_TILT_IMM 1\nconstexpr float PT_J_TILT[%d][16] = {\n", S.n_tilt);
        h += b;
        for (int k = 0; k < S.n_tilt; k++) {
            h += " {";
            for (int r = 0; r < 4; r++) {
                const float4 v = S.tilt[k][r];
                put_float(h, v.x); h += ","; put_float(h, v.y); h += ","; put_float(h, v.z); h += ","; put_float(h, v.w); h += ",";
            }
            h += "},\n";
        }
        h += "};\n";
    }
    if (S.n_huge > 0 && S.n_huge <= 16) {          // the huge-sphere table (FP64) as literals
        std::snprintf(b, sizeof b, "#define PT_J_HUGE_IMM 1\nconstexpr double PT_J_HUGE[%d][4] = {\n", S.n_huge);
        h += b;
        for (int k = 0; k < S.n_huge; k++) {
            std::snprintf(b, sizeof b, " {%a, %a, %a, %a},\n", S.huge[k][0], S.huge[k][1], S.huge[k][2], S.huge[k][3]);
            h += b;
        }
        h += "};\n";
        std::snprintf(b, sizeof b, "constexpr float PT_J_HUGEG[%d][8] = {\n", S.n_huge);      // the re-centred FP32 form
        h += b;
        for (int k = 0; k < S.n_huge; k++) {
            h += " {";
            for (int a = 0; a < 8; a++) { put_float(h, S.hugeg[k][a]); h += ","; }
            h += "},\n";
        }
        h += "};\nconstexpr float PT_J_huge_c[3] = {"; put_float(h, S.huge_c[0]); h += ","; put_float(h, S.huge_c[1]); h += ","; put_float(h, S.huge_c[2]); h += "};\n";
    }
    h += "constexpr float PT_J_sph_c[3] = {"; put_float(h, S.sph_c[0]); h += ","; put_float(h, S.sph_c[1]); h += ","; put_float(h, S.sph_c[2]); h += "};\n";
    return h;
}

// NVRTC: specialisation header + embedded kernel source -> sm_100a cubin.  No GPU needed (unit-tested on CPU).
int pt_jit_compile(const std::string &spec, std::vector<char> &cubin, std::string &log, double *seconds)
{
    Nvrtc &n = nvrtc();
    if (!n.ok) { log = n.why; return PT_ERR_STATE; }
    const auto t0 = std::chrono::steady_clock::now();
    std::string src = spec;
    src += PT_KERNEL_SRC;
    nvrtcProgram prog = nullptr;
    std::string name = "pt_kernel_jit.cu";
    if (const char *keep = std::getenv("PTB200_JIT_KEEP_SRC")) {   // write the translation unit where ncu --import-source finds it
        name = keep;
        if (FILE *f = std::fopen(keep, "w")) { std::fwrite(src.data(), 1, src.size(), f); std::fclose(f); }
    }
    if (n.CreateProgram(&prog, src.c_str(), name.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return PT_ERR_STATE; }
    std::vector<std::string> extra;                       // PTB200_JIT_OPTS: extra NVRTC options (tuning experiments)
    if (const char *e = std::getenv("PTB200_JIT_OPTS")) {
        std::string cur;
        for (const char *q = e;; q++) {
            if (*q == ' ' || *q == 0) { if (!cur.empty()) extra.push_back(cur); cur.clear(); if (!*q) break; }
            else cur += *q;
        }
    }
    // --fmad=false: the compiler never contracts a * b + c on its own, every FMA of the kernel is an explicit fmaf; that is what keeps this
    // build and the ahead-of-time one (nvcc -fmad=false) bit-identical whatever the surrounding code looks like
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "--fmad=false"};
    for (const std::string &x : extra) opts.push_back(x.c_str());
    const nvrtcResult rc = n.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); n.GetProgramLog(prog, &log[0]); }
    if (rc != NVRTC_SUCCESS) { n.DestroyProgram(&prog); if (log.empty()) log = "nvrtcCompileProgram failed"; return PT_ERR_STATE; }
    size_t cs = 0;
    if (n.GetCUBINSize(prog, &cs) != NVRTC_SUCCESS || cs == 0) { n.DestroyProgram(&prog); log = "no cubin"; return PT_ERR_STATE; }
    cubin.resize(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    if (const char *dump = std::getenv("PTB200_JIT_DUMP")) {       // keep the cubin for cuobjdump -sass
        if (FILE *f = std::fopen(dump, "wb")) { std::fwrite(cubin.data(), 1, cubin.size(), f); std::fclose(f); }
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return PT_OK;
}

// ---- disk cache of built modules: $PTB200_CACHE_DIR, else $XDG_CACHE_HOME/ptb200, else $HOME/.cache/ptb200 ("off" disables).
// Key = FNV-1a of the specialisation header, the embedded kernel source and the NVRTC options: any change of the
// library's device code or of the scene gives another file.
static std::string cache_dir()
{
    const char *e = std::getenv("PTB200_CACHE_DIR");
    std::string d;
    if (e) { if (!*e || std::string(e) == "off") return ""; d = e; }
    else if (const char *x = std::getenv("XDG_CACHE_HOME")) d = std::string(x) + "/ptb200";
    else if (const char *h = std::getenv("HOME")) d = std::string(h) + "/.cache/ptb200";
    else return "";
    std::string partial;
    for (size_t i = 0; i <= d.size(); i++) {             // mkdir -p
        if (i == d.size() || d[i] == '/') { if (!partial.empty()) ::mkdir(partial.c_str(), 0755); }
        if (i < d.size()) partial += d[i];
    }
    struct stat st;
    if (::stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return "";
    return d;
}

static std::string cache_path(const std::string &spec)
{
    const std::string dir = cache_dir();
    if (dir.empty()) return "";
    unsigned long long hsh = 1469598103934665603ull;
    auto mix = [&](const char *p, size_t n) { for (size_t i = 0; i < n; i++) { hsh ^= (unsigned char)p[i]; hsh *= 1099511628211ull; } };
    mix(spec.data(), spec.size());
    mix(PT_KERNEL_SRC, sizeof(PT_KERNEL_SRC));
    if (const char *o = std::getenv("PTB200_JIT_OPTS")) mix(o, std::strlen(o));
    { Nvrtc &n = nvrtc(); const int v[2] = {n.major, n.minor}; mix((const char *)v, sizeof v); }   // another compiler, another file
    char b[64];
    std::snprintf(b, sizeof b, "/ptb200-%016llx.cubin", hsh);
    return dir + b;
}

// Cubin for a specialisation: from the disk cache when present, else NVRTC (and the result is stored).
int pt_jit_build(const std::string &spec, std::vector<char> &cubin, std::string &log, double *seconds, bool *from_disk)
{
    if (from_disk) *from_disk = false;
    const std::string path = std::getenv("PTB200_JIT_KEEP_SRC") || std::getenv("PTB200_JIT_DUMP") ? std::string() : cache_path(spec);
    if (!path.empty()) {
        if (FILE *f = std::fopen(path.c_str(), "rb")) {
            std::fseek(f, 0, SEEK_END);
            const long sz = std::ftell(f);
            std::fseek(f, 0, SEEK_SET);
            if (sz > 1024) {
                cubin.resize((size_t)sz);
                const size_t got = std::fread(cubin.data(), 1, (size_t)sz, f);
                std::fclose(f);
                if (got == (size_t)sz && std::memcmp(cubin.data(), "\x7f" "ELF", 4) == 0) {
                    if (from_disk) *from_disk = true;
                    if (seconds) *seconds = 0;
                    return PT_OK;
                }
            } else std::fclose(f);
        }
    }
    int rc = pt_jit_compile(spec, cubin, log, seconds);
    if (rc == PT_OK && !path.empty()) {                   // write-then-rename: concurrent ranks never see a partial file
        char tmp[64];
        std::snprintf(tmp, sizeof tmp, ".tmp%d", (int)::getpid());
        const std::string t = path + tmp;
        if (FILE *f = std::fopen(t.c_str(), "wb")) {
            const bool ok = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
            std::fclose(f);
            if (!ok || std::rename(t.c_str(), path.c_str()) != 0) std::remove(t.c_str());
        }
    }
    return rc;
}

// Load a built cubin into the calling thread's CUDA context.
static PtJitKernel *load_module(pt_ctx *ctx, int mode, bool with_intersect, int rc, std::vector<char> &cubin, std::string &log, double secs, bool from_disk)
{
    PtJitKernel *jk = nullptr;
    if (rc == PT_OK) {
        jk = new PtJitKernel();
        jk->compile_seconds = secs;
        cudaError_t e = cudaLibraryLoadData(&jk->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (e == cudaSuccess) e = cudaLibraryGetKernel(&jk->kern, jk->lib, "k_bounce_jit");
        if (e == cudaSuccess && with_intersect && cudaLibraryGetKernel(&jk->kern_isect, jk->lib, "k_intersect_jit") != cudaSuccess) { jk->kern_isect = nullptr; cudaGetLastError(); }
        if (e != cudaSuccess) {
            log = std::string("loading the specialised module: ") + cudaGetErrorString(e);
            cudaGetLastError();
            delete jk;
            jk = nullptr;
        }
    }
    if (!jk) {
        ctx->jit_note = "generic kernel (" + log.substr(0, 300) + ")";
        if (std::getenv("PTB200_JIT_VERBOSE")) std::fprintf(stderr, "[ptb200] JIT unavailable: %s\n", log.c_str());
    } else if (std::getenv("PTB200_JIT_VERBOSE")) {
        std::fprintf(stderr, "[ptb200] scene-specialised k_bounce (mode %d) %s in %.2f s\n", mode, from_disk ? "loaded from the disk cache" : "compiled", secs);
    }
    return jk;
}

// The specialised kernel for the context's current scene, or nullptr (generic kernel) when JIT is off/unavailable.
// wait = true: build now if need be (large renders: the 0.5 s pay off at once).  wait = false (small renders): never
// block — the SECOND request for a specialisation starts its build on a host thread, later requests pick the kernel up
// once it is there, and until then the caller renders with the generic kernel (same image, bit for bit).
PtJitKernel *pt_jit_get(pt_ctx *ctx, int mode, bool stats, bool with_intersect, bool wait, int render_flags)
{
    if (!ctx->fp32_ok) return nullptr;
    const std::string spec = pt_jit_spec(*ctx->h_scene32, mode, stats, with_intersect, render_flags);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(spec);
    if (it != g_cache.end()) return it->second;          // may be nullptr: a failed build is not retried
    auto pit = g_pending.m.find(spec);
    if (pit == g_pending.m.end()) {
        if (wait) {
            std::vector<char> cubin;
            std::string log;
            double secs = 0;
            bool from_disk = false;
            const int rc = pt_jit_build(spec, cubin, log, &secs, &from_disk);
            return g_cache[spec] = load_module(ctx, mode, with_intersect, rc, cubin, log, secs, from_disk);
        }
        // buy after renting for the price: the build starts once the small renders of this specialisation have spent
        // about a compilation's worth of GPU time in the generic kernel (a process that renders once never compiles,
        // and none waits at exit for a build longer than it has been rendering)
        if (g_spent_ms[spec] < background_after_ms()) return nullptr;
        if (!nvrtc().ok) { ctx->jit_note = "generic kernel (" + nvrtc().why + ")"; return g_cache[spec] = nullptr; }
        if (!g_exit_hook) { g_exit_hook = true; std::atexit(join_pending_at_exit); }
        PendingBuild *pb = new PendingBuild();
        g_pending.m[spec] = pb;
        pb->th = std::thread([pb, spec]() {
            pb->rc = pt_jit_build(spec, pb->cubin, pb->log, &pb->secs, &pb->from_disk);
            pb->done.store(true, std::memory_order_release);
        });
        return nullptr;
    }
    PendingBuild *pb = pit->second;
    if (!wait && !pb->done.load(std::memory_order_acquire)) return nullptr;
    pb->th.join();
    PtJitKernel *jk = load_module(ctx, mode, with_intersect, pb->rc, pb->cubin, pb->log, pb->secs, pb->from_disk);
    g_pending.m.erase(pit);
    delete pb;
    g_spent_ms.erase(spec);
    return g_cache[spec] = jk;
}

// A small render of (scene, mode) ran the generic kernel for `ms`: counts towards starting its background build.
void pt_jit_account(pt_ctx *ctx, int mode, bool stats, double ms, int render_flags)
{
    if (!ctx->fp32_ok) return;
    const std::string spec = pt_jit_spec(*ctx->h_scene32, mode, stats, false, render_flags);
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_cache.find(spec) == g_cache.end() && g_pending.m.find(spec) == g_pending.m.end()) g_spent_ms[spec] += ms;
}
