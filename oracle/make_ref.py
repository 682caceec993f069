#!/usr/bin/env python3
"""make_ref.py — derive and build the patched-reference oracle in oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.  No reference source is committed: this script READS
/root/reference/src/smallpt.cpp where it lies, applies the textual patches P0-P6 of
SURVEY.md Appendix B to an in-memory copy, writes the result to oracle/_ref/smallpt_ref.cpp
and compiles it (utilities.h is included from the reference tree via -I).

Why patches are needed at all: at HEAD radiance() returns a debug colour
at src/smallpt.cpp:442, so the Monte Carlo code (:444-480) is dead; jitter/light draws use a
time-seeded libc rand() (non-reproducible, and `rand()*36` overflows int on glibc); the
OpenMP pragma is commented out; `Sphere` is abstract.

  P0  skip create_state_space (:517)                      -> main() is replaced (see MAIN)
  P1  delete the RL block :424-442
  P2  every rand()/RAND_MAX draw -> erand48(Xi)           (:365,366,533,534); q (:460) = mode constant
  P3  uniform sampler (:351-360) live under PT_MODE == 2
  P4  spp/mode/scene/size from argv, live OpenMP row loop, per-row Xi with explicit u16 cast
  P5  default bodies for Hitable::add_key/add_value; runtime scene table (scene B / C); light id variable
  P6  optional deterministic sincos (include/ptb200_detmath.h) under PT_DETSC

Everything else — Vec, Rectangle_*, Sphere, Camera, intersect(), hittingPoint(),
random_scattering(), light_sampling(), radiance(), clamp(), toInt(), erand48 — is the
reference's own text, compiled as is.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PT_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def replace_once(text, old, new, what):
    n = text.count(old)
    if n != 1:
        raise SystemExit(f"make_ref: patch '{what}': expected exactly 1 occurrence, found {n}")
    return text.replace(old, new)


MAIN = r'''
// ---- P4: driver (replaces main(), src/smallpt.cpp:502-557; loop body text kept, see :528-541) ----
#include <omp.h>
#include <string>
static void dump(const std::string &path, const double *p, size_t n) {
	FILE *f = fopen(path.c_str(), "wb"); fwrite(p, sizeof(double), n, f); fclose(f);
}
int main(int argc, char *argv[]) {
	int samps = argc > 1 ? atoi(argv[1]) : 16;
	PT_MODE = argc > 2 ? atoi(argv[2]) : 0;
	char scene = argc > 3 ? argv[3][0] : 'A';
	int w = argc > 4 ? atoi(argv[4]) : 512, h = argc > 5 ? atoi(argv[5]) : 512;
	std::string out = argc > 6 ? argv[6] : "";
	PT_DETSC = argc > 7 ? atoi(argv[7]) : 0;
	if (argc > 8 && atoi(argv[8]) > 0) omp_set_num_threads(atoi(argv[8]));
	if (scene == 'B') {			// sphere-era scene recovered from src/a.exe (SURVEY Appendix A)
		int k = 0;
		rect[k++] = new SphereP5(1e5, Vec(1e5 + 1, 40.8, 81.6), Vec(), Vec(.25, .75, .25), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(1e5 + 1, 40.8, 81.6), Vec(), Vec(.25, .75, .25), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(-1e5 + 99, 40.8, 81.6), Vec(), Vec(.75, .25, .25), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(50, 40.8, 1e5), Vec(), Vec(.75, .75, .75), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(50, 40.8, -1e5 + 170), Vec(), Vec(), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(50, 1e5, 81.6), Vec(), Vec(.75, .75, .75), DIFF);
		rect[k++] = new SphereP5(1e5, Vec(50, -1e5 + 81.6, 81.6), Vec(), Vec(.75, .75, .75), DIFF);
		rect[k++] = new SphereP5(16.5, Vec(27, 16.5, 47), Vec(), Vec(1, 1, 1) * .999, DIFF);
		rect[k++] = new SphereP5(16.5, Vec(73, 16.5, 78), Vec(), Vec(.75, .75, .75), DIFF);
		rect[k++] = new SphereP5(600, Vec(50, 681.6 - .27, 81.6), Vec(12, 12, 12), Vec(), DIFF);
		NUMBER_OBJ = k; LIGHT_ID = 9;
	} else if (scene == 'C') {	// rect walls + light + the two spheres of :297-298 (image_light_test.ppm)
		rect[7] = new SphereP5(16.5, Vec(27, 16.5, 47), Vec(), Vec(1, 1, 1) * .999, DIFF);
		rect[8] = new SphereP5(16.5, Vec(73, 16.5, 78), Vec(), Vec(.75, .75, .75), DIFF);
		NUMBER_OBJ = 9; LIGHT_ID = 6;
	}
	Vec *c = new Vec[w * h];		// :510
	double *mean = new double[(size_t)w * h * 3], *sumsq = new double[(size_t)w * h * 3];
	std::map<Key, QValue> *dict = new std::map<Key, QValue>;
	Camera cam(LOOKFROM, Vec(50, 40, 5), Vec(0, 1, 0), 65, float(w) / float(h));	// :521
	double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1)						// :526
	for (int y = 0; y < h; y++) {									// :528
		int i = y * w;
		Vec r;
		double path_length = 0; int counter_red = 0;
		unsigned short Xi[3] = { 0, 0, (unsigned short)((unsigned)y * (unsigned)y * (unsigned)y) };	// :530
		for (int x = 0; x < w; x++) {
			Vec m, sq;
			for (int s = 0; s < samps; s++) {
				float u = float(x - 0.5 + erand48(Xi)) / float(w);				// :533 (P2)
				float v = float((h - y - 1) - 0.5 + erand48(Xi)) / float(h);	// :534 (P2)
				Ray d = cam.get_ray(u, v);
				Vec L = radiance(Ray(cam.origin, d.d.norm()), 0, Xi, &path_length, dict, counter_red);
				r = r + L * (1. / samps);									// :536
				m = m + L; sq = sq + L.mult(L);
			}
			c[i] = c[i] + Vec(clamp(r.x), clamp(r.y), clamp(r.z));			// :538
			mean[3 * i] = m.x / samps; mean[3 * i + 1] = m.y / samps; mean[3 * i + 2] = m.z / samps;
			sumsq[3 * i] = sq.x; sumsq[3 * i + 1] = sq.y; sumsq[3 * i + 2] = sq.z;
			i++;
			r = Vec();
		}
	}
	double t1 = omp_get_wtime();
	if (!out.empty()) {
		double *flat = new double[(size_t)w * h * 3];
		for (int i = 0; i < w * h; i++) { flat[3 * i] = c[i].x; flat[3 * i + 1] = c[i].y; flat[3 * i + 2] = c[i].z; }
		dump(out + ".clamped.f64", flat, (size_t)w * h * 3);
		dump(out + ".mean.f64", mean, (size_t)w * h * 3);
		dump(out + ".sumsq.f64", sumsq, (size_t)w * h * 3);
		FILE *f = fopen((out + ".ppm").c_str(), "w");						// :548-551
		fprintf(f, "P3\n%d %d\n%d\n", w, h, 255);
		for (int i = 0; i < w * h; i++)
			fprintf(f, "%d %d %d ", toInt(c[i].x), toInt(c[i].y), toInt(c[i].z));
		fclose(f);
	}
	printf("{\"paths\": %.0f, \"render_ms\": %.3f, \"threads\": %d, \"w\": %d, \"h\": %d, \"spp\": %d, \"mode\": %d, \"scene\": \"%c\"}\n",
	       double(w) * h * samps, (t1 - t0) * 1e3, omp_get_max_threads(), w, h, samps, PT_MODE, scene);
	return 0;
}
'''


def derive(src):
    t = src
    # --- P5: runtime-sized scene, concrete Sphere
    t = replace_once(t, "const int NUMBER_OBJ = 17;",
                     "int NUMBER_OBJ = 17; int LIGHT_ID = 6; int PT_MODE = 0; int PT_DETSC = 0;\n"
                     '#include "ptb200_detmath.h"\n'
                     "#define PT_COS(a) (PT_DETSC ? pt_det_cos(a) : cos(a))\n"
                     "#define PT_SIN(a) (PT_DETSC ? pt_det_sin(a) : sin(a))", "P5 NUMBER_OBJ")
    t = replace_once(t, "virtual std::array<float, 3> add_key(Vec &pos) const = 0;",
                     "virtual std::array<float, 3> add_key(Vec &pos) const { return {0, 0, 0}; }", "P5 add_key")
    t = replace_once(t, "virtual std::array<float, 3> add_value(std::array<float, 3>& x_reduced) const = 0;",
                     "virtual std::array<float, 3> add_value(std::array<float, 3>& x_reduced) const { return {0, 0, 0}; }",
                     "P5 add_value")
    t = replace_once(t, "Hitable *rect[NUMBER_OBJ] = {", "typedef Sphere SphereP5;\nHitable *rect[64] = {", "P5 table")
    t = replace_once(t, "if (id != 6) {", "if (id != LIGHT_ID) {", "P5 light id")
    # --- P2: all draws from Xi
    t = replace_once(t, "double x_light = 32 + rand() * 36 / double(RAND_MAX);",
                     "double x_light = 32 + 36 * erand48(Xi);", "P2 x_light")
    t = replace_once(t, "double z_light = 63 + rand() * 36 / double(RAND_MAX);",
                     "double z_light = 63 + 36 * erand48(Xi);", "P2 z_light")
    t = replace_once(t, "double q = rand() / double(RAND_MAX);", "double q = (PT_MODE == 0) ? 0 : 2;", "P2 q")
    # --- P3 (+P6): uniform sampler under a flag, selectable sincos
    cos_line = "return (u * cos(r1) * r2s + v * sin(r1) * r2s + w * sqrt(1 - r2)).norm();"
    t = replace_once(t, cos_line,
                     "if (PT_MODE == 2) return (u*PT_COS(r1)*sqrt(r2*(2-r2)) + v*PT_SIN(r1)*sqrt(r2*(2-r2)) + w*(1-r2)).norm();\n"
                     "\treturn (u * PT_COS(r1) * r2s + v * PT_SIN(r1) * r2s + w * sqrt(1 - r2)).norm();", "P3 sampler")
    # --- P1: delete the RL block :424-442
    a = t.index("std::map<Key, QValue> &addrDict = *dict;\n\n\tKey key = rect[id]->add_key(x);")
    end_marker = "return Vec(addrDict[key][0], addrDict[key][1], addrDict[key][2]);"
    b = t.index(end_marker, a) + len(end_marker)
    t = t[:a] + "/* P1: RL block :424-442 removed */" + t[b:]
    # --- P0/P4: replace main()
    m = t.index("int main(int argc, char *argv[]) {")
    t = t[:m] + MAIN
    return t


def build(verbose=True):
    src_path = os.path.join(REF, "src", "smallpt.cpp")
    if not os.path.exists(src_path):
        if verbose:
            print(f"make_ref: {src_path} not present; keeping any prebuilt oracle/_ref/", file=sys.stderr)
        return False
    os.makedirs(OUT, exist_ok=True)
    with open(src_path) as f:
        src = f.read()
    derived = derive(src)
    cpp = os.path.join(OUT, "smallpt_ref.cpp")
    with open(cpp, "w") as f:
        f.write("// GENERATED by oracle/make_ref.py from the read-only reference; git-ignored; do not commit.\n")
        f.write(derived)
    inc = ["-I", os.path.join(REF, "src"), "-I", os.path.join(HERE, "..", "include")]
    flags = ["-O3", "-fopenmp", "-ffp-contract=off", "-w"]
    subprocess.check_call(["g++", *flags, *inc, cpp, "-o", os.path.join(OUT, "smallpt_ref")])
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", *inc,
                           os.path.join(HERE, "ref_harness.cpp"), "-o", os.path.join(OUT, "librefharness.so")])
    if verbose:
        print("make_ref: built oracle/_ref/smallpt_ref and oracle/_ref/librefharness.so")
    return True


if __name__ == "__main__":
    build()
