// pt_wavefront.cu — production engine (PT_ENGINE_FP32_PHILOX): an FP32 wavefront path tracer for sm_100a.
//
// What it replaces (reference src/smallpt.cpp): the triple loop :528-541, radiance() :419-496,
// intersect()/hittingPoint() :323-335/:371-377, the per-primitive intersect()/normal() members
// (:102-124, :145-167, :188-210, :229-253), random_scattering() :337-360 and light_sampling() :363-369.
//
// This file is the HOST side: queue/accumulator management, the launch loop, resolve.  The device code lives in
// pt_kernel.cuh (compiled here ahead of time, and by pt_jit.cu at run time with the scene's constants as immediates).
//
// Organisation (one kernel launch = up to 512 bounces of every path slot):
//   * k_bounce keeps the path state in registers across bounces: extend (closest hit), shade (emission, Russian
//     roulette, light sampling + shadow ray, BSDF sampling) and REGENERATION (a lane whose path ended takes the next
//     camera path at once: ray generation with uniform sub-pixel jitter; path indices are reserved per warp in chunks)
//     repeat inside the kernel; COMPACTION (warp ballot + block prefix sum + one atomic per block) of the survivors into
//     the output queue happens once per launch and matters in the tail of a render;
//   * between launches the survivors live in SoA float4 queues (3 x 16 B per path: origin+pixel, direction+sample,
//     throughput+depth/prev/E); loads and stores are 128-bit and fully coalesced;
//   * the scene sits in __constant__ memory (generic build) or in the instruction stream (specialised build), sorted
//     by primitive class; small-sphere tables are staged in shared memory for the conservative scan;
//   * randomness is Philox4x32-10 keyed by (pixel, sample, vertex), one block per vertex: the image does
//     not depend on queue order, chunking, bounces per launch or the number of GPUs;
//   * radiance is accumulated per pixel in 64-bit fixed point (2^-24) with integer atomics, which are
//     associative: the result is bit-reproducible run to run and across shardings; k_resolve / k_resolve_owned turn
//     the sums into FP64 (the latter writes only this rank's rows, possibly into another GPU's image).
//
// FP32 numerics: rectangles keep the reference's no-epsilon rule and its (k-o)/d, o+d*t
// forms for the winning hit so the self-hit "leak" statistics carry over; spheres use the
// perpendicular-distance discriminant; spheres with radius >= PT_HUGE_RADIUS get the c = |o-p|^2-r^2 term
// in FP64 and conjugate roots; the sphere a ray starts on is solved exactly (roots {0, 2b}).
#include <math_constants.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "pt_internal.h"
#include "pt_kernel.cuh"

namespace {

// fixed point -> double sums, restricted to the rows this rank owns (foreign rows stay zero)
__global__ void k_resolve(const unsigned long long *__restrict__ fix, const unsigned long long *__restrict__ fixsq,
                          double *__restrict__ sum, double *__restrict__ sumsq, size_t n, unsigned long long *stamp)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *stamp = global_timer_ns();
    if (i >= n) return;
    sum[i] = (double)fix[i] * PT_FIX_INV;
    if (sumsq && fixsq) sumsq[i] = (double)fixsq[i] * PT_FIX_INV;
}

// the same for the rows of this rank only (owned_rows_only): `sum` may be another GPU's image, written over NVLink
// peer memory — the gather of the row tiles is these stores
__global__ void k_resolve_owned(const unsigned long long *__restrict__ fix, double *__restrict__ sum, unsigned long long owned_pixels,
                                int w, int tile_rows, int rank, int world, unsigned long long *stamp)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *stamp = global_timer_ns();
    if (i >= owned_pixels * 3ull) return;
    const unsigned long long lp = i / 3ull;
    const unsigned int ch = (unsigned int)(i - lp * 3ull);
    const unsigned int row_local = (unsigned int)(lp / (unsigned int)w), x = (unsigned int)(lp - (unsigned long long)row_local * (unsigned int)w);
    const unsigned int tile = row_local / (unsigned int)tile_rows;
    const unsigned int y = (tile * (unsigned int)world + (unsigned int)rank) * (unsigned int)tile_rows + (row_local - tile * (unsigned int)tile_rows);
    const size_t idx = ((size_t)y * (size_t)w + x) * 3 + ch;
    sum[idx] = (double)fix[idx] * PT_FIX_INV;
}

// end-of-render time stamp (one thread)
__global__ void k_stamp(unsigned long long *stamp) { *stamp = global_timer_ns(); }

__global__ void k_philox(const uint32_t *__restrict__ ctr, const uint32_t *__restrict__ key, int n, uint32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
    // the render kernels use the round-key form: it must be the same function (a mismatch spoils the output, so the
    // known-answer test fails)
    uint32_t rk[10][2];
    philox_round_keys(key[2 * i], key[2 * i + 1], rk);
    const uint4 q = philox4x32_10_rk(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], rk);
    if (q.x != r.x || q.y != r.y || q.z != r.z || q.w != r.w) r = make_uint4(~r.x, ~r.y, ~r.z, ~r.w);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// FFMA-only microbenchmark: 8 independent chains per thread, register operands.
__global__ void __launch_bounds__(256) k_ffma_peak(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE> void launch_bounce(bool stats, bool glossy, int blocks, cudaStream_t s, const KParams &P, const PtJitKernel *jk, int jit_block)
{
    if (jk) {           // scene-specialised module (pt_jit.cu): same KParams, launched through its kernel handle with ITS block size
        void *args[] = {(void *)&P};
        cudaLaunchKernel((const void *)jk->kern, dim3(blocks * PT_BLOCK / jit_block), dim3(jit_block), args, 0, s);
        return;
    }
    // ahead-of-time build: DIFF-only scenes (every scene of the reference) run the instantiation without SPEC / REFR code;
    // scenes whose small spheres sit in the uniform grid run the GRID instantiation (built with every material)
    if (P.grid.n > 0) { if (stats) k_bounce<MODE, true, true, true><<<blocks, PT_BLOCK, 0, s>>>(P); else k_bounce<MODE, false, true, true><<<blocks, PT_BLOCK, 0, s>>>(P); }
    else if (stats) { if (glossy) k_bounce<MODE, true, true, false><<<blocks, PT_BLOCK, 0, s>>>(P); else k_bounce<MODE, true, false, false><<<blocks, PT_BLOCK, 0, s>>>(P); }
    else { if (glossy) k_bounce<MODE, false, true, false><<<blocks, PT_BLOCK, 0, s>>>(P); else k_bounce<MODE, false, false, false><<<blocks, PT_BLOCK, 0, s>>>(P); }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
static int ensure_queues(pt_ctx *ctx, int capacity, bool stats)
{
    if (ctx->q_capacity >= capacity && (!stats || ctx->q[0][3])) return PT_OK;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 4; b++) { if (ctx->q[a][b]) cudaFree(ctx->q[a][b]); ctx->q[a][b] = nullptr; }
    ctx->q_capacity = 0;
    // one allocation per array keeps every array 256 B aligned for the 128-bit accesses
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < (stats ? 4 : 3); b++)
            PT_CUDA(ctx, cudaMalloc(&ctx->q[a][b], sizeof(float4) * (size_t)capacity));
    if (ctx->d_warp_chunk) cudaFree(ctx->d_warp_chunk);
    ctx->d_warp_chunk = nullptr;
    PT_CUDA(ctx, cudaMalloc(&ctx->d_warp_chunk, sizeof(uint4) * (size_t)(capacity / 32)));
    ctx->q_capacity = capacity;
    return PT_OK;
}

// The layout of a render: a pure function of the context (SM count, scene materials) and the parameters.
void pt_fp32_plan(const pt_ctx *ctx, const pt_render_params *p, Fp32Plan &pl)
{
    pl = Fp32Plan();
    const int w = p->width, h = p->height;
    const int tile = p->tile_rows > 0 ? p->tile_rows : 8;
    const int world = p->world > 0 ? p->world : 1;
    const bool stats = p->collect_stats != 0;
    // rows owned by this rank
    const int n_tiles = (h + tile - 1) / tile;
    for (int k = p->rank; k < n_tiles; k += world) pl.owned_rows += (k * tile + tile <= h) ? tile : (h - k * tile);
    pl.owned_pixels = (unsigned long long)pl.owned_rows * w;
    if (pl.owned_pixels == 0 || p->spp <= 0) return;
    // pixel blocks (KParams::blk_pixels): every sample of a block before the next block, accumulators of a block L2-resident
    unsigned long long blk_target = 1ull << 19;                      // 512 Ki pixels = 12.6 MB of fixed-point accumulators
    if (const char *e = std::getenv("PTB200_BLOCK_PIXELS")) blk_target = std::max(1ll, std::atoll(e));
    pl.n_blk = (pl.owned_pixels + blk_target - 1) / blk_target;
    if (pl.n_blk > (unsigned long long)pl.owned_rows) pl.n_blk = (unsigned long long)pl.owned_rows;
    if (pl.n_blk * (unsigned long long)p->spp >= (1ull << 31)) pl.n_blk = 1;             // (block * runs per pixel + run is a 32-bit value)
    pl.blk_rows = ((unsigned long long)pl.owned_rows + pl.n_blk - 1) / pl.n_blk;
    pl.blk_pixels = pl.blk_rows * (unsigned long long)w;
    // queue capacity = path slots in flight = threads per launch.  Default: PT_DEFAULT_WAVES full waves of resident
    // blocks (no launch ends with a partially filled wave); the queues are touched once per launch, not per bounce.
    int cap = p->queue_capacity;
    const long long wave = (long long)ctx->sm_count * PT_BLOCKS_PER_SM * PT_BLOCK;
    if (cap <= 0) cap = (int)(PT_DEFAULT_WAVES * wave);
    {   // never more slots than paths (rounded up to whole blocks)
        const unsigned long long need = (pl.n_blk * pl.blk_pixels * (unsigned long long)p->spp + 1023) / 1024 * 1024;
        if ((unsigned long long)cap > need) cap = (int)need;
    }
    pl.cap = (cap + 1023) / 1024 * 1024;              // whole blocks for any block size up to 1024
    // REFR path splitting (:494-495): per-warp stacks of spawned branches, only for scenes that have such a material
    pl.want_spawn = ((ctx->h_scene32->refl_mask >> PT_REFR) & 1) && !stats && !std::getenv("PTB200_NO_SPLIT");
    // Sample runs (KParams::run_shift): 128 consecutive samples of a pixel per path index when a slot traces at least 2048
    // paths in this render, else single samples (measured on shares of C5: at 2336 paths per slot runs of 128 gain 3 %, at
    // 1168 runs of 64 gain 0.8 % on one GPU and lose as much in the 8-GPU step, runs of 32 nothing; on C2's 147 paths per
    // slot runs of 16 lose 10 %: a render of runs ends with every slot finishing the run it is in, and shorter runs end so
    // often that some lane of a warp needs the chunk step in most iterations anyway).  A lane that takes over spawned REFR
    // branches has no run of its own to come back to, so those scenes keep single samples too.
    {
        const unsigned long long per_slot = pl.owned_pixels * (unsigned long long)p->spp / (unsigned long long)pl.cap;
        unsigned long long run = per_slot >= 2048ull ? 128ull : 1ull;
        if (const char *e = std::getenv("PTB200_RUN")) run = (unsigned long long)std::max(1ll, std::atoll(e));
        if (pl.want_spawn) run = 1;
        while ((2ull << pl.run_shift) <= run && (2ull << pl.run_shift) <= (unsigned long long)p->spp && pl.run_shift < 12) pl.run_shift++;
    }
    pl.spp_runs = ((unsigned long long)p->spp + (1ull << pl.run_shift) - 1) >> pl.run_shift;
    pl.total = pl.n_blk * pl.blk_pixels * pl.spp_runs;     // path indices = runs, incl. the (< n_blk) skipped rows
    pl.use_magic = pl.owned_pixels < (1ull << 24) && w < 65536 && tile < 65536;
    pl.pix_magic = (unsigned long long)w * h < (1ull << 24) && w < 65536;
    pl.flags = (pl.blk_pixels >= 32ull ? PT_RF_WRAP_ONCE : 0) | (pl.use_magic && pl.pix_magic ? PT_RF_MAGIC : 0) | (world == 1 ? PT_RF_WORLD1 : 0)
             | (pl.n_blk == 1 ? PT_RF_ONE_BLOCK : 0) | (pl.run_shift > 0 ? PT_RF_RUNS : 0);
}

int pt_fp32_render(pt_ctx *ctx, const pt_render_params *p, double *d_sum, double *d_sumsq, cudaStream_t s)
{
    if (!ctx->fp32_ok) return pt_fail(ctx, PT_ERR_ARG, "scene does not fit the FP32 engine: " + ctx->fp32_why);
    const int w = p->width, h = p->height;
    const int tile = p->tile_rows > 0 ? p->tile_rows : 8;
    const int world = p->world > 0 ? p->world : 1;
    const bool stats = p->collect_stats != 0;
    if (p->mode == PT_MODE_NEE_CONE_SPHERE && ctx->h_scene32->n_lights == 0)
        return pt_fail(ctx, PT_ERR_ARG, "PT_MODE_NEE_CONE_SPHERE needs at least one emissive sphere");

    Fp32Plan pl;
    pt_fp32_plan(ctx, p, pl);
    const long long owned_rows = pl.owned_rows;
    const unsigned long long owned_pixels = pl.owned_pixels;
    const size_t n_acc = (size_t)w * h * 3;

    // fixed-point accumulators
    if (ctx->fix_elems < n_acc) {
        if (ctx->d_fix) cudaFree(ctx->d_fix);
        if (ctx->d_fixsq) cudaFree(ctx->d_fixsq);
        ctx->d_fix = ctx->d_fixsq = nullptr; ctx->fix_elems = 0;
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fix, n_acc * sizeof(unsigned long long)));
        PT_CUDA(ctx, cudaMalloc(&ctx->d_fixsq, n_acc * sizeof(unsigned long long)));
        ctx->fix_elems = n_acc;
    }
    const bool accumulate = p->accumulate != 0;
    if (accumulate) {       // render_common checked that the accumulators hold an FP32 image of this size
        if (stats && !ctx->fix_has_sq) return pt_fail(ctx, PT_ERR_STATE, "collect_stats = 1 cannot be added to accumulators without sums of squares");
    } else {
        PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fix, 0, n_acc * sizeof(unsigned long long), s));
        if (stats) PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fixsq, 0, n_acc * sizeof(unsigned long long), s));
        ctx->accum_spp = 0;
        ctx->fix_has_sq = stats;
    }
    if (!stats) ctx->fix_has_sq = false;
    ctx->accum_spp += p->spp;

    int it_total = 0;
    if (owned_pixels > 0 && p->spp > 0) {
        const unsigned long long n_blk = pl.n_blk, blk_rows = pl.blk_rows, blk_pixels = pl.blk_pixels, spp_runs = pl.spp_runs, total = pl.total;
        const int cap = pl.cap;
        const long long wave = (long long)ctx->sm_count * PT_BLOCKS_PER_SM * PT_BLOCK;
        const bool want_spawn = pl.want_spawn;
        const unsigned int run_shift = pl.run_shift;
        int rc = ensure_queues(ctx, cap, stats);
        if (rc) return rc;

        // per-iteration live counters (n[it]); counts[0] = cap (all dead => regenerate)
        const int max_it = 1 << 20;
        if (ctx->counts_len < max_it + 8) {
            if (ctx->d_counts) cudaFree(ctx->d_counts);
            ctx->d_counts = nullptr; ctx->counts_len = 0;
            PT_CUDA(ctx, cudaMalloc(&ctx->d_counts, sizeof(unsigned int) * (size_t)(max_it + 8)));
            ctx->counts_len = max_it + 8;
            ctx->counts_dirty = (size_t)ctx->counts_len;
        }
        {   // clear only the prefix the previous render dirtied (+ slack for the speculative batch)
            size_t dirty = ctx->counts_dirty + 256;
            if (dirty > (size_t)ctx->counts_len) dirty = (size_t)ctx->counts_len;
            // layout: [0..1] the generation counter (u64), [4..] n[it]
            PT_CUDA(ctx, cudaMemsetAsync(ctx->d_counts, 0, sizeof(unsigned int) * dirty, s));
        }
        if (want_spawn && ctx->spawn_warps < cap / 32) {
            if (ctx->d_spawn) cudaFree(ctx->d_spawn);
            ctx->d_spawn = nullptr; ctx->spawn_warps = 0;
            PT_CUDA(ctx, cudaMalloc(&ctx->d_spawn, sizeof(float4) * (3 * PT_SPAWN_SLOTS + 32) * (size_t)(cap / 32)));
            ctx->spawn_warps = cap / 32;
        }
        if (!ctx->d_launch_rec) PT_CUDA(ctx, cudaMalloc(&ctx->d_launch_rec, sizeof(LaunchRec) * (PT_MAX_LAUNCH_RECS + 1)));
        PT_CUDA(ctx, cudaMemsetAsync(ctx->d_warp_chunk, 0, sizeof(uint4) * (size_t)(cap / 32), s));   // n[0] = 0: every slot starts without a path
        const PtJitKernel *jk = ctx->jit;
        if (jk) {       // the specialised module has its own c_scene (lights, huge spheres, tilted planes, overflow)
            void *dptr = nullptr;
            size_t bytes = 0;
            cudaError_t e = cudaLibraryGetGlobal(&dptr, &bytes, jk->lib, "c_scene");
            if (e == cudaSuccess && bytes == sizeof(SceneF32)) e = cudaMemcpyAsync(dptr, ctx->h_scene32, sizeof(SceneF32), cudaMemcpyHostToDevice, s);
            else if (e == cudaSuccess) e = cudaErrorInvalidValue;
            if (e != cudaSuccess) { cudaGetLastError(); ctx->jit_note = std::string("generic kernel (c_scene of the specialised module: ") + cudaGetErrorString(e) + ")"; jk = nullptr; ctx->jit = nullptr; }
        }
        PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));

        KParams P{};
        P.gen_counter = (unsigned long long *)ctx->d_counts;
        P.warp_chunk = ctx->d_warp_chunk;
        P.capacity = (unsigned int)cap;
        P.iters = p->bounces_per_launch > 0 ? p->bounces_per_launch : PT_DEFAULT_ITERS;
        P.iters_tail = P.iters < PT_DEFAULT_ITERS_TAIL ? P.iters : PT_DEFAULT_ITERS_TAIL;
        P.iters_drain = p->bounces_per_launch > 0 ? P.iters : PT_DEFAULT_ITERS_DRAIN;
        P.drain_below = (unsigned int)wave;
        // with sample runs the survivors of the last path indices are slots in the middle of their runs (dozens of paths each, every
        // lane busy until its run ends): repacking them every few bounces only adds launches
        if (run_shift > 0 && p->bounces_per_launch <= 0) P.iters_tail = PT_DEFAULT_ITERS_DRAIN;
        // tuning overrides (tools/sweep_tail.py)
        if (const char *e = std::getenv("PTB200_ITERS_TAIL")) P.iters_tail = std::atoi(e) > 0 ? std::atoi(e) : P.iters_tail;
        if (const char *e = std::getenv("PTB200_ITERS_DRAIN")) P.iters_drain = std::atoi(e) > 0 ? std::atoi(e) : P.iters_drain;
        if (const char *e = std::getenv("PTB200_DRAIN_BELOW")) P.drain_below = (unsigned int)std::atoi(e);
        {   // indices a warp reserves per atomic: enough for about one launch, but small renders still spread over the GPU
            unsigned long long c = total / ((unsigned long long)(cap / 32) * 2ull);
            unsigned int chunk = 32;
            while (chunk < 256 && chunk * 2ull <= c) chunk *= 2;
            if (run_shift > 0) chunk = 32;          // (a run is dozens of paths: one per lane is plenty to reserve at a time)
            P.chunk = chunk;
            // sample runs: reserve ahead only while more than this many runs are left (default: 3 per slot)
            P.run_reserve = (unsigned long long)cap * (std::getenv("PTB200_RUN_RESERVE") ? (unsigned long long)std::max(0, std::atoi(std::getenv("PTB200_RUN_RESERVE"))) : 3ull);
            unsigned int sh = 0;
            while ((1ull << sh) < (unsigned long long)(cap / 32) / 2ull) sh++;     // fair share = remaining / (warps / 2)
            if (const char *e = std::getenv("PTB200_FAIR_DELTA")) { const int dlt = std::atoi(e); sh = (unsigned int)std::max(0, (int)sh + dlt); }
            P.fair_shift = sh;
        }
        P.total_paths = total;
        P.owned_pixels = (unsigned int)owned_pixels;
        P.blk_pixels = (unsigned int)blk_pixels; P.blk_rows = (unsigned int)blk_rows; P.n_blk = (unsigned int)n_blk; P.owned_rows = (unsigned int)owned_rows;
        P.inv_blk_pixels = 1.0 / (double)blk_pixels;
        P.run_shift = run_shift; P.run_mask = (1u << run_shift) - 1u; P.spp_runs = (unsigned int)spp_runs;
        P.pix_magic = pl.pix_magic ? 1 : 0;
        P.w = w; P.h = h; P.spp = p->spp; P.tile_rows = tile; P.rank = p->rank; P.world = world;
        P.magic_w = ((1ull << 40) + (unsigned long long)w - 1) / (unsigned long long)w;
        P.magic_tile = ((1ull << 40) + (unsigned long long)tile - 1) / (unsigned long long)tile;
        P.use_magic = pl.use_magic ? 1 : 0;
        P.grid = ctx->grid;
        P.key_low = PT_KEY_CODE_BITS;
        P.rect_tmin = p->robust_eps ? PT_EPS_F : 1.401298464e-45f;      // reference: no epsilon on rectangles (:106)
        P.wrap_once = blk_pixels >= 32ull ? 1 : 0;
        P.max_depth = p->max_depth > 0 ? p->max_depth : 4096;
        if (P.max_depth > 8000) P.max_depth = 8000;      // depth has 13 bits in a path record
        const pt_camera &c = ctx->cam;
        P.cam_o[0] = (float)c.origin.x; P.cam_o[1] = (float)c.origin.y; P.cam_o[2] = (float)c.origin.z;
        P.cam_base[0] = (float)(c.lower_left_corner.x - c.origin.x);
        P.cam_base[1] = (float)(c.lower_left_corner.y - c.origin.y);
        P.cam_base[2] = (float)(c.lower_left_corner.z - c.origin.z);
        P.cam_h[0] = (float)c.horizontal.x; P.cam_h[1] = (float)c.horizontal.y; P.cam_h[2] = (float)c.horizontal.z;
        P.cam_v[0] = (float)c.vertical.x; P.cam_v[1] = (float)c.vertical.y; P.cam_v[2] = (float)c.vertical.z;
        P.inv_w = 1.f / (float)w; P.inv_h = 1.f / (float)h;
        P.smp0 = (unsigned int)p->sample_offset;
        P.seed_lo = (unsigned int)p->seed; P.seed_hi = (unsigned int)(p->seed >> 32);
        philox_round_keys(P.seed_lo, P.seed_hi, P.philox_rk);
        P.fix = ctx->d_fix; P.fixsq = ctx->d_fixsq;
        P.mats = ctx->d_mats; P.sphf = ctx->d_sphf; P.stats = ctx->d_stats;
        P.spawn = want_spawn ? ctx->d_spawn : nullptr;
        P.spawn_slots = want_spawn ? PT_SPAWN_SLOTS : 0u;

        const int blocks = cap / PT_BLOCK;
        const int jit_block = jk ? pt_jit_block(*ctx->h_scene32) : PT_BLOCK;
        const bool glossy = (ctx->h_scene32->refl_mask & ((1 << PT_SPEC) | (1 << PT_REFR))) != 0;
        unsigned int *n_it = ctx->d_counts + 4;
        // Termination check without draining the pipeline: batch k+1 is enqueued before the live count
        // after batch k is read back (pinned slot + event per parity).
        unsigned int *h_n = ctx->h_pinned;                 // pinned slots + events live in the context
        cudaEvent_t *evb = ctx->ev_batch;
        int it = 0, nb = 0, rc2 = PT_OK;
        const int batch = std::getenv("PTB200_BATCH") ? std::max(1, std::atoi(std::getenv("PTB200_BATCH"))) : 2;
        bool done = false;
        while (!done) {
            for (int b = 0; b < batch && rc2 == PT_OK; b++, it++) {
                if (it >= max_it) { rc2 = pt_fail(ctx, PT_ERR_STATE, "wavefront iteration limit reached"); break; }
                for (int a = 0; a < 4; a++) { P.qin[a] = ctx->q[it & 1][a]; P.qout[a] = ctx->q[(it + 1) & 1][a]; }
                P.n_in = n_it + it; P.n_out = n_it + it + 1;
                P.first_launch = it == 0 ? 1 : 0;
                P.launch_rec = ctx->d_launch_rec + std::min(it, PT_MAX_LAUNCH_RECS - 1);
                P.launch_rec0 = ctx->d_launch_rec;
                switch (p->mode) {
                case PT_MODE_NEE_REF_RECT: launch_bounce<PT_MODE_NEE_REF_RECT>(stats, glossy, blocks, s, P, jk, jit_block); break;
                case PT_MODE_COS: launch_bounce<PT_MODE_COS>(stats, glossy, blocks, s, P, jk, jit_block); break;
                case PT_MODE_UNI: launch_bounce<PT_MODE_UNI>(stats, glossy, blocks, s, P, jk, jit_block); break;
                default: launch_bounce<PT_MODE_NEE_CONE_SPHERE>(stats, glossy, blocks, s, P, jk, jit_block); break;
                }
                ctx->stats.kernel_launches++;
            }
            if (rc2 != PT_OK) break;
            cudaMemcpyAsync(h_n + (nb & 1), n_it + it, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
            cudaEventRecord(evb[nb & 1], s);
            if (nb > 0) {
                cudaEventSynchronize(evb[(nb - 1) & 1]);
                if (h_n[(nb - 1) & 1] == 0) done = true;
            }
            nb++;
            cudaError_t e_ = cudaGetLastError();
            if (e_ != cudaSuccess) { rc2 = pt_fail(ctx, PT_ERR_CUDA, std::string(jk ? "k_bounce (scene-specialised): " : "k_bounce: ") + cudaGetErrorString(e_)); break; }
        }
        if (rc2 != PT_OK) { cudaStreamSynchronize(s); return rc2; }
        ctx->stats.iterations = (uint64_t)it;
        ctx->counts_dirty = (size_t)it + 8;
        it_total = it;
    }
    if (!ctx->d_stamps) PT_CUDA(ctx, cudaMalloc(&ctx->d_stamps, 2 * sizeof(unsigned long long)));
    if (p->owned_rows_only) {          // (render_common rejects owned_rows_only together with collect_stats)
        if (owned_pixels > 0)
            k_resolve_owned<<<(unsigned)((owned_pixels * 3ull + 255) / 256), 256, 0, s>>>(ctx->d_fix, d_sum, owned_pixels, w, tile, p->rank, world, ctx->d_stamps);
        else k_stamp<<<1, 1, 0, s>>>(ctx->d_stamps);
    } else {
        k_resolve<<<(unsigned)((n_acc + 255) / 256), 256, 0, s>>>(ctx->d_fix, stats ? ctx->d_fixsq : nullptr, d_sum, stats ? d_sumsq : nullptr, n_acc, ctx->d_stamps);
    }
    k_stamp<<<1, 1, 0, s>>>(ctx->d_stamps + 1);
    ctx->stats.kernel_launches += 2;
    PT_CUDA(ctx, cudaGetLastError());
    ctx->stats.queue_slots_io = 0;
    ctx->stats.main_kernel_ms = ctx->stats.tail_ms = ctx->stats.resolve_ms = 0.0;
    ctx->stats.tail_launches = 0;
    if (it_total > 0) {   // queue traffic of this render: launch k reads n[k] slots and writes n[k+1]
        std::vector<unsigned int> h(it_total + 1);
        const int n_rec = std::min(it_total, PT_MAX_LAUNCH_RECS);
        std::vector<LaunchRec> rec(n_rec);
        unsigned long long stamps[2] = {0, 0};
        PT_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->d_counts + 4, sizeof(unsigned int) * (size_t)(it_total + 1), cudaMemcpyDeviceToHost, s));
        PT_CUDA(ctx, cudaMemcpyAsync(rec.data(), ctx->d_launch_rec, sizeof(LaunchRec) * (size_t)n_rec, cudaMemcpyDeviceToHost, s));
        PT_CUDA(ctx, cudaMemcpyAsync(stamps, ctx->d_stamps, sizeof stamps, cudaMemcpyDeviceToHost, s));
        PT_CUDA(ctx, cudaStreamSynchronize(s));
        uint64_t io = 0;
        for (int k = 0; k <= it_total; k++) io += (k == 0 || k == it_total) ? h[k] : 2ull * h[k];
        ctx->stats.queue_slots_io = io;
        if (const char *dump = std::getenv("PTB200_DUMP_LAUNCHES")) {      // tuning aid: one line per launch (start in us, phase, live slots)
            if (FILE *f = std::fopen(dump, "a")) {
                std::fprintf(f, "# render %dx%d spp %d mode %d: %d launches\n", w, h, p->spp, p->mode, it_total);
                for (int k = 0; k < n_rec; k++)
                    std::fprintf(f, "%d %.2f %u %u %u\n", k, (double)(rec[k].t_start - rec[0].t_start) * 1e-3, rec[k].exhausted, rec[k].n_in, h[k + 1]);
                std::fprintf(f, "resolve %.2f end %.2f\n", (double)(stamps[0] - rec[0].t_start) * 1e-3, (double)(stamps[1] - rec[0].t_start) * 1e-3);
                std::fclose(f);
            }
        }
        // phases: launches are back to back on one stream, so a launch ends where the next one (or the resolve) starts
        int first_tail = n_rec;
        for (int k = 0; k < n_rec; k++) if (rec[k].exhausted) { first_tail = k; break; }
        const unsigned long long t0 = rec[0].t_start, t_tail = first_tail < n_rec ? rec[first_tail].t_start : stamps[0];
        ctx->stats.main_kernel_ms = (double)(t_tail - t0) * 1e-6;
        ctx->stats.tail_ms = (double)(stamps[0] - t_tail) * 1e-6;
        ctx->stats.resolve_ms = (double)(stamps[1] - stamps[0]) * 1e-6;
        ctx->stats.tail_launches = (uint64_t)(it_total - std::min(first_tail, it_total));
    }
    return PT_OK;
}

int pt_fp32_intersect(pt_ctx *ctx, const double *d_rays, int n, double *d_t, int *d_id, cudaStream_t s)
{
    if (!ctx->fp32_ok) return pt_fail(ctx, PT_ERR_ARG, "scene does not fit the FP32 engine: " + ctx->fp32_why);
    if (ctx->jit_mode >= 2 && ctx->grid.n == 0) {       // specialisation forced: answer with the closest_hit of the scene-specialised build
        const PtJitKernel *jk = pt_jit_get(ctx, PT_MODE_COS, false, true);
        void *dptr = nullptr;
        size_t bytes = 0;
        if (jk && jk->kern_isect && cudaLibraryGetGlobal(&dptr, &bytes, jk->lib, "c_scene") == cudaSuccess && bytes == sizeof(SceneF32)) {
            PT_CUDA(ctx, cudaMemcpyAsync(dptr, ctx->h_scene32, sizeof(SceneF32), cudaMemcpyHostToDevice, s));
            const MatF32 *mats = ctx->d_mats;
            const float4 *sphf = ctx->d_sphf;
            void *args[] = {(void *)&d_rays, (void *)&n, (void *)&d_t, (void *)&d_id, (void *)&mats, (void *)&sphf};
            PT_CUDA(ctx, cudaLaunchKernel((const void *)jk->kern_isect, dim3((n + 127) / 128), dim3(128), args, 0, s));
            ctx->stats.specialised = 1;
            return PT_OK;
        }
        cudaGetLastError();
    }
    ctx->stats.specialised = 0;
    PT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_scene, ctx->h_scene32, sizeof(SceneF32), 0, cudaMemcpyHostToDevice, s));
    if (ctx->grid.n > 0) k_intersect_fp32<true><<<(n + 127) / 128, 128, 0, s>>>(d_rays, n, d_t, d_id, ctx->d_mats, ctx->d_sphf, ctx->grid);
    else k_intersect_fp32<false><<<(n + 127) / 128, 128, 0, s>>>(d_rays, n, d_t, d_id, ctx->d_mats, ctx->d_sphf, ctx->grid);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp32_philox(pt_ctx *ctx, const uint32_t *d_ctr, const uint32_t *d_key, int n, uint32_t *d_out, cudaStream_t s, int width)
{
    (void)width;
    k_philox<<<(n + 127) / 128, 128, 0, s>>>(d_ctr, d_key, n, d_out);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_fp32_ffma_peak(pt_ctx *ctx, double *tflops, double *mhz)
{
    const int blocks = ctx->sm_count * 8, iters = 4096;
    float *d_out = nullptr;
    PT_CUDA(ctx, cudaMalloc(&d_out, sizeof(float) * blocks * 256));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k_ffma_peak<<<blocks, 256, 0, ctx->stream>>>(d_out, iters, 1.0001f, 0.0001f);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, ctx->stream);
        k_ffma_peak<<<blocks, 256, 0, ctx->stream>>>(d_out, iters, 1.0001f, 0.0001f);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    PT_CUDA(ctx, cudaGetLastError());
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * 256;
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *mhz = khz / 1000.0;
    return PT_OK;
}
