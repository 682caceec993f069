#!/usr/bin/env python3
"""bench.py — headline benchmark: Mpaths/s of the per-pixel Monte Carlo radiance loop (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our arm (B200, CUDA kernels through the C ABI)
  python bench.py --impl reference [...]                    the reference's own CPU code on the host cores

One "step" = one full render of the workload.  The headline workload is the SAME at every N — BASELINE.json
configs[4] ("C5": built-in scene A at 3840x2160, 1024 spp, the reference's explicit light sampling + Russian roulette;
it fits one GPU at ~0.28 s per render), so the driver's 1/2/4/8-GPU values form a strong-scaling curve.  At N > 1 the
image is sharded by interleaved 10-row tiles and assembled on rank 0.
`value` = paths of the whole job / device time (max over ranks), inputs resident on the device.
`e2e`   = the same metric through the C ABI with HOST buffers: scene upload (H2D), render, read-back of the FP64
          image to host memory (D2H; at N > 1 every rank DMAs its own rows into one shared host image), wall clock.
`secondary` (N = 1): every other BASELINE config on scenes A and B, each with its own roofline.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpaths/s"
WORKLOADS = {
    # name: (scene, width, height, spp, mode, description)
    "c1": ("A", 512, 512, 16, 1, "C1: built-in scene A 512x512, 16 spp, cosine-weighted"),
    "c1b": ("B", 512, 512, 16, 1, "C1/B: sphere-era scene B 512x512, 16 spp, cosine-weighted"),
    "c2": ("A", 512, 512, 512, 0, "C2: built-in scene A 512x512, 512 spp, NEE_REF_RECT + Russian roulette"),
    "c2b": ("B", 512, 512, 512, 3, "C2/B: sphere-era scene B 512x512, 512 spp, NEE_CONE_SPHERE + Russian roulette"),
    "c3": ("A", 512, 512, 32, 2, "C3: built-in scene A 512x512, 32 spp, uniform hemisphere"),
    "c3b": ("B", 512, 512, 32, 2, "C3/B: sphere-era scene B 512x512, 32 spp, uniform hemisphere"),
    "c4": ("synthetic", 1920, 1080, 256, 1, "C4: synthetic 256 spheres + tilted planes 1920x1080, 256 spp, cosine"),
    "c5": ("A", 3840, 2160, 1024, 0, "C5: built-in scene A 3840x2160, 1024 spp, NEE_REF_RECT + Russian roulette"),
}
SECONDARY = ["c1", "c1b", "c2", "c2b", "c3", "c3b", "c4"]
F_SHADE = {0: 150.0, 1: 110.0, 2: 110.0, 3: 150.0}     # algorithmic FLOPs per shaded bounce, SURVEY 8(d)


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def traffic_per_launch(workload, launches_per_step):
    """DRAM bytes per k_bounce launch — a PROFILE CONSTANT, not measured in this run: dram__bytes_read + dram__bytes_write of
    the committed ncu --set full capture of a main launch of this workload (profiles/r02_traffic.json; round 1's file held
    whole-step figures).  None when no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_per_launch"], "profiles/r02_traffic.json (ncu --set full capture of a main launch; a constant, not measured in this run)"
    except Exception:                   # noqa: BLE001
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_per_step"] / max(1, launches_per_step), "profiles/r01_traffic.json (round-1 capture, whole step / launches; a constant)"
    except Exception:                   # noqa: BLE001
        return None, None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every ~5 ms
    (nvidia-smi -lms cannot resolve a region that lasts tens of milliseconds)."""

    def __init__(self, index):
        self.index = index
        self.samples = []          # (t, sm_mhz, reasons bitmask)
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through its UUID/PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:          # noqa: BLE001
            self.err = str(e)

    def start(self):
        if not self.ok:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:           # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self, t0=None, t1=None):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": getattr(self, "err", "")}
        self._stop.set()
        self.thread.join(timeout=1)
        rows = [r for r in self.samples if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        sm = [r[1] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


def cpu_reference(workload, budget_s, threads=0):
    """Time the reference's own CPU implementation (oracle/_ref/smallpt_ref: src/smallpt.cpp + patches P0-P6) on
    a bounded sample of the workload: same scene/mode/resolution, reduced spp.  Falls back to the C port."""
    scene, w, h, spp, mode, _ = WORKLOADS[workload]
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "smallpt_ref")
    scene_arg = {"A": "A", "B": "B", "C": "C"}.get(scene)
    ncores = os.cpu_count() or 1
    threads = threads or ncores

    def run_ref(s):
        out = subprocess.check_output([ref_bin, str(s), str(mode), scene_arg, str(w), str(h), "", "0", str(threads)], text=True)
        return json.loads(out.strip().splitlines()[-1])

    def run_port(s):
        from _pkg import ptb
        from oracle import pyoracle as orc      # the CPU checker: cpu_baseline / --impl reference only
        L = orc.load_oracle()
        L.oracle_set_threads(threads)
        sc = ptb.builtin_scene(scene, w, h)
        st = orc.oracle_render(sc, ptb.params(w, h, s, mode=mode, engine=1))[3]
        return {"paths": float(st.paths), "render_ms": st.render_ms, "threads": st.threads}

    kind = "reference" if (os.path.exists(ref_bin) and scene_arg) else "port"
    run = run_ref if kind == "reference" else run_port
    res = run(2)                                                     # probe, then grow the sample to the budget
    s = 2
    for _ in range(3):
        rate = res["paths"] / max(res["render_ms"], 1e-3)            # paths per ms
        s_next = int(max(1, min(spp, budget_s * 1e3 * rate / (w * h))))
        if res["render_ms"] >= 0.5 * budget_s * 1e3 or s_next <= s:
            break
        s = s_next
        res = run(s)
    return {"value": res["paths"] / res["render_ms"] * 1e-3, "unit": METRIC, "cores": int(res["threads"]), "kind": kind,
            "sample": f"{WORKLOADS[workload][5].split(':')[0]} scene/mode/resolution at {s} spp ({res['paths']:.0f} paths, "
                      f"{res['render_ms'] / 1e3:.1f} s), OpenMP schedule(dynamic,1) over rows, host has {ncores} cores"}, res


def secondary(ptb, workload, peak_tf, device_index, stream, flush):
    """One warm-up + two timed renders of another configuration, device-timed by the library's CUDA events.
    The scene-specialised kernel is forced (pt_set_specialisation 2): a render this small would otherwise start on the
    generic kernel and get its specialised build in the background."""
    scene_name, w, h, spp, mode, desc = WORKLOADS[workload]
    scene = ptb.builtin_scene(scene_name, w, h)
    with ptb.Context(scene, device=device_index) as c:
        c.set_specialisation(2)
        p = ptb.params(w, h, spp, mode=mode, engine=ptb.PT_ENGINE_FP32_PHILOX, seed=0)
        best = None
        for i in range(4):
            flush.fill_(1)
            stream.synchronize()
            c.render(p)
            st = c.stats()
            if i > 0 and (best is None or st.render_ms < best.render_ms):
                best = st
        flops = float(best.rays) * scene.flops_per_ray() + float(best.shaded_vertices) * F_SHADE[mode]
        tf = flops / (best.render_ms * 1e-3) / 1e12
        return {"workload": desc, "value": best.paths / best.render_ms * 1e-3, "unit": METRIC, "mrays_per_s": best.rays / best.render_ms * 1e-3,
                "ms_per_step": best.render_ms, "rays_per_path": best.rays / best.paths, "launches_per_step": int(best.iterations),
                "phases_ms": {"generating": best.main_kernel_ms, "tail": best.tail_ms, "resolve": best.resolve_ms, "tail_launches": int(best.tail_launches)},
                "max_depth": int(best.max_depth_seen), "specialised": bool(best.specialised),
                "roofline": {"bound": "fp32", "kernel": "k_bounce", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                             "flops_per_ray": scene.flops_per_ray(), "flops_per_bounce": F_SHADE[mode]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--assembly", default="auto", choices=["auto", "nccl"], help="N > 1: auto = peer-memory stores when CUDA IPC works, else NCCL gather")
    ap.add_argument("--no-ffma-peak", action="store_true", help="roofline against the theoretical FP32 peak (keeps the microbenchmark out of an ncu launch list)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs (N = 1) / the NCCL-gather and single-GPU legs (N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    workload = args.workload
    scene_name, w, h, spp, mode, desc = WORKLOADS[workload]
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        info = None
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        for i in range(args.warmup + args.steps):
            info, res = cpu_reference(workload, per_step)
            if i >= args.warmup:
                vals.append((res["paths"], res["render_ms"]))
        paths = sum(v[0] for v in vals)
        ms = sum(v[1] for v in vals)
        value = paths / ms * 1e-3
        info["value"] = value
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / max(1, args.steps), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "note": "CPU: each step is a bounded sample (reduced spp) of the workload"},
                "cpu_baseline": info,
                "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import numpy as np
    import torch
    from _pkg import ptb
    from small_pathtracer_b200 import dist as pdist
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    # 10-row tiles: 2160 rows = 216 tiles, the same number of tiles (and rows) for every rank at 2, 4 and 8 GPUs
    tile_rows = 10 if world > 1 else 8
    if os.environ.get("BENCH_TILE_ROWS"):                    # tuning aid
        tile_rows = int(os.environ["BENCH_TILE_ROWS"])
    scene = ptb.builtin_scene(scene_name, w, h)
    ctx = ptb.Context(scene, device=local_rank)
    params = ptb.params(w, h, spp, mode=mode, engine=ptb.PT_ENGINE_FP32_PHILOX, seed=0, tile_rows=tile_rows, rank=rank, world=world)
    stream = torch.cuda.Stream(device)        # the kernels, the gather and the timing events all live on this stream
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # N > 1: every rank's resolve kernel stores its row tiles straight into rank 0's image over NVLink peer memory
    # (dist.SharedImage); where CUDA IPC is unavailable, pack + NCCL gather + de-interleave (dist.gather_rows)
    shared, assembly = None, "single GPU"
    if world > 1:
        try:
            if args.assembly == "nccl":
                raise RuntimeError("NCCL gather requested")
            shared = pdist.SharedImage(ctx, h, w, rank, world, dst=0)
            assembly = "peer-memory stores from the resolve kernel (CUDA IPC over NVLink) + one barrier"
        except RuntimeError:
            shared, assembly = None, "pack + NCCL gather to rank 0 + de-interleave"

    def step(use_shared=True):
        """one render of this rank's tiles + assembly on rank 0; returns (device ms, stats, image, ms until this rank's own
        render had finished)"""
        flush.fill_(1)                                                    # L2 flush between iterations
        stream.synchronize()
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if shared is not None and use_shared:
            e0.record(stream)
            full = shared.render(params, stream.cuda_stream, barrier=False)
            em.record(stream)
            if world > 1:
                dist.barrier()
            st = ctx.stats()
        else:
            local = torch.empty((h, w, 3), dtype=torch.float64, device=device)
            e0.record(stream)
            ctx.render_into(params, local.data_ptr(), stream.cuda_stream)
            em.record(stream)
            st = ctx.stats()
            full = pdist.gather_rows(local, h, tile_rows, rank, world, dst=0) if world > 1 else local
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), st, full, e0.elapsed_time(em)     # render (all k_bounce launches + resolve) + assembly, on `stream`

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)
    barrier()
    t_wall0 = time.perf_counter()
    ms_steps, stats, launches, kernel_ms, iters = [], None, 0, 0.0, 0
    ph = {"generating": 0.0, "tail": 0.0, "resolve": 0.0, "own_render": 0.0, "wait_and_barrier": 0.0}
    for _ in range(args.steps):
        ms, st, full, ms_own = step()
        ms_steps.append(ms)
        stats = st
        launches += st.kernel_launches
        kernel_ms += st.render_ms
        iters += st.iterations
        ph["generating"] += st.main_kernel_ms
        ph["tail"] += st.tail_ms
        ph["resolve"] += st.resolve_ms
        ph["own_render"] += ms_own
        ph["wait_and_barrier"] += ms - ms_own
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    total_ms = sum(ms_steps)
    my_paths = float(stats.paths) * args.steps
    my_rays = float(stats.rays) * args.steps
    my_shaded = float(stats.shaded_vertices) * args.steps
    agg = torch.tensor([total_ms, my_paths, my_rays, my_shaded, float(launches), kernel_ms], dtype=torch.float64, device=device)
    ph_keys = sorted(ph)
    ph_t = torch.tensor([ph[k] / args.steps for k in ph_keys], dtype=torch.float64, device=device)
    ph_min = ph_t.clone()
    if world > 1:
        tmax = agg[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sums = agg[1:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(ph_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ph_min, op=dist.ReduceOp.MIN)
        total_ms = float(tmax[0])
        paths, rays, shaded, launches_all, kernel_ms_all = (float(x) for x in sums)
    else:
        paths, rays, shaded, launches_all, kernel_ms_all = my_paths, my_rays, my_shaded, float(launches), kernel_ms
    phases = {k: {"max_over_ranks_ms": float(ph_t[i]), "min_over_ranks_ms": float(ph_min[i])} for i, k in enumerate(ph_keys)}
    if world > 1:                       # this rank's own render, rank by rank (which GPUs are the slow ones)
        own = torch.zeros(world, dtype=torch.float64, device=device)
        own[rank] = ph["own_render"] / args.steps
        dist.all_reduce(own, op=dist.ReduceOp.SUM)
        phases["own_render_by_rank_ms"] = [round(float(x), 3) for x in own]
    phases["note"] = ("per step; generating / tail / resolve from %globaltimer stamps the kernels take themselves (pt_stats), own_render = CUDA events "
                      "around this rank's pt_render_into, wait_and_barrier = the rest of the step (waiting for the slowest rank + the barrier / gather)")
    value = paths / total_ms * 1e-3
    mrays = rays / total_ms * 1e-3

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    # N = 1: pt_scene_upload + pt_render + pt_readback_view (one DMA into the context's pinned image).
    # N > 1: every rank renders its tiles and DMAs ITS rows into ONE host image in shared memory (pt_readback_owned),
    #        each over its own PCIe link; rank 0's host then holds the whole FP64 picture.
    host_img = None
    params_own = ptb.params(w, h, spp, mode=mode, engine=ptb.PT_ENGINE_FP32_PHILOX, seed=0, tile_rows=tile_rows, rank=rank, world=world, owned_rows_only=1)
    if world > 1:
        try:
            host_img = pdist.HostImage(ctx, h, w, rank, world, dst=0)
        except Exception:               # noqa: BLE001
            host_img = None
    host_pinned = torch.empty((h, w, 3), dtype=torch.float64, pin_memory=True) if (world > 1 and rank == 0 and host_img is None) else None

    def e2e_step():
        t0 = time.perf_counter()
        ctx.update_scene(scene)                                             # H2D: scene table + camera (pt_scene_upload, in place)
        if world > 1 and host_img is not None:
            ctx.render(params_own)                                          # this rank's tiles into its own accumulators (only its rows are resolved)
            ctx.readback_owned(host_img.array)                              # D2H: its rows into the shared host image
            dist.barrier()
            host = host_img.array if rank == 0 else None
        elif world > 1:
            if shared is not None:
                full = shared.render(params, stream.cuda_stream)
            else:
                full, local = pdist.render_sharded(ctx, params, device, dst=0)
            host = None
            if rank == 0:                                                   # D2H: assembled image (sum -> mean on the device)
                host_pinned.copy_(full / float(spp), non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
                host = host_pinned.numpy()
            dist.barrier()
        else:
            ctx.render(params)
            host, _ = ctx.readback_view()                                   # D2H inside pt_readback_view (context-owned pinned image)
        return time.perf_counter() - t0, host
    _, img_e2e = e2e_step()
    e2e_check = None
    if rank == 0 and world > 1 and full is not None:                        # the host-assembled picture is the device-assembled one
        e2e_check = bool(np.array_equal(np.asarray(img_e2e), full.cpu().numpy() * (1.0 / float(spp))))      # (the library scales by the reciprocal too)
    barrier()
    e2e_times = []
    for _ in range(max(1, min(args.steps, 10))):
        dt, _img = e2e_step()
        e2e_times.append(dt)
    e2e_t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = (paths / args.steps) * len(e2e_times) / float(e2e_t[0]) * 1e-6

    # N > 1, same job: (a) the NCCL-gather assembly, (b) the whole workload on rank 0 alone (strong-scaling reference)
    nccl_leg, single_ms = None, None
    if world > 1 and not args.no_secondary:
        if shared is not None:
            t_n = []
            for i in range(3):
                barrier()
                ms, _st, _f, _o = step(use_shared=False)
                if i > 0:
                    t_n.append(ms)
            t_nccl = torch.tensor([sum(t_n) / len(t_n)], dtype=torch.float64, device=device)
            dist.all_reduce(t_nccl, op=dist.ReduceOp.MAX)
            nccl_leg = {"assembly": "pack + NCCL gather to rank 0 + de-interleave", "ms_per_step": float(t_nccl[0]),
                        "value": (paths / args.steps) / float(t_nccl[0]) * 1e-3, "unit": METRIC}
        if rank == 0:
            p1 = ptb.params(w, h, spp, mode=mode, engine=ptb.PT_ENGINE_FP32_PHILOX, seed=0, tile_rows=tile_rows, rank=0, world=1)
            for _ in range(2):
                flush.fill_(1)
                stream.synchronize()
                ctx.render(p1)
                single_ms = ctx.stats().render_ms
        dist.barrier()
    import ctypes as C
    h2d = scene.n_spheres * C.sizeof(ptb.Sphere) + scene.n_planes * C.sizeof(ptb.Plane) + 4 * scene.n_objects + C.sizeof(ptb.Camera) + C.sizeof(ptb.Light)
    d2h = w * h * 3 * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline (rank 0's dominant kernel: k_bounce)
    peaks, peak_src = read_peaks()
    try:
        ffma_tf = None if args.no_ffma_peak else ctx.ffma_peak()[0]
    except Exception:
        ffma_tf = None
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count          # 148 on B200
    fp32_theory = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
    f_isect = scene.flops_per_ray()
    my_flops = float(stats.rays) * f_isect + float(stats.shaded_vertices) * F_SHADE[mode]     # per step, this rank
    k_ms = stats.main_kernel_ms + stats.tail_ms                                              # the k_bounce launches of a step (%globaltimer stamps)
    if not k_ms > 0:
        k_ms = stats.render_ms
    achieved_tf = my_flops / (k_ms * 1e-3) / 1e12
    per_launch_ms = k_ms / max(1, stats.iterations)
    qbytes = 48.0 * float(stats.queue_slots_io)                       # 48 B per path record read or written through the queues
    traffic, traffic_src = traffic_per_launch(workload, int(stats.iterations))
    roofline = {"bound": "fp32", "kernel": "k_bounce", "achieved": achieved_tf, "peak": ffma_tf or fp32_theory, "unit": "TFLOP/s",
                "frac": achieved_tf / (ffma_tf or fp32_theory),
                "peak_source": "FFMA-only microbenchmark measured in this run (pt_debug_ffma_peak)" if ffma_tf else f"{n_sm} SM x 128 lanes x 2 x sm_max_mhz",
                "peak_theoretical": fp32_theory, "frac_of_theoretical": achieved_tf / fp32_theory,
                "flops_per_ray": f_isect, "flops_per_bounce": F_SHADE[mode],
                "avg_launch_ms": per_launch_ms, "launches_per_step": int(stats.iterations),
                "traffic": traffic, "traffic_source": traffic_src,
                "queue": {"bound": "hbm", "achieved": qbytes / (k_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                          "frac": qbytes / (k_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peak_src,
                          "note": "path records through the wavefront queues x 48 B (a slot advances several bounces per launch in registers)"}}
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": n_gpus, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "engine": "FP32 Philox wavefront", "tile_rows": tile_rows, "parallelism": f"row-tiles x{world}", "assembly": assembly,
                       "l2": "256 MiB flush write between timed iterations", "paths_per_step": paths / args.steps,
                       "rays_per_path": rays / paths, "seed": 0,
                       "kernel": ("scene-specialised k_bounce (NVRTC build with the scene constants as immediates, compiled once during warm-up)"
                                  if stats.specialised else "generic k_bounce (scene in __constant__ memory)"),
                       "bounces_per_launch": 512},
            "mrays_per_s": mrays, "wall_ms_per_step": (t_wall1 - t_wall0) * 1e3 / args.steps,
            "clocks": clocks, "gpu_launches": int(launches_all), "phases": phases,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": ("pt_scene_upload (scene tables host->device, context re-used) + pt_render + pt_readback_view (FP64 mean image device->pinned host) per step, wall clock"
                             if world == 1 else
                             ("pt_scene_upload + pt_render of this rank's tiles + pt_readback_owned: every rank DMAs its rows into ONE shared, page-locked host image "
                              "over its own PCIe link + one barrier, wall clock, max over ranks" if host_img is not None else
                              "pt_scene_upload + sharded render + assembled image read back through rank 0, wall clock"))},
            "roofline": roofline}
    if e2e_check is not None:
        line["e2e"]["host_image_equals_device_image"] = e2e_check
    if single_ms:
        line["strong_scaling"] = {"single_gpu_ms_per_step": single_ms, "n_gpu_ms_per_step": total_ms / args.steps,
                                  "speedup": single_ms / (total_ms / args.steps), "n_gpus": n_gpus,
                                  "note": "same workload rendered by rank 0 alone in this job (render only, no gather) vs the sharded step incl. gather"}
    if nccl_leg:
        line["nccl_assembly"] = nccl_leg
    if n_gpus == 1 and not args.no_secondary:
        # every other BASELINE config, scenes A and B: same engine, same accounting, each with its own roofline
        line["secondary"] = {}
        for name in SECONDARY:
            if name == workload:
                continue
            try:
                line["secondary"][name] = secondary(ptb, name, ffma_tf or fp32_theory, local_rank, stream, flush)
            except Exception as e:                                       # noqa: BLE001
                line["secondary"][name] = {"error": str(e)}
    if n_gpus == 1 and not args.no_cpu_baseline:
        try:
            info, _ = cpu_reference(workload, args.cpu_budget)
            line["cpu_baseline"] = info
        except Exception as e:                                           # the CPU leg must never sink the GPU number
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
