"""A/B timing of NVRTC build options on bench.py workloads (specialised kernel):  python tools/ab_jit_opts.py WORKLOAD "OPTS_A" "OPTS_B" ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    import bench
    scene, w, h, spp, mode, desc = bench.WORKLOADS[sys.argv[2]]
    with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
        c.set_specialisation(int(os.environ.get("AB_SPEC", "2")))
        best = None
        for _ in range(5):
            # (AB_CAP / AB_ITERS / AB_WORLD: queue capacity, bounces per launch, and rank 0's share of a world of that many GPUs)
            c.render(ptb.params(w, h, spp, mode=mode, seed=0, queue_capacity=int(os.environ.get("AB_CAP", "0")), bounces_per_launch=int(os.environ.get("AB_ITERS", "0")),
                                world=int(os.environ.get("AB_WORLD", "1")), tile_rows=8))
            st = c.stats()
            if best is None or st.render_ms < best[0]:
                best = (st.render_ms, st.paths / st.render_ms * 1e-3, st.rays / st.render_ms * 1e-3, st.specialised,
                        f"main {st.main_kernel_ms:.3f} tail {st.tail_ms:.3f} resolve {st.resolve_ms:.3f} launches {st.iterations} ({st.tail_launches} tail) misses/path {st.miss_events / st.paths:.5f}")
    print(json.dumps(best))
else:
    wl = sys.argv[1]
    for opts in sys.argv[2:]:
        env = dict(os.environ, PTB200_JIT_OPTS="" if opts in ("-", "generic") else opts, PTB200_CACHE_DIR="off", AB_SPEC="0" if opts == "generic" else "2")
        env.pop("PTB200_JIT_BLOCK", None)
        for tok in opts.split():
            if tok.startswith("-DPT_BLOCK="):
                env["PTB200_JIT_BLOCK"] = tok.split("=")[1]          # the host launches the module with the block size it was built for
        for rep in range(2):
            out = subprocess.check_output([sys.executable, __file__, "child", wl], env=env, text=True).strip().splitlines()[-1]
            ms, mp, mr, spec, phases = json.loads(out)
            print(f"{wl:4s} {opts:24s} {ms:9.3f} ms  {mp:9.1f} Mpaths/s  {mr:9.1f} Mrays/s  specialised={spec}  {phases}", flush=True)
