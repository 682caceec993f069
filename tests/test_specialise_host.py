"""CPU: the scene-specialisation front end (pt_jit.cu).  NVRTC needs no GPU: the header generated for a scene must
carry the scene's constants and compile to an sm_100a cubin; a scene that differs in one constant gives another header."""
import re

import pytest

from conftest import ptb


def _have_nvrtc():
    try:
        ptb.specialise(ptb.builtin_scene("C", 64, 64), 1)
        return True
    except ptb.PtError as e:
        return "libnvrtc not found" not in str(e) and pytest.fail(str(e))


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
@pytest.mark.parametrize("scene,mode", [("A", 0), ("B", 1), ("synthetic", 2), ("B", 3)])
def test_specialised_source_compiles_for_sm100a(scene, mode):
    spec, cubin_bytes, seconds = ptb.specialise(ptb.builtin_scene(scene, 128, 96), mode)
    assert f"#define PT_J_MODE {mode}" in spec
    assert cubin_bytes > 10000
    assert seconds < 60


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
def test_cone_sampling_over_a_large_sphere_table_calls_closest_hit():
    # many sphere lights x an unrolled 200-sphere scan: inlined three times the kernel spills kilobytes per thread
    sc = ptb.builtin_scene("synthetic", 64, 64)
    assert "#define PT_NOINLINE_HIT 1" in ptb.specialise(sc, 3)[0]
    assert "PT_NOINLINE_HIT" not in ptb.specialise(sc, 1)[0]
    assert "PT_NOINLINE_HIT" not in ptb.specialise(ptb.builtin_scene("B", 64, 64), 3)[0]


def test_header_carries_the_scene_constants():
    sc = ptb.builtin_scene("A", 64, 64)
    spec, _, _ = ptb.specialise(sc, 0)
    # scene A (src/smallpt.cpp:287-311): 5 XZ, 6 XY and 6 YZ rectangles, no spheres, light = object 6 (the third XZ slot)
    assert "constexpr int PT_J_NSLOT[3] = {5, 6, 6};" in spec
    assert "#define PT_J_n_sph4 0" in spec and "#define PT_J_n_tilt 0" in spec
    assert "#define PT_J_light_code 2" in spec
    # k = 81.6 and 81.5 as exact FP32 hex floats, the ceiling and the light
    assert float.fromhex("0x1.466666p+6") == pytest.approx(81.6, rel=1e-7) and "0x1.466666p+6f" in spec and "0x1.46p+6f" in spec
    # light sampling constants of :365-367 (32 + 36 xi, 63 + 36 xi, y = 81.6, A = 1296)
    assert re.search(r"#define PT_J_lxw \(0x1\.2p\+5f\)", spec) and re.search(r"#define PT_J_larea \(0x1\.44p\+10f\)", spec)
    other, _, _ = ptb.specialise(ptb.builtin_scene("C", 64, 64), 0)
    assert other != spec


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
def test_module_for_a_render_layout_and_shadow_rays_without_slot_codes():
    # a module is built for the render's layout flags (pt_plan_info.layout_flags): the regeneration step then tests nothing
    # of it at run time.  C5's layout: blocks of >= 32 pixels, multiply-shift divisions, one GPU, row blocks, sample runs.
    sc = ptb.builtin_scene("A", 3840, 2160)
    flags = ptb.plan(sc, ptb.params(3840, 2160, 1024, mode=0)).layout_flags
    spec, cubin_bytes, _ = ptb.specialise(sc, 0, flags)
    for line in ("#define PT_BAKE_WORLD1 1", "#define PT_BAKE_RUNS 1", "#define PT_BAKE_WRAP_ONCE 1", "#define PT_BAKE_MAGIC 1"):
        assert line in spec
    assert "PT_NO_ROW_BLOCKS" not in spec and cubin_bytes > 10000
    small = ptb.specialise(sc, 0, ptb.plan(sc, ptb.params(512, 512, 512, mode=0)).layout_flags)[0]
    assert "#define PT_BAKE_RUNS 0" in small and "#define PT_NO_ROW_BLOCKS 1" in small
    eighth = ptb.specialise(sc, 0, ptb.plan(sc, ptb.params(3840, 2160, 1024, mode=0, tile_rows=10, rank=3, world=8)).layout_flags)[0]
    assert "#define PT_BAKE_WORLD1 0" in eighth and "#define PT_BAKE_RUNS 0" in eighth
    assert "PT_BAKE_" not in ptb.specialise(sc, 0)[0]                      # no layout given: everything stays a run-time test
    # shadow rays toward the rectangular light compete without slot codes only when no other rectangle comes near the light's
    # (scene A: the ceiling is 0.1 away, plenty) and only in the reference's NEE mode
    assert "#define PT_J_SHADOW_RAW 1" in spec and "PT_J_SHADOW_RAW" not in ptb.specialise(sc, 1)[0]
    near = ptb.builtin_scene("A", 64, 64)
    rects = list(near.planes)
    rects.append(ptb.rect(ptb.PT_PLANE_XZ, 40, 60, 70, 90, 81.5005, c=(.5, .5, .5)))     # a rectangle 5e-4 above the light
    crowded = ptb.Scene([], rects, list(range(len(rects))), near.light, near.camera)
    assert "PT_J_SHADOW_RAW" not in ptb.specialise(crowded, 0)[0]


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
def test_every_layout_flag_combination_compiles():
    # a module that failed to build would silently leave the render on the generic kernel: every combination of the five
    # layout flags must compile for sm_100a (scene A in the reference's NEE mode; the glass scene and the 256-sphere scene with
    # cone sampling - the lockstep / called-closest_hit variants - for the combinations a render can actually have)
    sc = ptb.builtin_scene("A", 64, 64)
    for flags in range(32):
        spec, cubin_bytes, _ = ptb.specialise(sc, 0, flags)
        assert cubin_bytes > 10000, flags
        assert ("#define PT_BAKE_RUNS 1" in spec) == bool(flags & 16) and ("#define PT_NO_ROW_BLOCKS 1" in spec) == bool(flags & 8)
    for scene, mode in (("G", 1), ("synthetic", 3), ("B", 3)):
        for flags in (1 | 2 | 4 | 8, 1 | 2, 1 | 2 | 4 | 16, 4):
            assert ptb.specialise(ptb.builtin_scene(scene, 64, 64), mode, flags)[1] > 10000, (scene, mode, flags)


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
def test_lockstep_modules_use_512_thread_blocks():
    # long immediate sphere tables are instruction-fetch bound: block-wide lockstep, 512 threads per block
    spec = ptb.specialise(ptb.builtin_scene("synthetic", 64, 64), 1)[0]
    assert "#define PT_LOCKSTEP 1" in spec and "#define PT_BLOCK 512" in spec
    other = ptb.specialise(ptb.builtin_scene("A", 64, 64), 1)[0]
    assert "PT_LOCKSTEP" not in other and "#define PT_BLOCK 1024" in other          # everything else: one block of 1024 threads per SM


@pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")
def test_disk_cache_serves_the_second_process(tmp_path):
    import os, subprocess, sys
    from conftest import ROOT
    code = ("import sys; sys.path.insert(0, %r); from _pkg import ptb; "
            "print(ptb.specialise(ptb.builtin_scene('C', 64, 64), 1)[1:])" % ROOT)
    env = dict(os.environ, PTB200_CACHE_DIR=str(tmp_path / "cache"))
    first = eval(subprocess.check_output([sys.executable, "-c", code], env=env, text=True))
    files = os.listdir(tmp_path / "cache")
    assert len(files) == 1 and files[0].startswith("ptb200-") and files[0].endswith(".cubin")
    second = eval(subprocess.check_output([sys.executable, "-c", code], env=env, text=True))
    assert first[0] == second[0] and first[1] > 0.05 and second[1] == 0.0      # same cubin, no NVRTC the second time
    assert open(tmp_path / "cache" / files[0], "rb").read(4) == b"\x7fELF"
