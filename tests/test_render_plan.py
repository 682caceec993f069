"""CPU: the layout of an FP32 render (pt_fp32_plan through the host-only pt_debug_plan): which rows a rank owns, row blocks,
path slots in flight, sample runs and the layout flags a scene-specialised module is built for.  No GPU needed."""
import pytest

from conftest import ptb

WRAP_ONCE, MAGIC, WORLD1, ONE_BLOCK, RUNS = 1, 2, 4, 8, 16
SLOTS_B200 = 148 * 4 * 256 * 6         # six waves of resident 256-thread blocks


def plan(scene, w, h, spp, **kw):
    sm = kw.pop("sm_count", 148)
    return ptb.plan(ptb.builtin_scene(scene, w, h), ptb.params(w, h, spp, **kw), sm)


def test_headline_config_runs_in_row_blocks_and_sample_runs():
    # C5 (BASELINE.json configs[4]): 8.3 M pixels -> 16 blocks of 135 rows, every sample of a block before the next one;
    # 9340 paths per slot -> runs of 128 samples; one GPU: no tile arithmetic
    pl = plan("A", 3840, 2160, 1024, mode=0)
    assert (pl.owned_rows, pl.owned_pixels) == (2160, 3840 * 2160)
    assert (pl.row_blocks, pl.block_rows) == (16, 135) and pl.path_slots == SLOTS_B200
    assert pl.run_length == 128 and pl.path_indices == 3840 * 2160 * 8
    assert pl.layout_flags == WRAP_ONCE | MAGIC | WORLD1 | RUNS
    assert pl.splits_refr_paths == 0


@pytest.mark.parametrize("world,run,blocks", [(2, 128, 8), (4, 128, 4), (8, 1, 2)])
def test_shares_of_the_headline_config(world, run, blocks):
    # 10-row tiles, k % world (bench.py): every rank owns 2160 / world rows; an eighth is too small for sample runs
    seen = 0
    for rank in range(world):
        pl = plan("A", 3840, 2160, 1024, mode=0, tile_rows=10, rank=rank, world=world)
        assert pl.owned_rows == 2160 // world and pl.run_length == run and pl.row_blocks == blocks
        assert pl.layout_flags == WRAP_ONCE | MAGIC | (RUNS if run > 1 else 0)
        assert pl.path_indices == pl.row_blocks * pl.block_rows * 3840 * (1024 // run)
        assert pl.row_blocks * pl.block_rows >= pl.owned_rows > (pl.row_blocks - 1) * pl.block_rows
        seen += pl.owned_rows
    assert seen == 2160


def test_small_renders_single_samples_one_block_and_never_more_slots_than_paths():
    pl = plan("A", 512, 512, 512, mode=0)                      # C2: 147 paths per slot
    assert pl.run_length == 1 and pl.row_blocks == 1 and pl.layout_flags == WRAP_ONCE | MAGIC | WORLD1 | ONE_BLOCK
    assert pl.path_indices == 512 * 512 * 512 and pl.path_slots == SLOTS_B200
    pl = plan("A", 40, 21, 8, mode=1)                          # 6720 paths: 7 blocks of 1024 threads
    assert pl.path_slots == 7168 and pl.path_indices == 6720
    pl = plan("A", 1, 1, 5, mode=0)                            # one pixel: a lane's index may wrap more than once per step
    assert pl.layout_flags & WRAP_ONCE == 0 and pl.path_slots == 1024
    pl = plan("A", 64, 64, 0, mode=0)                          # nothing to trace
    assert pl.path_indices == 0 and pl.owned_pixels == 64 * 64


def test_ragged_tiles_and_ranks_without_rows():
    # 21 rows in 8-row tiles over 5 ranks: ranks 0-1 own 8 rows, rank 2 the ragged 5, ranks 3-4 nothing
    rows = [plan("A", 40, 21, 8, mode=1, tile_rows=8, rank=r, world=5).owned_rows for r in range(5)]
    assert rows == [8, 8, 5, 0, 0]
    assert plan("A", 40, 21, 8, mode=1, tile_rows=8, rank=4, world=5).path_indices == 0


def test_images_beyond_2_pow_24_pixels_divide_instead_of_multiply_shift():
    assert plan("A", 4096, 4095, 1, mode=1).layout_flags & MAGIC == MAGIC               # the multiply-shift is exact for n < 2^24
    assert plan("A", 4096, 4096, 1, mode=1).layout_flags & MAGIC == 0
    assert plan("A", 5000, 3403, 2, mode=1).layout_flags & MAGIC == 0
    pl = plan("A", 5000, 3403, 2, mode=1)
    assert pl.row_blocks == 33 and pl.row_blocks * pl.block_rows >= 3403               # last block short by a few rows: those indices are skipped
    # a third of it is below 2^24 owned pixels, but pixel indices of the whole image are not
    assert plan("A", 5000, 3403, 2, mode=1, tile_rows=7, rank=1, world=3).layout_flags & MAGIC == 0


def test_scenes_that_split_paths_at_glass_keep_single_samples():
    # REFR path splitting (:494-495): lanes take over spawned branches and have no run of their own to come back to
    pl = plan("G", 3840, 2160, 1024, mode=1)
    assert pl.splits_refr_paths == 1 and pl.run_length == 1 and pl.layout_flags & RUNS == 0
    pl = plan("G", 3840, 2160, 1024, mode=1, collect_stats=1)   # statistics renders take one arm: runs are fine
    assert pl.splits_refr_paths == 0 and pl.run_length == 128
    assert plan("B", 3840, 2160, 1024, mode=3).run_length == 128


def test_run_length_override_and_spp_limit(monkeypatch):
    monkeypatch.setenv("PTB200_RUN", "16")
    pl = plan("A", 96, 60, 44, mode=0)
    assert pl.run_length == 16 and pl.path_indices == 96 * 60 * 3                       # ceil(44 / 16) runs per pixel
    monkeypatch.setenv("PTB200_RUN", "64")
    assert plan("A", 96, 60, 44, mode=0).run_length == 32                               # never longer than spp
    monkeypatch.setenv("PTB200_RUN", "1")
    assert plan("A", 3840, 2160, 1024, mode=0).run_length == 1


def test_other_gpus_scale_the_slots():
    assert plan("A", 3840, 2160, 1024, mode=0, sm_count=132).path_slots == 132 * 4 * 256 * 6
    assert plan("A", 3840, 2160, 1024, mode=0, queue_capacity=5000).path_slots == 5120


def test_library_and_python_plumbing_agree_on_row_ownership():
    # tile k -> rank k % world: the library's count of owned rows equals dist.owned_rows (what gather_rows, HostImage and the
    # tests use to place a rank's rows) for ragged heights, odd tiles and ranks without rows
    from small_pathtracer_b200 import dist as pdist
    import random
    rng = random.Random(5)
    for _ in range(60):
        h, w = rng.randint(1, 400), rng.randint(1, 64)
        tile, world = rng.randint(1, 17), rng.randint(1, 9)
        total = 0
        for rank in range(world):
            pl = plan("A", w, h, 2, mode=1, tile_rows=tile, rank=rank, world=world)
            rows = pdist.owned_rows(h, tile, rank, world)
            assert pl.owned_rows == len(rows) and pl.owned_pixels == len(rows) * w
            if len(rows):
                assert all(((int(y) // tile) % world) == rank for y in rows)
            total += pl.owned_rows
        assert total == h


def test_bad_arguments():
    with pytest.raises(ptb.PtError):
        ptb.plan(ptb.builtin_scene("A", 8, 8), ptb.params(0, 8, 1), 148)
    with pytest.raises(ptb.PtError):
        ptb.plan(ptb.builtin_scene("A", 8, 8), ptb.params(8, 8, 1), 0)
