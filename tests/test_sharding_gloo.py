"""CPU, world_size 2 over gloo: the N>1 host logic (row-tile ownership, pack, gather, de-interleave).
The per-rank renderer here is the CPU oracle honouring the same rank/world/tile_rows parameters as
pt_render_into; the GPU variant of this test is tests/test_gpu_sharding.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ptb, orc, ROOT


def test_owned_rows_partition():
    from small_pathtracer_b200 import dist as pdist
    for h, tile, world in ((512, 8, 2), (2160, 16, 8), (37, 8, 4), (5, 8, 2), (64, 1, 3)):
        rows = [pdist.owned_rows(h, tile, r, world) for r in range(world)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange(h))       # disjoint cover
        if h >= tile * world * 4:
            assert max(len(r) for r in rows) - min(len(r) for r in rows) <= tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, w, h, tile, out_path):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import ptb as P
    from oracle import pyoracle as orc
    from small_pathtracer_b200 import dist as pdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.load_oracle().oracle_set_threads(1)
    sc = P.builtin_scene("A", w, h)
    p = P.params(w, h, 4, mode=0, engine=1, tile_rows=tile, rank=rank, world=world)
    mean = orc.oracle_render(sc, p)[1]                     # foreign rows stay zero
    local = torch.from_numpy(mean * 4.0)
    full = pdist.gather_rows(local, h, tile, rank, world, dst=0)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("h,tile", [(37, 8), (32, 4)])
def test_two_rank_gather_is_bit_identical(tmp_path, h, tile):
    w = 24
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, tile, out), nprocs=2, join=True)
    got = np.load(out)
    sc = ptb.builtin_scene("A", w, h)
    want = orc.oracle_render(sc, ptb.params(w, h, 4, mode=0, engine=1))[1] * 4.0
    assert np.array_equal(got, want)
