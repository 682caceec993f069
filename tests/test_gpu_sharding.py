"""GPU: pt_render_into() into a torch tensor + the gather helper (world emulated on one GPU: ranks rendered
one after the other, the gather's pack/de-interleave exercised through the world==1 path and by hand)."""
import numpy as np
import pytest

from conftest import ptb

pytestmark = pytest.mark.gpu


def test_render_into_torch_tensor_and_assemble():
    import torch
    from small_pathtracer_b200 import dist as pdist
    w, h, spp, tile, world = 64, 50, 16, 8, 4
    sc = ptb.builtin_scene("A", w, h)
    dev = torch.device("cuda:0")
    with ptb.Context(sc) as c:
        c.render(ptb.params(w, h, spp, mode=0, seed=11))
        mean = c.readback()[0]
        img, _ = pdist.render_sharded(c, ptb.params(w, h, spp, mode=0, seed=11, tile_rows=tile, rank=0, world=1), dev)
        out = torch.zeros((h, w, 3), dtype=torch.float64, device=dev)
        for r in range(world):
            p = ptb.params(w, h, spp, mode=0, seed=11, tile_rows=tile, rank=r, world=world)
            local = torch.empty((h, w, 3), dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            c.render_into(p, local.data_ptr(), torch.cuda.current_stream().cuda_stream)
            rows = torch.as_tensor(pdist.owned_rows(h, tile, r, world), device=dev)
            out.index_copy_(0, rows, local.index_select(0, rows))
            assert c.accum_ptr() == local.data_ptr()
    assert np.array_equal(out.cpu().numpy() / spp, mean)
    assert np.array_equal(img.cpu().numpy() / spp, mean)


def test_owned_rows_only_renders_assemble_in_one_buffer():
    # the fused resolve + gather, ranks emulated on one GPU: every rank's resolve kernel writes only its own row tiles
    # into ONE image (a whole cudaMalloc allocation, wrapped as a torch tensor without a copy)
    import torch
    from small_pathtracer_b200 import dist as pdist
    w, h, spp, tile, world = 72, 37, 16, 8, 3       # spp a power of two: sum / spp == sum * (1 / spp) exactly
    sc = ptb.builtin_scene("A", w, h)
    with ptb.Context(sc) as c:
        c.render(ptb.params(w, h, spp, mode=1, seed=6))
        mean = c.readback()[0]
        ptr = c.device_alloc(h * w * 3 * 8)
        img = torch.as_tensor(pdist._CudaArray(ptr, (h, w, 3)), device=torch.device("cuda:0"))
        img.fill_(-1.0)                                       # rows nobody owns would stay -1
        torch.cuda.synchronize()
        for r in range(world):
            c.render_into(ptb.params(w, h, spp, mode=1, seed=6, tile_rows=tile, rank=r, world=world, owned_rows_only=1), ptr, 0)
        got = img.cpu().numpy() / spp
        del img
        c.device_free(ptr)
    assert np.array_equal(got, mean)


def _ipc_worker(rank, world, port, w, h, spp, tile, out_path):
    import os, sys
    import torch
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from _pkg import ptb as P
    from small_pathtracer_b200 import dist as pdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)                                  # both processes share the one GPU of the test box
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = P.builtin_scene("A", w, h)
    with P.Context(sc, device=0) as c:
        shared = pdist.SharedImage(c, h, w, rank, world, dst=0)
        for rep in range(2):                                  # the image is re-used across renders
            full = shared.render(P.params(w, h, spp, mode=0, seed=9 + rep, tile_rows=tile, rank=rank, world=world))
            if rank == 0:
                np.save(out_path + str(rep) + ".npy", full.cpu().numpy())
            dist.barrier()
        shared.close()
    dist.destroy_process_group()


def test_two_processes_assemble_through_cuda_ipc(tmp_path):
    # two processes, one GPU: rank 1 opens rank 0's image through its IPC handle and its resolve kernel stores into it
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    w, h, spp, tile = 64, 40, 8, 8
    out = str(tmp_path / "img")
    mp.spawn(_ipc_worker, args=(2, port, w, h, spp, tile, out), nprocs=2, join=True)
    sc = ptb.builtin_scene("A", w, h)
    with ptb.Context(sc) as c:
        for rep in range(2):
            c.render(ptb.params(w, h, spp, mode=0, seed=9 + rep))
            assert np.array_equal(np.load(out + str(rep) + ".npy") / spp, c.readback()[0])


def _host_worker(rank, world, port, w, h, spp, tile, out_path):
    import os, sys
    import torch
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from _pkg import ptb as P
    from small_pathtracer_b200 import dist as pdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = P.builtin_scene("A", w, h)
    with P.Context(sc, device=0) as c:
        host = pdist.HostImage(c, h, w, rank, world, dst=0)
        for rep in range(2):
            c.render(P.params(w, h, spp, mode=0, seed=30 + rep, tile_rows=tile, rank=rank, world=world))
            c.readback_owned(host.array)              # this rank's rows, DMA'd into the image every rank maps
            dist.barrier()
            if rank == 0:
                np.save(out_path + str(rep) + ".npy", np.array(host.array))
            dist.barrier()
        host.close()
    dist.destroy_process_group()


def test_two_processes_assemble_in_shared_host_memory(tmp_path):
    # the end-to-end path of a multi-GPU render: every rank copies its own row tiles device -> host into ONE image in POSIX
    # shared memory (pt_readback_owned); nothing funnels through rank 0's GPU.  7-row tiles over 45 rows: ragged last tile.
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    w, h, spp, tile = 64, 45, 8, 7
    out = str(tmp_path / "host")
    mp.spawn(_host_worker, args=(2, port, w, h, spp, tile, out), nprocs=2, join=True)
    sc = ptb.builtin_scene("A", w, h)
    with ptb.Context(sc) as c:
        for rep in range(2):
            c.render(ptb.params(w, h, spp, mode=0, seed=30 + rep))
            assert np.array_equal(np.load(out + str(rep) + ".npy"), c.readback()[0])


def test_single_process_multi_gpu_render():
    # pt_render_multi: one context per device, resolve kernels store into device 0's image over peer memory
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    w, h, spp = 96, 50, 16
    sc = ptb.builtin_scene("A", w, h)
    ctxs = [ptb.Context(sc, device=d) for d in range(n)]
    try:
        ctxs[0].render(ptb.params(w, h, spp, mode=0, seed=21))
        single, st1 = ctxs[0].readback()
        ptb.render_multi(ctxs, ptb.params(w, h, spp, mode=0, seed=21, tile_rows=4))
        multi, stn = ctxs[0].readback()
        assert np.array_equal(single, multi)
        assert stn.paths == st1.paths == w * h * spp and stn.rays == st1.rays
    finally:
        for c in ctxs:
            c.close()
