"""Row-tile sharding across GPUs + gather to rank 0.

One process per GPU (torchrun); torch.distributed is the plumbing.  Tile k (tile_rows rows) belongs to
rank k % world — interleaved so that ceiling/light rows (short paths) and floor rows (deep paths) spread
evenly.  Each rank renders its tiles with pt_render_into() straight into a torch tensor, packs its owned
rows and one gather (NCCL send/recv over NVLink; gloo in the CPU tests) brings them to rank 0, which
de-interleaves.  No reduction is involved: every pixel lives on exactly one rank, so the N-GPU image is
bit-identical to the 1-GPU image.
"""
import numpy as np


def owned_rows(height, tile_rows, rank, world):
    """Row indices owned by `rank` (ascending)."""
    tile_rows = tile_rows if tile_rows > 0 else 8
    rows = np.arange(height)
    return rows[(rows // tile_rows) % world == rank]


def gather_rows(local_full, height, tile_rows, rank, world, dst=0, group=None):
    """local_full: (H, W, 3) tensor holding this rank's owned rows (other rows ignored).
    Returns the assembled (H, W, 3) tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    tile_rows = tile_rows if tile_rows > 0 else 8
    counts = [len(owned_rows(height, tile_rows, r, world)) for r in range(world)]
    max_rows = max(counts)
    mine = torch.as_tensor(owned_rows(height, tile_rows, rank, world), device=local_full.device, dtype=torch.long)
    packed = local_full.new_zeros((max_rows,) + tuple(local_full.shape[1:]))
    if len(mine):
        packed[:len(mine)] = local_full.index_select(0, mine)
    if world == 1:
        parts = [packed]
    else:
        parts = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
        dist.gather(packed, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.zeros_like(local_full)
    for r in range(world):
        idx = torch.as_tensor(owned_rows(height, tile_rows, r, world), device=local_full.device, dtype=torch.long)
        if len(idx):
            out.index_copy_(0, idx, parts[r][:len(idx)])
    return out


def render_sharded(ctx, params, device, dst=0, group=None):
    """Render this rank's row tiles on `device` into a torch tensor (per-pixel SUM of sample radiance,
    FP64) and gather to rank `dst`.  Returns (image_sum on dst | None, local tensor)."""
    import torch
    # torch.empty: the library zeroes the buffer itself on the render stream (no torch kernel to race with)
    local = torch.empty((params.height, params.width, 3), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)
    stream.synchronize()
    # stream handle 0 (torch's legacy default stream) makes the library use its own non-blocking stream;
    # pt_render_into is synchronous on return either way
    ctx.render_into(params, local.data_ptr(), stream.cuda_stream)
    world = params.world if params.world > 0 else 1
    full = gather_rows(local, params.height, params.tile_rows, params.rank, world, dst=dst, group=group)
    return full, local


class _CudaArray:
    """A raw device pointer presented through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape, owner=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
        self._owner = owner


class SharedImage:
    """Fused resolve + gather over NVLink peer memory (one process per GPU on one node).

    Rank `dst` owns ONE (H, W, 3) float64 image in device memory (a whole cudaMalloc allocation, so that it has a CUDA IPC
    handle); the other ranks open it and `pt_render_into(..., owned_rows_only=1)` makes every rank's resolve kernel store
    its own row tiles straight into that image — there is no pack, no collective and no de-interleave.  The only
    synchronisation is one barrier after the renders.  Falls back (raises) where CUDA IPC is not available; callers then
    use `gather_rows`."""

    def __init__(self, ctx, height, width, rank, world, dst=0, group=None, buffers=2):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.dst, self.group = ctx, rank, world, dst, group
        self.shape = (height, width, 3)
        nbytes = height * width * 3 * 8
        # Two images, used alternately: after the barrier of render k the other ranks may already be storing render k+1
        # (into the OTHER image) while dst still reads image k; nobody can reach render k+2 before dst has entered the
        # barrier of render k+1, i.e. after it has let go of image k.  One image would be a write-after-read hazard.
        self.n_buf = max(1, int(buffers))
        self.ptrs = [None] * self.n_buf
        self.opened = False
        self.turn = 0
        err = 0
        handles = bytearray(64 * self.n_buf)
        if rank == dst:
            try:
                for b in range(self.n_buf):
                    self.ptrs[b] = ctx.device_alloc(nbytes)
                    handles[64 * b:64 * b + 64] = ctx.ipc_export(self.ptrs[b])
            except Exception:               # noqa: BLE001
                err = 1
        if world > 1:
            # the 64-byte handles travel as a tensor over whatever backend the group has (NCCL needs device tensors)
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.tensor(list(handles) + [err], dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=dst, group=group)
            t = t.cpu()
            err = int(t[64 * self.n_buf])
            if rank != dst and not err:
                try:
                    for b in range(self.n_buf):
                        self.ptrs[b] = ctx.ipc_open(bytes(t[64 * b:64 * b + 64].tolist()))
                    self.opened = True
                except Exception:           # noqa: BLE001
                    err = 1
            flag = torch.tensor([err], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            err = int(flag.item())
        if err:
            self.close()
            raise RuntimeError("CUDA IPC is not available between these processes")
        cur = torch.device("cuda", torch.cuda.current_device())
        self.tensors = [torch.as_tensor(_CudaArray(p, self.shape, owner=self), device=cur) for p in self.ptrs] if rank == dst else None

    @property
    def ptr(self):
        return self.ptrs[self.turn]

    def render(self, params, stream=0, barrier=True):
        """Every rank renders its row tiles into the shared image of this turn; returns the assembled tensor (per-pixel
        sums) on dst.  The tensor stays valid until the render after the next one (see __init__)."""
        import torch.distributed as dist
        self.turn = (self.turn + 1) % self.n_buf
        params.owned_rows_only = 1
        try:
            self.ctx.render_into(params, self.ptrs[self.turn], stream)       # synchronous on return: this rank's stores have landed
        finally:
            params.owned_rows_only = 0
        if self.world > 1 and barrier:
            dist.barrier(group=self.group)
        return self.tensors[self.turn] if self.tensors is not None else None

    def close(self):
        try:
            for b, p in enumerate(self.ptrs):
                if p is None:
                    continue
                if self.opened:
                    self.ctx.ipc_close(p)
                elif self.rank == self.dst:
                    self.tensors = None
                    self.ctx.device_free(p)
                self.ptrs[b] = None
        finally:
            self.ptrs = [None] * self.n_buf


class HostImage:
    """One (H, W, 3) float64 image in POSIX shared memory, mapped by every rank of the node and page-locked for DMA.

    The end-to-end path of a multi-GPU render: each rank copies ITS row tiles device -> host over its own PCIe link
    (Context.readback_owned) and the picture of src/smallpt.cpp:538 assembles in host memory; nothing funnels through
    rank 0's GPU.  The only synchronisation is the caller's barrier after the copies."""

    def __init__(self, ctx, height, width, rank, world, dst=0, group=None):
        import os
        import torch.distributed as dist
        self.ctx, self.rank, self.dst = ctx, rank, dst
        self.shape = (height, width, 3)
        nbytes = height * width * 3 * 8
        name = [None]
        if rank == dst:
            name[0] = f"/dev/shm/ptb200_img_{os.getpid()}_{id(self) & 0xFFFFFF:x}"
            with open(name[0], "wb") as f:
                f.truncate(nbytes)
        if world > 1:
            dist.broadcast_object_list(name, src=dst, group=group)
        self.path = name[0]
        self.array = np.memmap(self.path, dtype=np.float64, mode="r+", shape=self.shape)
        self.registered = False
        try:
            ctx.host_register(self.array)
            self.registered = True
        except Exception:               # noqa: BLE001   (pageable memory still works, through the driver's staging buffers)
            pass
        if world > 1:
            dist.barrier(group=group)
        if rank == dst:
            os.unlink(self.path)        # the mappings keep the segment alive; nothing is left behind in /dev/shm

    def close(self):
        if self.array is not None:
            if self.registered:
                try:
                    self.ctx.host_unregister(self.array)
                except Exception:       # noqa: BLE001
                    pass
            self.array = None
