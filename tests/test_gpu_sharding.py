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
