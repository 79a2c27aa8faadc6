"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo jobs exercise the slab
partition and the gather-to-the-CFD-rank collective with stand-in filters (no compute here: the
compute path exists only on the GPU; the slabs-equal-whole-plane property is a -m gpu test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeFilter:
    """stands in for DIGITAL_FILTER: field `which` at step t is the analytic plane which*1e6 + t*1e3 + j + k/1e4"""

    def __init__(self, Ny, NzG, k0, k1):
        self.Ny, self.Nz, self.k0, self.k1, self.t = Ny, k1 - k0, k0, k1, 0

    def filter(self, dt):
        self.t += 1

    def device_tensor(self, which):
        j = np.arange(self.Ny)[:, None]
        k = np.arange(self.k0, self.k1)[None, :]
        return torch.from_numpy(which * 1e6 + self.t * 1e3 + j + k / 1e4)


def _worker(rank, world, port, Ny, NzG, q):
    import sys
    sys.path.insert(0, ROOT)
    import _dfb_import  # noqa: F401
    from digital_filtering_b200 import parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sf = P.SlabFilter(dist, NzG, lambda k0, k1: _FakeFilter(Ny, NzG, k0, k1), dst=0)
        ok = True
        for step in range(2):
            sf.filter(1e-7)
            bufs = sf.gather([0, 3, 4], torch, "cpu")
            if rank == 0:
                for fi, which in enumerate([0, 3, 4]):
                    plane = sf.plane_on_dst(bufs, fi)
                    j = np.arange(Ny)[:, None]
                    k = np.arange(NzG)[None, :]
                    ok &= bool(np.array_equal(plane, which * 1e6 + (step + 1) * 1e3 + j + k / 1e4))
            else:
                ok &= bufs is None
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= float(t.item()) == float(world)
        q.put((rank, ok, sf.bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Ny,NzG", [(2, 12, 400), (3, 7, 51)])
def test_slab_gather_over_gloo(world, Ny, NzG):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Ny, NzG, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert len({tuple(b) for _, _, b in res}) == 1          # every rank derived the same partition


def test_slab_bounds_cover_the_plane():
    import _dfb_import  # noqa: F401
    from digital_filtering_b200 import parallel as P
    for Nz, w in ((8192, 8), (8192, 2), (400, 4), (51, 3), (2048, 8), (17, 16), (5, 5)):
        b = P.all_slab_bounds(Nz, w)
        assert b[0][0] == 0 and b[-1][1] == Nz
        assert all(a[1] == c[0] for a, c in zip(b[:-1], b[1:])) and all(k1 > k0 for k0, k1 in b)
    with pytest.raises(ValueError):
        P.slab_bounds(3, 4, 0)
