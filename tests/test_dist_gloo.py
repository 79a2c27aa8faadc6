"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo jobs exercise the slab partition, the broadcast of
the communicator id and the gather call sequence of parallel.SlabFilter with stand-in filters (no compute here: the compute
path and the NCCL hand-off exist only on the GPU -- tests/test_multi_gpu.py; slabs == whole plane is a -m gpu test)."""
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
    """stands in for DIGITAL_FILTER and for the library's communicator: field `which` at step t is the analytic plane
    which*1e6 + t*1e3 + j + k/1e4; the hand-off the library does with NCCL is done here with gloo point-to-point calls, with
    the library's call sequence (comm_unique_id on rank 0, comm_init everywhere, gather_begin / gather_end, gathered)."""

    def __init__(self, Ny, NzG, k0, k1):
        self.Ny, self.Nz, self.NzG, self.k0, self.k1, self.t = Ny, k1 - k0, NzG, k0, k1, 0

    def filter(self, dt):
        self.t += 1

    def field(self, which):
        j = np.arange(self.Ny)[:, None]
        k = np.arange(self.k0, self.k1)[None, :]
        return torch.from_numpy(which * 1e6 + self.t * 1e3 + j + k / 1e4)

    @staticmethod
    def comm_unique_id():
        return b"fake-id".ljust(128, b"\0")

    def comm_init(self, ident, rank, world):
        assert ident == b"fake-id".ljust(128, b"\0")
        self.rank, self.world = rank, world
        mine = torch.tensor([self.k0, self.k1])
        allb = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allb, mine)
        self.bounds = [(int(b[0]), int(b[1])) for b in allb]

    def comm_bounds(self):
        return self.bounds

    def gather_begin(self, dst):
        mine = torch.stack([self.field(w) for w in range(3)])            # u', v', w' on the wire
        self._reqs, self._parts = [], None
        if self.rank == dst:
            self._parts = [torch.empty((3, self.Ny, k1 - k0), dtype=torch.float64) for k0, k1 in self.bounds]
            self._parts[dst].copy_(mine)
            self._reqs = [dist.irecv(self._parts[r], r) for r in range(self.world) if r != dst]
        else:
            self._reqs = [dist.isend(mine, dst)]

    def gather_end(self):
        for r in self._reqs:
            r.wait()

    def gathered(self, which):
        from digital_filtering_b200 import parallel as P
        return P.assemble_plane([p[which].numpy() for p in self._parts], self.bounds, self.Ny, self.NzG)


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
            sf.gather_begin()
            sf.gather_end()
            for which in range(3):
                plane = sf.plane(which)
                if rank == 0:
                    j = np.arange(Ny)[:, None]
                    k = np.arange(NzG)[None, :]
                    ok &= bool(np.array_equal(plane, which * 1e6 + (step + 1) * 1e3 + j + k / 1e4))
                else:
                    ok &= plane is None
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
