"""Spanwise-slab partitioning of one inflow plane over the ranks of a torch.distributed job
(BASELINE.json config 4).

Each rank owns the columns [k0, k1) of the Nz-wide plane and runs its own DIGITAL_FILTER handle on
its own GPU.  Because the noise is addressed by the GLOBAL element index (include/dfb_rng_spec.h),
a rank regenerates the noise of the Nz_max columns either side of its slab and re-filters them in y
locally: the filter needs NO halo exchange, and the union of the slabs is bit-identical to the
single-GPU plane.  The only communication is the hand-off of the finished plane to the CFD rank:
one NCCL gather of the five fields (or an all-gather when every rank wants the plane).

torch.distributed is plumbing here (process group, NCCL); the data path is the library's own
device buffers, wrapped zero-copy.
"""
import numpy as np


def slab_bounds(Nz, world, rank, align=16):
    """Columns [k0, k1) of rank `rank`: near-equal slabs whose interior boundaries are multiples of
    `align` (keeps the device rows 128-byte aligned).  Every rank gets at least one column when
    Nz >= world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if Nz < world:
        raise ValueError("more ranks than columns")
    cuts = [0]
    for r in range(1, world):
        c = int(round(Nz * r / world / align)) * align
        c = min(max(c, cuts[-1] + 1), Nz - (world - r))
        cuts.append(c)
    cuts.append(Nz)
    return cuts[rank], cuts[rank + 1]


def all_slab_bounds(Nz, world, align=16):
    return [slab_bounds(Nz, world, r, align) for r in range(world)]


def assemble_plane(parts, bounds, Ny, Nz):
    """Row-major (Ny, Nz) plane from per-rank (Ny, k1-k0) slabs -- the consumer-side indexing of the
    gathered staging layout [rank][field][Ny][W_rank]."""
    out = np.empty((Ny, Nz), dtype=np.asarray(parts[0]).dtype)
    for p, (k0, k1) in zip(parts, bounds):
        out[:, k0:k1] = np.asarray(p).reshape(Ny, k1 - k0)
    return out


class SlabFilter:
    """One rank's share of a sharded plane + the gather to the CFD rank.

    dist:   an initialised torch.distributed module/process group (backend "nccl" on GPUs; the
            host-side logic is exercised with "gloo" on CPU in tests/test_dist_gloo.py)
    make_filter(k0, k1) -> object with .Ny, .Nz, .filter(dt), .device_tensor(which) (GPU) or
            .host_array(which) (tests)
    """

    def __init__(self, dist, Nz, make_filter, dst=0):
        self.dist = dist
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.NzG = Nz
        self.bounds = all_slab_bounds(Nz, self.world)
        self.k0, self.k1 = self.bounds[self.rank]
        self.filt = make_filter(self.k0, self.k1)
        self.Ny = self.filt.Ny

    def filter(self, dt):
        self.filt.filter(dt)

    def gather(self, fields, torch, device):
        """Gathers `fields` (list of `which` selectors) to rank dst.  Returns on dst a list (one per
        rank) of tensors [len(fields), Ny, W_rank]; elsewhere None.  One collective per call."""
        mine = torch.stack([self.filt.device_tensor(w) for w in fields])          # [F, Ny, W]
        if self.rank == self.dst:
            bufs = [torch.empty((len(fields), self.Ny, k1 - k0), dtype=mine.dtype, device=device) for k0, k1 in self.bounds]
        else:
            bufs = None
        # slabs may differ in width -> grouped point-to-point (the NCCL "gather" for ragged sizes)
        ops = []
        if self.rank == self.dst:
            bufs[self.dst].copy_(mine)
            for r in range(self.world):
                if r != self.dst:
                    ops.append(self.dist.P2POp(self.dist.irecv, bufs[r], r))
        else:
            ops.append(self.dist.P2POp(self.dist.isend, mine, self.dst))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        return bufs

    def gather_async(self, fields, torch, device):
        """The same gather, overlapped with the NEXT step: the fields are staged (device-to-device) on a communication stream as
        soon as the handle's stream has finished the step, the handle's stream only waits for that staging copy, and the NCCL
        transfer of the staged copy runs while the next filter(dt) computes.  Returns a ticket for gather_wait()."""
        hs = torch.cuda.ExternalStream(self.filt.stream(), device=device)
        if not hasattr(self, "_comm"):
            self._comm = torch.cuda.Stream(device=device)
        done = torch.cuda.Event()
        done.record(hs)                                              # step t is complete on the handle's stream
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(done)
            mine = torch.stack([self.filt.device_tensor(w) for w in fields])      # staged copy [F, Ny, W]
            staged = torch.cuda.Event()
            staged.record(self._comm)
            bufs, ops = None, []
            if self.rank == self.dst:
                bufs = [torch.empty((len(fields), self.Ny, k1 - k0), dtype=mine.dtype, device=device) for k0, k1 in self.bounds]
                bufs[self.dst].copy_(mine)
                for r in range(self.world):
                    if r != self.dst:
                        ops.append(self.dist.P2POp(self.dist.irecv, bufs[r], r))
            else:
                ops.append(self.dist.P2POp(self.dist.isend, mine, self.dst))
            reqs = self.dist.batch_isend_irecv(ops) if ops else []
        hs.wait_event(staged)                                        # step t+1 may overwrite the fields once they are staged
        return dict(reqs=reqs, bufs=bufs, keep=mine)

    def gather_wait(self, ticket, torch):
        """Completes a gather_async: the communication stream (and the caller's current stream) wait for the transfer."""
        with torch.cuda.stream(self._comm):
            for req in ticket["reqs"]:
                req.wait()
        torch.cuda.current_stream().wait_stream(self._comm)
        return ticket["bufs"]

    def plane_on_dst(self, bufs, field_index):
        """(Ny, Nz) numpy plane of one gathered field (dst rank only)."""
        return assemble_plane([b[field_index].cpu().numpy() for b in bufs], self.bounds, self.Ny, self.NzG)
