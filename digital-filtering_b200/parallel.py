"""Spanwise-slab partitioning of one inflow plane over the ranks of a job (BASELINE.json config 4) -- the thin
Python caller of the library's own communicator (dfb_comm_init / dfb_gather_begin / dfb_gather_end in include/dfb200.h).

Each rank owns the columns [k0, k1) of the Nz-wide plane and runs its own DIGITAL_FILTER handle on its own GPU.  Because the
noise is addressed by the GLOBAL element index (include/dfb_rng_spec.h), a rank regenerates the noise of the Nz_max columns
either side of its slab and re-filters them in y locally: the filter needs NO halo exchange, and the union of the slabs is
bit-identical to the single-GPU plane.  The only communication is the hand-off of the finished plane to the CFD rank, done
INSIDE libdfb200.so with NCCL (u', v', w' on the wire, T' and rho' rebuilt on the destination).  torch.distributed only carries
the 128-byte NCCL id from rank 0 to the others, exactly as MPI_Bcast would in a C++/Fortran CFD code.
"""
import numpy as np


def slab_bounds(Nz, world, rank, align=16):
    """Columns [k0, k1) of rank `rank`: near-equal slabs whose interior boundaries are multiples of
    `align` (keeps the device rows 128-byte aligned and the lane blocks of the z-sweep on the plane's grid).
    Every rank gets at least one column when Nz >= world."""
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
    """Row-major (Ny, Nz) plane from per-rank (Ny, k1-k0) slabs (host-side check of the device assembly)."""
    out = np.empty((Ny, Nz), dtype=np.asarray(parts[0]).dtype)
    for p, (k0, k1) in zip(parts, bounds):
        out[:, k0:k1] = np.asarray(p).reshape(Ny, k1 - k0)
    return out


class SlabFilter:
    """One rank's share of a sharded plane + the hand-off to the CFD rank.

    dist:   an initialised torch.distributed module (any backend: it only broadcasts the communicator id)
    make_filter(k0, k1) -> DIGITAL_FILTER for the slab (or a stand-in with the same comm_* / gather_* methods: the CPU tests)
    """

    def __init__(self, dist, Nz, make_filter, dst=0):
        self.dist = dist
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.NzG = Nz
        self.bounds = all_slab_bounds(Nz, self.world)
        self.k0, self.k1 = self.bounds[self.rank]
        self.filt = make_filter(self.k0, self.k1)
        self.Ny = self.filt.Ny
        ident = [self.filt.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        self.filt.comm_init(ident[0], self.rank, self.world)
        if list(self.filt.comm_bounds()) != [tuple(b) for b in self.bounds]:
            raise RuntimeError("the ranks disagree on the slab partition")

    def filter(self, dt):
        self.filt.filter(dt)

    def gather_begin(self):
        """enqueue the hand-off of the step just enqueued; the next filter(dt) may follow at once (it overlaps the transfer)"""
        self.filt.gather_begin(self.dst)

    def gather_end(self):
        self.filt.gather_end()

    def gather(self):
        self.gather_begin()
        self.gather_end()

    def plane(self, which):
        """destination rank: the gathered (Ny, Nz) field as a numpy array; elsewhere None"""
        return self.filt.gathered(which) if self.rank == self.dst else None
