"""tests/slab_worker.py -- one rank of the 2-GPU slab test (launched by tests/test_multi_gpu.py through torch.distributed.run):
parallel.SlabFilter (library-side NCCL hand-off) on a small plane; rank 0 checks every gathered field of every step, bit for bit,
against the same plane filtered whole on its own GPU.  The gather of step t overlaps step t+1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _dfb_import  # noqa: E402,F401
import digital_filtering_b200 as dfb  # noqa: E402
from digital_filtering_b200 import parallel as P, workloads as W  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")              # carries only the 128-byte communicator id: the data path is the library's NCCL
    plane = W.plane_profile(64, 1400, 16, 24)
    mk = lambda k0, k1: dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=3, device=local, k_begin=k0, k_end=k1), fetch=False)
    sf = P.SlabFilter(dist, plane["Nz"], mk)
    dts = [2e-7, 1e-7, 3e-7]
    got = []
    sf.gather()                                   # the constructor's step: T', rho' are zero (df.cpp:57-65)
    got.append([sf.plane(w) for w in range(5)])
    sf.filter(dts[0])
    sf.gather_begin()
    for s in range(3):
        if s + 1 < 3:
            sf.filter(dts[s + 1])                 # overlaps the transfer of step s
        sf.gather_end()
        got.append([sf.plane(w) for w in range(5)])
        if s + 1 < 3:
            sf.gather_begin()
    ok = True
    if rank == 0:
        whole = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=3, device=local))
        for s in range(4):
            if s:
                whole.filter(dts[s - 1])
            ref = [whole.u.fluc, whole.v.fluc, whole.w.fluc, whole.T_fluc, whole.rho_fluc]
            for w in range(5):
                ok &= bool(np.array_equal(got[s][w], ref[w]))
        assert not got[0][3].any() and got[1][3].any()
        want = {"p2p": 2, "nccl": 1}.get(os.environ.get("DFB_GATHER_TRANSPORT", "p2p"), 2)
        assert sf.filt.info(12) == want, (sf.filt.info(12), want)
        wire = sf.filt.gather_wire_bytes()
        assert wire == 24 * plane["Ny"] * (plane["Nz"] - (sf.k1 - sf.k0)), wire        # u', v', w' of the other ranks' slabs
        whole.close()
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sf.filt.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_OK" if flag.item() == 1.0 else "SLAB_FAIL", world)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
