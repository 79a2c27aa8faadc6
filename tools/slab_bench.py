"""tools/slab_bench.py -- BASELINE.json config 4: one 4096x8192 plane split into spanwise slabs over the
ranks of a torchrun job (one process per GPU), noise halos regenerated locally (no halo exchange),
NCCL used only to gather the finished plane to rank 0.  Reports ms/step without and with the gather,
and checks slab == whole-plane on a small plane first.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_bench.py [--workload 4096x8192_profile_N128] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _dfb_import  # noqa: E402,F401
import digital_filtering_b200 as dfb  # noqa: E402
from digital_filtering_b200 import parallel as P, workloads as W  # noqa: E402

FIELDS = [dfb.U_FLUC, dfb.V_FLUC, dfb.W_FLUC, dfb.T_FLUC, dfb.RHO_FLUC]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="4096x8192_profile_N128")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- correctness first: slabs + gather == the single-GPU plane, bit for bit ----
    small = W.plane_profile(64, 700, 16, 24)
    mk = lambda plane: (lambda k0, k1: dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=3, device=local, k_begin=k0, k_end=k1), fetch=False))
    sf = P.SlabFilter(dist, small["Nz"], mk(small))
    sf.filter(2e-7); sf.filter(2e-7)
    sf.filt.sync()
    bufs = sf.gather(FIELDS, torch, dev)
    ok = True
    if rank == 0:
        whole = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(small, seed=3, device=local))
        whole.filter(2e-7); whole.filter(2e-7)
        ref = [whole.u.fluc, whole.v.fluc, whole.w.fluc, whole.T_fluc, whole.rho_fluc]
        ok = all(np.array_equal(sf.plane_on_dst(bufs, i), ref[i]) for i in range(5))
        whole.close()
    sf.filt.close()

    # ---- config 4 ----
    plane = W.NAMED[args.workload]()
    sf = P.SlabFilter(dist, plane["Nz"], mk(plane))
    stream = torch.cuda.ExternalStream(sf.filt.stream(), device=local)

    def timed(with_gather):
        for _ in range(3):
            sf.filter(1e-7)
            if with_gather:
                with torch.cuda.stream(stream):
                    sf.gather(FIELDS, torch, dev)
        sf.filt.sync(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            sf.filter(1e-7)
            if with_gather:
                with torch.cuda.stream(stream):
                    sf.gather(FIELDS, torch, dev)
        b.record(stream)
        sf.filt.sync(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps

    def timed_overlapped():
        """filter(t+1) runs while the gather of step t is on the wire (SlabFilter.gather_async)"""
        prev = None
        for _ in range(3):
            sf.filter(1e-7)
            tk = sf.gather_async(FIELDS, torch, dev)
            if prev is not None:
                sf.gather_wait(prev, torch)
            prev = tk
        sf.gather_wait(prev, torch)
        sf.filt.sync(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        prev = None
        for _ in range(args.steps):
            sf.filter(1e-7)
            tk = sf.gather_async(FIELDS, torch, dev)
            if prev is not None:
                sf.gather_wait(prev, torch)
            prev = tk
        last = sf.gather_wait(prev, torch)
        torch.cuda.current_stream().wait_stream(stream)
        b.record(torch.cuda.current_stream())
        sf.filt.sync(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, last

    ms_nogather = timed(False)
    ms_gather = timed(True)
    ms_overlap, last = timed_overlapped()
    # the overlapped gather must deliver the same plane as the synchronous one (same step count on every rank)
    sf.filt.sync()
    ref_bufs = sf.gather(FIELDS, torch, dev)
    ok_overlap = True
    if rank == 0:
        ok_overlap = all(bool(torch.equal(x, y)) for x, y in zip(last, ref_bufs))
    if rank == 0:
        cells = plane["Ny"] * plane["Nz"]
        print(json.dumps(dict(workload=plane["name"], n_gpus=world, slab_equals_whole_plane=bool(ok), slabs=sf.bounds,
                              ms_per_step_no_gather=ms_nogather, ms_per_step_with_gather=ms_gather, ms_per_step_with_gather_overlapped=ms_overlap,
                              overlapped_gather_equals_synchronous=bool(ok_overlap),
                              cell_updates_per_s_no_gather=cells / (ms_nogather * 1e-3), cell_updates_per_s_with_gather=cells / (ms_gather * 1e-3),
                              gathered_bytes_per_step=40 * cells * (world - 1) // world)), flush=True)
    sf.filt.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
