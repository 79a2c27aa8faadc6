"""tools/batch_bench.py -- BASELINE.json config 5 on one GPU: P independent inflow planes (default 510x400 geometry of the
reference, distinct RNG stream groups = plane_id) advanced K timesteps each through the C ABI, results device-resident.
One handle (= one pair of CUDA streams) per plane, so the planes' kernels overlap on the GPU.
    python tools/batch_bench.py [--planes 8] [--steps 1000]
Under torchrun every rank runs its own P planes (weak scaling: 8 ranks x 8 planes = the 64 planes of config 5)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _dfb_import  # noqa: E402,F401
import digital_filtering_b200 as dfb  # noqa: E402
from digital_filtering_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=8)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", action="store_true", help="one dfb_filter_batch call per plane instead of a Python loop")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    files = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(files, "files", "RST.dat")):
        mk = lambda p: dfb.DFConfig(vel_fluc_file=os.path.join(files, "files", "RST.dat"), line_file=os.path.join(files, "line.dat"),
                                    seed=2026, plane_id=p, device=local)
        name = "reference default plane 510x400 (RST.dat + line.dat)"
    else:
        plane = W.plane_profile(510, 400, 212, 6)
        mk = lambda p: dfb.DFConfig.from_plane(plane, seed=2026, plane_id=p, device=local)
        name = plane["name"]
    planes = [dfb.DIGITAL_FILTER(mk(rank * args.planes + p), fetch=False) for p in range(args.planes)]
    cells = planes[0].n_cells
    dt = 1e-5
    for df in planes:
        for _ in range(5):
            df.filter(dt)
    for df in planes:
        df.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if args.batch:
        dts = np.full(args.steps, dt)
        for df in planes:
            df.filter_batch(dts)          # enqueues all steps of this plane; planes overlap on the GPU through their streams
    else:
        for _ in range(args.steps):
            for df in planes:
                df.filter(dt)
    for df in planes:
        df.sync()
    el = time.perf_counter() - t0
    t = torch.tensor([el], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    el = float(t.item())
    if rank == 0:
        print(json.dumps(dict(workload=name, n_gpus=world, planes_per_gpu=args.planes, steps=args.steps, batch_call=bool(args.batch), seconds=el,
                              plane_steps_per_s=world * args.planes * args.steps / el,
                              cell_updates_per_s=world * args.planes * args.steps * cells / el,
                              us_per_plane_step=1e6 * el / (args.planes * args.steps))), flush=True)
    for df in planes:
        df.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
