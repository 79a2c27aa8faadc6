"""tools/quick_gpu.py -- development probe (GPU box): fp64 DFMA peak + per-stage CUDA-event times of
the named workloads.  Not the benchmark (bench.py is); numbers land in gpurun_out/quick.json."""
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
    names = sys.argv[1:] or ["512x512_N32", "1024x2048_profile_N128", "1024x2048_saturated_N128"]
    out = {}
    tf, mhz = dfb.measure_fp64_peak()
    out["fp64_peak_tflops"] = tf
    out["implied_sm_mhz"] = mhz
    print("fp64 DFMA peak %.2f TFLOP/s (implied %.0f MHz at 64 DFMA/clk/SM)" % (tf, mhz), flush=True)
    for name in names:
        plane = W.NAMED[name]() if name in W.NAMED else None
        for variant in (0, 1):
            t0 = time.time()
            df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=1, kernel_variant=variant), fetch=False)
            t_create = time.time() - t0
            ny = [df.half_widths(f, 0) for f in range(3)]
            nz = [df.half_widths(f, 1) for f in range(3)]
            ty = int(sum((2 * a.astype(np.int64) + 1).sum() for a in ny))
            tz = int(sum((2 * a.astype(np.int64) + 1).sum() for a in nz))
            df.set_timing(True)
            ms = []
            for i in range(8 if variant == 0 else 3):
                df.filter(1e-7)
                ms.append(df.last_ms())
            med = {k: float(np.median([m[k] for m in ms[2:] or ms])) for k in ms[0]}
            rec = dict(create_s=t_create, tuned=df.tuned, cells=df.n_cells, taps_y=ty, taps_z=tz, ms=med,
                       y_tflops=2 * ty / (med["ysweep"] * 1e-3) / 1e12, z_tflops=2 * tz / (med["zsweep_epilogue"] * 1e-3) / 1e12,
                       step_tflops=2 * (ty + tz) / (med["step"] * 1e-3) / 1e12, mcells_per_s=df.n_cells / (med["step"] * 1e-3) / 1e6)
            out[f"{name}/variant{variant}"] = rec
            print(name, "variant", variant, json.dumps(rec), flush=True)
            df.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "quick.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
