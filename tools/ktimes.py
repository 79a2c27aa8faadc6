"""tools/ktimes.py -- development aid: per-kernel times (timing mode: kernels serialised, look-ahead noise off) and the production
ms/step of named workloads, tuned kernels only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
names = sys.argv[1:] or ["1024x2048_profile_N128", "1024x2048_saturated_N128", "4096x8192_profile_N128"]
for name in names:
    # a name like profile:1024:2048:128:64 is W.plane_profile(Ny, Nz, max N_y, max N_z)
    plane = W.plane_profile(*map(int, name.split(":")[1:])) if name.startswith("profile:") else W.NAMED[name]()
    df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=1), fetch=False)
    st = torch.cuda.ExternalStream(df.stream())
    for _ in range(30): df.filter(1e-7)
    df.sync()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(300): df.filter(1e-7)
        b.record(st); df.sync()
        best = min(best, a.elapsed_time(b) / 300)
    df.set_timing(True)
    ms = []
    for _ in range(10):
        df.filter(1e-7); ms.append(df.last_ms())
    med = {k: round(float(np.median([m[k] for m in ms[2:]])), 4) for k in ms[0]}
    print(name, "y_form", df.info(10), "rtiles", df.info(11), "step %.4f ms" % best, med, flush=True)
    df.close()
