"""tools/stepbench.py -- development aid: production ms/step (no per-kernel timing) of named workloads for the library in DFB_LIB."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
names = sys.argv[1:] or ["1024x2048_profile_N128", "1024x2048_saturated_N128"]
res = []
for name in names:
    df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.NAMED[name](), seed=1), fetch=False)
    st = torch.cuda.ExternalStream(df.stream())
    for _ in range(30): df.filter(1e-7)
    df.sync()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(400): df.filter(1e-7)
        b.record(st); df.sync()
        best = min(best, a.elapsed_time(b) / 400)
    res.append("%s %.4f" % (name.split("_")[1], best))
    df.close()
print(os.path.basename(os.environ.get("DFB_LIB", "default")), "ms/step:", "  ".join(res))
