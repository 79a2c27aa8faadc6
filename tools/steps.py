"""tools/steps.py -- development aid: run a few production steps of a named workload (the command profiled under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "1024x2048_profile_N128"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.NAMED[name](), seed=1), fetch=False)
for _ in range(n):
    df.filter(1e-7)
df.sync()
print("ok", name, n, "steps")
