"""tools/sanitize_small.py -- a few production steps of small planes through every tuned kernel form (run under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
for shape, kw in (((96, 700, 24, 40), {}), ((72, 530, 20, 16), dict(nplanes=3)), ((40, 96, 8, 6), {})):
    df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.plane_profile(*shape), seed=1), fetch=False, **kw)
    if not kw:
        df.stats_enable(True)
    for _ in range(4):
        df.filter(1e-7)
    df.sync()
    print("ok", shape, kw, "y form", df.info(10), "z form", df.info(7), flush=True)
    df.close()
