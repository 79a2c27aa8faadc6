"""tools/accuracy.py -- development aid (uses the test oracle): worst normwise deviation of the GPU path from the oracle restatement
under identical noise, for the recursive and the direct z-sweep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
import oracle as O
import test_gpu_parity as T
O.build()
for mode in ("1", "0"):
    os.environ["DFB_Z_MODE"] = mode
    for name, plane in (("profile 96x1100 N<=128", W.plane_profile(96, 1100, 128, 128)), ("saturated 64x1100 N=128", W.plane_saturated(64, 1100, 128)),
                        ("saturated 48x2048 N=254", W.plane_saturated(48, 2048, 254))):
        try:
            w = T.inject_and_step(dfb, O, plane, 0, 5, [2e-7, 2e-7, 1e-6])
            print("z mode", mode, name, {k: "%.2e" % v for k, v in w.items()}, flush=True)
        except Exception as e:
            print("z mode", mode, name, "FAILED", str(e)[:200], flush=True)
