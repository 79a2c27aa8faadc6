"""tools/soak.py -- development aid: many steps of the tuned kernels against the one-thread-per-cell kernels on the same noise stream
(generate mode, same seed): any race in the tuned path (work queues, completion stamps, programmatic launches, look-ahead noise)
would show up as a divergence."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "1024x2048_profile_N128"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
plane = W.NAMED[name]()
a = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=99), fetch=False)
b = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=99, kernel_variant=1), fetch=False)
worst = 0.0
for s in range(steps):
    a.filter(1e-7); b.filter(1e-7)
    if s % 50 == 49 or s == steps - 1:
        for w in (dfb.U_FLUC, dfb.V_FLUC, dfb.W_FLUC, dfb.T_FLUC, dfb.RHO_FLUC):
            x, y = a.get(w), b.get(w)
            r = float(np.max(np.abs(x - y)) / np.sqrt(np.mean(y * y)))
            worst = max(worst, r)
        print("step", s + 1, "worst deviation / rms so far %.2e" % worst, flush=True)
assert worst < 1e-12, worst
print("soak ok", name, steps, "steps")
