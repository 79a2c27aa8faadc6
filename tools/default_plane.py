"""tools/default_plane.py -- development aid: per-kernel and per-step times of the reference's default 510x400 plane (and a batch of it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import _dfb_import, digital_filtering_b200 as dfb
rst, ln = os.path.join(ROOT, "oracle/_ref/files/RST.dat"), os.path.join(ROOT, "oracle/_ref/line.dat")
for P in (1, 8):
    d = dfb.DIGITAL_FILTER(dfb.DFConfig(vel_fluc_file=rst, line_file=ln, seed=1), fetch=False, nplanes=P)
    st = torch.cuda.ExternalStream(d.stream())
    for _ in range(20): d.filter(1e-5)
    d.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(300): d.filter(1e-5)
    b.record(st); d.sync()
    step = a.elapsed_time(b) / 300
    d.set_timing(True)
    ms = []
    for _ in range(8):
        d.filter(1e-5); ms.append(d.last_ms())
    import numpy as np
    med = {k: float(np.median([m[k] for m in ms[2:]])) for k in ms[0]}
    print("planes", P, "y_form", d.info(10), "ytiles", d.info(11), "step %.4f ms (%.1f us/plane)" % (step, 1e3 * step / P), {k: round(v, 4) for k, v in med.items()})
    d.close()
