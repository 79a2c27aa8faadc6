"""tools/timeline.py -- development aid: when do noise / y-sweep / z-sweep of consecutive steps run (global timer)?"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["DFB_TIMELINE"] = "1"
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "1024x2048_profile_N128"
if name == "default":
    df = dfb.DIGITAL_FILTER(dfb.DFConfig(vel_fluc_file=os.path.join(ROOT, "oracle/_ref/files/RST.dat"), line_file=os.path.join(ROOT, "oracle/_ref/line.dat"), seed=1), fetch=False)
else:
    df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.NAMED[name](), seed=1), fetch=False)
print(name, "Ny x Nz", df.Ny, df.Nz, "Ny_max", [df.info(0, f) for f in range(3)], "Nz_max", [df.info(1, f) for f in range(3)])
L = dfb.lib(); L.dfb_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = (ctypes.c_uint64 * 512)()
for _ in range(20): df.filter(1e-7)
df.sync()
L.dfb_debug_timeline(df._h, out)
n0 = 21
for _ in range(12): df.filter(1e-7)
df.sync()
L.dfb_debug_timeline(df._h, out)
v = np.array(list(out), dtype=np.uint64).reshape(64, 4, 2)
rows = []
for s in range(64):
    if v[s, 1, 1] > 0: rows.append((int(v[s, 1, 0]), s))
rows.sort()
t0 = rows[0][0]
prev_zend = None
for _, s in rows:
    f = lambda k, e: (int(v[s, k, e]) - t0) / 1e3 if (v[s, k, 1] > 0 and v[s, k, 0] < 2**63) else float("nan")
    line = "slot %2d  noise %7.1f..%7.1f  y %7.1f..%7.1f  z %7.1f..%7.1f us" % (s, f(0, 0), f(0, 1), f(1, 0), f(1, 1), f(2, 0), f(2, 1))
    if prev_zend is not None: line += "   gap z->y %.1f  y->z %.1f  step %.1f" % (f(1, 0) - prev_zend, f(2, 0) - f(1, 1), f(2, 1) - prev_zend)
    prev_zend = f(2, 1)
    print(line)
