import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ.setdefault("DFB_DEBUG_Z", "16")
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "1024x2048_saturated_N128"
df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.NAMED[name](), seed=1), fetch=False)
L = dfb.lib(); L.dfb_debug_zprof.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = (ctypes.c_uint64 * 128)()
for _ in range(3): df.filter(1e-7)
L.dfb_debug_zprof(df._h, out)
df.set_timing(True)
df.filter(1e-7); ms = df.last_ms()
L.dfb_debug_zprof(df._h, out)
v = list(out); n = max(v[4], 1)
w = max(v[7], 1)
print("   warps", v[7], "avg life %.0f  max life %.0f cycles" % (v[5] / w, v[6]))
print(name, "z ms", ms["zsweep_epilogue"], "units", v[4], "per-unit cycles: wait %.0f  taps %.0f  epilogue %.0f  whole %.0f  | loop top %.0f (of which staging the next unit %.0f)" % (v[0]/n, v[1]/n, v[2]/n, v[3]/n, v[9]/n, v[8]/n))
B = 1 << 62
s0, s1, e0, e1 = B - v[11], v[10], B - v[13], v[12]
print("   global timer (ns): starts spread %d, ends spread %d, first start -> last end %d, avg life %.0f" % (s1 - s0, e1 - e0, e1 - s0, v[14] / w))
print("   warp lifetimes, 1 us bins:", {i: v[16 + i] for i in range(64) if v[16 + i]})
print("   units per warp:", {i: v[80 + i] for i in range(32) if v[80 + i]})
