"""tools/ztrace.py -- development aid: timeline of the z-sweep warps resident on SM 0 (DFB_DEBUG_Z=64)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["DFB_DEBUG_Z"] = "64"
import _dfb_import, digital_filtering_b200 as dfb
from digital_filtering_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "1024x2048_saturated_N128"
df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(W.NAMED[name](), seed=1), fetch=False)
L = dfb.lib(); L.dfb_debug_ztrace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_uint64 * 4096)()
for _ in range(3): df.filter(1e-7)
L.dfb_debug_ztrace(df._h, out, 4096)
df.filter(1e-7)
L.dfb_debug_ztrace(df._h, out, 4096)
v = np.array(list(out), dtype=np.uint64).astype(np.int64)
nslots = int(v[8])
st = v[2048:2048 + 148]; en = v[2304:2304 + 148]; cnt = v[2560:2560 + 148]
g0 = st.min()
print("per-SM first-warp start (us after earliest): min %.1f med %.1f max %.1f" % tuple(np.percentile((st - g0) / 1e3, [0, 50, 100])))
print("per-SM last-warp end   (us after earliest): min %.1f med %.1f max %.1f" % tuple(np.percentile((en - g0) / 1e3, [0, 50, 100])))
print("units per SM: min %d med %d max %d" % tuple(np.percentile(cnt, [0, 50, 100])))
df.set_timing(True); df.filter(1e-7); print("z kernel (events) ms", df.last_ms()["zsweep_epilogue"]); 
print("slots", nslots)
tr = v[16:16 + 15 * 256].reshape(15, 256)
t0 = min(int(tr[s, 0]) for s in range(min(nslots, 15)) if tr[s, 0] > 0)
for smsp in range(4):
    print("SMSP", smsp)
    for s in range(min(nslots, 15)):
        if int(tr[s, 255]) != smsp: continue
        u = tr[s, :252].reshape(63, 4)
        u = u[u[:, 0] > 0]
        print("  slot %2d: " % s + "  ".join("[%d wait %d taps %d epi %d]" % (a - t0, b - a, c - b, d - c) for a, b, c, d in u))
