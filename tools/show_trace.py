"""Print a tcgen05 attention bring-up timeline (RBM_TC_ATTN_TRACE dump): time in us relative to the first event."""
import sys
import numpy as np

NAMES = {1: "mma  wait ds_full", 2: "mma  dq issue begin", 3: "mma  dq issued", 4: "mma  S/dP issue begin", 5: "mma  S/dP issued",
         6: "soft S ready", 7: "soft dS done", 8: "aux  ops written", 9: "aux  dq read out", 10: "tma  issued", 11: "split done", 12: "mma  wait ops_full", 13: "mma  ops_full ok", 14: "aux  wait ops_free", 15: "aux  ops_free ok", 16: "aux  rows loaded"}
a = np.fromfile(sys.argv[1], dtype=np.uint64)
ev = a[1:]
ev = ev[ev != 0]
t = (ev >> np.uint64(24)).astype(np.int64)
e = ((ev >> np.uint64(16)) & np.uint64(0xff)).astype(int)
g = (ev & np.uint64(0xffff)).astype(int)
order = np.argsort(t, kind="stable")
t0 = t[order[0]]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mhz = 1965.0
for k in order[lo:hi]:
    arg = "%d.q%d" % (g[k] // 4, g[k] % 4) if e[k] in (6, 7, 8) else str(g[k])
    print("%9.3f us  %-22s %s" % ((t[k] - t0) / mhz, NAMES.get(e[k], str(e[k])), arg))
