"""Absolute-time event listing of one trace file (roles interleaved) for a window of the steady state."""
import sys, numpy as np
K = 4096
a = np.fromfile(sys.argv[1], np.uint64).reshape(6, K)
names = {0: "prod", 1: "mma", 2: "aux", 3: "store", 4: "epi0", 5: "epi1"}
codes = {0: {1: "span", 2: "issued"}, 1: {1: "span", 2: "accfree", 3: "opnd", 4: "issued"}, 2: {1: "want_slot", 2: "fill_issued"},
         3: {1: "ready", 2: "st_issued"}, 4: {1: "span", 2: "acc", 3: "slot", 4: "conv"}, 5: {1: "span", 2: "acc", 3: "slot", 4: "conv"}}
ev = []
for r in range(6):
    for x in a[r]:
        if x: ev.append((int(x >> 8), r, int(x & 0xff)))
ev.sort()
t0 = ev[0][0]
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
hi = float(sys.argv[3]) if len(sys.argv) > 3 else lo + 14
for t, r, c in ev:
    us = (t - t0) / 1e3
    if lo <= us <= hi:
        print("%9.3f us  %-6s %s" % (us, names[r], codes[r].get(c, c)))
