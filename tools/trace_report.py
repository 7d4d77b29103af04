"""Print the steady-state per-span timeline of CTA 0 from SVX_TRACE_DIR dumps (flat kernel: 6 roles)."""
import glob, os, sys
import numpy as np
K = 4096
names = {0: {1: "span", 2: "issued"}, 1: {1: "span", 2: "accfree", 3: "opnd", 4: "mma_issued"},
         3: {1: "ready", 2: "st_issued"}, 4: {1: "span", 2: "acc", 3: "slot", 4: "conv"}, 5: {1: "span", 2: "acc", 3: "slot", 4: "conv"}}
pat = sys.argv[3] if len(sys.argv) > 3 else ""
files = [f for f in sorted(glob.glob(sys.argv[1] + "/trace*.bin")) if pat in f]
for path in files[: int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
    a = np.fromfile(path, np.uint64).reshape(6, K)
    print("==", os.path.basename(path))
    firsts = [int(a[r][0] >> 8) for r in range(6) if a[r][0]]
    if not firsts:
        continue
    t0 = min(firsts)
    for role in (0, 1, 3, 4, 5):
        ev = [(int(x >> 8) - t0, int(x & 0xff)) for x in a[role] if x]
        if not ev:
            continue
        tiles, cur = [], []
        for t, c in ev:
            if c == 1 and cur:
                tiles.append(cur); cur = []
            cur.append((t, c))
        tiles.append(cur)
        n = len(tiles)
        starts = [tl[0][0] for tl in tiles]
        per = (starts[-1] - starts[len(starts) // 4]) / max(1, (len(starts) - 1 - len(starts) // 4))
        print(" role %d: %d units, steady %.0f ns/unit, total %d ns" % (role, n, per, ev[-1][0]))
        for tl in tiles[n // 2: n // 2 + 3]:
            print("    " + "  ".join("%s+%d" % (names[role].get(c, str(c)), t - tl[0][0]) for t, c in tl[:14]) + "   @%d" % tl[0][0])
