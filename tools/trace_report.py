"""Print the steady-state per-tile timeline of CTA 0 from SVX_TRACE_DIR dumps."""
import glob, os, sys
import numpy as np
K = 4096
names = {0: {1: "tile", 2: "auxfree", 3: "issued"}, 1: {1: "tile", 2: "accfree", 3: "opnd", 4: "mma_done"},
         2: {1: "tile", 2: "aux", 3: "acc", 4: "conv", 5: "st_iss", 6: "st_read"}, 3: {1: "tile", 2: "aux", 3: "acc", 4: "conv", 5: "st_iss", 6: "st_read"}}
for path in sorted(glob.glob(sys.argv[1] + "/trace*.bin"))[: int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
    a = np.fromfile(path, np.uint64).reshape(4, K)
    print("==", os.path.basename(path))
    t0 = min(int(a[r][0] >> 8) for r in range(4) if a[r][0])
    for role in range(4):
        ev = [(int(x >> 8) - t0, int(x & 0xff)) for x in a[role] if x]
        if not ev:
            continue
        # split into tiles at code 1
        tiles, cur = [], []
        for t, c in ev:
            if c == 1 and cur:
                tiles.append(cur); cur = []
            cur.append((t, c))
        tiles.append(cur)
        n = len(tiles)
        mid = tiles[n // 2: n // 2 + 4]
        starts = [tl[0][0] for tl in tiles]
        per = (starts[-1] - starts[len(starts) // 4]) / max(1, (len(starts) - 1 - len(starts) // 4))
        print(" role %d: %d tiles, steady %.0f ns/tile" % (role, n, per))
        for tl in mid:
            print("    " + "  ".join("%s+%d" % (names[role].get(c, str(c)), t - tl[0][0]) for t, c in tl) + "   @%d" % tl[0][0])
