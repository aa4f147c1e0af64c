"""Per-warp view of the epilogue stamps printed by B200M_TC_DEBUG & 1024 (library built with -DB200M_TC_TRACE):
average cycles per tile of each epilogue warp of CTA pair 0 in the TMEM loads, the hand-back, the filtering and the
way back to the next tile's accumulators, and how far the warp runs behind the first warp of its CTA.
    python tools/trace_warps.py gpurun_out/trace_c2_1280.log
"""
import re
import sys
from collections import defaultdict

epi = defaultdict(dict)
for line in open(sys.argv[1]):
    m = re.match(r"b200match trace epi cta (\d+) warp (\d+) tile (\d+): (-?\d+) (-?\d+) (-?\d+) (-?\d+)", line)
    if m:
        c, w, t = int(m.group(1)), int(m.group(2)), int(m.group(3))
        epi[(c, w)].setdefault(t, tuple(int(x) for x in m.groups()[3:]))
tiles = sorted(next(iter(epi.values())))
first = {(c, t): min(v[t][0] for (cc, w), v in epi.items() if cc == c) for c in (0, 1) for t in tiles}
print("cta warp sched | ld  hand-back  filter  to-next-tile  period | lag of 'accumulators seen' behind the CTA's first warp")
for (c, w) in sorted(epi):
    v = epi[(c, w)]
    n = len(tiles)
    ld = sum(v[t][1] - v[t][0] for t in tiles) / n
    ar = sum(v[t][2] - v[t][1] for t in tiles) / n
    fl = sum(v[t][3] - v[t][2] for t in tiles) / n
    nx = sum(v[t + 1][0] - v[t][3] for t in tiles[:-1]) / (n - 1)
    per = (v[tiles[-1]][0] - v[tiles[0]][0]) / (n - 1)
    lag = sum(v[t][0] - first[(c, t)] for t in tiles) / n
    print("%d %2d %d | %5.0f %5.0f %5.0f %5.0f %6.0f | %6.0f" % (c, w, w % 4, ld, ar, fl, nx, per, lag))
