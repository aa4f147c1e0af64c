"""Reads the cycle stamps printed by B200M_TC_DEBUG & 1024 and prints, per traced tile, the hand-off chain on the
leader SM of CTA pair 0 (all times in SM cycles relative to the first stamp; the peer SM has its own clock, so only
its durations are shown).
    python tools/trace_report.py gpurun_out/trace_c2_1024.log
"""
import re
import sys
from collections import defaultdict

mma, epi = {}, defaultdict(dict)
for line in open(sys.argv[1]):
    m = re.match(r"b200match trace mma tile (\d+): ([-\d ]+)$", line)
    if m:
        mma.setdefault(int(m.group(1)), tuple(int(x) for x in m.group(2).split()))   # first launch only
        continue
    m = re.match(r"b200match trace epi cta (\d+) warp (\d+) tile (\d+): (-?\d+) (-?\d+) (-?\d+) (-?\d+)", line)
    if m:
        c, w, t = int(m.group(1)), int(m.group(2)), int(m.group(3))
        epi[(c, w)].setdefault(t, tuple(int(x) for x in m.groups()[3:]))
tiles = sorted(mma)
base = mma[tiles[0]][0]
print("tile | MMA warp: operands, buffer back, mma1, mma2, mma3/4, commit(stage), commit(accumulator) | leader epilogue warps: tfull seen (min..max), ld done (max), "
      "arrive (max), filter done (max) | peer: ld, filter durations (max)")
prev_issue = None
for t in tiles:
    st = [x - base for x in mma[t]]
    a, b, c = st[0], st[1], st[-1]
    lead = [epi[(0, w)][t] for (cc, w) in epi if cc == 0 and t in epi[(cc, w)]]
    peer = [epi[(1, w)][t] for (cc, w) in epi if cc == 1 and t in epi[(cc, w)]]
    s0 = [x[0] - base for x in lead]
    s1 = [x[1] - base for x in lead]
    s2 = [x[2] - base for x in lead]
    s3 = [x[3] - base for x in lead]
    pl = max(x[1] - x[0] for x in peer) if peer else 0
    pf = max(x[3] - x[2] for x in peer) if peer else 0
    print("%4d | %s (period %s) | %6d..%6d %6d %6d %6d | %4d %4d" % (
        t, " ".join("%6d" % x for x in st), "-" if prev_issue is None else c - prev_issue, min(s0), max(s0), max(s1), max(s2), max(s3), pl, pf))
    prev_issue = c
