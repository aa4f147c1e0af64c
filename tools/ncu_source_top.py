"""Top SASS instructions of a kernel by executed count / stall samples from `ncu --page source --csv`.
    ncu -i rep --page source --csv > src.csv ; python tools/ncu_source_top.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try:
        data.append((r[col["Address"]], r[col["Source"]], int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0), r))
    except ValueError:
        pass
tot_s = sum(d[2] for d in data); tot_i = sum(d[3] for d in data)
print("total samples", tot_s, "total warp instructions", tot_i)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("-- by samples")
for i, d in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][2])[:n]
for i in sorted(order):
    d = data[i]
    st = sorted(((int(d[4][col[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% exec %10d  %-70s %s" % (i, 100.0 * d[2] / tot_s, d[3], d[1][:70], " ".join("%s=%d" % (c[6:], v) for v, c in st if v)))
