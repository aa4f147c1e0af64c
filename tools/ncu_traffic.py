"""profiles/ncu_traffic.json from ncu --set full reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
candidate kernel, averaged over the captured launches that carry real counters.  bench.py reads the JSON at run time
(roofline.traffic), so the number in a bench line is always the one a committed profile backs.
    python tools/ncu_traffic.py c3=gpurun_out/r02/prof_c3_cand.ncu-rep:profiles/r02_ncu_c3_cand.txt c2=...:..."""
import csv
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    wl, rest = arg.split("=", 1)
    rep, summary = rest.split(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    per_launch = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        if "tc_candidates" not in d.get("Kernel Name", ""):
            continue
        try:
            rd = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
        except (KeyError, ValueError):
            continue
        if math.isnan(rd) or math.isnan(wr):
            per_launch.append(None)
        else:
            per_launch.append({"grid": d.get("Grid Size"), "read": rd, "write": wr, "ms": float(d["gpu__time_duration.sum"])
                               * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0)})
    good = [x for x in per_launch if x]
    if not good:
        print(wl, "no launch with real DRAM counters in", rep)
        continue
    table[wl] = {"bytes_per_launch": sum(x["read"] + x["write"] for x in good) / len(good), "source": summary,
                 "launches_captured": len(per_launch), "launches_with_counters": len(good), "launches": good,
                 "note": "ncu --set full --clock-control none, 1 GPU; mean over the captured launches with real counters"}
    print(wl, table[wl]["bytes_per_launch"] / 1e9, "GB per launch from", len(good), "of", len(per_launch), "launches")
json.dump(table, open(out_path, "w"), indent=1)
