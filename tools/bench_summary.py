import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d["roofline"]
    print("%s: value %.0f q/s  ms/step %.2f  e2e %.0f (%.1f ms)  cand %.1f TF frac %.3f launch_ms %.3f  clocks %s" % (
        f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["achieved"], r["frac"], r["launch_ms"], d["clocks"]))
    print("   ", {k: round(v, 3) for k, v in r["breakdown_ms_per_step"].items()}, "cand/row %.1f overflow %.5f" % (r["candidates_per_row"], r["rows_overflowed_frac"]))
    if d.get("cpu_baseline"):
        print("    cpu", d["cpu_baseline"])
