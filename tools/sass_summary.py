"""profiles/sass_summary.txt: per kernel instantiation of the built objects, the SASS mnemonics that show which hardware
paths the code uses (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA = UTMALDG / UBLKCP, tcgen05.commit = UTCBAR, mbarrier =
SYNCS, register re-division = USETMAXREG), registers and spills.  Run after a build:  python tools/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lidar_global_registration_b200", "csrc")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "USETMAXREG", "FMNMX3", "MUFU", "STL", "LDL"]
out = open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w")
out.write("SASS mnemonic counts per kernel (cuobjdump -sass of the in-tree objects built by lidar_global_registration_b200/build.py)\n")
for obj in sorted(f for f in os.listdir(CSRC) if f.endswith(".o")):
    text = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    cur, counts = None, collections.OrderedDict()
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA.2CTA"):
                counts[cur]["UTCHMMA.2CTA"] += 1
    out.write("\n== %s ==\n" % obj)
    for fn, c in counts.items():
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(CUtensorMap_st.*", "", name).replace("(anonymous namespace)::", "")
        if len(name) > 110:
            name = name[:107] + "..."
        hot = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
        out.write("%-112s regs=%-4s instr=%-6d %s\n" % (name, regs.get(fn, "?"), sum(v for k, v in c.items() if k != "UTCHMMA.2CTA"), hot))
out.close()
print(open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()[:3000])
