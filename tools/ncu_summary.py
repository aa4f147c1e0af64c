"""Condense an .ncu-rep (ncu --set full) into the metrics the judge reads: python tools/ncu_summary.py rep out.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "sm__icc_request_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("kernel: %s  grid %s block %s\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for h, u, v in zip(hdr, units, r):
            if any(h.endswith(k) or h == k for k in KEYS) or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                f.write("  %-95s %-12s %s\n" % (h, u, v))
print(open(out).read()[:3000])
