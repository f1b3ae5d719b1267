"""Text summary of an .ncu-rep (ncu --set full) for profiles/: the metrics the roofline / stall discussion in DESIGN.md
uses.  usage: python tools/ncu_summary.py <file.ncu-rep> <profiles/out.txt> "<header line(s)>" """
import csv
import subprocess
import sys

rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio")
lines = [header]
for li, vals in enumerate(rows[2:]):
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines.append("--- launch %d: %s" % (li, name[:100]))
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
            try:
                if float(v.replace(",", "")) == 0.0 and "stalled" in h:
                    continue
            except ValueError:
                pass
            lines.append("%-92s %-16s %s" % (h, u, v))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:50]))
