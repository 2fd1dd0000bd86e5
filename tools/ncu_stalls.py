"""Print a per-kernel stall / throughput digest of an .ncu-rep (run here, no GPU needed).
usage: python tools/ncu_stalls.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
keys += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("----")
    for k in keys:
        if k in hdr:
            v = r[hdr.index(k)]
            if k.startswith("smsp__average_warps"):
                try:
                    if float(v) < 0.15: continue
                except ValueError: pass
                k = k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
            print(f"  {k[:60]:60s} {v[:70]}")
