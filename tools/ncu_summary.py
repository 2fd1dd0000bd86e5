"""Markdown table of the drtk_b200 kernels in an .ncu-rep (one row per distinct kernel, last captured instance).
usage: python tools/ncu_summary.py rep.ncu-rep > profiles/xxx.md   (run here, no GPU)"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
def col(r, k, scale=1.0, fmt="{:.3g}"):
    if k not in h: return "-"
    try: return fmt.format(float(r[h.index(k)]) * scale)
    except ValueError: return r[h.index(k)]
def unit(k): return units[h.index(k)] if k in h else ""
last = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if "unnamed>::" not in name: continue   # ncu prints our anonymous-namespace kernels as "unnamed>::name"
    short = name.split("unnamed>::")[1].split("(")[0]
    last[short] = r
print("| kernel | time | dram rd | dram wr | dram % peak | L1 data pipe % | issue % | warp instr | regs | occ % | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for k, r in last.items():
    t = f"{col(r, 'gpu__time_duration.sum')} {unit('gpu__time_duration.sum')}"
    rd = f"{col(r, 'dram__bytes_read.sum')} {unit('dram__bytes_read.sum')}"
    wr = f"{col(r, 'dram__bytes_write.sum')} {unit('dram__bytes_write.sum')}"
    print(f"| `{k}` | {t} | {rd} | {wr} | {col(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{col(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')} | "
          f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | {col(r, 'smsp__inst_executed.sum', fmt='{:.4g}')} | "
          f"{col(r, 'launch__registers_per_thread', fmt='{:.0f}')} | {col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
          f"{col(r, 'lts__t_sector_hit_rate.pct')} |")
