"""Per-SASS-instruction digest of one kernel in an .ncu-rep: executed warp instructions and stall samples
per instruction, plus cumulative shares (run here, no GPU).
usage: python tools/ncu_sass.py rep kernel_regex [instance] [min_pct]"""
import csv, io, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{sub}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if r[0] == "Address": cur["hdr"] = r; continue
    if cur and cur["hdr"] and r[0].startswith("0x"): cur["rows"].append(r)
b = blocks[inst]
h = b["hdr"]; iE = h.index("Instructions Executed"); iS = h.index("# Samples"); iT = h.index("Avg. Threads Executed")
base = int(b["rows"][0][0], 16)
totE = sum(float(r[iE]) for r in b["rows"]); totS = sum(float(r[iS]) for r in b["rows"])
print(f"# {b['name'][:120]}\n# total warp instr {totE:.0f}, samples {totS:.0f}, {len(blocks)} instances")
for r in b["rows"]:
    e, s = float(r[iE]), float(r[iS])
    if 100 * e / totE < minpct and 100 * s / totS < minpct: continue
    print(f"{int(r[0],16)-base:05x} e {100*e/totE:5.2f}% s {100*s/totS:5.2f}% thr {r[iT]:>3} | {r[1].strip()[:90]}")
