"""Per-CUDA-source-line digest (instructions executed, stall samples) of one kernel in an .ncu-rep (run here, no GPU).
usage: python tools/ncu_lines.py rep kernel_regex [topN]   (all captured instances are summed)"""
import collections, csv, io, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{sub}"],
                     capture_output=True, text=True).stdout
hdr, fname, agg = None, "?", collections.OrderedDict()
def num(x):
    try: return float(x)
    except ValueError: return 0.0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iE, iS = r.index("Instructions Executed"), r.index("# Samples"); continue
    if hdr and r[0].isdigit():
        k = (fname, int(r[0]))
        if k not in agg: agg[k] = [0.0, 0.0, r[1]]
        agg[k][0] += num(r[iE]); agg[k][1] += num(r[iS])
totE = sum(v[0] for v in agg.values()) or 1; totS = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {totE:.0f}, samples {totS:.0f}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:14]:14s} L{l:>4} inst {100*v[0]/totE:5.2f}% samp {100*v[1]/totS:5.2f}% | {v[2].strip()[:105]}")
