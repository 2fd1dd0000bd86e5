"""Per-source-line instruction / stall-sample digest of one kernel in an .ncu-rep (run here, no GPU).
usage: python tools/ncu_lines.py rep kernel_regex [topN]"""
import csv, io, os, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", f"regex:{sub}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur, body = None, "?", []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) >= len(hdr) - 1: body.append((cur, r))
ci = {}
for i, h in enumerate(hdr): ci.setdefault(h, i)
def num(r, k):
    try: return float(r[ci[k]])
    except (ValueError, KeyError, IndexError): return 0.0
tot_i = sum(num(r, "Instructions Executed") for _, r in body)
tot_s = sum(num(r, "# Samples") for _, r in body)
print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
body.sort(key=lambda fr: -num(fr[1], "# Samples"))
for f, r in body[:top]:
    st = {k: num(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    top2 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{f[:14]:14s}{r[0]:>4} inst {100*num(r,'Instructions Executed')/max(tot_i,1):5.1f}% samp {100*num(r,'# Samples')/max(tot_s,1):5.1f}% "
          f"{' '.join(f'{k[6:]}={v:.0f}' for k, v in top2):30s} | {r[1].strip()[:100]}")
