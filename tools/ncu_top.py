"""Top stall sites of an ncu report (SASS view): python tools/ncu_top.py <report.ncu-rep> [N] [kernel-name regex]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[S]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(float(r[ix[h]]) for r in body) for h in stalls}
print("stall mix:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
order = sorted(range(len(body)), key=lambda i: -float(body[i][S]))
for i in order[:N]:
    r = body[i]
    top = sorted(((float(r[ix[h]]), h) for h in stalls), reverse=True)[:2]
    print(f"{100*float(r[S])/tot:5.1f}%  #{i:5d} exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {top[0][1]}:{int(top[0][0])} {top[1][1]}:{int(top[1][0])}")
