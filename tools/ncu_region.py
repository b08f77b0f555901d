"""Samples per SASS region of an ncu report: python tools/ncu_region.py rep.ncu-rep start end [top]"""
import csv, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]); top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]; tot = sum(float(r[S]) for r in body)
reg = body[a:b]
print("region samples %.1f%% of total, %d instrs" % (100 * sum(float(r[S]) for r in reg) / tot, len(reg)))
ops = {}
for r in reg:
    op = r[ix["Source"]].strip().split()[0 if not r[ix["Source"]].strip().startswith("@") else 1]
    ops[op] = ops.get(op, 0) + 1
print("op histogram:", sorted(ops.items(), key=lambda kv: -kv[1])[:18])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(float(r[ix[h]]) for r in reg) for h in stalls}
rs = sum(agg.values()) or 1
print("stall mix:", {k: round(100 * v / rs, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.02 * rs})
order = sorted(range(len(reg)), key=lambda i: -float(reg[i][S]))
for i in order[:top]:
    r = reg[i]
    t = sorted(((float(r[ix[h]]), h) for h in stalls), reverse=True)[0]
    print(f"{100*float(r[S])/tot:5.2f}%  #{a+i:5d} exec={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:80]:80s} {t[1]}:{int(t[0])}")
