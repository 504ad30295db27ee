"""Top stalled SASS instructions of an .ncu-rep, with their stall reasons.  usage: python tools/ncu_sass.py rep [top_n] [min_idx max_idx]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.splitlines() if l and not l.startswith("==")]))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; si = hdr.index("# Samples"); src = hdr.index("Source")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) != len(hdr): continue
    try: recs.append((int(r[si] or 0), n, r))
    except ValueError: pass
tot = sum(s for s, _, _ in recs) or 1
# aggregate by opcode too
from collections import Counter
byop = Counter(); bystall = Counter()
for s, n, r in recs:
    op = r[src].split()[0] if r[src].split() else "?"
    if op.startswith("@"): op = r[src].split()[1]
    byop[op.split(".")[0]] += s
    for i, h in stall: bystall[h[6:]] += int(r[i] or 0)
print("total samples", tot)
print("by opcode:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in byop.most_common(14)))
print("by stall :", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in bystall.most_common(10)))
for s, n, r in sorted(recs, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
    print(f"{100*s/tot:5.1f}% #{n:5d} {r[src].strip()[:70]:70s} " + " ".join(f"{h}={v}" for v, h in st if v))
