"""Top stalled source lines of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.splitlines() if l and not l.startswith("==")]))
hdr, fname, recs = None, "", []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ci = {h: i for i, h in enumerate(hdr)}
        si = hdr.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0]:
        try:
            recs.append((int(r[si] or 0), fname, r))
        except ValueError:
            pass
tot = sum(s for s, _, _ in recs) or 1
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot}")
for s, f, r in sorted(recs, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:3]
    st = " ".join(f"{c[6:]}={v}" for v, c in st if v)
    print(f"{100 * s / tot:5.1f}% {s:7d} {f}:{r[0]:>4} {r[1].strip()[:100]}  [{st}]")
