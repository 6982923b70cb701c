#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump (stdin or file):
share of stall samples and of executed warp instructions per CUDA line, with the dominant stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
func = ""
fname = ""
hdr = None
lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if kfilter and kfilter not in func: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[2] != "-": continue            # SASS rows carry an address; aggregated CUDA-line rows have '-'
    d = dict(zip(hdr[4:], r[4:]))
    def num(k):
        try: return int(d.get(k, "0") or 0)
        except ValueError: return 0
    stalls = {k[6:]: num(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
    lines.append((fname, r[0], r[1].strip(), num("# Samples"), num("Instructions Executed"), d.get("Avg. Threads Executed", ""), stalls))
ts = sum(l[3] for l in lines) or 1
ti = sum(l[4] for l in lines) or 1
print(f"total samples {ts}, warp instructions {ti}")
for f, ln, src, s, i, thr_, st in lines:
    if 100.0 * s / ts < thr and 100.0 * i / ti < thr: continue
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    tops = " ".join(f"{k}={100.0 * v / max(s, 1):.0f}%" for k, v in top if v)
    print(f"{f}:{ln:>4} smp {100.0 * s / ts:5.1f}% ins {100.0 * i / ti:5.1f}% thr {thr_:>5} [{tops}] {src[:100]}")
