"""Per-instruction warp-stall summary of the first kernel in an .ncu-rep (needs --import-source on / SourceCounters).

    python profiles/stall_summary.py report.ncu-rep [top_n]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(h) or r[0] == "Address":   # a second kernel of the report starts: only the first one is summarised
        break
    data.append(r)
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
base = int(data[0][ia], 16)
tot = sum(int(r[isamp]) for r in data)
print(rows[0][1][:140])
print("samples", tot, "warp instructions executed", sum(int(r[iex]) for r in data))
agg = collections.Counter()
for r in data:
    for i in stall:
        agg[h[i]] += int(r[i])
print("stall totals:", [(k, v) for k, v in agg.most_common(10)])
for r in sorted(data, key=lambda r: -int(r[isamp]))[:top_n]:
    a = int(r[ia], 16) - base
    st = sorted(((int(r[i]), h[i][6:]) for i in stall), reverse=True)[:3]
    print(f"{a:#07x} {int(r[isamp]):6d} {int(r[iex]):9d}  {r[isrc].strip()[:64]:64s} {[s for s in st if s[0]]}")
