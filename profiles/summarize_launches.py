"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of one forward)."""
import collections, csv, re, sys

def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
        out.append((name, row.get("Grid Size", ""), ms))
    return out

if __name__ == "__main__":
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, ms in rows:
        agg[n][0] += 1; agg[n][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}, total {tot:.3f} ms (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:3d}  avg {v[1]/v[0]*1e3:8.1f} us  {k}")
    if len(sys.argv) > 2:
        print("first launches:")
        for r in rows[: int(sys.argv[2])]:
            print("  %-60s %-16s %8.1f us" % (r[0][:60], r[1], r[2] * 1e3))
