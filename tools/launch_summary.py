"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        n = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}%  n={c:4d} avg={t/c:9.1f} us  {n[:100]}")
    print(f"total {tot/1e3:.2f} ms")

if __name__ == "__main__":
    main(sys.argv[1])
