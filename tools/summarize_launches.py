"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    out = [f"# launch list summary: {path}", "", f"{len(rows)} launches, {total/1e6:.3f} ms total kernel time (ncu: cold-cache, serialised)", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
