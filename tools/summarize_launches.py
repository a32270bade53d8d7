"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name.  Usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md] [traffic.json]
With DRAM metrics present the table gains per-kernel DRAM traffic and `traffic.json` receives the per-launch DRAM bytes of the
GEMM kernel (bench.py reads it for `roofline.traffic`)."""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "s": 1e9}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per_id = defaultdict(dict)   # launch id -> {name, ns, rd, wr}
    for r in csv.DictReader(lines):
        m = r.get("Metric Name")
        val = float(r["Metric Value"].replace(",", "")) if r.get("Metric Value") not in (None, "", "n/a") else 0.0
        unit = r.get("Metric Unit", "")
        d = per_id[r["ID"]]
        d["name"] = r["Kernel Name"]
        if m == "gpu__time_duration.sum":
            d["ns"] = val * UNIT.get(unit, 1)
        elif m == "dram__bytes_read.sum":
            d["rd"] = val * BYTES.get(unit, 1)
        elif m == "dram__bytes_write.sum":
            d["wr"] = val * BYTES.get(unit, 1)
    rows = [d for d in per_id.values() if "ns" in d]
    have_dram = any("rd" in d for d in rows)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in rows:
        short = re.sub(r"\(.*", "", d["name"])
        short = re.sub(r"^void ", "", short)
        a = agg[short]
        a[0] += 1
        a[1] += d["ns"]
        a[2] += d.get("rd", 0.0) + d.get("wr", 0.0)
    total = sum(v[1] for v in agg.values())
    out = [f"# launch list summary: {path}", "",
           f"{len(rows)} launches, {total/1e6:.3f} ms total kernel time (ncu: cold-cache, serialised)", ""]
    if have_dram:
        out += ["| kernel | launches | total ms | share | avg us | DRAM MB / launch | DRAM GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
    else:
        out += ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |"
        if have_dram:
            line += f" {by/n/1e6:.2f} | {by/ns:.0f} |"
        out.append(line)
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    if len(sys.argv) > 3 and have_dram:
        g = [(k, v) for k, v in agg.items() if "gemm_bf16_tcgen05_kernel" in k]
        n = sum(v[0] for _, v in g)
        by = sum(v[2] for _, v in g)
        ns = sum(v[1] for _, v in g)
        json.dump({"kernel": "gemm_bf16_tcgen05_kernel", "launches": n, "dram_bytes_per_launch": by / max(n, 1),
                   "avg_launch_us_under_ncu": ns / max(n, 1) / 1e3, "gemm_share_of_kernel_time": ns / total,
                   "source": path, "how": "ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                   "dram__bytes_write.sum --clock-control none over the timed step of `bench.py --steps 1 --warmup 3`"},
                  open(sys.argv[3], "w"), indent=1)
    print(text)


if __name__ == "__main__":
    main()
