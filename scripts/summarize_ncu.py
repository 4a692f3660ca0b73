"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list per
kernel: launches, mean duration, mean DRAM bytes, achieved DRAM GB/s and the fraction of the measured HBM peak."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

UNIT = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1.0, "second": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    peak = 6460.9
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk)).get("hbm_gbs", peak))
        except Exception:
            pass
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    table(rows, peak, "all launches of the command (sample-group launches on their own streams included)")
    # bench.py runs its per-kernel pass (and the forward-only rollout) with the whole batch per launch on the caller's
    # stream, the sample groups on side streams: the first profiled launch belongs to the caller's stream
    main = rows[0]["Stream"]
    sel = [r for r in rows if r["Stream"] == main]
    if len(sel) != len(rows):
        print()
        table(sel, peak, "launches on the caller's stream only (whole batch per launch: the per-kernel pass of bench.py)")


def table(rows, peak, title):
    per = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        d = per.setdefault(name, dict(ids=set(), m=defaultdict(float)))
        d["ids"].add(r["ID"])
        val = float(r["Metric Value"].replace(",", ""))
        d["m"][r["Metric Name"]] += val * UNIT.get(r["Metric Unit"], 1.0)
    total = sum(d["m"]["gpu__time_duration.sum"] for d in per.values())
    print("%s\n" % title)
    print("| kernel | launches | mean time (us) | share of GPU time | DRAM read+write per launch (MB) | DRAM GB/s | %% of HBM peak (%.1f GB/s) |" % peak)
    print("|---|---|---|---|---|---|---|")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["m"]["gpu__time_duration.sum"]):
        n = len(d["ids"])
        t = d["m"]["gpu__time_duration.sum"]
        b = d["m"]["dram__bytes_read.sum"] + d["m"]["dram__bytes_write.sum"]
        gbs = b / t / 1e9 if t > 0 else 0.0
        print("| `%s` | %d | %.1f | %.1f %% | %.2f | %.0f | %.1f %% |" % (name[:90], n, t / n * 1e6, 100 * t / total, b / n / 1e6, gbs, 100 * gbs / peak))


if __name__ == "__main__":
    main(sys.argv[1])
