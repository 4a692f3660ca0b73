"""Turns the ncu launch list of `bench.py --config5-only --config5-maxit M` into profiles/r02_cg_global.{json,md}:
DRAM bytes per cell and CG iteration of pressure_cg_global_kernel and the achieved HBM fraction of the config-#5 kernels.

    python scripts/c5_profile_summary.py gpurun_out/c5_kernels.csv M [n] [batch]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, max_it = sys.argv[1], int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6460.9
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(path)) if r and r[0] == "ID")
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
             "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
    launches = {}
    for r in rows:
        d = launches.setdefault(r[0], {"kernel": r[ik]})
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    out, lines = {}, ["# Round 2: BASELINE configs[4] (periodic %d^2, batch %d) -- ncu DRAM traffic of the solver kernels" % (n, batch), "",
                      "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over "
                      "`python bench.py --config5-only --config5-maxit %d` (every CG launch runs exactly %d iterations).  "
                      "Peak = %.1f GB/s (MEASURED_PEAKS.json)." % (max_it, max_it, peak), "",
                      "| kernel | launch | time (ms) | DRAM read (MB) | DRAM write (MB) | GB/s | % of HBM peak | B / cell / iteration |", "|---|---|---|---|---|---|---|---|"]
    cells = batch * n * n
    per_it = []
    for k, d in sorted(launches.items(), key=lambda kv: int(kv[0])):
        name = d["kernel"]
        if "pressure_cg" not in name and "bicgstab" not in name:
            continue
        t = d.get("gpu__time_duration.sum", 0.0)
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / t / 1e9 if t else 0.0
        bpci = (rd + wr) / (cells * max_it) if "pressure_cg" in name else None
        if bpci is not None:
            per_it.append((bpci, gbs, t))
        short = name.split("(")[0].replace("void dpiso::", "")[:60]
        lines.append("| `%s` | %s | %.2f | %.1f | %.1f | %.0f | %.1f %% | %s |" % (short, k, t * 1e3, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak,
                                                                              "%.1f" % bpci if bpci is not None else ""))
    if per_it:
        out = {"dram_bytes_per_cell_iteration": sum(p[0] for p in per_it) / len(per_it),
               "achieved_gbs": sum(p[1] for p in per_it) / len(per_it), "hbm_frac": sum(p[1] for p in per_it) / len(per_it) / peak,
               "us_per_iteration": 1e6 * sum(p[2] for p in per_it) / len(per_it) / max_it, "grid": [n, n], "batch": batch,
               "iterations_per_launch": max_it, "source": "profiles/r02_cg_global.md (ncu dram__bytes, B200)"}
        lines += ["", "pressure_cg_global_kernel: %.1f B of DRAM traffic per cell and iteration (SURVEY 8(d) streaming model: 168 B), "
                  "%.0f GB/s = %.1f %% of the measured HBM peak, %.1f us per iteration of the whole batch." %
                  (out["dram_bytes_per_cell_iteration"], out["achieved_gbs"], 100 * out["hbm_frac"], out["us_per_iteration"])]
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_cg_global.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "r02_cg_global.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
