"""Turns gpurun_out/parity_records.jsonl (written by the `-m gpu` tests through tests/common.record) into the committed
parity table profiles/r02_parity.md: observed iteration differences and field errors of the CUDA path (and of the
reference's own kernels) against the oracle.

    python scripts/parity_report.py [records.jsonl] [out.md]
"""
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_records.jsonl")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_parity.md")
    recs = defaultdict(list)
    for line in open(src):
        line = line.strip()
        if line:
            r = json.loads(line)
            recs[r.pop("kind")].append(r)
    out = ["# Round 2: measured parity of the CUDA path against the oracle (B200)", "",
           "Written by `scripts/parity_report.py` from the records the `-m gpu` tests append to",
           "`gpurun_out/parity_records.jsonl` (`tests/common.record`).  Iteration counts of the pressure CG are quantised to",
           "the 5-iteration check cadence (SURVEY Q2); `two reductions` = the kernel run in the reference's reduction order",
           "(`dpiso_pressure_cg_set_reduction_order(1)`), `merged` = the default single reduction (deviation D2).", ""]
    # ---- CG iteration counts ---------------------------------------------------------------------------------------
    ref = {}
    for r in recs.get("cg_reference_kernel", []):
        ref[r["setup"]] = r
    if recs.get("cg"):
        out += ["## pressure CG, fp64: iterations per setup (3 samples each)", "",
                "| setup | reset | oracle | GPU merged | GPU two reductions | max abs diff merged | max abs diff two reductions | max x rel-L2 merged | reference kernel vs oracle (its, diff) |",
                "|---|---|---|---|---|---|---|---|---|"]
        by = defaultdict(list)
        for r in recs["cg"]:
            by[r["setup"]].append(r)
        worst = defaultdict(int)
        for name, rows in by.items():
            rows.sort(key=lambda r: r["sample"])
            d1 = max(abs(r["it_gpu"] - r["it_oracle"]) for r in rows)
            d2 = max(abs(r["it_gpu_two_reductions"] - r["it_oracle"]) for r in rows)
            reset = rows[0]["reset"]
            worst[("merged", reset <= 10)] = max(worst[("merged", reset <= 10)], d1)
            worst[("two", reset <= 10)] = max(worst[("two", reset <= 10)], d2)
            rk = ref.get(name)
            rk_s = "-" if rk is None else "%d vs %d (%+d)" % (rk["it_reference"], rk["it_oracle"], rk["it_reference"] - rk["it_oracle"])
            if rk is not None:
                worst[("ref", reset <= 10)] = max(worst[("ref", reset <= 10)], abs(rk["it_reference"] - rk["it_oracle"]))
            out.append("| %s | %d | %s | %s | %s | %d | %d | %.1e | %s |" % (
                name, reset, ", ".join(str(r["it_oracle"]) for r in rows), ", ".join(str(r["it_gpu"]) for r in rows),
                ", ".join(str(r["it_gpu_two_reductions"]) for r in rows), d1, d2, max(r["x_rel_l2"] for r in rows), rk_s))
        out += ["", "Worst observed |iterations - oracle|: " + "; ".join(
            "%s, reset %s: %d" % (k[0], "<= 10" if k[1] else "1000", v) for k, v in sorted(worst.items(), key=str)), ""]
    for r in recs.get("cg_reference_kernel", []):
        pass
    # ---- step / adjoint field errors ------------------------------------------------------------------------------------
    def table(kind, title, cols):
        rows = recs.get(kind)
        if not rows:
            return
        out.extend(["## " + title, "", "| setup | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)])
        by = defaultdict(list)
        for r in rows:
            by[r.get("setup", kind)].append(r)
        for name, rr in by.items():
            cells = []
            for c in cols:
                vals = [r[c] for r in rr if r.get(c) is not None]
                if not vals:
                    cells.append("-")
                elif isinstance(vals[0], (int, float)):
                    cells.append("%.2e" % max(vals) if isinstance(vals[0], float) else str(max(vals)))
                else:
                    cells.append(str(vals[0]))
            out.append("| %s | %s |" % (name, " | ".join(cells)))
        out.append("")
    table("step", "piso_step forward (3 steps x 2 samples): worst relative L2 against the oracle",
          ["cg_tol", "vel_rel_l2", "pres_rel_l2", "p1_rel_l2"])
    table("adjoint", "piso_step backward (2 samples): worst relative L2 of the gradients against oracle/adjoint.py",
          ["cg_tol", "g_vel", "g_pres", "g_forcing"])
    table("full_adjoint", "full-size fwd+adjoint (BASELINE configs 2-4): worst relative L2 over the checked samples",
          ["vel", "pres", "g_vel", "g_pres", "cg_adj_it", "cg_adj_it_oracle"])
    for kind in ("c1_ldc32_200steps", "c2_128_1000steps", "c5_1024_fwd_adjoint"):
        for r in recs.get(kind, []):
            out += ["## " + kind, "", "```", json.dumps(r, indent=1), "```", ""]
    its = [(r["setup"], r["cg2_it"], r["cg2_it_oracle"]) for r in recs.get("step", [])]
    if its:
        out += ["Second-corrector CG iterations inside the steps: worst |GPU - oracle| = %d over %d solves." % (
            max(abs(a - b) for _, a, b in its), len(its)), ""]
    bi = [(r["bicg_it"], r["bicg_it_oracle"]) for r in recs.get("step", [])]
    if bi:
        out += ["BiCGStab iterations inside the steps: worst |GPU - oracle| = %d over %d solves." % (
            max(max(abs(a[0] - b[0]), abs(a[1] - b[1])) for a, b in bi), 2 * len(bi)), ""]
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
