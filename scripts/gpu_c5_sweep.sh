# BASELINE configs[4] sweep on one GPU: forward + adjoint step, every CG launch capped at 300 iterations (fixed work), so
# the numbers compare kernels, not iteration counts.  One JSON line per configuration -> gpurun_out/c5_sweep.jsonl
mkdir -p gpurun_out
rm -f gpurun_out/c5_sweep.jsonl
for cfg in "1024 8" "1024 16" "1024 32" "1024 64" "2048 4" "2048 8" "2048 16"; do
  set -- $cfg
  timeout 600 python bench.py --config5-only --config5-maxit 300 --config5-n $1 --config5-batch $2 >> gpurun_out/c5_sweep.jsonl 2> gpurun_out/c5_sweep.err || echo "{\"failed\": \"$cfg\"}" >> gpurun_out/c5_sweep.jsonl
done
python - <<'PY'
import json
print("| grid | batch | ms per fwd+adjoint step | cell-updates/s | CG us per iteration (whole batch) | CG share | BiCGStab ms per launch | BiCGStab share |")
print("|---|---|---|---|---|---|---|---|")
for line in open("gpurun_out/c5_sweep.jsonl"):
    d = json.loads(line)
    if "failed" in d: print("| failed:", d["failed"], "|"); continue
    w = d["workload"].split("_")
    print("| %s | %s | %.0f | %.3g | %.1f | %.1f %% | %.1f | %.1f %% |" % (w[1], w[2].replace("batch", ""), d["ms_per_step"], d["cell_updates_per_s"],
          d["pressure_cg"]["us_per_iteration_whole_batch"], 100 * d["pressure_cg"]["share_of_step"], d["bicgstab"]["ms_per_launch"], 100 * d["bicgstab"]["share_of_step"]))
PY
