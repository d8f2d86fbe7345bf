#!/bin/bash
# the analytic variant at BASELINE cfg 4 on one GPU (not the headline: see DESIGN.md section 6b)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 200 python bench.py --analytic --steps 10 --warmup 3 > gpurun_out/r41_bench_cfg4_analytic.json 2> gpurun_out/r41_bench_cfg4_analytic.err
tail -c 600 gpurun_out/r41_bench_cfg4_analytic.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r41_bench_cfg4_analytic.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["phases_ms_per_step"], d["clocks"])
except Exception as e:
    print("no line:", e)
PY
