#!/bin/bash
# the headline bench line exactly as the driver runs it (no flags), one GPU
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/r42_bench_cfg4.json 2> gpurun_out/r42_bench_cfg4.err
tail -c 400 gpurun_out/r42_bench_cfg4.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r42_bench_cfg4.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "parity ok", d.get("parity_check", {}).get("ok"), d["clocks"])
except Exception as e:
    print("no line:", e)
PY
