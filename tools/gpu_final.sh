#!/bin/bash
# Round-end record on one B200: full GPU test suite, smoke, the bench line, the ncu launch list of the bench command and one
# `ncu --set full` capture of the dominant kernel at bench size.  Usage: gpu_final.sh TAG
cd "${GRAFT_REPO_ROOT:-.}"
T=${1:-r3}
mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/aar_rigs
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
T=$T python - <<'PY'
import json, os
try:
    d = json.loads(open("gpurun_out/%s_bench.json" % os.environ["T"]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["phases_ms_per_step"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-track --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
wc -l gpurun_out/${T}_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_jac_project -c 1 -f -o gpurun_out/${T}_project \
    python bench.py --steps 1 --warmup 1 --no-track --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_project.ncu-rep
