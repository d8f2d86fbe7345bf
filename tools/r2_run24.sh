#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
T=r24
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -15 gpurun_out/${T}_pytest_gpu.txt
for w in cfg1 cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_$w.json").read().strip().splitlines()[-1])
print("$w", "ms/step", d["ms_per_step"], "e2e ms/step", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
PY
  tail -2 gpurun_out/${T}_bench_$w.err
done
AAR_NO_GRAPH=1 timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1 no graph: e2e ms/step', d['e2e']['ms_per_step'])"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", "ms/step", d["ms_per_step"], "e2e ms/step", d["e2e"]["ms_per_step"], d["phases_ms_per_step"])
PY
