#!/bin/bash
# 8 GPUs of one box: NCCL parity (world 2 and 8), cfg 4 at N = 8 and N = 2, cfg 5 at N = 8
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s > gpurun_out/r30_pytest_multi.txt 2>&1; tail -3 gpurun_out/r30_pytest_multi.txt
python - <<'PY'
import sys, time
sys.path.insert(0, "automatic-ar_b200/python")
from aar_b200 import synth
t = time.time(); synth.make_config("cfg4"); synth.make_config("cfg5"); print("rigs cached", time.time() - t)
PY
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 8 --steps 10 --warmup 3 > gpurun_out/r30_bench_cfg4_n8.json 2> gpurun_out/r30_bench_cfg4_n8.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r30_bench_cfg4_n8.json").read().strip().splitlines()[-1]); print("cfg4 n8", d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"], d["e2e"]["create_s"])
PY
run 4 --steps 10 --warmup 3 > gpurun_out/r30_bench_cfg4_n4.json 2> gpurun_out/r30_bench_cfg4_n4.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r30_bench_cfg4_n4.json").read().strip().splitlines()[-1]); print("cfg4 n4", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
run 8 --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r30_bench_cfg5_n8.json 2> gpurun_out/r30_bench_cfg5_n8.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r30_bench_cfg5_n8.json").read().strip().splitlines()[-1]); print("cfg5 n8", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
tail -2 gpurun_out/r30_bench_cfg4_n8.err
