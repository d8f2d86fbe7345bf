#!/bin/bash
# 8 GPUs of one box: NCCL parity of the frame-sharded solve (world 2 and 8), cfg 4 and cfg 5 at N = 8
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s > gpurun_out/r12_pytest_multi.txt 2>&1; tail -4 gpurun_out/r12_pytest_multi.txt
python - <<'PY'
import sys, time
sys.path.insert(0, "automatic-ar_b200/python")
from aar_b200 import synth
t = time.time(); synth.make_config("cfg4"); synth.make_config("cfg5"); print("rigs cached", time.time() - t)
PY
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r12_bench_cfg5_n1.json 2> gpurun_out/r12_bench_cfg5_n1.err; tail -c 600 gpurun_out/r12_bench_cfg5_n1.json; tail -3 gpurun_out/r12_bench_cfg5_n1.err
timeout 600 python bench.py --workload cfg5 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r12_bench_cfg5_n8.json 2> gpurun_out/r12_bench_cfg5_n8.err; head -c 700 gpurun_out/r12_bench_cfg5_n8.json; tail -3 gpurun_out/r12_bench_cfg5_n8.err
timeout 600 python bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r12_bench_cfg4_n8.json 2> gpurun_out/r12_bench_cfg4_n8.err; head -c 400 gpurun_out/r12_bench_cfg4_n8.json; tail -3 gpurun_out/r12_bench_cfg4_n8.err
