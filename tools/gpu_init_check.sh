#!/bin/bash
# Initialisation path: GPU parity tests + the pipeline from raw detections at cfg 3 and cfg 4
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 900 python -m pytest tests/test_gpu_init.py tests/test_host_facade.py -x -q -m gpu > gpurun_out/init_pytest.txt 2>&1; tail -3 gpurun_out/init_pytest.txt
rm -f gpurun_out/init_pipeline.jsonl
for w in "cfg3 --consensus-max 256" "cfg4 --consensus-max 128 --max-iters 30"; do
  timeout 1500 python tools/pipeline_from_detections.py --workload $w >> gpurun_out/init_pipeline.jsonl 2>> gpurun_out/init_pipeline.err
done
cut -c1-700 gpurun_out/init_pipeline.jsonl; tail -3 gpurun_out/init_pipeline.err
