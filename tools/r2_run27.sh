#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 900 python -m pytest tests/test_gpu_init.py -x -q -m gpu > gpurun_out/r27_pytest_init.txt 2>&1; tail -3 gpurun_out/r27_pytest_init.txt
rm -f gpurun_out/r27_pipeline.jsonl
for w in "cfg3 --consensus-max 256" "cfg4 --frames 20000 --consensus-max 128" "cfg4 --consensus-max 128 --max-iters 30"; do
  timeout 1500 python tools/pipeline_from_detections.py --workload $w >> gpurun_out/r27_pipeline.jsonl 2>> gpurun_out/r27_pipeline.err
done
cat gpurun_out/r27_pipeline.jsonl | cut -c1-1000; tail -5 gpurun_out/r27_pipeline.err
