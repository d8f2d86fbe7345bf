#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
for w in "cfg1 --consensus-max 0" "cfg2 --consensus-max 0" "cfg3 --consensus-max 256" "cfg4 --frames 20000 --consensus-max 128"; do
  timeout 900 python tools/pipeline_from_detections.py --workload $w >> gpurun_out/r26_pipeline.jsonl 2>> gpurun_out/r26_pipeline.err
done
cat gpurun_out/r26_pipeline.jsonl | cut -c1-900; tail -5 gpurun_out/r26_pipeline.err
