#!/bin/bash
# Development aid: one gpurun call = [microbenchmarks] + parity subset (default build, env switches, variant libraries) +
# timing of all of them on cfg 4 / 20k frames + [ncu --set full of the Jacobian kernels].
# Usage: gpu_ab.sh TAG [divcheck] [dmma] [ncu] [env:K=V ...]
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-r1x}; shift
mkdir -p gpurun_out
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core"
LIBS=default
for opt in "$@"; do
  case $opt in
    dmma) timeout 60 ./automatic-ar_b200/dmma_peak | tee gpurun_out/${TAG}_dmma_peak.txt;;
    divcheck) timeout 120 ./automatic-ar_b200/divcheck > gpurun_out/${TAG}_divcheck.txt 2>&1; cat gpurun_out/${TAG}_divcheck.txt;;
    env:*) kv=${opt#env:}; echo "== parity subset with $kv"; env $kv timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" 2>&1 | tail -4; LIBS=$LIBS,default:$kv;;
  esac
done
for f in automatic-ar_b200/variants/*.so; do [ -f "$f" ] || continue; echo "== parity subset with $f"
  AAR_LIB=$PWD/$f timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" 2>&1 | tail -4; LIBS=$LIBS,$f; done
echo "== parity subset, default build"
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/${TAG}_pytest_subset.txt 2>&1
tail -15 gpurun_out/${TAG}_pytest_subset.txt
timeout 400 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 --libs $LIBS > gpurun_out/${TAG}_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/${TAG}_variants.txt
for opt in "$@"; do
  if [ "$opt" = ncu ]; then
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_jac_ -c 2 -f -o gpurun_out/${TAG}_jac \
        python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/${TAG}_ncu.log 2>&1
    tail -3 gpurun_out/${TAG}_ncu.log
  fi
done
