#!/bin/bash
# Round-end record on one B200: whole GPU test suite (parity records kept), smoke, bench lines of cfg 4 / 1 / 2 / 3 / 5, the reference arm,
# launch list of the bench command and `ncu --set full` captures of the big kernels at bench size.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
T=rfin
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.txt
cp gpurun_out/parity_achieved.jsonl gpurun_out/${T}_parity_achieved.jsonl
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err; tail -c 800 gpurun_out/${T}_bench_cfg4.json; tail -3 gpurun_out/${T}_bench_cfg4.err
for w in cfg1 cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; head -c 250 gpurun_out/${T}_bench_$w.json; echo; tail -2 gpurun_out/${T}_bench_$w.err
done
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err; head -c 250 gpurun_out/${T}_bench_cfg5.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; head -c 400 gpurun_out/${T}_bench_reference.json; echo
AAR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
wc -l gpurun_out/${T}_launches_cfg4.csv
AAR_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_jac_project|k_asm_pairs|k_asm_mruns|k_schur_syrk|k_pair_tab|k_schur_prepare|k_residual|k_backsub|k_reduced_solve" -c 9 -f -o gpurun_out/${T}_kernels_cfg4 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_kernels_cfg4.ncu-rep; tail -2 gpurun_out/${T}_ncu_full.log
