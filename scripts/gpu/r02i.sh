#!/bin/bash
# GPU evidence run (one B200): quick GPU tests, default bench (tree + C5), driver-argument bench, ncu launch
# list and `--set full` captures of the current kernels on the tree and on the ensemble.
mkdir -p gpurun_out
T=r02i
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not full_cycle" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log)
tail -n 4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_k20.json 2> gpurun_out/${T}_bench_k20.err; echo "k20 exit $?"
timeout 300 python scripts/e2e_breakdown.py 20 6 > gpurun_out/${T}_e2e.log 2>&1
CMD="python bench.py --steps 300 --warmup 20 --no-c5 --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 400 -c 4 -o gpurun_out/${T}_tree $CMD > gpurun_out/${T}_ncu_tree.log 2>&1
python scripts/run_c5_small.py 4096 60 > gpurun_out/${T}_c5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 64 -c 4 -o gpurun_out/${T}_c5 python scripts/run_c5_small.py 4096 60 > gpurun_out/${T}_ncu_c5.log 2>&1
ls -la gpurun_out | tail -n 12
