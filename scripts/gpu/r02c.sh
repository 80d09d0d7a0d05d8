#!/bin/bash
# GPU evidence run of round 2 (one B200): full GPU test suite, the default bench, the driver-argument
# bench, an ncu launch list and `--set full` captures of the two kernels on the tree and on the ensemble.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log)
tail -4 gpurun_out/r02c_pytest.log
timeout 600 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; echo "bench default exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 > gpurun_out/r02c_bench_k20.json 2> gpurun_out/r02c_bench_k20.err
CMD="python bench.py --steps 300 --warmup 20 --no-c5 --no-cpu-baseline"
$CMD > gpurun_out/r02c_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r02c_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 400 -c 4 -o gpurun_out/r02c_tree $CMD > gpurun_out/r02c_ncu_tree.log 2>&1
python scripts/run_c5_small.py 4096 60 > gpurun_out/r02c_c5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 64 -c 4 -o gpurun_out/r02c_c5 python scripts/run_c5_small.py 4096 60 > gpurun_out/r02c_ncu_c5.log 2>&1
ls -la gpurun_out | tail -12
