#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -k "not full_cycle" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest.log)
tail -6 gpurun_out/r02h_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err; echo "default exit $?"
timeout 300 python bench.py --no-c5 --no-cpu-baseline --theta 0.55 > gpurun_out/r02h_bench_theta055.json 2> gpurun_out/r02h_bench_theta055.err; echo "theta055 exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r02h_bench_k20.json 2> gpurun_out/r02h_bench_k20.err
