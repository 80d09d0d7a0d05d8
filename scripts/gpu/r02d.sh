#!/bin/bash
# dataflow validation + measurement
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log)
tail -6 gpurun_out/r02d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r02d_bench_k20.json 2> gpurun_out/r02d_bench_k20.err; echo "k20 exit $?"
timeout 300 python bench.py --no-c5 --no-cpu-baseline > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err; echo "default exit $?"
AF_NO_DATAFLOW=1 timeout 300 python bench.py --no-c5 --no-cpu-baseline > gpurun_out/r02d_bench_nodataflow.json 2> gpurun_out/r02d_bench_nodataflow.err; echo "nodataflow exit $?"
timeout 300 python bench.py --no-c5 --no-cpu-baseline --theta 0.55 > gpurun_out/r02d_bench_theta055.json 2> gpurun_out/r02d_bench_theta055.err; echo "theta055 exit $?"
AF_NO_DATAFLOW=1 timeout 300 python bench.py --no-c5 --no-cpu-baseline --theta 0.55 > gpurun_out/r02d_bench_theta055_nodataflow.json 2> gpurun_out/r02d_bench_theta055_nodataflow.err
