#!/bin/bash
# quick GPU tests + bench lines after a change (one B200)
mkdir -p gpurun_out
T=${1:-r02j}
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not full_cycle" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log)
tail -n 6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default exit $?"
timeout 300 python bench.py --no-c5 --no-cpu-baseline --theta 0.55 > gpurun_out/${T}_bench_theta055.json 2> gpurun_out/${T}_bench_theta055.err; echo "theta055 exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/${T}_bench_k20.json 2> gpurun_out/${T}_bench_k20.err; echo "k20 exit $?"
timeout 300 python scripts/e2e_breakdown.py 20 6 > gpurun_out/${T}_e2e.log 2>&1
