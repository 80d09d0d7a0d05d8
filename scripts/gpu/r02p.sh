#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02p}
(timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -x -k "plain_stream or failure or shipped" > gpurun_out/${T}_pytest_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_a.log)
tail -n 15 gpurun_out/${T}_pytest_a.log
ARTERYFE_B200_LIB=$PWD/bloodflow_b200/csrc/libarteryfe_b200_tl.so AF_TIMELINE_STEPS=64 timeout 300 python scripts/timeline.py 4096 > gpurun_out/${T}_timeline.log 2>&1
STEPS=8000 bash scripts/gpu/sweep_env.sh scripts/gpu/sweep4.txt ${T}_sweep
