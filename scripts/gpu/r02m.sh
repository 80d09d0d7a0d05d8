#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02m}
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not full_cycle" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log)
tail -n 6 gpurun_out/${T}_pytest.log
ARTERYFE_B200_LIB=$PWD/bloodflow_b200/csrc/libarteryfe_b200_tl.so AF_TIMELINE_STEPS=64 timeout 300 python scripts/timeline.py 4096 > gpurun_out/${T}_timeline.log 2>&1
timeout 300 python bench.py --no-c5 --no-cpu-baseline > gpurun_out/${T}_bench_cycle.json 2> gpurun_out/${T}_bench_cycle.err; echo "cycle exit $?"
timeout 300 python bench.py --no-c5 --no-cpu-baseline --theta 0.55 > gpurun_out/${T}_bench_theta055.json 2> gpurun_out/${T}_bench_theta055.err; echo "theta055 exit $?"
python - $T <<'P'
import json,sys
for f in ['cycle','theta055']:
    try:
        d=json.loads(open('gpurun_out/%s_bench_%s.json'%(sys.argv[1],f)).read().strip().splitlines()[-1])
        print(f,'%.3f us/step value %.4e e2e %.4e flags %s'%(d['ms_per_step']*1e3,d['value'],d['e2e']['value'],d['status']['flags']))
    except Exception as e: print(f,'ERR',e)
P
