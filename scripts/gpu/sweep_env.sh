#!/bin/bash
# step time of the order-8 tree under environment-variable variants (one B200): "VAR=val ..." per line of $1
mkdir -p gpurun_out
OUT=gpurun_out/${2:-sweep}.log
: > $OUT
while IFS= read -r line; do
  [ -z "$line" ] && continue
  for rep in 1 2; do
    res=$(env $line timeout 300 python bench.py --steps ${STEPS:-2000} --warmup 100 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f us/step value %.4e flags %s'%(d['ms_per_step']*1e3, d['value'], d['status']['flags']))")
    echo "$line | $res" | tee -a $OUT
  done
done < "$1"
