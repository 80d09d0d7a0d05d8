#!/bin/bash
# Final GPU evidence of round 2 (one B200).  ncu serialises kernels, so the persistent pair k_artery_p /
# k_junction_p (which wait for each other) cannot run under it: the captures are taken with AF_NO_PERSIST=1,
# i.e. of the per-step kernels k_artery / k_boundary, which execute the same device code.
mkdir -p gpurun_out
T=r02f
timeout 600 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 > gpurun_out/${T}_bench_k20.json 2> gpurun_out/${T}_bench_k20.err; echo "k20 exit $?"
timeout 300 python bench.py --theta 0.55 --no-c5 --no-cpu-baseline > gpurun_out/${T}_bench_theta055.json 2> gpurun_out/${T}_bench_theta055.err; echo "theta055 exit $?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_k20.json 2> gpurun_out/${T}_bench_reference.err; echo "reference exit $?"
AF_NO_PERSIST=1 timeout 300 python bench.py --no-c5 --no-cpu-baseline > gpurun_out/${T}_bench_per_step_launches.json 2>/dev/null
timeout 300 python scripts/e2e_breakdown.py 20 6 > gpurun_out/${T}_e2e.log 2>&1
ARTERYFE_B200_LIB=$PWD/bloodflow_b200/csrc/libarteryfe_b200_tl.so AF_TIMELINE_STEPS=64 timeout 300 python scripts/timeline.py 4096 > gpurun_out/${T}_timeline_persistent.log 2>&1
AF_NO_PERSIST=1 ARTERYFE_B200_LIB=$PWD/bloodflow_b200/csrc/libarteryfe_b200_tl.so AF_TIMELINE_STEPS=64 timeout 300 python scripts/timeline.py 4096 > gpurun_out/${T}_timeline_per_step.log 2>&1
timeout 120 python scripts/critical_path_probe.py 4000 > gpurun_out/${T}_critical_path.log 2>&1
export AF_NO_PERSIST=1
CMD="python bench.py --steps 300 --warmup 20 --no-c5 --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 400 -c 4 -o gpurun_out/${T}_tree $CMD > gpurun_out/${T}_ncu_tree.log 2>&1
unset AF_NO_PERSIST
python scripts/run_c5_small.py 4096 60 > gpurun_out/${T}_c5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_artery|k_boundary" -s 64 -c 4 -o gpurun_out/${T}_c5 python scripts/run_c5_small.py 4096 60 > gpurun_out/${T}_ncu_c5.log 2>&1
ls -la gpurun_out | grep ${T}_
