#!/bin/bash
# N-GPU bench lines exactly as the driver launches them (one box, N ranks over NCCL): reference arm, then ours
N=${1:-2}
mkdir -p gpurun_out
PORT=29517
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
timeout 600 $L bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_g${N}_ref.json 2> gpurun_out/r02_g${N}_ref.err; echo "ref exit $?"
timeout 900 $L bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_g${N}_k20.json 2> gpurun_out/r02_g${N}_k20.err; echo "k20 exit $?"
timeout 900 $L bench.py --gpus $N --no-c5 --no-cpu-baseline > gpurun_out/r02_g${N}_cycle.json 2> gpurun_out/r02_g${N}_cycle.err; echo "cycle exit $?"
tail -c 600 gpurun_out/r02_g${N}_k20.err
