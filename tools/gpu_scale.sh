#!/bin/bash
# Scaling runs on one box with G GPUs (gpurun --gpus G): the driver's bench contract at N = 1..G (replicas of config 2)
# and the sharded configs 4 and 5 (strong scaling).  tools/gpu_scale.sh G tag
G=${1:-2}; tag=${2:-r01}
mkdir -p gpurun_out
out=gpurun_out/scale_${tag}.jsonl; err=gpurun_out/scale_${tag}.err
: > $out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}" 2>> $err | grep '^{' | tee -a $out | cut -c1-400; }
python bench.py --steps 10 --no-cpu-baseline 2>> $err | grep '^{' | tee -a $out | cut -c1-400
for n in 2 4 8; do
  [ $n -le $G ] || continue
  tr $n bench.py --gpus $n --steps 10 --warmup 3
done
tr 2 bench.py --impl reference --gpus 2 --steps 2 --warmup 1
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  tr $n tools/bench_sharded.py --config c5 --collective fused --steps 5
  tr $n tools/bench_sharded.py --config c5 --collective nccl --steps 5
  tr $n tools/bench_sharded.py --config c4 --steps 5
done
tr $G tools/bench_sharded.py --config c5 --collective fused --log2-samples 25 --steps 2 --check
tr $G tools/bench_sharded.py --config c4 --log2-samples 22 --steps 2 --check
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $err | tail -8
