#!/bin/bash
# Multi-GPU check + measurement of the sharded configs (run with gpurun --gpus G): tools/gpu_sharded.sh G tag
G=${1:-2}; tag=${2:-r01}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) tools/bench_sharded.py "${@:2}" 2>> gpurun_out/sharded_${tag}.err | grep '^{' | tee -a gpurun_out/sharded_${tag}.jsonl; }
python -m pytest tests -m gpu -x -q > gpurun_out/tests_${tag}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_${tag}.log
run $G --config c5 --collective fused --log2-samples 25 --steps 2 --check
run $G --config c5 --collective nccl --log2-samples 25 --steps 2 --check
run $G --config c4 --log2-samples 22 --steps 2 --check
for g in $G; do
  run $g --config c5 --collective fused --steps 5
  run $g --config c5 --collective nccl --steps 5
  run $g --config c4 --steps 5
done
tail -5 gpurun_out/sharded_${tag}.err
