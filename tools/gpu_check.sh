#!/bin/bash
# One gpurun call: GPU parity tests, both bench arms, the ncu launch list and one full capture of the
# fused STFT kernel at the bench size.  Outputs land in gpurun_out/ (tag = $1).
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_${tag}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/tests_${tag}.log
tail -3 gpurun_out/tests_${tag}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${tag}.log
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cat gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>&1; echo "ref rc=$?"; cat gpurun_out/bench_ref_${tag}.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${tag}.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 3 -c 2 -f -o gpurun_out/prof_bench_${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
echo done
