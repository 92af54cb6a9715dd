#!/bin/bash
# Third (last) single-GPU pass of round 2, after the fused node kernel: tests, both bench arms, launch lists,
# one capture and the sanitizers over the kernels that are new.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -rf --no-header > $O/r02_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_bench_b200.json 2> $O/r02_bench_b200.err; echo "bench rc=$?"
OURS='regex:tc_|gram128|chol128|apply128|trinv128|colmax128|splitk_reduce|cast_shadow|finish_r12|peer_allreduce|gram32|chol32|apply32'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 2300 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/r02_bench_under_ncu.log 2>&1; echo "ncu bench rc=$?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1000 --csv --log-file $O/r02_launches_16384x16384.csv \
    python scripts/gpu_profile_run.py 16384 16384 1 > $O/r02_launches_16384x16384.log 2>&1; echo "launch list rc=$?"
timeout 100 ncu --set full --clock-control none -k regex:tc_node_kernel -c 1 -f -o $O/prof_node16k python scripts/gpu_profile_run.py 16384 16384 1 > $O/prof_node16k.log 2>&1; echo "ncu node rc=$?"
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool python scripts/gpu_sanitize_run.py square panel host > $O/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $O/r02_sanitize_$tool.log
done
