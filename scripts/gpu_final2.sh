#!/bin/bash
# Second measurement pass of round 2 (after the right-looking apply kernel, the wider cast-fused Gram grid and the
# shadow skip): tests, both bench arms, launch lists, the captures of the kernels that changed, sanitizers.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -rf --no-header > $O/r02_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_bench_b200.json 2> $O/r02_bench_b200.err; echo "bench rc=$?"
OURS='regex:tc_|gram128|chol128|apply128|trinv128|colmax128|splitk_reduce|cast_shadow|finish_r12|peer_allreduce|gram32|chol32|apply32'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 2700 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/r02_bench_under_ncu.log 2>&1; echo "ncu bench rc=$?"
for shape in "16384 16384" "1048576 1024" "131072 1024"; do
  set -- $shape
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1000 --csv --log-file $O/r02_launches_$1x$2.csv \
      python scripts/gpu_profile_run.py $1 $2 1 > $O/r02_launches_$1x$2.log 2>&1; echo "launch list $1x$2 rc=$?"
done
cap() {  # name demangled-regex count m n
  timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -c $3 -f -o $O/prof_$1 \
      python scripts/gpu_profile_run.py $4 $5 1 > $O/prof_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap apply16k 'apply128_kernel' 1 16384 16384
cap gramcast_1m 'tc_gram_cast_kernel' 3 1048576 1024
cap update_1m 'tc_update_kernel<\(int\)128' 1 1048576 1024
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python scripts/gpu_sanitize_run.py square panel panel32 host ormqr qdwh > $O/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $O/r02_sanitize_$tool.log
done
timeout 400 compute-sanitizer --tool memcheck python scripts/gpu_sanitize_run.py tall > $O/r02_sanitize_memcheck_tall.log 2>&1; echo "memcheck tall rc=$?"; tail -2 $O/r02_sanitize_memcheck_tall.log
timeout 400 compute-sanitizer --tool racecheck python scripts/gpu_sanitize_run.py tall > $O/r02_sanitize_racecheck_tall.log 2>&1; echo "racecheck tall rc=$?"; tail -2 $O/r02_sanitize_racecheck_tall.log
